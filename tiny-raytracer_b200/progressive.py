"""Progressive rendering and checkpoint/resume by sample range.

The reference has neither (no partial-image save, no sample-range resume: SURVEY.md §5). Here both fall out of the keyed
Philox streams: sample index s of pixel p is the same path whenever and wherever it is rendered, so a render can be
extended (`add_samples`), previewed (`image`) and checkpointed (`state` / `from_state`) at any sample count, and the
result equals a single pass over the same sample indices (f64 sums; <= 1 ulp after rounding, tested).
The f64 accumulator lives on the device in a torch tensor; pixels are computed by libtrt_b200.so only.
"""
from __future__ import annotations

import numpy as np

from . import api


class ProgressiveRenderer:
    def __init__(self, ctx: api.Context, scene: api.Scene, cam, max_bounces: int, background=(0, 0, 0), seed: int = 1,
                 variant: int = api.VARIANT_MEGAKERNEL):
        import torch
        self._torch = torch
        self.ctx, self.scene, self.cam = ctx, scene, cam
        self.max_bounces, self.background, self.seed, self.variant = int(max_bounces), tuple(background), int(seed), int(variant)
        self.width, self.height = cam.width, cam.height
        self.samples_done = 0
        dev = torch.device("cuda", ctx.device)
        n = self.width * self.height * 3
        self.acc = torch.zeros(n, dtype=torch.float64, device=dev)
        self._lin = torch.empty(n, dtype=torch.float32, device=dev)
        self._u8 = torch.empty(n, dtype=torch.uint8, device=dev)
        torch.cuda.synchronize(dev)
        scene.upload(ctx)

    def add_samples(self, n: int):
        """Render global sample indices [samples_done, samples_done + n) of every pixel into the accumulator."""
        if n <= 0:
            return None
        p = api.make_params(self.width, self.height, n, self.max_bounces, self.background, self.seed,
                            sample_offset=self.samples_done, spp_total=self.samples_done + n, variant=self.variant)
        st = self.scene.render_accumulate(self.ctx, self.cam, p, self.acc.data_ptr())
        self.samples_done += n
        return st

    def image(self) -> api.Image:
        """The image of the samples rendered so far: (float)(sum / samples_done), gamma 2.2, 8 bit."""
        if self.samples_done == 0:
            z = np.zeros((self.height, self.width, 3), np.float32)
            return api.Image(z, z.astype(np.uint8))
        self.ctx.synchronize()
        api.finalize(self.ctx, self.acc.data_ptr(), self.width, self.height, self.samples_done, self._lin.data_ptr(), self._u8.data_ptr())
        self.ctx.synchronize()
        shape = (self.height, self.width, 3)
        return api.Image(self._lin.cpu().numpy().reshape(shape), self._u8.cpu().numpy().reshape(shape))

    def state(self) -> dict:
        """Checkpoint: everything needed to continue elsewhere (the f64 sums and the sample count)."""
        self.ctx.synchronize()
        return {"sum64": self.acc.cpu().numpy().copy(), "samples_done": self.samples_done, "width": self.width, "height": self.height,
                "seed": self.seed, "max_bounces": self.max_bounces, "background": self.background}

    def load_state(self, st: dict):
        if (st["width"], st["height"]) != (self.width, self.height) or st["seed"] != self.seed or st["max_bounces"] != self.max_bounces:
            raise ValueError("checkpoint does not belong to this render")
        self.acc.copy_(self._torch.from_numpy(np.ascontiguousarray(st["sum64"], np.float64).reshape(-1)))
        self._torch.cuda.synchronize()
        self.samples_done = int(st["samples_done"])
