"""CPU tests of the N > 1 host logic: sample partitioning and the f64 accumulate-then-reduce decomposition, run with
world_size 2 and 3 on the gloo backend. The per-rank renderer here is the ORACLE standing in for a GPU (checker only);
what is under test is tiny_raytracer_b200.multigpu and the claim that partial f64 sums reduce to the single-rank image."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_partition_covers_all_samples():
    import tiny_raytracer_b200 as trt
    from tiny_raytracer_b200.multigpu import partition_samples
    for spp in (0, 1, 3, 7, 16, 1024, 4096, 1023):
        for world in (1, 2, 3, 4, 8):
            parts = [partition_samples(spp, world, r) for r in range(world)]
            assert parts[0][0] == 0 and sum(c for _, c in parts) == spp
            for (o0, c0), (o1, _) in zip(parts, parts[1:]):
                assert o0 + c0 == o1
            assert max(c for _, c in parts) - min(c for _, c in parts) <= 1
    with pytest.raises(ValueError):
        partition_samples(4, 2, 2)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, spp, out_path):
    import sys
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import tiny_raytracer_b200 as trt
    from tiny_raytracer_b200.multigpu import partition_samples, reduce_accumulators
    from oracle import pyoracle as orc
    desc = trt.scenes.cornell_spheres(40, 30, spp, 20)
    sc = orc.Scene.from_desc(desc); cam = orc.make_camera(**desc.camera)
    off, cnt = partition_samples(spp, world, rank)
    acc = torch.zeros(30 * 40 * 3, dtype=torch.float64)
    if cnt:
        p = orc.make_params(40, 30, cnt, 20, desc.background, desc.seed, sample_offset=off, spp_total=spp, n_threads=1)
        _, _, _, s64 = sc.render(cam, p, want_sum64=True)
        acc += torch.from_numpy(s64.reshape(-1))
    reduce_accumulators(acc, 0)
    if rank == 0:
        np.save(out_path, acc.numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("world,spp", [(2, 8), (3, 7), (2, 1)])
def test_spp_partition_reduce_equals_single_rank(tmp_path, world, spp):
    import tiny_raytracer_b200 as trt
    from oracle import pyoracle as orc
    out = str(tmp_path / "acc.npy")
    mp.spawn(_worker, args=(world, _free_port(), spp, out), nprocs=world, join=True)
    acc = np.load(out)
    desc = trt.scenes.cornell_spheres(40, 30, spp, 20)
    sc = orc.Scene.from_desc(desc); cam = orc.make_camera(**desc.camera)
    lin, _, _, s64 = sc.render(cam, orc.make_params(40, 30, spp, 20, desc.background, desc.seed), want_sum64=True)
    np.testing.assert_allclose(acc, s64.reshape(-1), rtol=1e-13, atol=0)
    got = (acc * (1.0 / spp)).astype(np.float32)
    ulp = np.abs(got.view(np.int32).astype(np.int64) - lin.reshape(-1).view(np.int32).astype(np.int64))
    assert ulp.max() <= 1
