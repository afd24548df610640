"""Pins the oracle END TO END against an artifact the reference's own renderer produced: /root/reference/output/output.png,
the image `cargo run` of /root/reference/src/main.rs:5-20 writes (Cornell box, 300x300, 300 spp, depth 20, background 0.001).

The reference's RNG is OS-seeded ChaCha12 (utils/random.rs:15-18) and the oracle's is Philox, so the two
images are two independent Monte-Carlo estimates of the same integrand: they cannot agree pixel by pixel, but they must be
statistically indistinguishable. The yardstick is the oracle against ITSELF with another seed:

  * PSNR(oracle seed 1, reference) within 0.3 dB of PSNR(oracle seed 1, oracle seed 2)   — the noise floor of 300 spp;
  * mean |difference| of the 10x10 block means (30x30-pixel blocks) within 1.2x of the seed-vs-seed value — catches a
    wrong BVH (missing primitives), wrong material, camera or tone map, which shift whole regions rather than add noise;
  * global mean within 1 %.
A bug in any of BVH build / traversal, quad test, Lambertian scatter, light emission, camera, accumulation or gamma moves
these far outside the band (e.g. dropping the `normal +` of lambertian.rs:18 costs 3 dB; a 2.0 gamma costs 6 dB).

The decoded PNG is committed as tests/golden/reference_output_300.npz (generator: tests/golden/make_reference_png_fixture.py)
because /root/reference does not exist on the GPU box; where it does exist the fixture is re-checked against the file."""
import importlib.util
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURE = os.path.join(HERE, "golden", "reference_output_300.npz")


def psnr8(a, b):
    mse = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)
    return 10.0 * np.log10(255.0 ** 2 / mse)


def block_means(img, n=10):
    h, w, _ = img.shape
    return img.astype(np.float64).reshape(n, h // n, n, w // n, 3).mean(axis=(1, 3))


def test_fixture_is_the_reference_png():
    ref = np.load(FIXTURE)["rgb8"]
    assert ref.shape == (300, 300, 3) and ref.dtype == np.uint8
    src = "/root/reference/output/output.png"
    if os.path.exists(src):
        spec = importlib.util.spec_from_file_location("mkfix", os.path.join(HERE, "golden", "make_reference_png_fixture.py"))
        m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
        assert np.array_equal(m.decode_png_rgb8(src), ref)


def test_oracle_render_is_statistically_the_reference_image(orc, trt):
    ref = np.load(FIXTURE)["rgb8"]
    desc = trt.scenes.cornell(300, 300, 300, 20)                 # main.rs:7-17: the scene, camera and Renderer::new arguments
    assert desc.background == (0.001, 0.001, 0.001) or tuple(np.float32(desc.background)) == tuple(np.float32([0.001] * 3))
    sc = orc.Scene.from_desc(desc); cam = orc.make_camera(**desc.camera)
    imgs = []
    for seed in (1, 2):
        _, u8, st, _ = sc.render(cam, orc.make_params(300, 300, 300, 20, desc.background, seed))
        assert st.paths == 300 * 300 * 300
        imgs.append(u8.reshape(300, 300, 3))
    a, b = imgs
    floor_db, ref_db = psnr8(a, b), psnr8(a, ref)
    assert abs(ref_db - floor_db) <= 0.3, "PSNR vs the reference's PNG %.2f dB, seed-vs-seed floor %.2f dB" % (ref_db, floor_db)
    blk_floor = np.abs(block_means(a) - block_means(b)).mean()
    blk_ref = np.abs(block_means(a) - block_means(ref)).mean()
    assert blk_ref <= 1.2 * blk_floor, "block means: %.3f vs reference, %.3f seed-vs-seed" % (blk_ref, blk_floor)
    assert abs(a.mean() - ref.mean()) <= 0.01 * ref.mean(), (a.mean(), ref.mean())
    # and the reference image is NOT explained by a visibly different integrand: the same test fails for 19 instead of 20 bounces?
    # (too subtle) — instead: a render without the two boxes (12 of the 18 quads missing) must fall outside the band
    keep = np.arange(desc.n_prims) < 6
    walls = trt.scenes.SceneDesc("walls", desc.materials, desc.prim_type[keep], desc.prim_data[keep], desc.prim_mat[keep],
                                 dict(desc.camera), 64, 20, desc.background, 1)
    _, wu8, _, _ = orc.Scene.from_desc(walls).render(cam, orc.make_params(300, 300, 64, 20, desc.background, 1))
    assert np.abs(block_means(wu8.reshape(300, 300, 3)) - block_means(ref)).mean() > 3.0 * blk_floor
