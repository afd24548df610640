"""CPU tests of the C-ABI library and the host-side mirror: the .so loads and exports every symbol include/trt.h
declares; host-only entry points (scene building, BVH build, camera) work without a GPU and agree with the oracle;
compute entry points fail loudly without a device (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(trt):
    hdr = open(os.path.join(ROOT, "include", "trt.h")).read()
    declared = set(re.findall(r"\b(trt_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(trt._lib.SYMBOLS), declared ^ set(trt._lib.SYMBOLS)
    lib = C.CDLL(trt.LIB_PATH)
    for name in declared:
        assert getattr(lib, name) is not None
    nm = subprocess.run(["nm", "-D", "--defined-only", trt.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (trt_[a-z0-9_]+)", nm))
    assert declared <= exported
    assert trt._lib.load().trt_abi_version() == 1


def test_wire_struct_sizes(trt):
    # #[repr(C)] SamplePoint 32 B (pointgen.rs:7-13) and SampledColor 20 B (imager.rs:9-15)
    assert C.sizeof(trt._lib.SamplePoint) == 32 and C.sizeof(trt._lib.SampledColor) == 20
    assert trt.SAMPLE_POINT_DTYPE.itemsize == 32 and trt.SAMPLED_COLOR_DTYPE.itemsize == 20


def test_camera_pod_equals_oracle_bitwise(trt, orc):
    for args in ((140.0, 0.6, (50, 50, -140), (50, 50, 0), (0, 1, 0), 40.0, 1920, 1080),
                 (10.0, 0.6, (13, 2, 3), (0, 0, 0), (0, 1, 0), 20.0, 1200, 675),
                 (3.4, 10.0, (-2, 2, 1), (0, 0, -1), (0, 1, 0), 20.0, 400, 300)):
        a, b = trt.make_camera(*args), orc.make_camera(*args)
        assert bytes(a) == bytes(b)


@pytest.mark.parametrize("kind", [0, 1])
def test_host_bvh_build_and_ranks(trt, orc, kind):
    for desc in (trt.scenes.cornell(), trt.scenes.final_scene(), trt.scenes.random_mixed(500, 2), trt.scenes.random_spheres(30000, 64, 64)):
        sc = trt.Scene.from_desc(desc, kind)
        info = sc.bvh_info()
        assert info["nodes"] == desc.n_prims - 1 and 0 < info["max_depth"] <= 64
        assert np.array_equal(sc.leaf_ranks(), orc.Scene.from_desc(desc).leaf_ranks())   # reference-topology leaf order
        if kind == 0:
            assert info["max_depth"] == orc.Scene.from_desc(desc).tree_stats()["max_depth"]
    sah = trt.Scene.from_desc(trt.scenes.final_scene(), 1).bvh_info()["sah_cost"]
    med = trt.Scene.from_desc(trt.scenes.final_scene(), 0).bvh_info()["sah_cost"]
    assert sah < med


def test_errors_are_values_not_crashes(trt):
    lib = trt._lib.load()
    sc = trt.Scene()
    with pytest.raises(trt.TrtError) as e:
        sc.add_sphere((0, 0, 0), 1.0, 7)                      # unknown material
    assert e.value.code == -1 and "material" in str(e.value)
    with pytest.raises(trt.TrtError):
        sc.bvh_info()                                          # not committed
    with pytest.raises(trt.TrtError):
        sc.commit(5)
    m = sc.add_material(1, (1, 1, 1), 9.0)                     # metal fuzz clamps silently like Metal::new
    sc.add_sphere((0, 0, 0), 1.0, m); sc.commit(1)
    assert sc.bvh_info()["nodes"] == 1 and sc.num_prims == 1
    assert lib.trt_scene_destroy(None) == 0 and lib.trt_ctx_destroy(None) == 0
    assert lib.trt_scene_commit(None, 0) == -1


def test_no_cpu_fallback(trt):
    """Without a usable GPU every compute entry point must fail loudly (skipped where a GPU exists)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(trt.TrtError) as e:
        trt.Context(0)
    assert e.value.code == -2 and "no CPU path" in str(e.value)


def test_product_does_not_reference_the_oracle():
    """The product path must never route through oracle/ (only tests, smoke() and bench.py's CPU legs may)."""
    pkg = os.path.join(ROOT, "tiny-raytracer_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".cuh", ".h", ".hpp", ".sh")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle/" not in src.replace("oracle/pyoracle.py", "").replace("`oracle/", "") or f == "scenes.py", f
                assert "import oracle" not in src and "from oracle" not in src and "liboracle" not in src, f
    ldd = subprocess.run(["ldd", os.path.join(pkg, "libtrt_b200.so")], capture_output=True, text=True).stdout
    assert "oracle" not in ldd


def test_reference_api_mirror(trt, tmp_path):
    w = trt.World.new()
    w.add_material("white", trt.Lambertian(trt.Vec3.new(0.73, 0.73, 0.73)))
    with pytest.raises(KeyError):
        w.add_material("white", trt.Light(trt.Vec3.new_diagonal(1.0)))      # world.rs:29-31 panics on a duplicate name
    assert w.get_material("nope") is None
    w.add_geometry(trt.Sphere(trt.Vec3.new(0, 0, 0), 1.0, w.get_material("white")))
    w.add_geometry(trt.new_box(trt.Vec3.new(25, 0, 50), trt.Vec3.new(55, 60, 80), w.get_material("white")))
    assert len(w.get_geometries(trt.Quad)) == 6 and len(w.get_geometries(trt.Sphere)) == 1
    d = w.to_desc()
    ref = trt.scenes.cornell()
    assert np.array_equal(d.prim_data[1:7], ref.prim_data[6:12])               # new_box order == main.rs:99-122
    cam = trt.Camera.new(1.0, 10.0, trt.Vec3.zero(), trt.Vec3.new(0, 0, 1), trt.Vec3.new(0, 1, 0), 90.0, 16, 9)
    assert cam.get_image_size() == (16, 9)
    img = trt.Image(np.zeros((4, 5, 3), np.float32), (np.arange(60).reshape(4, 5, 3) * 4).astype(np.uint8))
    p = tmp_path / "x.png"; img.save(str(p))
    raw = p.read_bytes()
    assert raw[:8] == b"\x89PNG\r\n\x1a\n" and img.size() == (5, 4)


def test_rust_bindings_name_every_symbol():
    """rust/trt-sys/src/lib.rs is source only (no Rust toolchain here); keep it in lock-step with include/trt.h."""
    hdr = open(os.path.join(ROOT, "include", "trt.h")).read()
    declared = set(re.findall(r"\b(trt_[a-z0-9_]+)\s*\(", hdr))
    rs = open(os.path.join(ROOT, "rust", "trt-sys", "src", "lib.rs")).read()
    bound = set(re.findall(r"pub fn (trt_[a-z0-9_]+)\s*\(", rs))
    assert declared == bound, declared ^ bound


def test_flag_constants_agree_across_bindings(trt):
    """trt_params.flags switches (include/trt.h) as seen from Python (_lib.py) and Rust (trt-sys): same bits everywhere, and no
    two single-bit switches share a bit with each other or with the two multi-bit fields (regen_min: bits 8..15, step: 16..17)."""
    import re
    hdr = open(os.path.join(ROOT, "include", "trt.h")).read()
    rs = open(os.path.join(ROOT, "rust", "trt-sys", "src", "lib.rs")).read()
    c_flags = {m.group(1): int(m.group(2), 0) for m in re.finditer(r"#define\s+TRT_FLAG_([A-Z0-9_]+)\s+(0x[0-9a-fA-F]+|\d+)u", hdr)}
    assert {"COUNT_EVENTS", "NO_TAIL_SPLIT", "BINARY_NODES", "COUNT_WALKED", "INLINE_PRIMARY", "PRIMARY_PASS", "NO_WIDE8"} <= set(c_flags), c_flags
    for name, bits in c_flags.items():
        assert bits and bits & (bits - 1) == 0, (name, bits)                      # one bit each
        assert not bits & 0x3FF00, (name, bits)                                   # outside the regen_min and step fields
        assert getattr(trt._lib, "FLAG_" + name) == bits, name
        m = re.search(r"pub const TRT_FLAG_%s: u32 = (0x[0-9a-fA-F]+|\d+);" % name, rs)
        assert m and int(m.group(1), 0) == bits, name
    assert len(set(c_flags.values())) == len(c_flags)
    assert trt._lib.flag_step(3) == 3 << 16 and "((n) & 3u) << 16" in hdr and "(n & 3) << 16" in rs


def test_rust_build_lists_every_translation_unit():
    """rust/trt-sys/build.rs had its own source list and missed trt_lbvh.cu. Both builds now read
    csrc/sources.txt; that file must name every .cu / .cpp under csrc/ and neither build script may name one itself."""
    csrc = os.path.join(ROOT, "tiny-raytracer_b200", "csrc")
    listed = [ln.split()[0] for ln in open(os.path.join(csrc, "sources.txt")) if ln.strip() and not ln.startswith("#")]
    on_disk = sorted(f for f in os.listdir(csrc) if f.endswith((".cu", ".cpp")))
    assert sorted(listed) == on_disk, (listed, on_disk)
    for script in (os.path.join(csrc, "build.sh"), os.path.join(ROOT, "rust", "trt-sys", "build.rs")):
        src = open(script).read()
        assert "sources.txt" in src, script
        assert not re.search(r"trt_[a-z]+\.(cu|cpp)\b", src), "%s names a translation unit itself" % script
    rs = open(os.path.join(ROOT, "rust", "raytracer-cuda", "src", "lib.rs")).read()
    assert "geometries_in_order()" in rs and "get_geometries::<Quad>()" not in rs      # insertion order across types is kept


def test_multi_gpu_entry_points_fail_as_values(trt):
    """trt_multi_* / trt_render_multi argument errors are status codes; without a GPU creation fails loudly (no CPU path)."""
    import torch
    lib = trt._lib.load()
    h = C.c_void_p()
    assert lib.trt_multi_create(None, 1, C.byref(h)) == -1
    arr = (C.c_int * 20)(*range(20))
    assert lib.trt_multi_create(arr, 0, C.byref(h)) == -1 and lib.trt_multi_create(arr, 17, C.byref(h)) == -1
    assert lib.trt_multi_destroy(None) == 0
    if not torch.cuda.is_available():
        assert lib.trt_multi_create(arr, 2, C.byref(h)) == -2 and b"no CPU path" in lib.trt_last_error()
        sc = trt.Scene.from_desc(trt.scenes.cornell(8, 8, 1, 2), trt.BVH_AUTO)
        with pytest.raises(trt.TrtError) as e:
            sc.render_multi([0, 1], trt.make_camera(**trt.scenes.cornell(8, 8, 1, 2).camera), trt.make_params(8, 8, 1, 2))
        assert e.value.code == -2


def test_cpp_mirror_compiles_and_fails_loudly_without_gpu(trt, tmp_path):
    """include/trt.hpp + examples/cornell.cpp (main.rs translated) compile and link against the C ABI."""
    exe = str(tmp_path / "cornell")
    pkg = os.path.join(ROOT, "tiny-raytracer_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "cornell.cpp"), "-o", exe, "-L" + pkg, "-ltrt_b200", "-Wl,-rpath," + pkg])
    exe2 = str(tmp_path / "cornell_multi")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "cornell_multi.cpp"), "-o", exe2, "-L" + pkg, "-ltrt_b200", "-Wl,-rpath," + pkg])
    import torch
    if not torch.cuda.is_available():
        r = subprocess.run([exe, str(tmp_path / "o.png")], capture_output=True, text=True)
        assert r.returncode == 1 and "no CPU path" in r.stderr
        r = subprocess.run([exe2, "2"], capture_output=True, text=True)
        assert r.returncode == 1 and "no CPU path" in r.stderr


def test_cpp_image_save_writes_a_valid_png(tmp_path):
    """trt::Image::save (utils/image.rs:66-90 writes a PNG): signature, chunk CRCs, zlib stream and pixels round-trip."""
    import struct, zlib
    src = tmp_path / "png.cpp"
    src.write_text('#include "trt.hpp"\nint main(int, char** v) { trt::Image im; im.width = 301; im.height = 219;\n'
                   '  im.rgb8.resize(301 * 219 * 3); for (size_t i = 0; i < im.rgb8.size(); ++i) im.rgb8[i] = (uint8_t)((i * 7) ^ (i >> 8));\n'
                   '  im.save(v[1]); return 0; }\n')
    exe, out, pkg = str(tmp_path / "png"), str(tmp_path / "o.png"), os.path.join(ROOT, "tiny-raytracer_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-I" + os.path.join(ROOT, "include"), str(src), "-o", exe,
                           "-L" + pkg, "-ltrt_b200", "-Wl,-rpath," + pkg])
    subprocess.check_call([exe, out])
    b = open(out, "rb").read()
    assert b[:8] == b"\x89PNG\r\n\x1a\n"
    pos, idat, size = 8, b"", None
    while pos < len(b):
        n, = struct.unpack(">I", b[pos:pos + 4]); tag = b[pos + 4:pos + 8]; data = b[pos + 8:pos + 8 + n]
        assert zlib.crc32(tag + data) == struct.unpack(">I", b[pos + 8 + n:pos + 12 + n])[0]
        if tag == b"IHDR": size = struct.unpack(">IIBBBBB", data)
        if tag == b"IDAT": idat += data
        pos += 12 + n
    assert size == (301, 219, 8, 2, 0, 0, 0)
    rows = np.frombuffer(zlib.decompress(idat), np.uint8).reshape(219, 1 + 3 * 301)
    i = np.arange(301 * 219 * 3)
    assert (rows[:, 0] == 0).all() and np.array_equal(rows[:, 1:].reshape(-1), ((i * 7) ^ (i >> 8)).astype(np.uint8))


def test_header_is_plain_c99_and_host_calls_work_from_c(tmp_path):
    """include/trt.h compiles as strict C99 (`-pedantic -Werror`) and the host-only half of the ABI (scene, commit, camera,
    error values) runs from a C program: examples/c_abi_host.c."""
    exe, pkg = str(tmp_path / "c_abi_host"), os.path.join(ROOT, "tiny-raytracer_b200")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "c_abi_host.c"), "-o", exe, "-L" + pkg, "-ltrt_b200", "-Wl,-rpath," + pkg])
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and "2 prims, 1 node(s), sample point 32 bytes" in r.stdout and "unknown material id" in r.stdout, r.stdout + r.stderr



def test_transform_kats_and_instancing(trt):
    """math/transform.rs:113-167 — the reference's own tests (translation, scale, the three rotations, the complex transform),
    with Vec3's approximate equality (|delta| < 1e-7 per component, vec3.rs:189-201 + :32-35) — on the Python mirror; plus
    instancing: a transformed new_box equals the box built in place, and a sphere refuses a non-similarity."""
    T, V = trt.Transform, trt.Vec3
    near = lambda a, b: all(abs(x - y) < 1e-7 for x, y in zip(a, b))
    rad = lambda d: float(np.float32(d) * (np.float32(np.pi) / np.float32(180.0)))        # f32::to_radians
    delta = V.new(1.0, -2.0, 3.0)
    for v in (V.zero(), V.new(10.0, 15.0, 20.0)):
        assert near(T.translate(delta) * v, (v.x + 1.0, v.y - 2.0, v.z + 3.0))
        assert near(T.scale(delta) * v, (v.x * 1.0, v.y * -2.0, v.z * 3.0))
    v1, v2 = V.new(1.0, 0.0, 0.0), V.new(0.0, 1.0, 1.0)
    r1, r2, r3 = T.rotate_x(rad(90)), T.rotate_y(rad(90)), T.rotate_z(rad(90))
    assert near(r1 * V.zero(), (0, 0, 0)) and near(r1 * v1, (1, 0, 0)) and near(r2 * v1, (0, 0, -1)) and near(r3 * v1, (0, 1, 0))
    assert near(r1 * v2, (0, -1, 1)) and near(r2 * v2, (1, 1, 0)) and near(r3 * v2, (-1, 0, 1))
    t = T.new(V.new(1.0, 2.0, 3.0), V.new(rad(90), rad(-90), rad(-90)), V.new(2.0, 3.0, 4.0))
    assert near(t * V.zero(), (1, 2, 3))
    got = t * V.new(1.0, 1.0, 0.0)
    assert all(abs(a - b) < 1e-5 for a, b in zip(got, (1.0, 5.0, 7.0)))       # three chained f32 rotations: 1e-6 off, the reference's 1e-7 test is for its own libm
    white = trt.Lambertian((0.73, 0.73, 0.73))
    w = trt.World.new()
    w.add_instance(trt.new_box((0, 0, 0), (30, 60, 30), white), T.translate((25, 0, 50)))
    d = w.to_desc()
    ref = trt.scenes.cornell()
    assert np.array_equal(d.prim_data, ref.prim_data[6:12])                    # == new_box((25,0,50), (55,60,80)) of main.rs:66-70
    s = T.new((1, 2, 3), (0.3, -0.2, 0.9), (2, 2, 2)).apply(trt.Sphere((1, 0, 0), 0.5, white))
    assert abs(s.radius - 1.0) < 1e-6
    with pytest.raises(ValueError):
        T.scale((1, 2, 1)).apply(trt.Sphere((0, 0, 0), 1.0, white))


def test_cpp_transform_mirror(tmp_path):
    """The C++ mirror of math/transform.rs (include/trt.hpp): the reference's complex_transform test and box instancing."""
    src = tmp_path / "t.cpp"
    src.write_text('#include "trt.hpp"\n#include <cstdio>\nint main() { using namespace trt; const float r = 90.0f * (3.14159265358979323846f / 180.0f);\n'
                   '  Transform t = Transform::make({1, 2, 3}, {r, -r, -r}, {2, 3, 4}); Vec3 a = t * Vec3(0, 0, 0), b = t * Vec3(1, 1, 0);\n'
                   '  World w; auto m = Lambertian({0.5f, 0.5f, 0.5f}); w.add_instance(new_box({0, 0, 0}, {30, 60, 30}, m), Transform::translate({25, 0, 50}));\n'
                   '  const Geometry& g = w.geometries()[0]; bool thrown = false; try { Transform::scale({1, 2, 1}).apply(Sphere({0, 0, 0}, 1, m)); } catch (const Error&) { thrown = true; }\n'
                   '  std::printf("%.5f %.5f %.5f | %.5f %.5f %.5f | %zu %.1f %.1f %.1f | %d\\n", a.x, a.y, a.z, b.x, b.y, b.z, w.geometries().size(), g.a.x, g.a.y, g.a.z, (int)thrown); return 0; }\n')
    exe, pkg = str(tmp_path / "t"), os.path.join(ROOT, "tiny-raytracer_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-I" + os.path.join(ROOT, "include"), str(src), "-o", exe, "-L" + pkg, "-ltrt_b200", "-Wl,-rpath," + pkg])
    out = subprocess.run([exe], capture_output=True, text=True).stdout.split("|")
    a = [float(x) for x in out[0].split()]; b = [float(x) for x in out[1].split()]; c = out[2].split()
    assert all(abs(x - y) < 1e-4 for x, y in zip(a, (1, 2, 3))) and all(abs(x - y) < 1e-4 for x, y in zip(b, (1, 5, 7)))
    assert c == ["6", "25.0", "0.0", "80.0"] and out[3].strip() == "1"         # first face of the box: corner (25, 0, 80), main.rs:99-103
