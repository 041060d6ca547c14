"""CPU tests (no GPU): pin the oracle.

1. the C restatement (oracle/restate.c) reproduces the reference's own known-answer test
   (rglv_triangle.t.cxx:205-228) and every golden frame under tests/golden/ (rendered by the
   unmodified reference, see make_golden.py) bit for bit;
2. where the compiled reference is present (oracle/_ref), restatement == reference on fresh seeded
   scenes, and the rcpps/rsqrtps table model == the instructions.
"""
import importlib.util
import os

import numpy as np
import pytest

import rsr_b200 as R
from oracle import restate
from rsr_b200 import scenes

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLDEN, "make_golden.py"))
make_golden = importlib.util.module_from_spec(spec)
spec.loader.exec_module(make_golden)

KAT = ["........", ".XX.....", ".XXXX...", "..X.....", "........", "........", "........", "........"]


def as_text(cov):
    return ["".join("X" if c else "." for c in row) for row in cov]


def test_fill_rule_known_answer_scalar_and_wide():
    """D3D top-left fill rule: tri (2,4)(6,2)(1,1) on 8x8 lights exactly 7 pixels"""
    g = np.load(os.path.join(GOLDEN, "kat_fill_rule.npz"))
    assert as_text(g["coverage"]) == KAT
    pts = g["points"].reshape(3, 2)
    xs = [int(16.0 * p[0]) for p in pts]
    ys = [int(16.0 * p[1]) for p in pts]
    for wide in (False, True):
        assert as_text(restate.raster_coverage(xs, ys, (0, 0, 8, 8), 8, 8, wide=wide)) == KAT


def test_fill_rule_shared_edge_no_double_hit_no_gap():
    """two triangles sharing an edge cover every pixel of a quad exactly once"""
    x = [3, 200, 200, 3]
    y = [5, 5, 130, 130]
    # front-facing winding for the rasteriser is the KAT's: (2,4) (6,2) (1,1)
    a = restate.raster_coverage([x[0] * 16, x[2] * 16, x[1] * 16], [y[0] * 16, y[2] * 16, y[1] * 16], (0, 0, 256, 144), 256, 144)
    b = restate.raster_coverage([x[0] * 16, x[3] * 16, x[2] * 16], [y[0] * 16, y[3] * 16, y[2] * 16], (0, 0, 256, 144), 256, 144)
    s = a.astype(int) + b.astype(int)
    assert s.max() == 1
    assert s[5:130, 3:200].min() == 1 and s.sum() == (130 - 5) * (200 - 3)


@pytest.mark.parametrize("name", sorted(make_golden.cases().keys()))
def test_restatement_reproduces_golden_frames(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    factory, kw = make_golden.cases()[name]
    size = tuple(int(v) for v in g["size"])
    rst = restate.RestateGPU(luts=(g["rcp"], g["rsqrt"]))
    out = np.zeros((size[1], size[0]), np.uint32)
    factory().record(rst, size, out, **kw)
    rst.Run()
    assert np.array_equal(out, g["frame"]), f"{np.count_nonzero(out != g['frame'])} pixels differ from the reference's frame"


def test_restatement_reproduces_the_bundled_colortest_scene():
    """SURVEY 8(d) C1 / BASELINE.json configs[0]: data/scene/colortest.lua (the reference's own OBJ loader, camera
    and renderer, tests/golden/make_bundled.py) at 640x360 -- the C restatement reproduces the committed frame"""
    from rsr_b200 import scenes
    g = np.load(os.path.join(GOLDEN, "colortest_c1.npz"))
    size = tuple(int(v) for v in g["size"])
    assert size == (640, 360) and g["idx"].size == 168 * 3
    rst = restate.RestateGPU(luts=(g["rcp"], g["rsqrt"]))
    out = np.zeros((size[1], size[0]), np.uint32)
    scenes.ColortestScene().record(rst, size, out)
    rst.Run()
    assert np.count_nonzero(g["frame"] != g["frame"][0, 0]) > 50000
    assert np.array_equal(out, g["frame"]), f"{np.count_nonzero(out != g['frame'])} pixels differ from the reference's frame"


def test_bundled_colortest_fixture_is_what_the_reference_loads(refgl):
    """(only where the reference tree is present) the committed vertex arrays / camera equal what the reference's
    LoadOBJ + MakeArray and LookAt / Perspective2 produce from data/mesh/colortest.obj today"""
    ref = os.environ.get("RSR_REFERENCE", "/root/reference")
    obj = os.path.join(ref, "data", "mesh", "colortest.obj")
    if not os.path.exists(obj):
        pytest.skip("no reference tree here")
    g = np.load(os.path.join(GOLDEN, "colortest_c1.npz"))
    (pos, nrm, kd), idx = refgl.load_obj_arrays(obj, "PND")
    assert np.array_equal(pos, g["pos"]) and np.array_equal(nrm, g["nrm"]) and np.array_equal(kd, g["kd"]) and np.array_equal(idx, g["idx"])
    view, proj = refgl.perspective_camera((88.0, 80.0, 93.0), 3.72, -0.35, 45.0, 640 / 360)
    assert np.array_equal(view, g["view"]) and np.array_equal(proj, g["proj"])


def test_lut_model_matches_golden_tables_on_same_cpu_family():
    """harvest is deterministic; (on another CPU family the tables may legitimately differ)"""
    a = restate.harvest_luts()
    b = restate.harvest_luts()
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert a[0][0] != 0 and a[1][1024] != 0


# ---- against the compiled reference (present in this container and shipped to the GPU box) --------

def test_rcp_rsqrt_model_vs_instructions(refgl):
    rng = np.random.default_rng(5)
    x = np.concatenate([rng.uniform(-1e6, 1e6, 200000), 10.0 ** rng.uniform(-30, 30, 200000),
                        np.array([1.0, 2.0, 0.5, 3.0, 1e-38, 1e38, 0.0, -0.0, np.inf, -np.inf])]).astype(np.float32)
    rcp, rsq = restate.harvest_luts()
    L = restate.lib()
    want_r = refgl.rcp(x).view(np.uint32)
    want_q = refgl.rsqrt(np.abs(x)).view(np.uint32)
    got_r = np.array([L.rst_rcp(float(v), rcp.ctypes.data) for v in x[:20000]], np.float32).view(np.uint32)
    got_q = np.array([L.rst_rsqrt(float(abs(v)), rsq.ctypes.data) for v in x[:20000]], np.float32).view(np.uint32)
    assert np.array_equal(got_r, want_r[:20000])
    assert np.array_equal(got_q, want_q[:20000])


def test_numpy_mipmap_equals_reference(refgl):
    base = scenes.hash_texture(64, 9)
    assert np.array_equal(scenes.make_mipmap(base), refgl.make_mipmap(base))


@pytest.mark.parametrize("case", ["grid", "grid_nearest", "cubes", "soup", "soup_many", "soup_obj2_cull", "soup_blend", "c2"])
def test_restatement_equals_compiled_reference(case, ref_gpu):
    sc, kw = {
        "grid": (scenes.WavyGridScene(n=30), {}),
        "grid_nearest": (scenes.WavyGridScene(n=30, bilinear=False), {}),
        "cubes": (scenes.CubesScene(instances=100), {}),
        "soup": (scenes.SoupScene(n=500, seed=13), {}),
        "soup_many": (scenes.SoupScene(n=100, seed=31, program=R.PROGRAM_MANY, instanced=5), {}),
        "soup_obj2_cull": (scenes.SoupScene(n=300, seed=15, program=R.PROGRAM_OBJ2, cull=R.GL_BACK), {}),
        "soup_blend": (scenes.SoupScene(n=300, seed=16, blend=True), {}),
        "c2": (scenes.BundledLikeScene(cubes=300), {"t": 0.5}),
    }[case]
    size = (640, 360)
    a = np.zeros((size[1], size[0]), np.uint32)
    b = np.zeros_like(a)
    sc.record(ref_gpu, size, a, **kw)
    ref_gpu.Run()
    rst = restate.RestateGPU()
    sc.record(rst, size, b, **kw)
    rst.Run()
    assert np.array_equal(a, b), f"{np.count_nonzero(a != b)} pixels differ"


def test_reference_is_thread_count_and_tile_size_independent(refgl):
    """SURVEY finding 2: the device may choose its own tile size"""
    sc = scenes.SoupScene(n=400, seed=17)
    frames = []
    g = refgl.RefGPU()
    for tb in ((8, 8), (4, 4), (2, 2), (16, 16)):
        out = np.zeros((360, 640), np.uint32)
        sc.record(g, (640, 360), out, tile_blocks=tb)
        g.Run()
        frames.append(out)
    g.close()
    for f in frames[1:]:
        assert np.array_equal(frames[0], f)
