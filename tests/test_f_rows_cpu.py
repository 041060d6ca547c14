"""CPU-side checks for the SURVEY 8(f) rows: the oracle functions behave as the reference documents them, and the
constant tables the CUDA library carries are the reference's."""
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_packed_marching_cubes_table_is_the_references(refgl):
    """rsrcu.cu's kMcTriHost: 16 nibbles per corner-sign case, 0xF terminated -- equal to rglv::tritable, and the edges a
    case's triangles use are exactly the edges rglv::cube_edge_flags marks (the kernel derives one from the other)"""
    src = open(os.path.join(ROOT, "rsr_b200", "csrc", "rsrcu.cu")).read()
    body = src[src.index("kMcTriHost[256] = {"):]
    body = body[:body.index("};")]
    rows = [int(m, 16) for m in re.findall(r"0x([0-9a-f]{16})ull", body)]
    assert len(rows) == 256
    flags, tri, conn = refgl.mc_tables()
    for i, row in enumerate(rows):
        nibbles = [(row >> (4 * k)) & 0xF for k in range(16)]
        want = [int(e) & 0xF if e >= 0 else 0xF for e in tri[i]]
        assert nibbles == want, f"case {i}"
        edges = 0
        for e in tri[i]:
            if e < 0:
                break
            edges |= 1 << int(e)
        assert edges == int(flags[i]) & 0xFFF, f"case {i}: edge flags"
    assert conn.tolist() == [[0, 1], [1, 2], [2, 3], [3, 0], [4, 5], [5, 6], [6, 7], [7, 4], [0, 4], [1, 5], [2, 6], [3, 7]]


def test_reference_kawase_keeps_a_constant_canvas_and_clamps_at_the_border(refgl):
    c = np.full((20, 28, 4), 0.375, np.float32)
    for dist in (0, 1, 3, 30):
        assert np.array_equal(refgl.kawase_blur(c, dist), c)
    # one bright pixel: dist 0 spreads it over the 4x4 neighbourhood weights (1 2 1 ...)/16 of the four 2x2 boxes
    p = np.zeros((9, 9, 4), np.float32)
    p[4, 4] = 16.0
    out = refgl.kawase_blur(p, 0)[:, :, 0]
    assert out.sum() == 16.0 and out[4, 4] == 4.0 and out[3, 3] == 1.0 and out[2, 4] == 0.0


def test_reference_march_surface_lies_on_the_field(refgl):
    t, precision, fork, rng = 0.8, 32, 2, 5.0
    pos, nrm, blocks = refgl.march_surface(t, precision, fork, rng)
    assert blocks and all(first % 4 == 0 and n % 3 == 0 and n > 0 for first, n in blocks)
    used = np.zeros(pos.shape[1], bool)
    for first, n in blocks:
        used[first:first + n] = True
    x, y, z = (pos[k][used].astype(np.float64) for k in range(3))
    distort = 0.6 * np.sin(5.0 * (x + t / 4.0)) * np.sin(2.0 * (y + t / 1.33))
    field = np.sqrt(x * x + y * y + z * z) - 3.0 + (distort * np.sin(t / 2.0) + 1.0)
    delta = 2 * rng / precision
    assert np.abs(field).max() < 0.25 * delta   # linear interpolation along a cell edge
    n = np.stack([nrm[k][used] for k in range(3)])
    assert np.allclose((n * n).sum(0), 1.0, atol=1e-5)


def test_reference_span_bars(refgl):
    canvas = np.zeros((64, 200), np.uint32)
    refgl.render_spans(canvas, 10, 5, 1.0, [(0.1, 0.5, 0x1234, 0), (0.2, 0.3, 0x99, 2)])
    # lane 0: columns left + 20 .. left + 100, brightness 1 -> 0 (the last pixel is black); lane 2 sits 2 x (8 + 2) rows lower
    assert canvas[5:13, 30:110].all() and not canvas[13:15].any() and canvas[25:33, 50:69].all() and not canvas[:, :30].any()
    assert (canvas[5:13] == canvas[5]).all() and canvas[5, 30] > canvas[5, 60] > canvas[5, 100]


def test_colortest_lua_through_the_node_graph_is_the_golden_frame(refgl):
    """data/scene/colortest.lua -> JSON -> the reference's node graph -> its CPU rasteriser at 640x360 equals the frame
    tests/golden/colortest_c1.npz holds (made earlier from hand-recorded GL calls of the same scene): the C1 gate is the
    bundled scene itself"""
    import pytest
    if not refgl.scenes_available():
        pytest.skip("bundled scenes not built (oracle/build_ref.sh with /root/reference present)")
    from oracle import scene_ref
    g = np.load(os.path.join(ROOT, "tests", "golden", "colortest_c1.npz"))
    frame = scene_ref.frames("colortest", (640, 360), (0.0,))[0]   # (in a process of its own: oracle/scene_ref.py)
    assert np.array_equal(frame, g["frame"])


def test_bundled_scenes_compile_link_and_render(refgl):
    import pytest
    if not refgl.scenes_available():
        pytest.skip("bundled scenes not built")
    from oracle import scene_ref
    for name in refgl.BUNDLED_SCENES:
        a, b = scene_ref.frames(name, (320, 180), (0.5, 0.5))
        assert len(np.unique(a)) > 20 and np.array_equal(a, b), name
