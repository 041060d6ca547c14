"""-m gpu: the pipeline states, attachment types and command sequences that the everyday scenes never reach,
each against the compiled, unmodified reference, bit for bit:

* the Envmap program's extra dispatch entries (src/viewer/shaders_envmap.cxx:18-31): 0x22 depth pre-pass
  (no colour write), 0x5a depth EQUAL without depth write + alpha blend, 0x50 no depth test at all
* a non-default Viewport (GPU::DSDO, rglv_gpu.hxx:265-270)
* RB_RGBAF32 colour attachment (rglv_gpu.cxx:296-344, :345-395): clear + every store kind
* separate colour / depth clears on RB_RGBF32 + RB_F32
* the reference's fill-rule known-answer test (rglv_triangle.t.cxx:205-228) through the whole CUDA path
* several stores of one kind in one frame, with draws in between
* submission order inside long tile lists that mix clip-fan triangles with unclipped ones
"""
import numpy as np
import pytest

import rsr_b200 as R
from parity import assert_identical, render_both
from rsr_b200 import scenes
from rsr_b200.scenes import SoupScene

pytestmark = pytest.mark.gpu


class EnvmapPasses:
    """depth pre-pass (0x22), then the same geometry with depth EQUAL, no depth write, alpha blend (0x5a), then a
    second soup without any depth test (0x50) -- optionally with ordinary 0x62 / 0x72 draws around them"""

    def __init__(self, n=400, seed=5, mix=True, near_cross=True):
        self.a = SoupScene(n=n, seed=seed, program=R.PROGRAM_ENVMAP, near_cross=near_cross)
        self.b = SoupScene(n=n // 2, seed=seed + 1, program=R.PROGRAM_ENVMAP, near_cross=near_cross)
        self.c = SoupScene(n=n // 2, seed=seed + 2, program=R.PROGRAM_AMY, near_cross=near_cross)
        self.mix = mix

    def record(self, gl, size, out, depth=None, tile_blocks=(8, 8), passes=("prepass", "equal", "nodepth")):
        scenes.begin(gl, size, clear=(0.3, 0.1, 0.2), tile_blocks=tile_blocks)
        if self.mix:
            self.c.draw(gl, size, color_write=True, depth_write=True, depth_test=True, depth_func=R.GL_LESS, blend=False)
        if "prepass" in passes:
            self.a.draw(gl, size, color_write=False, depth_write=True, depth_test=True, depth_func=R.GL_LESS, blend=False)
        if "equal" in passes:
            self.a.draw(gl, size, color_write=True, depth_write=False, depth_test=True, depth_func=R.GL_EQUAL, blend=True)
        if "nodepth" in passes:
            self.b.draw(gl, size, color_write=True, depth_write=False, depth_test=False, blend=True)
        if self.mix:
            self.c.draw(gl, size, color_write=True, depth_write=True, depth_test=True, depth_func=R.GL_LESS, blend=True)
        gl.ColorWriteMask(True); gl.DepthWriteMask(True); gl.Enable(R.GL_DEPTH_TEST); gl.DepthFunc(R.GL_LESS)
        scenes.finish(gl, out, True, depth)


@pytest.mark.parametrize("size", [(640, 360), (1920, 1080)])
@pytest.mark.parametrize("mix", [False, True])
def test_envmap_prepass_equal_and_no_depth_states(size, mix, ref_gpu, cuda_gpu):
    outs = render_both(EnvmapPasses(mix=mix), size, ref_gpu, cuda_gpu)
    assert np.unique(outs["ref"][0]).size > 500
    assert_identical(outs, f"envmap 0x22 -> 0x5a -> 0x50 (mix={mix})")


@pytest.mark.parametrize("passes", [("prepass",), ("prepass", "equal"), ("nodepth",), ("equal",)])
def test_envmap_states_one_by_one(passes, ref_gpu, cuda_gpu):
    """each state key on its own: 0x22 alone leaves the clear colour everywhere (and, with a following ordinary draw,
    its depth); 0x5a after a clear passes nowhere (depth != 1.0) except where fragments sit exactly on the far plane"""
    outs = render_both(EnvmapPasses(mix=("equal",) == passes), (640, 360), ref_gpu, cuda_gpu, passes=passes)
    assert_identical(outs, f"envmap passes {passes}")


def test_envmap_states_in_long_lists(ref_gpu, cuda_gpu):
    """the same states with enough small triangles per tile for several raster batches and the queued rasteriser"""
    class Dense(EnvmapPasses):
        def __init__(self):
            self.a = SoupScene(n=4000, seed=8, program=R.PROGRAM_ENVMAP, near_cross=False, tiny=True, spread=0.6)
            self.b = SoupScene(n=3000, seed=9, program=R.PROGRAM_ENVMAP, near_cross=False, tiny=True, spread=0.6)
            self.c = SoupScene(n=2000, seed=10, program=R.PROGRAM_AMY, near_cross=False, tiny=True, spread=0.6)
            self.mix = True
    assert_identical(render_both(Dense(), (640, 360), ref_gpu, cuda_gpu), "dense envmap passes")


def test_depth_func_without_a_dispatch_entry_is_an_error(cuda_gpu):
    """LEQUAL (key 0x66) is installed for no program (src/viewer/shaders.cxx:54-126): the reference exits,
    the library answers RSRCU_ERR_NO_PROGRAM"""
    out = np.zeros((360, 640), np.uint32)
    SoupScene(n=10, seed=1, depth_func=R.GL_LEQUAL).record(cuda_gpu, (640, 360), out)
    with pytest.raises(R.RsrError) as e:
        cuda_gpu.Run()
    assert e.value.code == 4
    # ... and so is a draw into RB_RGBAF32 (key 0x762)
    class Rgba:
        def record(self, gl, size, out):
            gl.Reset(size, (8, 8))
            gl.RenderbufferType(R.GL_COLOR_ATTACHMENT0, R.RB_RGBAF32)
            gl.RenderbufferType(R.GL_DEPTH_ATTACHMENT, R.RB_F32)
            gl.Clear(R.GL_COLOR_BUFFER_BIT | R.GL_DEPTH_BUFFER_BIT)
            SoupScene(n=10, seed=1).draw(gl, size)
            scenes.finish(gl, out)
    Rgba().record(cuda_gpu, (640, 360), out)
    with pytest.raises(R.RsrError) as e:
        cuda_gpu.Run()
    assert e.value.code == 4


@pytest.mark.parametrize("viewport", [(100, 40, 400, 250), (0, 0, 320, 180), (37, 11, 333, 201), (-60, -30, 800, 500)])
def test_non_default_viewport(viewport, ref_gpu, cuda_gpu):
    """DS/DO with integer halves of odd sizes and an origin; triangles outside the viewport but inside the target
    are still drawn (there is no viewport clipping in the reference, only the guard band of the target)"""
    scene = SoupScene(n=500, seed=33)
    outs = render_both(scene, (640, 360), ref_gpu, cuda_gpu, with_depth=True, attachments="split", viewport=viewport)
    assert np.unique(outs["ref"][0]).size > 300
    assert_identical(outs, f"viewport {viewport}")


def _aligned(shape):
    n = int(np.prod(shape))
    raw = np.zeros(n + 4, np.float32)
    ofs = (-raw.ctypes.data // 4) % 4
    return raw[ofs:ofs + n].reshape(shape)


class AttachmentCommands:
    """no draws: clears and every store kind for a given pair of attachment types"""

    def __init__(self, color_type, separate_clears):
        self.color_type, self.separate = color_type, separate_clears

    def record(self, gl, size, out, fp, half, quads, depth):
        gl.Reset(size, (8, 8))
        gl.RenderbufferType(R.GL_COLOR_ATTACHMENT0, self.color_type)
        gl.RenderbufferType(R.GL_DEPTH_ATTACHMENT, R.RB_F32)
        gl.ClearColor((0.25, 0.5, 0.75))
        gl.ClearDepth(0.625)
        if self.separate:
            gl.Clear(R.GL_DEPTH_BUFFER_BIT)
            gl.ClearColor((0.125, 0.375, 0.0625))
            gl.Clear(R.GL_COLOR_BUFFER_BIT)
        else:
            gl.Clear(R.GL_COLOR_BUFFER_BIT | R.GL_DEPTH_BUFFER_BIT)
        gl.UseProgram(R.PROGRAM_DEFAULT_POST)
        gl.StoreDepth(depth)
        gl.StoreColor(fp)
        gl.StoreColorHalf(half)
        gl.StoreColorQuads(quads)
        gl.StoreColor(out, True)


@pytest.mark.parametrize("separate", [False, True])
@pytest.mark.parametrize("color_type", [R.RB_RGBAF32, R.RB_RGBF32])
def test_rgbaf32_and_rgbf32_attachment_commands(color_type, separate, ref_gpu, cuda_gpu):
    """RB_RGBAF32 has no draw program in the reference's tables, but its clear and store commands are implemented
    (rglv_gpu.cxx:311-405): the quad-swizzled store's fourth plane is the clear colour's alpha"""
    w, h = 248, 120
    got = {}
    for name, gl in (("ref", ref_gpu), ("cuda", cuda_gpu)):
        bufs = dict(out=np.zeros((h, w), np.uint32), fp=_aligned((h, w, 4)), half=_aligned((h // 2, w // 2, 4)),
                    quads=_aligned((h // 2, w // 2, 4, 4)), depth=_aligned((h, w)))
        for k in ("fp", "half", "quads", "depth"):
            bufs[k][:] = -1.0
        AttachmentCommands(color_type, separate).record(gl, (w, h), **bufs)
        gl.Run()
        got[name] = bufs
    for k in ("out", "fp", "half", "quads", "depth"):
        a, b = got["ref"][k], got["cuda"][k]
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), f"{k} store differs (colour type {color_type}, separate clears {separate})"
    assert np.all(got["ref"]["depth"] == 0.625)


KAT = ["........", ".XX.....", ".XXXX...", "..X.....", "........", "........", "........", "........"]


class KatTriangle:
    """the reference's own fill-rule test triangle (rglv_triangle.t.cxx:205-228): (2,4) (6,2) (1,1) in device
    pixels on an 8x8 target.  It reaches the rasteriser through the whole pipeline here -- vertex program,
    pdiv, viewport, trunc(16 x) -- so the vertices sit 1/32 px inside their 28.4 cells: 1/w is rcpps + one Newton
    step = 1 - 2^-24 for w = 1, which would otherwise pull (6.0, 2.0) just below the cell boundary."""

    def __init__(self, order=(0, 1, 2)):
        e = 1.0 / 32.0
        dev = np.array([[2.0, 4.0], [6.0, 2.0], [1.0, 1.0]], np.float64) + e
        ndc_x = (dev[:, 0] - 4.0) / 4.0
        ndc_y = (4.0 - dev[:, 1]) / 4.0
        pos = np.stack([ndc_x, ndc_y, np.zeros(3)]).astype(np.float32)
        self.pos, self.uv = scenes.soa(pos), scenes.soa(np.zeros((2, 3), np.float32))
        self.idx = np.array(order, np.uint16)
        self.tex = scenes.make_mipmap(np.ones((4, 4, 4), np.float32))

    def record(self, gl, size, out, depth=None):
        scenes.begin(gl, size, clear=(0.0, 0.0, 0.0))
        gl.UseProgram(R.PROGRAM_AMY)
        gl.ViewMatrix(np.eye(4, dtype=np.float32))
        gl.ProjectionMatrix(np.eye(4, dtype=np.float32))
        gl.UseBuffer(0, self.pos)
        gl.UseBuffer(9, self.uv)
        gl.BindTexture(0, self.tex, 4, 4, 4, R.GL_NEAREST_MIPMAP_NEAREST)
        gl.DrawElements(3, self.idx, 0)
        scenes.finish(gl, out, True, depth)


@pytest.mark.parametrize("order", [(0, 1, 2), (1, 2, 0), (2, 1, 0)])
def test_fill_rule_known_answer_through_the_tile_kernel(order, ref_gpu, cuda_gpu):
    """D3D top-left fill rule: exactly the 7 pixels of rglv_triangle.t.cxx:219-228 are lit -- by the tile kernel,
    and by the reference's own 4-wide rasteriser driven the same way (either winding: back faces are re-wound)"""
    outs = render_both(KatTriangle(order), (8, 8), ref_gpu, cuda_gpu)
    for name in ("cuda", "ref"):
        img = outs[name][0]
        text = ["".join("X" if img[y, x] != 0 else "." for x in range(8)) for y in range(8)]
        assert text == KAT, f"{name}: {text}"
    assert_identical(outs, "KAT")


def test_two_stores_of_one_kind_in_one_frame(ref_gpu, cuda_gpu):
    """StoreColor, more draws, StoreColor again: each destination receives the image at its own point of the
    frame (the reference writes a store when the tile reaches it, rglv_gpu.cxx:384-395)"""
    w, h = 640, 360
    s1, s2 = SoupScene(n=200, seed=101), SoupScene(n=200, seed=102, program=R.PROGRAM_OBJ2)
    got = {}
    for name, gl in (("ref", ref_gpu), ("cuda", cuda_gpu)):
        first, second = np.zeros((h, w), np.uint32), np.zeros((h, w), np.uint32)
        f1, f2 = _aligned((h, w, 4)), _aligned((h, w, 4))
        scenes.begin(gl, (w, h))
        s1.draw(gl, (w, h))
        gl.UseProgram(R.PROGRAM_DEFAULT_POST)
        gl.StoreColor(first, True)
        gl.StoreColor(f1)
        s2.draw(gl, (w, h))
        gl.UseProgram(R.PROGRAM_DEFAULT_POST)
        gl.StoreColor(f2)
        gl.StoreColor(second, False)
        gl.Run()
        got[name] = (first, second, f1, f2)
    assert np.count_nonzero(got["ref"][0] != got["ref"][1]) > 1000
    for k in range(4):
        assert np.array_equal(got["ref"][k].view(np.uint32), got["cuda"][k].view(np.uint32)), f"store {k} differs"


def test_device_truecolor_follows_the_frame_launched_last(cuda_gpu):
    """rsrcu_device_truecolor after a replay points at the retained frame's own target, not at the ring slot"""
    import ctypes as C
    g = R.GPU(0)
    try:
        outs, live, handles = [], [], []
        for seed in (5, 6):
            out = np.zeros((360, 640), np.uint32)
            live.append(out)      # a retained frame keeps writing its recorded host destination at every replay
            SoupScene(n=200, seed=seed).record(g, (640, 360), out)
            g.Run()
            outs.append(out.copy())
            handles.append(g.Retain())
        scenes.WavyGridScene(n=10).record(g, (1920, 1080), np.zeros((1080, 1920), np.uint32))   # a larger frame regrows the ring's buffers
        g.Run()
        cudart = C.CDLL("libcudart.so")
        for k in (0, 1, 0):
            g.Replay(handles[k], sync=True)
            ptr, stride = g.device_truecolor()
            assert ptr and stride == 640
            back = np.zeros((360, 640), np.uint32)
            assert cudart.cudaMemcpy(back.ctypes.data_as(C.c_void_p), C.c_void_p(ptr), C.c_size_t(back.nbytes), 2) == 0
            assert np.array_equal(back, outs[k]), f"device frame after replay {k} differs"
        for hd in handles:
            g.Release(hd)
    finally:
        g.close()


def test_clip_fans_inside_long_runs_keep_submission_order(ref_gpu, cuda_gpu):
    """thousands of tiny triangles straddling the near plane, all inside a handful of tiles, alpha blended: warps of
    the binning kernels hold unclipped triangles and clip fans side by side, the tiles' lists are long (> 256
    entries: run merge / key ranges).  The reference draws a draw's clipped triangles after all its unclipped ones
    (rglv_gpu_impl.hxx:498-508); blending makes any other order visible."""
    scene = SoupScene(n=6000, seed=71, tiny=True, spread=0.04, zrange=(0.30, 0.62), blend=True)
    outs = render_both(scene, (640, 360), ref_gpu, cuda_gpu)
    st = cuda_gpu.stats()
    assert st["triangles_clipped"] > 500, st
    assert st["bin_entries"] > 3000, st
    assert_identical(outs, "clip fans in long lists")
    # and with depth test LESS ties instead of blending, several list cells per tile
    scene2 = SoupScene(n=6000, seed=72, tiny=True, spread=0.04, zrange=(0.30, 0.62), program=R.PROGRAM_OBJ2)
    assert_identical(render_both(scene2, (640, 360), ref_gpu, cuda_gpu), "clip fans in long lists, depth ties")
