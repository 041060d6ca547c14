"""-m gpu: the CUDA path (through the C ABI) against the compiled, unmodified reference renderer on
the same seeded inputs.

Tolerance (north_star): coverage and depth-test decisions identical -- zero differing pixels, depth
buffers compared bit for bit; shaded colour within 1 LSB of the 8-bit output.  These tests ask for
the stronger 0 LSB: with the host's rcpps/rsqrtps tables harvested, the arithmetic is the same.
"""
import numpy as np
import pytest

import rsr_b200 as R
from parity import assert_identical, render_both
from rsr_b200.scenes import BundledLikeScene, CubesScene, FillStressScene, GeometryStressScene, SoupScene, WavyGridScene

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cuda_direct():
    g = R.GPU(0, direct=True)
    yield g
    g.close()


@pytest.mark.parametrize("size", [(640, 360), (1920, 1080)])
@pytest.mark.parametrize("bilinear", [True, False])
def test_wavy_grid(size, bilinear, ref_gpu, cuda_gpu):
    scene = WavyGridScene(n=40, bilinear=bilinear)
    outs = render_both(scene, size, ref_gpu, cuda_gpu, with_depth=False)
    assert np.unique(outs["ref"][0]).size > 1000
    assert_identical(outs, "wavy grid")


def test_direct_abi_route(ref_gpu, cuda_direct):
    """every GL call through its own C-ABI entry point instead of the packed stream"""
    assert_identical(render_both(WavyGridScene(n=30), (640, 360), ref_gpu, cuda_direct), "direct ABI")
    assert_identical(render_both(CubesScene(instances=50), (640, 360), ref_gpu, cuda_direct), "direct ABI cubes")


@pytest.mark.parametrize("size", [(640, 360), (1920, 1080)])
def test_cubes_instanced_and_lit(size, ref_gpu, cuda_gpu):
    assert_identical(render_both(CubesScene(instances=300), size, ref_gpu, cuda_gpu), "cubes")


def test_c2_bundled_like_1080p(ref_gpu, cuda_gpu):
    scene = BundledLikeScene(cubes=600)
    for t in (0.0, 1.7):
        assert_identical(render_both(scene, (1920, 1080), ref_gpu, cuda_gpu, t=t), f"c2 t={t}")


@pytest.mark.parametrize("tile_blocks", [(8, 8), (4, 4), (16, 16), (2, 2), (3, 5)])
def test_reference_tile_size_does_not_matter(tile_blocks, ref_gpu, cuda_gpu):
    outs = render_both(WavyGridScene(n=20), (640, 360), ref_gpu, cuda_gpu, tile_blocks=tile_blocks)
    assert_identical(outs, f"tile blocks {tile_blocks}")


@pytest.mark.parametrize("seed", [3, 4, 5, 6])
def test_soup_clipping_near_plane_and_guard_band(seed, ref_gpu, cuda_gpu):
    """triangles crossing the near plane / guard band go through Sutherland-Hodgman"""
    scene = SoupScene(n=600, seed=seed)
    outs = render_both(scene, (640, 360), ref_gpu, cuda_gpu, with_depth=False)
    assert cuda_gpu.stats()["triangles_clipped"] > 10
    assert_identical(outs, f"soup seed {seed}")


@pytest.mark.parametrize("cull", [None, R.GL_BACK, R.GL_FRONT, R.GL_FRONT_AND_BACK])
def test_soup_culling_modes(cull, ref_gpu, cuda_gpu):
    assert_identical(render_both(SoupScene(n=400, seed=11, cull=cull), (640, 360), ref_gpu, cuda_gpu), f"cull {cull}")


@pytest.mark.parametrize("program", [R.PROGRAM_AMY, R.PROGRAM_OBJ1, R.PROGRAM_OBJ2, R.PROGRAM_OBJ2S, R.PROGRAM_DEPTH,
                                     R.PROGRAM_WIREFRAME, R.PROGRAM_ENVMAP, R.PROGRAM_PATTERN, R.PROGRAM_ALPHATEXTURE])
def test_programs(program, ref_gpu, cuda_gpu):
    scene = SoupScene(n=300, seed=21, program=program, near_cross=True)
    assert_identical(render_both(scene, (640, 360), ref_gpu, cuda_gpu), f"program {program}")


def test_many_instanced_soup(ref_gpu, cuda_gpu):
    scene = SoupScene(n=120, seed=31, program=R.PROGRAM_MANY, instanced=7)
    assert_identical(render_both(scene, (640, 360), ref_gpu, cuda_gpu), "instanced elements")
    assert_identical(render_both(scene, (640, 360), ref_gpu, cuda_gpu, arrays=True), "instanced arrays")


def test_draw_arrays(ref_gpu, cuda_gpu):
    assert_identical(render_both(SoupScene(n=300, seed=41), (640, 360), ref_gpu, cuda_gpu, arrays=True), "DrawArrays")


@pytest.mark.parametrize("program", [R.PROGRAM_AMY, R.PROGRAM_TEXT, R.PROGRAM_ALPHATEXTURE, R.PROGRAM_ENVMAP])
def test_alpha_blend(program, ref_gpu, cuda_gpu):
    scene = SoupScene(n=300, seed=51, program=program, blend=True)
    assert_identical(render_both(scene, (640, 360), ref_gpu, cuda_gpu), f"blend program {program}")


def test_split_attachments_and_depth_store(ref_gpu, cuda_gpu):
    """RB_RGBF32 + RB_F32 (state key 0x6e2), StoreDepth, StoreColor to float"""
    scene = SoupScene(n=400, seed=61)
    outs = render_both(scene, (640, 360), ref_gpu, cuda_gpu, with_depth=True, with_fp=True, attachments="split")
    assert_identical(outs, "split attachments")


def test_float_store_color_depth_attachment(ref_gpu, cuda_gpu):
    outs = render_both(SoupScene(n=300, seed=62), (640, 360), ref_gpu, cuda_gpu, with_fp=True)
    assert_identical(outs, "float store")


def test_linear_store_and_exposure_post(ref_gpu, cuda_gpu):
    scene = SoupScene(n=300, seed=71)
    assert_identical(render_both(scene, (640, 360), ref_gpu, cuda_gpu, gamma=False), "linear store")
    assert_identical(render_both(scene, (640, 360), ref_gpu, cuda_gpu, post=R.PROGRAM_EXPOSURE_POST, post_uniform=1.7), "exposure")


def test_tiny_triangles(ref_gpu, cuda_gpu):
    scene = SoupScene(n=5000, seed=81, tiny=True, near_cross=False, spread=1.0)
    assert_identical(render_both(scene, (640, 360), ref_gpu, cuda_gpu), "tiny")


@pytest.mark.parametrize("size", [(64, 64), (72, 40), (2048, 512), (1000, 600)])
def test_odd_target_sizes(size, ref_gpu, cuda_gpu):
    """partial edge tiles.  Widths stay multiples of 4: the reference's FilterTile stores 4 pixels at a
    time (rglr_algorithm.hxx:52-66) and writes out of bounds otherwise."""
    assert_identical(render_both(SoupScene(n=300, seed=91), size, ref_gpu, cuda_gpu), f"size {size}")


def test_empty_frame_and_empty_draw(ref_gpu, cuda_gpu):
    class Empty:
        def record(self, gl, size, out, depth=None):
            from rsr_b200.scenes import begin, finish
            begin(gl, size, clear=(0.5, 0.25, 0.125))
            finish(gl, out)
    assert_identical(render_both(Empty(), (640, 360), ref_gpu, cuda_gpu), "clear only")


def test_unknown_program_is_an_error(cuda_gpu):
    """the reference calls std::exit(1) for a missing dispatch entry (rglv_gpu.cxx:199-202)"""
    scene = SoupScene(n=10, seed=1, program=R.PROGRAM_TEXT)   # Text is installed with blending only
    out = np.zeros((360, 640), np.uint32)
    scene.record(cuda_gpu, (640, 360), out)
    with pytest.raises(R.RsrError) as e:
        cuda_gpu.Run()
    assert e.value.code == 4


@pytest.mark.parametrize("gamma", [False, True])
@pytest.mark.parametrize("tile_blocks", [(8, 8), (4, 4), (3, 5)])
def test_iq_post_program(gamma, tile_blocks, ref_gpu, cuda_gpu):
    """IQPostProgram (pow via sse_pow polynomials, vignette, sin-hash dither); its fragment coordinate is a
    running float sum that starts at each reference tile's edge, so the tile size matters here"""
    scene = SoupScene(n=300, seed=72)
    outs = render_both(scene, (640, 360), ref_gpu, cuda_gpu, gamma=gamma, post=R.PROGRAM_IQ_POST, tile_blocks=tile_blocks)
    assert_identical(outs, f"IQ post gamma={gamma} tiles={tile_blocks}")


def _aligned(shape):
    """float32 array whose data is 16-byte aligned (the reference writes it with streaming stores)"""
    n = int(np.prod(shape))
    raw = np.zeros(n + 4, np.float32)
    ofs = (-raw.ctypes.data // 4) % 4
    return raw[ofs:ofs + n].reshape(shape)


@pytest.mark.parametrize("attachments", [None, "split"])
@pytest.mark.parametrize("size", [(640, 360), (248, 120)])   # (the reference itself needs widths that are multiples of 4)
def test_half_size_and_quad_swizzled_float_stores(attachments, size, ref_gpu, cuda_gpu):
    """CMD_STORE_COLOR_HALF_LINEAR_FP (Downsample) and CMD_STORE_COLOR_FULL_QUADS_FP (quad-swizzled copy; the
    fourth plane holds depth for RB_COLOR_DEPTH and 1.0 for RB_RGBF32), bit for bit"""
    w, h = size
    scene = SoupScene(n=300, seed=77)
    got = {}
    for name, gl in (("ref", ref_gpu), ("cuda", cuda_gpu)):
        color = np.zeros((h, w), np.uint32)
        half = _aligned((h // 2, w // 2, 4))
        quads = _aligned((h // 2, w // 2, 4, 4))
        half[:] = -1.0
        quads[:] = -1.0
        scene.record(gl, size, color, half_out=half, quads_out=quads, attachments=attachments)
        gl.Run()
        got[name] = (color, half, quads)
    (rc, rh, rq), (cc, ch, cq) = got["ref"], got["cuda"]
    assert np.array_equal(rc, cc)
    assert np.unique(rh.view(np.uint32)).size > 100
    assert np.array_equal(rh.view(np.uint32), ch.view(np.uint32)), "half-size float store differs"
    assert np.array_equal(rq.view(np.uint32), cq.view(np.uint32)), "quad-swizzled float store differs"


def test_maximum_target_2048x2048(ref_gpu, cuda_gpu):
    """the largest target either renderer accepts (guard band ends at 2048, rglv_view_frustum.hxx:36-39): 64 x 64 device tiles,
    4096 CTAs, every tile-index field at its limit"""
    assert_identical(render_both(SoupScene(n=400, seed=17), (2048, 2048), ref_gpu, cuda_gpu, with_depth=True, attachments="split"), "2048x2048")
    assert_identical(render_both(CubesScene(instances=400), (2048, 2048), ref_gpu, cuda_gpu), "2048x2048 cubes")


def test_error_behaviour_matches_the_reference_contract(cuda_gpu):
    """what the reference treats as fatal (std::exit / assert) is an error code here, and the context stays usable"""
    from rsr_b200.scenes import begin, finish
    g = R.GPU(0, direct=True)
    try:
        # target beyond the guard band
        with pytest.raises(R.RsrError) as e:
            g.Reset((4096, 2160), (8, 8))
        assert e.value.code == 5   # RSRCU_ERR_UNSUPPORTED
        # odd dimensions cannot hold 2x2 quads
        with pytest.raises(R.RsrError):
            g.Reset((641, 360), (8, 8))
        # a stencil clear is not implemented (rglv_gpu.cxx:313-315)
        g.Reset((640, 360), (8, 8))
        g.ClearColor((0.0, 0.0, 0.0))
        with pytest.raises(R.RsrError):
            g.Clear(R.GL_STENCIL_BUFFER_BIT)
        # RB_COLOR_DEPTH needs colour and depth cleared together (rglv_gpu.cxx:317-319)
        with pytest.raises(R.RsrError):
            g.Clear(R.GL_COLOR_BUFFER_BIT)
        # a store canvas of the wrong size
        bad = np.zeros((100, 100), np.uint32)
        with pytest.raises(R.RsrError):
            g.StoreColor(bad)
    finally:
        g.close()
    # the shared context still renders
    out = np.zeros((360, 640), np.uint32)
    WavyGridScene(n=8).record(cuda_gpu, (640, 360), out)
    cuda_gpu.Run()
    assert np.unique(out).size > 10


def test_bench_size_fill_stress_matches_the_reference(ref_gpu, cuda_gpu):
    """the C4 workload exactly as bench.py renders it: 8 layers of 960x540 quads, 32 distinct 1024^2 mip-mapped
    textures sampled 1:1 with bilinear filtering, one 1920x1080 sub-frame"""
    scene = FillStressScene(layers=8, size=(1920, 1080), quads=(2, 2), tex_dim=1024)
    outs = render_both(scene, (1920, 1080), ref_gpu, cuda_gpu)
    assert np.unique(outs["ref"][0]).size > 100000
    assert_identical(outs, "c4 at bench size")
    assert cuda_gpu.stats()["fragments_shaded"] == 8 * 1920 * 1080


def test_bench_size_geometry_stress_matches_the_reference(ref_gpu, cuda_gpu):
    """the C3 workload exactly as bench.py renders it: 30 icospheres at 6 subdivisions = 2.46 M triangles of about a
    pixel, several list cells per tile (groups > 1), run-merged lists.  (Reference tiles of 4x4 blocks keep its
    unchecked 100 000-byte tile lists (rglv_gpu.hxx:29) within bounds; results do not depend on the tile size.)"""
    scene = GeometryStressScene(spheres=30, divs=6)
    outs = render_both(scene, (1920, 1080), ref_gpu, cuda_gpu, tile_blocks=(4, 4))
    st = cuda_gpu.stats()
    assert st["triangles_submitted"] == 2457600 and st["list_chunks_run_merge"] > 0, st
    assert_identical(outs, "c3 at bench size")
