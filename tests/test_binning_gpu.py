"""-m gpu: the list-ordering machinery of the tile kernel against the compiled reference.

Tile lists are appended with atomics (any order) and every tile CTA restores submission order from
the 32-bit order keys: rank sort for short cells (mode A), run merge for long cells (B), key ranges +
bitonic sort when runs interleave (C), and a queue of large triangles that bypasses binning
(rsr_b200/csrc/kernels.cuh, tile_kernel.cuh).  Depth-LESS ties and blending make any ordering
mistake visible, so every case below is compared bit for bit with the reference's single-threaded,
in-order binner (rglv_gpu.cxx:119-260).
"""
import numpy as np
import pytest

import rsr_b200 as R
from parity import assert_identical, render_both
from rsr_b200 import scenes
from rsr_b200.scenes import SoupScene

pytestmark = pytest.mark.gpu


class DenseSoup:
    """several draws of tiny, overlapping triangles squeezed into a few tiles: lists of thousands of entries"""

    def __init__(self, draws=4, n=8000, seed=5, spread=0.25, near_cross=False, blend=False, program=R.PROGRAM_AMY):
        self.parts = [SoupScene(n=n, seed=seed + 17 * k, spread=spread, near_cross=near_cross, tiny=not near_cross,
                                blend=blend, program=program) for k in range(draws)]
        self.triangles = draws * n

    def record(self, gl, size, out, depth=None, **kw):
        # 32 px reference tiles: the reference's per-tile command buffer holds 100 000 bytes
        # (rglv_gpu.hxx kMaxSizeInBytes) and a longer list simply overruns it
        scenes.begin(gl, size, tile_blocks=(4, 4))
        for p in self.parts:
            p.draw(gl, size)
        scenes.finish(gl, out, depth=depth)


def _fresh_gpu(monkeypatch, **env):
    for k, v in env.items():
        monkeypatch.setenv(k, str(v))
    return R.GPU(0)


@pytest.mark.parametrize("blend", [False, True])
def test_long_lists(blend, ref_gpu, cuda_gpu):
    """thousands of entries per tile; blending makes the order visible in every pixel.  (In a frame with
    one cell per tile the triangles that straddle tiles are one-entry runs: too many of them for the run
    merge here, so these lists are ordered by key ranges.)"""
    sc = DenseSoup(draws=3, n=5000, spread=0.06, blend=blend)
    outs = render_both(sc, (640, 360), ref_gpu, cuda_gpu)
    st = cuda_gpu.stats()
    assert st["bin_entries"] > 8000 and st["list_chunks_run_merge"] + st["list_chunks_key_range"] > 0
    assert_identical(outs, f"dense soup blend={blend}")


def test_long_lists_with_clip_fans_key_ranges(ref_gpu, cuda_gpu):
    """clip fan triangles come after every unclipped triangle of their draw, so their runs interleave
    with the others': mode C (key ranges, bitonic sort)"""
    sc = DenseSoup(draws=3, n=3000, near_cross=True, spread=0.06, blend=True)
    outs = render_both(sc, (640, 360), ref_gpu, cuda_gpu)
    st = cuda_gpu.stats()
    assert st["triangles_clipped"] > 100 and st["list_chunks_key_range"] > 0
    assert_identical(outs, "dense soup with clipping")


@pytest.mark.parametrize("shift", [9, 11])
def test_multi_cell_lists(shift, ref_gpu, monkeypatch):
    """frames with millions of triangles split every list into cells by triangle index range; forced here
    on a small frame (RSRCU_GROUP_SHIFT): cell offsets (K3), ordered walk, fans in their draw's last cell"""
    g = _fresh_gpu(monkeypatch, RSRCU_GROUP_SHIFT=shift)
    try:
        for sc, what in ((DenseSoup(draws=3, n=5000, spread=0.06, blend=True), "dense"),
                         (DenseSoup(draws=3, n=3000, near_cross=True, spread=0.06, blend=True), "dense clipped"),
                         (scenes.BundledLikeScene(cubes=500), "c2")):
            assert_identical(render_both(sc, (640, 360), ref_gpu, g), f"{what}, cells of {1 << shift} triangles")
            if what == "dense":
                assert g.stats()["list_chunks_run_merge"] > 0      # ordered walk -> one run per warp and cell
    finally:
        g.close()


class Layers:
    """n full-screen (or nearly) alpha-blended quads over a soup of small triangles: every quad is two
    large triangles that cover all tiles"""

    def __init__(self, n=12, seed=9, soup=2000):
        rng = np.random.default_rng(seed)
        self.quads = []
        for k in range(n):
            z = -2.0 - 0.01 * k
            s = rng.uniform(1.5, 3.0)
            q = np.array([[-s, s, s, -s], [-s, -s, s, s], [z, z, z, z]], np.float32)
            self.quads.append((scenes.soa(q), scenes.soa(rng.uniform(0, 1, (2, 4)).astype(np.float32))))
        self.idx = np.array([0, 1, 2, 0, 2, 3], np.uint16)
        self.tex = scenes.make_mipmap(scenes.hash_texture(64, seed))
        self.soup = SoupScene(n=soup, seed=seed + 1, spread=1.2, near_cross=False, blend=True) if soup else None
        self.triangles = 2 * n + soup

    def record(self, gl, size, out, depth=None):
        scenes.begin(gl, size)
        if self.soup:
            self.soup.draw(gl, size)
        gl.UseProgram(R.PROGRAM_AMY)
        gl.Enable(R.GL_BLEND)
        gl.ViewMatrix(scenes.translate(0, 0, 0))
        gl.ProjectionMatrix(scenes.perspective(60.0, size[0] / size[1], 1, 20))
        gl.BindTexture(0, self.tex, 64, 64, 64, R.GL_LINEAR_MIPMAP_NEAREST)
        for pos, uv in self.quads:
            gl.UseBuffer(0, pos); gl.UseBuffer(9, uv)
            gl.DrawElements(6, self.idx, 0)
        gl.Disable(R.GL_BLEND)
        if self.soup:
            self.soup.draw(gl, size)
        scenes.finish(gl, out, depth=depth)


def test_large_triangle_queue(ref_gpu, cuda_gpu):
    """triangles that cover more than 32 tiles bypass binning (every tile scans the queue) and are merged
    into the tile's list by order key, between the small triangles drawn before and after them"""
    assert_identical(render_both(Layers(n=12), (640, 360), ref_gpu, cuda_gpu), "12 blended layers")
    assert_identical(render_both(Layers(n=12), (1920, 1080), ref_gpu, cuda_gpu), "12 blended layers 1080p")


def test_large_queue_per_tile_overflow_raises_the_threshold(ref_gpu, monkeypatch):
    """more queued triangles over one tile than a tile CTA can hold (128): the library raises the threshold
    (eventually nothing is 'large' any more) and launches the frame again inside rsrcu_sync -- the caller only
    ever sees the exact frame"""
    g = _fresh_gpu(monkeypatch)
    try:
        sc = Layers(n=90, soup=300)
        a = np.zeros((360, 640), np.uint32)
        sc.record(g, (640, 360), a)
        g.Run()
        assert g.stats()["frames_retried"] >= 1
        b = np.zeros_like(a)
        sc.record(ref_gpu, (640, 360), b)
        ref_gpu.Run()
        assert np.array_equal(a, b)
    finally:
        g.close()


def test_everything_through_the_queue(ref_gpu, monkeypatch):
    """RSRCU_LARGE_TILES=1: every triangle that touches two tiles is 'large'; long cells with queued
    items use key ranges (mode C)"""
    g = _fresh_gpu(monkeypatch, RSRCU_LARGE_TILES=1)
    try:
        assert_identical(render_both(SoupScene(n=600, seed=3, blend=True), (640, 360), ref_gpu, g), "soup via queue")
        assert_identical(render_both(scenes.CubesScene(instances=200), (640, 360), ref_gpu, g), "cubes via queue")
    finally:
        g.close()
