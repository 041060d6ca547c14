"""the `$mc` node's marching cubes on the device (rsrcu_march_surface) against the reference's
rglv::march_sdf_vao driven by the node's block walk (oracle/ref_harness.cpp: ref_march_surface).

Tolerances: vertex POSITIONS, vertex order and the per-block table are bit-exact (integer / IEEE arithmetic restated in
the reference's order, its 4-wide sine approximation included).  NORMALS go through libm's sinf on the CPU and CUDA's
sinf on the GPU: |difference| <= 2e-5 per component.  A frame rendered from the device arrays therefore has the
reference's coverage and depth exactly and its colour within 1 LSB."""
import numpy as np
import pytest

import rsr_b200
from rsr_b200 import GL_COLOR_BUFFER_BIT, GL_DEPTH_BUFFER_BIT, PROGRAM_DEFAULT_POST, PROGRAM_OBJ2
from rsr_b200.scenes import perspective, rotate, translate

pytestmark = pytest.mark.gpu

NORMAL_TOL = 2e-5


def device_arrays(gpu, soa, total):
    return [gpu.read_device(p, total) for p in soa]


@pytest.mark.parametrize("t,precision,fork,rng", [(0.0, 32, 2, 5.0), (1.7, 16, 1, 5.0), (3.3, 32, 0, 4.0), (0.9, 64, 2, 5.0),
                                                  (2.4, 128, 2, 5.0), (5.1, 8, 0, 5.0), (0.4, 128, 4, 6.5)])
def test_march_surface_matches_the_reference(refgl, cuda_gpu, t, precision, fork, rng):
    pos, nrm, blocks = refgl.march_surface(t, precision, fork, rng)
    soa, dblocks, total = cuda_gpu.MarchSurface(t, precision, fork, rng)
    assert dblocks == blocks
    assert total == pos.shape[1] and total > 0
    got = device_arrays(cuda_gpu, soa, total)
    for k in range(3):
        assert np.array_equal(got[k].view(np.uint32), pos[k].view(np.uint32)), \
            f"position component {k}: {np.count_nonzero(got[k].view(np.uint32) != pos[k].view(np.uint32))} of {total} differ"
    err = max(float(np.abs(got[3 + k] - nrm[k]).max()) for k in range(3))
    assert err <= NORMAL_TOL, f"normals differ by {err}"
    # padding between blocks is zeros
    used = np.zeros(total, bool)
    for first, n in blocks:
        assert first % 4 == 0 and n % 3 == 0
        used[first:first + n] = True
    assert all(not got[k][~used].any() for k in range(6))


def test_far_away_surface_yields_nothing(cuda_gpu):
    soa, blocks, total = cuda_gpu.MarchSurface(0.0, 32, 2, 0.5)   # the sphere of radius ~4 lies outside [-0.5, 0.5]^3
    assert blocks == [] and total == 0


def test_bad_arguments_are_refused(cuda_gpu):
    for args in ((0.0, 33, 2, 5.0), (0.0, 32, 5, 5.0), (0.0, 32, 2, 0.0), (0.0, 128, 0, 5.0)):
        with pytest.raises(rsr_b200.RsrError):
            cuda_gpu.MarchSurface(*args)


def _frame(gl, size, draw_blocks, kd):
    w, h = size
    gl.Reset(size, (8, 8))
    gl.ClearColor((0.1, 0.1, 0.15))
    gl.ClearDepth(1.0)
    gl.Clear(GL_COLOR_BUFFER_BIT | GL_DEPTH_BUFFER_BIT)
    gl.UseProgram(PROGRAM_OBJ2)
    gl.ViewMatrix(translate(0, 0, -11) @ rotate(0.4, 0.3, 1.0, 0.2))
    gl.ProjectionMatrix(perspective(45.0, w / h, 1.0, 100.0))
    gl.UseUniforms(np.array([4.0, 6.0, 8.0, 0.0], np.float32))
    gl.UseBuffer(6, kd)   # (the `$mc` node's material binds nothing here; OBJ2's loader reads a colour array)
    draw_blocks(gl)
    gl.UseProgram(PROGRAM_DEFAULT_POST)


def test_frame_drawn_from_the_device_arrays(refgl, ref_gpu, cuda_gpu):
    """`$mc` Draw (mc.cxx:212-226): one DrawArrays per block with the position / normal arrays in slots 0 / 3 -- on the
    device the arrays are bound where the marching-cubes kernel wrote them, nothing is uploaded"""
    size = (640, 360)
    t = 1.25
    pos, nrm, blocks = refgl.march_surface(t, 32, 2, 5.0)
    from rsr_b200.scenes import soa
    kd = soa(np.tile(np.array([[0.9], [0.6], [0.3]], np.float32), (1, max(n for _, n in blocks) + 4)))

    def draw_ref(gl):
        for first, n in blocks:
            gl.UseBuffer(0, pos[:, first:first + ((n + 3) & ~3)])
            gl.UseBuffer(3, nrm[:, first:first + ((n + 3) & ~3)])
            gl.DrawArrays(n)
    # (depth of an RB_COLOR_DEPTH target: the fourth plane of the quad-swizzled store)
    want, want_quads = np.zeros((size[1], size[0]), np.uint32), np.zeros((size[1] // 2, size[0] // 2, 4, 4), np.float32)
    _frame(ref_gpu, size, draw_ref, kd)
    ref_gpu.StoreColorQuads(want_quads)
    ref_gpu.StoreColor(want, True)
    ref_gpu.Run()

    soa, dblocks, total = cuda_gpu.MarchSurface(t, 32, 2, 5.0)
    assert dblocks == blocks

    def draw_dev(gl):
        for first, n in dblocks:
            for k in range(3):
                gl.UseBufferDevice(k, soa[k] + 4 * first, n)
                gl.UseBufferDevice(3 + k, soa[3 + k] + 4 * first, n)
            gl.DrawArrays(n)
    got, got_quads = np.zeros_like(want), np.zeros_like(want_quads)
    _frame(cuda_gpu, size, draw_dev, kd)
    cuda_gpu.StoreColorQuads(got_quads)
    cuda_gpu.StoreColor(got, True)
    cuda_gpu.Run()
    st = cuda_gpu.stats()
    assert st["triangles_submitted"] == sum(n for _, n in blocks) // 3
    assert st["h2d_bytes"] < 64 * 1024 + kd.nbytes < 6 * 4 * total, st["h2d_bytes"]   # tables + the colour array: the surface's vertices did not cross PCIe
    got_depth, want_depth = got_quads[:, :, 3, :], want_quads[:, :, 3, :]
    assert np.array_equal(np.ascontiguousarray(got_depth).view(np.uint32), np.ascontiguousarray(want_depth).view(np.uint32)), "coverage / depth differ"
    ch = lambda a, s: ((a >> s) & 0xff).astype(np.int32)
    maxerr = max(int(np.abs(ch(got, s) - ch(want, s)).max()) for s in (0, 8, 16))
    assert maxerr <= 1, f"colour differs by {maxerr} LSB"
    assert np.count_nonzero(want_depth < 1.0) > 5000
