"""The compiled drop-in: the reference's OWN `rglv::GL` / `GLState` / packed command stream, recorded by
the reference's own code, rendered by librsrcu.so through the RunImpl replacement a maintainer would add
(rsr_b200/host/rglv_gpu_cuda.cxx, compiled against the reference tree into oracle/_ref/librsr_dropin.so by
oracle/build_ref.sh) -- against the pure reference (oracle/_ref/librsr_ref.so), bit for bit.

Both libraries are driven through the same C harness (oracle/ref_harness.cpp), i.e. the same sequence of `GL::`
calls; they differ only in the body of `GPU::RunImpl` (rglv_gpu.cxx:90-116).  Buffer extents are discovered by the
binding itself with an index scan, like the reference's binner does (rglv_gpu_impl.hxx:332-334)."""
import os
import subprocess

import numpy as np
import pytest

import rsr_b200 as R
from oracle import refgl
from rsr_b200 import scenes
from rsr_b200.scenes import BundledLikeScene, CubesScene, SoupScene, WavyGridScene

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_dropin_library_replaces_runimpl_and_binds_the_c_abi():
    """not gpu: the library exists (built where the reference tree is), defines GPU::RunImpl itself and imports the
    rsrcu_* entry points it forwards to"""
    if not refgl.dropin_available():
        pytest.skip("oracle/_ref/librsr_dropin.so not built (needs /root/reference and librsrcu.so)")
    out = subprocess.run(["nm", "-DC", refgl.DROPIN_PATH], capture_output=True, text=True, check=True).stdout
    assert any(" T " in line and "rqdq::rglv::GPU::RunImpl(" in line for line in out.splitlines())
    for sym in ("rsrcu_create", "rsrcu_begin_frame", "rsrcu_set_state", "rsrcu_bind_buffer", "rsrcu_bind_texture", "rsrcu_clear",
                "rsrcu_draw_elements", "rsrcu_draw_arrays", "rsrcu_store_color_tc", "rsrcu_store_color_fp", "rsrcu_store_color_quads",
                "rsrcu_store_depth", "rsrcu_end_frame", "rsrcu_sync", "rsrcu_sync_frame"):
        assert any(line.strip().startswith("U " + sym) for line in out.splitlines()), sym
    src = open(os.path.join(ROOT, "rsr_b200", "host", "rglv_gpu_cuda.cxx")).read()
    assert "void GPU::RunImpl(rclmt::jobsys::Job* job)" in src


@pytest.fixture(scope="module")
def dropin_gpu():
    if not refgl.dropin_available():
        pytest.skip("oracle/_ref/librsr_dropin.so not built")
    g = refgl.RefGPU(dropin=True)
    yield g
    g.close()


def both(scene, size, ref_gpu, dropin_gpu, **kw):
    w, h = size
    outs = []
    for gl in (ref_gpu, dropin_gpu):
        color = np.zeros((h, w), np.uint32)
        scene.record(gl, size, color, **kw)
        gl.Run()
        outs.append(color)
    return outs


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["wavy", "cubes", "c2", "soup_clip", "soup_blend", "arrays", "instanced_arrays", "obj2s", "iq_post"])
def test_reference_gl_through_the_cuda_runimpl(case, ref_gpu, dropin_gpu):
    size = (640, 360)
    kw = {}
    if case == "wavy":
        scene, size = WavyGridScene(n=40), (1920, 1080)
    elif case == "cubes":
        scene = CubesScene(instances=300)
    elif case == "c2":
        scene, size, kw = BundledLikeScene(cubes=600), (1920, 1080), {"t": 0.7}
    elif case == "soup_clip":
        scene = SoupScene(n=600, seed=3)
    elif case == "soup_blend":
        scene = SoupScene(n=300, seed=51, program=R.PROGRAM_TEXT, blend=True)
    elif case == "arrays":
        scene, kw = SoupScene(n=300, seed=41), {"arrays": True}
    elif case == "instanced_arrays":
        scene, kw = SoupScene(n=120, seed=31, program=R.PROGRAM_MANY, instanced=7), {"arrays": True}
    elif case == "obj2s":
        scene = SoupScene(n=300, seed=21, program=R.PROGRAM_OBJ2S)
    else:
        scene, kw = SoupScene(n=300, seed=72), {"post": R.PROGRAM_IQ_POST}
    a, b = both(scene, size, ref_gpu, dropin_gpu, **kw)
    assert np.unique(a).size > 100
    assert np.array_equal(a, b), f"{case}: {np.count_nonzero(a != b)} pixels differ between the reference and the drop-in"


def _aligned(shape):
    n = int(np.prod(shape))
    raw = np.zeros(n + 4, np.float32)
    ofs = (-raw.ctypes.data // 4) % 4
    return raw[ofs:ofs + n].reshape(shape)


@pytest.mark.gpu
def test_every_store_command_through_the_drop_in(ref_gpu, dropin_gpu):
    """RB_RGBF32 + RB_F32: StoreDepth, StoreColor (float, half-size, quad-swizzled, true colour)"""
    w, h = 640, 360
    scene = SoupScene(n=400, seed=61)
    got = []
    for gl in (ref_gpu, dropin_gpu):
        bufs = dict(out=np.zeros((h, w), np.uint32), depth=_aligned((h, w)), fp_out=_aligned((h, w, 4)),
                    half_out=_aligned((h // 2, w // 2, 4)), quads_out=_aligned((h // 2, w // 2, 4, 4)))
        scene.record(gl, (w, h), bufs["out"], bufs["depth"], fp_out=bufs["fp_out"], half_out=bufs["half_out"],
                     quads_out=bufs["quads_out"], attachments="split")
        gl.Run()
        got.append(bufs)
    for k in got[0]:
        assert np.array_equal(got[0][k].view(np.uint32), got[1][k].view(np.uint32)), f"{k} differs"


@pytest.mark.gpu
def test_drop_in_in_double_buffer_mode_has_the_reference_latency(ref_gpu):
    """doubleBuffer (rglv_gpu.cxx:16,111-112): a frame's canvas is complete when the NEXT Run has finished -- the
    binding keeps that contract (frame N is on the GPU while frame N+1 is recorded)"""
    if not refgl.dropin_available():
        pytest.skip("oracle/_ref/librsr_dropin.so not built")
    g = refgl.RefGPU(dropin=True, double_buffer=True)
    try:
        scene = CubesScene(instances=200)
        outs = [np.zeros((360, 640), np.uint32) for _ in range(4)]
        for i in range(4):
            scene.record(g, (640, 360), outs[i], t=0.3 * i)
            g.Run()
        g.Flush()
        for i in range(4):
            want = np.zeros((360, 640), np.uint32)
            scene.record(ref_gpu, (640, 360), want, t=0.3 * i)
            ref_gpu.Run()
            assert np.array_equal(outs[i], want), f"frame {i}"
    finally:
        g.L.ref_set_double_buffer(0)
        g.close()


@pytest.mark.gpu
@pytest.mark.parametrize("in_place", [False, True])
def test_frame_policy_large_arrays_follow_their_host_memory(refgl, ref_gpu, in_place):
    """RSRCU_UPLOAD_FRAME (the drop-in binding's default): an array is staged once per frame however often it is bound, and
    the host may change it between frames.  With rsrcu_set_pin_in_place arrays of 1 MiB or more are page-locked where
    they lie and copied by the copy engine instead of through the staging arena."""
    import rsr_b200
    from rsr_b200 import GL_LINEAR_MIPMAP_NEAREST, PROGRAM_AMY, PROGRAM_DEFAULT_POST
    from rsr_b200.scenes import WavyGridScene, begin, perspective, translate
    from oracle.refgl import make_mipmap
    size, dim = (640, 360), 512
    sc = WavyGridScene(n=16, tex_dim=64)
    rng = np.random.default_rng(12)
    tex = make_mipmap(rng.random((dim, dim, 4), dtype=np.float32))   # 8 MiB with its mip chain
    gpu = rsr_b200.GPU(0, direct=True)
    gpu.set_pin_in_place(in_place)

    def frame(gl, out, **kw):
        begin(gl, size)
        gl.UseProgram(PROGRAM_AMY)
        gl.ViewMatrix(translate(0, 0, -6))
        gl.ProjectionMatrix(perspective(45.0, size[0] / size[1], 1.0, 100.0))
        gl.UseBuffer(0, sc.pos); gl.UseBuffer(3, sc.nrm); gl.UseBuffer(9, sc.uv)
        for _ in range(3):   # bound again and again inside the frame, like a field of quads sharing a texture
            gl.BindTexture(0, tex, dim, dim, dim, GL_LINEAR_MIPMAP_NEAREST, **kw)
            gl.DrawElements(len(sc.idx), sc.idx, 0)
        gl.UseProgram(PROGRAM_DEFAULT_POST)
        gl.StoreColor(out, True)
    try:
        for i in range(5):
            tex[:dim] = rng.random((dim, dim, 4), dtype=np.float32)   # (the mip rows keep old content: all that matters is that both see the same bytes)
            want, got = np.zeros((size[1], size[0]), np.uint32), np.zeros((size[1], size[0]), np.uint32)
            frame(ref_gpu, want)
            ref_gpu.Run()
            frame(gpu, got, upload=rsr_b200.UPLOAD_FRAME)
            gpu.Run()
            assert np.array_equal(got, want), f"frame {i}: {np.count_nonzero(got != want)} pixels differ"
            if in_place:
                assert gpu.stats()["h2d_bytes"] < tex.nbytes // 4          # the texture did not go through the staging arena
            else:
                assert tex.nbytes <= gpu.stats()["h2d_bytes"] < 2 * tex.nbytes   # staged once, not three times
        if in_place:
            with pytest.raises(rsr_b200.RsrError):
                gpu.Retain()
    finally:
        gpu.set_pin_in_place(False)   # releases the pages before `tex` goes away
        gpu.close()
