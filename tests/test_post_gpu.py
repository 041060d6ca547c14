"""SURVEY 8(f)2 / 8(f)4 on the device: canvases that stay in HBM (render to texture, the shadow-map pass of a `$layer`
on a second context), the `$kawase` blur, the `$glow` combine, the telemetry span bars and the presentation blit --
each compared bit for bit with the reference's own functions (oracle/_ref/librsr_ref.so)."""
import numpy as np
import pytest

import rsr_b200
from rsr_b200 import (GL_COLOR_ATTACHMENT0, GL_COLOR_BUFFER_BIT, GL_CULL_FACE, GL_DEPTH_ATTACHMENT, GL_DEPTH_BUFFER_BIT,
                      GL_FRONT, GL_LESS, GL_NEAREST_MIPMAP_NEAREST, PROGRAM_AMY, PROGRAM_DEFAULT_POST, PROGRAM_OBJ2S, RB_F32, RB_RGBF32)
from rsr_b200.scenes import SoupScene, WavyGridScene, perspective, translate

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def test_kawase_blur_matches_reference(refgl, cuda_gpu):
    rng = np.random.default_rng(5)
    w, h = 322, 182   # not a multiple of the CTA footprint; borders take the clamped path
    src = rng.normal(0.0, 2.0, (h, w, 4)).astype(np.float32)
    src[::7, ::5] = -0.0
    src[3, 4] = 1e30
    a, b = cuda_gpu.Canvas("fp", w, h), cuda_gpu.Canvas("fp", w, h)
    a.write(src)
    for dist in (0, 1, 2, 5, 40):
        cuda_gpu.KawaseBlur(a, b, dist)
        got = b.read()
        want = refgl.kawase_blur(src, dist)
        assert np.array_equal(bits(got), bits(want)), f"dist {dist}: {np.count_nonzero(bits(got) != bits(want))} floats differ"
    a.free(); b.free()


def _soup_frame(gl, scene, size):
    gl.Reset(size, (8, 8))
    gl.ClearColor((0.1, 0.2, 0.3))
    gl.ClearDepth(1.0)
    gl.Clear(GL_COLOR_BUFFER_BIT | GL_DEPTH_BUFFER_BIT)
    scene.draw(gl, size)
    gl.UseProgram(PROGRAM_DEFAULT_POST)


@pytest.mark.parametrize("direct", [False, True])
@pytest.mark.parametrize("gamma", [True, False])
def test_glow_chain_stays_on_the_device(refgl, ref_gpu, direct, gamma):
    """`$gpu` -> `$buffers` (quads + half-size linear) -> `$kawase` (intensity 3) -> `$glow` -> true colour"""
    size = (640, 360)
    w, h = size
    scene = SoupScene(n=300, seed=21)
    # reference: every canvas in host memory, the reference's own filters
    quads = np.zeros((h // 2, w // 2, 4, 4), np.float32)
    half = np.zeros((h // 2, w // 2, 4), np.float32)
    _soup_frame(ref_gpu, scene, size)
    ref_gpu.StoreColorQuads(quads)
    ref_gpu.StoreColorHalf(half)
    ref_gpu.Run()
    blur = half
    for dist in range(3):
        blur = refgl.kawase_blur(blur, dist)
    want = refgl.glow_filter(quads, blur, gamma)

    gpu = rsr_b200.GPU(0, direct=direct)
    try:
        cq, ch = gpu.Canvas("quads", w, h), gpu.Canvas("fp", w // 2, h // 2)
        _soup_frame(gpu, scene, size)
        gpu.StoreToCanvas(cq)
        gpu.StoreToCanvas(ch, half=True)
        gpu.Run(sync=False)
        blurred = gpu.Kawase(ch, 3)
        got = np.zeros((h, w), np.uint32)
        gpu.Glow(cq, blurred, got, gamma)
        gpu.Sync()
        assert gpu.stats()["d2h_bytes"] == 0   # the frame itself copied nothing back
        assert np.array_equal(bits(cq.read()), bits(quads))
        assert np.array_equal(bits(blurred.read()), bits(blur))
        assert np.array_equal(got, want), f"{np.count_nonzero(got != want)} pixels differ"
        # the same chain into a device true-colour canvas, then the presentation blit
        tc, surface = gpu.Canvas("tc", w, h), gpu.Canvas("tc", w, h)
        gpu.Glow(cq, blurred, tc, gamma)
        gpu.Present(tc, surface)
        assert np.array_equal(surface.read(), want)
    finally:
        gpu.close()


def _shadow_pass(gl, scene, dim, light_view, light_proj):
    """the `$layer`'s light pass (src/viewer/node/gllayer.cxx:160-176)"""
    gl.Reset((dim, dim), (8, 8))
    gl.RenderbufferType(GL_COLOR_ATTACHMENT0, RB_RGBF32)
    gl.RenderbufferType(GL_DEPTH_ATTACHMENT, RB_F32)
    gl.ColorWriteMask(False)
    gl.DepthWriteMask(True)
    gl.DepthFunc(GL_LESS)
    gl.Enable(GL_CULL_FACE)
    gl.CullFace(GL_FRONT)
    gl.ClearDepth(1.0)
    gl.Clear(GL_DEPTH_BUFFER_BIT)
    gl.UseProgram(0)
    # IGl::DrawDepth (e.g. node/mc.cxx:195-210): matrices + the position buffer only
    gl.ViewMatrix(light_view)
    gl.ProjectionMatrix(light_proj)
    gl.UseBuffer(0, scene.pos)
    gl.DrawElements(len(scene.idx), scene.idx, 0)


def _lit_pass(gl, scene, size, light_view, light_proj, bind_shadow):
    w, h = size
    gl.Reset(size, (8, 8))
    gl.ClearColor((0.05, 0.05, 0.1))
    gl.ClearDepth(1.0)
    gl.Clear(GL_COLOR_BUFFER_BIT | GL_DEPTH_BUFFER_BIT)
    gl.UseProgram(PROGRAM_OBJ2S)
    gl.ViewMatrix(translate(0.1, -0.05, -1.0))
    gl.ProjectionMatrix(perspective(70.0, w / h, 0.5, 50.0))
    gl.UseBuffer(0, scene.pos)
    gl.UseBuffer(3, scene.nrm)
    gl.UseBuffer(6, scene.kd)
    m2s = (light_proj @ light_view).T.reshape(16)
    gl.UseUniforms(np.concatenate([m2s, [0.5, 2.0, 1.0], [0.1, -0.6, -0.8], [0.3]]).astype(np.float32))
    bind_shadow(gl)
    gl.DrawElements(len(scene.idx), scene.idx, 0)
    gl.UseProgram(PROGRAM_DEFAULT_POST)


def test_shadow_map_pass_on_a_second_context(refgl, ref_gpu):
    """gllayer.cxx:154-181: a second GPU renders the light's depth map (BaseProgram, key 0x6a2), the main pass samples
    it through texture unit 3.  Here the map never leaves the device and the two contexts are ordered by an event."""
    dim, size = 256, (640, 360)
    scene = SoupScene(n=250, seed=33, near_cross=False)
    lv = translate(0.3, 0.2, -2.0)
    lp = perspective(60.0, 1.0, 0.5, 60.0)
    # reference
    shadow_ref = refgl.RefGPU()
    shadow_ref.InstallShadowProgram()
    want_map = np.zeros((dim, dim), np.float32)
    _shadow_pass(shadow_ref, scene, dim, lv, lp)
    shadow_ref.StoreDepth(want_map)
    shadow_ref.Run()
    want = np.zeros((size[1], size[0]), np.uint32)
    _lit_pass(ref_gpu, scene, size, lv, lp, lambda gl: gl.BindTexture3(want_map, dim))
    ref_gpu.StoreColor(want, True)
    ref_gpu.Run()
    shadow_ref.close()
    assert want_map.min() < 1.0   # the light sees the soup

    main, light = rsr_b200.GPU(0), rsr_b200.GPU(0)
    try:
        depth_canvas = light.Canvas("depth", dim, dim)
        for _ in range(2):   # twice: the second round reuses every buffer
            _shadow_pass(light, scene, dim, lv, lp)
            light.StoreToCanvas(depth_canvas)
            light.Run(sync=False)
            main.WaitFor(light)
            got = np.zeros_like(want)
            _lit_pass(main, scene, size, lv, lp, lambda gl: gl.BindTexture3Device(depth_canvas))
            main.StoreColor(got, True)
            main.Run()
            light.WaitFor(main)   # the next light pass overwrites the map the main pass has just read
            assert main.stats()["h2d_bytes"] < dim * dim * 4   # the map was not uploaded
            assert np.array_equal(bits(depth_canvas.read()), bits(want_map)), "shadow map differs"
            assert np.array_equal(got, want), f"{np.count_nonzero(got != want)} pixels differ"
    finally:
        main.close(); light.close()


def test_render_to_texture_on_the_device(refgl, ref_gpu, cuda_gpu):
    """`$rendertotexture`: pass 1 stores its frame as a linear float canvas, pass 2 samples it as texture unit 0
    (320x180 is not a power of two: the nearest, unmipped sampler of rglr_texture_sampler.cxx:289-311)"""
    tsize, size = (320, 180), (640, 360)
    inner, outer = WavyGridScene(n=12, tex_dim=64), WavyGridScene(n=8, tex_dim=64, wave=0.2)

    def pass1(gl, store):
        from rsr_b200.scenes import begin
        begin(gl, tsize)
        gl.UseProgram(PROGRAM_AMY)
        gl.ViewMatrix(translate(0, 0, -6))
        gl.ProjectionMatrix(perspective(45.0, tsize[0] / tsize[1], 1.0, 100.0))
        gl.UseBuffer(0, inner.pos); gl.UseBuffer(3, inner.nrm); gl.UseBuffer(9, inner.uv)
        gl.BindTexture(0, inner.tex, inner.tex_dim, inner.tex_dim, inner.tex_dim, inner.filter)
        gl.DrawElements(len(inner.idx), inner.idx, 0)
        gl.UseProgram(PROGRAM_DEFAULT_POST)
        store(gl)

    def pass2(gl, bind, out):
        from rsr_b200.scenes import begin
        begin(gl, size)
        gl.UseProgram(PROGRAM_AMY)
        gl.ViewMatrix(translate(0, 0, -5))
        gl.ProjectionMatrix(perspective(45.0, size[0] / size[1], 1.0, 100.0))
        gl.UseBuffer(0, outer.pos); gl.UseBuffer(3, outer.nrm); gl.UseBuffer(9, outer.uv)
        bind(gl)
        gl.DrawElements(len(outer.idx), outer.idx, 0)
        gl.UseProgram(PROGRAM_DEFAULT_POST)
        gl.StoreColor(out, True)

    tex_ref = np.zeros((tsize[1], tsize[0], 4), np.float32)
    pass1(ref_gpu, lambda gl: gl.StoreColor(tex_ref))
    ref_gpu.Run()
    want = np.zeros((size[1], size[0]), np.uint32)
    pass2(ref_gpu, lambda gl: gl.BindTexture(0, tex_ref, tsize[0], tsize[1], tsize[0], GL_NEAREST_MIPMAP_NEAREST), want)
    ref_gpu.Run()

    canvas = cuda_gpu.Canvas("fp", *tsize)
    pass1(cuda_gpu, lambda gl: gl.StoreToCanvas(canvas))
    cuda_gpu.Run(sync=False)
    got = np.zeros_like(want)
    pass2(cuda_gpu, lambda gl: gl.BindTextureDevice(0, canvas, GL_NEAREST_MIPMAP_NEAREST), got)
    cuda_gpu.Run()
    assert np.array_equal(bits(canvas.read()), bits(tex_ref))
    assert np.array_equal(got, want), f"{np.count_nonzero(got != want)} pixels differ"
    assert len(np.unique(got)) > 50
    canvas.free()


def test_telemetry_span_bars_match_render_jobsys(refgl, cuda_gpu):
    w, h = 640, 360
    rng = np.random.default_rng(9)
    spans = []
    for lane in range(6):
        t = 0.0
        for _ in range(12):
            a = t + rng.uniform(0.0, 0.004)
            b = a + rng.choice([0.0, 1e-5, 0.002, 0.02])
            spans.append((a, b, int(rng.integers(0, 2 ** 32)), lane))
            t = b
    base = rng.integers(0, 2 ** 24, (h, w)).astype(np.uint32)
    want = base.copy()
    refgl.render_spans(want, 20, 40, 3.0, spans)
    tc = cuda_gpu.Canvas("tc", w, h)
    tc.write(base)
    cuda_gpu.DrawSpans(tc, 20, 40, 3.0, spans)
    got = tc.read()
    assert np.count_nonzero(want != base) > 1000
    assert np.array_equal(got, want), f"{np.count_nonzero(got != want)} pixels differ"
    tc.free()


def test_frame_spans_from_cuda_events_drawn_over_the_frame(cuda_gpu):
    """the overlay's input is the frame's own stage timings (CUDA events), not jobsys::measurements_pt"""
    size = (640, 360)
    scene = WavyGridScene(n=24)
    cuda_gpu.set_profiling(2)
    try:
        tc = cuda_gpu.Canvas("tc", *size)
        scene.record(cuda_gpu, size, None)
        cuda_gpu.Run()
        spans = cuda_gpu.FrameSpans()
        assert len(spans) >= 4 and all(b > a for a, b, _, _ in spans)
        assert spans[-1][3] == 5 and abs(spans[-1][1] * 1e3 - cuda_gpu.stage_ms()["frame"]) < 1e-3   # the tile stage ends the frame
        ptr, stride = cuda_gpu.device_truecolor()
        frame = cuda_gpu.read_device(ptr, size[0] * size[1], np.uint32).reshape(size[1], size[0])
        tc.write(frame)
        cuda_gpu.DrawSpans(tc, 20, 40, 1.0 / max(s[1] for s in spans) * 0.9, spans)
        over = tc.read()
        changed = over != frame
        assert changed[40:40 + 6 * 10].any() and not changed[:40].any() and not changed[40 + 6 * 10:].any()
        tc.free()
    finally:
        cuda_gpu.set_profiling(0)


@pytest.mark.parametrize("dim", [4, 16, 64, 256, 1024])
def test_device_mipmap_matches_reference(refgl, cuda_gpu, dim):
    """rglr::Texture::maybe_make_mipmap (rglr_texture.cxx:33-81) on a device texture"""
    rng = np.random.default_rng(dim)
    base = rng.normal(0.5, 1.0, (dim, dim, 4)).astype(np.float32)
    want = refgl.make_mipmap(base)
    tex = cuda_gpu.Canvas("tex", dim, dim)
    both = np.zeros((2 * dim, dim, 4), np.float32)
    both[:dim] = base
    tex.write(both)
    cuda_gpu.MakeMipmap(tex)
    got = tex.read()
    # (the reference leaves the texels right of each level untouched -- zeros in a fresh buffer, as here)
    assert np.array_equal(bits(got), bits(want)), f"{np.count_nonzero(bits(got) != bits(want))} floats differ"
    tex.free()


def test_render_to_a_mipmapped_texture_on_the_device(refgl, ref_gpu, cuda_gpu):
    """`$renderToTexture` with a power-of-two target (node/rendertotexture.cxx:70-91): StoreColor into the texture's base
    level, maybe_make_mipmap, then sampled bilinearly with LOD selection by the next pass -- all on the device"""
    from rsr_b200 import GL_LINEAR_MIPMAP_NEAREST
    from rsr_b200.scenes import begin
    dim, size = 256, (640, 360)
    inner, outer = WavyGridScene(n=12, tex_dim=64), WavyGridScene(n=8, tex_dim=64, wave=0.4)

    def pass1(gl, store):
        begin(gl, (dim, dim))
        gl.UseProgram(PROGRAM_AMY)
        gl.ViewMatrix(translate(0, 0, -6))
        gl.ProjectionMatrix(perspective(45.0, 1.0, 1.0, 100.0))
        gl.UseBuffer(0, inner.pos); gl.UseBuffer(3, inner.nrm); gl.UseBuffer(9, inner.uv)
        gl.BindTexture(0, inner.tex, inner.tex_dim, inner.tex_dim, inner.tex_dim, inner.filter)
        gl.DrawElements(len(inner.idx), inner.idx, 0)
        gl.UseProgram(PROGRAM_DEFAULT_POST)
        store(gl)

    def pass2(gl, bind, out):
        begin(gl, size)
        gl.UseProgram(PROGRAM_AMY)
        gl.ViewMatrix(translate(0, 0, -9) @ rsr_b200.scenes.rotate(0.9, 1.0, 0.2, 0.0))   # tilted and far: several LODs
        gl.ProjectionMatrix(perspective(45.0, size[0] / size[1], 1.0, 100.0))
        gl.UseBuffer(0, outer.pos); gl.UseBuffer(3, outer.nrm); gl.UseBuffer(9, outer.uv)
        bind(gl)
        gl.DrawElements(len(outer.idx), outer.idx, 0)
        gl.UseProgram(PROGRAM_DEFAULT_POST)
        gl.StoreColor(out, True)

    base_ref = np.zeros((dim, dim, 4), np.float32)
    pass1(ref_gpu, lambda gl: gl.StoreColor(base_ref))
    ref_gpu.Run()
    tex_ref = refgl.make_mipmap(base_ref)
    want = np.zeros((size[1], size[0]), np.uint32)
    pass2(ref_gpu, lambda gl: gl.BindTexture(0, tex_ref, dim, dim, dim, GL_LINEAR_MIPMAP_NEAREST), want)
    ref_gpu.Run()

    tex = cuda_gpu.Canvas("tex", dim, dim)
    pass1(cuda_gpu, lambda gl: gl.StoreToCanvas(tex))
    cuda_gpu.Run(sync=False)
    cuda_gpu.MakeMipmap(tex)
    got = np.zeros_like(want)
    pass2(cuda_gpu, lambda gl: gl.BindTextureDevice(0, tex, GL_LINEAR_MIPMAP_NEAREST), got)
    cuda_gpu.Run()
    assert np.array_equal(bits(tex.read()), bits(tex_ref))
    assert np.array_equal(got, want), f"{np.count_nonzero(got != want)} pixels differ"
    assert len(np.unique(got)) > 200
    tex.free()


def test_new_entry_points_refuse_bad_arguments(cuda_gpu):
    a, b = cuda_gpu.Canvas("fp", 64, 32), cuda_gpu.Canvas("fp", 64, 32)
    q, tc = cuda_gpu.Canvas("quads", 64, 32), cuda_gpu.Canvas("tc", 64, 32)
    L, h = cuda_gpu.L, cuda_gpu.h
    import ctypes as C
    vp = C.c_void_p
    bad = [
        lambda: L.rsrcu_kawase_blur(h, vp(a.ptr), 64, vp(a.ptr), 64, 64, 32, 1),                      # in place
        lambda: L.rsrcu_kawase_blur(h, vp(a.ptr), 32, vp(b.ptr), 64, 64, 32, 1),                      # stride < width
        lambda: L.rsrcu_kawase_blur(h, vp(a.ptr), 64, vp(b.ptr), 64, 64, 32, -1),                     # negative distance
        lambda: L.rsrcu_glow(h, vp(q.ptr), 32, vp(a.ptr), 64, 1, vp(tc.ptr), 1, 62, 32, 64),          # width not a multiple of 4
        lambda: L.rsrcu_glow(h, vp(q.ptr), 8, vp(a.ptr), 64, 1, vp(tc.ptr), 1, 64, 32, 64),           # quad stride too small
        lambda: L.rsrcu_make_mipmap(h, vp(a.ptr), 48),                                                # not a power of two
        lambda: L.rsrcu_canvas_free(h, vp(12345)),                                                     # not a canvas
        lambda: L.rsrcu_canvas_alloc(h, 0, C.byref(vp())),                                             # empty
        lambda: L.rsrcu_store_depth_device(h, vp(a.ptr)),                                              # outside a frame
        lambda: L.rsrcu_frame_spans(h, (rsr_b200.RsrSpan * 8)(), 8, C.byref(C.c_int())),              # no profiled frame... unless one ran
    ]
    for i, call in enumerate(bad[:-1]):
        assert call() != 0, f"call {i} was accepted"
        assert cuda_gpu.L.rsrcu_last_error()
    with pytest.raises(rsr_b200.RsrError):
        cuda_gpu.MarchSurface(0.0, 32, 2, -1.0)
    # a device store of the wrong size is refused when it is recorded
    g = rsr_b200.GPU(0, direct=True)
    try:
        g.Reset((128, 64))
        with pytest.raises(rsr_b200.RsrError, match="store canvas"):
            g.StoreToCanvas(a)      # 64x32 canvas, 128x64 target
    finally:
        g.close()
    for c in (a, b, q, tc):
        c.free()
