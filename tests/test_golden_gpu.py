"""-m gpu: the CUDA path against (a) the committed golden frames the reference rendered in the build
container -- with that CPU's rcpps/rsqrtps tables loaded through rsrcu_set_host_luts, so the check
does not depend on this box's CPU -- and (b) the C restatement on the same seeded inputs."""
import importlib.util
import os

import numpy as np
import pytest

import rsr_b200 as R
from oracle import restate
from rsr_b200 import scenes

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLDEN, "make_golden.py"))
make_golden = importlib.util.module_from_spec(spec)
spec.loader.exec_module(make_golden)


@pytest.mark.parametrize("name", sorted(make_golden.cases().keys()))
def test_cuda_reproduces_golden_frames(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    factory, kw = make_golden.cases()[name]
    size = tuple(int(v) for v in g["size"])
    gpu = R.GPU(0)
    try:
        gpu.set_host_luts(g["rcp"], g["rsqrt"])
        out = np.zeros((size[1], size[0]), np.uint32)
        factory().record(gpu, size, out, **kw)
        gpu.Run()
    finally:
        gpu.close()
    assert np.array_equal(out, g["frame"]), f"{np.count_nonzero(out != g['frame'])} pixels differ from the golden frame"


def test_cuda_reproduces_the_bundled_colortest_scene():
    """C1 parity gate (BASELINE.json configs[0]): data/scene/colortest.lua at 640x360 -- the reference's
    mesh, camera and frame from tests/golden/colortest_c1.npz"""
    g = np.load(os.path.join(GOLDEN, "colortest_c1.npz"))
    size = tuple(int(v) for v in g["size"])
    gpu = R.GPU(0)
    try:
        gpu.set_host_luts(g["rcp"], g["rsqrt"])
        out = np.zeros((size[1], size[0]), np.uint32)
        scenes.ColortestScene().record(gpu, size, out)
        gpu.Run()
    finally:
        gpu.close()
    assert np.array_equal(out, g["frame"]), f"{np.count_nonzero(out != g['frame'])} pixels differ from the reference's frame"


@pytest.mark.parametrize("size", [(640, 360), (1920, 1080)])
def test_bundled_colortest_scene_against_the_live_reference(size, ref_gpu, cuda_gpu):
    """configs[0] and configs[1]: the same bundled scene at 640x360 and 1920x1080 through both renderers on this box"""
    sc = scenes.ColortestScene()
    a, b = np.zeros((size[1], size[0]), np.uint32), np.zeros((size[1], size[0]), np.uint32)
    sc.record(ref_gpu, size, a)
    ref_gpu.Run()
    sc.record(cuda_gpu, size, b)
    cuda_gpu.Run()
    assert np.count_nonzero(a != a[0, 0]) > size[0] * size[1] // 5
    assert np.array_equal(a, b), f"{np.count_nonzero(a != b)} pixels differ"


def test_host_lut_harvest_equals_oracle_harvest(cuda_gpu):
    rcp, rsq = cuda_gpu.get_host_luts()
    want = restate.harvest_luts()
    assert np.array_equal(rcp, want[0]) and np.array_equal(rsq, want[1])


@pytest.mark.parametrize("case", ["c2_1080p", "soup_1080p"])
def test_cuda_equals_restatement_with_depth(case, cuda_gpu):
    """depth buffers bit for bit (the reference cannot store depth from RB_COLOR_DEPTH; the restatement can)"""
    sc, kw = {"c2_1080p": (scenes.BundledLikeScene(cubes=500), {"t": 0.25}),
              "soup_1080p": (scenes.SoupScene(n=800, seed=23), {})}[case]
    size = (1920, 1080)
    a, da = np.zeros((size[1], size[0]), np.uint32), np.zeros((size[1], size[0]), np.float32)
    b, db = np.zeros_like(a), np.zeros_like(da)
    sc.record(cuda_gpu, size, a, da, **kw)
    cuda_gpu.Run()
    rst = restate.RestateGPU()
    sc.record(rst, size, b, db, **kw)
    rst.Run()
    assert np.array_equal(a, b), f"{np.count_nonzero(a != b)} colour pixels differ"
    assert np.array_equal(da.view(np.uint32), db.view(np.uint32)), "depth buffers differ"
    assert cuda_gpu.stats()["fragments_shaded"] == rst.fragments


def test_determinism_and_idempotence(cuda_gpu):
    """same recorded frame twice -> same bytes (size-independent property)"""
    sc = scenes.BundledLikeScene(cubes=800)
    a = np.zeros((1080, 1920), np.uint32)
    sc.record(cuda_gpu, (1920, 1080), a, t=0.3, static=True)
    rec = cuda_gpu.Finish()
    cuda_gpu.Submit(rec)
    first = a.copy()
    a[:] = 0
    cuda_gpu.Submit(rec)
    assert np.array_equal(first, a)


def test_submission_order_decides_equal_depth(cuda_gpu, ref_gpu):
    """two coplanar quads with different textures: LESS keeps the first one drawn, everywhere"""
    class Coplanar:
        def __init__(self):
            q = np.array([[-1, 1, 1, -1], [-1, -1, 1, 1], [0, 0, 0, 0]], np.float32)
            self.pos, self.uv = scenes.soa(q), scenes.soa((q[:2] + 1) / 2)
            self.idx = np.array([0, 1, 2, 0, 2, 3], np.uint16)
            self.t0 = scenes.make_mipmap(scenes.hash_texture(32, 1, 0))
            self.t1 = scenes.make_mipmap(scenes.hash_texture(32, 1, 1))

        def record(self, gl, size, out, depth=None):
            scenes.begin(gl, size)
            gl.UseProgram(R.PROGRAM_AMY)
            gl.ViewMatrix(scenes.translate(0, 0, -3))
            gl.ProjectionMatrix(scenes.perspective(40.0, size[0] / size[1], 1, 10))
            gl.UseBuffer(0, self.pos); gl.UseBuffer(9, self.uv)
            for t in (self.t0, self.t1):
                gl.BindTexture(0, t, 32, 32, 32, R.GL_NEAREST_MIPMAP_NEAREST)
                gl.DrawElements(6, self.idx, 0)
            scenes.finish(gl, out)
    sc = Coplanar()
    a, b = np.zeros((360, 640), np.uint32), np.zeros((360, 640), np.uint32)
    sc.record(cuda_gpu, (640, 360), a); cuda_gpu.Run()
    sc.record(ref_gpu, (640, 360), b); ref_gpu.Run()
    assert np.array_equal(a, b)


def test_static_uploads_pin_their_host_memory(cuda_gpu):
    """RSRCU_UPLOAD_STATIC is keyed by host pointer + size: the Python mirror must keep such arrays alive,
    or a recycled address would silently serve stale vertices (regression)"""
    frames = []
    for z in (-3.0, -6.0):
        q = scenes.soa(np.array([[-1, 1, 1, -1], [-1, -1, 1, 1], [z, z, z, z]], np.float32))
        uv = scenes.soa(np.array([[0, 1, 1, 0], [0, 0, 1, 1]], np.float32))
        tex = scenes.make_mipmap(scenes.hash_texture(16, 3))
        out = np.zeros((180, 320), np.uint32)
        scenes.begin(cuda_gpu, (320, 180))
        cuda_gpu.UseProgram(R.PROGRAM_AMY)
        cuda_gpu.ProjectionMatrix(scenes.perspective(45.0, 320 / 180, 1, 20))
        cuda_gpu.UseBuffer(0, q, upload=R.UPLOAD_STATIC)
        cuda_gpu.UseBuffer(9, uv, upload=R.UPLOAD_STATIC)
        cuda_gpu.BindTexture(0, tex, 16, 16, 16, R.GL_NEAREST_MIPMAP_NEAREST, upload=R.UPLOAD_STATIC)
        cuda_gpu.DrawElements(6, np.array([0, 1, 2, 0, 2, 3], np.uint16), 0)
        scenes.finish(cuda_gpu, out)
        cuda_gpu.Run()
        frames.append(out)
        del q, uv, tex
    bg = frames[0][0, 0]
    assert np.count_nonzero(frames[0] != bg) > 2 * np.count_nonzero(frames[1] != bg) > 0


def test_two_deep_frame_pipelining(cuda_gpu):
    """frames submitted back to back without a full sync land, in order, in their own destinations"""
    sc = scenes.CubesScene(instances=150)
    size = (640, 360)
    want = []
    for i in range(5):
        out = np.zeros((size[1], size[0]), np.uint32)
        sc.record(cuda_gpu, size, out, t=0.3 * i)
        cuda_gpu.Run()
        want.append(out)
    bufs = [np.zeros((size[1], size[0]), np.uint32) for _ in range(5)]
    recs = []
    for i in range(5):
        sc.record(cuda_gpu, size, bufs[i], t=0.3 * i)
        recs.append(cuda_gpu.Finish())
    for i, rec in enumerate(recs):
        cuda_gpu.Submit(rec, sync=False)
        if i > 0:
            cuda_gpu.SyncFrame(1)
            assert np.array_equal(bufs[i - 1], want[i - 1]), f"frame {i - 1} not complete after SyncFrame(1)"
    cuda_gpu.Sync()
    assert all(np.array_equal(a, b) for a, b in zip(bufs, want))
    assert not np.array_equal(want[0], want[4])


def test_frame_overlap_mode_front_end_under_previous_tile_kernel():
    """rsrcu_set_overlap: K0-K5 of frame N+1 run on a second stream into the other work set while the tile kernel
    of frame N rasterises; alternating heavy / light / differently sized frames must still land bit-identical to the
    same frames rendered one at a time"""
    g = R.GPU(0)
    try:
        heavy = scenes.BundledLikeScene(cubes=400)
        light = scenes.WavyGridScene(n=12)
        cubes = scenes.CubesScene(instances=200)
        plan = [(heavy, (1920, 1080)), (light, (640, 360)), (cubes, (1920, 1080)), (light, (1920, 1080)), (heavy, (640, 360)),
                (cubes, (640, 360)), (heavy, (1920, 1080)), (light, (640, 360)), (cubes, (1920, 1080)), (heavy, (1920, 1080))]
        want = []
        for i, (sc, size) in enumerate(plan):
            out = np.zeros((size[1], size[0]), np.uint32)
            sc.record(g, size, out, t=0.25 * i)
            g.Run()
            want.append(out)
        g.set_overlap(True)
        for rounds in range(2):
            bufs = [np.zeros_like(w) for w in want]
            recs = []
            for i, (sc, size) in enumerate(plan):
                sc.record(g, size, bufs[i], t=0.25 * i)
                recs.append(g.Finish())
            for i, rec in enumerate(recs):
                g.Submit(rec, sync=False)
                if i > 1:
                    g.SyncFrame(2)
                    assert np.array_equal(bufs[i - 2], want[i - 2]), f"frame {i - 2} differs in overlap mode"
            g.Sync()
            for i, (a, b) in enumerate(zip(bufs, want)):
                assert np.array_equal(a, b), f"frame {i} differs in overlap mode"
        g.set_overlap(False)
        out = np.zeros_like(want[0])
        plan[0][0].record(g, plan[0][1], out, t=0.0)
        g.Run()
        assert np.array_equal(out, want[0])
    finally:
        g.close()


def test_submit_pool_three_contexts_render_the_sequence_bit_identical():
    """rsr_b200.SubmitPool: frames dealt round robin to three contexts / host threads land bit-identical to the same
    frames rendered one at a time on one context"""
    heavy = scenes.BundledLikeScene(cubes=300)
    light = scenes.WavyGridScene(n=12)
    plan = [(heavy, (1920, 1080)), (light, (640, 360)), (heavy, (640, 360)), (light, (1920, 1080))] * 4
    g = R.GPU(0)
    try:
        want = []
        for i, (sc, size) in enumerate(plan):
            out = np.zeros((size[1], size[0]), np.uint32)
            sc.record(g, size, out, t=0.25 * i)
            g.Run()
            want.append(out)
    finally:
        g.close()
    pool = R.SubmitPool(0, contexts=3)
    try:
        bufs = [np.zeros_like(w) for w in want]
        tickets = []
        for i, (sc, size) in enumerate(plan):
            sc.record(pool.gpus[0], size, bufs[i], t=0.25 * i)
            tickets.append(pool.submit(pool.gpus[0].Finish()))
            if i >= 9:
                pool.wait(tickets[i - 9])
                assert np.array_equal(bufs[i - 9], want[i - 9]), f"frame {i - 9} differs"
        pool.drain()
        for i, (a, b) in enumerate(zip(bufs, want)):
            assert np.array_equal(a, b), f"frame {i} differs"
    finally:
        pool.close()


@pytest.mark.parametrize("overlap", [False, True])
def test_retained_frames_replay_bit_identical(overlap):
    """rsrcu_retain_frame / rsrcu_replay_frame: the tables of a submitted frame (including its per-frame UPLOAD_ALWAYS
    data, whose device addresses move to the private copy) stay on the device; replays -- interleaved with other
    frames and with each other -- reproduce the frame bit for bit"""
    g = R.GPU(0)
    try:
        g.set_overlap(overlap)
        a_scene, b_scene = scenes.BundledLikeScene(cubes=300), scenes.CubesScene(instances=120)
        size = (1920, 1080)
        outs, want, handles = [], [], []
        for sc, kw in ((a_scene, {"t": 0.7}), (b_scene, {"t": 0.2}), (a_scene, {"t": 1.9, "static": True})):
            out = np.zeros((size[1], size[0]), np.uint32)
            sc.record(g, size, out, **kw)
            g.Run()
            handles.append(g.Retain())
            outs.append(out)
            want.append(out.copy())
        assert not np.array_equal(want[0], want[2])
        other = np.zeros((360, 640), np.uint32)
        for rounds in range(3):
            for o in outs:
                o[:] = 0
            for k in (2, 0, 1):
                g.Replay(handles[k])
            scenes.WavyGridScene(n=10).record(g, (640, 360), other)   # an ordinary frame in between
            g.Run()
            for k in range(3):
                assert np.array_equal(outs[k], want[k]), f"replay of frame {k} differs (round {rounds})"
        st = g.stats()
        for h in handles:
            g.Release(h)
    finally:
        g.close()


def test_tile_list_overflow_is_recovered_inside_sync(ref_gpu, monkeypatch):
    """more (triangle, tile) pairs than the list buffer holds: rsrcu_sync grows the buffer and launches the
    frame again from its device-resident tables; the caller gets the reference's pixels, never a truncated frame"""
    monkeypatch.setenv("RSRCU_LIST_CAPACITY", "2000")
    g = R.GPU(0)
    try:
        sc = scenes.CubesScene(instances=300)
        a = np.zeros((360, 640), np.uint32)
        sc.record(g, (640, 360), a)
        g.Run()
        st = g.stats()
        assert st["bin_entries"] > 2000 and st["frames_retried"] >= 1
        b = np.zeros_like(a)
        sc.record(ref_gpu, (640, 360), b)
        ref_gpu.Run()
        assert np.array_equal(a, b)
    finally:
        g.close()


def test_overflow_in_a_pipelined_frame_is_recovered_by_sync_frame(ref_gpu, monkeypatch):
    """three frames in flight (Submit(sync=False) + SyncFrame, the end-to-end path of bench.py), the first one
    overflows the tile lists: rsrcu_sync_frame launches exactly that frame again from the arena ring while the
    later frames are already queued; every host destination ends up with its own complete frame"""
    monkeypatch.setenv("RSRCU_LIST_CAPACITY", "2000")
    g = R.GPU(0)
    try:
        sc = scenes.CubesScene(instances=300)
        outs = [np.zeros((360, 640), np.uint32) for _ in range(4)]
        recs = []
        for i in range(4):
            sc.record(g, (640, 360), outs[i], t=0.5 * i)
            recs.append(g.Finish())
        for i, rec in enumerate(recs):
            g.Submit(rec, sync=False)
            if i >= 2:
                g.SyncFrame(2)
        g.Sync()
        assert g.stats()["frames_retried"] >= 1
        for i in range(4):
            want = np.zeros((360, 640), np.uint32)
            sc.record(ref_gpu, (640, 360), want, t=0.5 * i)
            ref_gpu.Run()
            assert np.array_equal(outs[i], want), f"pipelined frame {i} differs"
    finally:
        g.close()
