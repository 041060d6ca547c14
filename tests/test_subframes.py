"""sub-frame decomposition (>2048 px targets, multi-GPU shard unit): host logic on CPU incl. a 2-rank gloo
run; parity of every sub-frame and of the device-destination store on the GPU."""
import os
import subprocess
import sys

import numpy as np
import pytest

import rsr_b200 as R
from rsr_b200 import scenes
from rsr_b200.subframes import SubframePlan, crop_matrix

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_plan_grid_and_ownership():
    p = SubframePlan(7680, 4320, 8)
    assert (p.nx, p.ny, p.sub_w, p.sub_h) == (4, 4, 1920, 1080)
    assert [len(p.owned_by(r)) for r in range(8)] == [2] * 8
    assert sorted(s.index for r in range(8) for s in p.owned_by(r)) == list(range(16))
    p4 = SubframePlan(3840, 2160, 2)
    assert (p4.nx, p4.ny) == (2, 2) and [s.owner for s in p4.subframes] == [0, 1, 0, 1]
    with pytest.raises(AssertionError):
        SubframePlan(4000, 2160, 1, max_w=4000)          # wider than the reference's guard band
    # finer units (bench.py --split-sub-h 540): 4 x 8 half-height sub-frames, four per rank at 8 ranks
    ph = SubframePlan(7680, 4320, 8, 1920, 540)
    assert (ph.nx, ph.ny, ph.sub_w, ph.sub_h) == (4, 8, 1920, 540) and [len(ph.owned_by(r)) for r in range(8)] == [4] * 8
    img = {s_.index: np.full((s_.height, s_.width), s_.index, np.uint32) for s_ in ph.subframes}
    whole = ph.assemble(img)
    assert whole.shape == (4320, 7680) and all(int(whole[s_.y0, s_.x0]) == s_.index and int(whole[s_.y0 + 539, s_.x0 + 1919]) == s_.index for s_ in ph.subframes)


def test_crop_matrix_maps_subwindow_to_full_ndc():
    nx, ny = 4, 2
    for gx in range(nx):
        for gy in range(ny):
            m = crop_matrix(gx, gy, nx, ny).astype(np.float64)
            left, right = -1 + 2 * gx / nx, -1 + 2 * (gx + 1) / nx
            top, bottom = 1 - 2 * gy / ny, 1 - 2 * (gy + 1) / ny
            for (x, y, ex, ey) in ((left, top, -1, 1), (right, bottom, 1, -1)):
                v = m @ np.array([x * 3.0, y * 3.0, 0.5, 3.0])
                assert np.allclose([v[0] / v[3], v[1] / v[3]], [ex, ey])


def test_subframes_assemble_like_the_whole_frame_on_the_reference(ref_gpu):
    """sub-frames are not bit-identical to one big render (different float rounding), but must tile it"""
    plan = SubframePlan(640, 360, 1, max_w=320, max_h=180)
    sc = scenes.WavyGridScene(n=30)
    P = scenes.perspective(45.0, 640 / 360, 1.0, 100.0)
    imgs = {}
    for s in plan.subframes:
        out = np.zeros((s.height, s.width), np.uint32)
        sc.record(ref_gpu, (s.width, s.height), out, proj=plan.projection(P, s))
        ref_gpu.Run()
        imgs[s.index] = out
    whole = np.zeros((360, 640), np.uint32)
    sc.record(ref_gpu, (640, 360), whole)
    ref_gpu.Run()
    assert np.count_nonzero(plan.assemble(imgs) != whole) < 0.05 * whole.size


GLOO_WORKER = r'''
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from rsr_b200.subframes import SubframePlan
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
plan = SubframePlan(64, 32, world, max_w=32, max_h=16)
mine = plan.owned_by(rank)
# each rank "renders" its sub-frames: pixel value encodes (sub-frame index, y, x)
local = torch.stack([torch.from_numpy((s.index * 100000 + np.add.outer(np.arange(s.height) * 100, np.arange(s.width))).astype(np.int32)) for s in mine])
gathered = [torch.zeros_like(local) for _ in range(world)] if rank == 0 else None
dist.gather(local, gathered, dst=0)
if rank == 0:
    imgs = {}
    for r, part in enumerate(gathered):
        for k, s in enumerate(plan.owned_by(r)):
            imgs[s.index] = part[k].numpy().view(np.uint32)
    full = plan.assemble(imgs)
    for s in plan.subframes:
        blk = full[s.y0:s.y0 + s.height, s.x0:s.x0 + s.width].astype(np.int64)
        assert blk[0, 0] == s.index * 100000 and blk[-1, -1] == s.index * 100000 + (s.height - 1) * 100 + s.width - 1
    print("GATHER_OK", full.shape)
dist.destroy_process_group()
'''


def test_two_rank_gather_with_gloo(tmp_path):
    """the N>1 path (ownership, gather to the presenting rank, assembly) with world_size 2 on CPU"""
    script = tmp_path / "worker.py"
    script.write_text(GLOO_WORKER)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29617", str(script), ROOT], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "GATHER_OK (32, 64)" in r.stdout


@pytest.mark.gpu
def test_every_subframe_matches_the_reference(ref_gpu, cuda_gpu):
    plan = SubframePlan(1280, 720, 1, max_w=640, max_h=360)
    sc = scenes.CubesScene(instances=200)
    P = scenes.perspective(60.0, 1280 / 720, 1.0, 200.0)
    for s in plan.subframes:
        a, b = np.zeros((s.height, s.width), np.uint32), np.zeros((s.height, s.width), np.uint32)
        sc.record(ref_gpu, (s.width, s.height), a, proj=plan.projection(P, s)); ref_gpu.Run()
        sc.record(cuda_gpu, (s.width, s.height), b, proj=plan.projection(P, s)); cuda_gpu.Run()
        assert np.array_equal(a, b), f"sub-frame {s.index}: {np.count_nonzero(a != b)} pixels differ"
        assert cuda_gpu.stats()["triangles_clipped"] > 0 or s.index >= 0


@pytest.mark.gpu
def test_device_destination_store(cuda_gpu):
    import torch
    sc = scenes.BundledLikeScene(cubes=300)
    host = np.zeros((360, 640), np.uint32)
    sc.record(cuda_gpu, (640, 360), host, t=0.2, static=True); cuda_gpu.Run()
    dev = torch.zeros((360, 640), dtype=torch.int32, device="cuda:0")
    sc.record(cuda_gpu, (640, 360), None, t=0.2, static=True, device_out=(dev.data_ptr(), 640)); cuda_gpu.Run()
    assert np.array_equal(dev.cpu().numpy().view(np.uint32), host)


@pytest.mark.gpu
@pytest.mark.parametrize("which", ["c3_small", "c4_small", "c4_front_to_back"])
def test_stress_scenes_match_the_reference(which, ref_gpu, cuda_gpu):
    from parity import assert_identical, render_both
    sc, size = {"c3_small": (scenes.GeometryStressScene(spheres=6, divs=5, size=(640, 360), radius_px=60.0), (640, 360)),
                "c4_small": (scenes.FillStressScene(layers=4, size=(640, 360), tex_dim=512), (640, 360)),
                "c4_front_to_back": (scenes.FillStressScene(layers=4, size=(640, 360), tex_dim=512, front_to_back=True), (640, 360))}[which]
    assert_identical(render_both(sc, size, ref_gpu, cuda_gpu), which)


def test_cost_balanced_ownership():
    """longest-processing-time-first dealing: every sub-frame has exactly one owner and the heaviest rank is no
    worse than with round robin"""
    costs = [5.0, 1.0, 1.0, 5.0, 1.0, 9.0, 9.0, 1.0, 1.0, 9.0, 9.0, 1.0, 5.0, 1.0, 1.0, 5.0]
    for world in (2, 4, 8):
        owners = SubframePlan.balance(costs, world)
        assert len(owners) == 16 and set(owners) == set(range(world))
        load = [sum(c for c, o in zip(costs, owners) if o == r) for r in range(world)]
        rr = [sum(c for i, c in enumerate(costs) if i % world == r) for r in range(world)]
        assert max(load) <= max(rr)
        plan = SubframePlan(7680, 4320, world, 1920, 1080, owners=owners)
        assert sorted(s.index for r in range(world) for s in plan.owned_by(r)) == list(range(16))


@pytest.mark.gpu
def test_whole_draw_frustum_rejection_is_invisible(ref_gpu):
    """sub-frames of a large target see most draws entirely off screen: a draw whose cached bounding box lies beyond one
    guard-band plane is skipped on the host (every one of its triangles would be dropped by the reference's
    'all three vertices share a flag' test, rglv_gpu_impl.hxx:427-433).  Static buffers only; the frames stay
    bit-identical to the reference's, which processes every triangle."""
    g = R.GPU(0)
    try:
        # (1920x1080 sub-frames: the guard band reaches only 13 % beyond the sub-frame, rglv_view_frustum.hxx:36-39)
        plan = SubframePlan(3840, 2160, 1, max_w=1920, max_h=1080)
        sc = scenes.GeometryStressScene(spheres=40, divs=3, size=(3840, 2160), radius_px=60.0)
        culled_before = 0
        total_culled = 0
        for s in plan.subframes:
            a, b = np.zeros((s.height, s.width), np.uint32), np.zeros((s.height, s.width), np.uint32)
            proj = plan.projection(sc.projection(), s)
            sc.record(ref_gpu, (s.width, s.height), a, proj=proj); ref_gpu.Run()
            sc.record(g, (s.width, s.height), b, proj=proj, static=True); g.Run()
            assert np.unique(a).size > 50
            assert np.array_equal(a, b), f"sub-frame {s.index}: {np.count_nonzero(a != b)} pixels differ"
            st = g.stats()
            assert st["triangles_submitted"] == sc.triangles
            total_culled += st["draws_culled"] - culled_before
            culled_before = st["draws_culled"]
        assert total_culled > sc.draws            # on average more than a quarter of the draws per sub-frame
        # dynamic (UPLOAD_ALWAYS) buffers carry no bounding box: nothing is skipped, same pixels
        s = plan.subframes[0]
        a, b = np.zeros((s.height, s.width), np.uint32), np.zeros((s.height, s.width), np.uint32)
        proj = plan.projection(sc.projection(), s)
        sc.record(ref_gpu, (s.width, s.height), a, proj=proj); ref_gpu.Run()
        class NoStatic:        # (the scenes mark their buffers static when the GL object has a stats() method)
            def __getattr__(self, n):
                if n == "stats":
                    raise AttributeError(n)
                return getattr(g, n)
        dyn = NoStatic()
        sc.record(dyn, (s.width, s.height), b, proj=proj); g.Run()
        assert np.array_equal(a, b)
    finally:
        g.close()
