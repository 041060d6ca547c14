#!/usr/bin/env python3
"""bench.py -- frames/s of the frame-rendering hot path on B200 (see DESIGN.md "Measurement").

    python bench.py --gpus N --steps K --warmup W [--workload c2|c4|c3|c4_4k|c3_4k|c5] [--impl reference]

A step is one frame of the workload.  Default workload = BASELINE.json configs[1]: the bundled-scene-shaped frame
sequence at 1920x1080 (`rsr_b200.scenes.BundledLikeScene`: the reference's colortest mesh + seeded synthetic quads and
instanced cubes).  One JSON line is printed by rank 0:
  value      frames/s of the frame SEQUENCE with every input resident in HBM: K distinct frames of a ring whose
             per-frame data exceeds L2, submitted back to back with frame overlap (device timed, CUDA events around the
             K frames), aggregate over ranks (each rank renders its own frames: weak scaling)
  flushed_frame   the same frames one at a time with the L2 flushed before each (cold single-frame latency); the
             roofline's kernel time and the per-stage times come from this leg
  e2e        frames/s through the public API with HOST buffers: per frame the host rebuilds and uploads the instance
             matrices + state and reads the 1080p frame back (>= 200 frames whatever --steps says)
  parity     frame 0 of the workload rendered by the reference's own CPU rasteriser in the same run and compared
  roofline   SURVEY 8(d) algorithmic bytes / the tile kernel's time, against MEASURED_PEAKS.json
  sustained  the sequence leg for >= 2 s with its own clock samples
  c4_4k / c3_4k   (1 GPU) the fill / geometry stress configs of BASELINE.json at 3840x2160 = 2x2 sub-frames
  split_frame     (N > 1) the 7680x4320 frame split over the ranks by sub-frame ownership, NVLink peer stores
`--impl reference` times the reference's own CPU renderer (oracle/_ref) on this box's host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "frames_per_sec_1080p"
SIZE = (1920, 1080)


# ---------------------------------------------------------------------------------------------------------
# workloads
# ---------------------------------------------------------------------------------------------------------

class Workload:
    """a scene + the grid of <= 2048 px sub-frames it is rendered as (1x1 for the 1080p workloads)"""

    def __init__(self, key):
        from rsr_b200 import scenes
        from rsr_b200.subframes import SubframePlan
        self.key = key
        if key in ("c2", "c5"):
            self.scene, self.frame, self.name = scenes.BundledLikeScene(), SIZE, "c2_bundled_like_1920x1080"
        elif key == "c4":
            self.scene, self.frame = scenes.FillStressScene(layers=8, size=SIZE, quads=(2, 2)), SIZE
            self.name = "c4_fill_stress_8layers_bilinear_1to1_1920x1080_subframe"
        elif key == "c3":
            self.scene, self.frame = scenes.GeometryStressScene(spheres=30, divs=6, size=SIZE), SIZE
            self.name = "c3_geometry_stress_2p46Mtris_1920x1080_subframe"
        elif key == "c4_4k":
            self.frame = (3840, 2160)
            self.scene = scenes.FillStressScene(layers=8, size=self.frame, quads=(4, 4), fast_textures=True)
            self.name = "c4_fill_stress_8layers_128textures_bilinear_1to1_3840x2160_as_2x2_subframes"
        elif key == "c3_4k":
            self.frame = (3840, 2160)
            self.scene = scenes.GeometryStressScene(spheres=122, divs=6, size=self.frame)
            self.name = "c3_geometry_stress_9p99Mtris_3840x2160_as_2x2_subframes"
        else:
            raise SystemExit(f"unknown workload {key}")
        self.plan = SubframePlan(self.frame[0], self.frame[1], 1, 1920, 1080)
        self.subframes = self.plan.subframes
        self.sub_size = (self.plan.sub_w, self.plan.sub_h)
        self.single = len(self.subframes) == 1

    def record(self, gl, sf, out, t=0.0, **kw):
        """records sub-frame `sf` of the frame at time t into gl"""
        if self.single:
            self.scene.record(gl, self.sub_size, out, t=t, **kw)
        else:
            self.scene.record(gl, self.sub_size, out, t=t, proj=self.plan.projection(self.scene.projection(), sf), **kw)

    def config(self, args):
        return {"workload": self.name, "triangles_per_frame": self.scene.triangles, "draws_per_frame": getattr(self.scene, "draws", None),
                "width": self.frame[0], "height": self.frame[1], "subframes": len(self.subframes),
                "cache": "L2 flushed (256 MiB memset) before every timed frame",
                "resident_leg": "retained frame tables replayed (rsrcu_replay_frame)" if args.resident == "retained"
                                else "recorded stream decoded and uploaded every step (rsrcu_run_stream)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.path = f"/tmp/rsr_clocks_{os.getpid()}_{time.monotonic_ns()}.csv"
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.gpu)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None
        return self

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            p = [x.strip() for x in line.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1])); mx.append(float(p[2])); pw.append(float(p[3]))
            except ValueError:
                continue
            for n, v in zip(names, p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        try:
            os.remove(self.path)
        except OSError:
            pass
        if sm:
            out["sm_mhz"] = statistics.median(sm)
            out["sm_max_mhz"] = max(mx)
            out["power_w_max"] = max(pw)
        out["reasons"] = sorted(reasons)
        out["samples"] = len(sm)
        return out


def ncu_traffic(workload):
    """dram__bytes_read.sum + dram__bytes_write.sum of tile_kernel from the committed `ncu --set full` capture of
    the same workload (profiles/, tools/capture_profiles.sh; L2 flushed before the captured launch); None if none"""
    import glob
    import re
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", f"r*_tile_kernel_{workload}.txt")))
    if not files:
        return None
    total = 0.0
    for line in open(files[-1]):
        m = re.search(r"dram__bytes_(read|write)\.sum \[(\w+)\] = ([0-9.]+)", line)
        if m:
            total += float(m.group(3)) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[m.group(2)]
    return total or None


# ---------------------------------------------------------------------------------------------------------
# the reference's CPU renderer (cpu_baseline, parity gate, --impl reference)
# ---------------------------------------------------------------------------------------------------------

def time_reference(wl, steps, warmup, threads=None, budget_s=25.0, prime_s=1.0):
    """the reference's multithreaded CPU renderer on this box's cores, method of perf.cxx:216-235: priming frames,
    N timed frames, discard the worst 5 %, report the mean of the rest.  A frame of a > 2048 px workload is the sum
    of its sub-frames (the only way the reference can produce it)."""
    from oracle import refgl
    threads = refgl.init(threads or os.cpu_count())
    g = refgl.RefGPU(double_buffer=True)   # the reference's default: bin of frame N overlaps draw of N-1
    w, h = wl.sub_size
    out = np.zeros((h, w), np.uint32)
    # the geometry stress overflows the reference's unchecked 100 000-byte tile lists (rglv_gpu.hxx:29) at its default
    # 8x8-block tiles (it segfaults): it is timed with 4x4-block tiles, which is also the faster setting for it
    tiles = {"tile_blocks": (4, 4)} if wl.key.startswith("c3") else {}
    refgl.lib().ref_work_start()
    times = []
    try:
        # priming (perf.cxx primes too): a cold box -- first process after boot, worker threads asleep, nothing paged in --
        # times the reference up to 20 % low over its first few frames; it gets `prime_s` of untimed frames before the
        # `warmup` ones so that the figure it is compared by is its steady rate
        t_prime = time.perf_counter()
        i = 0
        while time.perf_counter() - t_prime < prime_s:
            for sf in wl.subframes:
                wl.record(g, sf, out, t=i / 60.0, **tiles)
                g.Run(manage_workers=False)
            i += 1
        t_begin = time.perf_counter()
        for i in range(warmup + steps):
            dt = 0.0
            for sf in wl.subframes:
                wl.record(g, sf, out, t=i / 60.0, **tiles)
                t0 = time.perf_counter()
                g.Run(manage_workers=False)
                dt += time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
            if time.perf_counter() - t_begin > budget_s and len(times) >= 3:
                break
    finally:
        refgl.lib().ref_work_end()
        g.close()
    times.sort()
    keep = times[:max(1, int(len(times) * 0.95))]
    ms = 1e3 * sum(keep) / len(keep)
    return {"ms_per_frame": ms, "fps": 1e3 / ms, "frames": len(times), "threads": threads}


def time_reference_isolated(key, steps, warmup, budget_s):
    """time_reference in a process of its own (`bench.py --impl reference`): the reference is timed with none of this
    process's threads (CUDA, submission pool) beside it, and a fault inside the reference's unchecked CPU code -- its tile
    lists are fixed 100 000-byte arrays, rglv_gpu.hxx:29 -- costs the baseline figure, not the measurement"""
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "LOCAL_WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT")}
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--workload", key, "--steps", str(steps),
           "--warmup", str(warmup), "--ref-budget", str(budget_s)]
    p = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=300)
    lines = [ln for ln in p.stdout.strip().splitlines() if ln.startswith("{")]
    if p.returncode != 0 or not lines:
        raise RuntimeError(f"reference process ended with code {p.returncode}: {p.stderr.strip()[-200:]}")
    d = json.loads(lines[-1])
    return {"ms_per_frame": d["ms_per_step"], "fps": d["value"], "frames": d["steps"], "threads": d["cpu_baseline"]["cores"]}


def parity_gate(wl, gpu):
    """frame 0 of the workload (every sub-frame) through the reference's CPU rasteriser and through the CUDA path on
    the same recorded calls: differing pixels and the largest 8-bit channel difference"""
    from oracle import refgl
    refgl.init(os.cpu_count())
    ref = refgl.RefGPU()
    w, h = wl.sub_size
    diff, max_lsb, pixels = 0, 0, 0
    try:
        for sf in wl.subframes:
            a, b = np.zeros((h, w), np.uint32), np.zeros((h, w), np.uint32)
            wl.record(ref, sf, a, t=0.0, tile_blocks=(4, 4))   # (4x4-block reference tiles keep its unchecked 100 000-byte lists in bounds; results do not depend on the tile size)
            ref.Run()
            wl.record(gpu, sf, b, t=0.0, tile_blocks=(4, 4), static=True)
            gpu.Run()
            diff += int(np.count_nonzero(a != b))
            pixels += a.size
            for s in (0, 8, 16):
                max_lsb = max(max_lsb, int(np.abs(((a >> s) & 0xff).astype(np.int32) - ((b >> s) & 0xff).astype(np.int32)).max()))
    finally:
        ref.close()
    return {"diff_pixels": diff, "max_lsb": max_lsb, "pixels_compared": pixels,
            "against": "the unmodified reference (oracle/_ref/librsr_ref.so) on this box's CPU, frame 0 of this workload"}


# ---------------------------------------------------------------------------------------------------------
# device-resident leg + roofline for one workload on this rank's GPU
# ---------------------------------------------------------------------------------------------------------

def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


class ResidentRun:
    """the workload's frame recorded once with every input static; a step replays it (all kernels run in full)"""

    def __init__(self, wl, gpu, args, torch, local_rank):
        self.wl, self.gpu, self.torch = wl, gpu, torch
        self.stream = torch.cuda.ExternalStream(gpu.stream(), device=local_rank)
        self.recs, self.retained = [], []
        for sf in wl.subframes:
            wl.record(gpu, sf, None, t=0.0, static=True)
            self.recs.append(gpu.Finish())
            if args.resident == "retained":
                gpu.Submit(self.recs[-1])
                self.retained.append(gpu.Retain())
        gpu.set_overlap(not wl.single)   # sub-frames of one frame: front end of the next one under the current tile kernel

    def step(self):
        if self.retained:
            for fr in self.retained:
                self.gpu.Replay(fr)
        else:
            for rec in self.recs:
                self.gpu.Submit(rec, sync=False)

    def close(self):
        self.gpu.Sync()
        for fr in self.retained:
            self.gpu.Release(fr)
        self.gpu.set_overlap(False)


class SequenceRun:
    """the frame-sequence leg: a ring of `ring` DISTINCT frames of the workload's animation (t = i / 60 s), each retained
    with its own state / draw tables, its own per-frame inputs (instance matrices) and its own output buffer, replayed
    back to back with frame overlap on (the front end of frame N+1 runs under the tile kernel of frame N -- the
    counterpart of the reference's doubleBuffer mode, which bins frame N+1 while its workers draw frame N).  The ring's
    per-frame data (outputs 8.3 MB + inputs each) exceeds the 126 MB L2, so no frame finds its own data cached; the
    scene-static meshes and textures are shared by the frames, as in any frame sequence."""

    def __init__(self, wl, gpu, torch, local_rank, ring):
        self.wl, self.gpu, self.torch = wl, gpu, torch
        self.stream = torch.cuda.ExternalStream(gpu.stream(), device=local_rank)
        self.retained = []
        for i in range(ring):
            wl.record(gpu, wl.subframes[0], None, t=i / 60.0, static=True)
            gpu.Submit(gpu.Finish())
            self.retained.append(gpu.Retain())
        self.pos = 0
        gpu.set_overlap(True)

    def frames(self, n):
        for _ in range(n):
            self.gpu.Replay(self.retained[self.pos % len(self.retained)])
            self.pos += 1

    def timed(self, n):
        """device time of n frames submitted back to back (CUDA events on the context's stream)"""
        torch = self.torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(self.stream):
            e0.record(self.stream)
        self.frames(n)
        self.gpu.Join()   # (overlap mode: tile kernels alternate between two streams)
        with torch.cuda.stream(self.stream):
            e1.record(self.stream)
        self.gpu.Sync()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    def close(self):
        self.gpu.Sync()
        for fr in self.retained:
            self.gpu.Release(fr)
        self.gpu.set_overlap(False)


def sequence_sustained(seq, seconds, local_rank):
    """the sequence leg for `seconds` (no host sync except to bound the queue): what a long run clocks at"""
    torch = seq.torch
    sampler = ClockSampler(local_rank).start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 0
    t0 = time.perf_counter()
    with torch.cuda.stream(seq.stream):
        e0.record(seq.stream)
    while time.perf_counter() - t0 < seconds:
        seq.frames(64)
        n += 64
        if n % 256 == 0:
            seq.gpu.Sync()   # bound the queue depth
    seq.gpu.Join()
    with torch.cuda.stream(seq.stream):
        e1.record(seq.stream)
    seq.gpu.Sync()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    return {"seconds": e0.elapsed_time(e1) / 1e3, "steps": n, "ms_per_step": ms, "frames_per_s": 1e3 / ms, "clocks": sampler.stop()}


def measure_resident(run, flush, steps, warmup, barrier, sampler=None):
    """per-step device time (CUDA events on the context's stream, L2 flushed before every step), tile-kernel time
    and frame statistics"""
    torch, gpu, stream = run.torch, run.gpu, run.stream
    gpu.set_profiling(1)
    for i in range(warmup):
        run.step()
        gpu.Sync()
    barrier()
    if sampler:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    tile_ms = []
    stats = None
    nsub = len(run.wl.subframes)
    for i in range(steps):
        with torch.cuda.stream(stream):
            flush.fill_(i & 0xff)
            ev[i][0].record(stream)
        if nsub == 1:
            run.step()
            gpu.Join()
            with torch.cuda.stream(stream):
                ev[i][1].record(stream)
            gpu.Sync()
            tile_ms.append(gpu.stage_ms()["tile"])
        else:
            run.step()
            gpu.Join()
            with torch.cuda.stream(stream):
                ev[i][1].record(stream)
            gpu.Sync()
        stats = gpu.stats()
    barrier()
    clocks = sampler.stop() if sampler else None
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    # per-stage breakdown and per-sub-frame statistics: a few extra, untimed frames with an event after every kernel
    gpu.set_profiling(2)
    stage_acc, nstage = {}, 4
    frags = entries = inbytes = 0
    tile_sum = 0.0
    for i in range(nstage):
        for k in range(nsub):
            with torch.cuda.stream(stream):
                flush.fill_(i & 0xff)
            if run.retained:
                gpu.Replay(run.retained[k])
            else:
                gpu.Submit(run.recs[k], sync=False)
            gpu.Sync()
            st = gpu.stats()
            if i == 0:
                frags += st["fragments_shaded"]; entries += st["bin_entries"]; inbytes += st["input_bytes"]
            for name, v in gpu.stage_ms().items():
                stage_acc[name] = stage_acc.get(name, 0.0) + v / nstage
    gpu.set_profiling(0)
    tile_avg = sum(tile_ms) / len(tile_ms) if tile_ms else stage_acc.get("tile", 0.0)
    return {"dev_ms": dev_ms, "tile_ms": tile_avg, "stage_ms": stage_acc, "clocks": clocks, "stats": stats,
            "fragments": frags, "bin_entries": entries, "input_bytes": inbytes}


def roofline_of(wl, m, ms_per_step):
    """B = 4 W H (resolved output) + draw inputs (bound SoA floats x vertices + indices + instance
    matrices) + 2 x 6 R (bin entries written then read) + 16 U (unique texels at the selected LOD), per frame;
    achieved = B / the tile kernel's time (summed over the frame's sub-frames), measured with CUDA events on the
    context's stream in this run"""
    pk = peaks()
    peak = float(pk.get("hbm_gbs", 6650.0))
    W, H = wl.frame
    nsub = len(wl.subframes)
    texels = wl.scene.unique_texels({"fragments_shaded": m["fragments"]}, wl.frame)
    B = 4 * W * H + m["input_bytes"] + 12 * m["bin_entries"] + 16 * texels
    kernel_ms = m["tile_ms"]
    achieved = B / (kernel_ms * 1e-3) / 1e9
    return {"bound": "hbm", "kernel": "tile_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "frac_of_whole_step": B / (ms_per_step * 1e-3) / 1e9 / peak,
            "traffic": ncu_traffic(wl.key), "algorithmic_bytes_per_launch": B,
            "algorithmic_bytes": {"output_4WH": 4 * W * H, "draw_inputs": m["input_bytes"], "bin_entries_12R": 12 * m["bin_entries"],
                                  "unique_texels_16U": 16 * texels},
            "kernel_ms": kernel_ms, "kernel_launches_per_step": nsub,
            "peak_source": "MEASURED_PEAKS.json hbm_gbs (burst copy)" if pk else "fallback 6650",
            "stage_ms": m["stage_ms"]}


def sustained_leg(run, flush, seconds, local_rank):
    """the same step back to back for `seconds` (no per-step host sync): what a long run clocks at"""
    torch, gpu, stream = run.torch, run.gpu, run.stream
    sampler = ClockSampler(local_rank).start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fl0, fl1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # the flush's own time is measured first and subtracted: the timed region is one unbroken stream of work
    torch.cuda.synchronize()
    with torch.cuda.stream(stream):
        fl0.record(stream)
        for i in range(20):
            flush.fill_(i)
        fl1.record(stream)
    torch.cuda.synchronize()
    flush_ms = fl0.elapsed_time(fl1) / 20
    n = 0
    t0 = time.perf_counter()
    with torch.cuda.stream(stream):
        e0.record(stream)
    while time.perf_counter() - t0 < seconds:
        for _ in range(16):
            with torch.cuda.stream(stream):
                flush.fill_(n & 0xff)
            run.step()
            n += 1
        if n % 64 == 0:
            torch.cuda.synchronize()   # bound the queue depth
    gpu.Join()
    with torch.cuda.stream(stream):
        e1.record(stream)
    gpu.Sync()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    clocks = sampler.stop()
    return {"seconds": e0.elapsed_time(e1) / 1e3, "steps": n, "ms_per_step_incl_flush": ms, "flush_ms": flush_ms,
            "ms_per_step": ms - flush_ms, "frames_per_s": 1e3 / max(ms - flush_ms, 1e-6), "clocks": clocks}


def e2e_leg(wl, gpu, frames, torch, dist, world, local_rank, barrier, contexts=2, reps=5):
    """end to end through the public API with HOST buffers.  Each frame of the sequence is recorded beforehand (that
    is the callers' job in the reference: node graph -> GL calls); the timed region is what replaces GPU::Run -- decode
    the recorded stream (rsrcu_run_stream), upload that frame's host buffers (instance matrices, state), kernels, read
    the frame back into pinned host memory (rsrcu_sync_frame).  Frames are submitted through rsr_b200.SubmitPool:
    `contexts` rendering contexts on this GPU, one host thread each, frames dealt round robin, three frames in flight per
    context, frame overlap on -- the reference renders with every host core (its job system); this leg uses `contexts`
    host threads.  Every frame is waited for and lands in its own slot of a ring of pinned host buffers."""
    import rsr_b200
    W, H = wl.sub_size
    K = max(1, contexts)
    pool = rsr_b200.SubmitPool(local_rank, contexts=K, overlap=True)
    recorder = pool.gpus[0]
    host_out = [[torch.empty((H, W), dtype=torch.int32).pin_memory().numpy().view(np.uint32) for _ in range(3)] for _ in range(K)]
    distinct = 33                       # per context: a 33-frame loop of the sequence (each frame with its own instance matrices)
    recs = []
    for k in range(K):
        row = []
        for j in range(distinct):
            wl.record(recorder, wl.subframes[0], host_out[k][j % 3], t=(j * K + k) / 60.0)
            row.append(recorder.Finish())
        recs.append(row)
    frame = lambda i: recs[i % K][(i // K) % distinct]   # frame i goes to context i % K as its (i // K)-th frame -> buffer (i // K) % 3
    for i in range(K * distinct):       # warm every context with one loop of the sequence (static uploads, buffer growth, every recorded frame's pages touched)
        ticket = pool.submit(frame(i))
        if i >= 3 * K:
            pool.wait(ticket - 3 * K)
    pool.drain()
    retried0 = sum(st["frames_retried"] for st in pool.stats())
    # the timed region (`frames` frames, 40 ms of wall clock at 5 k frames/s) is repeated and the MEDIAN repetition is the
    # value: one scheduling hiccup of the host no longer decides the figure; every repetition is listed in the record
    rep_s = []
    base = K * distinct                 # (frame numbers keep counting across repetitions: the pool deals frames round robin by its own count)
    for rep in range(max(1, reps)):
        barrier()
        t0 = time.perf_counter()
        for i in range(frames):
            ticket = pool.submit(frame(base + i))
            if i >= 3 * K:
                pool.wait(ticket - 3 * K)      # bounded queue: at most three frames per context outstanding
        pool.drain()
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        base += frames
        if os.environ.get("RSR_BENCH_DEBUG"):
            print(f"rank {os.environ.get('RANK', '0')}: e2e rep {rep}: {frames} frames in {e2e_s * 1e3:.1f} ms = {frames / e2e_s:.0f} frames/s", file=sys.stderr)
        t = torch.tensor([e2e_s], dtype=torch.float64, device=f"cuda:{local_rank}")
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        rep_s.append(float(t.item()))
    med_s = sorted(rep_s)[len(rep_s) // 2]
    st = pool.stats()[0]
    retried = sum(s_["frames_retried"] for s_ in pool.stats()) - retried0
    pool.close()
    return {"value": world * frames / med_s, "unit": "frames/s", "frames": frames,
            "repetitions_frames_per_s": [round(world * frames / x, 1) for x in rep_s],
            "h2d_bytes_per_step": st["h2d_bytes"], "d2h_bytes_per_step": st["d2h_bytes"],
            "frames_retried_for_overflow": retried, "host_threads": K,
            "timing": f"wall clock over `frames` frames of rsrcu_run_stream + rsrcu_sync_frame (stream decode, H2D, kernels, D2H), median of {len(rep_s)} repetitions, "
                      f"through rsr_b200.SubmitPool: {K} contexts / host threads, three frames in flight each, frame overlap on, max over ranks"}


def stress_subrecord(key, args, gpu, torch, local_rank, flush, barrier):
    """BASELINE.json configs[2] / [3] at their full size on one GPU: a 3840x2160 frame = 2x2 sub-frames of 1920x1080
    (the reference's guard band ends at 2048 px), every sub-frame submits the whole scene"""
    wl = Workload(key)
    par = parity_gate(wl, gpu)
    run = ResidentRun(wl, gpu, args, torch, local_rank)
    steps = max(10, min(args.steps, 30))
    m = measure_resident(run, flush, steps, 3, barrier)
    run.close()
    ms = m["dev_ms"] / steps
    fps = 1e3 / ms
    out = {"workload": wl.name, "frames_per_s": fps, "ms_per_frame": ms, "steps": steps,
           "mtris_per_s": wl.scene.triangles * fps / 1e6, "gpix_per_s": m["fragments"] * fps / 1e9,
           "triangles_per_frame": wl.scene.triangles, "triangle_setups_per_frame": wl.scene.triangles * len(wl.subframes),
           "fragments_per_frame": m["fragments"], "bin_entries_per_frame": m["bin_entries"],
           "parity": par, "roofline": roofline_of(wl, m, ms)}
    if not args.no_cpu_baseline:
        try:
            r = time_reference_isolated(wl.key, 6, 1, 12.0)
            out["cpu_baseline"] = {"value": r["fps"], "unit": "frames/s", "cores": r["threads"], "kind": "reference",
                                   "sample": f"{r['frames']} frames (sum over the 4 sub-frames), doubleBuffer=true"}
        except Exception as exc:
            out["cpu_baseline"] = {"value": None, "sample": f"unavailable: {exc}"}
    return out


# ---------------------------------------------------------------------------------------------------------
# the post chain and the geometry producer that follow / feed a frame on the device
# ---------------------------------------------------------------------------------------------------------

def f_rows_subrecord(gpu, torch, local_rank, flush, args):
    """Kawase blur, glow combine and marching cubes: device time per launch (CUDA events on the context's stream, 256 MiB
    L2 flush before every launch), algorithmic bytes per launch and the fraction of the measured HBM copy peak; each is
    compared with the reference's own function on this box first (bit for bit; marching-cubes normals to 2e-5)."""
    import numpy as np
    stream = torch.cuda.ExternalStream(gpu.stream(), device=local_rank)
    peak = float(peaks().get("hbm_gbs", 0.0)) or 6550.7
    rng = np.random.default_rng(3)
    try:
        from oracle import refgl
        have_ref = refgl.available()
        if have_ref:
            refgl.init(min(16, os.cpu_count() or 1))
    except Exception:
        have_ref = False

    def timed(fn, n=10):
        tot = 0.0
        for i in range(n + 2):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(stream):
                flush.fill_(i & 0xff)
                e0.record(stream)
            fn()
            with torch.cuda.stream(stream):
                e1.record(stream)
            e1.synchronize()
            if i >= 2:
                tot += e0.elapsed_time(e1)
        return tot / n

    def rec(ms, nbytes, **kw):
        gbs = nbytes / (ms * 1e-3) / 1e9
        return dict(ms=ms, algorithmic_bytes_per_launch=int(nbytes), achieved_gbs=gbs, peak_gbs=peak, frac=gbs / peak, **kw)

    out = {"what": f_rows_subrecord.__doc__.split(":")[0].strip(), "peak_source": "MEASURED_PEAKS.json hbm_gbs (burst copy)"}
    # -- `$kawase`: 16 B read + 16 B written per pixel and pass (the 16 overlapping taps are on-chip reuse) -----------
    for w, h in ((960, 540), (3840, 2160)):
        a, b = gpu.Canvas("fp", w, h), gpu.Canvas("fp", w, h)
        src = rng.random((h, w, 4), dtype=np.float32)
        a.write(src)
        par = None
        if have_ref and w == 960:
            gpu.KawaseBlur(a, b, 2)
            par = int(np.count_nonzero(b.read().view(np.uint32) != refgl.kawase_blur(src, 2).view(np.uint32)))
        ms = timed(lambda: gpu.KawaseBlur(a, b, 2))
        out[f"kawase_{w}x{h}"] = rec(ms, 32 * w * h, differing_floats_vs_reference=par)
        a.free(); b.free()
    # -- `$glow` at 1080p: 12 B (r, g, b planes of the quad canvas) + 4 B (one 16-byte blur pixel per quad) read, 4 B written per pixel
    w, h = 1920, 1080
    cq, cb, tc = gpu.Canvas("quads", w, h), gpu.Canvas("fp", w // 2, h // 2), gpu.Canvas("tc", w, h)
    quads = rng.random((h // 2, w // 2, 4, 4), dtype=np.float32)
    blur = rng.random((h // 2, w // 2, 4), dtype=np.float32)
    cq.write(quads); cb.write(blur)
    par = None
    if have_ref:
        gpu.Glow(cq, cb, tc, True)
        par = int(np.count_nonzero(tc.read() != refgl.glow_filter(quads, blur, True)))
    ms = timed(lambda: gpu.Glow(cq, cb, tc, True))
    out["glow_1920x1080"] = rec(ms, 20 * w * h, differing_pixels_vs_reference=par)
    for c in (cq, cb, tc):
        c.free()
    # -- `$mc`: precision 128, forkDepth 2 = 64 blocks of 32^3 cells; a blocking call (count pass, scan, counts to the host, emit pass)
    t, precision, fork, rngv = 1.25, 128, 2, 5.0
    gpu.MarchSurface(t, precision, fork, rngv)
    t0 = time.perf_counter()
    reps = 5
    for _ in range(reps):
        soa, blocks, total = gpu.MarchSurface(t, precision, fork, rngv)
    wall_ms = (time.perf_counter() - t0) / reps * 1e3
    tris = sum(n for _, n in blocks) // 3
    mc = {"precision": precision, "fork_depth": fork, "cells": precision ** 3, "triangles": tris, "ms_wall_per_call": wall_ms,
          "mcells_per_s": precision ** 3 / wall_ms / 1e3, "mtris_per_s": tris / wall_ms / 1e3}
    if have_ref:
        t0 = time.perf_counter()
        pos, nrm, rblocks = refgl.march_surface(t, precision, fork, rngv)
        ref_ms = (time.perf_counter() - t0) * 1e3
        got = [gpu.read_device(p, total) for p in soa]
        mc["parity"] = {"blocks_equal": rblocks == blocks,
                        "differing_position_floats": int(sum(np.count_nonzero(got[k].view(np.uint32) != pos[k].view(np.uint32)) for k in range(3))),
                        "max_normal_error": float(max(np.abs(got[3 + k] - nrm[k]).max() for k in range(3))), "normal_tolerance": 2e-5}
        mc["cpu_baseline"] = {"ms": ref_ms, "cores": 1, "kind": "reference",
                              "sample": "rglv::march_sdf_vao driven block after block by oracle/ref_harness.cpp on one core (the node spreads its 64 block jobs over the job system)"}
    out["march_surface"] = mc
    return out


def bundled_scenes_subrecord(seconds=2.5):
    """SURVEY 8(f)1 / BASELINE.json configs[1]: the reference's bundled data/scene files, unmodified, through its own
    front-end (Lua scene -> JSON -> node graph -> rglv::GL; oracle/ref_scene.cpp = perf.cxx's loop, doubleBuffer on) on
    the CPU reference and -- same node graph, GPU::RunImpl replaced by the C-ABI binding -- on the GPU, 1920x1080.  Each
    run is a process of its own (tools/scene_bench.py).  The node graph and the GL recording run on the host in both; the
    ratio is what a user of the reference sees when librsr's RunImpl (and, for glow scenes, the three post nodes of
    rsr_b200/host/post_nodes_cuda.cxx) change and nothing else.  dropin: every pointer GL recorded is staged once per
    frame (exact under the reference's own contract); dropin_static_assets: textures and index arrays declared immutable
    by the host (uploaded once), which holds for these three scenes; dropin_pin_in_place: arrays of 1 MiB or more copied
    by the copy engine from where they lie (rsrcu_set_pin_in_place: for hosts whose assets outlive their use)."""
    out = {"what": "bundled data/scene files through the reference's own node graph: CPU reference vs drop-in GPU::RunImpl, 1920x1080, doubleBuffer on",
           "scenes": {}}
    tool = os.path.join(ROOT, "tools", "scene_bench.py")
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "LOCAL_WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT")}
    for scene in ("colortest", "tucker-and-dino", "instanced-cubes"):
        rec = {}
        for lib in ("ref", "dropin", "dropin_static", "dropin_pinned"):
            if lib == "dropin_pinned" and scene != "tucker-and-dino":
                continue   # (only that scene binds arrays of 1 MiB or more per frame)
            try:
                cmd = [sys.executable, tool, "--lib", lib.split("_")[0], "--scene", scene, "--seconds", str(seconds)]
                if lib == "dropin_static":
                    cmd.append("--static-assets")
                env["RSRCU_PIN_IN_PLACE"] = "1" if lib == "dropin_pinned" else "0"
                for attempt in range(3):   # (the reference's binner occasionally faults on its own scenes: oracle/scene_ref.py)
                    p = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=120)
                    lines = [ln for ln in p.stdout.splitlines() if ln.startswith("{")]
                    if p.returncode == 0 and lines:
                        break
                if p.returncode != 0 or not lines:
                    raise RuntimeError(f"exit code {p.returncode}: {p.stderr.strip()[-160:]}")
                rec[lib] = json.loads(lines[-1])
            except Exception as exc:
                rec[lib] = {"error": repr(exc)}
        if "frames_per_s" in rec.get("ref", {}) and "frames_per_s" in rec.get("dropin", {}):
            out["scenes"][scene] = {"reference_frames_per_s": rec["ref"]["frames_per_s"], "cores": rec["ref"]["threads"],
                                    "dropin_frames_per_s": rec["dropin"]["frames_per_s"],
                                    "ratio": rec["dropin"]["frames_per_s"] / rec["ref"]["frames_per_s"],
                                    "dropin_static_assets_frames_per_s": rec.get("dropin_static", {}).get("frames_per_s"),
                                    "dropin_pin_in_place_frames_per_s": rec.get("dropin_pinned", {}).get("frames_per_s"),
                                    "frames_timed": [rec["ref"]["frames"], rec["dropin"]["frames"]],
                                    "frame_identical": rec["ref"]["crc32_frame_t1"] == rec["dropin"]["crc32_frame_t1"]}
        else:
            out["scenes"][scene] = rec
    return out


# ---------------------------------------------------------------------------------------------------------
# C5: 8K split-frame
# ---------------------------------------------------------------------------------------------------------

def split_frame(args, gpu, torch, dist, rank, local_rank, world, flush, barrier, scene=None):
    """C5 (BASELINE.json configs[4]): 7680x4320 frames = 4x4 sub-frames of 1920x1080 (the reference's guard band ends at
    2048 px), sub-frames dealt to the ranks; a step is a BATCH of frames submitted back to back.  Strong scaling: the
    frames are fixed.  Every rank's tile kernel resolves straight into the presenting GPU's frame buffer (CUDA-IPC
    mapping, peer stores over NVLink while it rasterises).  The presenter's frame is double buffered; behind its
    sub-frames of frame f every rank adds 1 to a counter in the presenter's memory, and before it starts frame f + 1
    it waits (on its stream, no host involvement, no collective) until frame f - 1 is complete on all ranks -- the
    buffer it is about to overwrite has been presented.  The presenter waits for the last frame of the batch.
    Rank 0 also renders the same batches alone (all 16 sub-frames on one GPU): the line carries its own single-GPU
    figure and efficiency.  --barrier nccl: an NCCL all-reduce per frame instead of the counters; --exchange nccl:
    render locally, NCCL gather, assemble (both single-buffered, for comparison)."""
    from rsr_b200 import scenes
    from rsr_b200.present import PresentedFrame
    from rsr_b200.subframes import SubframePlan
    scene = scene or scenes.BundledLikeScene()
    dev = f"cuda:{local_rank}"
    stream = torch.cuda.ExternalStream(gpu.stream(), device=local_rank)
    P = scenes.perspective(45.0, 7680 / 4320, 1.0, 400.0)
    SH = args.split_sub_h
    plan = SubframePlan(7680, 4320, world, 1920, SH)
    p2p = args.exchange == "p2p"
    flags = p2p and args.barrier == "flag"
    batch = max(1, args.batch_frames)
    nbuf = 2 if flags else 1
    gpu.set_overlap(True)            # sub-frames are independent frames: front end of the next one under the current tile kernel
    cur = torch.cuda.current_stream()
    pf = PresentedFrame(7680, 4320, rank, local_rank, world, dist, buffers=nbuf)
    gpu.EnablePeerAccess(pf.presenter_device)
    token = torch.zeros(1, dtype=torch.int32, device=dev)

    def retain(subframes, buf=0, local=None):
        out = []
        for k, sf in enumerate(subframes):
            target = (pf.pointer(sf.x0, sf.y0, buf), pf.stride_px) if local is None else (local[k].data_ptr(), 1920)
            scene.record(gpu, (sf.width, sf.height), None, t=0.0, static=True, proj=plan.projection(P, sf), device_out=target)
            gpu.Submit(gpu.Finish())
            out.append(gpu.Retain())
        return out

    # ---- ownership ------------------------------------------------------------------------------------------------
    owners = None
    costs_ms = None
    if world > 1 and args.balance == "cost":
        # sub-frames differ in cost (screen centre vs corners): rank 0 measures the tile kernel of each one (retained
        # replays, best of 3) and deals them longest first to the least loaded rank; every rank uses the same table
        costs = torch.zeros(len(plan.subframes), dtype=torch.float64, device=dev)
        if rank == 0:
            gpu.set_profiling(1)
            probes = retain(plan.subframes)
            for k, fr in enumerate(probes):
                best = 1e9
                for _ in range(3):
                    gpu.Replay(fr, sync=True)
                    best = min(best, gpu.stage_ms()["tile"])
                costs[k] = best
            for fr in probes:
                gpu.Release(fr)
            gpu.set_profiling(0)
        dist.broadcast(costs, src=0)
        costs_ms = [round(float(c), 4) for c in costs.tolist()]
        owners = SubframePlan.balance(costs_ms, world)
        plan = SubframePlan(7680, 4320, world, 1920, SH, owners=owners)

    frames_done = 0      # frames launched so far in the current phase (same arithmetic on every rank)

    def run_batch(sets, nranks, presenter):
        """one batch: `sets[b]` = this rank's retained sub-frames that resolve into presenter buffer b"""
        nonlocal frames_done
        for _ in range(batch):
            f = frames_done
            if flags:
                if f >= 2:
                    gpu.WaitCounters(pf.counter_pointer(0), nranks, f - 1)      # every rank has finished frame f - 2: its buffer is free
                for fr in sets[f % nbuf]:
                    gpu.Replay(fr)
                gpu.SignalCounter(pf.counter_pointer(rank))
            else:
                for fr in sets[0]:
                    gpu.Replay(fr)
                done = torch.cuda.Event()
                gpu.Join()
                with torch.cuda.stream(stream):
                    done.record(stream)
                cur.wait_event(done)                       # NCCL runs on torch's stream, after the render stream
                if p2p:
                    if nranks > 1:
                        dist.all_reduce(token)             # every rank's kernels (and with them their peer stores) have completed
                else:
                    if nranks > 1:
                        dist.gather(local, gathered, dst=0)
                    if rank == 0:
                        parts = gathered if nranks > 1 else [local]
                        for r, part in enumerate(parts):
                            for k, sf in enumerate(plan.owned_by(r) if nranks > 1 else plan.subframes):
                                pf.local[0, sf.y0:sf.y0 + sf.height, sf.x0:sf.x0 + sf.width].copy_(part[k])
            frames_done += 1
        if flags:
            if presenter:
                gpu.WaitCounters(pf.counter_pointer(0), nranks, frames_done)   # the last frame of the batch is complete on every rank
            done = torch.cuda.Event()
            gpu.Join()
            with torch.cuda.stream(stream):
                done.record(stream)
            cur.wait_event(done)

    def timed(sets, nranks, presenter, steps, warmup, sync_all):
        for _ in range(warmup):
            run_batch(sets, nranks, presenter)
        sync_all()
        t_ms = 0.0
        for i in range(steps):
            flush.fill_(i & 0xff)
            sync_all()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(stream):
                e0.record(stream)
            run_batch(sets, nranks, presenter)
            e1.record(cur)
            torch.cuda.synchronize()
            t_ms += e0.elapsed_time(e1)
            if os.environ.get("RSR_BENCH_DEBUG"):
                print(f"rank {rank} step {i}: {e0.elapsed_time(e1):.3f} ms", file=sys.stderr)
        sync_all()
        return t_ms / steps

    steps = max(5, min(args.steps, 50))
    # ---- the whole frame on ONE GPU (rank 0 alone): the strong-scaling baseline of this very run -------------
    single_ms = None
    checksum1 = None
    if rank == 0:
        alone = [retain(plan.subframes, b) for b in range(nbuf)] if p2p else None
        if not p2p:
            local = torch.zeros((len(plan.subframes), SH, 1920), dtype=torch.int32, device=dev)
            alone = [retain(plan.subframes, 0, local)]
        gpu.Sync()
        single_ms = timed(alone, 1, True, steps, 3, torch.cuda.synchronize) / batch
        gpu.Sync()
        checksum1 = int(pf.local[0].to(torch.int64).sum().item())
        for sset in alone:
            for fr in sset:
                gpu.Release(fr)
        pf.local.zero_()
        pf.local_counter.zero_()
        torch.cuda.synchronize()
    barrier()

    # ---- split over the ranks ------------------------------------------------------------------------------------
    frames_done = 0
    mine = plan.owned_by(rank)
    if p2p:
        sets = [retain(mine, b) for b in range(nbuf)]
    else:
        local = torch.zeros((len(mine), SH, 1920), dtype=torch.int32, device=dev)
        sets = [retain(mine, 0, local)]
        gathered = [torch.zeros_like(local) for _ in range(world)] if (rank == 0 and world > 1) else None
    gpu.Sync()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_local = timed(sets, world, rank == 0, steps, max(args.warmup, 3), barrier)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms_local], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / batch
    gpu.Sync()
    st = gpu.stats()
    for sset in sets:
        for fr in sset:
            gpu.Release(fr)
    gpu.set_overlap(False)
    rec = None
    if rank == 0:
        checksum = int(pf.local[0].to(torch.int64).sum().item())
        if world == 1:
            exchange = "none"
        elif flags:
            exchange = ("tile kernels store into the presenting GPU's double-buffered frame over NVLink (CUDA IPC peer memory); completion = a counter in the "
                        "presenter's memory that every rank increments behind its sub-frames of a frame (system-scope atomic over NVLink) and every rank's stream "
                        "waits on before it reuses a buffer: no host barrier, no collective")
        elif p2p:
            exchange = "tile kernels store into the presenting GPU's frame over NVLink (CUDA IPC peer memory) + 4-byte NCCL all-reduce per frame as barrier"
        else:
            exchange = "NCCL gather of resolved sub-frames to rank 0 + assembly copies"
        remote = sum(1 for s in plan.subframes if s.owner != 0)
        rec = {"metric": "frames_per_sec_8k_split_frame", "value": 1e3 / ms, "unit": "frames/s", "n_gpus": world, "steps": steps,
               "frames_per_step": batch, "ms_per_step": ms * batch, "ms_per_frame": ms, "scaling": "strong",
               "single_gpu_frames_per_s": 1e3 / single_ms, "single_gpu_ms_per_frame": single_ms,
               "speedup_vs_single_gpu": single_ms / ms, "efficiency": single_ms / ms / world,
               "nvlink_bytes_per_frame": remote * 1920 * SH * 4,
               "config": {"workload": f"c5_8k_split_frame_{plan.nx}x{plan.ny}_subframes_of_c2", "width": 7680, "height": 4320,
                          "subframes_per_rank": len(mine), "exchange": exchange,
                          "ownership": "round robin" if owners is None else f"cost balanced (longest first): {owners}",
                          "subframe_tile_ms": costs_ms,
                          "submission": f"batches of {batch} frames back to back, retained sub-frame tables replayed",
                          "cache": "L2 flushed before every timed batch"},
               "mtris_per_s": len(plan.subframes) * scene.triangles * 1e3 / ms / 1e6, "clocks": clocks,
               "frame_checksum": checksum, "frame_checksum_single_gpu": checksum1, "checksums_equal": checksum == checksum1,
               "gpu_launches": int(st["kernel_launches"]) * len(mine) * steps * batch}
    return rec


def pin_rank_to_cores(local_rank, world):
    """one process per GPU on one host: give every rank its own slice of the cores (submit thread, CUDA's helper
    threads and the interpreter of eight ranks otherwise migrate over the same cores)"""
    try:
        cores = sorted(os.sched_getaffinity(0))
        per = max(1, len(cores) // max(world, 1))
        mine = cores[local_rank * per:(local_rank + 1) * per] or cores
        os.sched_setaffinity(0, mine)
        return len(mine)
    except Exception:
        return None


def finish(code=0):
    """the line is out: flush and leave without running interpreter / library teardown (worker threads of the reference's
    job system, CUDA contexts of several host threads: nothing there is worth a crash after the measurement)"""
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(code)


def main():
    import faulthandler
    faulthandler.enable()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-subrecords", action="store_true", help="skip the c4_4k / c3_4k (1 GPU) and split_frame (N > 1) sub-records")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--e2e-frames", type=int, default=200)
    ap.add_argument("--e2e-reps", type=int, default=5, help="repetitions of the end-to-end timed region; the median is reported")
    ap.add_argument("--e2e-contexts", type=int, default=0, help="submission threads / contexts of the end-to-end leg (0: 2 at one or two GPUs, 1 per rank beyond: the ranks share the host's cores)")
    ap.add_argument("--sustained-seconds", type=float, default=2.0)
    ap.add_argument("--ref-budget", type=float, default=25.0, help="--impl reference: stop after this many seconds of timed frames (at least 3 frames)")
    ap.add_argument("--ring", type=int, default=24, help="distinct frames in the frame-sequence leg's ring")
    ap.add_argument("--resident", default="retained", choices=["retained", "stream"],
                    help="device-resident leg: replay retained frame tables, or decode + upload the recorded stream every step")
    ap.add_argument("--balance", default="cost", choices=["cost", "roundrobin"], help="split-frame: how sub-frames are dealt to the ranks")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"], help="split-frame: how resolved pixels reach the presenting GPU")
    ap.add_argument("--split-sub-h", type=int, default=1080, help="split-frame: sub-frame height (1080: 4x4 sub-frames of the 8K frame; 540: 4x8)")
    ap.add_argument("--batch-frames", type=int, default=8, help="split-frame: frames per step (a batch submitted back to back)")
    ap.add_argument("--barrier", default="flag", choices=["flag", "nccl"], help="split-frame with p2p stores: completion counter in the presenter's memory, or an NCCL all-reduce")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.workload == "f_rows":   # (developer shortcut: only the SURVEY 8(f) sub-record of the default line)
        import torch
        import rsr_b200
        torch.cuda.set_device(local_rank)
        gpu = rsr_b200.GPU(local_rank)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{local_rank}")
        print(json.dumps({"f_rows": f_rows_subrecord(gpu, torch, local_rank, flush, args), "bundled_scenes": bundled_scenes_subrecord()}), flush=True)
        finish(0)

    if args.impl == "reference":
        if rank != 0:
            return 0
        wl = Workload("c2" if args.workload == "c5" else args.workload)
        r = time_reference(wl, args.steps, args.warmup, budget_s=args.ref_budget)
        line = {"impl": "reference", "metric": METRIC, "value": r["fps"], "unit": "frames/s", "n_gpus": args.gpus,
                "steps": r["frames"], "warmup": args.warmup, "ms_per_step": r["ms_per_frame"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32+i32", "data": "synthetic", "config": wl.config(args),
                "mtris_per_s": wl.scene.triangles * r["fps"] / 1e6,
                "cpu_baseline": {"value": r["fps"], "unit": "frames/s", "cores": r["threads"], "kind": "reference",
                                 "sample": f"{r['frames']} frames of the same workload after 1 s of untimed priming frames, doubleBuffer=true, worst 5% dropped"},
                "e2e": {"value": r["fps"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), flush=True)
        finish(0)

    cores_per_rank = pin_rank_to_cores(local_rank, world) if (world > 1 and not os.environ.get("RSR_BENCH_NO_PIN")) else None
    import torch
    import rsr_b200
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    gpu = rsr_b200.GPU(local_rank)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{local_rank}")

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    if args.workload == "c5":
        rec = split_frame(args, gpu, torch, dist, rank, local_rank, world, flush, barrier)
        if rank == 0:
            rec.update({"warmup": args.warmup, "higher_is_better": True, "vs_baseline": None, "dtype": "f32+i32", "data": "synthetic"})
            print(json.dumps(rec), flush=True)
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        finish(0)

    wl = Workload(args.workload)
    config = wl.config(args)
    if cores_per_rank:
        config["host_cores_per_rank"] = cores_per_rank

    # ---- parity gate: frame 0 against the reference's CPU rasteriser, in this run -------------------------------
    parity = None
    if rank == 0 and not args.no_parity:
        try:
            parity = parity_gate(wl, gpu)
        except Exception as exc:  # oracle not shipped: say so, never fake
            parity = {"diff_pixels": None, "max_lsb": None, "against": f"unavailable: {exc}"}

    # ---- device-resident legs ---------------------------------------------------------------------------------
    # (a) frame sequence (single-sub-frame workloads): `value`.  (b) one frame at a time, L2 flushed before each, an
    # event after every kernel: the per-stage times, the tile kernel's launch duration for the roofline, and the
    # cold single-frame latency (`flushed_frame`); `value` of the workloads rendered as several sub-frames.
    sampler = ClockSampler(local_rank).start() if rank == 0 else None
    seq_ms = None
    sustained = None
    if wl.single and args.resident == "retained":
        seq = SequenceRun(wl, gpu, torch, local_rank, args.ring)
        seq.frames(max(args.warmup, args.ring))   # every frame of the ring once: buffers sized, pipeline primed
        gpu.Sync()
        barrier()
        seq_ms = seq.timed(args.steps)
        barrier()
        if args.sustained_seconds > 0 and rank == 0 and world == 1:
            sustained = sequence_sustained(seq, args.sustained_seconds, local_rank)
        seq.close()
        config["cache"] = (f"ring of {args.ring} distinct frames replayed back to back: per-frame outputs + inputs "
                           f"({args.ring} x {4 * wl.frame[0] * wl.frame[1] / 1e6:.1f} MB of outputs alone) exceed the 126 MB L2; "
                           "scene-static meshes / textures are shared by the frames as in any frame sequence")
        config["resident_leg"] = "frame sequence: retained frame tables replayed (rsrcu_replay_frame), frame overlap on"
    run = ResidentRun(wl, gpu, args, torch, local_rank)
    m = measure_resident(run, flush, args.steps, args.warmup, barrier, None)
    if sampler:
        m["clocks"] = sampler.stop()
    flushed_ms = m["dev_ms"]
    t = torch.tensor([seq_ms if seq_ms is not None else flushed_ms, flushed_ms], dtype=torch.float64, device=f"cuda:{local_rank}")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms_max, flushed_ms_max = float(t[0].item()), float(t[1].item())
    ms_per_step = dev_ms_max / args.steps
    value = world * args.steps / (dev_ms_max / 1e3)
    flushed_frame = {"ms": flushed_ms_max / args.steps, "frames_per_s": world * args.steps / (flushed_ms_max / 1e3),
                     "what": "one frame at a time, 256 MiB L2 flush before each, CUDA events around the frame's kernels (cold single-frame latency)"}
    if sustained is None and args.sustained_seconds > 0 and rank == 0 and world == 1:
        sustained = sustained_leg(run, flush, args.sustained_seconds, local_rank)
    run.close()
    barrier()

    # ---- end-to-end leg ------------------------------------------------------------------------------------------
    e2e_ctx = args.e2e_contexts or (2 if world <= 2 else 1)   # (measured: 1 / 2 / 3 / 4 contexts = 4.0 / 5.5 / 5.0 / 4.9 k frames/s on a fast host, 2.6 / 4.7 / 5.0 k on a slow one; 5.5 k is the PCIe read-back limit)
    e2e = e2e_leg(wl, gpu, max(args.e2e_frames, args.steps), torch, dist, world, local_rank, barrier, e2e_ctx, args.e2e_reps) if wl.single else None
    barrier()

    # ---- sub-records -----------------------------------------------------------------------------------------------
    extra = {}
    if args.workload == "c2" and not args.no_subrecords:
        if world == 1:
            for key in ("c4_4k", "c3_4k"):
                try:
                    extra[key] = stress_subrecord(key, args, gpu, torch, local_rank, flush, barrier)
                except Exception as exc:
                    extra[key] = {"error": repr(exc)}
            try:
                extra["f_rows"] = f_rows_subrecord(gpu, torch, local_rank, flush, args)
            except Exception as exc:
                extra["f_rows"] = {"error": repr(exc)}
            if not args.no_cpu_baseline:
                try:
                    extra["bundled_scenes"] = bundled_scenes_subrecord()
                except Exception as exc:
                    extra["bundled_scenes"] = {"error": repr(exc)}
        else:
            rec = split_frame(args, gpu, torch, dist, rank, local_rank, world, flush, barrier, scene=wl.scene)
            if rank == 0:
                extra["split_frame"] = rec

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        finish(0)

    stats = m["stats"]
    roofline = roofline_of(wl, m, ms_per_step)
    cpu = None
    if not args.no_cpu_baseline and world == 1:   # (rank 0 at N = 1 only: at N > 1 the ranks are pinned to slices of the cores)
        try:
            r = time_reference_isolated(wl.key, 30, 3, 20.0)
            cpu = {"value": r["fps"], "unit": "frames/s", "cores": r["threads"], "kind": "reference",
                   "sample": f"{r['frames']} frames of the same workload on the host cores after 1 s of untimed priming frames, doubleBuffer=true, worst 5% dropped"}
        except Exception as exc:  # oracle not shipped: say so, never fake
            cpu = {"value": None, "unit": "frames/s", "cores": 0, "kind": "reference", "sample": f"unavailable: {exc}"}

    line = {"metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32+i32", "data": "synthetic", "config": config,
            "mtris_per_s": wl.scene.triangles * value / 1e6,
            "gpix_per_s": m["fragments"] * value / 1e9,
            "fragments_per_frame": m["fragments"], "bin_entries_per_frame": m["bin_entries"],
            "flushed_frame": flushed_frame, "clocks": m["clocks"], "parity": parity, "e2e": e2e,
            "gpu_launches": int(stats["kernel_launches"]) * len(wl.subframes) * args.steps,
            "roofline": roofline, "sustained": sustained, "cpu_baseline": cpu}
    line.update(extra)
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    finish(0)


if __name__ == "__main__":
    sys.exit(main())
