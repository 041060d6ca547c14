#!/usr/bin/env python3
"""A bundled data/scene file through the reference's own front-end (oracle/ref_scene.cpp = perf.cxx's loop), on the CPU
reference or -- same node graph, GPU::RunImpl replaced by the C-ABI binding -- on the GPU.  One JSON line.
python tools/scene_bench.py --lib ref|dropin --scene tucker-and-dino [--size 1920x1080] [--seconds 3]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import refgl  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--lib", default="ref", choices=["ref", "dropin"])
ap.add_argument("--scene", default="tucker-and-dino")
ap.add_argument("--size", default="1920x1080")
ap.add_argument("--seconds", type=float, default=3.0)
ap.add_argument("--threads", type=int, default=0)
ap.add_argument("--static-assets", action="store_true", help="drop-in: textures and index arrays declared immutable (uploaded once): what a host sets for store-owned assets")
args = ap.parse_args()
w, h = (int(v) for v in args.size.split("x"))
threads = args.threads or os.cpu_count() or 1
if args.lib == "dropin":
    refgl.init_dropin(threads)
else:
    refgl.init(threads)
if args.lib == "dropin" and args.static_assets:
    refgl.dropin_lib().ref_dropin_set_upload_policy(3, 1, 1)   # buffers per frame, textures / indices static
sc = refgl.RefScene(args.scene, dropin=args.lib == "dropin")
sc.bench((w, h), 10)                       # priming (perf.cxx: 100 frames; the stores and caches settle in a few)
n, secs = 10, sc.bench((w, h), 10, t0=1.0)
while secs < args.seconds and n < 100000:  # then a run sized to `seconds`
    n = max(n + 1, int(n * min(10.0, 1.2 * args.seconds / max(secs, 1e-6))))
    secs = sc.bench((w, h), n, t0=1.0)
frame = sc.render((w, h), 1.0)
import zlib  # noqa: E402
print(json.dumps({"lib": args.lib, "scene": args.scene, "size": [w, h], "frames": n, "seconds": secs, "frames_per_s": n / secs,
                  "threads": threads, "static_assets": bool(args.static_assets), "crc32_frame_t1": f"{zlib.crc32(frame.tobytes()):08x}"}), flush=True)
os._exit(0)
