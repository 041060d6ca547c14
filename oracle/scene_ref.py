"""TEST INFRASTRUCTURE -- renders frames of a bundled scene through the UNMODIFIED reference (oracle/_ref/librsr_ref.so,
its own node graph and CPU rasteriser) in a process of its own:  python -m oracle.scene_ref NAME WxH t0,t1,... out.npy

Why a process: the reference's binner occasionally faults on its own bundled scenes (seen about once in eight runs of
tucker-and-dino.lua: SIGSEGV in GPUBinImpl<AmyProgram>::BinTriangles2P via DrawArrays1, rglv_gpu_impl.hxx:315-508 -- its
4-wide loop reads the lanes past a 6-vertex array's padding, and what it finds there decides which tile list it writes
to).  `frames()` retries a faulted run; a fault costs a retry, not the test session."""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def frames(name, size, times, attempts=5, threads=None):
    """[frame at t for t in times] of data/scene/NAME.lua at `size`, rendered by the reference in a subprocess"""
    with tempfile.TemporaryDirectory() as tmp:
        out = os.path.join(tmp, "frames.npy")
        cmd = [sys.executable, "-m", "oracle.scene_ref", name, f"{size[0]}x{size[1]}", ",".join(repr(float(t)) for t in times), out]
        if threads:
            cmd.append(str(int(threads)))
        last = None
        for _ in range(attempts):
            last = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
            if last.returncode == 0 and os.path.exists(out):
                return list(np.load(out))
        raise RuntimeError(f"reference render of {name} failed {attempts} times (last exit code {last.returncode}): {last.stderr[-300:]}")


def main(argv):
    from oracle import refgl
    name, (w, h) = argv[1], (int(v) for v in argv[2].split("x"))
    times = [float(t) for t in argv[3].split(",")]
    refgl.init(int(argv[5]) if len(argv) > 5 else min(8, os.cpu_count() or 1))
    sc = refgl.RefScene(name)
    out = np.stack([sc.render((w, h), t) for t in times])
    np.save(argv[4], out)
    sys.stdout.flush()
    os._exit(0)   # (no teardown: the job system's workers are still spinning)


if __name__ == "__main__":
    main(sys.argv)
