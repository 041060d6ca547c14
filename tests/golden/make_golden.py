#!/usr/bin/env python3
"""Generates tests/golden/*.npz by running the UNMODIFIED reference (oracle/_ref/librsr_ref.so, built from
/root/reference by oracle/build_ref.sh) in this container.  Each fixture stores the frame the reference
rendered plus the rcpps/rsqrtps tables of the CPU it ran on, so the frame can be reproduced bit for bit
anywhere: rsr_b200.GPU.set_host_luts() / oracle.restate.RestateGPU(luts=...).

    python tests/golden/make_golden.py
"""
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import refgl, restate  # noqa: E402
import rsr_b200 as R  # noqa: E402
from rsr_b200 import scenes  # noqa: E402

SIZE = (256, 144)


def cases():
    """name -> (scene factory, record kwargs); shared with the tests"""
    return {
        "wavy_bilinear": (lambda: scenes.WavyGridScene(n=24, tex_dim=64, bilinear=True), {}),
        "wavy_nearest": (lambda: scenes.WavyGridScene(n=24, tex_dim=64, bilinear=False), {}),
        "cubes_many_obj2": (lambda: scenes.CubesScene(instances=60), {}),
        "soup_clip": (lambda: scenes.SoupScene(n=300, seed=3, tex_dim=32), {}),
        "soup_blend_cull": (lambda: scenes.SoupScene(n=200, seed=7, tex_dim=32, blend=True, cull=R.GL_BACK), {}),
    }


def main():
    refgl.init(4)
    ref = refgl.RefGPU()
    rcp, rsq = restate.harvest_luts()
    # the tables must describe the instructions the reference just used
    x = np.float32(1.0) + np.arange(2048, dtype=np.float32) / np.float32(2048)
    assert np.array_equal(refgl.rcp(x).view(np.uint32), rcp)
    for name, (factory, kw) in cases().items():
        scene = factory()
        out = np.zeros((SIZE[1], SIZE[0]), np.uint32)
        scene.record(ref, SIZE, out, **kw)
        ref.Run()
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, frame=out, rcp=rcp, rsqrt=rsq, size=np.array(SIZE))
        print(f"{name}: crc32 {zlib.crc32(out.tobytes()):08x}, {np.unique(out).size} colours, {os.path.getsize(path)} bytes")
    # the reference's own known-answer test, rglv_triangle.t.cxx:205-228
    kat = ["........", ".XX.....", ".XXXX...", "..X.....", "........", "........", "........", "........"]
    got = refgl.raster_coverage([2.0, 4.0, 6.0, 2.0, 1.0, 1.0], 8, 8)
    assert ["".join("X" if c else "." for c in row) for row in got] == kat
    np.savez_compressed(os.path.join(HERE, "kat_fill_rule.npz"), points=np.array([2.0, 4.0, 6.0, 2.0, 1.0, 1.0], np.float32),
                        coverage=got)
    print("kat_fill_rule: ok")


if __name__ == "__main__":
    main()
