#!/usr/bin/env python3
"""Generates tests/golden/colortest_c1.npz: BASELINE.json configs[0], the reference's bundled scene
data/scene/colortest.lua (mesh data/mesh/colortest.obj, program OBJ2, camera colortest.lua:9-11) rendered at 640x360
by the UNMODIFIED reference in this container.

Everything comes out of the reference's own compiled code (oracle/_ref/librsr_ref.so): the mesh through
rglv::LoadOBJ + rglv::MakeArray(mesh, "PND") (what node/mesh.cxx:40 binds), the camera through LookAt / Perspective2 as
node/perspective.cxx:51-69 calls them, the frame through rglv::GL + rglv::GPU::Run.  The fixture stores the vertex
arrays and matrices (the GPU box has no /root/reference), the frame, and the rcpps / rsqrtps tables of the CPU that
rendered it, so the frame replays bit for bit anywhere (rsr_b200.GPU.set_host_luts).

    python tests/golden/make_bundled.py
"""
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import refgl, restate  # noqa: E402

REF = os.environ.get("RSR_REFERENCE", "/root/reference")
SIZE = (640, 360)


def main():
    (pos, nrm, kd), idx = refgl.load_obj_arrays(os.path.join(REF, "data", "mesh", "colortest.obj"), "PND")
    # data/scene/colortest.lua:9-11: Perspective{ position=Vec3(88, 80, 93), h=3.72, v=-0.35, fov=45.0 }
    view, proj = refgl.perspective_camera((88.0, 80.0, 93.0), 3.72, -0.35, 45.0, SIZE[0] / SIZE[1])
    # colortest.lua:12: color=sRGB(128,128,128); data/scene/extras.lua:79-85: Linear(s) = (s/255.0)^2.2333 in Lua doubles
    clear = np.array([(128 / 255.0) ** 2.2333] * 3, np.float64).astype(np.float32)
    path = os.path.join(HERE, "colortest_c1.npz")
    np.savez_compressed(path, pos=pos, nrm=nrm, kd=kd, idx=idx, view=view, proj=proj, clear=clear)   # (the scene reads these)

    from rsr_b200 import scenes
    refgl.init(4)
    ref = refgl.RefGPU()
    rcp, rsq = restate.harvest_luts()
    out = np.zeros((SIZE[1], SIZE[0]), np.uint32)
    scenes.ColortestScene().record(ref, SIZE, out)
    ref.Run()
    np.savez_compressed(path, pos=pos, nrm=nrm, kd=kd, idx=idx, view=view, proj=proj, clear=clear,
                        frame=out, rcp=rcp, rsqrt=rsq, size=np.array(SIZE))
    print(f"colortest_c1: {pos.shape[1]} vertices, {idx.size // 3} faces, crc32 {zlib.crc32(out.tobytes()):08x}, "
          f"{np.unique(out).size} colours, {np.count_nonzero(out != out[0, 0])} non-background pixels, {os.path.getsize(path)} bytes")


if __name__ == "__main__":
    main()
