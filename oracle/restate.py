"""TEST INFRASTRUCTURE -- ctypes binding of oracle/restate.c (the plain-C restatement of the hot path).

`RestateGPU` offers the same `rglv::GL` method names as `oracle.refgl.RefGPU` and `rsr_b200.GPU`, so a
scene function can be rendered three ways: unmodified reference, C restatement, CUDA product.
Only tests/, smoke() and bench.py's cpu_baseline leg may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "restate.c")
LIB_PATH = os.path.join(HERE, "_build", "librestate.so")

_lib = None


class RstDraw(C.Structure):
    _fields_ = [("vm", C.c_float * 16), ("pm", C.c_float * 16), ("uniforms", C.c_float * 32),
                ("program", C.c_int), ("culling", C.c_int), ("cullFace", C.c_int),
                ("depthTest", C.c_int), ("depthFunc", C.c_int), ("depthWrite", C.c_int), ("colorWrite", C.c_int), ("blend", C.c_int),
                ("width", C.c_int), ("height", C.c_int), ("tileW", C.c_int), ("tileH", C.c_int),
                ("vpx", C.c_int), ("vpy", C.c_int), ("vpw", C.c_int), ("vph", C.c_int),
                ("buffers", C.c_void_p * 16), ("tex", C.c_void_p), ("texDim", C.c_int), ("texFilter", C.c_int),
                ("rcpLut", C.c_void_p), ("rsqrtLut", C.c_void_p)]


def build(force: bool = False) -> str:
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(SRC):
        os.makedirs(os.path.dirname(LIB_PATH), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-msse2", "-ffp-contract=off", "-fno-fast-math", "-std=c11", "-shared", "-fPIC",
                               "-o", LIB_PATH, SRC, "-lm"])
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        vp, ci = C.c_void_p, C.c_int
        L.rst_harvest_luts.argtypes = [vp, vp]
        L.rst_rcp.argtypes = [C.c_float, vp]; L.rst_rcp.restype = C.c_float
        L.rst_rsqrt.argtypes = [C.c_float, vp]; L.rst_rsqrt.restype = C.c_float
        L.rst_raster_coverage.argtypes = [ci, vp, vp, ci, ci, ci, ci, ci, ci, vp]
        L.rst_draw.argtypes = [C.POINTER(RstDraw), ci, vp, ci, vp]; L.rst_draw.restype = C.c_uint64
        L.rst_clear.argtypes = [vp, ci, ci, vp, C.c_float]
        L.rst_store_tc.argtypes = [vp, ci, ci, ci, vp, ci]
        _lib = L
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def harvest_luts():
    rcp, rsq = np.zeros(2048, np.uint32), np.zeros(2048, np.uint32)
    lib().rst_harvest_luts(_ptr(rcp), _ptr(rsq))
    return rcp, rsq


def raster_coverage(x3, y3, rect, w, h, wide=True):
    xs, ys = np.ascontiguousarray(x3, np.int32), np.ascontiguousarray(y3, np.int32)
    out = np.zeros((h, w), np.uint8)
    lib().rst_raster_coverage(int(wide), _ptr(xs), _ptr(ys), rect[0], rect[1], rect[2], rect[3], w, h, _ptr(out))
    return out


SUPPORTED_PROGRAMS = (4, 6, 8)


class RestateGPU:
    def __init__(self, luts=None):
        self.L = lib()
        self.rcp, self.rsq = luts if luts is not None else harvest_luts()
        self.rcp = np.ascontiguousarray(self.rcp, np.uint32)
        self.rsq = np.ascontiguousarray(self.rsq, np.uint32)
        self.fragments = 0

    def close(self):
        pass

    def Reset(self, size, tile_blocks=(8, 8)):
        self.size = (int(size[0]), int(size[1]))
        self.tile = (8 * int(tile_blocks[0]), 8 * int(tile_blocks[1]))
        self.cmds = []
        self.s = dict(program=0, culling=0, cullFace=2, depthTest=1, depthFunc=0, depthWrite=1, colorWrite=1, blend=0,
                      clearColor=(0.0, 0.0, 0.0), clearDepth=1.0, vm=np.eye(4, dtype=np.float32).reshape(16),
                      pm=np.eye(4, dtype=np.float32).reshape(16), uniforms=np.zeros(32, np.float32),
                      buffers=[None] * 16, tex=None, texDim=8, texFilter=0, viewport=(0, 0, 0, 0))

    def _cap(self, cap, v):
        key = {1: "culling", 3: "blend", 4: "depthTest"}.get(cap)
        if key is None:
            raise ValueError("restatement: unsupported capability")
        self.s[key] = v

    def Enable(self, cap): self._cap(cap, 1)
    def Disable(self, cap): self._cap(cap, 0)
    def DepthFunc(self, v): self.s["depthFunc"] = v
    def DepthWriteMask(self, v): self.s["depthWrite"] = int(bool(v))
    def ColorWriteMask(self, v): self.s["colorWrite"] = int(bool(v))
    def CullFace(self, v): self.s["cullFace"] = v
    def Viewport(self, x, y, w, h): self.s["viewport"] = (x, y, w, h)
    def UseProgram(self, pid): self.s["program"] = int(pid)
    def RenderbufferType(self, attachment, t): assert t == 0, "restatement: RB_COLOR_DEPTH only"
    def ClearColor(self, rgb): self.s["clearColor"] = tuple(float(x) for x in rgb[:3])
    def ClearDepth(self, d): self.s["clearDepth"] = float(d)
    def ViewMatrix(self, m): self.s["vm"] = np.ascontiguousarray(np.asarray(m, np.float32).T.reshape(16))
    def ProjectionMatrix(self, m): self.s["pm"] = np.ascontiguousarray(np.asarray(m, np.float32).T.reshape(16))
    def NormalMatrix(self, m): pass

    def UseBuffer(self, slot, arr, **_):
        if arr is not None and np.asarray(arr).ndim == 2:
            for i, row in enumerate(np.asarray(arr)):
                self.UseBuffer(slot + i, row)
            return
        bufs = list(self.s["buffers"])
        bufs[slot] = None if arr is None else np.ascontiguousarray(arr, np.float32)
        self.s["buffers"] = bufs

    def UseUniforms(self, data):
        b = np.ascontiguousarray(data).view(np.uint8).reshape(-1)
        u = np.zeros(128, np.uint8)
        u[:b.size] = b
        self.s["uniforms"] = u.view(np.float32).copy()

    def BindTexture(self, unit, texels, width, height, stride, mode, **_):
        if unit == 0:
            assert width == height == stride and (width & (width - 1)) == 0, "restatement: pow2 square textures only"
            self.s["tex"] = np.ascontiguousarray(texels, np.float32)
            self.s["texDim"] = width
            self.s["texFilter"] = mode

    def BindTexture3(self, depth, dim, **_): pass

    def Clear(self, bits):
        assert bits == 3
        self.cmds.append(("clear", dict(self.s)))

    def _draw(self, count, idx, instances):
        assert self.s["program"] in SUPPORTED_PROGRAMS, f"restatement does not cover program {self.s['program']}"
        self.cmds.append(("draw", dict(self.s), int(count), None if idx is None else np.ascontiguousarray(idx, np.uint16), int(instances)))

    def DrawElements(self, count, indices, hint=0, **_): self._draw(count, indices, 0)
    def DrawArrays(self, count): self._draw(count, None, 0)
    def DrawElementsInstanced(self, count, indices, n, **_): self._draw(count, indices, n)
    def DrawArraysInstanced(self, count, n): self._draw(count, None, n)

    def StoreColor(self, dst, gamma=True):
        assert dst.dtype == np.uint32
        self.cmds.append(("store_tc", dict(self.s), dst, bool(gamma)))

    def StoreDepth(self, dst):
        self.cmds.append(("store_depth", dict(self.s), dst))

    def Run(self, **_):
        w, h = self.size
        fb = np.zeros((h, w, 4), np.float32)
        self.fragments = 0
        for cmd in self.cmds:
            s = cmd[1]
            if cmd[0] == "clear":
                rgb = np.array(s["clearColor"], np.float32)
                self.L.rst_clear(_ptr(fb), w, h, _ptr(rgb), s["clearDepth"])
            elif cmd[0] == "draw":
                d = RstDraw()
                d.vm[:] = s["vm"]; d.pm[:] = s["pm"]; d.uniforms[:] = s["uniforms"]
                d.program = s["program"]; d.culling = s["culling"]; d.cullFace = s["cullFace"]
                d.depthTest = s["depthTest"]; d.depthFunc = s["depthFunc"]; d.depthWrite = s["depthWrite"]
                d.colorWrite = s["colorWrite"]; d.blend = s["blend"]
                d.width, d.height = w, h
                d.tileW, d.tileH = self.tile
                d.vpx, d.vpy, d.vpw, d.vph = s["viewport"]
                for i, b in enumerate(s["buffers"]):
                    d.buffers[i] = None if b is None else b.ctypes.data
                if s["tex"] is not None:
                    d.tex = s["tex"].ctypes.data
                d.texDim = s["texDim"]; d.texFilter = s["texFilter"]
                d.rcpLut = self.rcp.ctypes.data; d.rsqrtLut = self.rsq.ctypes.data
                self.fragments += int(self.L.rst_draw(C.byref(d), cmd[2], _ptr(cmd[3]), cmd[4], _ptr(fb)))
            elif cmd[0] == "store_tc":
                dst = cmd[2]
                self.L.rst_store_tc(_ptr(fb), w, h, int(cmd[3]), _ptr(dst), dst.strides[0] // 4)
            elif cmd[0] == "store_depth":
                cmd[2][...] = fb[..., 3]
        self.fb = fb
