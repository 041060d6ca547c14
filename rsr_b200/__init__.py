"""rsr_b200 -- B200 (sm_100a) rasteriser behind the rsr `rglv::GL` / `rglv::GPU` drawing API.

`GPU` mirrors the reference's host interface for the frame-rendering hot path
(src/rgl/rglv/rglv_gl.hxx:182-344 `rglv::GL`, src/rgl/rglv/rglv_gpu.hxx:152-168 `rglv::GPU`)
with the reference's own method names and argument meaning, on top of the C ABI declared in
include/rsrcu.h (librsrcu.so, hand-written CUDA for sm_100a).  There is no CPU path: if the
library or a CUDA device is missing, construction raises.
"""
from __future__ import annotations

import ctypes as C
import os
import struct

import numpy as np

from . import build as _build

# -- constants: src/rgl/rglv/rglv_gl.hxx:20-62 ---------------------------------------------------
GL_CULL_FACE, GL_SCISSOR_TEST, GL_BLEND, GL_DEPTH_TEST = 1, 2, 3, 4
GL_FRONT, GL_BACK, GL_FRONT_AND_BACK = 1, 2, 3
GL_NEAREST_MIPMAP_NEAREST, GL_LINEAR_MIPMAP_NEAREST = 0, 1
GL_COLOR_BUFFER_BIT, GL_DEPTH_BUFFER_BIT, GL_STENCIL_BUFFER_BIT = 1, 2, 4
RGL_HINT_READ4, RGL_HINT_DENSE = 1, 2
GL_LESS, GL_LEQUAL, GL_EQUAL = 0, 1, 2
GL_DEPTH_ATTACHMENT, GL_STENCIL_ATTACHMENT, GL_COLOR_ATTACHMENT0 = 0, 1, 2
RB_COLOR_DEPTH, RB_RGBF32, RB_RGBAF32, RB_F32 = 0, 1, 2, 3

# program ids: src/viewer/shaders.hxx, shaders_envmap.hxx, shaders_wireframe.hxx
PROGRAM_DEFAULT_POST, PROGRAM_EXPOSURE_POST, PROGRAM_IQ_POST = 1, 2, 3
PROGRAM_AMY, PROGRAM_DEPTH, PROGRAM_MANY, PROGRAM_OBJ1, PROGRAM_OBJ2, PROGRAM_OBJ2S = 4, 5, 6, 7, 8, 9
PROGRAM_ENVMAP, PROGRAM_WIREFRAME, PROGRAM_TEXT, PROGRAM_PATTERN, PROGRAM_ALPHATEXTURE = 10, 11, 26, 41, 65

UPLOAD_ALWAYS, UPLOAD_STATIC, UPLOAD_DEVICE, UPLOAD_FRAME = 0, 1, 2, 3

RSRCU_OK = 0
ERROR_NAMES = {1: "NO_DEVICE", 2: "CUDA", 3: "INVALID", 4: "NO_PROGRAM", 5: "UNSUPPORTED", 6: "OVERFLOW"}


class RsrError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"rsrcu error {code} ({ERROR_NAMES.get(code, '?')}): {message}")
        self.code = code


class RsrState(C.Structure):
    """POD mirror of include/rsrcu.h `RsrState` (itself a mirror of rglv::GLState)."""
    _fields_ = [
        ("clear_color", C.c_float * 4),
        ("clear_depth", C.c_float),
        ("culling_enabled", C.c_int32),
        ("cull_face", C.c_int32),
        ("scissor_enabled", C.c_int32),
        ("scissor_origin", C.c_int32 * 2),
        ("scissor_size", C.c_int32 * 2),
        ("viewport_origin", C.c_int32 * 2),
        ("viewport_size", C.c_int32 * 2),
        ("blending_enabled", C.c_int32),
        ("color_write_mask", C.c_int32),
        ("depth_write_mask", C.c_int32),
        ("depth_test_enabled", C.c_int32),
        ("depth_func", C.c_int32),
        ("program_id", C.c_int32),
        ("color0_attachment_type", C.c_int32),
        ("depth_attachment_type", C.c_int32),
        ("view_matrix", C.c_float * 16),
        ("projection_matrix", C.c_float * 16),
        ("normal_matrix", C.c_float * 16),
        ("uniforms_valid", C.c_uint32),
        ("uniforms", C.c_float * 32),
    ]


class RsrStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("triangles_submitted", "triangles_binned", "triangles_clipped",
                                           "bin_entries", "fragments_shaded", "kernel_launches", "h2d_bytes", "d2h_bytes",
                                           "list_chunks_run_merge", "list_chunks_key_range", "host_record_ns", "host_submit_ns", "frames_retried", "input_bytes", "draws_culled")]


_lib = None


def library_path() -> str:
    return _build.LIB


def load_library():
    """Loads librsrcu.so (building it first when sources are newer and nvcc is present)."""
    global _lib
    if _lib is not None:
        return _lib
    if _build.needs_build():
        try:
            _build.build()
        except Exception as exc:  # no nvcc on this box: a prebuilt .so must be there
            if not os.path.exists(_build.LIB):
                raise RuntimeError(f"librsrcu.so is missing and cannot be built: {exc}") from exc
    L = C.CDLL(os.environ.get("RSRCU_LIB") or _build.LIB)   # RSRCU_LIB: tuning variants (build.build_variant)
    vp, ci, sz = C.c_void_p, C.c_int, C.c_size_t
    sig = {
        "rsrcu_create": [ci, C.POINTER(vp)],
        "rsrcu_destroy": [vp],
        "rsrcu_set_host_luts": [vp, vp, vp],
        "rsrcu_get_host_luts": [vp, vp, vp],
        "rsrcu_release_static": [vp],
        "rsrcu_begin_frame": [vp, ci, ci, ci, ci],
        "rsrcu_set_state": [vp, C.POINTER(RsrState)],
        "rsrcu_bind_buffer": [vp, ci, vp, sz, ci],
        "rsrcu_bind_texture": [vp, ci, vp, ci, ci, ci, ci, ci, ci],
        "rsrcu_bind_depth_texture": [vp, vp, ci, ci],
        "rsrcu_clear": [vp, ci],
        "rsrcu_draw_elements": [vp, ci, vp, ci, ci, ci],
        "rsrcu_draw_arrays": [vp, ci, ci],
        "rsrcu_store_color_tc": [vp, ci, vp, ci, ci, ci],
        "rsrcu_store_color_tc_device": [vp, ci, vp, ci, ci, ci],
        "rsrcu_store_color_fp": [vp, vp, ci, ci, ci, ci],
        "rsrcu_store_color_quads": [vp, vp, ci, ci, ci],
        "rsrcu_enable_peer_access": [vp, ci],
        "rsrcu_set_overlap": [vp, ci],
        "rsrcu_signal_counter": [vp, vp],
        "rsrcu_wait_counters": [vp, vp, ci, C.c_uint64],
        "rsrcu_retain_frame": [vp, C.POINTER(C.c_void_p)],
        "rsrcu_replay_frame": [vp, vp],
        "rsrcu_release_frame": [vp, vp],
        "rsrcu_store_depth": [vp, vp],
        "rsrcu_end_frame": [vp],
        "rsrcu_sync": [vp],
        "rsrcu_sync_frame": [vp, ci],
        "rsrcu_run_stream": [vp, vp, sz],
        "rsrcu_device_truecolor": [vp, C.POINTER(vp), C.POINTER(ci)],
        "rsrcu_stream": [vp, C.POINTER(vp)],
        "rsrcu_join": [vp],
        "rsrcu_get_stats": [vp, C.POINTER(RsrStats)],
        "rsrcu_set_profiling": [vp, ci],
        "rsrcu_get_stage_ms": [vp, vp],
        "rsrcu_canvas_alloc": [vp, sz, C.POINTER(vp)],
        "rsrcu_canvas_free": [vp, vp],
        "rsrcu_canvas_read": [vp, vp, vp, sz],
        "rsrcu_canvas_write": [vp, vp, vp, sz],
        "rsrcu_store_color_fp_device": [vp, vp, ci, ci, ci, ci],
        "rsrcu_store_color_quads_device": [vp, vp, ci, ci, ci],
        "rsrcu_store_depth_device": [vp, vp],
        "rsrcu_wait_for": [vp, vp],
        "rsrcu_kawase_blur": [vp, vp, ci, vp, ci, ci, ci, ci],
        "rsrcu_glow": [vp, vp, ci, vp, ci, ci, vp, ci, ci, ci, ci],
        "rsrcu_make_mipmap": [vp, vp, ci],
        "rsrcu_set_pin_in_place": [vp, ci],
        "rsrcu_march_surface": [vp, C.c_float, ci, ci, C.c_float, vp, vp, ci, C.POINTER(ci), C.POINTER(ci)],
        "rsrcu_draw_spans": [vp, vp, ci, ci, ci, ci, ci, C.c_float, vp, ci],
        "rsrcu_frame_spans": [vp, vp, ci, C.POINTER(ci)],
        "rsrcu_present": [vp, vp, ci, vp, ci, ci, ci],
    }
    for name, args in sig.items():
        fn = getattr(L, name)
        fn.argtypes = args
        fn.restype = ci
    L.rsrcu_last_error.restype = C.c_char_p
    L.rsrcu_last_error.argtypes = []
    _lib = L
    return L


EXPORTED_SYMBOLS = (
    "rsrcu_create", "rsrcu_destroy", "rsrcu_last_error", "rsrcu_set_host_luts", "rsrcu_get_host_luts",
    "rsrcu_release_static", "rsrcu_begin_frame", "rsrcu_set_state", "rsrcu_bind_buffer", "rsrcu_bind_texture",
    "rsrcu_bind_depth_texture", "rsrcu_clear", "rsrcu_draw_elements", "rsrcu_draw_arrays",
    "rsrcu_store_color_tc", "rsrcu_store_color_tc_device", "rsrcu_store_color_fp", "rsrcu_store_color_quads", "rsrcu_enable_peer_access", "rsrcu_set_overlap", "rsrcu_signal_counter", "rsrcu_wait_counters", "rsrcu_retain_frame", "rsrcu_replay_frame", "rsrcu_release_frame", "rsrcu_store_depth", "rsrcu_end_frame", "rsrcu_sync",
    "rsrcu_sync_frame", "rsrcu_run_stream",
    "rsrcu_device_truecolor", "rsrcu_stream", "rsrcu_join", "rsrcu_get_stats", "rsrcu_set_profiling", "rsrcu_get_stage_ms",
    "rsrcu_canvas_alloc", "rsrcu_canvas_free", "rsrcu_canvas_read", "rsrcu_canvas_write", "rsrcu_store_color_fp_device",
    "rsrcu_store_color_quads_device", "rsrcu_store_depth_device", "rsrcu_wait_for", "rsrcu_kawase_blur", "rsrcu_glow", "rsrcu_make_mipmap", "rsrcu_set_pin_in_place",
    "rsrcu_march_surface", "rsrcu_draw_spans", "rsrcu_frame_spans", "rsrcu_present",
)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


(OP_BEGIN_FRAME, OP_STATE, OP_BIND_BUFFER, OP_BIND_TEXTURE, OP_BIND_DEPTH, OP_CLEAR, OP_DRAW_ELEMENTS,
 OP_DRAW_ARRAYS, OP_STORE_TC, OP_STORE_FP, OP_STORE_DEPTH, OP_END_FRAME, OP_STORE_TC_DEV, OP_STORE_QUADS,
 OP_STORE_FP_DEV, OP_STORE_QUADS_DEV, OP_STORE_DEPTH_DEV) = range(1, 18)


class RsrSpan(C.Structure):
    """include/rsrcu.h `RsrSpan`: a jobsys::JobStat (start / end in seconds, raw bits for the colour) and its lane"""
    _fields_ = [("start", C.c_double), ("end", C.c_double), ("raw", C.c_uint32), ("lane", C.c_int32)]


class RsrMarchBlock(C.Structure):
    _fields_ = [("first_vertex", C.c_int32), ("vertex_count", C.c_int32)]


class DeviceCanvas:
    """Device memory standing in for one of the reference's canvases (rglr_canvas.hxx): kind 'fp' =
    FloatingPointCanvas (H, W, 4) float32, 'quads' = QFloat4Canvas (H/2, W/2, 4, 4) float32, 'depth' = (H, W) float32,
    'tc' = TrueColorCanvas (H, W) uint32.  Created by GPU.Canvas(); ptr is the device address."""

    def __init__(self, gpu, kind, width, height):
        self.gpu, self.kind, self.width, self.height = gpu, kind, int(width), int(height)
        # 'tex': a power-of-two square texture with room for its mip chain (2 x height rows), base level on top
        self.shape, self.dtype = {"fp": ((height, width, 4), np.float32), "tex": ((2 * height, width, 4), np.float32), "quads": ((height // 2, width // 2, 4, 4), np.float32),
                                  "depth": ((height, width), np.float32), "tc": ((height, width), np.uint32)}[kind]
        self.nbytes = int(np.prod(self.shape)) * 4
        p = C.c_void_p()
        gpu._check(gpu.L.rsrcu_canvas_alloc(gpu.h, self.nbytes, C.byref(p)))
        self.ptr = p.value

    def read(self) -> np.ndarray:
        out = np.empty(self.shape, self.dtype)
        self.gpu._check(self.gpu.L.rsrcu_canvas_read(self.gpu.h, C.c_void_p(self.ptr), _ptr(out), self.nbytes))
        return out

    def write(self, arr):
        a = np.ascontiguousarray(arr, dtype=self.dtype)
        assert a.shape == self.shape
        self.gpu._check(self.gpu.L.rsrcu_canvas_write(self.gpu.h, C.c_void_p(self.ptr), _ptr(a), self.nbytes))

    def free(self):
        if self.ptr:
            self.gpu._check(self.gpu.L.rsrcu_canvas_free(self.gpu.h, C.c_void_p(self.ptr)))
            self.ptr = None


def _addr(a) -> int:
    return 0 if a is None else a.ctypes.data


class RecordedFrame:
    """One frame as a packed command stream (include/rsrcu.h "packed command stream") plus the
    numpy arrays its pointers refer to.  Replayable with GPU.Submit()."""

    def __init__(self, data: bytes, keep: list, size):
        self.data = data
        self.buf = (C.c_char * len(data)).from_buffer_copy(data)
        self.keep = keep
        self.size = size


class GPU:
    """`rglv::GPU` + its recording `rglv::GL` context, rendered by the CUDA library.

    Usage follows the reference (node/gpu.cxx:107-161, node/truecolor.cxx:90-108):
    Reset(size, tileBlocks); state + draw calls; StoreColor(...); Run().

    Like the reference, GL calls are *recorded* (into the packed stream of include/rsrcu.h) and the
    whole frame is handed to the renderer by Run() -> rsrcu_run_stream.  With direct=True every GL
    call goes through its own C-ABI entry point instead (same result; used by the tests to cover
    both routes).
    """

    def __init__(self, device: int = 0, direct: bool = False):
        self.L = load_library()
        h = C.c_void_p()
        self._check(self.L.rsrcu_create(int(device), C.byref(h)))
        self.h = h
        self.device = device
        self.direct = direct
        self._keep = []
        self._static = {}   # address -> array: pins RSRCU_UPLOAD_STATIC memory so the address cannot be recycled
        self._rec = bytearray()
        self._state = RsrState()
        self._reset_state()
        self._dirty = True
        self.size = (0, 0)

    def _hold(self, arr, upload):
        if upload == UPLOAD_STATIC:
            self._static[arr.ctypes.data] = arr
        else:
            self._keep.append(arr)

    def release_static(self):
        """forget every cached static upload (and let the pinned host arrays go)"""
        self._check(self.L.rsrcu_release_static(self.h))
        self._static.clear()

    # -- plumbing -------------------------------------------------------------------------------
    def _check(self, rc: int):
        if rc != RSRCU_OK:
            raise RsrError(rc, self.L.rsrcu_last_error().decode("utf-8", "replace"))

    def close(self):
        if getattr(self, "h", None):
            self.L.rsrcu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _reset_state(self):
        """GLState::reset (rglv_gl.hxx:152-179) + the attachment defaults every caller sets"""
        s = self._state
        C.memset(C.byref(s), 0, C.sizeof(s))
        s.clear_color[:] = [0.0, 0.0, 0.0, 1.0]
        s.clear_depth = 1.0
        s.cull_face = GL_BACK
        s.color_write_mask = 1
        s.depth_write_mask = 1
        s.depth_test_enabled = 1
        s.depth_func = GL_LESS
        s.color0_attachment_type = RB_COLOR_DEPTH
        s.depth_attachment_type = RB_COLOR_DEPTH
        ident = np.eye(4, dtype=np.float32).reshape(16)
        s.view_matrix[:] = ident
        s.projection_matrix[:] = ident
        s.normal_matrix[:] = ident

    def _emit(self, op: int, payload: bytes):
        pad = (-len(payload)) % 8
        self._rec += struct.pack("<II", op, 8 + len(payload) + pad) + payload + b"\0" * pad

    def _flush_state(self):
        if self._dirty:
            if self.direct:
                self._check(self.L.rsrcu_set_state(self.h, C.byref(self._state)))
            else:
                self._emit(OP_STATE, bytes(self._state))
            self._dirty = False

    # -- rglv::GPU ------------------------------------------------------------------------------
    def Reset(self, size, tile_blocks=(8, 8)):
        self._keep = []
        self._rec = bytearray()
        self.size = (int(size[0]), int(size[1]))
        if self.direct:
            self._check(self.L.rsrcu_begin_frame(self.h, self.size[0], self.size[1], int(tile_blocks[0]), int(tile_blocks[1])))
        else:
            self._emit(OP_BEGIN_FRAME, struct.pack("<iiii", self.size[0], self.size[1], int(tile_blocks[0]), int(tile_blocks[1])))
        self._reset_state()
        self._dirty = True

    def Finish(self) -> RecordedFrame:
        """closes the recording and returns it (stream mode only)"""
        assert not self.direct
        self._emit(OP_END_FRAME, b"")
        rec = RecordedFrame(bytes(self._rec), self._keep, self.size)
        self._rec = bytearray()
        return rec

    def Submit(self, rec: RecordedFrame, sync: bool = True):
        """GPU::Run on a recorded frame: one C-ABI call (rsrcu_run_stream)"""
        self._check(self.L.rsrcu_run_stream(self.h, rec.buf, len(rec.data)))
        if sync:
            # (a frame whose device-side buffers overflowed is grown for and launched again inside rsrcu_sync)
            self.Sync()

    def Run(self, manage_workers: bool = True, sync: bool = True):
        """GPU::Run: end of recording -> kernels; `sync` waits and fills the store destinations"""
        if self.direct:
            self._check(self.L.rsrcu_end_frame(self.h))
            if sync:
                self.Sync()
        else:
            self.Submit(self.Finish(), sync)

    def Sync(self):
        self._check(self.L.rsrcu_sync(self.h))

    def SyncFrame(self, lag: int = 1):
        """wait until the frame `lag` submissions ago has landed in its store destinations"""
        self._check(self.L.rsrcu_sync_frame(self.h, lag))

    # -- rglv::GL -------------------------------------------------------------------------------
    def _cap(self, cap, value):
        s = self._state
        if cap == GL_CULL_FACE: s.culling_enabled = value
        elif cap == GL_SCISSOR_TEST: s.scissor_enabled = value
        elif cap == GL_BLEND: s.blending_enabled = value
        elif cap == GL_DEPTH_TEST: s.depth_test_enabled = value
        else: raise ValueError("unknown glEnable value")
        self._dirty = True

    def Enable(self, cap): self._cap(cap, 1)
    def Disable(self, cap): self._cap(cap, 0)

    def DepthFunc(self, v): self._state.depth_func = v; self._dirty = True
    def DepthWriteMask(self, v): self._state.depth_write_mask = int(bool(v)); self._dirty = True
    def ColorWriteMask(self, v): self._state.color_write_mask = int(bool(v)); self._dirty = True
    def CullFace(self, v): self._state.cull_face = v; self._dirty = True

    def Scissor(self, x, y, w, h):
        self._state.scissor_origin[:] = [x, y]; self._state.scissor_size[:] = [w, h]; self._dirty = True

    def Viewport(self, x, y, w, h):
        self._state.viewport_origin[:] = [x, y]; self._state.viewport_size[:] = [w, h]; self._dirty = True

    def UseProgram(self, pid): self._state.program_id = int(pid); self._dirty = True

    def RenderbufferType(self, attachment, t):
        if attachment == GL_DEPTH_ATTACHMENT: self._state.depth_attachment_type = t
        elif attachment == GL_COLOR_ATTACHMENT0: self._state.color0_attachment_type = t
        self._dirty = True

    def ClearColor(self, rgb):
        self._state.clear_color[:] = [float(rgb[0]), float(rgb[1]), float(rgb[2]), 1.0]; self._dirty = True

    def ClearDepth(self, d): self._state.clear_depth = float(d); self._dirty = True

    @staticmethod
    def _mat(m):
        """4x4 row-major numpy matrix (math convention) -> the reference's column-major ff[16]"""
        return np.ascontiguousarray(np.asarray(m, dtype=np.float32).T.reshape(16))

    def ViewMatrix(self, m): self._state.view_matrix[:] = self._mat(m); self._dirty = True
    def ProjectionMatrix(self, m): self._state.projection_matrix[:] = self._mat(m); self._dirty = True
    def NormalMatrix(self, m): self._state.normal_matrix[:] = self._mat(m); self._dirty = True

    def UseBuffer(self, slot, arr, upload=UPLOAD_ALWAYS):
        if arr is not None:
            a = np.asarray(arr)
            if a.ndim == 2:
                for i in range(a.shape[0]):
                    self.UseBuffer(slot + i, a[i], upload)
                return
            arr = np.ascontiguousarray(a, dtype=np.float32)
            self._hold(arr, upload)
        n = 0 if arr is None else arr.size
        if self.direct:
            self._check(self.L.rsrcu_bind_buffer(self.h, slot, _ptr(arr), n, upload))
        else:
            self._emit(OP_BIND_BUFFER, struct.pack("<iiQQ", slot, upload, _addr(arr), n))

    def UseUniforms(self, data):
        b = np.ascontiguousarray(data).view(np.uint8).reshape(-1)
        assert b.size <= 128
        buf = np.zeros(128, np.uint8)
        buf[:b.size] = b
        self._state.uniforms[:] = buf.view(np.float32)
        self._state.uniforms_valid = 1
        self._dirty = True

    def BindTexture(self, unit, texels, width, height, stride, mode, upload=UPLOAD_ALWAYS):
        t = np.ascontiguousarray(texels, dtype=np.float32)
        rows = t.size // (4 * stride)
        self._hold(t, upload)
        if self.direct:
            self._check(self.L.rsrcu_bind_texture(self.h, unit, _ptr(t), width, height, stride, mode, rows, upload))
        else:
            self._emit(OP_BIND_TEXTURE, struct.pack("<iiiiiiiiQ", unit, width, height, stride, mode, rows, upload, 0, _addr(t)))

    def BindTexture3(self, depth, dim, upload=UPLOAD_ALWAYS):
        t = np.ascontiguousarray(depth, dtype=np.float32)
        self._hold(t, upload)
        if self.direct:
            self._check(self.L.rsrcu_bind_depth_texture(self.h, _ptr(t), dim, upload))
        else:
            self._emit(OP_BIND_DEPTH, struct.pack("<iiQ", dim, upload, _addr(t)))

    def Clear(self, bits):
        self._flush_state()
        if self.direct:
            self._check(self.L.rsrcu_clear(self.h, bits))
        else:
            self._emit(OP_CLEAR, struct.pack("<ii", bits, 0))

    def _draw_elements(self, count, indices, hint, instances, upload):
        idx = np.ascontiguousarray(indices, dtype=np.uint16)
        self._hold(idx, upload)
        self._flush_state()
        if self.direct:
            self._check(self.L.rsrcu_draw_elements(self.h, int(count), _ptr(idx), int(hint), int(instances), upload))
        else:
            self._emit(OP_DRAW_ELEMENTS, struct.pack("<iiiiQ", int(count), int(hint), int(instances), upload, _addr(idx)))

    def _draw_arrays(self, count, instances):
        self._flush_state()
        if self.direct:
            self._check(self.L.rsrcu_draw_arrays(self.h, int(count), int(instances)))
        else:
            self._emit(OP_DRAW_ARRAYS, struct.pack("<ii", int(count), int(instances)))

    def DrawElements(self, count, indices, hint=0, upload=UPLOAD_ALWAYS): self._draw_elements(count, indices, hint, 0, upload)
    def DrawArrays(self, count): self._draw_arrays(count, 0)
    def DrawElementsInstanced(self, count, indices, instance_cnt, upload=UPLOAD_ALWAYS): self._draw_elements(count, indices, 0, instance_cnt, upload)
    def DrawArraysInstanced(self, count, instance_cnt): self._draw_arrays(count, instance_cnt)

    def StoreColor(self, dst, gamma: bool = True):
        """(H, W) uint32 -> CMD_STORE_COLOR_FULL_LINEAR_TC; (H, W, 4) float32 -> ..._LINEAR_FP;
        None -> true-colour resolve kept on the device (see device_truecolor)"""
        self._flush_state()
        if dst is None or dst.dtype == np.uint32:
            if dst is None:
                w, h = self.size
                stride = w
            else:
                h, w = dst.shape
                stride = dst.strides[0] // 4
                self._keep.append(dst)
            if self.direct:
                self._check(self.L.rsrcu_store_color_tc(self.h, int(bool(gamma)), _ptr(dst), w, h, stride))
            else:
                self._emit(OP_STORE_TC, struct.pack("<iiiiQ", int(bool(gamma)), w, h, stride, _addr(dst)))
        else:
            assert dst.dtype == np.float32 and dst.ndim == 3 and dst.shape[2] == 4
            h, w, _ = dst.shape
            self._keep.append(dst)
            if self.direct:
                self._check(self.L.rsrcu_store_color_fp(self.h, _ptr(dst), w, h, dst.strides[0] // 16, 0))
            else:
                self._emit(OP_STORE_FP, struct.pack("<iiiiQ", 0, w, h, dst.strides[0] // 16, _addr(dst)))

    def StoreColorHalf(self, dst):
        """(H/2, W/2, 4) float32 -> CMD_STORE_COLOR_HALF_LINEAR_FP (GL::StoreColor(dst, downsample=true))"""
        assert dst.dtype == np.float32 and dst.ndim == 3 and dst.shape[2] == 4
        self._flush_state()
        h, w, _ = dst.shape
        self._keep.append(dst)
        if self.direct:
            self._check(self.L.rsrcu_store_color_fp(self.h, _ptr(dst), w, h, dst.strides[0] // 16, 1))
        else:
            self._emit(OP_STORE_FP, struct.pack("<iiiiQ", 1, w, h, dst.strides[0] // 16, _addr(dst)))

    def StoreColorQuads(self, dst):
        """(H/2, W/2, 4, 4) float32 = [quad row][quad][r,g,b,a][lane] -> CMD_STORE_COLOR_FULL_QUADS_FP
        (GL::StoreColor(QFloat4Canvas*))"""
        assert dst.dtype == np.float32 and dst.ndim == 4 and dst.shape[2:] == (4, 4)
        self._flush_state()
        hq, wq = dst.shape[:2]
        self._keep.append(dst)
        if self.direct:
            self._check(self.L.rsrcu_store_color_quads(self.h, _ptr(dst), wq * 2, hq * 2, dst.strides[0] // 64))
        else:
            self._emit(OP_STORE_QUADS, struct.pack("<iiiiQ", 0, wq * 2, hq * 2, dst.strides[0] // 64, _addr(dst)))

    def StoreColorDevice(self, device_ptr: int, stride_px: int, gamma: bool = True):
        """true-colour resolve straight into caller-owned DEVICE memory (e.g. tensor.data_ptr())"""
        self._flush_state()
        w, h = self.size
        if self.direct:
            self._check(self.L.rsrcu_store_color_tc_device(self.h, int(bool(gamma)), C.c_void_p(device_ptr), w, h, stride_px))
        else:
            self._emit(OP_STORE_TC_DEV, struct.pack("<iiiiQ", int(bool(gamma)), w, h, stride_px, int(device_ptr)))

    def Retain(self) -> int:
        """snapshot of the frame submitted last (tables and per-frame data stay on the device): handle for Replay"""
        h = C.c_void_p()
        self._check(self.L.rsrcu_retain_frame(self.h, C.byref(h)))
        return h.value

    def Replay(self, frame: int, sync: bool = False):
        """launch a retained frame's kernels: no stream decode, no table build, no upload"""
        self._check(self.L.rsrcu_replay_frame(self.h, C.c_void_p(frame)))
        if sync:
            self.Sync()

    def Release(self, frame: int):
        self._check(self.L.rsrcu_release_frame(self.h, C.c_void_p(frame)))

    def set_overlap(self, enabled: bool):
        """front end of frame N+1 (second stream, own intermediate buffers) under the tile kernel of frame N"""
        self._check(self.L.rsrcu_set_overlap(self.h, int(bool(enabled))))

    def EnablePeerAccess(self, peer_device: int):
        """lets StoreColorDevice target memory of another GPU of the node (split-frame presentation over NVLink)"""
        self._check(self.L.rsrcu_enable_peer_access(self.h, int(peer_device)))

    def SignalCounter(self, device_ptr):
        """enqueue on this context's stream: add 1 to the 64-bit counter at `device_ptr` (possibly on another GPU) once
        every frame submitted so far has completed and its stores are visible system-wide"""
        self._check(self.L.rsrcu_signal_counter(self.h, C.c_void_p(device_ptr)))

    def WaitCounters(self, device_ptr, count, value):
        """enqueue on this context's stream a wait until each of the `count` counters at `device_ptr` has reached `value`"""
        self._check(self.L.rsrcu_wait_counters(self.h, C.c_void_p(device_ptr), int(count), int(value)))

    def StoreDepth(self, dst):
        assert dst.dtype == np.float32 and dst.flags.c_contiguous
        self._flush_state()
        self._keep.append(dst)
        if self.direct:
            self._check(self.L.rsrcu_store_depth(self.h, _ptr(dst)))
        else:
            self._emit(OP_STORE_DEPTH, struct.pack("<Q", _addr(dst)))

    # -- device canvases, post filters, geometry producers, presentation ---------------
    def Canvas(self, kind, width, height) -> "DeviceCanvas":
        return DeviceCanvas(self, kind, width, height)

    def UseBufferDevice(self, slot, device_ptr: int, n_floats: int):
        """GL::UseBuffer with an array that already lives on the device (marching-cubes output): nothing is uploaded"""
        if self.direct:
            self._check(self.L.rsrcu_bind_buffer(self.h, slot, C.c_void_p(device_ptr), int(n_floats), UPLOAD_DEVICE))
        else:
            self._emit(OP_BIND_BUFFER, struct.pack("<iiQQ", slot, UPLOAD_DEVICE, int(device_ptr), int(n_floats)))

    def BindTextureDevice(self, unit, canvas: "DeviceCanvas", mode):
        """GL::BindTexture on a frame another pass stored into a device canvas ('fp'): render to texture without a
        PCIe round trip.  Like the reference's `$rendertotexture`, the canvas is sampled without a mip chain unless it
        is a power-of-two square whose rows hold one (rows = 2 x height)."""
        assert canvas.kind in ("fp", "tex")
        w, h = canvas.width, canvas.height
        rows = 2 * h if canvas.kind == "tex" else h
        if self.direct:
            self._check(self.L.rsrcu_bind_texture(self.h, unit, C.c_void_p(canvas.ptr), w, h, w, mode, rows, UPLOAD_DEVICE))
        else:
            self._emit(OP_BIND_TEXTURE, struct.pack("<iiiiiiiiQ", unit, w, h, w, mode, rows, UPLOAD_DEVICE, 0, int(canvas.ptr)))

    def BindTexture3Device(self, canvas: "DeviceCanvas"):
        """GL::BindTexture3 on a depth canvas rendered on the device (the shadow map of a `$layer`, gllayer.cxx:154-181)"""
        assert canvas.kind == "depth" and canvas.width == canvas.height
        if self.direct:
            self._check(self.L.rsrcu_bind_depth_texture(self.h, C.c_void_p(canvas.ptr), canvas.width, UPLOAD_DEVICE))
        else:
            self._emit(OP_BIND_DEPTH, struct.pack("<iiQ", canvas.width, UPLOAD_DEVICE, int(canvas.ptr)))

    def StoreToCanvas(self, canvas: "DeviceCanvas", half: bool = False):
        """GL::StoreColor / StoreDepth into a device canvas: 'fp' (full size, or half size with half=True), 'quads', 'depth'"""
        self._flush_state()
        w, h = canvas.width, canvas.height
        if canvas.kind in ("fp", "tex"):   # ('tex': the frame is the base level of a mip-mapped texture)
            if self.direct:
                self._check(self.L.rsrcu_store_color_fp_device(self.h, C.c_void_p(canvas.ptr), w, h, w, int(half)))
            else:
                self._emit(OP_STORE_FP_DEV, struct.pack("<iiiiQ", int(half), w, h, w, int(canvas.ptr)))
        elif canvas.kind == "quads":
            if self.direct:
                self._check(self.L.rsrcu_store_color_quads_device(self.h, C.c_void_p(canvas.ptr), w, h, w // 2))
            else:
                self._emit(OP_STORE_QUADS_DEV, struct.pack("<iiiiQ", 0, w, h, w // 2, int(canvas.ptr)))
        elif canvas.kind == "depth":
            if self.direct:
                self._check(self.L.rsrcu_store_depth_device(self.h, C.c_void_p(canvas.ptr)))
            else:
                self._emit(OP_STORE_DEPTH_DEV, struct.pack("<Q", int(canvas.ptr)))
        else:
            raise ValueError("true-colour canvases are stored with StoreColorDevice")

    def set_pin_in_place(self, enabled: bool):
        """UPLOAD_FRAME arrays of 1 MiB or more are page-locked where they lie and copied by the copy engine (opt-in: the
        arrays must outlive their last use by this context; disable before freeing them)"""
        self._check(self.L.rsrcu_set_pin_in_place(self.h, int(bool(enabled))))

    def MakeMipmap(self, canvas: "DeviceCanvas"):
        """rglr::Texture::maybe_make_mipmap (rglr_texture.cxx:33-81) on a 'tex' canvas whose base level has been stored"""
        assert canvas.kind == "tex" and canvas.width == canvas.height
        self._check(self.L.rsrcu_make_mipmap(self.h, C.c_void_p(canvas.ptr), canvas.width))

    def WaitFor(self, producer: "GPU"):
        """everything submitted to `producer` so far happens before anything submitted to this context from now on"""
        self._check(self.L.rsrcu_wait_for(self.h, producer.h))

    def KawaseBlur(self, src: "DeviceCanvas", dst: "DeviceCanvas", dist: int):
        """rglr::KawaseBlurFilter (rglr_kawase.cxx:22-81) between two 'fp' canvases of one size"""
        assert src.kind == "fp" and dst.kind == "fp" and src.shape == dst.shape
        self._check(self.L.rsrcu_kawase_blur(self.h, C.c_void_p(src.ptr), src.width, C.c_void_p(dst.ptr), dst.width, src.width, src.height, int(dist)))

    def Kawase(self, src: "DeviceCanvas", intensity: int, scratch=None) -> "DeviceCanvas":
        """the `$kawase` node (node/kawase.cxx:83-129): `intensity` passes with dist 0 .. intensity - 1, ping-ponging
        two canvases; returns the canvas holding the result (src itself when intensity == 0)"""
        if intensity == 0:
            return src
        a, b = scratch if scratch else (self.Canvas("fp", src.width, src.height), self.Canvas("fp", src.width, src.height) if intensity > 1 else None)
        s, d = src, a
        first = True
        for dist in range(intensity):
            self.KawaseBlur(s, d, dist)
            if first:
                first = False
                s = b
            s, d = d, s
        return s

    def Glow(self, image: "DeviceCanvas", blur: "DeviceCanvas", dst, gamma: bool = True):
        """the `$glow` node (node/glow.cxx:146-160): (image + blur * 0.7) * 0.5 -> true colour.  dst: (H, W) uint32
        host array (written on return) or a 'tc' DeviceCanvas"""
        assert image.kind == "quads" and blur.kind == "fp"
        w, h = image.width, image.height
        if isinstance(dst, DeviceCanvas):
            self._check(self.L.rsrcu_glow(self.h, C.c_void_p(image.ptr), w // 2, C.c_void_p(blur.ptr), blur.width, int(bool(gamma)),
                                          C.c_void_p(dst.ptr), 1, w, h, dst.width))
        else:
            assert dst.dtype == np.uint32 and dst.shape == (h, w)
            self._check(self.L.rsrcu_glow(self.h, C.c_void_p(image.ptr), w // 2, C.c_void_p(blur.ptr), blur.width, int(bool(gamma)),
                                          _ptr(dst), 0, w, h, dst.strides[0] // 4))

    def MarchSurface(self, t: float, precision: int = 32, fork_depth: int = 2, rng: float = 5.0):
        """the `$mc` node's geometry (node/mc.cxx:171-300) produced on the device.  Returns (soa, blocks, total): six
        device addresses (x, y, z, nx, ny, nz arrays of `total` floats) and [(first_vertex, vertex_count)] per
        non-empty block, in the reference's draw order"""
        soa = (C.c_void_p * 6)()
        blocks = (RsrMarchBlock * 4096)()
        nb, total = C.c_int(0), C.c_int(0)
        self._check(self.L.rsrcu_march_surface(self.h, float(t), int(precision), int(fork_depth), float(rng), soa, blocks, 4096, C.byref(nb), C.byref(total)))
        return [int(p or 0) for p in soa], [(blocks[i].first_vertex, blocks[i].vertex_count) for i in range(nb.value)], total.value

    def read_device(self, device_ptr: int, count: int, dtype=np.float32) -> np.ndarray:
        out = np.empty(count, dtype)
        self._check(self.L.rsrcu_canvas_read(self.h, C.c_void_p(device_ptr), _ptr(out), out.nbytes))
        return out

    def DrawSpans(self, target, left, top, xscale, spans):
        """render_jobsys (viewer/jobsys_vis.cxx:26-90) into a 'tc' canvas: spans = [(start, end, raw, lane)]"""
        arr = (RsrSpan * max(1, len(spans)))()
        for i, s in enumerate(spans):
            arr[i] = RsrSpan(float(s[0]), float(s[1]), int(s[2]) & 0xffffffff, int(s[3]))
        self._check(self.L.rsrcu_draw_spans(self.h, C.c_void_p(target.ptr), target.width, target.width, target.height, int(left), int(top), float(xscale), arr, len(spans)))

    def FrameSpans(self):
        """the pipeline stages of the last profiled frame (set_profiling(2)) as spans, one lane per stage"""
        arr = (RsrSpan * 8)()
        n = C.c_int(0)
        self._check(self.L.rsrcu_frame_spans(self.h, arr, 8, C.byref(n)))
        return [(arr[i].start, arr[i].end, arr[i].raw, arr[i].lane) for i in range(n.value)]

    def Present(self, src: "DeviceCanvas", surface: "DeviceCanvas"):
        """device-to-device blit of a true-colour canvas into the presentation surface"""
        self._check(self.L.rsrcu_present(self.h, C.c_void_p(src.ptr), src.width, C.c_void_p(surface.ptr), surface.width, src.width, src.height))

    # -- beyond the reference surface -------------------------------------------------------------
    def stats(self) -> dict:
        st = RsrStats()
        self._check(self.L.rsrcu_get_stats(self.h, C.byref(st)))
        return {n: int(getattr(st, n)) for n, _ in RsrStats._fields_}

    def set_profiling(self, level):
        """0 / False: off; 1 / True: CUDA events around the tile kernel and the frame; 2: around every stage"""
        self._check(self.L.rsrcu_set_profiling(self.h, int(level)))

    def stage_ms(self) -> dict:
        ms = (C.c_float * 7)()
        self._check(self.L.rsrcu_get_stage_ms(self.h, ms))
        # slot 2 (a separate count kernel of an earlier design) is always 0 now: the setup kernel counts
        names = ("vertex", "setup_count", None, "bin_scan", "bin_fill", "tile", "frame")
        return {k: float(x) for k, x in zip(names, ms) if k}

    def device_truecolor(self):
        p, s = C.c_void_p(), C.c_int()
        self._check(self.L.rsrcu_device_truecolor(self.h, C.byref(p), C.byref(s)))
        return p.value, s.value

    def stream(self) -> int:
        p = C.c_void_p()
        self._check(self.L.rsrcu_stream(self.h, C.byref(p)))
        return p.value or 0

    def Join(self):
        """orders stream() behind every frame submitted so far (overlap mode runs tile kernels on two streams)"""
        self._check(self.L.rsrcu_join(self.h))

    def get_host_luts(self):
        rcp = np.zeros(2048, np.uint32)
        rsq = np.zeros(2048, np.uint32)
        self._check(self.L.rsrcu_get_host_luts(self.h, _ptr(rcp), _ptr(rsq)))
        return rcp, rsq

    def set_host_luts(self, rcp, rsq):
        rcp = np.ascontiguousarray(rcp, dtype=np.uint32)
        rsq = np.ascontiguousarray(rsq, dtype=np.uint32)
        assert rcp.size == 2048 and rsq.size == 2048
        self._check(self.L.rsrcu_set_host_luts(self.h, _ptr(rcp), _ptr(rsq)))


class SubmitPool:
    """Frame submission on several host threads: `contexts` rendering contexts on one GPU, each fed by its own thread,
    frames dealt round robin.  Recording a frame (the caller's job, as in the reference: node graph -> GL calls) stays on
    the caller's thread; decoding the recorded stream, building the frame's tables, launching its kernels and waiting
    for its read-back run on the pool's threads -- what the reference spreads over its job system's worker threads
    (src/rcl/rclmt/rclmt_jobsys.cxx), one frame per worker here instead of one tile per job.  Frames of one context
    complete in order; `wait(ticket)` returns when that frame has landed in its store destinations.  Every context
    keeps up to three frames in flight (upload / kernels / read-back) and uploads its own copy of static buffers."""

    def __init__(self, device: int = 0, contexts: int = 3, overlap: bool = True):
        import queue
        import threading
        self.gpus = [GPU(device) for _ in range(max(1, contexts))]
        for g in self.gpus:
            g.set_overlap(overlap)
        self._q = [queue.Queue() for _ in self.gpus]
        self._done = [0] * len(self.gpus)          # frames completed per context
        self._cv = threading.Condition()
        self._submitted = 0
        self._error = None
        self._threads = [threading.Thread(target=self._run, args=(k,), daemon=True) for k in range(len(self.gpus))]
        for t in self._threads:
            t.start()

    def _run(self, k):
        gpu, q = self.gpus[k], self._q[k]
        inflight = 0
        while True:
            rec = q.get()
            try:
                if rec is None or rec == "drain":
                    gpu.Sync()
                    with self._cv:
                        self._done[k] += inflight
                        inflight = 0
                        if rec == "drain":
                            self._drained[k] = True
                        self._cv.notify_all()
                    if rec is None:
                        return
                    continue
                gpu.Submit(rec, sync=False)
                inflight += 1
                if inflight > 2:
                    gpu.SyncFrame(2)
                    inflight -= 1
                    with self._cv:
                        self._done[k] += 1
                        self._cv.notify_all()
            except Exception as exc:  # surfaces in wait() / drain()
                with self._cv:
                    self._error = exc
                    self._cv.notify_all()
                return

    def submit(self, rec: RecordedFrame) -> int:
        """hands a recorded frame to the next context; returns its ticket"""
        t = self._submitted
        self._q[t % len(self.gpus)].put(rec)
        self._submitted += 1
        return t

    def wait(self, ticket: int):
        """blocks until frame `ticket` is complete in host memory (frames of its context before it included)"""
        k, n = ticket % len(self.gpus), ticket // len(self.gpus) + 1
        with self._cv:
            while self._done[k] < n and self._error is None:
                self._cv.wait()
            if self._error is not None:
                raise self._error

    def drain(self):
        """waits for every submitted frame"""
        self._drained = [False] * len(self.gpus)
        for q in self._q:
            q.put("drain")
        with self._cv:
            while not all(self._drained) and self._error is None:
                self._cv.wait()
            if self._error is not None:
                raise self._error

    def stats(self):
        return [g.stats() for g in self.gpus]

    def close(self):
        for q in self._q:
            q.put(None)
        for t in self._threads:
            t.join()
        for g in self.gpus:
            g.close()
