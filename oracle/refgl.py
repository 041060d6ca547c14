"""TEST INFRASTRUCTURE -- ctypes binding of oracle/_ref/librsr_ref.so (the unmodified reference).

`RefGPU` exposes the reference's `rglv::GPU` + `rglv::GL` surface (src/rgl/rglv/rglv_gl.hxx:182-344,
src/rgl/rglv/rglv_gpu.hxx:152-168) with the reference's own method names, so one scene-building
function can drive both this oracle and the CUDA product (`rsr_b200.GPU`) and the outputs can be
compared bit for bit.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs import
this module.  Nothing under rsr_b200/ does.
"""
from __future__ import annotations

import atexit
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_ref", "librsr_ref.so")

DROPIN_PATH = os.path.join(HERE, "_ref", "librsr_dropin.so")

_lib = None
_threads = None
_dropin = None
_dropin_threads = None


def build(force: bool = False) -> bool:
    """Build oracle/_ref/librsr_ref.so (and librsr_dropin.so) when the reference tree is present. Returns availability."""
    ref = os.environ.get("RSR_REFERENCE", "/root/reference")
    if os.path.isdir(os.path.join(ref, "src", "rgl", "rglv")) and (force or not os.path.exists(LIB_PATH) or not os.path.exists(DROPIN_PATH)):
        subprocess.check_call(["bash", os.path.join(HERE, "build_ref.sh")])
    return os.path.exists(LIB_PATH)


def available() -> bool:
    return os.path.exists(LIB_PATH)


def dropin_available() -> bool:
    return os.path.exists(DROPIN_PATH)


def _bind(path):
    L = C.CDLL(path)
    vp, ci, cf = C.c_void_p, C.c_int, C.c_float
    sigs = {
        "ref_init": (ci, [ci]),
        "ref_work_start": (None, []),
        "ref_work_end": (None, []),
        "ref_shutdown": (None, []),
        "ref_set_double_buffer": (None, [ci]),
        "ref_gpu_create": (vp, []),
        "ref_gpu_destroy": (None, [vp]),
        "ref_is_dropin": (ci, []),
        "ref_gpu_reset": (None, [vp, ci, ci, ci, ci]),
        "ref_gpu_run": (None, [vp]),
        "ref_gl_enable": (None, [vp, ci]),
        "ref_gl_disable": (None, [vp, ci]),
        "ref_gl_depth_func": (None, [vp, ci]),
        "ref_gl_depth_write_mask": (None, [vp, ci]),
        "ref_gl_color_write_mask": (None, [vp, ci]),
        "ref_gl_cull_face": (None, [vp, ci]),
        "ref_gl_scissor": (None, [vp, ci, ci, ci, ci]),
        "ref_gl_viewport": (None, [vp, ci, ci, ci, ci]),
        "ref_gl_use_program": (None, [vp, ci]),
        "ref_gl_renderbuffer_type": (None, [vp, ci, ci]),
        "ref_gl_clear_color": (None, [vp, cf, cf, cf]),
        "ref_gl_clear_depth": (None, [vp, cf]),
        "ref_gl_view_matrix": (None, [vp, vp]),
        "ref_gl_projection_matrix": (None, [vp, vp]),
        "ref_gl_normal_matrix": (None, [vp, vp]),
        "ref_gl_use_buffer": (None, [vp, ci, vp]),
        "ref_gl_uniforms": (None, [vp, vp, ci]),
        "ref_gl_bind_texture": (None, [vp, ci, vp, ci, ci, ci, ci]),
        "ref_gl_bind_texture3": (None, [vp, vp, ci]),
        "ref_gl_clear": (None, [vp, ci]),
        "ref_gl_draw_elements": (None, [vp, ci, vp, ci]),
        "ref_gl_draw_arrays": (None, [vp, ci]),
        "ref_gl_draw_elements_instanced": (None, [vp, ci, vp, ci]),
        "ref_gl_draw_arrays_instanced": (None, [vp, ci, ci]),
        "ref_gl_store_color_tc": (None, [vp, vp, ci, ci, ci, ci]),
        "ref_gl_store_color_fp": (None, [vp, vp, ci, ci, ci, ci]),
        "ref_gl_store_color_quads": (None, [vp, vp, ci, ci, ci]),
        "ref_gl_store_depth": (None, [vp, vp]),
        "ref_make_mipmap": (None, [vp, ci, vp]),
        "ref_rcp": (None, [vp, vp, ci]),
        "ref_rsqrt": (None, [vp, vp, ci]),
        "ref_oneover": (None, [vp, vp, ci]),
        "ref_mat4_mul": (None, [vp, vp, vp]),
        "ref_mat4_inverse": (None, [vp, vp]),
        "ref_look_at": (None, [vp, vp, vp, vp]),
        "ref_perspective2": (None, [cf, cf, cf, cf, vp]),
        "ref_perspective_camera": (None, [vp, cf, cf, cf, cf, vp, vp]),
        "ref_load_obj_arrays": (ci, [C.c_char_p, C.c_char_p, vp, ci, vp, ci, vp]),
        "ref_raster_coverage": (None, [vp, ci, ci, vp]),
        "ref_vraster_coverage": (None, [vp, vp, ci, ci, ci, ci, ci, ci, vp]),
        "ref_kawase_blur": (None, [vp, ci, vp, ci, ci, ci, ci]),
        "ref_glow_filter": (None, [vp, ci, vp, ci, ci, ci, vp, ci, ci, ci, ci]),
        "ref_gpu_install_shadow_program": (None, [vp]),
        "ref_mc_tables": (None, [vp, vp, vp]),
        "ref_march_surface": (ci, [cf, ci, ci, cf, vp, vp, ci, vp, vp, ci, vp]),
        "ref_render_spans": (None, [vp, ci, ci, ci, ci, ci, cf, vp, vp, vp, ci]),
        "ref_scene_load": (vp, [C.c_char_p, C.c_char_p]),
        "ref_scene_free": (None, [vp]),
        "ref_scene_render": (ci, [vp, ci, ci, ci, ci, cf, vp]),
        "ref_scene_bench": (C.c_double, [vp, ci, ci, ci, ci, ci, cf, cf, ci, vp]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    if L.ref_is_dropin():
        L.ref_dropin_flush.restype = None
        L.ref_dropin_flush.argtypes = [vp]
        L.ref_dropin_set_upload_policy.restype = None
        L.ref_dropin_set_upload_policy.argtypes = [ci, ci, ci]
    return L


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError(f"reference oracle not built: {LIB_PATH} (run oracle/build_ref.sh)")
        _lib = _bind(LIB_PATH)
    return _lib


def dropin_lib():
    """oracle/_ref/librsr_dropin.so: the reference's translation units with GPU::RunImpl replaced by the C-ABI binding
    (rsr_b200/host/rglv_gpu_cuda.cxx) -- the reference's own GL recording in front of librsrcu.so.  Needs a CUDA device."""
    global _dropin
    if _dropin is None:
        if not dropin_available():
            raise RuntimeError(f"drop-in library not built: {DROPIN_PATH} (run oracle/build_ref.sh after building librsrcu.so)")
        _dropin = _bind(DROPIN_PATH)
    return _dropin


def init(threads: int | None = None) -> int:
    """jobsys::init -- once per process; later calls return the thread count in use."""
    global _threads
    if _threads is None:
        n = threads or os.cpu_count() or 1
        _threads = lib().ref_init(int(n))
        atexit.register(lib().ref_shutdown)
    return _threads


def init_dropin(threads: int = 2) -> int:  # noqa: E302
    """the drop-in library carries its own copy of the job system (GPU::Run is a job): a small pool is enough"""
    global _dropin_threads
    if _dropin_threads is None:
        _dropin_threads = dropin_lib().ref_init(int(threads))
        atexit.register(dropin_lib().ref_shutdown)
    return _dropin_threads


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


class RefGPU:
    """The reference `rglv::GPU` (with `rqv::Install`ed programs) and its recording `GL` context."""

    def __init__(self, threads: int | None = None, double_buffer: bool = False, dropin: bool = False):
        if dropin:
            init_dropin()
            self.L = dropin_lib()
        else:
            init(threads)
            self.L = lib()
        self.dropin = dropin
        self.L.ref_set_double_buffer(1 if double_buffer else 0)
        self.h = self.L.ref_gpu_create()
        self._keep = []

    def close(self):
        if self.h:
            self.L.ref_gpu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- rglv::GPU --------------------------------------------------------------------------
    def Reset(self, size, tile_blocks=(8, 8)):
        self._keep = []
        self.size = (int(size[0]), int(size[1]))
        self.L.ref_gpu_reset(self.h, self.size[0], self.size[1], int(tile_blocks[0]), int(tile_blocks[1]))
        # GLState::reset() (rglv_gl.hxx:152-179) leaves the attachment types uninitialised; every
        # reference caller sets them right after Reset (node/gpu.cxx:132-133).  Do the same.
        self.RenderbufferType(2, 0)  # GL_COLOR_ATTACHMENT0 <- RB_COLOR_DEPTH
        self.RenderbufferType(0, 0)  # GL_DEPTH_ATTACHMENT  <- RB_COLOR_DEPTH

    def Flush(self):
        """drop-in, doubleBuffer mode: wait for the frame still in flight on the GPU"""
        if self.dropin:
            self.L.ref_dropin_flush(self.h)

    def Run(self, manage_workers: bool = True):
        if manage_workers:
            self.L.ref_work_start()
        self.L.ref_gpu_run(self.h)
        if manage_workers:
            self.L.ref_work_end()

    # -- rglv::GL ---------------------------------------------------------------------------
    def Enable(self, cap): self.L.ref_gl_enable(self.h, cap)
    def Disable(self, cap): self.L.ref_gl_disable(self.h, cap)
    def DepthFunc(self, v): self.L.ref_gl_depth_func(self.h, v)
    def DepthWriteMask(self, v): self.L.ref_gl_depth_write_mask(self.h, int(bool(v)))
    def ColorWriteMask(self, v): self.L.ref_gl_color_write_mask(self.h, int(bool(v)))
    def CullFace(self, v): self.L.ref_gl_cull_face(self.h, v)
    def Scissor(self, x, y, w, h): self.L.ref_gl_scissor(self.h, x, y, w, h)
    def Viewport(self, x, y, w, h): self.L.ref_gl_viewport(self.h, x, y, w, h)
    def UseProgram(self, pid): self.L.ref_gl_use_program(self.h, pid)
    def RenderbufferType(self, attachment, t): self.L.ref_gl_renderbuffer_type(self.h, attachment, t)
    def ClearColor(self, rgb): self.L.ref_gl_clear_color(self.h, float(rgb[0]), float(rgb[1]), float(rgb[2]))
    def ClearDepth(self, d): self.L.ref_gl_clear_depth(self.h, float(d))

    def _mat(self, m):
        """accepts a 4x4 row-major numpy matrix (math convention); the reference stores column-major"""
        a = _f32(np.asarray(m, dtype=np.float32).T.reshape(16))
        self._keep.append(a)
        return _ptr(a)

    def ViewMatrix(self, m): self.L.ref_gl_view_matrix(self.h, self._mat(m))
    def ProjectionMatrix(self, m): self.L.ref_gl_projection_matrix(self.h, self._mat(m))
    def NormalMatrix(self, m): self.L.ref_gl_normal_matrix(self.h, self._mat(m))

    def UseBuffer(self, slot, arr):
        """arr: float32 array (one SoA component) or None; a (3,N)/(2,N) array binds slot..slot+k"""
        if arr is None:
            self.L.ref_gl_use_buffer(self.h, slot, None)
            return
        a = np.asarray(arr)
        if a.ndim == 2:
            for i in range(a.shape[0]):
                self.UseBuffer(slot + i, a[i])
            return
        a = _f32(a)
        assert a.ctypes.data % 16 == 0 or a.size == 0, "SoA arrays must be 16-byte aligned"
        self._keep.append(a)
        self.L.ref_gl_use_buffer(self.h, slot, _ptr(a))

    def UseUniforms(self, data):
        b = np.ascontiguousarray(data).view(np.uint8).reshape(-1)
        assert b.size <= 128
        self._keep.append(b)
        self.L.ref_gl_uniforms(self.h, _ptr(b), int(b.size))

    def BindTexture(self, unit, texels, width, height, stride, mode):
        t = _f32(texels)
        self._keep.append(t)
        self.L.ref_gl_bind_texture(self.h, unit, _ptr(t), width, height, stride, mode)

    def BindTexture3(self, depth, dim):
        t = _f32(depth)
        self._keep.append(t)
        self.L.ref_gl_bind_texture3(self.h, _ptr(t), dim)

    def Clear(self, bits): self.L.ref_gl_clear(self.h, bits)

    def DrawElements(self, count, indices, hint=0):
        idx = np.ascontiguousarray(indices, dtype=np.uint16)
        self._keep.append(idx)
        self.L.ref_gl_draw_elements(self.h, int(count), _ptr(idx), int(hint))

    def DrawArrays(self, count): self.L.ref_gl_draw_arrays(self.h, int(count))

    def DrawElementsInstanced(self, count, indices, instance_cnt):
        idx = np.ascontiguousarray(indices, dtype=np.uint16)
        self._keep.append(idx)
        self.L.ref_gl_draw_elements_instanced(self.h, int(count), _ptr(idx), int(instance_cnt))

    def DrawArraysInstanced(self, count, instance_cnt):
        self.L.ref_gl_draw_arrays_instanced(self.h, int(count), int(instance_cnt))

    def StoreColor(self, dst: np.ndarray, gamma: bool = True):
        """dst: (H, W) uint32 -> CMD_STORE_COLOR_FULL_LINEAR_TC; (H, W, 4) float32 -> ..._LINEAR_FP"""
        if dst.dtype == np.uint32:
            h, w = dst.shape
            self._keep.append(dst)
            self.L.ref_gl_store_color_tc(self.h, _ptr(dst), w, h, dst.strides[0] // 4, int(bool(gamma)))
        else:
            assert dst.dtype == np.float32 and dst.ndim == 3 and dst.shape[2] == 4
            h, w, _ = dst.shape
            self._keep.append(dst)
            self.L.ref_gl_store_color_fp(self.h, _ptr(dst), w, h, dst.strides[0] // 16, 0)

    def StoreColorHalf(self, dst: np.ndarray):
        assert dst.dtype == np.float32 and dst.ndim == 3 and dst.shape[2] == 4
        h, w, _ = dst.shape
        self._keep.append(dst)
        self.L.ref_gl_store_color_fp(self.h, _ptr(dst), w, h, dst.strides[0] // 16, 1)

    def StoreColorQuads(self, dst: np.ndarray):
        """dst: (H/2, W/2, 4, 4) float32 = [quad row][quad][r,g,b,a][lane] -> CMD_STORE_COLOR_FULL_QUADS_FP"""
        assert dst.dtype == np.float32 and dst.ndim == 4 and dst.shape[2:] == (4, 4) and dst.ctypes.data % 16 == 0
        hq, wq = dst.shape[:2]
        self._keep.append(dst)
        self.L.ref_gl_store_color_quads(self.h, _ptr(dst), wq * 2, hq * 2, dst.strides[0] // 64)

    def StoreDepth(self, dst: np.ndarray):
        assert dst.dtype == np.float32 and dst.flags.c_contiguous
        self._keep.append(dst)
        self.L.ref_gl_store_depth(self.h, _ptr(dst))

    def InstallShadowProgram(self):
        """the depth-only BaseProgram of a `$layer`'s shadow-map GPU (src/viewer/node/gllayer.cxx:44-45)"""
        self.L.ref_gpu_install_shadow_program(self.h)


# -- free helpers ------------------------------------------------------------------------------

def make_mipmap(base: np.ndarray) -> np.ndarray:
    """(dim, dim, 4) float32 -> (2*dim, dim, 4) with the reference's stacked mip chain"""
    base = _f32(base)
    dim = base.shape[0]
    assert base.shape == (dim, dim, 4)
    out = np.zeros((2 * dim, dim, 4), dtype=np.float32)
    lib().ref_make_mipmap(_ptr(base), dim, _ptr(out))
    return out


def _map1(fn, x):
    a = _f32(x).reshape(-1)
    out = np.empty_like(a)
    fn(_ptr(a), _ptr(out), a.size)
    return out.reshape(np.shape(x))


def rcp(x): return _map1(lib().ref_rcp, x)
def rsqrt(x): return _map1(lib().ref_rsqrt, x)
def oneover(x): return _map1(lib().ref_oneover, x)


def mat4_mul(a, b):
    """row-major numpy in/out (math convention)"""
    aa, bb = _f32(np.asarray(a).T.reshape(16)), _f32(np.asarray(b).T.reshape(16))
    out = np.empty(16, np.float32)
    lib().ref_mat4_mul(_ptr(aa), _ptr(bb), _ptr(out))
    return out.reshape(4, 4).T.copy()


def mat4_inverse(a):
    aa = _f32(np.asarray(a).T.reshape(16))
    out = np.empty(16, np.float32)
    lib().ref_mat4_inverse(_ptr(aa), _ptr(out))
    return out.reshape(4, 4).T.copy()


def look_at(eye, center, up):
    """rglv::LookAt (rglv_math.cxx:66-86) -> 4x4 row-major numpy (math convention)"""
    e, c, u = (_f32(v).reshape(3) for v in (eye, center, up))
    out = np.empty(16, np.float32)
    lib().ref_look_at(_ptr(e), _ptr(c), _ptr(u), _ptr(out))
    return out.reshape(4, 4).T.copy()


def perspective2(fovy, aspect, znear, zfar):
    """rglv::Perspective2 (rglv_math.cxx:101-106) -> 4x4 row-major numpy"""
    out = np.empty(16, np.float32)
    lib().ref_perspective2(float(fovy), float(aspect), float(znear), float(zfar), _ptr(out))
    return out.reshape(4, 4).T.copy()


def perspective_camera(position, h, v, fov, aspect):
    """the viewer's $perspective node (src/viewer/node/perspective.cxx:51-69) -> (view, projection), 4x4 row-major numpy"""
    pos = _f32(position).reshape(3)
    vm, pm = np.empty(16, np.float32), np.empty(16, np.float32)
    lib().ref_perspective_camera(_ptr(pos), float(h), float(v), float(fov), float(aspect), _ptr(vm), _ptr(pm))
    return vm.reshape(4, 4).T.copy(), pm.reshape(4, 4).T.copy()


def load_obj_arrays(path, spec="PND", max_verts=65536):
    """rglv::LoadOBJ + rglv::MakeArray(mesh, spec) (rglv_obj.cxx:193-283, rglv_mesh_util.cxx:81-136): the three SoA
    vertex arrays the viewer's $mesh node binds to buffer slots 0 / 3 / 6 (each (3, nverts) float32, padded to a
    multiple of four vertices like the reference's) and the uint16 index list"""
    soa = np.zeros((9, max_verts), np.float32)
    idx = np.zeros(3 * 65536, np.uint16)
    nidx = C.c_int(0)
    nv = lib().ref_load_obj_arrays(os.fsencode(path), spec.encode(), _ptr(soa), max_verts, _ptr(idx), idx.size, C.byref(nidx))
    if nv < 0:
        raise RuntimeError(f"reference OBJ loader failed on {path}")
    return [np.ascontiguousarray(soa[3 * a:3 * a + 3, :nv]) for a in range(3)], idx[:nidx.value].copy()


def raster_coverage(xy, w, h):
    a = _f32(xy).reshape(6)
    out = np.zeros((h, w), np.uint8)
    lib().ref_raster_coverage(_ptr(a), w, h, _ptr(out))
    return out


def vraster_coverage(x3, y3, rect, w, h):
    xs = np.ascontiguousarray(x3, dtype=np.int32)
    ys = np.ascontiguousarray(y3, dtype=np.int32)
    out = np.zeros((h, w), np.uint8)
    lib().ref_vraster_coverage(_ptr(xs), _ptr(ys), rect[0], rect[1], rect[2], rect[3], w, h, _ptr(out))
    return out


# -- post filters, marching cubes, telemetry overlay ---------------------------

def kawase_blur(src: np.ndarray, dist: int) -> np.ndarray:
    """rglr::KawaseBlurFilter (rglr_kawase.cxx:22-81) over an (H, W, 4) float32 canvas"""
    s = _f32(src)
    h, w, _ = s.shape
    out = np.zeros_like(s)
    lib().ref_kawase_blur(_ptr(s), w, _ptr(out), w, w, h, int(dist))
    return out


def glow_filter(quads: np.ndarray, blur: np.ndarray, gamma: bool = True) -> np.ndarray:
    """rglr::Filter<GlowShader, sRGB|LinearColor> (rglr_algorithm.hxx:107-144, node/glow.cxx:24-39):
    quads (H/2, W/2, 4, 4) float32, blur (h, w, 4) float32 -> (H, W) uint32"""
    q = _f32(quads)
    b = _f32(blur)
    assert q.ctypes.data % 16 == 0
    hq, wq = q.shape[:2]
    bh, bw, _ = b.shape
    out = np.zeros((hq * 2, wq * 2), np.uint32)
    lib().ref_glow_filter(_ptr(q), wq, _ptr(b), bw, bh, bw, _ptr(out), wq * 2, hq * 2, wq * 2, int(bool(gamma)))
    return out


def mc_tables():
    """the reference's marching-cubes tables: (cube_edge_flags[256] int16, tritable[256][16] int8, edge_connection[12][2] uint8)"""
    flags = np.zeros(256, np.int16)
    tri = np.zeros((256, 16), np.int8)
    conn = np.zeros((12, 2), np.uint8)
    lib().ref_mc_tables(_ptr(flags), _ptr(tri), _ptr(conn))
    return flags, tri, conn


def march_surface(t: float, precision: int, fork_depth: int, rng: float, cap: int = 1 << 21):
    """`$mc` (node/mc.cxx:171-300 + rglv::march_sdf_vao) -> positions (3, n), normals (3, n), [(first_vertex, count)]"""
    pos = np.zeros((3, cap), np.float32)
    nrm = np.zeros((3, cap), np.float32)
    first = np.zeros(4096, np.int32)
    cnt = np.zeros(4096, np.int32)
    nb = C.c_int(0)
    total = lib().ref_march_surface(float(t), int(precision), int(fork_depth), float(rng), _ptr(pos), _ptr(nrm), cap,
                                    _ptr(first), _ptr(cnt), 4096, C.byref(nb))
    if total < 0:
        raise RuntimeError("ref_march_surface: capacity too small")
    blocks = [(int(first[i]), int(cnt[i])) for i in range(nb.value)]
    return np.ascontiguousarray(pos[:, :total]), np.ascontiguousarray(nrm[:, :total]), blocks


def render_spans(canvas: np.ndarray, left: int, top: int, xscale: float, spans):
    """render_jobsys (src/viewer/jobsys_vis.cxx:26-90) over spans = [(start, end, raw, lane)] into an (H, W) uint32 canvas"""
    assert canvas.dtype == np.uint32 and canvas.flags.c_contiguous
    se = np.ascontiguousarray([[s[0], s[1]] for s in spans], dtype=np.float64).reshape(-1)
    raw = np.ascontiguousarray([s[2] for s in spans], dtype=np.uint32)
    lane = np.ascontiguousarray([s[3] for s in spans], dtype=np.int32)
    h, w = canvas.shape
    lib().ref_render_spans(_ptr(canvas), w, h, w, int(left), int(top), float(xscale), _ptr(se), _ptr(raw), _ptr(lane), len(spans))


# -- the reference's scene front-end (Lua scene -> JSON -> node graph -> rglv::GPU) ----------------------

DATA_ROOT = os.path.join(HERE, "_ref")          # holds data/{scene,mesh,texture,font}, copied / generated by build_ref.sh
BUNDLED_SCENES = ("colortest", "tucker-and-dino", "instanced-cubes", "render-to-texture", "sdf-polygonization-1")


def scene_path(name: str) -> str:
    return os.path.join(DATA_ROOT, "data", "scene", name + ".json")


def scenes_available() -> bool:
    return available() and os.path.exists(scene_path("colortest"))


class RefScene:
    """A bundled data/scene/NAME.lua compiled by the reference's own node graph (src/viewer/compile.cxx, src/viewer/node/*)
    and rendered the way src/viewer/perf.cxx does.  dropin=True runs the same node graph in librsr_dropin.so, where
    rglv::GPU::RunImpl is the C-ABI binding in front of librsrcu.so: the bundled scene then renders on the GPU."""

    def __init__(self, name: str, dropin: bool = False):
        if dropin:
            init_dropin()
            self.L = dropin_lib()
        else:
            init()
            self.L = lib()
        self.name = name
        cwd = os.getcwd()
        os.chdir(DATA_ROOT)   # `$image` nodes open "data/texture/..." relative to the working directory
        try:
            self.h = self.L.ref_scene_load(os.fsencode(scene_path(name)), os.fsencode(os.path.join(DATA_ROOT, "data")))
        finally:
            os.chdir(cwd)
        if not self.h:
            raise RuntimeError(f"scene {name} did not compile / link")

    def render(self, size, t: float = 0.0, tile_blocks=(8, 8)) -> np.ndarray:
        w, h = size
        out = np.zeros((h, w), np.uint32)
        rc = self.L.ref_scene_render(self.h, w, h, int(tile_blocks[0]), int(tile_blocks[1]), float(t), _ptr(out))
        if rc != 0:
            raise RuntimeError(f"scene {self.name}: render failed ({rc})")
        return out

    def bench(self, size, frames: int, t0: float = 0.0, dt: float = 1.0 / 60.0, double_buffer: bool = True, tile_blocks=(8, 8)) -> float:
        """perf.cxx's timed loop: seconds for `frames` frames of the animation (the reference's doubleBuffer default)"""
        w, h = size
        out = np.zeros((h, w), np.uint32)
        secs = self.L.ref_scene_bench(self.h, w, h, int(tile_blocks[0]), int(tile_blocks[1]), int(frames), float(t0), float(dt),
                                      int(bool(double_buffer)), _ptr(out))
        if secs < 0:
            raise RuntimeError(f"scene {self.name}: no output node")
        return secs

    def close(self):
        if self.h:
            self.L.ref_scene_free(self.h)
            self.h = None
