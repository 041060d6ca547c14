"""Synthetic scenes + camera helpers shared by tests, smoke() and bench.py.

A scene is a function `build(gl, out, ...)` that records one frame into any object exposing the
reference's `rglv::GL` method names -- `rsr_b200.GPU` (CUDA) or `oracle.refgl.RefGPU` (the
unmodified reference) -- so both renderers see byte-identical inputs.  Mesh / texture
generation is input preparation, not part of the hot path; everything is seeded.
"""
from __future__ import annotations

import os

import numpy as np

from . import (GL_BACK, GL_BLEND, GL_COLOR_BUFFER_BIT, GL_CULL_FACE, GL_DEPTH_BUFFER_BIT, GL_FRONT,
               GL_LINEAR_MIPMAP_NEAREST, GL_NEAREST_MIPMAP_NEAREST, PROGRAM_AMY, PROGRAM_DEFAULT_POST,
               PROGRAM_MANY, PROGRAM_OBJ2)


# ---- aligned SoA arrays (the reference's LoadMD does _mm_load_ps on 4 floats at a time) --------

def aligned_f32(n: int, fill: float = 0.0) -> np.ndarray:
    """float32 array of length n padded to a multiple of 4, 64-byte aligned"""
    n4 = (int(n) + 3) & ~3
    raw = np.zeros(n4 * 4 + 64, dtype=np.uint8)
    off = (-raw.ctypes.data) % 64
    a = raw[off:off + n4 * 4].view(np.float32)
    a[:] = fill
    return a


def soa(arr: np.ndarray) -> np.ndarray:
    """(k, n) array -> (k, n4) aligned, zero padded"""
    arr = np.asarray(arr, dtype=np.float32)
    k, n = arr.shape
    n4 = (n + 3) & ~3
    raw = np.zeros(k * n4 * 4 + 64, dtype=np.uint8)
    off = (-raw.ctypes.data) % 64
    out = raw[off:off + k * n4 * 4].view(np.float32).reshape(k, n4)
    out[:, :n] = arr
    return out


# ---- matrices (row-major numpy, math convention; the GL wrappers transpose) ---------------------

def perspective(fovy_deg: float, aspect: float, znear: float, zfar: float) -> np.ndarray:
    f = 1.0 / np.tan(np.radians(fovy_deg) / 2.0)
    m = np.zeros((4, 4), np.float64)
    m[0, 0] = f / aspect
    m[1, 1] = f
    m[2, 2] = (zfar + znear) / (znear - zfar)
    m[2, 3] = 2.0 * zfar * znear / (znear - zfar)
    m[3, 2] = -1.0
    return m.astype(np.float32)


def frustum(l, r, b, t, n, f) -> np.ndarray:
    m = np.zeros((4, 4), np.float64)
    m[0, 0] = 2 * n / (r - l); m[0, 2] = (r + l) / (r - l)
    m[1, 1] = 2 * n / (t - b); m[1, 2] = (t + b) / (t - b)
    m[2, 2] = -(f + n) / (f - n); m[2, 3] = -2 * f * n / (f - n)
    m[3, 2] = -1.0
    return m.astype(np.float32)


def subframe_projection(fovy_deg, aspect, znear, zfar, gx, gy, nx, ny) -> np.ndarray:
    """off-axis frustum of sub-frame (gx, gy) of an nx x ny grid (gy = 0 is the TOP row)"""
    top = znear * np.tan(np.radians(fovy_deg) / 2.0)
    right = top * aspect
    l = -right + 2 * right * gx / nx
    r = -right + 2 * right * (gx + 1) / nx
    t = top - 2 * top * gy / ny
    b = top - 2 * top * (gy + 1) / ny
    return frustum(l, r, b, t, znear, zfar)


def orthographic(l, r, b, t, n, f) -> np.ndarray:
    m = np.eye(4, dtype=np.float64)
    m[0, 0] = 2 / (r - l); m[0, 3] = -(r + l) / (r - l)
    m[1, 1] = 2 / (t - b); m[1, 3] = -(t + b) / (t - b)
    m[2, 2] = -2 / (f - n); m[2, 3] = -(f + n) / (f - n)
    return m.astype(np.float32)


def translate(x, y, z) -> np.ndarray:
    m = np.eye(4, dtype=np.float32)
    m[:3, 3] = [x, y, z]
    return m


def scale(x, y=None, z=None) -> np.ndarray:
    y = x if y is None else y
    z = x if z is None else z
    return np.diag([x, y, z, 1.0]).astype(np.float32)


def rotate(theta, x, y, z) -> np.ndarray:
    v = np.array([x, y, z], np.float64)
    v /= np.linalg.norm(v)
    c, s = np.cos(theta), np.sin(theta)
    t = 1 - c
    x, y, z = v
    m = np.eye(4, dtype=np.float64)
    m[:3, :3] = [[t * x * x + c, t * x * y - s * z, t * x * z + s * y],
                 [t * x * y + s * z, t * y * y + c, t * y * z - s * x],
                 [t * x * z - s * y, t * y * z + s * x, t * z * z + c]]
    return m.astype(np.float32)


def look_at(eye, center, up=(0, 1, 0)) -> np.ndarray:
    eye, center, up = (np.asarray(v, np.float64) for v in (eye, center, up))
    f = center - eye
    f /= np.linalg.norm(f)
    s = np.cross(f, up)
    s /= np.linalg.norm(s)
    u = np.cross(s, f)
    m = np.eye(4, dtype=np.float64)
    m[0, :3], m[1, :3], m[2, :3] = s, u, -f
    m[:3, 3] = -m[:3, :3] @ eye
    return m.astype(np.float32)


# ---- meshes -------------------------------------------------------------------------------------

def grid_mesh(nx: int, ny: int, w: float, h: float, wave: float = 0.0):
    """(pos(3,N), normal(3,N), uv(2,N), indices) of an nx x ny vertex grid in the xy plane"""
    xs, ys = np.meshgrid(np.linspace(-w / 2, w / 2, nx), np.linspace(-h / 2, h / 2, ny))
    zs = wave * np.sin(xs * 2.0) * np.cos(ys * 3.0)
    pos = np.stack([xs.ravel(), ys.ravel(), zs.ravel()]).astype(np.float32)
    nrm = np.zeros_like(pos)
    nrm[2] = 1.0
    uv = np.stack([(xs.ravel() / w + 0.5), (ys.ravel() / h + 0.5)]).astype(np.float32)
    j, i = np.meshgrid(np.arange(ny - 1), np.arange(nx - 1), indexing="ij")
    a = (j * nx + i).ravel()
    idx = np.stack([a, a + 1, a + nx, a + 1, a + nx + 1, a + nx], axis=1).ravel().astype(np.uint16)
    assert pos.shape[1] <= 32768
    return pos, nrm, uv, idx


def cube_mesh(size: float = 1.0):
    """24-vertex cube with per-face normals and uv"""
    s = size / 2
    faces = [((0, 0, 1), (1, 0, 0), (0, 1, 0)), ((0, 0, -1), (-1, 0, 0), (0, 1, 0)),
             ((1, 0, 0), (0, 0, -1), (0, 1, 0)), ((-1, 0, 0), (0, 0, 1), (0, 1, 0)),
             ((0, 1, 0), (1, 0, 0), (0, 0, -1)), ((0, -1, 0), (1, 0, 0), (0, 0, 1))]
    pos, nrm, uv, idx = [], [], [], []
    for n, u, v in faces:
        n, u, v = (np.array(a, np.float32) for a in (n, u, v))
        base = len(pos)
        for (cu, cv) in ((-1, -1), (1, -1), (1, 1), (-1, 1)):
            pos.append((n + cu * u + cv * v) * s)
            nrm.append(n)
            uv.append(((cu + 1) / 2, (cv + 1) / 2))
        idx += [base, base + 1, base + 2, base, base + 2, base + 3]
    return (np.array(pos, np.float32).T.copy(), np.array(nrm, np.float32).T.copy(),
            np.array(uv, np.float32).T.copy(), np.array(idx, np.uint16))


def icosphere(divs: int, radius: float = 1.0):
    """subdivided icosahedron: (pos(3,N), normal(3,N), indices int32 (M,3))"""
    t = (1.0 + 5 ** 0.5) / 2.0
    v = np.array([[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0], [0, -1, t], [0, 1, t], [0, -1, -t], [0, 1, -t],
                  [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]], np.float64)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    f = np.array([[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4], [11, 10, 2],
                  [10, 7, 6], [7, 1, 8], [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9], [4, 9, 5],
                  [2, 4, 11], [6, 2, 10], [8, 6, 7], [9, 8, 1]], np.int64)
    for _ in range(divs):
        edges = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]])
        key = np.sort(edges, axis=1)
        uniq, inv = np.unique(key, axis=0, return_inverse=True)
        mid = v[uniq[:, 0]] + v[uniq[:, 1]]
        mid /= np.linalg.norm(mid, axis=1, keepdims=True)
        base = len(v)
        v = np.concatenate([v, mid])
        n = len(f)
        a, b, c = base + inv[:n], base + inv[n:2 * n], base + inv[2 * n:]
        f = np.concatenate([np.stack([f[:, 0], a, c], 1), np.stack([f[:, 1], b, a], 1),
                            np.stack([f[:, 2], c, b], 1), np.stack([a, b, c], 1)])
    pos = (v * radius).T.astype(np.float32)
    nrm = v.T.astype(np.float32)
    return pos, nrm, f.astype(np.int32)


def split_mesh_u16(pos, nrm, faces, max_verts=32768, hunks=4):
    """split an indexed mesh into `hunks` spatially coherent pieces addressable with uint16 indices
    (faces sorted by centroid longitude, re-indexed per piece)"""
    cen = pos[:, faces].mean(axis=2)                       # (3, M)
    order = np.argsort(np.arctan2(cen[2], cen[0]), kind="stable")
    faces = faces[order]
    out = []
    per = -(-len(faces) // hunks)
    for start in range(0, len(faces), per):
        sub = faces[start:start + per]
        used, inv = np.unique(sub.ravel(), return_inverse=True)
        assert len(used) <= max_verts, "hunk needs more than uint16-addressable vertices; raise `hunks`"
        out.append((pos[:, used].copy(), nrm[:, used].copy(), inv.astype(np.uint16)))
    return out


def hash_texture(dim: int, seed: int, tex_id: int = 0) -> np.ndarray:
    """(dim, dim, 4) float32 in [0,1): cheap integer hash of (seed, id, x, y, channel)"""
    y, x = np.meshgrid(np.arange(dim, dtype=np.uint64), np.arange(dim, dtype=np.uint64), indexing="ij")
    out = np.empty((dim, dim, 4), np.float32)
    for ch in range(4):
        h = (x * np.uint64(73856093)) ^ (y * np.uint64(19349663)) ^ np.uint64((seed * 83492791 + tex_id * 2654435761 + ch * 97) & 0xFFFFFFFF)
        h = (h ^ (h >> np.uint64(13))) * np.uint64(0x5bd1e995) & np.uint64(0xFFFFFFFF)
        h = h ^ (h >> np.uint64(15))
        out[..., ch] = (h & np.uint64(0xFFFFFF)).astype(np.float32) / np.float32(1 << 24)
    return out


def make_mipmap(base: np.ndarray) -> np.ndarray:
    """stacked mip chain as the reference builds it (rglr_texture.cxx:33-81): (2*dim, dim, 4).
    2x2 box: ((a + b) + c) + d, then / 4 -- in float32, same order."""
    base = np.ascontiguousarray(base, dtype=np.float32)
    dim = base.shape[0]
    out = np.zeros((2 * dim, dim, 4), np.float32)
    out[:dim] = base
    src, row, size = base, dim, dim
    while size > 1:
        s = ((src[0::2, 0::2] + src[0::2, 1::2]) + src[1::2, 0::2]) + src[1::2, 1::2]
        s = (s / np.float32(4.0)).astype(np.float32)
        size //= 2
        out[row:row + size, :size] = s
        src = s
        row += size
    return out


# ---- scenes -------------------------------------------------------------------------------------

def begin(gl, size, clear=(0.2, 0.3, 0.4), tile_blocks=(8, 8)):
    """what node/gpu.cxx:126-137 does at the start of every frame"""
    gl.Reset(size, tile_blocks)
    gl.ClearColor(clear)
    gl.ClearDepth(1.0)
    gl.Clear(GL_COLOR_BUFFER_BIT | GL_DEPTH_BUFFER_BIT)


def finish(gl, out, gamma=True, depth=None, program=PROGRAM_DEFAULT_POST):
    """node/truecolor.cxx:90-108"""
    gl.UseProgram(program)
    if depth is not None:
        gl.StoreDepth(depth)
    gl.StoreColor(out, gamma)


class WavyGridScene:
    """textured, depth-tested wavy grid (program Amy, bilinear or nearest)"""

    def __init__(self, n=40, tex_dim=256, seed=1, wave=0.5, bilinear=True):
        pos, nrm, uv, idx = grid_mesh(n, n, 8.0, 5.0, wave)
        self.pos, self.nrm, self.uv, self.idx = soa(pos), soa(nrm), soa(uv), idx
        self.tex = make_mipmap(hash_texture(tex_dim, seed))
        self.tex_dim = tex_dim
        self.filter = GL_LINEAR_MIPMAP_NEAREST if bilinear else GL_NEAREST_MIPMAP_NEAREST
        self.triangles = len(idx) // 3

    def record(self, gl, size, out, depth=None, t=0.0, tile_blocks=(8, 8), proj=None, cull=None):
        w, h = size
        begin(gl, size, tile_blocks=tile_blocks)
        gl.UseProgram(PROGRAM_AMY)
        gl.ViewMatrix(translate(0, 0, -6) @ rotate(0.3 + t, 0.2, 1.0, 0.1))
        gl.ProjectionMatrix(perspective(45.0, w / h, 1.0, 100.0) if proj is None else proj)
        if cull is not None:
            gl.Enable(GL_CULL_FACE)
            gl.CullFace(cull)
        gl.UseBuffer(0, self.pos)
        gl.UseBuffer(3, self.nrm)
        gl.UseBuffer(9, self.uv)
        gl.BindTexture(0, self.tex, self.tex_dim, self.tex_dim, self.tex_dim, self.filter)
        gl.DrawElements(len(self.idx), self.idx, 0)
        finish(gl, out, True, depth)


class CubesScene:
    """instanced cubes (program Many) + lit cubes (program OBJ2), fixed seed"""

    def __init__(self, instances=300, seed=1):
        pos, nrm, uv, idx = cube_mesh(1.0)
        self.pos, self.nrm, self.uv, self.idx = soa(pos), soa(nrm), soa(uv), idx
        rng = np.random.default_rng(seed)
        self.kd = soa(rng.random((3, pos.shape[1])).astype(np.float32))
        mats = []
        for i in range(instances):
            p = rng.uniform(-12, 12, 3)
            p[2] = rng.uniform(-30, -6)
            m = translate(*p) @ rotate(rng.uniform(0, 6.28), *rng.uniform(-1, 1, 3)) @ scale(rng.uniform(0.3, 1.5))
            mats.append(m.T.reshape(16))   # column-major, as rmlm::mat4::ff
        self.mats = aligned_f32(instances * 16)
        self.mats[:instances * 16] = np.concatenate(mats)
        self.instances = instances
        self.triangles = (len(idx) // 3) * instances + (len(idx) // 3) * 8

    def record(self, gl, size, out, depth=None, t=0.0, tile_blocks=(8, 8), proj=None):
        w, h = size
        begin(gl, size, clear=(0.05, 0.05, 0.08), tile_blocks=tile_blocks)
        proj = perspective(60.0, w / h, 1.0, 200.0) if proj is None else proj
        gl.Enable(GL_CULL_FACE)
        gl.CullFace(GL_BACK)
        gl.UseProgram(PROGRAM_MANY)
        gl.UseUniforms(np.array([0.25], np.float32))
        gl.ViewMatrix(rotate(0.1 * t, 0, 1, 0))
        gl.ProjectionMatrix(proj)
        gl.UseBuffer(0, self.pos)
        gl.UseBuffer(3, self.nrm)
        gl.UseBuffer(9, self.uv)
        gl.UseBuffer(15, self.mats)
        gl.DrawElementsInstanced(len(self.idx), self.idx, self.instances)
        gl.UseProgram(PROGRAM_OBJ2)
        gl.UseBuffer(6, self.kd)
        for k in range(8):
            ang = 0.7 * k + t
            gl.ViewMatrix(translate(4 * np.cos(ang), 3 * np.sin(ang), -9 - k) @ rotate(ang, 1, 1, 0) @ scale(2.0))
            gl.DrawElements(len(self.idx), self.idx, 0)
        finish(gl, out, True, depth)


def colortest_fixture():
    """the reference's data/mesh/colortest.obj as its own loader turns it into vertex arrays (LoadOBJ + MakeArray "PND"),
    plus the camera of data/scene/colortest.lua -- tests/golden/colortest_c1.npz, written by tests/golden/make_bundled.py
    from the compiled reference (the GPU box has no reference tree)"""
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "colortest_c1.npz")
    return np.load(path)


class ColortestScene:
    """BASELINE.json configs[0] / SURVEY 8(d) C1: data/scene/colortest.lua -- one lit mesh (data/mesh/colortest.obj,
    168 faces), program OBJ2 (no lights in the scene: node/mesh.cxx:97-99), Perspective camera at (88, 80, 93),
    background sRGB(128,128,128), Default post program, sRGB store.  The GL calls are the ones the viewer's nodes make:
    node/gpu.cxx:126-137 (reset, clear), node/mesh.cxx:101-108 (matrices, buffers 0 / 3 / 6, DrawElements),
    node/truecolor.cxx:93-104 (post program, StoreColor)."""

    def __init__(self):
        f = colortest_fixture()
        self.pos, self.nrm, self.kd = (soa(f[k]) for k in ("pos", "nrm", "kd"))
        self.idx = f["idx"].astype(np.uint16)
        self.view, self.proj, self.clear = f["view"], f["proj"], tuple(float(v) for v in f["clear"])
        self.triangles = len(self.idx) // 3
        self.draws = 1

    def projection(self):
        return self.proj

    def record(self, gl, size, out, depth=None, t=0.0, tile_blocks=(8, 8), proj=None, static=False, gamma=True):
        up = {"upload": 1} if (static and hasattr(gl, "stats")) else {}
        begin(gl, size, clear=self.clear, tile_blocks=tile_blocks)
        gl.UseProgram(PROGRAM_OBJ2)
        gl.ViewMatrix(self.view)          # vmat * mmat, mmat = identity (node/gllayer.cxx:136)
        gl.ProjectionMatrix(self.proj if proj is None else proj)
        gl.UseBuffer(0, self.pos, **up)
        gl.UseBuffer(3, self.nrm, **up)
        gl.UseBuffer(6, self.kd, **up)
        gl.DrawElements(len(self.idx), self.idx, 0, **up)
        finish(gl, out, gamma, depth)


class BundledLikeScene:
    """C2: the shape of the reference's bundled scenes at once --
    the lit OBJ2 mesh of data/scene/colortest.lua (the real data/mesh/colortest.obj, 168 faces, as the reference's
    loader produces it: tests/golden/colortest_c1.npz), drawn twice; a seeded 24x24 field of
    bilinear-textured Amy quads with two 512^2 mip-mapped textures, one draw per quad
    (data/scene/tucker-and-dino.lua), and 3 x 3000 instanced cubes, program Many
    (data/scene/instanced-cubes.lua; matrices from a fixed seed instead of std::random_device,
    node/many.cxx:56).  ~110 k triangles, 580 draws."""

    def __init__(self, seed=1, cubes=3000, groups=3, field=24):
        rng = np.random.default_rng(seed)
        # lit mesh: data/mesh/colortest.obj (positions, normals, diffuse colours, indices from the reference's loader)
        fx = colortest_fixture()
        p = fx["pos"]
        self.m_pos, self.m_nrm, self.m_kd = soa(fx["pos"]), soa(fx["nrm"]), soa(fx["kd"])
        self.m_idx = fx["idx"].astype(np.uint16)
        lo, hi = p[:, :-4].min(axis=1), p[:, :-4].max(axis=1)   # (the last four vertices are MakeArray's padding)
        self.m_model = scale(3.2 / float((hi - lo).max())) @ translate(*(-(lo + hi) / 2))   # into a 3.2-unit box around the origin
        # textured quads
        q = np.array([[0, 1, 1, 0], [0, 0, 1, 1], [0, 0, 0, 0]], np.float32)
        self.q_pos = soa(q)
        self.q_uv = soa(q[:2])
        self.q_idx = np.array([0, 1, 2, 0, 2, 3], np.uint16)
        self.tex = [make_mipmap(hash_texture(512, seed + 10, i)) for i in range(2)]
        self.field = field
        # instanced cubes
        cp, cn, cuv, cidx = cube_mesh(1.0)
        self.c_pos, self.c_nrm, self.c_uv, self.c_idx = soa(cp), soa(cn), soa(cuv), cidx
        self.cubes = cubes
        self.groups = groups
        self.base = []
        for g in range(groups):
            pos = rng.uniform(-60, 60, (cubes, 3)).astype(np.float32)
            pos[:, 2] = rng.uniform(-160, -20, cubes)
            axis = rng.uniform(-1, 1, (cubes, 3))
            ang = rng.uniform(0, 6.28, cubes)
            sc = rng.uniform(0.4, 1.6, cubes)
            self.base.append((pos, axis, ang, sc))
        self.triangles = (len(self.m_idx) // 3) * 2 + field * field * 2 + groups * cubes * (len(cidx) // 3)
        self.draws = 2 + field * field + groups
        self.mats = self.instance_matrices(0.0)
        self.vertex_record_bytes = 48 * (groups * cubes * cp.shape[1] + field * field * 4) + 80 * 2 * p.shape[1]

    def unique_texels(self, stats, size=(1920, 1080)):
        """unique texels touched at the LOD the sampler selects.  Every quad maps the whole
        [0,1]^2 of its 512^2 texture onto (0.9 x 0.5) world units at z = -30, so the per-pixel texel step and with
        it the coarse LOD (rglr_texture_sampler.cxx:25-42) follow from the projection; both textures are then
        touched at that one level, counted once however many quads and taps read them"""
        h = size[1]
        px_per_unit = (h / 2.0) / (np.tan(np.radians(45.0) / 2.0) * 30.0)
        du = 512.0 / (0.9 * px_per_unit)
        lod = max(0, min(9, int(np.floor(np.log2(du * du))) >> 1))
        return 2 * (512 >> lod) ** 2

    def instance_matrices(self, t):
        """per-frame CPU work of the $many node (node/many.cxx:188-226): rebuild instance matrices"""
        out = []
        for g in range(self.groups):
            pos, axis, ang, sc = self.base[g]
            a = ang + t * (0.5 + 0.1 * g)
            ax = axis / np.linalg.norm(axis, axis=1, keepdims=True)
            c, s = np.cos(a), np.sin(a)
            x, y, z = ax[:, 0], ax[:, 1], ax[:, 2]
            tt = 1 - c
            m = np.zeros((self.cubes, 4, 4), np.float32)
            m[:, 0, 0] = tt * x * x + c; m[:, 0, 1] = tt * x * y - s * z; m[:, 0, 2] = tt * x * z + s * y
            m[:, 1, 0] = tt * x * y + s * z; m[:, 1, 1] = tt * y * y + c; m[:, 1, 2] = tt * y * z - s * x
            m[:, 2, 0] = tt * x * z - s * y; m[:, 2, 1] = tt * y * z + s * x; m[:, 2, 2] = tt * z * z + c
            m[:, :3, :3] *= sc[:, None, None]
            m[:, :3, 3] = pos
            m[:, 3, 3] = 1.0
            a16 = aligned_f32(self.cubes * 16)
            a16[:self.cubes * 16] = m.transpose(0, 2, 1).reshape(-1)   # column-major
            out.append(a16)
        return out

    def record(self, gl, size, out, depth=None, t=0.0, tile_blocks=(8, 8), proj=None, static=False, gamma=True, device_out=None):
        """static=True marks every buffer immutable (device-resident bench leg); otherwise the
        instance matrices are rebuilt and re-uploaded each frame like the reference's $many node.
        device_out=(device pointer, stride px): resolve into caller-owned device memory instead"""
        w, h = size
        up = {"upload": 1} if hasattr(gl, "stats") else {}
        dyn = up if static else {}
        mats = self.mats if static else self.instance_matrices(t)
        begin(gl, size, clear=(0.222, 0.222, 0.333), tile_blocks=tile_blocks)
        proj = perspective(45.0, w / h, 1.0, 400.0) if proj is None else proj
        gl.ProjectionMatrix(proj)

        # instanced cubes
        gl.Enable(GL_CULL_FACE)
        gl.CullFace(GL_BACK)
        gl.UseProgram(PROGRAM_MANY)
        gl.UseUniforms(np.array([0.5], np.float32))
        gl.ViewMatrix(rotate(0.05 * t, 0, 1, 0))
        gl.UseBuffer(0, self.c_pos, **up)
        gl.UseBuffer(3, self.c_nrm, **up)
        gl.UseBuffer(9, self.c_uv, **up)
        for g in range(self.groups):
            gl.UseBuffer(15, mats[g], **dyn)
            gl.DrawElementsInstanced(len(self.c_idx), self.c_idx, self.cubes, **up)

        # lit meshes
        gl.UseProgram(PROGRAM_OBJ2)
        gl.UseBuffer(0, self.m_pos, **up)
        gl.UseBuffer(3, self.m_nrm, **up)
        gl.UseBuffer(6, self.m_kd, **up)
        gl.UseBuffer(9, None)
        gl.UseBuffer(10, None)
        for k in range(2):
            gl.ViewMatrix(translate(-3.0 + 6.0 * k, 1.5 * np.sin(t + k), -9.0) @ rotate(0.7 * t + k, 0.3, 1.0, 0.2) @ self.m_model)
            gl.DrawElements(len(self.m_idx), self.m_idx, 0, **up)

        # textured quad field, orthographic-like placement in front of the camera
        gl.Disable(GL_CULL_FACE)
        gl.UseProgram(PROGRAM_AMY)
        gl.UseBuffer(0, self.q_pos, **up)
        gl.UseBuffer(3, None); gl.UseBuffer(4, None); gl.UseBuffer(5, None)
        gl.UseBuffer(6, None); gl.UseBuffer(7, None); gl.UseBuffer(8, None)
        gl.UseBuffer(9, self.q_uv, **up)
        n = self.field
        ox, oy = np.sin(t) * 0.5, np.cos(t) * 0.5
        for j in range(n):
            for i in range(n):
                k = (i + j) & 1
                gl.BindTexture(0, self.tex[k], 512, 512, 512, GL_LINEAR_MIPMAP_NEAREST, **up)
                gl.ViewMatrix(translate((i - n / 2) * 1.0 + ox, (j - n / 2) * 0.55 + oy - 1.0, -30.0 - 0.01 * (i + j)) @ scale(0.9, 0.5, 1.0))
                gl.DrawElements(6, self.q_idx, 0, **up)
        if device_out is not None:
            gl.UseProgram(PROGRAM_DEFAULT_POST)
            gl.StoreColorDevice(device_out[0], device_out[1], gamma)
        else:
            finish(gl, out, gamma, depth)


class SoupScene:
    """seeded random triangle soup in clip-ish space: slivers, degenerates, triangles crossing the
    near plane and the guard band, back faces -- the cases the reference's binner/clipper special-case
    (rglv_gpu_impl.hxx:427-494, :678-793)"""

    def __init__(self, n=600, seed=3, spread=1.6, near_cross=True, program=PROGRAM_AMY, cull=None, blend=False,
                 instanced=0, depth_func=None, tex_dim=64, bilinear=True, tiny=False, zrange=None):
        rng = np.random.default_rng(seed)
        c = rng.uniform(-spread, spread, (n, 1, 3))
        zr = zrange if zrange is not None else (-12.0, 0.5 if near_cross else -1.5)
        c[:, :, 2] = rng.uniform(zr[0], zr[1], (n, 1))
        ext = rng.choice([0.02, 0.2, 1.0, 4.0], size=(n, 1, 1), p=[0.3, 0.4, 0.2, 0.1]) if not tiny else 0.02
        tri = c + rng.normal(0, 1, (n, 3, 3)) * ext
        # a few exact degenerates and slivers
        tri[::50, 2] = tri[::50, 1]
        tri[7::60, 2] = tri[7::60, 0] + (tri[7::60, 1] - tri[7::60, 0]) * 0.5 + 1e-4
        pos = tri.reshape(-1, 3).T
        self.pos = soa(pos)
        self.nrm = soa(rng.normal(0, 1, pos.shape))
        self.kd = soa(rng.random(pos.shape))
        self.uv = soa(rng.uniform(-0.5, 2.5, (2, pos.shape[1])))
        self.idx = np.arange(3 * n, dtype=np.uint16)
        self.tex = make_mipmap(hash_texture(tex_dim, seed))
        self.tex_dim = tex_dim
        self.filter = GL_LINEAR_MIPMAP_NEAREST if bilinear else GL_NEAREST_MIPMAP_NEAREST
        self.program, self.cull, self.blend, self.depth_func = program, cull, blend, depth_func
        self.instanced = instanced
        if instanced:
            mats = [(translate(*rng.uniform(-2, 2, 3)) @ rotate(rng.uniform(0, 6), *rng.uniform(-1, 1, 3))).T.reshape(16)
                    for _ in range(instanced)]
            self.mats = aligned_f32(instanced * 16)
            self.mats[:instanced * 16] = np.concatenate(mats)
        self.shadow = rng.random((256, 256)).astype(np.float32)
        self.triangles = n * max(1, instanced)

    def record(self, gl, size, out, depth=None, tile_blocks=(8, 8), arrays=False, gamma=True, fp_out=None,
               attachments=None, post=PROGRAM_DEFAULT_POST, post_uniform=None, half_out=None, quads_out=None, viewport=None):
        from . import GL_COLOR_ATTACHMENT0, GL_DEPTH_ATTACHMENT, RB_F32, RB_RGBF32
        gl.Reset(size, tile_blocks)
        if attachments == "split":
            gl.RenderbufferType(GL_COLOR_ATTACHMENT0, RB_RGBF32)
            gl.RenderbufferType(GL_DEPTH_ATTACHMENT, RB_F32)
        if viewport is not None:
            gl.Viewport(*viewport)
        gl.ClearColor((0.1, 0.2, 0.3))
        gl.ClearDepth(1.0)
        gl.Clear(GL_COLOR_BUFFER_BIT | GL_DEPTH_BUFFER_BIT)
        self.draw(gl, size, arrays)
        gl.UseProgram(post)
        if post_uniform is not None:
            gl.UseUniforms(np.array([post_uniform], np.float32))
        if depth is not None:
            gl.StoreDepth(depth)
        if fp_out is not None:
            gl.StoreColor(fp_out)
        if half_out is not None:
            gl.StoreColorHalf(half_out)
        if quads_out is not None:
            gl.StoreColorQuads(quads_out)
        gl.StoreColor(out, gamma)

    def draw(self, gl, size, arrays=False, color_write=None, depth_write=None, depth_test=None, depth_func=None, blend=None):
        """the soup's state and draw call alone (inside a frame somebody else begins and finishes); the keyword
        arguments override the pipeline state for this draw (depth pre-pass / EQUAL passes)"""
        from . import GL_BLEND, GL_DEPTH_TEST, PROGRAM_OBJ2S, PROGRAM_PATTERN
        w, h = size
        if color_write is not None:
            gl.ColorWriteMask(color_write)
        if depth_write is not None:
            gl.DepthWriteMask(depth_write)
        if depth_test is not None:
            (gl.Enable if depth_test else gl.Disable)(GL_DEPTH_TEST)
        gl.UseProgram(self.program)
        gl.ViewMatrix(translate(0.1, -0.05, -1.0) @ rotate(0.2, 0.1, 1.0, 0.3))
        gl.ProjectionMatrix(perspective(70.0, w / h, 0.5, 50.0))
        if self.cull is not None:
            gl.Enable(GL_CULL_FACE)
            gl.CullFace(self.cull)
        else:
            gl.Disable(GL_CULL_FACE)
        if self.blend if blend is None else blend:
            gl.Enable(GL_BLEND)
        else:
            gl.Disable(GL_BLEND)
        if depth_func is not None or self.depth_func is not None:
            gl.DepthFunc(depth_func if depth_func is not None else self.depth_func)
        gl.UseBuffer(0, self.pos)
        gl.UseBuffer(3, self.nrm)
        gl.UseBuffer(6, self.kd)
        gl.UseBuffer(9, self.uv)
        gl.BindTexture(0, self.tex, self.tex_dim, self.tex_dim, self.tex_dim, self.filter)
        if self.program == PROGRAM_MANY:
            gl.UseUniforms(np.array([0.75], np.float32))
        if self.program == PROGRAM_PATTERN:
            gl.UseUniforms(np.array([0.25, 0.5, 0, 0, 0, float(h), 0, 0], np.float32))
        if self.program == PROGRAM_OBJ2S:
            m2s = (perspective(60.0, 1.0, 0.5, 60.0) @ translate(0.3, 0.2, -2.0)).T.reshape(16)
            u = np.concatenate([m2s, [0.5, 2.0, 1.0], [0.1, -0.6, -0.8], [0.3]]).astype(np.float32)
            gl.UseUniforms(u)
            gl.BindTexture3(self.shadow, 256)
        if self.program == 5:  # Depth program samples the depth texture
            gl.BindTexture3(self.shadow, 256)
        if self.instanced:
            gl.UseBuffer(15, self.mats)
            if arrays:
                gl.DrawArraysInstanced(len(self.idx), self.instanced)
            else:
                gl.DrawElementsInstanced(len(self.idx), self.idx, self.instanced)
        elif arrays:
            gl.DrawArrays(len(self.idx))
        else:
            gl.DrawElements(len(self.idx), self.idx, 0)


class FillStressScene:
    """C4: `layers` full-screen layers drawn back to front (every fragment passes LESS and is shaded),
    each layer a grid of quads, every quad bound to its own distinct 1024^2 mip-mapped RGBA32F texture
    sampled 1 texel : 1 pixel with bilinear filtering (program Amy).  No texel is reused across
    quads, so texture traffic really comes from HBM.  `size` is one <= 2048 px (sub-)frame."""

    def __init__(self, layers=8, size=(1920, 1080), quads=(2, 2), seed=11, tex_dim=1024, front_to_back=False, fast_textures=False):
        self.layers, self.size, self.quads, self.tex_dim = layers, size, quads, tex_dim
        w, h = size
        qw, qh = w // quads[0], h // quads[1]
        assert qw <= tex_dim and qh <= tex_dim
        self.items = []
        for layer in range(layers):
            for qy in range(quads[1]):
                for qx in range(quads[0]):
                    x0, y0 = qx * qw, qy * qh
                    z = -(2.0 + layer) if front_to_back else -(2.0 + (layers - 1 - layer))
                    pos = np.array([[x0, x0 + qw, x0 + qw, x0], [y0, y0, y0 + qh, y0 + qh], [z, z, z, z]], np.float32)
                    uv = np.array([[0, qw / tex_dim, qw / tex_dim, 0], [0, 0, qh / tex_dim, qh / tex_dim]], np.float32)
                    tid = len(self.items)
                    if fast_textures:
                        # (128 textures at 4K: a seeded generator per texture instead of the integer hash, same layout)
                        base = np.random.default_rng(seed * 1000003 + tid).random((tex_dim, tex_dim, 4), dtype=np.float32)
                    else:
                        base = hash_texture(tex_dim, seed, tid)
                    self.items.append((soa(pos), soa(uv), make_mipmap(base)))
        self.idx = np.array([0, 1, 2, 0, 2, 3], np.uint16)
        self.triangles = 2 * len(self.items)
        self.draws = len(self.items)
        self.vertex_record_bytes = 48 * 4 * len(self.items)

    def unique_texels(self, stats, size=None):
        # 1:1 mapping by construction: every shaded fragment touches its own texel (+ shared borders)
        w, h = size if size is not None else self.size
        return self.layers * w * h

    def projection(self):
        """of the whole frame (`size` of the constructor); a sub-frame uses crop @ projection (rsr_b200.subframes)"""
        return orthographic(0, self.size[0], 0, self.size[1], 1.0, 20.0)

    def record(self, gl, size, out, depth=None, t=0.0, tile_blocks=(8, 8), static=False, gamma=True, proj=None, device_out=None):
        assert proj is not None or tuple(size) == tuple(self.size)
        up = {"upload": 1} if hasattr(gl, "stats") else {}
        begin(gl, size, clear=(0.0, 0.0, 0.0), tile_blocks=tile_blocks)
        gl.UseProgram(PROGRAM_AMY)
        gl.ViewMatrix(np.eye(4, dtype=np.float32))
        gl.ProjectionMatrix(self.projection() if proj is None else proj)
        for pos, uv, tex in self.items:
            gl.UseBuffer(0, pos, **up)
            gl.UseBuffer(9, uv, **up)
            gl.BindTexture(0, tex, self.tex_dim, self.tex_dim, self.tex_dim, GL_LINEAR_MIPMAP_NEAREST, **up)
            gl.DrawElements(6, self.idx, 0, **up)
        if device_out is not None:
            gl.UseProgram(PROGRAM_DEFAULT_POST)
            gl.StoreColorDevice(device_out[0], device_out[1], gamma)
        else:
            finish(gl, out, gamma, depth)


class GeometryStressScene:
    """C3: a field of finely subdivided icospheres, each triangle ~1 pixel, back faces culled, program
    Amy with uv = normal.xy; sphere centres from a fixed seed.  The mesh is split into uint16-indexable
    hunks (<= 32768 vertices each) like the reference's 4-hunk icosphere (rglv_icosphere.cxx:16-145)."""

    def __init__(self, spheres=30, divs=6, seed=7, size=(1920, 1080), radius_px=90.0, tex_dim=256):
        p, n, f = icosphere(divs, 1.0)
        self.hunks = []
        for hp, hn, hidx in split_mesh_u16(p, n, f, 32768):
            self.hunks.append((soa(hp), soa(hn), soa(hn[:2] * 0.5 + 0.5), hidx))
        rng = np.random.default_rng(seed)
        self.size = size
        self.fov, self.zn, self.zf = 45.0, 1.0, 400.0
        w, h = size
        # world radius that projects to radius_px at depth z:  r = radius_px * 2 z tan(fov/2) / h
        self.centres = []
        for _ in range(spheres):
            z = rng.uniform(20.0, 120.0)
            r = radius_px * 2 * z * np.tan(np.radians(self.fov) / 2) / h
            half_h = z * np.tan(np.radians(self.fov) / 2)
            half_w = half_h * w / h
            self.centres.append((rng.uniform(-half_w, half_w), rng.uniform(-half_h, half_h), -z, r, rng.uniform(0, 6.28)))
        self.tex = make_mipmap(hash_texture(tex_dim, seed))
        self.tex_dim = tex_dim
        self.triangles = spheres * len(f)
        self.draws = spheres * len(self.hunks)
        self.vertex_record_bytes = 48 * spheres * p.shape[1]

    def unique_texels(self, stats, size=None):
        # uv = normal.xy / 2 + 0.5: a sphere's visible hemisphere covers the unit disc of the texture at LOD 0
        # (radius_px > tex_dim / 4), every sphere the same texels
        return min(stats["fragments_shaded"], int(self.tex_dim * self.tex_dim * np.pi / 4))

    def projection(self):
        return perspective(self.fov, self.size[0] / self.size[1], self.zn, self.zf)

    def record(self, gl, size, out, depth=None, t=0.0, tile_blocks=(8, 8), proj=None, static=False, gamma=True, device_out=None):
        w, h = size
        up = {"upload": 1} if hasattr(gl, "stats") else {}
        begin(gl, size, clear=(0.02, 0.02, 0.05), tile_blocks=tile_blocks)
        gl.UseProgram(PROGRAM_AMY)
        gl.Enable(GL_CULL_FACE)
        gl.CullFace(GL_BACK)
        gl.ProjectionMatrix(perspective(self.fov, w / h, self.zn, self.zf) if proj is None else proj)
        gl.BindTexture(0, self.tex, self.tex_dim, self.tex_dim, self.tex_dim, GL_LINEAR_MIPMAP_NEAREST, **up)
        for (cx, cy, cz, r, ang) in self.centres:
            gl.ViewMatrix(translate(cx, cy, cz) @ rotate(ang + 0.2 * t, 0.3, 1.0, 0.1) @ scale(r))
            for hp, hn, huv, hidx in self.hunks:
                gl.UseBuffer(0, hp, **up)
                gl.UseBuffer(3, hn, **up)
                gl.UseBuffer(9, huv, **up)
                gl.DrawElements(len(hidx), hidx, 0, **up)
        if device_out is not None:
            gl.UseProgram(PROGRAM_DEFAULT_POST)
            gl.StoreColorDevice(device_out[0], device_out[1], gamma)
        else:
            finish(gl, out, gamma, depth)
