"""the C++ mirror of the reference API (rsr_b200/host/rglv_cuda.hxx): compiles and links on CPU; on a
GPU box its frame hashes equal the Python route's."""
import os
import subprocess

import numpy as np
import pytest

import rsr_b200 as R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "oracle", "_build", "cpp_demo")


def build_demo():
    R.load_library()
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-o", EXE, os.path.join(ROOT, "tools", "cpp_demo.cpp"),
                           "-L" + os.path.join(ROOT, "rsr_b200"), "-lrsrcu", "-Wl,-rpath," + os.path.join(ROOT, "rsr_b200")])


def fnv1a(a):
    h = 1469598103934665603
    for b in a.tobytes():
        h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


def test_cpp_mirror_compiles_and_fails_loudly_without_gpu():
    import torch
    build_demo()
    r = subprocess.run([EXE], capture_output=True, text=True, timeout=120)
    if torch.cuda.is_available():
        assert r.returncode == 0 and r.stdout.startswith("fnv1a ")
    else:
        assert r.returncode == 3 and "no CPU fallback" in r.stdout


@pytest.mark.gpu
def test_cpp_mirror_equals_python_route(cuda_gpu, ref_gpu):
    from rsr_b200 import scenes
    build_demo()
    r = subprocess.run([EXE], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    got = int(r.stdout.split()[1], 16)

    W, H, D = 320, 180, 16
    q = np.array([[-1, 1, 1, -1], [-1, -1, 1, 1], [0, 0.3, 0, -0.3]], np.float32)
    uv = np.array([[0, 1, 1, 0], [0, 0, 1, 1]], np.float32)
    tex = np.zeros((2 * D, D, 4), np.float32)
    y, x, c = np.meshgrid(np.arange(D), np.arange(D), np.arange(4), indexing="ij")
    tex[:D] = ((x * 7 + y * 13 + c * 5) % 32).astype(np.float32) / np.float32(32.0)
    view = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, -3], [0, 0, 0, 1]], np.float32)
    f = np.float32(1.0) / np.tan(np.float32(0.39269908))
    proj = np.zeros((4, 4), np.float32)
    proj[0, 0] = np.float32(f / np.float32(16.0 / 9.0)); proj[1, 1] = f
    proj[2, 2] = np.float32(-11.0 / 9.0); proj[2, 3] = np.float32(-20.0 / 9.0); proj[3, 2] = -1
    outs = []
    for gl in (cuda_gpu, ref_gpu):
        out = np.zeros((H, W), np.uint32)
        scenes.begin(gl, (W, H))
        gl.UseProgram(R.PROGRAM_AMY)
        gl.ViewMatrix(view); gl.ProjectionMatrix(proj)
        gl.UseBuffer(0, scenes.soa(q)); gl.UseBuffer(9, scenes.soa(uv))
        gl.BindTexture(0, tex, D, D, D, R.GL_LINEAR_MIPMAP_NEAREST)
        gl.DrawElements(6, np.array([0, 1, 2, 0, 2, 3], np.uint16), 0)
        scenes.finish(gl, out)
        gl.Run()
        outs.append(out)
    assert np.array_equal(outs[0], outs[1])
    assert fnv1a(outs[1]) == got
    # the second line: the glow chain on device canvases, against the reference's own filters on the reference's stores
    from oracle import refgl
    quads, half = np.zeros((H // 2, W // 2, 4, 4), np.float32), np.zeros((H // 2, W // 2, 4), np.float32)
    gl = ref_gpu
    scenes.begin(gl, (W, H))
    gl.UseProgram(R.PROGRAM_AMY)
    gl.ViewMatrix(view); gl.ProjectionMatrix(proj)
    gl.UseBuffer(0, scenes.soa(q)); gl.UseBuffer(9, scenes.soa(uv))
    gl.BindTexture(0, tex, D, D, D, R.GL_LINEAR_MIPMAP_NEAREST)
    gl.DrawElements(6, np.array([0, 1, 2, 0, 2, 3], np.uint16), 0)
    gl.UseProgram(R.PROGRAM_DEFAULT_POST)
    gl.StoreColorQuads(quads); gl.StoreColorHalf(half)
    gl.Run()
    want = refgl.glow_filter(quads, refgl.kawase_blur(refgl.kawase_blur(half, 0), 1), True)
    assert fnv1a(want) == int(r.stdout.split()[3], 16)
    assert fnv1a(outs[0]) == got
