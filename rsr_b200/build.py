"""Builds rsr_b200/librsrcu.so (the C-ABI library of include/rsrcu.h) in-tree with nvcc for sm_100a."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "librsrcu.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17", "-diag-suppress", "550",
    # the reference is SSE code with separate mul/add: never contract into FMA, keep IEEE div/sqrt,
    # keep denormals (see csrc/dev_math.cuh)
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC,-O2,-ffp-contract=off,-fno-fast-math",
]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))]


STAMP = LIB + ".stamp"


def source_hash() -> str:
    """content hash of everything the library is built from (file times do not survive a copy to another box)"""
    import hashlib
    h = hashlib.sha1(" ".join(NVCC_FLAGS).encode())
    for p in sources() + [os.path.join(HERE, "..", "include", "rsrcu.h")]:
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    try:
        with open(STAMP) as f:
            return f.read().strip() != source_hash()
    except OSError:
        return False   # a shipped library without its stamp: take it as it is


def build(force: bool = False, verbose: bool = False) -> str:
    """nvcc -gencode arch=compute_100a,code=sm_100a ... -> rsr_b200/librsrcu.so"""
    if not force and not needs_build():
        return LIB
    obj = os.path.join(HERE, "host_luts.o")
    subprocess.check_call(["g++", "-O2", "-msse2", "-fPIC", "-std=c++17", "-c",
                           os.path.join(CSRC, "host_luts.cpp"), "-o", obj])
    cmd = [_nvcc(), *NVCC_FLAGS, "-shared", "-o", LIB, os.path.join(CSRC, "rsrcu.cu"), obj]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    subprocess.check_call(cmd)
    with open(STAMP, "w") as f:
        f.write(source_hash())
    return LIB


def build_variant(name: str, defines: list[str]) -> str:
    """Developer knob: builds rsr_b200/variants/librsrcu_<name>.so with extra -D flags (kernel tuning
    experiments; select it at run time with RSRCU_LIB=<path>)."""
    vdir = os.path.join(HERE, "variants")
    os.makedirs(vdir, exist_ok=True)
    out = os.path.join(vdir, f"librsrcu_{name}.so")
    obj = os.path.join(HERE, "host_luts.o")
    if not os.path.exists(obj):
        build(force=True)
    subprocess.check_call([_nvcc(), *NVCC_FLAGS, *[f"-D{d}" for d in defines], "-shared", "-o", out,
                           os.path.join(CSRC, "rsrcu.cu"), obj])
    return out


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
