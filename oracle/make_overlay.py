#!/usr/bin/env python3
"""Generate GCC-compatible overlay headers for the reference build (test infrastructure).

The reference (usrlocalben/rsr) is an MSVC/clang-cl project.  Two of its headers use
constructs g++ rejects.  This script reads them from the read-only reference tree and writes
patched copies into oracle/_ref/overlay/ (git-ignored build output, never committed):

* src/rml/rmlv/rmlv_soa.hxx  -- qfloat2/3/4 keep `mvec4f` (which has constructors) inside
  anonymous structs inside anonymous unions ("member with constructor not allowed in anonymous
  aggregate").  Each such union is rewritten to one anonymous union per component plus an
  empty, offset-0 `v` accessor so `v[i]` keeps working.  Object layout is unchanged.
* src/rgl/rglv/rglv_mesh_store.hxx -- `throw std::exception("...")` is MSVC-only.
* dropin/rglv_gpu.cxx -- src/rgl/rglv/rglv_gpu.cxx with `#ifndef RSR_CUDA_RUNIMPL` around the body of
  GPU::RunImpl (lines 90-116), for the compiled drop-in test (see build_ref.sh).

Usage: make_overlay.py <reference_root> <overlay_out_dir>
"""
import os
import re
import sys

ACCESSOR = (
    "struct VIdx_ {\n"
    "\t\tmvec4f& operator[](int i) { return reinterpret_cast<mvec4f*>(this)[i]; }\n"
    "\t\tconst mvec4f& operator[](int i) const { return reinterpret_cast<const mvec4f*>(this)[i]; } };\n"
)


def patch_soa(text: str) -> str:
    # matches:  union { struct { mvec4f a, b, ..; }; struct { mvec4f s, t, ..; }; [comment] mvec4f v[N]; };
    pat = re.compile(
        r"union\s*\{\s*((?:(?://)?struct\s*\{\s*mvec4f\s+[a-z, ]+;\s*\};\s*)+)mvec4f\s+v\[(\d)\];\s*\};",
        re.S)

    def repl(m):
        groups = []
        for line in m.group(1).splitlines():
            line = line.strip()
            if not line or line.startswith("//"):
                continue
            names = re.search(r"mvec4f\s+([a-z, ]+);", line).group(1)
            groups.append([n.strip() for n in names.split(",")])
        n = int(m.group(2))
        out = ["[[no_unique_address]] VIdx_ v;"]
        for i in range(n):
            members = " ".join(f"mvec4f {g[i]};" for g in groups)
            out.append(f"union {{ {members} }};")
        return "\n\t".join(out)

    text, cnt = pat.subn(repl, text)
    assert cnt == 3, f"expected 3 unions in rmlv_soa.hxx, patched {cnt}"
    # the accessor type must be declared once, before the first struct
    text = text.replace("using qfloat = mvec4f;", "using qfloat = mvec4f;\n\n" + ACCESSOR, 1)
    return text


def main():
    ref, out = sys.argv[1], sys.argv[2]
    jobs = [
        ("src/rml/rmlv/rmlv_soa.hxx", patch_soa),
        ("src/rgl/rglv/rglv_mesh_store.hxx",
         lambda t: t.replace("throw std::exception(", "throw std::runtime_error(")
                    .replace("#pragma once", "#pragma once\n#include <stdexcept>", 1)),
    ]
    for rel, fn in jobs:
        src = open(os.path.join(ref, rel), encoding="utf-8", errors="replace").read()
        dst = os.path.join(out, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        with open(dst, "w", encoding="utf-8") as f:
            f.write(fn(src))
    # the drop-in build (librsr_dropin.so): rglv_gpu.cxx with the body of GPU::RunImpl compiled out under
    # -DRSR_CUDA_RUNIMPL, so that rsr_b200/host/rglv_gpu_cuda.cxx can supply it; everything else in the file
    # (Install, Reset, Retile, BinImpl, DrawImpl) is compiled as it is
    src = open(os.path.join(ref, "src/rgl/rglv/rglv_gpu.cxx"), encoding="utf-8", errors="replace").read()
    a = src.index("void GPU::RunImpl(rclmt::jobsys::Job* job) {")
    b = src.index("void GPU::BinImpl() {")
    patched = src[:a] + "#ifndef RSR_CUDA_RUNIMPL\n" + src[a:b] + "#endif  // RSR_CUDA_RUNIMPL\n\n" + src[b:]
    dst = os.path.join(out, "dropin", "rglv_gpu.cxx")
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    with open(dst, "w", encoding="utf-8") as f:
        f.write(patched)
    # an empty <intrin.h> (MSVC-only header; x86intrin.h comes from the shim)
    open(os.path.join(out, "intrin.h"), "w").close()


if __name__ == "__main__":
    main()
