#!/usr/bin/env bash
# TEST INFRASTRUCTURE -- builds the parity oracle from the UNMODIFIED reference sources.
#
# Compiles the reference's own hot-path translation units where they lie under
# $RSR_REFERENCE (default /root/reference, read-only) plus oracle/ref_harness.cpp into
#   oracle/_ref/librsr_ref.so          (git-ignored; travels to the GPU box with gpurun)
#   oracle/_ref/rglv_triangle_test     (the reference's own rglv_triangle.t.cxx, run as a check)
#   oracle/_ref/librsr_dropin.so       (the reference with GPU::RunImpl replaced by the C-ABI binding: drop-in test)
# Nothing from the reference tree is copied into the repository; the two g++-compat overlay
# headers are generated into oracle/_ref/overlay/ by make_overlay.py at build time.
# The reference's own build system (bazel / MSVC) is not used.
#
# Flags follow SURVEY.md 8(c): -O2 -msse4.1 -ffp-contract=off, no -march=native, no fast-math.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${RSR_REFERENCE:-/root/reference}"
OUT="$HERE/_ref"
if [ ! -d "$REF/src/rgl/rglv" ]; then
  echo "build_ref.sh: reference tree not found at $REF (expected on the GPU box; using prebuilt $OUT)" >&2
  exit 0
fi
mkdir -p "$OUT/obj"
python3 "$HERE/make_overlay.py" "$REF" "$OUT/overlay"

CXX="${CXX:-g++}"
FLAGS=(-std=c++17 -O2 -msse4.1 -ffp-contract=off -fpermissive -w -Wno-psabi -fPIC -DPLATFORM_NULL -DNDEBUG
       -I"$OUT/overlay" -I"$REF" -I"$REF/3rdparty/fmt/include"
       -include "$HERE/ref_shim.h" '-D__declspec(x)=__attribute__((aligned(16)))')

SRCS=(
  src/rgl/rglv/rglv_gpu.cxx
  src/rgl/rglv/rglv_gl.cxx
  src/rgl/rglv/rglv_math.cxx
  src/rgl/rglr/rglr_algorithm.cxx
  src/rgl/rglr/rglr_canvas_util.cxx
  src/rgl/rglr/rglr_texture.cxx
  src/rgl/rglr/rglr_texture_sampler.cxx
  src/rcl/rclmt/rclmt_jobsys.cxx
  src/rcl/rclmt/rclmt_barrier.cxx
  src/rml/rmlm/rmlm_mat4.cxx
  src/rml/rmlv/rmlv_vec.cxx
  src/rgl/rglv/rglv_obj.cxx
  src/rgl/rglv/rglv_mesh.cxx
  src/rgl/rglv/rglv_mesh_util.cxx
  src/rgl/rglv/rglv_material.cxx
  src/rgl/rglr/rglr_kawase.cxx
  src/rgl/rglv/rglv_marching_cubes.cxx
  src/viewer/jobsys_vis.cxx
  src/viewer/compile.cxx
  src/viewer/fontloader.cxx
  src/rcl/rclx/rclx_gason_util.cxx
  src/rcl/rclma/rclma_framepool.cxx
  src/rml/rmlg/rmlg_noise.cxx
  src/rgl/rglv/rglv_icosphere.cxx
  src/rgl/rglv/rglv_camera.cxx
  src/rgl/rglv/rglv_mesh_store.cxx
  src/rgl/rglr/rglr_texture_load.cxx
  src/rgl/rglr/rglr_texture_store.cxx
  3rdparty/gason/gason.cpp
  3rdparty/picopng/picopng.cpp
  src/viewer/shaders.cxx
  src/viewer/shaders_envmap.cxx
  src/viewer/shaders_wireframe.cxx
  3rdparty/fmt/src/format.cc
)
# the scene front-end: every node of src/viewer/node
for f in "$REF"/src/viewer/node/*.cxx; do SRCS+=("src/viewer/node/$(basename "$f")"); done
OBJS=()
pids=()
# the vendored Lua 5.4 (3rdparty/lua): the library for `$writer`'s font loader, the interpreter for scene.lua -> JSON
LUAOBJS=()
for f in "$REF"/3rdparty/lua/*.c; do
  b="$(basename "$f" .c)"
  case "$b" in lua|luac) continue;; esac
  o="$OUT/obj/lua_$b.o"
  LUAOBJS+=("$o")
  if [ ! -f "$o" ] || [ "$f" -nt "$o" ]; then gcc -O2 -fPIC -w -DLUA_USE_POSIX -c "$f" -o "$o" & pids+=($!); fi
done
for s in "${SRCS[@]}"; do
  o="$OUT/obj/$(echo "$s" | tr '/' '_').o"
  OBJS+=("$o")
  if [ ! -f "$o" ] || [ "$REF/$s" -nt "$o" ] || [ "$HERE/ref_shim.h" -nt "$o" ]; then
    extra=()
    case "$s" in src/viewer/node/many.cxx|src/viewer/node/particles.cxx) extra=(-DRSR_FIXED_SEED);; esac
    "$CXX" "${FLAGS[@]}" "${extra[@]}" -c "$REF/$s" -o "$o" &
    pids+=($!)
  fi
done
ho="$OUT/obj/ref_harness.o"
"$CXX" "${FLAGS[@]}" -c "$HERE/ref_harness.cpp" -o "$ho" &
pids+=($!)
so="$OUT/obj/ref_scene.o"
OBJS+=("$so")
"$CXX" "${FLAGS[@]}" -c "$HERE/ref_scene.cpp" -o "$so" &
pids+=($!)
# POSIX stand-ins for the Windows-only string / path helpers the reference's OBJ loader calls (see the file's header)
po="$OUT/obj/ref_posix_util.o"
OBJS+=("$po")
"$CXX" "${FLAGS[@]}" -c "$HERE/ref_posix_util.cpp" -o "$po" &
pids+=($!)
for p in "${pids[@]}"; do wait "$p"; done
OBJS+=("${LUAOBJS[@]}")
"$CXX" -shared -o "$OUT/librsr_ref.so" "${OBJS[@]}" "$ho" -lpthread -ldl -Wl,--no-undefined

# ---- bundled scenes as inputs: the reference's data/scene/*.lua -> JSON through its own host.lua ------------------
# (outputs only under oracle/_ref/, git-ignored like the libraries: they travel to the GPU box with gpurun)
gcc -O2 -w -DLUA_USE_POSIX "$REF/3rdparty/lua/lua.c" "${LUAOBJS[@]}" -o "$OUT/lua" -lm -ldl
mkdir -p "$OUT/data/scene" "$OUT/data/mesh" "$OUT/data/texture" "$OUT/data/font"
cp -f "$REF"/data/mesh/*.obj "$REF"/data/mesh/*.mtl "$OUT/data/mesh/"
cp -f "$REF"/data/texture/*.png "$OUT/data/texture/"
cp -f "$REF"/data/font/* "$OUT/data/font/" 2>/dev/null || true
for sc in colortest tucker-and-dino instanced-cubes render-to-texture sdf-polygonization-1 particles oldschool plusrqdq auraforlaura writer; do
  if [ -f "$REF/data/scene/$sc.lua" ]; then
    (cd "$REF/data/scene" && "$OUT/lua" sc.lua "$sc.lua" > "$OUT/data/scene/$sc.json") || echo "build_ref.sh: scene $sc did not convert" >&2
  fi
done

# ---- the compiled drop-in: the reference's own GL / GLState / command stream in front of librsrcu.so ------------
# Same translation units, except that GPU::RunImpl's body (rglv_gpu.cxx:90-116) is compiled out and supplied by
# rsr_b200/host/rglv_gpu_cuda.cxx, which forwards the recorded stream to the C ABI of include/rsrcu.h.
# tests/test_dropin_gpu.py renders through both libraries and compares bit for bit.
RSRCU_DIR="$(cd "$HERE/../rsr_b200" && pwd)"
if [ -f "$RSRCU_DIR/librsrcu.so" ]; then
  DOBJS=()
  for o in "${OBJS[@]}"; do
    # (GPU::RunImpl comes from rglv_gpu_cuda.cxx; the `$buffers` / `$kawase` / `$glow` nodes from post_nodes_cuda.cxx)
    case "$o" in *src_rgl_rglv_rglv_gpu.cxx.o|*src_viewer_node_buffers.cxx.o|*src_viewer_node_kawase.cxx.o|*src_viewer_node_glow.cxx.o) ;; *) DOBJS+=("$o");; esac
  done
  "$CXX" "${FLAGS[@]}" -DRSR_CUDA_RUNIMPL -c "$OUT/overlay/dropin/rglv_gpu.cxx" -o "$OUT/obj/dropin_rglv_gpu.o" &
  "$CXX" "${FLAGS[@]}" -DRSR_CUDA_RUNIMPL -I"$HERE/../include" -c "$RSRCU_DIR/host/rglv_gpu_cuda.cxx" -o "$OUT/obj/dropin_rglv_gpu_cuda.o" &
  "$CXX" "${FLAGS[@]}" -DRSR_CUDA_RUNIMPL -c "$HERE/ref_harness.cpp" -o "$OUT/obj/dropin_ref_harness.o" &
  "$CXX" "${FLAGS[@]}" -DRSR_CUDA_RUNIMPL -I"$HERE/../include" -c "$RSRCU_DIR/host/post_nodes_cuda.cxx" -o "$OUT/obj/dropin_post_nodes_cuda.o" &
  wait
  "$CXX" -shared -o "$OUT/librsr_dropin.so" "${DOBJS[@]}" "$OUT/obj/dropin_rglv_gpu.o" "$OUT/obj/dropin_rglv_gpu_cuda.o" \
      "$OUT/obj/dropin_post_nodes_cuda.o" "$OUT/obj/dropin_ref_harness.o" -L"$RSRCU_DIR" -lrsrcu -Wl,-rpath,'$ORIGIN/../../rsr_b200' -lpthread
  echo "build_ref.sh: built $OUT/librsr_dropin.so"
else
  echo "build_ref.sh: rsr_b200/librsrcu.so not built yet, skipping librsr_dropin.so" >&2
fi

# the reference's own rasteriser test (plain main(), no gtest): fill-rule KAT + UV interpolation
"$CXX" "${FLAGS[@]}" "$REF/src/rgl/rglv/rglv_triangle.t.cxx" \
    "$OUT/obj/src_rgl_rglr_rglr_algorithm.cxx.o" "$OUT/obj/3rdparty_fmt_src_format.cc.o" \
    -o "$OUT/rglv_triangle_test" -lpthread
"$OUT/rglv_triangle_test" > "$OUT/rglv_triangle_test.log" 2>&1 \
  && echo "build_ref.sh: reference rglv_triangle.t.cxx PASSED" \
  || { echo "build_ref.sh: reference rglv_triangle.t.cxx FAILED"; cat "$OUT/rglv_triangle_test.log"; exit 1; }
echo "build_ref.sh: built $OUT/librsr_ref.so"
