#!/bin/bash
# TEST INFRASTRUCTURE ONLY. Compiles the UNMODIFIED reference sources where they lie under
# /root/reference (never copied into this repo) into oracle/_ref/libggnn_ref.so for sm_100a,
# using the glog stand-in in oracle/ref_shim (glog/gflags cannot be fetched offline), and links
# the dump/timing driver oracle/ref_driver.cpp against it. Outputs go only to oracle/_ref/
# (git-ignored, but shipped to the GPU box by gpurun). The reference's own CMake build is not used.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${GGNN_REFERENCE:-/root/reference}"
OUT="$HERE/_ref"
OBJ="$OUT/obj"
[ -d "$REF/src/ggnn" ] || { echo "reference sources not found at $REF (expected on the build container only)"; exit 0; }
mkdir -p "$OBJ"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-std=c++20 --expt-relaxed-constexpr -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a
       -Xcompiler -fPIC -I"$REF/include" -I"$HERE/ref_shim")
SRCS=$(cd "$REF/src/ggnn" && ls base/*.cu base/*.cpp query/*.cu construction/*.cu)
pids=()
for s in $SRCS; do
  o="$OBJ/$(echo "$s" | tr '/' '_').o"
  if [ ! -f "$o" ] || [ "$REF/src/ggnn/$s" -nt "$o" ]; then
    ( "$NVCC" "${FLAGS[@]}" -x cu -c "$REF/src/ggnn/$s" -o "$o" ) &
    pids+=($!)
    # at most 8 parallel compiles
    while [ "$(jobs -rp | wc -l)" -ge "${JOBS:-8}" ]; do sleep 0.5; done
  fi
done
for p in "${pids[@]:-}"; do [ -n "$p" ] && wait "$p"; done
"$NVCC" -shared -o "$OUT/libggnn_ref.so" "$OBJ"/*.o -lcurand -lcudart
if [ -f "$HERE/ref_driver.cpp" ]; then
  "$NVCC" "${FLAGS[@]}" -x cu "$HERE/ref_driver.cpp" -o "$OUT/ref_driver" \
      -L"$OUT" -lggnn_ref -lcurand -Xlinker -rpath -Xlinker '$ORIGIN'
fi
# Host-only driver of the reference's ResultMerger / Evaluator (no GPU needed): tools/gen_host_golden.py runs it to make
# tests/golden/host_merge_eval.npz.  Plain g++: the program defines cudaPeekAtLastError itself (see its header comment).
if [ -f "$HERE/ref_host_check.cpp" ]; then
  g++ -std=c++20 -O2 -I"$REF/include" -I"$HERE/ref_shim" -I/usr/local/cuda/include "$HERE/ref_host_check.cpp" \
      -o "$OUT/ref_host_check" -rdynamic -L"$OUT" -lggnn_ref -L/usr/local/cuda/lib64 -lcudart -Wl,-rpath,'$ORIGIN'
fi
# The hybrid (compiled): the same unmodified reference objects, except that its two thin
# host -> CUDA launcher files (query_kernels.cu, graph_construction.cu) are replaced by
# integration/reference_launchers/*.cu, which call libggnn_b200.so through the C ABI.
REPO="$(cd "$HERE/.." && pwd)"
if [ -f "$REPO/ggnn_b200/libggnn_b200.so" ]; then
  HOBJ="$OUT/obj_hybrid"
  mkdir -p "$HOBJ"
  for f in query_kernels_b200 graph_construction_b200; do
    src="$REPO/integration/reference_launchers/$f.cu"
    if [ ! -f "$HOBJ/$f.o" ] || [ "$src" -nt "$HOBJ/$f.o" ] || [ "$REPO/include/ggnn_b200.h" -nt "$HOBJ/$f.o" ]; then
      "$NVCC" "${FLAGS[@]}" -I"$REPO/include" -c "$src" -o "$HOBJ/$f.o" &
    fi
  done
  wait
  REF_OBJS=$(ls "$OBJ"/*.o | grep -v -e 'query_query_kernels' -e 'construction_graph_construction')
  "$NVCC" -shared -o "$OUT/libggnn_ref_hybrid.so" $REF_OBJS "$HOBJ"/*.o -L"$REPO/ggnn_b200" -lggnn_b200 -lcurand -lcudart \
      -Xlinker -rpath -Xlinker '$ORIGIN/../../ggnn_b200'
  if [ -f "$HERE/ref_driver.cpp" ]; then
    "$NVCC" "${FLAGS[@]}" -x cu "$HERE/ref_driver.cpp" -o "$OUT/ref_driver_hybrid" \
        -L"$OUT" -lggnn_ref_hybrid -L"$REPO/ggnn_b200" -lggnn_b200 -lcurand -Xlinker -rpath -Xlinker '$ORIGIN' \
        -Xlinker -rpath -Xlinker '$ORIGIN/../../ggnn_b200'
  fi
  echo "built $OUT/libggnn_ref_hybrid.so"
fi
echo "built $OUT/libggnn_ref.so"
