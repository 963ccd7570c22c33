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
echo "built $OUT/libggnn_ref.so"
