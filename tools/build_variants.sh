#!/bin/bash
# A/B builds of libggnn_b200.so with different -D switches into gpurun_variants/ (shipped to the GPU box by gpurun, not
# committed): tools/ab_query.py times each.   usage: bash tools/build_variants.sh name1 "-DX=0 -DY=1" name2 "..." ...
set -euo pipefail
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
OUT="$ROOT/gpurun_variants"
mkdir -p "$OUT"
NVCC=/usr/local/cuda/bin/nvcc
FLAGS=(-std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC --expt-relaxed-constexpr)
while [ $# -ge 2 ]; do
  name="$1"; defs="$2"; shift 2
  ( d="$OUT/obj_$name"; mkdir -p "$d"
    for f in "$ROOT"/ggnn_b200/csrc/*.cu; do
      b=$(basename "$f")
      if [ "$b" = "query.cu" ] || [ ! -f "$d/$b.o" ]; then $NVCC "${FLAGS[@]}" $defs -c "$f" -o "$d/$b.o" & fi
    done; wait
    $NVCC -shared -o "$OUT/lib_$name.so" "$d"/*.o -lcurand -lcudart -gencode arch=compute_100a,code=sm_100a
    echo "built $OUT/lib_$name.so ($defs)" ) &
done
wait
