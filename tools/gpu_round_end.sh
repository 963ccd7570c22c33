#!/usr/bin/env bash
# Everything a round needs from the GPU box in ONE gpurun call (every call pays 25-50 s of fixed box time):
#   gpurun --timeout 2400 -- 'ROUND=02 bash tools/gpu_round_end.sh [quick]'
# Outputs land in gpurun_out/r${ROUND}_*; copy the summaries into profiles/ afterwards (tools/ncu_summary.py,
# tools/ncu_ops.py, tools/ncu_lines.py read the .ncu-rep / CSVs on the CPU).  "quick" skips the ncu --set full captures
# and the reference arm.
set -u
R=${ROUND:-02}
O=gpurun_out
mkdir -p $O
quick=${1:-}
python -c "import __graft_entry__ as g; g.smoke()" > $O/r${R}_smoke.log 2>&1; echo "smoke rc=$?"
timeout 1200 python -m pytest tests -m gpu -q --durations=8 > $O/r${R}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r${R}_pytest_gpu.log
timeout 900 python bench.py > $O/r${R}_bench_ours.json 2> $O/r${R}_bench_ours.err; echo "bench rc=$?"; cut -c1-260 $O/r${R}_bench_ours.json
if [ -z "$quick" ]; then
  timeout 600 python bench.py --impl reference > $O/r${R}_bench_reference.json 2> $O/r${R}_bench_reference.err; echo "reference rc=$?"; cut -c1-200 $O/r${R}_bench_reference.json
fi
timeout 120 python tools/bf_i8_check.py > $O/r${R}_bf_i8_check.log 2>&1; echo "bf_i8_check rc=$?"; tail -3 $O/r${R}_bf_i8_check.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r${R}_launches_bench.csv \
  python bench.py --steps 2 --warmup 3 --batches-per-step 4 --no-config4 --no-extras > $O/r${R}_ncu_list.log 2>&1; echo "launch list rc=$?"
if [ -z "$quick" ]; then
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:query_kernel -s 4 -c 1 -f -o $O/r${R}_query_kernel \
    python bench.py --steps 2 --warmup 3 --batches-per-step 4 --streams 1 --e2e-depth 1 --no-config4 --no-extras > $O/r${R}_ncu_full.log 2>&1; echo "full capture rc=$?"
  timeout 300 ncu --set full --clock-control none -k regex:tc_gemm_kernel -s 7 -c 1 -f -o $O/r${R}_bf_tc_gemm \
    python tools/bf_tc_check.py > $O/r${R}_ncu_bf.log 2>&1; echo "bf capture rc=$?"
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:i8_gemm_kernel -c 1 -f -o $O/r${R}_bf_i8_gemm \
    python tools/bf_i8_check.py 5 > $O/r${R}_ncu_bf_i8.log 2>&1; echo "bf_i8 capture rc=$?"
  timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv \
    --log-file $O/r${R}_launches_build.csv python tools/build_profile.py > $O/r${R}_ncu_build.log 2>&1; echo "build list rc=$?"
fi
