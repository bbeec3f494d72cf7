#!/bin/bash
# ncu evidence for one round (run under gpurun, one GPU):  bash tools/ncu_capture.sh <tag> [kernel ...]
#   1. launch list of the bench command (gpu__time_duration per launch; cold-cache, serialised -> compare SHARES)
#   2. one `--set full` capture per named kernel out of tools/kernels_bench.py
# Each ncu pass runs only after the same command exited 0 without ncu (B200_PROFILING.md).
set -u
tag=${1:-r1}; shift || true
kernels=("$@")
[ ${#kernels[@]} -eq 0 ] && kernels=(head_fwd_kernel head_bwd_kernel cpl_fwd_kernel cpl_bwd_kernel specaug_kernel)
out=gpurun_out
mkdir -p $out
BENCH="python bench.py --steps 2 --warmup 3 --episodes 32 --skip-cpu --skip-kernels"
$BENCH > $out/${tag}_bench_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $out/${tag}_launches.csv \
    $BENCH > $out/${tag}_launches.log 2>&1
echo "launch list rc=$?"
KB="python tools/kernels_bench.py"
$KB > $out/${tag}_kb_plain.log 2>&1 || { echo "kernels_bench failed"; exit 1; }
for k in "${kernels[@]}"; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -f -o $out/${tag}_$k $KB > $out/${tag}_ncu_$k.log 2>&1
  echo "$k rc=$?"
done
