#!/bin/bash
# ncu evidence for one round (run under gpurun, one GPU):  bash tools/ncu_capture.sh <tag> [kernel ...]
#   1. launch list of the bench command (gpu__time_duration per launch; cold-cache, serialised -> compare SHARES)
#   2. one `--set full` capture per named kernel out of tools/kernels_bench.py
# Each ncu pass runs only after the same command exited 0 without ncu (B200_PROFILING.md).
set -u
tag=${1:-r1}; shift || true
kernels=("$@")
[ ${#kernels[@]} -eq 0 ] && kernels=(head_warp_kernel:4:head_fwd head_warp_kernel:27:head_bwd cpl_warp_kernel:4:cpl_fwd cpl_warp_kernel:27:cpl_bwd specaug_tile_kernel:4:specaug_tile angular_warp_kernel:2:angular_fwd angular_warp_kernel:10:angular_bwd head_tma_fwd_kernel:50:head_tma_20w5s_d256 head_tma_fwd_kernel:4:head_tma_20w5s_d64 fusion_fwd_kernel:27:fusion_fwd_n409600 fusion_bwd_kernel:27:fusion_bwd_n409600 vote_kernel:4:eval_vote)
out=gpurun_out
mkdir -p $out
BENCH="python bench.py --steps 2 --warmup 3 --episodes 32 --skip-cpu --skip-kernels --skip-eval --skip-tf32"
$BENCH > $out/${tag}_bench_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --graph-profiling node -c 8000 --csv --log-file $out/${tag}_launches.csv \
    $BENCH > $out/${tag}_launches.log 2>&1
echo "launch list rc=$?"
KB="python tools/kernels_bench.py"
$KB > $out/${tag}_kb_plain.log 2>&1 || { echo "kernels_bench failed"; exit 1; }
for spec in "${kernels[@]}"; do
  IFS=: read -r k skip label <<< "$spec"
  skip=${skip:-4}; label=${label:-$k}
  ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -f -o $out/${tag}_$label $KB > $out/${tag}_ncu_$label.log 2>&1
  echo "$label rc=$?"
done
# fused encoder stage 1 at the training shape
S1="python tools/stage1_bench.py"
$S1 > $out/${tag}_s1_plain.log 2>&1 && for k in stage1_fwd_nhwc_kernel stage1_bwd_nhwc_kernel stage1_moments_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o $out/${tag}_$k $S1 > $out/${tag}_ncu_$k.log 2>&1
  echo "$k rc=$?"
done
