#!/bin/bash
# `ncu --set full` of named kernels out of tools/kernels_bench.py:  bash tools/ncu_kernels.sh <tag> <kernel[:skip[:label]]> ...
# (ncu matches the function name without template arguments; `skip` picks the launch: kernels_bench runs each
#  entry 3 + 20 times, so e.g. cpl_warp_kernel:4:fwd and cpl_warp_kernel:27:bwd)
set -u
tag=$1; shift
out=gpurun_out; mkdir -p $out
KB="python tools/kernels_bench.py"
$KB > $out/${tag}_kb_plain.log 2>&1 || { echo "kernels_bench failed"; tail -5 $out/${tag}_kb_plain.log; exit 1; }
head -12 $out/${tag}_kb_plain.log
for spec in "$@"; do
  IFS=: read -r k skip label <<< "$spec"
  skip=${skip:-4}; label=${label:-$k}
  ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -f -o $out/${tag}_$label $KB > $out/${tag}_ncu_$label.log 2>&1
  echo "$label rc=$?"
done
