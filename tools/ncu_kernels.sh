#!/bin/bash
# `ncu --set full` of named kernels out of tools/kernels_bench.py:  bash tools/ncu_kernels.sh <tag> <kernel> [kernel ...]
set -u
tag=$1; shift
out=gpurun_out; mkdir -p $out
KB="python tools/kernels_bench.py"
$KB > $out/${tag}_kb_plain.log 2>&1 || { echo "kernels_bench failed"; tail -5 $out/${tag}_kb_plain.log; exit 1; }
head -12 $out/${tag}_kb_plain.log
for k in "$@"; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -f -o $out/${tag}_$k $KB > $out/${tag}_ncu_$k.log 2>&1
  echo "$k rc=$?"
done
