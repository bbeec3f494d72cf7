#!/bin/bash
# compute-sanitizer over the libafsl kernels at small sizes (run under gpurun, one GPU):
#     bash tools/sanitize.sh [memcheck|racecheck|initcheck|synccheck]      (default: memcheck)
# ONE tool per gpurun call (B200_PROFILING.md: several sanitizer tools in one call have wedged the GPU).  The target is the
# GPU parity suite restricted to the kernel-level tests (head, CPL, angular, SpecAugment, vote, view fusion, normalise) -
# every libafsl entry point, every kernel family, through the C ABI - at its small shapes.  (Round 2: the pool answered
# "compute-sanitizer is closed on this pool", rc 86 - see profiles/r2_sanitizer_note.txt; the script is kept for pools where it runs.)  The program must have exited 0
# without the sanitizer first (same call, `&&`).  Log -> gpurun_out/sanitize_<tool>.log; copy the summary to profiles/.
set -u
tool=${1:-memcheck}
out=gpurun_out
mkdir -p $out
SEL='head_vs_reference or cpl_vs_reference or angular_vs or specaug_vs_reference or vote_vs_reference or view_fusion_vs_reference or l2_normalize or head_edge_cases or head_ragged or many_way_tensor'
CMD=(python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "$SEL" -p no:cacheprovider)
"${CMD[@]}" > $out/sanitize_plain.log 2>&1 || { echo "plain run failed"; tail -5 $out/sanitize_plain.log; exit 1; }
timeout 1500 compute-sanitizer --tool "$tool" --target-processes all --error-exitcode 99 --log-file $out/sanitize_${tool}.log \
    "${CMD[@]}" > $out/sanitize_${tool}_pytest.log 2>&1
rc=$?
echo "compute-sanitizer --tool $tool rc=$rc"
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Error|hazard" $out/sanitize_${tool}.log | sort | uniq -c | head -20
tail -3 $out/sanitize_${tool}_pytest.log
