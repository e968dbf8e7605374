#!/bin/bash
# compute-sanitizer over the wavefront queues: memcheck, racecheck, initcheck, synccheck
# on one small render per mode x policy x pipeline (tools/sanitize_case.py).  GPU box:  bash tools/sanitize.sh [WxH]
# Output: gpurun_out/sanitizer_<tool>.log (+ a one-line verdict each in gpurun_out/sanitizer_summary.txt)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
SZ=${1:-96x54}
CS=$(command -v compute-sanitizer || echo /usr/local/cuda/bin/compute-sanitizer)
: > gpurun_out/sanitizer_summary.txt
for form in 4 8; do          # both forms of the acceleration tree (EZRT_ACCEL): 4-wide exact boxes (default), W8
for tool in memcheck racecheck initcheck synccheck; do
    log=gpurun_out/sanitizer_${tool}_accel${form}.log
    extra=""
    [ "$tool" = memcheck ] && extra="--leak-check no"
    EZRT_ACCEL=$form EZRT_AUTO_BUILD=0 timeout 900 "$CS" --tool $tool $extra --print-limit 20 python tools/sanitize_case.py "$SZ" > "$log" 2>&1
    rc=$?
    echo "EZRT_ACCEL=$form $tool rc=$rc: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|sanitize_case:' "$log" | tr '\n' ' ')" >> gpurun_out/sanitizer_summary.txt
done
done
cat gpurun_out/sanitizer_summary.txt
