#!/bin/bash
# ncu --set full captures of one timed step of the bench command (GPU box).  One capture per workload; the raw metric page is
# exported on the box so that tools/ncu_summaries.py can turn it into profiles/ncu_<workload>_r2_summary.md and
# profiles/ncu_dram_r2.json here.     bash tools/ncu_capture.sh [workloads...]      (default: c3 c4)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export EZRT_AUTO_BUILD=0
for WL in ${@:-c3 c4}; do
    SKIP=9; [ "$WL" = c4 ] && SKIP=13         # accel / shade / deferred-lane shade (/ shadow / nee) launches of the warm-up step
    CMD="python bench.py --workload $WL --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-parity --extra-workloads ''"
    ncu --set full --clock-control none --import-source on -k regex:"k_extend_accel|k_extend_w8|k_shade|k_shadow_accel|k_shadow_w8|k_nee" -s $SKIP -c $SKIP \
        -f -o gpurun_out/prof_${WL}_r2 python bench.py --workload $WL --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-parity --extra-workloads "" > gpurun_out/ncu_${WL}_r2.log 2>&1
    ncu -i gpurun_out/prof_${WL}_r2.ncu-rep --page raw --csv > gpurun_out/prof_${WL}_r2_raw.csv 2>/dev/null
    [ "${KEEP_REP:-0}" = 1 ] || rm -f gpurun_out/prof_${WL}_r2.ncu-rep    # gpurun copies at most 64 MiB back: the csv export is what the summaries read
    echo "$WL: $(grep -c . gpurun_out/prof_${WL}_r2_raw.csv) csv lines; $CMD"
    # every launch of one warm-up + one timed step with its device time (shares of a step)
    ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_${WL}_r2.csv \
        python bench.py --workload $WL --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-parity --extra-workloads "" > /dev/null 2>&1
done
