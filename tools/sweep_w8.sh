#!/bin/bash
# A/B runs of the accel kernels on the GPU box (short bench runs, headline only).  bash tools/sweep_w8.sh [workload]
cd "$(dirname "$0")/.."
WL=${1:-c3}
run() { echo "== $*"; env "$@" EZRT_AUTO_BUILD=0 python bench.py --workload $WL --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --no-parity --extra-workloads "" 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('  %.0f Mrays/s  %.2f ms/step  kernel_ms %s deferred %.5f' % (d['value'], d['ms_per_step'], {k: round(v/d['steps'],2) for k,v in d['kernel_ms'].items()}, d['deferred_ray_fraction']))"; }
run EZRT_ACCEL=4
run EZRT_ACCEL=4 EZRT_W4_COLLAPSE=greedy
run EZRT_ACCEL=8
