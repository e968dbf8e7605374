#!/bin/bash
# A/B sweep of the W8 traversal tunables on the GPU box (short bench runs, C3 headline only).  bash tools/sweep_w8.sh [workload]
cd "$(dirname "$0")/.."
WL=${1:-c3}
run() { echo "== $*"; env "$@" EZRT_AUTO_BUILD=0 python bench.py --workload $WL --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --no-parity --extra-workloads "" 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('  %.0f Mrays/s  %.2f ms/step  kernel_ms %s deferred %.5f' % (d['value'], d['ms_per_step'], {k: round(v/d['steps'],2) for k,v in d['kernel_ms'].items()}, d['deferred_ray_fraction']))"; }
run EZRT_TRI_W=1
run EZRT_TRI_W=2
run EZRT_TRI_W=3
run EZRT_TRI_W=4
run EZRT_TRI_W=2 EZRT_REFILL_T=16
run EZRT_TRI_W=2 EZRT_REFILL_T=28
run EZRT_TRI_W=2 EZRT_CHUNK=64
run EZRT_ACCEL=4
