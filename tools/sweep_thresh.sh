#!/bin/bash
# sweep of the persistent-traversal vote thresholds of the 4-wide accel kernel (GPU box, C3 headline only)
cd "$(dirname "$0")/.."
run() { echo "== $*"; env "$@" EZRT_AUTO_BUILD=0 python bench.py --workload c3 --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --no-parity --extra-workloads "" 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('  %.0f Mrays/s  %.2f ms/step  extend %.2f' % (d['value'], d['ms_per_step'], d['kernel_ms']['extend']/d['steps']))"; }
run X=0
run EZRT_LEAF_T=6
run EZRT_LEAF_T=9
run EZRT_LEAF_T=16
run EZRT_INNER_T=12
run EZRT_INNER_T=20
run EZRT_REFILL_T=20
run EZRT_REFILL_T=28
run EZRT_LEAF_T=9 EZRT_INNER_T=20
run EZRT_CHUNK=64
