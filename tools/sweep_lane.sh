#!/bin/bash
# A/B of the deferred lane (GPU box): EZRT_DEFERRED_LANE=0 traces the accel policy's deferred rays in line between the accel
# kernel and k_shade (round 1), 1 (default) on a side stream beside k_shade.  C3 headline + C2 and C4 (parity: tests/test_gpu_configs.py runs with the default).
cd "$(dirname "$0")/.."
run() { echo "== $*"; env "$@" EZRT_AUTO_BUILD=0 python bench.py --steps 16 --warmup 3 --no-cpu-baseline --no-parity --no-e2e --extra-workloads c2,c4 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); w=d['workloads']
f=lambda x: '%.0f Mrays/s (%.3f ms/step; extend %.2f shade %.2f ms/step)' % (x['value'], x['ms_per_step'], x['kernel_ms']['extend']/x['steps'], x['kernel_ms']['shade']/x['steps'])
print('  c3', f(d)); print('  c2', f(w['c2'])); print('  c4', f(w['c4']))"; }
run EZRT_DEFERRED_LANE=0
run EZRT_DEFERRED_LANE=1
run EZRT_DEFERRED_LANE=0
run EZRT_DEFERRED_LANE=1
