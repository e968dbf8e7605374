#!/bin/bash
# A/B of k_shade build variants (GPU box): shade milliseconds per step on C3 and C4
cd "$(dirname "$0")/.."
run() { echo "== $*"; env "$@" EZRT_AUTO_BUILD=0 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e --no-parity --extra-workloads c4 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); w=d['workloads']['c4']
print('  c3 %.0f Mrays/s shade %.2f ms/step | c4 %.0f Mrays/s shade %.2f ms/step' % (d['value'], d['kernel_ms']['shade']/d['steps'], w['value'], w['kernel_ms']['shade']/w['steps']))"; }
run X=0
run EZRT_LIB_VARIANT=rg
run X=0
run EZRT_LIB_VARIANT=rg
