#!/bin/bash
# bounce-ray sort (octant | Morton cell counting sort, EZRT_SORT_RAYS=1) with the accel kernels -- GPU box
cd "$(dirname "$0")/.."
run() { echo "== $*"; env "$@" EZRT_AUTO_BUILD=0 python bench.py --workload c3 --steps 8 --warmup 3 --no-cpu-baseline --no-e2e --no-parity --extra-workloads "c4" 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); w=d['workloads']; print('  c3 %.0f Mrays/s kernel_ms/step %s | c4 %.0f' % (d['value'], {k: round(v/d['steps'],2) for k,v in d['kernel_ms'].items()}, w['c4']['value']))"; }
run X=0
run EZRT_SORT_RAYS=1
