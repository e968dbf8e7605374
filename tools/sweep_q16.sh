#!/bin/bash
# A/B on the GPU box: 96-byte quantised nodes for the bounce / shadow launches (default) vs the 128-byte exact nodes everywhere (EZRT_ACCEL_Q16=0)
cd "$(dirname "$0")/.."
run() { echo "== $*"; env "$@" EZRT_AUTO_BUILD=0 python bench.py --workload c3 --steps 8 --warmup 3 --cpu-reps 1 --no-e2e --extra-workloads "c2,c4" 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); w=d['workloads']; print('  c3 %.0f Mrays/s extend %.2f ms/step parity %s | c2 %.0f %s | c4 %.0f extend %.2f shadow %.2f %s' % (d['value'], d['kernel_ms']['extend']/d['steps'], d['parity'] and d['parity']['differing'], w['c2']['value'], w['c2']['parity'] and w['c2']['parity']['differing'], w['c4']['value'], w['c4']['kernel_ms']['extend']/w['c4']['steps'], w['c4']['kernel_ms']['shadow']/w['c4']['steps'], w['c4']['parity'] and w['c4']['parity']['differing']))"; }
run X=0
run EZRT_ACCEL_Q16=0
run X=0
run EZRT_ACCEL_Q16=0
