#!/bin/bash
# A/B on the GPU box: software prefetch of the next node record (pf1) and of the first triangles of the next leaf (pf2) in the 4-wide accel kernel
cd "$(dirname "$0")/.."
run() { echo "== $*"; env "$@" EZRT_AUTO_BUILD=0 python bench.py --workload c3 --steps 8 --warmup 3 --no-cpu-baseline --no-e2e --no-parity --extra-workloads "c4" 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); w=d['workloads']['c4']; print('  c3 %.0f Mrays/s extend %.2f shade %.2f ms/step | c4 %.0f extend %.2f shadow %.2f' % (d['value'], d['kernel_ms']['extend']/d['steps'], d['kernel_ms']['shade']/d['steps'], w['value'], w['kernel_ms']['extend']/w['steps'], w['kernel_ms']['shadow']/w['steps']))"; }
run X=0
run EZRT_LIB_VARIANT=pf1
run EZRT_LIB_VARIANT=pf2
run X=0
