#!/bin/bash
# e2e (host-buffer) A/B on the GPU box: lastFrame upload beside the kernels (default) vs on the render stream
cd "$(dirname "$0")/.."
run() { echo "== $*"; env "$@" EZRT_AUTO_BUILD=0 python bench.py --workload c3 --steps 10 --warmup 3 --no-cpu-baseline --no-parity --extra-workloads "c4" 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); w=d['workloads']['c4']; print('  c3 value %.0f e2e %.0f (%.2f ms/step) | c4 value %.0f e2e %.0f' % (d['value'], d['e2e']['value'], d['e2e']['ms_per_step'], w['value'], w['e2e']['value']))"; }
run EZRT_RENDER_OVERLAP=1
run EZRT_RENDER_OVERLAP=0
run EZRT_RENDER_OVERLAP=1
run EZRT_RENDER_OVERLAP=0
run EZRT_REFILL_CAM=1 EZRT_CHUNK=32
run EZRT_REFILL_CAM=8
run EZRT_REFILL_CAM=16
run EZRT_CHUNK=64 EZRT_LEAF_T=16 EZRT_REFILL_T=20
run EZRT_CHUNK=128
