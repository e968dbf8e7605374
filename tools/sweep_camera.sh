#!/bin/bash
# A/B on the GPU box: camera pass traced pixel-major (default: a warp = the samples of one pixel) vs frame-major (an 8x4 block of one frame),
# work chunks of the camera pass / the other passes, refill threshold of the camera pass
cd "$(dirname "$0")/.."
run() { echo "== $*"; env "$@" EZRT_AUTO_BUILD=0 python bench.py --workload c3 --steps 8 --warmup 3 --no-cpu-baseline --no-e2e --no-parity --extra-workloads "c2,c4" 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); w=d['workloads']; print('  c3 %.0f Mrays/s extend %.2f ms/step | c2 %.0f | c4 %.0f' % (d['value'], d['kernel_ms']['extend']/d['steps'], w['c2']['value'], w['c4']['value']))"; }
run X=0
run EZRT_CAMERA_ORDER=frame
run EZRT_CHUNK_CAM=128
run EZRT_CHUNK_CAM=32
run EZRT_CHUNK=64
run EZRT_REFILL_CAM=16
run EZRT_REFILL_CAM=28
run X=0
