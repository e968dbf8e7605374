#!/bin/bash
# strong scaling on ONE fixed 3840x2160 image (BASELINE configs[4]) -- run on an 8-GPU box:  bash tools/run_round2_scaling.sh "8 4 2"
cd "$(dirname "$0")/.."
export EZRT_AUTO_BUILD=0
python bench.py --image 3840x2160 --steps 10 --warmup 3 --no-cpu-baseline --no-parity --extra-workloads "" > gpurun_out/bench_r2_n1_4k.json 2> gpurun_out/bench_r2_n1_4k.err
python -c "
import json; d=json.load(open('gpurun_out/bench_r2_n1_4k.json')); print('N=1 3840x2160: value %.0f e2e %.0f' % (d['value'], d['e2e']['value']))"
PORT=29520
for N in ${1:-8 4 2}; do
  PORT=$((PORT+1))
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_r2_n$N.json 2> gpurun_out/bench_r2_n$N.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_r2_n$N.json'))
print('N=$N image %s: value %.0f e2e %.0f parity differing %s linf %s rank max/mean %.4f step_ms %s cpu %.2f Mrays/s on %d threads' % (d['config']['image'], d['value'], d['e2e']['value'], d['parity']['differing'], d['parity']['linf'], d['rank_rays']['max_over_mean'], d['step_ms_rank0'], d['cpu_baseline']['value'], d['cpu_baseline']['threads_used']))"
done
