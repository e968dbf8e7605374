export EZRT_AUTO_BUILD=0
python -m pytest tests -m gpu -x -q 2>&1 | tail -6
EZRT_ACCEL=8 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/bench_r2_n1.json 2> gpurun_out/bench_r2_n1.err; tail -3 gpurun_out/bench_r2_n1.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r2_n1.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'parity',d['parity'],'\nroofline',json.dumps(d['roofline'])[:1200],'\ncpu',d['cpu_baseline'],'\nkernel_ms',d['kernel_ms'])
for k,v in d.get('workloads',{}).items(): print(k, v['value'], v['e2e']['value'], v['parity'], v['kernel_ms'])
"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r2_ref.json 2>/dev/null; cat gpurun_out/bench_r2_ref.json | cut -c1-600
