export EZRT_AUTO_BUILD=0
for i in 1 2 3 4; do
EZRT_VERBOSE=$([ $i = 1 ] && echo 1 || echo 0) python bench.py --steps 16 --warmup 3 --no-cpu-baseline --no-parity --extra-workloads "c4" 2> gpurun_out/stab_$i.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); w=d['workloads']['c4']; print('c3 value %.0f e2e %.0f step_ms %s | c4 value %.0f e2e %.0f step_ms %s' % (d['value'], d['e2e']['value'], {k: round(v,2) for k,v in d['step_ms_rank0'].items()}, w['value'], w['e2e']['value'], {k: round(v,2) for k,v in w['step_ms_rank0'].items()}))"
done
grep "ezrt_scene_create" gpurun_out/stab_1.err | head -14
