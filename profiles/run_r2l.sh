python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/r2l_bench_n2.json 2> gpurun_out/r2l_bench_n2.err; tail -3 gpurun_out/r2l_bench_n2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2l_bench_n2.json'))
print('c2', d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'])
for k,v in d.get('secondary',{}).items():
    if 'error' in v: print(k, v); continue
    print(k, v['value'], v['unit'], v['ms_per_step'], 'e2e', v['e2e']['value'], v['clocks'].get('samples'), v.get('sync_ms_blocking'), v['config'].get('parallelism'))
PY
python tests/multi_gpu_check.py 2>&1 | tail -3
