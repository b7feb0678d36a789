mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 > gpurun_out/r2q_bench_n8.json 2> gpurun_out/r2q_bench_n8.err; tail -2 gpurun_out/r2q_bench_n8.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2q_bench_n8.json'))
print('c2', d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])
for k,v in d.get('secondary',{}).items():
    if 'error' in v: print(k, v); continue
    print(k, v['value'], v['unit'], v['ms_per_step'], 'e2e', v['e2e']['value'], v['clocks'].get('samples'), v.get('sync_ms_blocking'))
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514 tests/multi_gpu_check.py 2>&1 | grep "OK" | tee gpurun_out/r2q_multi_gpu_check_n8.txt
