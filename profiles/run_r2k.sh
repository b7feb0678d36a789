set -x
python -m pytest tests/test_sff_gpu.py tests/test_bench_contract.py -x -q -m gpu 2>&1 | tail -12
python bench.py > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err; tail -3 gpurun_out/r2k_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2k_bench.json'))
print('c2', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['bound'], d['roofline']['frac'], d['roofline']['peak'])
for k,v in d.get('secondary',{}).items():
    if 'error' in v: print(k, v); continue
    print(k, v['value'], v['unit'], v['ms_per_step'], 'e2e', v['e2e']['value'], v['clocks'].get('samples'), v.get('per_mode_ms'), v.get('sync_ms_blocking'), v.get('record'))
PY
