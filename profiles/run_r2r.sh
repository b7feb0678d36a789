mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -6
python bench.py > gpurun_out/r2r_bench.json 2> gpurun_out/r2r_bench.err; tail -2 gpurun_out/r2r_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2r_bench.json'))
print('c2', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['bound'], d['roofline']['frac'], d['roofline']['traffic'], d['clocks'], d['cpu_baseline']['value'])
for k,v in d.get('secondary',{}).items():
    if 'error' in v: print(k, v); continue
    print(k, v['value'], v['unit'], v['ms_per_step'], 'e2e', v['e2e']['value'], v['clocks'].get('samples'), v.get('per_mode_ms'), v.get('sync_ms_blocking'), v['roofline'].get('frac'), v['roofline'].get('traffic'))
PY
python bench.py --impl reference --steps 2 --warmup 1 | cut -c1-300
python __graft_entry__.py --smoke
