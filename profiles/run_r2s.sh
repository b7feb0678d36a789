mkdir -p gpurun_out
python -m pytest tests/test_core_parity_gpu.py tests/test_golden_gpu.py tests/test_fullsize_gpu.py tests/test_edge_cases_gpu.py tests/test_bench_contract.py -x -q -m gpu 2>&1 | tail -4
for w in c3 c2dff c1 c2; do python bench.py --workload $w --no-cpu --no-secondary --steps 3 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('$w', d['value'], d['ms_per_step'], d['clocks']['samples'], d['config']['kernel']['name'], d['config']['kernel']['threads'], d['config']['kernel']['ctas_per_sm'])"; done
FFM_KERNEL=ped python bench.py --workload c2dff --no-cpu --steps 3 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('c2dff ped', d['value'], d['ms_per_step'])"
FFM_KERNEL=cell python bench.py --workload c1 --no-cpu --steps 3 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('c1 cell', d['value'], d['ms_per_step'])"
env FFM_CLUSTER=4 TAG=cl4 python profiles/quickbench_c3.py 148 | tail -1 | cut -c1-120
