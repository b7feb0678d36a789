mkdir -p gpurun_out
python -m pytest tests/test_core_parity_gpu.py tests/test_golden_gpu.py tests/test_fullsize_gpu.py tests/test_edge_cases_gpu.py tests/test_distribution_gpu.py -x -q -m gpu 2>&1 | tail -4
python bench.py --workload c3 --no-cpu > gpurun_out/r2y_bench_c3.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r2y_bench_c3.json')); print('c3', d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['kernel'], d['roofline']['frac'])"
ncu --set full --clock-control none -k regex:ffm_cell_rollout -s 3 -c 1 -o gpurun_out/prof_r2y_c3_cl2 -f python bench.py --workload c3 --no-cpu --steps 2 > gpurun_out/r2y_ncu_c3.log 2>&1; tail -1 gpurun_out/r2y_ncu_c3.log
