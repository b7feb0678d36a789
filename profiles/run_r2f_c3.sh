set -x
python -m pytest tests/test_core_parity_gpu.py tests/test_golden_gpu.py tests/test_edge_cases_gpu.py -x -q -m gpu 2>&1 | tail -15
python -m pytest tests/test_fullsize_gpu.py -x -q -m gpu -k c3 2>&1 | tail -15
for cfg in "FFM_KERNEL=ped" "FFM_CLUSTER=4 FFM_THREADS=512" "FFM_CLUSTER=4 FFM_THREADS=1024" "FFM_CLUSTER=8 FFM_THREADS=512" "FFM_CLUSTER=8 FFM_THREADS=1024" "FFM_CLUSTER=8 FFM_THREADS=1024 FFM_SCORE_GLOBAL=1"; do
  env $cfg TAG="$cfg" timeout 300 python profiles/quickbench_c3.py 148 2>&1 | tail -1 | tee -a gpurun_out/r2f_c3_variants.jsonl
done
