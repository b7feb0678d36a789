for cfg in "FFM_CLUSTER=4 FFM_THREADS=512" "FFM_CLUSTER=8 FFM_THREADS=512"; do
  echo "== $cfg"; env $cfg FFM_B200_LIB=$PWD/ffm_b200/libffm_b200_timing.so timeout 300 python profiles/quickbench_c3.py 33 400 2>&1 | tail -2 | head -1
  env $cfg TAG="$cfg" timeout 300 python profiles/quickbench_c3.py 148 2>&1 | tail -1 | tee -a gpurun_out/r2j_c3_variants.jsonl
done
python -m pytest tests/test_fullsize_gpu.py tests/test_distribution_gpu.py -x -q -m gpu 2>&1 | tail -5
