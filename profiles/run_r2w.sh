for cfg in "FFM_CLUSTER=2 FFM_THREADS=1024" "FFM_CLUSTER=4 FFM_FIELDS_GLOBAL=1" "FFM_CLUSTER=8 FFM_FIELDS_GLOBAL=1" "FFM_CLUSTER=4 FFM_FIELDS_GLOBAL=1 FFM_THREADS=1024"; do
  env $cfg TAG="$cfg" timeout 300 python profiles/quickbench_c3.py 148 2>&1 | tail -1 | tee -a gpurun_out/r2w_c3_variants.jsonl | cut -c1-330
done
