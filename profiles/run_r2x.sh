export FFM_B200_LIB=$PWD/ffm_b200/libffm_b200_timing.so
for cfg in "FFM_CLUSTER=2 FFM_THREADS=1024" "FFM_CLUSTER=2 FFM_THREADS=512"; do
  echo "== $cfg"; env $cfg timeout 300 python profiles/quickbench_c3.py 74 2048 2>&1 | grep "phase cycles" | tail -1
done
