export FFM_B200_LIB=$PWD/ffm_b200/libffm_b200_timing.so
for cfg in "FFM_CLUSTER=4 FFM_THREADS=512" "FFM_CLUSTER=4 FFM_THREADS=1024" "FFM_CLUSTER=8 FFM_THREADS=512"; do
  echo "== $cfg"; env $cfg timeout 300 python profiles/quickbench_c3.py 33 400 2>&1 | tail -2
done
echo "== c2dff"; python bench.py --workload c2dff --no-cpu --steps 1 --episodes 1184 2>&1 | grep "phase cycles" | tail -1
echo "== c2"; python bench.py --workload c2 --no-cpu --steps 1 --episodes 1776 2>&1 | grep "phase cycles" | tail -1
