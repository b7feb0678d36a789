set -x
export FFM_CLUSTER=4 FFM_THREADS=512
python profiles/quickbench_c3.py 33 200 | tail -1
ncu --set full --clock-control none --import-source on -k regex:ffm_cell_rollout -c 1 -o gpurun_out/prof_r2g_c3_cl4 -f python profiles/quickbench_c3.py 33 200 > gpurun_out/ncu_r2g.log 2>&1
tail -3 gpurun_out/ncu_r2g.log
