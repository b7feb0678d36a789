set -x
mkdir -p gpurun_out
python bench.py --no-secondary --no-cpu > gpurun_out/r2p_bench_c2.json 2> gpurun_out/r2p_bench_c2.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2p_launches_c2.csv python bench.py --no-secondary --no-cpu --steps 2 > gpurun_out/r2p_ncu_c2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ffm_cell_rollout -s 3 -c 1 -o gpurun_out/prof_r2p_c2 -f python bench.py --no-secondary --no-cpu --steps 2 > gpurun_out/r2p_ncu_c2_full.log 2>&1
python bench.py --workload c3 --no-cpu --steps 3 > gpurun_out/r2p_bench_c3.json 2> gpurun_out/r2p_bench_c3.err
ncu --set full --clock-control none --import-source on -k regex:ffm_core_rollout -s 3 -c 1 -o gpurun_out/prof_r2p_c3 -f python bench.py --workload c3 --no-cpu --steps 2 > gpurun_out/r2p_ncu_c3_full.log 2>&1
FFM_CLUSTER=4 ncu --set full --clock-control none --import-source on -k regex:ffm_cell_rollout -s 1 -c 1 -o gpurun_out/prof_r2p_c3_cl4 -f python profiles/quickbench_c3.py 33 400 > gpurun_out/r2p_ncu_c3cl4_full.log 2>&1
ls -la gpurun_out/*.ncu-rep
