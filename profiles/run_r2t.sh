mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:ffm_unified_rollout -s 20 -c 1 -o gpurun_out/prof_r2t_c4 -f python bench.py --workload c4 --no-cpu --steps 1 > gpurun_out/r2t_ncu_c4.log 2>&1
tail -2 gpurun_out/r2t_ncu_c4.log
