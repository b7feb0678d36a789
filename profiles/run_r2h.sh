set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -8
for cfg in "FFM_KERNEL=ped" "FFM_CLUSTER=4 FFM_THREADS=512" "FFM_CLUSTER=4 FFM_THREADS=1024" "FFM_CLUSTER=8 FFM_THREADS=512"; do
  env $cfg TAG="$cfg" timeout 300 python profiles/quickbench_c3.py 148 2>&1 | tail -1 | tee -a gpurun_out/r2h_c3_variants.jsonl
done
for w in c2dff c1 c2; do python bench.py --workload $w --no-cpu --steps 3 > gpurun_out/r2h_$w.json 2> gpurun_out/r2h_$w.err; python -c "
import json,sys; d=json.load(open('gpurun_out/r2h_$w.json')); print('$w', d['value'], d['ms_per_step'], d['config']['kernel'])"; done
for w in c2dff c1; do FFM_KERNEL=ped python bench.py --workload $w --no-cpu --steps 3 > gpurun_out/r2h_${w}_ped.json 2> gpurun_out/r2h_${w}_ped.err; python -c "
import json,sys; d=json.load(open('gpurun_out/r2h_${w}_ped.json')); print('$w ped', d['value'], d['ms_per_step'], d['config']['kernel'])"; done
