python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 tests/multi_gpu_check.py 2>&1 | grep "OK\|Error\|error" | head
for w in c5train; do python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 2 --workload $w --no-cpu 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('$w', d['n_gpus'], d['value'], d['ms_per_step'], d.get('tables'))"; done
python -m pytest tests/test_mcq_gpu.py -x -q -m gpu 2>&1 | tail -2
