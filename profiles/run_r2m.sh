python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tests/multi_gpu_check.py 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -12
