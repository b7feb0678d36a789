python -m pytest tests/test_unified_gpu.py tests/test_golden_gpu.py tests/test_dropin_gpu.py tests/test_batched_learning_gpu.py tests/test_unified_training_gpu.py -x -q -m gpu 2>&1 | tail -3
for w in c4 c1 c3 c2dff; do python bench.py --workload $w --no-cpu --steps 3 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('$w', d['value'], d['ms_per_step'])"; done
