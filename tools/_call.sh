timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "forward_first_layer_folded" 2>&1 | tail -4
