for i in 1 2 3; do timeout 600 python -m pytest tests/test_gpu_next.py -m gpu -q -x -k "explain or stream" 2>&1 | tail -2; done
