for i in 1 2; do timeout 600 python -m pytest tests/test_gpu_modules.py -m gpu -q -s -k "test_autograd_vs_oracle" 2>&1 | grep -E "max err" | sed 's/cancelling.*//' | cut -c1-200; done
