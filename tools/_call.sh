timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "wgrad" 2>&1 | tail -2
timeout 300 python tools/layer_bench.py --family mnist --batch 4096 --only "Dx.dx.1" --passes wgrad 2>&1 | grep "dx.1 "
timeout 300 python tools/layer_bench.py --family mnist --batch 4096 --only "G.layers.8" --passes wgrad 2>&1 | grep "layers.8 "
