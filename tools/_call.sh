timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "first_layer_folded" 2>&1 | tail -2
for v in "ICF_CM_DBG=0" "ICF_CM_DBG=1"; do
echo "--- $v"; env $v timeout 300 python tools/layer_bench.py --family mnist --batch 4096 --only "Dx.dx.1" --passes fwd 2>&1 | grep "dx.1 "
done
timeout 300 python tools/layer_bench.py --family mnist --batch 4096 --only "E.layers.0" --passes fwd 2>&1 | grep "layers.0 "
timeout 300 python tools/layer_bench.py --family mnist --batch 4096 --only "G.layers.8" --passes dgrad 2>&1 | grep "layers.8 "
