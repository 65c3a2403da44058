timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "forward_first_layer_folded" 2>&1 | tail -3
for v in "ICF_CM_DBG=0" "ICF_CM_DBG=2"; do
echo "--- $v"; env $v timeout 300 python tools/layer_bench.py --family mnist --batch 4096 --only "Dx.dx.1" --passes fwd 2>&1 | grep "dx.1 "
done
