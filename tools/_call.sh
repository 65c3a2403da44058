timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "first_layer_folded" 2>&1 | tail -5
timeout 300 python tools/layer_bench.py --family mnist --batch 4096 --only "E.layers.0" --passes fwd 2>&1 | grep "layers.0 "
ICF_CM=0 timeout 300 python tools/layer_bench.py --family mnist --batch 4096 --only "E.layers.0" --passes fwd 2>&1 | grep "layers.0 "
python bench.py --steps 10 --warmup 3 --skip-cpu --skip-torch > gpurun_out/r02_bench15.json 2> gpurun_out/r02_bench15.err; tail -3 gpurun_out/r02_bench15.err
python - <<PY
import json
d=json.load(open("gpurun_out/r02_bench15.json"))
print({k:d[k] for k in ("value","ms_per_step")}, d["e2e"]["value"], d["roofline"]["frac"], d["counterfactual"]["value"], d["counterfactual"]["e2e"]["value"])
k=d["kernels_ms_per_step"]
for n,v in list(k.items())[:12]: print(n, v)
PY
