timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "wgrad" 2>&1 | tail -2
python -m pytest tests -m gpu -q -x 2>&1 | tail -2
python bench.py --steps 10 --warmup 3 --skip-cpu --skip-torch --skip-cf > gpurun_out/r02_bench17.json 2> gpurun_out/r02_bench17.err; tail -3 gpurun_out/r02_bench17.err
python - <<PY
import json
d=json.load(open("gpurun_out/r02_bench17.json"))
print({k:d[k] for k in ("value","ms_per_step")}, d["e2e"]["value"], d["roofline"]["frac"])
k=d["kernels_ms_per_step"]
for n,v in list(k.items())[:5]: print(n, v)
r=json.load(open("gpurun_out/per_layer_mnist.json"))
for x in r:
    if "wgrad" in x["layer"] or "32x24x24->5" in x["layer"]: print(x["layer"], x["n"], x["ms"], x["roofline_frac"])
PY
