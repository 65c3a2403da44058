python -m pytest tests -m gpu -q -x --durations=3 > gpurun_out/r02_pytest16.log 2>&1; tail -4 gpurun_out/r02_pytest16.log
python bench.py --steps 10 --warmup 3 --skip-cpu --skip-torch --skip-cf > gpurun_out/r02_bench14.json 2> gpurun_out/r02_bench14.err; tail -3 gpurun_out/r02_bench14.err
python - <<PY
import json
d=json.load(open("gpurun_out/r02_bench14.json"))
print({k:d[k] for k in ("value","ms_per_step","launches_per_step","step_tensor_frac")}, d["e2e"]["value"], d["roofline"]["frac"])
r=json.load(open("gpurun_out/per_layer_mnist.json"))
r.sort(key=lambda x:-x["ms"])
for x in r[:14]: print(x["layer"], x["n"], x["ms"], x["roofline_frac"])
for x in r:
    if "1x28x28" in x["layer"] or "32x28x28" in x["layer"]: print(x)
PY
