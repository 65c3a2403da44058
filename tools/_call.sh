python -m pytest tests -m gpu -q -x --durations=3 > gpurun_out/r02_pytest10.log 2>&1; tail -5 gpurun_out/r02_pytest10.log
python bench.py --steps 10 --warmup 3 --skip-cpu --skip-torch > gpurun_out/r02_bench7.json 2> gpurun_out/r02_bench7.err; tail -3 gpurun_out/r02_bench7.err
python - <<PY
import json
d=json.load(open("gpurun_out/r02_bench7.json"))
print({k:d[k] for k in ("value","ms_per_step","launches_per_step","step_tensor_frac")}, d["e2e"]["value"], d["roofline"]["frac"], d["counterfactual"]["value"])
k=d["kernels_ms_per_step"]
for n,v in list(k.items())[:12]: print(n, v)
PY
