python -m pytest tests -m gpu -q -x --durations=3 > gpurun_out/r02_pytest7.log 2>&1; tail -6 gpurun_out/r02_pytest7.log
python bench.py --steps 10 --warmup 3 --skip-cpu --skip-torch > gpurun_out/r02_bench4.json 2> gpurun_out/r02_bench4.err; tail -3 gpurun_out/r02_bench4.err
python - <<PY
import json
d=json.load(open("gpurun_out/r02_bench4.json"))
print({k:d[k] for k in ("value","ms_per_step","launches_per_step","step_tensor_frac")}, d["e2e"]["value"], d["roofline"]["frac"], d["counterfactual"]["value"])
k=d["kernels_ms_per_step"]
for n,v in k.items(): print(n, v)
PY
for F in whalecalls esrf_acoustic; do
  timeout 600 python bench.py --family $F --steps 5 --warmup 3 --skip-cpu --skip-torch > gpurun_out/r02_bench4_$F.json 2> gpurun_out/r02_bench4_$F.err; tail -2 gpurun_out/r02_bench4_$F.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02_bench4_$F.json"))
    print("$F", {k:d[k] for k in ("value","ms_per_step","launches_per_step","step_tensor_frac")})
    for r in json.load(open("gpurun_out/per_layer_$F.json"))[:8]: print("   ", r)
except Exception as e: print("$F failed", e)
PY
done
