python -m pytest tests -m gpu -q -x --durations=3 > gpurun_out/r02_pytest9.log 2>&1; tail -8 gpurun_out/r02_pytest8.log
python bench.py --steps 10 --warmup 3 --skip-cpu --skip-torch > gpurun_out/r02_bench6.json 2> gpurun_out/r02_bench6.err; tail -3 gpurun_out/r02_bench6.err
python - <<PY
import json
d=json.load(open("gpurun_out/r02_bench6.json"))
print({k:d[k] for k in ("value","ms_per_step","launches_per_step","step_tensor_frac")}, d["e2e"]["value"], d["roofline"]["frac"], d["counterfactual"]["value"])
PY
for F in audio_mnist whalecalls esrf_acoustic; do
  timeout 600 python bench.py --family $F --steps 5 --warmup 3 --skip-cpu --skip-torch > gpurun_out/r02_bench6_$F.json 2> gpurun_out/r02_bench6_$F.err; tail -2 gpurun_out/r02_bench6_$F.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02_bench6_$F.json"))
    print("$F", {k:d[k] for k in ("value","ms_per_step","launches_per_step","step_tensor_frac")}, d["counterfactual"]["value"])
    for r in json.load(open("gpurun_out/per_layer_$F.json"))[:10]: print("   ", r)
except Exception as e: print("$F failed", e)
PY
done
