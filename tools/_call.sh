for f in whalecalls audio_mnist; do
python bench.py --family $f --steps 5 --warmup 3 --skip-cpu --skip-torch > gpurun_out/r02_bench_$f.json 2> gpurun_out/r02_bench_$f.err; tail -1 gpurun_out/r02_bench_$f.err; cp gpurun_out/per_layer_$f.json gpurun_out/r02_per_layer_$f.json 2>/dev/null
python - <<PY
import json
d=json.load(open("gpurun_out/r02_bench_$f.json"))
print("$f", {k:d[k] for k in ("value","ms_per_step")}, d["step_tensor_frac"], d["roofline"]["frac"], d.get("counterfactual",{}).get("value"), d.get("counterfactual",{}).get("roofline",{}).get("frac"))
PY
done
python -m pytest tests/test_gpu_modules.py -m gpu -q -x 2>&1 | tail -2
