timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "forward_first_layer_folded" 2>&1 | tail -4
for f in audio_mnist whalecalls; do
python bench.py --family $f --steps 5 --warmup 3 --skip-cpu --skip-torch > gpurun_out/r02_bench_$f.json 2> gpurun_out/r02_bench_$f.err; tail -1 gpurun_out/r02_bench_$f.err; cp gpurun_out/per_layer_$f.json gpurun_out/r02_per_layer_$f.json 2>/dev/null
python - <<PY
import json
d=json.load(open("gpurun_out/r02_bench_$f.json"))
print("$f", {k:d[k] for k in ("value","ms_per_step")}, d["roofline"]["frac"], d.get("counterfactual",{}).get("value"))
r=json.load(open("gpurun_out/r02_per_layer_$f.json"))
for x in r:
    if "win5" in x["layer"]: print(x)
PY
done
timeout 900 python -m pytest tests/test_gpu_modules.py -m gpu -q -x -k "audio or whale" 2>&1 | tail -2
