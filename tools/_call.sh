(time python bench.py) > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; tail -4 gpurun_out/r02_bench_final.err
(time python bench.py --impl reference --steps 3 --warmup 1) > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err; tail -4 gpurun_out/r02_bench_ref.err; cut -c1-400 gpurun_out/r02_bench_ref.json
for f in audio_mnist whalecalls esrf_acoustic; do
python bench.py --family $f --steps 5 --warmup 3 --skip-cpu --skip-torch > gpurun_out/r02_bench_$f.json 2> gpurun_out/r02_bench_$f.err; tail -1 gpurun_out/r02_bench_$f.err; cp gpurun_out/per_layer_$f.json gpurun_out/r02_per_layer_$f.json 2>/dev/null
python - <<PY
import json
d=json.load(open("gpurun_out/r02_bench_$f.json"))
print("$f", {k:d[k] for k in ("value","ms_per_step")}, d["roofline"]["frac"], d.get("counterfactual",{}).get("value"))
PY
done
