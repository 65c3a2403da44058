python bench.py --steps 20 --warmup 5 --skip-cpu --skip-torch --skip-cf > gpurun_out/r02_bench18.json 2> gpurun_out/r02_bench18.err; tail -1 gpurun_out/r02_bench18.err
cp gpurun_out/per_layer_mnist.json gpurun_out/r02_per_layer_mnist_clean.json
python - <<PY
import json
d=json.load(open("gpurun_out/r02_bench18.json"))
print({k:d[k] for k in ("value","ms_per_step")}, d["roofline"]["frac"], d["roofline"]["layer_rows_at_or_above_half_roofline"])
r=json.load(open("gpurun_out/r02_per_layer_mnist_clean.json"))
r.sort(key=lambda x:-x["ms"])
for x in r[:16]: print(x["layer"], x["n"], x["ms"], x["roofline_frac"])
PY
