python -m pytest tests/test_gpu_next.py -m gpu -q -x 2>&1 | tail -30 > gpurun_out/r02_pytest13.log; tail -30 gpurun_out/r02_pytest13.log
python -m pytest tests/test_gpu_ops.py tests/test_gpu_modules.py -m gpu -q -x -k "feat or forward or mnist" 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 --skip-cpu --skip-torch --skip-cf > gpurun_out/r02_bench10.json 2> gpurun_out/r02_bench10.err; tail -3 gpurun_out/r02_bench10.err
python - <<PY
import json
d=json.load(open("gpurun_out/r02_bench10.json"))
print({k:d[k] for k in ("value","ms_per_step","launches_per_step","step_tensor_frac")}, d["e2e"]["value"], d["roofline"]["frac"])
k=d["kernels_ms_per_step"]
for n,v in list(k.items())[:8]: print(n, v)
PY
