timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -x 2>&1 | tail -2
timeout 300 python tools/layer_bench.py --family mnist --batch 4096 --only "Dx.dx.9" --passes dgrad 2>&1 | grep "dx.9 "
ICF_TC_DUAL=0 timeout 300 python tools/layer_bench.py --family mnist --batch 4096 --only "Dx.dx.9" --passes dgrad 2>&1 | grep "dx.9 "
python -m pytest tests -m gpu -q -x 2>&1 | tail -2
python bench.py --steps 20 --warmup 5 --skip-cpu --skip-torch --skip-cf > gpurun_out/r02_bench19.json 2> gpurun_out/r02_bench19.err; tail -1 gpurun_out/r02_bench19.err
python - <<PY
import json
d=json.load(open("gpurun_out/r02_bench19.json"))
print({k:d[k] for k in ("value","ms_per_step")}, d["roofline"]["frac"], d["roofline"]["layer_rows_at_or_above_half_roofline"])
PY
