python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; tail -2 gpurun_out/r02_bench_final.err
python - <<PY
import json
d=json.load(open("gpurun_out/r02_bench_final.json"))
print({k:d[k] for k in ("value","ms_per_step")}, d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["layer_rows_at_or_above_half_roofline"], d["counterfactual"]["value"], d["counterfactual"]["e2e"]["value"], d["cudnn_baseline"]["bf16_cl"]["value"], d["cpu_baseline"]["value"])
PY
export SKIP_LIST=1
export KERNELS="conv_cm_kernel<.bool.1,..bool.0,..int.1> conv_sx_kernel"
bash tools/profile_round.sh r02j
CMD="python bench.py --steps 2 --warmup 3 --skip-cf --skip-cpu --skip-torch --no-graph"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 2000 -c 1400 --csv \
    --log-file gpurun_out/launches_dram_r02j.csv $CMD > gpurun_out/ncu_list_dram_j.log 2>&1
echo "rc=$?"
