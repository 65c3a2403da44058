export SKIP_LIST=1
export KERNELS="conv_cm_kernel<.bool.1,..bool.0,..int.1> conv_cm_kernel<.bool.1,..bool.1 conv_cm_kernel<.bool.1,..bool.0,..int.2>"
bash tools/profile_round.sh r02i
CMD="python bench.py --steps 2 --warmup 3 --skip-cf --skip-cpu --skip-torch --no-graph"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 2000 -c 1400 --csv \
    --log-file gpurun_out/launches_dram_r02i.csv $CMD > gpurun_out/ncu_list_dram_i.log 2>&1
echo "rc=$?"; wc -l gpurun_out/launches_dram_r02i.csv
python -m pytest tests -m gpu -q -x 2>&1 | tail -2
