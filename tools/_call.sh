CMD="python bench.py --steps 2 --warmup 3 --skip-cf --skip-cpu --skip-torch --no-graph"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 2000 -c 1400 --csv \
    --log-file gpurun_out/launches_dram_r02.csv $CMD > gpurun_out/ncu_list_dram.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/ncu_list_dram.log; wc -l gpurun_out/launches_dram_r02.csv
python -m pytest tests/test_gpu_ops.py -m gpu -q -k "taps or bn_fold" 2>&1 | tail -3
