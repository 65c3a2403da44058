#!/bin/bash
# Profiling recipe of a round (run on the GPU box through gpurun; one GPU).  Writes into gpurun_out/:
#   plain.log                 the bench command without a profiler (must exit 0 first)
#   launches_<tag>.csv        ncu launch list (gpu__time_duration.sum, no clock control) of the same command
#   prof_<tag>_<k>.ncu-rep    one `--set full` capture per kernel family, a steady-state launch each
# usage: tools/profile_round.sh <tag>
set -u
TAG=${1:-r01}
CMD="python bench.py --steps 2 --warmup 3 --skip-cf --skip-cpu --skip-torch --no-graph"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain.log; exit 1; }
if [ -z "${SKIP_LIST:-}" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 1400 --csv \
    --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu_list.log 2>&1
  echo "launch list rc=$?"
fi
i=0
# patterns are regular expressions over the demangled name, e.g. conv_ws_kernel<(int)32, (int)3>(...): '.' stands for
# the parentheses and blanks
for K in ${KERNELS:-conv_ws_kernel<.int.32,..int.3> wgrad_tc_kernel<.int.64> conv_tc_kernel<.int.256 act_backward_v8<.int.4>}; do
  i=$((i + 1))
  timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$K" -s 6 -c 1 -f \
    -o gpurun_out/prof_${TAG}_$i $CMD > gpurun_out/ncu_full_$i.log 2>&1
  echo "capture $i ($K) rc=$? $(ls -la gpurun_out/prof_${TAG}_$i.ncu-rep 2>/dev/null | awk '{print $5}') bytes"
done
