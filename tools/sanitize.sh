#!/bin/bash
# compute-sanitizer passes over the hot path (run on the GPU box through gpurun; one GPU).  Logs -> gpurun_out/sanitizer_<tool>.log
# (copied to profiles/ once read).  Each pass is bounded by its own timeout: the instrumented tcgen05/TMA kernels run 10-100x slower.
# usage: tools/sanitize.sh [tools...]   default: memcheck racecheck synccheck initcheck
set -u
TOOLS=${*:-memcheck racecheck synccheck initcheck}
SMOKE='import __graft_entry__ as g; g.smoke()'
for T in $TOOLS; do
  EXTRA=""
  [ "$T" = initcheck ] && EXTRA="--track-unused-memory no"
  timeout ${SAN_TIMEOUT:-420} compute-sanitizer --tool $T $EXTRA --print-limit 40 --error-exitcode 3 \
      --log-file gpurun_out/sanitizer_${T}_smoke.log python -c "$SMOKE" > gpurun_out/sanitizer_${T}_smoke.out 2>&1
  echo "$T smoke rc=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitizer_${T}_smoke.log | tail -1)"
  if [ -n "${SAN_PYTEST:-}" ]; then
    timeout ${SAN_TIMEOUT:-420} compute-sanitizer --tool $T $EXTRA --print-limit 40 --error-exitcode 3 \
        --log-file gpurun_out/sanitizer_${T}_ops.log python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "$SAN_PYTEST" \
        > gpurun_out/sanitizer_${T}_ops.out 2>&1
    echo "$T ops rc=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitizer_${T}_ops.log | tail -1)"
  fi
done
