#!/bin/bash
# Run every GPU test file in its own process (a faulting kernel must not poison the others).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
rc=0
for f in ${1:-tests/test_gpu_*.py}; do
  name=$(basename $f .py)
  timeout 600 python -m pytest $f -q -m gpu --timeout 300 -x --no-header -p no:cacheprovider > gpurun_out/$name.log 2>&1
  code=$?
  echo "$name exit $code: $(tail -1 gpurun_out/$name.log)"
  [ $code -ne 0 ] && rc=1
done
exit $rc
