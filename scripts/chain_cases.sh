#!/bin/bash
# Run every case of tests/test_gpu_chain.py in its own process (a trapped kernel kills the CUDA context).
mkdir -p gpurun_out
ids=$(python -m pytest tests/test_gpu_chain.py --collect-only -q -m gpu 2>/dev/null | grep "::" )
for id in $ids; do
  name=$(echo $id | sed 's/.*\[\(.*\)\]/\1/')
  AICAM_CHAIN_DEBUG=1 timeout 120 python -m pytest "$id" -q -m gpu --no-header -p no:cacheprovider -x > gpurun_out/chain_$name.log 2>&1
  echo "$name exit $?: $(grep -E 'passed|failed|error' gpurun_out/chain_$name.log | tail -1) | $(grep -h 'conv_chain:' gpurun_out/chain_$name.log | head -1) | $(grep -h -E 'elements off|aicam error|timeout' gpurun_out/chain_$name.log | head -1 | cut -c1-200)"
done
