#!/bin/bash
# round 2, GPU call 16: co-residency parity tests with the start barrier, three times
mkdir -p gpurun_out
for i in 1 2 3; do
  timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "config2 or config5" > gpurun_out/r02p_parity_$i.log 2>&1; echo "pytest rc $?" >> gpurun_out/r02p_parity_$i.log
  tail -2 gpurun_out/r02p_parity_$i.log
done
