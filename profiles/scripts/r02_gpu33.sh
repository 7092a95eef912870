#!/bin/bash
# round 2, GPU call 33: compute-sanitizer (memcheck, racecheck, synccheck) over the accounting kernels' parity tests (a subset of sizes)
mkdir -p gpurun_out
SEL='(staged and (70001 or 4096 or 769)) or (fast_path and (513 or 33)) or (reduce_host and (4097 or 257) and not 1048576)'
for tool in memcheck racecheck synccheck; do
  timeout 500 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_acct.py -m gpu -q -x -k "$SEL" > gpurun_out/r02ae_sanitizer_$tool.log 2>&1
  echo "$tool rc $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|hazard" gpurun_out/r02ae_sanitizer_$tool.log | tail -3
done
