#!/bin/bash
# run 11: why is the reference-hooked storm sometimes faster than ours/unhooked?  + config2 parity twice more
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02k_smi.txt
timeout 600 python profiles/scripts/r02_rate_probe.py > gpurun_out/r02k_rate_probe.log 2>&1
for i in 1 2; do
  timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "config2" > gpurun_out/r02k_config2_$i.log 2>&1
done
tail -5 gpurun_out/r02k_rate_probe.log; tail -3 gpurun_out/r02k_config2_*.log
