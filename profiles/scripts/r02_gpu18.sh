#!/bin/bash
# round 2, GPU call 18: TMA-staged many-slot kernel -- parity, sweep; the parity tests the -x run of call 17 did not reach
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_acct.py -m gpu -x -q > gpurun_out/r02r_acct.log 2>&1; echo "pytest rc $?" >> gpurun_out/r02r_acct.log
tail -3 gpurun_out/r02r_acct.log
timeout 600 python profiles/scripts/r02_sweep_staged.py > gpurun_out/r02r_sweep_staged.jsonl 2> gpurun_out/r02r_sweep_staged.err; echo "sweep rc $?"
tail -3 gpurun_out/r02r_sweep_staged.err
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "yield or truth or scrape or graph" > gpurun_out/r02r_parity_rest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r02r_parity_rest.log
tail -3 gpurun_out/r02r_parity_rest.log
