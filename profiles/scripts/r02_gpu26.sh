#!/bin/bash
# round 2, GPU call 26: crossover between the register-staged and the TMA-staged kernel
mkdir -p gpurun_out
timeout 600 python profiles/scripts/r02_sweep_crossover.py > gpurun_out/r02z_crossover.jsonl 2> gpurun_out/r02z_crossover.err; echo "rc $?"
tail -2 gpurun_out/r02z_crossover.err; cut -c1-210 gpurun_out/r02z_crossover.jsonl
