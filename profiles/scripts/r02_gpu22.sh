#!/bin/bash
# round 2, GPU call 22: TMA-staged kernel, update variants: group size 1/2/4, software-pipelined forwarding
mkdir -p gpurun_out
timeout 600 python profiles/scripts/r02_sweep_staged_ilp.py > gpurun_out/r02v_sweep_staged_ilp.jsonl 2> gpurun_out/r02v_sweep_staged_ilp.err; echo "sweep rc $?"
tail -3 gpurun_out/r02v_sweep_staged_ilp.err; cut -c1-230 gpurun_out/r02v_sweep_staged_ilp.jsonl
