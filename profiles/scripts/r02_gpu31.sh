#!/bin/bash
# round 2, GPU call 31: 2 KB ring buffers, 9-10 warps per SM at 48-64 slots
mkdir -p gpurun_out
timeout 600 python profiles/scripts/r02_sweep_rows4.py > gpurun_out/r02ad_rows4.jsonl 2> gpurun_out/r02ad_rows4.err; echo "rc $?"
tail -2 gpurun_out/r02ad_rows4.err; cut -c1-250 gpurun_out/r02ad_rows4.jsonl
