#!/bin/bash
# round 2, GPU call 20: full ncu capture of the TMA-staged kernel at 64 and 32 slots
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:gemhook_acct_reduce_staged -s 3 -c 1 -o gpurun_out/r02_prof_staged_64slots python bench.py --only-roofline --steps 3 --warmup 3 --nslots 64 > gpurun_out/r02t_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemhook_acct_reduce_staged -s 3 -c 1 -o gpurun_out/r02_prof_staged_32slots python bench.py --only-roofline --steps 3 --warmup 3 --nslots 32 > gpurun_out/r02t_ncu2.log 2>&1
ls -la gpurun_out/*.ncu-rep; tail -2 gpurun_out/r02t_ncu1.log
