#!/bin/bash
# round 2, GPU call 23: final kernel set -- parity of every kernel, slot sweep with the host's own selection, ncu at 64 slots
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_acct.py -m gpu -x -q > gpurun_out/r02w_acct.log 2>&1; echo "pytest rc $?" >> gpurun_out/r02w_acct.log
tail -3 gpurun_out/r02w_acct.log
timeout 600 python profiles/scripts/r02_sweep_final2.py > gpurun_out/r02w_sweep_final2.jsonl 2> gpurun_out/r02w_sweep_final2.err; echo "sweep rc $?"
tail -3 gpurun_out/r02w_sweep_final2.err; cut -c1-200 gpurun_out/r02w_sweep_final2.jsonl
ncu --set full --clock-control none --import-source on -k regex:gemhook_acct_reduce -s 3 -c 2 -o gpurun_out/r02_prof_acct_64slots_staged python bench.py --only-roofline --steps 3 --warmup 3 --nslots 64 > gpurun_out/r02w_ncu.log 2>&1
tail -2 gpurun_out/r02w_ncu.log
