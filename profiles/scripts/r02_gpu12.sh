#!/bin/bash
# run 12: ledger time vs the client's own unblocked time, three stacks, configs[1]
mkdir -p gpurun_out
timeout 900 python profiles/scripts/r02_ledger_vs_busy.py 3 > gpurun_out/r02l_ledger_vs_busy.log 2>&1
tail -20 gpurun_out/r02l_ledger_vs_busy.log
