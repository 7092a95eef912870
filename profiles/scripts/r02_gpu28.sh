#!/bin/bash
# round 2, GPU call 28: one-warp kernel with the running totals pre-loaded (plain read-modify-write instead of atomics with
# return) -- parity, live hook tests, its device time under ncu in a hooked storm (flush forced every 2 records)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_acct.py tests/test_gpu_hook.py -m gpu -q -x > gpurun_out/r02ab_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r02ab_pytest.log
tail -3 gpurun_out/r02ab_pytest.log
T=$(mktemp -d); printf '1\nbench/c0 1.0 1.0 8589934592\n' > $T/quota.txt
GEMHOOK_FLUSH_RECORDS=2 GEMHOOK_SEG_MIN_US=0 GEMHOOK_POOL=$T/pool GEMHOOK_QUOTA_FILE=$T/quota.txt POD_NAME=bench/c0 ncu --target-processes all --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02ab_launches_hooked_storm.csv env LD_PRELOAD=$PWD/kubeshare_b200/lib/libgemhook.so.1 kubeshare_b200/bin/gem-storm --mode storm --steps 1 --warmup 1 --step-launches 1024 --sync-every 256 > gpurun_out/r02ab_ncu.log 2>&1
grep gemhook gpurun_out/r02ab_launches_hooked_storm.csv | cut -d, -f5,15 | head -12
