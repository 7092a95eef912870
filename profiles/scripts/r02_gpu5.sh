#!/bin/bash
# round 2, GPU call 5: LDS.128 / 16-column kernel variants, fixed parity tests, DEBUG=1 reference hook crash diagnosis
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02_build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_acct.py -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r02e_pytest_acct.log
python profiles/scripts/r02_sweep.py ILP4 quick > gpurun_out/r02e_sweep.jsonl 2> gpurun_out/r02e_sweep.err
for v in "ILP1:-DGEMHOOK_ILP=1" "ILP2:-DGEMHOOK_ILP=2" "C16_ILP1:-DGEMHOOK_COLS=16 -DGEMHOOK_ILP=1" "C16_ILP2:-DGEMHOOK_COLS=16 -DGEMHOOK_ILP=2" "C16_ILP4:-DGEMHOOK_COLS=16" "C16_ILP1_U16:-DGEMHOOK_COLS=16 -DGEMHOOK_ILP=1 -DGEMHOOK_UNROLL=16" "ILP1_U16:-DGEMHOOK_ILP=1 -DGEMHOOK_UNROLL=16"; do
  tag=${v%%:*}; flags=${v#*:}
  make -s -C kubeshare_b200/csrc clean > /dev/null 2>&1; make -s -C kubeshare_b200/csrc VARIANT="$flags" > gpurun_out/r02e_build_$tag.log 2>&1
  python profiles/scripts/r02_sweep.py $tag quick >> gpurun_out/r02e_sweep.jsonl 2>> gpurun_out/r02e_sweep.err
  if [ $tag = C16_ILP1 ]; then timeout 600 python -m pytest tests/test_gpu_acct.py -m gpu -q -x 2>&1 | tail -3 > gpurun_out/r02e_pytest_acct_c16.log; fi
done
make -s -C kubeshare_b200/csrc clean > /dev/null 2>&1; make -s -C kubeshare_b200/csrc > /dev/null 2>&1
timeout 1800 python -m pytest tests/test_gpu_parity.py -m gpu -q -s 2>&1 | tail -120 > gpurun_out/r02e_pytest_parity.log
# why does the reference's DEBUG=1 hook flavour crash?
gcc -O1 -g -fPIC -shared -o /tmp/segv_trace.so profiles/scripts/segv_trace.c
mkdir -p /kubeshare/library /kubeshare/log; echo 127.0.0.1 > /kubeshare/library/schedulerIP.txt
T=$(mktemp -d); printf '1\nbench/c0 1.0 1.0 8589934592\n' > $T/quota.txt
oracle/_ref/gem-schd -p $T -f quota.txt -P 50911 -q 300 -m 20 -w 10000 > /dev/null 2>&1 &
S=$!; sleep 0.5
POD_NAME=bench/c0 POD_MANAGER_PORT=50912 SCHEDULER_IP=127.0.0.1 SCHEDULER_PORT=50911 oracle/_ref/gem-pmgr > /dev/null 2>&1 &
P=$!; sleep 0.5
LD_PRELOAD=/tmp/segv_trace.so:oracle/_ref/libgemhook_ref_dbg.so.1 POD_NAME=bench/c0 POD_MANAGER_PORT=50912 timeout 60 kubeshare_b200/bin/gem-storm --mode storm --steps 1 --warmup 0 --step-launches 4096 > gpurun_out/r02e_dbg_hook.log 2>&1
echo "rc=$?" >> gpurun_out/r02e_dbg_hook.log; tail -5 /kubeshare/log/hook.log >> gpurun_out/r02e_dbg_hook.log 2>&1
kill $P $S
tail -3 gpurun_out/r02e_pytest_acct.log gpurun_out/r02e_pytest_acct_c16.log
python - <<PY
import json
for l in open("gpurun_out/r02e_sweep.jsonl"):
    d=json.loads(l); print(d["tag"], d["nslots"], d["env"], d["ms"], d["gbps"], d["frac"], d["grid"])
PY
tail -2 gpurun_out/r02e_sweep.err
grep -E "passed|failed|ledgers|graph replays|Error|assert |EMA" gpurun_out/r02e_pytest_parity.log | cut -c1-1500
head -40 gpurun_out/r02e_dbg_hook.log | cut -c1-300
