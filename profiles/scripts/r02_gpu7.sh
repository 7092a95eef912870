#!/bin/bash
# round 2, GPU call 7: final kernel (shuffle fold for few slots) -- parity, slot sweep, ncu captures; hook A/B at two sync rates
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02_build.log 2>&1
cp kubeshare_b200/csrc/build/acct_kernels.ptxas.txt gpurun_out/r02g_ptxas.txt
timeout 900 python -m pytest tests/test_gpu_acct.py tests/test_gpu_hook.py -m gpu -q -x 2>&1 | tail -5 > gpurun_out/r02g_pytest.log
python profiles/scripts/r02_sweep.py final > gpurun_out/r02g_sweep.jsonl 2> gpurun_out/r02g_sweep.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_roofline_leg.csv python bench.py --only-roofline --steps 3 --warmup 3 > gpurun_out/r02g_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemhook_acct_reduce -s 3 -c 2 -o gpurun_out/r02_prof_acct_2slots python bench.py --only-roofline --steps 3 --warmup 3 > gpurun_out/r02g_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemhook_acct_reduce -s 3 -c 2 -o gpurun_out/r02_prof_acct_64slots python bench.py --only-roofline --steps 3 --warmup 3 --nslots 64 > gpurun_out/r02g_ncu3.log 2>&1
T=$(mktemp -d); printf '1\nbench/c0 1.0 1.0 8589934592\n' > $T/quota.txt
GEMHOOK_FLUSH_RECORDS=2 GEMHOOK_SEG_MIN_US=0 GEMHOOK_POOL=$T/pool GEMHOOK_QUOTA_FILE=$T/quota.txt POD_NAME=bench/c0 ncu --target-processes all --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_launches_hooked_storm.csv env LD_PRELOAD=$PWD/kubeshare_b200/lib/libgemhook.so.1 kubeshare_b200/bin/gem-storm --mode storm --steps 1 --warmup 1 --step-launches 1024 --sync-every 256 > gpurun_out/r02g_ncu4.log 2>&1
# A/B: un-hooked vs round-1 hook vs this hook, one client, sync every 1024 and every 16 launches
for se in 1024 16; do
 for rep in 1 2 3; do
  for lib in none profiles/ab/libgemhook_r1.so.1 kubeshare_b200/lib/libgemhook.so.1; do
    rm -f $T/pool
    if [ $lib = none ]; then
      taskset -c 2 kubeshare_b200/bin/gem-storm --mode storm --steps 10 --warmup 3 --sync-every $se | python -c "import sys,json; d=json.load(sys.stdin); print('AB sync$se unhooked', round(d['launches']/d['event_ms']*1e3))" >> gpurun_out/r02g_ab.log
    else
      LD_PRELOAD=$lib GEMHOOK_POOL=$T/pool GEMHOOK_QUOTA_FILE=$T/quota.txt POD_NAME=bench/c0 taskset -c 2 kubeshare_b200/bin/gem-storm --mode storm --steps 10 --warmup 3 --sync-every $se | python -c "import sys,json; d=json.load(sys.stdin); print('AB sync$se $lib', round(d['launches']/d['event_ms']*1e3))" >> gpurun_out/r02g_ab.log
    fi
  done
 done
done
tail -3 gpurun_out/r02g_pytest.log
python - <<PY
import json
for l in open("gpurun_out/r02g_sweep.jsonl"):
    d=json.loads(l); print(d["tag"], d["nslots"], d["n"], d["env"], d["ms"], d["gbps"], d["frac"], d["grid"])
PY
tail -2 gpurun_out/r02g_sweep.err
grep -E "gemhook" gpurun_out/r02_launches_hooked_storm.csv | awk -F'","' '{print $5, $NF}' | tr -d '"' | head -12
cat gpurun_out/r02g_ab.log
