#!/bin/bash
# round 2, GPU call 6: everything at HEAD -- full -m gpu suite, smoke(), ncu launch lists + full captures of the accounting
# kernel (2 and 64 slots) and of the one-warp fast path, DEBUG=1 reference flavour, both bench arms
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02_build.log 2>&1
timeout 2400 python -m pytest tests -m gpu -q -s 2>&1 | tail -60 > gpurun_out/r02f_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02f_smoke.log 2>&1
# launch list of the roofline leg (kernel share) and full captures
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_roofline_leg.csv python bench.py --only-roofline --steps 3 --warmup 3 > gpurun_out/r02f_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemhook_acct_reduce -s 3 -c 2 -o gpurun_out/r02_prof_acct_2slots python bench.py --only-roofline --steps 3 --warmup 3 > gpurun_out/r02f_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemhook_acct_reduce -s 3 -c 2 -o gpurun_out/r02_prof_acct_64slots python bench.py --only-roofline --steps 3 --warmup 3 --nslots 64 > gpurun_out/r02f_ncu3.log 2>&1
# the live hook's regime: flush forced every 2 records so that our kernel shows up often
T=$(mktemp -d); printf '1\nbench/c0 1.0 1.0 8589934592\n' > $T/quota.txt
GEMHOOK_FLUSH_RECORDS=2 GEMHOOK_SEG_MIN_US=0 GEMHOOK_POOL=$T/pool GEMHOOK_QUOTA_FILE=$T/quota.txt POD_NAME=bench/c0 ncu --target-processes all --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_launches_hooked_storm.csv env LD_PRELOAD=$PWD/kubeshare_b200/lib/libgemhook.so.1 kubeshare_b200/bin/gem-storm --mode storm --steps 1 --warmup 1 --step-launches 1024 --sync-every 256 > gpurun_out/r02f_ncu4.log 2>&1
rm -f $T/pool
GEMHOOK_POOL=$T/pool GEMHOOK_QUOTA_FILE=$T/quota.txt POD_NAME=bench/c0 ncu --target-processes all --metrics gpu__time_duration.sum --clock-control none -c 70000 --csv --log-file gpurun_out/r02_launches_hooked_step.csv env LD_PRELOAD=$PWD/kubeshare_b200/lib/libgemhook.so.1 kubeshare_b200/bin/gem-storm --mode storm --steps 1 --warmup 0 --step-launches 65536 --sync-every 1024 > gpurun_out/r02f_ncu5.log 2>&1
python - <<PY > gpurun_out/r02_hooked_step_summary.json
import csv, json, collections
tot=collections.Counter(); cnt=collections.Counter()
rows=[r for r in csv.reader(open("gpurun_out/r02_launches_hooked_step.csv")) if len(r)>5]
hdr=None
for r in rows:
    if "Kernel Name" in r: hdr=r; continue
    if hdr is None: continue
    d=dict(zip(hdr,r))
    if d.get("Metric Name")!="gpu__time_duration.sum": continue
    k=d["Kernel Name"]; v=float(d["Metric Value"].replace(",","")); u=d.get("Metric Unit","")
    if u in ("us","usecond"): v*=1e3
    elif u in ("ms","msecond"): v*=1e6
    tot[k]+=v; cnt[k]+=1
print(json.dumps({"launches": dict(cnt), "device_ns": dict(tot), "share_of_device_time": {k: tot[k]/sum(tot.values()) for k in tot}}))
PY
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02f_bench_ref.json 2> gpurun_out/r02f_bench_ref.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r02f_bench_ours.json 2> gpurun_out/r02f_bench_ours.log
grep -E "passed|failed|^ledgers|graph replays|EMA|yield-on-idle|truth |live scrape|FAILED|Error" gpurun_out/r02f_pytest_gpu.log | cut -c1-700
cat gpurun_out/r02f_smoke.log | tail -2
cat gpurun_out/r02_hooked_step_summary.json
grep -E "gemhook" gpurun_out/r02_launches_hooked_storm.csv | awk -F'","' '{print $5, $NF}' | tr -d '"' | head -12
tail -2 gpurun_out/r02f_ncu2.log gpurun_out/r02f_ncu3.log | cut -c1-200
cut -c1-2500 gpurun_out/r02f_bench_ours.json; echo; cut -c1-800 gpurun_out/r02f_bench_ref.json; tail -4 gpurun_out/r02f_bench_ours.log | cut -c1-300
