#!/bin/bash
# round 2, GPU call 4: ILP-grouped bin update (variants), parity tests with exit-clipped ledgers, r1-vs-r2 hook A/B
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02_build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_acct.py -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r02d_pytest_acct.log
python profiles/scripts/r02_sweep.py ILP4_U8 quick > gpurun_out/r02d_sweep.jsonl 2> gpurun_out/r02d_sweep.err
for v in "ILP2_U8:-DGEMHOOK_ILP=2" "ILP1_U8:-DGEMHOOK_ILP=1" "ILP4_U16:-DGEMHOOK_UNROLL=16" "ILP8_U8:-DGEMHOOK_ILP=8" "ILP8_U16:-DGEMHOOK_ILP=8 -DGEMHOOK_UNROLL=16"; do
  tag=${v%%:*}; flags=${v#*:}
  make -s -C kubeshare_b200/csrc clean > /dev/null 2>&1; make -s -C kubeshare_b200/csrc VARIANT="$flags" > gpurun_out/r02d_build_$tag.log 2>&1
  python profiles/scripts/r02_sweep.py $tag quick >> gpurun_out/r02d_sweep.jsonl 2>> gpurun_out/r02d_sweep.err
done
make -s -C kubeshare_b200/csrc clean > /dev/null 2>&1; make -s -C kubeshare_b200/csrc > /dev/null 2>&1
timeout 1800 python -m pytest tests/test_gpu_parity.py -m gpu -q -s 2>&1 | tail -120 > gpurun_out/r02d_pytest_parity.log
kubeshare_b200/bin/gem-storm --mode probe > gpurun_out/r02d_probe.json 2>&1
# A/B: round-1 hook vs this hook, one client, same box, alternating
T=$(mktemp -d); printf '1\nbench/c0 1.0 1.0 8589934592\n' > $T/quota.txt
for rep in 1 2 3; do
  for lib in none profiles/ab/libgemhook_r1.so.1 kubeshare_b200/lib/libgemhook.so.1; do
    rm -f $T/pool
    if [ $lib = none ]; then
      taskset -c 2 kubeshare_b200/bin/gem-storm --mode storm --steps 20 --warmup 3 | python -c "import sys,json; d=json.load(sys.stdin); print('AB unhooked', round(d['launches']/d['event_ms']*1e3))" >> gpurun_out/r02d_ab.log
    else
      LD_PRELOAD=$lib GEMHOOK_POOL=$T/pool GEMHOOK_QUOTA_FILE=$T/quota.txt POD_NAME=bench/c0 taskset -c 2 kubeshare_b200/bin/gem-storm --mode storm --steps 20 --warmup 3 | python -c "import sys,json; d=json.load(sys.stdin); print('AB $lib', round(d['launches']/d['event_ms']*1e3))" >> gpurun_out/r02d_ab.log
    fi
  done
done
tail -3 gpurun_out/r02d_pytest_acct.log
python - <<PY
import json
for l in open("gpurun_out/r02d_sweep.jsonl"):
    d=json.loads(l); print(d["tag"], d["nslots"], d["env"], d["ms"], d["gbps"], d["frac"], d["grid"])
PY
tail -2 gpurun_out/r02d_sweep.err
grep -E "passed|failed|^ledgers|^F?ledgers|graph replays|Error|assert " gpurun_out/r02d_pytest_parity.log | cut -c1-1200
cat gpurun_out/r02d_probe.json gpurun_out/r02d_ab.log
