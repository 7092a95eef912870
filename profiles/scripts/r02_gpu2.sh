#!/bin/bash
# round 2, GPU call 2: full -m gpu suite on the lock-free pool, graph-mode diagnostics, both bench arms
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02_build.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q -s -x --deselect tests/test_gpu_parity.py 2>&1 | tail -40 > gpurun_out/r02b_pytest_old.log
timeout 1800 python -m pytest tests/test_gpu_parity.py -m gpu -q -s 2>&1 | tail -150 > gpurun_out/r02b_pytest_parity.log
# graph mode: what is accounted, with and without per-K segment ticks / short tokens
T=$(mktemp -d); printf '1\nbench/c0 1.0 1.0 8589934592\n' > $T/quota.txt
for v in "GEMHOOK_BASE_QUOTA_MS=5 GEMHOOK_MIN_QUOTA_MS=2 GEMHOOK_SEG_MIN_US=100 GEMHOOK_SEG_LAUNCHES=16" "GEMHOOK_BASE_QUOTA_MS=300" "GEMHOOK_BASE_QUOTA_MS=5 GEMHOOK_MIN_QUOTA_MS=2"; do
  rm -f $T/pool $T/stats.json
  env $v LD_PRELOAD=kubeshare_b200/lib/libgemhook.so.1 GEMHOOK_POOL=$T/pool GEMHOOK_QUOTA_FILE=$T/quota.txt POD_NAME=bench/c0 GEMHOOK_STATS_FILE=$T/stats.json \
    kubeshare_b200/bin/gem-storm --mode graph --step-launches 64 --rounds 20 --spin-us 20 >> gpurun_out/r02b_graph.log 2>&1
  echo "$v" >> gpurun_out/r02b_graph.log; cat $T/stats.json >> gpurun_out/r02b_graph.log
done
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02b_bench_ref.json 2> gpurun_out/r02b_bench_ref.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r02b_bench_ours.json 2> gpurun_out/r02b_bench_ours.log
tail -3 gpurun_out/r02b_pytest_old.log; grep -E "passed|failed|^ledgers|graph replays|pool_kill" gpurun_out/r02b_pytest_parity.log | cut -c1-900
cat gpurun_out/r02b_graph.log | cut -c1-700
cut -c1-1500 gpurun_out/r02b_bench_ref.json; cut -c1-3000 gpurun_out/r02b_bench_ours.json; tail -3 gpurun_out/r02b_bench_ours.log
