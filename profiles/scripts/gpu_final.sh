cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 16 --warmup 3 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.log; tail -5 gpurun_out/bench_ours.log
python bench.py --impl reference --steps 16 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.log; tail -5 gpurun_out/bench_ref.log
ncu --set full --clock-control none --import-source on -k regex:gemhook_acct_reduce -s 3 -c 2 -o gpurun_out/prof_acct3 python bench.py --only-roofline --steps 3 --warmup 3 > gpurun_out/ncu2.log 2>&1; tail -1 gpurun_out/ncu2.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_roofline3.csv python bench.py --only-roofline --steps 3 --warmup 3 > gpurun_out/ncu1.log 2>&1
mkdir -p /tmp/gh; printf '1\nbench/c0 1.0 1.0 8589934592\n' > /tmp/gh/quota.txt; rm -f /tmp/gh/pool
GEMHOOK_FLUSH_RECORDS=4 GEMHOOK_SEG_MIN_US=0 GEMHOOK_POOL=/tmp/gh/pool GEMHOOK_QUOTA_FILE=/tmp/gh/quota.txt POD_NAME=bench/c0 ncu --target-processes all --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/launches_storm3.csv env LD_PRELOAD=$PWD/kubeshare_b200/lib/libgemhook.so.1 kubeshare_b200/bin/gem-storm --mode storm --steps 1 --warmup 0 --step-launches 8192 --sync-every 1024 > gpurun_out/ncu3.log 2>&1; tail -1 gpurun_out/ncu3.log | cut -c1-200
