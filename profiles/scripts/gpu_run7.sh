cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out /tmp/gh /kubeshare/library /kubeshare/log; echo 127.0.0.1 > /kubeshare/library/schedulerIP.txt
S=kubeshare_b200/bin/gem-storm
printf '1\nbench/c0 1.0 1.0 8589934592\n' > /tmp/gh/quota.txt
for SE in 16 128 1024; do
  for rep in 1 2; do
  U=$($S --mode storm --steps 4 --warmup 2 --step-launches 65536 --sync-every $SE | python -c "import json,sys;print(json.load(sys.stdin)['wall_s'])")
  rm -f /tmp/gh/pool
  O=$(GEMHOOK_POOL=/tmp/gh/pool GEMHOOK_QUOTA_FILE=/tmp/gh/quota.txt POD_NAME=bench/c0 GEMHOOK_STATS_FILE=/tmp/gh/st.json LD_PRELOAD=$PWD/kubeshare_b200/lib/libgemhook.so.1 $S --mode storm --steps 4 --warmup 2 --step-launches 65536 --sync-every $SE | python -c "import json,sys;print(json.load(sys.stdin)['wall_s'])")
  SEG=$(python -c "import json;d=json.load(open('/tmp/gh/st.json'));print(d['segments'],d['gpu_ns']/1e6)")
  oracle/_ref/gem-schd -p /tmp/gh -f quota.txt -P 49901 -q 300 -m 20 -w 10000 > /dev/null 2>&1 &
  SCHD=$!; sleep 0.3
  POD_NAME=bench/c0 POD_MANAGER_PORT=50061 SCHEDULER_IP=127.0.0.1 SCHEDULER_PORT=49901 oracle/_ref/gem-pmgr > /dev/null 2>&1 &
  PMGR=$!; sleep 0.3
  R=$(POD_NAME=bench/c0 POD_MANAGER_PORT=50061 LD_PRELOAD=$PWD/oracle/_ref/libgemhook_ref.so.1 $S --mode storm --steps 4 --warmup 2 --step-launches 65536 --sync-every $SE 2>/dev/null | python -c "import json,sys;print(json.load(sys.stdin)['wall_s'])")
  kill $PMGR $SCHD; wait $PMGR $SCHD 2>/dev/null; sleep 1.2
  echo "sync_every=$SE unhooked=$U ours=$O ref=$R segs/gpu_ms=$SEG"
  done
done
python -m pytest tests/test_gpu_hook.py -x -q -m gpu 2>&1 | tail -3
