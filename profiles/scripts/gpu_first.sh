set -x
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv
nproc; lscpu | grep -E "Model name|^CPU\(s\)|Thread|Socket" 
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests/test_gpu_acct.py -x -q -m gpu 2>&1 | tail -15
S=kubeshare_b200/bin/gem-storm
$S --mode probe
$S --mode storm --steps 4 --warmup 2 --step-launches 65536 > gpurun_out/storm_unhooked.json; cat gpurun_out/storm_unhooked.json | cut -c1-260
mkdir -p /tmp/gh; printf '1\nbench/c0 1.0 1.0 8589934592\n' > /tmp/gh/quota.txt
GEMHOOK_LOG=1 GEMHOOK_POOL=/tmp/gh/pool GEMHOOK_QUOTA_FILE=/tmp/gh/quota.txt POD_NAME=bench/c0 GEMHOOK_STATS_FILE=gpurun_out/stats_hooked.json LD_PRELOAD=$PWD/kubeshare_b200/lib/libgemhook.so.1 $S --mode storm --steps 4 --warmup 2 --step-launches 65536 > gpurun_out/storm_hooked.json; cut -c1-260 gpurun_out/storm_hooked.json; cat gpurun_out/stats_hooked.json
# reference hook with reference daemons
printf '1\nbench/c0 1.0 1.0 8589934592\n' > /tmp/gh/resource.txt
oracle/_ref/gem-schd -p /tmp/gh -f resource.txt -P 49901 -q 300 -m 20 -w 10000 > gpurun_out/schd.log 2>&1 &
SCHD=$!
sleep 0.5
POD_NAME=bench/c0 POD_MANAGER_PORT=50061 SCHEDULER_IP=127.0.0.1 SCHEDULER_PORT=49901 oracle/_ref/gem-pmgr > gpurun_out/pmgr.log 2>&1 &
PMGR=$!
sleep 0.5
mkdir -p /kubeshare/library /kubeshare/log; echo 127.0.0.1 > /kubeshare/library/schedulerIP.txt
POD_NAME=bench/c0 POD_MANAGER_PORT=50061 LD_PRELOAD=$PWD/oracle/_ref/libgemhook_ref.so.1 timeout 120 $S --mode storm --steps 4 --warmup 2 --step-launches 65536 > gpurun_out/storm_ref.json 2> gpurun_out/ref_hook.log; cut -c1-260 gpurun_out/storm_ref.json; tail -3 gpurun_out/ref_hook.log
# our hook over TCP to the same reference daemons
GEMHOOK_LOG=1 GEMHOOK_SCHEDULER_IP=127.0.0.1 POD_NAME=bench/c0 POD_MANAGER_PORT=50061 GEMHOOK_STATS_FILE=gpurun_out/stats_tcp.json LD_PRELOAD=$PWD/kubeshare_b200/lib/libgemhook.so.1 timeout 120 $S --mode storm --steps 4 --warmup 2 --step-launches 65536 > gpurun_out/storm_tcp.json; cut -c1-260 gpurun_out/storm_tcp.json; cat gpurun_out/stats_tcp.json
kill $PMGR $SCHD
