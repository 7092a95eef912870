cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python -m pytest tests/test_gpu_acct.py tests/test_gpu_hook.py -x -q -m gpu 2>&1 | tail -3
python bench.py --only-roofline --steps 10 --warmup 3 2>/dev/null | grep "^{" | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('big %.1f GB/s'%d['ring_2p26']['gbps'], 'small avg %.1f us min %.1f us'%(d['ring_2p20']['avg_ms']*1e3, d['ring_2p20']['min_ms']*1e3))"
mkdir -p /tmp/gh; printf '1\nbench/c0 1.0 1.0 8589934592\n' > /tmp/gh/quota.txt; rm -f /tmp/gh/pool
GEMHOOK_FLUSH_RECORDS=2 GEMHOOK_SEG_MIN_US=0 GEMHOOK_POOL=/tmp/gh/pool GEMHOOK_QUOTA_FILE=/tmp/gh/quota.txt POD_NAME=bench/c0 ncu --target-processes all --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_small.csv env LD_PRELOAD=$PWD/kubeshare_b200/lib/libgemhook.so.1 kubeshare_b200/bin/gem-storm --mode storm --steps 1 --warmup 1 --step-launches 1024 --sync-every 256 > /dev/null 2>&1
grep gemhook_acct_reduce gpurun_out/launches_small.csv | awk -F'","' '{print $NF}' | tr -d '"' | tr '\n' ' '
python kubeshare_b200/tools/config5.py --gpus 1 --iters 30 2>/dev/null | grep "^{" > gpurun_out/config5_1gpu.json; python -c "
import json; d=json.load(open('gpurun_out/config5_1gpu.json'))
for k in ('unhooked','ours','reference'):
    if k in d: print(k, round(d[k]['aggregate_launches_per_s']), [ (round(x['jain_delivered_over_entitled'],3), round(x['jain_completion_time'],3), [round(w,2) for w in x['client_wall_s']]) for x in d[k]['per_device']])"
