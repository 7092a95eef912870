cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
nvidia-smi -L | wc -l
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 16 --warmup 3 --clients 1,2 --skip-roofline > gpurun_out/bench_8gpu.json 2> gpurun_out/bench_8gpu.log; grep -c "clients=" gpurun_out/bench_8gpu.log; tail -3 gpurun_out/bench_8gpu.log; python -c "
import json; d=json.load(open('gpurun_out/bench_8gpu.json')); print(d['value'], d['n_gpus'], d['overhead_pct'], d['clients']['1'], d['e2e'])"
python kubeshare_b200/tools/config5.py --gpus 8 --iters 30 2>/dev/null | grep "^{" > gpurun_out/config5_8gpu.json; python -c "
import json; d=json.load(open('gpurun_out/config5_8gpu.json'))
for k in ('unhooked','ours','reference'):
    if k in d: print(k, round(d[k]['aggregate_launches_per_s']), [round(x['launches_per_s']) for x in d[k]['per_device']], [round(x['jain_completion_time'],3) for x in d[k]['per_device']])"
