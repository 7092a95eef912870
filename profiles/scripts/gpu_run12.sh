cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
lscpu | grep -E "Thread|Core|Socket|NUMA node\(s\)"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 4 --steps 16 --warmup 3 --clients 1,2 --skip-roofline --skip-baseline > gpurun_out/bench_4gpu.json 2> gpurun_out/bench_4gpu.log; grep "clients=" gpurun_out/bench_4gpu.log; python -c "
import json; d=json.load(open('gpurun_out/bench_4gpu.json')); print(d['value'], d['n_gpus'], d['overhead_pct'], d['clients']['1']['overhead_pct'], d['host'])"
