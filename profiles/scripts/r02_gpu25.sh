#!/bin/bash
# round 2, GPU call 25 (--gpus 2): the driver's multi-GPU launch line, both arms, short run -- does the line parse, are both replicas real
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02y_smi.txt
for arm in reference ours; do
  extra=""; [ $arm = reference ] && extra="--impl reference"
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 6 --warmup 3 $extra > gpurun_out/r02y_bench_${arm}_2gpu.json 2> gpurun_out/r02y_bench_${arm}_2gpu.err
  echo "$arm rc $? bytes $(wc -c < gpurun_out/r02y_bench_${arm}_2gpu.json) lines $(wc -l < gpurun_out/r02y_bench_${arm}_2gpu.json)"
  python -c "
import json,sys
d=json.loads(open('gpurun_out/r02y_bench_${arm}_2gpu.json').read().strip().splitlines()[-1])
print('$arm', 'n_gpus', d['n_gpus'], 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'clients', {k:round(v['hooked_launches_per_s']) for k,v in d['clients'].items()})
"
done
tail -3 gpurun_out/r02y_bench_ours_2gpu.err
