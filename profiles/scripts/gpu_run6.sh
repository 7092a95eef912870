cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
nvidia-smi -L
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 16 --warmup 3 --clients 1,2 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.log; tail -5 gpurun_out/bench_2gpu.log; cut -c1-1500 gpurun_out/bench_2gpu.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 16 --warmup 3 --clients 1,2 > gpurun_out/bench_2gpu_ref.json 2> gpurun_out/bench_2gpu_ref.log; tail -3 gpurun_out/bench_2gpu_ref.log; cut -c1-600 gpurun_out/bench_2gpu_ref.json
