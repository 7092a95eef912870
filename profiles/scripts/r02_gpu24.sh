#!/bin/bash
# round 2, GPU call 24: the driver's round-end sequence at HEAD -- GPU tests, smoke(), both bench arms
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv > gpurun_out/r02x_smi.txt
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/r02x_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r02x_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02x_smoke.log 2>&1
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02x_bench_reference.json 2> gpurun_out/r02x_bench_reference.err
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02x_bench_ours.json 2> gpurun_out/r02x_bench_ours.err
tail -4 gpurun_out/r02x_pytest.log; tail -2 gpurun_out/r02x_smoke.log; cut -c1-600 gpurun_out/r02x_bench_reference.json; cut -c1-1200 gpurun_out/r02x_bench_ours.json
