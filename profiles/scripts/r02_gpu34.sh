#!/bin/bash
# round 2, GPU call 34: final HEAD -- whole -m gpu suite, then our bench arm with the driver's flags
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -s > gpurun_out/r02af_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r02af_pytest.log
tail -3 gpurun_out/r02af_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02af_bench_ours.json 2> gpurun_out/r02af_bench_ours.err; echo "bench rc $?"
cut -c1-400 gpurun_out/r02af_bench_ours.json
