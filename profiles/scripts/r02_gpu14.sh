#!/bin/bash
# round 2, GPU call 14: the whole -m gpu suite at HEAD (twice for the co-residency parity tests)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/r02n_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r02n_pytest.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "config2 or config5" > gpurun_out/r02n_parity_again.log 2>&1; echo "pytest rc $?" >> gpurun_out/r02n_parity_again.log
tail -4 gpurun_out/r02n_pytest.log; tail -3 gpurun_out/r02n_parity_again.log
