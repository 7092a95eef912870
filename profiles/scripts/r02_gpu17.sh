#!/bin/bash
# round 2, GPU call 17: the whole -m gpu suite at HEAD + smoke()
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/r02q_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r02q_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02q_smoke.log 2>&1
tail -4 gpurun_out/r02q_pytest.log; tail -2 gpurun_out/r02q_smoke.log
