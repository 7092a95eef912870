#!/bin/bash
# round 2, GPU call 1: full -m gpu suite (with the new parity / truth tests), primitive probe
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02_build.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q -s -x --deselect tests/test_gpu_parity.py 2>&1 | tail -60 > gpurun_out/r02_pytest_old.log
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -s 2>&1 | tail -150 > gpurun_out/r02_pytest_parity.log
kubeshare_b200/bin/gem-storm --mode probe > gpurun_out/r02_probe.json 2>&1
tail -5 gpurun_out/r02_pytest_old.log; grep -E "passed|failed|truth |ledgers|EMA|live scrape" gpurun_out/r02_pytest_parity.log | cut -c1-400
