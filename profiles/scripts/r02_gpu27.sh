#!/bin/bash
# round 2, GPU call 27: at HEAD -- the whole -m gpu suite, smoke(), fresh ncu captures (launch list of the roofline leg; full
# captures at 2 slots = register-staged kernel, 32 slots = TMA-staged with 32 columns)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/r02aa_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r02aa_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02aa_smoke.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02aa_launches_roofline_leg.csv python bench.py --only-roofline --steps 3 --warmup 3 > gpurun_out/r02aa_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemhook_acct_reduce -s 3 -c 2 -o gpurun_out/r02aa_prof_acct_2slots python bench.py --only-roofline --steps 3 --warmup 3 > gpurun_out/r02aa_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemhook_acct_reduce -s 3 -c 2 -o gpurun_out/r02aa_prof_acct_32slots python bench.py --only-roofline --steps 3 --warmup 3 --nslots 32 > gpurun_out/r02aa_ncu3.log 2>&1
tail -3 gpurun_out/r02aa_pytest.log; tail -1 gpurun_out/r02aa_smoke.log; ls -la gpurun_out/r02aa_*.ncu-rep gpurun_out/r02aa_launches_roofline_leg.csv
