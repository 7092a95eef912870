#!/bin/bash
# round 2, GPU call 9: harness A/B after the one-cache-line fast path; ledger parity with entry dumps
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02_build.log 2>&1
run() { tag=$1; c=$2; shift 2
  env "$@" python bench.py --clients $c --headline-clients $c --reps 4 --steps 10 --warmup 3 --skip-roofline --skip-baseline --skip-other 2> gpurun_out/r02i_$tag.log | python -c "
import sys, json
d = json.loads(sys.stdin.read()); c = d['clients'][str($c)]
print('$tag', 'clients', $c, 'unhooked', c['unhooked_launches_per_s'], 'hooked', c['hooked_launches_per_s'], 'overhead_pct', c['overhead_pct'], 'reps', d['reps_values'], d.get('ledger_gaps'))" >> gpurun_out/r02i_ab.log
}
run r2_c1 1 A=1
run r1_c1 1 GEMBENCH_HOOK=$PWD/profiles/ab/libgemhook_r1.so.1
run r2_c1_again 1 A=1
run r2_c2 2 A=1
run r1_c2 2 GEMBENCH_HOOK=$PWD/profiles/ab/libgemhook_r1.so.1
cat gpurun_out/r02i_ab.log
for i in 1 2; do timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "config2" 2>&1 | grep -E "^ledger|passed|failed|assert |Error" | cut -c1-1800; done
