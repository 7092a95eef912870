#!/bin/bash
# round 2, GPU call 8: where do 1-3 % go at 1-2 clients in the bench harness?  r1 hook vs r2 hook vs r2 without accounting
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02_build.log 2>&1
run() { # tag clients env...
  tag=$1; c=$2; shift 2
  env "$@" python bench.py --clients $c --headline-clients $c --reps 4 --steps 10 --warmup 3 --skip-roofline --skip-baseline --skip-other 2> gpurun_out/r02h_$tag.log | python -c "
import sys, json
d = json.loads(sys.stdin.read()); c = d['clients'][str($c)]
print('$tag', 'clients', $c, 'unhooked', c['unhooked_launches_per_s'], 'hooked', c['hooked_launches_per_s'], 'overhead_pct', c['overhead_pct'], 'reps', d['reps_values'])" >> gpurun_out/r02h_ab.log
}
for c in 1 2; do
  run r2_c$c $c A=1
  run r1_c$c $c GEMBENCH_HOOK=$PWD/profiles/ab/libgemhook_r1.so.1
  run r2dry_c$c $c GEMHOOK_DRY_RUN=1
  run r2nomerge_c$c $c GEMHOOK_SEG_MIN_US=100000000
done
run r2_c4 4 A=1
run r1_c4 4 GEMBENCH_HOOK=$PWD/profiles/ab/libgemhook_r1.so.1
cat gpurun_out/r02h_ab.log
python - <<PY
import json
d=json.load(open("gpurun_out/bench_detail_ours_storm_n1.json"))
PY
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_hook.py -m gpu -q -s 2>&1 | tail -40 > gpurun_out/r02h_pytest.log
grep -E "passed|failed|^ledgers|F?ledgers|Error|assert " gpurun_out/r02h_pytest.log | cut -c1-900
