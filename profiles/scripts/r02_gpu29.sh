#!/bin/bash
# round 2, GPU call 29: the other two workloads through bench.py's one-line schema at HEAD (short runs) + kernel parity once more
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_acct.py -m gpu -q -x > gpurun_out/r02ac_acct.log 2>&1; tail -1 gpurun_out/r02ac_acct.log
for w in bursty mnist; do
  timeout 600 python bench.py --workload $w --steps 5 --warmup 3 --reps 2 --skip-roofline > gpurun_out/r02ac_bench_$w.json 2> gpurun_out/r02ac_bench_$w.err
  echo "$w rc $? bytes $(wc -c < gpurun_out/r02ac_bench_$w.json)"
  python -c "
import json
d=json.loads(open('gpurun_out/r02ac_bench_$w.json').read().strip().splitlines()[-1])
print('$w', d['metric'], round(d['value'],1), d['unit'], 'overhead', d.get('overhead_pct'), 'jain', d.get('jain_fairness'), 'cpu_baseline', (d.get('cpu_baseline') or {}).get('value'))
"
done
