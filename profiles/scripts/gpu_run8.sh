cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python bench.py --steps 16 --warmup 3 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.log; tail -5 gpurun_out/bench_ours.log; python -c "
import json; d=json.load(open('gpurun_out/bench_ours.json'))
print({k:(round(v['unhooked_launches_per_s']),round(v['hooked_launches_per_s']),round(v['overhead_pct'],2)) for k,v in d['clients'].items()})
print(d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['traffic'], d['cpu_baseline'], d['e2e'], d['gpu_launches'], d['clocks'])"
