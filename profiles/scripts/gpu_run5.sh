cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for i in 1 2 3; do python bench.py --steps 16 --warmup 3 --clients 1 --headline-clients 1 --skip-roofline --skip-baseline 2>/dev/null | grep "^{" | python -c "
import json,sys; d=json.loads(sys.stdin.read()); c=d['clients']['1']; print('run', c['unhooked_launches_per_s'], c['hooked_launches_per_s'], c['overhead_pct'], d['hook_stats']['1'])"; done
python -m pytest tests/test_gpu_hook.py -x -q -m gpu 2>&1 | tail -3
