cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 16 --warmup 3 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.log; tail -4 gpurun_out/bench_ours.log
python bench.py --impl reference --steps 16 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.log; tail -4 gpurun_out/bench_ref.log
