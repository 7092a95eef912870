cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python -m pytest tests/test_gpu_hook.py -x -q -m gpu -k "pytorch or host_sum or modern" 2>&1 | tail -15
