"""CPU: bench.py's N>1 path (one rank per GPU, replicas, gloo gather) on the stub driver, world_size 2."""
import json
import os
import subprocess as sp
import sys

import kubeshare_b200 as kb
import wireproto as wp


def run_bench(nproc, extra_env=None):
    env = dict(os.environ, LD_LIBRARY_PATH=kb.STUB_DIR + ":" + os.environ.get("LD_LIBRARY_PATH", ""),
               GEMBENCH_STEP_LAUNCHES="4096", STUB_KERNEL_US="1")
    env.update(extra_env or {})
    args = ["bench.py", "--gpus", str(nproc), "--steps", "2", "--warmup", "3", "--clients", "1,2", "--reps", "2", "--skip-roofline",
            "--skip-baseline"]
    if nproc > 1:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc), "--master-addr",
               "127.0.0.1", "--master-port", str(wp.free_port())] + args
    else:
        cmd = [sys.executable] + args
    p = sp.run(cmd, cwd=kb.ROOT, env=env, stdout=sp.PIPE, stderr=sp.PIPE, timeout=600)
    assert p.returncode == 0, p.stderr.decode()[-3000:]
    lines = [l for l in p.stdout.decode().splitlines() if l.startswith("{")]
    assert len(lines) == 1, "exactly one JSON line from rank 0"
    return json.loads(lines[0])


def test_two_replicas_aggregate():
    one = run_bench(1)
    two = run_bench(2)
    for line in (one, two):
        for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                  "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "clients"):
            assert k in line, k
        assert line["metric"] == "hooked_launches_per_s" and line["unit"] == "launches/s" and line["scaling"] == "weak"
        assert "workload" in line["config"] and line["vs_baseline"] is None
        assert set(line["clients"]) == {"1", "2"}
    assert (one["n_gpus"], two["n_gpus"]) == (1, 2)
    # two independent replicas process twice the launches in about the same time
    # summed over replicas, not the slower one's alone (that would read ~1.0); wide: eight shared CPU cores run both replicas
    assert 1.15 < two["value"] / one["value"] < 3.2
    assert two["config"]["parallelism"].startswith("replicas x2")
