"""Overhead of the hook on a cudart application (PyTorch): small-kernel-heavy loop, timed with CUDA events."""
import json
import os
import subprocess as sp
import sys
import tempfile

SCRIPT = r'''
import json, torch, time
torch.manual_seed(0)
x = torch.randn(256, 1024, device="cuda")
ws = [torch.randn(1024, 1024, device="cuda") * 0.03 for _ in range(4)]
def step(x):
    for w in ws:
        x = torch.relu(x @ w) + 0.1          # 3 small kernels per layer
    return x
for _ in range(200):
    y = step(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.time(); e0.record()
for i in range(4000):
    y = step(x)
    if i % 100 == 99:
        float(y[0, 0])                        # a DtoH copy every 100 iterations (cuMemcpyDtoH path)
e1.record(); torch.cuda.synchronize()
print(json.dumps({"event_ms": e0.elapsed_time(e1), "wall_s": time.time() - t0, "launches": 4000 * 12}))
'''

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
HOOK = os.path.join(ROOT, "kubeshare_b200", "lib", "libgemhook.so.1")
out = {}
for name, extra in (("unhooked", None), ("ours", {}), ("ours_extra_hooks", {"GEMHOOK_EXTRA_HOOKS": "1"}), ("unhooked2", None)):
    with tempfile.TemporaryDirectory() as tmp:
        env = {k: v for k, v in os.environ.items() if not k.startswith("GEMHOOK_") and k != "LD_PRELOAD"}
        if extra is not None:
            with open(os.path.join(tmp, "q.txt"), "w") as f:
                f.write("1\nbench/c0 1.0 1.0 17179869184\n")
            env.update(LD_PRELOAD=HOOK, GEMHOOK_POOL=os.path.join(tmp, "pool"), GEMHOOK_QUOTA_FILE=os.path.join(tmp, "q.txt"),
                       POD_NAME="bench/c0", GEMHOOK_STATS_FILE=os.path.join(tmp, "st.json"), **extra)
        p = sp.run([sys.executable, "-c", SCRIPT], env=env, stdout=sp.PIPE, stderr=sp.PIPE, timeout=600)
        assert p.returncode == 0, p.stderr.decode()[-1500:]
        r = json.loads(p.stdout.decode().strip().splitlines()[-1])
        if extra is not None:
            st = json.load(open(os.path.join(tmp, "st.json")))
            r["hook"] = {k: st[k] for k in ("launches", "slow_path", "token_requests", "host_syncs", "segments", "gpu_ns")}
        out[name] = r
base = (out["unhooked"]["event_ms"] + out["unhooked2"]["event_ms"]) / 2
for k in ("ours", "ours_extra_hooks"):
    out[k]["overhead_pct"] = (out[k]["event_ms"] - base) / base * 100
print(json.dumps(out))
