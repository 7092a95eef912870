"""Round 2: gemhook_acct_reduce (register double buffer, 16-byte cells) as a function of the number of client slots,
the launch shape, and -- for the one-warp fast path -- the record count.  One JSON object per line.
  python profiles/scripts/r02_sweep.py TAG      (TAG names the build variant, e.g. U8 / U16)"""
import json
import os
import sys

import torch

sys.path.insert(0, ".")
import kubeshare_b200 as kb

tag = sys.argv[1] if len(sys.argv) > 1 else "default"
torch.cuda.init()
torch.zeros(1, device="cuda")
PEAK = 6566.7
try:
    PEAK = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]
except Exception:  # noqa: BLE001
    pass


def records(n, nslots):
    rec = torch.empty((n, 4), dtype=torch.int32, device="cuda")
    idx = torch.arange(n, device="cuda", dtype=torch.int64)
    rec[:, 0] = ((idx * 2654435761) >> 7).remainder(nslots).to(torch.int32)
    rec[:, 1] = 7
    rec[:, 2] = 1000
    rec[:, 3] = 0
    return rec


def run(nslots, n, env=None, reps=8, flush=None):
    for k in ("GEMHOOK_ACCT_WARPS", "GEMHOOK_ACCT_BLOCKS_PER_SM", "GEMHOOK_ACCT_SMALL"):
        os.environ.pop(k, None)
    os.environ.update(env or {})
    rec = records(n, nslots)
    a = kb.Acct(nslots)
    ts = []
    for i in range(reps):
        if flush is not None:
            flush.fill_(i & 0xFF)
            torch.cuda.synchronize()
        ts.append(a.reduce_device(rec.data_ptr(), n, timed=True))
    ts = ts[3:]
    tot, _ = a.totals()
    assert int(tot[:, 2].sum()) == reps * n, (tot[:, 2], reps * n)
    ms = sum(ts) / len(ts)
    out = {"tag": tag, "nslots": nslots, "n": n, "env": env or {}, "ms": round(ms, 5), "min_ms": round(min(ts), 5),
           "gbps": round(16 * n / ms / 1e6, 1), "frac": round(16 * n / ms / 1e6 / PEAK, 4), "grid": a.grid_for(n)}
    a.close()
    del rec
    print(json.dumps(out), flush=True)
    return out


big = 1 << 26
quick = len(sys.argv) > 2 and sys.argv[2] == "quick"
for ns in ((2, 16, 32, 48, 64) if quick else (1, 2, 8, 16, 24, 32, 48, 64)):
    run(ns, big)
if quick:
    run(64, big, {"GEMHOOK_ACCT_WARPS": "5"})
    run(64, big, {"GEMHOOK_ACCT_WARPS": "4"})
    sys.exit(0)
# launch-shape overrides where shared memory is the constraint
for env in ({"GEMHOOK_ACCT_WARPS": "4"}, {"GEMHOOK_ACCT_WARPS": "3", "GEMHOOK_ACCT_BLOCKS_PER_SM": "2"}, {"GEMHOOK_ACCT_WARPS": "2", "GEMHOOK_ACCT_BLOCKS_PER_SM": "3"}):
    run(64, big, env)
for env in ({"GEMHOOK_ACCT_WARPS": "8", "GEMHOOK_ACCT_BLOCKS_PER_SM": "1"}, {"GEMHOOK_ACCT_WARPS": "4", "GEMHOOK_ACCT_BLOCKS_PER_SM": "4"}):
    run(2, big, env)
    run(32, big, env)
# the live hook's regime: tiny flushes, L2 flushed between launches
fl = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for n in (2, 64, 512, 1024):
    for ns in (2, 64):
        run(ns, n, {"GEMHOOK_ACCT_SMALL": "1"}, reps=13, flush=fl)
        run(ns, n, {"GEMHOOK_ACCT_SMALL": "0"}, reps=13, flush=fl)
for n in (1 << 14, 1 << 20):
    run(2, n, reps=13, flush=fl)
