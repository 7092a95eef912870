"""Round 2: TMA-staged kernel, variants of the bin update: group size 1/2/4 (independent read-modify-write chains per lane,
same-slot records merged in registers) and 0 = software-pipelined with forwarding.  python profiles/scripts/r02_sweep_staged_ilp.py"""
import json
import os
import sys

import torch

sys.path.insert(0, ".")
import kubeshare_b200 as kb

torch.cuda.init()
torch.zeros(1, device="cuda")
PEAK = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6566.7
KEYS = ("GEMHOOK_ACCT_WARPS", "GEMHOOK_ACCT_BLOCKS_PER_SM", "GEMHOOK_ACCT_SMALL", "GEMHOOK_ACCT_STAGED", "GEMHOOK_ACCT_STAGES", "GEMHOOK_ACCT_STAGED_ILP", "GEMHOOK_ACCT_STAGED_COLS")


def records(n, nslots, seed=0):
    rec = torch.empty((n, 4), dtype=torch.int32, device="cuda")
    idx = torch.arange(n, device="cuda", dtype=torch.int64) + seed
    rec[:, 0] = ((idx * 2654435761) >> 7).remainder(nslots + 1).to(torch.int32)   # (one value out of range: trash row)
    rec[:, 1] = (idx & 0xFFFF).to(torch.int32)
    rec[:, 2] = (idx * 977).to(torch.int32)
    rec[:, 3] = (idx & 3).to(torch.int32)
    return rec


def run(nslots, n, env=None, reps=8, check=None):
    for k in KEYS:
        os.environ.pop(k, None)
    os.environ.update(env or {})
    rec = records(n, nslots)
    a = kb.Acct(nslots)
    ts = [a.reduce_device(rec.data_ptr(), n, timed=True) for _ in range(reps)]
    ts = ts[3:]
    tot, _ = a.totals()
    t = [[int(x) for x in row] for row in tot.tolist()] if hasattr(tot, "tolist") else tot
    ms = sum(ts) / len(ts)
    out = {"nslots": nslots, "n": n, "env": env or {}, "ms": round(ms, 5), "min_ms": round(min(ts), 5),
           "gbps": round(16 * n / ms / 1e6, 1), "frac": round(16 * n / ms / 1e6 / PEAK, 4), "grid": a.grid_for(n)}
    if check is not None:
        out["equal_to_register_staged"] = (t == check)
    a.close()
    del rec
    print(json.dumps(out), flush=True)
    return t


big = 1 << 26
for ns in (32, 48, 64):
    ref = run(ns, big, {"GEMHOOK_ACCT_STAGED": "1", "GEMHOOK_ACCT_STAGED_ILP": "0"})
    for ilp in (0, 2):
        for w in (8, 7, 6):
            run(ns, big, {"GEMHOOK_ACCT_STAGED": "1", "GEMHOOK_ACCT_STAGED_ILP": str(ilp), "GEMHOOK_ACCT_STAGED_COLS": "16",
                          "GEMHOOK_ACCT_WARPS": str(w)}, check=ref)
for n in (513, 4097, (1 << 20) + 77):
    ref = run(64, n, {"GEMHOOK_ACCT_STAGED": "0", "GEMHOOK_ACCT_SMALL": "0"}, reps=4)
    run(64, n, {"GEMHOOK_ACCT_STAGED": "1", "GEMHOOK_ACCT_SMALL": "0", "GEMHOOK_ACCT_STAGED_ILP": "0", "GEMHOOK_ACCT_STAGED_COLS": "16"}, reps=4, check=ref)
