"""Roofline of gemhook_acct_reduce as a function of the number of client slots (N = 2^26 records)."""
import json
import sys

import torch

sys.path.insert(0, ".")
import kubeshare_b200 as kb

torch.cuda.init()
torch.zeros(1, device="cuda")
n = 1 << 26
out = {}
for nslots in (1, 2, 4, 8, 16, 20, 21, 32, 64):
    rec = torch.empty((n, 4), dtype=torch.int32, device="cuda")
    idx = torch.arange(n, device="cuda", dtype=torch.int64)
    rec[:, 0] = ((idx * 2654435761) >> 7).remainder(nslots).to(torch.int32)
    rec[:, 1] = 7
    rec[:, 2] = 1000
    rec[:, 3] = 0
    del idx
    a = kb.Acct(nslots)
    ts = [a.reduce_device(rec.data_ptr(), n, timed=True) for _ in range(8)][3:]
    tot, _ = a.totals()
    assert int(tot[:, 2].sum()) == 8 * n
    ms = sum(ts) / len(ts)
    out[nslots] = {"ms": ms, "gbps": 16 * n / ms / 1e6, "grid": a.grid_for(n)}
    a.close()
    del rec
print(json.dumps(out))
