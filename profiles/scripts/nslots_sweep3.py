"""Compare the bin-column kernels with the match/redux variant across slot counts (N = 2^26)."""
import json
import os
import sys

import torch

sys.path.insert(0, ".")
import kubeshare_b200 as kb

torch.cuda.init()
torch.zeros(1, device="cuda")
n = 1 << 26
import numpy as np
for nslots in (1, 2, 8, 16, 20, 32, 64):
    rec = torch.empty((n, 4), dtype=torch.int32, device="cuda")
    idx = torch.arange(n, device="cuda", dtype=torch.int64)
    rec[:, 0] = ((idx * 2654435761) >> 7).remainder(nslots).to(torch.int32)
    rec[:, 1] = (idx % 70001).to(torch.int32)
    rec[:, 2] = (idx % 1000003).to(torch.int32)
    rec[:, 3] = (idx % 5).to(torch.int32)
    del idx
    row = {"nslots": nslots}
    ref = None
    for kern in ("cols", "mr")  # "mr" = the removed match/redux variant (GEMHOOK_ACCT_KERNEL is ignored now):
        os.environ.pop("GEMHOOK_ACCT_KERNEL", None)
        if kern == "mr":
            os.environ["GEMHOOK_ACCT_KERNEL"] = "mr"
        a = kb.Acct(nslots)
        ts = [a.reduce_device(rec.data_ptr(), n, timed=True) for _ in range(6)][2:]
        tot, _ = a.totals()
        if ref is None:
            ref = tot.copy()
        else:
            assert (tot == ref).all(), "mr kernel disagrees with the column kernel"
        row[kern] = round(16 * n / (sum(ts) / len(ts)) / 1e6, 1)
        row[kern + "_grid"] = a.grid_for(n)
        a.close()
    print(json.dumps(row), flush=True)
    del rec
