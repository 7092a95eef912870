"""8-column bins at large slot counts: columns x warps x shared-memory cap (N = 2^26)."""
import itertools, json, os, sys
import torch
sys.path.insert(0, ".")
import kubeshare_b200 as kb
torch.cuda.init(); torch.zeros(1, device="cuda")
n = 1 << 26
for nslots in (20, 32, 48, 64):
    rec = torch.empty((n, 4), dtype=torch.int32, device="cuda")
    idx = torch.arange(n, device="cuda", dtype=torch.int64)
    rec[:, 0] = ((idx * 2654435761) >> 7).remainder(nslots).to(torch.int32)
    rec[:, 1] = 7; rec[:, 2] = 1000; rec[:, 3] = 0
    del idx
    best = None
    for cols, warps, cap in itertools.product((8, 16), (8, 4, 2), (96, 128, 160, 208)):
        if warps * nslots * cols * 20 > 220 * 1024:
            continue
        os.environ.update(GEMHOOK_ACCT_COLS=str(cols), GEMHOOK_ACCT_WARPS=str(warps), GEMHOOK_ACCT_SMEM_CAP_KB=str(cap))
        a = kb.Acct(nslots)
        ts = [a.reduce_device(rec.data_ptr(), n, timed=True) for _ in range(6)][2:]
        tot, _ = a.totals()
        assert int(tot[:, 2].sum()) == 6 * n
        g = 16 * n / (sum(ts) / len(ts)) / 1e6
        row = {"nslots": nslots, "cols": cols, "warps": warps, "cap": cap, "grid": a.grid_for(n), "gbps": round(g, 1)}
        print(json.dumps(row), flush=True)
        if best is None or g > best["gbps"]:
            best = row
        a.close()
    print("BEST", json.dumps(best), flush=True)
    del rec
