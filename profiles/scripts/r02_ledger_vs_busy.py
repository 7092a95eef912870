"""configs[1] through the three stacks: per client, token time the ledger delivered vs the time the client itself was not
blocked in a launch (gem-storm --track-blocked).  Usage: r02_ledger_vs_busy.py [repetitions]"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import importlib.util
_spec = importlib.util.spec_from_file_location("tp", os.path.join(ROOT, "tests", "test_gpu_parity.py"))
tp = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(tp)
tp._kubeshare_dirs()
W = ["--mode", "storm", "--steps", 32, "--warmup", 0, "--step-launches", 65536, "--sync-every", 1024, "--track-blocked"]
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    for which in ("reference", "ours-tcp", "pool"):
        spans, outs, st, trace = tp.run_arm(which, [0.5, 0.5], W)
        acc = spans.pop("accumulated_ms", None)
        t0 = spans.pop("schd_t0", None)
        if acc is None:
            acc, _ = tp._delivered(spans, outs, t0)
        for c, o in enumerate(outs):
            busy = (o["t_last"] - o["t_first"] - o["blocked_s"]) * 1e3
            print("rep %d %-9s c%d delivered %.1f busy %.1f ratio %.4f blocked %.1f rate_in_token %.1f K/s" % (
                rep, which, c, acc[c], busy, acc[c] / busy, o["blocked_s"] * 1e3, o["launches"] / busy), flush=True)
