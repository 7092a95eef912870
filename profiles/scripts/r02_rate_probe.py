"""Launch-rate probe: the same one-client storm un-hooked (with a host-side delay sweep), under the reference hook,
under ours (TCP and pool), interleaved, several times.  Prints launches/s over the timed steps and the fastest step."""
import json, os, subprocess as sp, sys, tempfile, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import importlib.util
import kubeshare_b200 as kb
_spec = importlib.util.spec_from_file_location("tp", os.path.join(ROOT, "tests", "test_gpu_parity.py"))
tp = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(tp)

W = ["--mode", "storm", "--steps", 12, "--warmup", 2, "--step-launches", 65536, "--sync-every", 1024]


def rate(o):
    return o["launches"] / o["wall_s"], 65536 / min(o["step_s"]), 65536 / statistics.median(o["step_s"])


def unhooked(pace=0, sync_every=1024):
    with tempfile.TemporaryDirectory() as tmp:
        out = os.path.join(tmp, "o.json")
        env = {k: v for k, v in os.environ.items() if not k.startswith("GEMHOOK_") and k != "LD_PRELOAD"}
        w = [str(x) for x in W]
        w[w.index("--sync-every") + 1] = str(sync_every)
        sp.run([kb.STORM_PATH, *w, "--pace-ns", str(pace), "--out", out], env=env, check=True)
        return json.load(open(out))


tp._kubeshare_dirs()
rows = []
for rep in range(3):
    for pace in (0, 20, 40, 60, 80, 120, 160, 240):
        r = rate(unhooked(pace))
        rows.append(("unhooked pace %3d ns" % pace, r))
        print("rep %d unhooked pace %3d ns: %.1f K/s  best step %.1f  median step %.1f" % (rep, pace, r[0] / 1e3, r[1] / 1e3, r[2] / 1e3), flush=True)
    for which in ("reference", "ours-tcp", "pool"):
        spans, outs, st, _ = tp.run_arm(which, [1.0], W)
        r = rate(outs[0])
        print("rep %d %-9s: %.1f K/s  best step %.1f  median step %.1f" % (rep, which, r[0] / 1e3, r[1] / 1e3, r[2] / 1e3), flush=True)
for se in (256, 4096, 65536):
    for pace in (0, 60, 120):
        r = rate(unhooked(pace, se))
        print("sync_every %5d unhooked pace %3d ns: %.1f K/s best %.1f" % (se, pace, r[0] / 1e3, r[1] / 1e3), flush=True)
