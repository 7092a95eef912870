#!/usr/bin/env python
"""BASELINE.json configs[4]: N x B200, one scheduler per device, 4 clients per device with mixed priority running
MNIST-shaped conv kernels (gem-storm --mode mnist: N=64, 1->32->64 channels, 28x28, 3x3; 100 launches + one DtoH per
iteration).  K8s `priority` never reaches Gemini (only pkg/scheduler uses it, reference pkg/scheduler/pod.go:179-199);
SURVEY.md 8d models priority 0 / 100 as min-fraction 0.1 / 0.4.

Reports, per implementation (ours = shared credit pool, reference = oracle/_ref hook + gem-pmgr + gem-schd):
aggregate launches/s over all devices and Jain's fairness index of delivered/entitled GPU time per client, where
delivered = the client's share of the device-timed run it spent inside its timed region and entitled = its
min-fraction share.

    python kubeshare_b200/tools/scenario.py --gpus 8 --iters 40 [--impl ours|reference|both]
"""
import argparse
import json
import os
import shutil
import subprocess as sp
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (reuses free_ports / pin / paths)

FRACS = [0.1, 0.1, 0.4, 0.4]  # priority 0, 0, 100, 100
WORK = ["--mode", "mnist"]    # overridden by --config 3 (bursty spin-kernel trace, 4 x min-fraction 0.25)


def quota_text():
    rows = ["bench/c%d %s 1.0 %d" % (i, repr(f), bench.GIB8) for i, f in enumerate(FRACS)]
    return "%d\n%s\n" % (len(FRACS), "\n".join(rows))


def run_device(gpu, impl, iters, core_base):
    tmp = tempfile.mkdtemp(prefix="gemcfg5_")
    daemons, procs = [], []
    env0 = {k: v for k, v in os.environ.items() if not k.startswith("GEMHOOK_") and k != "LD_PRELOAD"}
    env0["CUDA_VISIBLE_DEVICES"] = str(gpu)
    with open(os.path.join(tmp, "quota.txt"), "w") as f:
        f.write(quota_text())
    ports = []
    if impl == "reference":
        os.makedirs("/kubeshare/library", exist_ok=True)
        os.makedirs("/kubeshare/log", exist_ok=True)
        with open("/kubeshare/library/schedulerIP.txt", "w") as f:
            f.write("127.0.0.1\n")
        sport = bench.free_ports(1)[0]
        daemons.append(sp.Popen([os.path.join(bench.REFDIR, "gem-schd"), "-p", tmp, "-f", "quota.txt", "-P", str(sport), "-q", "300",
                                 "-m", "20", "-w", "10000"], stdout=sp.DEVNULL, stderr=sp.DEVNULL))
        time.sleep(0.4)
        for i, port in enumerate(bench.free_ports(len(FRACS))):
            ports.append(port)
            e = dict(env0, POD_NAME="bench/c%d" % i, POD_MANAGER_PORT=str(port), SCHEDULER_IP="127.0.0.1", SCHEDULER_PORT=str(sport))
            daemons.append(sp.Popen([os.path.join(bench.REFDIR, "gem-pmgr")], env=e, stdout=sp.DEVNULL, stderr=sp.DEVNULL))
        time.sleep(0.4)
    for i in range(len(FRACS)):
        e = dict(env0)
        if impl in ("ours", "ours-yield"):
            if impl == "ours-yield":
                e["GEMHOOK_YIELD_ON_IDLE"] = "1"
            e.update(LD_PRELOAD=bench.HOOK, GEMHOOK_POOL=os.path.join(tmp, "pool"), GEMHOOK_QUOTA_FILE=os.path.join(tmp, "quota.txt"),
                     POD_NAME="bench/c%d" % i, GEMHOOK_STATS_FILE=os.path.join(tmp, "stats.%d.json" % i))
        elif impl == "reference":
            e.update(LD_PRELOAD=os.path.join(bench.REFDIR, "libgemhook_ref.so.1"), POD_NAME="bench/c%d" % i, POD_MANAGER_PORT=str(ports[i]))
        procs.append(sp.Popen([bench.STORM, *WORK, "--iters", str(iters), "--rounds", str(iters), "--client-id", str(i), "--nclients", str(len(FRACS)),
                               "--barrier-dir", tmp, "--out", os.path.join(tmp, "out.%d.json" % i)], env=e, stderr=sp.PIPE,
                              preexec_fn=bench.pin(core_base + i)))
    return tmp, daemons, procs


def collect(tmp, daemons, procs, impl):
    try:
        for p in procs:
            _, err = p.communicate(timeout=1800)
            if p.returncode != 0:
                raise RuntimeError("client failed: %s" % (err or b"").decode()[-400:])
        res = [json.load(open(os.path.join(tmp, "out.%d.json" % i))) for i in range(len(FRACS))]
        stats = []
        if impl in ("ours", "ours-yield"):
            stats = [json.load(open(os.path.join(tmp, "stats.%d.json" % i))) for i in range(len(FRACS))]
        return res, stats
    finally:
        for d in daemons:
            d.kill()
            d.wait()
        shutil.rmtree(tmp, ignore_errors=True)


def jain(xs):
    s, s2 = sum(xs), sum(x * x for x in xs)
    return s * s / (len(xs) * s2) if s2 else 0.0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--iters", type=int, default=40)
    ap.add_argument("--impl", default="both", choices=["ours", "ours-yield", "reference", "both", "unhooked"])
    ap.add_argument("--config", type=int, default=5, choices=[3, 5],
                    help="5: MNIST-shaped conv, fractions 0.1/0.1/0.4/0.4; 3: bursty trace (--iters = rounds), 4 x 0.25")
    args = ap.parse_args()
    global FRACS, WORK
    if args.config == 3:
        FRACS, WORK = [0.25] * 4, ["--mode", "bursty"]
    sp.check_call([sys.executable, os.path.join(ROOT, "__graft_entry__.py")], stdout=sys.stderr)
    if args.config == 3:
        out = {"config": "configs[2]: %d x B200, 4 clients, gpu_request 0.25 / gpu_limit 1.0, bursty trace seed 0xB200: %d rounds of "
                         "{U{16..4096} launches of a ~5 us spin kernel; sync; sleep Exp(2 ms)}" % (args.gpus, args.iters)}
    else:
        out = {"config": "configs[4]: %d x B200, 4 clients/device, min-fractions %s, limit 1.0, mnist-shaped conv, %d iterations x 100 launches + DtoH"
                         % (args.gpus, FRACS, args.iters)}
    impls = ["unhooked", "ours", "ours-yield", "reference"] if args.impl == "both" else [args.impl]
    ncpu = len(bench.physical_cores())
    for impl in impls:
        if impl == "reference" and not os.path.exists(os.path.join(bench.REFDIR, "libgemhook_ref.so.1")):
            continue
        runs = [run_device(g, impl, args.iters, 1 + g * max(1, ncpu // max(1, args.gpus))) for g in range(args.gpus)]
        per_dev = []
        for g, (tmp, daemons, procs) in enumerate(runs):
            res, stats = collect(tmp, daemons, procs, impl)
            launches = sum(r["launches"] for r in res)
            span = max(r["event_ms"] for r in res) / 1e3
            # delivered GPU time: ours -> device-reduced SM-time of the client; others -> its timed-region wall time share
            if stats:
                delivered = [s["gpu_ns"] / 1e9 for s in stats]
            else:
                delivered = [r["wall_s"] for r in res]
            norm = [d / f for d, f in zip(delivered, FRACS)]
            tok = [{k: s_.get(k) for k in ("token_requests", "slow_path", "token_wait_ms", "segments")} for s_ in stats]
            per_dev.append({"gpu": g, "launches": launches, "span_s": span, "launches_per_s": launches / span, "hook": tok,
                            "client_wall_s": [r["wall_s"] for r in res], "delivered_s": delivered,
                            "jain_delivered_over_entitled": jain(norm), "jain_completion_time": jain([1.0 / r["wall_s"] for r in res])})
        out[impl] = {"aggregate_launches_per_s": sum(d["launches"] for d in per_dev) / max(d["span_s"] for d in per_dev),
                     "per_device": per_dev}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
