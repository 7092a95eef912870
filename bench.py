#!/usr/bin/env python
"""bench.py -- hook overhead % and launches/s at 1/2/4/8 co-resident clients vs un-hooked (BASELINE.json).

    python bench.py --gpus N --steps K --warmup W            # our hook (libgemhook.so.1, credit pool)
    python bench.py --impl reference --gpus N ...            # the UNMODIFIED reference hook + gem-pmgr + gem-schd
                                                             # (oracle/_ref, built from /root/reference)

A "step" is one pass of the hot path over one batch of synthetic input: every co-resident client issues
STEP_LAUNCHES (65536) launches of noop<<<1,32>>> on the default stream with cuCtxSynchronize every 1024
(SURVEY.md 8d config 2).  The default K=16 is the 1 M-launch storm (2^20 launches per client).  The
headline workload is BASELINE.json configs[1]: 2 clients, gpu_request 0.5 each, on one B200; the 1/2/4/8
client sweep is reported under "clients".  Each client is a separate process (kubeshare_b200/bin/gem-storm,
CUDA driver API) -- exactly how pods share a GPU -- pinned to its own host core.

  value      aggregate hooked launches/s, timed on the device with CUDA events around the K timed steps in
             every client (max over clients and ranks), accounting records resident in the device ring
  e2e        the same launches through the LD_PRELOAD boundary timed on the host clock from the first
             client's start to the last client's end, including the host->device copies of the accounting
             records and the device->host publication of the totals page that happen inside the steps
  roofline   the sm_100a accounting kernel (gemhook_acct_reduce) on a 2^26-record (1 GiB > L2) device-resident
             ring, CUDA events on the accounting stream, algorithmic bytes = 16 B/record
  cpu_baseline  the reference hook stack on the same box / same workload (bounded sample), kind "reference"

Multi-GPU: the path does not shard (one gem-scheduler + hook set per device): --gpus N runs N
independent replicas, one rank per GPU, no collective in the data path; value sums the replicas' launches
over the slowest replica's time.
"""
import argparse
import json
import os
import shutil
import statistics
import subprocess as sp
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
STORM = os.path.join(ROOT, "kubeshare_b200", "bin", "gem-storm")
HOOK = os.path.join(ROOT, "kubeshare_b200", "lib", "libgemhook.so.1")
REFDIR = os.path.join(ROOT, "oracle", "_ref")
STEP_LAUNCHES = int(os.environ.get("GEMBENCH_STEP_LAUNCHES", "65536"))  # override only for the CPU stub tests
SYNC_EVERY = 1024
GIB8 = 8589934592


def log(*a):
    print("[bench]", *a, file=sys.stderr, flush=True)


# --------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi sampling DURING the timed regions (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.gpu = gpu
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = sp.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                  "--format=csv,noheader,nounits", "-lms", "200"], stdout=sp.PIPE, stderr=sp.DEVNULL,
                                 text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except sp.TimeoutExpired:
                self.proc.kill()

    def summary(self, windows):
        sm, mx, reasons = [], 0.0, set()
        for ts, line in self.rows:
            if windows and not any(a <= ts <= b for a, b in windows):
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx = max(mx, float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------------------------------- storms
def quota_text(nclients):
    """Quota file in the order gem-schd READS it: name request(min) limit(max) mem (SURVEY.md 8b trap)."""
    req = 1.0 / nclients
    rows = ["bench/c%d %s 1.0 %d" % (i, repr(req), GIB8) for i in range(nclients)]
    return "%d\n%s\n" % (nclients, "\n".join(rows))


def free_ports(n):
    """n distinct currently-free TCP ports (fresh per run: gem-schd binds without SO_REUSEADDR)."""
    import socket

    socks = []
    for _ in range(n):
        s = socket.socket()
        s.bind(("127.0.0.1", 0))
        socks.append(s)
    ports = [s.getsockname()[1] for s in socks]
    for s in socks:
        s.close()
    return ports


_PHYS = None


def physical_cores():
    """One logical CPU per physical core (first hyper-thread sibling), so pinned clients never share a core."""
    global _PHYS
    if _PHYS is None:
        seen, out = set(), []
        try:
            avail = sorted(os.sched_getaffinity(0))
        except AttributeError:
            avail = list(range(os.cpu_count() or 1))
        for c in avail:
            try:
                sib = open("/sys/devices/system/cpu/cpu%d/topology/thread_siblings_list" % c).read().strip()
            except OSError:
                sib = str(c)
            if sib not in seen:
                seen.add(sib)
                out.append(c)
        _PHYS = out or avail
    return _PHYS


def pin(core):
    """`core` is an index into the list of physical cores (wraps around)."""
    phys = physical_cores()
    cpu = phys[core % len(phys)]

    def f():
        try:
            os.sched_setaffinity(0, {cpu})
        except OSError:
            pass
    return f


def run_clients(nclients, steps, warmup, gpu, mode, core_base, step_launches=STEP_LAUNCHES, timeout=900):
    """mode: 'unhooked' | 'ours' | 'reference'.  Returns dict with per-client results and aggregates."""
    tmp = tempfile.mkdtemp(prefix="gembench_")
    daemons = []
    try:
        env0 = dict(os.environ, CUDA_VISIBLE_DEVICES=str(gpu))
        for k in ("LD_PRELOAD", "GEMHOOK_POOL", "GEMHOOK_QUOTA_FILE", "POD_NAME"):
            env0.pop(k, None)
        with open(os.path.join(tmp, "quota.txt"), "w") as f:
            f.write(quota_text(nclients))
        ports = []
        if mode == "reference":
            # the reference hook hard-codes /kubeshare/library/schedulerIP.txt (reference hook.cpp:162, 233-237)
            os.makedirs("/kubeshare/library", exist_ok=True)
            os.makedirs("/kubeshare/log", exist_ok=True)
            with open("/kubeshare/library/schedulerIP.txt", "w") as f:
                f.write("127.0.0.1\n")
            sport = free_ports(1)[0]
            schd = sp.Popen([os.path.join(REFDIR, "gem-schd"), "-p", tmp, "-f", "quota.txt", "-P", str(sport), "-q", "300",
                             "-m", "20", "-w", "10000"], stdout=sp.DEVNULL, stderr=sp.DEVNULL,
                            preexec_fn=pin(core_base + 2 * nclients))
            daemons.append(schd)
            time.sleep(0.4)
            for i, port in enumerate(free_ports(nclients)):
                ports.append(port)
                e = dict(env0, POD_NAME="bench/c%d" % i, POD_MANAGER_PORT=str(port), SCHEDULER_IP="127.0.0.1",
                         SCHEDULER_PORT=str(sport))
                daemons.append(sp.Popen([os.path.join(REFDIR, "gem-pmgr")], env=e, stdout=sp.DEVNULL, stderr=sp.DEVNULL,
                                        preexec_fn=pin(core_base + nclients + i)))
            time.sleep(0.4)
        procs = []
        for i in range(nclients):
            e = dict(env0)
            if mode == "ours":
                e.update(LD_PRELOAD=HOOK, GEMHOOK_POOL=os.path.join(tmp, "pool"), GEMHOOK_QUOTA_FILE=os.path.join(tmp, "quota.txt"),
                         POD_NAME="bench/c%d" % i, GEMHOOK_STATS_FILE=os.path.join(tmp, "stats.%d.json" % i))
            elif mode == "reference":
                e.update(LD_PRELOAD=os.path.join(REFDIR, "libgemhook_ref.so.1"), POD_NAME="bench/c%d" % i,
                         POD_MANAGER_PORT=str(ports[i]))
            cmd = [STORM, "--mode", "storm", "--steps", str(steps), "--warmup", str(warmup), "--step-launches",
                   str(step_launches), "--sync-every", str(SYNC_EVERY), "--client-id", str(i), "--nclients", str(nclients),
                   "--barrier-dir", tmp, "--out", os.path.join(tmp, "out.%d.json" % i)]
            procs.append(sp.Popen(cmd, env=e, stdout=sp.DEVNULL, stderr=sp.PIPE, preexec_fn=pin(core_base + i)))
        wall0 = time.time()
        errs = []
        for p in procs:
            try:
                _, err = p.communicate(timeout=timeout)
            except sp.TimeoutExpired:
                p.kill()
                _, err = p.communicate()
                errs.append("timeout")
            if p.returncode != 0:
                errs.append("rc=%s %s" % (p.returncode, (err or b"").decode()[-300:]))
        wall1 = time.time()
        if errs:
            raise RuntimeError("%s clients failed: %s" % (mode, errs))
        res = [json.load(open(os.path.join(tmp, "out.%d.json" % i))) for i in range(nclients)]
        stats = []
        if mode == "ours":
            for i in range(nclients):
                try:
                    stats.append(json.load(open(os.path.join(tmp, "stats.%d.json" % i))))
                except (OSError, ValueError):
                    stats.append({})
        launches = sum(r["launches"] for r in res)
        dev_s = max(r["event_ms"] for r in res) / 1e3
        host_s = max(r["t1"] for r in res) - min(r["t0"] for r in res)
        return {"clients": nclients, "launches": launches, "device_s": dev_s, "host_s": host_s,
                "launches_per_s_device": launches / dev_s, "launches_per_s_host": launches / host_s,
                "per_client_wall_s": [r["wall_s"] for r in res], "stats": stats, "window": (wall0, wall1)}
    finally:
        for d in daemons:
            d.kill()
            d.wait()
        shutil.rmtree(tmp, ignore_errors=True)


# --------------------------------------------------------------------------------------------- roofline
def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        return None


def roofline_kernel(steps, warmup, nslots=2):
    """Time gemhook_acct_reduce on a device-resident ring through the C ABI (events on its own stream)."""
    import numpy as np
    import torch

    import kubeshare_b200 as kb

    torch.cuda.init()
    torch.zeros(1, device="cuda")
    out = {}
    acct = kb.Acct(nslots)
    try:
        for label, n in (("ring_2p26", 1 << 26), ("ring_2p20", 1 << 20)):
            rec = torch.empty((n, 4), dtype=torch.int32, device="cuda")
            idx = torch.arange(n, device="cuda", dtype=torch.int64)
            rec[:, 0] = ((idx * 2654435761) >> 7).remainder(nslots).to(torch.int32)
            rec[:, 1] = 1024
            rec[:, 2] = (2_000_000 + (idx % 4096)).to(torch.int32)
            rec[:, 3] = 0
            del idx
            torch.cuda.synchronize()
            flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda") if n < (1 << 24) else None
            times = []
            for i in range(warmup + steps):
                if flush is not None:
                    flush.fill_(i & 0xFF)  # write 256 MiB > 126 MB L2 between timed launches of the small ring
                    torch.cuda.synchronize()
                ms = acct.reduce_device(rec.data_ptr(), n, timed=True)
                if i >= warmup:
                    times.append(ms)
            tot, _ = acct.totals()
            expect = (warmup + steps) * n
            assert int(tot[:, 2].sum()) == expect, "accounting kernel lost records: %s vs %d" % (tot[:, 2], expect)
            acct.reset()
            avg_ms = sum(times) / len(times)
            out[label] = {"records": n, "bytes": 16 * n, "avg_ms": avg_ms, "min_ms": min(times),
                          "gbps": 16 * n / (avg_ms * 1e-3) / 1e9, "grid": acct.grid_for(n), "launches": len(times)}
            del rec
        # CPU path of the same reduction (oracle), bounded sample: 2^24 records
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import orc

        OL = orc.load()
        n_cpu = 1 << 24
        r = np.zeros(n_cpu, np.dtype([("slot", "<u4"), ("launches", "<u4"), ("elapsed_ns", "<u8")]))
        r["slot"] = np.arange(n_cpu, dtype=np.uint32) % nslots
        r["launches"] = 1024
        r["elapsed_ns"] = 2_000_000
        t = time.time(); orc.acct_reduce(OL, r, nslots); t1 = time.time() - t
        threads = min(os.cpu_count() or 1, 64)
        t = time.time(); orc.acct_reduce(OL, r, nslots, threads=threads); tn = time.time() - t
        out["cpu_oracle"] = {"records": n_cpu, "gbps_1_thread": 16 * n_cpu / t1 / 1e9, "threads": threads,
                             "gbps_all_threads": 16 * n_cpu / tn / 1e9}
        out["kernel_launches"] = acct.kernel_launches
    finally:
        acct.close()
    return out


# --------------------------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=16)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--clients", default="1,2,4,8", help="co-resident client counts to sweep")
    ap.add_argument("--headline-clients", type=int, default=2)
    ap.add_argument("--skip-roofline", action="store_true")
    ap.add_argument("--skip-baseline", action="store_true", help="skip the reference cpu_baseline leg")
    ap.add_argument("--only-roofline", action="store_true", help="run just the accounting-kernel leg (for ncu)")
    args = ap.parse_args()
    if args.warmup < 3:
        log("warm-up raised to 3 (timing rules)")
        args.warmup = 3

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    n_gpus_reported = world
    if args.impl == "reference" and world > 1:
        # contract: under torchrun the reference arm runs on rank 0 alone; the other ranks exit without work
        if rank != 0:
            return
        world = 1
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        dist = dist_mod
        dist.init_process_group("gloo")  # host-side gather only: there is no collective in the data path
    gpu = local
    os.environ["CUDA_VISIBLE_DEVICES"] = str(gpu)  # this rank (and torch below) sees only its own GPU

    import __graft_entry__ as ge

    if rank == 0:
        ge.build()
    if dist:
        dist.barrier()

    if args.only_roofline:
        os.environ["CUDA_VISIBLE_DEVICES"] = str(gpu)
        print(json.dumps(roofline_kernel(args.steps, args.warmup)))
        return

    if args.impl == "reference" and not os.path.exists(os.path.join(REFDIR, "libgemhook_ref.so.1")):
        if rank == 0:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref not built (needs /root/reference at build time)"}))
        return

    ncpu = os.cpu_count() or 1
    nphys = len(physical_cores())
    cores_per_rank = max(1, nphys // max(n_gpus_reported, 1))
    core_base = rank * cores_per_rank + 1  # indices into physical_cores(): ranks never share a physical core
    sweep = sorted({int(c) for c in args.clients.split(",") if c} | {args.headline_clients})
    sampler = ClockSampler(gpu)
    sampler.start()
    windows = []
    mode = "ours" if args.impl == "ours" else "reference"
    results = {}
    t_start = time.time()
    for c in sweep:
        if dist:
            dist.barrier()
        un = run_clients(c, args.steps, args.warmup, gpu, "unhooked", core_base)
        if dist:
            dist.barrier()
        hk = run_clients(c, args.steps, args.warmup, gpu, mode, core_base)
        windows += [un["window"], hk["window"]]
        results[c] = {"unhooked": un, "hooked": hk}
        log("rank %d clients=%d unhooked %.0f/s hooked(%s) %.0f/s (device-timed)" % (
            rank, c, un["launches_per_s_device"], mode, hk["launches_per_s_device"]))

    roof = None
    cpu_ref = None
    if args.impl == "ours" and rank == 0:
        if not args.skip_roofline:
            w0 = time.time()
            roof = roofline_kernel(args.steps, args.warmup)
            windows.append((w0, time.time()))
        if os.path.exists(os.path.join(REFDIR, "libgemhook_ref.so.1")) and not args.skip_baseline:
            try:  # bounded sample of the same workload through the reference stack
                k_ref = min(args.steps, 4)
                cpu_ref = run_clients(args.headline_clients, k_ref, 3, gpu, "reference", core_base)
                cpu_ref["sample_steps"] = k_ref
            except Exception as e:  # noqa: BLE001 -- the baseline leg must not take the product arm down
                log("reference baseline leg failed: %r" % (e,))
    sampler.stop()
    clocks = sampler.summary(windows)

    # ---- gather replicas
    mine = {c: {k: {kk: vv for kk, vv in v.items() if kk not in ("window",)} for k, v in r.items()} for c, r in results.items()}
    allr = [mine]
    if dist:
        allr = [None] * world
        dist.all_gather_object(allr, mine)
    if rank != 0:
        return

    def agg(c, which):
        launches = sum(r[c][which]["launches"] for r in allr)
        dev_s = max(r[c][which]["device_s"] for r in allr)
        host_s = max(r[c][which]["host_s"] for r in allr)
        return launches, dev_s, host_s

    sweep_out = {}
    for c in sweep:
        lu, du, hu = agg(c, "unhooked")
        lh, dh, hh = agg(c, "hooked")
        sweep_out[str(c)] = {
            "unhooked_launches_per_s": lu / du, "hooked_launches_per_s": lh / dh, "frac_of_unhooked": (lh / dh) / (lu / du),
            "overhead_pct": (dh - du) / du * 100.0, "hooked_launches_per_s_e2e": lh / hh,
            "unhooked_launches_per_s_e2e": lu / hu, "overhead_pct_e2e": (hh - hu) / hu * 100.0}
    hc = args.headline_clients
    lh, dh, hh = agg(hc, "hooked")
    total_steps_ms = dh * 1e3 / args.steps
    stats = [s for r in allr for s in r[hc]["hooked"].get("stats", [])]
    acct_kernels = sum(s.get("acct_kernels", 0) for s in stats)
    segments = sum(s.get("segments", 0) for s in stats)
    line = {
        "metric": "hooked_launches_per_s", "value": lh / dh, "unit": "launches/s", "n_gpus": n_gpus_reported, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_steps_ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": "configs[1]: %d co-resident clients per B200, gpu_request %.3g each / gpu_limit 1.0, each "
                               "%d x %d noop<<<1,32>>> launches on the default stream, cuCtxSynchronize every %d; quota file "
                               "in gem-schd column order; base/min quota 300/20 ms, window 10 s" % (
                                   hc, 1.0 / hc, args.steps, STEP_LAUNCHES, SYNC_EVERY),
                   "clients": hc, "step_launches": STEP_LAUNCHES, "sync_every": SYNC_EVERY, "transport": "shared credit pool" if mode == "ours" else "tcp gem-pmgr/gem-schd",
                   "parallelism": "replicas x%d (path does not shard)" % world,
                   "l2": "roofline ring 1 GiB > 126 MB L2; 16 MiB ring timed after a 256 MiB L2 flush"},
        "overhead_pct": sweep_out[str(hc)]["overhead_pct"],
        "overhead_pct_single_client_quota_1": sweep_out.get("1", {}).get("overhead_pct"),
        "clients": sweep_out,
        "e2e": {"value": lh / hh, "unit": "launches/s",
                "h2d_bytes_per_step": (16 * segments // max(1, (args.steps + args.warmup))) if mode == "ours" else 0,
                "d2h_bytes_per_step": (acct_kernels * (32 + 24 * hc) // max(1, (args.steps + args.warmup))) if mode == "ours" else 0},
        "hook_stats": {str(c): [{k: s.get(k) for k in ("token_requests", "token_wait_ms", "slow_path", "segments", "acct_kernels", "gpu_ns", "accumulated_token_ms", "quota_ms")}
                                 for r in allr for s in r[c]["hooked"].get("stats", [])] for c in sweep},
        "gpu_launches": int(acct_kernels + (roof or {}).get("kernel_launches", 0)) if mode == "ours" else 0,
        "clocks": clocks, "host": {"cpus": ncpu, "physical_cores": nphys, "client_cores": "one pinned PHYSICAL core per client (no hyper-thread siblings shared), daemons on their own cores"},
        "wall_s": time.time() - t_start,
    }
    if mode == "ours":
        peaks = measured_peaks()
        peak = (peaks or {}).get("hbm_gbs", 6650.0)
        if roof:
            big = roof["ring_2p26"]
            traffic = None
            try:  # dram read+write bytes per launch of the same kernel/size from the committed ncu --set full capture
                tj = json.load(open(os.path.join(ROOT, "profiles", "acct_reduce_traffic.json")))
                if tj.get("records") == big["records"]:
                    traffic = tj["traffic_bytes_per_launch"]
            except (OSError, ValueError, KeyError):
                pass
            line["roofline"] = {"bound": "hbm", "kernel": "gemhook_acct_reduce", "achieved": big["gbps"], "peak": peak,
                                "unit": "GB/s", "frac": big["gbps"] / peak, "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6.65 TB/s (of fallback)",
                                "traffic": traffic, "algorithmic_bytes": big["bytes"], "records": big["records"], "avg_ms": big["avg_ms"], "grid": big["grid"],
                                "ring_2p20": roof["ring_2p20"], "cpu_oracle": roof["cpu_oracle"]}
        if cpu_ref:
            line["cpu_baseline"] = {"value": cpu_ref["launches_per_s_host"], "unit": "launches/s", "kind": "reference",
                                    "cores": 2 * hc + 1, "sample": "%d clients x (3 warm-up + %d timed) steps x %d launches through oracle/_ref libgemhook_ref.so.1 + gem-pmgr + gem-schd" % (
                                        hc, cpu_ref["sample_steps"], STEP_LAUNCHES)}
        else:
            line["cpu_baseline"] = {"value": None, "unit": "launches/s", "kind": "reference", "cores": 0, "sample": "oracle/_ref unavailable"}
    else:
        line["impl"] = "reference"
        line["cpu_baseline"] = {"value": lh / dh, "unit": "launches/s", "kind": "reference", "cores": 2 * hc + 1,
                                "sample": "%d clients x (%d warm-up + %d timed) steps x %d launches" % (hc, args.warmup, args.steps, STEP_LAUNCHES)}
        line["e2e"] = {"value": lh / dh, "unit": "launches/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    print(json.dumps(line))


if __name__ == "__main__":
    try:
        main()
    except Exception as e:  # noqa: BLE001
        if "--impl" in sys.argv and "reference" in sys.argv and int(os.environ.get("RANK", "0")) == 0:
            # contract: the reference arm never fails the driver; say why it could not run
            print(json.dumps({"impl": "reference", "unavailable": "reference stack failed on this box: %r" % (e,)}))
            sys.exit(0)
        raise
