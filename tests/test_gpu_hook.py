"""GPU: the live hook on a real B200 -- device accounting vs host sum, gpu_mem cap with real allocations,
burst-edge token renewals, and accumulated GPU-ms against the reference stack's ledger (+-1 %)."""
import glob
import json
import os
import signal
import subprocess as sp
import tempfile
import time

import pytest

import kubeshare_b200 as kb
import orc
import wireproto as wp

pytestmark = pytest.mark.gpu
REF = os.path.join(kb.ROOT, "oracle", "_ref")
GIB8 = 8589934592


def env_pool(tmp, pod="bench/c0", quota=None, **kw):
    quota = quota or "1\nbench/c0 1.0 1.0 %d\n" % GIB8
    with open(os.path.join(tmp, "quota.txt"), "w") as f:
        f.write(quota)
    env = {k: v for k, v in os.environ.items() if not k.startswith("GEMHOOK_") and k != "LD_PRELOAD"}
    env.update(LD_PRELOAD=kb.LIB_PATH, GEMHOOK_POOL=os.path.join(tmp, "pool"), GEMHOOK_QUOTA_FILE=os.path.join(tmp, "quota.txt"),
               POD_NAME=pod, GEMHOOK_STATS_FILE=os.path.join(tmp, "stats.%d.json"))
    env.update({k: str(v) for k, v in kw.items()})
    return env


def storm(env, *args, timeout=300):
    p = sp.run([kb.STORM_PATH, *map(str, args)], env=env, stdout=sp.PIPE, stderr=sp.PIPE, timeout=timeout)
    assert p.returncode == 0, p.stderr.decode()[-2000:]
    return json.loads(p.stdout)


def stats(tmp):
    return [json.load(open(f)) for f in sorted(glob.glob(os.path.join(tmp, "stats.*.json")))]


@pytest.mark.parametrize("seg", [0, 256])
def test_device_reduced_sm_time_equals_host_sum(seg):
    with tempfile.TemporaryDirectory() as tmp:
        res = storm(env_pool(tmp, GEMHOOK_SEG_LAUNCHES=seg, GEMHOOK_FLUSH_RECORDS=16), "--mode", "storm", "--steps", 4,
                    "--warmup", 1, "--step-launches", 16384, "--sync-every", 1024)
        st = stats(tmp)[0]
        assert st["launches"] == 5 * 16384
        assert st["segments"] >= (20 if seg == 0 else 300)   # bursts younger than GEMHOOK_SEG_MIN_US merge
        assert st["acct_kernels"] >= 2
        assert st["gpu_ns"] == st["gpu_ns_host"] > 0          # sm_100a reduction == host-side sum, bit-exact
        # the bursts cover the storm: SM-time within the wall time, and most of it
        assert 0.5 * res["event_ms"] < st["gpu_ns"] / 1e6 * 4 / 5 < 1.02 * res["event_ms"]


def test_config4_real_allocations_bit_exact_oom_point():
    O = orc.load()
    with tempfile.TemporaryDirectory() as tmp:
        res = storm(env_pool(tmp), "--mode", "memsweep")
    limit, used, first_fail = GIB8, 0, -1
    for row in res["sweep1"]:
        ok = O.orc_mem_prehook_allows(row["bytes"], used, limit)
        assert row["rc"] == (0 if ok else 2), row
        used += row["bytes"] if ok else 0
        if not ok and first_fail < 0:
            first_fail = row["i"]
        assert (row["free"], row["total"]) == (limit - used, limit)
    assert res["first_fail"] == first_fail == 8
    used = 0
    for sz, rc, free in res["sweep2"]:
        ok = O.orc_mem_prehook_allows(sz, used, limit)
        assert rc == (0 if ok else 2)
        used += sz if ok else 0
        assert free == limit - used
    assert res["free_end"] == limit


def test_config3_bursty_tokens_renew_only_at_burst_edges():
    with tempfile.TemporaryDirectory() as tmp:
        quota = "4\n" + "".join("bench/c%d 0.25 1.0 %d\n" % (i, GIB8) for i in range(4))
        procs = []
        for i in range(4):
            env = env_pool(tmp, pod="bench/c%d" % i, quota=quota)
            procs.append(sp.Popen([kb.STORM_PATH, "--mode", "bursty", "--rounds", "150", "--client-id", str(i), "--nclients", "4",
                                   "--barrier-dir", tmp, "--out", os.path.join(tmp, "out%d.json" % i)], env=env, stderr=sp.PIPE))
        for p in procs:
            _, err = p.communicate(timeout=600)
            assert p.returncode == 0, err.decode()[-1000:]
        st = stats(tmp)
        outs = [json.load(open(os.path.join(tmp, "out%d.json" % i))) for i in range(4)]
        assert len(st) == 4
        for s in st:
            o = [x for x in outs if x["launches"] == s["launches"]]
            assert o, "every launch intercepted"
            assert s["slow_path"] <= 150 + s["token_requests"] + 2   # one slow path per burst (+ tracker-forced edges)
            assert s["token_requests"] <= s["slow_path"] + 1          # renewals happen only on the slow path
            assert s["gpu_ns"] == s["gpu_ns_host"] > 0
        # 4 clients x ~150 bursts x ~2000 launches x 5 us: everybody got GPU time, nobody starved
        busy = [s["gpu_ns"] for s in st]
        assert min(busy) > 0.3 * max(busy)


def _ledger_total(path, pod):
    led = json.load(open(path))
    return sum(e["end"] - e["start"] for e in led if e["container"] == pod) * 1e3, len(led)


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "gem-schd-dbg")), reason="oracle/_ref not built")
def test_accumulated_gpu_ms_within_1pct_of_reference_ledger():
    """Same storm through (a) the reference hook and (b) our hook, each against fresh UNMODIFIED gem-pmgr +
    gem-schd(_DEBUG); gem-schd's own ledger dump (scheduler.cpp:693-714) is the judge: sum(end-start)."""
    try:
        os.makedirs("/kubeshare/library", exist_ok=True)
        os.makedirs("/kubeshare/log", exist_ok=True)
        with open("/kubeshare/library/schedulerIP.txt", "w") as f:
            f.write("127.0.0.1\n")
    except OSError:
        pytest.skip("cannot create /kubeshare/library (the reference hook hard-codes it)")
    totals = {}
    for which in ("reference", "ours"):
        with tempfile.TemporaryDirectory() as tmp:
            with open(os.path.join(tmp, "cfg.txt"), "w") as f:
                f.write("1\nbench/c0 1.0 1.0 %d\n" % GIB8)
            sport, pport = wp.free_port(), wp.free_port()
            schd = sp.Popen([os.path.join(REF, "gem-schd-dbg"), "-p", tmp, "-f", "cfg.txt", "-P", str(sport), "-q", "300", "-m",
                             "20", "-w", "10000", "-v", "1"], cwd=tmp, stdout=sp.DEVNULL, stderr=sp.DEVNULL)
            time.sleep(0.5)
            pmgr = sp.Popen([os.path.join(REF, "gem-pmgr")], stdout=sp.DEVNULL, stderr=sp.DEVNULL,
                            env=dict(os.environ, POD_NAME="bench/c0", POD_MANAGER_PORT=str(pport), SCHEDULER_IP="127.0.0.1",
                                     SCHEDULER_PORT=str(sport)))
            time.sleep(0.5)
            try:
                env = {k: v for k, v in os.environ.items() if not k.startswith("GEMHOOK_") and k != "LD_PRELOAD"}
                env.update(POD_NAME="bench/c0", POD_MANAGER_PORT=str(pport))
                if which == "ours":
                    env.update(LD_PRELOAD=kb.LIB_PATH, GEMHOOK_SCHEDULER_IP="127.0.0.1",
                               GEMHOOK_STATS_FILE=os.path.join(tmp, "stats.%d.json"))
                else:
                    env.update(LD_PRELOAD=os.path.join(REF, "libgemhook_ref.so.1"))
                res = storm(env, "--mode", "storm", "--steps", 30, "--warmup", 2, "--step-launches", 65536)
                time.sleep(0.2)
                schd.send_signal(signal.SIGINT)
                schd.wait(timeout=20)
                dumps = glob.glob(os.path.join(tmp, "*.json"))
                dumps = [d for d in dumps if os.path.basename(d)[0].isdigit()]
                assert dumps, "gem-schd did not dump its ledger"
                total_ms, n = _ledger_total(dumps[0], "bench/c0")
                led = [e for e in json.load(open(dumps[0])) if e["container"] == "bench/c0"]
                closed_ms = sum(e["end"] - e["start"] for e in led[:-1]) * 1e3
                extra = stats(tmp)[0] if which == "ours" else {}
                totals[which] = {"ledger_ms": total_ms, "closed_ms": closed_ms, "tokens": n, "wall_ms": res["wall_s"] * 1e3,
                                 "stats": extra}
            finally:
                pmgr.kill()
                schd.kill()
                pmgr.wait()
                schd.wait()
    # third arm: the same storm with the credit pool doing gem-pmgr's + gem-schd's job (no daemon at all)
    with tempfile.TemporaryDirectory() as tmp:
        res = storm(env_pool(tmp), "--mode", "storm", "--steps", 30, "--warmup", 2, "--step-launches", 65536)
        L = kb.lib()
        p = L.gemhook_pool_open(os.path.join(tmp, "pool").encode(), 0, 0, 0, 0, 0)
        n = L.gemhook_pool_history(p, None, None, None, 0)
        import ctypes as C
        sl, a, b = (C.c_int * n)(), (C.c_double * n)(), (C.c_double * n)()
        L.gemhook_pool_history(p, sl, a, b, n)
        L.gemhook_pool_close(p)
        spans = [(a[i], b[i]) for i in range(n)]
        totals["pool"] = {"tokens": n, "closed_ms": sum(e - s for s, e in spans[:-1]), "wall_ms": res["wall_s"] * 1e3}
    ref, ours = totals["reference"], totals["ours"]
    print("ledger totals:", json.dumps(totals))
    # closed tokens only (the last, still-held token is clipped at exit by the pool but kept at full quota by
    # gem-schd): same number of tokens, same closed time within 1 %
    assert totals["pool"]["tokens"] == ref["tokens"] == ours["tokens"]
    assert abs(totals["pool"]["closed_ms"] - ref["closed_ms"]) <= 0.01 * ref["closed_ms"], totals
    # identical launch trace, identical policy, identical daemons: the ledgers agree within 1 %
    assert abs(ours["ledger_ms"] - ref["ledger_ms"]) <= 0.01 * ref["ledger_ms"], totals
    # and our own per-token view of the same quantity matches what gem-schd recorded for us (within 1 %)
    # (closed tokens as the hook saw them + the token still held at exit, which the ledger carries at full quota)
    mine = ours["stats"]["accumulated_token_ms"] + ours["stats"]["quota_ms"]
    assert abs(mine - ours["ledger_ms"]) <= 0.01 * ours["ledger_ms"], (mine, ours)


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "libgemhook_ref.so.1")), reason="oracle/_ref not built")
def test_reference_hook_served_by_gem_arbiter_next_to_a_native_client():
    """Drop-in in the other direction: the UNMODIFIED reference hook speaks TCP to gem-arbiter while a native hook
    arbitrates in the same pool; both finish, both appear in one ledger."""
    try:
        os.makedirs("/kubeshare/library", exist_ok=True)
        os.makedirs("/kubeshare/log", exist_ok=True)
        with open("/kubeshare/library/schedulerIP.txt", "w") as f:
            f.write("127.0.0.1\n")
    except OSError:
        pytest.skip("cannot create /kubeshare/library (the reference hook hard-codes it)")
    arbiter = os.path.join(kb.HERE, "bin", "gem-arbiter")
    with tempfile.TemporaryDirectory() as tmp:
        with open(os.path.join(tmp, "cfg.txt"), "w") as f:
            f.write("2\nbench/native 0.5 1.0 %d\nbench/legacy 0.5 1.0 %d\n" % (GIB8, GIB8))
        port = wp.free_port()
        pool = os.path.join(tmp, "pool")
        arb = sp.Popen([arbiter, "--pool", pool, "-p", tmp, "-f", "cfg.txt", "-P", str(port), "-q", "300", "-m", "20", "-w",
                        "10000"], stderr=sp.DEVNULL)
        time.sleep(0.5)
        try:
            base = {k: v for k, v in os.environ.items() if not k.startswith("GEMHOOK_") and k != "LD_PRELOAD"}
            e_native = dict(base, LD_PRELOAD=kb.LIB_PATH, GEMHOOK_POOL=pool, POD_NAME="bench/native",
                            GEMHOOK_STATS_FILE=os.path.join(tmp, "stats.%d.json"))
            e_legacy = dict(base, LD_PRELOAD=os.path.join(REF, "libgemhook_ref.so.1"), POD_NAME="bench/legacy",
                            POD_MANAGER_PORT=str(port))
            args = ["--mode", "storm", "--steps", "6", "--warmup", "1", "--step-launches", "65536", "--nclients", "2",
                    "--barrier-dir", tmp]
            procs = [sp.Popen([kb.STORM_PATH, *args, "--client-id", "0", "--out", os.path.join(tmp, "o0.json")], env=e_native, stderr=sp.PIPE),
                     sp.Popen([kb.STORM_PATH, *args, "--client-id", "1", "--out", os.path.join(tmp, "o1.json")], env=e_legacy, stderr=sp.PIPE)]
            for p in procs:
                _, err = p.communicate(timeout=300)
                assert p.returncode == 0, err.decode()[-1500:]
            outs = [json.load(open(os.path.join(tmp, "o%d.json" % i))) for i in range(2)]
            assert all(o["launches"] == 6 * 65536 for o in outs)
            L = kb.lib()
            p = L.gemhook_pool_open(pool.encode(), 0, 0, 0, 0, 0)
            acc = {n: L.gemhook_pool_accumulated_ms(p, L.gemhook_pool_find(p, n.encode())) for n in ("bench/native", "bench/legacy")}
            L.gemhook_pool_close(p)
            print("ledger:", acc)
            assert all(v > 100.0 for v in acc.values()), acc
        finally:
            arb.kill()
            arb.wait()


def test_modern_entry_points_on_real_driver():
    """cuLaunchKernelEx + cuMemAllocAsync/cuMemFreeAsync + cuStreamSynchronize against the real libcuda."""
    with tempfile.TemporaryDirectory() as tmp:
        res = storm(env_pool(tmp, quota="1\nbench/c0 1.0 1.0 5000\n", GEMHOOK_EXTRA_HOOKS=1), "--mode", "modern")
        st = stats(tmp)[0]
    assert st["launches"] == 100 and res["rc"] == [0, 0, 2]
    assert (res["free"], res["free_after"], res["total"]) == (2000, 4000, 5000)
    assert st["slow_path"] >= 5 and st["gpu_ns"] == st["gpu_ns_host"]
    # virtual memory management (PyTorch expandable segments): cuMemCreate is charged, cuMemRelease gives it back
    cap = 64 << 20
    with tempfile.TemporaryDirectory() as tmp:
        res = storm(env_pool(tmp, quota="1\nbench/c0 1.0 1.0 %d\n" % cap), "--mode", "modern")
    g = res["vmm"]["gran"]
    assert g >= 4096 and res["vmm"]["rc"] == [0, 2]                 # one granule fits, 4096 granules do not
    assert res["vmm"]["free_held"] == cap - 1000 - 3000 - g        # a (1000) and c (3000) are still allocated
    assert res["vmm"]["free_released"] == cap - 1000 - 3000


def test_pytorch_application_under_the_hook():
    """A real cudart application (PyTorch): cudart binds the driver through dlsym + cuGetProcAddress_v2, both
    interposed.  The gpu_mem cap shows up as torch's OOM, mem_get_info is virtualised, every kernel passes the gate."""
    import sys

    script = r'''
import json, sys, torch
x = torch.ones(1 << 20, device="cuda")
for _ in range(1500):
    x = x + 1
torch.cuda.synchronize()
free, total = torch.cuda.mem_get_info()
big = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")      # 1 GiB fits under the 2 GiB cap
oom = False
try:
    torch.empty(3 << 30, dtype=torch.uint8, device="cuda")          # 3 GiB does not
except torch.cuda.OutOfMemoryError:
    oom = True
print(json.dumps({"x0": float(x[0]), "total": total, "free_le_total": free <= total, "oom": oom}))
'''
    with tempfile.TemporaryDirectory() as tmp:
        env = env_pool(tmp, quota="1\nbench/c0 1.0 1.0 %d\n" % (2 << 30))
        p = sp.run([sys.executable, "-c", script], env=env, stdout=sp.PIPE, stderr=sp.PIPE, timeout=300)
        assert p.returncode == 0, p.stderr.decode()[-2000:]
        res = json.loads(p.stdout.decode().strip().splitlines()[-1])
        st = stats(tmp)[0]
    assert res["x0"] == 1501.0 and res["oom"] is True
    assert res["total"] == 2 << 30 and res["free_le_total"]
    assert st["launches"] >= 1500 and st["allocs_denied"] >= 1
    assert st["gpu_ns"] == st["gpu_ns_host"] > 0


def test_binding_paths_on_real_driver():
    """direct symbol, dlsym, cuGetProcAddress (legacy stream and per-thread stream) against the real libcuda."""
    with tempfile.TemporaryDirectory() as tmp:
        res = storm(env_pool(tmp, quota="1\nbench/c0 1.0 1.0 5000\n"), "--mode", "resolve")
        st = stats(tmp)[0]
    assert st["launches"] == 40 and res["gpa_is_hooked"] == 1 and res["ptsz_distinct"] == 1
    assert res["rc"] == [0, 0, 2] and (res["free"], res["total"]) == (2000, 5000)
