"""CPU: properties of the lock-free credit pool beyond the policy goldens (tests/test_pool_policy.py): slot reuse and
over-long names on quota-file reload, the gpu_mem rule when a reload lowers the limit below what is in use,
processes of one pod sharing a token (no release while a sibling lives, one request in flight per pod), observers that
never disturb the state, recycling of what dead processes leave behind."""
import ctypes as C
import os
import signal
import subprocess as sp
import sys
import tempfile
import threading
import time

import kubeshare_b200 as kb

L = kb.lib()


def counters(p):
    a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
    L.gemhook_pool_counters(p, C.byref(a), C.byref(b), C.byref(c))
    return a.value, b.value, c.value


def test_slots_of_departed_pods_are_reused_when_the_table_is_full():
    """the 33rd rewrite of a 2-pod file used to fail with 'more than 64 clients'."""
    p = L.gemhook_pool_open(None, 1, 300.0, 20.0, 10000.0, 1)
    for gen in range(200):
        text = "2\nns/gen%d-a 0.5 1.0 1000\nns/gen%d-b 0.5 1.0 2000\n" % (gen, gen)
        assert L.gemhook_pool_load_config(p, text.encode(), 0) == 2, (gen, kb.last_error())
        a = L.gemhook_pool_find(p, ("ns/gen%d-a" % gen).encode())
        assert 0 <= a < 64 and L.gemhook_pool_nslots(p) <= 64
        u, lim = C.c_uint64(), C.c_uint64()
        L.gemhook_pool_mem_info(p, a, C.byref(u), C.byref(lim))
        assert (u.value, lim.value) == (0, 1000)
    assert L.gemhook_pool_find(p, b"ns/gen0-a") == -1          # long gone, its slot serves somebody else
    # a pod that still holds memory, a token or a request is never evicted
    busy = L.gemhook_pool_find(p, b"ns/gen199-a")
    assert L.gemhook_pool_mem_reserve(p, busy, 10) == 1
    for gen in range(200, 300):
        L.gemhook_pool_load_config(p, ("1\nns/other%d 1.0 1.0 5\n" % gen).encode(), 0)
    assert L.gemhook_pool_find(p, b"ns/gen199-a") == busy
    L.gemhook_pool_close(p)


def test_over_long_pod_name_skips_that_row_only():
    p = L.gemhook_pool_open(None, 1, 300.0, 20.0, 10000.0, 1)
    long_name = "namespace-" + "x" * 70 + "/pod"       # 'namespace/name' as pkg/scheduler/pod.go:455 injects it can exceed 63 bytes
    text = "3\nns/a 0.5 1.0 10\n%s 0.5 1.0 20\nns/c 0.25 1.0 30\n" % long_name
    assert L.gemhook_pool_load_config(p, text.encode(), 0) == 3   # the file is well-formed: the other rows are loaded
    assert L.gemhook_pool_find(p, b"ns/a") == 0 and L.gemhook_pool_find(p, b"ns/c") == 1 and L.gemhook_pool_nslots(p) == 2
    assert b"longer than" in L.gemhook_last_error()
    L.gemhook_pool_close(p)


def test_limit_lowered_below_usage_denies_instead_of_wrapping():
    """limit 1000, reserve 800, reload with limit 500 -> a reserve of 10^12 used to be granted."""
    p = L.gemhook_pool_open(None, 1, 300.0, 20.0, 10000.0, 1)
    L.gemhook_pool_load_config(p, b"1\nns/a 1.0 1.0 1000\n", 0)
    assert L.gemhook_pool_mem_reserve(p, 0, 800) == 1
    L.gemhook_pool_load_config(p, b"1\nns/a 1.0 1.0 500\n", 0)
    assert L.gemhook_pool_mem_reserve(p, 0, 10 ** 12) == 0 and L.gemhook_pool_mem_reserve(p, 0, 1) == 0
    L.gemhook_pool_mem_release(p, 0, 800)
    assert L.gemhook_pool_mem_reserve(p, 0, 500) == 1 and L.gemhook_pool_mem_reserve(p, 0, 1) == 0
    L.gemhook_pool_close(p)


def test_processes_of_one_pod_share_the_token():
    """release is per pod, not per process; one request per pod in flight."""
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "pool").encode()
        a = L.gemhook_pool_open(path, 1, 300.0, 20.0, 10000.0, 0)
        b = L.gemhook_pool_open(path, 0, 0, 0, 0, 0)
        c = L.gemhook_pool_open(path, 0, 0, 0, 0, 0)
        L.gemhook_pool_load_config(a, b"2\nns/pod 0.5 1.0 100\nns/other 0.5 1.0 100\n", 0)
        assert L.gemhook_pool_attach(a, 0) >= 0 and L.gemhook_pool_attach(b, 0) >= 0 and L.gemhook_pool_attach(c, 1) >= 0
        fwd = C.c_int()
        assert L.gemhook_pool_acquire_ex(a, 0, 0.0, 0.0, C.byref(fwd)) == 300.0 and fwd.value == 1
        # the sibling is answered from the pod's token (remaining quota), not from the scheduler
        q = L.gemhook_pool_acquire_ex(b, 0, 0.0, 1.0, C.byref(fwd))
        assert fwd.value == 0 and 0 < q <= 300.0
        info = kb.SlotInfo()
        L.gemhook_pool_release(b, 0)                       # sibling `a` is alive: the pod keeps its token
        L.gemhook_pool_slot_info(a, 0, C.byref(info))
        assert info.holds_token == 1
        # the other pod asks meanwhile and has to wait for the deadline ...
        got = {}
        t = threading.Thread(target=lambda: got.setdefault("q", L.gemhook_pool_acquire(c, 1, 0.0, 0.0)))
        t0 = time.time()
        t.start()
        time.sleep(0.05)
        assert "q" not in got
        L.gemhook_pool_detach(b)
        L.gemhook_pool_release(a, 0)                       # ... or for the LAST process of the pod to hand it back
        t.join(timeout=5)
        assert got.get("q") == 300.0 and time.time() - t0 < 0.25
        for h in (c, b, a):
            L.gemhook_pool_close(h)


def test_observers_do_not_commit_anything():
    p = L.gemhook_pool_open(None, 1, 300.0, 20.0, 1000.0, 1)
    L.gemhook_pool_load_config(p, b"2\nA 0.5 1.0 1\nB 0.5 1.0 1\n", 0)
    who, q, slp = C.c_int(), C.c_double(), C.c_double()
    L.gemhook_pool_request(p, 0, 10.0, 0.0, 0.0)
    assert L.gemhook_pool_schedule(p, 10.0, C.byref(who), C.byref(q), C.byref(slp)) == 1
    before = counters(p)[0]
    info = kb.SlotInfo()
    for _ in range(100):
        L.gemhook_pool_slot_info(p, 0, C.byref(info))
        L.gemhook_pool_usage(p, 0, 5000.0)                 # would prune the ledger if it were published
        L.gemhook_pool_history(p, None, None, None, 0)
        L.gemhook_pool_accumulated_ms(p, 0)
        L.gemhook_pool_find(p, b"B")
        L.gemhook_pool_others_waiting(p, 0)
        assert L.gemhook_pool_schedule(p, 20.0, C.byref(who), C.byref(q), C.byref(slp)) == -2   # a decision that changes nothing
    assert counters(p)[0] == before
    assert L.gemhook_pool_history(p, None, None, None, 0) == 1
    L.gemhook_pool_close(p)


def test_what_a_killed_observer_leaves_behind_is_recycled():
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "pool")
        p = L.gemhook_pool_open(path.encode(), 1, 300.0, 20.0, 10000.0, 0)
        L.gemhook_pool_load_config(p, b"1\nA 1.0 1.0 1000\n", 0)
        child = sp.Popen([sys.executable, "-c", """
import sys, time
sys.path.insert(0, %r)
import kubeshare_b200 as kb
L = kb.lib()
p = L.gemhook_pool_open(%r.encode(), 0, 0, 0, 0, 0)
print("ready", flush=True)
time.sleep(60)
""" % (kb.ROOT, path)], stdout=sp.PIPE)
        assert child.stdout.readline().strip() == b"ready"
        child.send_signal(signal.SIGKILL)
        child.wait()
        assert L.gemhook_pool_reap(p) == 0                 # no CLIENT died ...
        # ... and the tool's liveness entry is free again: 300 more handles fit into the 256-entry table one after another
        for _ in range(300):
            h = L.gemhook_pool_open(path.encode(), 0, 0, 0, 0, 0)
            assert h
            L.gemhook_pool_close(h)
        L.gemhook_pool_close(p)


def test_racing_clients_reload_a_quota_file_version_exactly_once():
    """Eight clients notice the same (re)written quota file at the same moment: one of them reloads, the others see that
    the configuration already carries that file's stamp -- a second reload would reset every client's adaptive quota
    again and show up in the quota sequence (it did, once, in the EMA-replay test on the GPU box)."""
    with tempfile.TemporaryDirectory() as tmp:
        qf = os.path.join(tmp, "quota.txt")
        with open(qf, "w") as f:
            f.write("2\nns/a 0.5 1.0 10\nns/b 0.5 1.0 20\n")
        path = os.path.join(tmp, "pool").encode()
        handles = [L.gemhook_pool_open(path, 1, 300.0, 20.0, 10000.0, 0) for _ in range(8)]
        for round_ in range(5):
            if round_:
                time.sleep(0.01)
                with open(qf, "w") as f:
                    f.write("2\nns/a 0.5 1.0 %d\nns/b 0.5 1.0 20\n" % (10 + round_))
            before = counters(handles[0])[0]
            out = [None] * 8
            go = threading.Barrier(8)

            def work(i):
                go.wait()
                out[i] = L.gemhook_pool_sync_quota_file(handles[i], qf.encode(), 0)

            th = [threading.Thread(target=work, args=(i,)) for i in range(8)]
            for t in th:
                t.start()
            for t in th:
                t.join()
            assert sorted(out) == [0] * 7 + [1], out
            assert counters(handles[0])[0] == before + 1       # one committed transaction
            u, lim = C.c_uint64(), C.c_uint64()
            L.gemhook_pool_mem_info(handles[3], 0, C.byref(u), C.byref(lim))
            assert lim.value == 10 + round_
        for h in handles:
            L.gemhook_pool_close(h)


def test_pool_clock_places_ledger_entries_on_the_readers_clock():
    """gemhook_pool_now_ms is the ledger's time base (ms since the pool was created, CLOCK_MONOTONIC -- gem-schd's
    ms_since_start, scheduler.cpp:107-109): a reader that notes its own monotonic clock next to it can tell when a ledger entry
    began and ended.  Two handles on one file read the same clock."""
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "pool").encode()
        p = L.gemhook_pool_open(path, 1, 300.0, 20.0, 10000.0, 0)
        q = L.gemhook_pool_open(path, 0, 0, 0, 0, 0)              # an observer on the same file
        assert L.gemhook_pool_load_config(p, b"1\nns/a 1.0 1.0 1000\n", 0) == 1
        t0 = time.monotonic() - L.gemhook_pool_now_ms(p) / 1e3   # the pool's origin on this process's clock
        a, b = L.gemhook_pool_now_ms(p), L.gemhook_pool_now_ms(q)
        assert 0.0 <= a <= b <= a + 50.0
        before = (time.monotonic() - t0) * 1e3
        quota = L.gemhook_pool_acquire(p, 0, 0.0, 0.0)
        after = (time.monotonic() - t0) * 1e3
        assert quota > 0
        n = L.gemhook_pool_history(q, None, None, None, 0)
        sl, s, e = (C.c_int * n)(), (C.c_double * n)(), (C.c_double * n)()
        L.gemhook_pool_history(q, sl, s, e, n)
        assert n == 1 and before - 1.0 <= s[0] <= after + 1.0, (before, s[0], after)   # granted between the two stamps
        assert e[0] == s[0] + quota                                                     # Record(): end = start + quota
        L.gemhook_pool_close(q)
        L.gemhook_pool_close(p)
