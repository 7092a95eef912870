"""CPU: the shared credit pool's token policy and gpu_mem counter (csrc/gh_pool.cpp through the C ABI)
against golden ledgers from the reference's scheduler.o under a virtual clock, the live gem-schd known
answers, the live gem-pmgr memory counter, and the oracle.  Bit-exact."""
import ctypes as C
import json
import os

import pytest

import kubeshare_b200 as kb
import orc

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "ref_golden.json")))


def _ms(t_ns, start_ns):
    return ((t_ns - start_ns) // 1000) / 1e3


def history(L, p, names):
    n = L.gemhook_pool_history(p, None, None, None, 0)
    s, a, b = (C.c_int * n)(), (C.c_double * n)(), (C.c_double * n)()
    L.gemhook_pool_history(p, s, a, b, n)
    return [[names[s[i]], a[i], b[i]] for i in range(n)]


@pytest.mark.parametrize("sc", G["schd"], ids=lambda s: "seed%d" % s["seed"])
def test_pool_policy_reproduces_reference_ledger(sc):
    L = kb.lib()
    p = L.gemhook_pool_open(None, 1, sc["base"], sc["min"], sc["window"], sc["start_ns"])
    assert p
    assert L.gemhook_pool_load_config(p, sc["config_text"].encode(), 0) == len(sc["clients"])
    names = {}
    for name, mn, mx, mem in sc["clients"]:
        idx = L.gemhook_pool_find(p, name.encode())
        assert idx >= 0
        names[idx] = name
        u, lim = C.c_uint64(), C.c_uint64()
        L.gemhook_pool_mem_info(p, idx, C.byref(u), C.byref(lim))
        assert (u.value, lim.value) == (0, mem)
    slot_of = {v: k for k, v in names.items()}
    start = sc["start_ns"]
    who, q, slp = C.c_int(), C.c_double(), C.c_double()
    sleeps = 0
    for i, st in enumerate(sc["steps"]):
        now = _ms(st["t_ns"], start)
        for name, overuse, burst in st["requests"]:
            assert L.gemhook_pool_request(p, slot_of[name], now, overuse, burst) == 0
        if st["selected"] is not None:
            wake = list(st["wakeups_ns"])
            # the virtual-clock driver calls select_candidate back to back; the daemon's wait for the previous
            # holder (scheduler.cpp:501-521) is not part of it -> declare the previous token timed out
            L.gemhook_pool_expire_token(p)
            while True:
                rc = L.gemhook_pool_schedule(p, now, C.byref(who), C.byref(q), C.byref(slp))
                if rc == 1:
                    break
                assert rc == 0 and wake, "step %d: rc=%d" % (i, rc)
                nxt = wake.pop(0)
                now = _ms(nxt, start)
                sleeps += 1
            assert not wake
            assert names[who.value] == st["selected"], "step %d" % i
            assert q.value == st["quota"], "step %d" % i
        if "history" in st:
            assert history(L, p, names) == st["history"], "step %d" % i
    L.gemhook_pool_close(p)


def test_live_schd_known_answers():
    g = G["live_schd"]
    L = kb.lib()
    p = L.gemhook_pool_open(None, 1, g["base"], g["min"], g["window"], 1)
    L.gemhook_pool_load_config(p, g["config"].encode(), 0)
    a = L.gemhook_pool_find(p, b"ns/a")
    who, q, slp = C.c_int(), C.c_double(), C.c_double()
    now = 0.0
    for c in g["calls"]:
        if c["op"] == "mem_limit":
            u, lim = C.c_uint64(), C.c_uint64()
            L.gemhook_pool_mem_info(p, a, C.byref(u), C.byref(lim))
            assert (u.value, lim.value) == (c["used"], c["total"])
        elif c["op"] == "quota":
            now += 10.0
            L.gemhook_pool_request(p, a, now, c["overuse"], c["burst"])
            assert L.gemhook_pool_schedule(p, now, C.byref(who), C.byref(q), C.byref(slp)) == 1
            assert q.value == c["quota"]
    L.gemhook_pool_close(p)


def test_one_outstanding_token_per_gpu():
    """scheduler.cpp:501-521: after a grant nobody else is served until the holder returns or times out."""
    L = kb.lib()
    p = L.gemhook_pool_open(None, 1, 300.0, 20.0, 10000.0, 1)
    L.gemhook_pool_load_config(p, b"2\nA 0.5 1.0 1\nB 0.5 1.0 1\n", 0)
    who, q, slp = C.c_int(), C.c_double(), C.c_double()
    L.gemhook_pool_request(p, 0, 100.0, 0.0, 0.0)
    L.gemhook_pool_request(p, 1, 100.0, 0.0, 0.0)
    assert L.gemhook_pool_schedule(p, 100.0, C.byref(who), C.byref(q), C.byref(slp)) == 1 and who.value == 0
    assert L.gemhook_pool_schedule(p, 150.0, C.byref(who), C.byref(q), C.byref(slp)) == -2
    assert slp.value == 250.0  # until A's deadline (100 + 300)
    # A returns early at t=200: B is served at once
    L.gemhook_pool_request(p, 0, 200.0, 0.0, 50.0)
    assert L.gemhook_pool_schedule(p, 200.0, C.byref(who), C.byref(q), C.byref(slp)) == 1 and who.value == 1
    # B never returns: A is served after B's deadline
    assert L.gemhook_pool_schedule(p, 499.0, C.byref(who), C.byref(q), C.byref(slp)) == -2
    assert L.gemhook_pool_schedule(p, 500.0, C.byref(who), C.byref(q), C.byref(slp)) == 1 and who.value == 0
    assert L.gemhook_pool_schedule(p, 500.0, C.byref(who), C.byref(q), C.byref(slp)) == -2
    # release hands the token back without waiting for the deadline
    L.gemhook_pool_release(p, 0)
    assert L.gemhook_pool_schedule(p, 501.0, C.byref(who), C.byref(q), C.byref(slp)) == -1  # nobody waiting
    L.gemhook_pool_close(p)


def test_accumulated_gpu_ms_matches_oracle_full_history():
    """'Accumulated GPU-ms' (SURVEY.md 8a): sum over the FULL ledger of end-start per client."""
    import random

    L, O = kb.lib(), orc.load()
    p = L.gemhook_pool_open(None, 1, 300.0, 20.0, 2000.0, 1)
    h = O.orc_schd_new(300.0, 20.0, 2000.0)
    cfg = b"3\nA 0.2 0.6 1\nB 0.3 1.0 1\nC 0.1 0.5 1\n"
    L.gemhook_pool_load_config(p, cfg, 0)
    O.orc_schd_load_config(h, cfg)
    rng = random.Random(5)
    now = 0.0
    who, q, slp = C.c_int(), C.c_double(), C.c_double()
    nm = C.create_string_buffer(64)
    names = [b"A", b"B", b"C"]
    holder = None
    for _ in range(800):
        now += rng.choice([0.5, 3.0, 20.0, 90.0, 400.0]) * rng.random()
        k = holder if (holder is not None and rng.random() < 0.7) else rng.randrange(3)
        over, burst = rng.choice([0.0, 0.0, 1.5]), rng.choice([0.0, 30.0, 250.0, 900.0])
        L.gemhook_pool_request(p, k, now, over, burst)
        O.orc_schd_request(h, names[k], now, over, burst)
        while True:
            rc = L.gemhook_pool_schedule(p, now, C.byref(who), C.byref(q), C.byref(slp))
            if rc == -2:  # token outstanding: the oracle models only select_candidate, so wait it out
                now += slp.value
                continue
            orc_rc = O.orc_schd_select(h, now, nm, C.byref(slp)) if rc in (0, 1) else None
            if rc == 0:
                assert orc_rc == 0
                now += max(slp.value, 0.001)
                continue
            break
        if rc == 1:
            assert orc_rc == 1 and nm.value == names[who.value]
            assert O.orc_schd_grant(h, nm.value, now) == q.value
            holder = who.value
    for k in range(3):
        assert L.gemhook_pool_accumulated_ms(p, k) == pytest.approx(O.orc_schd_accumulated_ms(h, names[k]), rel=1e-12)
        assert L.gemhook_pool_usage(p, k, now) == O.orc_schd_usage(h, names[k], now)
    L.gemhook_pool_close(p)
    O.orc_schd_free(h)


def test_gpu_mem_counter_matches_live_reference_pmgr():
    g = G["live_pmgr_mem"]
    L = kb.lib()
    p = L.gemhook_pool_open(None, 1, 300.0, 20.0, 10000.0, 1)
    L.gemhook_pool_load_config(p, ("1\nns/pod 1.0 1.0 %d\n" % g["limit"]).encode(), 0)
    u, lim = C.c_uint64(), C.c_uint64()
    held = {0: 0, 1: 0}
    for op in g["ops"]:
        if op["op"] == "alloc":
            L.gemhook_pool_mem_info(p, 0, C.byref(u), C.byref(lim))
            assert (u.value, lim.value) == (op["used_before"], op["total"])
            ok = L.gemhook_pool_mem_reserve(p, 0, op["bytes"])
            assert ok == op["verdict"]
            if ok:
                held[op["conn"]] += op["bytes"]
        elif op["op"] == "free":
            L.gemhook_pool_mem_release(p, 0, op["bytes"])
            held[op["conn"]] -= op["bytes"]
        else:  # a connection closes: its bytes are reclaimed (pod-manager.cpp:533-545)
            L.gemhook_pool_mem_release(p, 0, held[op["conn"]])
        L.gemhook_pool_mem_info(p, 0, C.byref(u), C.byref(lim))
        assert u.value == op["used_after"]
    L.gemhook_pool_close(p)


def test_mem_cap_edge_cases_bit_exact():
    L, O = kb.lib(), orc.load()
    limit = 8589934592
    p = L.gemhook_pool_open(None, 1, 300.0, 20.0, 10000.0, 1)
    L.gemhook_pool_load_config(p, ("1\nx 1.0 1.0 %d\n" % limit).encode(), 0)
    used = 0
    for b in [0, 1, limit - 2, 1, 1, 2**64 - 1, 2**63, 0]:
        want = O.orc_mem_prehook_allows(b, used, limit)
        assert L.gemhook_pool_mem_reserve(p, 0, b) == want, (b, used)
        if want:
            used += b
    assert used == limit
    L.gemhook_pool_close(p)


def test_array_byte_rules():
    L, O = kb.lib(), orc.load()
    for w, h, d, ch, fmt, is3d in [(640, 480, 0, 4, 0x01, 0), (640, 480, 0, 1, 0x20, 0), (33, 7, 5, 2, 0x10, 1),
                                   (1024, 0, 0, 4, 0x03, 0), (8, 8, 0, 4, 0x0a, 1), (8, 8, 8, 3, 0x09, 1),
                                   (8, 8, 8, 3, 0x99, 1)]:
        assert L.gemhook_array_bytes(w, h, d, ch, fmt, is3d) == O.orc_array_bytes(w, h, d, ch, fmt, is3d)


def test_quota_file_column_order_trap():
    """kubeshare-config writes `limit request`, gem-schd reads `min max`."""
    L = kb.lib()
    who, q, slp = C.c_int(), C.c_double(), C.c_double()
    for swap, expect_max in ((0, 0.25), (1, 1.0)):
        p = L.gemhook_pool_open(None, 1, 300.0, 20.0, 10000.0, 1)
        L.gemhook_pool_load_config(p, b"1\nns/pod 1.0 0.25 1024\n", swap)
        # quota clamps at max_frac * window: feed a huge burst twice
        for t in (10.0, 15000.0, 30000.0):  # spaced so that earlier tokens have left the window
            L.gemhook_pool_request(p, 0, t, 0.0, 1e9)
            assert L.gemhook_pool_schedule(p, t, C.byref(who), C.byref(q), C.byref(slp)) == 1
        assert q.value == expect_max * 10000.0
        L.gemhook_pool_close(p)


def test_pod_level_token_rule_matches_reference_pmgr():
    """gem-pmgr's forwarding rule (pod-manager.cpp:316-473): golden from the live binary + oracle on random traces."""
    import random

    g = G["live_pmgr_forward"]
    L, O = kb.lib(), orc.load()
    p = L.gemhook_pool_open(None, 1, 300.0, 20.0, 10000.0, 1)
    L.gemhook_pool_load_config(p, b"1\nns/pod 1.0 1.0 123456789\n", 0)
    fo, fb, rem = C.c_double(), C.c_double(), C.c_double()
    now = 0
    for st in g["steps"]:
        now += 1
        fwd = L.gemhook_pool_pod_launch(p, 0, now, st["overuse"], st["burst"], C.byref(fo), C.byref(fb), C.byref(rem))
        if st["forwarded"] is None:
            assert fwd == 0 and rem.value <= prev_quota
        else:
            assert fwd == 1 and (fo.value, fb.value) == (st["forwarded"]["overuse"], st["forwarded"]["burst"])
            assert L.gemhook_pool_pod_granted(p, 0, now, st["schd_quota"]) == st["reply_quota"]
            prev_quota = st["schd_quota"]
    L.gemhook_pool_close(p)

    # random trace, two processes of one pod (two attachments through two handles on one file)
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "pool").encode()
        a = L.gemhook_pool_open(path, 1, 300.0, 20.0, 10000.0, 1)
        b = L.gemhook_pool_open(path, 0, 0, 0, 0, 0)
        L.gemhook_pool_load_config(a, b"1\nns/pod 1.0 1.0 1\n", 0)
        assert L.gemhook_pool_attach(a, 0) >= 0 and L.gemhook_pool_attach(b, 0) >= 0
        o = O.orc_pmgr_new(1, 0)
        O.orc_pmgr_connect(o, 0)
        O.orc_pmgr_connect(o, 1)
        rng = random.Random(12)
        now_us = 0
        e1, e2, e3 = C.c_double(), C.c_double(), C.c_double()
        forwards = 0
        for _ in range(3000):
            now_us += rng.randrange(1, 400_000)
            k = rng.randrange(2)
            over, burst = rng.choice([0.0, 0.0, 2.25]), rng.choice([0.0, 5.0, 80.0, 700.0])
            r1 = L.gemhook_pool_pod_launch([a, b][k], 0, now_us, over, burst, C.byref(fo), C.byref(fb), C.byref(rem))
            r2 = O.orc_pmgr_kernel_launch(o, k, now_us * 1000, over, burst, C.byref(e1), C.byref(e2), C.byref(e3))
            assert r1 == r2
            if r1:
                forwards += 1
                assert (fo.value, fb.value) == (e1.value, e2.value)
                q = rng.choice([20.0, 300.0, 1234.5])
                now_us += rng.randrange(10, 5000)
                assert L.gemhook_pool_pod_granted(a, 0, now_us, q) == O.orc_pmgr_schd_reply(o, now_us * 1000, q)
            else:
                assert rem.value == e3.value
        assert 100 < forwards < 2900
        L.gemhook_pool_close(b)
        L.gemhook_pool_close(a)
        O.orc_pmgr_free(o)


def test_dead_process_is_reaped_bytes_and_token():
    """A client killed with SIGKILL leaves bytes and maybe the token behind; peers reclaim both
    (reference: reclaim on socket close, pod-manager.cpp:533-545; token by timeout, scheduler.cpp:507-510)."""
    import signal
    import subprocess
    import sys
    import tempfile
    import time

    L = kb.lib()
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "pool")
        p = L.gemhook_pool_open(path.encode(), 1, 300.0, 20.0, 10000.0, 0)
        L.gemhook_pool_load_config(p, b"2\nA 0.5 1.0 1000\nB 0.5 1.0 1000\n", 0)
        child = subprocess.Popen([sys.executable, "-c", """
import sys, time, ctypes as C
sys.path.insert(0, %r)
import kubeshare_b200 as kb
L = kb.lib()
p = L.gemhook_pool_open(%r.encode(), 0, 0, 0, 0, 0)
assert L.gemhook_pool_attach(p, 0) >= 0
assert L.gemhook_pool_mem_reserve(p, 0, 700) == 1
q = L.gemhook_pool_acquire(p, 0, 0.0, 0.0)
print('held', q, flush=True)
time.sleep(60)
""" % (kb.ROOT, path)], stdout=subprocess.PIPE, text=True)
        assert child.stdout.readline().startswith("held")
        u, lim = C.c_uint64(), C.c_uint64()
        L.gemhook_pool_mem_info(p, 0, C.byref(u), C.byref(lim))
        assert u.value == 700
        assert L.gemhook_pool_reap(p) == 0          # alive: nothing to reclaim
        child.send_signal(signal.SIGKILL)
        child.wait()
        time.sleep(0.05)
        assert L.gemhook_pool_reap(p) == 1
        L.gemhook_pool_mem_info(p, 0, C.byref(u), C.byref(lim))
        assert u.value == 0
        # B can take the token at once although A held a 300 ms quota and never returned it
        assert L.gemhook_pool_attach(p, 1) >= 0
        t0 = time.time()
        assert L.gemhook_pool_acquire(p, 1, 0.0, 0.0) == 300.0
        assert time.time() - t0 < 0.2
        L.gemhook_pool_detach(p)
        L.gemhook_pool_close(p)


def test_quota_file_sync_reloads_only_on_change():
    import tempfile
    import time

    L = kb.lib()
    with tempfile.TemporaryDirectory() as tmp:
        qf = os.path.join(tmp, "q.txt")
        with open(qf, "w") as f:
            f.write("1\nns/a 0.5 1.0 100\n")
        p = L.gemhook_pool_open(os.path.join(tmp, "pool").encode(), 1, 300.0, 20.0, 10000.0, 0)
        assert L.gemhook_pool_sync_quota_file(p, qf.encode(), 0) == 1
        assert L.gemhook_pool_sync_quota_file(p, qf.encode(), 0) == 0      # unchanged: one stat, no reload
        q2 = L.gemhook_pool_open(os.path.join(tmp, "pool").encode(), 0, 0, 0, 0, 0)
        assert L.gemhook_pool_sync_quota_file(q2, qf.encode(), 0) == 0     # the stamp lives in the pool: peers see it
        time.sleep(0.01)
        with open(qf, "w") as f:
            f.write("2\nns/a 0.5 1.0 777\nns/new 0.25 0.5 4242\n")
        assert L.gemhook_pool_sync_quota_file(q2, qf.encode(), 0) == 1
        assert L.gemhook_pool_sync_quota_file(p, qf.encode(), 0) == 0
        u, lim = C.c_uint64(), C.c_uint64()
        L.gemhook_pool_mem_info(p, L.gemhook_pool_find(p, b"ns/new"), C.byref(u), C.byref(lim))
        assert lim.value == 4242
        L.gemhook_pool_mem_info(p, 0, C.byref(u), C.byref(lim))
        assert lim.value == 777
        with open(qf, "w") as f:
            f.write("3\nns/a 0.5 1.0 777\n")                              # half-written: count says 3, one row
        assert L.gemhook_pool_sync_quota_file(p, qf.encode(), 0) == -1
        assert L.gemhook_pool_sync_quota_file(p, b"/nonexistent/file", 0) == -1
        L.gemhook_pool_close(q2)
        L.gemhook_pool_close(p)


def test_quota_file_edge_cases():
    """The file starts life as "0" (launcher-multigpus.sh:26-31) and is rewritten whole by kubeshare-config."""
    L = kb.lib()
    p = L.gemhook_pool_open(None, 1, 300.0, 20.0, 10000.0, 1)
    assert L.gemhook_pool_load_config(p, b"0\n", 0) == 0 and L.gemhook_pool_nslots(p) == 0
    assert L.gemhook_pool_load_config(p, b"", 0) == -1                      # no count at all
    assert L.gemhook_pool_load_config(p, b"2\nns/a 0.5 1.0 10\n", 0) == -1  # count says 2, one row (torn write)
    assert L.gemhook_pool_find(p, b"ns/a") == 0                              # ... the complete row was still taken
    assert L.gemhook_pool_load_config(p, b"1\nns/a 0.25 0.75 99\n", 0) == 1  # re-read replaces the row in place
    assert L.gemhook_pool_nslots(p) == 1
    info = kb.SlotInfo()
    L.gemhook_pool_slot_info(p, 0, info)
    assert (info.min_frac, info.max_frac, info.mem_limit) == (0.25, 0.75, 99)
    # 64 clients fit, the 65th does not
    rows = "".join("ns/p%02d 0.01 1.0 1\n" % i for i in range(70))
    assert L.gemhook_pool_load_config(p, ("70\n" + rows).encode(), 0) == -1
    assert L.gemhook_pool_nslots(p) == 64
    assert L.gemhook_pool_find(p, b"ns/p62") == 63 and L.gemhook_pool_find(p, b"ns/p69") == -1
    L.gemhook_pool_close(p)
