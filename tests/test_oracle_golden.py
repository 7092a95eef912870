"""Pin the CPU oracle (oracle/gemini_oracle.cpp) against the REFERENCE's own code.

tests/golden/ref_golden.json was produced by tests/golden/make_golden.py from the reference's
object code under a virtual clock and from the live gem-schd / gem-pmgr binaries.  Every value
here must be reproduced bit-exactly (doubles compared with ==).
"""
import ctypes as C
import json
import os

import pytest

import orc
import wireproto as wp

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "ref_golden.json")))


@pytest.fixture(scope="module")
def L():
    return orc.load()


# ----------------------------------------------------------------------------- wire
@pytest.mark.parametrize("case", G["wire"], ids=lambda c: c["name"][:12])
def test_wire_requests_match_reference_bytes(L, case):
    for r in case["requests"]:
        buf = (C.c_uint8 * 80)()
        used = L.orc_wire_request(buf, case["name"].encode(), r["id"], r["type"], r["overuse"], r["burst"],
                                  r["bytes"], r["alloc"])
        assert bytes(buf).hex() == r["hex"]
        assert used <= 80
        # the test-side python codec must agree too (it drives the live daemons)
        assert wp.pack_request(case["name"], r["id"], r["type"], r["overuse"], r["burst"], r["bytes"],
                               r["alloc"]).hex() == r["hex"]
        nm = C.create_string_buffer(80)
        nl, rid, ty = C.c_uint64(), C.c_int32(), C.c_int32()
        off = L.orc_wire_parse_request(buf, nm, C.byref(nl), C.byref(rid), C.byref(ty))
        assert (nm.value.decode(), nl.value, rid.value, ty.value, off) == (
            r["parsed_name"], r["parsed_len"], r["parsed_id"], r["parsed_type"], r["payload_off"])


def test_wire_responses_match_reference_bytes(L):
    for r in G["wire"][0]["responses"]:
        buf = (C.c_uint8 * 40)()
        n = L.orc_wire_response(buf, r["type"], r["id"], r.get("quota", 0.0), r.get("used", 0), r.get("total", 0),
                                r.get("verdict", 0))
        assert bytes(buf).hex() == r["hex"] and n == r["len"]
        assert wp.pack_response(r["type"], r["id"], r.get("quota", 0.0), r.get("used", 0), r.get("total", 0),
                                r.get("verdict", 0)).hex() == r["hex"]


# ----------------------------------------------------------------------------- predictor
@pytest.mark.parametrize("trace", G["predictor"], ids=lambda t: "seed%d" % t["seed"])
def test_predictor_trace(L, trace):
    p = L.orc_pred_new(trace["thres"])
    try:
        for i, op in enumerate(trace["ops"]):
            t = op["t_ns"]
            if op["op"] == "start":
                L.orc_pred_record_start(p, t)
            elif op["op"] == "stop":
                L.orc_pred_record_stop(p, t)
            elif op["op"] == "interrupt":
                L.orc_pred_interrupt(p)
            got = (L.orc_pred_predict_unmerged(p, t), L.orc_pred_predict_merged(p, t),
                   L.orc_pred_ongoing_unmerged(p), L.orc_pred_ongoing_merged(p))
            assert got == (op["unmerged"], op["merged"], op["on_u"], op["on_m"]), "op %d %r" % (i, op)
    finally:
        L.orc_pred_free(p)


# ----------------------------------------------------------------------------- scheduler
def _ms(t_ns, start_ns):
    return ((t_ns - start_ns) // 1000) / 1e3  # ms_since_start(), scheduler.cpp:107-109


@pytest.mark.parametrize("sc", G["schd"], ids=lambda s: "seed%d" % s["seed"])
def test_scheduler_trace(L, sc):
    h = L.orc_schd_new(sc["base"], sc["min"], sc["window"])
    try:
        assert L.orc_schd_load_config(h, sc["config_text"].encode()) == len(sc["clients"])
        for name, mn, mx, mem in sc["clients"]:
            assert L.orc_schd_has_client(h, name.encode())
            assert L.orc_schd_mem_limit(h, name.encode()) == mem
        start = sc["start_ns"]
        nsleeps = 0
        for i, st in enumerate(sc["steps"]):
            now = _ms(st["t_ns"], start)
            for name, overuse, burst in st["requests"]:
                assert L.orc_schd_request(h, name.encode(), now, overuse, burst) == 0
            if st["selected"] is not None:
                nm = C.create_string_buffer(128)
                slp = C.c_double()
                wake = list(st["wakeups_ns"])
                while True:
                    rc = L.orc_schd_select(h, now, nm, C.byref(slp))
                    if rc == 1:
                        break
                    assert rc == 0 and wake, "oracle wants to sleep but the reference did not (step %d)" % i
                    nxt = wake.pop(0)
                    # the reference slept until `nxt` (get_timespec_after, scheduler.cpp:76-88, 385)
                    assert abs((_ms(nxt, start) - now) - slp.value) < 2e-3 or slp.value <= 0
                    now = _ms(nxt, start)
                    nsleeps += 1
                assert not wake, "reference slept more often than the oracle (step %d)" % i
                assert nm.value.decode() == st["selected"], "step %d" % i
                assert _ms(st["t_after_ns"], start) == now
                q = L.orc_schd_grant(h, st["selected"].encode(), now)
                assert q == st["quota"], "step %d" % i
            if "history" in st:
                assert orc.schd_history(L, h) == st["history"], "step %d" % i
        if sc["seed"] in (3, 21, 22):
            assert nsleeps > 0  # the throttling path (all candidates at their limit) is exercised
    finally:
        L.orc_schd_free(h)


def test_live_schd_known_answers(L):
    """BASELINE.md 2 / SURVEY.md 8c known answers from the running reference gem-schd."""
    g = G["live_schd"]
    h = L.orc_schd_new(g["base"], g["min"], g["window"])
    L.orc_schd_load_config(h, g["config"].encode())
    now = 0.0
    for c in g["calls"]:
        if c["op"] == "mem_limit":
            assert (0, L.orc_schd_mem_limit(h, b"ns/a")) == (c["used"], c["total"])
        elif c["op"] == "mem_update":
            assert c["verdict"] == 1  # scheduler.cpp:442-455 always answers 1
        else:
            now += 10.0
            L.orc_schd_request(h, b"ns/a", now, c["overuse"], c["burst"])
            nm = C.create_string_buffer(128)
            slp = C.c_double()
            assert L.orc_schd_select(h, now, nm, C.byref(slp)) == 1
            assert L.orc_schd_grant(h, b"ns/a", now) == c["quota"]
    L.orc_schd_free(h)
    # the three BASELINE.md answers, spelled out
    qs = [c["quota"] for c in g["calls"] if c["op"] == "quota"]
    assert qs[:3] == [250.0, 127.5, 100.0]


def test_select_candidate_hand_kat(L):
    """A=[10,110], B=[60,160], now=200 (< window) -> usage 75 / 75."""
    h = L.orc_schd_new(1e9, 1e9, 10000.0)  # quota = base = huge so that Record() ends are set by hand below
    L.orc_schd_set_client(h, b"A", 0.5, 1.0, 0)
    L.orc_schd_set_client(h, b"B", 0.5, 1.0, 0)
    # build the ledger through the public path: grant at t=10 / t=60, then return early at 110 / 160
    for name, t0, t1 in ((b"A", 10.0, 110.0), (b"B", 60.0, 160.0)):
        L.orc_schd_request(h, name, t0, 0.0, 0.0)
        nm = C.create_string_buffer(64)
        s = C.c_double()
        assert L.orc_schd_select(h, t0, nm, C.byref(s)) == 1
        L.orc_schd_grant(h, name, t0)
    # update_return_time clamps end to `now` when the client comes back (scheduler.cpp:128)
    L.orc_schd_request(h, b"A", 110.0, 0.0, 0.0)
    L.orc_schd_request(h, b"B", 160.0, 0.0, 0.0)
    assert orc.schd_history(L, h) == [["A", 10.0, 110.0], ["B", 60.0, 160.0]]
    assert L.orc_schd_usage(h, b"A", 200.0) == 75.0
    assert L.orc_schd_usage(h, b"B", 200.0) == 75.0
    L.orc_schd_free(h)


def test_schd_priority_rules(L):
    # schd-priority.cpp:19-26
    assert L.orc_schd_priority(10, 10, 5, 20) == 1      # both under-served: larger missing share first
    assert L.orc_schd_priority(5, 20, 10, 10) == 0
    assert L.orc_schd_priority(1, 100, -1, 0) == 1      # under-served beats over-served
    assert L.orc_schd_priority(-1, 0, 1, 100) == 0
    assert L.orc_schd_priority(-5, 10, -1, 20) == 1     # both over-served: least usage first
    assert L.orc_schd_priority(0, 30, 0, 20) == 0


# ----------------------------------------------------------------------------- pod manager
def test_pmgr_memory_counter_matches_live_reference(L):
    g = G["live_pmgr_mem"]
    p = L.orc_pmgr_new(g["limit"], 0)
    L.orc_pmgr_connect(p, 0)
    L.orc_pmgr_connect(p, 1)
    used, lim = C.c_uint64(), C.c_uint64()
    rejected = 0
    for op in g["ops"]:
        if op["op"] == "alloc":
            L.orc_pmgr_mem_info(p, C.byref(used), C.byref(lim))
            assert (used.value, lim.value) == (op["used_before"], op["total"])
            v = L.orc_pmgr_mem_update(p, op["conn"], op["bytes"], 1)
            assert v == op["verdict"]
            # hook-side pre-check agrees with the pmgr verdict for a single-threaded client
            assert L.orc_mem_prehook_allows(op["bytes"], op["used_before"], op["total"]) == op["verdict"]
            rejected += (v == 0)
        elif op["op"] == "free":
            assert L.orc_pmgr_mem_update(p, op["conn"], op["bytes"], 0) == op["verdict"] == 1
        else:
            L.orc_pmgr_disconnect(p, op["conn"])
        L.orc_pmgr_mem_info(p, C.byref(used), C.byref(lim))
        assert used.value == op["used_after"]
    assert rejected >= 10
    L.orc_pmgr_free(p)


def test_pmgr_forwarding_rule_matches_live_reference(L):
    g = G["live_pmgr_forward"]
    assert g["hello"]["type"] == wp.REQ_MEM_LIMIT  # retrieve_mem_info, pod-manager.cpp:126-160
    p = L.orc_pmgr_new(g["limit"], 0)
    L.orc_pmgr_connect(p, 7)
    now = 0
    fo, fb, rem = C.c_double(), C.c_double(), C.c_double()
    for st in g["steps"]:
        now += 1000  # 1 us later: elapsed is truncated to whole microseconds by the reference
        fwd = L.orc_pmgr_kernel_launch(p, 7, now, st["overuse"], st["burst"], C.byref(fo), C.byref(fb), C.byref(rem))
        if st["forwarded"] is None:
            assert fwd == 0
            # live run: reply = pod_quota - elapsed with real elapsed time; same rule, elapsed differs
            assert st["reply_quota"] <= st_prev_quota and st_prev_quota - st["reply_quota"] < 5000.0
            assert rem.value <= st_prev_quota
        else:
            assert fwd == 1
            assert (fo.value, fb.value) == (st["forwarded"]["overuse"], st["forwarded"]["burst"])
            assert L.orc_pmgr_schd_reply(p, now, st["schd_quota"]) == st["reply_quota"]
            st_prev_quota = st["schd_quota"]
    L.orc_pmgr_free(p)
