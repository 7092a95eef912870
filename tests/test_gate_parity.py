"""CPU: the product's predictor and launch gate (csrc/gh_gate.cpp, through the C ABI) against
(a) golden traces produced by the reference's own predictor.o and (b) the oracle on seeded launch traces.
Doubles are compared with == (the arithmetic is the reference's: whole microseconds / 1e3)."""
import ctypes as C
import json
import os
import random

import pytest

import kubeshare_b200 as kb
import orc

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "ref_golden.json")))


@pytest.mark.parametrize("trace", G["predictor"], ids=lambda t: "seed%d" % t["seed"])
def test_predictor_matches_reference_object_code(trace):
    L = kb.lib()
    p = L.gemhook_predictor_new(trace["thres"])
    for i, op in enumerate(trace["ops"]):
        t = op["t_ns"]
        if op["op"] == "start":
            L.gemhook_predictor_record_start(p, t)
        elif op["op"] == "stop":
            L.gemhook_predictor_record_stop(p, t)
        elif op["op"] == "interrupt":
            L.gemhook_predictor_interrupt(p)
        got = (L.gemhook_predictor_predict_unmerged(p, t), L.gemhook_predictor_predict_merged(p, t),
               L.gemhook_predictor_ongoing_unmerged(p), L.gemhook_predictor_ongoing_merged(p))
        assert got == (op["unmerged"], op["merged"], op["on_u"], op["on_m"]), "op %d %r" % (i, op)
    L.gemhook_predictor_free(p)


def test_estimate_full_burst():
    L, O = kb.lib(), orc.load()
    for b, w in [(0.0, 0.0), (1e-10, 5.0), (3.5, 0.0), (3.5, 1.999), (3.5, 2.0), (250.0, 100.0)]:
        assert L.gemhook_estimate_full_burst(b, w) == O.orc_estimate_full_burst(b, w)


def replay(seed, n_events, quota_policy):
    """Drive product gate and oracle hook with the same synthetic launch trace; compare every output.

    The trace models an application: bursts of launches separated by host syncs and idle gaps, token
    renewals answered by `quota_policy`, tracker drains firing at token expiry or on demand."""
    L, O = kb.lib(), orc.load()
    g, h = L.gemhook_gate_new(), O.orc_hook_new()
    rng = random.Random(seed)
    t = 5_000_000_000_123  # ns
    token_deadline = None
    ev_token_t = t
    renewals = []
    d1, d2, e1, e2 = C.c_double(), C.c_double(), C.c_double(), C.c_double()

    def drain(now):
        elapsed = C.c_float((now - ev_token_t) / 1e6).value
        L.gemhook_gate_tracker_fire(g, now, elapsed)
        O.orc_hook_tracker_fire(h, now, elapsed)

    for _ in range(n_events):
        r = rng.random()
        t += rng.choice([rng.randrange(200, 3000), rng.randrange(1000, 60_000), rng.randrange(10**5, 5 * 10**6),
                         rng.randrange(10**6, 4 * 10**8)])
        # token expiry fires the tracker asynchronously
        if token_deadline is not None and t >= token_deadline and not L.gemhook_gate_tracker_complete(g):
            drain(token_deadline + rng.randrange(0, 2_000_000))
            token_deadline = None
        if r < 0.75:  # a launch
            a, b = L.gemhook_gate_launch_begin(g, t), O.orc_hook_launch_begin(h, t)
            assert a == b
            if a:
                if not L.gemhook_gate_tracker_complete(g):
                    assert not O.orc_hook_tracker_complete(h)
                    t += rng.randrange(1000, 500_000)
                    drain(t)
                    token_deadline = None
                L.gemhook_gate_renew_request(g, t, C.byref(d1), C.byref(d2))
                O.orc_hook_renew_request(h, t, C.byref(e1), C.byref(e2))
                assert (d1.value, d2.value) == (e1.value, e2.value)
                q = quota_policy(rng, d2.value)
                t += rng.randrange(2000, 300_000)  # token round trip
                L.gemhook_gate_renew_granted(g, t, q)
                O.orc_hook_renew_granted(h, t, q)
                ev_token_t = t
                token_deadline = t + int(max(q, 0.0) * 1e6)
                renewals.append((d1.value, d2.value, q))
            t += rng.randrange(100, 900)
            L.gemhook_gate_launch_end(g, t)
            O.orc_hook_launch_end(h, t)
        else:  # a synchronising call
            L.gemhook_gate_host_sync(g, t)
            O.orc_hook_host_sync(h, t)
        assert L.gemhook_gate_quota_ms(g) == O.orc_hook_quota_ms(h)
        assert L.gemhook_gate_overuse_ms(g) == O.orc_hook_overuse_ms(h)
        assert L.gemhook_gate_tracker_complete(g) == O.orc_hook_tracker_complete(h)
    L.gemhook_gate_free(g)
    O.orc_hook_free(h)
    return renewals


@pytest.mark.parametrize("seed", [1, 2, 3, 0xB200])
def test_gate_matches_oracle_on_seeded_traces(seed):
    ema = {"q": 300.0}

    def policy(rng, burst):  # gem-schd's get_quota shape: EMA toward the reported burst
        if burst < 1e-9:
            ema["q"] = 300.0
        else:
            ema["q"] = min(max(0.5 * burst + 0.5 * ema["q"], 20.0), 10000.0)
        return ema["q"]

    ren = replay(seed, 6000, policy)
    assert len(ren) > 20
    assert any(o > 0 for o, _, _ in ren), "overuse path never exercised"
    assert any(b > 0 for _, b, _ in ren), "burst estimate never non-zero"


def test_first_launch_always_renews():
    """quota_time starts at 0 and request_start at the epoch, so the first launch renews (hook.cpp:766-768)."""
    L = kb.lib()
    g = L.gemhook_gate_new()
    assert L.gemhook_gate_launch_begin(g, 123_456_789_000) == 1
    a, b = C.c_double(), C.c_double()
    L.gemhook_gate_renew_request(g, 123_456_789_000, C.byref(a), C.byref(b))
    assert (a.value, b.value) == (0.0, 0.0)
    L.gemhook_gate_renew_granted(g, 123_456_800_000, 300.0)
    L.gemhook_gate_launch_end(g, 123_456_800_500)
    assert L.gemhook_gate_is_open(g) == 1
    assert L.gemhook_gate_launch_begin(g, 123_456_900_000) == 0  # burst ongoing: no check at all
    L.gemhook_gate_free(g)
