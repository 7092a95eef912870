"""GPU parity of the sm_100a accounting reduction against the CPU oracle -- bit-exact (u64 sums).

Calls go through the C ABI (include/gemhook.h 2d) of libgemhook.so.1; torch only supplies the CUDA
context and device memory.
"""
import numpy as np
import pytest

import kubeshare_b200 as kb
import orc

pytestmark = pytest.mark.gpu

REC = np.dtype([("slot", "<u4"), ("launches", "<u4"), ("elapsed_ns", "<u8")])


def make_records(n, nslots, seed, out_of_range=0.05, big=False):
    rng = np.random.default_rng(seed)
    r = np.zeros(n, REC)
    r["slot"] = rng.integers(0, nslots, n, dtype=np.uint32)
    if out_of_range and n:
        bad = rng.random(n) < out_of_range
        r["slot"][bad] = rng.integers(nslots, 2**32, int(bad.sum()), dtype=np.uint64).astype(np.uint32)
    r["launches"] = rng.integers(0, 2**32 if big else 4097, n, dtype=np.uint64).astype(np.uint32)
    r["elapsed_ns"] = rng.integers(0, 2**62 if big else 50_000_000, n, dtype=np.uint64)
    return r


def oracle(L, r, nslots):
    ns, la, rc = orc.acct_reduce(L, r, nslots)
    return np.stack([ns, la, rc], axis=1)


@pytest.fixture(scope="module")
def torch_cuda():
    import torch

    assert torch.cuda.is_available()
    torch.cuda.init()
    torch.zeros(1, device="cuda")  # make the primary context current on this thread
    return torch


@pytest.fixture(scope="module")
def OL():
    return orc.load()


@pytest.mark.parametrize("nslots", [1, 2, 8, 21, 64])
@pytest.mark.parametrize("n", [0, 1, 31, 32, 255, 256, 257, 4097, 100_003, 1 << 20])
def test_reduce_host_matches_oracle(torch_cuda, OL, nslots, n):
    a = kb.Acct(nslots, ring_capacity=1 << 18)  # n > capacity exercises the chunked path
    try:
        r = make_records(n, nslots, seed=n * 131 + nslots)
        got = a.reduce_host(r)
        assert (got == oracle(OL, r, nslots)).all()
        # running totals: a second batch accumulates on top
        r2 = make_records(n // 2 + 3, nslots, seed=7 * n + nslots, big=True)
        got2 = a.reduce_host(r2)
        want2 = oracle(OL, r, nslots) + oracle(OL, r2, nslots)  # uint64 wrap-around is part of the contract
        assert (got2 == want2).all()
    finally:
        a.close()


def test_reduce_device_resident_and_page(torch_cuda, OL):
    torch = torch_cuda
    nslots, n = 4, (1 << 22) + 77
    r = make_records(n, nslots, seed=99)
    d = torch.from_numpy(r.view(np.uint8).reshape(-1).copy()).cuda()
    a = kb.Acct(nslots)
    try:
        ms = a.reduce_device(d.data_ptr(), n, timed=True)
        assert ms > 0
        tot, epoch = a.totals()
        assert epoch == 1
        assert (tot == oracle(OL, r, nslots)).all()
        a.reset()
        tot, _ = a.totals()
        assert (tot == 0).all()
        # linearity / split invariance at a size the oracle would take long on: reduce halves separately
        a.reduce_device(d.data_ptr(), n // 2)
        a.reduce_device(d.data_ptr() + (n // 2) * 16, n - n // 2)
        tot2, epoch2 = a.totals()
        assert (tot2 == oracle(OL, r, nslots)).all()
        assert a.kernel_launches >= 4
    finally:
        a.close()


def test_full_size_properties(torch_cuda):
    """BASELINE-size ring (2^26 records = 1 GiB): constant records make the answer closed-form."""
    torch = torch_cuda
    n, nslots = 1 << 26, 8
    rec = torch.empty((n, 4), dtype=torch.int32, device="cuda")
    idx = torch.arange(n, device="cuda", dtype=torch.int64)
    rec[:, 0] = (idx % nslots).to(torch.int32)          # slot
    rec[:, 1] = 3                                        # launches
    rec[:, 2] = 1000                                     # elapsed_ns low word
    rec[:, 3] = 0                                        # high word
    a = kb.Acct(nslots)
    try:
        a.reduce_device(rec.data_ptr(), n)
        tot, _ = a.totals()
        per = n // nslots
        assert (tot[:, 0] == per * 1000).all() and (tot[:, 1] == per * 3).all() and (tot[:, 2] == per).all()
    finally:
        a.close()


def test_misaligned_pointer_is_rejected(torch_cuda):
    a = kb.Acct(2)
    try:
        t = torch_cuda.zeros(64, dtype=torch_cuda.uint8, device="cuda")
        with pytest.raises(RuntimeError):
            a.reduce_device(t.data_ptr() + 8, 1)
    finally:
        a.close()


# ---- round 2: one-warp fast path, in-kernel bin flush, gpu_mem mirror, shared-pinned read ------------------------------
@pytest.mark.parametrize("nslots", [1, 3, 64])
@pytest.mark.parametrize("n", [2, 33, 511, 512, 513, 5000])
def test_fast_path_boundary_and_equivalence(torch_cuda, OL, nslots, n, monkeypatch):
    """n <= 512 runs gemhook_acct_reduce_small (one warp, no ticket); the result must not depend on which kernel ran."""
    r = make_records(n, nslots, seed=n * 17 + nslots, big=True)
    want = oracle(OL, r, nslots)
    for small in ("1", "0"):
        monkeypatch.setenv("GEMHOOK_ACCT_SMALL", small)
        a = kb.Acct(nslots, ring_capacity=1 << 14)
        try:
            assert a.grid_for(n) == (1 if (small == "1" and n <= 512) else a.grid_for(n))
            assert (a.reduce_host(r) == want).all()
            assert (a.reduce_host(r) == want + want).all()     # running totals through the atomics' return values
        finally:
            a.close()


@pytest.mark.parametrize("nslots,warps,stages,cols", [(3, 8, 2, 32), (17, 8, 3, 32), (33, 6, 4, 32), (64, 4, 5, 32), (64, 1, 8, 32),
                                                       (5, 8, 4, 16), (40, 8, 2, 16), (64, 8, 2, 16), (64, 3, 7, 16)])
@pytest.mark.parametrize("n", [513, 767, 768, 769, 4096, 70_001, (1 << 21) + 255])
def test_tma_staged_kernel_matches_oracle_and_the_register_staged_one(torch_cuda, OL, nslots, warps, stages, cols, n, monkeypatch):
    """gemhook_acct_reduce_staged[_c16] (above 22 client slots by default; 16 columns above 37): the same bins fed from
    per-warp rings of 4 KB buffers filled with cp.async.bulk, software-pipelined bin update with forwarding.  Forced on for
    small slot counts too, both column counts, ring depths 2..8, sizes that are not a multiple of the 256-record tile, fewer
    tiles than warps, the in-kernel flush every 2 tiles, runs of equal slots (forwarding) -- bit-exact against the oracle,
    hence equal to the register-staged kernel."""
    monkeypatch.setenv("GEMHOOK_ACCT_SMALL", "0")
    monkeypatch.setenv("GEMHOOK_ACCT_FLUSH_EVERY", "2")
    r = make_records(n, nslots, seed=n + 31 * nslots + stages, big=True)
    r["slot"][n // 3: n // 3 + 3000] = r["slot"][n // 3]          # a long run of one slot: every update forwards
    r["slot"][n // 2: n // 2 + 4096: 2] = (nslots - 1)             # and an alternating pattern
    want = oracle(OL, r, nslots)
    for staged in ("1", "0"):
        monkeypatch.setenv("GEMHOOK_ACCT_STAGED", staged)
        monkeypatch.setenv("GEMHOOK_ACCT_STAGED_COLS", str(cols))
        monkeypatch.setenv("GEMHOOK_ACCT_WARPS", str(warps))
        monkeypatch.setenv("GEMHOOK_ACCT_STAGES", str(stages))
        a = kb.Acct(nslots, ring_capacity=1 << 22)
        try:
            assert (a.reduce_host(r) == want).all(), staged
            assert (a.reduce_host(r) == want + want).all(), staged
        finally:
            a.close()


@pytest.mark.parametrize("every", [1, 3])
def test_in_kernel_bin_flush_is_exact(torch_cuda, OL, every, monkeypatch):
    """The packed (count << 48 | launches) half of a bin cell holds < 2^16 records per column; the kernel folds its bins
    into u64 accumulators every FLUSH_EVERY tiles (8000 in production: never reached below 2^31 records per launch).
    Forced to 1 / 3 tiles here, with launches = 2^32 - 1 so the packed sum is as large as it gets."""
    monkeypatch.setenv("GEMHOOK_ACCT_FLUSH_EVERY", str(every))
    nslots, n = 5, 300_007
    r = make_records(n, nslots, seed=every, big=True)
    r["launches"] = 2**32 - 1
    a = kb.Acct(nslots)
    try:
        assert (a.reduce_host(r) == oracle(OL, r, nslots)).all()
    finally:
        a.close()


def test_gpu_mem_mirror_and_shared_pinned_read(torch_cuda):
    """north_star (b): the pod's gpu_mem counter is a CAS word in the shared-pinned pool; every reduce launch mirrors
    (used, limit) into device memory and the totals page, and device code can read the pool words themselves."""
    import ctypes as C

    L = kb.lib()
    p = L.gemhook_pool_open(None, 1, 300.0, 20.0, 10000.0, 1)
    L.gemhook_pool_load_config(p, b"2\nns/a 0.5 1.0 1000000\nns/b 0.5 1.0 77\n", 0)
    assert L.gemhook_pool_mem_reserve(p, 0, 123456) == 1
    a = kb.Acct(2)
    try:
        u, lim = C.c_uint64(), C.c_uint64()
        L.gemhook_pool_mem_info(p, 0, C.byref(u), C.byref(lim))
        L.gemhook_acct_set_mem(a.h, 0, u.value, lim.value)
        for n in (10, 5000):                       # fast path and main kernel both publish the mirror
            a.reduce_host(make_records(n, 2, seed=n))
            for from_device in (0, 1):
                sl, us, li, ep = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_uint64()
                assert L.gemhook_acct_read_mem(a.h, from_device, C.byref(sl), C.byref(us), C.byref(li), C.byref(ep)) == 0
                assert (sl.value, us.value, li.value) == (0, 123456, 1000000) and ep.value >= 1
        # the device reads the SAME words the host arbitrates on (zero-copy through the registered mapping)
        words = (C.c_uint64 * 4)()
        assert L.gemhook_acct_peek_host_words(a.h, L.gemhook_pool_shared_words(p, 0), words) == 0, kb.last_error()
        assert (words[0], words[1]) == (123456, 1000000)
        L.gemhook_pool_mem_release(p, 0, 456)
        assert L.gemhook_acct_peek_host_words(a.h, L.gemhook_pool_shared_words(p, 0), words) == 0
        assert words[0] == 123000
    finally:
        a.close()
        L.gemhook_pool_close(p)
