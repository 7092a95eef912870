"""kubeshare_b200 -- B200-native replacement of KubeShare/Gemini's GPU-sharing hook (one hot path).

The product is native: ``lib/libgemhook.so.1`` (LD_PRELOAD shim + C ABI, ``include/gemhook.h``) with the
sm_100a accounting kernel embedded.  This Python module is only a thin ctypes binding of that C ABI for
the tests, ``bench.py`` and ``__graft_entry__.py`` -- it contains no compute and no fallback: if the
native library is missing it raises.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB_PATH = os.path.join(HERE, "lib", "libgemhook.so.1")
STORM_PATH = os.path.join(HERE, "bin", "gem-storm")
STUB_DIR = os.path.join(ROOT, "tests", "stub")

MAX_SLOTS = 64
REQ_QUOTA, REQ_MEM_LIMIT, REQ_MEM_UPDATE = 0, 1, 2


class Request(C.Structure):
    _fields_ = [("name", C.c_char * 72), ("req_id", C.c_int32), ("type", C.c_int32), ("overuse_ms", C.c_double),
                ("burst_ms", C.c_double), ("bytes", C.c_uint64), ("is_alloc", C.c_int32)]


class Response(C.Structure):
    _fields_ = [("req_id", C.c_int32), ("quota_ms", C.c_double), ("mem_used", C.c_uint64),
                ("mem_total", C.c_uint64), ("verdict", C.c_int32)]


class Record(C.Structure):
    _fields_ = [("slot", C.c_uint32), ("launches", C.c_uint32), ("elapsed_ns", C.c_uint64)]


class SlotInfo(C.Structure):
    _fields_ = [("name", C.c_char * 64), ("min_frac", C.c_double), ("max_frac", C.c_double), ("mem_limit", C.c_uint64),
                ("mem_used", C.c_uint64), ("gpu_ns", C.c_uint64), ("launches", C.c_uint64), ("tokens", C.c_uint64),
                ("quota_ms", C.c_double), ("accumulated_ms", C.c_double), ("holds_token", C.c_int32), ("waiting", C.c_int32)]


class Stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("launches", "fast_path", "slow_path", "token_requests", "host_syncs",
                                          "segments", "acct_kernels", "gpu_ns", "mem_used", "mem_limit",
                                          "allocs_denied")] + \
               [(n, C.c_double) for n in ("quota_ms", "overuse_ms", "token_wait_ms", "accumulated_token_ms")]


_lib = None


def lib():
    """Load libgemhook.so.1 (RTLD_LOCAL: its interposers do not affect this process)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("native library missing: %s (run `python __graft_entry__.py` to build)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    d, i64, u64, i32, u32, vp, cp, sz = (C.c_double, C.c_int64, C.c_uint64, C.c_int32, C.c_uint32, C.c_void_p,
                                          C.c_char_p, C.c_size_t)
    pd, pi = C.POINTER(C.c_double), C.POINTER(C.c_int)
    sig = {
        "gemhook_hooked_symbols": (C.POINTER(cp), [C.POINTER(sz)]),
        "gemhook_wire_pack_request": (C.c_int, [C.POINTER(Request), vp]),
        "gemhook_wire_unpack_request": (C.c_int, [vp, C.POINTER(Request)]),
        "gemhook_wire_pack_response": (C.c_int, [i32, C.POINTER(Response), vp]),
        "gemhook_wire_unpack_response": (C.c_int, [i32, vp, C.POINTER(Response)]),
        "gemhook_gate_new": (vp, []), "gemhook_gate_free": (None, [vp]),
        "gemhook_gate_launch_begin": (C.c_int, [vp, i64]),
        "gemhook_gate_renew_request": (None, [vp, i64, pd, pd]),
        "gemhook_gate_renew_granted": (None, [vp, i64, d]),
        "gemhook_gate_launch_end": (None, [vp, i64]),
        "gemhook_gate_host_sync": (None, [vp, i64]),
        "gemhook_gate_tracker_fire": (None, [vp, i64, C.c_float]),
        "gemhook_gate_tracker_complete": (C.c_int, [vp]),
        "gemhook_gate_quota_ms": (d, [vp]), "gemhook_gate_overuse_ms": (d, [vp]),
        "gemhook_gate_is_open": (C.c_int, [vp]), "gemhook_gate_expire": (None, [vp]),
        "gemhook_gate_predicted_window_ms": (d, [vp, i64]),
        "gemhook_estimate_full_burst": (d, [d, d]),
        "gemhook_predictor_new": (vp, [d]), "gemhook_predictor_free": (None, [vp]),
        "gemhook_predictor_record_start": (None, [vp, i64]), "gemhook_predictor_record_stop": (None, [vp, i64]),
        "gemhook_predictor_interrupt": (None, [vp]),
        "gemhook_predictor_ongoing_unmerged": (C.c_int, [vp]), "gemhook_predictor_ongoing_merged": (C.c_int, [vp]),
        "gemhook_predictor_predict_unmerged": (d, [vp, i64]), "gemhook_predictor_predict_merged": (d, [vp, i64]),
        "gemhook_pool_open": (vp, [cp, C.c_int, d, d, d, i64]), "gemhook_pool_close": (None, [vp]),
        "gemhook_pool_load_config": (C.c_int, [vp, cp, C.c_int]),
        "gemhook_pool_sync_quota_file": (C.c_int, [vp, cp, C.c_int]),
        "gemhook_pool_find": (C.c_int, [vp, cp]), "gemhook_pool_nslots": (C.c_int, [vp]),
        "gemhook_pool_request": (C.c_int, [vp, C.c_int, d, d, d]),
        "gemhook_pool_schedule": (C.c_int, [vp, d, pi, pd, pd]),
        "gemhook_pool_usage": (d, [vp, C.c_int, d]),
        "gemhook_pool_history": (sz, [vp, pi, pd, pd, sz]),
        "gemhook_pool_accumulated_ms": (d, [vp, C.c_int]),
        "gemhook_pool_now_ms": (d, [vp]),
        "gemhook_pool_acquire": (d, [vp, C.c_int, d, d]),
        "gemhook_pool_acquire_ex": (d, [vp, C.c_int, d, d, pi]),
        "gemhook_pool_release": (None, [vp, C.c_int]),
        "gemhook_pool_expire_token": (None, [vp]), "gemhook_pool_others_waiting": (C.c_int, [vp, C.c_int]),
        "gemhook_pool_slot_info": (C.c_int, [vp, C.c_int, C.POINTER(SlotInfo)]),
        "gemhook_pool_counters": (None, [vp, C.POINTER(u64), C.POINTER(u64), C.POINTER(u64)]),
        "gemhook_pool_attach": (C.c_int, [vp, C.c_int]), "gemhook_pool_detach": (None, [vp]),
        "gemhook_pool_reap": (C.c_int, [vp]),
        "gemhook_pool_pod_launch": (C.c_int, [vp, C.c_int, i64, d, d, pd, pd, pd]),
        "gemhook_pool_pod_granted": (d, [vp, C.c_int, i64, d]),
        "gemhook_pool_mem_reserve": (C.c_int, [vp, C.c_int, u64]),
        "gemhook_pool_mem_release": (None, [vp, C.c_int, u64]),
        "gemhook_pool_mem_info": (None, [vp, C.c_int, C.POINTER(u64), C.POINTER(u64)]),
        "gemhook_array_bytes": (u64, [u64, u64, u64, u32, u32, C.c_int]),
        "gemhook_mipmap_bytes": (u64, [u64, u64, u64, u32, u32, u32]),
        "gemhook_call_counts": (sz, [C.POINTER(C.POINTER(cp)), C.POINTER(C.POINTER(u64))]),
        "gemhook_acct_create": (vp, [u32, sz]), "gemhook_acct_destroy": (None, [vp]),
        "gemhook_acct_reduce_host": (C.c_int, [vp, vp, sz, vp]),
        "gemhook_acct_reduce_device": (C.c_int, [vp, u64, sz, C.POINTER(C.c_float)]),
        "gemhook_acct_read_totals": (C.c_int, [vp, vp, C.POINTER(u64)]),
        "gemhook_acct_sync": (C.c_int, [vp]), "gemhook_acct_reset": (C.c_int, [vp]),
        "gemhook_acct_kernel_launches": (u64, [vp]), "gemhook_acct_stream": (u64, [vp]),
        "gemhook_acct_grid_for": (u32, [vp, sz]),
        "gemhook_acct_launch_shape": (None, [vp, C.POINTER(C.c_uint32)]),
        "gemhook_acct_set_mem": (None, [vp, u32, u64, u64]),
        "gemhook_acct_read_mem": (C.c_int, [vp, C.c_int, C.POINTER(u64), C.POINTER(u64), C.POINTER(u64), C.POINTER(u64)]),
        "gemhook_acct_peek_host_words": (C.c_int, [vp, vp, C.POINTER(u64)]),
        "gemhook_pool_shared_words": (vp, [vp, C.c_int]),
        "gemhook_get_stats": (C.c_int, [C.POINTER(Stats)]), "gemhook_flush": (C.c_int, []),
        "gemhook_last_error": (cp, []), "gemhook_version": (cp, []),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)  # AttributeError here = the C ABI lost a symbol include/gemhook.h declares
        f.restype, f.argtypes = res, args
    _lib = L
    return L


def abi_symbols():
    """Every gemhook_* function include/gemhook.h declares (parsed from the header)."""
    import re

    text = open(os.path.join(ROOT, "include", "gemhook.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gemhook_[a-z0-9_]+)\s*\(", text)))


def last_error():
    return (lib().gemhook_last_error() or b"").decode()


class Acct:
    """Device accounting object bound to the current CUDA context (torch's primary context in the tests)."""

    def __init__(self, nslots, ring_capacity=1 << 20):
        self.L = lib()
        self.nslots = nslots
        self.h = self.L.gemhook_acct_create(nslots, ring_capacity)
        if not self.h:
            raise RuntimeError("gemhook_acct_create failed: " + last_error())

    def reduce_host(self, records_np):
        import numpy as np

        raw = np.ascontiguousarray(records_np).view(np.uint8).reshape(-1)
        out = np.zeros(self.nslots * 3, np.uint64)
        if self.L.gemhook_acct_reduce_host(self.h, raw.ctypes.data, raw.size // 16, out.ctypes.data) != 0:
            raise RuntimeError(last_error())
        return out.reshape(self.nslots, 3)

    def reduce_device(self, dptr, n, timed=False):
        ms = C.c_float(0)
        if self.L.gemhook_acct_reduce_device(self.h, int(dptr), int(n), C.byref(ms) if timed else None) != 0:
            raise RuntimeError(last_error())
        return ms.value

    def totals(self):
        import numpy as np

        self.sync()
        out = np.zeros(self.nslots * 3, np.uint64)
        ep = C.c_uint64()
        if self.L.gemhook_acct_read_totals(self.h, out.ctypes.data, C.byref(ep)) != 0:
            raise RuntimeError(last_error())
        return out.reshape(self.nslots, 3), ep.value

    def sync(self):
        if self.L.gemhook_acct_sync(self.h) != 0:
            raise RuntimeError(last_error())

    def reset(self):
        if self.L.gemhook_acct_reset(self.h) != 0:
            raise RuntimeError(last_error())

    def grid_for(self, n):
        return self.L.gemhook_acct_grid_for(self.h, n)

    def launch_shape(self):
        out = (C.c_uint32 * 6)()
        self.L.gemhook_acct_launch_shape(self.h, out)
        return dict(zip(("warps", "wave_blocks", "smem_bytes", "stages", "cols", "small_smem"), out))

    @property
    def kernel_launches(self):
        return self.L.gemhook_acct_kernel_launches(self.h)

    def close(self):
        if self.h:
            self.L.gemhook_acct_destroy(self.h)
            self.h = None
