"""ctypes binding of oracle/libgemoracle.so (TEST INFRASTRUCTURE; see oracle/gemini_oracle.h)."""
import ctypes as C
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SO = os.path.join(ROOT, "oracle", "libgemoracle.so")


class AcctRecord(C.Structure):
    _fields_ = [("slot", C.c_uint32), ("launches", C.c_uint32), ("elapsed_ns", C.c_uint64)]


def load():
    if not os.path.exists(_SO):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "oracle"])
    L = C.CDLL(_SO)
    d, i64, u64, i32, u32, vp, cp = C.c_double, C.c_int64, C.c_uint64, C.c_int32, C.c_uint32, C.c_void_p, C.c_char_p
    pd = C.POINTER(C.c_double)
    sig = {
        "orc_wire_request": (C.c_size_t, [vp, cp, i32, i32, d, d, u64, i32]),
        "orc_wire_parse_request": (C.c_size_t, [vp, vp, C.POINTER(u64), C.POINTER(i32), C.POINTER(i32)]),
        "orc_wire_response": (C.c_size_t, [vp, i32, i32, d, u64, u64, i32]),
        "orc_pred_new": (vp, [d]), "orc_pred_free": (None, [vp]),
        "orc_pred_record_start": (None, [vp, i64]), "orc_pred_record_stop": (None, [vp, i64]),
        "orc_pred_interrupt": (None, [vp]),
        "orc_pred_ongoing_unmerged": (C.c_int, [vp]), "orc_pred_ongoing_merged": (C.c_int, [vp]),
        "orc_pred_predict_unmerged": (d, [vp, i64]), "orc_pred_predict_merged": (d, [vp, i64]),
        "orc_estimate_full_burst": (d, [d, d]),
        "orc_hook_new": (vp, []), "orc_hook_free": (None, [vp]),
        "orc_hook_launch_begin": (C.c_int, [vp, i64]),
        "orc_hook_renew_request": (None, [vp, i64, pd, pd]),
        "orc_hook_renew_granted": (None, [vp, i64, d]),
        "orc_hook_launch_end": (None, [vp, i64]),
        "orc_hook_host_sync": (None, [vp, i64]),
        "orc_hook_tracker_fire": (None, [vp, i64, C.c_float]),
        "orc_hook_tracker_complete": (C.c_int, [vp]),
        "orc_hook_quota_ms": (d, [vp]), "orc_hook_overuse_ms": (d, [vp]),
        "orc_us_since": (i64, [i64, i64]),
        "orc_mem_prehook_allows": (C.c_int, [u64, u64, u64]),
        "orc_array_bytes": (u64, [u64, u64, u64, u32, u32, C.c_int]),
        "orc_pmgr_new": (vp, [u64, i64]), "orc_pmgr_free": (None, [vp]),
        "orc_pmgr_connect": (None, [vp, C.c_int]), "orc_pmgr_disconnect": (None, [vp, C.c_int]),
        "orc_pmgr_mem_update": (C.c_int, [vp, C.c_int, u64, C.c_int]),
        "orc_pmgr_mem_info": (None, [vp, C.POINTER(u64), C.POINTER(u64)]),
        "orc_pmgr_kernel_launch": (C.c_int, [vp, C.c_int, i64, d, d, pd, pd, pd]),
        "orc_pmgr_schd_reply": (d, [vp, i64, d]),
        "orc_schd_new": (vp, [d, d, d]), "orc_schd_free": (None, [vp]),
        "orc_schd_set_client": (None, [vp, cp, d, d, u64]),
        "orc_schd_load_config": (C.c_int, [vp, cp]),
        "orc_schd_has_client": (C.c_int, [vp, cp]), "orc_schd_mem_limit": (u64, [vp, cp]),
        "orc_schd_request": (C.c_int, [vp, cp, d, d, d]),
        "orc_schd_select": (C.c_int, [vp, d, vp, pd]),
        "orc_schd_grant": (d, [vp, cp, d]),
        "orc_schd_usage": (d, [vp, cp, d]),
        "orc_schd_history_len": (C.c_size_t, [vp]),
        "orc_schd_history_get": (C.c_int, [vp, C.c_size_t, vp, pd, pd]),
        "orc_schd_accumulated_ms": (d, [vp, cp]),
        "orc_schd_priority": (C.c_int, [d, d, d, d]),
        "orc_acct_reduce": (None, [vp, C.c_size_t, u32, vp, vp, vp]),
        "orc_acct_reduce_mt": (None, [vp, C.c_size_t, u32, vp, vp, vp, C.c_int]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype, f.argtypes = res, args
    return L


def schd_history(L, h):
    out = []
    nm = C.create_string_buffer(128)
    a, b = C.c_double(), C.c_double()
    for i in range(L.orc_schd_history_len(h)):
        L.orc_schd_history_get(h, i, nm, C.byref(a), C.byref(b))
        out.append([nm.value.decode(), a.value, b.value])
    return out


def acct_reduce(L, records_np, nslots, threads=0):
    """records_np: numpy structured/uint32 view of 16-byte records. Returns (ns, launches, records) uint64 arrays."""
    import numpy as np

    raw = np.ascontiguousarray(records_np).view(np.uint8).reshape(-1)
    n = raw.size // 16
    ns = np.zeros(nslots, np.uint64)
    la = np.zeros(nslots, np.uint64)
    rc = np.zeros(nslots, np.uint64)
    if threads and threads > 1:
        L.orc_acct_reduce_mt(raw.ctypes.data, n, nslots, ns.ctypes.data, la.ctypes.data, rc.ctypes.data, threads)
    else:
        L.orc_acct_reduce(raw.ctypes.data, n, nslots, ns.ctypes.data, la.ctypes.data, rc.ctypes.data)
    return ns, la, rc
