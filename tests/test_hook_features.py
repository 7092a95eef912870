"""CPU: round-2 hook features on the stub driver -- stream-capture safety, cuStreamDestroy, token trace, per-symbol
call counters (CU_HOOK_DEBUG), opt-in accounting of managed / mipmapped / pinned-host memory and communicate() retry.  (Live publication of the
device-reduced usage needs kernels that run: tests/test_gpu_hook.py.)"""
import json
import os
import socket
import subprocess as sp
import sys
import tempfile
import threading

import kubeshare_b200 as kb
import wireproto as wp
from test_preload_stub import base_env, hooked_env, run_storm, stats_files


def test_no_event_ever_lands_on_a_capturing_stream():
    """torch.cuda.graph()-style capture under the hook.  Short tokens expire during the capture, so the
    launch slow path, the tracker thread and the sync pre-hook all run while a stream is capturing."""
    with tempfile.TemporaryDirectory() as tmp:
        env = hooked_env(tmp, GEMHOOK_BASE_QUOTA_MS=5, GEMHOOK_MIN_QUOTA_MS=2, GEMHOOK_SEG_MIN_US=100, GEMHOOK_SEG_LAUNCHES=16)
        res = run_storm(env, "--mode", "graph", "--step-launches", 64, "--rounds", 5, "--spin-us", 5)
        stub = json.load(open(os.path.join(tmp, "stub.json")))
        st = stats_files(tmp)[0]
    assert [res[k] for k in ("begin", "launch", "end", "instantiate", "replay", "sync", "destroy")] == [0] * 7
    assert stub["events_on_capturing_streams"] == 0
    assert stub["stream_destroys"] == 1           # the interposed cuStreamDestroy_v2 reached the driver
    assert st["launches"] == 64 + 64 + 5 + 64     # pre-capture + captured + graph launches + tail: all pass the gate
    assert st["token_requests"] >= 3              # tokens did expire during the capture


def test_token_trace_and_call_counters():
    with tempfile.TemporaryDirectory() as tmp:
        env = hooked_env(tmp, GEMHOOK_BASE_QUOTA_MS=4, GEMHOOK_MIN_QUOTA_MS=1, CU_HOOK_DEBUG=1, GEMHOOK_LOG=0,
                         GEMHOOK_TOKEN_TRACE=os.path.join(tmp, "trace.%d.jsonl"))
        res = run_storm(env, "--mode", "storm", "--steps", 3, "--warmup", 1, "--step-launches", 4000, "--sync-every", 500)
        st = stats_files(tmp)[0]
        tr = [json.loads(l) for f in os.listdir(tmp) if f.startswith("trace.") for l in open(os.path.join(tmp, f))]
    assert res["launches"] == 12000
    assert len(tr) == st["token_requests"] >= 3
    assert (tr[0]["overuse_ms"], tr[0]["burst_ms"]) == (0.0, 0.0) and tr[0]["quota_ms"] == 4.0   # initialize(): base quota
    assert all(t["pod"] == "bench/c0" and t["quota_ms"] > 0 for t in tr)
    assert [t["t_ms"] for t in tr] == sorted(t["t_ms"] for t in tr)
    calls = st["calls"]                               # reference hookInfo::call_count (hook.cpp:87-100)
    assert calls["cuLaunchKernel"] in (15999, 16000)   # the very first call loads the configuration
    assert calls["cuCtxSynchronize"] == st["host_syncs"] >= 32
    assert calls["cuMemAlloc"] == 0 and calls["cuMemGetInfo"] == 0


def test_counters_stay_off_without_cu_hook_debug():
    with tempfile.TemporaryDirectory() as tmp:
        run_storm(hooked_env(tmp), "--mode", "storm", "--steps", 1, "--warmup", 0, "--step-launches", 100)
        assert "calls" not in stats_files(tmp)[0]


def test_opt_in_accounting_of_managed_mipmapped_and_host_memory():
    cap = 1 << 20
    quota = "1\nbench/c0 1.0 1.0 %d\n" % cap
    with tempfile.TemporaryDirectory() as tmp:   # default = the reference: none of the three is charged
        res = run_storm(hooked_env(tmp, quota=quota), "--mode", "optin")
        assert res["rc"] == [0, 0, 0] and res["free"] == [cap] * 5
    with tempfile.TemporaryDirectory() as tmp:
        res = run_storm(hooked_env(tmp, quota=quota, GEMHOOK_ACCOUNT_MANAGED=1, GEMHOOK_ACCOUNT_HOST=1), "--mode", "optin")
        mip = 8 * 8 * 4 + 4 * 4 * 4 + 2 * 2 * 4      # three levels of an 8x8 float array
        assert res["free"] == [cap, cap - 4096, cap - 4096 - mip, cap - 4096 - mip - 2048, cap]
        L = kb.lib()
        assert L.gemhook_mipmap_bytes(8, 8, 0, 1, 0x20, 3) == mip
    with tempfile.TemporaryDirectory() as tmp:   # and the cap bites: 4096 managed bytes do not fit under 4000
        res = run_storm(hooked_env(tmp, quota="1\nbench/c0 1.0 1.0 4000\n", GEMHOOK_ACCOUNT_MANAGED=1), "--mode", "optin")
        assert res["rc"][0] == 2 and res["free"][1] == 4000


class DeafOncePodManager(threading.Thread):
    """gem-pmgr that swallows the first REQ_MEM_LIMIT: the hook's receive times out and communicate() must send the
    same 80 bytes again on the same socket (reference hook.cpp:312-324, comm.cpp:124-134)."""

    def __init__(self):
        super().__init__(daemon=True)
        self.lsock = socket.socket()
        self.lsock.bind(("127.0.0.1", 0))
        self.lsock.listen(2)
        self.port = self.lsock.getsockname()[1]
        self.raw, self.connections, self.deaf = [], 0, True

    def run(self):
        try:
            c, _ = self.lsock.accept()
            self.connections += 1
            while True:
                buf = wp.recv_exact(c, wp.REQ_LEN)
                self.raw.append(buf)
                r = wp.unpack_request(buf)
                if r["type"] == wp.REQ_QUOTA:
                    c.sendall(wp.pack_response(wp.REQ_QUOTA, r["id"], quota=50.0))
                elif r["type"] == wp.REQ_MEM_LIMIT:
                    if self.deaf:
                        self.deaf = False
                        continue
                    c.sendall(wp.pack_response(wp.REQ_MEM_LIMIT, r["id"], used=0, total=8192))
                else:
                    c.sendall(wp.pack_response(wp.REQ_MEM_UPDATE, r["id"], verdict=1))
        except (ConnectionError, OSError):
            pass


def test_failed_exchange_is_retried_on_the_same_socket():
    pm = DeafOncePodManager()
    pm.start()
    with tempfile.TemporaryDirectory() as tmp:
        env = base_env(tmp, LD_PRELOAD=kb.LIB_PATH, GEMHOOK_SCHEDULER_IP="127.0.0.1", POD_MANAGER_PORT=pm.port, POD_NAME="bench/c0",
                       GEMHOOK_RPC_TIMEOUT_S=1)
        res = run_storm(env, "--mode", "resolve")
    assert res["rc"] == [0, 0, 0] and res["total"] == 8192
    reqs = [wp.unpack_request(b) for b in pm.raw]
    lim = [r["id"] for r in reqs if r["type"] == wp.REQ_MEM_LIMIT]
    assert lim[0] == lim[1], "the swallowed REQ_MEM_LIMIT was not sent again"   # identical bytes, same req_id
    assert pm.raw[[r["type"] for r in reqs].index(wp.REQ_MEM_LIMIT)] == pm.raw[[r["type"] for r in reqs].index(wp.REQ_MEM_LIMIT) + 1]
    assert pm.connections == 1


def test_accounting_kernel_launch_shapes_fit_the_sm_for_every_slot_count():
    """gh_acct.cpp picks kernel, warps, ring depth and bin columns from the slot count.  On the CPU stub driver (any
    cuModuleGetFunction succeeds) walk 1..64 slots and check the arithmetic: the block fits into 227 KB of shared memory,
    the register-staged kernel runs up to 22 slots, the TMA-staged one beyond with >= 2 buffers per warp and >= 48 KB in
    flight per SM, 16 columns once eight warps with 32 columns no longer fit."""
    code = r"""
import json, sys
sys.path.insert(0, %r)
import ctypes as C
import kubeshare_b200 as kb
cu = C.CDLL("libcuda.so.1")          # the stub driver (LD_LIBRARY_PATH): a context has to be current
dev, ctx = C.c_int(), C.c_void_p()
assert cu.cuInit(0) == 0 and cu.cuDeviceGet(C.byref(dev), 0) == 0 and cu.cuCtxCreate_v2(C.byref(ctx), 0, dev) == 0
out = {}
for ns in range(1, 65):
    a = kb.Acct(ns, ring_capacity=1 << 10)
    out[ns] = a.launch_shape()
    a.close()
print(json.dumps(out))
""" % kb.ROOT
    env = {k: v for k, v in os.environ.items() if not k.startswith("GEMHOOK_") and k != "LD_PRELOAD"}
    env["LD_LIBRARY_PATH"] = kb.STUB_DIR + ":" + env.get("LD_LIBRARY_PATH", "")
    p = sp.run([sys.executable, "-c", code], env=env, stdout=sp.PIPE, stderr=sp.PIPE, timeout=120)
    assert p.returncode == 0, p.stderr.decode()[-2000:]
    shapes = {int(k): v for k, v in json.loads(p.stdout.decode().strip().splitlines()[-1]).items()}
    for ns, s in shapes.items():
        assert s["smem_bytes"] <= 227 * 1024, (ns, s)
        assert 1 <= s["warps"] <= 8 and s["wave_blocks"] >= 1
        bins32 = (ns + 1) * 512 + ns * 24
        if ns <= 22:
            assert s["stages"] == 0 and s["cols"] == 32 and s["smem_bytes"] == s["warps"] * bins32, (ns, s)
        else:
            assert 2 <= s["stages"] <= 3 and s["warps"] * s["stages"] * 4096 >= 48 * 1024, (ns, s)
            assert s["cols"] == (32 if 8 * (bins32 + 2 * 4104) + 16 <= 227 * 1024 else 16), (ns, s)
            per_warp = (ns + 1) * s["cols"] * 16 + ns * 24
            assert s["smem_bytes"] == ((s["warps"] * per_warp + 15) & ~15) + s["warps"] * s["stages"] * 4104, (ns, s)
    assert shapes[64] == dict(shapes[64], warps=8, stages=2, cols=16) and shapes[37]["cols"] == 32 and shapes[38]["cols"] == 16
