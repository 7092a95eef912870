"""CPU: gem-arbiter, the native gem-schd + gem-pmgr replacement serving legacy TCP hooks from the shared pool
(SURVEY.md 8f-1).  Legacy side = the test's protocol client (codec pinned against the reference's bytes)."""
import json
import os
import subprocess as sp
import tempfile
import time

import pytest

import kubeshare_b200 as kb
import wireproto as wp

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "ref_golden.json")))
ARBITER = os.path.join(kb.HERE, "bin", "gem-arbiter")


class Arbiter:
    def __init__(self, tmp, cfg, base=300, minq=20, window=10000, extra=()):
        self.tmp = tmp
        self.cfg_path = os.path.join(tmp, "cfg.txt")
        with open(self.cfg_path, "w") as f:
            f.write(cfg)
        self.port = wp.free_port()
        self.pool = os.path.join(tmp, "pool")
        self.proc = sp.Popen([ARBITER, "--pool", self.pool, "-p", tmp, "-f", "cfg.txt", "-P", str(self.port), "-q", str(base),
                              "-m", str(minq), "-w", str(window), *extra], stderr=sp.PIPE)
        deadline = time.time() + 10
        while time.time() < deadline:
            line = self.proc.stderr.readline().decode()
            if "listening" in line:
                return
        raise RuntimeError("gem-arbiter did not start")

    def close(self):
        self.proc.kill()
        self.proc.wait()


def test_memory_counter_trace_matches_live_reference_pmgr():
    g = G["live_pmgr_mem"]
    with tempfile.TemporaryDirectory() as tmp:
        a = Arbiter(tmp, "1\nns/pod 1.0 1.0 %d\n" % g["limit"])
        try:
            conns = [wp.Client("127.0.0.1", a.port, "ns/pod"), wp.Client("127.0.0.1", a.port, "ns/pod")]
            for op in g["ops"]:
                c = conns[op["conn"]] if op["op"] != "disconnect" else None
                if op["op"] == "alloc":
                    assert c.mem_limit() == (op["used_before"], op["total"])
                    assert c.mem_update(op["bytes"], 1) == op["verdict"]
                elif op["op"] == "free":
                    assert c.mem_update(op["bytes"], 0) == op["verdict"]
                else:
                    conns[0].close()
                    time.sleep(0.2)
                other = conns[1] if op["op"] == "disconnect" else c
                assert other.mem_limit()[0] == op["used_after"]
            conns[1].close()
        finally:
            a.close()


def test_quota_answers_follow_the_pod_manager_rule_then_the_scheduler_policy():
    with tempfile.TemporaryDirectory() as tmp:
        a = Arbiter(tmp, "2\nns/a 0.5 1.0 8589934592\nns/b 0.25 1.0 1073741824\n", base=250, minq=100)
        try:
            c = wp.Client("127.0.0.1", a.port, "ns/a")
            assert c.quota(0.0, 0.0) == 250.0                      # forwarded: no pod token yet -> base quota
            q = c.quota(0.25, 7.0)                                 # fits the pod token: remaining quota, no scheduling
            assert 240.0 < q < 250.0
            assert c.quota(3.5, 2.0e9) == 10000.0                  # does not fit: forwarded, EMA clamped to max_frac*window
            assert c.mem_limit() == (0, 8589934592)
            # unknown client: ignored without a reply, exactly like gem-schd (scheduler.cpp:411-414)
            u = wp.Client("127.0.0.1", a.port, "ns/unknown", timeout=0.5)
            with pytest.raises(Exception):
                u.mem_limit()
        finally:
            a.close()


def test_quota_file_reload_and_mirror():
    with tempfile.TemporaryDirectory() as tmp:
        mirror = os.path.join(tmp, "mirror.txt")
        a = Arbiter(tmp, "1\nns/a 0.5 1.0 100\n", extra=("--mirror", mirror))
        try:
            assert open(mirror).read() == "1\nns/a 0.5 1.0 100\n"
            new = "2\nns/a 0.5 1.0 100\nns/late 0.5 1.0 4242\n"
            with open(a.cfg_path, "w") as f:  # kubeshare-config rewrites the file; close() triggers IN_CLOSE_WRITE
                f.write(new)
            deadline = time.time() + 5
            while time.time() < deadline and open(mirror).read() != new:
                time.sleep(0.05)
            assert open(mirror).read() == new
            c = wp.Client("127.0.0.1", a.port, "ns/late")
            assert c.mem_limit() == (0, 4242)
        finally:
            a.close()


def test_legacy_tcp_client_and_native_pool_client_share_one_token():
    """A legacy hook (TCP -> gem-arbiter) and a native hook (GEMHOOK_POOL) on the same GPU: one ledger."""
    with tempfile.TemporaryDirectory() as tmp:
        cfg = "2\nbench/native 0.5 1.0 8589934592\nbench/legacy 0.5 1.0 8589934592\n"
        a = Arbiter(tmp, cfg, base=40, minq=5)
        try:
            env = dict(os.environ, LD_LIBRARY_PATH=kb.STUB_DIR + ":" + os.environ.get("LD_LIBRARY_PATH", ""),
                       LD_PRELOAD=kb.LIB_PATH, GEMHOOK_POOL=a.pool, POD_NAME="bench/native",
                       GEMHOOK_STATS_FILE=os.path.join(tmp, "stats.json"))
            env.pop("GEMHOOK_QUOTA_FILE", None)  # the pool was populated by the arbiter
            native = sp.Popen([kb.STORM_PATH, "--mode", "storm", "--steps", "4", "--warmup", "0", "--step-launches", "40000",
                               "--sync-every", "500"], env=env, stdout=sp.PIPE, stderr=sp.PIPE)
            legacy = wp.Client("127.0.0.1", a.port, "bench/legacy")
            grants = 0
            while native.poll() is None:
                q = legacy.quota(0.0, 30.0)
                assert q > 0
                grants += 1
                time.sleep(min(q, 20.0) / 1e3)   # "use" part of the token, then come back
            out, err = native.communicate()
            assert native.returncode == 0, err.decode()[-800:]
            st = json.load(open(os.path.join(tmp, "stats.json")))
            assert st["launches"] == 160000 and st["token_requests"] >= 3 and grants >= 3
            L = kb.lib()
            p = L.gemhook_pool_open(a.pool.encode(), 0, 0, 0, 0, 0)
            acc = [L.gemhook_pool_accumulated_ms(p, L.gemhook_pool_find(p, n)) for n in (b"bench/native", b"bench/legacy")]
            assert all(x > 0 for x in acc), acc
            L.gemhook_pool_close(p)
            legacy.close()
        finally:
            a.close()


def test_poolctl_usage_export():
    """per-client usage out of the pool, JSON and Prometheus text."""
    ctl = os.path.join(kb.HERE, "bin", "gem-poolctl")
    with tempfile.TemporaryDirectory() as tmp:
        pool, qf = os.path.join(tmp, "pool"), os.path.join(tmp, "q.txt")
        with open(qf, "w") as f:
            f.write("2\nns/a 1.0 0.5 1000\nns/b 1.0 0.25 2000\n")   # kubeshare-config order: limit request
        assert sp.check_output([ctl, pool, "load", qf, "limit_request"]).strip() == b"2"
        L = kb.lib()
        p = L.gemhook_pool_open(pool.encode(), 0, 0, 0, 0, 0)
        assert L.gemhook_pool_mem_reserve(p, 1, 1234) == 1
        q = L.gemhook_pool_acquire(p, 0, 0.0, 0.0)
        assert q == 300.0
        info = kb.SlotInfo()
        assert L.gemhook_pool_slot_info(p, 1, info) == 0
        assert (info.name, info.min_frac, info.max_frac, info.mem_used, info.mem_limit) == (b"ns/b", 0.25, 1.0, 1234, 2000)
        dump = json.loads(sp.check_output([ctl, pool, "dump"]))
        assert [d["pod"] for d in dump] == ["ns/a", "ns/b"]
        assert (dump[0]["request"], dump[0]["limit"], dump[0]["tokens"], dump[0]["holds_token"]) == (0.5, 1.0, 1, 1)
        assert dump[1]["mem_used"] == 1234
        prom = sp.check_output([ctl, pool, "prom"]).decode()
        assert 'gemhook_mem_used_bytes{pod="ns/b",request="0.25",limit="1"} 1234' in prom
        assert 'gemhook_token_seconds_total{pod="ns/a"} 0.300000' in prom
        L.gemhook_pool_close(p)


def test_port_file_listeners_follow_kubeshare_config():
    """Launcher parity: podmanagerport/<UUID> rows appear -> listeners open; rows vanish -> they close."""
    with tempfile.TemporaryDirectory() as tmp:
        pf = os.path.join(tmp, "ports")
        p1, p2 = wp.free_port(), wp.free_port()
        with open(pf, "w") as f:
            f.write("1\nns/a %d\n" % p1)
        a = Arbiter(tmp, "2\nns/a 0.5 1.0 100\nns/b 0.5 1.0 200\n", extra=("--port-file", pf))
        try:
            time.sleep(0.3)
            assert wp.Client("127.0.0.1", p1, "ns/a").mem_limit() == (0, 100)
            with open(pf, "w") as f:
                f.write("2\nns/a %d\nns/b %d\n" % (p1, p2))
            deadline = time.time() + 5
            ok = False
            while time.time() < deadline and not ok:
                try:
                    ok = wp.Client("127.0.0.1", p2, "ns/b", timeout=1).mem_limit() == (0, 200)
                except OSError:
                    time.sleep(0.05)
            assert ok
            with open(pf, "w") as f:
                f.write("1\nns/b %d\n" % p2)
            deadline = time.time() + 5
            closed = False
            while time.time() < deadline and not closed:
                try:
                    c = wp.Client("127.0.0.1", p1, "ns/a", timeout=1)
                    c.mem_limit()
                    c.close()
                    time.sleep(0.05)
                except (OSError, ConnectionError):
                    closed = True
            assert closed
            assert wp.Client("127.0.0.1", a.port, "ns/a").mem_limit() == (0, 100)   # the static -P listener stays
        finally:
            a.close()


def test_launcher_script_starts_one_arbiter_per_gpu_uuid():
    """f4 launcher parity: tools/launcher-multigpus-b200.sh replaces launcher-multigpus.sh + launcher.py (reference
    launcher-multigpus.sh:21-42): for every GPU UUID nvidia-smi reports it pre-creates the `0` files, starts one
    gem-arbiter that owns that GPU's pool, mirrors the quota file onto the hostPath and opens the pods' ports."""
    launcher = os.path.join(kb.HERE, "tools", "launcher-multigpus-b200.sh")
    with tempfile.TemporaryDirectory() as tmp:
        cfg, ports, lib, fake = (os.path.join(tmp, d) for d in ("config", "ports", "library", "bin"))
        for d in (cfg, ports, lib, fake):
            os.makedirs(d)
        with open(os.path.join(fake, "nvidia-smi"), "w") as f:     # two GPUs
            f.write("#!/bin/sh\necho GPU-aaaa\necho GPU-bbbb\n")
        os.chmod(os.path.join(fake, "nvidia-smi"), 0o755)
        port = wp.free_port()
        with open(os.path.join(cfg, "GPU-aaaa"), "w") as f:        # kubeshare-config's column order: limit request
            f.write("1\nns/pod 1.0 0.5 4096\n")
        with open(os.path.join(ports, "GPU-aaaa"), "w") as f:
            f.write("1\nns/pod %d\n" % port)
        env = dict(os.environ, PATH=fake + ":" + os.environ["PATH"])
        p = sp.Popen(["bash", launcher, cfg, ports, lib], env=env, stderr=sp.PIPE, start_new_session=True)
        try:
            deadline = time.time() + 10
            while time.time() < deadline and not (os.path.exists(os.path.join(lib, "gemhook-GPU-bbbb.pool")) and
                                                  os.path.exists(os.path.join(lib, "config-GPU-aaaa"))):
                time.sleep(0.05)
            assert open(os.path.join(cfg, "GPU-bbbb")).read().strip() == "0"      # pre-created like launcher-multigpus.sh:26-31
            assert open(os.path.join(lib, "config-GPU-aaaa")).read() == "1\nns/pod 1.0 0.5 4096\n"   # hostPath mirror
            for _ in range(100):                                                   # the pod's port is served by GPU a's arbiter
                try:
                    c = wp.Client("127.0.0.1", port, "ns/pod")
                    break
                except OSError:
                    time.sleep(0.05)
            assert c.mem_limit() == (0, 4096)
            assert c.quota(0.0, 0.0) == 300.0                                      # launcher.py:77-80 defaults
            c.close()
            L = kb.lib()
            h = L.gemhook_pool_open(os.path.join(lib, "gemhook-GPU-aaaa.pool").encode(), 0, 0, 0, 0, 0)
            import ctypes as C
            info = kb.SlotInfo()
            assert L.gemhook_pool_slot_info(h, 0, C.byref(info)) == 0
            assert (info.min_frac, info.max_frac) == (0.5, 1.0)                    # --columns limit_request applied
            L.gemhook_pool_close(h)
        finally:
            os.killpg(p.pid, 15)
            p.wait(timeout=10)
