#!/usr/bin/env python
"""Generate tests/golden/ref_golden.json from the REFERENCE's own code.

Runs only in the build container (needs /root/reference and `make -C oracle ref golden-bin`):

  * oracle/_ref/ref_golden  -- the reference's comm.o / predictor.o / scheduler.o driven under a
    virtual clock (oracle/ref_golden.cpp): wire bytes, predictor traces, select_candidate /
    get_quota / Record ledgers.
  * oracle/_ref/gem-schd, gem-pmgr -- the live reference daemons on loopback: known answers for
    REQ_MEM_LIMIT / REQ_MEM_UPDATE / REQ_QUOTA, gem-pmgr's pod-wide gpu_mem counter
    incl. reclaim on disconnect, and gem-pmgr's forwarding rule observed with a fake scheduler.

The output is committed; the GPU box never needs /root/reference.
    python tests/golden/make_golden.py
"""
import json
import os
import random
import socket
import subprocess as sp
import sys
import tempfile
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import wireproto as wp  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref")
DRV = os.path.join(REF, "ref_golden")


def drv(*args):
    out = sp.run([DRV, *map(str, args)], check=True, stdout=sp.PIPE, stderr=sp.DEVNULL).stdout
    return json.loads(out)


def wait_port(port, timeout=10.0):
    t0 = time.time()
    while time.time() - t0 < timeout:
        try:
            socket.create_connection(("127.0.0.1", port), timeout=0.5).close()
            return
        except OSError:
            time.sleep(0.05)
    raise RuntimeError("port %d never opened" % port)


def start_schd(tmp, cfg_text, base, minq, window):
    with open(os.path.join(tmp, "cfg.txt"), "w") as f:
        f.write(cfg_text)
    port = wp.free_port()
    p = sp.Popen([os.path.join(REF, "gem-schd"), "-p", tmp, "-f", "cfg.txt", "-P", str(port), "-q", str(base),
                  "-m", str(minq), "-w", str(window)], stdout=sp.DEVNULL, stderr=sp.DEVNULL)
    # do not probe-connect: every accepted connection spawns a thread in gem-schd, harmless but noisy
    time.sleep(0.5)
    return p, port


def live_schd():
    """BASELINE.md 2 known answers + a longer EMA sequence, straight from the reference binary."""
    cfg = "2\nns/a 0.5 1.0 8589934592\nns/b 0.25 1.0 1073741824\n"
    out = {"config": cfg, "base": 250, "min": 100, "window": 10000, "calls": []}
    with tempfile.TemporaryDirectory() as tmp:
        p, port = start_schd(tmp, cfg, 250, 100, 10000)
        try:
            c = wp.Client("127.0.0.1", port, "ns/a")
            used, total = c.mem_limit()
            out["calls"].append({"op": "mem_limit", "used": used, "total": total})
            out["calls"].append({"op": "mem_update", "bytes": 4096, "alloc": 1, "verdict": c.mem_update(4096, 1)})
            for burst in [0.0, 5.0, 10.0, 400.0, 400.0, 20000.0, 20000.0, 20000.0, 0.0, 30.0]:
                q = c.quota(0.0, burst)
                out["calls"].append({"op": "quota", "overuse": 0.0, "burst": burst, "quota": q})
                time.sleep(0.01)
            c.close()
        finally:
            p.kill()
            p.wait()
    return out


def live_pmgr_mem():
    """gem-pmgr's authoritative counter (pod-manager.cpp:295-313, 501-504, 533-545)."""
    limit = 8589934592
    cfg = "1\nns/pod 1.0 1.0 %d\n" % limit
    rng = random.Random(4)
    out = {"limit": limit, "ops": []}
    with tempfile.TemporaryDirectory() as tmp:
        ps, sport = start_schd(tmp, cfg, 300, 20, 10000)
        pport = wp.free_port()
        env = dict(os.environ, POD_NAME="ns/pod", POD_MANAGER_PORT=str(pport), SCHEDULER_IP="127.0.0.1",
                   SCHEDULER_PORT=str(sport))
        pm = sp.Popen([os.path.join(REF, "gem-pmgr")], env=env, stdout=sp.DEVNULL, stderr=sp.DEVNULL)
        try:
            time.sleep(0.5)
            conns = [wp.Client("127.0.0.1", pport, "ns/pod"), wp.Client("127.0.0.1", pport, "ns/pod")]
            held = [[], []]
            for i in range(160):
                k = rng.randrange(2)
                c = conns[k]
                r = rng.random()
                if r < 0.68 or not held[k]:
                    size = rng.choice([rng.randrange(1, 64 << 20), 256 << 20, 1 << 30, (1 << 30) + 1, 2 << 30, rng.randrange(1, 4096)])
                    used, total = c.mem_limit()
                    v = c.mem_update(size, 1)
                    if v:
                        held[k].append(size)
                    out["ops"].append({"conn": k, "op": "alloc", "bytes": size, "used_before": used, "total": total,
                                       "verdict": v})
                else:
                    size = held[k].pop(rng.randrange(len(held[k])))
                    v = c.mem_update(size, 0)
                    out["ops"].append({"conn": k, "op": "free", "bytes": size, "verdict": v})
                used, total = c.mem_limit()
                out["ops"][-1]["used_after"] = used
            # reclaim on disconnect
            conns[0].close()
            time.sleep(0.3)
            used, total = conns[1].mem_limit()
            out["ops"].append({"conn": 0, "op": "disconnect", "used_after": used})
            conns[1].close()
        finally:
            pm.kill()
            pm.wait()
            ps.kill()
            ps.wait()
    return out


def live_pmgr_forward():
    """gem-pmgr's hook_kernel_launch forwarding rule (pod-manager.cpp:316-473) against a fake scheduler."""
    out = {"limit": 123456789, "steps": []}
    lsock = socket.socket()
    lsock.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
    lsock.bind(("127.0.0.1", 0))
    lsock.listen(4)
    sport = lsock.getsockname()[1]
    pport = wp.free_port()
    env = dict(os.environ, POD_NAME="ns/pod", POD_MANAGER_PORT=str(pport), SCHEDULER_IP="127.0.0.1",
               SCHEDULER_PORT=str(sport))
    pm = sp.Popen([os.path.join(REF, "gem-pmgr")], env=env, stdout=sp.DEVNULL, stderr=sp.DEVNULL)
    try:
        s, _ = lsock.accept()
        first = wp.unpack_request(wp.recv_exact(s, wp.REQ_LEN))
        out["hello"] = first
        s.sendall(wp.pack_response(wp.REQ_MEM_LIMIT, first["id"], used=0, total=out["limit"]))
        time.sleep(0.3)
        c = wp.Client("127.0.0.1", pport, "ns/pod")
        # (overuse, burst, quota the fake scheduler will answer with, or None when pmgr must answer locally)
        script = [(0.0, 0.0, 1.0e9), (0.25, 7.0, None), (3.5, 2.0e9, 500.0), (1.0, 100.0, None)]
        for overuse, burst, reply in script:
            req = wp.pack_request("ns/pod", c.next_id, wp.REQ_QUOTA, overuse=overuse, burst=burst)
            c.next_id += 1
            c.sock.sendall(req)
            step = {"overuse": overuse, "burst": burst, "forwarded": None}
            if reply is not None:
                fwd_raw = wp.recv_exact(s, wp.REQ_LEN)
                fwd = wp.unpack_request(fwd_raw)
                step["forwarded"] = {"name": fwd["name"], "overuse": fwd["overuse"], "burst": fwd["burst"]}
                s.sendall(wp.pack_response(wp.REQ_QUOTA, fwd["id"], quota=reply))
                step["schd_quota"] = reply
            rsp = wp.unpack_response(wp.recv_exact(c.sock, wp.RSP_LEN), wp.REQ_QUOTA)
            step["reply_quota"] = rsp["quota"]
            out["steps"].append(step)
        c.close()
    finally:
        pm.kill()
        pm.wait()
        lsock.close()
    return out


def main():
    if not os.path.exists(DRV):
        sys.exit("build first: make -C oracle ref golden-bin")
    g = {"_made_by": "tests/golden/make_golden.py from /root/reference (KubeShare a862311c / Gemini 953052b9)"}
    g["wire"] = [drv("wire", n) for n in ["ns/a", "default/mnist-pod-0123456789", "x" * 47, "bench/c0"]]
    g["predictor"] = [drv("predictor", seed, 400, thres) for seed, thres in [(7, 2.0), (11, 0.0), (0xB200, 2.0)]]
    g["schd"] = []
    with tempfile.TemporaryDirectory() as tmp:
        cfgs = {
            "gemini3.txt": "3\nclient1 0.1 0.5 1073741824\nclient2 0.2 0.8 1073741824\nclient3 0.4 0.5 2147483648\n",
            "half2.txt": "2\nbench/c0 0.5 1.0 8589934592\nbench/c1 0.5 1.0 8589934592\n",
            "tight3.txt": "3\nns/a 0.1 0.3 1\nns/b 0.1 0.2 2\nns/c 0.2 0.4 3\n",
            "quarter4.txt": "4\nbench/c0 0.25 1.0 1\nbench/c1 0.25 1.0 2\nbench/c2 0.25 1.0 3\nbench/c3 0.25 1.0 4\n",
        }
        for fn, txt in cfgs.items():
            with open(os.path.join(tmp, fn), "w") as f:
                f.write(txt)
        # (seed, steps, cfg, base, min, window, mean gap ms, history every)
        for seed, steps, fn, base, minq, win, gap, he in [
            (3, 120, "gemini3.txt", 300, 20, 10000, 50, 10),
            (5, 150, "half2.txt", 300, 20, 2000, 40, 10),
            (9, 150, "quarter4.txt", 250, 100, 1500, 120, 10),
            (0xB200, 200, "quarter4.txt", 300, 20, 3000, 15, 20),
            (21, 200, "tight3.txt", 100, 20, 1000, 30, 20),
            (22, 150, "tight3.txt", 300, 20, 10000, 80, 15),
        ]:
            d = drv("schd", seed, steps, tmp, fn, base, minq, win, gap, he)
            d["config_text"] = cfgs[fn]
            g["schd"].append(d)
    g["live_schd"] = live_schd()
    g["live_pmgr_mem"] = live_pmgr_mem()
    g["live_pmgr_forward"] = live_pmgr_forward()
    path = os.path.join(HERE, "ref_golden.json")
    with open(path, "w") as f:
        json.dump(g, f, separators=(",", ":"))
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
