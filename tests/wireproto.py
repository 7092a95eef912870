"""Test-side codec for the Gemini token protocol (80-byte requests / 40-byte responses).

Independent Python statement of reference Gemini/src/comm.cpp:26-120 (layout in SURVEY.md 8b),
used by the tests to talk to the live reference daemons (oracle/_ref/gem-schd, gem-pmgr), to our
native arbiter, and to play a fake scheduler / fake pod manager against both hooks.
"""
import socket
import struct

REQ_QUOTA, REQ_MEM_LIMIT, REQ_MEM_UPDATE = 0, 1, 2
REQ_LEN, RSP_LEN = 80, 40


def pack_request(name, req_id, rtype, overuse=0.0, burst=0.0, nbytes=0, is_alloc=0):
    nb = name.encode()
    out = struct.pack("<Q", len(nb)) + nb + b"\0" + struct.pack("<ii", req_id, rtype)
    if rtype == REQ_QUOTA:
        out += struct.pack("<dd", overuse, burst)
    elif rtype == REQ_MEM_UPDATE:
        out += struct.pack("<Qi", nbytes, is_alloc)
    assert len(out) <= REQ_LEN, "name too long for the fixed 80-byte request"
    return out + b"\0" * (REQ_LEN - len(out))


def unpack_request(buf):
    (n,) = struct.unpack_from("<Q", buf, 0)
    name = buf[8:8 + n].decode()
    pos = 8 + n + 1
    req_id, rtype = struct.unpack_from("<ii", buf, pos)
    pos += 8
    d = {"name": name, "id": req_id, "type": rtype}
    if rtype == REQ_QUOTA:
        d["overuse"], d["burst"] = struct.unpack_from("<dd", buf, pos)
    elif rtype == REQ_MEM_UPDATE:
        d["bytes"], d["alloc"] = struct.unpack_from("<Qi", buf, pos)
    return d


def pack_response(rtype, req_id, quota=0.0, used=0, total=0, verdict=0):
    out = struct.pack("<i", req_id)
    if rtype == REQ_QUOTA:
        out += struct.pack("<d", quota)
    elif rtype == REQ_MEM_LIMIT:
        out += struct.pack("<QQ", used, total)
    elif rtype == REQ_MEM_UPDATE:
        out += struct.pack("<i", verdict)
    return out + b"\0" * (RSP_LEN - len(out))


def unpack_response(buf, rtype):
    (req_id,) = struct.unpack_from("<i", buf, 0)
    d = {"id": req_id}
    if rtype == REQ_QUOTA:
        (d["quota"],) = struct.unpack_from("<d", buf, 4)
    elif rtype == REQ_MEM_LIMIT:
        d["used"], d["total"] = struct.unpack_from("<QQ", buf, 4)
    elif rtype == REQ_MEM_UPDATE:
        (d["verdict"],) = struct.unpack_from("<i", buf, 4)
    return d


def recv_exact(sock, n):
    buf = b""
    while len(buf) < n:
        chunk = sock.recv(n - len(buf))
        if not chunk:
            raise ConnectionError("peer closed")
        buf += chunk
    return buf


class Client:
    """A hook-like protocol client."""

    def __init__(self, host, port, name, timeout=20.0):
        self.name = name
        self.next_id = 0
        self.sock = socket.create_connection((host, port), timeout=timeout)
        self.sock.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)

    def call(self, rtype, **kw):
        req = pack_request(self.name, self.next_id, rtype, **kw)
        self.next_id += 1
        self.sock.sendall(req)
        return unpack_response(recv_exact(self.sock, RSP_LEN), rtype), req

    def quota(self, overuse, burst):
        return self.call(REQ_QUOTA, overuse=overuse, burst=burst)[0]["quota"]

    def mem_limit(self):
        r = self.call(REQ_MEM_LIMIT)[0]
        return r["used"], r["total"]

    def mem_update(self, nbytes, is_alloc):
        return self.call(REQ_MEM_UPDATE, nbytes=nbytes, is_alloc=is_alloc)[0]["verdict"]

    def close(self):
        try:
            self.sock.close()
        except OSError:
            pass


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p
