"""`xgboost.tracker.RabitTracker` replacement (distributed.py:183-190,247-257) plus the worker-side client.

The tracker is a tiny TCP hub: workers connect, announce their task id, receive (rank, world) -- ranks follow the
sorted task ids when sortby="task", as the container requests -- and later relay host-side object broadcasts
(the container's RabitHelper.synchronize) through it.  Bulk numeric traffic never goes here: histograms travel
over NCCL inside the engine.
"""
import base64
import json
import socket
import struct
import threading
import time

# Wire format: 8-byte big-endian length + UTF-8 JSON.  No pickle: the tracker listens on the cluster-facing address
# (distributed.py:183-190 binds it to the master's IP), so a frame must never be able to execute code on its reader.
MAX_FRAME = 64 << 20          # broadcast payloads are small membership records (RabitHelper.synchronize)
MAX_HELLO = 4096
HELLO_TIMEOUT_S = 10.0
MAGIC = "b200xgb-tracker-1"


def _encode(obj):
    """JSON with two extensions: bytes and tuples survive the round trip."""
    if isinstance(obj, (bytes, bytearray)):
        return {"__bytes__": base64.b64encode(bytes(obj)).decode("ascii")}
    if isinstance(obj, tuple):
        return {"__tuple__": [_encode(x) for x in obj]}
    if isinstance(obj, list):
        return [_encode(x) for x in obj]
    if isinstance(obj, dict):
        for k in obj:
            if not isinstance(k, str):
                raise TypeError("collective.broadcast: dictionary keys must be strings (got %r)" % (k,))
        return {k: _encode(v) for k, v in obj.items()}
    if obj is None or isinstance(obj, (str, bool, int, float)):
        return obj
    try:                                  # numpy scalars and similar
        return _encode(obj.item())
    except Exception:
        raise TypeError("collective.broadcast carries JSON-serialisable values (dict / list / tuple / str / number / bool / None / "
                        "bytes); got %s" % type(obj).__name__)


def _decode(obj):
    if isinstance(obj, list):
        return [_decode(x) for x in obj]
    if isinstance(obj, dict):
        if set(obj) == {"__bytes__"}:
            return base64.b64decode(obj["__bytes__"])
        if set(obj) == {"__tuple__"}:
            return tuple(_decode(x) for x in obj["__tuple__"])
        return {k: _decode(v) for k, v in obj.items()}
    return obj


def _send(sock, obj):
    data = json.dumps(_encode(obj), separators=(",", ":")).encode("utf-8")
    if len(data) > MAX_FRAME:
        raise ValueError("tracker frame of %d bytes exceeds the %d byte limit" % (len(data), MAX_FRAME))
    sock.sendall(struct.pack("!Q", len(data)) + data)


def _recv(sock, limit=MAX_FRAME):
    hdr = b""
    while len(hdr) < 8:
        chunk = sock.recv(8 - len(hdr))
        if not chunk:
            raise ConnectionError("tracker connection closed")
        hdr += chunk
    (n,) = struct.unpack("!Q", hdr)
    if n > limit:
        raise ConnectionError("tracker frame of %d bytes exceeds the %d byte limit" % (n, limit))
    buf = bytearray()
    while len(buf) < n:
        chunk = sock.recv(min(1 << 20, n - len(buf)))
        if not chunk:
            raise ConnectionError("tracker connection closed")
        buf += chunk
    try:
        return _decode(json.loads(bytes(buf).decode("utf-8")))
    except (ValueError, UnicodeDecodeError) as e:
        raise ConnectionError("malformed tracker frame: %s" % e)


def _alive(sock):
    """False once the peer has closed its end (a zero-byte peek); a quiet but open link counts as alive."""
    try:
        sock.setblocking(False)
        try:
            return sock.recv(1, socket.MSG_PEEK) != b""
        except (BlockingIOError, InterruptedError):
            return True
        except OSError:
            return False
    finally:
        try:
            sock.setblocking(True)
        except OSError:
            pass


class RabitTracker:
    def __init__(self, n_workers, host_ip="127.0.0.1", port=0, sortby="host", timeout=0):
        self.n_workers = int(n_workers)
        self.host_ip = host_ip
        self.sortby = sortby
        self.timeout = timeout
        self._sock = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
        self._sock.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
        self._sock.bind(("" if host_ip in ("0.0.0.0", "") else host_ip, int(port)))
        self.port = self._sock.getsockname()[1]
        self._sock.listen(max(16, self.n_workers * 2))
        self._thread = None
        self._done = threading.Event()
        self._error = None

    def worker_args(self):
        return {"dmlc_tracker_uri": self.host_ip, "dmlc_tracker_port": self.port}

    def start(self):
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()

    def _run(self):
        try:
            conns = []                         # (task id, host, arrival order, socket)
            order = 0
            while len(conns) < self.n_workers:
                c, addr = self._sock.accept()
                try:
                    c.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
                    c.settimeout(HELLO_TIMEOUT_S)
                    hello = _recv(c, limit=MAX_HELLO)
                    if not isinstance(hello, dict) or hello.get("magic") != MAGIC or not isinstance(hello.get("task_id"), str):
                        raise ConnectionError("bad handshake")
                    c.settimeout(None)
                except (ConnectionError, OSError, socket.timeout):
                    try:                       # port probe, stray client or a peer that went silent: drop it, keep listening
                        c.close()
                    except OSError:
                        pass
                    continue
                task = hello["task_id"]
                # a worker that retried its CommunicatorContext left a dead link behind: the new one takes its place.  Live
                # links with the same task id are kept (the container derives the id from hosts.index(host), which repeats
                # when several workers share a host name, test/unit/test_distributed.py:26).
                for i, old in enumerate(conns):
                    if old[0] == task and not _alive(old[3]):
                        try:
                            old[3].close()
                        except OSError:
                            pass
                        conns.pop(i)
                        break
                conns.append((task, addr[0], order, c))
                order += 1
            if self.sortby == "task":
                conns.sort(key=lambda t: (t[0], t[2]))
            else:
                conns.sort(key=lambda t: (t[1], t[2]))
            socks = [c for *_, c in conns]
            for rank, s in enumerate(socks):
                _send(s, {"rank": rank, "world": self.n_workers})
            alive = set(range(self.n_workers))
            while alive:
                msgs = {}
                for r in sorted(alive):
                    try:
                        m = _recv(socks[r])
                        if not isinstance(m, dict) or m.get("op") not in ("bcast", "barrier", "bye"):
                            raise ConnectionError("malformed request")
                        msgs[r] = m
                    except (ConnectionError, OSError):
                        msgs[r] = {"op": "bye"}
                ops = {m["op"] for m in msgs.values()}
                if ops == {"bye"} or "bye" in ops:
                    for r in list(alive):
                        if msgs[r]["op"] == "bye":
                            alive.discard(r)
                            try:
                                socks[r].close()
                            except OSError:
                                pass
                    if not alive:
                        break
                    continue
                if ops == {"bcast"}:
                    root = next(iter(msgs.values())).get("root")
                    if not isinstance(root, int) or root not in msgs:
                        raise RuntimeError("tracker: broadcast from unknown root %r" % (root,))
                    payload = msgs[root].get("data")
                    for r in alive:
                        _send(socks[r], {"data": payload})
                elif ops == {"barrier"}:
                    for r in alive:
                        _send(socks[r], {"ok": True})
                else:
                    raise RuntimeError("tracker: mismatched collective ops %s" % ops)
        except Exception as e:  # surfaced by wait_for
            self._error = e
        finally:
            self._done.set()

    def wait_for(self, timeout=None):
        ok = self._done.wait(timeout if timeout and timeout > 0 else None)
        if self._error is not None:
            raise RuntimeError("tracker failed: %s" % self._error)
        if not ok:
            raise TimeoutError("tracker timed out")

    def join(self):
        self.wait_for()

    def free(self):
        try:
            self._sock.close()
        except OSError:
            pass


class TrackerClient:
    def __init__(self, uri, port, task_id, timeout=300.0):
        self.uri, self.port, self.task_id, self.timeout = uri, int(port), task_id, timeout
        self.sock = None
        self.rank, self.world = 0, 1

    def connect(self):
        deadline = time.time() + max(self.timeout, 1.0)
        last = None
        while True:
            try:
                self.sock = socket.create_connection((self.uri, self.port), timeout=10.0)
                break
            except OSError as e:
                last = e
                if time.time() > deadline:
                    raise ConnectionError("cannot reach the tracker at %s:%d: %s" % (self.uri, self.port, last))
                time.sleep(0.2)
        self.sock.settimeout(None)
        self.sock.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
        _send(self.sock, {"magic": MAGIC, "task_id": str(self.task_id)})
        info = _recv(self.sock)
        self.rank, self.world = info["rank"], info["world"]

    def broadcast(self, data, root):
        _send(self.sock, {"op": "bcast", "root": root, "data": data if self.rank == root else None})
        return _recv(self.sock)["data"]

    def barrier(self):
        _send(self.sock, {"op": "barrier"})
        _recv(self.sock)

    def close(self):
        if self.sock is not None:
            try:
                _send(self.sock, {"op": "bye"})
                self.sock.close()
            except OSError:
                pass
            self.sock = None
