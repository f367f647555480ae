"""Minimal UBJSON reader/writer (xgboost model dialect) -- TEST INFRASTRUCTURE.

Independent of the product's C++ model IO (csrc/model_io.cc) so that each can check the other.
Schema followed: SURVEY.md section 8(c); fixture: the reference's
test/resources/abalone/models/libsvm_pickled/xgboost-model (copied to tests/golden/).
Markers handled: { [ $ # S i U I l L d D T F Z C ; big-endian payloads.
"""
import struct

import numpy as np

_NUM = {b"i": (">b", 1), b"U": (">B", 1), b"I": (">h", 2), b"l": (">i", 4), b"L": (">q", 8),
        b"d": (">f", 4), b"D": (">d", 8)}
_NPT = {b"i": ">i1", b"U": ">u1", b"I": ">i2", b"l": ">i4", b"L": ">i8", b"d": ">f4", b"D": ">f8"}


class _R:
    def __init__(self, b):
        self.b = b
        self.i = 0

    def take(self, n):
        v = self.b[self.i:self.i + n]
        self.i += n
        return v

    def peek(self):
        return self.b[self.i:self.i + 1]

    def num(self, m):
        fmt, n = _NUM[m]
        return struct.unpack(fmt, self.take(n))[0]

    def length(self):
        return self.num(self.take(1))

    def string(self):
        n = self.length()
        return self.take(n).decode("utf-8")

    def value(self, m=None):
        m = m or self.take(1)
        if m in _NUM:
            return self.num(m)
        if m == b"S":
            return self.string()
        if m == b"C":
            return self.take(1).decode()
        if m == b"T":
            return True
        if m == b"F":
            return False
        if m == b"Z":
            return None
        if m == b"[":
            return self.array()
        if m == b"{":
            return self.obj()
        raise ValueError("ubjson: bad marker %r at %d" % (m, self.i))

    def array(self):
        typ = None
        cnt = None
        if self.peek() == b"$":
            self.take(1)
            typ = self.take(1)
        if self.peek() == b"#":
            self.take(1)
            cnt = self.length()
        if typ is not None:
            if typ in _NPT:
                dt = np.dtype(_NPT[typ])
                a = np.frombuffer(self.take(cnt * dt.itemsize), dtype=dt).astype(dt.newbyteorder("="))
                return a
            return [self.value(typ) for _ in range(cnt)]
        out = []
        if cnt is not None:
            return [self.value() for _ in range(cnt)]
        while self.peek() != b"]":
            out.append(self.value())
        self.take(1)
        return out

    def obj(self):
        cnt = None
        if self.peek() == b"#":
            self.take(1)
            cnt = self.length()
        out = {}
        if cnt is not None:
            for _ in range(cnt):
                k = self.string()
                out[k] = self.value()
            return out
        while self.peek() != b"}":
            k = self.string()
            out[k] = self.value()
        self.take(1)
        return out


def loads(b):
    return _R(bytes(b)).value()


def load(path):
    with open(path, "rb") as f:
        return loads(f.read())


def _wlen(n):
    return b"L" + struct.pack(">q", n)


def _wstr(s):
    e = s.encode("utf-8")
    return _wlen(len(e)) + e


def dumps(v):
    if isinstance(v, dict):
        return b"{" + b"".join(_wstr(k) + dumps(x) for k, x in v.items()) + b"}"
    if isinstance(v, np.ndarray):
        code = {"f4": b"d", "f8": b"D", "i4": b"l", "i8": b"L", "u1": b"U", "i1": b"i", "i2": b"I"}[v.dtype.str[1:]]
        return b"[$" + code + b"#" + _wlen(v.size) + v.astype(v.dtype.newbyteorder(">")).tobytes()
    if isinstance(v, (list, tuple)):
        return b"[" + b"".join(dumps(x) for x in v) + b"]"
    if isinstance(v, bool):
        return b"T" if v else b"F"
    if v is None:
        return b"Z"
    if isinstance(v, str):
        return b"S" + _wstr(v)
    if isinstance(v, (int, np.integer)):
        return b"L" + struct.pack(">q", int(v))
    if isinstance(v, (float, np.floating)):
        return b"d" + struct.pack(">f", float(v))
    raise TypeError(type(v))


def model_from_xgb_json(doc):
    """xgboost model document (parsed UBJSON/JSON) -> oracle.Model-style flat arrays."""
    from .gbt_oracle import Model
    learner = doc["learner"]
    gb = learner["gradient_booster"]["model"]
    trees = gb["trees"]
    m = Model()
    offs = [0]
    keys = {"left": "left_children", "right": "right_children", "parent": "parents", "split_index": "split_indices",
            "default_left": "default_left", "split_cond": "split_conditions", "base_weight": "base_weights",
            "loss_chg": "loss_changes", "sum_hess": "sum_hessian"}
    acc = {k: [] for k in keys}
    for t in trees:
        n = int(t["tree_param"]["num_nodes"])
        offs.append(offs[-1] + n)
        for k, src in keys.items():
            acc[k].append(np.asarray(t[src]))
    dt = {"left": np.int32, "right": np.int32, "parent": np.int32, "split_index": np.int32, "default_left": np.uint8,
          "split_cond": np.float32, "base_weight": np.float32, "loss_chg": np.float32, "sum_hess": np.float32}
    for k in keys:
        m[k] = np.ascontiguousarray(np.concatenate(acc[k]) if acc[k] else np.zeros(0), dt[k])
    m["split_bin"] = np.full(len(m["left"]), -1, np.int32)
    m["tree_offset"] = np.asarray(offs, np.int64)
    m["tree_info"] = np.ascontiguousarray(np.asarray(gb["tree_info"]), np.int32)
    lmp = learner["learner_model_param"]
    bs = lmp["base_score"]
    m["base_score"] = float(bs.strip("[]")) if isinstance(bs, str) else float(bs)
    m["num_class"] = max(1, int(lmp.get("num_class", "0")))
    m["num_feature"] = int(lmp["num_feature"])
    m["objective"] = learner["objective"]["name"]
    return m
