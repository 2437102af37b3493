"""Torch7 `.t7` snapshots of the experiment scripts.

The training scripts write, every test interval (F:1223-1232),

    torch.save(snapshot_iterNNNNNN_epochE.EE.t7, {it = it, opt = opt, epoch = epoch,
                                                  params = {[1] = embed_param, [2] = rnn_param, [3] = mult_param}})

and `experiments/Ours_ResNet/Eval.lua` reads it back with torch.load and copies the three flat vectors into freshly flattened
parameters (EV:114, EV:345-347).  The three vectors are the `getParameters()` results of F:322-324, so their element order is
nngraph's module traversal order.  This module has

* a reader / writer of Torch7's binary serialisation (the subset such a file holds: nil, number, string, boolean, table,
  torch.*Tensor / torch.*Storage including CudaTensor, with shared-object back references), and
* the permutation between nngraph's order and librau's flat layout (`rau_param_offset`, include/rau.h).

WHAT IS PINNED AND WHAT IS NOT.  The binary format is restated from knowledge of torch7's File.lua / Tensor.c / Storage.c
(the rocks are not vendored in the reference); tests pin it with a byte string assembled independently of
this writer.  The traversal order below is DERIVED, not observed: nn.gModule lists its modules in the order of
`fg:topsort()`, which (torch/graph Graph:topsort -> Node:dfs) is a post-order depth-first walk from the output node over
each node's inputs in the order they were passed to the module call; nn.Container:parameters() then concatenates each
module's {weight, bias}.  Walking F:231-307, A:4-74 and D:14-71 that way gives NNGRAPH_ORDER.  No released snapshot is
available offline to confirm it; a flat vector carries no segment markers, so a real file can confirm only the total sizes
(`infer_sizes`).  The groups `embed` and `rnn` come out identical to librau's layout; `mult` differs in three places.

Torch7's Lua host does not need the reader (torch.load is native there): lua/rau/snapshot.lua applies the same permutation.
"""
from __future__ import annotations

import struct
from collections import OrderedDict

import numpy as np

TYPE_NIL, TYPE_NUMBER, TYPE_STRING, TYPE_TABLE, TYPE_TORCH, TYPE_BOOLEAN = 0, 1, 2, 3, 4, 5
TYPE_FUNCTION, TYPE_LEGACY_RECUR_FUNCTION, TYPE_RECUR_FUNCTION = 6, 7, 8

_STORAGE_DTYPES = {
    "torch.FloatStorage": np.float32, "torch.DoubleStorage": np.float64, "torch.CudaStorage": np.float32,
    "torch.LongStorage": np.int64, "torch.IntStorage": np.int32, "torch.ByteStorage": np.uint8,
    "torch.CudaDoubleStorage": np.float64, "torch.CudaLongStorage": np.int64,
}
_TENSOR_STORAGE = {
    "torch.FloatTensor": "torch.FloatStorage", "torch.DoubleTensor": "torch.DoubleStorage",
    "torch.CudaTensor": "torch.CudaStorage", "torch.LongTensor": "torch.LongStorage", "torch.IntTensor": "torch.IntStorage",
    "torch.ByteTensor": "torch.ByteStorage", "torch.CudaDoubleTensor": "torch.CudaDoubleStorage",
    "torch.CudaLongTensor": "torch.CudaLongStorage",
}


class T7Tensor:
    """A deserialised torch.*Tensor: `array` is the strided view materialised as a numpy array."""

    def __init__(self, cls: str, array: np.ndarray):
        self.cls, self.array = cls, array

    def __repr__(self):
        return f"T7Tensor({self.cls}, shape={self.array.shape})"


class T7Reader:
    """torch.load (binary mode, the default of torch.save): File:readObject."""

    def __init__(self, data: bytes):
        self.b, self.pos, self.objects = memoryview(data), 0, {}

    def _take(self, fmt):
        n = struct.calcsize(fmt)
        if self.pos + n > len(self.b):
            raise ValueError("truncated .t7 stream")
        v = struct.unpack_from(fmt, self.b, self.pos)
        self.pos += n
        return v[0]

    def _int(self):
        return self._take("<i")

    def _long(self):
        return self._take("<q")

    def _string(self):
        n = self._int()
        s = bytes(self.b[self.pos:self.pos + n])
        if len(s) != n:
            raise ValueError("truncated .t7 stream")
        self.pos += n
        return s.decode("latin-1")

    def read(self):
        t = self._int()
        if t == TYPE_NIL:
            return None
        if t == TYPE_NUMBER:
            return self._take("<d")
        if t == TYPE_BOOLEAN:
            return self._int() == 1
        if t == TYPE_STRING:
            return self._string()
        if t in (TYPE_FUNCTION, TYPE_LEGACY_RECUR_FUNCTION, TYPE_RECUR_FUNCTION):
            raise ValueError("the stream holds a Lua function; snapshots of the experiment scripts do not")
        if t not in (TYPE_TABLE, TYPE_TORCH):
            raise ValueError(f"unknown .t7 type tag {t} at byte {self.pos - 4}")
        index = self._int()
        if index in self.objects:           # a second reference to an object already read
            return self.objects[index]
        if t == TYPE_TABLE:
            out = OrderedDict()
            self.objects[index] = out
            for _ in range(self._int()):
                k = self.read()
                out[_key(k)] = self.read()
            return out
        version = self._string()
        cls = self._string() if version.startswith("V ") else version       # (pre-versioning files name the class directly)
        if cls in _STORAGE_DTYPES:
            n = self._long()
            dt = np.dtype(_STORAGE_DTYPES[cls])
            arr = np.frombuffer(self.b, dtype=dt, count=n, offset=self.pos).copy()
            self.pos += n * dt.itemsize
            self.objects[index] = arr
            return arr
        if cls in _TENSOR_STORAGE:
            nd = self._int()
            size = [self._long() for _ in range(nd)]
            stride = [self._long() for _ in range(nd)]
            offset = self._long() - 1                                        # storageOffset is 1-based in the file
            holder = T7Tensor(cls, np.zeros(0, dtype=_STORAGE_DTYPES[_TENSOR_STORAGE[cls]]))
            self.objects[index] = holder
            storage = self.read()
            if storage is None or nd == 0:
                return holder
            item = storage.dtype.itemsize
            holder.array = np.lib.stride_tricks.as_strided(storage[offset:], shape=size,
                                                           strides=[s * item for s in stride]).copy()
            return holder
        raise ValueError(f"unsupported torch class {cls!r} in the stream")


def _key(k):
    if isinstance(k, float) and k == int(k):
        return int(k)
    return k


class T7Writer:
    """torch.save (binary): File:writeObject for the object kinds a snapshot holds."""

    def __init__(self):
        self.chunks, self.next_index = [], 1

    def _int(self, v):
        self.chunks.append(struct.pack("<i", v))

    def _long(self, v):
        self.chunks.append(struct.pack("<q", v))

    def _string(self, s):
        raw = s.encode("latin-1")
        self._int(len(raw))
        self.chunks.append(raw)

    def write(self, obj):
        if obj is None:
            self._int(TYPE_NIL)
        elif isinstance(obj, bool):
            self._int(TYPE_BOOLEAN)
            self._int(1 if obj else 0)
        elif isinstance(obj, (int, float, np.integer, np.floating)):
            self._int(TYPE_NUMBER)
            self.chunks.append(struct.pack("<d", float(obj)))
        elif isinstance(obj, str):
            self._int(TYPE_STRING)
            self._string(obj)
        elif isinstance(obj, dict):
            self._int(TYPE_TABLE)
            self._int(self._index())
            self._int(len(obj))
            for k, v in obj.items():
                self.write(k)
                self.write(v)
        elif isinstance(obj, (list, tuple)):
            self.write({i + 1: v for i, v in enumerate(obj)})               # a Lua array: keys 1..n
        elif isinstance(obj, T7Tensor):
            self._tensor(obj.cls, obj.array)
        elif isinstance(obj, np.ndarray):
            cls = {np.dtype(np.float32): "torch.FloatTensor", np.dtype(np.float64): "torch.DoubleTensor",
                   np.dtype(np.int64): "torch.LongTensor"}[obj.dtype]
            self._tensor(cls, obj)
        else:
            raise TypeError(f"cannot serialise {type(obj)} as Torch7")

    def _index(self):
        i = self.next_index
        self.next_index += 1
        return i

    def _tensor(self, cls, arr):
        arr = np.ascontiguousarray(arr, dtype=_STORAGE_DTYPES[_TENSOR_STORAGE[cls]])
        self._int(TYPE_TORCH)
        self._int(self._index())
        self._string("V 1")
        self._string(cls)
        self._int(arr.ndim)
        for s in arr.shape:
            self._long(s)
        stride, acc = [], 1
        for s in reversed(arr.shape):
            stride.append(acc)
            acc *= s
        for s in reversed(stride):
            self._long(s)
        self._long(1)                                                        # storageOffset (1-based)
        self._int(TYPE_TORCH)                                                # the storage object
        self._int(self._index())
        self._string("V 1")
        self._string(_TENSOR_STORAGE[cls])
        self._long(arr.size)
        self.chunks.append(arr.tobytes())

    def getvalue(self) -> bytes:
        return b"".join(self.chunks)


def t7_load(path):
    with open(path, "rb") as f:
        return T7Reader(f.read()).read()


def t7_save(path, obj):
    w = T7Writer()
    w.write(obj)
    with open(path, "wb") as f:
        f.write(w.getvalue())


# ------------------------------------------------------------------ flat layouts
# librau's layout per group (include/rau.h, rau_param_offset) and nngraph's (derived above): names in flat order
LIBRAU_ORDER = {
    "embed": ["E"],
    "rnn": ["l1.Wi", "l1.bi", "l1.Wh", "l1.bh", "l2.Wi", "l2.bi", "l2.Wh", "l2.bh"],
    "mult": ["Wq", "bq", "Wh", "bh", "Wi", "bi", "Wqa", "bqa", "Wa", "ba", "ws", "Wm", "bm", "Wp", "bp", "Wx", "bx", "Whh", "bhh",
             "Wo", "bo", "Ws", "bso", "wd", "bs", "bd"],
}
NNGRAPH_ORDER = {
    "embed": ["E"],                                                           # LookupTable.weight (F:204)
    "rnn": LIBRAU_ORDER["rnn"],                                               # per layer i2h {W, b}, h2h {W, b} (D:43-44)
    # q_embed {Wq bq Wh bh} (F:233-234) | i_embed {Wi bi} (F:240) | attbycontent: the CAddTable lists ifeatatt first, so
    # the 1x1 conv {Wa ba} (F:247) precedes the query Linear {Wqa bqa} (F:246), then the score conv {ws bs} (F:251) |
    # attbymemory {Wm bm} (F:287) | classifier: {Wp bp} (F:271), attlstm {Wx bx Whh bhh} (A:6-7), {Wo bo} (F:279),
    # {Ws bso} (F:280), {wd bd} (F:281)
    "mult": ["Wq", "bq", "Wh", "bh", "Wi", "bi", "Wa", "ba", "Wqa", "bqa", "ws", "bs", "Wm", "bm", "Wp", "bp", "Wx", "bx", "Whh",
             "bhh", "Wo", "bo", "Ws", "bso", "wd", "bd"],
}


def tensor_shapes(cfg, group):
    """{name: shape} of one group for a configuration with fields V, embed, Hq, nlayer, C, S, M, A, H, N."""
    Q = 2 * cfg.Hq * cfg.nlayer
    if group == "embed":
        return {"E": (cfg.V, cfg.embed)}
    if group == "rnn":
        out = {}
        for L in range(1, cfg.nlayer + 1):
            i = cfg.embed if L == 1 else cfg.Hq
            out.update({f"l{L}.Wi": (4 * cfg.Hq, i), f"l{L}.bi": (4 * cfg.Hq,), f"l{L}.Wh": (4 * cfg.Hq, cfg.Hq), f"l{L}.bh": (4 * cfg.Hq,)})
        return out
    return {"Wq": (cfg.M, Q), "bq": (cfg.M,), "Wh": (cfg.M, cfg.H), "bh": (cfg.M,), "Wi": (cfg.M, cfg.C), "bi": (cfg.M,),
            "Wqa": (cfg.A, cfg.M), "bqa": (cfg.A,), "Wa": (cfg.A, cfg.M), "ba": (cfg.A,), "ws": (1, cfg.A), "bs": (1,),
            "Wm": (cfg.S, cfg.H), "bm": (cfg.S,), "Wp": (cfg.M, cfg.S), "bp": (cfg.M,), "Wx": (4 * cfg.H, cfg.M), "bx": (4 * cfg.H,),
            "Whh": (4 * cfg.H, cfg.H), "bhh": (4 * cfg.H,), "Wo": (cfg.M, cfg.H), "bo": (cfg.M,), "Ws": (cfg.N, cfg.M),
            "bso": (cfg.N,), "wd": (1, cfg.M), "bd": (1,)}


def _order(cfg, group, which):
    names = list((NNGRAPH_ORDER if which == "nngraph" else LIBRAU_ORDER)[group])
    if group == "rnn":
        names = [n for n in names if int(n[1]) <= cfg.nlayer]
    return names


def permutation(cfg, group, src="nngraph", dst="librau"):
    """index array `perm` with dst_flat = src_flat[perm]."""
    shapes = tensor_shapes(cfg, group)
    off, pos = {}, 0
    for name in _order(cfg, group, src):
        off[name] = pos
        pos += int(np.prod(shapes[name]))
    parts = [np.arange(off[name], off[name] + int(np.prod(shapes[name])), dtype=np.int64) for name in _order(cfg, group, dst)]
    perm = np.concatenate(parts)
    assert perm.size == pos
    return perm


def from_nngraph(cfg, group, flat):
    """a getParameters() vector of the reference -> librau's flat layout"""
    return np.asarray(flat)[permutation(cfg, group, "nngraph", "librau")]


def to_nngraph(cfg, group, flat):
    return np.asarray(flat)[permutation(cfg, group, "librau", "nngraph")]


def infer_sizes(cfg, n_embed, n_rnn, n_mult):
    """(V, N) from the lengths of a snapshot's three vectors (vocab and answer counts are data dependent, LD:1389-1416);
    raises when the lengths cannot come from this architecture -- the only check a marker-less flat vector allows."""
    if n_embed % cfg.embed:
        raise ValueError(f"embed vector of {n_embed} floats is not a multiple of the embedding width {cfg.embed}")
    V = n_embed // cfg.embed

    class _C:
        pass
    c = _C()
    c.__dict__.update({k: getattr(cfg, k) for k in ("embed", "Hq", "nlayer", "C", "S", "M", "A", "H")})
    c.V, c.N = V, 0
    want_rnn = sum(int(np.prod(s)) for s in tensor_shapes(c, "rnn").values())
    if n_rnn != want_rnn:
        raise ValueError(f"rnn vector has {n_rnn} floats, this architecture has {want_rnn}")
    fixed = sum(int(np.prod(s)) for s in tensor_shapes(c, "mult").values())
    rest = n_mult - fixed
    if rest <= 0 or rest % (cfg.M + 1):
        raise ValueError(f"mult vector of {n_mult} floats does not fit C = {cfg.C}: {rest} floats left for the answer head")
    return V, rest // (cfg.M + 1)


def save_snapshot(path, cfg, params, it, epoch, opt=None, cuda=True):
    """What F:1223-1232 writes: params = (embed, rnn, mult) flat vectors in LIBRAU's layout; stored in nngraph's order as
    torch.CudaTensor (training on GPU) or torch.FloatTensor."""
    cls = "torch.CudaTensor" if cuda else "torch.FloatTensor"
    vecs = {i + 1: T7Tensor(cls, to_nngraph(cfg, g, np.asarray(p, dtype=np.float32))) for i, (g, p) in
            enumerate(zip(("embed", "rnn", "mult"), params))}
    t7_save(path, OrderedDict([("it", it), ("opt", dict(opt or {})), ("epoch", epoch), ("params", vecs)]))


def load_snapshot(path, cfg):
    """-> dict(it, epoch, opt, params=[embed, rnn, mult] float32 arrays in LIBRAU's layout, V, N).  `cfg` supplies the
    architecture constants (C in particular: 512 or 2048); V and N are read off the vector lengths."""
    snap = t7_load(path)
    vecs = [np.asarray(snap["params"][i].array, dtype=np.float32).reshape(-1) for i in (1, 2, 3)]
    V, N = infer_sizes(cfg, *(v.size for v in vecs))

    class _C:
        pass
    c = _C()
    c.__dict__.update({k: getattr(cfg, k) for k in ("embed", "Hq", "nlayer", "C", "S", "M", "A", "H")})
    c.V, c.N = V, N
    params = [from_nngraph(c, g, v) for g, v in zip(("embed", "rnn", "mult"), vecs)]
    return dict(it=snap.get("it"), epoch=snap.get("epoch"), opt=snap.get("opt"), params=params, V=V, N=N)
