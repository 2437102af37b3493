"""Torch7 `.t7` snapshot reader / writer and the nngraph <-> librau layout permutation (F:1223-1232,
EV:114, EV:345-347).  CPU only: host logic."""
import io
import struct

import numpy as np
import pytest

from oracle import rau_oracle as O
from rau_vqa_b200.utils import snapshot as S


def _hand_built_snapshot():
    """A byte string assembled here with struct.pack, independently of T7Writer: the table the scripts save, holding a
    number, a nested option table (string, number, boolean), and params = {CudaTensor, FloatTensor, a second reference to
    the first tensor}."""
    b = io.BytesIO()
    i32 = lambda v: b.write(struct.pack("<i", v))
    i64 = lambda v: b.write(struct.pack("<q", v))
    f64 = lambda v: b.write(struct.pack("<d", v))

    def s(x):
        i32(len(x)); b.write(x.encode())

    def key(x):
        i32(2); s(x)

    i32(3); i32(1); i32(4)                         # TABLE, index 1, 4 entries
    key("it"); i32(1); f64(1234.0)
    key("epoch"); i32(1); f64(2.5)
    key("opt"); i32(3); i32(2); i32(3)             # nested table, index 2
    key("alg_name"); i32(2); s("Ours_Full")
    key("nhop"); i32(1); f64(8.0)
    key("visatt"); i32(5); i32(0)                  # boolean false
    key("params"); i32(3); i32(3); i32(3)          # table index 3 with keys 1, 2, 3 (numbers)
    i32(1); f64(1.0)
    i32(4); i32(4); s("V 1"); s("torch.CudaTensor")
    i32(1); i64(6); i64(1); i64(1)                 # nDim 1, size 6, stride 1, offset 1
    i32(4); i32(5); s("V 1"); s("torch.CudaStorage"); i64(6)
    b.write(np.arange(6, dtype=np.float32).tobytes())
    i32(1); f64(2.0)
    i32(4); i32(6); s("V 1"); s("torch.FloatTensor")
    i32(2); i64(2); i64(2); i64(1); i64(2); i64(3)   # 2 x 2 view, strides (1, 2): a TRANSPOSED view, offset 3 (1-based)
    i32(4); i32(7); s("V 1"); s("torch.FloatStorage"); i64(8)
    b.write((10 + np.arange(8, dtype=np.float32)).tobytes())
    i32(1); f64(3.0)
    i32(4); i32(4)                                 # back reference to object 4 (the CudaTensor)
    return b.getvalue()


def test_reader_on_a_hand_built_stream():
    snap = S.T7Reader(_hand_built_snapshot()).read()
    assert snap["it"] == 1234.0 and snap["epoch"] == 2.5
    assert snap["opt"] == {"alg_name": "Ours_Full", "nhop": 8.0, "visatt": False}
    p = snap["params"]
    assert p[1].cls == "torch.CudaTensor"
    np.testing.assert_array_equal(p[1].array, np.arange(6, dtype=np.float32))
    np.testing.assert_array_equal(p[2].array, np.array([[12, 14], [13, 15]], dtype=np.float32))   # strided view honoured
    assert p[3] is p[1]                            # the shared object is ONE object, as torch.load returns it


def test_writer_reader_round_trip_and_byte_layout():
    w = S.T7Writer()
    obj = {"it": 7, "name": "x", "flag": True, "none": None, "t": np.arange(12, dtype=np.float32).reshape(3, 4),
           "list": [1.5, "a"]}
    w.write(obj)
    data = w.getvalue()
    assert data[:8] == struct.pack("<ii", 3, 1)    # TYPE_TABLE, first object index 1
    back = S.T7Reader(data).read()
    assert back["it"] == 7.0 and back["name"] == "x" and back["flag"] is True and back["none"] is None
    np.testing.assert_array_equal(back["t"].array, obj["t"])
    assert back["list"] == {1: 1.5, 2: "a"}
    with pytest.raises(ValueError):
        S.T7Reader(data[:-3]).read()               # truncated stream


def test_layout_permutation_is_a_bijection_and_moves_exactly_three_blocks():
    cfg = O.RauConfig(V=300, C=2048, N=1000)
    for g in O.GROUPS:
        n = O.group_size(cfg, g)
        perm = S.permutation(cfg, g, "nngraph", "librau")
        assert perm.size == n and np.array_equal(np.sort(perm), np.arange(n))
        inv = S.permutation(cfg, g, "librau", "nngraph")
        assert np.array_equal(perm[inv], np.arange(n))
    assert np.array_equal(S.permutation(cfg, "embed"), np.arange(O.group_size(cfg, "embed")))
    assert np.array_equal(S.permutation(cfg, "rnn"), np.arange(O.group_size(cfg, "rnn")))
    # mult: fill every tensor with its own id in librau's layout (the oracle's views ARE that layout), go to nngraph order
    # and check the documented sequence
    flat = np.zeros(O.group_size(cfg, "mult"), dtype=np.float32)
    names = [n for n, _ in O.mult_param_shapes(cfg)]
    assert names == S.LIBRAU_ORDER["mult"]
    assert dict(O.mult_param_shapes(cfg)) == {k: tuple(v) for k, v in S.tensor_shapes(cfg, "mult").items()}
    for i, (name, v) in enumerate(O.views(cfg, "mult", flat).items()):
        v[...] = i
    ng = S.to_nngraph(cfg, "mult", flat)
    seq = [names[int(v)] for v in ng[np.r_[True, ng[1:] != ng[:-1]]]]
    assert seq == S.NNGRAPH_ORDER["mult"]
    np.testing.assert_array_equal(S.from_nngraph(cfg, "mult", ng), flat)


def test_snapshot_round_trip_through_a_file(tmp_path):
    cfg = O.RauConfig(V=123, C=512, N=77)
    params = O.init_params(cfg, seed=5, dtype=np.float32)
    path = tmp_path / "snapshot_iter000010_epoch0.50.t7"
    S.save_snapshot(str(path), cfg, [params[g] for g in O.GROUPS], it=10, epoch=0.5, opt={"nhop": 8, "alg_name": "Ours_Full"})
    raw = S.t7_load(str(path))
    assert raw["params"][3].cls == "torch.CudaTensor" and raw["params"][3].array.ndim == 1
    # the file holds nngraph's order: its mult vector differs from ours, its Wa block sits where nngraph puts it
    ng = raw["params"][3].array
    assert not np.array_equal(ng, params["mult"])
    off_wa = sum(int(np.prod(S.tensor_shapes(cfg, "mult")[n])) for n in S.NNGRAPH_ORDER["mult"][:6])
    np.testing.assert_array_equal(ng[off_wa:off_wa + cfg.A * cfg.M], O.views(cfg, "mult", params["mult"])["Wa"].reshape(-1))
    blank = O.RauConfig(V=1, C=512, N=1)           # the loader reads V and N off the vector lengths
    snap = S.load_snapshot(str(path), blank)
    assert (snap["V"], snap["N"], snap["it"], snap["epoch"]) == (123, 77, 10.0, 0.5)
    assert snap["opt"]["alg_name"] == "Ours_Full"
    for g, p in zip(O.GROUPS, snap["params"]):
        np.testing.assert_array_equal(p, params[g])
    with pytest.raises(ValueError):
        S.load_snapshot(str(path), O.RauConfig(V=1, C=2048, N=1))      # a VGG snapshot does not fit the ResNet architecture
