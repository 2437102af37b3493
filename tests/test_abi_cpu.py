"""Host-side checks that need no GPU: librau.so builds/loads, exports every symbol include/rau.h declares, its
layout functions agree with the oracle's parameter inventory, and it refuses to run without a B200."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from conftest import small_cfg
from helpers import lib_cfg
from oracle import rau_oracle as O


@pytest.fixture(scope="module")
def lib():
    from rau_vqa_b200 import _ffi
    return _ffi.load()


def test_library_exports_every_declared_symbol(lib):
    from rau_vqa_b200 import _ffi
    names = _ffi.declared_functions()
    assert len(names) >= 30
    h = ctypes.CDLL(_ffi.LIBNAME)
    missing = [n for n in names if not hasattr(h, n)]
    assert not missing, missing


def test_library_has_no_torch_or_cublas_dependency():
    from rau_vqa_b200 import _ffi
    out = subprocess.run(["ldd", _ffi.LIBNAME], capture_output=True, text=True).stdout
    for bad in ("libtorch", "libc10", "libcublas", "libcudnn"):
        assert bad not in out, out


def test_sass_is_sm100a():
    from rau_vqa_b200 import _ffi
    cuobjdump = "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("no cuobjdump")
    out = subprocess.run([cuobjdump, "-lelf", _ffi.LIBNAME], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out[:500]


@pytest.mark.parametrize("kw", [dict(), dict(C=2048), dict(N=1000, nHop=1)])
def test_group_sizes_match_survey_inventory(lib, kw):
    cfg = O.RauConfig(**kw)
    lc = lib_cfg(cfg)
    for g in O.GROUPS:
        assert lc.group_size(g) == O.group_size(cfg, g)
    if not kw:
        assert lc.group_size("rnn") == 3563520 and lc.group_size("mult") == 5429142
    if kw.get("C") == 2048:
        assert lc.group_size("mult") == 6215574


def test_param_offsets_match_oracle_layout(lib):
    cfg = small_cfg()
    lc = lib_cfg(cfg)
    for g in O.GROUPS:
        off = 0
        for name, shp in O.group_shapes(cfg, g):
            assert lc.param_offset(g, name) == off, (g, name)
            off += int(np.prod(shp))
    assert lc.param_offset("mult", "nope") == -1


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import rau_vqa_b200 as R
    with pytest.raises(R.RauError) as e:
        R.Context(0)
    assert "no CPU fallback" in str(e.value)


def test_header_is_ffi_parseable_without_preprocessor():
    from rau_vqa_b200 import _ffi
    text = _ffi.header_cdef()
    assert "#" not in text and "extern" not in text


def test_batch_struct_layout_and_feed_formats(lib):
    """rau_batch grew a trailing `feats_f16` pointer (fp16 features read directly by the feature pack): the fields every
    existing caller fills keep their offsets, a zero-filled struct means "float32 features only", and the feed formats
    keep their numbers (they cross the ABI as plain ints)."""
    from rau_vqa_b200._ffi import ffi
    b = ffi.new("rau_batch*")
    assert b.feats == ffi.NULL and b.feats_f16 == ffi.NULL and b.B == 0 and b.max_len == 0
    off = {f: ffi.offsetof("rau_batch", f) for f in ("B", "B_global", "feats", "tokens", "lengths", "max_len", "labels", "feats_f16")}
    assert [off[f] for f in ("B", "B_global", "feats", "tokens", "lengths", "max_len", "labels")] == [0, 4, 8, 16, 24, 32, 40]
    assert off["feats_f16"] == 48 and ffi.sizeof("rau_batch") == 56
    assert (lib.RAU_FEED_F32, lib.RAU_FEED_F16, lib.RAU_FEED_F16_DIRECT) == (0, 1, 2)
    from rau_vqa_b200 import feed as F
    assert (F.FEED_F32, F.FEED_F16, F.FEED_F16_DIRECT) == (0, 1, 2)
