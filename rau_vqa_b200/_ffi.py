"""cffi binding of librau.so.

The declarations are not repeated here: include/rau.h is handed to ``ffi.cdef`` after dropping preprocessor
lines, exactly what the LuaJIT host does with ``ffi.cdef`` (lua/rau_ffi.lua).  There is no CPU
fallback: if the library is missing, or no B200 is present when a context is requested, this raises.
"""
from __future__ import annotations

import os
import re
import threading

import cffi

_HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(_HERE, "..", "include", "rau.h")
LIBNAME = os.path.join(_HERE, "librau.so")


def header_cdef(path: str = HEADER) -> str:
    """include/rau.h as FFI-parseable C: no #directives, no extern "C" wrapper."""
    text = open(path).read()
    text = re.sub(r"#ifdef __cplusplus.*?#endif", "", text, flags=re.S)
    text = "\n".join(l for l in text.splitlines() if not l.lstrip().startswith("#"))
    return text


def declared_functions(path: str = HEADER):
    """Names of every entry point the header declares (used by the symbol-export test)."""
    return sorted(set(re.findall(r"\b(rau_[a-z0-9_]+)\s*\(", header_cdef(path))))


ffi = cffi.FFI()
ffi.cdef(header_cdef())

_lib = None
_lock = threading.Lock()


class RauError(RuntimeError):
    def __init__(self, status: int, msg: str):
        super().__init__(f"librau status {status}: {msg}")
        self.status = status


def load(build_if_missing: bool = True):
    """dlopen librau.so (building it with nvcc first when the sources are newer and nvcc exists)."""
    global _lib
    with _lock:
        if _lib is None:
            if build_if_missing:
                from . import build as _build
                try:
                    _build.build()
                except Exception:
                    if not os.path.exists(LIBNAME):
                        raise
            if not os.path.exists(LIBNAME):
                raise RuntimeError(f"{LIBNAME} is missing: build it with `python -m rau_vqa_b200.build` (no CPU fallback)")
            _lib = ffi.dlopen(LIBNAME)
    return _lib


def check(status: int):
    if status != 0:
        raise RauError(status, ffi.string(load().rau_last_error()).decode())
    return status
