"""ctypes binding of libkmap_b200.so (the C ABI declared in include/kmap_b200.h).

The signatures are read from the header itself, so the binding cannot drift from the declaration.  There is no
CPU fallback: if the shared library is missing, or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes
import os
import re
from pathlib import Path

_PKG = Path(__file__).resolve().parent
HEADER = _PKG.parent / "include" / "kmap_b200.h"
LIB_PATH = Path(os.environ.get("KMAP_B200_LIB", _PKG / "libkmap_b200.so"))      # (the override is for tuning sweeps)

_CTYPES = {
    "int": ctypes.c_int, "int64_t": ctypes.c_int64, "uint32_t": ctypes.c_uint32, "uint64_t": ctypes.c_uint64,
    "int32_t": ctypes.c_int32, "float": ctypes.c_float, "uint8_t": ctypes.c_uint8,
}


class KmapError(RuntimeError):
    pass


def parse_header(path: Path = HEADER):
    """[(name, restype, [argtypes])] for every function prototype in the header."""
    text = re.sub(r"/\*.*?\*/", "", path.read_text(), flags=re.S)
    text = re.sub(r"^\s*#.*$", "", text, flags=re.M)
    protos = []
    for m in re.finditer(r"([A-Za-z_][\w\s\*]*?)\b(kmap_\w+)\s*\(([^)]*)\)\s*;", text):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        if "*" in ret:
            restype = ctypes.c_char_p
        else:
            restype = _CTYPES[ret.replace("const", "").strip()]
        argtypes = []
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                if "*" in a:
                    argtypes.append(ctypes.c_void_p)
                else:
                    argtypes.append(_CTYPES[a.replace("const", "").split()[0]])
        protos.append((name, restype, argtypes))
    return protos


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise KmapError(f"{LIB_PATH} is missing: build it with `make -C kmap_b200/csrc` "
                            "(or `python -c 'import __graft_entry__ as g; g.build()'`). There is no CPU fallback.")
        handle = ctypes.CDLL(str(LIB_PATH))
        for name, restype, argtypes in parse_header():
            fn = getattr(handle, name)          # AttributeError if the .so does not export a declared symbol
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = handle
    return _lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = lib().kmap_last_error().decode(errors="replace")
        raise KmapError(f"{what or 'libkmap_b200'} failed with code {rc}: {msg}")
