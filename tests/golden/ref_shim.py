"""Import shim that lets the UNMODIFIED reference (`/root/reference/src/kmap`) run in a container
that has neither taichi nor biopython / matplotlib / logomaker.

Used ONLY by `tests/golden/make_golden.py` (run once, here, where /root/reference exists) to produce the
committed golden fixtures. Nothing in the product, the tests or the bench imports this at run time.

What is substituted, and why it does not change integer results:

* `taichi`  -> a tiny interpreter: `@ti.kernel` / `@ti.func` run the decorated function body as ordinary
  Python, `ti.u8/u32/u64` are the numpy scalar types (fixed-width, wrapping arithmetic like Taichi's),
  `ti.cast(x, t)` is `t(x)`.  The reference kernels only use `<<`, `>>`, `+`, `-`, `^`, `&`, `!=` on those
  types, so the interpreted result is the same machine arithmetic.  Out-of-bounds *reads* (the reference's
  `kmer2hash` reads `arr[st_pos+i]` past the end before overwriting the result with `invalid_hash`,
  taichi_core.py:14-22) return 0 instead of raising; the value read is never observable.
* `Bio.SeqIO.parse(handle_or_path, "fasta")` -> a minimal FASTA reader (header lines start with '>',
  sequence = concatenation of the following lines with whitespace stripped), which is what biopython
  yields for plain FASTA.
* `matplotlib`, `logomaker`, `Bio.Align` -> inert mocks (plotting / alignment are outside the hot path).
"""
from __future__ import annotations

import sys
import types
from unittest import mock

import numpy as np

REF_SRC = "/root/reference/src"
REF_TESTS = "/root/reference"


class _OOBSafe:
    """ndarray proxy: integer reads past the end give 0 (see module docstring)."""

    __slots__ = ("a",)

    def __init__(self, a):
        self.a = a

    def __getitem__(self, idx):
        try:
            return self.a[idx]
        except IndexError:
            return self.a.dtype.type(0)

    def __setitem__(self, idx, val):
        self.a[idx] = val

    def __len__(self):
        return len(self.a)


def _make_taichi():
    ti = types.ModuleType("taichi")

    def func(f):
        return f

    def kernel(f):
        def run(*args, **kwargs):
            args = [(_OOBSafe(a) if isinstance(a, np.ndarray) else a) for a in args]
            with np.errstate(over="ignore"):
                return f(*args, **kwargs)
        run.__wrapped__ = f
        return run

    ti.func = func
    ti.kernel = kernel
    for name, t in dict(u8=np.uint8, u16=np.uint16, u32=np.uint32, u64=np.uint64, i8=np.int8, i16=np.int16,
                        i32=np.int32, i64=np.int64, f32=np.float32, f64=np.float64).items():
        setattr(ti, name, t)
    tys = types.SimpleNamespace(ndarray=lambda *a, **k: object, u8=np.uint8, u32=np.uint32, u64=np.uint64,
                                f32=np.float32, i32=np.int32)
    ti.types = tys
    ti.cast = lambda x, t: t(x)
    ti.log = np.log
    ti.sqrt = np.sqrt
    ti.init = lambda *a, **k: None
    ti.set_logging_level = lambda *a, **k: None
    ti.ERROR = "error"
    ti.cpu = "cpu"
    ti.cuda = "cuda"
    ti.cfg = types.SimpleNamespace(arch="cpu")
    ti.field = lambda *a, **k: (_ for _ in ()).throw(NotImplementedError("ti.field not shimmed"))
    alg = types.ModuleType("taichi.algorithms")
    alg.parallel_sort = lambda *a, **k: (_ for _ in ()).throw(NotImplementedError)
    ti.algorithms = alg
    return ti, alg


class _Rec:
    def __init__(self, name, seq):
        self.id = self.name = name
        self.description = name
        self.seq = seq


def _fasta_parse(handle, fmt="fasta"):
    assert fmt == "fasta"
    close = False
    if isinstance(handle, (str, bytes)) or hasattr(handle, "__fspath__"):
        handle = open(handle, "r")
        close = True
    try:
        name, chunks = None, []
        for line in handle:
            line = line.rstrip("\r\n")
            if line.startswith(">"):
                if name is not None:
                    yield _Rec(name, "".join(chunks))
                name, chunks = (line[1:].split() or [""])[0], []
            elif name is not None:
                chunks.append("".join(line.split()))
        if name is not None:
            yield _Rec(name, "".join(chunks))
    finally:
        if close:
            handle.close()


def install():
    """Install the stand-in modules and return the imported reference modules (kc, md, tc)."""
    ti, alg = _make_taichi()
    sys.modules["taichi"] = ti
    sys.modules["taichi.algorithms"] = alg

    bio = types.ModuleType("Bio")
    seqio = types.ModuleType("Bio.SeqIO")
    seqio.parse = _fasta_parse
    bio.SeqIO = seqio
    sys.modules["Bio"] = bio
    sys.modules["Bio.SeqIO"] = seqio
    for m in ("Bio.Align", "Bio.Seq", "Bio.SeqRecord", "matplotlib", "matplotlib.pyplot", "matplotlib.colors",
              "matplotlib.cm", "logomaker"):
        sys.modules[m] = mock.MagicMock(name=m)
    bio.Align = sys.modules["Bio.Align"]
    bio.Seq = sys.modules["Bio.Seq"]
    bio.SeqRecord = sys.modules["Bio.SeqRecord"]

    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)
    if REF_TESTS not in sys.path:
        sys.path.insert(0, REF_TESTS)
    import kmap.kmer_count as kc
    import kmap.motif_discovery as md
    import kmap.taichi_core as tc
    # plotting / alignment are outside the hot path: make them no-ops
    md._align_conseq = lambda *a, **k: None
    md._draw_logo = lambda *a, **k: None
    return kc, md, tc
