"""CPU-side checks: the shared library loads and exports every symbol the header declares; the product never
imports the oracle; argument validation works without a GPU."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol():
    from kmap_b200 import _lib
    protos = _lib.parse_header()
    names = [p[0] for p in protos]
    assert len(names) == len(set(names)) and len(names) >= 35
    handle = _lib.lib()
    for name in names:
        assert hasattr(handle, name), name
    assert handle.kmap_version() >= 100
    assert handle.kmap_valid_words(0) >= 1 and handle.kmap_packed_words(64) == 2 * handle.kmap_valid_words(64)
    assert handle.kmap_compact_scratch_words(8) > 0 and handle.kmap_dedup_work_words(10) == 24


def test_header_cites_reference_for_every_entry_point():
    text = (ROOT / "include" / "kmap_b200.h").read_text()
    assert text.count("kmer_count.py:") + text.count("taichi_core.py:") + text.count("motif_discovery.py:") >= 20


def test_bad_arguments_are_reported_not_crashed():
    from kmap_b200 import _lib
    L = _lib.lib()
    assert L.kmap_kmer2hash_u32(None, 10, 16, None, None) == -1          # k too large for 32-bit hashes
    assert b"k out of range" in L.kmap_last_error()
    assert L.kmap_count_dense(None, None, 10, 16, None, None) == -1
    assert L.kmap_kmer2hash_u32(None, 0, 8, None, None) == 0              # empty input is a no-op


def test_product_never_touches_the_oracle():
    for path in list((ROOT / "kmap_b200").rglob("*.py")) + list((ROOT / "kmap_b200").rglob("*.cu")) + \
            list((ROOT / "kmap_b200").rglob("*.cuh")):
        text = path.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), path
        assert "kmap_oracle" not in text, path


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import numpy as np
    from kmap_b200 import kmer_count as kc
    from kmap_b200._lib import KmapError
    with pytest.raises(KmapError):
        kc.comp_kmer_hash_taichi(np.zeros(10, dtype=np.uint8), 4)
