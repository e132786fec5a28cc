"""Counter-based synthetic reads (bench / tests).  NumPy twin of csrc/synth.cu: every random decision is
splitmix64(splitmix64(seed + read_id) + slot), so any read range is reproducible on the host and on any GPU.

Shapes follow SURVEY.md section 8d: cfg2 = HT-SELEX-like 1e6 x 40 bp with AATCGATAGC (40 %) and AGGACCTACGTAC (40 %);
cfg3 = ChIP-like 1e8 x 100 bp with GTACGTAGGTCCTA in 10 % of the reads; 5 % per-base mutation of the planted motif
(the reference's own generator, tests/kmap_tests.py:75-114).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)
SLOT_MOTIF = 1 << 20
SLOT_N = 1 << 21


def splitmix64(z: np.ndarray) -> np.ndarray:
    z = (z + np.uint64(0x9E3779B97F4A7C15)).astype(np.uint64)
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def draw(seed: int, read: np.ndarray, slot) -> np.ndarray:
    with np.errstate(over="ignore"):
        return splitmix64(splitmix64(np.uint64(seed) + read.astype(np.uint64)) + np.asarray(slot, dtype=np.uint64))


def unit24(u: np.ndarray) -> np.ndarray:
    return (u >> np.uint64(40)).astype(np.float32) * np.float32(1.0 / 16777216.0)


@dataclass
class SynthSpec:
    seed: int
    read_len: int
    motifs: List[str] = field(default_factory=list)
    motif_frac: List[float] = field(default_factory=list)   # fraction of reads carrying each motif
    mut_rate: float = 0.05
    n_rate: float = 0.0

    def cum_frac(self) -> np.ndarray:
        return np.cumsum(np.asarray(self.motif_frac, dtype=np.float32), dtype=np.float32)


CFG2 = SynthSpec(seed=20240412, read_len=40, motifs=["AATCGATAGC", "AGGACCTACGTAC"], motif_frac=[0.4, 0.4])
CFG2_N = SynthSpec(seed=20240412, read_len=40, motifs=["AATCGATAGC", "AGGACCTACGTAC"], motif_frac=[0.4, 0.4], n_rate=0.001)
CFG3 = SynthSpec(seed=20240413, read_len=100, motifs=["GTACGTAGGTCCTA"], motif_frac=[0.1])


def generate_numpy(spec: SynthSpec, read0: int, n_reads: int) -> Tuple[np.ndarray, np.ndarray]:
    """(seq uint8[n_reads*(L+1)] in the input.bin layout, borders int64[n_reads, 2])"""
    L = spec.read_len
    reads = np.arange(read0, read0 + n_reads, dtype=np.uint64)
    pos = np.arange(L, dtype=np.int64)
    with np.errstate(over="ignore"):
        words = draw(spec.seed, reads[:, None], (pos >> 5)[None, :])
        base = ((words >> (np.uint64(2) * (pos & 31).astype(np.uint64))[None, :]) & np.uint64(3)).astype(np.uint8)
        if spec.motifs:
            u = unit24(draw(spec.seed, reads, SLOT_MOTIF))
            cum = spec.cum_frac()
            which = np.full(n_reads, -1, dtype=np.int64)
            for i in range(len(cum) - 1, -1, -1):
                which[u < cum[i]] = i
            for i, motif in enumerate(spec.motifs):
                M = len(motif)
                rows = np.flatnonzero(which == i)
                if L < M or len(rows) == 0:
                    continue
                start = (draw(spec.seed, reads[rows], SLOT_MOTIF + 1) % np.uint64(L - M + 1)).astype(np.int64)
                codes = np.array(["ACGT".index(c) for c in motif], dtype=np.uint8)
                for j in range(M):
                    keep = unit24(draw(spec.seed, reads[rows], SLOT_MOTIF + 16 + j)) >= np.float32(spec.mut_rate)
                    base[rows[keep], start[keep] + j] = codes[j]
        if spec.n_rate > 0:
            un = unit24(draw(spec.seed, reads[:, None], (SLOT_N + pos)[None, :]))
            base[un < np.float32(spec.n_rate)] = 255
    seq = np.full((n_reads, L + 1), 255, dtype=np.uint8)
    seq[:, :L] = base
    borders = np.zeros((n_reads, 2), dtype=np.int64)
    borders[:, 0] = np.arange(n_reads, dtype=np.int64) * (L + 1)
    borders[:, 1] = borders[:, 0] + L
    return seq.reshape(-1), borders


def generate_device(spec: SynthSpec, read0: int, n_reads: int, want_borders: bool = True):
    """(seq uint8 device tensor, borders int64[n_reads, 2] device tensor or None) produced by the CUDA generator"""
    import torch
    from . import engine as E
    from ._lib import check, lib
    L = spec.read_len
    seq = E.empty(n_reads * (L + 1), torch.uint8)
    borders = E.empty(2 * n_reads, torch.int64) if want_borders else None
    nm = len(spec.motifs)
    motifs = np.zeros(max(nm, 1) * 32, dtype=np.uint8)
    for i, m in enumerate(spec.motifs):
        assert len(m) <= 32
        motifs[32 * i:32 * i + len(m)] = ["ACGT".index(c) for c in m]
    mlen = np.array([len(m) for m in spec.motifs] or [0], dtype=np.int32)
    cum = spec.cum_frac() if nm else np.zeros(1, dtype=np.float32)
    m_d, l_d, c_d = E.to_device(motifs), E.to_device(mlen), E.to_device(cum)
    check(lib().kmap_synth_reads(spec.seed, read0, n_reads, L, m_d.data_ptr(), l_d.data_ptr(), c_d.data_ptr(), nm,
                                 float(spec.mut_rate), float(spec.n_rate), 0, seq.data_ptr(),
                                 None if borders is None else borders.data_ptr(), torch.cuda.current_stream().cuda_stream),
          "kmap_synth_reads")
    return seq, (None if borders is None else borders.view(n_reads, 2))
