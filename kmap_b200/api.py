"""Batch entry points over host buffers: the calls a user (or the reference's drivers) makes when the reads live in
NumPy arrays.  They upload once, keep everything on the device across all k, and return the reference's data structures.

Multi-GPU: one process per GPU (torch.distributed, NCCL).  Reads are independent units, so each rank counts a contiguous
range of reads into a private dense table; the only exchange step is an in-place integer all-reduce of the 4^k table
(`TableAllReduce`).  The distance matrix is partitioned by row blocks and needs no collective.
"""
from __future__ import annotations

from typing import Dict, Iterable, Optional, Sequence, Tuple

import numpy as np
import torch

from . import engine as E
from ._lib import KmapError


class TableAllReduce:
    """In-place sum of a dense count table over all ranks.  uint32 counts travel as their int32 bit patterns: two's
    complement addition is the same modular sum, so the merged table is bit-identical for any number of ranks."""

    def __init__(self, group=None):
        import torch.distributed as dist
        if not dist.is_initialized():
            raise KmapError("TableAllReduce needs an initialised torch.distributed process group")
        self.dist, self.group = dist, group

    def __call__(self, table: torch.Tensor) -> torch.Tensor:
        self.dist.all_reduce(table, op=self.dist.ReduceOp.SUM, group=self.group)
        return table

    @property
    def rank(self) -> int:
        return self.dist.get_rank(self.group)


def _pinned_like(t: torch.Tensor) -> torch.Tensor:
    return torch.empty(t.shape, dtype=t.dtype, pin_memory=True)


def _to_host_pinned(t: torch.Tensor, dtype) -> np.ndarray:
    host = _pinned_like(t)
    host.copy_(t, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    a = host.numpy()
    return a.view(dtype) if a.dtype != np.dtype(dtype) else a


def upload_reads(seq_np_arr: np.ndarray, boarder_mat: Optional[np.ndarray], validate: bool = True) -> E.SeqOnDevice:
    """input.bin + input.seqboarder.bin (host) -> packed device form.  Pinned inputs are copied asynchronously."""
    E.require_cuda()
    seq_np_arr = np.asarray(seq_np_arr)
    if seq_np_arr.dtype != np.uint8:
        raise KmapError("seq_np_arr must be uint8")
    borders = None
    if boarder_mat is not None:
        b = np.ascontiguousarray(np.asarray(boarder_mat, dtype=np.int64)).reshape(-1, 2)
        if validate:
            E.check_borders_tile(b, len(seq_np_arr))
        borders = torch.from_numpy(b).to("cuda", non_blocking=True) if len(b) else E.empty(0, torch.int64).view(0, 2)
    seq_d = torch.from_numpy(np.ascontiguousarray(seq_np_arr)).to("cuda", non_blocking=True) if len(seq_np_arr) else \
        E.empty(0, torch.uint8)
    return E.SeqOnDevice.from_device_u8(seq_d, borders)


def count_kmers(seq_np_arr: np.ndarray, boarder_mat: np.ndarray, k_list: Iterable[int], rep_mode: bool = False,
                revcom_mode: bool = True, validate: bool = True, table_allreduce: Optional[TableAllReduce] = None,
                lists_on: Optional[int] = None) -> Dict[int, Tuple[np.ndarray, np.ndarray]]:
    """First-round counts of find_motif for every k (reference motif_discovery.py:627-640): per k the
    `(uniq_kh_arr uint32, uniq_kh_cnt_arr int32)` pair that the reference pickles into kmer_count/k{k}.pkl, in the
    reference's order.  With `table_allreduce` each rank passes ITS shard of the reads and the tables are merged before
    compaction; `lists_on=r` returns the lists on rank r only (others get {})."""
    dev = upload_reads(seq_np_arr, boarder_mat, validate)
    out: Dict[int, Tuple[np.ndarray, np.ndarray]] = {}
    table = None
    for k in k_list:
        n_cells = 1 << (2 * k)
        if table is None or table.numel() < n_cells:
            table = E.empty(n_cells, torch.int32)
        tk = table[:n_cells]
        dev.count(k, dedup=not rep_mode, table=tk, zero=True)
        if table_allreduce is not None:
            table_allreduce(tk)
            if lists_on is not None and table_allreduce.rank != lists_on:
                continue
        kh, cnt = E.compact_merge(tk, k, revcom_mode)
        out[k] = (_to_host_pinned(kh, np.uint32), _to_host_pinned(cnt, np.int32))
    return out


def hamdist_matrix_rows(kh: np.ndarray, labels: np.ndarray, head_len: Sequence[int], kmer_len: int, rank: int, world: int):
    """Row-block partition of the sampled k-mer distance matrix (motif_discovery.py:759-808) for multi-GPU runs:
    returns (row0, row1, uint8 device tensor [(row1-row0), n]) of this rank's slab; no collective is involved."""
    from .motif_discovery import hamdist_matrix_u8
    n = len(kh)
    row0, row1 = n * rank // world, n * (rank + 1) // world
    return row0, row1, hamdist_matrix_u8(kh, labels, head_len, kmer_len, row0, row1)
