"""Device-side engine: torch tensors for the buffers, libkmap_b200 for every kernel.

`SeqOnDevice` is the packed, device-resident form of `input.bin.pkl` + `input.seqboarder.bin.pkl`
(reference kmer_count.py:326-347) and carries the operations of the scan_motif counting path:
count (with or without per-read de-duplication), order-exact compaction to the reference's
`(uniq_kh_arr, uniq_kh_cnt_arr)`, Hamming-ball sums, masking, occurrence scan.

PyTorch is plumbing only (allocation, H2D/D2H, streams, torch.distributed); no torch op computes a result.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from ._lib import KmapError, check, lib

MISSING_VAL = 255


def require_cuda() -> torch.device:
    if not torch.cuda.is_available():
        raise KmapError("kmap_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def to_device(a: np.ndarray, dtype=None) -> torch.Tensor:
    """host ndarray -> device tensor with the same bytes (uint32/uint64 travel as int32/int64 views)."""
    dev = require_cuda()
    a = np.ascontiguousarray(a if dtype is None else a.astype(dtype, copy=False))
    view = {np.dtype(np.uint32): np.int32, np.dtype(np.uint64): np.int64, np.dtype(np.uint16): np.int16}.get(a.dtype)
    src = a.view(view) if view is not None else a
    if src.size == 0:
        return torch.empty(0, dtype=torch.from_numpy(np.empty(0, src.dtype)).dtype, device=dev)
    return torch.from_numpy(src).to(dev, non_blocking=False)


PINNED_D2H_MIN_BYTES = 1 << 20


def to_host(t: torch.Tensor, dtype=None) -> np.ndarray:
    """device tensor -> host ndarray [reinterpreted as `dtype` (same item size)].  Results of a megabyte or more land in
    pinned memory (torch's caching host allocator keeps the blocks between calls): a pageable copy runs at a few GB/s,
    a pinned one at link speed -- the per-read results of the occurrence scan are ~1 GB per consensus at 1e8 reads."""
    if t.is_cuda and t.numel() * t.element_size() >= PINNED_D2H_MIN_BYTES:
        try:
            host = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        except RuntimeError:
            host = None
        if host is not None:
            host.copy_(t, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            a = host.numpy()
        else:
            a = t.cpu().numpy()
    else:
        a = t.cpu().numpy()
    return a.view(dtype) if dtype is not None and a.dtype != np.dtype(dtype) else a


def empty(n: int, dtype: torch.dtype) -> torch.Tensor:
    return torch.empty(max(int(n), 0), dtype=dtype, device=require_cuda())


def zeros(n: int, dtype: torch.dtype) -> torch.Tensor:
    return torch.zeros(max(int(n), 0), dtype=dtype, device=require_cuda())


def alloc_tables(kmin: int, kmax: int, zero: bool = False):
    """(flat, {k: view}) -- the dense tables of levels kmin..kmax as views of ONE buffer (largest first), so that the
    multi-GPU merge is a single all-reduce call instead of one per level."""
    total = sum(1 << (2 * k) for k in range(kmin, kmax + 1))
    flat = zeros(total, torch.int32) if zero else empty(total, torch.int32)
    views, o = {}, 0
    for k in range(kmax, kmin - 1, -1):
        views[k] = flat[o:o + (1 << (2 * k))]
        o += 1 << (2 * k)
    return flat, views


def check_borders_tile(borders: np.ndarray, n: int):
    """The fused de-duplicating count walks reads, so every non-separator position must belong to exactly one read:
    borders must be the layout preproc writes (kmer_count.py:335-343): st_0 = 0, en_i + 1 = st_{i+1}, en_last = n-1."""
    b = np.asarray(borders)
    if b.size == 0:
        if n != 0:
            raise KmapError("empty border matrix for a non-empty sequence array")
        return
    if b.ndim != 2 or b.shape[1] != 2:
        raise KmapError("boarder_mat must be n_seq x 2")
    ok = b[0, 0] == 0 and b[-1, 1] == n - 1 and np.all(b[:, 1] >= b[:, 0]) and np.all(b[1:, 0] == b[:-1, 1] + 1)
    if not ok:
        raise KmapError("boarder_mat does not tile the sequence array as `kmap preproc` writes it "
                        "([start, separator index] per read, reads back to back)")


class SeqOnDevice:
    """Packed reads resident in HBM.  packed: 2 bits/base, valid: 1 bit/base, borders: int64[n_seq, 2]."""

    def __init__(self, n: int, packed: torch.Tensor, valid: torch.Tensor, borders: Optional[torch.Tensor], n_seq: int,
                 seq_u8: Optional[torch.Tensor] = None):
        self.n, self.packed, self.valid, self.borders, self.n_seq, self.seq_u8 = n, packed, valid, borders, n_seq, seq_u8
        self._flag_scratch = None
        self._work = None
        self._valid0 = None
        self._part = None

    # level-k counts of tables beyond L2 go through key partitioning (csrc/partition.cu); below this k the table is
    # L2 resident and direct atomics win (measured)
    PARTITION_MIN_K = 13
    # de-duplicated level-k counts from this k on go through the per-read scan + hidden-window count of the all-k path
    # (one filter atomic per window in shared memory) instead of one hash-set insertion + global RED per window
    DEDUP_SCAN_MIN_K = 12

    # schemes for a level-k table beyond L2 (include/kmap_b200.h)
    PREFIX_PASSES, SORTED = 0, 1

    def _part_scratch(self, k: int, scheme: int = 1) -> torch.Tensor:
        need = lib().kmap_partition_scratch_bytes(self.n, k)
        if self._part is None or self._part.numel() < need:
            self._part = None                      # release before growing
            self._part = empty(need, torch.uint8)
        return self._part

    # ---- construction ---------------------------------------------------------------------------------------
    @classmethod
    def from_device_u8(cls, seq_u8: torch.Tensor, borders: Optional[torch.Tensor], keep_u8: bool = False,
                       capacity: Optional[int] = None) -> "SeqOnDevice":
        """capacity: positions to size the packed buffers for (>= len(seq_u8)) when the object will be `rebind`-ed."""
        L = lib()
        n = int(seq_u8.numel())
        cap = max(n, int(capacity or 0))
        packed = empty(L.kmap_packed_words(cap), torch.int32)
        valid = empty(L.kmap_valid_words(cap), torch.int32)
        check(L.kmap_pack2bit(_ptr(seq_u8), n, _ptr(packed), _ptr(valid), _stream_ptr()), "kmap_pack2bit")
        n_seq = 0 if borders is None else int(borders.shape[0])
        return cls(n, packed, valid, borders, n_seq, seq_u8 if keep_u8 else None)

    def rebind(self, seq_u8: torch.Tensor, borders: Optional[torch.Tensor]):
        """re-use this object (packed buffers and every scratch buffer) for another chunk of reads of at most the same size"""
        L = lib()
        n = int(seq_u8.numel())
        if L.kmap_valid_words(n) > self.valid.numel():
            raise KmapError("rebind: the chunk is larger than the buffers of this SeqOnDevice")
        check(L.kmap_pack2bit(_ptr(seq_u8), n, _ptr(self.packed), _ptr(self.valid), _stream_ptr()), "kmap_pack2bit")
        self.n, self.borders, self.n_seq = n, borders, (0 if borders is None else int(borders.shape[0]))
        self.seq_u8, self._valid0 = None, None

    def rebind_packed(self, n: int, packed: torch.Tensor, valid: torch.Tensor, borders: Optional[torch.Tensor]):
        """re-use this object (every scratch buffer) for another chunk of at most the same size that is ALREADY in the packed
        form (uploaded packed, or packed into these buffers by the caller)"""
        L = lib()
        if valid.numel() < L.kmap_valid_words(n) or packed.numel() < L.kmap_packed_words(n):
            raise KmapError("rebind_packed: buffers too small for the chunk")
        if self._dupmask_words() and valid.numel() > self._dupmask_words():
            self._dupmask = None
        self.n, self.packed, self.valid, self.borders = n, packed, valid, borders
        self.n_seq = 0 if borders is None else int(borders.shape[0])
        self.seq_u8, self._valid0 = None, None

    def _dupmask_words(self) -> int:
        d = getattr(self, "_dupmask", None)
        return 0 if d is None else int(d.numel())

    @classmethod
    def from_numpy(cls, seq_np_arr: np.ndarray, boarder_mat: Optional[np.ndarray] = None, keep_u8: bool = False,
                   validate: bool = True) -> "SeqOnDevice":
        seq_np_arr = np.asarray(seq_np_arr)
        if seq_np_arr.dtype != np.uint8:
            raise KmapError("seq_np_arr must be uint8 (A0 C1 G2 T3, 255 missing)")
        borders = None
        if boarder_mat is not None:
            b = np.ascontiguousarray(np.asarray(boarder_mat, dtype=np.int64))
            if validate:
                check_borders_tile(b, len(seq_np_arr))
            borders = to_device(b.reshape(-1, 2))
        return cls.from_device_u8(to_device(seq_np_arr), borders, keep_u8)

    @classmethod
    def from_fasta(cls, fasta_file, keep_u8: bool = False, rank: int = 0, world: int = 1) -> "SeqOnDevice":
        """straight from the FASTA file: parsed, encoded and packed on the device, no host copy of input.bin.
        world > 1: only the contiguous read range of `rank` is kept (borders rebased to it); `all_borders` then holds the
        border matrix of the whole file."""
        seq, borders = fasta_to_device(fasta_file)
        if world <= 1:
            return cls.from_device_u8(seq, borders, keep_u8)
        n_seq = int(borders.shape[0])
        r0, r1 = n_seq * rank // world, n_seq * (rank + 1) // world
        if r1 > r0:
            ends = borders[[r0, r1 - 1]].cpu()
            p0, p1 = int(ends[0, 0]), int(ends[1, 1]) + 1
            shard = cls.from_device_u8(seq[p0:p1].clone(), borders[r0:r1] - p0, keep_u8)
        else:
            shard = cls.from_device_u8(seq[:0].clone(), borders[:0].clone(), keep_u8)
        shard.all_borders = borders
        return shard

    # ---- masking state ----------------------------------------------------------------------------------------
    def snapshot_valid(self):
        self._valid0 = self.valid.clone()

    def restore_valid(self):
        self.valid.copy_(self._valid0)

    # ---- counting ---------------------------------------------------------------------------------------------
    def count(self, k: int, dedup: bool, table: Optional[torch.Tensor] = None, zero: bool = True,
              partitioned: Optional[bool] = None, scheme: Optional[int] = None) -> torch.Tensor:
        """dense forward table uint32[4^k] (held as int32 bits).  dedup=True fuses remove_duplicate_hash_per_seq.
        partitioned: None = choose by k (PARTITION_MIN_K / DEDUP_SCAN_MIN_K), True = force the partitioned count (plain counts
        with 9 <= k <= 14), False = force the direct kernels (one global atomic per counted window)."""
        L = lib()
        if not 1 <= k <= 15:
            raise KmapError(f"dense counting supports 1 <= k <= 15 (got {k}); k >= 16 uses 64-bit hashes: count_sorted")
        n_cells = 1 << (2 * k)
        via_scan = dedup and zero and partitioned is not False and k >= self.DEDUP_SCAN_MIN_K
        if table is None:
            table = empty(n_cells, torch.int32) if via_scan else zeros(n_cells, torch.int32)
        elif zero and not via_scan:
            check(L.kmap_fill_u32(_ptr(table), n_cells, 0, _stream_ptr()), "kmap_fill_u32")
        if via_scan:
            # per-read scan (dedup_scan_kernel) + level-k count with the repeats hidden: the kernels of the all-k count
            # (csrc/count_all.cu, csrc/partition.cu) with kmin = kmax = k; count_all zeroes the table itself
            self.count_all(k, k, True, tables={k: table})
        elif dedup:
            if self.borders is None:
                raise KmapError("per-read de-duplication needs the border matrix")
            need = L.kmap_dedup_work_words(self.n_seq)
            if self._work is None or self._work.numel() < need:
                self._work = empty(need, torch.int32)
            # when accumulating into a caller's table a failed first attempt could not be undone: give the bitmap upfront
            bitmap = None if zero else zeros(max(n_cells // 32, 1), torch.int32)
            rc = L.kmap_count_dense_dedup(_ptr(self.packed), _ptr(self.valid), self.n, _ptr(self.borders), self.n_seq, k,
                                          _ptr(table), _ptr(self._work), _ptr(bitmap), _stream_ptr())
            if rc == -3:  # KMAP_ERR_NEED_SCRATCH: a very long read; retry the long reads with the bitmap
                check(L.kmap_fill_u32(_ptr(table), n_cells, 0, _stream_ptr()), "kmap_fill_u32")
                bitmap = zeros(max(n_cells // 32, 1), torch.int32)
                rc = L.kmap_count_dense_dedup(_ptr(self.packed), _ptr(self.valid), self.n, _ptr(self.borders), self.n_seq,
                                              k, _ptr(table), _ptr(self._work), _ptr(bitmap), _stream_ptr())
            check(rc, "kmap_count_dense_dedup")
        elif (partitioned if partitioned is not None else k >= self.PARTITION_MIN_K) and 9 <= k <= 14 and zero:
            scratch = self._part_scratch(k)
            check(L.kmap_count_dense_partitioned(_ptr(self.packed), _ptr(self.valid), self.n, k, _ptr(table), _ptr(scratch),
                                                 scratch.numel(), _stream_ptr()), "kmap_count_dense_partitioned")
        else:
            check(L.kmap_count_dense(_ptr(self.packed), _ptr(self.valid), self.n, k, _ptr(table), _stream_ptr()),
                  "kmap_count_dense")
        return table

    def count_all(self, kmin: int, kmax: int, dedup: bool, tables: Optional[dict] = None, n_partitions: int = 0,
                  phase_events: Optional[Sequence[torch.cuda.Event]] = None, partitioned: Optional[bool] = None,
                  scheme: Optional[int] = None, merge=None) -> dict:
        """Dense forward tables for every k in [kmin, kmax] from ONE update per window at level kmax (csrc/count_all.cu);
        identical to {k: self.count(k, dedup)}.  Returns {k: int32-bit-pattern tensor of 4^k cells}.
        scheme: how the level-kmax table is built when 12 <= kmax <= 14 -- SORTED (default: measured fastest)
        or PREFIX_PASSES (global atomics in n_partitions key-prefix passes; also what `partitioned=False` /
        n_partitions > 0 select).
        merge (api.TableAllReduce): this object holds ONE RANK's shard of the reads; the tables returned are those of the
        whole input, all-reduced over the ranks from inside the count as they become final (kmap_count_all_k_sharded)."""
        L = lib()
        if not (1 <= kmin <= kmax <= 15):
            raise KmapError("count_all needs 1 <= kmin <= kmax <= 15")
        if tables is None:
            tables = alloc_tables(kmin, kmax)[1]
        for k in range(kmin, kmax + 1):
            if k not in tables:
                tables[k] = empty(1 << (2 * k), torch.int32)
        ptrs = (ctypes.c_void_p * (kmax - kmin + 1))(*[tables[k].data_ptr() for k in range(kmin, kmax + 1)])
        dupmask = work = bitmap = None
        if dedup:
            if self.borders is None:
                raise KmapError("per-read de-duplication needs the border matrix")
            if getattr(self, "_dupmask", None) is None:
                self._dupmask = empty(self.valid.numel(), torch.int32)
            need = L.kmap_dedup_work_words(self.n_seq)
            if self._work is None or self._work.numel() < need:
                self._work = empty(need, torch.int32)
            dupmask, work = self._dupmask, self._work
        part, part_bytes = None, 0
        if scheme is None:
            scheme = self.SORTED if (partitioned if partitioned is not None else n_partitions <= 0) else self.PREFIX_PASSES
        if not 12 <= kmax <= 14:
            scheme = self.PREFIX_PASSES
        if scheme != self.PREFIX_PASSES:
            part = self._part_scratch(kmax, scheme)
            part_bytes = part.numel()
        ev = None
        if phase_events is not None:                    # 6 torch events (already recorded once so that the handles exist)
            if len(phase_events) != 6:
                raise KmapError("phase_events must hold 6 events")
            ev = (ctypes.c_void_p * 6)(*[e.cuda_event for e in phase_events])
        def call(bm):
            args = (_ptr(self.packed), _ptr(self.valid), self.n, _ptr(self.borders), self.n_seq, kmin, kmax, int(dedup), ptrs,
                    _ptr(dupmask), _ptr(work), _ptr(bm), int(n_partitions), int(scheme), _ptr(part), part_bytes, ev, _stream_ptr())
            if merge is None:
                return L.kmap_count_all_k(*args)
            comm, comm_stream = merge.native()
            if getattr(merge, "scatter", False):
                return L.kmap_count_all_k_scattered(*args, comm, comm_stream.cuda_stream, merge.rank, merge.world)
            return L.kmap_count_all_k_sharded(*args, comm, comm_stream.cuda_stream)
        if merge is not None and dedup:    # (a retry on one rank would leave the others waiting in a collective: give the bitmap upfront)
            bitmap = zeros(max((1 << (2 * kmax)) // 32, 1), torch.int32)
        rc = call(bitmap)
        if rc == -3 and merge is None:   # a read beyond the block path: rerun with the bitmap scratch (tables are re-zeroed by the call)
            bitmap = zeros(max((1 << (2 * kmax)) // 32, 1), torch.int32)
            rc = call(bitmap)
        check(rc, "kmap_count_all_k")
        return tables

    def count_sorted(self, k: int, dedup: bool) -> Tuple[torch.Tensor, torch.Tensor]:
        """(kh, cnt) = ascending distinct k-mer hashes (uint64 bits in int64) and their int64 counts by the sort /
        run-length path (csrc/sorted.cu): what count_uniq_hash returns (kmer_count.py:476-491), for any 1 <= k <= 31.
        dedup=True applies remove_duplicate_hash_per_seq (kmer_count.py:743-760) to the keys first."""
        L = lib()
        if not 1 <= k <= 31:
            raise KmapError(f"k-mer hashes are 64-bit: 1 <= k <= 31 (got {k})")
        keys = empty(self.n, torch.int64)
        check(L.kmap_window_keys_u64(_ptr(self.packed), _ptr(self.valid), self.n, k, _ptr(keys), _stream_ptr()), "kmap_window_keys_u64")
        if dedup:
            if self.borders is None:
                raise KmapError("per-read de-duplication needs the border matrix")
            work = empty(L.kmap_dedup_keys_work_words(self.n_seq), torch.int32)
            check(L.kmap_dedup_hash_per_read_u64(_ptr(keys), self.n, _ptr(self.borders), self.n_seq, _ptr(work), _stream_ptr()),
                  "kmap_dedup_hash_per_read_u64")
        return sort_count_keys(keys, 2 * k)

    # ---- masking ----------------------------------------------------------------------------------------------
    def mask(self, k: int, consensus_kh: Sequence[int], max_dist: Sequence[int]):
        L = lib()
        m = len(consensus_kh)
        if m == 0 or self.n == 0:
            return
        if not 1 <= k <= 31:
            raise KmapError(f"mask_input: 1 <= k <= 31 (got {k})")
        wide = k > 16                               # 64-bit hashes: the plain kernel (csrc/mask.cu, kmap_mask_u64)
        if wide:
            cons = to_device(np.asarray([int(c) for c in consensus_kh], dtype=np.uint64))
        else:
            cons = to_device(np.asarray([int(c) & 0xFFFFFFFF for c in consensus_kh], dtype=np.uint32))
        d = to_device(np.asarray([int(x) for x in max_dist], dtype=np.int32))
        if self._flag_scratch is None:
            self._flag_scratch = empty(self.valid.numel(), torch.int32)
        pre = self.valid if m <= 16 else self.valid.clone()   # every consensus is compared on the pre-mask windows
        fn = L.kmap_mask_u64 if wide else L.kmap_mask
        for i in range(0, m, 16):
            mm = min(16, m - i)
            check(fn(_ptr(self.packed), _ptr(pre), _ptr(self.valid), self.n, k, cons[i:].data_ptr(),
                     d[i:].data_ptr(), mm, _ptr(self._flag_scratch), _stream_ptr()), "kmap_mask")

    def masked_seq_to_numpy(self, out: np.ndarray):
        """write the masking state into `out` (the caller's seq_np_arr): 255 wherever a base is no longer valid."""
        L = lib()
        if self.n == 0:
            return out
        if self.seq_u8 is None:
            self.seq_u8 = to_device(out)
        check(L.kmap_apply_valid_to_seq(_ptr(self.seq_u8), self.n, _ptr(self.valid), _stream_ptr()), "kmap_apply_valid_to_seq")
        out[:] = self.seq_u8.cpu().numpy()
        return out


# ---- table -> reference lists ----------------------------------------------------------------------------------
_scratch_cache = {}


def _scratch(words: int) -> torch.Tensor:
    dev = torch.cuda.current_device()
    t = _scratch_cache.get(dev)
    if t is None or t.numel() < words:
        t = empty(words, torch.int64)
        _scratch_cache[dev] = t
    return t


def key_range(k: int, rank: int, world: int) -> Tuple[int, int]:
    """[cell_lo, cell_hi) of the level-k table that `rank` of `world` compacts: contiguous, aligned to the 2048-cell tiles of
    the compaction kernels, tiling [0, 4^k)"""
    n_cells = 1 << (2 * k)
    n_tiles = (n_cells + 2047) // 2048
    lo, hi = n_tiles * rank // world * 2048, n_tiles * (rank + 1) // world * 2048
    return min(lo, n_cells), min(hi, n_cells)


def compact_merge(table: torch.Tensor, k: int, revcom: bool, upper_bound: Optional[int] = None,
                  cell_range: Optional[Tuple[int, int]] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """(kh uint32-bits, cnt int32) device tensors in the reference's order (kmer_count.py:476-491, 643-685).
    With `upper_bound` (a safe capacity) the size query is skipped: one counting pass + one writing pass.
    cell_range = (lo, hi): only the entries whose forward hash lies in [lo, hi) (`key_range`): the slices of consecutive
    ranges concatenate to the whole list."""
    L = lib()
    scratch = _scratch(L.kmap_compact_scratch_words(k))
    n_out = ctypes.c_int64(0)
    lo, hi = cell_range if cell_range is not None else (0, 1 << (2 * k))
    if upper_bound is not None:
        upper_bound = min(int(upper_bound), hi - lo)
    if upper_bound is not None and upper_bound <= (1 << 29):
        kh, cnt = empty(upper_bound, torch.int32), empty(upper_bound, torch.int32)
        if upper_bound > 0:
            check(L.kmap_compact_merge_range(_ptr(table), k, int(revcom), lo, hi, _ptr(scratch), _ptr(kh), _ptr(cnt), upper_bound,
                                             ctypes.byref(n_out), _stream_ptr()), "kmap_compact_merge")
        return kh[:n_out.value], cnt[:n_out.value]
    check(L.kmap_compact_merge_range(_ptr(table), k, int(revcom), lo, hi, _ptr(scratch), None, None, 0, ctypes.byref(n_out), _stream_ptr()),
          "kmap_compact_merge(size)")
    n = n_out.value
    kh, cnt = empty(n, torch.int32), empty(n, torch.int32)
    if n:
        check(L.kmap_compact_merge_range(_ptr(table), k, int(revcom), lo, hi, _ptr(scratch), _ptr(kh), _ptr(cnt), n, ctypes.byref(n_out),
                                         _stream_ptr()), "kmap_compact_merge")
    return kh, cnt


def hamball_sums(table: torch.Tensor, k: int, cand: Sequence[int], d: int, revcom: bool) -> np.ndarray:
    """int64 ball sums for each candidate hash by neighbour enumeration on the dense table (motif_discovery.py:666-673)."""
    L = lib()
    m = len(cand)
    if m == 0:
        return np.zeros(0, dtype=np.int64)
    c = to_device(np.asarray([int(x) for x in cand], dtype=np.uint32))
    sums = empty(m, torch.int64)
    check(L.kmap_hamball_sum(_ptr(table), k, _ptr(c), m, int(d), int(revcom), _ptr(sums), _stream_ptr()), "kmap_hamball_sum")
    return sums.cpu().numpy()


def hamball_sums_list(kh: torch.Tensor, cnt: torch.Tensor, k: int, cand: Sequence[int], d: int, revcom: bool) -> np.ndarray:
    L = lib()
    out = np.zeros(len(cand), dtype=np.int64)
    for i in range(0, len(cand), 16):
        part = list(cand[i:i + 16])
        c = to_device(np.asarray([int(x) for x in part], dtype=np.uint32))
        sums = empty(len(part), torch.int64)
        check(L.kmap_hamball_sum_list(_ptr(kh), _ptr(cnt), int(kh.numel()), k, _ptr(c), len(part), int(d), int(revcom),
                                      _ptr(sums), _stream_ptr()), "kmap_hamball_sum_list")
        out[i:i + 16] = sums.cpu().numpy()
    return out


def topk_candidates(cnt: torch.Tensor, kk: int) -> Tuple[np.ndarray, np.ndarray]:
    """(values, indices) of the kk largest counts of a device list (int32 or int64), ordered by (value descending, index
    ascending); fewer than kk when the list is shorter.  Blocks of kmap_topk_candidates_* merged on the host."""
    L = lib()
    n = int(cnt.numel())
    n_blocks = max(1, min(148 * 4, (n + 255) // 256))
    wide = cnt.dtype == torch.int64
    out_val = empty(n_blocks * kk, torch.int64 if wide else torch.int32)
    out_idx = empty(n_blocks * kk, torch.int64)
    fn = L.kmap_topk_candidates_i64 if wide else L.kmap_topk_candidates_i32
    check(fn(_ptr(cnt), n, int(kk), _ptr(out_val), _ptr(out_idx), n_blocks, _stream_ptr()), "kmap_topk_candidates")
    val, idx = out_val.cpu().numpy(), out_idx.cpu().numpy()
    keep = idx >= 0
    val, idx = val[keep], idx[keep]
    order = np.lexsort((idx, -val.astype(np.int64)))[:kk]
    return val[order], idx[order]


def exclusive_scan_u32(counts: torch.Tensor) -> torch.Tensor:
    L = lib()
    n = int(counts.numel())
    out = empty(n + 1, torch.int64)
    scratch = _scratch(L.kmap_list_scratch_words(n))
    check(L.kmap_exclusive_scan_u32(_ptr(counts), n, _ptr(out), _ptr(scratch), _stream_ptr()), "kmap_exclusive_scan_u32")
    return out


def occurrence_scan_device(seq: SeqOnDevice, k: int, conseq_kh: int, d: int, revcom: bool):
    """per read: (min_dist uint8[n_seq] (255 = none), offsets int64[n_seq+1], positions int32[total]) as DEVICE tensors"""
    L = lib()
    n_seq = seq.n_seq
    min_dist = empty(n_seq, torch.uint8)
    n_hit = empty(n_seq, torch.int32)
    f_count, f_fill = (L.kmap_occurrence_count_u64, L.kmap_occurrence_fill_u64) if k > 16 else \
                      (L.kmap_occurrence_count, L.kmap_occurrence_fill)
    check(f_count(_ptr(seq.packed), _ptr(seq.valid), _ptr(seq.borders), n_seq, k, int(conseq_kh), int(d),
                  int(revcom), _ptr(min_dist), _ptr(n_hit), _stream_ptr()), "kmap_occurrence_count")
    offsets = exclusive_scan_u32(n_hit)
    total = int(offsets[-1].item()) if n_seq else 0
    pos = empty(total, torch.int32)
    if total:
        check(f_fill(_ptr(seq.packed), _ptr(seq.valid), _ptr(seq.borders), n_seq, k, int(conseq_kh), int(d),
                     int(revcom), _ptr(min_dist), _ptr(offsets), _ptr(pos), _stream_ptr()), "kmap_occurrence_fill")
    return min_dist, offsets, pos


def occurrence_scan(seq: SeqOnDevice, k: int, conseq_kh: int, d: int, revcom: bool):
    """per read: (min_dist uint8[n_seq] (255 = none), offsets int64[n_seq+1], positions int32[total]) on the host."""
    min_dist, offsets, pos = occurrence_scan_device(seq, k, conseq_kh, d, revcom)
    return to_host(min_dist), to_host(offsets), to_host(pos)


def sum_counts(cnt: torch.Tensor) -> int:
    """exact total of a device count list (int32 or int64): the `sum()` of find_motif (motif_discovery.py:648)"""
    L = lib()
    out = empty(1, torch.int64)
    fn = L.kmap_sum_counts_i64 if cnt.dtype == torch.int64 else L.kmap_sum_counts_i32
    check(fn(_ptr(cnt), int(cnt.numel()), _ptr(out), _stream_ptr()), "kmap_sum_counts")
    return int(out.item())


COOC_MAX_MOTIFS = 31          # csrc/consumers.cu CO_MAX: one presence bit per motif, bit 31 flags a read left to the host


def co_occurrence_scan(scan_dev):
    """get_motif_co_occurence_mat's integer parts from the occurrence-scan results of m <= 31 consensus sequences on the device
    (csrc/consumers.cu): (counts int64[m, m] -- [i][i] reads listing motif i, [i][j] (i < j) reads listing both --,
    {(i, j): (read index int64[], twice the difference of the median positions int32[]) in read order},
    reads left to the host int64[]: those with a cell of more than 20 positions, which are in none of the counts)."""
    L = lib()
    m = len(scan_dev)
    n_seq = int(scan_dev[0][1].numel()) - 1
    offs = (ctypes.c_void_p * m)(*[t[1].data_ptr() for t in scan_dev])
    poss = (ctypes.c_void_p * m)(*[_ptr(t[2]) if t[2].numel() else None for t in scan_dev])
    present = empty(n_seq, torch.int32)
    counts_d = empty(m * m, torch.int64)
    check(L.kmap_cooc_reads(offs, poss, m, n_seq, _ptr(present), _ptr(counts_d), _stream_ptr()), "kmap_cooc_reads")
    counts = counts_d.cpu().numpy().reshape(m, m)
    over = torch.nonzero(present < 0).flatten().cpu().numpy().astype(np.int64) if n_seq else np.zeros(0, dtype=np.int64)
    pairs = {}
    flags = empty(n_seq, torch.int32)
    for i in range(m):
        for j in range(i + 1, m):
            c = int(counts[i, j])
            if c == 0:
                pairs[(i, j)] = (np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.int32))
                continue
            check(L.kmap_cooc_pair_flags(_ptr(present), n_seq, i, j, _ptr(flags), _stream_ptr()), "kmap_cooc_pair_flags")
            at = exclusive_scan_u32(flags)
            reads, diff2 = empty(c, torch.int64), empty(c, torch.int32)
            check(L.kmap_cooc_pair_fill(scan_dev[i][1].data_ptr(), _ptr(scan_dev[i][2]), scan_dev[j][1].data_ptr(), _ptr(scan_dev[j][2]),
                                        _ptr(flags), _ptr(at), n_seq, _ptr(reads), _ptr(diff2), _stream_ptr()), "kmap_cooc_pair_fill")
            pairs[(i, j)] = (to_host(reads), to_host(diff2))
    return counts, pairs, over


# ---- sort / run-length path (csrc/sorted.cu): uint64 hashes, int64 counts ------------------------------------------------
def sort_count_keys(keys: torch.Tensor, key_bits: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """np.unique(return_counts) of a uint64 key array on the device (all-ones keys dropped).  `keys` is consumed."""
    L = lib()
    n = int(keys.numel())
    if n == 0:
        return empty(0, torch.int64), empty(0, torch.int64)
    tmp = empty(n, torch.int64)
    scratch = empty(L.kmap_sort_scratch_words(n), torch.int64)
    n_valid, n_uniq = ctypes.c_int64(0), ctypes.c_int64(0)
    check(L.kmap_sort_keys_u64(_ptr(keys), _ptr(tmp), n, int(key_bits), _ptr(scratch), ctypes.byref(n_valid), ctypes.byref(n_uniq),
                               _stream_ptr()), "kmap_sort_keys_u64")
    kh, cnt = empty(n_uniq.value, torch.int64), empty(n_uniq.value, torch.int64)
    if n_uniq.value:
        check(L.kmap_rle_u64(_ptr(keys), n_valid.value, _ptr(scratch), _ptr(tmp), _ptr(kh), _ptr(cnt), n_uniq.value, _stream_ptr()),
              "kmap_rle_u64")
    return kh, cnt


def merge_revcom_sorted(kh: torch.Tensor, cnt: torch.Tensor, k: int, want_summed: bool = False, keep_higher: bool = False):
    """merge_revcom (kmer_count.py:643-685) on an ascending unique (uint64 kh, int64 cnt) device list -> (kh, cnt) in the
    reference's order [, the summed counts the reference leaves in the caller's array]."""
    L = lib()
    n = int(kh.numel())
    if n == 0:
        return (kh, cnt, cnt) if want_summed else (kh, cnt)
    scratch = empty(L.kmap_merge_sorted_scratch_words(n), torch.int64)
    n_out = ctypes.c_int64(0)
    out_kh, out_cnt = empty(n, torch.int64), empty(n, torch.int64)      # survivors <= n: one call
    summed = empty(n, torch.int64) if want_summed else None
    check(L.kmap_merge_revcom_sorted_u64(_ptr(kh), _ptr(cnt), n, k, int(keep_higher), _ptr(scratch), _ptr(out_kh), _ptr(out_cnt), n, ctypes.byref(n_out),
                                         _ptr(summed), _stream_ptr()), "kmap_merge_revcom_sorted_u64")
    out = (out_kh[:n_out.value], out_cnt[:n_out.value])
    return out + (summed,) if want_summed else out


def hamball_sums_list64(kh: torch.Tensor, cnt: torch.Tensor, k: int, cand: Sequence[int], d: int, revcom: bool) -> np.ndarray:
    L = lib()
    out = np.zeros(len(cand), dtype=np.int64)
    for i in range(0, len(cand), 16):
        part = list(cand[i:i + 16])
        c = to_device(np.asarray([int(x) for x in part], dtype=np.uint64))
        sums = empty(len(part), torch.int64)
        check(L.kmap_hamball_sum_list_u64(_ptr(kh), _ptr(cnt), int(kh.numel()), k, _ptr(c), len(part), int(d), int(revcom),
                                          _ptr(sums), _stream_ptr()), "kmap_hamball_sum_list_u64")
        out[i:i + 16] = sums.cpu().numpy()
    return out


# ---- preproc ingest: FASTA text -> device-resident input.bin / input.seqboarder.bin contents --------------------------
_FIRST_HEADER = None


def read_fasta_bytes(fasta_file) -> np.ndarray:
    """the file's bytes (gunzipped when the name ends in .gz, like kmer_count.py:318-323), from the first header line on"""
    import gzip
    import re
    global _FIRST_HEADER
    if _FIRST_HEADER is None:
        _FIRST_HEADER = re.compile(rb"(?:\A|[\r\n])>")
    if str(fasta_file).endswith(".gz"):
        with gzip.open(fasta_file, "rb") as fh:
            raw = np.frombuffer(bytearray(fh.read()), dtype=np.uint8)
    else:
        raw = np.fromfile(fasta_file, dtype=np.uint8)
    m = _FIRST_HEADER.search(memoryview(raw))
    if m is None:
        return raw[:0]
    return raw[m.end() - 1:]


def _fasta_chunks_to_device(chunks) -> Tuple[torch.Tensor, torch.Tensor]:
    """core of the ingest: `chunks` yields (device uint8 tensor, is_final) in file order"""
    L = lib()
    state = (ctypes.c_int64 * 4)(0, 0, 0, 10)
    parts, starts = [], []
    for chunk, final in chunks:
        nc = int(chunk.numel())
        scratch = empty(L.kmap_fasta_scratch_words(nc), torch.int64)
        out = (ctypes.c_int64 * 4)()
        check(L.kmap_fasta_scan(_ptr(chunk) if nc else None, nc, state, _ptr(scratch), out, _stream_ptr()), "kmap_fasta_scan")
        new_rec = out[1] - state[1]
        n_out = (out[0] - state[0]) + new_rec - (1 if state[1] == 0 and out[1] > 0 else 0) + (1 if final and out[1] > 0 else 0)
        origin = state[0] + max(state[1] - 1, 0)
        seq_part = empty(n_out, torch.uint8)
        rec_start = empty(new_rec, torch.int64)
        check(L.kmap_fasta_emit(_ptr(chunk) if nc else None, nc, state, _ptr(scratch), _ptr(seq_part) if n_out else None, origin,
                                _ptr(rec_start) if new_rec else None, int(final), out, _stream_ptr()), "kmap_fasta_emit")
        parts.append(seq_part)
        starts.append(rec_start)
        state = out
    seq = parts[0] if len(parts) == 1 else torch.cat(parts)
    rec_start = starts[0] if len(starts) == 1 else torch.cat(starts)
    n_rec = int(state[1])
    total = int(state[0]) + n_rec if n_rec else 0
    if int(seq.numel()) != total or int(rec_start.numel()) != n_rec:
        raise KmapError(f"fasta ingest: {seq.numel()} bytes / {rec_start.numel()} records emitted, {total} / {n_rec} expected")
    borders = empty(2 * n_rec, torch.int64)
    check(L.kmap_borders_from_starts(_ptr(rec_start), n_rec, total, _ptr(borders), _stream_ptr()), "kmap_borders_from_starts")
    return seq, borders.view(n_rec, 2)


def fasta_text_to_device(text: np.ndarray, chunk_bytes: int = 1 << 28) -> Tuple[torch.Tensor, torch.Tensor]:
    """(seq uint8[n_pos], borders int64[n_seq, 2]) on the device from FASTA text that starts at its first header line
    (csrc/fasta.cu; reference kmer_count.py:244-263, 326-347).  The text travels in chunks of `chunk_bytes`."""
    dev = require_cuda()
    text = np.ascontiguousarray(text, dtype=np.uint8)
    n = len(text)
    step = max(int(chunk_bytes), 16)

    def chunks():
        pos = 0
        while True:
            end = min(pos + step, n)
            yield (torch.from_numpy(text[pos:end]).to(dev) if end > pos else torch.empty(0, dtype=torch.uint8, device=dev)), end >= n
            pos = end
            if end >= n:
                return
    return _fasta_chunks_to_device(chunks())


def fasta_to_device(fasta_file, chunk_bytes: int = 1 << 28) -> Tuple[torch.Tensor, torch.Tensor]:
    """the file -> (seq, borders) on the device.  A plain file is read straight into two alternating pinned staging buffers
    (no pageable copy of the text); a .gz file is inflated on the host first (kmer_count.py:318-323)."""
    import os
    import re
    if str(fasta_file).endswith(".gz"):
        return fasta_text_to_device(read_fasta_bytes(fasta_file), chunk_bytes)
    dev = require_cuda()
    size = os.path.getsize(fasta_file)
    step = max(int(chunk_bytes), 1 << 16)
    first_header = re.compile(rb"(?:\A|[\r\n])>")
    later_header = re.compile(rb"[\r\n]>")            # (the byte carried over from the previous block is not a line start)

    def chunks():
        with open(fasta_file, "rb", buffering=0) as fh:
            # skip whatever precedes the first header line
            pos, tail = 0, b""
            start = None
            while pos < size and start is None:
                block = fh.read(1 << 20)
                if not block:
                    break
                m = (first_header if pos == 0 else later_header).search(tail + block)
                if m is not None:
                    start = pos - len(tail) + m.end() - 1
                tail = block[-1:]
                pos += len(block)
            if start is None:
                yield torch.empty(0, dtype=torch.uint8, device=dev), True
                return
            fh.seek(start)
            pos = start
            cap = min(step, size - start)
            stage = [torch.empty(cap, dtype=torch.uint8).pin_memory() for _ in range(2 if size - start > cap else 1)]
            busy = [None] * len(stage)
            i = 0
            while True:
                buf = stage[i % len(stage)]
                if busy[i % len(stage)] is not None:
                    busy[i % len(stage)].synchronize()          # the copy that last used this buffer is done
                got = fh.readinto(memoryview(buf.numpy())[:min(cap, size - pos)])
                got = got or 0
                pos += got
                final = pos >= size or got == 0
                d = torch.empty(got, dtype=torch.uint8, device=dev)
                if got:
                    d.copy_(buf[:got], non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record()
                    busy[i % len(stage)] = ev
                yield d, final
                if final:
                    return
                i += 1
    return _fasta_chunks_to_device(chunks())
