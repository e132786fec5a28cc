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
from ._lib import KmapError, check as _check, lib as _lib


class _DeviceBlob:
    """n int32 cells at a raw device pointer (memory owned by libkmap_b200), for torch.as_tensor"""

    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<i4", "data": (int(ptr), False), "version": 3}


class TableAllReduce:
    """In-place sum of a dense count table over all ranks.  uint32 counts travel as their int32 bit patterns: two's
    complement addition is the same modular sum, so the merged table is bit-identical for any number of ranks.
    Device tensors go through the library's own exchange step (`kmap_table_allreduce`, include/kmap_b200.h: an NCCL
    communicator created from a unique id that torch.distributed broadcasts -- the process group is plumbing only); host
    tensors (the gloo tests of the host logic) through torch.distributed."""

    def __init__(self, group=None, scatter: bool = False):
        """scatter=True: counts merged from inside SeqOnDevice.count_all are left SCATTERED over the ranks by key range
        (reduce-scatter instead of all-reduce, `kmap_count_all_k_scattered`): rank r owns cells `owned_range(k)` of every level,
        its other cells hold partial sums.  Half the exchange volume; for consumers that work on key ranges."""
        import torch.distributed as dist
        if not dist.is_initialized():
            raise KmapError("TableAllReduce needs an initialised torch.distributed process group")
        self.dist, self.group = dist, group
        self.scatter = bool(scatter)
        self._comm = None
        self._stream = None
        self._region = None            # peer-memory exchange (csrc/peer.cu): (device pointer, table cells) of this rank's region
        self._peer_refused = None      # why the peer-memory exchange is not in use (None: not tried yet / in use)

    def owned_range(self, k: int) -> Tuple[int, int]:
        """[lo, hi) of the level-k table this rank owns after a scattered merge"""
        cells = 1 << (2 * k)
        if cells % self.world:
            raise KmapError(f"{self.world} ranks do not divide 4^{k} cells")
        return cells // self.world * self.rank, cells // self.world * (self.rank + 1)

    def reduce_scatter(self, table: torch.Tensor) -> torch.Tensor:
        """in place: this rank's block of `table` (numel / world cells at block index rank) becomes the sum over the ranks"""
        comm, _ = self.native()
        _check(_lib().kmap_table_reduce_scatter(table.data_ptr(), table.numel(), self.rank, self.world, comm,
                                                torch.cuda.current_stream().cuda_stream), "kmap_table_reduce_scatter")
        return table

    # ---- the native communicator (lazy: the first device tensor creates it; collective over the group) ----------------
    def native(self):
        """(comm handle, exchange stream) of libkmap_b200 on the current device"""
        if self._comm is None:
            import ctypes
            L = _lib()
            rank, world = self.dist.get_rank(self.group), self.dist.get_world_size(self.group)
            ident = (ctypes.c_uint8 * 128)()
            if rank == 0:
                _check(L.kmap_comm_unique_id(ident), "kmap_comm_unique_id")
            box = [bytes(ident)]
            self.dist.broadcast_object_list(box, src=self.dist.get_global_rank(self.group, 0) if self.group is not None else 0,
                                            group=self.group)
            ident = (ctypes.c_uint8 * 128).from_buffer_copy(box[0])
            comm = ctypes.c_void_p()
            _check(L.kmap_comm_init(ident, rank, world, ctypes.byref(comm)), "kmap_comm_init")
            self._comm = comm
            self._stream = torch.cuda.Stream()
        return self._comm, self._stream

    # ---- the peer-memory exchange: tables that live in a region every rank of the node has mapped ---------------------
    def alloc_tables(self, kmin: int, kmax: int, zero: bool = False):
        """(flat, {k: view}) like engine.alloc_tables, for counts merged through this object.  On one NVLink node (2..8 ranks)
        the tables are views of this rank's PEER REGION, and every merge of them -- `count_all(merge=self)`, `self(table)` --
        takes the one-byte-per-cell exchange over peer memory (csrc/peer.cu, kmap_comm_attach_peers) instead of the NCCL
        all-reduce; bit-identical.  The region holds ONE set of tables: a second call hands out the same memory.  Collective
        on first use (every rank must call it with the same levels).  KMAP_PEER_EXCHANGE=0 keeps NCCL."""
        total = sum(1 << (2 * k) for k in range(kmin, kmax + 1))
        if not self._ensure_region(total):
            return E.alloc_tables(kmin, kmax, zero)
        ptr, _ = self._region
        flat = torch.as_tensor(_DeviceBlob(ptr, total), device=E.require_cuda())
        if flat.data_ptr() != ptr:
            raise KmapError("torch copied the peer region instead of wrapping it")
        if zero:
            flat.zero_()
        views, o = {}, 0
        for k in range(kmax, kmin - 1, -1):
            views[k] = flat[o:o + (1 << (2 * k))]
            o += 1 << (2 * k)
        return flat, views

    @property
    def peer_exchange(self) -> bool:
        return self._region is not None

    def _ensure_region(self, cells: int) -> bool:
        import ctypes
        import os
        import socket
        cells = (int(cells) + 15) // 16 * 16
        if self._region is not None and self._region[1] >= cells:
            return True
        if self._peer_refused is not None:
            return False
        L = _lib()
        comm, _ = self.native()
        rank, world = self.rank, self.world
        why = None
        if os.environ.get("KMAP_PEER_EXCHANGE", "1") == "0":
            why = "KMAP_PEER_EXCHANGE=0"
        elif not 2 <= world <= 8:
            why = f"{world} ranks (the peer-memory exchange takes 2..8 ranks of one node)"
        else:
            hosts = [None] * world
            self.dist.all_gather_object(hosts, (socket.gethostname(), torch.cuda.current_device()), group=self.group)
            if len({h for h, _ in hosts}) != 1:
                why = "the ranks are on more than one node"
            elif len({d for _, d in hosts}) != world:
                why = "two ranks share a device"
            elif not all(torch.cuda.can_device_access_peer(torch.cuda.current_device(), d) for _, d in hosts
                         if d != torch.cuda.current_device()):
                why = "no peer access between the devices"
        if why is None:
            if self._region is not None:                      # (grown: every rank drops its mapping before any region is freed)
                torch.cuda.synchronize()
                _check(L.kmap_comm_detach_peers(comm), "kmap_comm_detach_peers")
                self.dist.barrier(group=self.group)
                _check(L.kmap_peer_region_free(self._region[0]), "kmap_peer_region_free")
                self._region = None
            region, handle = ctypes.c_void_p(), (ctypes.c_uint8 * 64)()
            ok = L.kmap_peer_region_alloc(cells, ctypes.byref(region), handle) == 0
            box = [None] * world
            self.dist.all_gather_object(box, bytes(handle) if ok else None, group=self.group)
            if any(b is None for b in box):
                why = "a rank could not allocate its region: " + L.kmap_last_error().decode(errors="replace")
            else:
                handles = (ctypes.c_uint8 * (64 * world)).from_buffer_copy(b"".join(box))
                ok = L.kmap_comm_attach_peers(comm, rank, world, region, cells, handles) == 0
                oks = [None] * world
                self.dist.all_gather_object(oks, ok, group=self.group)
                if all(oks):
                    self._region = (region.value, cells)
                    return True
                why = "a rank could not map its peers: " + L.kmap_last_error().decode(errors="replace")
                if ok:
                    L.kmap_comm_detach_peers(comm)
                self.dist.barrier(group=self.group)
            if ok or region.value:
                L.kmap_peer_region_free(region)
        self._peer_refused = why
        return False

    def check(self):
        """raises if a rank failed to reach a barrier of the peer-memory exchange (synchronises the exchange stream)"""
        if self._comm is not None and self._region is not None:
            import ctypes
            status = ctypes.c_int(0)
            cur = torch.cuda.current_stream()
            cur.wait_stream(self._stream)               # (exchanges run on the exchange stream or, for self(table), on the current one)
            _check(_lib().kmap_comm_peer_status(self._comm, ctypes.byref(status), cur.cuda_stream), "kmap_comm_peer_status")

    def close(self):
        if self._comm is not None:
            torch.cuda.synchronize()
            region = self._region
            self._region = None
            if region is not None:
                _check(_lib().kmap_comm_detach_peers(self._comm), "kmap_comm_detach_peers")
                self.dist.barrier(group=self.group)          # nobody frees a region a peer still has mapped
                _check(_lib().kmap_peer_region_free(region[0]), "kmap_peer_region_free")
            _check(_lib().kmap_comm_destroy(self._comm), "kmap_comm_destroy")
            self._comm = None

    def __call__(self, table: torch.Tensor) -> torch.Tensor:
        if table.is_cuda:
            comm, _ = self.native()
            _check(_lib().kmap_table_allreduce(table.data_ptr(), table.numel(), comm, torch.cuda.current_stream().cuda_stream),
                   "kmap_table_allreduce")
        else:
            self.dist.all_reduce(table, op=self.dist.ReduceOp.SUM, group=self.group)
        return table

    def sum_int(self, value: int, device=None) -> int:
        """sum of one host integer over the ranks (e.g. the window count that bounds the merged lists)"""
        t = torch.tensor([int(value)], dtype=torch.int64, device=device if device is not None else "cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
        return int(t.item())

    @property
    def rank(self) -> int:
        return self.dist.get_rank(self.group)

    @property
    def world(self) -> int:
        return self.dist.get_world_size(self.group)


class DistContext:
    """One process per GPU (launched by `python -m torch.distributed.run ... -m kmap_b200 scan_motif ...`): what the
    reference-shaped drivers need from the process group.  Reads shard by contiguous ranges (reads are the independent
    units of the counting path, kmer_count.py:755-759), dense tables merge by an integer all-reduce, per-read results are
    gathered on rank 0, which alone writes files.  With one process everything is a no-op."""

    def __init__(self, dist=None, group=None):
        self.dist, self.group = dist, group
        self.rank = dist.get_rank(group) if dist is not None else 0
        self.world = dist.get_world_size(group) if dist is not None else 1
        self.table_allreduce = TableAllReduce(group) if self.world > 1 else None

    @classmethod
    def from_env(cls) -> "DistContext":
        """WORLD_SIZE > 1 in the environment (torchrun): join (or create, backend nccl when CUDA is there, else gloo) the
        default process group and bind this process to cuda:LOCAL_RANK."""
        import os
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return cls(dist) if dist.get_world_size() > 1 else cls()
        if int(os.environ.get("WORLD_SIZE", "1")) <= 1:
            return cls()
        if torch.cuda.is_available():
            local = int(os.environ.get("LOCAL_RANK", "0"))
            torch.cuda.set_device(local)
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        else:
            dist.init_process_group("gloo")
        return cls(dist)

    @property
    def is_root(self) -> bool:
        return self.rank == 0

    def allreduce(self, table: torch.Tensor) -> torch.Tensor:
        return table if self.table_allreduce is None else self.table_allreduce(table)

    def agree(self, obj):
        """rank 0's value on every rank (decisions that depend on files rank 0 may be writing)"""
        if self.world == 1:
            return obj
        box = [obj]
        self.dist.broadcast_object_list(box, src=0, group=self.group)
        return box[0]

    def gather(self, obj):
        """[obj of rank 0, obj of rank 1, ...] on rank 0, None elsewhere"""
        if self.world == 1:
            return [obj]
        out = [None] * self.world if self.rank == 0 else None
        self.dist.gather_object(obj, out, dst=0, group=self.group)
        return out

    def barrier(self):
        if self.world > 1:
            self.dist.barrier(group=self.group)


def concat_occurrence_shards(parts):
    """per-shard results of the occurrence scan, in rank order -> the result for all reads:
    parts[r] = (min_dist uint8[n_r], offsets int64[n_r + 1], positions int32[t_r]) (engine.occurrence_scan)."""
    min_dist = np.concatenate([p_[0] for p_ in parts])
    offs, base = [np.zeros(1, dtype=np.int64)], 0
    for _, o, _ in parts:
        o = np.asarray(o, dtype=np.int64)
        offs.append(o[1:] + base)
        base += int(o[-1]) if len(o) else 0
    return min_dist, np.concatenate(offs), np.concatenate([np.asarray(p_[2], dtype=np.int32) for p_ in parts])


def merge_sorted_counts_over_ranks(kh: torch.Tensor, cnt: torch.Tensor, ctx: DistContext, key_bits: int = 64):
    """k >= 16 has no dense table to all-reduce (kmer_count.py:359-365: uint64 hashes): every rank sorts and run-length
    encodes its own reads (SeqOnDevice.count_sorted), the (hash, count) lists are all-gathered, and every rank builds the
    same merged list: distinct hashes of all ranks ascending, counts added up -- what count_uniq_hash returns for the
    whole input (kmer_count.py:476-491)."""
    if ctx.world == 1:
        return kh, cnt
    L = _lib()
    dist = ctx.dist
    n_local = torch.tensor([int(kh.numel())], dtype=torch.int64, device=kh.device)
    sizes = [torch.zeros_like(n_local) for _ in range(ctx.world)]
    dist.all_gather(sizes, n_local, group=ctx.group)
    sizes = [int(x.item()) for x in sizes]
    cap = max(max(sizes), 1)
    pad_kh, pad_cnt = torch.zeros(cap, dtype=torch.int64, device=kh.device), torch.zeros(cap, dtype=torch.int64, device=kh.device)
    pad_kh[:kh.numel()] = kh
    pad_cnt[:cnt.numel()] = cnt
    all_kh = [torch.empty_like(pad_kh) for _ in range(ctx.world)]
    all_cnt = [torch.empty_like(pad_cnt) for _ in range(ctx.world)]
    dist.all_gather(all_kh, pad_kh, group=ctx.group)
    dist.all_gather(all_cnt, pad_cnt, group=ctx.group)
    lists = [(a[:m], c[:m]) for a, c, m in zip(all_kh, all_cnt, sizes) if m]
    if not lists:
        return kh[:0], cnt[:0]
    uniq, _ = E.sort_count_keys(torch.cat([a for a, _ in lists]), key_bits)     # distinct hashes of all ranks, ascending
    total = torch.zeros(uniq.numel(), dtype=torch.int64, device=kh.device)
    stream = torch.cuda.current_stream().cuda_stream
    for a, c in lists:
        a, c = a.contiguous(), c.contiguous()
        _check(L.kmap_list_add_counts_u64(uniq.data_ptr(), uniq.numel(), a.data_ptr(), c.data_ptr(), a.numel(), total.data_ptr(), stream),
               "kmap_list_add_counts_u64")
    return uniq, total


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
                lists_on: Optional[int] = None, chunk_positions: int = 1 << 29,
                host_pack: Optional[bool] = None) -> Dict[int, Tuple[np.ndarray, np.ndarray]]:
    """First-round counts of find_motif for every k (reference motif_discovery.py:627-640): per k the
    `(uniq_kh_arr uint32, uniq_kh_cnt_arr int32)` pair that the reference pickles into kmer_count/k{k}.pkl, in the
    reference's order.  With `table_allreduce` each rank passes ITS shard of the reads and the tables are merged before
    compaction; `lists_on=r` returns the lists on rank r only (others get {}); `lists_on="sharded"` returns on every rank the
    slice of each list whose forward hashes fall into the rank's key range (`engine.key_range`): the slices of ranks 0, 1, ..
    concatenate to the whole list, and compaction and the device-to-host copies run on all ranks at once.
    Inputs larger than `chunk_positions` are streamed through the device in chunks of whole reads (`count_tables_streamed`):
    the host cores re-encode chunks into the packed form the device works on (0.375 B/position over the link instead of 1)
    while the copy engine ships other chunks as they are, and every chunk is counted while the next ones travel.
    host_pack=False ships every chunk as it is (None: on unless KMAP_HOST_PACK=0)."""
    out: Dict[int, Tuple[np.ndarray, np.ndarray]] = {}
    ks_all = sorted(set(int(k) for k in k_list))
    wide = [k for k in ks_all if k >= 16]          # uint64 hashes, no dense table: sort / run-length path (csrc/sorted.cu)
    ks = [k for k in ks_all if k < 16]
    if wide:
        # no dense table to all-reduce: every rank sorts / run-length encodes its shard, the lists are merged over the ranks
        ctx = DistContext(table_allreduce.dist, table_allreduce.group) if table_allreduce is not None else None
        dev = upload_reads(seq_np_arr, boarder_mat, validate)
        for k in wide:
            kh, cnt = dev.count_sorted(k, dedup=not rep_mode)
            if ctx is not None:
                kh, cnt = merge_sorted_counts_over_ranks(kh, cnt, ctx, 2 * k)
            if lists_on is not None and ctx is not None and ctx.rank != (0 if lists_on == "sharded" else lists_on):
                continue
            if revcom_mode:
                kh, cnt = E.merge_revcom_sorted(kh, cnt, k)
            out[k] = (E.to_host(kh, np.uint64), E.to_host(cnt, np.int64))
        del dev
    if not ks:
        return out
    contiguous = ks == list(range(ks[0], ks[-1] + 1))
    bounds = _chunk_bounds(seq_np_arr, boarder_mat, chunk_positions) if contiguous else None
    flat = None
    merged_inside = False
    if bounds is not None and len(bounds) > 2:
        # (the layout check of `validate` is inherent here: _chunk_bounds verified the ends and the cuts, and the stride encoder
        #  verifies every row inside a chunk, kmap_host_border_strides)
        world_local = table_allreduce.world if table_allreduce is not None else 1
        flat, tables, n_total = count_tables_streamed(seq_np_arr, boarder_mat, ks[0], ks[-1], rep_mode, bounds, host_pack,
                                                      host_pack_threads(world_local),
                                                      alloc_totals=table_allreduce.alloc_tables if table_allreduce is not None else None)
    else:
        dev = upload_reads(seq_np_arr, boarder_mat, validate)
        n_total = dev.n
        if contiguous:
            # (sharded: the tables live in the peer region of the exchange, so the merge takes the peer-memory path)
            flat, tables = (table_allreduce.alloc_tables if table_allreduce is not None else E.alloc_tables)(ks[0], ks[-1])
            # one update per window at kmax, the rest derived; sharded: merged over the ranks from inside the count
            dev.count_all(ks[0], ks[-1], dedup=not rep_mode, tables=tables, merge=table_allreduce)
            merged_inside = table_allreduce is not None
        else:
            tables = {k: dev.count(k, dedup=not rep_mode) for k in ks}
        del dev
    if table_allreduce is not None:
        if flat is not None:
            if not merged_inside:
                table_allreduce(flat)                              # every level in one call
        else:
            for k in ks:
                table_allreduce(tables[k])
        # the merged table holds the distinct k-mers of EVERY shard: the capacity bound is the total window count
        n_total = table_allreduce.sum_int(n_total, device=tables[ks[0]].device)
        table_allreduce.check()                                    # (peer-memory exchange: every rank reached every barrier)
        if lists_on is not None and lists_on != "sharded" and table_allreduce.rank != lists_on:
            return out
    sharded_lists = table_allreduce is not None and lists_on == "sharded"
    # compaction of table k+1 overlaps the device-to-host copy of the lists of k (copy stream + pinned buffers)
    copy_stream = _copy_stream()
    pending = []
    for k in ks:
        cells = E.key_range(k, table_allreduce.rank, table_allreduce.world) if sharded_lists else None
        kh, cnt = E.compact_merge(tables[k], k, revcom_mode, upper_bound=_upper_bound(n_total, k), cell_range=cells)
        done = torch.cuda.Event()
        done.record()
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(done)
            h_kh, h_cnt = _pinned_like(kh), _pinned_like(cnt)
            h_kh.copy_(kh, non_blocking=True)
            h_cnt.copy_(cnt, non_blocking=True)
        pending.append((k, kh, cnt, h_kh, h_cnt))
    copy_stream.synchronize()
    for k, kh, cnt, h_kh, h_cnt in pending:
        out[k] = (h_kh.numpy().view(np.uint32), h_cnt.numpy())
    return out


_copy_streams = {}


def _copy_stream(which: int = 0) -> torch.cuda.Stream:
    key = (torch.cuda.current_device(), which)
    if key not in _copy_streams:
        _copy_streams[key] = torch.cuda.Stream()
    return _copy_streams[key]


def _upper_bound(n: int, k: int) -> int:
    """no table can hold more distinct k-mers than cells or than windows"""
    return int(min(1 << (2 * k), max(n - k + 1, 0)))


def _first_read_at_or_after(b: np.ndarray, pos: int) -> int:
    """index of the first read whose start is >= pos (bisection on the border matrix itself: no strided copy)"""
    lo, hi = 0, len(b)
    while lo < hi:
        mid = (lo + hi) // 2
        if b[mid, 0] < pos:
            lo = mid + 1
        else:
            hi = mid
    return lo


def _chunk_bounds(seq_np_arr: np.ndarray, boarder_mat, chunk_positions: int):
    """read indices [r_0 = 0 < r_1 < ... < r_m = n_seq] cutting the input into chunks of about chunk_positions positions,
    or None when the input is small or the border matrix is not the back-to-back layout `kmap preproc` writes at the cuts
    (kmer_count.py:335-343)."""
    n = len(seq_np_arr)
    if boarder_mat is None or n <= chunk_positions:
        return None
    b = np.asarray(boarder_mat).reshape(-1, 2)
    n_seq = len(b)
    if n_seq < 2 or b[0, 0] != 0 or b[-1, 1] != n - 1:
        return None
    m = -(-n // chunk_positions)
    cuts = [0]
    for i in range(1, m):
        r = _first_read_at_or_after(b, n * i // m)
        if cuts[-1] < r < n_seq:
            if b[r, 0] != b[r - 1, 1] + 1:
                return None
            cuts.append(r)
    cuts.append(n_seq)
    return cuts


_staging = {}


def _os_environ_get(key, default):
    import os
    return os.environ.get(key, default)

last_stream_trace = None
last_stream_stats = None


def _pinned_staging(tag: str, n: int, dtype: torch.dtype) -> torch.Tensor:
    """pinned host staging buffer of at least n elements, kept between calls (pinning memory costs ~0.3 s per GB)"""
    key = (tag, dtype)
    buf = _staging.get(key)
    if buf is None or buf.numel() < n:
        _staging[key] = None
        buf = _staging[key] = torch.empty(n, dtype=dtype, pin_memory=True)
    return buf


def host_pack_threads(world_local: int = 1) -> int:
    """host threads one process may use for re-encoding (all hardware threads, shared by the ranks of the box)"""
    import os
    n = int(os.environ.get("KMAP_HOST_THREADS", "0"))
    if n <= 0:                       # all hardware threads but two: the feeder and the counting thread need to run too
        n = max(1, int(_lib().kmap_host_threads()) - 2)
    return max(1, n // max(1, world_local))


def count_tables_streamed(seq_np_arr: np.ndarray, boarder_mat: np.ndarray, kmin: int, kmax: int, rep_mode: bool, cuts,
                          host_pack: Optional[bool] = None, n_threads: int = 0, alloc_totals=None):
    """Dense forward tables of every k in [kmin, kmax] with the reads streamed through the device chunk by chunk.
    Reads are independent units (per-read de-duplication never crosses a read, kmer_count.py:755-759), so the table of the
    whole input is the sum of the chunk tables (SeqOnDevice.count_all per chunk + kmap_add_u32).
    The call is bound by the host-to-device link when input.bin travels at one byte per base, so two feeders share the
    chunks: a host thread re-encodes chunks from the FRONT into the packed form (csrc/host_pack.cpp, all host cores, 0.375
    B/position, into pinned staging buffers) and ships them packed; the main thread ships chunks from the BACK as they are
    (one raw copy in flight at a time; packed on the device) so that the link never waits for the encoder.  They meet
    wherever the two rates put them.  The border rows travel as one uint32 stride per read (kmap_host_border_strides) and
    are rebuilt on the device by a prefix sum.  host_pack=False: raw chunks only.  Returns (flat buffer, {k: table view}, n)."""
    import queue
    import threading
    E.require_cuda()
    L = _lib()
    if host_pack is None:
        import os
        host_pack = os.environ.get("KMAP_HOST_PACK", "1") != "0"
    b = np.asarray(boarder_mat).reshape(-1, 2)
    if b.dtype != np.int64 or not b.flags.c_contiguous:
        b = np.ascontiguousarray(b, dtype=np.int64)
    seq_np = np.ascontiguousarray(seq_np_arr)
    seq_t = torch.from_numpy(seq_np)
    n_threads = n_threads or host_pack_threads()
    spans = []
    for r0, r1 in zip(cuts[:-1], cuts[1:]):
        spans.append((r0, r1, int(b[r0, 0]), int(b[r1 - 1, 1]) + 1))
    n_chunks = len(spans)
    max_pos = max(p1 - p0 for _, _, p0, p1 in spans)
    max_rows = max(r1 - r0 for r0, r1, _, _ in spans)
    vw, pw = int(L.kmap_valid_words(max_pos)), int(L.kmap_packed_words(max_pos))
    # device side: slots used in the order the chunks are enqueued; enough of them that neither the link nor the count
    # waits for the other over a few chunks (a slot is 0.375 B/position, + 1 B/position if a raw chunk ever lands in it)
    n_slots = min(int(_os_environ_get('KMAP_STREAM_SLOTS', '6')), n_chunks)
    d_u8 = [None] * n_slots
    d_packed = [E.empty(pw, torch.int32) for _ in range(n_slots)]
    d_valid = [E.empty(vw, torch.int32) for _ in range(n_slots)]
    d_strides = [E.empty(max_rows, torch.int32) for _ in range(n_slots)]
    d_offsets = E.empty(max_rows + 1, torch.int64)
    d_borders = torch.empty((max_rows, 2), dtype=torch.int64, device="cuda")
    d_scan = E.empty(int(L.kmap_list_scratch_words(max_rows)), torch.int64)
    # host side: two staging slots for the encoder, two stride slots for the raw feeder
    N_STAGED = 4
    h_packed = [_pinned_staging(f"packed{i}", pw, torch.int32) for i in range(N_STAGED)]
    h_valid = [_pinned_staging(f"valid{i}", vw, torch.int32) for i in range(N_STAGED)]
    max_raw = int(_os_environ_get('KMAP_STREAM_MAX_RAW', '1')) if host_pack else 2      # raw copies queued on the link at a time
    h_strides = [_pinned_staging(f"strides{i}", max_rows, torch.int32) for i in range(N_STAGED + max_raw + 1)]
    compute = torch.cuda.current_stream()
    # two copy streams: the packed chunks do not queue behind the (2.7 x longer) raw copies, so the encoder gets its staging
    # slots back as soon as the link has taken them
    copy = _copy_stream()
    copy_packed = _copy_stream(1)
    copy.wait_stream(compute)
    copy_packed.wait_stream(compute)
    flat, totals = (alloc_totals or E.alloc_tables)(kmin, kmax)
    part_flat, part = E.alloc_tables(kmin, kmax)

    lock = threading.Lock()
    nxt = {"front": 0, "back": n_chunks - 1}
    import os as _os
    import time as _time
    trace = [] if _os.environ.get("KMAP_STREAM_TRACE") else None       # (debugging aid: host-side timeline of the three threads)
    t_origin = _time.perf_counter()

    def mark(what, i=-1):
        if trace is not None:
            trace.append((round(1e3 * (_time.perf_counter() - t_origin), 2), threading.current_thread().name, what, i))
    mark("setup done")

    def take(side):
        with lock:
            if nxt["front"] > nxt["back"]:
                return None
            i = nxt[side]
            nxt[side] += 1 if side == "front" else -1
            return i

    def strides_of(i, out: torch.Tensor):
        r0, r1, p0, _ = spans[i]
        rc = L.kmap_host_border_strides(b.ctypes.data + 16 * r0, r1 - r0, p0, out.data_ptr(), n_threads)
        if rc != 0:
            raise KmapError("boarder_mat rows are not back to back inside a chunk (kmer_count.py:335-343 layout expected)")

    ready = queue.Queue()
    staged_free = [threading.Event() for _ in range(N_STAGED)]    # set once the copy out of the staging slot has been queued
    for e in staged_free:
        e.set()
    staged_copy_done = [None] * N_STAGED
    failed = []

    def encoder():
        try:
            j = 0
            while True:
                slot = j % N_STAGED
                staged_free[slot].wait()
                if staged_copy_done[slot] is not None:
                    staged_copy_done[slot].synchronize()
                i = take("front")
                if i is None:
                    break
                staged_free[slot].clear()
                _, _, p0, p1 = spans[i]
                mark("pack begin", i)
                _check(L.kmap_host_pack2bit(seq_np.ctypes.data + p0, p1 - p0, h_packed[slot].data_ptr(), h_valid[slot].data_ptr(), n_threads),
                       "kmap_host_pack2bit")
                strides_of(i, h_strides[slot])
                mark("pack end", i)
                ready.put((i, slot))
                j += 1
        except BaseException as exc:           # surfaced by the main thread
            failed.append(exc)
        finally:
            ready.put(None)

    enc = None
    if host_pack:
        enc = threading.Thread(target=encoder, name="kmap-host-pack", daemon=True)
        enc.start()
    else:
        ready.put(None)

    # The count of a chunk ends with a host synchronisation (the per-read scan reports reads beyond its on-chip paths), so
    # it runs on a thread of its own: this thread keeps the link busy meanwhile.
    N_SLOTS = len(d_packed)
    slots_free = threading.Semaphore(N_SLOTS)
    released = []                    # per enqueued chunk: event recorded when its count has finished (device slot free again)
    to_count = queue.Queue()
    device_index = torch.cuda.current_device()

    def counter():
        chunk = None
        counted = 0
        try:
            torch.cuda.set_device(device_index)
            with torch.cuda.stream(compute):
                while True:
                    item = to_count.get()
                    if item is None:
                        break
                    i, slot, uploaded, raw = item
                    r0, r1, p0, p1 = spans[i]
                    mark("count begin (host)", i)
                    if trace is not None:
                        uploaded.synchronize()
                        mark("upload landed", i)
                    compute.wait_event(uploaded)
                    if raw:
                        _check(L.kmap_pack2bit(d_u8[slot].data_ptr(), p1 - p0, d_packed[slot].data_ptr(), d_valid[slot].data_ptr(),
                                               compute.cuda_stream), "kmap_pack2bit")
                    _check(L.kmap_borders_from_strides(d_strides[slot].data_ptr(), r1 - r0, d_offsets.data_ptr(), d_borders.data_ptr(),
                                                       d_scan.data_ptr(), compute.cuda_stream), "kmap_borders_from_strides")
                    borders_d = d_borders[:r1 - r0]
                    if chunk is None:
                        chunk = E.SeqOnDevice(p1 - p0, d_packed[slot], d_valid[slot], borders_d, r1 - r0)
                    else:
                        chunk.rebind_packed(p1 - p0, d_packed[slot], d_valid[slot], borders_d)
                    chunk.count_all(kmin, kmax, dedup=not rep_mode, tables=totals if counted == 0 else part)
                    if counted > 0:
                        _check(L.kmap_add_u32(flat.data_ptr(), part_flat.data_ptr(), flat.numel(), compute.cuda_stream), "kmap_add_u32")
                    counted += 1
                    ev = torch.cuda.Event()
                    ev.record(compute)
                    released.append(ev)
                    slots_free.release()
                    if trace is not None:
                        ev.synchronize()
                    mark("count end", i)
        except BaseException as exc:
            failed.append(exc)
            for _ in range(n_chunks + N_SLOTS):          # let the feeder run to its end
                slots_free.release()

    cnt_thread = threading.Thread(target=counter, name="kmap-count", daemon=True)
    cnt_thread.start()

    n_enq = 0

    def device_slot(stream):
        nonlocal n_enq
        slots_free.acquire()                                # the count that used this slot N_SLOTS chunks ago has been queued ...
        if n_enq >= N_SLOTS and not failed:
            stream.wait_event(released[n_enq - N_SLOTS])    # ... and the copy waits until it has finished
        n_enq += 1
        return (n_enq - 1) % N_SLOTS

    raw_events = []                  # events behind the raw copies in flight
    copy_events = []
    n_raw = 0
    encoder_finished = False
    import time
    try:
        while not failed:
            raw_events = [e for e in raw_events if not e.query()]
            chunks_left = nxt["front"] <= nxt["back"]
            can_raw = chunks_left and len(raw_events) < max_raw
            item = False
            if not encoder_finished:
                try:
                    item = ready.get_nowait() if can_raw else ready.get(timeout=0.0002)
                except queue.Empty:
                    item = False
                if item is None:
                    encoder_finished = True
            elif not can_raw:
                if not chunks_left:
                    break
                time.sleep(0.0001)                          # nothing to do until a raw copy in flight has landed
            if item:                                        # a chunk re-encoded by the host: ship it packed
                i, hs = item
                r0, r1, p0, p1 = spans[i]
                slot = device_slot(copy_packed)
                with torch.cuda.stream(copy_packed):
                    if trace is not None:
                        ev0 = torch.cuda.Event(enable_timing=True)
                        ev0.record(copy_packed)
                    n_vw = int(L.kmap_valid_words(p1 - p0))
                    d_packed[slot][:2 * n_vw].copy_(h_packed[hs][:2 * n_vw], non_blocking=True)
                    d_valid[slot][:n_vw].copy_(h_valid[hs][:n_vw], non_blocking=True)
                    d_strides[slot][:r1 - r0].copy_(h_strides[hs][:r1 - r0], non_blocking=True)
                    ev = torch.cuda.Event(enable_timing=trace is not None)
                    ev.record(copy_packed)
                if trace is not None:
                    copy_events.append(("packed", i, ev0, ev))
                staged_copy_done[hs] = ev
                staged_free[hs].set()
                mark("packed copy enqueued", i)
                to_count.put((i, slot, ev, False))
                continue
            if can_raw:                                     # the link has room: ship one chunk as it is
                i = take("back")
                if i is None:
                    continue
                r0, r1, p0, p1 = spans[i]
                hs = N_STAGED + n_raw % (max_raw + 1)
                n_raw += 1
                strides_of(i, h_strides[hs])
                slot = device_slot(copy)
                if d_u8[slot] is None:
                    d_u8[slot] = E.empty(max_pos, torch.uint8)
                with torch.cuda.stream(copy):
                    if trace is not None:
                        ev0 = torch.cuda.Event(enable_timing=True)
                        ev0.record(copy)
                    d_u8[slot][:p1 - p0].copy_(seq_t[p0:p1], non_blocking=True)
                    d_strides[slot][:r1 - r0].copy_(h_strides[hs][:r1 - r0], non_blocking=True)
                    ev = torch.cuda.Event(enable_timing=trace is not None)
                    ev.record(copy)
                if trace is not None:
                    copy_events.append(("raw", i, ev0, ev))
                raw_events.append(ev)
                mark("raw copy enqueued", i)
                to_count.put((i, slot, ev, True))
    finally:
        with lock:                                          # (on an error: no more chunks for the encoder)
            nxt["front"] = nxt["back"] + 1
        for e in staged_free:
            e.set()
        to_count.put(None)
        cnt_thread.join()
        if enc is not None:
            enc.join()
    if failed:
        raise failed[0]
    compute.synchronize()
    mark("all counted")
    global last_stream_stats
    h2d = sum((p1 - p0) if i >= n_chunks - n_raw else 12 * int(L.kmap_valid_words(p1 - p0)) for i, (_, _, p0, p1) in enumerate(spans))
    last_stream_stats = {"chunks": n_chunks, "raw_chunks": n_raw, "stream_ms": round(1e3 * (_time.perf_counter() - t_origin), 1),
                         "h2d_bytes": int(h2d + 4 * len(b))}
    if trace is not None:
        global last_stream_trace
        first = copy_events[0][2] if copy_events else None
        for kind, i, e0, e1 in copy_events:
            trace.append((round(first.elapsed_time(e0), 2), "copy-engine", f"{kind} copy {round(e0.elapsed_time(e1), 2)} ms", i))
        last_stream_trace = trace
    counted = len(released)
    assert counted == n_chunks
    return flat, totals, len(seq_np_arr)


def hamdist_matrix_rows(kh: np.ndarray, labels: np.ndarray, head_len: Sequence[int], kmer_len: int, rank: int, world: int):
    """Row-block partition of the sampled k-mer distance matrix (motif_discovery.py:759-808) for multi-GPU runs:
    returns (row0, row1, uint8 device tensor [(row1-row0), n]) of this rank's slab; no collective is involved."""
    from .motif_discovery import hamdist_matrix_u8
    n = len(kh)
    row0, row1 = row_range(n, rank, world)
    return row0, row1, hamdist_matrix_u8(kh, labels, head_len, kmer_len, row0, row1)


# ---- sharding (host logic; reads are the independent units of the counting path) --------------------------------------
def read_range(n_seq: int, rank: int, world: int) -> Tuple[int, int]:
    """contiguous range of reads [r0, r1) owned by `rank`; the ranges of all ranks tile [0, n_seq)"""
    if not (0 <= rank < world):
        raise KmapError(f"rank {rank} outside world of {world}")
    return n_seq * rank // world, n_seq * (rank + 1) // world


def shard_reads(seq_np_arr: np.ndarray, boarder_mat: np.ndarray, rank: int, world: int) -> Tuple[np.ndarray, np.ndarray]:
    """This rank's slice of input.bin / input.seqboarder.bin (kmer_count.py:326-347 layout): the bytes of its reads,
    separators included, and the border rows rebased to the slice.  Views, no copies of the sequence."""
    b = np.asarray(boarder_mat, dtype=np.int64).reshape(-1, 2)
    r0, r1 = read_range(len(b), rank, world)
    if r1 == r0:
        return seq_np_arr[:0], b[:0].copy()
    p0, p1 = int(b[r0, 0]), int(b[r1 - 1, 1]) + 1          # [first base of read r0, separator of read r1-1]
    return seq_np_arr[p0:p1], b[r0:r1] - p0


def row_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """row block of the distance matrix owned by `rank` (motif_discovery.py:759-808 has no cross-row dependency)"""
    return read_range(n, rank, world)
