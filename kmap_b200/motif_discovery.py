"""Drop-in for the counting path of the reference's `kmap/motif_discovery.py` (find_motif, ex_hamball, motif
occurrence, sampled k-mer distance matrix, the scan_motif / ex_hamball drivers and their file layouts).  Same
names, signatures and files; integer work runs in libkmap_b200 (sm_100a CUDA) on device-resident packed reads.

Reference line numbers cited below are in /root/reference/src/kmap/motif_discovery.py.  Plotting, KDE position
densities, co-occurrence networks and consensus alignment are outside this package's scope.
"""
from __future__ import annotations

import pickle
import warnings
from pathlib import Path
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import engine as E
from ._lib import KmapError, check, lib
ALL_K_MAX = 14       # largest k whose first count comes out of the all-k pass (SeqOnDevice.count_all through csrc/partition.cu)

from .kmer_count import (FileNameDict, MotifDef, cal_hamming_dist, cal_hamming_dist_head, cal_hamming_dist_tail,
                         dna2arr, gen_motif_def_dict, get_cnt_dtype, get_hash_dtype, get_revcom_hash_arr, hash2kmer,
                         init_motif_def_dict, kmer2hash, mask_ham_ball, reverse_complement, revcom_hash)


def write_lines(str_list: List, outfile):
    with open(outfile, "w+") as fh:
        for line in str_list:
            fh.write(line + "\n")


# ======================================================================================================================
# find_motif (:594-702)
# ======================================================================================================================
class _CountState:
    """Counts of the current (possibly masked) sequence: dense forward table on the device (k <= 15) or 64-bit merged
    lists from the sort path (k >= 16), + the reference's merged lists.  The host copies of the lists are fetched lazily:
    the first round needs them (k{k}.pkl, n_total_kmer), a later round only when numpy itself has to make the top-k
    selection (np.argpartition has to see exactly those arrays)."""

    def __init__(self, k, revcom, table=None, kh_dev=None, cnt_dev=None, wide=False):
        self.k, self.revcom, self.table, self.kh_dev, self.cnt_dev, self.wide = k, revcom, table, kh_dev, cnt_dev, wide
        self._kh = self._cnt = None

    @property
    def n(self) -> int:
        return int(self.kh_dev.numel()) if self._kh is None else len(self._kh)

    @property
    def kh(self) -> np.ndarray:
        if self._kh is None:
            self._kh = E.to_host(self.kh_dev, np.uint64 if self.wide else np.uint32).astype(get_hash_dtype(self.k), copy=False)
        return self._kh

    @property
    def cnt(self) -> np.ndarray:
        if self._cnt is None:
            self._cnt = E.to_host(self.cnt_dev, np.int64 if self.wide else np.int32).astype(get_cnt_dtype(self.k), copy=False)
        return self._cnt

    def kh_at(self, inds) -> np.ndarray:
        """the hashes at a few list positions (without bringing the whole list to the host)"""
        if self._kh is not None:
            return self._kh[inds]
        hd = np.uint64 if self.wide else np.uint32
        vals = [E.to_host(self.kh_dev[int(i):int(i) + 1], hd)[0] for i in inds]
        return np.array(vals, dtype=hd).astype(get_hash_dtype(self.k), copy=False)

    @classmethod
    def from_table(cls, table, k, revcom):
        kh_dev, cnt_dev = E.compact_merge(table, k, revcom)
        return cls(k, revcom, table, kh_dev, cnt_dev)

    @classmethod
    def from_sorted(cls, dev, k, revcom, dedup):
        kh_dev, cnt_dev = dev.count_sorted(k, dedup)
        if revcom:
            kh_dev, cnt_dev = E.merge_revcom_sorted(kh_dev, cnt_dev, k)
        return cls(k, revcom, None, kh_dev, cnt_dev, wide=True)

    @classmethod
    def from_lists(cls, kh, cnt, k, revcom):
        wide = k >= 16
        st = cls(k, revcom, wide=wide)
        st._kh, st._cnt = np.asarray(kh), np.asarray(cnt)
        st.kh_dev = E.to_device(st._kh.astype(np.uint64 if wide else np.uint32, copy=False))
        st.cnt_dev = E.to_device(st._cnt.astype(np.int64 if wide else np.int32, copy=False))
        return st

    def ball_sums(self, cand, d):
        if self.table is not None:
            return E.hamball_sums(self.table, self.k, cand, d, self.revcom)
        if self.wide:
            return E.hamball_sums_list64(self.kh_dev, self.cnt_dev, self.k, cand, d, self.revcom)
        return E.hamball_sums_list(self.kh_dev, self.cnt_dev, self.k, cand, d, self.revcom)


def _top_k_indices(state: "_CountState", top_k: int):
    """(indices, unambiguous): the top_k entries of state.cnt that `np.argpartition(cnt, -top_k)[-top_k:]` selects (:657).
    On a long list the selection runs on the device (kmap_topk_candidates_*; numpy's introselect over 1.7e7 counts is 0.2 s
    per trial and was 58 % of scan_motif at 1e6 reads).  The device result is used only when it cannot differ from
    numpy's as a SET: the top_k-th count must be strictly larger than the next one (which of several equal counts
    introselect keeps is not reproducible); the ORDER numpy would return only matters when two candidates tie on their
    ball sums, which the caller checks (unambiguous=True asks for that check)."""
    n = state.n if hasattr(state, "n") else len(state.cnt)
    if state.cnt_dev is not None and n >= (1 << 16) and top_k + 1 <= 8 and int(state.cnt_dev.numel()) == n:
        val, idx = E.topk_candidates(state.cnt_dev, top_k + 1)
        if len(val) == top_k + 1 and val[top_k - 1] > val[top_k]:
            return idx[:top_k].astype(np.intp), True
    return np.array(np.argpartition(state.cnt, -top_k)[-top_k:]), False


def find_motif_on_device(dev: E.SeqOnDevice, kmer_len: int, max_ham_dist, p_unif, ratio_mu, ratio_std, ratio_cutoff,
                         top_k=5, n_trial=10, merge_revcom_mode=True, rep_mode=False, first_lists=None,
                         first_table: Optional[torch.Tensor] = None, debug=False, sorted_path: Optional[bool] = None,
                         table_buffer: Optional[torch.Tensor] = None, ctx=None, want_first: bool = True):
    """Core of find_motif on a device-resident sequence (mutates dev.valid).  Returns (result dict, (uniq_kh, uniq_cnt)
    of the first round, or None with want_first=False: the lists then stay on the device).  `first_lists` plays the role
    of a pre-existing k{k}.pkl (:621-624); `first_table` lets a
    caller that counted every k in one pass (SeqOnDevice.count_all) hand in the forward table -- already merged over the
    ranks.  sorted_path: count by sorting 64-bit keys (csrc/sorted.cu) instead of a dense table; None = only where there
    is no dense table (k >= 16).  ctx (api.DistContext): `dev` holds this rank's shard of the reads; every count is merged
    over the ranks (dense tables: integer all-reduce; sorted lists: all-gather + merge), the selection, the ball sums and
    the decisions are then replicated, and every rank masks its own reads."""
    from scipy.stats import norm
    k = kmer_len
    use_sorted = (k >= 16) if sorted_path is None else bool(sorted_path)
    sharded = ctx is not None and ctx.world > 1

    def count_sorted_state(dedup):
        if not sharded:
            return _CountState.from_sorted(dev, k, merge_revcom_mode, dedup=dedup)
        from .api import merge_sorted_counts_over_ranks
        kh_dev, cnt_dev = merge_sorted_counts_over_ranks(*dev.count_sorted(k, dedup), ctx, 2 * k)
        if merge_revcom_mode:
            kh_dev, cnt_dev = E.merge_revcom_sorted(kh_dev, cnt_dev, k)
        return _CountState(k, merge_revcom_mode, None, kh_dev, cnt_dev, wide=True)

    if first_lists is not None:
        state = _CountState.from_lists(first_lists[0], first_lists[1], k, merge_revcom_mode)
    elif use_sorted:
        state = count_sorted_state(dedup=not rep_mode)
    else:
        if first_table is not None:
            table = first_table
        else:       # (table_buffer: a caller that walks several k re-uses one allocation of at least 4^k cells)
            buf = table_buffer[:1 << (2 * k)] if table_buffer is not None and table_buffer.numel() >= (1 << (2 * k)) else None
            table = dev.count(k, dedup=not rep_mode, table=buf)
            if sharded:
                ctx.allreduce(table)
        state = _CountState.from_table(table, k, merge_revcom_mode)
    first = (state.kh, state.cnt) if want_first else None
    # exact total, summed where the counts are
    n_total_kmer = E.sum_counts(state.cnt_dev) if state.cnt_dev is not None else int(np.sum(state.cnt, dtype=np.int64))

    found = {}
    for i_trial in range(n_trial):
        if top_k > state.n:
            break
        top_k_inds, unambiguous = _top_k_indices(state, top_k)
        if len(top_k_inds) == 0:
            break
        cand = state.kh_at(top_k_inds)
        hamball_cnt_arr = np.zeros(top_k)
        hamball_cnt_arr[:] = state.ball_sums(cand, max_ham_dist)
        if unambiguous and np.count_nonzero(hamball_cnt_arr == hamball_cnt_arr.max()) > 1:
            # equal ball sums: the winner is the first one in numpy's own order of the top-k (np.argmax below)
            top_k_inds = np.array(np.argpartition(state.cnt, -top_k)[-top_k:])
            cand = state.kh_at(top_k_inds)
            hamball_cnt_arr[:] = state.ball_sums(cand, max_ham_dist)
        if debug:
            print(f"{i_trial= }")
        best = np.argmax(hamball_cnt_arr)
        consensus_kh = cand[best]
        hamball_proportion = (hamball_cnt_arr[best] + 0.0) / n_total_kmer
        hamball_ratio = hamball_proportion / p_unif
        if not hamball_ratio > ratio_cutoff:
            break
        found[consensus_kh] = (hamball_proportion, hamball_ratio,
                               norm.logsf(hamball_ratio, loc=ratio_mu, scale=ratio_std) / np.log(10))
        cons = [int(consensus_kh)]
        if merge_revcom_mode:
            cons.append(int(revcom_hash(consensus_kh, k)))
        dev.mask(k, cons, [max_ham_dist] * len(cons))
        if use_sorted:                                                # recounts are never de-duplicated (:695-696)
            state = count_sorted_state(dedup=False)
        else:
            if state.table is None:         # the first round came from a k{k}.pkl: no table yet
                buf = table_buffer[:1 << (2 * k)] if table_buffer is not None and table_buffer.numel() >= (1 << (2 * k)) else None
                table = dev.count(k, dedup=False, table=buf)
            else:
                table = dev.count(k, dedup=False, table=state.table)
            if sharded:
                ctx.allreduce(table)
            state = _CountState.from_table(table, k, merge_revcom_mode)
    return found, first


def find_motif(seq_np_arr, kmer_len: int, max_ham_dist, p_unif, ratio_mu, ratio_std, ratio_cutoff, top_k=5, n_trial=10,
               merge_revcom_mode=True, rep_mode=False, save_kmer_cnt_flag=True, kmer_cnt_pkl_file: Path = None,
               boarder_pkl_file: Path = None, debug=False) -> dict:
    """:594-702, same signature and side effects: mutates seq_np_arr (masking), reads/writes k{k}.pkl."""
    if boarder_pkl_file:
        assert Path(boarder_pkl_file).exists()
    first_lists = None
    boarder_mat = None
    if save_kmer_cnt_flag and kmer_cnt_pkl_file and Path(kmer_cnt_pkl_file).exists():
        with open(Path(kmer_cnt_pkl_file), "rb") as fh:
            kmer_len_from_pkl_file, uniq_kh_arr, uniq_kh_cnt_arr = pickle.load(fh)
            assert kmer_len == kmer_len_from_pkl_file
        first_lists = (uniq_kh_arr, uniq_kh_cnt_arr)
    else:
        with open(boarder_pkl_file, "rb") as fh:
            boarder_mat = pickle.load(fh)
    dev = E.SeqOnDevice.from_numpy(seq_np_arr, boarder_mat, keep_u8=True)
    found, first = find_motif_on_device(dev, kmer_len, max_ham_dist, p_unif, ratio_mu, ratio_std, ratio_cutoff, top_k,
                                        n_trial, merge_revcom_mode, rep_mode, first_lists, None, debug)
    if save_kmer_cnt_flag and kmer_cnt_pkl_file and not Path(kmer_cnt_pkl_file).exists():
        with open(kmer_cnt_pkl_file, "wb") as fh:
            pickle.dump([kmer_len, first[0], first[1]], fh)
    if found:
        dev.masked_seq_to_numpy(seq_np_arr)
    return found


# ======================================================================================================================
# Hamming ball extraction + count matrix (:489-530, 924-986)
# ======================================================================================================================
def _hamball_extract(uniq_kh_arr, uniq_kh_cnt_arr, conseq_kh: int, kmer_len: int, max_ham_dist: int, revcom_mode: bool,
                     want_list=True):
    L = lib()
    n = len(uniq_kh_arr)
    import ctypes
    if kmer_len >= 16:                      # uint64 hashes, int64 counts (kmer_count.py:351-365)
        kh_d = E.to_device(np.asarray(uniq_kh_arr).astype(np.uint64, copy=False))
        cnt_d = E.to_device(np.asarray(uniq_kh_cnt_arr).astype(np.int64, copy=False))
        scratch = E._scratch(L.kmap_list_scratch_words(n))
        cnt_mat = E.empty(4 * kmer_len, torch.int64)
        n_out = ctypes.c_int64(0)
        stream = torch.cuda.current_stream().cuda_stream
        m = n if want_list else 0
        out_kh, out_cnt = E.empty(m, torch.int64), E.empty(m, torch.int64)
        check(L.kmap_hamball_extract_u64(kh_d.data_ptr(), cnt_d.data_ptr(), n, kmer_len, int(conseq_kh), int(max_ham_dist),
                                         int(revcom_mode), scratch.data_ptr(), out_kh.data_ptr() if m else None,
                                         out_cnt.data_ptr() if m else None, m, ctypes.byref(n_out), cnt_mat.data_ptr(), stream),
              "kmap_hamball_extract_u64")
        mat = cnt_mat.cpu().numpy().reshape(4, kmer_len).astype(int)
        if not want_list:
            return None, None, mat
        return E.to_host(out_kh[:n_out.value], np.uint64), E.to_host(out_cnt[:n_out.value], np.int64), mat
    kh_d = E.to_device(np.asarray(uniq_kh_arr).astype(np.uint32, copy=False))
    cnt_d = E.to_device(np.asarray(uniq_kh_cnt_arr).astype(np.int32, copy=False))
    scratch = E._scratch(L.kmap_list_scratch_words(n))
    cnt_mat = E.empty(4 * kmer_len, torch.int64)
    n_out = ctypes.c_int64(0)
    stream = torch.cuda.current_stream().cuda_stream
    # size query = a run without list output (the matrix is complete after it)
    check(L.kmap_hamball_extract(kh_d.data_ptr(), cnt_d.data_ptr(), n, kmer_len, int(conseq_kh), int(max_ham_dist),
                                 int(revcom_mode), scratch.data_ptr(), None, None, 0, ctypes.byref(n_out),
                                 cnt_mat.data_ptr(), stream), "kmap_hamball_extract")
    mat = cnt_mat.cpu().numpy().reshape(4, kmer_len).astype(int)
    if not want_list:
        return None, None, mat
    m = n_out.value
    out_kh, out_cnt = E.empty(m, torch.int32), E.empty(m, torch.int32)
    if m:
        check(L.kmap_hamball_extract(kh_d.data_ptr(), cnt_d.data_ptr(), n, kmer_len, int(conseq_kh), int(max_ham_dist),
                                     int(revcom_mode), scratch.data_ptr(), out_kh.data_ptr(), out_cnt.data_ptr(), m,
                                     ctypes.byref(n_out), cnt_mat.data_ptr(), stream), "kmap_hamball_extract")
    return E.to_host(out_kh, np.uint32), E.to_host(out_cnt, np.int32), mat


def ex_hamball_kh_arr(res_dir: str, conseq: str, max_ham_dist: int = -1, motif_def_file: str = None, revcom_mode=True):
    """:924-975"""
    conseq = conseq.upper()
    assert all([e in ("A", "C", "G", "T") for e in conseq])
    kmer_len = len(conseq)
    conseq_kh = kmer2hash(conseq)
    rc_conseq_kh = revcom_hash(conseq_kh, kmer_len)
    if revcom_mode:
        assert conseq_kh <= rc_conseq_kh
    assert Path(motif_def_file).exists()
    assert Path(res_dir).exists()
    kmer_cnt_file = Path(res_dir) / FileNameDict["kmer_count_dir"] / f"k{kmer_len}.pkl"
    with open(kmer_cnt_file, "rb") as fh:
        res_list = pickle.load(fh)
    assert res_list[0] == kmer_len
    uniq_kh_arr, uniq_kh_cnt_arr = res_list[1], res_list[2]
    if max_ham_dist == -1:
        max_ham_dist = init_motif_def_dict(motif_def_file)[kmer_len].max_ham_dist
    kh, cnt, _ = _hamball_extract(uniq_kh_arr, uniq_kh_cnt_arr, int(conseq_kh), kmer_len, max_ham_dist, revcom_mode)
    return kh.astype(uniq_kh_arr.dtype, copy=False), cnt.astype(uniq_kh_cnt_arr.dtype, copy=False)


def cal_cnt_mat(uniq_kh_arr, uniq_kh_cnt_arr, kmer_len):
    """:978-986: int64[4, k], row = base code, column = position.  Runs the ball kernel with the ball = everything."""
    if len(uniq_kh_arr) == 0:
        return np.zeros((4, kmer_len), dtype=int)
    _, _, mat = _hamball_extract(uniq_kh_arr, uniq_kh_cnt_arr, 0, kmer_len, kmer_len, False, want_list=False)
    return mat


def _ex_hamball(res_dir: str, conseq: str, return_type: str, output_file: str, max_ham_dist: int = -1):
    """:489-530"""
    import tomllib
    config_file_path = Path(res_dir) / FileNameDict["config_file"]
    assert config_file_path.exists()
    with open(config_file_path, "rb") as fh:
        config_dict = tomllib.load(fh)
    assert return_type in ("hash", "kmer", "matrix")
    motif_def_file_path = Path(res_dir) / FileNameDict["motif_def_file"]
    revcom_mode = config_dict["kmer_count"]["revcom_mode"]
    uniq_kh_arr, uniq_kh_cnt_arr = ex_hamball_kh_arr(res_dir, conseq, max_ham_dist, motif_def_file_path, revcom_mode)
    kmer_len = len(conseq)
    with open(output_file, "w+") as fh:
        if return_type == "hash":
            for kh, cnt in zip(uniq_kh_arr, uniq_kh_cnt_arr):
                fh.write(f"{kh},{cnt}\n")
        elif return_type == "kmer":
            for kh, cnt in zip(uniq_kh_arr, uniq_kh_cnt_arr):
                fh.write(f"{hash2kmer(kh, kmer_len)},{cnt}\n")
        else:
            np.savetxt(fh, cal_cnt_mat(uniq_kh_arr, uniq_kh_cnt_arr, kmer_len), delimiter=",", fmt="%d")
    print(f"Extract Hamming ball [type={return_type}] save in {output_file}.")


# ======================================================================================================================
# motif occurrence (:1345-1477)
# ======================================================================================================================
def motif_occurence_table(dev: E.SeqOnDevice, conseq_list: Sequence[str], motif_def_dict: dict, revcom_mode=True,
                          keep_device: bool = False):
    """All reads x all consensus sequences in one device pass per consensus.  Returns a list (per consensus) of
    (min_dist[n_seq], offsets[n_seq+1], positions[total]) on the host [, the same list as device tensors]."""
    out, out_dev = [], []
    for conseq in conseq_list:
        k = len(conseq)
        t = E.occurrence_scan_device(dev, k, int(kmer2hash(conseq)), motif_def_dict[k].max_ham_dist, revcom_mode)
        out.append(tuple(E.to_host(x) for x in t))
        if keep_device:
            out_dev.append(t)
    return (out, out_dev) if keep_device else out


def _cells_for_read(per_conseq, r):
    cells, flag = [], False
    for min_dist, offsets, pos in per_conseq:
        lo, hi = offsets[r], offsets[r + 1]
        if hi == lo:
            cells.append("")
            continue
        locs = pos[lo:hi]
        if len(locs) > 20:                                   # :1467-1469 (the reference's unseeded random pick)
            locs = np.sort(locs[np.random.choice(len(locs), 20, replace=False)])
        flag = True
        cells.append(",".join(map(str, locs)))
    return flag, ";".join(cells)


def get_motif_occurence(seq_np_arr: np.ndarray, conseq_list: List[str], motif_def_dict: dict, revcom_mode=True):
    """:1422-1477 for ONE read given without separator.  Returns (motif_flag, 'p,p;..;..')."""
    arr = np.concatenate([np.asarray(seq_np_arr, dtype=np.uint8), np.array([255], dtype=np.uint8)])
    dev = E.SeqOnDevice.from_numpy(arr, np.array([[0, len(arr) - 1]], dtype=np.int64))
    return _cells_for_read(motif_occurence_table(dev, conseq_list, motif_def_dict, revcom_mode), 0)


def motif_occurence_lines(dev: E.SeqOnDevice, borders: np.ndarray, conseq_list, motif_def_dict, revcom_mode=True) -> List[str]:
    """lines of a *.motif_occurence.csv (:1409-1418): header, then one row per read that has at least one hit"""
    lines = ["seq_ind;" + ";".join(f"motif_{i}_{conseq_list[i]}" for i in range(len(conseq_list))) + ";seq_len"]
    per_conseq = motif_occurence_table(dev, conseq_list, motif_def_dict, revcom_mode)
    if not per_conseq:
        return lines
    has_hit = np.zeros(dev.n_seq, dtype=bool)
    for min_dist, offsets, pos in per_conseq:
        has_hit |= np.diff(offsets) > 0
    lens = borders[:, 1] - borders[:, 0]
    for r in np.flatnonzero(has_hit):
        flag, cells = _cells_for_read(per_conseq, r)
        lines.append(f"{r};{cells};{lens[r]}")
    return lines


def write_motif_occurence_file(per_conseq, borders: np.ndarray, conseq_list, output_file,
                               picked_rows: Optional[dict] = None) -> List[Tuple[int, int]]:
    """The *.motif_occurence.csv of :1409-1418 from the scan results, rows formatted natively (kmap_write_occurrence_rows;
    ~5 us per row in Python is what scan_motif would otherwise wait for).  A read with more than 20 positions in some cell
    needs the reference's random pick (:1467-1469, numpy's global RNG): those rows go through the Python formatter, in read
    order, so the RNG stream is consumed exactly as the reference consumes it.  Returns per consensus
    (reads with the motif, listed positions) -- what get_motif_seq_num parses back out of the file.
    picked_rows (dict): filled with {read: [cell string per consensus]} for the rows that went through the random pick, for
    the consumers that work from the scan results instead of the file (co_occurrence_from_scan, pos_density_from_scan)."""
    import ctypes
    L = lib()
    with open(output_file, "w+") as fh:
        fh.write("seq_ind;" + ";".join(f"motif_{i}_{conseq_list[i]}" for i in range(len(conseq_list))) + ";seq_len\n")
    m = len(per_conseq)
    if m == 0:
        return []
    n_seq = len(borders)
    lens = np.ascontiguousarray(borders[:, 1] - borders[:, 0], dtype=np.int64)
    offs = [np.ascontiguousarray(o, dtype=np.int64) for _, o, _ in per_conseq]
    poss = [np.ascontiguousarray(p_, dtype=np.int32) for _, _, p_ in per_conseq]
    counts = [np.diff(o) for o in offs]
    over = np.zeros(n_seq, dtype=bool)
    for c in counts:
        over |= c > 20
    off_ptrs = (ctypes.c_void_p * m)(*[o.ctypes.data for o in offs])
    pos_ptrs = (ctypes.c_void_p * m)(*[p_.ctypes.data if len(p_) else None for p_ in poss])
    path = str(output_file).encode()

    def native_rows(r0, r1):
        if r1 > r0:
            rc = L.kmap_write_occurrence_rows(path, 1, m, off_ptrs, pos_ptrs, lens.ctypes.data, r0, r1)
            if rc < 0:
                check(int(rc), "kmap_write_occurrence_rows")
    r = 0
    for f in np.flatnonzero(over):
        native_rows(r, int(f))
        _, cells = _cells_for_read(per_conseq, int(f))
        if picked_rows is not None:
            picked_rows[int(f)] = cells.split(";")
        with open(output_file, "a") as fh:
            fh.write(f"{int(f)};{cells};{lens[f]}\n")
        r = int(f) + 1
    native_rows(r, n_seq)
    return [(int(np.count_nonzero(c)), int(np.minimum(c, 20).sum())) for c in counts]


def gen_motif_occurence_file(conseq_list: List[str], motif_def_dict: dict, input_fasta_file: Path, output_file: Path,
                             revcom_mode=True, _dev_cache=None, _ctx=None, _return_scan=False):
    """:1396-1419.  The reference re-parses the FASTA file per call; the encoded arrays are identical to input.bin
    (same upper-casing and code table), so a cached device copy may be passed by the driver.  Returns per consensus
    (reads with the motif, listed positions) [, with _return_scan a dict for the consumers of the file: `scan` = the scan
    results per consensus (min_dist, offsets, positions) on the host, `lens` = read lengths, `picked_rows` = the rows
    formatted with the random pick of 20 positions, `cooc` = the integer parts of the co-occurrence step computed on the
    device from the scan results (engine.co_occurrence_scan), read indices in whole-file numbering].
    _ctx (api.DistContext, world > 1): `_dev_cache` holds this rank's reads and the border matrix of the whole file; the
    per-read results are gathered in rank order and rank 0 writes the file (the other ranks return ([], None))."""
    assert Path(input_fasta_file).exists()
    if _dev_cache is not None:
        dev, borders = _dev_cache
    else:
        dev = E.SeqOnDevice.from_fasta(input_fasta_file)
        borders = E.to_host(dev.borders, np.int64).reshape(-1, 2)
    per_conseq, per_conseq_dev = motif_occurence_table(dev, conseq_list, motif_def_dict, revcom_mode, keep_device=True)
    world = _ctx.world if _ctx is not None else 1
    cooc = None
    if _return_scan and 1 <= len(conseq_list) <= E.COOC_MAX_MOTIFS:
        cooc = [E.co_occurrence_scan(per_conseq_dev)]
    del per_conseq_dev
    if world > 1:
        from .api import concat_occurrence_shards
        parts = _ctx.gather((per_conseq, cooc))
        if not _ctx.is_root:
            return ([], None) if _return_scan else []
        per_conseq = [concat_occurrence_shards([parts[r][0][j] for r in range(world)]) for j in range(len(conseq_list))]
        if cooc is not None:                      # shard-local read indices -> whole-file numbering, in rank order
            cooc, base = [], 0
            for r in range(world):
                counts, pairs, over = parts[r][1][0]
                cooc.append((counts, {key: (reads + base, d2) for key, (reads, d2) in pairs.items()}, over + base))
                base += len(parts[r][0][0][0]) if parts[r][0] else 0
    picked_rows = {}
    stats = write_motif_occurence_file(per_conseq, borders, conseq_list, output_file, picked_rows)
    if not _return_scan:
        return stats
    lens = np.ascontiguousarray(borders[:, 1] - borders[:, 0], dtype=np.int64)
    return stats, {"scan": per_conseq, "lens": lens, "picked_rows": picked_rows, "cooc": cooc}


def get_motif_seq_num(occurence_file_path: Path, motif_index: int) -> Tuple[int, int]:
    """:1345-1393"""
    lines_with_motif = total_occurrences = 0
    with open(occurence_file_path, "r") as fh:
        next(fh)
        for row in fh:
            cell = row.rstrip("\n").split(";")[motif_index + 1].strip()
            if cell == "":
                continue
            lines_with_motif += 1
            total_occurrences += len(cell.split(","))
    return lines_with_motif, total_occurrences


# ======================================================================================================================
# consumers of final.motif_occurence.csv that write DATA files (:1189-1343): co-occurrence matrices / distances and the
# motif position densities.  Host functions like the reference's (they parse the CSV this path writes); the plots the
# reference draws from the same numbers stay in the reference package.
# ======================================================================================================================
def _occurrence_rows(occurence_file_path: Path, n_cols: Optional[int] = None):
    with open(occurence_file_path, "r", newline="") as fh:
        header = next(fh).rstrip("\r\n").split(";")
        if n_cols is not None:
            assert len(header) == n_cols
        for line in fh:
            yield line.rstrip("\r\n").split(";")


def get_motif_co_occurence_mat(occurence_file_path: Path, n_conseq: int):
    """:1189-1254.  (co-occurrence counts int[n, n] with the per-motif read counts on the diagonal, median |distance|
    matrix float[n, n] (1e6 where two motifs never share a read), {(i, j): [median position of j - median position of i
    per shared read, in file order]})."""
    assert n_conseq > 0
    res_mat = np.zeros((n_conseq, n_conseq), dtype=int)
    dist_mat = np.zeros((n_conseq, n_conseq), dtype=float)
    individual_counts = np.zeros(n_conseq, dtype=int)
    dist_dict = {(i, j): [] for i in range(n_conseq) for j in range(i + 1, n_conseq)}
    for row in _occurrence_rows(occurence_file_path, n_conseq + 2):
        motif_inds = [i for i, e in enumerate(row[1:-1]) if e.strip() != ""]
        for i in motif_inds:
            individual_counts[i] += 1
        if len(motif_inds) <= 1:
            continue
        med = {}
        for i in motif_inds:
            v = sorted(int(x) for x in row[i + 1].split(","))
            h = len(v) // 2
            med[i] = float(v[h]) if len(v) % 2 else (v[h - 1] + v[h]) / 2.0          # np.median of the positions
        for a in range(len(motif_inds)):
            for b in range(a + 1, len(motif_inds)):
                ii, jj = motif_inds[a], motif_inds[b]
                res_mat[ii, jj] += 1
                res_mat[jj, ii] += 1
                dist_dict[(ii, jj)].append(med[jj] - med[ii])
    np.fill_diagonal(res_mat, individual_counts)
    for i in range(n_conseq):
        for j in range(i + 1, n_conseq):
            dist_mat[i, j] = dist_mat[j, i] = 1e6 if len(dist_dict[(i, j)]) == 0 else np.median(np.abs(dist_dict[(i, j)]))
    return res_mat, dist_mat, dist_dict


def write_co_occurence_dist_arr(output_file: Path, dist_dict, conseq_list: List[str]):
    """:1147-1163"""
    names = [f"m{i}_{s}_{reverse_complement(s)}" for i, s in enumerate(conseq_list)]
    with open(output_file, "w") as fh:
        for i, j in dist_dict:
            vals = dist_dict[(i, j)]
            if len(vals) == 0:
                continue
            fh.write(names[i] + "-" + names[j] + "\n")
            fh.write("\t".join(f"{n:.2f}" for n in vals) + "\n")


def write_co_occurence_mat(output_file: Path, dist_mat: np.ndarray, conseq_list: List[str]):
    """:1166-1187"""
    assert len(conseq_list) == len(dist_mat)
    rc_names = [f"m{i}_{reverse_complement(s)}" for i, s in enumerate(conseq_list)]
    names = [f"m{i}_{s}" for i, s in enumerate(conseq_list)]
    with open(output_file, "w") as fh:
        fh.write("\t".join(["RC"] + names) + "\n")
        for i, arr in enumerate(dist_mat):
            arr = np.around(arr, decimals=2)
            fh.write(rc_names[i] + "\t" + "\t".join([str(x) for x in arr]) + "\n")


def _density_add_batch(density, counts, centres, x_arr, x_step):
    """density += for every read of the batch, in order, the mean of the normal densities (sd = x_step) centred at
    centres[r, :counts[r]]: the arithmetic of :1318-1326 (`sum(norm(xi, scale=x_step).pdf(x_arr) for xi in ...) / len(...)`
    accumulated read by read; scipy's pdf is exp(-y^2 / 2) / sqrt(2 pi) / scale with y = (x - loc) / scale) in the same
    order; only the exponentials are batched."""
    norm_c = np.sqrt(2 * np.pi)
    acc = None
    for j in range(int(counts.max())):              # left to right inside a read, like sum() over the generator
        y = (x_arr[None, :] - centres[:, j:j + 1]) / x_step
        pdf = np.exp(-y ** 2 / 2.0) / norm_c / x_step
        if acc is None:
            acc = 0 + pdf
        else:
            acc = np.where((counts > j)[:, None], acc + pdf, acc)
    acc = acc / counts[:, None]
    for row in acc:                                 # read by read, like `density +=`
        density += row


def get_motif_pos_density(occurence_file_path: Path, motif_index: int, kmer_len: int, x_step=0.01, x_arr=None):
    """:1256-1343.  (reads with the motif, listed positions, density over x_arr): every read adds the mean of normal
    densities (sd = x_step) centred at its relative motif positions loc / (seq_len - k + 1)."""
    if x_arr is None:
        x_arr = np.arange(0, 1, x_step)
    x_arr = np.asarray(x_arr)
    density = np.zeros_like(x_arr)
    lines_with_motif = total_occurrences = 0
    batch_pos = []

    def flush():
        if not batch_pos:
            return
        counts = np.array([len(p_) for p_ in batch_pos])
        centres = np.zeros((len(batch_pos), int(counts.max())))
        for r, p_ in enumerate(batch_pos):
            centres[r, :len(p_)] = p_
        _density_add_batch(density, counts, centres, x_arr, x_step)
        batch_pos.clear()

    for row in _occurrence_rows(occurence_file_path):
        cell = row[motif_index + 1].strip()
        if cell == "":
            continue
        seq_len = float(row[-1].strip())
        locs = [int(n) for n in cell.split(",")]
        batch_pos.append([(loc + 0.0) / (seq_len - kmer_len + 1) for loc in locs])
        lines_with_motif += 1
        total_occurrences += len(locs)
        if len(batch_pos) >= 4096:
            flush()
    flush()
    return lines_with_motif, total_occurrences, density


def pos_density_from_scan(scan_info: dict, motif_index: int, kmer_len: int, x_step=0.01, x_arr=None):
    """get_motif_pos_density (:1256-1343) from the results of the occurrence scan (the dict gen_motif_occurence_file returns
    with _return_scan) instead of final.motif_occurence.csv parsed back: same three results, same float arithmetic in the
    same order; the rows with more than 20 positions in a cell use the positions the file writer picked."""
    if x_arr is None:
        x_arr = np.arange(0, 1, x_step)
    x_arr = np.asarray(x_arr)
    density = np.zeros_like(x_arr)
    _, offsets, pos = scan_info["scan"][motif_index]
    lens, picked = scan_info["lens"], scan_info["picked_rows"]
    cnt_all = np.diff(offsets)
    reads = np.flatnonzero(cnt_all > 0)
    lines_with_motif, total_occurrences = len(reads), 0
    picked_reads = np.array(sorted(picked), dtype=np.int64)
    for b0 in range(0, len(reads), 4096):
        R = reads[b0:b0 + 4096]
        counts = cnt_all[R].copy()
        start = offsets[R]
        denom = (lens[R] - kmer_len + 1).astype(np.float64)           # float(seq_len) - kmer_len + 1
        swap = R[np.isin(R, picked_reads)] if len(picked_reads) else ()
        for r in swap:                                                # random pick made while the file was written
            counts[np.searchsorted(R, r)] = len(picked[int(r)][motif_index].split(","))
        centres = np.zeros((len(R), int(counts.max())))
        for j in range(centres.shape[1]):
            sel = counts > j
            centres[sel, j] = (pos[start[sel] + j] + 0.0) / denom[sel]
        for r in swap:
            at = int(np.searchsorted(R, r))
            locs = [int(n) for n in picked[int(r)][motif_index].split(",")]
            centres[at, :] = 0.0
            centres[at, :len(locs)] = [(loc + 0.0) / denom[at] for loc in locs]
        total_occurrences += int(counts.sum())
        _density_add_batch(density, counts, centres, x_arr, x_step)
    return lines_with_motif, total_occurrences, density


def co_occurrence_from_scan(scan_info: dict, n_conseq: int):
    """get_motif_co_occurence_mat (:1189-1254) from the results of the occurrence scan: the per-motif and per-pair read
    counts, and for every pair the per-read difference of the median positions in read order, come from the device
    (csrc/consumers.cu via engine.co_occurrence_scan); the reads whose row went through the random pick of 20 positions
    are added from the picked rows.  Same three results as get_motif_co_occurence_mat on the file."""
    assert n_conseq > 0
    res_mat = np.zeros((n_conseq, n_conseq), dtype=int)
    dist_mat = np.zeros((n_conseq, n_conseq), dtype=float)
    individual_counts = np.zeros(n_conseq, dtype=int)
    keys = [(i, j) for i in range(n_conseq) for j in range(i + 1, n_conseq)]
    reads_of, vals_of = {key: [] for key in keys}, {key: [] for key in keys}
    n_over = 0
    for counts, pairs, over in scan_info["cooc"]:
        individual_counts += np.diag(counts).astype(int)
        res_mat += np.triu(counts, 1).astype(int)
        n_over += len(over)
        for key in keys:
            reads_of[key].append(pairs[key][0])
            vals_of[key].append(pairs[key][1] / 2.0)
    picked = scan_info["picked_rows"]
    assert n_over == len(picked)
    extra = {key: ([], []) for key in keys}
    for r in sorted(picked):
        cells = picked[r]
        motif_inds = [i for i, e in enumerate(cells) if e.strip() != ""]
        for i in motif_inds:
            individual_counts[i] += 1
        med = {}
        for i in motif_inds:
            v = sorted(int(x) for x in cells[i].split(","))
            h = len(v) // 2
            med[i] = float(v[h]) if len(v) % 2 else (v[h - 1] + v[h]) / 2.0
        for a_ in range(len(motif_inds)):
            for b_ in range(a_ + 1, len(motif_inds)):
                ii, jj = motif_inds[a_], motif_inds[b_]
                res_mat[ii, jj] += 1
                extra[(ii, jj)][0].append(r)
                extra[(ii, jj)][1].append(med[jj] - med[ii])
    res_mat = res_mat + res_mat.T
    np.fill_diagonal(res_mat, individual_counts)
    dist_dict = {}
    for key in keys:
        reads = np.concatenate(reads_of[key] + [np.asarray(extra[key][0], dtype=np.int64)])
        vals = np.concatenate(vals_of[key] + [np.asarray(extra[key][1], dtype=np.float64)])
        if len(extra[key][0]):
            vals = vals[np.argsort(reads, kind="stable")]
        dist_dict[key] = vals.tolist()
        i, j = key
        dist_mat[i, j] = dist_mat[j, i] = 1e6 if len(vals) == 0 else np.median(np.abs(dist_dict[key]))
    return res_mat, dist_mat, dist_dict


# ======================================================================================================================
# sampled k-mer distance matrix (:705-808)
# ======================================================================================================================
def _convert_to_block_mat(uniq_dist_mat: np.ndarray, block_size_arr: np.ndarray) -> np.ndarray:
    """:705-730"""
    assert np.issubdtype(block_size_arr.dtype, np.integer)
    assert np.all(block_size_arr > 0)
    return np.repeat(np.repeat(uniq_dist_mat, block_size_arr, axis=0), block_size_arr, axis=1)


def _convert_to_block_arr(arr: np.ndarray, block_size_arr: np.ndarray) -> np.ndarray:
    """:733-757"""
    assert np.issubdtype(block_size_arr.dtype, np.integer)
    assert np.all(block_size_arr > 0)
    assert len(arr) == len(block_size_arr)
    return np.repeat(arr, block_size_arr)


HAMDIST_MMA_MIN_N = 2048            # below this the matrix is a few hundred tiles: the XOR/popcount kernel, no operand preparation
HAMDIST_MMA_MAX_SLOTS = 32          # K = 128 int8 = 32 slots of one base: k bases + the tail bases of every shorter consensus


def hamdist_formulation(n: int, kmer_len: int, head_len: Sequence[int]) -> str:
    """which kernel hamdist_matrix_u8 uses.  Measured on B200 (profiles/r02_ncu_hamdist_*.txt, 1e10 pairs): the int8 one-hot
    GEMM on the tcgen05 tensor cores writes the matrix in 2.12 ms with or without head overrides (they are extra K columns),
    the XOR/popcount kernel in 2.63 / 2.89 ms (bound by the POPC pipe: 16 per clock and SM) -- so the tensor cores are kept
    wherever the formulation applies: 32-bit hashes (k <= 16), override columns that fit into K = 128, a matrix large enough
    to fill the machine.  Everything else (uint64 hashes, many short consensus sequences, small samples) stays on popcount."""
    slots = kmer_len + sum(kmer_len - max(int(h), 0) for h in head_len if int(h) < kmer_len)
    if kmer_len <= 16 and n >= HAMDIST_MMA_MIN_N and slots <= HAMDIST_MMA_MAX_SLOTS and len(head_len) <= 127:
        return "onehot_mma"
    return "popcount"


def hamdist_matrix_u8(kh: np.ndarray, labels: np.ndarray, head_len: Sequence[int], kmer_len: int, row0: int = 0,
                      row1: Optional[int] = None, out: Optional[torch.Tensor] = None, impl: Optional[str] = None) -> torch.Tensor:
    """rows [row0,row1) of the pairwise distance matrix as a uint8 device tensor [(row1-row0), n] (md:759-808).
    impl: None = hamdist_formulation(); "popcount" (csrc/hamdist.cu) or "onehot_mma" (csrc/hamdist_mma.cu) force one; both
    write the same bytes."""
    L = lib()
    n = len(kh)
    row1 = n if row1 is None else row1
    if impl is None:
        impl = hamdist_formulation(n, kmer_len, head_len)
    if impl == "onehot_mma":
        return hamdist_matrix_onehot_mma(kh, labels, head_len, kmer_len, row0, row1, out)
    if impl != "popcount":
        raise KmapError(f"unknown distance-matrix formulation {impl!r}")
    hd = get_hash_dtype(kmer_len)
    kh_d = E.to_device(np.asarray(kh).astype(hd, copy=False))
    lab_d = E.to_device(np.asarray(labels).astype(np.int32, copy=False))
    hl_d = E.to_device(np.asarray(list(head_len), dtype=np.int32)) if len(head_len) else None
    if out is None:
        out = E.empty((row1 - row0) * n, torch.uint8)
    fn = L.kmap_hamdist_matrix_u32 if hd == np.uint32 else L.kmap_hamdist_matrix_u64
    check(fn(kh_d.data_ptr(), lab_d.data_ptr(), n, kmer_len, None if hl_d is None else hl_d.data_ptr(), len(head_len),
             row0, row1, out.data_ptr(), torch.cuda.current_stream().cuda_stream), "kmap_hamdist_matrix")
    return out.view(row1 - row0, n)


def hamdist_matrix_onehot_mma(kh: np.ndarray, labels: np.ndarray, head_len: Sequence[int], kmer_len: int, row0: int = 0,
                              row1: Optional[int] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """The rows of hamdist_matrix_u8 (k <= 16) computed as an int8 one-hot GEMM on the tcgen05 tensor cores
    (csrc/hamdist_mma.cu): one-hot A x complement-one-hot B^T is the distance itself, the head override of same-label pairs
    is a few extra K columns, the tiles leave through TMA stores.  Bit-identical to the XOR/popcount kernel."""
    L = lib()
    n = len(kh)
    row1 = n if row1 is None else row1
    if kmer_len > 16:
        raise KmapError("the one-hot GEMM formulation covers k <= 16 (32-bit hashes)")
    kh_d = E.to_device(np.asarray(kh).astype(np.uint32, copy=False))
    lab_d = E.to_device(np.asarray(labels).astype(np.int32, copy=False))
    hl_d = E.to_device(np.asarray(list(head_len), dtype=np.int32)) if len(head_len) else None
    if out is None:
        out = E.empty((row1 - row0) * n, torch.uint8)
    nbytes = int(L.kmap_hamdist_mma_scratch_bytes(n))
    scratch = E.empty(nbytes, torch.uint8)
    check(L.kmap_hamdist_matrix_onehot_mma(kh_d.data_ptr(), lab_d.data_ptr(), n, kmer_len, None if hl_d is None else hl_d.data_ptr(),
                                           len(head_len), row0, row1, out.data_ptr(), scratch.data_ptr(), nbytes,
                                           torch.cuda.current_stream().cuda_stream), "kmap_hamdist_matrix_onehot_mma")
    return out.view(row1 - row0, n)


def cal_samp_kmer_hamdist_mat(samp_kh_arr: np.ndarray, samp_cnts: np.ndarray, samp_label_arr: np.ndarray,
                              conseq_list: List[str], kmer_len: int, uniq_dist_flag=False, dtype=int) -> np.ndarray:
    """:759-808.  The expansion by samp_cnts is fused: the kernel computes the expanded matrix directly from the
    repeated key / label arrays.  `dtype=int` reproduces the reference's int64 result; pass np.uint8 to keep the
    compact device format (1 B per pair) for large samples."""
    samp_kh_arr = np.asarray(samp_kh_arr)
    assert len(samp_kh_arr) == len(np.unique(samp_kh_arr))
    for conseq in conseq_list:
        assert len(conseq) <= kmer_len
    head_len = [len(c) for c in conseq_list]
    labels = np.asarray(samp_label_arr)
    if uniq_dist_flag:
        kh, lab = samp_kh_arr, labels
    else:
        cnts = np.asarray(samp_cnts)
        assert np.issubdtype(cnts.dtype, np.integer) and np.all(cnts > 0)
        kh, lab = np.repeat(samp_kh_arr, cnts), np.repeat(labels, cnts)
    if len(kh) == 0:
        return np.zeros((0, 0), dtype=dtype)
    mat = hamdist_matrix_u8(kh, lab, head_len, kmer_len).cpu().numpy()
    return mat if np.dtype(dtype) == np.uint8 else mat.astype(dtype)


# ======================================================================================================================
# k-mer sampling for the visualisation (:812-921)
# ======================================================================================================================
def label_kmers(uniq_kh_arr, conseq_list: List[str], kmer_len: int, motif_def_dict: dict, revcom_mode=True):
    """:849-892, the deterministic part of sample_disp_kmer in one device pass (kmap_label_kmers_*): (k-mers with the
    reverse-complement-closer ones flipped, label per k-mer; label == len(conseq_list) means "no motif")."""
    hd = get_hash_dtype(kmer_len)
    kh = np.asarray(uniq_kh_arr)
    n, n_conseq = len(kh), len(conseq_list)
    if n == 0:
        return kh.copy(), np.zeros(0, dtype=np.intp)
    ckh = [kmer2hash(c) for c in conseq_list]
    rckh = [revcom_hash(h, len(c)) for h, c in zip(ckh, conseq_list)]
    if revcom_mode:
        assert all(int(a) <= int(b) for a, b in zip(ckh, rckh))
    L = lib()
    kh_d = E.to_device(kh.astype(hd, copy=True))
    c_d = E.to_device(np.array([int(x) for x in ckh], dtype=hd))
    rc_d = E.to_device(np.array([int(x) for x in rckh], dtype=hd))
    len_d = E.to_device(np.array([len(c) for c in conseq_list], dtype=np.int32))
    dmax_d = E.to_device(np.array([motif_def_dict[len(c)].max_ham_dist for c in conseq_list], dtype=np.int32))
    label_d = E.empty(n, torch.int32)
    fn = L.kmap_label_kmers_u32 if hd == np.uint32 else L.kmap_label_kmers_u64
    check(fn(kh_d.data_ptr(), n, kmer_len, c_d.data_ptr(), rc_d.data_ptr(), len_d.data_ptr(), dmax_d.data_ptr(), n_conseq,
             int(motif_def_dict[kmer_len].max_ham_dist), int(revcom_mode), label_d.data_ptr(),
             torch.cuda.current_stream().cuda_stream), "kmap_label_kmers")
    return E.to_host(kh_d, hd).astype(kh.dtype, copy=False), label_d.cpu().numpy().astype(np.intp)


def sample_disp_kmer(conseq_list: List[str], kmer_len: int, motif_def_dict: dict, kmer_count_dir: Path,
                     n_total_sample=5000, n_motif_kmer=2500, revcom_mode=True) -> Tuple:
    conseq_list = [s for s in conseq_list if 2 < len(s) <= kmer_len]
    assert len(conseq_list) > 0
    assert all(len(a) >= len(b) for a, b in zip(conseq_list, conseq_list[1:]))
    with open(Path(kmer_count_dir) / f"k{kmer_len}.pkl", "rb") as fh:
        res_list = pickle.load(fh)
    assert res_list[0] == kmer_len
    uniq_kh_arr, uniq_kh_cnt_arr = res_list[1], res_list[2]
    total = int(np.sum(uniq_kh_cnt_arr, dtype=np.int64))
    sampling_flag = True
    if n_total_sample > total:
        warnings.warn(f"The number of samples n_sample={n_total_sample} is larger than the original "
                      f"data n_seq={total}, process and return original data.")
        sampling_flag = False
    n_conseq = len(conseq_list)
    uniq_kh_arr, label_arr = label_kmers(uniq_kh_arr, conseq_list, kmer_len, motif_def_dict, revcom_mode)
    if not sampling_flag:
        return uniq_kh_arr, uniq_kh_cnt_arr, label_arr, conseq_list
    sample_cnt_arr = np.bincount(label_arr, weights=uniq_kh_cnt_arr)
    motif_weights = sample_cnt_arr[:-1] / sum(sample_cnt_arr[:-1])
    sample_cnt_arr[:-1] = np.around(n_motif_kmer * motif_weights)
    sample_cnt_arr[-1] = n_total_sample - sum(sample_cnt_arr[0:-1])
    sample_cnt_arr = sample_cnt_arr.astype(int)
    assert len(sample_cnt_arr) == n_conseq + 1
    samp_inds, samp_cnts = [], []
    for c in range(n_conseq + 1):
        c_inds = np.where(label_arr == c)[0]
        ws = uniq_kh_cnt_arr[c_inds]
        ws = ws / np.sum(ws, dtype=np.int64)       # (:908 uses the builtin sum: the same integer, 0.1 s per 1e7 counts slower)
        tmpcnts = np.random.multinomial(sample_cnt_arr[c], ws, size=1).squeeze()
        samp_inds.append(c_inds[tmpcnts > 0])
        samp_cnts.append(tmpcnts[tmpcnts > 0])
    samp_inds = np.concatenate(samp_inds)
    samp_cnts = np.concatenate(samp_cnts)
    return uniq_kh_arr[samp_inds], samp_cnts, label_arr[samp_inds], conseq_list


# ======================================================================================================================
# merge candidates of different lengths (:533-591) -- host string logic
# ======================================================================================================================
def merge_consensus_seqs(conseq_list: List[str]) -> List[str]:
    def related(long_kmer, short_kmer):
        # a (len-1)-substring of short_kmer occurs in long_kmer
        return short_kmer[:-1] in long_kmer or short_kmer[1:] in long_kmer

    remaining = sorted(conseq_list, key=len, reverse=True)
    final_conseq_list = []
    while remaining:
        cur = remaining[0]
        rc_cur = reverse_complement(cur)

        def hit(s):
            return related(cur, s) or related(rc_cur, s)

        one_shorter = next((s for s in remaining if len(s) == len(cur) - 1 and hit(s)), None)
        two_shorter = next((s for s in remaining if len(s) == len(cur) - 2 and hit(s)), None)
        if one_shorter and two_shorter:
            final_conseq_list.append(one_shorter)
            remaining = [s for s in remaining if not hit(s)]
        else:
            remaining = remaining[1:]
    return final_conseq_list


# ======================================================================================================================
# scan_motif driver (:187-486) -- same files, same resume-by-file-existence behaviour
# ======================================================================================================================
def _scan_motif(res_dir: str, debug=False):
    import tomllib
    res = Path(res_dir)
    config_file_path = res / FileNameDict["config_file"]
    motif_def_file_path = res / FileNameDict["motif_def_file"]
    proc_fasta_file_path = res / FileNameDict["processed_fasta_file"]
    assert config_file_path.exists()
    assert motif_def_file_path.exists()
    assert proc_fasta_file_path.exists()
    with open(config_file_path, "rb") as fh:
        config_dict = tomllib.load(fh)
    motif_def_dict = gen_motif_def_dict(config_dict, debug=debug)
    min_k = config_dict["kmer_count"]["min_k"]
    max_k = config_dict["kmer_count"]["max_k"]
    revcom_mode = config_dict["kmer_count"]["revcom_mode"]
    rep_mode = config_dict["general"]["repetitive_mode"]
    md_cfg = config_dict["motif_discovery"]

    mask_noise_seq_list = []
    if md_cfg["noise_kmer_file"] != "None":
        assert Path(md_cfg["noise_kmer_file"]).exists()
        with open(Path(md_cfg["noise_kmer_file"]), "r") as fh:
            mask_noise_seq_list = [ln.strip() for ln in fh if ln.strip()]
    from .api import DistContext, concat_occurrence_shards, shard_reads
    ctx = DistContext.from_env()          # one process per GPU under torchrun; a single process otherwise
    with open(proc_fasta_file_path, "rb") as fh:
        seq_np_arr = pickle.load(fh)
    if mask_noise_seq_list:
        seq_np_arr = mask_ham_ball(seq_np_arr, motif_def_dict, mask_noise_seq_list, [0 for _ in mask_noise_seq_list])
    boarder_pkl_file = res / FileNameDict["processed_fasta_seqboarder_file"]
    with open(boarder_pkl_file, "rb") as fh:
        boarder_mat = pickle.load(fh)
    n_all_seq = len(boarder_mat)
    if ctx.world > 1:                     # this rank's contiguous range of reads (reads are independent units)
        seq_np_arr, boarder_mat = shard_reads(seq_np_arr, boarder_mat, ctx.rank, ctx.world)

    top_k, n_trial = md_cfg["top_k"], md_cfg["n_trial"]
    save_kmer_cnt_flag = md_cfg["save_kmer_cnt_flag"]
    candidate_conseq_list = []
    kmer_count_dir = res / FileNameDict["kmer_count_dir"]
    if save_kmer_cnt_flag and ctx.is_root:
        kmer_count_dir.mkdir(exist_ok=True)
    input_fasta_file = Path(config_dict["general"]["input_fasta_file"])

    # one device copy of the (noise-masked) reads serves every k; the occurrence scans use the unmasked FASTA content
    dev = E.SeqOnDevice.from_numpy(seq_np_arr, boarder_mat)
    dev.snapshot_valid()
    occ_dev = None

    def occurrence_dev():
        nonlocal occ_dev
        if occ_dev is None:
            assert input_fasta_file.exists()
            fa_dev = E.SeqOnDevice.from_fasta(input_fasta_file, rank=ctx.rank, world=ctx.world)   # parsed on the device
            all_borders = getattr(fa_dev, "all_borders", fa_dev.borders)
            occ_dev = (fa_dev, E.to_host(all_borders, np.int64).reshape(-1, 2) if ctx.is_root else None)
        return occ_dev

    candidate_conseq_file = res / FileNameDict["candidate_conseq_file"]
    if ctx.agree(candidate_conseq_file.exists()):
        print(f"{candidate_conseq_file} already exist, re-use it.")
    else:
        store_flag = md_cfg["store_conseq_occur_info_flag"]
        header = "kmer_len,conseq_hash,conseq,conseq_rc,hamball_proportion,hamball_ratio,log10_p_value"
        if store_flag:
            header += ",n_motif_reads,n_all_reads,motif_reads_prop,motif_occurrence,motif_occurrence_per_motif_read"
        rows = [header]
        # k{k}.pkl files of an earlier run replace the first counts (:621-624); rank 0 looks, every rank follows
        have_pkl = ctx.agree({k: bool(save_kmer_cnt_flag and (kmer_count_dir / f"k{k}.pkl").exists())
                              for k in range(min_k, max_k + 1)})
        # The first-round counts of every k with a dense table in ONE pass over the reads (:262-273 calls find_motif, and with
        # it comp_kmer_hash + remove_duplicate_hash_per_seq + count_uniq_hash, once per k): SeqOnDevice.count_all updates
        # only the table of the largest k per window and derives the smaller ones (csrc/count_all.cu, csrc/partition.cu).
        # Under torchrun the tables of the shards are merged by one all-reduce of the flat buffer.
        first_tables = {}
        ks_first = [k for k in range(min_k, min(max_k, ALL_K_MAX) + 1) if not have_pkl[k]]
        if ks_first:
            alloc = ctx.table_allreduce.alloc_tables if ctx.table_allreduce is not None else E.alloc_tables
            flat, first_tables = alloc(ks_first[0], ks_first[-1])      # (sharded: in the peer region of the exchange, csrc/peer.cu)
            dev.count_all(ks_first[0], ks_first[-1], dedup=not rep_mode, tables=first_tables, merge=ctx.table_allreduce)
            if ctx.table_allreduce is not None:
                ctx.table_allreduce.check()
        k_single = [k for k in range(min_k, min(max_k, 15) + 1) if k not in first_tables]
        table_buffer = E.empty(1 << (2 * max(k_single)), torch.int32) if k_single else None   # one allocation for those k
        for kmer_len in range(min_k, max_k + 1):
            dev.restore_valid()
            m = motif_def_dict[kmer_len]
            kmer_cnt_file = kmer_count_dir / f"k{kmer_len}.pkl"
            first_lists = None
            if have_pkl[kmer_len]:
                with open(kmer_cnt_file, "rb") as fh:
                    k_from_file, kh0, cnt0 = pickle.load(fh)
                    assert kmer_len == k_from_file
                first_lists = (kh0, cnt0)
            consensus_kh_dict, first = find_motif_on_device(dev, kmer_len, m.max_ham_dist, m.p_uniform, m.ratio_mu,
                                                            m.ratio_std, m.ratio_cutoff, top_k, n_trial, revcom_mode,
                                                            rep_mode, first_lists, first_tables.get(kmer_len), debug,
                                                            table_buffer=table_buffer, ctx=ctx,
                                                            want_first=bool(save_kmer_cnt_flag and not have_pkl[kmer_len]))
            if save_kmer_cnt_flag and not have_pkl[kmer_len] and ctx.is_root:
                with open(kmer_cnt_file, "wb") as fh:
                    pickle.dump([kmer_len, first[0], first[1]], fh)
            tmp_candidate_conseq_list = [hash2kmer(kh, kmer_len) for kh in consensus_kh_dict]
            occ_stats = []
            if store_flag:
                tmp_occurence_file = kmer_count_dir / f"k{kmer_len}.motif_occurence.csv"
                occ_stats = gen_motif_occurence_file(tmp_candidate_conseq_list, motif_def_dict, input_fasta_file,
                                                     tmp_occurence_file, revcom_mode, _dev_cache=occurrence_dev(), _ctx=ctx)
            for i, kmer_seq in enumerate(tmp_candidate_conseq_list):
                candidate_conseq_list.append(kmer_seq)
                if not ctx.is_root:
                    continue
                kh = kmer2hash(kmer_seq)
                prop, ratio, log10_p_value = consensus_kh_dict[kh]
                n_motif_seq, n_motif_occurrence = -n_all_seq, -n_all_seq
                if store_flag:
                    n_motif_seq, n_motif_occurrence = occ_stats[i]      # == get_motif_seq_num(tmp_occurence_file, i) (:1345-1393)
                motif_seq_prop = float(n_motif_seq) / n_all_seq
                motif_per_motif_seq = float(n_motif_occurrence) / n_motif_seq
                row = (f"{kmer_len},{kh},{kmer_seq},{reverse_complement(kmer_seq)},{prop:0.8f},"
                       f"{ratio:0.4f},{log10_p_value:0.4f}")
                if store_flag:
                    row += (f",{n_motif_seq},{n_all_seq},{motif_seq_prop:0.4f},{n_motif_occurrence},"
                            f"{motif_per_motif_seq:0.2f}")
                rows.append(row)
        del first_tables, table_buffer
        print(f"kmer counting finished for k={min_k}...{max_k}. Candidate consensus sequences generated.")
        if ctx.is_root:
            write_lines(rows, candidate_conseq_file)
        ctx.barrier()

    final_conseq_file = res / FileNameDict["final_conseq_file"]
    if ctx.agree(final_conseq_file.exists()):
        final_conseq_list = ctx.agree(final_conseq_file.read_text().splitlines() if ctx.is_root else None)
        print(f"{final_conseq_file} already exist, re-use it.")
    else:
        final_conseq_list = ctx.agree(merge_consensus_seqs(candidate_conseq_list) if ctx.is_root else None)
        if ctx.is_root:
            write_lines(final_conseq_list, final_conseq_file)

    final_conseq_info_file = res / FileNameDict["final_conseq_info_file"]
    if not ctx.is_root:
        pass
    elif final_conseq_info_file.exists():
        print(f"{final_conseq_info_file} already exist, re-use it.")
    else:
        final_conseq_list = final_conseq_file.read_text().splitlines()
        cand_lines = candidate_conseq_file.read_text().splitlines()
        elements = cand_lines[0].split(",")
        elements[1] = elements[0]
        elements[0] = "motif_id"
        info = [",".join(elements)]
        motif_ind = 0
        for conseq in final_conseq_list:
            for line in cand_lines:
                if "," + conseq + "," in line:
                    elements = line.split(",")
                    elements[1] = elements[0]
                    elements[0] = str(motif_ind)
                    motif_ind += 1
                    info.append(",".join(elements))
        write_lines(info, final_conseq_info_file)
        print("Final consensus sequences generated.")

    occurence_file = res / FileNameDict["motif_occurence_file"]
    _, final_scan = gen_motif_occurence_file(final_conseq_list, motif_def_dict, input_fasta_file, occurence_file, revcom_mode,
                                             _dev_cache=occurrence_dev(), _ctx=ctx, _return_scan=True)
    if not ctx.is_root:                   # everything below is host work on files: rank 0 alone
        ctx.barrier()
        return

    # the DATA files of the density / co-occurrence steps (:364-425); the pdf figures the reference draws from the same
    # numbers (matplotlib) are not produced here
    if md_cfg["motif_pos_density_flag"] and final_conseq_list:
        x_step = 0.01
        x_arr = np.arange(0, 1.0 + x_step, x_step)
        dens = [pos_density_from_scan(final_scan, i, len(conseq), x_step=x_step, x_arr=x_arr)[2]        # :1256-1343
                for i, conseq in enumerate(final_conseq_list)]
        with open(res / FileNameDict["motif_pos_density_file"], "wb") as fh:
            pickle.dump([x_arr, np.vstack(dens)], fh)
        print("motif position distribution generated (figures: reference package).")
    if md_cfg["motif_co_occurence_flag"] and final_conseq_list:
        co_occur_dir = res / FileNameDict["co_occur_dir"]
        co_occur_dir.mkdir(exist_ok=True)
        co_occur_mat_file = co_occur_dir / FileNameDict["co_occur_mat_file"]
        if co_occur_mat_file.exists():
            print(f"{co_occur_mat_file}, re-use it!")
        else:
            if final_scan["cooc"] is not None:     # integer parts on the device, from the scan results (csrc/consumers.cu)
                co_occur_mat, loc_dist_mat, loc_dist_dict = co_occurrence_from_scan(final_scan, len(final_conseq_list))
            else:                                  # more than 31 final motifs: the file, like the reference (:1189-1254)
                co_occur_mat, loc_dist_mat, loc_dist_dict = get_motif_co_occurence_mat(occurence_file, len(final_conseq_list))
            co_sum_mat = np.diag(co_occur_mat) + np.diag(co_occur_mat).reshape((-1, 1))
            co_occur_norm_mat = 2 * co_occur_mat / co_sum_mat
            write_co_occurence_mat(co_occur_mat_file, co_occur_mat + 0.0, final_conseq_list)
            write_co_occurence_mat(co_occur_dir / FileNameDict["co_occur_mat_norm_file"], co_occur_norm_mat, final_conseq_list)
            write_co_occurence_mat(co_occur_dir / FileNameDict["co_occur_dist_mat_file"], loc_dist_mat, final_conseq_list)
            write_co_occurence_dist_arr(co_occur_dir / FileNameDict["co_occur_dist_data_file"], loc_dist_dict, final_conseq_list)
        print("motif co-occurence matrix generated (figures: reference package).")

    if md_cfg["sample_kmer_flag"] and not save_kmer_cnt_flag:
        print(f"kmers cannot be sampled when {save_kmer_cnt_flag=}, skip kmer sampling!")
    sample_kmer_pkl_file = res / FileNameDict["sample_kmer_pkl_file"]
    sample_kmer_txt_file = res / FileNameDict["sample_kmer_txt_file"]
    if sample_kmer_pkl_file.exists():
        print(f"sample kmer file {sample_kmer_pkl_file} exists, skip sampling!")
    elif md_cfg["sample_kmer_flag"] and save_kmer_cnt_flag and final_conseq_list:
        n_total_sample, n_motif_sample = md_cfg["n_total_sample"], md_cfg["n_motif_sample"]
        kmer_len = max(len(conseq) for conseq in final_conseq_list)
        samp_kh_arr, samp_cnts, samp_label_arr, conseq_list = sample_disp_kmer(
            final_conseq_list, kmer_len, motif_def_dict, kmer_count_dir=kmer_count_dir, n_total_sample=n_total_sample,
            n_motif_kmer=n_motif_sample, revcom_mode=revcom_mode)
        with open(sample_kmer_pkl_file, "wb") as fh:
            pickle.dump([samp_kh_arr, samp_cnts, samp_label_arr, conseq_list], fh)
        lines = []
        for kh, cnt, label in zip(samp_kh_arr, samp_cnts, samp_label_arr):
            lines.extend([f"{hash2kmer(kh, kmer_len)}\t{label}"] * int(cnt))
        write_lines(lines, sample_kmer_txt_file)
        print(f"kmers are sampled for visualization. {kmer_len= }, {n_total_sample= }, {n_motif_sample= }")
        hamdist_mat = cal_samp_kmer_hamdist_mat(samp_kh_arr, samp_cnts, samp_label_arr, conseq_list, kmer_len,
                                                uniq_dist_flag=False)
        label_arr = _convert_to_block_arr(samp_label_arr, samp_cnts)
        with open(res / FileNameDict["sample_kmer_hamdist_mat_file"], "wb") as fh:
            pickle.dump([kmer_len, hamdist_mat, label_arr], fh)
        print("Hamming distance matrix of sampled kmers are generated.")

    if md_cfg["gen_hamball_flag"]:
        out_dir_path = res / FileNameDict["hamball_dir"]
        out_dir_path.mkdir(exist_ok=True)
        for i, conseq in enumerate(final_conseq_list):
            output_cntmat_file = str(out_dir_path / f"cntmat_motif{i}_{conseq}.csv")
            if Path(output_cntmat_file).exists():
                print(f"motif matrix file {output_cntmat_file} exist, skip generating.")
                continue
            _ex_hamball(res_dir, conseq, "matrix", output_cntmat_file, max_ham_dist=motif_def_dict[len(conseq)].max_ham_dist)
        print("Motif count matrix extracted (logos are drawn by the reference package's draw_logo).")
    print("All tasks of scan motif finished.")
    ctx.barrier()
