"""CPU oracle for KMAP's scan_motif counting path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A NumPy restatement of the reference algorithm (chengl7-lab/kmap v0.0.7).  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may import this module;
`kmap_b200/` never does (tests/test_no_oracle_in_product.py enforces it).

Parity status: PINNED.  `tests/golden/*.pkl` were produced by running the unmodified reference source under
an import shim (tests/golden/ref_shim.py + make_golden.py); tests/test_oracle_golden.py checks every function
below against those outputs and against the reference's own known-answer tests
(tests/kmap_tests.py:173-189, 212-238, 241-266, 268-284, 434-441; tests/test_kmer_count.py:51-71).

Every function cites the reference file:line it follows (paths relative to /root/reference/src/kmap).
Where the reference runs a Taichi kernel (a parallel-for over a NumPy array) the restatement is the
equivalent vectorised NumPy expression; where the reference calls NumPy (`np.unique`, `np.intersect1d`,
`np.delete`, `np.argpartition`) the same NumPy call is made, because its tie/ordering behaviour is part of
the observable result.
"""
from __future__ import annotations

from dataclasses import dataclass
from pathlib import Path
from typing import Dict, List, Sequence, Tuple

import numpy as np

MISSING_VAL = 255  # kmer_count.py:58


# ----------------------------------------------------------------------------------------------------------
# dtype rules  (kmer_count.py:351-370)
def get_cnt_dtype(kmer_len: int):
    return np.int32 if kmer_len < 16 else np.int64


def get_hash_dtype(kmer_len: int):
    if 0 < kmer_len < 16:
        return np.uint32
    if kmer_len < 32:
        return np.uint64
    raise Exception(f"max_kmer_len=31, kmer_len={kmer_len} is greater the maximum value.")


def get_invalid_hash(dtype):
    return dtype(np.iinfo(dtype).max)


# ----------------------------------------------------------------------------------------------------------
# encoding  (kmer_count.py:244-263, 308-347)
_ENC = np.full(256, MISSING_VAL, dtype=np.uint8)
for _c, _v in zip(b"ACGT", range(4)):
    _ENC[_c] = _v


def dna2arr(dna_str: str, dtype=np.uint8, append_missing_val_flag: bool = True) -> np.ndarray:
    """A0 C1 G2 T3, anything else 255; optional trailing 255 separator (kmer_count.py:244-263).
    Case-sensitive like the reference (callers upper-case first, kmer_count.py:316)."""
    raw = np.frombuffer(dna_str.encode("latin-1", "replace"), dtype=np.uint8)
    body = _ENC[raw].astype(dtype)
    if append_missing_val_flag:
        return np.concatenate([body, np.array([MISSING_VAL], dtype=dtype)])
    return body


def arr2dna(arr: np.ndarray) -> str:
    """kmer_count.py:238-241"""
    lut = np.full(256, ord("?"), dtype=np.uint8)
    lut[[0, 1, 2, 3, MISSING_VAL]] = np.frombuffer(b"ACGTN", dtype=np.uint8)
    return lut[np.asarray(arr, dtype=np.uint8)].tobytes().decode()


def read_fasta(path) -> List[Tuple[str, str]]:
    """(name, sequence) records of a plain / gzipped FASTA file, as Bio.SeqIO.parse(fh, 'fasta') yields them
    (kmer_count.py:308-323): header lines start with '>', sequence lines are concatenated."""
    import gzip
    opener = gzip.open if str(path).endswith(".gz") else open
    recs, name, chunks = [], None, []
    with opener(path, "rt") as fh:
        for line in fh:
            line = line.rstrip("\r\n")
            if line.startswith(">"):
                if name is not None:
                    recs.append((name, "".join(chunks)))
                name, chunks = (line[1:].split() or [""])[0], []
            elif name is not None:
                chunks.append("".join(line.split()))
    if name is not None:
        recs.append((name, "".join(chunks)))
    return recs


def fasta_to_binary(path) -> Tuple[np.ndarray, np.ndarray]:
    """input.bin / input.seqboarder.bin contents (kmer_count.py:326-347): every read upper-cased, encoded,
    followed by one 255; borders[i] = [start, start+L] (second column = index of the separator)."""
    arrs = [dna2arr(s.upper()) for _, s in read_fasta(path)]
    borders = np.zeros((len(arrs), 2), dtype=int)
    p = 0
    for i, a in enumerate(arrs):
        borders[i, 0] = p
        borders[i, 1] = p + len(a) - 1
        p += len(a)
    seq = np.concatenate(arrs) if arrs else np.zeros(0, dtype=np.uint8)
    return seq, borders


# ----------------------------------------------------------------------------------------------------------
# scalar helpers  (kmer_count.py:416-446, 626-640)
def kmer2hash(kmer: str) -> np.uint64:
    assert len(kmer) < 32, "kmer should be shorted than 32 bases"
    kh = 0
    for ch in kmer:
        kh = (kh << 2) + "ACGT".index(ch)
    return np.uint64(kh)


def hash2kmer(hashkey, k: int) -> str:
    h = int(hashkey)
    return "".join("ACGT"[(h >> (2 * (k - 1 - i))) & 3] for i in range(k))


def reverse_complement(seq: str) -> str:
    """kmer_count.py:266-268"""
    comp = {"A": "T", "T": "A", "C": "G", "G": "C"}
    return "".join(comp[b] for b in reversed(seq))


def revcom_hash(in_hash, kmer_len: int):
    """complement = mask - h, then reverse the 2-bit groups (kmer_count.py:626-640); returns the hash dtype."""
    hd = get_hash_dtype(kmer_len)
    mask = (1 << (2 * kmer_len)) - 1
    com = (mask - int(hd(in_hash))) & int(np.iinfo(hd).max)
    ret = com & 3
    for _ in range(kmer_len - 1):
        com >>= 2
        ret = ((ret << 2) & int(np.iinfo(hd).max)) + (com & 3)
    return hd(ret & int(np.iinfo(hd).max))


# ----------------------------------------------------------------------------------------------------------
# Taichi kernels restated  (taichi_core.py:3-224)
def comp_kmer_hash(seq_np_arr: np.ndarray, kmer_len: int) -> np.ndarray:
    """One hash per array position: MSB-first 2-bit pack of seq[i:i+k]; all-ones if the window leaves the
    array or touches a 255 (taichi_core.py:3-31 via kmer_count.py:449-473)."""
    hd = get_hash_dtype(kmer_len)
    n = len(seq_np_arr)
    padded = np.concatenate([np.asarray(seq_np_arr, dtype=np.uint8), np.full(kmer_len, MISSING_VAL, np.uint8)])
    h = np.zeros(n, dtype=hd)
    bad = np.zeros(n, dtype=bool)
    two = hd(2)
    for j in range(kmer_len):
        col = padded[j:j + n]
        bad |= col == MISSING_VAL
        h = (h << two) + col.astype(hd)
    h[bad] = get_invalid_hash(hd)
    return h


def _popcount2bit(x: np.ndarray, n_groups: int) -> np.ndarray:
    """number of non-zero 2-bit groups among the low n_groups groups (taichi_core.py:63-72)."""
    dt = x.dtype.type
    nbits = x.dtype.itemsize * 8
    low = (1 << (2 * n_groups)) - 1 if 2 * n_groups < nbits else (1 << nbits) - 1
    evn = int("01" * (nbits // 2), 2)
    m = (x | (x >> dt(1))) & dt(evn & low)
    return np.bitwise_count(m).astype(np.uint8)


def cal_hamming_dist(kh_arr: np.ndarray, consensus_kh, kmer_len: int) -> np.ndarray:
    """taichi_core.py:75-104 via kmer_count.py:494-515.  Works on empty arrays."""
    hd = get_hash_dtype(kmer_len)
    kh_arr = np.asarray(kh_arr).astype(hd, copy=False)
    target = np.array([consensus_kh]).astype(hd)[0]
    return _popcount2bit(kh_arr ^ target, kmer_len)


def cal_hamming_dist_head(kh_arr, consensus_kh, kmer_len: int, consensus_len: int) -> np.ndarray:
    """first `consensus_len` bases of each k-mer vs the consensus (taichi_core.py:108-124, 144-160)."""
    assert consensus_len <= kmer_len
    hd = get_hash_dtype(kmer_len)
    kh_arr = np.asarray(kh_arr).astype(hd, copy=False)
    target = np.array([consensus_kh]).astype(hd)[0]
    return _popcount2bit((kh_arr >> hd(2 * (kmer_len - consensus_len))) ^ target, consensus_len)


def cal_hamming_dist_tail(kh_arr, consensus_kh, kmer_len: int, consensus_len: int) -> np.ndarray:
    """last `consensus_len` bases (taichi_core.py:127-141, 163-177)."""
    assert consensus_len <= kmer_len
    hd = get_hash_dtype(kmer_len)
    kh_arr = np.asarray(kh_arr).astype(hd, copy=False)
    target = np.array([consensus_kh]).astype(hd)[0]
    return _popcount2bit(kh_arr ^ target, consensus_len)


def get_revcom_hash_arr(in_hash_arr: np.ndarray, kmer_len: int) -> np.ndarray:
    """taichi_core.py:181-224 via kmer_count.py:613-623"""
    hd = get_hash_dtype(kmer_len)
    src = np.asarray(in_hash_arr)
    a = src.astype(hd)          # the kernel computes in the hash dtype; out = np.empty_like(in) (kmer_count.py:617)
    com = hd((1 << (2 * kmer_len)) - 1) - a
    ret = com & hd(3)
    for _ in range(kmer_len - 1):
        com = com >> hd(2)
        ret = (ret << hd(2)) + (com & hd(3))
    return ret.astype(src.dtype)


# ----------------------------------------------------------------------------------------------------------
# counting  (kmer_count.py:476-491, 643-685, 743-760)
def remove_duplicate_hash_per_seq(hash_arr: np.ndarray, boarder_mat: np.ndarray, invalid_hash) -> np.ndarray:
    """Within every read [st, en) keep the first occurrence of each hash, set the others to invalid.
    In place + returned.  Same per-read loop as the reference (kmer_count.py:743-760)."""
    assert boarder_mat.shape[1] == 2
    for st, en in boarder_mat:
        blank = np.full(en - st, invalid_hash, dtype=hash_arr.dtype)
        vals, first = np.unique(hash_arr[st:en], return_index=True)
        blank[first] = vals
        hash_arr[st:en] = blank
    return hash_arr


def count_uniq_hash(hash_arr: np.ndarray, kmer_len: int) -> Tuple[np.ndarray, np.ndarray]:
    """np.unique(return_counts) minus the invalid hash; counts in the count dtype (kmer_count.py:476-491)."""
    hd = get_hash_dtype(kmer_len)
    uniq, cnt = np.unique(hash_arr, return_counts=True)
    keep = uniq != get_invalid_hash(hd)
    return uniq[keep], cnt[keep].astype(get_cnt_dtype(kmer_len))


def merge_revcom(uniq_kmer_hash_arr, uniq_kh_cnt_arr, kmer_len: int, keep_lower_hash_flag: bool = True):
    """kmer_count.py:643-685.  cnt[h] += cnt[rc h] for every h whose rc is present (palindromes pair with
    themselves -> doubled); entries on the wrong side of a present pair are deleted; lone entries on the
    wrong side are relabelled to their rc.  Result order = ascending forward hash of the survivors.
    Mutates the caller's count array exactly like the reference's `+=` (kmer_count.py:661)."""
    rc = get_revcom_hash_arr(uniq_kmer_hash_arr, kmer_len)
    _, nat_inds, rc_inds = np.intersect1d(uniq_kmer_hash_arr, rc, return_indices=True)
    uniq_kh_cnt_arr[nat_inds] += uniq_kh_cnt_arr[rc_inds]
    if keep_lower_hash_flag:
        wrong = uniq_kmer_hash_arr[nat_inds] > rc[nat_inds]
    else:
        wrong = uniq_kmer_hash_arr[nat_inds] < rc[nat_inds]
    drop = nat_inds[wrong]
    kh = np.delete(uniq_kmer_hash_arr, drop)
    rc = np.delete(rc, drop)
    cnt = np.delete(uniq_kh_cnt_arr, drop)
    swap = kh > rc if keep_lower_hash_flag else kh < rc
    kh[swap] = rc[swap]
    return kh, cnt


# ----------------------------------------------------------------------------------------------------------
# masking  (kmer_count.py:580-610, 688-723)
def mask_input(seq_np_arr: np.ndarray, kmer_len: int, consensus_kh_arr, max_hamball_dist_arr) -> np.ndarray:
    """Hash every position once; for each consensus flag positions with dist <= d *on that pre-mask hash
    array* and overwrite seq[i:min(i+k,n)] with 255.  Invalid positions are not skipped: their all-ones hash
    is compared like T..T (kmer_count.py:592-607).  In place + returned."""
    n = len(seq_np_arr)
    kh_hash_arr = comp_kmer_hash(seq_np_arr, kmer_len)
    for consensus_kh, max_d in zip(consensus_kh_arr, max_hamball_dist_arr):
        if n == 0:
            continue
        dist = cal_hamming_dist(kh_hash_arr, consensus_kh, kmer_len)
        if np.min(dist) > max_d:
            continue
        flagged = np.flatnonzero(dist <= max_d)
        delta = np.zeros(n + 1, dtype=np.int64)            # == the reference's per-position slice assignment
        np.add.at(delta, flagged, 1)
        np.add.at(delta, np.minimum(flagged + kmer_len, n), -1)
        seq_np_arr[np.cumsum(delta[:n]) > 0] = MISSING_VAL
    return seq_np_arr


def mask_input_loop(seq_np_arr, kmer_len, consensus_kh_arr, max_hamball_dist_arr):
    """Same as mask_input but with the reference's per-position Python loop (kmer_count.py:599-602); used to
    cross-check the vectorised form and for the reference-faithful CPU timing."""
    kh_hash_arr = comp_kmer_hash(seq_np_arr, kmer_len)
    for consensus_kh, max_d in zip(consensus_kh_arr, max_hamball_dist_arr):
        dist = cal_hamming_dist(kh_hash_arr, consensus_kh, kmer_len)
        if len(dist) == 0 or np.min(dist) > max_d:
            continue
        for i, flag in enumerate(dist <= max_d):
            if flag:
                j = i + kmer_len if i + kmer_len < len(seq_np_arr) else len(seq_np_arr)
                seq_np_arr[i:j] = MISSING_VAL
    return seq_np_arr


def mask_ham_ball(seq_np_arr, motif_def_dict, consensus_seq_list, max_ham_dist_list=()):
    """kmer_count.py:688-723: group user consensus strings by length (ascending), mask each group."""
    lens = np.array([len(c) for c in consensus_seq_list])
    if len(max_ham_dist_list) == 0:
        max_ham_dist_list = [motif_def_dict[int(L)].max_ham_dist for L in lens]
    assert len(max_ham_dist_list) == len(consensus_seq_list)
    for L in np.unique(lens):
        idx = np.where(lens == L)[0]
        khs = np.array([kmer2hash(consensus_seq_list[i]) for i in idx])
        ds = np.array([max_ham_dist_list[i] for i in idx])
        seq_np_arr = mask_input(seq_np_arr, int(L), khs, ds)
    return seq_np_arr


# ----------------------------------------------------------------------------------------------------------
# motif definition table  (kmer_count.py:221-235, 726-740)
@dataclass
class MotifDef:
    kmer_len: int
    p_uniform: float
    max_ham_dist: int
    ratio_mu: float
    ratio_std: float
    ratio_cutoff: float


def init_motif_def_dict(motif_def_file, p_value_cutoff: float = 1e-10) -> dict:
    import pandas as pd
    from scipy.stats import norm
    out = {"p_value_cutoff": p_value_cutoff}
    for _, row in pd.read_csv(motif_def_file).iterrows():
        k = int(row["kmer_len"])
        cutoff = norm.ppf(1 - p_value_cutoff, loc=row["ratio_mu"], scale=row["ratio_std"])
        out[k] = MotifDef(k, row["p_uniform"], int(row["max_ham_dist"]), row["ratio_mu"], row["ratio_std"], cutoff)
    return out


# ----------------------------------------------------------------------------------------------------------
# find_motif  (motif_discovery.py:594-702)
def first_count(seq_np_arr, kmer_len, boarder_mat, merge_revcom_mode=True, rep_mode=False):
    """motif_discovery.py:627-640: hash, per-read de-dup unless repetitive mode, unique-count, rc merge."""
    hash_arr = comp_kmer_hash(seq_np_arr, kmer_len)
    if not rep_mode:
        hash_arr = remove_duplicate_hash_per_seq(hash_arr, boarder_mat, get_invalid_hash(get_hash_dtype(kmer_len)))
    kh, cnt = count_uniq_hash(hash_arr, kmer_len)
    if merge_revcom_mode:
        kh, cnt = merge_revcom(kh, cnt, kmer_len, keep_lower_hash_flag=True)
    return kh, cnt


def recount(seq_np_arr, kmer_len, merge_revcom_mode=True):
    """motif_discovery.py:695-699: after masking NO per-read de-dup is applied."""
    kh, cnt = count_uniq_hash(comp_kmer_hash(seq_np_arr, kmer_len), kmer_len)
    if merge_revcom_mode:
        kh, cnt = merge_revcom(kh, cnt, kmer_len, keep_lower_hash_flag=True)
    return kh, cnt


def hamball_count(uniq_kh_arr, uniq_kh_cnt_arr, kh, kmer_len, max_ham_dist, merge_revcom_mode=True) -> int:
    """motif_discovery.py:667-673"""
    dist = cal_hamming_dist(uniq_kh_arr, kh, kmer_len)
    if merge_revcom_mode:
        dist = np.minimum(dist, cal_hamming_dist(uniq_kh_arr, revcom_hash(kh, kmer_len), kmer_len))
    return int(np.sum(uniq_kh_cnt_arr[dist <= max_ham_dist], dtype=np.int64))


def find_motif(seq_np_arr, kmer_len, max_ham_dist, p_unif, ratio_mu, ratio_std, ratio_cutoff, top_k=5, n_trial=10,
               merge_revcom_mode=True, rep_mode=False, boarder_mat=None, first_count_result=None, trace=None):
    """Statement-by-statement restatement of motif_discovery.py:594-702 without the pickle I/O:
    `first_count_result` plays the role of an existing k{k}.pkl (motif_discovery.py:621-624).
    Mutates seq_np_arr.  Returns (dict consensus_hash -> (proportion, ratio, log10_p), (uniq_kh, uniq_cnt) of the
    first round == the k{k}.pkl payload).  n_total_kmer is the exact integer sum."""
    from scipy.stats import norm
    if first_count_result is not None:
        uniq_kh_arr, uniq_kh_cnt_arr = first_count_result
    else:
        uniq_kh_arr, uniq_kh_cnt_arr = first_count(seq_np_arr, kmer_len, boarder_mat, merge_revcom_mode, rep_mode)
    first = (uniq_kh_arr, uniq_kh_cnt_arr)
    n_total_kmer = int(np.sum(uniq_kh_cnt_arr, dtype=np.int64))
    res = {}
    for i_trial in range(n_trial):
        if top_k > len(uniq_kh_cnt_arr):
            break
        top_k_inds = np.array(np.argpartition(uniq_kh_cnt_arr, -top_k)[-top_k:])
        if len(top_k_inds) == 0:
            break
        hamball_cnt_arr = np.zeros(top_k)
        for i, ind in enumerate(top_k_inds):
            hamball_cnt_arr[i] = hamball_count(uniq_kh_arr, uniq_kh_cnt_arr, uniq_kh_arr[ind], kmer_len, max_ham_dist,
                                               merge_revcom_mode)
        best = np.argmax(hamball_cnt_arr)
        consensus_kh = uniq_kh_arr[top_k_inds[best]]
        proportion = (hamball_cnt_arr[best] + 0.0) / n_total_kmer
        ratio = proportion / p_unif
        if trace is not None:
            trace.append(dict(trial=i_trial, top_k_inds=top_k_inds.copy(), top_k_kh=uniq_kh_arr[top_k_inds].copy(),
                              hamball=hamball_cnt_arr.copy(), consensus=consensus_kh, ratio=ratio))
        if ratio > ratio_cutoff:
            res[consensus_kh] = (proportion, ratio, norm.logsf(ratio, loc=ratio_mu, scale=ratio_std) / np.log(10))
            if merge_revcom_mode:
                rc = revcom_hash(consensus_kh, kmer_len)
                seq_np_arr = mask_input(seq_np_arr, kmer_len, np.array([consensus_kh, rc]),
                                        np.array([max_ham_dist, max_ham_dist]))
            else:
                seq_np_arr = mask_input(seq_np_arr, kmer_len, np.array([consensus_kh]), np.array([max_ham_dist]))
            uniq_kh_arr, uniq_kh_cnt_arr = recount(seq_np_arr, kmer_len, merge_revcom_mode)
        else:
            break
    return res, first


# ----------------------------------------------------------------------------------------------------------
# Hamming-ball extraction + count matrix  (motif_discovery.py:924-986)
def ex_hamball_from_arrays(uniq_kh_arr, uniq_kh_cnt_arr, conseq: str, max_ham_dist: int, revcom_mode=True):
    """motif_discovery.py:936-975 on already-loaded k{k}.pkl arrays (the input array is copied, the reference
    mutates its private unpickled copy)."""
    conseq = conseq.upper()
    assert all(e in "ACGT" for e in conseq)
    k = len(conseq)
    ckh = kmer2hash(conseq)
    rc_ckh = revcom_hash(ckh, k)
    if revcom_mode:
        assert ckh <= rc_ckh
    kh = np.array(uniq_kh_arr, copy=True)
    dist = cal_hamming_dist(kh, ckh, k)
    rc_flag = np.zeros(len(kh), dtype=bool)
    if revcom_mode:
        rc_dist = cal_hamming_dist(kh, rc_ckh, k)
        rc_flag = rc_dist < dist
        dist = np.minimum(dist, rc_dist)
    in_ball = dist <= max_ham_dist
    if revcom_mode:
        idx = np.where(rc_flag & in_ball)[0]
        kh[idx] = get_revcom_hash_arr(kh[idx], k)
    return kh[in_ball], np.asarray(uniq_kh_cnt_arr)[in_ball]


def cal_cnt_mat(uniq_kh_arr, uniq_kh_cnt_arr, kmer_len: int) -> np.ndarray:
    """int64[4, k]; row = base code, column = position from the 5' end (motif_discovery.py:978-986)."""
    cnt_mat = np.zeros((4, kmer_len), dtype=int)
    kh = np.asarray(uniq_kh_arr).astype(np.uint64)
    cnt = np.asarray(uniq_kh_cnt_arr).astype(np.int64)
    for pos in range(kmer_len):
        base = ((kh >> np.uint64(2 * (kmer_len - 1 - pos))) & np.uint64(3)).astype(np.int64)
        np.add.at(cnt_mat[:, pos], base, cnt)
    return cnt_mat


# ----------------------------------------------------------------------------------------------------------
# per-read motif occurrence  (motif_discovery.py:1396-1477, 1345-1393)
def get_motif_occurence(seq_np_arr, conseq_list: Sequence[str], motif_def_dict: dict, revcom_mode=True, rng=None):
    """One read (no separator).  Positions 0..L-k whose min(fwd, rc) distance is <= d, restricted to those at
    the read's minimum distance; more than 20 -> a random 20 (np.random in the reference; `rng` here).
    Reproduces the negative-slice quirk for L < k (motif_discovery.py:1447)."""
    locs_out, any_flag = [], False
    for conseq in conseq_list:
        k = len(conseq)
        d_max = motif_def_dict[k].max_ham_dist
        ckh = kmer2hash(conseq)
        rc_ckh = revcom_hash(ckh, k)
        hash_arr = comp_kmer_hash(seq_np_arr, k)
        hash_arr = hash_arr[0:(len(seq_np_arr) - k + 1)]
        dist = cal_hamming_dist(hash_arr, ckh, k)
        if revcom_mode:
            dist = np.minimum(dist, cal_hamming_dist(hash_arr, rc_ckh, k))
        locs = np.where(dist <= d_max)[0]
        if len(locs) == 0:
            locs_out.append("")
            continue
        locs = locs[dist[locs] == np.min(dist[locs])]
        if len(locs) > 20:
            pick = (rng or np.random).choice(len(locs), 20, replace=False)
            locs = np.sort(locs[pick])
        any_flag = True
        locs_out.append(",".join(map(str, locs)))
    return any_flag, ";".join(locs_out)


def motif_occurence_lines(reads: Sequence[str], conseq_list, motif_def_dict, revcom_mode=True, rng=None) -> List[str]:
    """Lines of a *.motif_occurence.csv (motif_discovery.py:1409-1418): header + one row per read with a hit."""
    lines = ["seq_ind;" + ";".join(f"motif_{i}_{c}" for i, c in enumerate(conseq_list)) + ";seq_len"]
    for i, r in enumerate(reads):
        arr = dna2arr(r.upper(), append_missing_val_flag=False)
        flag, s = get_motif_occurence(arr, conseq_list, motif_def_dict, revcom_mode, rng)
        if flag:
            lines.append(f"{i};{s};{len(arr)}")
    return lines


def get_motif_seq_num(lines: Sequence[str], motif_index: int) -> Tuple[int, int]:
    """(#reads with the motif, #listed positions) from occurrence-file lines (motif_discovery.py:1377-1393)."""
    n_reads = n_occ = 0
    for row in lines[1:]:
        cell = row.split(";")[motif_index + 1].strip()
        if cell == "":
            continue
        n_reads += 1
        n_occ += len(cell.split(","))
    return n_reads, n_occ


# ----------------------------------------------------------------------------------------------------------
# sampled k-mer distance matrix  (motif_discovery.py:705-808)
def convert_to_block_mat(uniq_dist_mat: np.ndarray, block_size_arr: np.ndarray) -> np.ndarray:
    """motif_discovery.py:705-730 (np.repeat on both axes == the reference's block slice assignment)."""
    assert np.issubdtype(block_size_arr.dtype, np.integer) and np.all(block_size_arr > 0)
    return np.repeat(np.repeat(uniq_dist_mat, block_size_arr, axis=0), block_size_arr, axis=1)


def convert_to_block_arr(arr: np.ndarray, block_size_arr: np.ndarray) -> np.ndarray:
    """motif_discovery.py:733-757"""
    assert np.issubdtype(block_size_arr.dtype, np.integer) and np.all(block_size_arr > 0)
    assert len(arr) == len(block_size_arr)
    return np.repeat(arr, block_size_arr)


def cal_samp_kmer_hamdist_mat(samp_kh_arr, samp_cnts, samp_label_arr, conseq_list, kmer_len, uniq_dist_flag=False):
    """motif_discovery.py:759-808: all-pairs distance at k; pairs sharing the label of a conseq shorter than k
    use only the first len(conseq) bases; expanded by samp_cnts unless uniq_dist_flag.  dtype int (int64)."""
    samp_kh_arr = np.asarray(samp_kh_arr)
    assert len(samp_kh_arr) == len(np.unique(samp_kh_arr))
    n = len(samp_kh_arr)
    for c in conseq_list:
        assert len(c) <= kmer_len
    hd = get_hash_dtype(kmer_len)
    kh = samp_kh_arr.astype(hd)
    mat = _popcount2bit(kh[:, None] ^ kh[None, :], kmer_len).astype(int)
    for li, conseq in enumerate(conseq_list):
        c = len(conseq)
        if c == kmer_len:
            continue
        idx = np.where(np.asarray(samp_label_arr) == li)[0]
        sub = np.right_shift(samp_kh_arr[idx], 2 * (kmer_len - c)).astype(get_hash_dtype(c))
        mat[np.ix_(idx, idx)] = _popcount2bit(sub[:, None] ^ sub[None, :], c).astype(int)
    if uniq_dist_flag:
        return mat
    return convert_to_block_mat(mat, np.asarray(samp_cnts))


# ----------------------------------------------------------------------------------------------------------
# k-mer labelling for the sample  (motif_discovery.py:812-903, deterministic part)
def label_kmers(uniq_kh_arr, conseq_list, kmer_len, motif_def_dict, revcom_mode=True):
    """Returns (aligned_kh_arr, label_arr): nearest-conseq label via head/tail partial distances, rc-closer
    members reverse-complemented (motif_discovery.py:849-892)."""
    kh = np.array(uniq_kh_arr, copy=True)
    n_conseq = len(conseq_list)
    dist_mat = np.zeros((n_conseq, len(kh)), dtype=int)
    rc_flag = np.zeros((n_conseq, len(kh)), dtype=bool)
    for i, conseq in enumerate(conseq_list):
        ckh = kmer2hash(conseq)
        d = cal_hamming_dist_head(kh, ckh, kmer_len, len(conseq))
        if revcom_mode:
            rc_ckh = revcom_hash(ckh, len(conseq))
            assert ckh <= rc_ckh
            rd = cal_hamming_dist_tail(kh, rc_ckh, kmer_len, len(conseq))
            rc_flag[i] = rd < d
            d = np.minimum(d, rd)
        dist_mat[i] = d
    for i, conseq in enumerate(conseq_list):
        dist_mat[i][dist_mat[i] > motif_def_dict[len(conseq)].max_ham_dist] = kmer_len
    min_d = np.min(dist_mat, axis=0)
    label = np.argmin(dist_mat, axis=0)
    label[min_d > motif_def_dict[kmer_len].max_ham_dist] = n_conseq
    if revcom_mode:
        for i in range(n_conseq):
            idx = np.where((label == i) & rc_flag[i])[0]
            kh[idx] = get_revcom_hash_arr(kh[idx], kmer_len)
    return kh, label


# ----------------------------------------------------------------------------------------------------------
# merge candidates of different k  (motif_discovery.py:533-591)  -- tiny string logic, host only
def merge_consensus_seqs(conseq_list: List[str]) -> List[str]:
    def shares_k_minus_1(long_s, short_s):
        return short_s[:-1] in long_s or short_s[1:] in long_s

    pool = sorted(conseq_list, key=len, reverse=True)
    final = []
    while pool:
        cur = pool[0]
        rc_cur = reverse_complement(cur)
        hit = lambda s: shares_k_minus_1(cur, s) or shares_k_minus_1(rc_cur, s)
        sub1 = next((s for s in pool if len(s) == len(cur) - 1 and hit(s)), None)
        sub2 = next((s for s in pool if len(s) == len(cur) - 2 and hit(s)), None)
        if sub1 and sub2:
            final.append(sub1)
            pool = [s for s in pool if not hit(s)]
        else:
            pool = pool[1:]
    return final
