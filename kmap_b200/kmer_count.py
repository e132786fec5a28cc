"""Drop-in for the hot-path half of the reference's `kmap/kmer_count.py`: same function names, argument order,
dtypes and in-place / return conventions, bodies replaced by calls into libkmap_b200
(hand-written sm_100a CUDA) through ctypes.  NumPy in, NumPy out; every function raises if no GPU / library.

Reference line numbers cited below are in /root/reference/src/kmap/kmer_count.py.
"""
from __future__ import annotations

import pickle
from dataclasses import dataclass, fields
from pathlib import Path
from typing import Dict, List, Tuple

import numpy as np
import torch

from . import engine as E
from ._lib import KmapError, check, lib

MISSING_VAL = 255  # :58

FileNameDict = {  # :26-53 (file layout of a result directory; kept identical)
    "default_config_file": "default_config.toml",
    "config_file": "config.toml",
    "default_motif_def_file": "default_motif_def_table.csv",
    "motif_def_file": "motif_def_table.csv",
    "processed_fasta_file": "input.bin.pkl",
    "processed_fasta_seqboarder_file": "input.seqboarder.bin.pkl",
    "motif_pos_density_file": "motif_pos_density.np.pkl",
    "motif_pos_density_plot_dir": "motif_pos_density",
    "kmer_count_dir": "kmer_count",
    "conseq_similarity_dir": "conseq_similarity",
    "co_occur_dir": "co_occurence",
    "co_occur_dist_mat_file": "co_occurence_motif_dist_mat.tsv",
    "co_occur_dist_data_file": "co_occurence_motif_dist_data.txt",
    "co_occur_mat_file": "co_occurence_mat.tsv",
    "co_occur_mat_norm_file": "co_occurence_mat.norm.tsv",
    "co_occur_network_fig": "co_occur_network.pdf",
    "motif_occurence_file": "final.motif_occurence.csv",
    "hamball_dir": "hamming_balls",
    "candidate_conseq_file": "candidate_conseq.csv",
    "final_conseq_file": "final_conseq.txt",
    "final_conseq_info_file": "final_conseq.info.csv",
    "sample_kmer_pkl_file": "sample_kmers.pkl",
    "sample_kmer_txt_file": "sample_kmers.tsv",
    "sample_kmer_hamdist_mat_file": "sample_kmer_hamdist_mat.pkl",
    "ld_data_file": "low_dim_data.tsv",
    "ld_fig_file_stem": "ld_data",
}

_PKG_DIR = Path(__file__).resolve().parent


# ---- dtype rules (:351-370) ---------------------------------------------------------------------------------
def get_cnt_dtype(kmer_len: int):
    return np.int32 if kmer_len < 16 else np.int64


def get_hash_dtype(kmer_len):
    if 0 < kmer_len < 16:
        return np.uint32
    elif kmer_len < 32:
        return np.uint64
    raise Exception(f"max_kmer_len=31, kmer_len={kmer_len} is greater the maximum value.")


def get_invalid_hash(dtype):
    return dtype(np.iinfo(dtype).max)


# ---- host scalars / strings (:238-268, 416-446, 626-640) ----------------------------------------------------------
_ENCODE = np.full(256, MISSING_VAL, dtype=np.uint8)
_ENCODE[np.frombuffer(b"ACGT", dtype=np.uint8)] = np.arange(4, dtype=np.uint8)
_DECODE = np.full(256, ord("?"), dtype=np.uint8)
_DECODE[[0, 1, 2, 3, MISSING_VAL]] = np.frombuffer(b"ACGTN", dtype=np.uint8)


def dna2arr(dna_str, dtype=np.uint8, append_missing_val_flag=True) -> np.ndarray:
    codes = _ENCODE[np.frombuffer(dna_str.encode("latin-1", "replace"), dtype=np.uint8)]
    res = np.empty(len(codes) + (1 if append_missing_val_flag else 0), dtype=dtype)
    res[:len(codes)] = codes
    if append_missing_val_flag:
        res[-1] = MISSING_VAL
    return res


def arr2dna(dna_np_arr: np.ndarray) -> str:
    return _DECODE[np.asarray(dna_np_arr, dtype=np.uint8)].tobytes().decode()


def reverse_complement(seq):
    return seq[::-1].translate(str.maketrans("ACGT", "TGCA")) if set(seq) <= set("ACGT") else \
        "".join({"A": "T", "T": "A", "C": "G", "G": "C"}[b] for b in reversed(seq))


def kmer2hash(kmer: str) -> np.uint64:
    assert len(kmer) < 32, "kmer should be shorted than 32 bases"
    value = 0
    for base in kmer:
        value = value * 4 + {"A": 0, "C": 1, "G": 2, "T": 3}[base]
    return np.uint64(value)


def hash2kmer(hashkey, k: int) -> str:
    value = int(hashkey)
    return "".join("ACGT"[(value >> (2 * (k - 1 - j))) & 3] for j in range(k))


def revcom_hash(in_hash, kmer_len: int):
    """scalar reverse complement in the hash dtype (host arithmetic; :626-640)"""
    hd = get_hash_dtype(kmer_len)
    width = np.iinfo(hd).bits
    comp = (((1 << (2 * kmer_len)) - 1) - int(hd(in_hash))) % (1 << width)
    out = 0
    for _ in range(kmer_len):
        out = ((out << 2) | (comp & 3)) % (1 << width)
        comp >>= 2
    return hd(out)


# ---- array primitives: the ten integer Taichi kernels ---------------------------------------------------------------
def _stream():
    return torch.cuda.current_stream().cuda_stream


def comp_kmer_hash_taichi(seq_np_arr: np.ndarray, kmer_len: int) -> np.ndarray:
    """:449-473.  One hash per array position (separators included); invalid = all ones.  Name kept for drop-in use."""
    hd = get_hash_dtype(kmer_len)
    seq_np_arr = np.asarray(seq_np_arr)
    if seq_np_arr.dtype != np.uint8:
        raise KmapError("seq_np_arr must be uint8")
    n = len(seq_np_arr)
    seq_d = E.to_device(seq_np_arr)
    if hd == np.uint32:
        out = E.empty(n, torch.int32)
        check(lib().kmap_kmer2hash_u32(seq_d.data_ptr(), n, kmer_len, out.data_ptr(), _stream()), "kmap_kmer2hash_u32")
    else:
        out = E.empty(n, torch.int64)
        check(lib().kmap_kmer2hash_u64(seq_d.data_ptr(), n, kmer_len, out.data_ptr(), _stream()), "kmap_kmer2hash_u64")
    return E.to_host(out, hd)


comp_kmer_hash = comp_kmer_hash_taichi


def _dist_call(kh_arr, consensus_kh, kmer_len, consensus_len, which):
    hd = get_hash_dtype(kmer_len)
    kh_arr = np.asarray(kh_arr)
    if kh_arr.dtype != hd:
        kh_arr = kh_arr.astype(hd)
    n = len(kh_arr)
    target = int(np.array([consensus_kh]).astype(hd)[0])
    out = E.empty(n, torch.uint8)
    if n:
        kh_d = E.to_device(kh_arr)
        sfx = "u32" if hd == np.uint32 else "u64"
        L = lib()
        if which == "full":
            rc = getattr(L, f"kmap_ham_dist_{sfx}")(kh_d.data_ptr(), n, target, kmer_len, out.data_ptr(), _stream())
        else:
            rc = getattr(L, f"kmap_ham_dist_{which}_{sfx}")(kh_d.data_ptr(), n, target, kmer_len, consensus_len,
                                                            out.data_ptr(), _stream())
        check(rc, f"kmap_ham_dist_{which}")
    return out.cpu().numpy()


def cal_hamming_dist(kh_arr: np.ndarray, consensus_kh, kmer_len: int) -> np.ndarray:
    """:494-515 (works on empty arrays)"""
    return _dist_call(kh_arr, consensus_kh, kmer_len, kmer_len, "full")


def cal_hamming_dist_head(kh_arr, consensus_kh, kmer_len: int, consensus_len: int) -> np.ndarray:
    """:518-546"""
    assert consensus_len <= kmer_len
    return _dist_call(kh_arr, consensus_kh, kmer_len, consensus_len, "head")


def cal_hamming_dist_tail(kh_arr, consensus_kh, kmer_len: int, consensus_len: int) -> np.ndarray:
    """:549-577"""
    assert consensus_len <= kmer_len
    return _dist_call(kh_arr, consensus_kh, kmer_len, consensus_len, "tail")


def get_revcom_hash_arr(in_hash_arr: np.ndarray, kmer_len: int) -> np.ndarray:
    """:613-623; the result has the dtype of the input array (np.empty_like in the reference)"""
    hd = get_hash_dtype(kmer_len)
    src = np.asarray(in_hash_arr)
    a = src.astype(hd, copy=False)
    n = len(a)
    if n == 0:
        return np.empty_like(src)
    a_d = E.to_device(a)
    out = torch.empty_like(a_d)
    fn = lib().kmap_revcom_u32 if hd == np.uint32 else lib().kmap_revcom_u64
    check(fn(a_d.data_ptr(), n, kmer_len, out.data_ptr(), _stream()), "kmap_revcom")
    return E.to_host(out, hd).astype(src.dtype, copy=False)


def remove_duplicate_hash_per_seq(hash_arr: np.ndarray, boarder_mat: np.ndarray, invalid_hash) -> np.ndarray:
    """:743-760.  In place + returned: inside each read keep the first occurrence of every hash."""
    assert boarder_mat.shape[1] == 2
    if hash_arr.dtype not in (np.dtype(np.uint32), np.dtype(np.uint64)):
        raise KmapError("remove_duplicate_hash_per_seq expects the uint32 / uint64 hash array of comp_kmer_hash_taichi")
    n = len(hash_arr)
    if n == 0 or len(boarder_mat) == 0:
        return hash_arr
    h_d = E.to_device(hash_arr)
    b_d = E.to_device(np.ascontiguousarray(boarder_mat, dtype=np.int64))
    if hash_arr.dtype == np.uint64:          # k >= 16 (csrc/sorted.cu)
        work = E.empty(lib().kmap_dedup_keys_work_words(len(boarder_mat)), torch.int32)
        check(lib().kmap_dedup_hash_per_read_u64(h_d.data_ptr(), n, b_d.data_ptr(), len(boarder_mat), work.data_ptr(), _stream()),
              "kmap_dedup_hash_per_read_u64")
        hash_arr[:] = E.to_host(h_d, np.uint64)
        return hash_arr
    check(lib().kmap_dedup_hash_per_read_u32(h_d.data_ptr(), n, b_d.data_ptr(), len(boarder_mat), _stream()),
          "kmap_dedup_hash_per_read_u32")
    hash_arr[:] = E.to_host(h_d, np.uint32)
    return hash_arr


def _require_dense(kmer_len, what):
    if not 1 <= kmer_len <= 15:
        raise KmapError(f"{what}: dense tables cover 1 <= k <= 15 and the sort path 16 <= k <= 31 (got {kmer_len})")


def count_uniq_hash(hash_arr: np.ndarray, kmer_len):
    """:476-491.  (ascending unique hashes without the invalid hash, counts in the count dtype)"""
    hash_arr = np.asarray(hash_arr)
    if 16 <= kmer_len <= 31:                 # no dense table: sort + run-length encoding (csrc/sorted.cu)
        if hash_arr.dtype != np.uint64:
            raise KmapError("count_uniq_hash expects the uint64 hash array of comp_kmer_hash_taichi for k >= 16")
        kh, cnt = E.sort_count_keys(E.to_device(hash_arr.copy()), 2 * kmer_len)
        return E.to_host(kh, np.uint64), E.to_host(cnt, np.int64).astype(get_cnt_dtype(kmer_len), copy=False)
    _require_dense(kmer_len, "count_uniq_hash")
    if hash_arr.dtype != np.uint32:
        raise KmapError("count_uniq_hash expects the uint32 hash array of comp_kmer_hash_taichi")
    table = E.zeros(1 << (2 * kmer_len), torch.int32)
    h_d = E.to_device(hash_arr)
    check(lib().kmap_count_hashes_u32(h_d.data_ptr(), len(hash_arr), kmer_len, table.data_ptr(), _stream()), "kmap_count_hashes_u32")
    kh, cnt = E.compact_merge(table, kmer_len, revcom=False)
    return E.to_host(kh, np.uint32), E.to_host(cnt, np.int32).astype(get_cnt_dtype(kmer_len), copy=False)


def merge_revcom(uniq_kmer_hash_arr: np.ndarray, uniq_kh_cnt_arr: np.ndarray, kmer_len: int, keep_lower_hash_flag=True) -> Tuple:
    """:643-685.  Sums the counts of reverse-complement pairs (a palindrome is its own partner: doubled), keeps the
    lower hash of a pair (the higher one with keep_lower_hash_flag=False), relabels lone k-mers to min (max) of (h, rc h);
    result order = ascending forward hash of the survivors.
    Like the reference it also updates the caller's count array in place (the `+=` at :661)."""
    kh = np.asarray(uniq_kmer_hash_arr)
    n = len(kh)
    if n == 0:
        return kh.copy(), np.asarray(uniq_kh_cnt_arr).copy()
    if 16 <= kmer_len <= 31:                 # sorted-list formulation (csrc/sorted.cu)
        if n > 1 and not np.all(kh[1:] > kh[:-1]):
            raise KmapError("merge_revcom for k >= 16 expects the ascending unique hashes count_uniq_hash returns")
        out_kh, out_cnt, summed = E.merge_revcom_sorted(E.to_device(kh.astype(np.uint64, copy=False)),
                                                        E.to_device(np.asarray(uniq_kh_cnt_arr).astype(np.int64, copy=False)),
                                                        kmer_len, want_summed=True, keep_higher=not keep_lower_hash_flag)
        uniq_kh_cnt_arr[:] = E.to_host(summed, np.int64).astype(uniq_kh_cnt_arr.dtype, copy=False)
        return (E.to_host(out_kh, np.uint64).astype(kh.dtype, copy=False),
                E.to_host(out_cnt, np.int64).astype(uniq_kh_cnt_arr.dtype, copy=False))
    _require_dense(kmer_len, "merge_revcom")
    L = lib()
    kh_d = E.to_device(kh.astype(np.uint32, copy=False))
    cnt_d = E.to_device(np.asarray(uniq_kh_cnt_arr).astype(np.int32, copy=False))
    table = E.zeros(1 << (2 * kmer_len), torch.int32)
    check(L.kmap_scatter_counts(kh_d.data_ptr(), cnt_d.data_ptr(), n, kmer_len, table.data_ptr(), _stream()), "kmap_scatter_counts")
    out_kh, out_cnt = E.compact_merge(table, kmer_len, revcom=1 if keep_lower_hash_flag else 2)
    check(L.kmap_list_add_rc_counts(kh_d.data_ptr(), cnt_d.data_ptr(), n, kmer_len, table.data_ptr(), _stream()),
          "kmap_list_add_rc_counts")
    uniq_kh_cnt_arr[:] = E.to_host(cnt_d, np.int32).astype(uniq_kh_cnt_arr.dtype, copy=False)
    return (E.to_host(out_kh, np.uint32).astype(kh.dtype, copy=False),
            E.to_host(out_cnt, np.int32).astype(uniq_kh_cnt_arr.dtype, copy=False))


def mask_input(seq_np_arr: np.ndarray, kmer_len: int, consensus_kh_arr: np.ndarray, max_hamball_dist_arr: np.ndarray):
    """:580-610.  Mutates seq_np_arr in place and returns it.  Windows are compared on the PRE-mask array for every
    consensus; invalid windows behave like T..T (the reference compares their all-ones hash)."""
    if not 1 <= kmer_len <= 31:
        raise KmapError(f"mask_input: 1 <= k <= 31 (got {kmer_len})")
    if len(seq_np_arr) == 0 or len(consensus_kh_arr) == 0:
        return seq_np_arr
    dev = E.SeqOnDevice.from_numpy(seq_np_arr, None, keep_u8=True)
    dev.mask(kmer_len, [int(c) for c in consensus_kh_arr], [int(d) for d in max_hamball_dist_arr])
    return dev.masked_seq_to_numpy(seq_np_arr)


def mask_ham_ball(seq_np_arr: np.ndarray, motif_def_dict: dict, consensus_seq_list: List[str],
                  max_ham_dist_list: List[int] = ()) -> np.ndarray:
    """:688-723"""
    len_list = np.array([len(conseq) for conseq in consensus_seq_list])
    if len(max_ham_dist_list) == 0:
        max_ham_dist_list = [motif_def_dict[int(n)].max_ham_dist for n in len_list]
    assert len(max_ham_dist_list) == len(consensus_seq_list)
    for uniq_len in np.unique(len_list):
        inds = np.where(len_list == uniq_len)[0]
        khs = np.array([kmer2hash(consensus_seq_list[i]) for i in inds])
        ds = np.array([max_ham_dist_list[i] for i in inds])
        seq_np_arr = mask_input(seq_np_arr, int(uniq_len), khs, ds)
    return seq_np_arr


# ---- motif definition table and config (:104-136, 221-235, 726-740) ------------------------------------------------
@dataclass
class MotifDef:
    kmer_len: int
    p_uniform: float
    max_ham_dist: int
    ratio_mu: float
    ratio_std: float
    ratio_cutoff: float

    @classmethod
    def get_field_names(cls):
        return ",".join(f.name for f in fields(cls))

    def __str__(self):
        return ",".join(str(getattr(self, f.name)) for f in fields(self))


def init_motif_def_dict(motif_def_file, p_value_cutoff=1e-10) -> dict:
    import pandas as pd
    from scipy.stats import norm
    table = {"p_value_cutoff": p_value_cutoff}
    for _, row in pd.read_csv(motif_def_file).iterrows():
        k = int(row["kmer_len"])
        cutoff = norm.ppf(1 - p_value_cutoff, loc=row["ratio_mu"], scale=row["ratio_std"])
        table[k] = MotifDef(k, row["p_uniform"], int(row["max_ham_dist"]), row["ratio_mu"], row["ratio_std"], cutoff)
    return table


def read_default_config_file(debug=False):
    import tomllib
    with open(_PKG_DIR / FileNameDict["default_config_file"], "rb") as fh:
        cfg = tomllib.load(fh)
    if debug:
        print(cfg)
    return cfg


def gen_motif_def_dict(config_dict: dict, debug=False) -> Dict:
    motif_def_file = config_dict["motif_discovery"]["motif_def_file"]
    if motif_def_file == "default":
        motif_def_file = _PKG_DIR / FileNameDict["default_motif_def_file"]
    else:
        assert Path(motif_def_file).exists()
    out = init_motif_def_dict(motif_def_file, p_value_cutoff=config_dict["motif_discovery"]["p_value_cutoff"])
    if debug:
        print(out)
    return out


# ---- preproc: FASTA -> input.bin.pkl / input.seqboarder.bin.pkl (:139-347) -----------------------------------------
def fasta_to_arrays(fasta_file) -> Tuple[np.ndarray, np.ndarray]:
    """the two arrays preproc pickles (:326-347): uint8 codes with a 255 after every read; int borders [start, separator].
    Parsed and encoded on the device (csrc/fasta.cu) instead of two Bio.SeqIO passes with a per-base Python loop."""
    seq, borders = E.fasta_to_device(fasta_file)
    return E.to_host(seq, np.uint8), E.to_host(borders, np.int64).reshape(-1, 2).astype(int, copy=False)


def proc_input(input_fasta_file: str, res_dir=".", out_bin_file_name: str = "input.bin.pkl",
               out_boarder_bin_file_name: str = "input.seqboarder.bin.pkl", debug=True):
    """:182-218"""
    assert Path(input_fasta_file).exists()
    assert Path(res_dir).exists()
    assert out_bin_file_name.endswith(".pkl")
    seq, borders = fasta_to_arrays(input_fasta_file)
    input_binary_file = str(Path(res_dir) / out_bin_file_name)
    if debug:
        print(f"Convert input file={input_fasta_file} into binary file {input_binary_file}. buffer_size={len(seq)/2**30}GB.")
    with open(input_binary_file, "wb") as fh:
        pickle.dump(seq, fh)
    with open(Path(res_dir) / out_boarder_bin_file_name, "wb") as fh:
        pickle.dump(borders, fh)
    print(f"input binary file {input_binary_file} generated.\n")


def _preproc(fasta_file: str, res_dir=".", debug=False):
    """:139-179"""
    import tomllib
    import tomli_w
    assert Path(fasta_file).exists()
    Path(res_dir).mkdir(exist_ok=True)
    config_file_path = Path(res_dir) / FileNameDict["config_file"]
    if config_file_path.exists():
        with open(config_file_path, "rb") as fh:
            config_dict = tomllib.load(fh)
    else:
        config_dict = read_default_config_file(debug=debug)
    if not config_file_path.exists() or config_dict["general"].get("input_fasta_file") is None:
        config_dict["general"]["input_fasta_file"] = fasta_file
        config_dict["general"]["res_dir"] = res_dir
        with open(config_file_path, "wb") as fh:
            tomli_w.dump(config_dict, fh)
    motif_def_dict = gen_motif_def_dict(config_dict, debug=debug)
    kmer_len_list = sorted(e for e in motif_def_dict if isinstance(e, int))
    with open(Path(res_dir) / FileNameDict["motif_def_file"], "w+") as fh:
        fh.write(MotifDef.get_field_names() + "\n")
        for k in kmer_len_list:
            fh.write(str(motif_def_dict[k]) + "\n")
    proc_input(config_dict["general"]["input_fasta_file"], config_dict["general"]["res_dir"],
               out_bin_file_name=FileNameDict["processed_fasta_file"],
               out_boarder_bin_file_name=FileNameDict["processed_fasta_seqboarder_file"], debug=debug)
    return config_dict, motif_def_dict
