"""CPU checks (oracle only) of the two table identities the device count relies on, so that the
algorithm -- not only its CUDA implementation -- is pinned to the reference's semantics (`comp_kmer_hash_taichi`
kmer_count.py:449-473, `count_uniq_hash` kmer_count.py:476-491, `remove_duplicate_hash_per_seq` kmer_count.py:743-760):

  1. T_k = (4:1 reduction of T_(k+1) over the LAST base) + C_k [- the repeats whose extension is new, in de-duplicating mode],
     with C_k[h] = number of valid runs of at least k bases whose last k bases are h     (csrc/count_all.cu)
  2. C_k = (4:1 reduction of C_(k+1) over the FIRST base) + the runs of exactly k bases  (fold_corrections_kernel, csrc/partition.cu)
"""
import numpy as np
import pytest

from oracle import kmap_oracle as O


def _dense(seq, borders, k, dedup):
    h = O.comp_kmer_hash(seq, k)
    if dedup:
        h = O.remove_duplicate_hash_per_seq(h, borders, O.get_invalid_hash(O.get_hash_dtype(k)))
    kh, cnt = O.count_uniq_hash(h, k)
    t = np.zeros(4 ** k, dtype=np.int64)
    t[kh.astype(np.int64)] = cnt
    return t


def _run_end_tables(seq, kmin, kmax):
    """C[v][h] for kmin <= v <= kmax and E[v][h] = the same for runs of exactly v bases"""
    C = {v: np.zeros(4 ** v, dtype=np.int64) for v in range(kmin, kmax + 1)}
    E = {v: np.zeros(4 ** v, dtype=np.int64) for v in range(kmin, kmax + 1)}
    valid = np.concatenate([seq != 255, [False]])
    start = None
    for i, ok in enumerate(valid):
        if ok and start is None:
            start = i
        if not ok and start is not None:
            length = i - start
            for v in range(kmin, min(length, kmax) + 1):
                h = 0
                for b in seq[i - v:i]:
                    h = h * 4 + int(b)
                C[v][h] += 1
                if v == length:
                    E[v][h] += 1
            start = None
    return C, E


def _reads(seed):
    rng = np.random.default_rng(seed)
    reads = ["A" * 30, "CA" * 20, "", "N", "ACG", "ACGTTGCA" * 6, "GGGGGGGGGGGGGGGGGGGGAGGGGGGGGGGGGGGGGGGG", "ACGTNNACGTACNACGTACGT"]
    for _ in range(200):
        n = int(rng.integers(0, 40))
        r = rng.integers(0, 4, n)
        s = "".join("ACGT"[x] for x in r)
        if n and rng.random() < 0.3:
            p = int(rng.integers(0, n))
            s = s[:p] + "N" + s[p + 1:]
        reads.append(s)
    arrs = [O.dna2arr(r) for r in reads]
    lens = np.array([len(a) for a in arrs])
    ends = np.cumsum(lens)
    return np.concatenate(arrs), np.stack([ends - lens, ends - 1], axis=1).astype(np.int64)


@pytest.mark.parametrize("seed", [1, 2])
def test_tables_of_neighbouring_levels_rep_mode(seed):
    seq, borders = _reads(seed)
    kmin, kmax = 2, 7
    C, _ = _run_end_tables(seq, kmin, kmax)
    T = {k: _dense(seq, borders, k, dedup=False) for k in range(kmin, kmax + 1)}
    for k in range(kmin, kmax):
        assert np.array_equal(T[k], T[k + 1].reshape(-1, 4).sum(axis=1) + C[k]), k


@pytest.mark.parametrize("seed", [1, 2])
def test_run_end_tables_fold_over_the_first_base(seed):
    seq, _ = _reads(seed)
    kmin, kmax = 2, 7
    C, E = _run_end_tables(seq, kmin, kmax)
    for v in range(kmin, kmax):
        assert np.array_equal(C[v], C[v + 1].reshape(4, -1).sum(axis=0) + E[v]), v
    assert sum(int(E[v].sum()) for v in E) > 0 and int(C[kmax].sum()) > 0          # both kinds of run are present


def test_dedup_tables_differ_from_the_fold_only_where_kmers_repeat_inside_a_read():
    """de-duplicating mode: T_k - fold(T_(k+1)) - C_k = -(windows whose k-mer repeats an earlier window of the read while
    their (k+1)-mer, if there is one, is new) + 0 elsewhere: never positive, and zero for reads without repeats"""
    seq, borders = _reads(3)
    kmin, kmax = 2, 7
    C, _ = _run_end_tables(seq, kmin, kmax)
    T = {k: _dense(seq, borders, k, dedup=True) for k in range(kmin, kmax + 1)}
    some = False
    for k in range(kmin, kmax):
        d = T[k] - T[k + 1].reshape(-1, 4).sum(axis=1) - C[k]
        assert (d <= 0).all(), k
        some = some or bool((d < 0).any())
    assert some
