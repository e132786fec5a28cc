"""Parity of the CUDA path (through the C ABI / the reference-shaped Python API) with the oracle and with the
golden outputs of the unmodified reference.  Everything here is integer work: the bar is bit-exact.
Run on the B200 box:  python -m pytest tests -m gpu -x -q
"""
import numpy as np
import pytest

from pathlib import Path

from helpers import assert_occurrence_text_equal
from oracle import kmap_oracle as O

GOLDEN_DIR = Path(__file__).resolve().parent / "golden"

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def K():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import kmap_b200.kmer_count as kc
    return kc


@pytest.fixture(scope="module")
def MD(K):
    import kmap_b200.motif_discovery as md
    return md


@pytest.fixture(scope="module")
def ENG(K):
    import kmap_b200.engine as E
    return E


def rand_reads(rng, n_reads, lmin, lmax, p_n=0.0, special=()):
    reads = list(special)
    for _ in range(n_reads):
        L = int(rng.integers(lmin, lmax + 1))
        s = rng.integers(0, 4, L).astype(np.uint8)
        if p_n > 0:
            s[rng.random(L) < p_n] = 255
        reads.append(O.arr2dna(s))
    arrs = [O.dna2arr(r) for r in reads]
    seq = np.concatenate(arrs)
    lens = np.array([len(a) for a in arrs])
    ends = np.cumsum(lens)
    borders = np.stack([ends - lens, ends - 1], axis=1).astype(np.int64)
    return reads, seq, borders


def dense_table_from_oracle(seq, borders, k, dedup):
    h = O.comp_kmer_hash(seq, k)
    if dedup:
        h = O.remove_duplicate_hash_per_seq(h, borders, np.uint32(0xFFFFFFFF))
    u, c = O.count_uniq_hash(h, k)
    t = np.zeros(4 ** k, dtype=np.uint32)
    t[u] = c
    return t


# ---------------------------------------------------------------------------------------------------------------
def test_primitives_match_reference_vectors(K, unit_vectors):
    for g in unit_vectors["primitives"]:
        k = g["k"]
        h = K.comp_kmer_hash_taichi(g["seq"], k)
        assert h.dtype == g["hash"].dtype and np.array_equal(h, g["hash"]), k
        assert np.array_equal(K.cal_hamming_dist(g["kh"], g["target"][0], k), g["dist"]), k
        rc = K.get_revcom_hash_arr(g["revcom_in"], k)
        assert rc.dtype == g["revcom"].dtype and np.array_equal(rc, g["revcom"]), k
        assert np.array_equal(np.array([K.revcom_hash(x, k) for x in g["revcom_in"][:20]]), g["revcom_scalar"])
        for key in g:
            if key.startswith("head_"):
                cl = int(key[5:])
                ct = g[f"ct_{cl}"][0]
                assert np.array_equal(K.cal_hamming_dist_head(g["kh"], ct, k, cl), g[key]), (k, cl)
                assert np.array_equal(K.cal_hamming_dist_tail(g["kh"], ct, k, cl), g[f"tail_{cl}"]), (k, cl)
        if k < 16:
            dd = K.remove_duplicate_hash_per_seq(g["hash"].copy(), g["borders"], np.uint32(0xFFFFFFFF))
            assert np.array_equal(dd, g["dedup"]), k
    assert len(K.cal_hamming_dist(np.zeros(0, dtype=np.uint32), 5, 8)) == 0        # empty input (md:786 last row)


def test_primitives_random_vs_oracle(K):
    rng = np.random.default_rng(5)
    reads, seq, borders = rand_reads(rng, 300, 0, 90, p_n=0.02, special=["A" * 300, "ACGT" * 700, "", "N"])
    for k in (1, 2, 7, 12, 15, 16, 23, 31):
        assert np.array_equal(K.comp_kmer_hash_taichi(seq, k), O.comp_kmer_hash(seq, k)), k
    for k in (8, 15):
        h = O.comp_kmer_hash(seq, k)
        want = O.remove_duplicate_hash_per_seq(h.copy(), borders, np.uint32(0xFFFFFFFF))
        assert np.array_equal(K.remove_duplicate_hash_per_seq(h.copy(), borders, np.uint32(0xFFFFFFFF)), want), k
    for k, hd in ((14, np.uint32), (27, np.uint64)):
        kh = rng.integers(0, 4 ** k, 5000, dtype=np.uint64).astype(hd)
        t = hd(rng.integers(0, 4 ** k, dtype=np.uint64))
        assert np.array_equal(K.cal_hamming_dist(kh, t, k), O.cal_hamming_dist(kh, t, k))
        assert np.array_equal(K.get_revcom_hash_arr(kh, k), O.get_revcom_hash_arr(kh, k))


def test_dedup_low_complexity_reads(K, unit_vectors):
    g = unit_vectors["dedup_lowcomplex"]
    h = K.comp_kmer_hash_taichi(g["seq"], 8)
    assert np.array_equal(h, g["hash"])
    assert np.array_equal(K.remove_duplicate_hash_per_seq(h, g["borders"], np.uint32(0xFFFFFFFF)), g["dedup"])


def test_count_k3_known_answer(K, unit_vectors):
    g = unit_vectors["count_k3"]
    h = K.comp_kmer_hash_taichi(K.dna2arr(g["seq"]), 3)
    u, c = K.count_uniq_hash(h, 3)
    assert np.array_equal(u, g["uniq"]) and np.array_equal(c, g["cnt"]) and c.dtype == g["cnt"].dtype and c.sum() == 96


def test_merge_revcom_vectors(K, unit_vectors):
    for g in unit_vectors["merge_revcom"]:
        kh, cnt = g["kh"].copy(), g["cnt"].copy()
        mk, mc = K.merge_revcom(kh, cnt, g["k"])
        assert np.array_equal(mk, g["out_kh"]) and np.array_equal(mc, g["out_cnt"]), g["k"]
        assert mk.dtype == g["out_kh"].dtype and mc.dtype == g["out_cnt"].dtype
        if "mutated_cnt" in g:
            assert np.array_equal(cnt, g["mutated_cnt"])


def test_merge_revcom_keep_higher_flag(K, r2_vectors):
    """keep_lower_hash_flag=False (kmer_count.py:668-683) on the dense (k < 16) and the sorted-list (k >= 16) path"""
    for g in r2_vectors["merge_revcom_flag"]:
        kh, cnt = g["kh"].copy(), g["cnt"].copy()
        mk, mc = K.merge_revcom(kh, cnt, g["k"], keep_lower_hash_flag=g["keep_lower"])
        assert np.array_equal(mk, g["out_kh"]) and np.array_equal(mc, g["out_cnt"]), (g["k"], g["keep_lower"])
        assert mk.dtype == g["out_kh"].dtype and mc.dtype == g["out_cnt"].dtype
        assert np.array_equal(cnt, g["mutated_cnt"]), (g["k"], g["keep_lower"])


def test_mask_known_answers(K, unit_vectors, motif_def_file):
    g = unit_vectors["mask_ham_ball"]
    mdd = K.init_motif_def_dict(motif_def_file)
    assert K.arr2dna(K.mask_ham_ball(K.dna2arr(g["s1"])[:-1], mdd, g["c1"], g["d1"])) == g["r1"]
    assert K.arr2dna(K.mask_ham_ball(K.dna2arr(g["s2"])[:-1], mdd, g["c2"])) == g["r2"]
    for c in unit_vectors["mask_input"]:
        k, d = c["k"], c["d"]
        kh = K.kmer2hash(c["conseq"])
        a = c["before"].copy()
        out = K.mask_input(a, k, np.array([kh, K.revcom_hash(kh, k)]), np.array([d, d]))
        assert out is a and np.array_equal(a, c["after"]), c["conseq"]


@pytest.mark.parametrize("k", [1, 4, 8, 11, 14, 15])
def test_fused_count_tables_vs_oracle(ENG, k):
    rng = np.random.default_rng(100 + k)
    # short reads (warp path), 300-7000 bp reads (block path), one 30 kb read (bitmap path), repeats, N's, empties
    special = ["A" * 80, "CA" * 60, "", "N", "ACG", "ACGTTGCA" * 150, ("ACGTAGCTAGCTAGGATCGAT" * 1500)[:30000]]
    reads, seq, borders = rand_reads(rng, 400, 0, 120, p_n=0.01, special=special)
    reads2, seq2, borders2 = rand_reads(rng, 6, 300, 7000, p_n=0.001)
    seq = np.concatenate([seq, seq2])
    borders = np.concatenate([borders, borders2 + borders[-1, 1] + 1])
    dev = ENG.SeqOnDevice.from_numpy(seq, borders)
    for dedup in (False, True):
        got = ENG.to_host(dev.count(k, dedup=dedup), np.uint32)
        want = dense_table_from_oracle(seq, borders, k, dedup)
        assert np.array_equal(got, want), (k, dedup, int(np.abs(got.astype(np.int64) - want).sum()))


@pytest.mark.parametrize("k", [9, 11, 12, 13, 14])
def test_partitioned_count_equals_direct_count(ENG, k):
    """key partitioning + shared-memory counters (csrc/partition.cu) == one global atomic per window == oracle; the input
    spans several partition tiles, has N's, and a homopolymer long enough to fold the 16-bit shared counters (>= 32768)"""
    rng = np.random.default_rng(77 + k)
    special = ["A" * 70000, "CA" * 60, "", "N", "ACG", "T" * 40000 + "G" + "T" * 33000]
    reads, seq, borders = rand_reads(rng, 3000, 0, 150, p_n=0.01, special=special)
    dev = ENG.SeqOnDevice.from_numpy(seq, borders)
    got = ENG.to_host(dev.count(k, dedup=False, partitioned=True, scheme=ENG.SeqOnDevice.SORTED), np.uint32)
    direct = ENG.to_host(dev.count(k, dedup=False, partitioned=False), np.uint32)
    assert np.array_equal(got, direct), (k, int(np.abs(got.astype(np.int64) - direct.astype(np.int64)).sum()))
    assert np.array_equal(got, dense_table_from_oracle(seq, borders, k, False))
    assert got[0] >= 69000 and got[4 ** k - 1] >= 39000
    # a second call re-uses the scratch and must not depend on its previous content
    again = ENG.to_host(dev.count(k, dedup=False, partitioned=True), np.uint32)
    assert np.array_equal(again, got)


def test_partitioned_count_tiny_and_empty(ENG):
    for reads in (["ACGTACGTACGTACG"], ["ACGT"], ["N" * 50], ["ACGTTGCAACGTTGCAAC", "", "GGGGGGGGGGGGGGGGGGGG"]):
        arrs = [O.dna2arr(r) for r in reads]
        seq = np.concatenate(arrs)
        lens = np.array([len(a) for a in arrs])
        ends = np.cumsum(lens)
        borders = np.stack([ends - lens, ends - 1], axis=1).astype(np.int64)
        dev = ENG.SeqOnDevice.from_numpy(seq, borders)
        for k in (9, 14):
            got = ENG.to_host(dev.count(k, dedup=False, partitioned=True), np.uint32)
            assert np.array_equal(got, dense_table_from_oracle(seq, borders, k, False))


@pytest.mark.parametrize("kmin,kmax,parts", [(8, 14, 0), (8, 14, 3), (1, 6, 0), (5, 5, 1), (11, 15, 5), (3, 9, 2), (9, 13, 0),
                                             (12, 12, 0)])
def test_count_all_k_equals_per_k_counts(ENG, kmin, kmax, parts):
    """the hierarchical all-k count (one atomic pass at kmax + 4:1 reductions + corrections) == independent per-k counts
    == oracle, in both modes, incl. repeats inside reads, N's, reads shorter than k, block-path and bitmap-path reads"""
    rng = np.random.default_rng(1000 + 17 * kmin + kmax)
    special = ["A" * 80, "CA" * 60, "", "N", "ACG", "ACGTTGCA" * 30, "ACGTACGTAC" * 9 + "T", "GGGGGGGGGGGGGGGGGGGGAGGGGGGGGGGGGGGGGGGG",
               ("ACGTAGCTAGCTAGGATCGAT" * 500)[:9000], "ACGTTGCAAC" * 40]
    reads, seq, borders = rand_reads(rng, 500, 0, 130, p_n=0.02, special=special)
    reads2, seq2, borders2 = rand_reads(rng, 3, 300, 2000, p_n=0.002)
    seq = np.concatenate([seq, seq2])
    borders = np.concatenate([borders, borders2 + borders[-1, 1] + 1])
    # tandem repeats so that k-mers repeat inside reads with different extensions
    dev = ENG.SeqOnDevice.from_numpy(seq, borders)
    for dedup, scheme in ((True, None), (False, None), (True, ENG.SeqOnDevice.PREFIX_PASSES)):
        tabs = dev.count_all(kmin, kmax, dedup, n_partitions=parts, scheme=scheme)
        for k in range(kmin, kmax + 1):
            got = ENG.to_host(tabs[k], np.uint32)
            if k in (kmin, kmax, (kmin + kmax) // 2):
                want = dense_table_from_oracle(seq, borders, k, dedup)
            else:
                want = ENG.to_host(dev.count(k, dedup=dedup), np.uint32)
            assert np.array_equal(got, want), (k, dedup, int(np.abs(got.astype(np.int64) - want.astype(np.int64)).sum()))


@pytest.mark.parametrize("k", [2, 5, 6, 9, 12])
def test_compact_merge_order_exact(ENG, k):
    rng = np.random.default_rng(k)
    reads, seq, borders = rand_reads(rng, 500, 5, 60, p_n=0.01, special=["ACGT" * 10, "AATT" * 8, "GC" * 20])
    dev = ENG.SeqOnDevice.from_numpy(seq, borders)
    for dedup in (True, False):
        table = dev.count(k, dedup=dedup)
        h = O.comp_kmer_hash(seq, k)
        if dedup:
            h = O.remove_duplicate_hash_per_seq(h, borders, np.uint32(0xFFFFFFFF))
        u, c = O.count_uniq_hash(h, k)
        kh, cnt = ENG.compact_merge(table, k, revcom=False)
        assert np.array_equal(ENG.to_host(kh, np.uint32), u) and np.array_equal(ENG.to_host(cnt, np.int32), c)
        mk, mc = O.merge_revcom(u.copy(), c.copy(), k)
        kh, cnt = ENG.compact_merge(table, k, revcom=True)
        assert np.array_equal(ENG.to_host(kh, np.uint32), mk) and np.array_equal(ENG.to_host(cnt, np.int32), mc)


@pytest.mark.parametrize("k,d", [(4, 0), (6, 1), (8, 2), (10, 3), (12, 4), (13, 5)])
def test_hamball_sums_enumeration_and_list(ENG, k, d):
    rng = np.random.default_rng(k * 10 + d)
    n = 4000 if k >= 10 else 600
    reads, seq, borders = rand_reads(rng, n, 20, 60, special=["ACGT" * 12, "GGATCC" * 9])
    dev = ENG.SeqOnDevice.from_numpy(seq, borders)
    table = dev.count(k, dedup=True)
    u, c = O.count_uniq_hash(O.remove_duplicate_hash_per_seq(O.comp_kmer_hash(seq, k), borders, np.uint32(0xFFFFFFFF)), k)
    pal = [int(O.kmer2hash(("ACGT" * 4)[:k]))] if k % 2 == 0 else []
    cand = [int(x) for x in rng.choice(u, 6)] + pal + [0, 4 ** k - 1]
    for revcom in (True, False):
        if revcom:
            mk, mc = O.merge_revcom(u.copy(), c.copy(), k)
        else:
            mk, mc = u, c
        want = np.array([O.hamball_count(mk, mc, np.uint32(x), k, d, revcom) for x in cand], dtype=np.int64)
        got = ENG.hamball_sums(table, k, cand, d, revcom)
        assert np.array_equal(got, want), (k, d, revcom)
        got2 = ENG.hamball_sums_list(ENG.to_device(mk), ENG.to_device(mc), k, cand, d, revcom)
        assert np.array_equal(got2, want), (k, d, revcom)


@pytest.mark.parametrize("idx", range(16))
def test_find_motif_small_all_modes(MD, K, small_cases, motif_def_file, idx, tmp_path):
    import pickle
    mdd = K.init_motif_def_dict(motif_def_file)
    c = small_cases["cases"][idx]
    k = c["k"]
    m = mdd[k]
    bfile = tmp_path / "b.pkl"
    with open(bfile, "wb") as fh:
        pickle.dump(small_cases["borders"], fh)
    pkl = tmp_path / f"k{k}.pkl"
    seq = small_cases["seq"].copy()
    res = MD.find_motif(seq, k, m.max_ham_dist, m.p_uniform, m.ratio_mu, m.ratio_std, m.ratio_cutoff, 5, 10, c["revcom"],
                        c["rep"], save_kmer_cnt_flag=True, kmer_cnt_pkl_file=pkl, boarder_pkl_file=bfile)
    with open(pkl, "rb") as fh:
        kk, ukh, ucnt = pickle.load(fh)
    assert kk == k and np.array_equal(ukh, c["uniq_kh"]) and np.array_equal(ucnt, c["uniq_cnt"])
    assert ukh.dtype == c["uniq_kh"].dtype and ucnt.dtype == c["uniq_cnt"].dtype
    assert [int(x) for x in res] == [int(x) for x in c["consensus"]]
    assert np.array_equal(np.array([list(v) for v in res.values()], dtype=np.float64).reshape(-1, 3), c["stats"])
    assert np.array_equal(seq, c["masked"])
    # second call resumes from the pickle (md:621-624) and must give the same answer
    seq2 = small_cases["seq"].copy()
    res2 = MD.find_motif(seq2, k, m.max_ham_dist, m.p_uniform, m.ratio_mu, m.ratio_std, m.ratio_cutoff, 5, 10, c["revcom"],
                         c["rep"], save_kmer_cnt_flag=True, kmer_cnt_pkl_file=pkl, boarder_pkl_file=bfile)
    assert [int(x) for x in res2] == [int(x) for x in c["consensus"]] and np.array_equal(seq2, c["masked"])


@pytest.mark.parametrize("k", [6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16])
def test_find_motif_testfa(MD, K, testfa, motif_def_file, k, tmp_path):
    import pickle
    mdd = K.init_motif_def_dict(motif_def_file)
    c = next(x for x in testfa["find_motif"] if x["k"] == k)
    m = mdd[k]
    bfile = tmp_path / "b.pkl"
    with open(bfile, "wb") as fh:
        pickle.dump(testfa["borders"], fh)
    pkl = tmp_path / f"k{k}.pkl"
    seq = testfa["input_bin"].copy()
    res = MD.find_motif(seq, k, m.max_ham_dist, m.p_uniform, m.ratio_mu, m.ratio_std, m.ratio_cutoff, 5, 10, True, False,
                        save_kmer_cnt_flag=True, kmer_cnt_pkl_file=pkl, boarder_pkl_file=bfile)
    with open(pkl, "rb") as fh:
        kk, ukh, ucnt = pickle.load(fh)
    assert np.array_equal(ukh, c["uniq_kh"]) and np.array_equal(ucnt, c["uniq_cnt"])
    assert ukh.dtype == c["uniq_kh"].dtype and ucnt.dtype == c["uniq_cnt"].dtype
    assert [int(x) for x in res] == [int(x) for x in c["consensus"]]
    assert np.array_equal(np.array([list(v) for v in res.values()], dtype=np.float64).reshape(-1, 3), c["stats"])
    assert np.array_equal(seq, c["masked"])


def test_ex_hamball_and_cnt_mat(MD, testfa):
    fm = {c["k"]: c for c in testfa["find_motif"]}
    for g in testfa["ex_hamball"]:
        cs = g["conseq"]
        k = len(cs)
        d = g.get("d", 5)
        kh, cnt, mat = MD._hamball_extract(fm[k]["uniq_kh"], fm[k]["uniq_cnt"], int(O.kmer2hash(cs)), k, d, g.get("revcom", True))
        assert np.array_equal(kh, g["kh"]) and np.array_equal(cnt, g["cnt"]) and np.array_equal(mat, g["cnt_mat"])
        cm = MD.cal_cnt_mat(g["kh"], g["cnt"], k)
        assert cm.dtype == g["cnt_mat"].dtype and np.array_equal(cm, g["cnt_mat"])


def test_occurrence_unit_cases(MD, K, unit_vectors, motif_def_file):
    g = unit_vectors["occurrence"]
    mdd = K.init_motif_def_dict(motif_def_file)
    for case in g["cases"]:
        arr = K.dna2arr(case["read"], append_missing_val_flag=False)
        np.random.seed(1)
        flag, s = MD.get_motif_occurence(arr, g["conseqs"], mdd, case["revcom_mode"])
        if case["locs"].count(",") >= 19 and s != case["locs"]:       # > 20 hits: random pick in the reference
            assert flag == case["flag"] and s.count(",") == case["locs"].count(",")
        else:
            assert (flag, s) == (case["flag"], case["locs"]), case


def test_hamdist_matrix_vectors(MD, unit_vectors, testfa):
    for g in unit_vectors["hamdist_mat"]:
        um = MD.cal_samp_kmer_hamdist_mat(g["kh"], g["cnts"], g["labels"], g["conseq_list"], g["k"], uniq_dist_flag=True)
        bm = MD.cal_samp_kmer_hamdist_mat(g["kh"], g["cnts"], g["labels"], g["conseq_list"], g["k"])
        assert str(bm.dtype) == g["ref_dtype"]
        assert np.array_equal(um, g["uniq"]) and np.array_equal(bm, g["block"]), g["k"]
    kh, cnts, labels, conseq_list = testfa["sample_kmers"]
    g = testfa["hamdist"]
    mat = MD.cal_samp_kmer_hamdist_mat(kh, cnts, labels, conseq_list, g["k"])
    assert np.array_equal(mat, g["mat"]) and np.array_equal(MD._convert_to_block_arr(labels, cnts), g["labels"])


def test_hamdist_row_blocks_and_properties(MD):
    # row-block partition (multi-GPU layout): any row range equals the corresponding slice of the full matrix
    rng = np.random.default_rng(3)
    k, n = 14, 1616
    kh = np.unique(rng.integers(0, 4 ** k, n + 50, dtype=np.uint64))[:n].astype(np.uint32)
    rng.shuffle(kh)
    labels = rng.integers(0, 3, n)
    full = MD.hamdist_matrix_u8(kh, labels, [14, 12], k).cpu().numpy()
    assert np.array_equal(full, O.cal_samp_kmer_hamdist_mat(kh, np.ones(n, dtype=int), labels, ["A" * 14, "A" * 12], k).astype(np.uint8))
    assert np.all(full.diagonal() == 0) and np.array_equal(full, full.T)
    for r0, r1 in ((0, 1), (5, 700), (700, 1616), (1615, 1616)):
        part = MD.hamdist_matrix_u8(kh, labels, [14, 12], k, r0, r1).cpu().numpy()
        assert np.array_equal(part, full[r0:r1])


@pytest.mark.parametrize("k,n", [(14, 1616), (14, 777), (8, 300), (16, 2048), (1, 50), (12, 5000)])
def test_hamdist_onehot_tcgen05_gemm_equals_popcount_kernel(MD, k, n):
    """the int8 one-hot GEMM on the tcgen05 tensor cores (csrc/hamdist_mma.cu, the formulation BASELINE config 5 asks to be
    measured against XOR/popcount) writes the same bytes as the XOR/popcount kernel and the oracle: head overrides as extra
    K columns (K = 64, 96, 128) and recomputed in the epilogue when they do not fit, ragged edges (n not a multiple of the
    128 x 256 tile or of 16: TMA clipping / the warps' own stores), row blocks that do not start at a tile boundary"""
    rng = np.random.default_rng(100 + k + n)
    kh = rng.integers(0, 4 ** k, n, dtype=np.uint64).astype(np.uint32)
    labels = rng.integers(0, 4, n)
    head = [k, max(1, k - 2), max(1, k - 5)]                     # label 3 has no consensus entry: never overrides
    want = MD.hamdist_matrix_u8(kh, labels, head, k, impl="popcount").cpu().numpy()
    got = MD.hamdist_matrix_onehot_mma(kh, labels, head, k).cpu().numpy()
    assert np.array_equal(got, want)
    if n <= 2048 and len(np.unique(kh)) == n:
        ref = O.cal_samp_kmer_hamdist_mat(kh, np.ones(n, dtype=int), labels, ["A" * h for h in head], k).astype(np.uint8)
        assert np.array_equal(got, ref)
    for r0, r1 in ((0, 1), (3, n // 2), (n // 2, n), (n - 1, n), (129, min(n, 400))):
        if r0 < r1 <= n:
            part = MD.hamdist_matrix_onehot_mma(kh, labels, head, k, r0, r1).cpu().numpy()
            assert np.array_equal(part, want[r0:r1]), (r0, r1)
    # what the product picks: the GEMM where it applies and pays (k <= 16, override columns within K = 128, n >= 2048)
    assert MD.hamdist_formulation(n, k, head) == ("onehot_mma" if n >= 2048 else "popcount")
    assert MD.hamdist_formulation(100000, 17, []) == "popcount" and MD.hamdist_formulation(100000, 14, [14, 2, 3]) == "popcount"
    assert np.array_equal(MD.hamdist_matrix_u8(kh, labels, head, k).cpu().numpy(), want)
    plain = MD.hamdist_matrix_onehot_mma(kh, np.full(n, -1), [], k).cpu().numpy()
    assert np.array_equal(plain, MD.hamdist_matrix_u8(kh, np.full(n, -1), [], k, impl="popcount").cpu().numpy())
    if k >= 12:       # so many tail bases that the override columns do not fit into K = 128: the epilogue recomputes those pairs
        many = [k, 2, 3, k - 2]
        assert np.array_equal(MD.hamdist_matrix_onehot_mma(kh, labels, many, k).cpu().numpy(),
                              MD.hamdist_matrix_u8(kh, labels, many, k, impl="popcount").cpu().numpy())
        # K = 128 exactly used / one label only
        for heads in ([k, k - 4, k - 4, k - 4][:4], [k - 1]):
            assert np.array_equal(MD.hamdist_matrix_onehot_mma(kh, labels, heads, k).cpu().numpy(),
                                  MD.hamdist_matrix_u8(kh, labels, heads, k, impl="popcount").cpu().numpy()), heads


def test_synth_device_matches_numpy(ENG):
    from kmap_b200 import synth
    for spec in (synth.CFG2, synth.CFG2_N, synth.CFG3):
        seq_d, b_d = synth.generate_device(spec, 1000, 3000)
        seq, b = synth.generate_numpy(spec, 1000, 3000)
        assert np.array_equal(seq_d.cpu().numpy(), seq) and np.array_equal(b_d.cpu().numpy(), b)


def test_scan_motif_workflow_testfa(MD, K, testfa, tmp_path):
    """README workflow (preproc + scan_motif, k=8..14) on test.fa: every text output equals the reference's."""
    import tomli_w
    import tomllib
    seq, borders = testfa["input_bin"], testfa["borders"]
    fa = tmp_path / "test.fa"
    with open(fa, "w") as fh:
        for i, (st, en) in enumerate(borders):
            fh.write(f">r{i}\n{K.arr2dna(seq[st:en])}\n")
    res_dir = tmp_path / "res"
    res_dir.mkdir()
    cfg = tomllib.loads(testfa["text_files"]["config.toml"])
    cfg["general"]["input_fasta_file"] = str(fa)
    cfg["general"]["res_dir"] = str(res_dir)
    with open(res_dir / "config.toml", "wb") as fh:
        tomli_w.dump(cfg, fh)
    K._preproc(str(fa), str(res_dir))
    import pickle
    with open(res_dir / "input.bin.pkl", "rb") as fh:
        assert np.array_equal(pickle.load(fh), seq)
    with open(res_dir / "input.seqboarder.bin.pkl", "rb") as fh:
        b = pickle.load(fh)
        assert np.array_equal(b, borders) and b.dtype == borders.dtype
    assert (res_dir / "motif_def_table.csv").read_text() == testfa["text_files"]["motif_def_table.csv"]
    np.random.seed(20240414)
    MD._scan_motif(str(res_dir))
    tf = testfa["text_files"]
    for name in ("candidate_conseq.csv", "final_conseq.txt", "final_conseq.info.csv",
                 "hamming_balls/cntmat_motif0_AGGACCTACGTAC.csv", "hamming_balls/cntmat_motif1_AATCGATAGCGAA.csv"):
        assert (res_dir / name).read_text() == tf[name], name
    for name in [n for n in tf if n.endswith("motif_occurence.csv")]:
        assert_occurrence_text_equal((res_dir / name).read_text().splitlines(), tf[name])
    fm = {c["k"]: c for c in testfa["find_motif"]}
    for k in range(8, 15):
        with open(res_dir / "kmer_count" / f"k{k}.pkl", "rb") as fh:
            kk, ukh, ucnt = pickle.load(fh)
        assert kk == k and np.array_equal(ukh, fm[k]["uniq_kh"]) and np.array_equal(ucnt, fm[k]["uniq_cnt"])
    # the sample is drawn with numpy's global RNG from identical arrays in the identical call order
    with open(res_dir / "sample_kmers.pkl", "rb") as fh:
        skh, scnt, slab, sconseq = pickle.load(fh)
    gkh, gcnt, glab, gconseq = testfa["sample_kmers"]
    assert sconseq == gconseq and np.array_equal(skh, gkh) and np.array_equal(scnt, gcnt) and np.array_equal(slab, glab)
    with open(res_dir / "sample_kmer_hamdist_mat.pkl", "rb") as fh:
        kk, mat, lab = pickle.load(fh)
    assert kk == testfa["hamdist"]["k"] and str(mat.dtype) == testfa["hamdist"]["ref_dtype"]
    assert np.array_equal(mat, testfa["hamdist"]["mat"]) and np.array_equal(lab, testfa["hamdist"]["labels"])
    assert (res_dir / "sample_kmers.tsv").read_text() == tf["sample_kmers.tsv"]


def test_streamed_count_kmers_equals_single_upload(ENG):
    """api.count_kmers with the reads streamed through the device in chunks (some re-encoded by the host cores and shipped
    packed, some shipped as they are and packed on the device, whichever feeder gets to them first; or all shipped as they
    are) returns the same lists as one upload, in both modes; reads of every path (warp / block / bitmap) sit in different
    chunks"""
    from kmap_b200 import api
    rng = np.random.default_rng(4242)
    special = ["A" * 80, "CA" * 60, "", "N", "ACGTTGCA" * 150, ("ACGTAGCTAGCTAGGATCGAT" * 1500)[:30000], "ACGTTGCAAC" * 40]
    reads, seq, borders = rand_reads(rng, 3000, 0, 130, p_n=0.01, special=special)
    for rep_mode in (False, True):
        one = api.count_kmers(seq, borders, range(8, 15), rep_mode=rep_mode, chunk_positions=1 << 40)
        many = api.count_kmers(seq, borders, range(8, 15), rep_mode=rep_mode, chunk_positions=len(seq) // 7)
        raw = api.count_kmers(seq, borders, range(8, 15), rep_mode=rep_mode, chunk_positions=len(seq) // 23, host_pack=False)
        assert api._chunk_bounds(seq, borders, len(seq) // 7) is not None
        for k in range(8, 15):
            assert np.array_equal(one[k][0], many[k][0]) and np.array_equal(one[k][1], many[k][1]), (k, rep_mode)
            assert np.array_equal(one[k][0], raw[k][0]) and np.array_equal(one[k][1], raw[k][1]), (k, rep_mode)
        want = O.merge_revcom(*O.count_uniq_hash(O.comp_kmer_hash(seq, 11) if rep_mode else
                                                 O.remove_duplicate_hash_per_seq(O.comp_kmer_hash(seq, 11), borders, np.uint32(0xFFFFFFFF)), 11), 11)
        assert np.array_equal(many[11][0], want[0]) and np.array_equal(many[11][1], want[1])


def test_host_encoders_equal_device_pack(ENG):
    """kmap_host_pack2bit (csrc/host_pack.cpp) writes the words kmap_pack2bit writes on the device, and the border matrix
    rebuilt on the device from the per-read strides equals the matrix (kmer_count.py:335-343 layout)"""
    import torch
    from kmap_b200._lib import lib
    L = lib()
    rng = np.random.default_rng(8)
    for n_reads in (0, 1, 700, 40000):
        _, seq, borders = rand_reads(rng, n_reads, 0, 150, p_n=0.02, special=["ACGT"])
        seq = seq.copy()
        seq[rng.random(len(seq)) < 0.001] = 7                          # any byte above 3 is a missing value
        n = len(seq)
        dev = ENG.SeqOnDevice.from_numpy(seq, borders)
        nw = int(L.kmap_valid_words(n))
        packed, valid = np.zeros(2 * nw, np.uint32), np.zeros(nw, np.uint32)
        for th in (1, 5):
            assert L.kmap_host_pack2bit(seq.ctypes.data, n, packed.ctypes.data, valid.ctypes.data, th) == 0
            assert np.array_equal(packed, ENG.to_host(dev.packed, np.uint32)[:2 * nw])
            assert np.array_equal(valid, ENG.to_host(dev.valid, np.uint32)[:nw])
        m = len(borders)
        strides = np.zeros(m, np.uint32)
        shifted = borders + 12345
        assert L.kmap_host_border_strides(shifted.ctypes.data, m, 12345, strides.ctypes.data, 3) == 0
        d_str = ENG.to_device(strides)
        off = ENG.empty(m + 1, torch.int64)
        out = torch.empty((m, 2), dtype=torch.int64, device="cuda")
        scr = ENG.empty(int(L.kmap_list_scratch_words(m)), torch.int64)
        ENG.check(L.kmap_borders_from_strides(d_str.data_ptr(), m, off.data_ptr(), out.data_ptr(), scr.data_ptr(), None), "borders_from_strides")
        assert np.array_equal(out.cpu().numpy(), borders)


# ---- preproc ingest (csrc/fasta.cu) -------------------------------------------------------------------------------------
def _random_fasta_text(rng, n_rec, max_lines, max_len, alphabet="ACGTacgtNn >\t x-"):
    parts = []
    if rng.random() < 0.3:
        parts.append(["junk before the first header\n", "\n\n", "ACGT\r\n", " "][int(rng.integers(0, 4))])
    eols = ["\n", "\r\n", "\r", "\n\n", ""]
    for _ in range(n_rec):
        parts.append(">" + "".join(rng.choice(list("abc >x\t"), int(rng.integers(0, 12)))) + eols[int(rng.integers(0, 3))])
        for _ in range(int(rng.integers(0, max_lines + 1))):
            parts.append("".join(rng.choice(list(alphabet), int(rng.integers(0, max_len + 1)))) + eols[int(rng.integers(0, 5))])
    txt = "".join(parts)
    return txt.rstrip("\r\n") if rng.random() < 0.3 else txt


def test_fasta_ingest_matches_oracle(ENG, K, testfa, tmp_path):
    """FASTA text -> input.bin / input.seqboarder.bin on the device == the oracle's restatement of kmer_count.py:244-347:
    CRLF / CR line ends, blank lines, lower case, N and other letters, white space inside lines, empty records, '>' inside a
    sequence line, no trailing newline, text before the first header; lines and records that straddle tiles and chunks."""
    rng = np.random.default_rng(77)
    cases = [b"", b"no header at all\nACGT\n", b">", b">\n", b">a", b">a\n\n>b\n>c\nAC", b">r\nACGT", b"\n>r\r\nac gt\r\nNN\r\n",
             b">x\n" + b"ACGT" * 3000 + b"\n>y\n" + b"T" * 5000, b">" + b"h" * 9000 + b"\nGATTACA\n",
             b"x>y" * 400000 + b"\n>r1\nACGT\n>r2\nTTGA\n", b"j" * ((1 << 20) - 1) + b">\n>q\nAC\n", b"j" * ((1 << 20) - 1) + b"\n>q\nAC\n"]
    cases += [_random_fasta_text(rng, int(rng.integers(0, 8)), 4, 14).encode() for _ in range(150)]
    cases += [_random_fasta_text(rng, int(rng.integers(50, 400)), 3, 300).encode() for _ in range(6)]
    for i, raw in enumerate(cases):
        fa = tmp_path / f"c{i}.fa"
        fa.write_bytes(raw)
        want_seq, want_b = O.fasta_to_binary(fa)
        text = ENG.read_fasta_bytes(fa)
        for chunk in (1 << 28, 16, 4096 + 16, 7 * 16):
            if chunk < 4096 and len(raw) > 20000:
                continue
            seq_d, b_d = ENG.fasta_text_to_device(text, chunk_bytes=chunk)
            assert np.array_equal(seq_d.cpu().numpy(), want_seq), (i, chunk)
            assert np.array_equal(b_d.cpu().numpy(), np.asarray(want_b, dtype=np.int64).reshape(-1, 2)), (i, chunk)
        for chunk in (1 << 28, 1 << 16):             # the file path: pinned staging buffers, header search across blocks
            seq_d, b_d = ENG.fasta_to_device(fa, chunk_bytes=chunk)
            assert np.array_equal(seq_d.cpu().numpy(), want_seq), (i, chunk)
            assert np.array_equal(b_d.cpu().numpy(), np.asarray(want_b, dtype=np.int64).reshape(-1, 2)), (i, chunk)
    # test.fa itself (rebuilt from the golden arrays, 60 bases per line) and its gzipped copy, through the public function
    seq, borders = testfa["input_bin"], testfa["borders"]
    lines = []
    for r, (st, en) in enumerate(borders):
        s = K.arr2dna(seq[st:en])
        lines.append(f">read{r} some description")
        lines.extend(s[j:j + 60] for j in range(0, len(s), 60))
    fa = tmp_path / "test.fa"
    fa.write_text("\n".join(lines) + "\n")
    import gzip
    with gzip.open(tmp_path / "test.fa.gz", "wb") as fh:
        fh.write(fa.read_bytes())
    for f in (fa, tmp_path / "test.fa.gz"):
        got_seq, got_b = K.fasta_to_arrays(str(f))
        assert np.array_equal(got_seq, seq) and np.array_equal(got_b, borders) and got_b.dtype == borders.dtype
    dev = ENG.SeqOnDevice.from_fasta(fa)
    ref = ENG.SeqOnDevice.from_numpy(seq, borders)
    assert dev.n == ref.n and dev.n_seq == ref.n_seq
    assert np.array_equal(dev.packed.cpu().numpy(), ref.packed.cpu().numpy()) and np.array_equal(dev.valid.cpu().numpy(), ref.valid.cpu().numpy())


def test_fasta_ingest_large_synthetic(ENG):
    """2e5 synthetic reads written as FASTA text: the device parse gives back exactly the arrays the text was made from"""
    from kmap_b200 import synth
    seq, borders = synth.generate_numpy(synth.CFG2, 0, 200000)
    L = int(borders[0, 1] - borders[0, 0])
    body = np.frombuffer(b"ACGT", dtype=np.uint8)[np.minimum(seq.reshape(-1, L + 1)[:, :L], 3)]
    body = np.where(seq.reshape(-1, L + 1)[:, :L] == 255, ord("N"), body).astype(np.uint8)
    rec = np.concatenate([np.full((len(body), 1), ord(">"), np.uint8), np.full((len(body), 1), ord("r"), np.uint8),
                          np.full((len(body), 1), 10, np.uint8), body, np.full((len(body), 1), 10, np.uint8)], axis=1)
    text = rec.reshape(-1)
    for chunk in (1 << 28, 1 << 20):
        seq_d, b_d = ENG.fasta_text_to_device(text, chunk_bytes=chunk)
        assert np.array_equal(seq_d.cpu().numpy(), seq) and np.array_equal(b_d.cpu().numpy(), borders)


# ---- sort / run-length path (csrc/sorted.cu): uint64 hashes, int64 counts ------------------------------------------------
def _oracle_counts(seq, borders, k, dedup):
    h = O.comp_kmer_hash(seq, k)
    if dedup:
        h = O.remove_duplicate_hash_per_seq(h, borders, O.get_invalid_hash(h.dtype.type))
    return O.count_uniq_hash(h, k)


def test_radix_sort_and_run_lengths_vs_numpy(ENG):
    """sort + run-length encoding == np.unique(return_counts) for even / odd pass counts, many duplicates, all-ones keys"""
    import torch
    rng = np.random.default_rng(11)
    for n, bits, n_distinct in [(1, 8, 1), (37, 8, 5), (5000, 16, 300), (100001, 34, 20000), (1 << 20, 62, 1 << 18), (300000, 64, 1000)]:
        pool = rng.integers(0, 1 << min(bits, 63), n_distinct, dtype=np.uint64)
        if bits == 64:
            pool |= np.uint64(1) << np.uint64(63)
            pool = pool[pool != np.uint64(0xFFFFFFFFFFFFFFFF)]
        keys = pool[rng.integers(0, len(pool), n)]
        keys[rng.random(n) < 0.2] = np.uint64(0xFFFFFFFFFFFFFFFF)
        want_u, want_c = np.unique(keys[keys != np.uint64(0xFFFFFFFFFFFFFFFF)], return_counts=True)
        kh, cnt = ENG.sort_count_keys(ENG.to_device(keys.copy()), bits)
        assert np.array_equal(ENG.to_host(kh, np.uint64), want_u) and np.array_equal(ENG.to_host(cnt, np.int64), want_c), (n, bits)
    kh, cnt = ENG.sort_count_keys(ENG.to_device(np.full(1000, 0xFFFFFFFFFFFFFFFF, dtype=np.uint64)), 40)
    assert kh.numel() == 0 and cnt.numel() == 0


@pytest.mark.parametrize("k", [5, 14, 16, 17, 21, 31])
def test_sorted_count_and_merge_vs_oracle(ENG, K, k):
    """count_sorted (keys from the packed reads, per-read de-duplication on the warp / block paths, sort, run lengths) and
    the sorted-list merge_revcom == the oracle, for 64-bit hashes (k >= 16) and, forced, for small k (palindromes)"""
    rng = np.random.default_rng(100 + k)
    special = ["A" * 80, "CA" * 60, "", "N", "ACGT", "ACGTTGCA" * 150, ("ACGTAGCTAGCTAGGATCGAT" * 1500)[:30000],
               "ACGTTGCAAC" * 40, "AATT" * 20, "GAATTC" * 30]
    reads, seq, borders = rand_reads(rng, 1500, 0, 130, p_n=0.01, special=special)
    dev = ENG.SeqOnDevice.from_numpy(seq, borders)
    for dedup in (False, True):
        want_u, want_c = _oracle_counts(seq, borders, k, dedup)
        kh, cnt = dev.count_sorted(k, dedup)
        got_u, got_c = ENG.to_host(kh, np.uint64), ENG.to_host(cnt, np.int64)
        assert np.array_equal(got_u, want_u.astype(np.uint64)) and np.array_equal(got_c, want_c.astype(np.int64)), (k, dedup)
        mu, mc = O.merge_revcom(want_u.copy(), want_c.copy(), k)
        mkh, mcnt = ENG.merge_revcom_sorted(kh, cnt, k)
        assert np.array_equal(ENG.to_host(mkh, np.uint64), mu.astype(np.uint64)), (k, dedup)
        assert np.array_equal(ENG.to_host(mcnt, np.int64), mc.astype(np.int64)), (k, dedup)
    if k >= 16:    # the reference-shaped array functions on uint64 hashes
        h = K.comp_kmer_hash_taichi(seq, k)
        assert h.dtype == np.uint64 and np.array_equal(h, O.comp_kmer_hash(seq, k))
        hd = K.remove_duplicate_hash_per_seq(h.copy(), borders, K.get_invalid_hash(h.dtype.type))
        assert np.array_equal(hd, O.remove_duplicate_hash_per_seq(h.copy(), borders, O.get_invalid_hash(h.dtype.type)))
        u, c = K.count_uniq_hash(hd, k)
        wu, wc = O.count_uniq_hash(hd, k)
        assert u.dtype == wu.dtype and c.dtype == wc.dtype and np.array_equal(u, wu) and np.array_equal(c, wc)
        c_in, wc_in = c.copy(), wc.copy()
        m1 = K.merge_revcom(u, c_in, k)
        m0 = O.merge_revcom(wu, wc_in, k)
        assert np.array_equal(m1[0], m0[0]) and np.array_equal(m1[1], m0[1]) and np.array_equal(c_in, wc_in)
        assert m1[0].dtype == m0[0].dtype and m1[1].dtype == m0[1].dtype


@pytest.mark.parametrize("k,d", [(16, 5), (18, 4), (24, 6)])
def test_hamball_lists_u64_vs_oracle(ENG, MD, k, d):
    rng = np.random.default_rng(k)
    centre = rng.integers(0, 4, k)
    kms = []
    for _ in range(4000):
        s = centre.copy()
        idx = rng.choice(k, int(rng.integers(0, d + 3)), replace=False)
        s[idx] = rng.integers(0, 4, len(idx))
        kms.append(int(O.kmer2hash(O.arr2dna(s.astype(np.uint8)))))
    kms += [int(x) for x in rng.integers(0, 4 ** k, 3000, dtype=np.uint64)]
    kh = np.unique(np.array(kms, dtype=np.uint64))
    cnt = rng.integers(1, 50, len(kh)).astype(np.int64)
    c0 = int(O.kmer2hash(O.arr2dna(centre.astype(np.uint8))))
    cands = [c0, int(O.revcom_hash(np.uint64(c0), k)), int(kh[5]), int(kh[-1])]
    kh_d, cnt_d = ENG.to_device(kh), ENG.to_device(cnt)
    for revcom in (True, False):
        got = ENG.hamball_sums_list64(kh_d, cnt_d, k, cands, d, revcom)
        want = [O.hamball_count(kh, cnt, np.uint64(c), k, d, revcom) for c in cands]
        assert [int(x) for x in got] == [int(x) for x in want], (k, d, revcom)
    cs = O.hash2kmer(min(c0, int(O.revcom_hash(np.uint64(c0), k))), k)
    for revcom in (True, False):
        wkh, wcnt = O.ex_hamball_from_arrays(kh, cnt, cs, d, revcom)
        gkh, gcnt, gmat = MD._hamball_extract(kh, cnt, int(O.kmer2hash(cs)), k, d, revcom)
        assert np.array_equal(gkh, wkh) and np.array_equal(gcnt, wcnt)
        assert np.array_equal(gmat, O.cal_cnt_mat(wkh, wcnt, k))
    assert np.array_equal(MD.cal_cnt_mat(kh, cnt, k), O.cal_cnt_mat(kh, cnt, k))


def test_find_motif_sorted_path_equals_dense_path(ENG, MD, K, small_cases, motif_def_file):
    """the two counting paths are interchangeable: find_motif through the sort path (forced) == through the dense table"""
    mdd = K.init_motif_def_dict(motif_def_file)
    for idx in (0, 5, 10, 15):
        c = small_cases["cases"][idx]
        k, m = c["k"], mdd[c["k"]]
        out = []
        for sorted_path in (False, True):
            dev = ENG.SeqOnDevice.from_numpy(small_cases["seq"], small_cases["borders"], keep_u8=True)
            found, first = MD.find_motif_on_device(dev, k, m.max_ham_dist, m.p_uniform, m.ratio_mu, m.ratio_std, m.ratio_cutoff, 5, 10,
                                                   c["revcom"], c["rep"], sorted_path=sorted_path)
            seq = small_cases["seq"].copy()
            dev.masked_seq_to_numpy(seq)
            out.append((found, first, seq))
        (f0, a0, s0), (f1, a1, s1) = out
        assert [int(x) for x in f0] == [int(x) for x in f1] == [int(x) for x in c["consensus"]]
        assert all(f0[a] == f1[b] for a, b in zip(f0, f1))
        assert np.array_equal(a0[0], a1[0]) and np.array_equal(a0[1], a1[1]) and a0[0].dtype == a1[0].dtype and a0[1].dtype == a1[1].dtype
        assert np.array_equal(s0, s1) and np.array_equal(s0, c["masked"])


@pytest.mark.parametrize("k,conseqs", [(14, ["GTACGTAGGTCCTA", "AATCGATAGCGA", "ACGTAG"]), (16, ["AGTACGTAGGTCCTCA", "AATCGATAGCGAAG"]),
                                      (9, ["ACCTACGTA"]), (12, ["AAAAAAAAAAAA", "CCCCCCCC"])])
def test_label_kmers_vs_oracle(MD, K, motif_def_file, k, conseqs):
    """the fused labelling kernel of sample_disp_kmer == the oracle's statement-by-statement restatement (md:849-892)"""
    mdd = K.init_motif_def_dict(motif_def_file)
    conseqs = [c if O.kmer2hash(c) <= O.revcom_hash(O.kmer2hash(c), len(c)) else O.reverse_complement(c) for c in conseqs]
    rng = np.random.default_rng(k)
    hd = K.get_hash_dtype(k)
    pool = [rng.integers(0, 4 ** k, 4000, dtype=np.uint64)]
    for c in conseqs:                       # k-mers around each consensus and around its reverse complement, at every offset
        for s in (c, O.reverse_complement(c)):
            base = np.array([int(x) for x in O.dna2arr(s, append_missing_val_flag=False)])
            for _ in range(600):
                full = rng.integers(0, 4, k)
                off = int(rng.integers(0, k - len(s) + 1))
                full[off:off + len(s)] = base
                idx = rng.choice(k, int(rng.integers(0, 4)), replace=False)
                full[idx] = rng.integers(0, 4, len(idx))
                pool.append(np.array([int(O.kmer2hash(O.arr2dna(full.astype(np.uint8))))], dtype=np.uint64))
    kh = np.unique(np.concatenate(pool)).astype(hd)
    for revcom in (True, False):
        want_kh, want_label = O.label_kmers(kh, conseqs, k, mdd, revcom)
        got_kh, got_label = MD.label_kmers(kh, conseqs, k, mdd, revcom)
        assert got_kh.dtype == want_kh.dtype and np.array_equal(got_kh, want_kh), (k, revcom)
        assert np.array_equal(got_label, want_label), (k, revcom)
        assert len(set(got_label.tolist())) >= 2


def test_count_kmers_api_mixed_dense_and_wide_k(ENG):
    """api.count_kmers across the dense / sorted boundary: k = 14, 15 from dense tables, k = 16, 17 from the sort path,
    every list equal to the oracle's first-round (merged) counts with the reference's dtypes"""
    from kmap_b200 import api
    rng = np.random.default_rng(99)
    reads, seq, borders = rand_reads(rng, 800, 0, 120, p_n=0.01, special=["A" * 70, "ACGTACGTTTGCA" * 9, "GAATTC" * 12])
    for rep_mode in (False, True):
        got = api.count_kmers(seq, borders, [14, 15, 16, 17], rep_mode=rep_mode)
        for k in (14, 15, 16, 17):
            u, c = _oracle_counts(seq, borders, k, not rep_mode)
            wu, wc = O.merge_revcom(u, c, k)
            assert got[k][0].dtype == wu.dtype and got[k][1].dtype == wc.dtype, k
            assert np.array_equal(got[k][0], wu) and np.array_equal(got[k][1], wc), (k, rep_mode)


def test_scan_motif_workflow_stock_k_range(MD, K, testfa, testfa_stock_k, tmp_path):
    """README workflow with the STOCK k range of default_config.toml (6..16): k = 15 on the 4 GiB dense table, k = 16 through
    the uint64 sort path (three accepted consensus sequences, so mask + recount run there too); every text output, the
    k15 / k16 pickles, the sample and its distance matrix equal the unmodified reference's."""
    import pickle
    import tomli_w
    import tomllib
    g = testfa_stock_k
    seq, borders = testfa["input_bin"], testfa["borders"]
    fa = tmp_path / "test.fa"
    with open(fa, "w") as fh:
        for i, (st, en) in enumerate(borders):
            fh.write(f">r{i}\n{K.arr2dna(seq[st:en])}\n")
    res_dir = tmp_path / "res"
    res_dir.mkdir()
    cfg = tomllib.loads(g["text_files"]["config.toml"])
    assert cfg["kmer_count"]["min_k"] == 6 and cfg["kmer_count"]["max_k"] == 16
    cfg["motif_discovery"]["motif_pos_density_flag"] = True          # stock settings: the data files of these two steps
    cfg["motif_discovery"]["motif_co_occurence_flag"] = True         # are compared with the reference functions' below
    cfg["general"]["input_fasta_file"] = str(fa)
    cfg["general"]["res_dir"] = str(res_dir)
    with open(res_dir / "config.toml", "wb") as fh:
        tomli_w.dump(cfg, fh)
    K._preproc(str(fa), str(res_dir))
    np.random.seed(20240415)
    MD._scan_motif(str(res_dir))
    tf = g["text_files"]
    for name in [n for n in tf if n.endswith((".txt", ".tsv")) or n.startswith("hamming_balls/") or n in
                 ("candidate_conseq.csv", "final_conseq.info.csv", "motif_def_table.csv")]:
        assert (res_dir / name).read_text() == tf[name], name
    for name in [n for n in tf if n.endswith("motif_occurence.csv")]:
        assert_occurrence_text_equal((res_dir / name).read_text().splitlines(), tf[name])
    for k in (15, 16):
        with open(res_dir / "kmer_count" / f"k{k}.pkl", "rb") as fh:
            kk, ukh, ucnt = pickle.load(fh)
        want = g["kmer_count"][k]
        assert kk == k and ukh.dtype == want["uniq_kh"].dtype and ucnt.dtype == want["uniq_cnt"].dtype
        assert np.array_equal(ukh, want["uniq_kh"]) and np.array_equal(ucnt, want["uniq_cnt"])
    with open(res_dir / "sample_kmers.pkl", "rb") as fh:
        skh, scnt, slab, sconseq = pickle.load(fh)
    gkh, gcnt, glab, gconseq = g["sample_kmers"]
    assert sconseq == gconseq and np.array_equal(skh, gkh) and np.array_equal(scnt, gcnt) and np.array_equal(slab, glab)
    with open(res_dir / "sample_kmer_hamdist_mat.pkl", "rb") as fh:
        kk, mat, lab = pickle.load(fh)
    assert kk == g["hamdist"]["k"] and str(mat.dtype) == g["hamdist"]["ref_dtype"]
    assert np.array_equal(mat, g["hamdist"]["mat"]) and np.array_equal(lab, g["hamdist"]["labels"])
    # density / co-occurrence data files (reference motif_discovery.py:364-425) against what the reference's own functions
    # returned for the same final.motif_occurence.csv (tests/golden/consumers.pkl.gz)
    import gzip
    with gzip.open(GOLDEN_DIR / "consumers.pkl.gz", "rb") as fh:
        cons = next(c for c in pickle.load(fh) if c["name"] == "testfa_stock_k.pkl.gz")
    for name, text in cons["files"].items():
        assert (res_dir / "co_occurence" / name).read_text() == text, name
    with open(res_dir / "motif_pos_density.np.pkl", "rb") as fh:
        x_arr, dens = pickle.load(fh)
    assert np.array_equal(x_arr, cons["x_arr"]) and np.array_equal(dens, np.vstack([d[2] for d in cons["density"]]))


def test_topk_candidates_and_find_motif_selection(ENG, MD):
    """the device top-k candidates == the k largest counts (value descending, index ascending), for int32 and int64 lists, and
    _top_k_indices falls back to numpy's own selection whenever the boundary is tied"""
    import torch
    rng = np.random.default_rng(5)
    for dtype, n in [(np.int32, 10), (np.int32, 70000), (np.int32, 3_000_000), (np.int64, 200_000)]:
        cnt = rng.integers(0, 1000, n).astype(dtype)
        cnt[rng.integers(0, n, 4)] += 5000
        for kk in (1, 6, 8):
            val, idx = ENG.topk_candidates(ENG.to_device(cnt), kk)
            order = np.lexsort((np.arange(n), -cnt.astype(np.int64)))[:kk]
            assert np.array_equal(idx, order) and np.array_equal(val, cnt[order]), (dtype, n, kk)

    class S:       # what _top_k_indices reads of a _CountState
        pass
    st = S()
    st.cnt = rng.integers(0, 50, 100000).astype(np.int32)
    st.cnt[[7, 99, 5000, 77777, 1234]] = [900, 800, 700, 600, 500]
    st.cnt_dev = ENG.to_device(st.cnt)
    inds, unambiguous = MD._top_k_indices(st, 5)
    assert unambiguous and sorted(inds.tolist()) == sorted([7, 99, 5000, 77777, 1234])
    assert sorted(inds.tolist()) == sorted(np.argpartition(st.cnt, -5)[-5:].tolist())
    st.cnt[4321] = 500                       # a tie at the boundary: numpy decides
    st.cnt_dev = ENG.to_device(st.cnt)
    inds, unambiguous = MD._top_k_indices(st, 5)
    assert not unambiguous and np.array_equal(inds, np.argpartition(st.cnt, -5)[-5:])


@pytest.mark.parametrize("k", [17, 20, 31])
def test_wide_k_mask_occurrence_and_find_motif_vs_oracle(ENG, MD, K, k):
    """17 <= k <= 31 (64-bit hashes end to end): mask_input, the occurrence scan and the whole find_motif loop (count by
    sorting, ball sums on lists, mask, recount) == the oracle; the motif definition is made up (the stock table has no
    accepted k there), so that consensus sequences ARE accepted and masked"""
    rng = np.random.default_rng(1000 + k)
    motif = "".join("ACGT"[b] for b in rng.integers(0, 4, k))
    reads = []
    for _ in range(1200):
        L = int(rng.integers(0, 90))
        s = rng.integers(0, 4, L).astype(np.uint8)
        s[rng.random(L) < 0.01] = 255
        r = O.arr2dna(s)
        if rng.random() < 0.5 and L > k + 2:
            mm = list(motif)
            for j in rng.choice(k, int(rng.integers(0, 3)), replace=False):
                mm[j] = "ACGT"[int(rng.integers(0, 4))]
            mm = "".join(mm) if rng.random() < 0.5 else O.reverse_complement("".join(mm))
            off = int(rng.integers(0, L - k))
            r = r[:off] + mm + r[off + k:]
        reads.append(r)
    reads += ["T" * 60, "A" * 50, "N" * 5, ""]
    arrs = [O.dna2arr(r) for r in reads]
    seq = np.concatenate(arrs)
    lens = np.array([len(a) for a in arrs])
    ends = np.cumsum(lens)
    borders = np.stack([ends - lens, ends - 1], axis=1).astype(np.int64)
    ckh = O.kmer2hash(motif)
    cons = np.array([ckh, O.revcom_hash(ckh, k), O.kmer2hash("T" * (k - 1) + "G")], dtype=np.uint64)
    ds = np.array([2, 2, 1])
    want = O.mask_input(seq.copy(), k, cons, ds)
    got = K.mask_input(seq.copy(), k, cons, ds)
    assert np.array_equal(got, want) and (want != seq).any()
    # occurrence scan: every read, consensus of length k
    m_def = K.MotifDef(kmer_len=k, p_uniform=1e-6, max_ham_dist=2, ratio_mu=1.0, ratio_std=0.1, ratio_cutoff=1.5)
    mdd = {k: m_def}
    dev = ENG.SeqOnDevice.from_numpy(seq, borders)
    cmin = motif if ckh <= O.revcom_hash(ckh, k) else O.reverse_complement(motif)
    lines = MD.motif_occurence_lines(dev, borders, [cmin], mdd, True)
    np.random.seed(3)
    want_lines = O.motif_occurence_lines(reads, [cmin], mdd, True)
    assert_occurrence_text_equal(lines, "\n".join(want_lines))
    # the whole find_motif loop
    for rep in (False, True):
        work = seq.copy()
        want_found, want_first = O.find_motif(work, k, 2, m_def.p_uniform, m_def.ratio_mu, m_def.ratio_std, m_def.ratio_cutoff,
                                              boarder_mat=borders, rep_mode=rep)
        dev = ENG.SeqOnDevice.from_numpy(seq, borders, keep_u8=True)
        found, first = MD.find_motif_on_device(dev, k, 2, m_def.p_uniform, m_def.ratio_mu, m_def.ratio_std, m_def.ratio_cutoff,
                                               rep_mode=rep)
        assert len(want_found) >= 1 and [int(x) for x in found] == [int(x) for x in want_found]
        assert all(found[a] == want_found[b] for a, b in zip(found, want_found))
        assert np.array_equal(first[0], want_first[0]) and np.array_equal(first[1], want_first[1])
        out = seq.copy()
        dev.masked_seq_to_numpy(out)
        assert np.array_equal(out, work)


def test_large_input_identities(ENG):
    """size-independent identities on an input far beyond what the oracle can count (5e6 synthetic ChIP-like reads x 100 bp,
    5.05e8 positions, generated on the device): repetitive mode -- every table sums to the number of windows of its level;
    de-duplicated mode -- the derived tables of the all-k count (routed level 13, fused lower levels) equal direct per-k
    counts, and a 4:1 fold of level k+1 never exceeds level k by more than the run-end windows; sort path at k = 16 --
    counts sum to the windows, keys ascend"""
    import torch
    from kmap_b200 import synth
    n_reads, L = 5_000_000, 100
    seq_d, borders_d = synth.generate_device(synth.CFG3, 0, n_reads)
    dev = ENG.SeqOnDevice.from_device_u8(seq_d, borders_d)
    del seq_d
    rep = dev.count_all(8, 14, dedup=False)
    for k in range(8, 15):
        assert int(rep[k].to(torch.int64).sum().item()) == n_reads * (L - k + 1), k
    ded = dev.count_all(8, 14, dedup=True)
    for k in (8, 11, 13):
        direct = dev.count(k, dedup=True)
        assert torch.equal(direct, ded[k]), k
        del direct
    for k in range(8, 14):
        folded = ded[k + 1].view(-1, 4).sum(dim=1, dtype=torch.int64)
        extra = ded[k].to(torch.int64) - folded            # +1 per read end, -1 per repeat whose extension is new
        assert int(extra.sum().item()) <= n_reads and int(extra.abs().sum().item()) <= 2 * n_reads, k
        assert int(ded[k].to(torch.int64).sum().item()) <= n_reads * (L - k + 1)
    kh, cnt = dev.count_sorted(16, dedup=False)
    assert int(cnt.sum().item()) == n_reads * (L - 16 + 1)
    assert bool((kh[1:] > kh[:-1]).all().item()) and int(cnt.min().item()) >= 1


def test_consumers_from_scan_equal_consumers_of_the_file(MD, K, motif_def_file, tmp_path):
    """co-occurrence counts / per-read median differences (csrc/consumers.cu) and the position density computed from the
    occurrence-scan results == the reference-shaped functions that parse final.motif_occurence.csv back
    (motif_discovery.py:1189-1343; those are pinned to the reference's own outputs in test_oracle_golden.py).  Reads with
    more than 20 positions in a cell (random pick, :1467-1469) take the host route: poly-A / CA-repeat reads force it."""
    import kmap_b200.engine as E
    rng = np.random.default_rng(11)
    mdd = K.init_motif_def_dict(motif_def_file)
    conseqs = ["ACGTACGT", "AAAAAAAA", "CACACACAC", "GGATCCGGATCC", "TTTTGGGGCCCC"]
    special = ["A" * 60, "CA" * 40, "A" * 30 + "ACGTACGT" + "CA" * 20, "ACGTACGTTTTTGGGGCCCC", "ACGT", "ACGTACGTACGTACGT" + "A" * 12]
    reads = list(special)
    for _ in range(3000):
        L = int(rng.integers(6, 90))
        s = rng.integers(0, 4, L).astype(np.uint8)
        for c in conseqs:                                     # plant (possibly mutated) copies so that motifs share reads
            if rng.random() < 0.35 and L > len(c) + 2:
                at = int(rng.integers(0, L - len(c)))
                m = K.dna2arr(c, append_missing_val_flag=False).copy()
                if rng.random() < 0.5:
                    m[int(rng.integers(0, len(m)))] = int(rng.integers(0, 4))
                s[at:at + len(m)] = m
        reads.append(K.arr2dna(s))
    fa = tmp_path / "in.fa"
    with open(fa, "w") as fh:
        for i, r in enumerate(reads):
            fh.write(f">r{i}\n{r}\n")
    for revcom in (True, False):
        out = tmp_path / f"occ_{int(revcom)}.csv"
        np.random.seed(5)
        stats, info = MD.gen_motif_occurence_file(conseqs, mdd, fa, out, revcom, _return_scan=True)
        assert len(info["picked_rows"]) >= 3 and info["cooc"] is not None
        assert stats == [MD.get_motif_seq_num(out, i) for i in range(len(conseqs))]
        co_f, dist_f, dd_f = MD.get_motif_co_occurence_mat(out, len(conseqs))
        co_s, dist_s, dd_s = MD.co_occurrence_from_scan(info, len(conseqs))
        assert co_s.dtype == co_f.dtype and np.array_equal(co_s, co_f)
        assert np.array_equal(dist_s, dist_f)
        assert dd_s == dd_f and any(len(v) for v in dd_f.values())
        for i, c in enumerate(conseqs):
            a = MD.get_motif_pos_density(out, i, len(c))
            b = MD.pos_density_from_scan(info, i, len(c))
            assert a[:2] == b[:2] and np.array_equal(a[2], b[2]), c
    # the exact total of a count list (the `sum()` of find_motif, :648)
    import torch
    c32 = torch.randint(0, 2 ** 31 - 1, (1_000_003,), dtype=torch.int32, device="cuda")
    assert E.sum_counts(c32) == int(c32.to(torch.int64).sum().item())
    c64 = torch.randint(0, 2 ** 40, (77_777,), dtype=torch.int64, device="cuda")
    assert E.sum_counts(c64) == int(c64.sum().item())
