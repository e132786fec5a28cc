#!/usr/bin/env python
"""Generate the golden fixtures in this directory by running the UNMODIFIED reference
(`/root/reference/src/kmap`) under `ref_shim.py` (a taichi interpreter + FASTA reader stand-in).

Run once in the build container (the GPU box has no /root/reference):

    python tests/golden/make_golden.py            # everything (several minutes: kernels run as Python loops)
    python tests/golden/make_golden.py unit       # only the unit-level vectors

The fixtures pin `oracle/kmap_oracle.py` (tests/test_oracle_golden.py) and, through the same files,
the CUDA path (tests/test_gpu_*.py).
"""
from __future__ import annotations

import gzip
import hashlib
import json
import pickle
import shutil
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
import ref_shim  # noqa: E402

kc, md, tc = ref_shim.install()
MOTIF_DEF = "/root/reference/src/kmap/default_motif_def_table.csv"
TEST_FA = "/root/reference/tests/test.fa"


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def rand_seq_arr(rng, n_reads, lmin, lmax, p_n=0.0, special=()):
    """uint8 array in the reference's input.bin layout + borders, built through the reference's dna2arr."""
    reads = []
    for i in range(n_reads):
        L = int(rng.integers(lmin, lmax + 1))
        s = "".join("ACGT"[b] for b in rng.integers(0, 4, L))
        if p_n > 0:
            s = "".join(("N" if rng.random() < p_n else c) for c in s)
        reads.append(s)
    reads = list(special) + reads
    arrs = [kc.dna2arr(s) for s in reads]
    seq = np.concatenate(arrs)
    borders = np.zeros((len(reads), 2), dtype=int)
    p = 0
    for i, a in enumerate(arrs):
        borders[i] = (p, p + len(a) - 1)
        p += len(a)
    return reads, seq, borders


# ----------------------------------------------------------------------------------------------------------
def gen_unit():
    out = {}
    rng = np.random.default_rng(20240412)

    # --- tests/kmap_tests.py:173-189  count k=3 on a string with N runs, vs the independent KmerCounter
    import tests.inimotif as im
    seq = ("TTTTCGTNCACGACGCTACCTTAAAGCATCCTTCTNTGATACCATAGANNNNNGCAGCTCCTTATCGTTTTAGCTTTCGTATTCGTCTAATCGTCTTTTACT"
           "CGACGAAAA")
    kcnt = im.KmerCounter(3, revcom_flag=False, unique_kmer_in_seq_mode=False)
    kcnt.dtype = lambda v: np.uint64(int(v) & 0xFFFFFFFFFFFFFFFF)  # numpy>=2: dtype(-1) raises; same wrap as numpy 1.x
    indep = kcnt.scan_seq(seq)
    arr = kc.dna2arr(seq)
    h = kc.comp_kmer_hash_taichi(arr, 3)
    u, c = kc.count_uniq_hash(h, 3)
    assert {int(a): int(b) for a, b in zip(u, c)} == {int(a): int(b) for a, b in indep.items()}
    out["count_k3"] = dict(seq=seq, hash=h, uniq=u, cnt=c,
                           indep_keys=np.array(sorted(int(a) for a in indep), dtype=np.uint64),
                           indep_vals=np.array([int(indep[a]) for a in sorted(indep)], dtype=np.int64))

    # --- tests/kmap_tests.py:268-284  mask_ham_ball known answers
    mdd = kc.init_motif_def_dict(MOTIF_DEF)
    s1 = "AAAAAAAAAAAAAAAAAAAAAACTAGCTGCCAGTCCCCCCCCCCC"
    r1 = kc.arr2dna(kc.mask_ham_ball(kc.dna2arr(s1)[:-1], mdd, ["AAA", "CCCC"], [0, 0]))
    assert r1 == "NNNNNNNNNNNNNNNNNNNNNNCTAGCTGCCAGTNNNNNNNNNNN"
    s2 = "AAAAAAAAAAAAAAAAAAAAAACTAGCTGGGGGGGGGGGGGGGGGGGGGGGGGGCCAGTCCCCCCCCCCC"
    r2 = kc.arr2dna(kc.mask_ham_ball(kc.dna2arr(s2)[:-1], mdd, ["AAAAAAA", "CCCCCCCC", "GGGGGGGGG"]))
    assert r2 == "NNNNNNNNNNNNNNNNNNNNNNNTANNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNANNNNNNNNNNNNN"
    out["mask_ham_ball"] = dict(s1=s1, r1=r1, c1=["AAA", "CCCC"], d1=[0, 0],
                                s2=s2, r2=r2, c2=["AAAAAAA", "CCCCCCCC", "GGGGGGGGG"])

    # --- mask_input directly, incl. the invalid-hash-as-T..T quirk (kmer_count.py:592-607)
    mi = []
    for k, conseq, d in [(5, "TTTTC", 2), (5, "TTTTT", 1), (6, "AAAAAA", 1), (8, "ACGTACGT", 3), (4, "TTTA", 1)]:
        reads, seqarr, borders = rand_seq_arr(rng, 40, 3, 30, p_n=0.03, special=["TTTTTTTTTTTT", "AAAAAAAAAA", "ACG"])
        kh = kc.kmer2hash(conseq)
        rc = kc.revcom_hash(kh, k)
        before = seqarr.copy()
        after = kc.mask_input(seqarr, k, np.array([kh, rc]), np.array([d, d]))
        mi.append(dict(k=k, conseq=conseq, d=d, before=before, after=after.copy()))
    out["mask_input"] = mi

    # --- tests/kmap_tests.py:212-238  merge_revcom
    kh_arr = np.array([0, 2, 10, 11, 17, 18, 19, 23, 27, 33, 36, 38, 41, 43, 46, 51, 53, 57, 59])
    cnt_arr = np.ones_like(kh_arr)
    mk, mc = kc.merge_revcom(kh_arr.copy(), cnt_arr.copy(), 3, keep_lower_hash_flag=True)
    for x in (10, 17, 36):
        assert mc[mk == x] == 2
    assert mc.sum() == 19
    mr = [dict(k=3, kh=kh_arr, cnt=cnt_arr, out_kh=mk, out_cnt=mc)]
    for k, n in [(4, 1000), (4, 150), (5, 700), (6, 3000), (8, 20000), (7, 5)]:
        raw = rng.integers(0, 4 ** k, n).astype(np.uint32)
        u, c = np.unique(raw, return_counts=True)
        c = c.astype(np.int32)
        c_in = c.copy()
        mk, mc = kc.merge_revcom(u.copy(), c, k, keep_lower_hash_flag=True)
        mr.append(dict(k=k, kh=u, cnt=c_in, out_kh=mk, out_cnt=mc, mutated_cnt=c.copy()))
    out["merge_revcom"] = mr

    # --- tests/kmap_tests.py:241-266  kmer2hash / hash2kmer
    k2h = {}
    for kmer in ["ACTGA", "ACTACTGGAGGACCTACGTAAGCCACGA", "T" * 14, "GTACGTAGGTCCTA"]:
        khh = kc.kmer2hash(kmer)
        assert kc.hash2kmer(khh, len(kmer)) == kmer
        k2h[kmer] = int(khh)
    out["kmer2hash"] = k2h

    # --- 1:1 primitives on random data, u32 and u64 (taichi_core.py:3-224)
    prim = []
    for k in (3, 8, 11, 14, 15, 16, 20, 31):
        hd = kc.get_hash_dtype(k)
        n = 300
        reads, seqarr, borders = rand_seq_arr(rng, 8, 1, 70, p_n=0.04)
        h = kc.comp_kmer_hash_taichi(seqarr, k)
        hi = 4 ** k
        kh = rng.integers(0, hi, n, dtype=np.uint64).astype(hd)
        kh[:3] = (0, hi - 1, np.iinfo(hd).max)          # incl. the invalid hash as an operand
        target = hd(rng.integers(0, hi, dtype=np.uint64))
        rec = dict(k=k, seq=seqarr, borders=borders, hash=h, kh=kh, target=np.array([target]),
                   dist=kc.cal_hamming_dist(kh, target, k),
                   revcom=kc.get_revcom_hash_arr(kh[:2].copy(), k) if True else None,
                   revcom_in=kh[:2].copy())
        vkh = kh[kh < hi]
        rec["revcom_in"] = vkh
        rec["revcom"] = kc.get_revcom_hash_arr(vkh, k)
        rec["revcom_scalar"] = np.array([kc.revcom_hash(x, k) for x in vkh[:20]], dtype=hd)
        for cl in sorted({max(1, k - 5), max(1, k - 2), k}):
            ct = hd(rng.integers(0, 4 ** cl, dtype=np.uint64))
            rec[f"head_{cl}"] = kc.cal_hamming_dist_head(kh, ct, k, cl)
            rec[f"tail_{cl}"] = kc.cal_hamming_dist_tail(kh, ct, k, cl)
            rec[f"ct_{cl}"] = np.array([ct])
        # dedup per read
        inv = kc.get_invalid_hash(hd)
        rec["dedup"] = kc.remove_duplicate_hash_per_seq(h.copy(), borders, inv)
        prim.append(rec)
    out["primitives"] = prim

    # --- dedup on low-complexity reads (tests/test_kmer_count.py:51-71 intent)
    reads, seqarr, borders = rand_seq_arr(rng, 5, 20, 40, special=["A" * 42, "CA" * 25, "ACGACGACGACGNACGACGACG", "AC"])
    h = kc.comp_kmer_hash_taichi(seqarr, 8)
    dd = kc.remove_duplicate_hash_per_seq(h.copy(), borders, kc.get_invalid_hash(np.uint32))
    b0, b1 = borders[0], borders[1]
    assert dd[b0[0]] == kc.kmer2hash("A" * 8) and np.all(dd[b0[0] + 1:b0[1]] == 0xFFFFFFFF)
    assert dd[b1[0]] == kc.kmer2hash("CA" * 4) and dd[b1[0] + 1] == kc.kmer2hash("AC" * 4)
    assert np.all(dd[b1[0] + 2:b1[1]] == 0xFFFFFFFF)
    out["dedup_lowcomplex"] = dict(k=8, seq=seqarr, borders=borders, hash=h, dedup=dd)

    # --- get_motif_occurence (motif_discovery.py:1422-1477), incl. N, L<k (negative slice quirk), L==k
    occ = []
    conseqs = ["ACGTAC", "TTTTTTTT", "GGACCTACGTAC", "AAAAA"]
    reads = ["ACGTACGTACGTAC", "TTTTTTTTTTT", "ACG", "ACGTA", "ACGTAC", "NNACGTACNN", "GTACGT", "AAAAAAAAAAAAAAAAAAAAAAAA",
             "AGGACCTACGTACTTTGTACGTAGGTCCT", "TTTTTTT", "TTTT", "T", "CCCCCCCCC", "ACGTNCGTACGAACGTAC", "AAAAATTTTT"]
    for i in range(25):
        L = int(rng.integers(1, 45))
        reads.append("".join("ACGTN"[b] for b in rng.choice(5, L, p=[0.24, 0.24, 0.24, 0.24, 0.04])))
    for revcom_mode in (True, False):
        for r in reads:
            a = kc.dna2arr(r, append_missing_val_flag=False)
            np.random.seed(1)
            flag, s = md.get_motif_occurence(a, conseqs, mdd, revcom_mode)
            occ.append(dict(read=r, revcom_mode=revcom_mode, flag=bool(flag), locs=s))
    out["occurrence"] = dict(conseqs=conseqs, cases=occ)

    # --- block expansion (tests/kmap_tests.py:434-441)
    um = rng.integers(0, 100, (3, 3))
    out["block"] = dict(cnts=np.array([2, 1, 4]), mat=um, out=md._convert_to_block_mat(um, np.array([2, 1, 4])),
                        arr_out=md._convert_to_block_arr(np.array([7, 8, 9]), np.array([2, 1, 4])))

    # --- cal_samp_kmer_hamdist_mat (motif_discovery.py:759-808) with a shorter second conseq (head override)
    hm = []
    for (k, conseq_list, n) in [(14, ["GTACGTAGGTCCTA", "AATCGATAGCGA"], 260), (12, ["ACGTACGTACGT"], 90),
                                (16, ["GTACGTAGGTCCTAAC", "AATCGATAGC"], 120), (10, ["AATCGATAGC", "ACGTAC", "GGA"], 150)]:
        hd = kc.get_hash_dtype(k)
        khs, labels = [], []
        for li, cs in enumerate(conseq_list):
            for _ in range(n // 4):
                s = list(cs + "".join("ACGT"[b] for b in rng.integers(0, 4, k - len(cs))))
                for _m in range(int(rng.integers(0, 4))):
                    s[int(rng.integers(0, len(cs)))] = "ACGT"[int(rng.integers(0, 4))]
                khs.append(int(kc.kmer2hash("".join(s))))
                labels.append(li)
        while len(khs) < n:
            khs.append(int(rng.integers(0, 4 ** k, dtype=np.uint64)))
            labels.append(len(conseq_list))
        khs, idx = np.unique(np.array(khs, dtype=np.uint64), return_index=True)
        labels = np.array(labels)[idx]
        perm = rng.permutation(len(khs))
        khs, labels = khs[perm].astype(hd), labels[perm]
        cnts = rng.integers(1, 4, len(khs))
        um = md.cal_samp_kmer_hamdist_mat(khs, cnts, labels, conseq_list, k, uniq_dist_flag=True)
        bm = md.cal_samp_kmer_hamdist_mat(khs, cnts, labels, conseq_list, k, uniq_dist_flag=False)
        assert um.dtype == np.int64 and um.max() < 256
        hm.append(dict(k=k, conseq_list=conseq_list, kh=khs, cnts=cnts, labels=labels,
                       uniq=um.astype(np.uint8), block=bm.astype(np.uint8), ref_dtype=str(bm.dtype)))
    out["hamdist_mat"] = hm

    with gzip.open(HERE / "unit_vectors.pkl.gz", "wb") as fh:
        pickle.dump(out, fh, protocol=4)
    print("unit vectors written")


# ----------------------------------------------------------------------------------------------------------
def _find_motif_case(seq, borders, k, revcom, rep, mdd, tmp, top_k=5, n_trial=10):
    m = mdd[k]
    bfile = Path(tmp) / "b.pkl"
    with open(bfile, "wb") as fh:
        pickle.dump(borders, fh)
    pkl = Path(tmp) / f"k{k}_{int(revcom)}{int(rep)}.pkl"
    if pkl.exists():
        pkl.unlink()
    work = seq.copy()
    res = md.find_motif(work, k, m.max_ham_dist, m.p_uniform, m.ratio_mu, m.ratio_std, m.ratio_cutoff, top_k, n_trial,
                        revcom, rep, save_kmer_cnt_flag=True, kmer_cnt_pkl_file=pkl, boarder_pkl_file=bfile)
    with open(pkl, "rb") as fh:
        kk, ukh, ucnt = pickle.load(fh)
    return dict(k=k, revcom=revcom, rep=rep, uniq_kh=ukh, uniq_cnt=ucnt, masked=work,
                consensus=np.array([int(x) for x in res.keys()], dtype=np.uint64),
                stats=np.array([list(v) for v in res.values()], dtype=np.float64).reshape(-1, 3))


def gen_small():
    """find_motif on small seeded inputs in all four (revcom, repetitive) modes."""
    rng = np.random.default_rng(7)
    mdd = kc.init_motif_def_dict(MOTIF_DEF)
    cases = []
    tmp = tempfile.mkdtemp()
    # planted motif so that something is accepted; N's, short reads, poly-T/A, repeats
    def planted(n, motif, lmin, lmax):
        rs = []
        for _ in range(n):
            L = int(rng.integers(lmin, lmax + 1))
            s = [("ACGT"[b]) for b in rng.integers(0, 4, L)]
            if L >= len(motif) and rng.random() < 0.6:
                p = int(rng.integers(0, L - len(motif) + 1))
                for i, b in enumerate(motif):
                    if rng.random() > 0.05:
                        s[p + i] = b
            if rng.random() < 0.1:
                s[int(rng.integers(0, L))] = "N"
            rs.append("".join(s))
        return rs
    special = ["T" * 30, "A" * 25, "CA" * 20, "ACG", "", "N", "ACGTACGTAC" * 4]
    reads = special + planted(400, "AATCGATAGC", 4, 50)
    arrs = [kc.dna2arr(s) for s in reads]
    seq = np.concatenate(arrs)
    borders = np.zeros((len(reads), 2), dtype=int)
    p = 0
    for i, a in enumerate(arrs):
        borders[i] = (p, p + len(a) - 1)
        p += len(a)
    for k in (5, 6, 8, 10):
        for revcom in (True, False):
            for rep in (False, True):
                t = time.time()
                cases.append(_find_motif_case(seq, borders, k, revcom, rep, mdd, tmp))
                print(f"small k={k} revcom={revcom} rep={rep}: {len(cases[-1]['consensus'])} motifs, {time.time()-t:.1f}s",
                      flush=True)
    shutil.rmtree(tmp)
    with gzip.open(HERE / "small_find_motif.pkl.gz", "wb") as fh:
        pickle.dump(dict(reads=reads, seq=seq, borders=borders, cases=cases), fh, protocol=4)


# ----------------------------------------------------------------------------------------------------------
def gen_testfa():
    """README workflow on tests/test.fa: preproc + scan_motif (k=8..14) + ex_hamball, through the reference drivers."""
    import tomllib
    import tomli_w
    res_dir = Path(tempfile.mkdtemp()) / "res"
    res_dir.mkdir()
    with open("/root/reference/src/kmap/default_config.toml", "rb") as fh:
        cfg = tomllib.load(fh)
    cfg["kmer_count"]["min_k"] = 8
    cfg["kmer_count"]["max_k"] = 14
    cfg["motif_discovery"]["motif_pos_density_flag"] = False     # float KDE + plots: out of scope
    cfg["motif_discovery"]["motif_co_occurence_flag"] = False    # plots: out of scope
    cfg["motif_discovery"]["n_total_sample"] = 600               # keeps the committed matrix small
    cfg["motif_discovery"]["n_motif_sample"] = 300
    cfg["general"]["input_fasta_file"] = TEST_FA
    cfg["general"]["res_dir"] = str(res_dir)
    with open(res_dir / "config.toml", "wb") as fh:
        tomli_w.dump(cfg, fh)
    t = time.time()
    kc._preproc(TEST_FA, str(res_dir))
    print(f"preproc {time.time()-t:.1f}s", flush=True)
    np.random.seed(20240414)
    t = time.time()
    md._scan_motif(str(res_dir))
    print(f"scan_motif {time.time()-t:.1f}s", flush=True)

    out = {}
    with open(res_dir / "input.bin.pkl", "rb") as fh:
        out["input_bin"] = pickle.load(fh)
    with open(res_dir / "input.seqboarder.bin.pkl", "rb") as fh:
        out["borders"] = pickle.load(fh)
    out["kmer_count"] = {}
    for k in range(8, 15):
        with open(res_dir / "kmer_count" / f"k{k}.pkl", "rb") as fh:
            kk, ukh, ucnt = pickle.load(fh)
        out["kmer_count"][k] = dict(uniq_kh=ukh, uniq_cnt=ucnt)
    text = {}
    for p in sorted(res_dir.rglob("*")):
        if p.suffix in (".csv", ".txt", ".tsv", ".toml"):
            text[str(p.relative_to(res_dir))] = p.read_text()
    out["text_files"] = text
    with open(res_dir / "sample_kmers.pkl", "rb") as fh:
        out["sample_kmers"] = pickle.load(fh)
    with open(res_dir / "sample_kmer_hamdist_mat.pkl", "rb") as fh:
        kk, mat, lab = pickle.load(fh)
    assert mat.max() < 256
    out["hamdist"] = dict(k=kk, mat=mat.astype(np.uint8), ref_dtype=str(mat.dtype), labels=lab)

    # find_motif per k directly, keeping the masked array and the stats (not written to any file by the driver)
    mdd = kc.init_motif_def_dict(MOTIF_DEF)
    tmp = tempfile.mkdtemp()
    fm = []
    for k in (6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16):
        t = time.time()
        fm.append(_find_motif_case(out["input_bin"], out["borders"], k, True, False, mdd, tmp))
        c = fm[-1]
        print(f"test.fa k={k}: n_uniq={len(c['uniq_kh'])} consensus={[kc.hash2kmer(x, k) for x in c['consensus']]} "
              f"{time.time()-t:.1f}s", flush=True)
    out["find_motif"] = fm

    # ex_hamball hash/kmer outputs for the final consensus sequences
    finals = text["final_conseq.txt"].split()
    exh = []
    for cs in finals:
        ukh, ucnt = md.ex_hamball_kh_arr(str(res_dir), cs, -1, str(res_dir / "motif_def_table.csv"), True)
        exh.append(dict(conseq=cs, kh=ukh, cnt=ucnt, cnt_mat=md.cal_cnt_mat(ukh, ucnt, len(cs))))
        ukh, ucnt = md.ex_hamball_kh_arr(str(res_dir), cs, 2, str(res_dir / "motif_def_table.csv"), False)
        exh.append(dict(conseq=cs, d=2, revcom=False, kh=ukh, cnt=ucnt, cnt_mat=md.cal_cnt_mat(ukh, ucnt, len(cs))))
    out["ex_hamball"] = exh
    # the k{k}.pkl payloads written by the driver are the first-round counts find_motif returned: keep one copy
    for c in fm:
        if c["k"] in out["kmer_count"]:
            kcnt = out["kmer_count"][c["k"]]
            assert np.array_equal(kcnt["uniq_kh"], c["uniq_kh"]) and np.array_equal(kcnt["uniq_cnt"], c["uniq_cnt"])
            assert kcnt["uniq_kh"].dtype == c["uniq_kh"].dtype and kcnt["uniq_cnt"].dtype == c["uniq_cnt"].dtype
    del out["kmer_count"]
    with gzip.open(HERE / "testfa.pkl.gz", "wb") as fh:
        pickle.dump(out, fh, protocol=4)
    print("test.fa goldens written; final conseqs:", finals)


def gen_testfa_stock_k():
    """The same README workflow with the STOCK k range of default_config.toml (min_k = 6, max_k = 16): k = 16 runs the
    reference's uint64 / int64 code paths.  Only what the driver writes is kept (text files, sample, distance matrix,
    k{k}.pkl payloads of k = 15, 16)."""
    import tomllib
    import tomli_w
    res_dir = Path(tempfile.mkdtemp()) / "res"
    res_dir.mkdir()
    with open("/root/reference/src/kmap/default_config.toml", "rb") as fh:
        cfg = tomllib.load(fh)
    assert cfg["kmer_count"]["min_k"] == 6 and cfg["kmer_count"]["max_k"] == 16
    cfg["motif_discovery"]["motif_pos_density_flag"] = False     # float KDE + plots: out of scope
    cfg["motif_discovery"]["motif_co_occurence_flag"] = False    # plots: out of scope
    cfg["motif_discovery"]["n_total_sample"] = 400               # keeps the committed matrix small
    cfg["motif_discovery"]["n_motif_sample"] = 200
    cfg["general"]["input_fasta_file"] = TEST_FA
    cfg["general"]["res_dir"] = str(res_dir)
    with open(res_dir / "config.toml", "wb") as fh:
        tomli_w.dump(cfg, fh)
    t = time.time()
    kc._preproc(TEST_FA, str(res_dir))
    np.random.seed(20240415)
    md._scan_motif(str(res_dir))
    print(f"stock-k scan_motif {time.time()-t:.1f}s", flush=True)
    out = {}
    text = {}
    for p in sorted(res_dir.rglob("*")):
        if p.suffix in (".csv", ".txt", ".tsv", ".toml"):
            text[str(p.relative_to(res_dir))] = p.read_text()
    out["text_files"] = text
    out["kmer_count"] = {}
    for k in (15, 16):
        with open(res_dir / "kmer_count" / f"k{k}.pkl", "rb") as fh:
            kk, ukh, ucnt = pickle.load(fh)
        out["kmer_count"][k] = dict(uniq_kh=ukh, uniq_cnt=ucnt)
    with open(res_dir / "sample_kmers.pkl", "rb") as fh:
        out["sample_kmers"] = pickle.load(fh)
    with open(res_dir / "sample_kmer_hamdist_mat.pkl", "rb") as fh:
        kk, mat, lab = pickle.load(fh)
    assert mat.max() < 256
    out["hamdist"] = dict(k=kk, mat=mat.astype(np.uint8), ref_dtype=str(mat.dtype), labels=lab)
    with gzip.open(HERE / "testfa_stock_k.pkl.gz", "wb") as fh:
        pickle.dump(out, fh, protocol=4)
    print("stock-k goldens written; final conseqs:", text["final_conseq.txt"].split(), "sample k =", kk)


def gen_consumers():
    """The data-producing consumers of final.motif_occurence.csv in the reference's scan_motif (motif_discovery.py:364-425):
    get_motif_pos_density, get_motif_co_occurence_mat and the three writers, called directly (their plotting siblings are
    not run) on the occurrence files recorded above and on a synthetic file with three motifs and empty / multi-hit cells."""
    out = []
    srcs = []
    for name in ("testfa.pkl.gz", "testfa_stock_k.pkl.gz"):
        with gzip.open(HERE / name, "rb") as fh:
            g = pickle.load(fh)
        finals = g["text_files"]["final_conseq.txt"].split()
        srcs.append((name, g["text_files"]["final.motif_occurence.csv"], finals))
    rng = np.random.default_rng(5)
    conseqs = ["ACGTAC", "TTGACA", "GGATCC"]
    lines = ["seq_ind;" + ";".join(f"motif_{i}_{c}" for i, c in enumerate(conseqs)) + ";seq_len"]
    for r in range(400):
        L = int(rng.integers(20, 120))
        cells = []
        for j in range(3):
            n = int(rng.choice([0, 0, 1, 1, 2, 3, 5]))
            cells.append(",".join(map(str, sorted(rng.choice(L - 6, n, replace=False).tolist()))) if n else "")
        if any(cells):
            lines.append(f"{r};" + ";".join(cells) + f";{L}")
    srcs.append(("synthetic3", "\n".join(lines) + "\n", conseqs))
    tmp = Path(tempfile.mkdtemp())
    for name, text, finals in srcs:
        f = tmp / "occ.csv"
        f.write_text(text)
        n = len(finals)
        co_mat, dist_mat, dist_dict = md.get_motif_co_occurence_mat(f, n)
        co_sum = np.diag(co_mat) + np.diag(co_mat).reshape((-1, 1))
        norm = 2 * co_mat / co_sum
        files = {}
        for key, mat in (("co_occurence_mat.tsv", co_mat + 0.0), ("co_occurence_mat.norm.tsv", norm),
                         ("co_occurence_motif_dist_mat.tsv", dist_mat)):
            md.write_co_occurence_mat(tmp / key, mat, finals)
            files[key] = (tmp / key).read_text()
        md.write_co_occurence_dist_arr(tmp / "d.txt", dist_dict, finals)
        files["co_occurence_motif_dist_data.txt"] = (tmp / "d.txt").read_text()
        x_step = 0.01
        x_arr = np.arange(0, 1.0 + x_step, x_step)
        dens = []
        for i, c in enumerate(finals):
            dens.append(md.get_motif_pos_density(f, i, len(c), x_step=x_step, x_arr=x_arr))
        out.append(dict(name=name, occurence_text=text, conseqs=finals, co_mat=co_mat, dist_mat=dist_mat,
                        dist_dict={k: list(map(float, v)) for k, v in dist_dict.items()}, files=files, x_arr=x_arr,
                        density=[(a, b, d) for a, b, d in dens]))
        print(name, "co_mat", co_mat.tolist(), "density sums", [float(d[2].sum()) for d in dens])
    with gzip.open(HERE / "consumers.pkl.gz", "wb") as fh:
        pickle.dump(out, fh, protocol=4)


if __name__ == "__main__":
    what = sys.argv[1:] or ["unit", "small", "testfa"]
    if "unit" in what:
        gen_unit()
    if "small" in what:
        gen_small()
    if "testfa" in what:
        gen_testfa()
    if "stock" in what:
        gen_testfa_stock_k()
    if "consumers" in what:
        gen_consumers()
