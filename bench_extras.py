"""Secondary measurements reported under "extras" by `bench.py --extras`: the other kernels of the scan_motif counting
path on the same device-resident workload.  CUDA-event timed, after warm-up."""
from __future__ import annotations

import numpy as np
import torch


def _time(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


def fill_rate_gbs(n_bytes=10_000_000_000, reps=3):
    """write-only rate of this GPU: kmap_fill_u32 over n_bytes (the ceiling of a kernel that only writes, like the distance
    matrix: the copy peak of MEASURED_PEAKS.json counts a read AND a write)"""
    from kmap_b200 import engine as E
    from kmap_b200._lib import check, lib
    buf = E.empty(n_bytes // 4, torch.int32)
    ms, _ = _time(lambda: check(lib().kmap_fill_u32(buf.data_ptr(), buf.numel(), 0x01010101, torch.cuda.current_stream().cuda_stream), "kmap_fill_u32"),
                  reps=reps)
    del buf
    return n_bytes / ms / 1e6, ms


def run_piece2(dev, tables, n_reads, L, peak_gbs):
    """Piece 2 of the north star on the bench's own device-resident workload (BASELINE config 4): Hamming-ball aggregation at
    k = 14, d = 5 over the candidates find_motif looks at (the 5 most frequent merged k-mers) plus 1 000 random present k-mers,
    in both formulations (neighbour enumeration on the dense table / scan of the merged list = the reference's, md:666-673),
    ball extraction with the 4 x 14 count matrices (md:924-986), order-exact compaction + reverse-complement merge
    (kc:476-491, 643-685), mask_input (kc:580-610), the occurrence scan (md:1422-1477) and the whole find_motif loop (md:594-702)."""
    import ctypes
    from kmap_b200 import engine as E
    from kmap_b200._lib import check, lib
    from kmap_b200 import motif_discovery as MD
    from kmap_b200.kmer_count import init_motif_def_dict, kmer2hash, revcom_hash
    from pathlib import Path
    res = {}
    k, d = 14, 5
    table = tables[k]
    ms, (kh, cnt) = _time(lambda: E.compact_merge(table, k, True, upper_bound=1 << 28))
    n_m = int(kh.numel())
    moved = 6 * 4 ** k * 4 + 8 * n_m      # F read 3x, G = F[rc] written once + read twice, 8 B per merged entry written
    res["compact_merge_k14"] = {"ms": ms, "n_merged": n_m, "bytes_moved": moved, "GBs": moved / ms / 1e6, "frac_of_hbm": moved / ms / 1e6 / peak_gbs}
    # candidates: the 5 most frequent merged k-mers (what a find_motif trial looks at) + 1000 random present ones
    val, idx = E.topk_candidates(cnt, 5)
    rng = np.random.default_rng(20240415)
    pick = np.concatenate([idx, rng.integers(0, n_m, 1000)]).astype(np.int64)
    cand = [int(x) & 0xFFFFFFFF for x in kh[torch.from_numpy(pick).cuda()].cpu().numpy().view(np.uint32)]
    ms_e5, s5 = _time(lambda: E.hamball_sums(table, k, cand[:5], d, True), reps=5)
    ms_e, sums_e = _time(lambda: E.hamball_sums(table, k, cand, d, True), reps=2)
    ms_l, sums_l = _time(lambda: E.hamball_sums_list(kh, cnt, k, cand, d, True), reps=1, warm=0)
    pairs = 2 * len(cand) * n_m            # reference formulation: every merged k-mer against each candidate and its rc
    res["hamball_sum_k14_d5"] = {
        "candidates": len(cand), "n_merged": n_m, "formulations_agree": bool(np.array_equal(sums_e, sums_l)),
        "enumeration_ms": ms_e, "enumeration_top5_ms": ms_e5, "list_scan_ms": ms_l,
        "pairs_reference_formulation": pairs, "pairs_per_s_enumeration": pairs / ms_e * 1e3, "pairs_per_s_list_scan": pairs / ms_l * 1e3,
        "list_scan_GBs": 8.0 * n_m * (len(cand) / 16.0) / ms_l / 1e6,
        "list_scan_frac_of_hbm": 8.0 * n_m * np.ceil(len(cand) / 16.0) / ms_l / 1e6 / peak_gbs,
        "note": "list scan = the reference's formulation (16 candidates per pass over the 8 B/entry merged list); enumeration "
                "gathers the <= 2 x 578 257 ball members of each candidate from the dense table"}
    # ball extraction + 4 x 14 count matrices of the 5 candidates, from the merged list on the device (md:924-986)
    Lb = lib()
    scratch = E._scratch(Lb.kmap_list_scratch_words(n_m))
    cnt_mat = E.empty(4 * k, torch.int64)
    n_out = ctypes.c_int64(0)
    cap = 2 * 578_257 + 16
    okh, ocnt = E.empty(cap, torch.int32), E.empty(cap, torch.int32)

    def extract_all():
        sizes = []
        for c in cand[:5]:
            c = min(c, int(revcom_hash(c, k)))
            check(Lb.kmap_hamball_extract(kh.data_ptr(), cnt.data_ptr(), n_m, k, c, d, 1, scratch.data_ptr(), okh.data_ptr(), ocnt.data_ptr(), cap,
                                          ctypes.byref(n_out), cnt_mat.data_ptr(), torch.cuda.current_stream().cuda_stream), "kmap_hamball_extract")
            sizes.append(n_out.value)
        return sizes
    ms, sizes = _time(extract_all, reps=2)
    res["hamball_extract_k14_d5"] = {"ms_per_consensus": ms / 5, "ball_members": sizes, "GBs": 5 * 2 * 8.0 * n_m / ms / 1e6,
                                     "note": "two passes over the merged list per consensus (count, write) + the 4 x 14 matrix"}
    # mask_input with one consensus + its reverse complement over the whole input
    c = int(kmer2hash("GTACGTAGGTCCTA"))
    rc = int(revcom_hash(c, k))
    dev.snapshot_valid()

    def do_mask():
        dev.restore_valid()
        dev.mask(k, [c, rc], [d, d])
    ms, _ = _time(do_mask)
    ms_restore, _ = _time(dev.restore_valid)
    dev.restore_valid()
    res["mask_k14_d5"] = {"ms": ms - ms_restore, "positions": dev.n, "Gpositions_per_s": dev.n / (ms - ms_restore) / 1e6,
                          "algorithmic_GBs": 0.5 * dev.n / (ms - ms_restore) / 1e6}
    ms, (mind, offs, pos) = _time(lambda: E.occurrence_scan(dev, k, c, d, True), reps=1)
    res["occurrence_scan_k14_d5"] = {"ms_incl_d2h": ms, "reads": n_reads, "reads_with_hit": int((np.diff(offs) > 0).sum()), "hits": int(len(pos))}
    # the whole find_motif loop at k = 14 on the resident reads, first count handed in (what scan_motif does per k)
    mdd = init_motif_def_dict(Path(MD.__file__).resolve().parent / "default_motif_def_table.csv")
    m = mdd[k]
    torch.cuda.synchronize()
    import time as _t
    t0 = _t.perf_counter()
    found, _ = MD.find_motif_on_device(dev, k, m.max_ham_dist, m.p_uniform, m.ratio_mu, m.ratio_std, m.ratio_cutoff, first_table=table.clone())
    torch.cuda.synchronize()
    res["find_motif_k14"] = {"ms": (_t.perf_counter() - t0) * 1e3, "consensus": [MD.hash2kmer(int(h), k) for h in found],
                             "note": "top-5 selection, ball sums, z-test, mask, recount per accepted consensus (md:594-702); wall clock"}
    dev.restore_valid()
    return res


def run_extras(dev, tables, n_reads, L, peak_gbs=6455.3):
    from kmap_b200 import engine as E
    from kmap_b200.motif_discovery import hamdist_matrix_u8
    from kmap_b200.kmer_count import kmer2hash, revcom_hash
    res = {}
    k, d = 14, 5
    table = tables[k]
    # order-exact compaction + reverse-complement merge of the k=14 table (count_uniq_hash + merge_revcom)
    ms, (kh, cnt) = _time(lambda: E.compact_merge(table, k, True, upper_bound=1 << 28))
    n_m = int(kh.numel())
    # bytes moved: F read by the permutation, the count pass and the write pass; G = F[rc] written once and read twice;
    # 8 B per merged entry written
    moved = 6 * 4 ** k * 4 + 8 * n_m
    res["compact_merge_k14"] = {"ms": ms, "n_merged": n_m, "table_GB": 4 ** k * 4 / 1e9, "bytes_moved": moved,
                                "GBs": moved / ms / 1e6, "frac_of_hbm": moved / ms / 1e6 / peak_gbs}
    # Hamming-ball sums for top_k = 5 candidates (fwd + rc), k = 14, d = 5
    cnt_host = cnt.cpu().numpy()
    top = np.argpartition(cnt_host, -5)[-5:]
    cand = [int(x) & 0xFFFFFFFF for x in kh[torch.from_numpy(top).cuda()].cpu().numpy().view(np.uint32)]
    ms_e, sums_e = _time(lambda: E.hamball_sums(table, k, cand, d, True), reps=5)
    ms_l, sums_l = _time(lambda: E.hamball_sums_list(kh, cnt, k, cand, d, True), reps=3)
    assert np.array_equal(sums_e, sums_l), "enumeration and list formulations disagree"
    pairs = 2 * len(cand) * n_m          # reference formulation: every merged k-mer against each candidate and its rc
    res["hamball_sum_k14_d5_top5"] = {
        "enumeration_ms": ms_e, "list_scan_ms": ms_l, "pairs_reference_formulation": pairs,
        "pairs_per_s_enumeration": pairs / ms_e * 1e3, "pairs_per_s_list_scan": pairs / ms_l * 1e3,
        "list_scan_GBs": 8.0 * n_m / ms_l / 1e6, "list_scan_frac_of_hbm": 8.0 * n_m / ms_l / 1e6 / peak_gbs}
    # mask_input with one consensus + its reverse complement over the whole input, then restore
    c = int(kmer2hash("GTACGTAGGTCCTA"))
    rc = int(revcom_hash(c, k))
    dev.snapshot_valid()

    def do_mask():
        dev.restore_valid()
        dev.mask(k, [c, rc], [d, d])
    ms, _ = _time(do_mask)
    ms_restore, _ = _time(dev.restore_valid)
    dev.restore_valid()
    n_pos = dev.n
    res["mask_k14_d5"] = {"ms": ms - ms_restore, "positions": n_pos, "Gpositions_per_s": n_pos / (ms - ms_restore) / 1e6,
                          "algorithmic_GBs": 0.5 * n_pos / (ms - ms_restore) / 1e6}
    # occurrence scan of one consensus over all reads (count + scan + fill)
    ms, (mind, offs, pos) = _time(lambda: E.occurrence_scan(dev, k, c, d, True), reps=2)
    res["occurrence_scan_k14_d5"] = {"ms_incl_d2h": ms, "reads": n_reads, "reads_with_hit": int((np.diff(offs) > 0).sum()),
                                     "hits": int(len(pos))}
    # sampled k-mer distance matrix: 100 000 distinct 14-mers, second consensus of length 12 (head override), uint8 output
    rng = np.random.default_rng(20240414)
    n = 100_000
    khs = np.unique(rng.integers(0, 4 ** k, int(n * 1.01), dtype=np.uint64))[:n].astype(np.uint32)
    rng.shuffle(khs)
    labels = rng.integers(0, 3, n).astype(np.int32)
    out = E.empty(n * n, torch.uint8)
    ms, _ = _time(lambda: hamdist_matrix_u8(khs, labels, [14, 12], k, 0, n, out=out), reps=5)
    res["hamdist_matrix_100k_k14"] = {"ms": ms, "pairs": n * n, "pairs_per_s": n * n / ms * 1e3, "write_GBs": n * n / ms / 1e6,
                                      "frac_of_hbm": n * n / ms / 1e6 / peak_gbs,
                                      "note": "1 B written per pair (uint8); includes the H2D of the 100k keys/labels"}
    del out
    # The same matrix through an int8 one-hot GEMM (4k = 56 MACs per pair) as the library offers it: cuBLASLt int8 x int8 ->
    # int32 (torch._int_mm) on a slab of rows, then an epilogue pass matches -> uint8 distance.  A comparator for the design
    # decision of DESIGN.md section 4.7 only; nothing in the product calls it.  (A fused tcgen05 kernel with a uint8 epilogue
    # would skip the int32 round trip and then meet the same 1 B/pair store bound as the popcount kernel.)
    try:
        slab = 8192
        codes = torch.from_numpy(((khs[:, None].astype(np.uint64) >> (2 * (k - 1 - np.arange(k, dtype=np.uint64)))[None, :]) & np.uint64(3)).astype(np.int64)).cuda()
        onehot = torch.zeros((n, 4 * k), dtype=torch.int8, device="cuda")
        onehot.scatter_(1, codes + 4 * torch.arange(k, device="cuda")[None, :], 1)
        bt = onehot.t().contiguous()
        a = onehot[:slab].contiguous()
        ms_g, matches = _time(lambda: torch._int_mm(a, bt), reps=5)
        ms_e, d8 = _time(lambda: (k - matches).to(torch.uint8), reps=5)
        want = hamdist_matrix_u8(khs, np.full(n, 2, dtype=np.int32), [14, 12], k, 0, slab)
        ok = bool(torch.equal(d8, want))
        res["hamdist_int8_onehot_gemm_cublaslt"] = {
            "slab_rows": slab, "gemm_ms_slab": ms_g, "epilogue_ms_slab": ms_e, "full_matrix_ms_extrapolated": (ms_g + ms_e) * n / slab,
            "int32_bytes_written_per_pair": 4, "equals_popcount_kernel": ok,
            "note": "library int8 GEMM (cuBLASLt via torch._int_mm) + elementwise epilogue; comparator only"}
        del onehot, bt, a, matches, d8, want
    except Exception as exc:                                   # the comparator must never break the bench
        res["hamdist_int8_onehot_gemm_cublaslt"] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
    return res
