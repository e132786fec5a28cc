"""Measurements of the SURVEY 8f rows built this round (CUDA events, after warm-up; CPU legs = the oracle port on a sample):
  * preproc ingest: FASTA text -> input.bin / borders on the device (csrc/fasta.cu)
  * sort / run-length counting path (csrc/sorted.cu) against the dense path at a k both cover, and at k = 16
Writes one JSON object to gpurun_out/next.json and prints it."""
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from kmap_b200 import engine as E, synth  # noqa: E402


def ev_time(fn, reps=3, warm=1):
    for _ in range(warm):
        out = fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


def fasta_text(seq, borders):
    L = int(borders[0, 1] - borders[0, 0])
    body = np.frombuffer(b"ACGT", dtype=np.uint8)[np.minimum(seq.reshape(-1, L + 1)[:, :L], 3)]
    n = len(body)
    rec = np.concatenate([np.full((n, 1), ord(">"), np.uint8), np.full((n, 1), ord("r"), np.uint8), np.full((n, 1), 10, np.uint8),
                          body, np.full((n, 1), 10, np.uint8)], axis=1)
    return rec.reshape(-1)


def main():
    n_ingest = int(float(sys.argv[1])) if len(sys.argv) > 1 else 4_000_000
    n_sort = int(float(sys.argv[2])) if len(sys.argv) > 2 else 1_000_000
    res = {}
    torch.cuda.set_device(0)
    # ---- ingest ------------------------------------------------------------------------------------------------------
    spec = synth.CFG3
    seq_d, borders_d = synth.generate_device(spec, 0, n_ingest)
    seq, borders = seq_d.cpu().numpy(), borders_d.cpu().numpy()
    text = fasta_text(seq, borders)
    del seq_d, borders_d
    ms_e2e, (s_d, b_d) = ev_time(lambda: E.fasta_text_to_device(text, chunk_bytes=1 << 28))
    assert np.array_equal(s_d.cpu().numpy(), seq) and np.array_equal(b_d.cpu().numpy(), borders)
    fa_path = Path("/tmp/kmap_bench_ingest.fa")
    text.tofile(fa_path)
    ms_file, (s_d, b_d) = ev_time(lambda: E.fasta_to_device(fa_path))
    assert np.array_equal(s_d.cpu().numpy(), seq) and np.array_equal(b_d.cpu().numpy(), borders)
    fa_path.unlink()
    text_d = torch.from_numpy(text).cuda()
    L = E.lib()
    import ctypes
    nc = int(text_d.numel())
    scratch = E.empty(L.kmap_fasta_scratch_words(nc), torch.int64)
    seq_out = E.empty(len(seq), torch.uint8)
    rec = E.empty(len(borders), torch.int64)
    st_in = (ctypes.c_int64 * 4)(0, 0, 0, 10)
    st_out = (ctypes.c_int64 * 4)()

    def parse_resident():
        E.check(L.kmap_fasta_scan(text_d.data_ptr(), nc, st_in, scratch.data_ptr(), st_out, E._stream_ptr()), "scan")
        E.check(L.kmap_fasta_emit(text_d.data_ptr(), nc, st_in, scratch.data_ptr(), seq_out.data_ptr(), 0, rec.data_ptr(), 1, st_out,
                                  E._stream_ptr()), "emit")
    ms_dev, _ = ev_time(parse_resident, reps=5)
    # CPU: the oracle's restatement of the reference's per-record loop on a sample of the same text
    from oracle import kmap_oracle as O
    import tempfile
    n_cpu = min(n_ingest, 200_000)
    with tempfile.NamedTemporaryFile("wb", suffix=".fa", delete=False) as fh:
        fh.write(text[: n_cpu * (len(text) // n_ingest)].tobytes())
    t0 = time.perf_counter()
    O.fasta_to_binary(fh.name)
    cpu_s = time.perf_counter() - t0
    bytes_algo = len(text) + len(seq) + 16 * len(borders)       # text read once (twice by the two passes), outputs written
    res["fasta_ingest"] = {
        "reads": n_ingest, "text_bytes": int(len(text)), "device_ms_text_resident": ms_dev,
        "text_GBs_resident": len(text) / ms_dev / 1e6, "algorithmic_GBs": (2 * len(text) + len(seq) + 16 * len(borders)) / ms_dev / 1e6,
        "ms_from_host_text": ms_e2e, "text_GBs_from_host": len(text) / ms_e2e / 1e6,
        "ms_from_file_page_cache": ms_file, "text_GBs_from_file": len(text) / ms_file / 1e6,
        "cpu_oracle_text_MBs": n_cpu * (len(text) // n_ingest) / cpu_s / 1e6, "cpu_sample_reads": n_cpu, "bytes_model": int(bytes_algo)}
    del text_d, scratch, seq_out, rec, s_d, b_d
    # ---- sorted path vs dense path -----------------------------------------------------------------------------------------
    spec = synth.CFG2
    seq_d, borders_d = synth.generate_device(spec, 0, n_sort)
    dev = E.SeqOnDevice.from_device_u8(seq_d, borders_d)
    n_bases = n_sort * spec.read_len
    out = {"reads": n_sort, "read_len": spec.read_len}
    for k in (14, 16, 20):
        for dedup in (True, False):
            ms, (kh, cnt) = ev_time(lambda: dev.count_sorted(k, dedup))
            ms_m, (mkh, mcnt) = ev_time(lambda: E.merge_revcom_sorted(kh, cnt, k))
            out[f"sorted_k{k}_{'dedup' if dedup else 'rep'}"] = {"count_ms": ms, "merge_ms": ms_m, "n_unique": int(kh.numel()),
                                                                   "n_merged": int(mkh.numel()), "Gbases_per_s": n_bases / ms / 1e6}
            if k == 14:
                ms_d, table = ev_time(lambda: dev.count(k, dedup=dedup))
                ms_c, (dkh, dcnt) = ev_time(lambda: E.compact_merge(table, k, True))
                assert np.array_equal(dkh.cpu().numpy().view(np.uint32).astype(np.uint64), mkh.cpu().numpy().view(np.uint64))
                assert np.array_equal(dcnt.cpu().numpy().astype(np.int64), mcnt.cpu().numpy())
                out[f"dense_k{k}_{'dedup' if dedup else 'rep'}"] = {"count_ms": ms_d, "merge_ms": ms_c, "Gbases_per_s": n_bases / ms_d / 1e6}
    # phases of the sorted count at k = 16 (dedup)
    k = 16
    keys = E.empty(dev.n, torch.int64)
    work = E.empty(L.kmap_dedup_keys_work_words(dev.n_seq), torch.int32)
    ms_keys, _ = ev_time(lambda: E.check(L.kmap_window_keys_u64(dev.packed.data_ptr(), dev.valid.data_ptr(), dev.n, k, keys.data_ptr(), E._stream_ptr())))
    ms_dd, _ = ev_time(lambda: E.check(L.kmap_dedup_hash_per_read_u64(keys.data_ptr(), dev.n, dev.borders.data_ptr(), dev.n_seq, work.data_ptr(), E._stream_ptr())))
    ms_sort, _ = ev_time(lambda: E.sort_count_keys(keys.clone(), 2 * k))
    ms_clone, _ = ev_time(lambda: keys.clone())
    out["phases_k16_dedup_ms"] = {"keys": ms_keys, "dedup": ms_dd, "sort_rle": ms_sort - ms_clone,
                                  "sort_GBs_model": (2 * k + 7) // 8 * 16 * dev.n / (ms_sort - ms_clone) / 1e6}
    # CPU: oracle count at k = 16 on a sample
    n_cpu = min(n_sort, 20000)
    s_np, b_np = synth.generate_numpy(spec, 0, n_cpu)
    t0 = time.perf_counter()
    h = O.comp_kmer_hash(s_np, 16)
    h = O.remove_duplicate_hash_per_seq(h, b_np, O.get_invalid_hash(h.dtype.type))
    O.merge_revcom(*O.count_uniq_hash(h, 16), 16)
    out["cpu_oracle_k16_dedup_Gbases_per_s"] = n_cpu * spec.read_len / (time.perf_counter() - t0) / 1e9
    res["sorted_path"] = out
    Path("gpurun_out").mkdir(exist_ok=True)
    if len(sys.argv) <= 3:                            # (a third argument = a run under a profiler: numbers are not kept)
        Path("gpurun_out/next.json").write_text(json.dumps(res, indent=1))
    print(json.dumps(res))


if __name__ == "__main__":
    main()
