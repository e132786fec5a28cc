"""Bit-exact parity on BASELINE's own configurations at their full size ("bit-exact parity asserted
on cfg1, cfg2 and a prefix of cfg3"): the CUDA path against digests of the ORACLE's results recorded by
tests/golden/make_golden_cfg.py (the oracle needs minutes per configuration; the inputs are regenerated bit-identically
here by the counter-based generator, whose device and NumPy twins are compared in test_synth_device_matches_numpy).

  cfg2 / cfg2-N : 1e6 HT-SELEX-like reads x 40 bp (0.1 % N in cfg2-N), k = 8..14, both repetitive modes, the whole find_motif
                  loop: first merged lists, accepted consensus k-mers with their float statistics, the masked reads
  cfg3 prefix   : the first 1e6 reads x 100 bp (3 083 partition tiles, dedup scan, routed level 13): forward and merged
                  lists of every k through count_all, the per-k count, and the streamed host-buffer API

Run on the B200 box:  python -m pytest tests -m gpu -x -q
"""
import hashlib
import json
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

DIGESTS = json.loads((Path(__file__).resolve().parent / "golden" / "cfg_digests.json").read_text())
KS = list(range(8, 15))


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def table_checksum(table) -> int:
    """sum over h of h * T[h] mod 2^64 -- the per-level checksum bench.py prints (`checks.table_checksums`)"""
    import torch
    t = table.view(torch.int32).to(torch.int64) & 0xFFFFFFFF
    idx = torch.arange(t.numel(), dtype=torch.int64, device=t.device)
    return int((t * idx).sum().item()) & 0xFFFFFFFFFFFFFFFF


@pytest.fixture(scope="module")
def ENG():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import kmap_b200.engine as E
    return E


@pytest.fixture(scope="module")
def cfg3_prefix(ENG):
    from kmap_b200 import synth
    want = DIGESTS["cfg3_prefix"]
    seq_d, borders_d = synth.generate_device(synth.CFG3, 0, want["n_reads"])
    seq = seq_d.cpu().numpy()
    assert sha(seq) == want["sha_input"]
    return seq_d, borders_d, seq, borders_d.cpu().numpy()


@pytest.mark.parametrize("rep_mode", [False, True])
def test_cfg3_prefix_count_all_vs_oracle(ENG, cfg3_prefix, rep_mode):
    """count_all (dedup scan + bucket histogram + partition + per-bucket count + derived levels) over 3 083 tiles: the
    forward (hash, count) lists, the merged lists and the table checksums of every k equal the oracle's"""
    seq_d, borders_d, _, _ = cfg3_prefix
    want = DIGESTS["cfg3_prefix"]
    dev = ENG.SeqOnDevice.from_device_u8(seq_d, borders_d)
    tables = dev.count_all(8, 14, dedup=not rep_mode)
    for k in KS:
        w = want[f"rep{int(rep_mode)}_k{k}"]
        kh, cnt = ENG.compact_merge(tables[k], k, revcom=False)
        assert kh.numel() == w["n_fwd"], k
        assert sha(ENG.to_host(kh, np.uint32)) == w["sha_fwd_kh"] and sha(ENG.to_host(cnt, np.int32)) == w["sha_fwd_cnt"], k
        assert table_checksum(tables[k]) == w["weighted_checksum"], k
        kh, cnt = ENG.compact_merge(tables[k], k, revcom=True)
        assert kh.numel() == w["n_merged"], k
        assert sha(ENG.to_host(kh, np.uint32)) == w["sha_merged_kh"] and sha(ENG.to_host(cnt, np.int32)) == w["sha_merged_cnt"], k


@pytest.mark.parametrize("k", [8, 12, 13, 14])
def test_cfg3_prefix_per_k_count_vs_oracle(ENG, cfg3_prefix, k):
    """SeqOnDevice.count(k, dedup) -- what find_motif's recounts and single-k callers run -- on the same prefix"""
    seq_d, borders_d, _, _ = cfg3_prefix
    dev = ENG.SeqOnDevice.from_device_u8(seq_d, borders_d)
    for rep_mode in (False, True):
        w = DIGESTS["cfg3_prefix"][f"rep{int(rep_mode)}_k{k}"]
        kh, cnt = ENG.compact_merge(dev.count(k, dedup=not rep_mode), k, revcom=True)
        assert sha(ENG.to_host(kh, np.uint32)) == w["sha_merged_kh"] and sha(ENG.to_host(cnt, np.int32)) == w["sha_merged_cnt"], rep_mode


@pytest.mark.parametrize("rep_mode", [False, True])
def test_cfg3_prefix_streamed_api_vs_oracle(ENG, cfg3_prefix, rep_mode):
    """api.count_kmers from HOST buffers, streamed through the device in 5 chunks of whole reads"""
    from kmap_b200 import api
    _, _, seq, borders = cfg3_prefix
    res = api.count_kmers(seq, borders, KS, rep_mode=rep_mode, chunk_positions=len(seq) // 5 + 1)
    for k in KS:
        w = DIGESTS["cfg3_prefix"][f"rep{int(rep_mode)}_k{k}"]
        assert res[k][0].dtype == np.uint32 and res[k][1].dtype == np.int32
        assert sha(res[k][0]) == w["sha_merged_kh"] and sha(res[k][1]) == w["sha_merged_cnt"], k


def _find_motif_config(ENG, name, spec, rep_mode):
    import kmap_b200.kmer_count as kc
    import kmap_b200.motif_discovery as md
    from kmap_b200 import synth
    if name not in DIGESTS:
        pytest.skip(f"no digests for {name} (tests/golden/make_golden_cfg.py {name})")
    want = DIGESTS[name]
    seq_d, borders_d = synth.generate_device(spec, 0, want["n_reads"])
    seq0 = seq_d.cpu().numpy()
    assert sha(seq0) == want["sha_input"]
    mdd = kc.init_motif_def_dict(Path(kc.__file__).resolve().parent / "default_motif_def_table.csv")
    dev = ENG.SeqOnDevice.from_device_u8(seq_d, borders_d)
    dev.snapshot_valid()
    # what scan_motif does: every first count in one all-k pass, then the find_motif loop per k on the same device copy
    first_tables = dev.count_all(8, 14, dedup=not rep_mode)
    for k in KS:
        w = want[f"rep{int(rep_mode)}_k{k}"]
        m = mdd[k]
        dev.restore_valid()
        found, first = md.find_motif_on_device(dev, k, m.max_ham_dist, m.p_uniform, m.ratio_mu, m.ratio_std, m.ratio_cutoff,
                                               rep_mode=rep_mode, first_table=first_tables[k])
        assert len(first[0]) == w["n_first"] and int(np.sum(first[1], dtype=np.int64)) == w["sum_cnt"], k
        assert sha(first[0]) == w["sha_kh"] and sha(first[1]) == w["sha_cnt"], k
        got = [[int(h), float(v[0]), float(v[1]), float(v[2])] for h, v in found.items()]
        assert [g[0] for g in got] == [x[0] for x in w["found"]], (k, got, w["found"])
        for g, x in zip(got, w["found"]):
            assert g[1] == x[1] and g[2] == x[2], (k, g, x)                       # integer ratios of exact counts
            assert g[3] == pytest.approx(x[3], rel=1e-12, abs=1e-12), (k, g, x)     # scipy logsf
        work = seq0.copy()
        dev.seq_u8 = None                       # (masked_seq_to_numpy masks the device copy in place: start from the clean reads)
        dev.masked_seq_to_numpy(work)
        assert sha(work) == w["sha_masked_seq"], k


@pytest.mark.parametrize("rep_mode", [False, True])
def test_cfg2_find_motif_vs_oracle(ENG, rep_mode):
    from kmap_b200 import synth
    _find_motif_config(ENG, "cfg2", synth.CFG2, rep_mode)


@pytest.mark.parametrize("rep_mode", [False, True])
def test_cfg2n_find_motif_vs_oracle(ENG, rep_mode):
    from kmap_b200 import synth
    _find_motif_config(ENG, "cfg2n", synth.CFG2_N, rep_mode)


def test_cfg2_host_api_find_motif_one_k(ENG, tmp_path):
    """the reference-shaped entry point (find_motif on host arrays + pkl files, :594-702) at cfg2 size for one k"""
    import pickle
    import kmap_b200.kmer_count as kc
    import kmap_b200.motif_discovery as md
    from kmap_b200 import synth
    want = DIGESTS["cfg2"]
    seq, borders = synth.generate_numpy(synth.CFG2, 0, want["n_reads"])
    bfile = tmp_path / "input.seqboarder.bin.pkl"
    with open(bfile, "wb") as fh:
        pickle.dump(borders, fh)
    mdd = kc.init_motif_def_dict(Path(kc.__file__).resolve().parent / "default_motif_def_table.csv")
    k = 13
    m, w = mdd[k], want[f"rep0_k{k}"]
    found = md.find_motif(seq, k, m.max_ham_dist, m.p_uniform, m.ratio_mu, m.ratio_std, m.ratio_cutoff,
                          kmer_cnt_pkl_file=tmp_path / f"k{k}.pkl", boarder_pkl_file=bfile)
    assert [int(h) for h in found] == [x[0] for x in w["found"]]
    assert sha(seq) == w["sha_masked_seq"]
    with open(tmp_path / f"k{k}.pkl", "rb") as fh:
        kk, kh, cnt = pickle.load(fh)
    assert kk == k and sha(kh) == w["sha_kh"] and sha(cnt) == w["sha_cnt"]
