"""Host-side logic that needs no GPU: how api.count_kmers cuts the reads into chunks for the streamed upload."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from kmap_b200 import api, synth  # noqa: E402
from oracle import kmap_oracle as O  # noqa: E402


def test_chunk_bounds_cut_at_read_starts_and_cover_everything():
    seq, borders = synth.generate_numpy(synth.CFG2_N, 0, 1000)          # 1000 reads x 40 bp (+ separators)
    n = len(seq)
    cuts = api._chunk_bounds(seq, borders, n // 6)
    assert cuts[0] == 0 and cuts[-1] == len(borders) and cuts == sorted(set(cuts)) and len(cuts) >= 6
    for r0, r1 in zip(cuts[:-1], cuts[1:]):
        p0, p1 = int(borders[r0, 0]), int(borders[r1 - 1, 1]) + 1
        assert seq[p1 - 1] == 255 and (p0 == 0 or seq[p0 - 1] == 255)     # a chunk starts after a separator, ends on one
        assert p1 - p0 <= n // 6 + 2 * 41
    # chunk tables add up to the table of the whole input (reads are independent units, kmer_count.py:755-759)
    k = 5
    whole = np.zeros(4 ** k, dtype=np.int64)
    u, c = O.count_uniq_hash(O.remove_duplicate_hash_per_seq(O.comp_kmer_hash(seq, k), borders, np.uint32(0xFFFFFFFF)), k)
    whole[u] = c
    acc = np.zeros(4 ** k, dtype=np.int64)
    for r0, r1 in zip(cuts[:-1], cuts[1:]):
        p0, p1 = int(borders[r0, 0]), int(borders[r1 - 1, 1]) + 1
        b = borders[r0:r1] - p0
        u, c = O.count_uniq_hash(O.remove_duplicate_hash_per_seq(O.comp_kmer_hash(seq[p0:p1], k), b, np.uint32(0xFFFFFFFF)), k)
        acc[u] += c
    assert np.array_equal(acc, whole)


def test_chunk_bounds_declines_small_or_foreign_layouts():
    seq, borders = synth.generate_numpy(synth.CFG2_N, 0, 200)
    assert api._chunk_bounds(seq, borders, len(seq)) is None             # fits one chunk
    assert api._chunk_bounds(seq, None, 100) is None
    gappy = borders.copy()
    gappy[:, 1] -= 1                                                     # not the back-to-back layout preproc writes
    assert api._chunk_bounds(seq, gappy, len(seq) // 4) is None
    shifted = borders.copy()
    shifted[0, 0] = 1
    assert api._chunk_bounds(seq, shifted, len(seq) // 4) is None


def test_read_fasta_bytes_starts_at_first_header(tmp_path):
    """host side of the device FASTA ingest: the text handed to the parser starts at the first header line
    (text before it is ignored like Bio.SeqIO does), plain and gzipped"""
    import gzip
    from kmap_b200 import engine as E
    cases = {b">a\nACGT\n": b">a\nACGT\n", b"junk\n>a\nAC": b">a\nAC", b"x>y\r>b\nT": b">b\nT", b"no header": b"", b"": b"",
             b"\n\n>": b">"}
    for i, (raw, want) in enumerate(cases.items()):
        p = tmp_path / f"c{i}.fa"
        p.write_bytes(raw)
        assert bytes(E.read_fasta_bytes(p)) == want
        with gzip.open(tmp_path / f"c{i}.fa.gz", "wb") as fh:
            fh.write(raw)
        assert bytes(E.read_fasta_bytes(tmp_path / f"c{i}.fa.gz")) == want


def test_native_occurrence_writer_equals_python_formatter(tmp_path):
    """kmap_write_occurrence_rows (host-only C++) writes the *.motif_occurence.csv rows exactly as the per-read Python
    formatter (reference motif_discovery.py:1409-1418, 1472-1475), including the rows whose > 20 hits need the random
    pick from numpy's global RNG (same stream consumption), and the returned counts equal get_motif_seq_num of the file"""
    from kmap_b200 import motif_discovery as MD
    rng = np.random.default_rng(0)
    n_seq, m = 3000, 3
    lens = rng.integers(0, 200, n_seq)
    ends = np.cumsum(lens + 1)
    borders = np.stack([ends - lens - 1, ends - 1], axis=1).astype(np.int64)
    per = []
    for j in range(m):
        cnt = rng.integers(0, 4, n_seq) * (rng.random(n_seq) < 0.3)
        if j == 1:
            cnt[[5, 700, 2999]] = 25
        off = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
        pos = np.concatenate([np.sort(rng.choice(300, c, replace=False)) for c in cnt]).astype(np.int32)
        per.append((None, off, pos))
    conseqs = ["ACG", "TTGA", "CC"]
    np.random.seed(7)
    lines = ["seq_ind;" + ";".join(f"motif_{i}_{conseqs[i]}" for i in range(m)) + ";seq_len"]
    has = np.zeros(n_seq, bool)
    for _, o, _p in per:
        has |= np.diff(o) > 0
    for r in np.flatnonzero(has):
        _, cells = MD._cells_for_read(per, r)
        lines.append(f"{r};{cells};{lens[r]}")
    np.random.seed(7)
    stats = MD.write_motif_occurence_file(per, borders, conseqs, tmp_path / "o.csv")
    assert (tmp_path / "o.csv").read_text() == "\n".join(lines) + "\n"
    assert stats == [MD.get_motif_seq_num(tmp_path / "o.csv", i) for i in range(m)]
    assert MD.write_motif_occurence_file([], borders, [], tmp_path / "e.csv") == []
    assert (tmp_path / "e.csv").read_text() == "seq_ind;;seq_len\n"


def _cooc_of_scan_numpy(per, m, n_seq):
    """what engine.co_occurrence_scan returns (csrc/consumers.cu), restated with numpy for the host-side merge test"""
    cnts = np.stack([np.diff(o) for _, o, _ in per])
    over = np.flatnonzero((cnts > 20).any(axis=0)).astype(np.int64)
    ok = ~(cnts > 20).any(axis=0)
    counts = np.zeros((m, m), dtype=np.int64)
    pairs = {}

    def med2(j, r):
        o, p = per[j][1], per[j][2]
        v = p[o[r]:o[r + 1]]
        h = len(v) // 2
        return 2 * int(v[h]) if len(v) % 2 else int(v[h - 1]) + int(v[h])
    for i in range(m):
        counts[i, i] = np.count_nonzero((cnts[i] > 0) & ok)
        for j in range(i + 1, m):
            reads = np.flatnonzero((cnts[i] > 0) & (cnts[j] > 0) & ok).astype(np.int64)
            counts[i, j] = len(reads)
            pairs[(i, j)] = (reads, np.array([med2(j, r) - med2(i, r) for r in reads], dtype=np.int32))
    return counts, pairs, over


def test_consumers_from_scan_merge_logic(tmp_path):
    """co_occurrence_from_scan / pos_density_from_scan (host side: device results + the rows of the random pick, one or
    several shards) == the reference-shaped functions on the file the same scan results were written to"""
    from kmap_b200 import motif_discovery as MD
    rng = np.random.default_rng(3)
    n_seq, m = 2500, 4
    lens = rng.integers(30, 200, n_seq)
    ends = np.cumsum(lens + 1)
    borders = np.stack([ends - lens - 1, ends - 1], axis=1).astype(np.int64)
    per = []
    for j in range(m):
        cnt = rng.integers(0, 5, n_seq) * (rng.random(n_seq) < 0.4)
        if j in (1, 2):
            cnt[[5 + j, 700, 2499]] = 23
        off = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
        pos = np.concatenate([np.sort(rng.choice(25, c, replace=False)) for c in cnt]).astype(np.int32)
        per.append((None, off, pos))
    conseqs = ["ACGT", "TTGA", "CCAGG", "GGGTTT"]
    picked = {}
    np.random.seed(9)
    MD.write_motif_occurence_file(per, borders, conseqs, tmp_path / "o.csv", picked)
    assert sorted(picked) == [6, 7, 700, 2499]
    want = MD.get_motif_co_occurence_mat(tmp_path / "o.csv", m)
    info = {"scan": per, "lens": (borders[:, 1] - borders[:, 0]).astype(np.int64), "picked_rows": picked}
    # one shard, and three shards gathered in rank order
    cuts = [0, 900, 901, n_seq]
    shards = []
    for a, b in zip(cuts[:-1], cuts[1:]):
        sub = [(None, o[a:b + 1] - o[a], p[o[a]:o[b]]) for _, o, p in per]
        counts, pairs, over = _cooc_of_scan_numpy(sub, m, b - a)
        shards.append((counts, {k_: (r + a, d) for k_, (r, d) in pairs.items()}, over + a))
    for cooc in ([_cooc_of_scan_numpy(per, m, n_seq)], shards):
        got = MD.co_occurrence_from_scan(dict(info, cooc=cooc), m)
        assert got[0].dtype == want[0].dtype and np.array_equal(got[0], want[0])
        assert np.array_equal(got[1], want[1]) and got[2] == want[2]
    for i, c in enumerate(conseqs):
        a = MD.get_motif_pos_density(tmp_path / "o.csv", i, len(c))
        b = MD.pos_density_from_scan(info, i, len(c))
        assert a[:2] == b[:2] and np.array_equal(a[2], b[2])


def test_hamdist_formulation_choice():
    """which distance-matrix kernel the product picks (motif_discovery.hamdist_formulation): the tcgen05 GEMM for 32-bit hashes
    when the override columns fit into K = 128 and the sample is large, the XOR/popcount kernel otherwise"""
    from kmap_b200 import motif_discovery as MD
    assert MD.hamdist_formulation(100_000, 14, [14, 12]) == "onehot_mma"          # BASELINE config 5
    assert MD.hamdist_formulation(100_000, 14, []) == "onehot_mma"
    assert MD.hamdist_formulation(100_000, 16, [16, 12, 12, 12, 12]) == "onehot_mma"   # 16 + 4 * 4 = 32 slots: K = 128 exactly
    assert MD.hamdist_formulation(100_000, 16, [16, 12, 12, 12, 11]) == "popcount"     # 33 slots
    assert MD.hamdist_formulation(100_000, 17, [17]) == "popcount"                # uint64 hashes
    assert MD.hamdist_formulation(2047, 14, [14]) == "popcount" and MD.hamdist_formulation(2048, 14, [14]) == "onehot_mma"
