"""Small invocations of the kernels added or restructured late in the round, for compute-sanitizer (memcheck / racecheck)."""
import sys
import tempfile
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from kmap_b200 import engine as E, synth  # noqa: E402
from kmap_b200 import motif_discovery as MD, kmer_count as K  # noqa: E402

n_reads = int(float(sys.argv[1])) if len(sys.argv) > 1 else 3000
seq, borders = synth.generate_numpy(synth.CFG2_N, 0, n_reads)
dev = E.SeqOnDevice.from_numpy(seq, borders)
tabs = dev.count_all(8, 14, dedup=True)                 # dedup_scan (batches of 32 reads), routed level-13 corrections
tabs2 = dev.count_all(9, 13, dedup=False)
kh, cnt = E.compact_merge(tabs[14], 14, True)
print("merged k=14:", int(kh.numel()))
for k in (16, 21):
    a, b = dev.count_sorted(k, dedup=True)              # keys, per-read de-duplication, radix sort, run lengths
    m = E.merge_revcom_sorted(a, b, k)
    print("sorted k", k, int(a.numel()), int(m[0].numel()))
    s = E.hamball_sums_list64(m[0], m[1], k, [int(x) for x in E.to_host(m[0][:3], np.uint64)], 3, True)
txt = b">a\nACGTN\nacgt\r\n>b desc\n\n>c\nGGGTTTAAACCC" * 700
with tempfile.NamedTemporaryFile("wb", suffix=".fa", delete=False) as fh:
    fh.write(txt)
for chunk in (1 << 28, 1 << 16):
    s_d, b_d = E.fasta_to_device(fh.name, chunk_bytes=chunk)
print("fasta:", int(s_d.numel()), tuple(b_d.shape))
val, idx = E.topk_candidates(cnt, 6)
mdd = K.init_motif_def_dict(ROOT / "kmap_b200" / "default_motif_def_table.csv")
khh = E.to_host(kh, np.uint32)
lab = MD.label_kmers(khh, ["AATCGATAGC"], 14, mdd, True)
# round 2: consumers of the occurrence scan, host encoders + border strides, streamed count, both distance-matrix formulations
from kmap_b200 import api  # noqa: E402
conseqs = ["AATCGATAGC", "AGGACCTACGTAC"]
scan = [E.occurrence_scan_device(dev, len(c), int(K.kmer2hash(c)), mdd[len(c)].max_ham_dist, True) for c in conseqs]
counts, pairs, over = E.co_occurrence_scan(scan)
print("co-occurrence:", counts.tolist(), len(over), E.sum_counts(cnt))
res = api.count_kmers(seq, borders, range(8, 15), chunk_positions=len(seq) // 5 + 1)
print("streamed:", {k_: len(v[0]) for k_, v in res.items()})
rng = np.random.default_rng(0)
kk = rng.integers(0, 4 ** 14, 3000, dtype=np.uint64).astype(np.uint32)
lb = rng.integers(0, 3, 3000)
a1 = MD.hamdist_matrix_u8(kk, lb, [14, 12], 14, 5, 2900, impl="popcount")
a2 = MD.hamdist_matrix_u8(kk, lb, [14, 12], 14, 5, 2900, impl="onehot_mma")
a3 = MD.hamdist_matrix_u8(kk, lb, [14, 2, 3], 14, impl="onehot_mma")
print("hamdist equal:", bool(torch.equal(a1, a2)), tuple(a3.shape))
torch.cuda.synchronize()
print("done")
