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
torch.cuda.synchronize()
print("done")
