"""One all-k count (plus one per-k count) on a small input, for ncu captures."""
import sys
sys.path.insert(0, ".")
import torch
from kmap_b200 import engine as E, synth
n_reads = int(float(sys.argv[1])) if len(sys.argv) > 1 else 2_000_000
dedup = (sys.argv[2] != "rep") if len(sys.argv) > 2 else True
seq_d, b_d = synth.generate_device(synth.CFG3, 0, n_reads)
dev = E.SeqOnDevice.from_device_u8(seq_d, b_d)
tables = {k: E.zeros(1 << (2 * k), torch.int32) for k in range(8, 15)}
for _ in range(2):
    dev.count_all(8, 14, dedup, tables)
torch.cuda.synchronize()
print("done")
