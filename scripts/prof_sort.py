"""One sort + run-length count of the k = 16 window keys of 1e6 synthetic reads x 40 bp (the sort path of csrc/sorted.cu);
target of ncu launch lists / captures."""
import sys
sys.path.insert(0, ".")
import torch
from kmap_b200 import engine as E, synth
from kmap_b200._lib import check, lib
n_reads = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
L = lib()
seq_d, b_d = synth.generate_device(synth.CFG2, 0, n_reads)
dev = E.SeqOnDevice.from_device_u8(seq_d, b_d)
k = 16
keys = E.empty(dev.n, torch.int64)
check(L.kmap_window_keys_u64(dev.packed.data_ptr(), dev.valid.data_ptr(), dev.n, k, keys.data_ptr(), E._stream_ptr()))
work = E.empty(L.kmap_dedup_keys_work_words(dev.n_seq), torch.int32)
check(L.kmap_dedup_hash_per_read_u64(keys.data_ptr(), dev.n, dev.borders.data_ptr(), dev.n_seq, work.data_ptr(), E._stream_ptr()))
for _ in range(2):
    kh, cnt = E.sort_count_keys(keys.clone(), 2 * k)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
kc = keys.clone()
e0.record()
kh, cnt = E.sort_count_keys(kc, 2 * k)
e1.record()
torch.cuda.synchronize()
print("keys", dev.n, "unique", int(kh.numel()), "sort + run lengths", round(e0.elapsed_time(e1), 3), "ms")
