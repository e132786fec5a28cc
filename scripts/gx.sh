python - <<'PY'
import sys, time; sys.path.insert(0, ".")
import torch, numpy as np
from kmap_b200 import engine as E, synth
from kmap_b200._lib import check, lib
from kmap_b200.kmer_count import kmer2hash
seq_d, b_d = synth.generate_device(synth.CFG3, 0, 100_000_000)
dev = E.SeqOnDevice.from_device_u8(seq_d, b_d); del seq_d
L = lib(); k, d = 14, 5; c = int(kmer2hash("GTACGTAGGTCCTA")); n_seq = dev.n_seq
min_dist = E.empty(n_seq, torch.uint8); n_hit = E.empty(n_seq, torch.int32)
def T(name, fn, reps=2):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): r = fn()
    torch.cuda.synchronize(); print(f"{name}: {(time.perf_counter()-t0)/reps*1e3:.1f} ms", flush=True); return r
sp = torch.cuda.current_stream().cuda_stream
T("occurrence_count", lambda: check(L.kmap_occurrence_count(dev.packed.data_ptr(), dev.valid.data_ptr(), dev.borders.data_ptr(), n_seq, k, c, d, 1, min_dist.data_ptr(), n_hit.data_ptr(), sp)))
offsets = T("scan", lambda: E.exclusive_scan_u32(n_hit))
total = int(offsets[-1].item()); pos = E.empty(total, torch.int32)
T("occurrence_fill", lambda: check(L.kmap_occurrence_fill(dev.packed.data_ptr(), dev.valid.data_ptr(), dev.borders.data_ptr(), n_seq, k, c, d, 1, min_dist.data_ptr(), offsets.data_ptr(), pos.data_ptr(), sp)))
T("D2H (3 x .cpu())", lambda: (min_dist.cpu(), offsets.cpu(), pos.cpu()), reps=1)
PY
