"""Where does the end-to-end time of api.count_kmers go?"""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from kmap_b200 import engine as E, synth, api
n_reads = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
seq_d, b_d = synth.generate_device(synth.CFG3, 0, n_reads)
seq_host = torch.empty(seq_d.numel(), dtype=torch.uint8, pin_memory=True); seq_host.copy_(seq_d)
b_host = torch.empty(b_d.shape, dtype=torch.int64, pin_memory=True); b_host.copy_(b_d)
torch.cuda.synchronize(); del seq_d, b_d
seq_np, b_np = seq_host.numpy(), b_host.numpy()
def T(label, fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize()
    print(f"{label:40s} {1e3*(time.perf_counter()-t0):9.1f} ms", flush=True); return r
print("from_numpy pinned?", torch.from_numpy(seq_np).is_pinned())
for rep in range(2):
    s = T("H2D seq (from_numpy.to)", lambda: torch.from_numpy(seq_np).to("cuda", non_blocking=True))
    b = T("H2D borders", lambda: torch.from_numpy(b_np).to("cuda", non_blocking=True))
    dev = T("pack", lambda: E.SeqOnDevice.from_device_u8(s, b))
    del s
    tabs = T("count_all 8..14 dedup", lambda: dev.count_all(8, 14, True))
    lists = {}
    for k in range(8, 15):
        lists[k] = T(f"compact_merge k={k}", lambda: E.compact_merge(tabs[k], k, True))
    for k in (8, 14):
        T(f"D2H lists k={k} (pinned)", lambda: (api._to_host_pinned(lists[k][0], np.uint32), api._to_host_pinned(lists[k][1], np.int32)))
    T("whole api.count_kmers", lambda: api.count_kmers(seq_np, b_np, range(8, 15), validate=False))
    del dev, tabs, lists
