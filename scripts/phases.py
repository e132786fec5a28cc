"""Phase times of kmap_count_all_k (library-recorded events) + equality against the global-atomic prefix-pass scheme."""
import sys
sys.path.insert(0, ".")
import torch
from kmap_b200 import engine as E, synth

n_reads = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
check = len(sys.argv) > 2 and sys.argv[2] == "check"
seq_d, b_d = synth.generate_device(synth.CFG3, 0, n_reads)
dev = E.SeqOnDevice.from_device_u8(seq_d, b_d)
del seq_d
tables = {k: E.zeros(1 << (2 * k), torch.int32) for k in range(8, 15)}
for dedup in (True, False):
    res = []
    for it in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ph = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        for e in ph:
            e.record()
        e0.record()
        dev.count_all(8, 14, dedup, tables, phase_events=ph)
        e1.record()
        torch.cuda.synchronize()
        res.append((e0.elapsed_time(ph[0]), ph[0].elapsed_time(ph[1]), ph[1].elapsed_time(ph[2]), ph[2].elapsed_time(ph[3]), e0.elapsed_time(e1)))
    r = res[-1]
    print(f"dedup={dedup}: zero {r[0]:.2f} scan {r[1]:.2f} count_kmax {r[2]:.2f} derive {r[3]:.2f} total {r[4]:.2f} ms "
          f"(min total {min(x[4] for x in res[1:]):.2f})", flush=True)
    if check:
        ref = {k: tables[k].clone() for k in range(8, 15)}
        dev.count_all(8, 14, dedup, tables, partitioned=False)
        print("  equal to prefix-pass scheme:", all(torch.equal(ref[k], tables[k]) for k in ref), flush=True)
