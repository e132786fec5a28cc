"""Exploration: partitioned (shared-memory) level-k count vs global-atomic schemes, per-kernel times via CUDA events."""
import sys, time
sys.path.insert(0, ".")
import torch
from kmap_b200 import engine as E, synth

n_reads = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
seq_d, b_d = synth.generate_device(synth.CFG3, 0, n_reads)
dev = E.SeqOnDevice.from_device_u8(seq_d, b_d)
del seq_d
tables = {k: E.zeros(1 << (2 * k), torch.int32) for k in range(8, 15)}

def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

for k in (11, 12, 13, 14):
    a = timeit(lambda: dev.count(k, dedup=False, table=tables[k], partitioned=True))
    ref = tables[k].clone()
    b = timeit(lambda: dev.count(k, dedup=False, table=tables[k], partitioned=False), reps=1)
    print(f"rep count k={k}: partitioned {a:8.2f} ms | direct atomics {b:8.2f} ms | equal={torch.equal(ref, tables[k])}", flush=True)
for dedup in (True, False):
    a = timeit(lambda: dev.count_all(8, 14, dedup, tables, partitioned=True))
    ref = {k: tables[k].clone() for k in (8, 11, 14)}
    b = timeit(lambda: dev.count_all(8, 14, dedup, tables, partitioned=False))
    ok = all(torch.equal(ref[k], tables[k]) for k in ref)
    print(f"count_all 8..14 dedup={dedup}: partitioned {a:8.2f} ms | prefix passes {b:8.2f} ms | equal={ok}", flush=True)
    for kmax in (12, 13):
        a = timeit(lambda: dev.count_all(8, kmax, dedup, tables, partitioned=True))
        b = timeit(lambda: dev.count_all(8, kmax, dedup, tables, partitioned=False))
        print(f"count_all 8..{kmax} dedup={dedup}: partitioned {a:8.2f} ms | prefix passes {b:8.2f} ms", flush=True)
