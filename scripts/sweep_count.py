"""Exploration: time the all-k hierarchical count for several partition counts, against the per-k kernels."""
import sys, time, json
sys.path.insert(0, ".")
import torch
from kmap_b200 import engine as E, synth

n_reads = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
spec = synth.CFG3
seq_d, b_d = synth.generate_device(spec, 0, n_reads)
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

ref = {}
for dedup in (True, False):
    for k in (12, 13, 14):
        ms = timeit(lambda: dev.count(k, dedup=dedup, table=tables[k]))
        print(f"per-k dedup={dedup} k={k}: {ms:.1f} ms", flush=True)
    ref[dedup] = {k: tables[k].clone() for k in (12, 13, 14)}
    for parts in (0, 1, 4, 16, 64):
        ms = timeit(lambda: dev.count_all(8, 14, dedup, tables, n_partitions=parts))
        ok = all(torch.equal(tables[k], ref[dedup][k]) for k in (12, 13, 14))
        print(f"all-k dedup={dedup} partitions={parts}: {ms:.1f} ms  equal_to_per_k={ok}", flush=True)
    for kmax in (12, 13):
        ms = timeit(lambda: dev.count_all(8, kmax, dedup, tables, n_partitions=0))
        print(f"all-k dedup={dedup} k=8..{kmax} auto partitions: {ms:.1f} ms", flush=True)
