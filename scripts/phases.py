"""Phase times of kmap_count_all_k (library-recorded events) per level-kmax scheme + equality against the prefix-pass scheme."""
import os
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
S = E.SeqOnDevice
for dedup in (True, False):
    ref = None
    for scheme, name in [x for x in ((S.SLOTTED, "slotted"), (S.SORTED, "sorted")) if x[1] in os.environ.get("SCHEMES", "slotted,sorted")]:
        res = []
        for it in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ph = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
            for e in ph:
                e.record()
            e0.record()
            dev.count_all(8, 14, dedup, tables, phase_events=ph, scheme=scheme)
            e1.record()
            torch.cuda.synchronize()
            order = [e0, ph[0], ph[1]] + ([ph[4]] if scheme == S.SORTED else []) + [ph[5], ph[2], ph[3], e1]
            res.append([a.elapsed_time(b) for a, b in zip(order[:-1], order[1:])] + [e0.elapsed_time(e1)])
        r = res[-1]
        names = ["zero", "scan"] + (["hist"] if scheme == S.SORTED else []) + ["partition", "count", "derive", "tail", "total"]
        print(f"dedup={dedup} {name:8s}: " + " ".join(f"{n} {v:.2f}" for n, v in zip(names, r)) +
              f" ms (min total {min(x[-1] for x in res[1:]):.2f})", flush=True)
        if check:
            if ref is None:
                ref = {k: tables[k].clone() for k in range(8, 15)}
                dev.count_all(8, 14, dedup, tables, partitioned=False)
                print("  equal to prefix-pass scheme:", all(torch.equal(ref[k], tables[k]) for k in ref), flush=True)
            else:
                print("  equal to the first scheme:", all(torch.equal(ref[k], tables[k]) for k in ref), flush=True)
