mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "mask or find_motif or scan_motif or smoke" 2>&1 | tail -2
python - <<'PY'
import sys; sys.path.insert(0, ".")
import torch
from kmap_b200 import engine as E, synth
from kmap_b200.kmer_count import kmer2hash, revcom_hash
seq_d, b_d = synth.generate_device(synth.CFG3, 0, 100_000_000)
dev = E.SeqOnDevice.from_device_u8(seq_d, b_d); del seq_d
for k, d, cs in ((14, 5, "GTACGTAGGTCCTA"), (8, 2, "CCTACGTA")):
    c = int(kmer2hash(cs)); rc = int(revcom_hash(c, k))
    dev.snapshot_valid()
    def run():
        dev.restore_valid(); dev.mask(k, [c, rc], [d, d])
    run(); torch.cuda.synchronize()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    for _ in range(3): run()
    e1.record()
    for _ in range(3): dev.restore_valid()
    e2.record(); torch.cuda.synchronize()
    ms = (e0.elapsed_time(e1) - e1.elapsed_time(e2)) / 3
    print(f"mask k={k} d={d}: {ms:.2f} ms  ({dev.n/ms/1e6:.0f} Gpositions/s, {0.5*dev.n/ms/1e6:.0f} GB/s algorithmic)")
    dev.restore_valid()
PY
