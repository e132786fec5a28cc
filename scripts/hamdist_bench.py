"""Distance-matrix kernel alone (cfg5: 100 000 distinct 14-mers), CUDA-event timed; also the target of ncu captures."""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
from kmap_b200 import engine as E
from kmap_b200._lib import check, lib
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
rng = np.random.default_rng(20240414); k = 14
khs = np.unique(rng.integers(0, 4 ** k, int(n * 1.01), dtype=np.uint64))[:n].astype(np.uint32); rng.shuffle(khs)
out = E.empty(n * n, torch.uint8)
L = lib()
for name, labels, heads in (("no override", np.full(n, 2, np.int32), [14, 14]), ("labels 0/1/2, head 12 for label 1", rng.integers(0, 3, n).astype(np.int32), [14, 12])):
    kh_d, lab_d, hl_d = E.to_device(khs), E.to_device(labels), E.to_device(np.asarray(heads, np.int32))
    run = lambda: check(L.kmap_hamdist_matrix_u32(kh_d.data_ptr(), lab_d.data_ptr(), n, k, hl_d.data_ptr(), len(heads), 0, n, out.data_ptr(),
                                                  torch.cuda.current_stream().cuda_stream))
    run(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{name}: {ms:.3f} ms  {n*n/ms/1e6:.0f} GB/s written  {n*n/ms*1e3:.3e} pairs/s")
    # the same matrix as an int8 one-hot GEMM on the tcgen05 tensor cores (csrc/hamdist_mma.cu)
    nb = int(L.kmap_hamdist_mma_scratch_bytes(n))
    scr = E.empty(nb, torch.uint8)
    out2 = E.empty(n * n, torch.uint8)
    run2 = lambda: check(L.kmap_hamdist_matrix_onehot_mma(kh_d.data_ptr(), lab_d.data_ptr(), n, k, hl_d.data_ptr(), len(heads), 0, n, out2.data_ptr(),
                                                          scr.data_ptr(), nb, torch.cuda.current_stream().cuda_stream))
    run2(); torch.cuda.synchronize()
    e0.record()
    for _ in range(reps): run2()
    e1.record(); torch.cuda.synchronize()
    ms2 = e0.elapsed_time(e1) / reps
    print(f"  one-hot tcgen05 GEMM: {ms2:.3f} ms  {n*n/ms2/1e6:.0f} GB/s written  {n*n/ms2*1e3:.3e} pairs/s  identical={bool(torch.equal(out, out2))}")
    del out2, scr
