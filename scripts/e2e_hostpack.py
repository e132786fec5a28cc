"""End-to-end api.count_kmers at cfg3 with the two feeders of count_tables_streamed: host re-encoding on/off, chunk size,
host threads; + the bare rate of the host encoder.  Usage: python scripts/e2e_hostpack.py [n_reads]"""
import os, sys, time, json
sys.path.insert(0, ".")
import numpy as np, torch
from kmap_b200 import engine as E, synth, api
from kmap_b200._lib import lib
L = lib()
n_reads = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
seq_d, b_d = synth.generate_device(synth.CFG3, 0, n_reads)
seq_host = torch.empty(seq_d.numel(), dtype=torch.uint8, pin_memory=True); seq_host.copy_(seq_d)
b_host = torch.empty(b_d.shape, dtype=torch.int64, pin_memory=True); b_host.copy_(b_d)
torch.cuda.synchronize(); del seq_d, b_d
seq_np, b_np = seq_host.numpy(), b_host.numpy()
out = {"n_reads": n_reads, "host_threads": int(L.kmap_host_threads())}
n = min(len(seq_np), 1 << 31)
nw = int(L.kmap_valid_words(n))
hp = torch.empty(2 * nw, dtype=torch.int32, pin_memory=True); hv = torch.empty(nw, dtype=torch.int32, pin_memory=True)
rates = {}
for th in (1, 4, 8, 16, 32):
    if th > 2 * out["host_threads"]:
        continue
    L.kmap_host_pack2bit(seq_np.ctypes.data, n, hp.data_ptr(), hv.data_ptr(), th)
    t0 = time.perf_counter()
    L.kmap_host_pack2bit(seq_np.ctypes.data, n, hp.data_ptr(), hv.data_ptr(), th)
    rates[th] = n / (time.perf_counter() - t0) / 1e9
out["host_pack_GBs_by_threads"] = rates
print(json.dumps(out), flush=True)
del hp, hv
ref = None
res = {}
for name, kw in [("slots6_raw1", dict(slots=6, max_raw=1)), ("slots6_raw2", dict(slots=6, max_raw=2)), ("slots6_raw1_16thr", dict(slots=6, max_raw=1, threads=16))]:
    os.environ["KMAP_HOST_THREADS"] = str(kw.pop("threads", 14))
    os.environ["KMAP_STREAM_SLOTS"] = str(kw.pop("slots"))
    os.environ["KMAP_STREAM_MAX_RAW"] = str(kw.pop("max_raw"))
    kw.setdefault("chunk_positions", 1 << 29)
    kw["host_pack"] = True
    ts = []
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = api.count_kmers(seq_np, b_np, range(8, 15), validate=False, **kw)
        torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    chk = {k: int(np.sum(r[k][1], dtype=np.int64)) ^ int(np.bitwise_xor.reduce(r[k][0])) for k in r}
    if ref is None:
        ref = chk
    res[name] = {"ms": [round(1e3 * t, 1) for t in ts], "same_lists_digest": chk == ref, "last": api.last_stream_stats}
    print(name, res[name], flush=True)
    del r
os.environ["KMAP_STREAM_TRACE"] = "1"
os.environ["KMAP_HOST_THREADS"] = "14"
torch.cuda.synchronize(); t0 = time.perf_counter()
r = api.count_kmers(seq_np, b_np, range(8, 15), validate=False, host_pack=True, chunk_positions=1 << 29)
torch.cuda.synchronize(); print("traced call", round(1e3 * (time.perf_counter() - t0), 1), "ms")
for row in api.last_stream_trace:
    if row[1] in ("copy-engine", "kmap-host-pack"):
        print(row)
out["e2e_ms"] = res
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/e2e_hostpack.json", "w"), indent=1)
