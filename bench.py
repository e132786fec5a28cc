#!/usr/bin/env python
"""Headline benchmark of kmap_b200: Gbases/s of k-mers counted for k = 8..14 (BASELINE.json metric) on the
ChIP-like synthetic workload (1e8 reads x 100 bp), sharded by reads over N GPUs of one
node with an NCCL all-reduce of every dense 4^k table.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [...]                          # the reference's CPU algorithm (oracle port)

One step = one pass of the counting hot path over the whole input: for every k in 8..14 zero the uint32[4^k] table,
count every read's distinct k-mers into it (the reference's default, non-repetitive mode: comp_kmer_hash +
remove_duplicate_hash_per_seq + count_uniq_hash fused), and for N > 1 all-reduce the table.  `value` is measured
with the packed reads resident in HBM; `e2e` is the same work through the public host-buffer API
(`kmap_b200.api.count_kmers`: pinned uint8 input.bin + int64 borders H2D, pack, count, order-exact compaction +
reverse-complement merge, the merged (kh, cnt) lists D2H).  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

KMIN, KMAX = 8, 14
METRIC = "gbases_per_s_kmers_counted_k8_14"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--reads", type=float, default=1e8, help="total reads over all GPUs (cfg3: 1e8)")
    ap.add_argument("--read-len", type=int, default=100)
    ap.add_argument("--mode", default="dedup", choices=["dedup", "rep"], help="dedup = reference default (repetitive_mode=false)")
    ap.add_argument("--algo", default="allk", choices=["allk", "perk"],
                    help="allk = one atomic pass at k=14 + 4:1 table reductions (csrc/count_all.cu); perk = 7 independent passes")
    ap.add_argument("--partitions", type=int, default=0, help="key-range passes for the k=14 table (0 = auto)")
    ap.add_argument("--cpu-sample-reads", type=int, default=40000)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-crosscheck", action="store_true", help="skip the full-size all-k vs direct per-k table comparison")
    ap.add_argument("--no-hamdist", action="store_true", help="skip the distance-matrix leg (10 GB / n_gpus of output per GPU)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-piece2", action="store_true", help="skip the Hamming-ball / compaction / mask / find_motif leg")
    ap.add_argument("--no-workflow", action="store_true", help="skip the cfg2 preproc + scan_motif wall-clock leg")
    ap.add_argument("--extras", action="store_true", help="also time compaction / Hamming-ball / mask / distance-matrix kernels")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


FOLD_RUN_ENDS_DEFAULT = "2"      # the library's default of KMAP_FOLD_RUN_ENDS (csrc/partition.cu: kmap_fold_run_ends)


def launches_per_step(args, dedup):
    """kernels of libkmap_b200 launched per step (memsets not counted)"""
    if args.algo != "allk":
        return KMAX - KMIN + 1
    derive = KMAX - KMIN
    if args.partitions <= 0:     # dedup_scan + hist + 3 scan kernels + partition + bucket_count + bucket_segments + derive
        fold = derive if int(os.environ.get("KMAP_FOLD_RUN_ENDS", FOLD_RUN_ENDS_DEFAULT)) >= 1 else 0    # + one fold_corrections_kernel per lower level
        return (1 if dedup else 0) + 7 + derive + fold
    return (1 if dedup else 0) + 5 + max(1, args.partitions) + derive      # + terminal-correction launches + prefix passes


def table_checksum(table):
    """sum over h of h * T[h] mod 2^64: a position-weighted checksum of a dense table.  Equal checksums on every level prove
    what equal sums cannot: that the merged tables of 1, 2, 4 and 8 GPUs hold the same counts in the same cells."""
    import torch
    total, step = 0, 1 << 26
    for lo in range(0, table.numel(), step):
        t = table[lo:lo + step].view(torch.int32).to(torch.int64) & 0xFFFFFFFF
        idx = torch.arange(lo, lo + t.numel(), dtype=torch.int64, device=t.device)
        total += int((t * idx).sum().item())
    return total & 0xFFFFFFFFFFFFFFFF


def n_kmers_total(n_reads, L):
    return sum(n_reads * max(0, L - k + 1) for k in range(KMIN, KMAX + 1))


def algorithmic_bytes_count(n_reads, L, k):
    """SURVEY.md section 8d per-unit figure for one counting launch: 0.375 B per base (2-bit code + validity bit) +
    8 B per counted k-mer (uint32 read-modify-write).  The 8 * 4^k table zero/read-out term belongs to the memset and
    the compaction, not to this kernel."""
    return 0.375 * n_reads * L + 8.0 * n_reads * max(0, L - k + 1)


# ----------------------------------------------------------------------------------------------------------------------
def cpu_reference_step(seq, borders, mode):
    """the reference's CPU algorithm for the same step (oracle port): per k hash every position, de-duplicate per read
    (the reference's per-read Python loop), np.unique-count.  Returns seconds."""
    from oracle import kmap_oracle as O
    t0 = time.perf_counter()
    for k in range(KMIN, KMAX + 1):
        h = O.comp_kmer_hash(seq, k)
        if mode == "dedup":
            h = O.remove_duplicate_hash_per_seq(h, borders, np.uint32(0xFFFFFFFF))
        O.count_uniq_hash(h, k)
    return time.perf_counter() - t0


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from kmap_b200 import synth
    spec = synth.CFG3 if args.read_len == 100 else synth.SynthSpec(seed=20240413, read_len=args.read_len)
    n_sample = int(min(args.cpu_sample_reads, args.reads))
    seq, borders = synth.generate_numpy(spec, 0, n_sample)
    for _ in range(args.warmup):
        cpu_reference_step(seq, borders, args.mode)
    times = [cpu_reference_step(seq, borders, args.mode) for _ in range(args.steps)]
    t = float(np.mean(times))
    bases = n_sample * args.read_len
    value = bases * (KMAX - KMIN + 1) / t / 1e9
    sample = f"first {n_sample} reads x {args.read_len} bp of the same synthetic workload, k={KMIN}..{KMAX}, mode={args.mode}"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Gbases/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "u32", "data": "synthetic",
        "config": {"workload": f"cfg3 ChIP-like {int(args.reads)} reads x {args.read_len} bp, dense 4^k counting k={KMIN}..{KMAX}, "
                               f"{args.mode} mode (bounded CPU sample)", "sample": sample},
        "cpu_baseline": {"value": value, "unit": "Gbases/s", "cores": 1, "kind": "port", "sample": sample,
                         "host_cpus": os.cpu_count()},
        "e2e": {"value": value, "unit": "Gbases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def run_workflow_cfg2(n_reads=1_000_000):
    """BASELINE config 2 through the reference-shaped drivers: 1e6 synthetic HT-SELEX-like reads x 40 bp as a FASTA file ->
    `kmap preproc` -> `kmap scan_motif` (k = 8..14, stock settings otherwise); wall clock per stage."""
    import shutil
    import tempfile
    try:
        import tomli_w
        import torch
        from kmap_b200 import kmer_count as K, motif_discovery as MD, synth
        spec = synth.CFG2
        seq, _ = synth.generate_numpy(spec, 0, n_reads)
        Lr = spec.read_len
        body = np.frombuffer(b"ACGT", dtype=np.uint8)[np.minimum(seq.reshape(-1, Lr + 1)[:, :Lr], 3)]
        rec = np.concatenate([np.full((n_reads, 1), ord(">"), np.uint8), np.full((n_reads, 1), ord("r"), np.uint8),
                              np.full((n_reads, 1), 10, np.uint8), body, np.full((n_reads, 1), 10, np.uint8)], axis=1)
        tmp = Path(tempfile.mkdtemp())
        fa = tmp / "reads.fa"
        rec.reshape(-1).tofile(fa)
        res = tmp / "res"
        res.mkdir()
        cfg = K.read_default_config_file()
        cfg["kmer_count"]["min_k"], cfg["kmer_count"]["max_k"] = 8, 14
        cfg["motif_discovery"]["motif_pos_density_flag"] = False
        cfg["motif_discovery"]["motif_co_occurence_flag"] = False
        cfg["general"]["input_fasta_file"] = str(fa)
        cfg["general"]["res_dir"] = str(res)
        with open(res / "config.toml", "wb") as fh:
            tomli_w.dump(cfg, fh)
        out = {"reads": n_reads, "read_len": Lr, "fasta_MB": fa.stat().st_size / 1e6}
        t = time.perf_counter()
        K._preproc(str(fa), str(res))
        torch.cuda.synchronize()
        out["preproc_s"] = time.perf_counter() - t
        np.random.seed(1)
        import contextlib
        import io
        t = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()):
            MD._scan_motif(str(res))
        torch.cuda.synchronize()
        out["scan_motif_s"] = time.perf_counter() - t
        out["final_conseq"] = (res / "final_conseq.txt").read_text().split()
        out["candidate_rows"] = len((res / "candidate_conseq.csv").read_text().splitlines()) - 1
        shutil.rmtree(tmp, ignore_errors=True)
        return out
    except Exception as exc:
        return {"error": f"{type(exc).__name__}: {exc}"[:400]}


# ----------------------------------------------------------------------------------------------------------------------
def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import torch.distributed as dist
    from kmap_b200 import api, engine as E, synth
    from kmap_b200._lib import check, lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    n_gpus = world
    L = args.read_len
    n_total = int(args.reads)
    r0 = n_total * rank // world
    r1 = n_total * (rank + 1) // world
    n_local = r1 - r0
    spec = synth.CFG3 if L == 100 else synth.SynthSpec(seed=20240413, read_len=L)
    dedup = args.mode == "dedup"

    # ---- device-resident input (generated on the device by the counter-based generator) -----------------------------
    seq_d, borders_d = synth.generate_device(spec, r0, n_local)
    dev = E.SeqOnDevice.from_device_u8(seq_d, borders_d)
    # N > 1: the tables live in this rank's peer region, so the merge is the one-byte-per-cell exchange over NVLink peer
    # memory (csrc/peer.cu); the NCCL all-reduce of the same step is timed beside it further down
    merge_comm = api.TableAllReduce() if world > 1 else None
    flat_tables, tables = merge_comm.alloc_tables(KMIN, KMAX, zero=True) if world > 1 else E.alloc_tables(KMIN, KMAX, zero=True)
    torch.cuda.synchronize()

    count_events = []
    phase_events = []

    def step(record=False):
        if args.algo == "allk":
            ph = None
            if record:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ph = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
                for e in ph:
                    e.record()                      # creates the cudaEvent_t handles the library re-records
                e0.record()
            # N > 1: the tables are all-reduced from inside the count as they become final (kmap_count_all_k_sharded)
            dev.count_all(KMIN, KMAX, dedup, tables, n_partitions=args.partitions, phase_events=ph, merge=merge_comm)
            if record:
                e1.record()
                count_events.append(("all", e0, e1))
                phase_events.append((e0, ph, e1))
            return
        for k in range(KMIN, KMAX + 1):
            if record:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            dev.count(k, dedup=dedup, table=tables[k], zero=True)
            if record:
                e1.record()
                count_events.append((k, e0, e1))
            if world > 1:
                merge_comm(tables[k])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        step(record=True)
    ev1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms_local = ev0.elapsed_time(ev1)
    t = torch.tensor([ms_local], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_per_step = ms_total / args.steps
    value = n_total * L * (KMAX - KMIN + 1) / (ms_per_step * 1e-3) / 1e9

    # per-k count time (memset + kernel) on this rank, averaged over the timed steps
    per_k = {}
    for k, e0, e1 in count_events:
        per_k.setdefault(k, []).append(e0.elapsed_time(e1))
    per_k_ms = {k: float(np.mean(v)) for k, v in per_k.items()}

    # sanity of the timed work (integer identities, not timing): sum of the rep-mode table == number of valid windows
    checks = {}
    tot14 = int(tables[KMAX].to(torch.int64).sum().item())
    checks["sum_table_k14"] = tot14
    # per-level position-weighted checksums of the (merged) tables: identical for every number of GPUs iff the tables are
    checks["table_checksums"] = {str(k): table_checksum(tables[k]) for k in range(KMIN, KMAX + 1)}
    # N > 1: the same step with the merged tables left scattered over the ranks by key range (reduce-scatter instead of
    # all-reduce: half the exchange volume, kmap_count_all_k_scattered).  Reported beside `value`, which keeps the all-reduce
    # the north star names.  The owned ranges of all ranks together must give the checksums of the all-reduced tables.
    exchange = None
    if world > 1:
        merge_comm.check()
        exchange = {"product": "peer memory, one byte per cell (csrc/peer.cu)" if merge_comm.peer_exchange else "NCCL all-reduce",
                    "peer_exchange_refused": merge_comm._peer_refused}
    if args.algo == "allk" and world > 1 and merge_comm.peer_exchange:
        # the same step with the tables outside the peer region: the NCCL all-reduce of round 1 / the first half of round 2
        _, nc_tables = E.alloc_tables(KMIN, KMAX, zero=True)
        for _ in range(max(1, args.warmup)):
            dev.count_all(KMIN, KMAX, dedup, nc_tables, n_partitions=args.partitions, merge=merge_comm)
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(args.steps):
            dev.count_all(KMIN, KMAX, dedup, nc_tables, n_partitions=args.partitions, merge=merge_comm)
        s1.record()
        barrier()
        ts = torch.tensor([s0.elapsed_time(s1) / args.steps], device="cuda", dtype=torch.float64)
        dist.all_reduce(ts, op=dist.ReduceOp.MAX)
        same = all(bool(torch.equal(nc_tables[k], tables[k])) for k in range(KMIN, KMAX + 1))
        flag = torch.tensor([int(same)], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        exchange["nccl_allreduce_ms_per_step"] = float(ts.item())
        exchange["nccl_allreduce_value"] = n_total * L * (KMAX - KMIN + 1) / (float(ts.item()) * 1e-3) / 1e9
        exchange["tables_identical"] = bool(flag.item())
        checks["peer_exchange_equals_nccl_allreduce"] = bool(flag.item())
        del nc_tables
    scattered = None
    if args.algo == "allk" and world > 1 and (1 << (2 * KMIN)) % world == 0:
        rs_comm = api.TableAllReduce(scatter=True)
        _, rs_tables = rs_comm.alloc_tables(KMIN, KMAX, zero=True)
        for _ in range(max(1, args.warmup)):
            dev.count_all(KMIN, KMAX, dedup, rs_tables, n_partitions=args.partitions, merge=rs_comm)
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(args.steps):
            dev.count_all(KMIN, KMAX, dedup, rs_tables, n_partitions=args.partitions, merge=rs_comm)
        s1.record()
        barrier()
        ts = torch.tensor([s0.elapsed_time(s1) / args.steps], device="cuda", dtype=torch.float64)
        dist.all_reduce(ts, op=dist.ReduceOp.MAX)
        same = True
        for k in range(KMIN, KMAX + 1):
            lo, hi = rs_comm.owned_range(k)
            same = same and bool(torch.equal(rs_tables[k][lo:hi], tables[k][lo:hi]))
        flag = torch.tensor([int(same)], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        scattered = {"ms_per_step": float(ts.item()), "value": n_total * L * (KMAX - KMIN + 1) / (float(ts.item()) * 1e-3) / 1e9,
                     "unit": "Gbases/s", "owned_ranges_equal_allreduced_tables": bool(flag.item()),
                     "note": "reduce-scatter by key range: rank r owns cells [r 4^k / N, (r+1) 4^k / N) of every level"}
        checks["scattered_merge_equals_allreduce"] = bool(flag.item())
        rs_comm.close()
        del rs_tables
    if args.algo == "allk" and world == 1 and not args.no_crosscheck:
        # full-size parity property: the tables the all-k algorithm DERIVES (level 14 -> 13 -> .. -> 8, with the run-end and
        # repeat corrections of every level on the way) equal independent direct counts by the per-k kernels
        for kc in (KMIN, KMAX - 1):
            direct = dev.count(kc, dedup=dedup)
            checks[f"allk_table_k{kc}_equals_direct_count"] = bool(torch.equal(direct, tables[kc]))
            del direct

    # ---- roofline of the dominant kernel (the counting kernel), measured live with CUDA events ----------------------
    peaks_file = ROOT / "MEASURED_PEAKS.json"
    if peaks_file.exists():
        peak, peak_src = json.loads(peaks_file.read_text())["hbm_gbs"], "MEASURED_PEAKS.json hbm_gbs (of measured)"
    else:
        peak, peak_src = 6650.0, "B200_PROFILING.md fallback (of fallback)"
    alg_bytes = sum(algorithmic_bytes_count(n_local, L, k) for k in range(KMIN, KMAX + 1))
    kern_s = sum(per_k_ms.values()) * 1e-3
    traffic = None
    tfile = ROOT / "profiles" / "count_traffic.json"      # dram__bytes_read+write of the dominant kernel, one ncu --set full capture
    if tfile.exists():
        try:
            tj = json.loads(tfile.read_text())
            # the capture is taken on a smaller input (ncu replays the kernel ~40 times): scale per position
            traffic = tj["dram_bytes_per_launch"] * (n_local * (L + 1)) / tj["positions"]
        except Exception:
            traffic = None
    if args.algo == "allk":
        # phases of kmap_count_all_k from the events the library records on the launching stream
        partitioned = args.partitions <= 0
        names = ["zero", "dedup_scan", "bucket_hist", "partition", "bucket_count", "derive", "tail"] if partitioned else \
                ["zero", "dedup_scan", "count_kmax", "derive", "tail"]
        ph_ms = {k: [] for k in names}
        for e0, ph, e1 in phase_events:
            order = [e0, ph[0], ph[1], ph[4], ph[5], ph[2], ph[3], e1] if partitioned else [e0, ph[0], ph[1], ph[2], ph[3], e1]
            for name, a, b in zip(names, order[:-1], order[1:]):
                ph_ms[name].append(a.elapsed_time(b))
        ph_ms = {k: float(np.mean(v)) for k, v in ph_ms.items()}
        n_pos_local = dev.n
        n_win = n_local * max(0, L - KMAX + 1)
        if partitioned:
            # dominant kernel: partition_kernel, one launch per step.  Algorithmic bytes of THIS launch: it reads the packed
            # bases, the validity bits and (dedup mode) the hidden-window bits once, the per-tile bucket counts + offsets of
            # pass 1 (6 B x 4096 buckets per tile of 32768 positions = 0.75 B per position) and writes one 16-bit key suffix
            # per counted k=14 window
            in_b = (0.375 + (0.125 if dedup else 0.0) + 0.75) * n_pos_local
            b_launch = in_b + 2.0 * n_win
            t_launch = ph_ms["partition"] * 1e-3
            kernel = "partition_kernel<4> (1 launch per step, k=14)"
            # the three launches that together are `count at k=14` against SURVEY 8d's figure for one counting pass
            t_pipe = (ph_ms["bucket_hist"] + ph_ms["partition"] + ph_ms["bucket_count"]) * 1e-3
            b_pipe = algorithmic_bytes_count(n_local, L, KMAX)
            pipeline = {"launches": "bucket_hist_kernel + bucket_scan_kernel + partition_kernel + bucket_count_kernel",
                        "algorithmic_bytes": b_pipe, "ms": t_pipe * 1e3, "achieved_GBs": b_pipe / t_pipe / 1e9,
                        "frac": b_pipe / t_pipe / 1e9 / peak,
                        "note": "SURVEY 8d per-unit figure for one counting pass (0.375 B/base + 8 B per k-mer) over the three "
                                "launches that build the k=14 table"}
        else:
            n_pass = max(1, args.partitions)
            b_launch = 0.375 * n_pos_local + 8.0 * n_win / n_pass
            t_launch = ph_ms["count_kmax"] * 1e-3 / n_pass
            kernel = f"count_prefix_kernel ({n_pass} launches per step, k={KMAX})"
            pipeline = None
        achieved = b_launch / t_launch / 1e9
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                    "kernel": kernel, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": b_launch, "avg_launch_ms": t_launch * 1e3,
                    "share_of_step": t_launch * 1e3 / (kern_s * 1e3), "phases_ms": ph_ms,
                    "phases_note": "bucket_hist = histogram pass + run-end detection; bucket_count = per-bucket count + the reductions of "
                                   "the folded run-end corrections" + (" + the exchange of the tables over the ranks" if world > 1 else "")
                                   + "; zero = fills on the counting stream (the level-14 table is filled beside the per-read scan)",
                    "k14_count_pipeline": pipeline,
                    "step_model": {"note": "SURVEY 8d model for 7 independent per-k passes (0.375 B/base + 8 B/k-mer each) over the "
                                           "measured step time; the all-k algorithm updates the k=14 table only and derives the "
                                           "smaller tables by 4:1 reductions, so this can exceed what 7 passes could reach",
                                   "algorithmic_bytes_per_step": alg_bytes, "achieved_GBs": alg_bytes / kern_s / 1e9,
                                   "frac": alg_bytes / kern_s / 1e9 / peak},
                    "frac_of_8TBs_nominal": achieved / 8000.0, "algo": "allk",
                    "bound_note": "the level-14 count is bound by the shared-memory / LSU pipe (three shared-memory atomics per window "
                                  "over the three launches, a random STS and 7 sectors per store request in the write-out), not "
                                  "by HBM: see DESIGN.md section 4.5"}
    else:
        achieved = alg_bytes / kern_s / 1e9
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                    "kernel": "count_dedup_warp_kernel" if dedup else "count_dense_kernel", "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": alg_bytes / len(per_k_ms), "avg_launch_ms": kern_s * 1e3 / len(per_k_ms),
                    "per_k_ms": {str(k): v for k, v in per_k_ms.items()}, "frac_of_8TBs_nominal": achieved / 8000.0, "algo": "perk"}

    # ---- second half of the metric: pairwise Hamming-distance matrix of 100 000 sampled 14-mers (BASELINE config 5), rows
    # partitioned over the ranks in contiguous blocks, no collective (cal_samp_kmer_hamdist_mat, motif_discovery.py:759-808)
    hamdist = None
    if not args.no_hamdist:
        from kmap_b200.motif_discovery import hamdist_formulation, hamdist_matrix_u8
        rng = np.random.default_rng(20240414)
        n_h, k_h = 100_000, 14
        khs = np.unique(rng.integers(0, 4 ** k_h, int(n_h * 1.01), dtype=np.uint64))[:n_h].astype(np.uint32)
        rng.shuffle(khs)
        labels_h = rng.integers(0, 3, n_h).astype(np.int32)
        row0, row1 = api.row_range(n_h, rank, world)

        def time_matrix(impl, out_buf):
            hamdist_matrix_u8(khs, labels_h, [14, 12], k_h, row0, row1, out=out_buf, impl=impl)         # warm-up
            barrier()
            h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            h0.record()
            for _ in range(5):
                hamdist_matrix_u8(khs, labels_h, [14, 12], k_h, row0, row1, out=out_buf, impl=impl)
            h1.record()
            barrier()
            th = torch.tensor([h0.elapsed_time(h1) / 5], device="cuda", dtype=torch.float64)
            if world > 1:
                dist.all_reduce(th, op=dist.ReduceOp.MAX)
            return float(th.item())

        # the product path (hamdist_formulation picks the kernel) and both formulations by name: XOR/popcount (csrc/hamdist.cu)
        # and the int8 one-hot GEMM on the tcgen05 tensor cores (csrc/hamdist_mma.cu); the outputs must be bit-identical
        out_h = E.empty((row1 - row0) * n_h, torch.uint8)
        out_g = E.empty((row1 - row0) * n_h, torch.uint8)
        chosen = hamdist_formulation(n_h, k_h, [14, 12])
        ms_h = time_matrix(None, out_h)
        ms_pop = time_matrix("popcount", out_g)
        same = bool(torch.equal(out_g, out_h))
        ms_mma = time_matrix("onehot_mma", out_g)
        same = same and bool(torch.equal(out_g, out_h))
        checks["hamdist_onehot_gemm_equals_popcount"] = same
        diag = out_h.view(row1 - row0, n_h)[torch.arange(min(row1 - row0, 1000), device="cuda"), torch.arange(row0, row0 + min(row1 - row0, 1000), device="cuda")]
        checks["hamdist_diag_zero"] = bool((diag == 0).all().item())
        hamdist = {"metric": "hamdist_pairs_per_s", "value": n_h * n_h / ms_h * 1e3, "unit": "pairs/s", "ms": ms_h, "pairs": n_h * n_h,
                   "n_kmers": n_h, "k": k_h, "rows_per_gpu": row1 - row0, "partition": "contiguous row blocks, no collective",
                   "formulation": chosen, "bytes_written_per_gpu": (row1 - row0) * n_h, "write_GBs_per_gpu": (row1 - row0) * n_h / ms_h / 1e6,
                   "frac_of_hbm_copy_peak": (row1 - row0) * n_h / ms_h / 1e6 / peak,
                   "formulations_ms": {"popcount": ms_pop, "onehot_tcgen05_gemm": ms_mma, "identical_output": same},
                   "mma": "tcgen05.mma.cta_group::1.kind::i8, M=128 N=256 K=64..128, S32 accumulators in TMEM (2 x 256 columns), "
                          "cp.async.bulk operand tiles, TMA tensor stores",
                   "note": "uint8 output, 1 B per pair; same-label pairs of the 12-base consensus use the head distance (extra K columns "
                           "in the GEMM, masked flags in the popcount kernel); every figure includes the H2D of the keys and labels and, "
                           "for the GEMM, the operand preparation; the write-only fill rate of this GPU is in fill_rate_GBs"}
        del out_h, out_g

    # ---- end-to-end through the public API with host buffers -------------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        seq_host = torch.empty(seq_d.numel(), dtype=torch.uint8, pin_memory=True)
        seq_host.copy_(seq_d)
        borders_host = torch.empty(borders_d.shape, dtype=torch.int64, pin_memory=True)
        borders_host.copy_(borders_d)
        torch.cuda.synchronize()
        del seq_d
        dev.seq_u8 = None
        seq_np, borders_np = seq_host.numpy(), borders_host.numpy()
        comm = merge_comm

        def e2e_step():
            res = api.count_kmers(seq_np, borders_np, range(KMIN, KMAX + 1), rep_mode=not dedup, revcom_mode=True, validate=False,
                                  table_allreduce=comm, lists_on="sharded" if world > 1 else None)
            return sum(a.nbytes + b.nbytes for a, b in res.values()) if res else 0

        e2e_step()
        barrier()
        t0 = time.perf_counter()
        n_e2e = max(1, min(args.steps, 3))
        d2h = 0
        for _ in range(n_e2e):
            d2h = e2e_step()
        barrier()
        dt = (time.perf_counter() - t0) / n_e2e
        tt = torch.tensor([dt], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        stats = api.last_stream_stats or {}
        h2d = stats.get("h2d_bytes", seq_np.nbytes + borders_np.nbytes)     # bytes that crossed the link (chunks re-encoded by the host travel packed)
        io_bytes = torch.tensor([h2d, d2h], device="cuda", dtype=torch.int64)
        if world > 1:
            dist.all_reduce(io_bytes)                  # bytes copied by all ranks together
        e2e = {"value": n_total * L * (KMAX - KMIN + 1) / dt / 1e9, "unit": "Gbases/s",
               "h2d_bytes_per_step": int(io_bytes[0].item()), "d2h_bytes_per_step": int(io_bytes[1].item()),
               "ms_per_step": dt * 1e3, "steps": n_e2e, "host_input_bytes": int(seq_np.nbytes + borders_np.nbytes) * world,
               "feeders": {k_: stats.get(k_) for k_ in ("chunks", "raw_chunks")},
               "transport": "chunks of reads re-encoded by the host cores into the packed form (0.375 B/position, csrc/host_pack.cpp) "
                            "or shipped as they are, whichever feeder reaches them first; border rows as uint32 strides",
               "api": "kmap_b200.api.count_kmers(seq_np_arr, boarder_mat, k=8..14) -> {k: (uniq_kh_arr, uniq_kh_cnt_arr)}"
                      + ("; every rank uploads its shard of the reads and returns the key-range slice of every merged list "
                         "(lists_on='sharded': the slices of ranks 0..N-1 concatenate to the reference's list); "
                         "h2d / d2h bytes = sums over the ranks" if world > 1 else "")}

    # ---- piece 2 (Hamming-ball aggregation, compaction, mask, occurrence scan, find_motif) on the same resident workload ----
    piece2 = None
    if rank == 0 and not args.no_piece2 and L == 100 and n_local * 8 >= n_total:
        from bench_extras import fill_rate_gbs, run_piece2
        try:
            piece2 = run_piece2(dev, tables, n_local, L, peak)
            if hamdist is not None:
                gbs, ms_fill = fill_rate_gbs()
                hamdist["fill_rate_GBs"] = gbs
                hamdist["frac_of_fill_rate"] = hamdist["write_GBs_per_gpu"] / gbs
        except Exception as exc:                                   # a secondary measurement must never break the headline line
            piece2 = {"error": f"{type(exc).__name__}: {exc}"[:400]}
    workflow = None
    if rank == 0 and world == 1 and not args.no_workflow:
        workflow = run_workflow_cfg2()

    extras = None
    if args.extras and rank == 0:
        from bench_extras import run_extras
        extras = run_extras(dev, tables, n_local, L)

    # ---- CPU baseline on a bounded sample (rank 0, N = 1 only) --------------------------------------------------------------
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu:
        n_sample = int(min(args.cpu_sample_reads, n_total))
        seq_s, borders_s = synth.generate_numpy(spec, 0, n_sample)
        tcpu = cpu_reference_step(seq_s, borders_s, args.mode)
        cpu_baseline = {"value": n_sample * L * (KMAX - KMIN + 1) / tcpu / 1e9, "unit": "Gbases/s", "cores": 1, "kind": "port",
                        "sample": f"first {n_sample} reads x {L} bp of the same workload, k={KMIN}..{KMAX}, {args.mode} mode, "
                                  f"NumPy oracle port of the reference (single-threaded np.unique + per-read Python loop)",
                        "seconds": tcpu, "host_cpus": os.cpu_count()}

    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": "Gbases/s", "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u32",
            "data": "synthetic",
            "config": {"workload": f"cfg3 ChIP-like {n_total} reads x {L} bp (10% carry GTACGTAGGTCCTA, 5% mutation), dense 4^k "
                                   f"counting k={KMIN}..{KMAX}, {args.mode} mode, reads sharded over {n_gpus} GPU(s) with NCCL table merge",
                       "l2": "inputs larger than L2 (packed reads + borders = %.1f GB per GPU)" % ((dev.packed.numel() * 4 + dev.valid.numel() * 4 + n_local * 16) / 1e9),
                       "parallelism": f"reads x{n_gpus}"},
            "clocks": clocks, "e2e": e2e, "exchange": exchange, "scattered_merge": scattered, "gpu_launches": args.steps * (launches_per_step(args, dedup)
                                          # + the exchange over peer memory: narrow / reduce / widen and the three barriers between them
                                          + (6 if world > 1 and args.algo == "allk" and merge_comm.peer_exchange else 0)),
            "roofline": roofline, "cpu_baseline": cpu_baseline, "hamdist": hamdist, "hamball": piece2, "workflow_cfg2": workflow,
            "checks": checks,
        }
        if extras:
            out["extras"] = extras
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
