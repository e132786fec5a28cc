"""The exchange step alone (torchrun, one rank per GPU): in-place sum of a level-14 sized table (2^28 cells, + the level-13
sized one) over the ranks through (a) the peer-memory exchange of csrc/peer.cu and (b) the NCCL all-reduce, device time as
the max over the ranks.  Cells hold what a shard's table holds (Poisson counts of mean 32 / world, a few large cells).
    python -m torch.distributed.run --nproc-per-node N scripts/bench_exchange.py [cells_log2]"""
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    import torch
    import torch.distributed as dist
    from kmap_b200 import api, engine as E
    ctx = api.DistContext.from_env()
    rank, world = ctx.rank, ctx.world
    ar = ctx.table_allreduce
    bits = int(sys.argv[1]) if len(sys.argv) > 1 else 28
    n = 1 << bits
    flat, _ = ar.alloc_tables(14, 14) if bits == 28 else (None, None)
    if flat is None:
        raise SystemExit("only 2^28 cells")
    g = torch.Generator(device="cuda").manual_seed(7 + rank)
    src = torch.poisson(torch.full((n,), 32.0 / world, device="cuda"), generator=g).to(torch.int32)
    src[12345::1000003] = 1 << 20
    plain = torch.empty_like(src)

    def timed(fn, reps=10):
        for _ in range(3):
            fn()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / reps], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # (the input is restored by a device copy inside the timed region of both arms: 2 GiB of local traffic, the same for both)
    t_copy = timed(lambda: plain.copy_(src))
    t_peer = timed(lambda: (flat.copy_(src), ar(flat)))
    t_nccl = timed(lambda: (plain.copy_(src), ar(plain)))
    same = bool(torch.equal(flat, plain))
    if rank == 0:
        gb = n * 4 / 1e9
        print({"ranks": world, "cells": n, "peer_ms": t_peer - t_copy, "nccl_ms": t_nccl - t_copy, "copy_ms": t_copy, "identical": same,
               "nccl_busbw_GBs": gb * 2 * (world - 1) / world / ((t_nccl - t_copy) * 1e-3),
               "peer_link_GBs_per_direction": n * (world - 1) / world / 1e9 / ((t_peer - t_copy) * 1e-3)})
    ar.check()
    ctx.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
