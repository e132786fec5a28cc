"""Host logic of the multi-GPU path, exercised with world_size = 2 on CPU (gloo): read sharding, the integer table
all-reduce wrapper and the row-block partition.  The per-shard tables come from the oracle here (no GPU in this suite);
tests/test_gpu_parity.py checks the same identity with the CUDA kernels on one device."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    from kmap_b200 import api, synth
    from oracle import kmap_oracle as O
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        spec = synth.CFG2_N
        seq, borders = synth.generate_numpy(spec, 0, 301)            # odd number of reads: uneven shards
        s, b = api.shard_reads(seq, borders, rank, world)
        k = 6
        h = O.remove_duplicate_hash_per_seq(O.comp_kmer_hash(s, k), b, np.uint32(0xFFFFFFFF))
        u, c = O.count_uniq_hash(h, k)
        table = np.zeros(4 ** k, dtype=np.uint32)
        table[u] = c
        if rank == 0:
            table[5] += np.uint32(0xFFFFFFF0)                        # modular uint32 sums must survive the int32 transport
        if rank == 1:
            table[5] += np.uint32(0x00000020)
        t = torch.from_numpy(table.view(np.int32).copy())
        api.TableAllReduce()(t)
        q.put((rank, t.numpy().view(np.uint32).copy(), api.row_range(1001, rank, world)))
    finally:
        dist.destroy_process_group()


def test_sharded_count_and_table_allreduce_world2():
    import torch.multiprocessing as mp
    from kmap_b200 import synth
    from oracle import kmap_oracle as O
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    seq, borders = synth.generate_numpy(synth.CFG2_N, 0, 301)
    k = 6
    h = O.remove_duplicate_hash_per_seq(O.comp_kmer_hash(seq, k), borders, np.uint32(0xFFFFFFFF))
    u, c = O.count_uniq_hash(h, k)
    want = np.zeros(4 ** k, dtype=np.uint32)
    want[u] = c
    want[5] += np.uint32(0x10)                                        # 0xFFFFFFF0 + 0x20 mod 2^32
    rows = {}
    for rank, table, rr in got:
        assert np.array_equal(table, want), f"rank {rank}: merged table differs from the single-shard count"
        rows[rank] = rr
    assert rows[0][0] == 0 and rows[0][1] == rows[1][0] and rows[1][1] == 1001


@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
def test_shards_tile_the_input(world):
    from kmap_b200 import api, synth
    seq, borders = synth.generate_numpy(synth.CFG2, 0, 37)
    pieces, n_reads = [], 0
    for rank in range(world):
        s, b = api.shard_reads(seq, borders, rank, world)
        pieces.append(s)
        n_reads += len(b)
        if len(b):
            assert b[0, 0] == 0 and b[-1, 1] == len(s) - 1 and np.all(s[b[:, 1]] == 255)
    assert n_reads == 37 and np.array_equal(np.concatenate(pieces), seq)
    with pytest.raises(Exception):
        api.read_range(10, world, world)


def test_more_ranks_than_reads():
    from kmap_b200 import api, synth
    seq, borders = synth.generate_numpy(synth.CFG2, 0, 2)
    sizes = [len(api.shard_reads(seq, borders, r, 5)[1]) for r in range(5)]
    assert sum(sizes) == 2 and max(sizes) == 1


def _ctx_worker(rank, world, port, q):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch.distributed as dist
    from kmap_b200 import api
    ctx = api.DistContext.from_env()                                  # no CUDA here: joins with gloo
    try:
        agreed = ctx.agree({"have": rank == 0, "k": [8, 9]})          # rank 0's view of the files wins
        # per-shard occurrence-scan results (min_dist, offsets, positions) of ranks with 3 and 2 reads
        part = (np.array([1, 255, 0], np.uint8), np.array([0, 2, 2, 3], np.int64), np.array([4, 9, 1], np.int32)) if rank == 0 else \
               (np.array([255, 2], np.uint8), np.array([0, 0, 4], np.int64), np.array([0, 5, 6, 7], np.int32))
        gathered = ctx.gather(part)
        merged = api.concat_occurrence_shards(gathered) if ctx.is_root else None
        ctx.barrier()
        q.put((rank, ctx.world, ctx.is_root, agreed, merged))
    finally:
        dist.destroy_process_group()


def test_dist_context_agree_gather_and_occurrence_concat_world2():
    """host side of `torchrun ... -m kmap_b200 scan_motif`: rank 0's decisions are broadcast, per-read scan results of the
    shards are concatenated in rank order with rebased offsets (occurrence scan shards by reads)"""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_ctx_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = {r[0]: r for r in (q.get(timeout=120) for _ in procs)}
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert got[0][1] == got[1][1] == 2 and got[0][2] and not got[1][2]
    assert got[0][3] == got[1][3] == {"have": True, "k": [8, 9]}
    assert got[1][4] is None
    md, off, pos = got[0][4]
    assert md.tolist() == [1, 255, 0, 255, 2] and off.tolist() == [0, 2, 2, 3, 3, 7] and pos.tolist() == [4, 9, 1, 0, 5, 6, 7]


def test_dist_context_single_process_is_a_noop():
    from kmap_b200 import api
    ctx = api.DistContext()
    assert ctx.world == 1 and ctx.rank == 0 and ctx.is_root and ctx.table_allreduce is None
    assert ctx.agree(5) == 5 and ctx.gather("x") == ["x"]
    ctx.barrier()
