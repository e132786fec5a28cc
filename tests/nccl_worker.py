"""Worker of tests/test_gpu_multi.py: launched by torch.distributed.run with one rank per GPU (NCCL).
    mode `count`      every rank counts ITS shard through api.count_kmers(table_allreduce=...); rank 0 compares the merged
                      lists with the oracle's lists of the whole input (kmer_count.py:476-491, 643-685) and, through the
                      same call on one GPU, with the single-rank lists
    mode `scan_motif` kmap scan_motif under torchrun on tests/test.fa: rank 0 compares every output with the reference goldens
"""
import gzip
import os
import pickle
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    import torch
    import torch.distributed as dist
    from kmap_b200 import api, synth
    mode, out_dir = sys.argv[1], Path(sys.argv[2])
    ctx = api.DistContext.from_env()
    rank, world = ctx.rank, ctx.world
    assert world > 1 and torch.cuda.current_device() == int(os.environ["LOCAL_RANK"])
    if mode == "count":
        from oracle import kmap_oracle as O
        n_reads = int(sys.argv[3])
        def with_long_reads():
            """reads of every path of the per-read scan (warp / block / bitmap) in ONE of the two shards only: the rank without
            them must still take part in the exchange of their tables"""
            rng = np.random.default_rng(99)
            reads = ["A" * 80, "CA" * 60, "ACGTTGCA" * 150, ("ACGTAGCTAGCTAGGATCGAT" * 1500)[:30000], "ACGTTGCAAC" * 40, "", "N"]
            reads += [O.arr2dna(rng.integers(0, 4, int(rng.integers(0, 130))).astype(np.uint8)) for _ in range(2500)]
            arrs = [O.dna2arr(r) for r in reads]
            lens = np.array([len(a) for a in arrs])
            ends = np.cumsum(lens)
            return np.concatenate(arrs), np.stack([ends - lens, ends - 1], axis=1).astype(np.int64)
        inputs = [synth.generate_numpy(synth.CFG3, 0, n_reads), synth.generate_numpy(synth.CFG2_N, 0, n_reads), with_long_reads()]
        from kmap_b200 import engine as E
        scat = api.TableAllReduce(scatter=True)
        # the peer-memory exchange (csrc/peer.cu) on a table of every kind of cell: bytes that fit, the escape value itself,
        # words far outside a byte, sums that wrap -- against the modular sum computed from the gathered tables
        ar = ctx.table_allreduce
        flat, _ = ar.alloc_tables(8, 9)
        assert ar.peer_exchange, ar._peer_refused
        g = torch.Generator(device="cuda").manual_seed(1234 + rank)
        n_cells = flat.numel()
        vals = torch.randint(-130, 131, (n_cells,), device="cuda", generator=g, dtype=torch.int64)
        big = torch.randint(0, 1 << 32, (n_cells,), device="cuda", generator=g, dtype=torch.int64)
        pick = torch.randint(0, 50, (n_cells,), device="cuda", generator=g)
        vals = torch.where(pick == 0, big, vals)
        vals[:4096] = big[:4096]                               # a run of nothing but escapes
        vals[4096:8192] = -128
        mine = (vals & 0xFFFFFFFF)
        every = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(every, mine)
        want = (sum(every) & 0xFFFFFFFF)
        for lo, hi in ((0, n_cells), (1 << 16, (1 << 16) + 4096 * world), (16, n_cells - 16)):
            flat.copy_(torch.where(mine >= (1 << 31), mine - (1 << 32), mine).to(torch.int32))
            ar(flat[lo:hi])
            torch.cuda.synchronize()
            got = flat.to(torch.int64) & 0xFFFFFFFF
            assert torch.equal(got[lo:hi], want[lo:hi]), ("peer exchange", lo, hi)
            assert torch.equal(got[:lo], mine[:lo]) and torch.equal(got[hi:], mine[hi:]), ("cells outside the exchanged range", lo, hi)
        sflat, _ = scat.alloc_tables(8, 9)
        sflat.copy_(torch.where(mine >= (1 << 31), mine - (1 << 32), mine).to(torch.int32))
        scat.reduce_scatter(sflat)
        torch.cuda.synchronize()
        blk = n_cells // world
        got = sflat.to(torch.int64) & 0xFFFFFFFF
        assert torch.equal(got[rank * blk:(rank + 1) * blk], want[rank * blk:(rank + 1) * blk]), "peer reduce-scatter"
        ar.check()
        scat.check()
        for seq, borders in inputs:
            s, b = api.shard_reads(seq, borders, rank, world)
            for rep_mode in (False, True):
                # the merge left scattered by key range (kmap_count_all_k_scattered: reduce-scatter + reductions on the owned
                # ranges only): every rank's owned range of every level equals that range of the all-reduced tables
                dev = api.upload_reads(s, b)
                full = dev.count_all(8, 14, dedup=not rep_mode, merge=ctx.table_allreduce)
                part_t = dev.count_all(8, 14, dedup=not rep_mode, merge=scat)
                for k in range(8, 15):
                    lo, hi = scat.owned_range(k)
                    assert torch.equal(part_t[k][lo:hi], full[k][lo:hi]), (k, rep_mode, "scattered merge")
                # the same two merges with the tables in the peer regions: one byte per cell over NVLink peer memory
                peer_t = dev.count_all(8, 14, dedup=not rep_mode, tables=ar.alloc_tables(8, 14)[1], merge=ar)
                peer_s = dev.count_all(8, 14, dedup=not rep_mode, tables=scat.alloc_tables(8, 14)[1], merge=scat)
                for k in range(8, 15):
                    lo, hi = scat.owned_range(k)
                    assert torch.equal(peer_t[k], full[k]), (k, rep_mode, "peer-memory exchange vs NCCL all-reduce")
                    assert torch.equal(peer_s[k][lo:hi], full[k][lo:hi]), (k, rep_mode, "peer-memory scattered merge")
                ar.check()
                del peer_t, peer_s
                t13 = full[13].clone()
                scat.reduce_scatter(t13)                       # (x world on the owned block: every rank holds the merged table)
                lo, hi = scat.owned_range(13)
                assert torch.equal(t13[lo:hi].to(torch.int64), full[13][lo:hi].to(torch.int64) * world)
                del dev, full, part_t, t13
                ks = [8, 11, 13, 14]
                got = api.count_kmers(s, b, list(range(8, 15)) + [16], rep_mode=rep_mode, table_allreduce=ctx.table_allreduce, lists_on=0)
                # every rank compacts and returns one key range of every list: the slices concatenate to the whole list
                part = api.count_kmers(s, b, range(8, 15), rep_mode=rep_mode, table_allreduce=ctx.table_allreduce, lists_on="sharded")
                parts = ctx.gather(part)
                if rank != 0:
                    assert got == {}
                    continue
                for k in range(8, 15):
                    for j in (0, 1):
                        assert np.array_equal(np.concatenate([p_[k][j] for p_ in parts]), got[k][j]), (k, rep_mode, "sharded lists")
                one = api.count_kmers(seq, borders, range(8, 15), rep_mode=rep_mode)
                for k in range(8, 15):
                    assert np.array_equal(got[k][0], one[k][0]) and np.array_equal(got[k][1], one[k][1]), (k, rep_mode, "vs 1 GPU")
                for k in ks + [16]:
                    h = O.comp_kmer_hash(seq, k)
                    if not rep_mode:
                        h = O.remove_duplicate_hash_per_seq(h, borders, O.get_invalid_hash(O.get_hash_dtype(k)))
                    want = O.merge_revcom(*O.count_uniq_hash(h, k), k)
                    g = got[k]
                    assert g[0].dtype == want[0].dtype and np.array_equal(g[0], want[0]) and np.array_equal(g[1], want[1]), (k, rep_mode)
        (out_dir / "count.ok").write_text("ok")
    elif mode == "scan_motif":
        import tomli_w
        import tomllib
        import kmap_b200.kmer_count as K
        import kmap_b200.motif_discovery as MD
        from helpers import assert_occurrence_text_equal
        golden = sys.argv[3]

        def load(name):
            with gzip.open(ROOT / "tests" / "golden" / (name + ".pkl.gz"), "rb") as fh:
                return pickle.load(fh)
        base = load("testfa")
        g = base if golden == "testfa" else load(golden)
        seq, borders = base["input_bin"], base["borders"]
        fa, res_dir = out_dir / "test.fa", out_dir / "res"
        if rank == 0:
            with open(fa, "w") as fh:
                for i, (st, en) in enumerate(borders):
                    fh.write(f">r{i}\n{K.arr2dna(seq[st:en])}\n")
            res_dir.mkdir()
            cfg = tomllib.loads(g["text_files"]["config.toml"])
            cfg["general"]["input_fasta_file"] = str(fa)
            cfg["general"]["res_dir"] = str(res_dir)
            with open(res_dir / "config.toml", "wb") as fh:
                tomli_w.dump(cfg, fh)
            K._preproc(str(fa), str(res_dir))
        ctx.barrier()
        np.random.seed(20240414 if golden == "testfa" else 20240415)      # (the seeds the goldens were recorded with)
        MD._scan_motif(str(res_dir))
        if rank == 0:
            tf = g["text_files"]
            for name in [n for n in tf if n in ("candidate_conseq.csv", "final_conseq.txt", "final_conseq.info.csv") or n.startswith("hamming_balls/")]:
                assert (res_dir / name).read_text() == tf[name], name
            for name in [n for n in tf if n.endswith("motif_occurence.csv")]:
                assert_occurrence_text_equal((res_dir / name).read_text().splitlines(), tf[name])
            lists = {c["k"]: c for c in g["find_motif"] if 8 <= c["k"] <= 14} if golden == "testfa" else g["kmer_count"]
            for k, c in lists.items():
                with open(res_dir / "kmer_count" / f"k{k}.pkl", "rb") as fh:
                    kk, ukh, ucnt = pickle.load(fh)
                assert kk == k and ukh.dtype == c["uniq_kh"].dtype and ucnt.dtype == c["uniq_cnt"].dtype, k
                assert np.array_equal(ukh, c["uniq_kh"]) and np.array_equal(ucnt, c["uniq_cnt"]), k
            with open(res_dir / "sample_kmers.pkl", "rb") as fh:
                skh, scnt, slab, sconseq = pickle.load(fh)
            gkh, gcnt, glab, gconseq = g["sample_kmers"]
            assert sconseq == gconseq and np.array_equal(skh, gkh) and np.array_equal(scnt, gcnt) and np.array_equal(slab, glab)
            with open(res_dir / "sample_kmer_hamdist_mat.pkl", "rb") as fh:
                kk, mat, lab = pickle.load(fh)
            assert kk == g["hamdist"]["k"] and np.array_equal(mat, g["hamdist"]["mat"]) and np.array_equal(lab, g["hamdist"]["labels"])
            (out_dir / "scan_motif.ok").write_text("ok")
    ctx.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
