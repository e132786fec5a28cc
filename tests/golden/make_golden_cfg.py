#!/usr/bin/env python
"""Digests of the ORACLE's results on BASELINE's own configurations at full size (too slow to recompute inside the GPU
tests: the per-read Python loop of remove_duplicate_hash_per_seq alone takes ~10 s per k at 1e6 reads):

  cfg2 / cfg2-N : 1e6 synthetic HT-SELEX-like reads x 40 bp (kmap_b200.synth.CFG2 / CFG2_N), k = 8..14, both
                  repetitive modes: the whole find_motif loop (first merged lists, accepted consensus k-mers with their
                  float statistics, the masked sequence array)
  cfg3 prefix   : the first 1e6 reads x 100 bp of the ChIP-like workload (synth.CFG3), k = 8..14, both modes: the forward
                  (hash, count) lists and the merged lists of the first count

    python tests/golden/make_golden_cfg.py [cfg2] [cfg2n] [cfg3]      # writes / updates cfg_digests.json

The oracle (oracle/kmap_oracle.py) is itself pinned to the unmodified reference by tests/test_oracle_golden.py; the
inputs are regenerated bit-identically on the GPU box by the counter-based generator (host twin: synth.generate_numpy).
"""
from __future__ import annotations

import hashlib
import json
import sys
import time
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))
from kmap_b200 import synth  # noqa: E402
from oracle import kmap_oracle as O  # noqa: E402

OUT = HERE / "cfg_digests.json"
KS = list(range(8, 15))


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def find_motif_digests(spec, n_reads):
    mdd = O.init_motif_def_dict(ROOT / "kmap_b200" / "default_motif_def_table.csv")
    seq0, borders = synth.generate_numpy(spec, 0, n_reads)
    out = {"n_reads": n_reads, "sha_input": sha(seq0)}
    for rep_mode in (False, True):
        for k in KS:
            t0 = time.time()
            m = mdd[k]
            work = seq0.copy()
            found, first = O.find_motif(work, k, m.max_ham_dist, m.p_uniform, m.ratio_mu, m.ratio_std, m.ratio_cutoff,
                                        rep_mode=rep_mode, boarder_mat=borders)
            out[f"rep{int(rep_mode)}_k{k}"] = {
                "n_first": int(len(first[0])), "sum_cnt": int(np.sum(first[1], dtype=np.int64)),
                "sha_kh": sha(first[0]), "sha_cnt": sha(first[1]),
                "found": [[int(h), float(v[0]), float(v[1]), float(v[2])] for h, v in found.items()],
                "sha_masked_seq": sha(work)}
            print(f"  rep={rep_mode} k={k}: {len(first[0])} merged k-mers, {len(found)} consensus, {time.time() - t0:.1f} s", flush=True)
    return out


def count_digests(spec, n_reads):
    seq, borders = synth.generate_numpy(spec, 0, n_reads)
    out = {"n_reads": n_reads, "sha_input": sha(seq)}
    for rep_mode in (False, True):
        for k in KS:
            t0 = time.time()
            h = O.comp_kmer_hash(seq, k)
            if not rep_mode:
                h = O.remove_duplicate_hash_per_seq(h, borders, O.get_invalid_hash(O.get_hash_dtype(k)))
            u, c = O.count_uniq_hash(h, k)
            weighted = int(np.sum(u.astype(np.uint64) * c.astype(np.uint64), dtype=np.uint64))      # mod 2^64, as bench.py's checks
            rec = {"n_fwd": int(len(u)), "sum_fwd": int(np.sum(c, dtype=np.int64)), "sha_fwd_kh": sha(u), "sha_fwd_cnt": sha(c),
                   "weighted_checksum": weighted}
            mk, mc = O.merge_revcom(u.copy(), c.copy(), k)
            rec.update({"n_merged": int(len(mk)), "sha_merged_kh": sha(mk), "sha_merged_cnt": sha(mc)})
            out[f"rep{int(rep_mode)}_k{k}"] = rec
            print(f"  rep={rep_mode} k={k}: {len(u)} forward / {len(mk)} merged k-mers, {time.time() - t0:.1f} s", flush=True)
    return out


def main():
    which = sys.argv[1:] or ["cfg2", "cfg2n", "cfg3"]
    res = json.loads(OUT.read_text()) if OUT.exists() else {}
    for name in which:
        print(name, flush=True)
        if name == "cfg2":
            res["cfg2"] = find_motif_digests(synth.CFG2, 1_000_000)
        elif name == "cfg2n":
            res["cfg2n"] = find_motif_digests(synth.CFG2_N, 1_000_000)
        elif name == "cfg3":
            res["cfg3_prefix"] = count_digests(synth.CFG3, 1_000_000)
        OUT.write_text(json.dumps(res, indent=1, sort_keys=True) + "\n")
    print("wrote", OUT)


if __name__ == "__main__":
    main()
