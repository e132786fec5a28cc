#!/usr/bin/env python
"""Round-2 additions to the golden fixtures, produced like make_golden.py by running the UNMODIFIED reference
(`/root/reference/src/kmap`) under ref_shim.py.  Run in the build container:

    python tests/golden/make_golden_r2.py         # writes r2_vectors.pkl.gz

Contents: merge_revcom with keep_lower_hash_flag=False (kmer_count.py:668-683) for uint32 and uint64 hashes.
"""
from __future__ import annotations

import gzip
import pickle
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
import ref_shim  # noqa: E402

kc, md, tc = ref_shim.install()


def main():
    rng = np.random.default_rng(20241017)
    out = {}
    mr = []
    for k, n in [(3, 40), (4, 1000), (4, 150), (5, 700), (6, 3000), (8, 20000), (7, 5), (16, 2000), (17, 1500)]:
        hd = kc.get_hash_dtype(k)
        cd = kc.get_cnt_dtype(k)
        if k >= 16:   # few distinct values so that reverse-complement pairs are present
            base = rng.integers(0, 4 ** k, n // 3, dtype=np.uint64)
            raw = np.concatenate([base, kc.get_revcom_hash_arr(base[: n // 4].astype(hd), k).astype(np.uint64),
                                  rng.integers(0, 4 ** k, n // 3, dtype=np.uint64)]).astype(hd)
        else:
            raw = rng.integers(0, 4 ** k, n).astype(hd)
        u, c = np.unique(raw, return_counts=True)
        c = c.astype(cd)
        for flag in (False, True):
            c_in = c.copy()
            c_work = c.copy()
            mk, mc = kc.merge_revcom(u.copy(), c_work, k, keep_lower_hash_flag=flag)
            mr.append(dict(k=k, keep_lower=flag, kh=u, cnt=c_in, out_kh=mk, out_cnt=mc, mutated_cnt=c_work.copy()))
    out["merge_revcom_flag"] = mr
    with gzip.open(HERE / "r2_vectors.pkl.gz", "wb") as fh:
        pickle.dump(out, fh, protocol=4)
    print("wrote r2_vectors.pkl.gz:", {k: len(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
