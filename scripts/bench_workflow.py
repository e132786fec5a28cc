"""BASELINE config 2 end to end through the reference-shaped drivers: synthetic HT-SELEX-like reads (1e6 x 40 bp, two planted
motifs) written as FASTA -> `preproc` -> `scan_motif` (k = 8..14, stock settings otherwise), wall-clock per stage and the
cProfile top of scan_motif.  Writes gpurun_out/workflow[_<spec>].json.
Usage: python scripts/bench_workflow.py [n_reads] [cfg2|cfg3]   (cfg3 = ChIP-like 100-bp reads, one planted 14-mer: the
workload of the headline bench; run it at >= 1e7 reads under `ncu --metrics gpu__time_duration.sum` for the launch list that
shows which counting kernels the CLI path runs)."""
import cProfile
import io
import json
import pstats
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    n_reads = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
    import tomli_w
    import torch
    from kmap_b200 import kmer_count as K, motif_discovery as MD, synth
    torch.cuda.set_device(0)
    spec_name = sys.argv[2] if len(sys.argv) > 2 else "cfg2"
    spec = {"cfg2": synth.CFG2, "cfg3": synth.CFG3}[spec_name]
    seq, borders = synth.generate_numpy(spec, 0, n_reads)
    L = spec.read_len
    body = np.frombuffer(b"ACGT", dtype=np.uint8)[np.minimum(seq.reshape(-1, L + 1)[:, :L], 3)]
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
    out = {"spec": spec_name, "reads": n_reads, "read_len": L, "fasta_MB": fa.stat().st_size / 1e6}
    t = time.perf_counter()
    K._preproc(str(fa), str(res))
    torch.cuda.synchronize()
    out["preproc_s"] = time.perf_counter() - t
    np.random.seed(1)
    pr = cProfile.Profile()
    t = time.perf_counter()
    pr.enable()
    MD._scan_motif(str(res))
    pr.disable()
    torch.cuda.synchronize()
    out["scan_motif_s"] = time.perf_counter() - t
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28)
    # CPU leg: the oracle port of find_motif (per-read Python de-duplication, np.unique, list scans, Python mask loop) for ONE
    # k on a sample of the same reads; the reference's drivers repeat it for 7 values of k and add per-read occurrence scans
    from oracle import kmap_oracle as O
    n_cpu = min(n_reads, 20000)
    s_np, b_np = synth.generate_numpy(spec, 0, n_cpu)
    mdd = K.init_motif_def_dict(ROOT / "kmap_b200" / "default_motif_def_table.csv")
    m = mdd[10]
    t = time.perf_counter()
    O.find_motif(s_np.copy(), 10, m.max_ham_dist, m.p_uniform, m.ratio_mu, m.ratio_std, m.ratio_cutoff, boarder_mat=b_np)
    out["cpu_oracle_find_motif_k10_s"] = time.perf_counter() - t
    out["cpu_sample_reads"] = n_cpu
    out["final_conseq"] = (res / "final_conseq.txt").read_text().split()
    out["candidate_rows"] = len((res / "candidate_conseq.csv").read_text().splitlines()) - 1
    Path("gpurun_out").mkdir(exist_ok=True)
    tag = "" if spec_name == "cfg2" else "_" + spec_name
    Path(f"gpurun_out/workflow{tag}.json").write_text(json.dumps(out, indent=1))
    Path(f"gpurun_out/workflow{tag}_profile.txt").write_text(s.getvalue())
    print(json.dumps(out))
    print(s.getvalue()[:6000])


if __name__ == "__main__":
    main()
