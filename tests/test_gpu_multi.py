"""N-GPU == 1-GPU == oracle, on hardware: two NCCL ranks (one process per GPU, launched like the driver launches bench.py)
run the sharded product path and rank 0 compares the merged results.  Skipped on a box with one GPU
(`gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`)."""
import os
import socket
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _torchrun(n, args, timeout=900):
    import torch
    if not torch.cuda.is_available() or torch.cuda.device_count() < n:
        pytest.skip(f"needs {n} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), str(ROOT / "tests" / "nccl_worker.py")] + [str(a) for a in args]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT, env=dict(os.environ, PYTHONPATH=str(ROOT)))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-6000:]


def test_sharded_count_kmers_two_ranks_vs_oracle_and_one_gpu(tmp_path):
    """api.count_kmers(shard, table_allreduce=...) on 2 ranks: merged lists == oracle lists of the whole input == 1-GPU lists
    (k = 8..14 dense, k = 16 sorted lists merged over the ranks), both repetitive modes; 3 001 reads: the shard has far
    fewer windows than 4^14 cells while the merged table holds the k-mers of both shards (the capacity case)"""
    _torchrun(2, ["count", tmp_path, 3001])
    assert (tmp_path / "count.ok").exists()


@pytest.mark.parametrize("golden", ["testfa", "testfa_stock_k"])
def test_scan_motif_two_ranks_equals_reference_goldens(tmp_path, golden):
    """`torchrun --nproc-per-node 2 -m kmap_b200 scan_motif`: reads sharded, tables all-reduced, occurrence rows gathered;
    every output file equals the unmodified reference's (k = 8..14, and the stock range 6..16 with the sorted k = 16 path)"""
    _torchrun(2, ["scan_motif", tmp_path, golden])
    assert (tmp_path / "scan_motif.ok").exists()
