"""Run the worker of tests/test_distributed_gpu.py::test_two_rank_update_matches_oracle with N ranks (diagnosis: N=1 separates
what is distributed from what is not).   python scripts/run_rank_worker.py N precision fused"""
import os
import subprocess
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.test_distributed_gpu import WORKER

n, precision, fused = sys.argv[1], sys.argv[2], sys.argv[3]
with tempfile.TemporaryDirectory() as d:
    path = os.path.join(d, "worker.py")
    open(path, "w").write(WORKER)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
                        "--master-port", "29517", path, precision, fused], capture_output=True, text=True, timeout=300)
    print("\n".join(l for l in r.stdout.splitlines() if l.startswith("{")) or (r.stdout[-2000:] + r.stderr[-2000:]))
