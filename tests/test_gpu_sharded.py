"""Multi-GPU parity (needs >= 2 B200 on the box; skipped otherwise): one process per GPU, fact
table row-range sharded, partial aggregates merged inside the library with NCCL."""
import os
import socket
import subprocess
import sys

import pytest

from common import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_plans_match_reference(world):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", str(port),
                        os.path.join(ROOT, "tests", "sharded_worker.py")],
                       capture_output=True, text=True, timeout=900)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"sharded_worker_n{world}.log"), "w") as f:
        f.write(r.stdout + "\n--- stderr ---\n" + r.stderr[-4000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "sharded q3" in r.stdout
    assert "partitioned micro_join_avg" in r.stdout
    assert "sharded error agreement" in r.stdout
    assert "broadcast q3" in r.stdout and "shared-build q3" in r.stdout
