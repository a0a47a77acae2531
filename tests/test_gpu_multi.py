"""Real multi-process NCCL parity: the partitioned EXCHANGE path on >= 2 GPUs against the serial oracle.
Skipped on a single-GPU box (the emulated-rank test in test_gpu_parity.py and the gloo test in
test_partition_host.py cover the same logic there)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _ngpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("kind", ["heat", "heat_big", "elasticity", "elasticity_q1"])
def test_nccl_exchange_matches_serial_oracle(tmp_path, kind):
    n = _ngpus()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 4 if n >= 4 else 2
    out = tmp_path / "res.json"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29600 + (os.getpid() % 200)), os.path.join(HERE, "mp_exchange_worker.py"), str(out), kind]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    res = json.load(open(out))
    assert res["pattern_exact"] and res["dofs_once"]
    assert res["nz_err"] <= 1e-12 and res["f_err"] <= 1e-12, res
