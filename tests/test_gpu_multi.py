"""k-rank vs 1-rank parity on real GPUs (needs >= 2 devices; skipped otherwise): scripts/check_multi.py under torchrun."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("args", [["30000"], ["0", "48"]])
def test_two_ranks_bit_identical_to_one(args):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "scripts", "check_multi.py")] + args
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1]
    cmp_ = json.loads(line)["compare"]
    assert all(v["bit_identical"] for v in cmp_.values()), cmp_
