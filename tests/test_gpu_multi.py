"""k-rank vs 1-rank parity on real GPUs (skipped with fewer than 2 devices): scripts/check_multi.py under torchrun with
as many ranks as the box has GPUs (2, 4 or 8).  Both step drivers -- kept neighbour lists and scanning kernels -- must
return the SAME BITS as their single-GPU form on the same grid geometry."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ranks():
    import torch
    n = torch.cuda.device_count()
    return 8 if n >= 8 else 4 if n >= 4 else 2 if n >= 2 else 0


@pytest.mark.parametrize("path,steps", [("lists", 6), ("scan", 2)])
@pytest.mark.parametrize("args", [["30000"], ["0", "48"]])
def test_ranks_bit_identical_to_one(args, path, steps):
    k = _ranks()
    if k == 0:
        pytest.skip("needs at least 2 GPUs")
    env = dict(os.environ, CHECK_PATH=path, CHECK_STEPS=str(steps))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(k), "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "scripts", "check_multi.py")] + args
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1]
    cmp_ = json.loads(line)["compare"]
    assert all(v["bit_identical"] for v in cmp_.values()), cmp_


@pytest.mark.parametrize("args", [["30000"], ["40000", "plummer"]])
def test_range_decomposition_bit_identical_to_one(args):
    """ranges of sorted slots on one global cell list (ranges.RangeRunner; nested grids for the Plummer sphere) == the
    single-GPU FastStepper, bit for bit"""
    k = _ranks()
    if k == 0:
        pytest.skip("needs at least 2 GPUs")
    env = dict(os.environ, CHECK_PATH="lists", CHECK_STEPS="5", CHECK_DECOMP="range")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(k), "--master-addr", "127.0.0.1",
           "--master-port", "29534", os.path.join(ROOT, "scripts", "check_multi.py")] + args
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1]
    cmp_ = json.loads(line)["compare"]
    assert all(v["bit_identical"] for v in cmp_.values()), cmp_


@pytest.mark.skipif(os.environ.get("SPHB_TEST_MULTI_GRAVITY") is None,
                    reason="SlabRunner._stage_gravity was written after the round's GPU budget was spent and has not run on "
                           "hardware yet: set SPHB_TEST_MULTI_GRAVITY=1 to run it")
def test_ranks_with_the_reference_default_gravity():
    """the reference's default G = 39.4227 (pyfluid.py:35) on k ranks: xi and its correction through the slab kernels, the
    pair sum over all-gathered records.  The pair sum runs in another order than on one GPU: 1e-12, not bit for bit"""
    k = _ranks()
    if k == 0:
        pytest.skip("needs at least 2 GPUs")
    env = dict(os.environ, CHECK_PATH="scan", CHECK_STEPS="2", CHECK_G="39.4227")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(k), "--master-addr", "127.0.0.1",
           "--master-port", "29535", os.path.join(ROOT, "scripts", "check_multi.py"), "12000"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1]
    cmp_ = json.loads(line)["compare"]
    assert all(v["max_abs_diff"] <= 1e-12 * max(v["scale"], 1e-300) for v in cmp_.values()), cmp_
