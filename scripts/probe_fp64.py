"""DFMA throughput probe on the current GPU (roofline denominator for the FP64-bound gather kernels)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "python-sph_b200")]
import torch  # noqa: E402
import engine  # noqa: E402

e = engine.Engine()
out = {}
for blocks in (148 * 4, 148 * 8, 148 * 16):
    out[str(blocks)] = max(e.fp64_peak(blocks, 20000) for _ in range(3))
print(json.dumps({"gpu": torch.cuda.get_device_name(0), "fp64_dfma_tflops": out}))
