"""A few launches of the classical-ADMM kernel at BASELINE.json configs[1] (batch 65536, n = 100, complex64 in,
complex128 out, 5 iterations): the small command line ncu wraps to profile k_classic_p (see profiles/).
    python tools/classic_run.py [n_iter]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import admmnet_b200 as pkg  # noqa: E402
from bench import tile_signals  # noqa: E402

it = int(sys.argv[1]) if len(sys.argv) > 1 else 5
dev = torch.device("cuda")
y, b, s = tile_signals(65536, seed=77)
yd, bd = torch.from_numpy(y).to(dev), torch.from_numpy(b).to(dev)
o = torch.empty(65536, 100, dtype=torch.complex128, device=dev)
for _ in range(3):
    pkg.admm_for_us_batched(yd, bd, 1.0, it, out=o)
torch.cuda.synchronize()
print("ok", float(o.abs().sum()))
