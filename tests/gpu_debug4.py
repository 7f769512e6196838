import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch, numpy as np
import admmnet_b200 as pkg
from test_gpu_parity import _arrow_eigh
n = 33
torch.manual_seed(n); B = 8
h = torch.randn(B, n) * 0.1
phi = torch.randn(B, n, dtype=torch.complex64)
c0 = torch.full((B,), 1.8)
ev, U, ok = _arrow_eigh(h, phi, c0)
print(ok.tolist())
for i in range(B):
    if ok[i] != 1:
        print(i, ev[i].tolist())
