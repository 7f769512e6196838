import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.gpu_debug import eigh_gpu
torch.manual_seed(0)
for d in (3, 4, 8, 9, 17, 101):
    X = torch.randn(2, d, d, dtype=torch.complex64)
    A = 0.5 * (X + X.transpose(1, 2).conj())
    ev, U, _, st = eigh_gpu(A)
    res = (A @ U - U * ev.unsqueeze(1).to(torch.complex64)).abs().amax(dim=(1, 2))
    orth = (U.transpose(1, 2).conj() @ U - torch.eye(d)).abs().amax(dim=(1, 2))
    print(d, 'resid', res.numpy(), 'orth', orth.numpy())
    if d <= 4:
        w, Ut = torch.linalg.eigh(A[0].to(torch.complex128))
        print(' ev', ev[0].numpy(), 'true', w.numpy())
        print(' |U| gpu\n', U[0].abs().numpy(), '\n |U| true (cols sorted by ev)\n', Ut.abs().numpy())
