import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.gpu_debug import eigh_gpu
torch.manual_seed(0)
d = 101
eye = torch.eye(d, dtype=torch.complex64)
diag = torch.diag(torch.linspace(-3, 5, d)).to(torch.complex64)
arrow = torch.diag(torch.full((d,), 0.01)).to(torch.complex64)
v = torch.randn(d - 1, dtype=torch.complex64)
arrow[:-1, -1] = v; arrow[-1, :-1] = v.conj(); arrow[-1, -1] = 1.8
rank1 = torch.outer(v.new_ones(d), v.new_ones(d))
A = torch.stack([eye, diag, arrow, rank1, torch.zeros(d, d, dtype=torch.complex64)])
ev, U, _, st = eigh_gpu(A)
print('status', st)
for i, name in enumerate(['eye', 'diag', 'arrow', 'rank1', 'zeros']):
    res = (A[i] @ U[i] - U[i] * ev[i][None].to(torch.complex64)).abs().amax().item()
    G = U[i].conj().T @ U[i]
    orth = (G - torch.eye(d)).abs().amax().item()
    cn = U[i].abs().pow(2).sum(0).sqrt()
    w = torch.linalg.eigvalsh(A[i].to(torch.complex128)).float()
    print(name, 'resid %.2e orth %.2e' % (res, orth), 'colnorm min/max', cn.min().item(), cn.max().item(), 'nan', torch.isnan(U[i].real).sum().item(),
          'lamerr', (ev[i].sort()[0] - w).abs().max().item())
    if orth > 1e-3:
        bad = (G - torch.eye(d)).abs().amax(0)
        print('  bad cols', torch.nonzero(bad > 1e-3).flatten()[:20].tolist(), 'ev there', ev[i][torch.nonzero(bad > 1e-3).flatten()[:8]].tolist())
