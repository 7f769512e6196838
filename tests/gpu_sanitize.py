"""Small end-to-end run for compute-sanitizer (memcheck): every kernel once, tiny batch."""
import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import admmnet_b200
from oracle import signals
from tests.gpu_debug import eigh_gpu
torch.manual_seed(0)
for d in (5, 40, 101, 122):
    X = torch.randn(2, d, d, dtype=torch.complex64)
    A = 0.5 * (X + X.transpose(1, 2).conj())
    ev, U, _, st = eigh_gpu(A)
    print(d, st, float((A @ U - U * ev.unsqueeze(1).to(torch.complex64)).abs().amax()))
net = admmnet_b200.PhiEstADMMNet(10, 10, 3, 3)
net.chunk = 3
y, b, s, _ = signals.generate(7, seed=1)
phi = net(torch.from_numpy(y), torch.from_numpy(b), torch.from_numpy(s))
print('phi', float(phi.abs().max()))
r = admmnet_b200.alt_peak_search_batched(phi, 10, 10, dict(xstep=0.02, ystep=0.02, iter=2), topl=3)
print('peaks', r['count'].tolist())
p, it = admmnet_b200.admm_for_us(y[0].astype(np.complex128), b[0].astype(np.complex128), 10, 10, 1.0, 1.0)
print('classic', it, float(np.abs(p).max()))
