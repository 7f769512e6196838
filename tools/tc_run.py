"""Runs the f(A) tap (tridiagonalisation, tridiagonal solver, tail kernel) a few times on random Hermitian matrices:
the small, self-contained command line that ncu wraps to profile k_tail_tc (see profiles/).
    python tools/tc_run.py [B] [d] [reps]"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import admmnet_b200  # noqa: E402
from admmnet_b200 import _capi  # noqa: E402
from admmnet_b200.params import pack_state_dict  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 8
d = int(sys.argv[2]) if len(sys.argv) > 2 else 101
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
L = _capi.lib()
dev = torch.device("cuda")
torch.manual_seed(0)
net = admmnet_b200.PhiEstADMMNet(10, 10, 3, 10)
P = pack_state_dict(net.state_dict(), 100, 10).to(dev)
g = torch.Generator().manual_seed(1)
X = torch.randn(B, d, d, dtype=torch.complex64, generator=g) * (3.0 / d ** 0.5)
A = (0.5 * (X + X.transpose(1, 2).conj())).to(dev).contiguous()
nb = C.c_size_t()
_capi.check(L.admmnet_eigh_workspace_bytes(B, d, 0, C.byref(nb)))
ws = torch.empty(nb.value, dtype=torch.uint8, device=dev)
G = torch.empty(B, d * (d + 1) // 2, dtype=torch.complex64, device=dev)
st = torch.zeros(1, dtype=torch.int32, device=dev)
for _ in range(reps):
    _capi.check(L.admmnet_eigh_batched(A.data_ptr(), B, d, None, None, G.data_ptr(), P[3].data_ptr(), ws.data_ptr(),
                                       nb.value, 0, torch.cuda.current_stream().cuda_stream, st.data_ptr()))
torch.cuda.synchronize()
assert int(st.item()) == 0
print("ok", B, d, float(G.abs().mean()))
