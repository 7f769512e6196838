import ctypes as C, sys, os, torch, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import admmnet_b200
from admmnet_b200 import _capi
L = _capi.lib()
dev = torch.device("cuda")
B, d = 256, 101
g = torch.Generator().manual_seed(1)
X = torch.randn(B, d, d, dtype=torch.complex64, generator=g) * (3.0 / d ** 0.5)
A = (0.5 * (X + X.transpose(1, 2).conj())).to(dev).contiguous()
nb = C.c_size_t()
_capi.check(L.admmnet_eigh_workspace_bytes(B, d, 0, C.byref(nb)))
def run():
    ws = torch.empty(nb.value, dtype=torch.uint8, device=dev)
    ev = torch.empty(B, d, dtype=torch.float32, device=dev)
    U = torch.empty(B, d, d, dtype=torch.complex64, device=dev)
    st = torch.zeros(1, dtype=torch.int32, device=dev)
    _capi.check(L.admmnet_eigh_batched(A.data_ptr(), B, d, ev.data_ptr(), U.data_ptr(), None, None, ws.data_ptr(),
                                       nb.value, 0, torch.cuda.current_stream().cuda_stream, st.data_ptr()))
    torch.cuda.synchronize()
    return ev.cpu(), U.cpu(), int(st.item())
r = [run() for _ in range(4)]
for i in range(1, 4):
    print("eigh run", i, "ev equal", torch.equal(r[0][0], r[i][0]), "U equal", torch.equal(r[0][1], r[i][1]),
          "max dU", float((r[0][1] - r[i][1]).abs().max()), "status", r[i][2])
# training step determinism
sys.path.insert(0, os.path.join(ROOT))
from tests.test_training import _load, _model
from admmnet_b200.autograd import PhiAlignmentLoss
from admmnet_b200.training import make_optimizer, train_step
z, sd, _ = _load()
K = int(z["K"])
y, b, s, pt = (torch.from_numpy(z[k]).cuda() for k in ("y", "b", "sigma", "phi_true"))
crit = PhiAlignmentLoss()
ms = [_model(sd, K, "cuda") for _ in range(2)]
grads = []
for m in ms:
    m.zero_grad()
    phi = m(y, b, s)
    loss = crit(phi, pt)[0]
    loss.backward()
    grads.append({n: p.grad.clone() for n, p in m.named_parameters() if p.grad is not None})
    print("loss", float(loss), "phi sum", float(phi.abs().sum()))
bad = [(n, float((grads[0][n] - grads[1][n]).abs().max()), float(grads[0][n].abs().max())) for n in grads[0] if not torch.equal(grads[0][n], grads[1][n])]
print("params with non-identical grads:", len(bad), "of", len(grads[0]))
for t in bad[:10]: print(t)
