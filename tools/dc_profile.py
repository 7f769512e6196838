"""Tuning aid for k_dc (csrc/dc_kernels.cu): device time of the tridiagonal eigen-solver (fused divide & conquer
against the k_ql + k_rotf pair) on a batch of random Hermitian matrices through the f(A) tap, and with
ADMMNET_DC_PROF=1 the clock cycles of CTA 0 per (level, phase).
    python tools/dc_profile.py [B] [d]    |   ... sweep   (batch-size sweep)   |   ... net [B]   (the net's own matrices)"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CODE = r'''
import ctypes as C, sys, torch
sys.path.insert(0, %r)
from admmnet_b200 import _capi
from admmnet_b200.params import pack_state_dict
import admmnet_b200
L = _capi.lib()
B, d = %d, %d
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
def run():
    _capi.check(L.admmnet_eigh_batched(A.data_ptr(), B, d, None, None, G.data_ptr(), P[3].data_ptr(), ws.data_ptr(),
                                       nb.value, 0, torch.cuda.current_stream().cuda_stream, st.data_ptr()))
for _ in range(2): run()
torch.cuda.synchronize()
buf = (C.c_longlong * 128)()
L.admmnet_dc_profile_read(buf)          # clear
nk = L.admmnet_profile_kinds()
L.admmnet_profile_begin()
for _ in range(3): run()
ms = (C.c_double * nk)(); ln = (C.c_longlong * nk)()
_capi.check(L.admmnet_profile_end(ms, ln))
names = [L.admmnet_profile_kind_name(i).decode() for i in range(nk)]
print({names[i]: round(ms[i] / 3, 3) for i in range(nk) if ln[i]}, "status", int(st.item()))
L.admmnet_dc_profile_read(buf)
v = list(buf)
if v[96]:
    ph = ["tables+z", "sort/defl", "close", "secular", "gu-eis", "norms", "W", "gemm+copy"]
    tot = sum(v[8:64])
    print("signals", v[96], "cycles/signal", tot // v[96])
    print("level " + " ".join("%%10s" %% p for p in ph) + "      total   sec.its/warp")
    for lev in range(1, 8):
        row = v[8 * lev: 8 * lev + 8]
        if sum(row):
            print("%%5d " %% lev + " ".join("%%10d" %% (c // v[96]) for c in row) + " %%10d   %%.2f" %% (sum(row) // v[96], v[72 + lev] / max(1, v[80 + lev])))
    print("phase " + " ".join("%%10d" %% (sum(v[8 * lev + p] for lev in range(1, 8)) // v[96]) for p in range(8)))
'''

CODE_NET = r'''
import ctypes as C, sys, torch
sys.path.insert(0, %r)
from admmnet_b200 import _capi
import admmnet_b200
L = _capi.lib()
B = %d
dev = torch.device("cuda")
torch.manual_seed(0)
net = admmnet_b200.PhiEstADMMNet(10, 10, 3, 10).eval()
y, b, s = admmnet_b200.generate_signals(B, 10, 10, 3, snr_w=20.0, snr_demod=7.0, seed=1234, device=dev)
with torch.no_grad():
    net.forward_device(y, b, s)
torch.cuda.synchronize()
buf = (C.c_longlong * 128)()
L.admmnet_dc_profile_read(buf)
nk = L.admmnet_profile_kinds()
L.admmnet_profile_begin()
with torch.no_grad():
    net.forward_device(y, b, s)
ms = (C.c_double * nk)(); ln = (C.c_longlong * nk)()
_capi.check(L.admmnet_profile_end(ms, ln))
names = [L.admmnet_profile_kind_name(i).decode() for i in range(nk)]
print({names[i]: round(ms[i], 3) for i in range(nk) if ln[i]})
L.admmnet_dc_profile_read(buf)
v = list(buf)
if v[96]:
    ph = ["tables+z", "sort/defl", "close", "secular", "gu-eis", "norms", "W", "gemm+copy"]
    tot = sum(v[8:64])
    print("signals", v[96], "cycles/signal", tot // v[96])
    print("level " + " ".join("%%10s" %% p for p in ph) + "      total   sec.its/warp")
    for lev in range(1, 8):
        row = v[8 * lev: 8 * lev + 8]
        if sum(row):
            print("%%5d " %% lev + " ".join("%%10d" %% (c // v[96]) for c in row) + " %%10d   %%.2f" %% (sum(row) // v[96], v[72 + lev] / max(1, v[80 + lev])))
'''

if __name__ == "__main__":
    d = int(sys.argv[2]) if len(sys.argv) > 2 else 101
    if len(sys.argv) > 1 and sys.argv[1] == "net":        # the matrices of the unrolled net itself (8 general layers of one chunk)
        B = int(sys.argv[2]) if len(sys.argv) > 2 else 148 * 16
        for env in ({"ADMMNET_DCK": "0"}, {"ADMMNET_DCK": "1", "ADMMNET_DC_PROF": "1"}):
            print(env, flush=True)
            subprocess.run([sys.executable, "-c", CODE_NET % (ROOT, B)], env=dict(os.environ, **env), check=False)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "sweep":      # batch-size sweep: where does the fused kernel win?
        for B in (1, 64, 512, 2368, 4736, 9472, 16384):
            for env in ({"ADMMNET_DCK": "0"}, {"ADMMNET_DCK": "1"}):
                print("B =", B, env, flush=True)
                subprocess.run([sys.executable, "-c", CODE % (ROOT, B, d)], env=dict(os.environ, **env), check=False)
        sys.exit(0)
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 16
    envs = ({"ADMMNET_DCK": "0"}, {"ADMMNET_DCK": "1", "ADMMNET_DC_PROF": "1"})
    for env in envs:
        print(env, flush=True)
        subprocess.run([sys.executable, "-c", CODE % (ROOT, B, d)], env=dict(os.environ, **env), check=False)
