"""Tuning aid for k_tail_tc: (1) device time of the tail kernel (tensor-core form against the SIMT form) on a batch of
random Hermitian matrices through the f(A) tap, (2) with ADMMNET_TC_PROF=1, clock cycles per phase of CTA 0.
    python tools/tc_profile.py [B] [d]"""
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
buf = (C.c_longlong * 16)()
L.admmnet_tail_tc_profile_read(buf)          # clear
nk = L.admmnet_profile_kinds()
L.admmnet_profile_begin()
for _ in range(3): run()
ms = (C.c_double * nk)(); ln = (C.c_longlong * nk)()
_capi.check(L.admmnet_profile_end(ms, ln))
names = [L.admmnet_profile_kind_name(i).decode() for i in range(nk)]
print({names[i]: round(ms[i] / 3, 3) for i in range(nk) if ln[i]})
L.admmnet_tail_tc_profile_read(buf)
v = list(buf)
if v[12]:
    ph = ["wait", "s1", "gram", "ysolve", "vtile", "gemm1", "psplit_ytile", "gemm2", "resplit", "wstage", "rebuild", "epilogue"]
    tot = sum(v[:12])
    print("signals", v[12], "cycles/signal", tot // v[12])
    for nme, c in zip(ph, v[:12]): print("  %%-14s %%8d  %%5.1f %%%%" %% (nme, c // v[12], 100.0 * c / tot))
'''

if __name__ == "__main__":
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 16
    d = int(sys.argv[2]) if len(sys.argv) > 2 else 101
    envs = ({"ADMMNET_TAILTC": "0"}, {"ADMMNET_TAILTC": "1", "ADMMNET_TC_PROF": "1"})
    if len(sys.argv) > 3:                       # e.g.  ... 2368 101 ADMMNET_TRD=0 ADMMNET_TRD=1
        envs = [dict(kv.split("=") for kv in a.split(",")) for a in sys.argv[3:]]
    for env in envs:
        print(env, flush=True)
        subprocess.run([sys.executable, "-c", CODE % (ROOT, B, d)], env=dict(os.environ, **env), check=False)
