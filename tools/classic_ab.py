"""A/B of the classical-ADMM kernels at BASELINE.json configs[1] (batch 65536, n = 100, complex64 in, complex128 out):
the persistent bulk-copy form (default) against the plain kernels with 32 / 16 / 8 lanes per signal; L2 flushed before
every timed launch.  Also checks that the forms agree.     python tools/classic_ab.py"""
import os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = '''
import sys, torch, numpy as np
sys.path.insert(0, %r)
import admmnet_b200 as pkg
from bench import tile_signals
dev = torch.device("cuda")
y, b, s = tile_signals(65536, seed=77)
yd, bd = torch.from_numpy(y).to(dev), torch.from_numpy(b).to(dev)
o = torch.empty(65536, 100, dtype=torch.complex128, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for it in (5, 100):
    pkg.admm_for_us_batched(yd, bd, 1.0, it, out=o)
    ts = []
    for _ in range(10):
        flush.zero_()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); pkg.admm_for_us_batched(yd, bd, 1.0, it, out=o); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): pkg.admm_for_us_batched(yd, bd, 1.0, it, out=o)
    e1.record(); torch.cuda.synchronize()
    ms = float(np.median(ts))
    print("iters", it, "flushed median ms %%.4f" %% ms, "-> %%.0f GB/s" %% (65536 * 3204 / ms / 1e6), "back-to-back ms %%.4f" %% (e0.elapsed_time(e1) / 20),
          "checksum %%.15e" %% float(o.abs().sum()))
# ragged sizes against the 32-lane kernel's formula in torch fp64
for B in (1, 31, 33, 1000):
    yy, bb = yd[:B].to(torch.complex128), bd[:B].to(torch.complex128)
    D = (bb.abs() ** 2); dyb = yy * bb.conj(); rho = 0.7
    phi = torch.zeros_like(yy)
    for _ in range(5):
        Dv = dyb + rho * D * phi
        phi = Dv - D * (rho * Dv.sum(1, keepdim=True) / (1 + rho * D.sum(1, keepdim=True)))
    got = pkg.admm_for_us_batched(yd[:B], bd[:B], rho, 5)
    print("B", B, "max rel err vs torch fp64", float(((got - phi).abs().amax(1) / phi.abs().amax(1)).max()))
''' % ROOT
for env in ({"ADMMNET_CLASSIC_PV": "0"}, {"ADMMNET_CLASSIC_PV": "1"}, {"ADMMNET_CLASSIC_PV": "2"}, {"ADMMNET_CLASSIC_PV": "3"},
            {"ADMMNET_CLASSIC_P": "0"}, {"ADMMNET_CLASSIC_GL": "16"}, {"ADMMNET_CLASSIC_GL": "8"}):
    print(env, flush=True)
    subprocess.run([sys.executable, "-c", code], env=dict(os.environ, **env))
