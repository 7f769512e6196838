import os, sys, subprocess, json
code = '''
import sys, torch, numpy as np
sys.path.insert(0, ".")
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
    print("iters", it, "flushed median ms", float(np.median(ts)), "back-to-back ms", e0.elapsed_time(e1) / 20)
'''
for gl in ("8", "16", "32"):
    env = dict(os.environ, ADMMNET_CLASSIC_GL=gl)
    print("GL", gl, flush=True)
    subprocess.run([sys.executable, "-c", code], env=env)
