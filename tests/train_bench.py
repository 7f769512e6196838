"""Dev script: training step throughput (K=10, batch 256 = trainPhi.py's batch size) on one GPU."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from admmnet_b200.admm_net import PhiEstADMMNet
from admmnet_b200.autograd import PhiAlignmentLoss
from admmnet_b200.training import train_step
from admmnet_b200.generate import generate_signals

for B in (256, 1024, 4096):
    torch.manual_seed(0)
    model = PhiEstADMMNet(10, 10, 3, 10).cuda().train()
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3, weight_decay=1e-4)
    crit = PhiAlignmentLoss()
    y, b, s = generate_signals(B, 10, 10, 3, seed=1)
    pt = y / (b + 1e-8)
    for _ in range(3):
        train_step(model, crit, opt, y, b, s, pt)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = 5
    for _ in range(n):
        loss, _ = train_step(model, crit, opt, y, b, s, pt)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"B={B} train step {ms:.1f} ms  {B / ms * 1e3:.0f} signals/s  loss {float(loss):.4f}  peak mem {torch.cuda.max_memory_allocated() / 2**30:.2f} GiB")
