"""Dev script (torchrun --nproc-per-node N): data-parallel training steps over NCCL — every rank trains on its share
of each global batch, gradients are averaged with one flat all-reduce, parameters must stay identical on all ranks."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from admmnet_b200.admm_net import PhiEstADMMNet
from admmnet_b200.autograd import PhiAlignmentLoss
from admmnet_b200.generate import generate_signals
from admmnet_b200.sharding import shard_range
from admmnet_b200.training import make_optimizer, train_step

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
torch.manual_seed(0)
model = PhiEstADMMNet(10, 10, 3, 10).cuda()
opt, sch = make_optimizer(model)
crit = PhiAlignmentLoss()
GB = 256 * world                                    # weak scaling: trainPhi.py's batch of 256 per GPU
y, b, s = generate_signals(GB, 10, 10, 3, seed=5)   # same seed on every rank -> same global batch
pt = y / (b + 1e-8)
lo, hi = shard_range(GB, rank, world)
for it in range(3):
    train_step(model, crit, opt, y[lo:hi], b[lo:hi], s[lo:hi], pt[lo:hi])
torch.cuda.synchronize()
dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
n = 5
for it in range(n):
    loss, _ = train_step(model, crit, opt, y[lo:hi], b[lo:hi], s[lo:hi], pt[lo:hi])
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / n], device="cuda")
dist.all_reduce(ms, op=dist.ReduceOp.MAX)
flat = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
ref = flat.clone()
dist.broadcast(ref, 0)
same = torch.tensor([float(torch.equal(flat, ref))], device="cuda")
dist.all_reduce(same, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"world {world}: train step {ms.item():.1f} ms, {GB / ms.item() * 1e3:.0f} signals/s, loss {float(loss):.4f}, "
          f"parameters identical on all ranks: {bool(same.item())}")
dist.destroy_process_group()
