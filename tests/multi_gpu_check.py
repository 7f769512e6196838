"""torchrun --nproc-per-node 2 tests/multi_gpu_check.py — exact whole-batch semantics across ranks
(norm_scope='global', one scalar all-reduce per active layer over NCCL) against the golden vectors of the
reference and against the single-GPU result; 'shard' scope against the per-shard oracle."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import admmnet_b200  # noqa: E402
from admmnet_b200.sharding import shard_range, sharded_forward  # noqa: E402
from oracle import net_oracle, signals  # noqa: E402
from tests.helpers import load_net_case, rel_err  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True
    for tag in ("init_k10", "pert_k5"):
        z, sd = load_net_case(tag)
        K = int(z["K"])
        net = admmnet_b200.PhiEstADMMNet(10, 10, 3, K).eval()
        net.load_state_dict(sd)
        y, b, s = (torch.from_numpy(z[k]) for k in ("y", "b", "sigma"))
        lo, hi = shard_range(y.shape[0], rank, world)
        phi = sharded_forward(net, y[lo:hi], b[lo:hi], s[lo:hi], "global").cpu().numpy()
        e = rel_err(phi, z["phi_batch"][lo:hi]).max()
        print(f"rank {rank} {tag} global-scope vs reference whole-batch golden: {e:.2e}")
        ok &= e < 1e-4
        phis = sharded_forward(net, y[lo:hi], b[lo:hi], s[lo:hi], "shard").cpu().numpy()
        ref = net_oracle.forward(sd, y[lo:hi], b[lo:hi], s[lo:hi], 10, 10, K).numpy()
        e2 = rel_err(phis, ref).max()
        print(f"rank {rank} {tag} shard-scope vs oracle on the shard: {e2:.2e}")
        ok &= e2 < 1e-4
    # bigger batch: global scope over 2 ranks == single-GPU whole batch
    torch.manual_seed(0)
    net = admmnet_b200.PhiEstADMMNet(10, 10, 3, 10).eval()
    y, b, s, _ = signals.generate(3000, seed=7)
    y, b, s = (torch.from_numpy(a) for a in (y, b, s))
    lo, hi = shard_range(3000, rank, world)
    net.chunk = 512
    phi = sharded_forward(net, y[lo:hi], b[lo:hi], s[lo:hi], "global")
    with torch.no_grad():
        full = net(y.cuda(), b.cuda(), s.cuda())[lo:hi]
    e3 = rel_err(phi.cpu().numpy(), full.cpu().numpy()).max()
    print(f"rank {rank} B=3000 global-scope (2 ranks) vs single-GPU whole batch: {e3:.2e}")
    ok &= e3 < 2e-5
    t = torch.tensor([1.0 if ok else 0.0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MULTI_GPU_CHECK", "PASS" if t.item() == 1.0 else "FAIL")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
