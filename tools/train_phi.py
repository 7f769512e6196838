"""Command-line counterpart of the reference's trainPhi.py (trainPhi.py:13-261) on the B200 path.

    python tools/train_phi.py --data ./ofdm_dataset --generate 100000 --epochs 50
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/train_phi.py --data ./ofdm_dataset

Datasets and checkpoints use the reference's formats (admmnet_b200.dataset), so either side can read the other's."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from admmnet_b200.admm_net import PhiEstADMMNet
from admmnet_b200.dataset import generate_dataset, load_split
from admmnet_b200.training import fit


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--data", default="./ofdm_dataset")
    ap.add_argument("--generate", type=int, default=0, help="write a fresh dataset of this many samples first")
    ap.add_argument("--checkpoint-dir", default="./checkpoints")
    ap.add_argument("--num-layers", type=int, default=10)
    ap.add_argument("--batch-size", type=int, default=256)
    ap.add_argument("--epochs", type=int, default=1000)
    ap.add_argument("--lr", type=float, default=5e-3)
    ap.add_argument("--weight-decay", type=float, default=1e-3)
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args()

    distributed = "RANK" in os.environ and int(os.environ.get("WORLD_SIZE", "1")) > 1
    rank = int(os.environ.get("RANK", "0"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    if distributed:
        dist.init_process_group("nccl")
    if args.generate and rank == 0:
        generate_dataset(args.data, total_samples=args.generate, seed=args.seed)
    if distributed:
        dist.barrier()
    train, val = load_split(args.data, "train"), load_split(args.data, "val")
    os.makedirs(args.checkpoint_dir, exist_ok=True)
    config = {"data_dir": args.data, "batch_size": args.batch_size, "num_layers": args.num_layers, "M": 10, "N": 10,
              "L_max": 3, "epochs": args.epochs, "lr": args.lr, "weight_decay": args.weight_decay,
              "checkpoint_dir": args.checkpoint_dir, "seed": args.seed}
    torch.manual_seed(args.seed)                 # identical initial weights on every rank
    model = PhiEstADMMNet(num_layers=args.num_layers, M=10, N=10, L=3).cuda()
    history = fit(model, train, val, config, log=print if rank == 0 else (lambda *_: None))
    if rank == 0:
        with open(os.path.join(args.checkpoint_dir, "training_history.json"), "w") as fh:
            json.dump({k: [float(x) for x in v] for k, v in history.items()}, fh, indent=2)
    if distributed:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
