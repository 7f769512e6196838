"""Dataset and checkpoint formats of the reference, so its artefacts load here unchanged and
ours load there.

  * dataset directory  (generate_data.py:223-256, 361-408): <dir>/{train,val,test}/<field>.npy with the 13 fields of
    DatasetGeneratorCreatePhi (y_real/imag, b_real/imag, tau, f, C_real/imag, L_true, sigma, phi_real/imag, ser) plus
    dataset_config.json / dataset_info.npz;
  * checkpoint  (trainPhi.py:238-246): torch.save dict with epoch, model/optimizer/scheduler state_dicts, best_val_loss,
    config, history.

`generate_dataset` writes such a directory from the DEVICE generator (gen_kernels.cu) with labels from the CUDA
classical solver (the reference labels phi with admm_for_us, generate_data.py:454) instead of the per-sample Python loop.
"""
import json
import os

import numpy as np
import torch

from . import _capi
from .admm import admm_for_us_batched, executed_iterations

FIELDS = ("y_real", "y_imag", "b_real", "b_imag", "tau", "f", "C_real", "C_imag", "L_true", "sigma", "phi_real",
          "phi_imag", "ser")
SPLITS = ("train", "val", "test")


def split_sizes(total_samples, train_ratio=0.7, val_ratio=0.15):
    """generate_data.py:54-57."""
    n_train = int(total_samples * train_ratio)
    n_val = int(total_samples * val_ratio)
    return n_train, n_val, total_samples - n_train - n_val


def generate_split(n_samples, Nb=10, Nd=10, L_max=3, snr_range=(5, 25), snr_demod=7.0, seed=0, device=None):
    """One split as the reference's dict of numpy arrays (generate_data.py:361-408), computed on the GPU."""
    _capi.require_cuda()
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    n = Nb * Nd
    y = torch.empty(n_samples, n, dtype=torch.complex64, device=dev)
    b = torch.empty_like(y)
    sigma = torch.empty(n_samples, dtype=torch.float32, device=dev)
    ser = torch.empty(n_samples, dtype=torch.float32, device=dev)
    truth = torch.empty(n_samples, L_max, 4, dtype=torch.float64, device=dev)
    if n_samples:
        _capi.check(_capi.lib().admmnet_generate_dataset(
            y.data_ptr(), b.data_ptr(), sigma.data_ptr(), truth.data_ptr(), ser.data_ptr(), n_samples, Nb, Nd, L_max,
            float(snr_range[0]), float(snr_range[1]), float(snr_demod), seed,
            torch.cuda.current_stream(dev).cuda_stream))
        # labels: admm_for_us(y, b, Nd, Nb, 1, sigma, {'max_iter': 100, ...})  (generate_data.py:448-454)
        phi = admm_for_us_batched(y, b, rho=1.0, n_iter=executed_iterations({"max_iter": 100})).to(torch.complex64)
    else:
        phi = torch.empty_like(y)
    t = truth.to(torch.float32).cpu().numpy()
    yc, bc, pc = y.cpu().numpy(), b.cpu().numpy(), phi.cpu().numpy()
    return {
        "y_real": yc.real.copy(), "y_imag": yc.imag.copy(), "b_real": bc.real.copy(), "b_imag": bc.imag.copy(),
        "tau": t[:, :, 0].copy(), "f": t[:, :, 1].copy(), "C_real": t[:, :, 2].copy(), "C_imag": t[:, :, 3].copy(),
        "L_true": np.full((n_samples,), L_max, dtype=np.int32), "sigma": sigma.cpu().numpy(),
        "phi_real": pc.real.copy(), "phi_imag": pc.imag.copy(), "ser": ser.cpu().numpy(),
    }


def save_dataset(data_dir, splits, config):
    """generate_data.py:223-256."""
    os.makedirs(data_dir, exist_ok=True)
    for name, data in splits.items():
        d = os.path.join(data_dir, name)
        os.makedirs(d, exist_ok=True)
        for key, arr in data.items():
            np.save(os.path.join(d, f"{key}.npy"), arr)
    with open(os.path.join(data_dir, "dataset_config.json"), "w") as fh:
        json.dump(config, fh, indent=2)
    np.savez(os.path.join(data_dir, "dataset_info.npz"), **config)


def generate_dataset(data_dir, total_samples=10000, Nb=10, Nd=10, L_max=3, snr_range=(5, 25), train_ratio=0.7,
                     val_ratio=0.15, seed=0, device=None):
    """generate_complete_dataset (generate_data.py:46-84) on the device."""
    sizes = split_sizes(total_samples, train_ratio, val_ratio)
    splits = {name: generate_split(sz, Nb, Nd, L_max, snr_range, seed=seed + i, device=device)
              for i, (name, sz) in enumerate(zip(SPLITS, sizes))}
    config = {"Nb": Nb, "Nd": Nd, "L_max": L_max, "snr_range": list(snr_range), "total_samples": total_samples,
              "train_samples": sizes[0], "val_samples": sizes[1], "test_samples": sizes[2],
              "created_date": str(np.datetime64("now"))}
    save_dataset(data_dir, splits, config)
    return splits


def load_split(data_dir, split="train", pin=False):
    """The tensors of create_pytorch_dataloader's TensorDataset (generate_data.py:465-516), in its order:
    (y, b, tau, f, C, L_true, sigma, phi); phi is absent for datasets written by the base OFDMDatasetGenerator."""
    d = os.path.join(data_dir, split)
    if not os.path.isdir(d):
        raise ValueError(f"split {split} does not exist under {data_dir}")
    a = {os.path.splitext(fn)[0]: np.load(os.path.join(d, fn)) for fn in os.listdir(d) if fn.endswith(".npy")}
    cplx = lambda k: torch.complex(torch.from_numpy(a[k + "_real"]).float(), torch.from_numpy(a[k + "_imag"]).float())
    out = [cplx("y"), cplx("b"), torch.from_numpy(a["tau"]).float(), torch.from_numpy(a["f"]).float(), cplx("C"),
           torch.from_numpy(a["L_true"]).long(), torch.from_numpy(a["sigma"]).float()]
    if "phi_real" in a:
        out.append(cplx("phi"))
    if pin and torch.cuda.is_available():
        out = [t.pin_memory() for t in out]
    return tuple(out)


def save_checkpoint(path, epoch, model, optimizer, scheduler, best_val_loss, config, history):
    """trainPhi.py:238-246."""
    torch.save({"epoch": epoch, "model_state_dict": model.state_dict(), "optimizer_state_dict": optimizer.state_dict(),
                "scheduler_state_dict": scheduler.state_dict(), "best_val_loss": best_val_loss, "config": config,
                "history": history}, path)


def load_checkpoint(path, model, optimizer=None, scheduler=None, map_location=None):
    """trainPhi.py:127-133; -> (start_epoch, best_val_loss, checkpoint dict)."""
    ck = torch.load(path, map_location=map_location, weights_only=False)
    model.load_state_dict(ck["model_state_dict"])
    if optimizer is not None:
        optimizer.load_state_dict(ck["optimizer_state_dict"])
    if scheduler is not None:
        scheduler.load_state_dict(ck["scheduler_state_dict"])
    return ck["epoch"] + 1, ck["best_val_loss"], ck
