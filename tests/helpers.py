"""Shared helpers for the parity tests (test infrastructure)."""
import ast
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_net_case(tag):
    z = np.load(os.path.join(GOLDEN, f"net_{tag}.npz"))
    sd = {k[4:].replace("__", "."): torch.from_numpy(z[k]) for k in z.files if k.startswith("sd__")}
    return z, sd


def rel_err(a, b):
    """per-signal max-norm relative error."""
    a = np.asarray(a)
    b = np.asarray(b)
    return np.abs(a - b).max(axis=-1) / np.maximum(np.abs(b).max(axis=-1), 1e-30)


def parse_opts(arr):
    return ast.literal_eval(str(arr))
