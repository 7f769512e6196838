"""Drop-in mirror of the reference's `admm` module API: `admm_for_us` (admm.py:6-114).

The reference loop — dense inverse (77-79), ECOS projection (117-148), SVD "PSD projection" (151-179),
dual update (88-92), stopping rule (95-112) — evaluates, for every input reachable through this
signature, the linear recursion  phi_k = M^{-1}(y/b + rho*phi_{k-1}),  M = diag(1/|b|^2) + rho*11^T,
for  K_exec = min(max_iter, max(min_iter, 2))  iterations (pinned by
tests/golden/classic.npz).  That recursion runs on the GPU in fp64 (csrc/classic_kernels.cu).
"""
import numpy as np
import torch

from . import _capi


def executed_iterations(opts=None, use_min_iter=True, min_iter=5):
    max_iter = 500 if opts is None else opts.get("max_iter", 500)        # admm.py:36-45
    return min(max_iter, max(min_iter, 2) if use_min_iter else 2)


def admm_for_us_batched(y, b, rho=1.0, n_iter=5, out=None):
    """y, b: [B, n] complex64/complex128 torch tensors on the GPU -> phi complex128 [B, n] (device)."""
    _capi.require_cuda()
    if y.shape != b.shape or y.dim() != 2:
        raise ValueError("y and b must both be [B, n]")
    if y.dtype != b.dtype or y.dtype not in (torch.complex64, torch.complex128):
        raise ValueError("y and b must share dtype complex64 or complex128")
    y, b = y.contiguous(), b.contiguous()
    B, n = y.shape
    if out is None:
        out = torch.empty(B, n, dtype=torch.complex128, device=y.device)
    stream = torch.cuda.current_stream(y.device).cuda_stream
    _capi.check(_capi.lib().admm_classic_forward(y.data_ptr(), b.data_ptr(), int(y.dtype == torch.complex128), B, n,
                                                 float(rho), int(n_iter), out.data_ptr(), stream))
    return out


def admm_for_us(y, b, xbase, ybase, lambda_val, sigma, opts=None, use_min_iter=True, min_iter=5):
    """Same signature and return contract as admm.py:6: (phi complex128 (n,), iter_count).
    xbase/ybase/lambda_val/sigma are accepted for compatibility; they never reach phi."""
    rho = 1.0 if opts is None else opts.get("rho", 1.0)
    y = np.asarray(y).flatten().astype(np.complex128)          # admm.py:48-49
    b = np.asarray(b).flatten().astype(np.complex128)
    if y.shape != b.shape:
        raise ValueError("y and b must have the same number of elements")
    n_iter = executed_iterations(opts, use_min_iter, min_iter)
    _capi.require_cuda()
    dev = torch.device("cuda", torch.cuda.current_device())
    yt = torch.from_numpy(y).to(dev).reshape(1, -1)
    bt = torch.from_numpy(b).to(dev).reshape(1, -1)
    phi = admm_for_us_batched(yt, bt, rho, n_iter)
    return phi[0].cpu().numpy(), n_iter
