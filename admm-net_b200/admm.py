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


# The reference's stopping test (admm.py:95-112) compares a primal residual that is pure SVD round-off
# (~1e-15 * ||block||_F) and a dual residual that is exactly 0 with
#     eta_pri = eta_abs*sqrt(n+1) + eta_rel*max(||G||_F, ||block||_F).
# Tolerances comfortably above that round-off stop the loop at the first tested iteration; tolerances of exactly 0
# never stop it (max_iter iterations); anything in between depends on LAPACK's rounding and is refused.
_ROUNDOFF_SAFE = 1e-11


def executed_iterations(opts=None, use_min_iter=True, min_iter=5, block_norm=None, n=None):
    """Iteration count admm_for_us executes.  With `block_norm` (= ||[[0,phi],[phi^H,1/lambda^2]]||_F) and n given,
    the tolerances eta_abs/eta_rel are honoured as described above; without them the default-tolerance count."""
    opts = opts or {}
    max_iter = opts.get("max_iter", 500)                                  # admm.py:36-45
    first_test = max(min_iter, 2) if use_min_iter else 2
    if block_norm is None:
        return min(max_iter, first_test)
    eta_abs, eta_rel = opts.get("eta_abs", 1e-5), opts.get("eta_rel", 1e-5)
    eta_pri = eta_abs * np.sqrt(n + 1) + eta_rel * block_norm
    if eta_pri >= _ROUNDOFF_SAFE * block_norm:
        return min(max_iter, first_test)
    if eta_abs == 0 and eta_rel == 0:
        return max_iter
    raise ValueError("admm_for_us: eta_abs/eta_rel below ~1e-11 relative (but not both 0) make the reference's "
                     "stopping iteration depend on SVD round-off; not reproducible, refused")


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
    _capi.require_cuda()
    dev = torch.device("cuda", torch.cuda.current_device())
    yt = torch.from_numpy(y).to(dev).reshape(1, -1)
    bt = torch.from_numpy(b).to(dev).reshape(1, -1)
    n_iter = executed_iterations(opts, use_min_iter, min_iter)
    phi = admm_for_us_batched(yt, bt, rho, n_iter)[0].cpu().numpy()
    # honour eta_abs / eta_rel (admm.py:95-112): the block matrix the test looks at is [[0, phi],[phi^H, 1/lambda^2]]
    block_norm = float(np.sqrt(2.0 * np.sum(np.abs(phi) ** 2) + 1.0 / float(lambda_val) ** 4))
    n_exec = executed_iterations(opts, use_min_iter, min_iter, block_norm, y.size)
    if n_exec != n_iter:
        n_iter = n_exec
        phi = admm_for_us_batched(yt, bt, rho, n_iter)[0].cpu().numpy()
    return phi, n_iter
