"""CPU restatement of the reference's classical ADMM solver (TEST INFRASTRUCTURE ONLY).

Follows /root/reference/admm.py:
    admm_for_us           admm.py:6-114   (loop, slices 71-74, phi-update 77-79, Z-update 88-92,
                                           min-iter gate 95-96, stopping rule 99-112)
    admm_for_us_H_cvx_0   admm.py:117-148 (ECOS solve of  min ||h - t||  s.t. A*||h||_inf + sum(h) <= 1)
    admm_for_us_G_svd     admm.py:151-179

Third-party arithmetic absent from /root/reference: cvxpy + ECOS (no version pinned anywhere in the
reference; call site admm.py:129-141).  It is replaced here by the exact Euclidean projection onto
the same convex set (the optimisation problem has a unique solution, ECOS returns it to ~1e-8).
Parity for the H-update is therefore UNPINNED against ECOS itself; it is moot on every input reachable
through the public signature because t == 0 exactly (G and Z vanish on the [:n,:n] block diagonal,
SURVEY.md §8a-9), so h == 0 for any correct solver.  Everything else is pinned by
tests/golden/classic_*.npz, produced by running the reference's own admm_for_us with that single
function patched (tests/golden/make_golden.py).

`admm_linear_recursion` is the closed form the CUDA kernel evaluates; the tests
check it against the literal loop.
"""
import numpy as np
from scipy.linalg import svd


def project_linf_sum(t, A):
    """argmin_h ||h - t||_2  s.t.  A*max|h_i| + sum(h_i) <= 1   (t real, A >= 0)."""
    t = np.asarray(t, dtype=float)
    if A * np.max(np.abs(t)) + np.sum(t) <= 1.0:
        return t.copy()

    def inner(nu):
        v = t - nu
        a = np.abs(v)
        # s minimises nu*A*s + 0.5*sum(max(|v|-s,0)^2): root of sum(max(|v|-s,0)) = nu*A
        if a.sum() <= nu * A:
            s = 0.0
        else:
            lo, hi = 0.0, a.max()
            for _ in range(200):
                s = 0.5 * (lo + hi)
                if np.maximum(a - s, 0).sum() > nu * A:
                    lo = s
                else:
                    hi = s
            s = 0.5 * (lo + hi)
        h = np.clip(v, -s, s)
        return h, A * np.max(np.abs(h)) + h.sum()

    lo, hi = 0.0, np.max(np.abs(t)) + 1.0
    while inner(hi)[1] > 1.0:
        hi *= 2
    for _ in range(200):
        nu = 0.5 * (lo + hi)
        if inner(nu)[1] > 1.0:
            lo = nu
        else:
            hi = nu
    return inner(hi)[0]


def h_update(GK_hat, ZK_hat, rho, xbase, ybase, sigma):
    """admm.py:117-148 with the ECOS call replaced by the exact projection (real h => project Re t)."""
    n = xbase * ybase
    diag_GZ = np.diag(GK_hat + ZK_hat / rho)
    A = 2 * np.sqrt(n) * sigma + sigma ** 2
    return np.diag(project_linf_sum(diag_GZ.real, A))


def g_update_svd(HK, phiK, lambda_val, ZK, rho):
    """admm.py:151-179"""
    n = HK.shape[0]
    sd = np.zeros((n + 1, n + 1), dtype=complex)
    sd[:n, :n] = HK
    sd[:n, n] = phiK
    sd[n, :n] = phiK.conj().T
    sd[n, n] = 1.0 / (lambda_val ** 2)
    sd = sd - ZK / rho
    U, S, Vh = svd(sd)
    S[S < 0] = 0
    Sm = np.zeros_like(sd, dtype=complex)
    np.fill_diagonal(Sm, S)
    return U @ Sm @ Vh


def admm_for_us(y, b, xbase, ybase, lambda_val, sigma, opts=None, use_min_iter=True, min_iter=5):
    """Literal restatement of admm.py:6-114 (without the two print calls)."""
    rho, max_iter, eta_abs, eta_rel = 1.0, 500, 1e-5, 1e-5
    if opts is not None:
        rho = opts.get("rho", rho)
        max_iter = opts.get("max_iter", max_iter)
        eta_abs = opts.get("eta_abs", eta_abs)
        eta_rel = opts.get("eta_rel", eta_rel)
    y = np.asarray(y).flatten()
    b = np.asarray(b).flatten()
    n = y.shape[0]
    GK = np.zeros((n + 1, n + 1), dtype=complex)
    ZK = np.zeros((n + 1, n + 1), dtype=complex)
    HK = np.zeros((n, n), dtype=complex)
    phiK = np.zeros(n, dtype=complex)
    it = 0
    for it in range(1, max_iter + 1):
        HK_pre = np.zeros((n, n), dtype=complex) if it == 1 else HK.copy()
        GK_hat, gK = GK[:n, :n], GK[:n, n]
        ZK_hat, zetaK = ZK[:n, :n], ZK[:n, n]
        diag_inv = np.linalg.inv(np.diag(b * np.conj(b))) + rho * np.ones(n)      # admm.py:78 (broadcast!)
        phiK = np.linalg.inv(diag_inv) @ (np.linalg.inv(np.diag(b)) @ y + rho * gK + zetaK)
        HK = h_update(GK_hat, ZK_hat, rho, xbase, ybase, sigma)
        GK = g_update_svd(HK, phiK, lambda_val, ZK, rho)
        blk = np.vstack([np.hstack([HK, phiK.reshape(-1, 1)]),
                         np.hstack([phiK.conj().T, 1.0 / (lambda_val ** 2)])])
        ZK = ZK + rho * (GK - blk)
        if use_min_iter and it < min_iter:
            continue
        if it > 1:
            eta_pri = eta_abs * np.sqrt(n + 1) + eta_rel * max(np.linalg.norm(GK, "fro"), np.linalg.norm(blk, "fro"))
            eta_dual = eta_abs * np.sqrt(n) + eta_rel * np.linalg.norm(ZK, "fro")
            if np.linalg.norm(GK - blk, "fro") <= eta_pri and np.linalg.norm(rho * (HK - HK_pre), "fro") <= eta_dual:
                break
    return phiK, it


def executed_iterations(max_iter, use_min_iter=True, min_iter=5):
    """Iteration count admm_for_us returns for every input."""
    return min(max_iter, max(min_iter, 2) if use_min_iter else 2)


def admm_linear_recursion(y, b, rho, n_iter):
    """phi_k = M^{-1}(y/b + rho*phi_{k-1}), M = diag(1/|b|^2) + rho*11^T, by Sherman-Morrison.
    y,b: [B,n] complex128 -> [B,n] complex128."""
    y = np.asarray(y, dtype=complex)
    b = np.asarray(b, dtype=complex)
    D = (b * np.conj(b)).real                            # diag of inv(diag(1/|b|^2))
    den = 1.0 + rho * D.sum(axis=-1, keepdims=True)
    phi = np.zeros_like(y)
    for _ in range(n_iter):
        v = y / b + rho * phi
        Dv = D * v
        phi = Dv - rho * D * (Dv.sum(axis=-1, keepdims=True) / den)
    return phi
