"""CPU emulation of k_tail_tc's algorithm (test infrastructure): blocked compact-WY back-transformation of the real
eigenvectors of T with every complex product written as the real-plane GEMMs the kernel issues (same operand roles,
signs and hi/lo split terms), followed by the Hermitian rebuild G = W^H-form with W = sqrt(l') U^T.

    X = U^T (rows = eigenvector index, columns = coordinate), starts as Z^T (real)
    block j (reflectors k0 .. k0+nb-1, applied from the LAST block to the first):
        T_j   upper triangular, T[i,i] = tau_i, T[:i,i] = -tau_i T[:i,:i] (V[:, :i]^H v_i)      (zlarft 'F','C')
        Y_j = V_j T_j
        P   = X conj(V_j)          Pr = Xr Vr + Xi Vi,   Pi = Xi Vr - Xr Vi
        X  -= P Y_j^T              Xr -= Pr Yr^T - Pi Yi^T,   Xi -= Pr Yi^T + Pi Yr^T
    G[a,b] = sum_i W[i,a] conj(W[i,b]),   W = sqrt(l')_i X[i,:]
        Gr = Wr^T Wr + Wi^T Wi,   Gi = Wi^T Wr - Wr^T Wi
"""
import numpy as np


def tf32_trunc(x):
    return (np.ascontiguousarray(x, dtype=np.float32).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def split(x):
    x = np.ascontiguousarray(x, dtype=np.float32)
    hi = tf32_trunc(x)
    lo = tf32_trunc((x - hi).astype(np.float32))        # the tensor core sees the leading 10 bits of lo
    return hi, lo


def mm3(A, B, exact=False):
    """A @ B.T the way the kernel does it: hi*hi + hi*lo + lo*hi with fp32 accumulation (exact: fp64 reference)."""
    if exact:
        return A.astype(np.float64) @ B.astype(np.float64).T
    ah, al = split(A)
    bh, bl = split(B)
    return (ah @ bh.T + ah @ bl.T + al @ bh.T).astype(np.float32)


def block_plan(d, nb_max=24):
    """reflectors 0..d-2 in nblk uniform blocks of at most nb_max"""
    nref = d - 1
    nblk = (nref + nb_max - 1) // nb_max
    nb = (nref + nblk - 1) // nblk
    return [(k0, min(k0 + nb, nref)) for k0 in range(0, nref, nb)]


def larft(V, tau):
    """T of H_0 H_1 ... H_{m-1} = I - V T V^H (forward, columnwise)."""
    m = V.shape[1]
    T = np.zeros((m, m), dtype=V.dtype)
    for i in range(m):
        T[i, i] = tau[i]
        if i:
            T[:i, i] = -tau[i] * (T[:i, :i] @ (V[:, :i].conj().T @ V[:, i]))
    return T


def back_transform_blocked(Zt, Vfull, tau, exact=False, nb_max=24):
    """Zt [d,d] real (rows = eigenvectors of T); Vfull [d,d-1] unit-lower reflectors; returns X = U^T (complex)."""
    d = Zt.shape[0]
    dt = np.float64 if exact else np.float32
    Xr, Xi = Zt.astype(dt).copy(), np.zeros((d, d), dt)
    for (k0, k1) in reversed(block_plan(d, nb_max)):
        V = Vfull[:, k0:k1]
        T = larft(V.astype(np.complex128), tau[k0:k1].astype(np.complex128))
        Y = V.astype(np.complex128) @ T
        if not exact:
            V = V.astype(np.complex64)
            Y = Y.astype(np.complex64)
        a0 = (k0 + 1) & ~7                                      # coordinate range touched by this block, 8-aligned
        Vr, Vi = np.ascontiguousarray(V.real[a0:].T), np.ascontiguousarray(V.imag[a0:].T)     # [nb, K] K-major
        Yr, Yi = np.ascontiguousarray(Y.real[a0:]), np.ascontiguousarray(Y.imag[a0:])         # [N, nb] K-major
        Pr = mm3(Xr[:, a0:], Vr, exact) + mm3(Xi[:, a0:], Vi, exact)
        Pi = mm3(Xi[:, a0:], Vr, exact) - mm3(Xr[:, a0:], Vi, exact)
        Xr[:, a0:] = Xr[:, a0:] - mm3(Pr, Yr, exact) + mm3(Pi, Yi, exact)
        Xi[:, a0:] = Xi[:, a0:] - mm3(Pr, Yi, exact) - mm3(Pi, Yr, exact)
    return Xr, Xi


def rebuild(Xr, Xi, lamp, exact=False):
    s = np.sqrt(lamp.astype(Xr.dtype))[:, None]
    Wr, Wi = (s * Xr).T.copy(), (s * Xi).T.copy()               # [a, i]: K-major over i
    Gr = mm3(Wr, Wr, exact) + mm3(Wi, Wi, exact)
    Gi = mm3(Wi, Wr, exact) - mm3(Wr, Wi, exact)
    return Gr, Gi


def hetrd_lower(A):
    """LAPACK zhetrd('L') convention in plain numpy: A = Q T Q^H, Q = H_0 ... H_{d-2}, H_k = I - tau_k v_k v_k^H,
    v_k[k+1] = 1, zeros above; returns (dd, ee, V [d, d-1], tau)."""
    A = A.astype(np.complex128).copy()
    d = A.shape[0]
    V = np.zeros((d, d - 1), np.complex128)
    tau = np.zeros(d - 1, np.complex128)
    ee = np.zeros(d - 1)
    for k in range(d - 1):
        x = A[k + 1:, k].copy()
        alpha = x[0]
        xn = np.linalg.norm(x[1:])
        if xn == 0 and alpha.imag == 0:
            t, beta, v = 0.0, alpha.real, np.zeros_like(x)
            v[0] = 1
        else:
            beta = -np.copysign(np.sqrt(abs(alpha) ** 2 + xn ** 2), alpha.real)
            t = (beta - alpha.real) / beta - 1j * alpha.imag / beta
            v = x / (alpha - beta)
            v[0] = 1
        V[k + 1:, k], tau[k], ee[k] = v, t, beta
        H = np.eye(d, dtype=np.complex128)
        H[k + 1:, k + 1:] -= t * np.outer(v, v.conj())
        A = H.conj().T @ A @ H
    return np.real(np.diag(A)).copy(), ee, V, tau


def selftest(d=101, seed=0):
    import scipy.linalg as sl
    rng = np.random.default_rng(seed)
    M = rng.normal(size=(d, d)) + 1j * rng.normal(size=(d, d))
    A = (M + M.conj().T) / 2
    dd, ee, V, tau = hetrd_lower(A)
    lam, Z = sl.eigh_tridiagonal(dd, ee)
    Q = np.eye(d, dtype=np.complex128)
    for k in range(d - 1):
        Q = Q @ (np.eye(d) - tau[k] * np.outer(V[:, k], V[:, k].conj()))
    U = Q @ Z
    assert np.abs(A @ U - U * lam).max() < 1e-10 * np.abs(lam).max() * d
    out = {}
    for exact in (True, False):
        Xr, Xi = back_transform_blocked(Z.T.copy(), V, tau, exact)
        X = Xr + 1j * Xi
        out["U_err_exact" if exact else "U_err_3xtf32"] = np.abs(X - U.T).max()
        lamp = np.abs(lam) + 0.1
        Gr, Gi = rebuild(Xr, Xi, lamp, exact)
        G = (U * lamp) @ U.conj().T
        out["G_err_exact" if exact else "G_err_3xtf32"] = np.abs((Gr + 1j * Gi) - G).max() / np.abs(G).max()
    return out


if __name__ == "__main__":
    for d in (101, 65, 17):
        print(d, selftest(d))
