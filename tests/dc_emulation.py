"""CPU emulation (numpy, float32 arithmetic throughout) of the divide & conquer tridiagonal eigen-solver that
csrc/dc_kernels.cu runs on the GPU (k_dc): full binary tree down to 1 x 1 leaves, every merge = rank-one update
D + rho z z^T solved through its secular equation in coordinates shifted to the nearer pole, Gu-Eisenstat
re-derivation of z from the computed roots (orthogonal eigenvectors for any pole spacing), deflation of negligible z
and of (nearly) equal poles as LAPACK's xLAED2.  The emulation follows the kernel step by step (same formulas, same
tolerances) so that the algorithm's accuracy can be studied without a GPU; tests/test_host_cpu.py runs it on random,
clustered, graded and degenerate tridiagonals.  (The kernel's first merge level - blocks of one or two rows - is a closed-form 2 x 2 Jacobi
rotation, mirrored by merge2_closed_form; that its merge products skip the structural zeros of the other child's
columns does not change the mathematics and is not mirrored.)
"""
import numpy as np

F = np.float32
EPS = F(5.9604645e-08)        # 2^-24


def _levels(d):
    """ranges per level: level 0 = single indices; level l merges children split at p."""
    nl = 0
    while (1 << nl) < d:
        nl += 1
    out = []
    for l in range(1, nl + 1):
        nb = 1 << (nl - l)
        rs = []
        for r in range(nb):
            lo, hi = (r * d) // nb, ((r + 1) * d) // nb
            p = ((2 * r + 1) * d) // (2 * nb)
            if lo < p < hi:
                rs.append((lo, p, hi))
        out.append(rs)
    return out


def secular_roots(sd, sz2, rho):
    """k poles sd (ascending, distinct), weights sz2 = z^2 > 0, rho > 0.  Root j of  -1/rho + sum z2/(l - d) = 0 in
    (sd[j], sd[j+1]) (last: right of sd[k-1]).  Returns (org index, mu): l_j = sd[org_j] + mu_j."""
    k = len(sd)
    a0 = F(-1.0) / rho
    orgs = np.zeros(k, np.int64)
    mus = np.zeros(k, F)
    zsum = F(np.sum(sz2, dtype=F))
    for j in range(k):
        last = j == k - 1

        def ev(o, x):
            den = x - (sd - sd[o])                      # l - d_i in shifted coordinates
            r = F(1.0) / den
            t = sz2 * r
            left = np.arange(k) <= j
            psi, phi = F(np.sum(t[left], dtype=F)), F(np.sum(t[~left], dtype=F))
            wl, wr = -F(np.sum((t * r)[left], dtype=F)), -F(np.sum((t * r)[~left], dtype=F))
            return F(a0 + (psi + phi)), wl, wr, F(abs(psi) + abs(phi))

        if last:
            o = j
            lo, hi = F(rho * sz2[j]), F(rho * zsum * F(1.0001)) + F(1e-30)
            lo = F(lo * F(0.9999))
            x = F(0.5) * (lo + hi)
            dL = dR = F(0)
        else:
            gap = F(sd[j + 1] - sd[j])
            o = j
            g, _, _, _ = ev(o, F(0.5) * gap)
            if g > 0:                                    # g decreasing: root right of the midpoint -> origin j+1
                o = j + 1
                lo, hi, dL, dR = F(-0.5) * gap, F(0), -gap, F(0)
                x = lo
            else:
                lo, hi, dL, dR = F(0), F(0.5) * gap, F(0), gap
                x = hi
        for it in range(48):
            g, wl, wr, sa = ev(o, x)
            if g > 0:
                lo = x
            else:
                hi = x
            if abs(g) <= F(1.2e-7) * (F(8) * sa + abs(a0)):
                break
            if last:
                w, D = wl + wr, x
                den = g + w * D
                eta = -g * D / den if den != 0 else F(0)
            else:
                DL, DR = x - dL, x - dR
                s, S = -wl * DL * DL, -wr * DR * DR
                C = g + wl * DL + wr * DR
                a1, a0q = C * (DL + DR) + s + S, DL * DR * g
                disc = max(F(a1 * a1 - F(4) * C * a0q), F(0))
                q = a1 + np.copysign(np.sqrt(disc, dtype=F), a1)
                eta = F(-2) * a0q / q if q != 0 else F(0)
                xn = x + eta
                if not (lo < xn < hi) and C != 0 and eta != 0:
                    eta = a0q / (C * eta)
            xn = F(x + eta)
            if not (lo < xn < hi):
                xn = F(0.5) * (lo + hi)
            conv = xn == x or abs(xn - x) <= F(6e-8) * abs(xn) or (hi - lo) <= F(1.2e-7) * max(abs(lo), abs(hi))
            x = xn
            if conv:
                break
        orgs[j], mus[j] = o, x
    return orgs, mus


def merge(lam, Q, lo, p, hi, beta, stats=None):
    """Merge the solved blocks [lo,p) and [p,hi) torn at p (off-diagonal beta); lam, Q updated in place."""
    k0 = hi - lo
    rho = F(abs(beta))
    if rho == 0:
        return
    sgn = F(1.0) if beta >= 0 else F(-1.0)
    z = np.concatenate([Q[p - 1, lo:p], sgn * Q[p, p:hi]]).astype(F) * F(0.70710678)
    rho = F(2.0) * rho
    dloc = lam[lo:hi].copy()
    perm = np.argsort(dloc, kind="stable")
    sd, sz = dloc[perm], z[perm]
    cols = lo + perm                                     # column of Q for each sorted position
    tol = F(8.0) * EPS * max(F(np.abs(sd).max()), F(np.abs(sz).max()))
    # deflation scan (xLAED2): negligible z, then (nearly) equal poles by a Givens rotation
    keep = []
    prev = -1
    for t in range(k0):
        if rho * abs(sz[t]) <= tol:
            sz[t] = 0
            continue
        if prev >= 0:
            s_, c_ = sz[prev], sz[t]
            tau = F(np.hypot(c_, s_))
            tt = sd[t] - sd[prev]
            c, s = c_ / tau, -s_ / tau
            if abs(tt * c * s) <= tol:
                # rotate columns prev, t of Q: z[prev] -> 0, z[t] -> tau
                qp, qt = Q[lo:hi, cols[prev]].copy(), Q[lo:hi, cols[t]].copy()
                Q[lo:hi, cols[prev]] = c * qp + s * qt
                Q[lo:hi, cols[t]] = -s * qp + c * qt
                dp, dt = sd[prev], sd[t]
                sd[prev] = dp * c * c + dt * s * s
                sd[t] = dp * s * s + dt * c * c
                sz[t], sz[prev] = tau, 0
                keep.remove(prev)
                if stats is not None:
                    stats["givens"] = stats.get("givens", 0) + 1
        keep.append(t)
        prev = t
    keep = np.array(keep, dtype=np.int64)
    k = len(keep)
    if stats is not None:
        stats["deflated"] = stats.get("deflated", 0) + (k0 - k)
        stats["total"] = stats.get("total", 0) + k0
    newlam = sd.copy()                                   # deflated entries keep their (possibly rotated) poles
    if k > 0:
        ksd, ksz, kcol = sd[keep], sz[keep], cols[keep]
        # the rotation can leave ksd slightly out of order / equal: enforce strict ascent by the deflation itself
        orgs, mus = secular_roots(ksd, ksz * ksz, rho)
        # Gu-Eisenstat: zhat_i^2 = (l_{k-1} - d_i)/rho * prod_{j<i} (d_i - l_j)/(d_i - d_j) * prod_{j=i}^{k-2} (l_j - d_i)/(d_{j+1} - d_i)
        zh = np.zeros(k, F)
        for i in range(k):
            di = ksd[i]
            prod = ((ksd[orgs[k - 1]] - di) + mus[k - 1]) / rho
            for j in range(i):
                prod = prod * (((di - ksd[orgs[j]]) - mus[j]) / (di - ksd[j]))
            for j in range(i, k - 1):
                prod = prod * (((ksd[orgs[j]] - di) + mus[j]) / (ksd[j + 1] - di))
            zh[i] = np.copysign(np.sqrt(max(prod, F(0)), dtype=F), ksz[i])
        W = np.zeros((k, k), F)
        for j in range(k):
            den = (ksd - ksd[orgs[j]]) - mus[j]          # d_i - l_j
            w = zh / den
            W[:, j] = w / F(np.sqrt(np.sum(w * w, dtype=F), dtype=F))
            newlam[keep[j]] = ksd[orgs[j]] + mus[j]
        Qn = (Q[lo:hi][:, kcol].astype(F) @ W).astype(F)
        Q[lo:hi, kcol] = Qn
    lam[cols] = newlam


def merge2_closed_form(lam, Q, lo, beta):
    """The first merge level as csrc/dc_kernels.cu does it: a block of two leaves is the 2 x 2 problem
    [[l0 + |b|, b], [b, l1 + |b|]], diagonalised by one Jacobi rotation (float32).  Columns of Q = eigenvectors."""
    b = F(beta)
    if b == 0:
        return
    aa, cc = F(lam[lo] + abs(b)), F(lam[lo + 1] + abs(b))
    theta = F((cc - aa) / (F(2) * b))
    t = F(np.copysign(F(1), theta) / (abs(theta) + np.sqrt(F(theta * theta + F(1)), dtype=F)))
    cs = F(F(1) / np.sqrt(F(t * t + F(1)), dtype=F))
    sn = F(t * cs)
    lam[lo], lam[lo + 1] = F(aa - t * b), F(cc + t * b)
    Q[lo, lo], Q[lo + 1, lo] = cs, -sn          # eigenvector of lam[lo]   (kernel: row lo of Q^T = (cs, -sn))
    Q[lo, lo + 1], Q[lo + 1, lo + 1] = sn, cs   # eigenvector of lam[lo+1]


def dc_eigh(dT, eT, stats=None, closed_form_level1=True):
    """Symmetric tridiagonal (diagonal dT [d], off-diagonal eT [d-1]) -> (lam [d] unsorted, Q [d][d]) in float32."""
    d = len(dT)
    dT, eT = np.asarray(dT, F), np.asarray(eT, F)
    lam = dT.copy()
    lam[:-1] -= np.abs(eT)
    lam[1:] -= np.abs(eT)
    Q = np.eye(d, dtype=F)
    for li, rs in enumerate(_levels(d)):
        for (lo, p, hi) in rs:
            if li == 0 and closed_form_level1:
                assert hi - lo == 2 and p == lo + 1
                merge2_closed_form(lam, Q, lo, eT[p - 1])
            else:
                merge(lam, Q, lo, p, hi, eT[p - 1], stats)
    return lam, Q


def check(dT, eT):
    d = len(dT)
    T = np.diag(np.asarray(dT, np.float64)) + np.diag(np.asarray(eT, np.float64), 1) + np.diag(np.asarray(eT, np.float64), -1)
    st = {}
    lam, Q = dc_eigh(dT, eT, st)
    nrm = max(np.abs(np.linalg.eigvalsh(T)).max(), 1e-30)
    res = np.abs(T @ Q - Q * lam).max() / nrm
    orth = np.abs(Q.T.astype(np.float64) @ Q - np.eye(d)).max()
    ev = np.abs(np.sort(lam) - np.linalg.eigvalsh(T)).max() / nrm
    return res, orth, ev, st


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    for name, (dT, eT) in {
        "random": (rng.normal(size=101) * 3, rng.normal(size=100)),
        "wilkinson": (np.abs(np.arange(101) - 50.0), np.ones(100)),
        "graded": (10.0 ** -np.linspace(0, 6, 101), 10.0 ** -np.linspace(0, 6, 101)[:-1] * 0.5),
        "clustered": (np.ones(101) + 1e-5 * rng.normal(size=101), 1e-3 * rng.normal(size=100)),
        "toeplitz": (2 * np.ones(101), -np.ones(100)),
        "zeros_e": (rng.normal(size=101), np.zeros(100)),
        "some_zero_e": (rng.normal(size=101), rng.normal(size=100) * (rng.random(100) > 0.3)),
        "identity": (np.ones(101), np.zeros(100)),
        "small": (rng.normal(size=5), rng.normal(size=4)),
        "two": (np.array([1.0, 2.0]), np.array([0.5])),
    }.items():
        print(name, check(dT, eT))
