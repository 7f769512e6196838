"""fp32 numpy emulation of csrc/arrow_kernels.cu (same formulas, same order of operations, one root at a time):
lets the layer-0 arrowhead algorithm — bracketing, shifted coordinates, osculatory two-pole iteration, Gu-Eisenstat
re-derivation of |phi_i| — be exercised on the CPU over adversarial inputs.  Test infrastructure only."""
import numpy as np

f32 = np.float32


def _eval(sd, sz2, js, org, a0, x):
    r = f32(1) / (x - (sd - org))
    t = sz2 * r
    psi, phi = t[:js].sum(dtype=f32), t[js:].sum(dtype=f32)
    q, h = (t[:js] * r[:js]).sum(dtype=f32), (t[js:] * r[js:]).sum(dtype=f32)
    return (a0 - x) + (psi + phi), -q - f32(.5), -h - f32(.5), abs(psi) + abs(phi)


def roots(sd, sz2, c0, maxit=48):
    """-> (origin pole, offset) of the n+1 roots, or None if one did not converge (the kernel then declines)."""
    n = len(sd)
    zn2 = sz2.sum(dtype=f32)
    org, xs = np.zeros(n + 1, f32), np.zeros(n + 1, f32)
    for j in range(n + 1):
        first, last = j == 0, j == n
        dL = dR = f32(0)
        if first or last:
            o = sd[0] if first else sd[n - 1]
            ap = f32(c0) - o
            rt = np.sqrt(ap * ap + 4 * zn2)
            if first:
                x = -2 * zn2 / (ap + rt) if ap > 0 else f32(.5) * (ap - rt)
                lo, hi = x * f32(1.0001) - f32(1e-30), f32(0)
            else:
                x = 2 * zn2 / (rt - ap) if ap < 0 else f32(.5) * (ap + rt)
                lo, hi = f32(0), x * f32(1.0001) + f32(1e-30)
        else:
            gap = sd[j] - sd[j - 1]
            o = sd[j - 1]
            g, _, _, _ = _eval(sd, sz2, j, o, f32(c0) - o, f32(.5) * gap)
            if g > 0:
                o, lo, hi, dL, dR = sd[j], -f32(.5) * gap, f32(0), -gap, f32(0)
                x = lo
            else:
                lo, hi, dL, dR = f32(0), f32(.5) * gap, f32(0), gap
                x = hi
        a0 = f32(c0) - o
        ok = False
        for _ in range(maxit):
            g, wl, wr, sa = _eval(sd, sz2, j, o, a0, x)
            if g > 0:
                lo = x
            else:
                hi = x
            if abs(g) <= f32(1.2e-7) * (8 * sa + abs(a0) + abs(x)):
                ok = True
                break
            if first or last:
                den = g + (wl + wr) * x
                eta = -g * x / den if den != 0 else f32(0)
            else:
                DL, DR = x - dL, x - dR
                s, S = -wl * DL * DL, -wr * DR * DR
                C = g + wl * DL + wr * DR
                a1, a0q = C * (DL + DR) + s + S, DL * DR * g
                q = a1 + np.copysign(np.sqrt(max(a1 * a1 - 4 * C * a0q, f32(0))), a1)
                eta = -2 * a0q / q if q != 0 else f32(0)
                xn = x + eta
                if not (lo < xn < hi) and C != 0 and eta != 0:
                    eta = a0q / (C * eta)
            xn = x + eta
            if not (lo < xn < hi):
                xn = f32(.5) * (lo + hi)
            done = xn == x or abs(xn - x) <= f32(6e-8) * abs(xn) or hi - lo <= f32(1.2e-7) * max(abs(lo), abs(hi))
            x = xn
            if done:
                ok = True
                break
        if not ok or not np.isfinite(x):
            return None
        org[j], xs[j] = o, x
    return org, xs


def arrow_eigh(h, phi, c0):
    """eigen-decomposition of [[diag(h), phi],[phi^H, c0]] the way k_arrow does it -> (lam ascending, U) or None."""
    n = len(h)
    order = np.argsort(h, kind="stable")
    sd = h[order].astype(f32)
    ph = phi[order]
    sz2 = (ph.real.astype(f32) ** 2 + ph.imag.astype(f32) ** 2).astype(f32)
    if np.any(np.diff(sd) <= 0) or np.any(sz2 <= 0):
        return None
    r = roots(sd, sz2, c0)
    if r is None:
        return None
    org, xs = r
    zh2 = np.zeros(n, f32)
    for i in range(n):
        si = sd[i]
        prod = ((si - org[0]) - xs[0]) * ((org[n] - si) + xs[n])
        for j in range(1, i + 1):
            prod = f32(prod * (((si - org[j]) - xs[j]) / (si - sd[j - 1])))
        for j in range(i + 1, n):
            prod = f32(prod * (((org[j] - si) + xs[j]) / (sd[j] - si)))
        zh2[i] = max(prod, f32(0))
    zph = (np.sqrt(zh2) * (ph / np.abs(ph))).astype(np.complex64)
    U = np.zeros((n + 1, n + 1), np.complex64)
    for j in range(n + 1):
        den = ((org[j] - sd) + xs[j]).astype(f32)
        nu = f32(1) / np.sqrt(f32(1) + np.sum(zh2 / (den * den), dtype=f32))
        U[order, j] = zph * (nu / den)
        U[n, j] = nu
    return (org + xs).astype(f32), U
