"""CPU restatement of the reference's grid peak search (TEST INFRASTRUCTURE ONLY).

Follows /root/reference/utils/peakSearchUtils.py:
    peak_search_func   peakSearchUtils.py:9-33   (+ utils/mathUtils.py:4-21 vander_vec)
    peak_search        peakSearchUtils.py:37-60
    alt_peak_search    peakSearchUtils.py:63-173
and restates the one third-party routine on the path that is absent from /root/reference:
    skimage.morphology.local_maxima(image, connectivity=2)   (scikit-image, version unpinned by the
    reference — no requirements file; call site peakSearchUtils.py:118).
Published algorithm (skimage/morphology/extrema.py + _extrema_cy.pyx): pad the image with its global
minimum (allow_borders=True), mark every pixel with no strictly greater 8-neighbour as a candidate,
flood-fill each plateau of equal values; a plateau is a maximum iff none of its pixels has a strictly
greater neighbour and it does not reach the padding ring with the padding's value.
The only KAT the reference holds for it is the 4x5 plateau image at peakSearchUtils.py:427-436
(expected mask derivable by hand: exactly the 2x2 block of 5s) -> tests/test_oracle.py.  Beyond that
the local_maxima restatement is PARITY-UNPINNED against scikit-image itself (not installed here).
tests/golden/peaks_*.npz hold outputs of the reference's own alt_peak_search/peak_search run with
this local_maxima injected as `skimage.morphology` (tests/golden/make_golden.py).
"""
import numpy as np


def vander_vec(x, y, length):
    """mathUtils.py:4-21"""
    fre = np.linspace(x, y, length)
    return np.exp(1j * 2 * np.pi * fre).reshape(-1, 1)


def peak_search_func(phi, x, x_base, y, y_base):
    """peakSearchUtils.py:9-33"""
    s = vander_vec(0, (y_base - 1) * y, y_base)
    d = vander_vec(0, (x_base - 1) * x, x_base)
    a = np.kron(s, np.conj(d))
    return np.abs(np.dot(phi.conj().T, a)) ** 2


def peak_search(phi, X, x_base, Y, y_base):
    """peakSearchUtils.py:37-60 (literal double loop)."""
    out = np.zeros((Y.shape[0], X.shape[1]))
    for i in range(Y.shape[0]):
        for j in range(X.shape[1]):
            out[i, j] = np.squeeze(peak_search_func(phi, X[i, j], x_base, Y[i, j], y_base))
    return out


def peak_search_separable(phi, xs, x_base, ys, y_base):
    """Same surface on the tensor grid ys x xs as two small matrix products
    (Z = |S conj(Phi) conj(D)^T|^2); used by tests to cross-check kernels
    on grids where the literal loop is too slow."""
    S = np.stack([vander_vec(0, (y_base - 1) * y, y_base)[:, 0] for y in ys])      # [Gy,yb]
    D = np.stack([vander_vec(0, (x_base - 1) * x, x_base)[:, 0] for x in xs])      # [Gx,xb]
    P = np.asarray(phi).reshape(y_base, x_base)
    return np.abs(S @ np.conj(P) @ np.conj(D).T) ** 2


def local_maxima(image, connectivity=2):
    """8-connected, plateau-aware, allow_borders=True (see module docstring)."""
    assert connectivity == 2 and image.ndim == 2
    img = np.asarray(image)
    H, W = img.shape
    pad = np.full((H + 2, W + 2), img.min(), dtype=img.dtype)
    pad[1:-1, 1:-1] = img
    offs = [(-1, -1), (-1, 0), (-1, 1), (0, -1), (0, 1), (1, -1), (1, 0), (1, 1)]
    out = np.zeros((H, W), dtype=bool)
    visited = np.zeros((H + 2, W + 2), dtype=bool)
    for i in range(1, H + 1):
        for j in range(1, W + 1):
            if visited[i, j]:
                continue
            h = pad[i, j]
            # quick reject: some neighbour strictly greater
            if any(pad[i + di, j + dj] > h for di, dj in offs):
                continue
            # flood fill the plateau
            stack, plateau, is_max = [(i, j)], [], True
            visited[i, j] = True
            while stack:
                ci, cj = stack.pop()
                plateau.append((ci, cj))
                for di, dj in offs:
                    ni, nj = ci + di, cj + dj
                    v = pad[ni, nj]
                    if v > h:
                        is_max = False
                    elif v == h:
                        if ni == 0 or nj == 0 or ni == H + 1 or nj == W + 1:
                            is_max = False          # plateau reaches the padding ring
                        elif not visited[ni, nj]:
                            visited[ni, nj] = True
                            stack.append((ni, nj))
            if is_max:
                for ci, cj in plateau:
                    out[ci - 1, cj - 1] = True
    return out


DEFAULT_OPTS = {"xmin": 0, "xmax": 1, "xstep": 0.01, "ymin": -0.5, "ymax": 0.5, "ystep": 0.01,
                "reducefactor": 0.1, "iter": 1}


def alt_peak_search(func_opts, opts=None, surface=peak_search):
    """peakSearchUtils.py:63-173"""
    so = {**DEFAULT_OPTS, **(opts or {})}
    phi, xb, yb = func_opts["phi"], func_opts["xbase"], func_opts["ybase"]
    xmin, xmax, xstep = so["xmin"], so["xmax"], so["xstep"]
    ymin, ymax, ystep = so["ymin"], so["ymax"], so["ystep"]
    rf, iters = so["reducefactor"], so["iter"]
    ax = np.arange(xmin, xmax - xstep, xstep)
    ay = np.arange(ymin, ymax - xstep, ystep)          # peakSearchUtils.py:106 (xstep, sic)
    if len(ax) == 0 or len(ay) == 0:
        return np.zeros((0, 3))
    AX, AY = np.meshgrid(ax, ay)
    Zs = surface(phi, AX, xb, AY, yb)
    r, c = np.where(local_maxima(Zs, connectivity=2))
    P = len(r)
    res = np.zeros((P, 3))
    res[:, 0] = AX[r, c]
    res[:, 1] = AY[r, c]
    lx, ly = xstep, ystep
    for _ in range(iters):
        lx, ly = rf * lx, rf * ly
        for k in range(P):
            x0, x1 = max(xmin, res[k, 0] - lx), min(xmax - lx, res[k, 0] + lx)
            y0, y1 = max(ymin, res[k, 1] - ly), min(ymax - ly, res[k, 1] + ly)
            if x0 >= x1 or y0 >= y1:
                continue
            gx, gy = np.arange(x0, x1, lx), np.arange(y0, y1, ly)
            if len(gx) == 0 or len(gy) == 0:
                continue
            GX, GY = np.meshgrid(gx, gy)
            Zl = surface(phi, GX, xb, GY, yb)
            m = np.max(Zl)
            pos = np.where(Zl == m)
            if len(pos[0]) > 0:
                res[k, 0] = GX[pos[0][0], pos[1][0]]
                res[k, 1] = GY[pos[0][0], pos[1][0]]
                res[k, 2] = m
    return res


def top_l(peaks, L):
    """What every caller does with the result (main_for_net.py:119-126, main.py:114-120):
    stable sort by height, descending, keep the first L rows."""
    return np.array(sorted(peaks, key=lambda p: p[2], reverse=True)[:L]).reshape(-1, 3)
