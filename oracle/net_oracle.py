"""CPU restatement of the reference's unrolled ADMM-Net forward (TEST INFRASTRUCTURE ONLY).

This module is the parity oracle for the CUDA path.  It restates, op for op and in the
reference's dtypes (fp32 / complex64, torch CPU), what
/root/reference/admm_net.py::PhiEstADMMNet.forward (admm_net.py:742-764) computes:

    PhiLayer.forward                admm_net.py:79-105
    HLayer.forward / projection     admm_net.py:134-194
    GLayer (build/eigh/map/rebuild) admm_net.py:262-354
    ZLayer (constraint/step)        admm_net.py:388-474

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import it.  The product path (admm-net_b200/) never does.

Parity pinning: the reference ships no golden vectors for this path.  The
restatement is pinned instead against outputs of the reference itself, generated in the build
container by tests/golden/make_golden.py (which imports /root/reference/admm_net.py unmodified)
and committed under tests/golden/.  tests/test_oracle.py checks this file against those vectors.
"""
import math

import torch
import torch.nn.functional as F

EPS = 1e-8  # admm_net.py:74,114,211,360


def _sp(x):
    return F.softplus(x)


def lambda_inv_scalar(lambda_param, eps=EPS):
    """c = 1/(softplus(lambda)^2+eps) taken through .item() (admm_net.py:269-271, 424-426):
    computed in fp32 tensor arithmetic, then a Python float, then an fp32 fill."""
    lam = _sp(lambda_param)
    return (1.0 / (lam ** 2 + eps)).item()


def phi_update(y, b, G, Z, rho_param):
    """admm_net.py:90-103"""
    g = G[:, :-1, -1]
    zeta = Z[:, :-1, -1]
    b_sq = torch.abs(b) ** 2 + EPS
    rho = _sp(rho_param)
    weight = b_sq / (1 + rho * b_sq)
    return weight * (y / (b + EPS) + rho * g + zeta)


def h_update(G, Z, sigma, n, p):
    """admm_net.py:146-192; p holds rho, projection_weight, W1,b1,W2,b2. Returns h [B,n] (real)."""
    rho = _sp(p["rho"])
    T = G[:, :n, :n] + Z[:, :n, :n] / (rho + EPS)
    t = torch.diagonal(T, dim1=1, dim2=2).real
    A = 2 * torch.sqrt(torch.tensor(n).float()).to(sigma.device) * sigma + sigma ** 2
    A = A.view(-1, 1)
    corr = torch.tanh(F.linear(F.relu(F.linear(t, p["W1"], p["b1"])), p["W2"], p["b2"]))
    tc = t + 0.1 * corr
    linf = torch.max(torch.abs(tc), dim=1, keepdim=True)[0]
    tr = torch.sum(tc, dim=1, keepdim=True)
    cv = A * linf + tr
    scale = torch.sigmoid(p["projection_weight"]) / (cv + EPS)
    scale = torch.clamp(scale, max=1.0)
    return tc * scale


def block_matrix(phi, h, c):
    """[[diag(h), phi],[phi^H, c]] as complex64 (admm_net.py:273-284, 428-439)."""
    B, n = phi.shape
    M = torch.zeros(B, n + 1, n + 1, dtype=torch.complex64, device=phi.device)
    idx = torch.arange(n, device=phi.device)
    M[:, idx, idx] = h.to(torch.complex64)
    M[:, :n, n] = phi
    M[:, n, :n] = phi.conj()
    M[:, n, n] = c
    return M


def eig_map(values, p):
    """admm_net.py:310-334 (vectorised over the eigenvalue index; same arithmetic per element)."""
    thr = torch.sigmoid(p["threshold"])
    base = _sp(values - thr)
    x = values.abs().unsqueeze(-1)                              # [B,d,1]
    hid = F.relu(F.linear(x, p["V1"], p["c1"]))                 # [B,d,16]
    scale = torch.sigmoid(F.linear(hid, p["V2"], p["c2"])).squeeze(-1)
    return base * scale


def g_update(phi, h, Z, p, eigh=None):
    """admm_net.py:237-354. Returns (G, A_h, values, values_corrected)."""
    c0 = lambda_inv_scalar(p["lambda_param"])
    blk = block_matrix(phi, h, c0)
    rho = _sp(p["rho"])
    A = blk - (1.0 / (rho + EPS)) * Z
    Ah = 0.5 * (A + A.transpose(1, 2).conj())
    if eigh is None:
        vals, vecs = torch.linalg.eigh(Ah)
    else:
        vals, vecs = eigh(Ah)
    vc = eig_map(vals, p)
    G = torch.bmm(vecs, torch.bmm(torch.diag_embed(vc.to(torch.complex64)), vecs.transpose(1, 2).conj()))
    G = 0.5 * (G + G.transpose(1, 2).conj())
    return G, Ah, vals, vc


def z_update(phi, h, G, Z, k, p, mean_r=None):
    """admm_net.py:388-474. mean_r overrides the batch mean (used to emulate other norm scopes)."""
    c1 = lambda_inv_scalar(p["lambda_param"])
    C = block_matrix(phi, h, c1)
    R = G - C
    rho = _sp(p["rho"])
    r = torch.norm(R, dim=[1, 2], p="fro")
    B = r.shape[0]
    k_norm = torch.tensor(k / 10.0, device=r.device).repeat(B)
    rho_norm = torch.full((B,), rho.item(), device=r.device)
    m = r.mean() if mean_r is None else mean_r
    res_norm = r / (m + EPS)
    feat = torch.stack([k_norm, rho_norm, res_norm], dim=1)
    sf = torch.sigmoid(F.linear(F.relu(F.linear(feat, p["U1"], p["d1"])), p["U2"], p["d2"]))
    sf = 0.5 + 1.5 * sf
    alpha = rho * sf.squeeze(1)
    return Z + alpha.unsqueeze(-1).unsqueeze(-1) * R, r, alpha


def layer_params(sd, k, device=None):
    """Pick layer k's tensors out of a reference-format state_dict (keys: SURVEY.md §8a)."""
    g = lambda name: sd[name].detach().float().to(device) if device is not None else sd[name].detach().float()
    return dict(
        phi=dict(rho=g(f"phiLayers.{k}.rho")),
        h=dict(rho=g(f"hLayers.{k}.rho"), projection_weight=g(f"hLayers.{k}.projection_weight"),
               W1=g(f"hLayers.{k}.correction_net.0.weight"), b1=g(f"hLayers.{k}.correction_net.0.bias"),
               W2=g(f"hLayers.{k}.correction_net.2.weight"), b2=g(f"hLayers.{k}.correction_net.2.bias")),
        g=dict(lambda_param=g(f"gLayers.{k}.lambda_param"), rho=g(f"gLayers.{k}.rho"),
               threshold=g(f"gLayers.{k}.threshold"),
               V1=g(f"gLayers.{k}.value_net.0.weight"), c1=g(f"gLayers.{k}.value_net.0.bias"),
               V2=g(f"gLayers.{k}.value_net.2.weight"), c2=g(f"gLayers.{k}.value_net.2.bias")),
        z=dict(rho=g(f"zLayers.{k}.rho"), lambda_param=g(f"zLayers.{k}.lambda_param"),
               U1=g(f"zLayers.{k}.residual_scale_net.0.weight"), d1=g(f"zLayers.{k}.residual_scale_net.0.bias"),
               U2=g(f"zLayers.{k}.residual_scale_net.2.weight"), d2=g(f"zLayers.{k}.residual_scale_net.2.bias")),
    )


@torch.no_grad()
def forward(sd, y, b, sigma, M, N, num_layers, eigh=None, taps=None, chunk=None, means=None):
    """PhiEstADMMNet.forward (admm_net.py:742-764).  y,b complex64 [B,n]; sigma fp32 [B] or [B,1].

    chunk: if given, the batch is processed in independent chunks of that many signals
    (norm_scope='chunk' of the CUDA path: the ZLayer batch mean is taken per chunk).
    means: optional per-layer values that replace the batch mean of admm_net.py:459 (used to check a SUBSET of a
    large batch: the subset's signals then see the statistic of the batch they were computed in)."""
    if chunk is not None and y.shape[0] > chunk:
        outs = [forward(sd, y[i:i + chunk], b[i:i + chunk], sigma[i:i + chunk], M, N, num_layers, eigh)
                for i in range(0, y.shape[0], chunk)]
        return torch.cat(outs, 0)
    n = M * N
    B = y.shape[0]
    G = torch.zeros(B, n + 1, n + 1, device=y.device)
    Z = torch.zeros(B, n + 1, n + 1, device=y.device)
    phi = None
    for k in range(num_layers):
        p = layer_params(sd, k, y.device)
        phi = phi_update(y, b, G, Z, p["phi"]["rho"])
        h = h_update(G, Z, sigma, n, p["h"])
        G, Ah, vals, vc = g_update(phi, h, Z, p["g"], eigh)
        mr = None if means is None or k >= len(means) else torch.tensor(float(means[k]), dtype=torch.float32, device=y.device)
        Z, r, alpha = z_update(phi, h, G, Z, k, p["z"], mean_r=mr)
        if taps is not None:
            taps.append(dict(phi=phi.clone(), h=h.clone(), A=Ah, vals=vals, vc=vc, G=G.clone(), r=r, alpha=alpha,
                             Z=Z.clone()))
    return phi
