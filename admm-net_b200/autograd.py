"""Differentiable (training) path of the unrolled net — SURVEY.md §8f rank 1, first step.

The reference trains PhiEstADMMNet through PyTorch autograd (trainPhi.py:148-179); 82 % of a forward is
`torch.linalg.eigh`, and gradients only flow through the eigenVALUES because the eigenvectors are detached
(admm_net.py:303-306).  Here the eigen-decomposition is the hand-written CUDA solver (`admmnet_eigh_batched`, the
same kernels as the inference path) wrapped in an autograd Function whose backward is the eigenvalue term of the
Hermitian eigh derivative, gA = U diag(g_lambda) U^H; the small element-wise / MLP ops of the four layers
(admm_net.py:79-105, 134-194, 262-354, 388-474) stay ordinary differentiable torch ops on the GPU.  The inference
fast path (fused kernels, no autograd) is untouched.
"""
import ctypes as C
import math

import torch
import torch.nn.functional as F

from . import _capi

EPS = 1e-8


_STATUS = {}          # device -> int32[1]: sticky OR of the eigen-solver status words of every BatchedEigh call


def eigh_status(device=None, reset=True):
    """Status accumulated by BatchedEigh since the last reset (one device->host read; 0 = every solve converged).
    The training step checks it once per step instead of synchronising inside every layer."""
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    st = _STATUS.get(str(dev))
    if st is None:
        return 0
    v = int(st.item())
    if reset and v:
        st.zero_()
    return v


class BatchedEigh(torch.autograd.Function):
    """(values fp32 [B,d], vectors c64 [B,d,d]) of Hermitian c64 [B,d,d]; vectors carry no gradient.
    No host synchronisation (CUDA-graph capturable): convergence is reported through eigh_status()."""

    @staticmethod
    def forward(ctx, A):
        _capi.require_cuda()
        if not A.is_cuda or A.dtype != torch.complex64:
            raise _capi.AdmmnetError("BatchedEigh needs a complex64 CUDA tensor (no CPU fallback)")
        L = _capi.lib()
        A = A.contiguous()
        B, d, _ = A.shape
        nb = C.c_size_t()
        _capi.check(L.admmnet_eigh_workspace_bytes(B, d, 0, C.byref(nb)))
        ws = torch.empty(nb.value, dtype=torch.uint8, device=A.device)
        vals = torch.empty(B, d, dtype=torch.float32, device=A.device)
        vecs = torch.empty(B, d, d, dtype=torch.complex64, device=A.device)
        key = str(A.device)
        if key not in _STATUS:
            _STATUS[key] = torch.zeros(1, dtype=torch.int32, device=A.device)
        status = _STATUS[key]                      # the kernels only ever OR bits into it
        with torch.cuda.device(A.device):
            _capi.check(L.admmnet_eigh_batched(A.data_ptr(), B, d, vals.data_ptr(), vecs.data_ptr(), None, None,
                                               ws.data_ptr(), nb.value, 0,
                                               torch.cuda.current_stream(A.device).cuda_stream, status.data_ptr()))
        ctx.save_for_backward(vecs)
        ctx.mark_non_differentiable(vecs)
        return vals, vecs

    @staticmethod
    def backward(ctx, g_vals, g_vecs):
        (U,) = ctx.saved_tensors
        return (U * g_vals.to(U.dtype).unsqueeze(1)) @ U.transpose(1, 2).conj()


def _cuda_eigh(A):
    return BatchedEigh.apply(A)


def _block(phi, h, c):
    """[[diag(h), phi],[phi^H, c]]  (admm_net.py:273-284 / 428-439)."""
    B = phi.shape[0]
    H = torch.diag_embed(h)
    # c: the reference's `.item()` scalar, kept as a detached 0-dim device tensor (same fp32 value, no gradient, no
    # host synchronisation)
    corner = (c if torch.is_tensor(c) else torch.tensor(c, device=h.device)).to(h.dtype).reshape(1, 1, 1).expand(B, 1, 1)
    top = torch.cat([H, phi.unsqueeze(-1)], dim=2)
    bottom = torch.cat([phi.conj().unsqueeze(1), corner], dim=2)
    return torch.cat([top, bottom], dim=1)


def forward_train(model, y, b, sigma, _eigh=None):
    """PhiEstADMMNet.forward (admm_net.py:742-764) as a differentiable graph.  `_eigh` is a test hook (the CPU
    tests inject torch.linalg.eigh to check the graph against the reference's gradients without a GPU)."""
    eigh = _cuda_eigh if _eigh is None else _eigh
    n = model.M * model.N
    B = y.shape[0]
    dev = y.device
    sigma = sigma.reshape(-1).to(torch.float32)
    G = torch.zeros(B, n + 1, n + 1, device=dev)
    Z = torch.zeros(B, n + 1, n + 1, device=dev)
    phi = None
    for k in range(model.num_layers):
        pl, hl, gl, zl = model.phiLayers[k], model.hLayers[k], model.gLayers[k], model.zLayers[k]
        # ---- PhiLayer
        b_sq = torch.abs(b) ** 2 + EPS
        rho = F.softplus(pl.rho)
        phi = b_sq / (1 + rho * b_sq) * (y / (b + EPS) + rho * G[:, :-1, -1] + Z[:, :-1, -1])
        if k == model.num_layers - 1:
            break                                   # the last layer's H/G/Z never reach phi
        # ---- HLayer
        rho_h = F.softplus(hl.rho)
        t = torch.diagonal(G[:, :n, :n] + Z[:, :n, :n] / (rho_h + EPS), dim1=1, dim2=2).real
        Asig = (2 * math.sqrt(n) * sigma + sigma ** 2).view(-1, 1)
        tc = t + 0.1 * hl.correction_net(t)
        cv = Asig * tc.abs().max(dim=1, keepdim=True)[0] + tc.sum(dim=1, keepdim=True)
        h = tc * torch.clamp(torch.sigmoid(hl.projection_weight) / (cv + EPS), max=1.0)
        # ---- GLayer
        c0 = (1.0 / (F.softplus(gl.lambda_param) ** 2 + EPS)).detach()          # `.item()` in admm_net.py:269-271
        A = _block(phi, h, c0) - (1.0 / (F.softplus(gl.rho) + EPS)) * Z
        A = 0.5 * (A + A.transpose(1, 2).conj())
        vals, vecs = eigh(A)
        vecs = vecs.detach()
        lam = F.softplus(vals - torch.sigmoid(gl.threshold)) * gl.value_net(vals.abs().unsqueeze(-1)).squeeze(-1)
        Gn = vecs @ (lam.to(torch.complex64).unsqueeze(-1) * vecs.transpose(1, 2).conj())
        G = 0.5 * (Gn + Gn.transpose(1, 2).conj())
        # ---- ZLayer
        c1 = (1.0 / (F.softplus(zl.lambda_param) ** 2 + EPS)).detach()          # `.item()` in admm_net.py:424-426
        R = G - _block(phi, h, c1)
        rho_z = F.softplus(zl.rho)
        r = torch.norm(R, dim=[1, 2], p="fro")
        feats = torch.stack([torch.full((B,), k / 10.0, device=dev), rho_z.detach().expand(B),     # `.item()`, :458
                             r / (r.mean() + EPS)], dim=1)
        alpha = rho_z * (0.5 + 1.5 * zl.residual_scale_net(feats)).squeeze(1)
        Z = Z + alpha.unsqueeze(-1).unsqueeze(-1) * R
    return phi


class PhiAlignmentLoss(torch.nn.Module):
    """loss.py:62-98: amplitude MSE + 0.5 * wrapped-phase MSE."""

    def __init__(self, amplitude_weight=1.0, phase_weight=0.5, spectral_weight=0.2, distribution_weight=0.3):
        super().__init__()
        self.amplitude_weight, self.phase_weight = amplitude_weight, phase_weight
        self.spectral_weight, self.distribution_weight = spectral_weight, distribution_weight

    def forward(self, phi_final, phi_true):
        amplitude_loss = F.mse_loss(torch.abs(phi_final), torch.abs(phi_true))
        diff = torch.angle(phi_final) - torch.angle(phi_true)
        diff = (diff + torch.pi) % (2 * torch.pi) - torch.pi
        phase_loss = F.mse_loss(diff, torch.zeros_like(diff))
        total = self.amplitude_weight * amplitude_loss + self.phase_weight * phase_loss
        return total, {"total_loss": total, "amplitude_loss": amplitude_loss, "phase_loss": phase_loss}


def basic_parameter_loss(tau_pred, f_pred, tau_true, f_true, confidences, L_true):
    """loss.py:6-30 without the per-sample Python loop: for a sample with L true targets the MSE of the first L
    (tau, f) pairs plus 0.1 * MSE(confidence[:L], 1); with L = 0 the sum of squared confidences; batch mean."""
    B, Lmax = tau_pred.shape
    L = L_true.to(tau_pred.device).reshape(B, 1)
    mask = (torch.arange(Lmax, device=tau_pred.device).unsqueeze(0) < L).to(tau_pred.dtype)
    cnt = L.clamp(min=1).to(tau_pred.dtype)
    tau_loss = (mask * (tau_pred - tau_true) ** 2).sum(1, keepdim=True) / cnt
    f_loss = (mask * (f_pred - f_true) ** 2).sum(1, keepdim=True) / cnt
    conf_loss = (mask * (confidences - 1.0) ** 2).sum(1, keepdim=True) / cnt
    with_targets = tau_loss + f_loss + 0.1 * conf_loss
    without = (confidences ** 2).sum(1, keepdim=True)
    return torch.where(L > 0, with_targets, without).sum() / B


class BasicANMLoss(torch.nn.Module):
    """loss.py:33-59 (train.py's criterion): parameter loss + lambda_reg * mean ||phi||."""

    def __init__(self, lambda_reg=1e-4):
        super().__init__()
        self.lambda_reg = lambda_reg

    def forward(self, model_outputs, ground_truth):
        param_loss = basic_parameter_loss(model_outputs["tau_est"], model_outputs["f_est"], ground_truth["tau_true"],
                                          ground_truth["f_true"], model_outputs["confidences"], ground_truth["L_true"])
        reg_loss = self.lambda_reg * torch.mean(torch.norm(model_outputs["phi_final"], dim=1))
        total = param_loss + reg_loss
        return total, {"total_loss": total, "param_loss": param_loss, "reg_loss": reg_loss}
