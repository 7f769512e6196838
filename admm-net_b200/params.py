"""Pack a reference-format state_dict (admm_net.py:727-739 keys) into the per-layer float
records the kernels read (csrc/common.cuh, enum ParamOff).  The scalar pre-computations use the same
fp32 torch ops the reference applies on every forward (softplus / sigmoid / .item()), so the device
sees bit-identical scalars."""
import torch
import torch.nn.functional as F

EPS = 1e-8
# offsets: keep in sync with csrc/common.cuh
P_RHO_PHI, P_RHO_H_EPS, P_SIG_PW, P_C0, P_INV_RHO_G, P_THR, P_C1Z, P_RHO_Z, P_KNORM, P_ZD2, P_VC2 = range(11)
P_V1, P_VC1, P_V2, P_ZU1, P_ZD1, P_ZU2, P_HB1, P_HW1T = 16, 32, 48, 64, 160, 192, 224, 288


def param_stride(n):
    return (288 + 129 * n + 3) & ~3


def _lam_inv(p):
    lam = F.softplus(p)
    return torch.tensor((1.0 / (lam ** 2 + EPS)).item(), dtype=torch.float32)      # admm_net.py:269-271 / 424-426


@torch.no_grad()
def pack_state_dict(sd, n, num_layers):
    """-> float32 CPU tensor [K, param_stride(n)]"""
    st = param_stride(n)
    out = torch.zeros(num_layers, st, dtype=torch.float32)
    g = lambda name: sd[name].detach().to("cpu", torch.float32)
    for k in range(num_layers):
        r = out[k]
        r[P_RHO_PHI] = F.softplus(g(f"phiLayers.{k}.rho"))
        r[P_RHO_H_EPS] = F.softplus(g(f"hLayers.{k}.rho")) + EPS
        r[P_SIG_PW] = torch.sigmoid(g(f"hLayers.{k}.projection_weight"))
        r[P_C0] = _lam_inv(g(f"gLayers.{k}.lambda_param"))
        r[P_INV_RHO_G] = 1.0 / (F.softplus(g(f"gLayers.{k}.rho")) + EPS)
        r[P_THR] = torch.sigmoid(g(f"gLayers.{k}.threshold"))
        r[P_C1Z] = _lam_inv(g(f"zLayers.{k}.lambda_param"))
        r[P_RHO_Z] = F.softplus(g(f"zLayers.{k}.rho"))
        r[P_KNORM] = torch.tensor(k / 10.0)                                       # admm_net.py:457
        r[P_ZD2] = g(f"zLayers.{k}.residual_scale_net.2.bias").reshape(())
        r[P_VC2] = g(f"gLayers.{k}.value_net.2.bias").reshape(())
        r[P_V1:P_V1 + 16] = g(f"gLayers.{k}.value_net.0.weight").reshape(16)
        r[P_VC1:P_VC1 + 16] = g(f"gLayers.{k}.value_net.0.bias")
        r[P_V2:P_V2 + 16] = g(f"gLayers.{k}.value_net.2.weight").reshape(16)
        r[P_ZU1:P_ZU1 + 96] = g(f"zLayers.{k}.residual_scale_net.0.weight").reshape(96)      # [32][3]
        r[P_ZD1:P_ZD1 + 32] = g(f"zLayers.{k}.residual_scale_net.0.bias")
        r[P_ZU2:P_ZU2 + 32] = g(f"zLayers.{k}.residual_scale_net.2.weight").reshape(32)
        r[P_HB1:P_HB1 + 64] = g(f"hLayers.{k}.correction_net.0.bias")
        W1 = g(f"hLayers.{k}.correction_net.0.weight")                            # [64, n]
        W2 = g(f"hLayers.{k}.correction_net.2.weight")                            # [n, 64]
        assert W1.shape == (64, n) and W2.shape == (n, 64), "state_dict does not match M*N"
        r[P_HW1T:P_HW1T + 64 * n] = W1.t().contiguous().reshape(-1)               # [n][64]
        r[P_HW1T + 64 * n:P_HW1T + 128 * n] = W2.t().contiguous().reshape(-1)     # [64][n]
        r[P_HW1T + 128 * n:P_HW1T + 129 * n] = g(f"hLayers.{k}.correction_net.2.bias")
    return out


HD = 128


def head_param_count(n, L):
    base = (2 * n * HD + HD) + 3 * (HD * HD + HD) + 2 * n * HD + (HD * 64 + 64) + (64 * 32 + 32) + (32 * 16 + 16)
    return base + L * 2 * (16 * 32 + 32 + 32 + 1) + (16 * 16 + 16 + 16 + 1)


@torch.no_grad()
def pack_head(sd, n, L, prefix="peakSearchLayer."):
    """Pack PeakSearchLayer's parameters (admm_net.py:496-554) for csrc/head_kernels.cu.  Weights are stored
    transposed ([in][out]); the attention keys/values only depend on the parameters (key = value =
    position_projection(position_encoder), admm_net.py:595-603), so K and V are projected here once with the same
    fp32 torch ops nn.MultiheadAttention applies."""
    g = lambda name: sd[prefix + name].detach().to("cpu", torch.float32)
    parts = []
    T = lambda w: w.t().contiguous().reshape(-1)
    parts += [T(g("feature_extractor.0.weight")), g("feature_extractor.0.bias")]
    parts += [T(g("feature_extractor.2.weight")), g("feature_extractor.2.bias")]
    Wi, bi = g("attention.in_proj_weight"), g("attention.in_proj_bias")
    Wq, Wk, Wv = Wi[:HD], Wi[HD:2 * HD], Wi[2 * HD:]
    bq, bk, bv = bi[:HD], bi[HD:2 * HD], bi[2 * HD:]
    parts += [T(Wq), bq]
    pos = F.linear(g("position_encoder"), g("position_projection.weight"), g("position_projection.bias"))   # [n, HD]
    K = F.linear(pos, Wk, bk)                                      # [n, HD]
    V = F.linear(pos, Wv, bv)
    parts += [K.t().contiguous().reshape(-1), V.contiguous().reshape(-1)]
    parts += [T(g("attention.out_proj.weight")), g("attention.out_proj.bias")]
    for i in (0, 2, 4):
        parts += [T(g(f"peak_extractor.{i}.weight")), g(f"peak_extractor.{i}.bias")]
    for t in range(L):
        for kind in ("tau_regressor", "f_regressor"):
            parts += [T(g(f"{kind}.{t}.0.weight")), g(f"{kind}.{t}.0.bias"), g(f"{kind}.{t}.2.weight").reshape(-1),
                      g(f"{kind}.{t}.2.bias").reshape(-1)]
    parts += [T(g("confidence_net.0.weight")), g("confidence_net.0.bias"), g("confidence_net.2.weight").reshape(-1),
              g("confidence_net.2.bias").reshape(-1)]
    out = torch.cat([p.reshape(-1) for p in parts])
    assert out.numel() == head_param_count(n, L), (out.numel(), head_param_count(n, L))
    return out
