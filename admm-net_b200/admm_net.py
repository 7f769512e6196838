"""Drop-in mirror of the reference's `admm_net` module API for the unrolled forward path.

Same constructor signatures, attribute names and `state_dict` keys as /root/reference/admm_net.py
(PhiLayer 71-105, HLayer 108-205, GLayer 208-354, ZLayer 357-490, PhiEstADMMNet 724-764), so
`load_state_dict(checkpoint['model_state_dict'])` accepts reference checkpoints and optimiser param
groups that match on 'phiLayers'/'hLayers'/'gLayers'/'zLayers' (trainPhi.py:106-111) keep working.
The layer modules only HOLD parameters; the arithmetic of `PhiEstADMMNet.forward` runs in the sm_100a
kernels behind include/admmnet_b200.h (no PyTorch/CPU fallback).

Differences from the reference, all documented in DESIGN.md:
  * in train() mode with grad enabled the forward is the differentiable graph of autograd.py (CUDA eigen-solver +
    torch element-wise ops).  In eval() mode the values come from the fused inference kernels; with grad enabled the
    result still carries a grad_fn (as the reference's does, test/test_time_net.py:94-100): its backward rebuilds the
    differentiable graph on demand (_FusedForward), so `.backward()` works and costs nothing until it is called;
  * `norm_scope` / `chunk` attributes control how the ZLayer batch mean (admm_net.py:459) is scoped:
      'batch' (default) = the whole batch of the call, exactly like the reference;
      'chunk'           = independent chunks of `chunk` signals (throughput mode, no coupling);
  * `ADMMNet`'s learned PeakSearchLayer head (admm_net.py:494-630) runs with eval-mode semantics.
"""
import ctypes as C
import torch
import torch.nn as nn

from . import _capi
from .params import pack_state_dict, param_stride


class _FusedForward(torch.autograd.Function):
    """Fast-path values with the reference's autograd contract: forward runs the fused inference kernels
    (PhiEstADMMNet.forward_device); backward re-runs the layers as the differentiable graph of autograd.forward_train
    on the saved inputs and back-propagates through it (recompute on demand, like activation checkpointing).  The
    gradients are therefore exactly those of the train()-mode graph."""

    @staticmethod
    def forward(ctx, model, y, b, sigma, *params):
        ctx.model = model
        ctx.save_for_backward(y, b, sigma)
        return model.forward_device(y, b, sigma)

    @staticmethod
    def backward(ctx, g):
        from .autograd import forward_train
        model = ctx.model
        ins = [t.detach().requires_grad_(ctx.needs_input_grad[1 + i]) for i, t in enumerate(ctx.saved_tensors)]
        params = list(model.parameters())
        wanted = [t for t in ins if t.requires_grad] + [p for p in params if p.requires_grad]
        with torch.enable_grad():
            out = forward_train(model, *ins)
            got = iter(torch.autograd.grad(out, wanted, g, allow_unused=True))
        g_ins = [next(got) if t.requires_grad else None for t in ins]
        g_par = [next(got) if p.requires_grad else None for p in params]
        return (None, *g_ins, *g_par)


class PhiLayer(nn.Module):
    """Parameter holder for admm_net.py:71-105."""

    def __init__(self, epsilon=1e-8):
        super().__init__()
        self.rho = nn.Parameter(torch.tensor(1.0))
        self.epsilon = epsilon


class HLayer(nn.Module):
    """Parameter holder for admm_net.py:108-205."""

    def __init__(self, M, N, epsilon=1e-8):
        super().__init__()
        self.M, self.N = M, N
        self.dim = M * N
        self.epsilon = epsilon
        self.rho = nn.Parameter(torch.tensor(1.0))
        self.projection_weight = nn.Parameter(torch.tensor(1.0))
        self.correction_net = nn.Sequential(nn.Linear(self.dim, 64), nn.ReLU(), nn.Linear(64, self.dim), nn.Tanh())


class GLayer(nn.Module):
    """Parameter holder for admm_net.py:208-354."""

    def __init__(self, M, N, epsilon=1e-8, use_learnable_threshold=True):
        super().__init__()
        self.M, self.N = M, N
        self.dim = M * N + 1
        self.epsilon = epsilon
        self.lambda_param = nn.Parameter(torch.tensor(0.1))
        self.rho = nn.Parameter(torch.tensor(1.0))
        if use_learnable_threshold:
            self.threshold = nn.Parameter(torch.tensor(0.0))
        else:
            self.threshold = torch.tensor(0.0)
        self.value_net = nn.Sequential(nn.Linear(1, 16), nn.ReLU(), nn.Linear(16, 1), nn.Sigmoid())


class ZLayer(nn.Module):
    """Parameter holder for admm_net.py:357-490 (step_adjust_net exists in the state_dict, is never used)."""

    def __init__(self, M, N, epsilon=1e-8):
        super().__init__()
        self.M, self.N = M, N
        self.dim_h = M * N
        self.dim_z = M * N + 1
        self.epsilon = epsilon
        self.rho = nn.Parameter(torch.tensor(1.0))
        self.lambda_param = nn.Parameter(torch.tensor(1.0))
        self.residual_scale_net = nn.Sequential(nn.Linear(3, 32), nn.ReLU(), nn.Linear(32, 1), nn.Sigmoid())
        self.step_adjust_net = nn.Sequential(nn.Linear(3, 8), nn.ReLU(), nn.Linear(8, 1), nn.Sigmoid())


class _Workspace:
    """Caller-owned device buffers for one (B, chunk, n, K, rcap) configuration."""

    def __init__(self, B, chunk, n, K, rcap, device):
        L = _capi.lib()
        nbytes = C.c_size_t()
        _capi.check(L.admmnet_forward_workspace_bytes(B, chunk, n, K, rcap, C.byref(nbytes)))
        self.key = (B, chunk, n, K, rcap, str(device))
        self.nbytes = nbytes.value
        self.buf = torch.empty(self.nbytes, dtype=torch.uint8, device=device)
        self.B, self.chunk, self.n, self.K, self.rcap = B, chunk, n, K, rcap
        rs, mn, stt = C.c_void_p(), C.c_void_p(), C.c_void_p()
        _capi.check(L.admmnet_ws_scalars(self.ptr, self.nbytes, B, chunk, n, K, rcap, C.byref(rs), C.byref(mn),
                                         C.byref(stt)))
        base = self.buf.data_ptr()
        self.rsum = self.buf[rs.value - base: rs.value - base + 8 * (K + 1)].view(torch.float64)
        self.mean = self.buf[mn.value - base: mn.value - base + 4 * (K + 1)].view(torch.float32)

    @property
    def ptr(self):
        return self.buf.data_ptr()


class PhiEstADMMNet(nn.Module):
    """admm_net.py:724-764.  forward(y, b, sigma) -> phi  (complex64 [B, M*N])."""

    default_chunk = 16384

    def __init__(self, M, N, L=3, num_layers=10):
        super().__init__()
        self.num_layers = num_layers
        self.M, self.N, self.L = M, N, L
        self.phiLayers = nn.ModuleList([PhiLayer() for _ in range(num_layers)])
        self.hLayers = nn.ModuleList([HLayer(M, N) for _ in range(num_layers)])
        self.gLayers = nn.ModuleList([GLayer(M, N) for _ in range(num_layers)])
        self.zLayers = nn.ModuleList([ZLayer(M, N) for _ in range(num_layers)])
        self.norm_scope = "batch"
        self.chunk = self.default_chunk
        self.rcap = 0
        self.check_status = True
        self._packed = None
        self._packed_key = None
        self._ws = None

    # ------------------------------------------------------------------ parameters
    def packed_params(self, device):
        key = (str(device),) + tuple((p.data_ptr(), p._version) for p in self.parameters())
        if self._packed is None or self._packed_key != key:
            self._packed = pack_state_dict(self.state_dict(), self.M * self.N, self.num_layers).to(device)
            self._packed_key = key
        return self._packed

    def workspace(self, B, chunk, device):
        n, K = self.M * self.N, self.num_layers
        key = (B, chunk, n, K, self.rcap, str(device))
        if self._ws is None or self._ws.key != key:
            self._ws = None
            self._ws = _Workspace(B, chunk, n, K, self.rcap, device)
        return self._ws

    # ------------------------------------------------------------------ forward
    def _prep(self, y, b, sigma):
        if y.dim() != 2 or y.shape != b.shape or y.shape[1] != self.M * self.N:
            raise ValueError(f"y and b must be [B, {self.M * self.N}] (got {tuple(y.shape)}, {tuple(b.shape)})")
        B = y.shape[0]
        if sigma.numel() != B:
            raise ValueError("sigma must hold one value per signal ([B] or [B,1])")
        _capi.require_cuda()
        dev = y.device if y.is_cuda else torch.device("cuda", torch.cuda.current_device())
        keep = torch.is_grad_enabled()      # inputs that require grad stay attached (the reference differentiates through them)
        to = lambda t, dt: (t if keep and t.requires_grad else t.detach()).to(device=dev, dtype=dt, non_blocking=True).contiguous()
        return to(y, torch.complex64), to(b, torch.complex64), to(sigma.reshape(-1), torch.float32), dev

    def forward_device(self, y, b, sigma, out=None):
        """Device-resident fast path: y,b complex64 [B,n] and sigma float32 [B] already on the GPU."""
        L = _capi.lib()
        B, n, K = y.shape[0], self.M * self.N, self.num_layers
        dev = y.device
        P = self.packed_params(dev)
        if out is None:
            out = torch.empty(B, n, dtype=torch.complex64, device=dev)
        chunk = min(self.chunk, B)
        # the library keys its lane streams by cudaGetDevice(): make the tensors' device current for the calls
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            if self.norm_scope == "batch":
                ws = self.workspace(B, chunk, dev)
                _capi.check(L.admmnet_forward(y.data_ptr(), b.data_ptr(), sigma.data_ptr(), B, chunk, self.M, self.N,
                                              K, P.data_ptr(), out.data_ptr(), ws.ptr, ws.nbytes, self.rcap, stream))
                self._status(ws, stream, B, chunk)
            elif self.norm_scope == "chunk":
                # one workspace sized for a full chunk serves the ragged tail too (the C ABI only needs
                # ws_bytes >= what (Bc, Bc) asks for)
                ws = self.workspace(chunk, chunk, dev)
                bad = 0
                for off in range(0, B, chunk):
                    Bc = min(chunk, B - off)
                    _capi.check(L.admmnet_forward(y[off:].data_ptr(), b[off:].data_ptr(), sigma[off:].data_ptr(), Bc,
                                                  Bc, self.M, self.N, K, P.data_ptr(), out[off:].data_ptr(), ws.ptr,
                                                  ws.nbytes, self.rcap, stream))
                    bad |= self._status(ws, stream, Bc, Bc, raise_=False)   # every forward resets the status word
                if bad:
                    raise _capi.AdmmnetError(f"eigen-solver reported status {bad} (QL non-convergence / stream overflow)")
            else:
                raise ValueError("norm_scope must be 'batch' or 'chunk' (use sharding.sharded_forward for 'global')")
        return out

    def _status(self, ws, stream, B=None, chunk=None, raise_=True):
        if not self.check_status:
            return 0
        st = C.c_int(0)
        _capi.check(_capi.lib().admmnet_status(ws.ptr, ws.nbytes, ws.B if B is None else B,
                                               ws.chunk if chunk is None else chunk, ws.n, ws.K, ws.rcap, stream,
                                               C.byref(st)))
        if st.value and raise_:
            raise _capi.AdmmnetError(f"eigen-solver reported status {st.value} (QL non-convergence / stream overflow)")
        return st.value

    def forward(self, y, b, sigma):
        needs_grad = torch.is_grad_enabled() and (any(p.requires_grad for p in self.parameters())
                                                  or any(t.requires_grad for t in (y, b, sigma)))
        if needs_grad and self.training:
            # training: differentiable graph around the CUDA eigen-solver (autograd.py)
            from .autograd import forward_train
            src_dev = y.device
            yd, bd, sd, _ = self._prep(y, b, sigma)
            out = forward_train(self, yd, bd, sd)
            return out if src_dev.type == "cuda" else out.to(src_dev)
        src_dev = y.device
        yd, bd, sd, _ = self._prep(y, b, sigma)
        out = self._forward_values(yd, bd, sd, needs_grad)
        return out if src_dev.type == "cuda" else out.to(src_dev)

    def _forward_values(self, yd, bd, sd, needs_grad):
        """Fused kernels; with grad enabled the result carries the lazily rebuilt graph (_FusedForward)."""
        if needs_grad and self.norm_scope == "batch":
            return _FusedForward.apply(self, yd, bd, sd, *self.parameters())
        return self.forward_device(yd, bd, sd)


class PeakSearchLayer(nn.Module):
    """Parameter holder for the learned regression head (admm_net.py:494-554): same sub-module names, shapes and
    construction order as the reference, so `state_dict`s are interchangeable and a seeded init matches."""

    def __init__(self, M, N, L=3, hidden_dim=128, num_heads=4):
        super().__init__()
        if hidden_dim != 128 or num_heads != 4:
            raise ValueError("the B200 head kernel is built for hidden_dim=128, num_heads=4 (the reference's defaults)")
        self.M, self.N, self.L_max = M, N, L
        self.dim = M * N
        self.feature_extractor = nn.Sequential(nn.Linear(2 * self.dim, hidden_dim), nn.ReLU(),
                                               nn.Linear(hidden_dim, hidden_dim), nn.ReLU())
        tau_grid, f_grid = torch.meshgrid(torch.linspace(0, 1, M), torch.linspace(-0.5, 0.5, N), indexing="ij")
        self.position_encoder = nn.Parameter(torch.stack([tau_grid.flatten(), f_grid.flatten()], dim=1),
                                             requires_grad=True)                       # admm_net.py:556-568
        self.position_projection = nn.Linear(2, hidden_dim)
        self.attention = nn.MultiheadAttention(embed_dim=hidden_dim, num_heads=num_heads, batch_first=True, dropout=0.1)
        self.peak_extractor = nn.Sequential(nn.Linear(hidden_dim, hidden_dim // 2), nn.ReLU(),
                                            nn.Linear(hidden_dim // 2, hidden_dim // 4), nn.ReLU(),
                                            nn.Linear(hidden_dim // 4, hidden_dim // 8), nn.ReLU())
        self.tau_regressor = nn.ModuleList([nn.Sequential(nn.Linear(hidden_dim // 8, 32), nn.ReLU(), nn.Linear(32, 1),
                                                          nn.Sigmoid()) for _ in range(L)])
        self.f_regressor = nn.ModuleList([nn.Sequential(nn.Linear(hidden_dim // 8, 32), nn.ReLU(), nn.Linear(32, 1),
                                                        nn.Tanh()) for _ in range(L)])
        self.confidence_net = nn.Sequential(nn.Linear(hidden_dim // 8, 16), nn.ReLU(), nn.Linear(16, 1), nn.Sigmoid())

    def forward(self, phi, b=None):
        """admm_net.py:570-630 as a differentiable torch graph (the TRAINING path of the head, attention dropout
        included in train mode); inference goes through the fused kernel, ADMMNet.head_device."""
        B = phi.shape[0]
        x = self.feature_extractor(torch.cat([phi.real, phi.imag], dim=1))
        pos = self.position_projection(self.position_encoder.unsqueeze(0).repeat(B, 1, 1))
        attended, _ = self.attention(query=x.unsqueeze(1), key=pos, value=pos)
        x_peak = self.peak_extractor(x + attended.squeeze(1))
        taus, fs, confs = [], [], []
        for t in range(self.L_max):
            feat = x_peak + torch.tensor(t / self.L_max, device=phi.device)
            taus.append(self.tau_regressor[t](feat))
            fs.append(self.f_regressor[t](feat))
            confs.append(self.confidence_net(feat))
        return torch.cat(taus, dim=1), torch.cat(fs, dim=1), torch.cat(confs, dim=1)


class ADMMNet(PhiEstADMMNet):
    """admm_net.py:767-816.  forward(y, b, sigma) -> (tau_est, f_est, confidences, phi): the unrolled loop is shared
    with PhiEstADMMNet; in eval mode the PeakSearchLayer head (admm_net.py:570-630) runs in csrc/head_kernels.cu (no
    attention dropout), in train() mode with grad enabled the whole forward is differentiable (forward_differentiable)."""

    def __init__(self, M, N, L=3, num_layers=10):
        super().__init__(M, N, L, num_layers)
        self.peakSearchLayer = PeakSearchLayer(M, N, L)
        self._head = None
        self._head_key = None

    def packed_head(self, device):
        from .params import pack_head
        ps = list(self.peakSearchLayer.parameters())
        key = (str(device),) + tuple((p.data_ptr(), p._version) for p in ps)
        if self._head is None or self._head_key != key:
            self._head = pack_head(self.state_dict(), self.M * self.N, self.L).to(device)
            self._head_key = key
        return self._head

    def head_device(self, phi):
        """phi complex64 [B,n] on the GPU -> (tau, f, conf) float32 [B,L] on the GPU."""
        B, n, L = phi.shape[0], self.M * self.N, self.L
        dev = phi.device
        H = self.packed_head(dev)
        tau, f, conf = (torch.empty(B, L, dtype=torch.float32, device=dev) for _ in range(3))
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            _capi.check(_capi.lib().admmnet_peak_head(phi.contiguous().data_ptr(), B, n, L, H.data_ptr(),
                                                      tau.data_ptr(), f.data_ptr(), conf.data_ptr(), stream))
        return tau, f, conf

    def forward_differentiable(self, y, b, sigma, _eigh=None):
        """train.py:177-201: the unrolled layers as the autograd graph around the CUDA eigen-solver (autograd.py) and
        the head as torch modules.  `_eigh` is the CPU test hook of forward_train."""
        from .autograd import forward_train
        phi = forward_train(self, y, b, sigma, _eigh)
        tau, f, conf = self.peakSearchLayer(phi, b)
        return tau, f, conf, phi

    def forward(self, y, b, sigma):
        needs_grad = torch.is_grad_enabled() and (any(p.requires_grad for p in self.parameters())
                                                  or any(t.requires_grad for t in (y, b, sigma)))
        src_dev = y.device
        yd, bd, sd, _ = self._prep(y, b, sigma)
        if needs_grad and self.training:
            outs = self.forward_differentiable(yd, bd, sd)
            return outs if src_dev.type == "cuda" else tuple(t.to(src_dev) for t in outs)
        if needs_grad:
            # eval mode with grad enabled: fused unrolled layers with the lazily rebuilt graph, head as torch modules
            # (eval semantics: attention dropout off) so that the outputs are differentiable like the reference's
            phi = self._forward_values(yd, bd, sd, True)
            tau, f, conf = self.peakSearchLayer(phi, bd)
        else:
            phi = self.forward_device(yd, bd, sd)
            tau, f, conf = self.head_device(phi)
        outs = (tau, f, conf, phi)
        return outs if src_dev.type == "cuda" else tuple(t.to(src_dev) for t in outs)
