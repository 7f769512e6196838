"""Generate the committed golden vectors by running the REFERENCE ITSELF (build container only).

    python tests/golden/make_golden.py            # needs /root/reference (read-only mount)

Nothing here runs on the GPU box: the outputs (*.npz next to this file) are committed and are what
tests/ compare the oracle restatements (oracle/*.py) and the CUDA path against.

What is run unmodified from /root/reference:
  * admm_net.PhiEstADMMNet.forward                      -> net_*.npz
  * admm.admm_for_us  (cvxpy is not installed: an empty `cvxpy` module is injected so the file
    imports, and the single function admm_for_us_H_cvx_0 — the ECOS call — is replaced by the exact
    projection in oracle/classic_oracle.py)             -> classic.npz
  * utils.peakSearchUtils.{peak_search,alt_peak_search} (matplotlib stubbed; skimage is not
    installed: `skimage.morphology.local_maxima` is provided by oracle/peak_oracle.py)   -> peaks.npz
Inputs: data/data.npz of the reference (main_for_net.py:68-75, data_type==2 branch, np.random.seed(0)
before the noise draw at line 81) and oracle/signals.py batches.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

from oracle import classic_oracle, peak_oracle, signals  # noqa: E402


def import_reference():
    # --- stubs for absent third-party modules
    for name in ["matplotlib", "matplotlib.pyplot", "matplotlib.colors", "mpl_toolkits", "mpl_toolkits.mplot3d",
                 "matplotlib.cm", "matplotlib.ticker"]:
        sys.modules.setdefault(name, types.ModuleType(name))
    sk = types.ModuleType("skimage")
    skm = types.ModuleType("skimage.morphology")
    skm.local_maxima = peak_oracle.local_maxima
    sk.morphology = skm
    sys.modules["skimage"] = sk
    sys.modules["skimage.morphology"] = skm
    sys.modules["cvxpy"] = types.ModuleType("cvxpy")
    # the repository ships drop-in modules with the reference's names (admm_net, admm, utils.*): make sure the
    # REFERENCE's files are the ones imported here.  `utils` needs care: ours is a regular package and would
    # shadow the reference's namespace package whatever the sys.path order.
    upkg = types.ModuleType("utils")
    upkg.__path__ = [os.path.join(REF, "utils")]
    sys.modules["utils"] = upkg
    import admm_net  # noqa
    import admm  # noqa
    import utils.peakSearchUtils as psu  # noqa
    import utils.mathUtils as mu  # noqa
    admm.admm_for_us_H_cvx_0 = classic_oracle.h_update
    for m in (admm_net, admm, psu, mu):
        assert os.path.abspath(m.__file__).startswith(REF), m.__file__
    return admm_net, admm, psu, mu


def sd_to_npz(sd):
    return {k.replace(".", "__"): v.detach().cpu().numpy() for k, v in sd.items()}


def perturb_(model, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if p.ndim == 0:
                p.add_(0.3 * torch.randn((), generator=g))
            else:
                p.mul_(1 + 0.2 * torch.randn(p.shape, generator=g)).add_(0.05 * torch.randn(p.shape, generator=g))


def data_npz_case(mu):
    """main_for_net.py:16-97 with data_type == 2 and np.random.seed(0)."""
    Nb = Nd = 10
    L = 3
    f = np.array([-0.25, 0, 0.14])
    tau = np.array([0.45, 0.25, 0.63])
    C = np.array([-0.5 + 1j, 0.6 - 0.2j, 0.3 + 0.7j])
    S = np.zeros((Nb, L), dtype=complex)
    D = np.zeros((Nd, L), dtype=complex)
    for i in range(L):
        S[:, i] = mu.vander_vec(0, (Nb - 1) * f[i], Nb).reshape(-1)
        D[:, i] = mu.vander_vec(0, (Nd - 1) * tau[i], Nd).reshape(-1)
    Psi = mu.kr(S, np.conj(D)) @ C.reshape(-1, 1)
    df = np.load(os.path.join(REF, "data", "data.npz"))
    sig, e = df["sig"], df["e"]
    b = sig - e
    np.random.seed(0)
    real_y = np.diag(b + e) @ Psi
    w = np.sqrt(1 / 2) * (np.random.randn(Nb * Nd, 1) + 1j * np.random.randn(Nb * Nd, 1))
    w_var = np.linalg.norm(real_y) ** 2 / (10 ** (20 / 10) * Nb * Nd)
    y = (real_y + np.sqrt(w_var) * w).flatten()
    sigma = np.linalg.norm(e / b) + 1
    return y, b, sigma, dict(f=f, tau=tau)


def run_net(admm_net, model, y, b, s):
    """Reference forward + per-layer phi / residual taps through forward hooks."""
    K = model.num_layers
    phis, Zs, Gs = [], [], []
    hooks = []
    for k in range(K):
        hooks.append(model.phiLayers[k].register_forward_hook(lambda m, i, o: phis.append(o.detach().clone())))
        hooks.append(model.gLayers[k].register_forward_hook(lambda m, i, o: Gs.append(o.detach().clone())))
        hooks.append(model.zLayers[k].register_forward_hook(lambda m, i, o: Zs.append(o.detach().clone())))
    with torch.no_grad():
        out = model(y, b, s)
    for h in hooks:
        h.remove()
    n = y.shape[1]
    taps = dict(phi_layers=torch.stack(phis).numpy(),
                g_col=torch.stack([G[:, :n, n] for G in Gs]).numpy(),          # G[:n,n] per layer
                g_diag=torch.stack([torch.diagonal(G, dim1=1, dim2=2).real for G in Gs]).numpy(),
                z_col=torch.stack([Z[:, :n, n] for Z in Zs]).numpy(),
                z_fro=torch.stack([torch.linalg.norm(Z, dim=(1, 2)) for Z in Zs]).numpy())
    return out.numpy(), taps


def main():
    admm_net, admm, psu, mu = import_reference()
    torch.set_num_threads(1)

    # ------------------------------------------------------------------ net forward
    y1, b1, s1, truth1 = data_npz_case(mu)
    yb, bb, sb, _ = signals.generate(6, seed=11)
    y_all = np.concatenate([y1[None].astype(np.complex64), yb])
    b_all = np.concatenate([b1[None].astype(np.complex64), bb])
    s_all = np.concatenate([np.float32([sigma_ := s1]), sb]).astype(np.float32)
    ty, tb, ts = map(torch.from_numpy, (y_all, b_all, s_all))
    for tag, K, seed, pert in [("init_k10", 10, 0, None), ("pert_k10", 10, 1, 7), ("pert_k5", 5, 2, 8)]:
        torch.manual_seed(seed)
        model = admm_net.PhiEstADMMNet(10, 10, 3, K).eval()
        if pert is not None:
            perturb_(model, pert)
        out = {}
        # whole batch (batch-mean coupling across the 7 signals) ...
        phi, taps = run_net(admm_net, model, ty, tb, ts)
        out["phi_batch"] = phi
        for k_, v_ in taps.items():
            out["batch_" + k_] = v_
        # ... and the data.npz signal alone, sigma shaped [1,1] as main_for_net.py:93 does
        phi1, taps1 = run_net(admm_net, model, ty[:1], tb[:1], ts[:1].reshape(1, 1))
        out["phi_single"] = phi1
        np.savez(os.path.join(HERE, f"net_{tag}.npz"), y=y_all, b=b_all, sigma=s_all, K=K, M=10, N=10, **out,
                 **{"sd__" + k: v for k, v in sd_to_npz(model.state_dict()).items()})
        print(tag, "phi max", np.abs(phi).max())
        if tag == "init_k10":
            phi_for_peaks = phi

    # ------------------------------------------------------------------ gradients (trainPhi.py:167-172 with loss.py:62-98)
    import loss as ref_loss
    assert os.path.abspath(ref_loss.__file__).startswith(REF)
    torch.manual_seed(4)
    tm = admm_net.PhiEstADMMNet(10, 10, 3, 4).train()
    perturb_(tm, 10)
    torch.manual_seed(5)
    phi_true = torch.view_as_complex(torch.randn(7, 100, 2) * 0.2).to(torch.complex64)
    crit = ref_loss.PhiAlignmentLoss()
    out_phi = tm(ty, tb, ts)
    total, parts = crit(out_phi, phi_true)
    total.backward()
    grads = {n_.replace(".", "__"): (p_.grad.numpy() if p_.grad is not None else np.zeros(0, np.float32))
             for n_, p_ in tm.named_parameters()}
    np.savez(os.path.join(HERE, "train_grads_k4.npz"), y=y_all, b=b_all, sigma=s_all, K=4, phi_true=phi_true.numpy(),
             loss=float(total), amplitude_loss=float(parts["amplitude_loss"]), phase_loss=float(parts["phase_loss"]),
             phi=out_phi.detach().numpy(), **{"grad__" + k_: v_ for k_, v_ in grads.items()},
             **{"sd__" + k_: v_ for k_, v_ in sd_to_npz(tm.state_dict()).items()})
    print("train loss", float(total), "params with grad", sum(v_.size > 0 for v_ in grads.values()), "of", len(grads))

    # ------------------------------------------------------------------ train.py: ADMMNet + BasicANMLoss gradients
    # (eval mode = attention dropout off, so the numbers are reproducible; grad enabled)
    torch.manual_seed(6)
    fm = admm_net.ADMMNet(10, 10, 3, 3).eval()
    perturb_(fm, 11)
    torch.manual_seed(7)
    tau_t, f_t = torch.rand(7, 3) * 0.8 + 0.1, torch.rand(7, 3) * 0.8 - 0.4
    L_t = torch.tensor([3, 2, 0, 1, 3, 3, 2])
    tau_e, f_e, conf_e, phi_e = fm(ty, tb, ts)
    crit2 = ref_loss.BasicANMLoss()
    tot2, parts2 = crit2({"tau_est": tau_e, "f_est": f_e, "confidences": conf_e, "phi_final": phi_e},
                         {"tau_true": tau_t, "f_true": f_t, "L_true": L_t, "y": ty, "b": tb})
    tot2.backward()
    grads2 = {n_.replace(".", "__"): (p_.grad.numpy() if p_.grad is not None else np.zeros(0, np.float32))
              for n_, p_ in fm.named_parameters()}
    np.savez(os.path.join(HERE, "train_full_grads_k3.npz"), y=y_all, b=b_all, sigma=s_all, K=3, tau_true=tau_t.numpy(),
             f_true=f_t.numpy(), L_true=L_t.numpy(), loss=float(tot2), param_loss=float(parts2["param_loss"]),
             reg_loss=float(parts2["reg_loss"]), tau=tau_e.detach().numpy(), f=f_e.detach().numpy(),
             conf=conf_e.detach().numpy(), **{"grad__" + k_: v_ for k_, v_ in grads2.items()},
             **{"sd__" + k_: v_ for k_, v_ in sd_to_npz(fm.state_dict()).items()})
    print("ADMMNet train loss", float(tot2), "params with grad", sum(v_.size > 0 for v_ in grads2.values()), "of", len(grads2))

    # ------------------------------------------------------------------ full ADMMNet (unrolled loop + PeakSearchLayer head)
    torch.manual_seed(3)
    full = admm_net.ADMMNet(10, 10, 3, 3).eval()
    perturb_(full, 9)
    with torch.no_grad():
        tau_e, f_e, conf_e, phi_e = full(ty, tb, ts)
        tau_h, f_h, conf_h = full.peakSearchLayer(phi_e)          # head alone on the reference's own phi
    np.savez(os.path.join(HERE, "admmnet_full_k3.npz"), y=y_all, b=b_all, sigma=s_all, K=3, M=10, N=10, L=3,
             tau=tau_e.numpy(), f=f_e.numpy(), conf=conf_e.numpy(), phi=phi_e.numpy(),
             **{"sd__" + k: v for k, v in sd_to_npz(full.state_dict()).items()})
    print("ADMMNet tau", tau_e[0].numpy(), "f", f_e[0].numpy(), "conf", conf_e[0].numpy())

    # ------------------------------------------------------------------ classical ADMM
    import contextlib
    import io
    ys, bs, ss, _ = signals.generate(3, seed=21)
    cy = np.concatenate([y1[None], ys.astype(complex)])
    cb = np.concatenate([b1[None], bs.astype(complex)])
    cs = np.concatenate([[s1], ss.astype(float)])
    res = []
    for opts, umi, mi in [(dict(eta_abs=1e-7, eta_rel=1e-7, max_iter=100), True, 5),   # main.py:88-95
                          (None, True, 5), (dict(max_iter=3), True, 5), (dict(rho=0.5, max_iter=50), False, 5),
                          (dict(rho=2.0), True, 7)]:
        for i in range(len(cy)):
            with contextlib.redirect_stdout(io.StringIO()):
                phi, it = admm.admm_for_us(cy[i], cb[i], 10, 10, 1.0, cs[i], opts, umi, mi)
            res.append((i, opts, umi, mi, phi, it))
    np.savez(os.path.join(HERE, "classic.npz"), y=cy, b=cb, sigma=cs,
             case_sig=np.array([r[0] for r in res]),
             case_rho=np.array([(r[1] or {}).get("rho", 1.0) for r in res]),
             case_max_iter=np.array([(r[1] or {}).get("max_iter", 500) for r in res]),
             case_use_min_iter=np.array([r[2] for r in res]), case_min_iter=np.array([r[3] for r in res]),
             phi=np.array([r[4] for r in res]), iters=np.array([r[5] for r in res]))
    print("classic iters", [r[5] for r in res])
    phi_classic = res[0][4]

    # ------------------------------------------------------------------ peak search
    pk = {}
    cases = [("net0", phi_for_peaks[0].astype(np.complex64), dict(xstep=0.01, ystep=0.01, iter=3)),     # main_for_net.py:112-117
             ("net3", phi_for_peaks[3].astype(np.complex64), dict(xstep=0.01, ystep=0.01, iter=3)),
             ("classic0", phi_classic, dict(xstep=0.01, ystep=0.01, iter=3)),                            # main.py:102-112
             ("classic0_default", phi_classic, None),
             ("net1_coarse", phi_for_peaks[1].astype(np.complex64), dict(xstep=0.04, ystep=0.02, iter=2)),
             ("net2_window", phi_for_peaks[2].astype(np.complex64),
              dict(xmin=0.2, xmax=0.8, ymin=-0.3, ymax=0.4, xstep=0.02, ystep=0.02, iter=2, reducefactor=0.2)),
             ("zero", np.zeros(100, dtype=np.complex64), dict(xstep=0.05, ystep=0.05, iter=1)),
             ("empty", phi_classic, dict(xmin=0.5, xmax=0.5))]
    for name, phi, opts in cases:
        r = psu.alt_peak_search({"phi": phi, "xbase": 10, "ybase": 10}, opts)
        pk[f"{name}__phi"] = phi
        pk[f"{name}__opts"] = np.array(repr(opts))
        pk[f"{name}__peaks"] = r
        print(name, r.shape)
    # one full coarse surface from the reference's literal double loop
    ax = np.arange(0, 1 - 0.01, 0.01)
    ay = np.arange(-0.5, 0.5 - 0.01, 0.01)
    AX, AY = np.meshgrid(ax, ay)
    pk["surface_net0"] = psu.peak_search(phi_for_peaks[0].astype(np.complex64), AX, 10, AY, 10)
    # the reference's own plateau KAT (peakSearchUtils.py:427-436): expected mask by hand
    np.savez(os.path.join(HERE, "peaks.npz"), **pk)


if __name__ == "__main__":
    main()
