"""Stage-by-stage diagnostics of the CUDA path (run on the GPU box; prints, no asserts)."""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import admmnet_b200  # noqa: E402
from admmnet_b200 import _capi  # noqa: E402
from oracle import classic_oracle, net_oracle, peak_oracle, signals  # noqa: E402
from tests.helpers import load_net_case, parse_opts, rel_err  # noqa: E402

dev = torch.device("cuda", 0)
L = _capi.lib()


def eigh_gpu(A, params=None, want_vecs=True, want_fn=False):
    B, d, _ = A.shape
    nb = C.c_size_t()
    _capi.check(L.admmnet_eigh_workspace_bytes(B, d, 0, C.byref(nb)))
    ws = torch.empty(nb.value, dtype=torch.uint8, device=dev)
    ev = torch.empty(B, d, dtype=torch.float32, device=dev)
    U = torch.empty(B, d, d, dtype=torch.complex64, device=dev) if want_vecs else None
    G = torch.empty(B, d * (d + 1) // 2, dtype=torch.complex64, device=dev) if want_fn else None
    st = torch.zeros(1, dtype=torch.int32, device=dev)
    Ad = A.to(dev).contiguous()
    _capi.check(L.admmnet_eigh_batched(Ad.data_ptr(), B, d, ev.data_ptr(), U.data_ptr() if U is not None else None,
                                       G.data_ptr() if G is not None else None,
                                       params.data_ptr() if params is not None else None, ws.data_ptr(), nb.value, 0,
                                       torch.cuda.current_stream().cuda_stream, st.data_ptr()))
    torch.cuda.synchronize()
    return ev.cpu(), (U.cpu() if U is not None else None), (G.cpu() if G is not None else None), int(st.item())


def main():
    torch.manual_seed(0)
    print("device", torch.cuda.get_device_name(0))
    for d in (5, 16, 33, 101, 128):
        B = 6
        X = torch.randn(B, d, d, dtype=torch.complex64)
        A = 0.5 * (X + X.transpose(1, 2).conj())
        ev, U, _, st = eigh_gpu(A)
        res = (A @ U - U * ev.unsqueeze(1).to(torch.complex64)).abs().amax().item()
        orth = (U.transpose(1, 2).conj() @ U - torch.eye(d)).abs().amax().item()
        w = torch.linalg.eigvalsh(A.to(torch.complex128)).float()
        lerr = (ev.sort(dim=1)[0] - w).abs().amax().item()
        print(f"eigh d={d}: status {st} resid {res:.2e} orth {orth:.2e} lam err {lerr:.2e} |A| {A.abs().amax():.2f}")
    # forward
    for tag in ("init_k10", "pert_k10", "pert_k5"):
        z, sd = load_net_case(tag)
        K = int(z["K"])
        net = admmnet_b200.PhiEstADMMNet(10, 10, 3, K)
        net.load_state_dict(sd)
        y, b, s = (torch.from_numpy(z[k]) for k in ("y", "b", "sigma"))
        phi = net(y, b, s).numpy()
        print(tag, "batch rel err vs reference golden", rel_err(phi, z["phi_batch"]))
        phi1 = net(y[:1], b[:1], s[:1].reshape(1, 1)).numpy()
        print(tag, "single rel err", rel_err(phi1, z["phi_single"]))
        for kk in (1, 2, 3):
            if kk > K:
                continue
            netk = admmnet_b200.PhiEstADMMNet(10, 10, 3, kk)
            netk.load_state_dict({k_: v for k_, v in sd.items() if int(k_.split(".")[1]) < kk})
            pk = netk(y, b, s).numpy()
            print(tag, f"K={kk} prefix rel err", rel_err(pk, z["batch_phi_layers"][kk - 1]).max())
    # bigger batch vs oracle
    torch.manual_seed(0)
    net = admmnet_b200.PhiEstADMMNet(10, 10, 3, 10)
    y, b, s, _ = signals.generate(64, seed=3)
    yt, bt, stt = (torch.from_numpy(a) for a in (y, b, s))
    t0 = time.time()
    phi = net(yt, bt, stt)
    torch.cuda.synchronize()
    t1 = time.time()
    ref = net_oracle.forward(net.state_dict(), yt, bt, stt, 10, 10, 10)
    print("B=64 vs oracle rel err max", rel_err(phi.numpy(), ref.numpy()).max(), "gpu wall", t1 - t0)
    # classic
    z = np.load(os.path.join(ROOT, "tests/golden/classic.npz"))
    for c in (0, 8, 12, 16):
        i = int(z["case_sig"][c])
        opts = dict(rho=float(z["case_rho"][c]), max_iter=int(z["case_max_iter"][c]))
        phi, it = admmnet_b200.admm_for_us(z["y"][i], z["b"][i], 10, 10, 1.0, float(z["sigma"][i]), opts,
                                           bool(z["case_use_min_iter"][c]), int(z["case_min_iter"][c]))
        print("classic case", c, "iters", it, int(z["iters"][c]), "rel err", rel_err(phi, z["phi"][c]))
    # peaks
    z = np.load(os.path.join(ROOT, "tests/golden/peaks.npz"))
    names = sorted({k.split("__")[0] for k in z.files if "__" in k})
    for name in names:
        opts = parse_opts(z[f"{name}__opts"])
        got = admmnet_b200.alt_peak_search({"phi": z[f"{name}__phi"], "xbase": 10, "ybase": 10}, opts)
        exp = z[f"{name}__peaks"]
        same = got.shape == exp.shape and np.array_equal(got[:, :2], exp[:, :2])
        herr = np.abs(got[:, 2] - exp[:, 2]).max() / max(exp[:, 2].max(), 1e-300) if same and len(exp) else 0.0
        print("peaks", name, got.shape, exp.shape, "positions identical", same, "height rel err", herr)
        if not same and got.shape == exp.shape:
            bad = np.where((got[:, :2] != exp[:, :2]).any(1))[0]
            print("   first mismatches", bad[:5], got[bad[:3]], exp[bad[:3]])
    ax = np.arange(0, 1 - 0.01, 0.01)
    ay = np.arange(-0.5, 0.5 - 0.01, 0.01)
    AX, AY = np.meshgrid(ax, ay)
    surf = admmnet_b200.peak_search(z["net0__phi"], AX, 10, AY, 10)
    print("surface rel err", np.abs(surf - z["surface_net0"]).max() / z["surface_net0"].max())
    r = admmnet_b200.alt_peak_search_batched(z["net0__phi"][None], 10, 10, dict(xstep=0.01, ystep=0.01, iter=3), topl=3,
                                             return_surface=True)
    print("batched surface rel err", np.abs(r["surface"][0].cpu().numpy() - z["surface_net0"]).max() / z["surface_net0"].max())
    print("top3", r["top"][0].cpu().numpy(), "\nexp", peak_oracle.top_l(z["net0__peaks"], 3))


if __name__ == "__main__":
    main()
