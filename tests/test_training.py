"""Training path: the differentiable graph and its gradients against the reference's own
autograd (tests/golden/train_grads_k4.npz, written by tests/golden/make_golden.py from trainPhi.py's loss/backward),
and the flat gradient all-reduce on two gloo ranks."""
import json
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden", "train_grads_k4.npz")


def _load():
    z = np.load(GOLD)
    sd = {k[4:].replace("__", "."): torch.from_numpy(z[k]) for k in z.files if k.startswith("sd__")}
    grads = {k[6:].replace("__", "."): z[k] for k in z.files if k.startswith("grad__")}
    return z, sd, grads


def _model(sd, K, device="cpu"):
    from admmnet_b200.admm_net import PhiEstADMMNet
    m = PhiEstADMMNet(10, 10, 3, K)
    m.load_state_dict(sd)
    return m.to(device).train()


def _compare_grads(model, grads, tol):
    """per-parameter max-norm error; parameters whose reference gradient is below 1e-4 of the largest one (the
    H-layer MLPs: 1e-7..1e-9 against 1e-2) are measured against that floor, they are fp32 noise in the reference."""
    worst = 0.0
    n_live = 0
    floor = 1e-4 * max(np.abs(v).max() for v in grads.values() if v.size)
    for name, p in model.named_parameters():
        ref = grads[name]
        if ref.size == 0:                      # the reference leaves .grad None: dead parameter
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, name
            continue
        n_live += 1
        assert p.grad is not None, name
        g = p.grad.detach().cpu().numpy()
        scale = max(np.abs(ref).max(), floor)
        err = np.abs(g - ref).max() / scale
        worst = max(worst, err)
        assert err < tol, (name, err)
    assert n_live == 55
    return worst


def _cpu_eigh(A):
    return torch.linalg.eigh(A)


def test_training_graph_gradients_match_reference_cpu():
    """graph wiring only: torch.linalg.eigh injected through the test hook (the product default is the CUDA solver)."""
    from admmnet_b200.autograd import PhiAlignmentLoss, forward_train
    z, sd, grads = _load()
    model = _model(sd, int(z["K"]))
    y, b, s, pt = (torch.from_numpy(z[k]) for k in ("y", "b", "sigma", "phi_true"))
    phi = forward_train(model, y, b, s, _eigh=_cpu_eigh)
    loss, parts = PhiAlignmentLoss()(phi, pt)
    loss.backward()
    assert np.abs(phi.detach().numpy() - z["phi"]).max() < 1e-4 * np.abs(z["phi"]).max()
    assert abs(float(loss.detach()) - float(z["loss"])) < 1e-4 * float(z["loss"])
    assert abs(float(parts["phase_loss"].detach()) - float(z["phase_loss"])) < 1e-4 * float(z["phase_loss"])
    _compare_grads(model, grads, 2e-3)


def test_eval_mode_lazy_backward_wiring_cpu(monkeypatch):
    """admm_net._FusedForward (eval mode with grad enabled): the values come from the fast path, backward re-runs the
    differentiable graph.  CPU wiring test: the fused kernels are stood in for by the oracle's forward and the CUDA
    eigen-solver by torch.linalg.eigh, so only the autograd plumbing is exercised - grad_fn on the result, gradients
    equal to the train-mode graph's (the reference's autograd, golden vectors), input gradients, nothing saved under
    no_grad."""
    from admmnet_b200 import admm_net, autograd
    from admmnet_b200.autograd import PhiAlignmentLoss
    from oracle import net_oracle
    z, sd, grads = _load()
    K = int(z["K"])
    model = _model(sd, K).eval()
    y, b, s, pt = (torch.from_numpy(z[k]) for k in ("y", "b", "sigma", "phi_true"))
    monkeypatch.setattr(autograd, "_cuda_eigh", _cpu_eigh)
    monkeypatch.setattr(admm_net.PhiEstADMMNet, "_prep", lambda self, y_, b_, s_: (y_, b_, s_.reshape(-1), None))
    calls = []

    def fake_forward_device(self, y_, b_, s_, out=None):
        calls.append(1)
        with torch.no_grad():
            return net_oracle.forward(self.state_dict(), y_, b_, s_, 10, 10, K)
    monkeypatch.setattr(admm_net.PhiEstADMMNet, "forward_device", fake_forward_device)
    phi = model(y, b, s)
    assert phi.requires_grad and phi.grad_fn is not None and len(calls) == 1
    assert np.abs(phi.detach().numpy() - z["phi"]).max() < 1e-4 * np.abs(z["phi"]).max()
    loss, _ = PhiAlignmentLoss()(phi, pt)
    loss.backward()
    assert len(calls) == 1                                   # backward rebuilt the graph, it did not call the fast path
    _compare_grads(model, grads, 2e-3)
    yg = y.clone().requires_grad_(True)
    model(yg, b, s).abs().sum().backward()
    assert yg.grad is not None and float(yg.grad.abs().max()) > 0
    with torch.no_grad():
        out = model(y, b, s)
    assert not out.requires_grad and out.grad_fn is None


def test_training_path_refuses_cpu_tensors():
    from admmnet_b200 import _capi
    from admmnet_b200.autograd import BatchedEigh
    with pytest.raises(_capi.AdmmnetError):
        BatchedEigh.apply(torch.eye(5, dtype=torch.complex64)[None])


def _ar_worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from admmnet_b200.training import allreduce_gradients
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.Linear(3, 2))
    x = torch.arange(8, dtype=torch.float32).view(2, 4) + rank
    net[0](x).sum().backward()                 # net[1] gets no gradient: must stay None, buffer offsets must hold
    allreduce_gradients(net)
    q.put((rank, net[0].weight.grad.numpy().copy(), net[0].bias.grad.numpy().copy(), net[1].weight.grad is None))
    dist.barrier()
    dist.destroy_process_group()


def test_flat_gradient_allreduce_two_ranks_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_ar_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    x0 = np.arange(8, dtype=np.float32).reshape(2, 4)
    want_w = np.tile(((x0.sum(0)) + (x0 + 1).sum(0)) / 2, (3, 1))
    for rank, gw, gb, none_kept in res:
        assert np.allclose(gw, want_w)
        assert np.allclose(gb, 2.0)
        assert none_kept


@pytest.mark.gpu
def test_training_gradients_match_reference_gpu():
    from admmnet_b200.autograd import PhiAlignmentLoss
    z, sd, grads = _load()
    model = _model(sd, int(z["K"]), "cuda")
    y, b, s, pt = (torch.from_numpy(z[k]).cuda() for k in ("y", "b", "sigma", "phi_true"))
    phi = model(y, b, s)
    assert phi.requires_grad
    loss, _ = PhiAlignmentLoss()(phi, pt)
    loss.backward()
    assert np.abs(phi.detach().cpu().numpy() - z["phi"]).max() < 1e-4 * np.abs(z["phi"]).max()
    assert abs(float(loss.detach()) - float(z["loss"])) < 1e-4 * float(z["loss"])
    worst = _compare_grads(model, grads, 1e-3)          # measured on B200: 2.6e-4
    print("worst relative gradient error", worst)


@pytest.mark.gpu
def test_train_step_lowers_the_loss_gpu():
    from admmnet_b200.autograd import PhiAlignmentLoss
    from admmnet_b200.training import train_step
    z, sd, _ = _load()
    model = _model(sd, int(z["K"]), "cuda")
    y, b, s, pt = (torch.from_numpy(z[k]).cuda() for k in ("y", "b", "sigma", "phi_true"))
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3, weight_decay=1e-4)
    crit = PhiAlignmentLoss()
    losses = [float(train_step(model, crit, opt, y, b, s, pt)[0]) for _ in range(8)]
    assert losses[-1] < losses[0]
    # eval mode still takes the fused inference path and agrees with the graph
    model.eval()
    with torch.no_grad():
        fast = model(y, b, s)
    model.train()
    slow = model(y, b, s).detach()
    assert float((fast - slow).abs().max() / slow.abs().max()) < 1e-4


@pytest.mark.gpu
def test_eval_mode_forward_is_differentiable_like_the_reference_gpu():
    """test/test_time_net.py:94-100 calls the model in eval() with grad enabled and gets a differentiable phi.  Here the
    values come from the fused kernels and the graph is rebuilt when backward is called (_FusedForward): same values as
    no_grad, a grad_fn on the result, and the gradients of the train()-mode graph; ADMMNet likewise."""
    from admmnet_b200.autograd import PhiAlignmentLoss
    z, sd, _ = _load()
    K = int(z["K"])
    y, b, s, pt = (torch.from_numpy(z[k]).cuda() for k in ("y", "b", "sigma", "phi_true"))
    crit = PhiAlignmentLoss()
    m_eval, m_train = _model(sd, K, "cuda").eval(), _model(sd, K, "cuda").train()
    phi = m_eval(y, b, s)
    assert phi.requires_grad and phi.grad_fn is not None
    with torch.no_grad():
        assert torch.equal(phi.detach(), m_eval(y, b, s))
    crit(phi, pt)[0].backward()
    crit(m_train(y, b, s), pt)[0].backward()
    live = 0
    for (n1, p1), (_, p2) in zip(m_eval.named_parameters(), m_train.named_parameters()):
        assert (p1.grad is None) == (p2.grad is None), n1
        if p1.grad is not None:
            live += 1
            assert float((p1.grad - p2.grad).abs().max()) <= 2e-3 * (float(p2.grad.abs().max()) + 1e-12), n1
    assert live > 40
    # gradient with respect to an input
    yg = y.clone().requires_grad_(True)
    m_eval(yg, b, s).abs().sum().backward()
    assert yg.grad is not None and float(yg.grad.abs().max()) > 0
    # host tensors in, host tensors out, still differentiable
    assert m_eval(y.cpu(), b.cpu(), s.cpu()).requires_grad
    # ADMMNet: all four outputs differentiable in eval mode
    import admmnet_b200
    torch.manual_seed(1)
    net = admmnet_b200.ADMMNet(10, 10, 3, 3).cuda().eval()
    tau, f, conf, phi4 = net(y, b, s)
    (tau.sum() + f.sum() + conf.sum() + phi4.abs().sum()).backward()
    assert net.peakSearchLayer.tau_regressor[0][0].weight.grad is not None and net.phiLayers[0].rho.grad is not None
    with torch.no_grad():
        t2, f2, c2, p2 = net(y, b, s)
    assert torch.equal(p2, phi4.detach()) and float((t2 - tau.detach()).abs().max()) < 1e-5


# ------------------------------------------------------------------ dataset / checkpoint formats
def _fake_split(d, n=6, nn=100, L=3, with_phi=True):
    rng = np.random.default_rng(0)
    os.makedirs(d)
    arrs = {"y_real": (n, nn), "y_imag": (n, nn), "b_real": (n, nn), "b_imag": (n, nn), "tau": (n, L), "f": (n, L),
            "C_real": (n, L), "C_imag": (n, L), "sigma": (n,), "ser": (n,)}
    if with_phi:
        arrs.update({"phi_real": (n, nn), "phi_imag": (n, nn)})
    out = {k: rng.standard_normal(s).astype(np.float32) for k, s in arrs.items()}
    out["L_true"] = np.full((n,), L, dtype=np.int32)
    for k, v in out.items():
        np.save(os.path.join(d, k + ".npy"), v)
    return out


def test_load_split_reads_the_reference_layout(tmp_path):
    from admmnet_b200.dataset import load_split
    raw = _fake_split(str(tmp_path / "train"))
    y, b, tau, f, C, L_true, sigma, phi = load_split(str(tmp_path), "train")
    assert y.dtype == torch.complex64 and L_true.dtype == torch.int64 and sigma.dtype == torch.float32
    assert np.array_equal(y.numpy(), raw["y_real"] + 1j * raw["y_imag"])
    assert np.array_equal(phi.numpy(), raw["phi_real"] + 1j * raw["phi_imag"])
    assert np.array_equal(C.numpy(), raw["C_real"] + 1j * raw["C_imag"])
    _fake_split(str(tmp_path / "val"), with_phi=False)          # base-class datasets carry no phi
    assert len(load_split(str(tmp_path), "val")) == 7
    with pytest.raises(ValueError):
        load_split(str(tmp_path), "test")


def test_checkpoint_layout_round_trip(tmp_path):
    from admmnet_b200.dataset import load_checkpoint, save_checkpoint
    from admmnet_b200.training import make_optimizer
    z, sd, _ = _load()
    model = _model(sd, int(z["K"]))
    opt, sch = make_optimizer(model)
    assert opt.param_groups[0]["lr"] == 2.5e-3 and len(opt.param_groups[0]["params"]) == len(list(model.parameters()))
    path = str(tmp_path / "best_model.pth")
    save_checkpoint(path, 4, model, opt, sch, 0.25, {"lr": 5e-3}, {"train_loss": [1.0]})
    ck = torch.load(path, weights_only=False)
    assert set(ck) == {"epoch", "model_state_dict", "optimizer_state_dict", "scheduler_state_dict", "best_val_loss",
                       "config", "history"}                       # trainPhi.py:238-246
    assert set(ck["model_state_dict"]) == set(sd)               # the reference's own state_dict keys
    fresh = _model({k: torch.zeros_like(v) for k, v in sd.items()}, int(z["K"]))
    opt2, sch2 = make_optimizer(fresh)
    start, best, _ = load_checkpoint(path, fresh, opt2, sch2)
    assert (start, best) == (5, 0.25)
    for k, v in fresh.state_dict().items():
        assert torch.equal(v, sd[k])


def test_batches_split_every_global_batch_over_the_ranks():
    from admmnet_b200.training import _batches
    x = torch.arange(10)
    order = torch.arange(10)
    got = [[t[0].tolist() for t in _batches((x,), 4, order, r, 2)] for r in range(2)]
    assert got[0] == [[0, 1], [4, 5], [8]] and got[1] == [[2, 3], [6, 7], [9]]
    # a tail batch smaller than the world: every rank still yields the same number of steps (empty shard)
    got = [[t[0].tolist() for t in _batches((x[:9],), 4, order[:9], r, 2)] for r in range(2)]
    assert got[0] == [[0, 1], [4, 5], [8]] and got[1] == [[2, 3], [6, 7], []]


def _ragged_worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from admmnet_b200.training import _batches, train_step
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    net = torch.nn.Linear(4, 1, bias=False)

    class M(torch.nn.Module):                      # stand-in with the (y, b, sigma) call signature of the net
        def __init__(self):
            super().__init__()
            self.lin = net

        def forward(self, y, b, sigma):
            return self.lin(y)

    model = M()
    crit = lambda out, tgt: (((out - tgt) ** 2).sum() / out.shape[0], {})
    opt = torch.optim.SGD(model.parameters(), lr=0.0)          # lr 0: only the averaged .grad matters
    X = torch.arange(20, dtype=torch.float32).view(5, 4) / 10
    T = torch.ones(5, 1)
    grads = []
    for (xb, tb) in _batches((X, T), 4, torch.arange(5), rank, world):     # batches of 4 and 1: rank 1's tail is empty
        train_step(model, crit, opt, xb, None, None, tb, max_norm=1e9)
        grads.append(net.weight.grad.numpy().copy())
    q.put((rank, grads))
    dist.barrier()
    dist.destroy_process_group()


def test_ragged_and_empty_shards_give_the_global_batch_gradient_gloo():
    """a tail batch smaller than the world must neither hang nor mis-weight the average."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_ragged_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    torch.manual_seed(0)
    net = torch.nn.Linear(4, 1, bias=False)
    X = torch.arange(20, dtype=torch.float32).view(5, 4) / 10
    want = []
    for lo, hi in ((0, 4), (4, 5)):                 # per-sample mean over the GLOBAL batch
        net.zero_grad()
        (((net(X[lo:hi]) - 1.0) ** 2).sum() / (hi - lo)).backward()
        want.append(net.weight.grad.numpy().copy())
    for r in range(2):
        assert len(res[r]) == 2
        for g, w in zip(res[r], want):
            assert np.allclose(g, w, rtol=1e-5, atol=1e-6)


@pytest.mark.gpu
def test_generate_dataset_writes_reference_format_gpu(tmp_path):
    from admmnet_b200.dataset import FIELDS, generate_dataset, load_split
    from oracle import classic_oracle
    generate_dataset(str(tmp_path), total_samples=200, seed=3)
    for split, n in (("train", 140), ("val", 30), ("test", 30)):
        for k in FIELDS:
            a = np.load(tmp_path / split / f"{k}.npy")
            assert a.shape[0] == n and a.dtype == (np.int32 if k == "L_true" else np.float32), (split, k)
    y, b, tau, f, C, L_true, sigma, phi = load_split(str(tmp_path), "train")
    assert float(tau.min()) >= 0.1 and float(tau.max()) <= 0.9 and float(f.abs().max()) <= 0.4
    assert np.allclose(np.abs(b.numpy()), 1.0, atol=1e-6)
    ser = np.load(tmp_path / "train" / "ser.npy")
    assert ((ser > 0) == (sigma.numpy() > 1.0)).all() and 0 < ser.mean() < 30          # QPSK at 7 dB: a few percent
    # labels are the classical solver's phi (generate_data.py:454)
    for i in (0, 57, 139):
        ref = classic_oracle.admm_for_us(y[i].numpy().astype(np.complex128), b[i].numpy().astype(np.complex128), 10, 10,
                                         1, float(sigma[i]), {"eta_abs": 1e-7, "eta_rel": 1e-7, "max_iter": 100})[0]
        assert np.abs(phi[i].numpy() - ref).max() < 2e-6 * np.abs(ref).max()
    # the noise SNR is drawn per signal from [5, 25) dB: residual power spans the range
    info = json.load(open(tmp_path / "dataset_config.json"))
    assert info["train_samples"] == 140 and info["snr_range"] == [5, 25]


@pytest.mark.gpu
def test_fit_trains_checkpoints_and_resumes_gpu(tmp_path):
    from admmnet_b200.admm_net import PhiEstADMMNet
    from admmnet_b200.dataset import generate_dataset, load_split
    from admmnet_b200.training import fit
    generate_dataset(str(tmp_path / "data"), total_samples=160, seed=1)
    train, val = load_split(str(tmp_path / "data"), "train"), load_split(str(tmp_path / "data"), "val")
    os.makedirs(tmp_path / "ck")
    cfg = {"batch_size": 56, "epochs": 3, "lr": 5e-3, "weight_decay": 1e-3, "checkpoint_dir": str(tmp_path / "ck")}
    torch.manual_seed(0)
    model = PhiEstADMMNet(10, 10, 3, 3).cuda()
    hist = fit(model, train, val, cfg, log=lambda *_: None)
    assert len(hist["train_loss"]) == 3 and hist["train_loss"][-1] < hist["train_loss"][0]
    ck = torch.load(tmp_path / "ck" / "best_model.pth", weights_only=False)
    assert ck["best_val_loss"] == min(hist["val_loss"])
    cfg["epochs"] = ck["epoch"] + 2
    hist2 = fit(PhiEstADMMNet(10, 10, 3, 3).cuda(), train, val, cfg, log=lambda *_: None)   # resumes after the best epoch
    assert len(hist2["train_loss"]) == 1


# ------------------------------------------------------------------ train.py: full ADMMNet + BasicANMLoss
GOLD_FULL = os.path.join(ROOT, "tests", "golden", "train_full_grads_k3.npz")


def _load_full():
    z = np.load(GOLD_FULL)
    sd = {k[4:].replace("__", "."): torch.from_numpy(z[k]) for k in z.files if k.startswith("sd__")}
    grads = {k[6:].replace("__", "."): z[k] for k in z.files if k.startswith("grad__")}
    return z, sd, grads


def _full_step(model, z, device, eigh=None):
    from admmnet_b200.autograd import BasicANMLoss
    y, b, s = (torch.from_numpy(z[k]).to(device) for k in ("y", "b", "sigma"))
    tau, f, conf, phi = model.forward_differentiable(y, b, s, _eigh=eigh)
    truth = {"tau_true": torch.from_numpy(z["tau_true"]).to(device), "f_true": torch.from_numpy(z["f_true"]).to(device),
             "L_true": torch.from_numpy(z["L_true"]).to(device)}
    loss, parts = BasicANMLoss()({"tau_est": tau, "f_est": f, "confidences": conf, "phi_final": phi}, truth)
    loss.backward()
    return tau, f, conf, loss, parts


def _check_full(model, z, grads, tau, f, conf, loss, parts, tol):
    assert np.abs(tau.detach().cpu().numpy() - z["tau"]).max() < 2e-4
    assert np.abs(f.detach().cpu().numpy() - z["f"]).max() < 2e-4
    assert np.abs(conf.detach().cpu().numpy() - z["conf"]).max() < 2e-4
    assert abs(float(loss.detach()) - float(z["loss"])) < 2e-4 * float(z["loss"])
    assert abs(float(parts["reg_loss"].detach()) - float(z["reg_loss"])) < 1e-4 * float(z["reg_loss"])
    floor = 1e-4 * max(np.abs(v).max() for v in grads.values() if v.size)
    live = 0
    for name, p in model.named_parameters():
        ref = grads[name]
        if ref.size == 0:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, name
            continue
        live += 1
        err = np.abs(p.grad.detach().cpu().numpy() - ref).max() / max(np.abs(ref).max(), floor)
        assert err < tol, (name, err)
    assert live == 82


def test_admmnet_training_graph_matches_reference_cpu():
    """ADMMNet (unrolled layers + learned head) with BasicANMLoss, eval mode so the attention dropout is off:
    outputs, loss and all 82 live gradients against the reference's autograd (samples with 0, 1, 2, 3 targets)."""
    from admmnet_b200.admm_net import ADMMNet
    z, sd, grads = _load_full()
    model = ADMMNet(10, 10, 3, int(z["K"]))
    model.load_state_dict(sd)
    model.eval()
    out = _full_step(model, z, "cpu", eigh=_cpu_eigh)
    _check_full(model, z, grads, *out, tol=2e-3)


def test_basic_parameter_loss_matches_the_per_sample_loop():
    from admmnet_b200.autograd import basic_parameter_loss
    torch.manual_seed(0)
    tp, fp, cf = torch.rand(9, 3), torch.rand(9, 3) - 0.5, torch.rand(9, 3)
    tt, ft = torch.rand(9, 3), torch.rand(9, 3) - 0.5
    L = torch.tensor([0, 1, 2, 3, 3, 0, 2, 1, 3])
    want = 0.0
    for i in range(9):                                     # loss.py:13-28, literally
        n = int(L[i])
        if n == 0:
            want = want + torch.sum(cf[i] ** 2)
        else:
            want = want + (torch.nn.functional.mse_loss(tp[i, :n], tt[i, :n]) +
                           torch.nn.functional.mse_loss(fp[i, :n], ft[i, :n]) +
                           0.1 * torch.nn.functional.mse_loss(cf[i, :n], torch.ones(n)))
    assert abs(float(basic_parameter_loss(tp, fp, tt, ft, cf, L)) - float(want / 9)) < 1e-6


@pytest.mark.gpu
def test_admmnet_training_gradients_match_reference_gpu():
    from admmnet_b200.admm_net import ADMMNet
    z, sd, grads = _load_full()
    model = ADMMNet(10, 10, 3, int(z["K"]))
    model.load_state_dict(sd)
    model = model.cuda().eval()
    out = _full_step(model, z, "cuda")
    _check_full(model, z, grads, *out, tol=5e-3)
    # train() mode goes through the same graph from the module call (dropout on: only shapes and grads are checked)
    model.train()
    model.zero_grad()
    y, b, s = (torch.from_numpy(z[k]).cuda() for k in ("y", "b", "sigma"))
    tau, f, conf, phi = model(y, b, s)
    assert tau.requires_grad and phi.requires_grad and tau.shape == (7, 3)


def test_save_dataset_round_trip_and_no_cpu_generation(tmp_path):
    from admmnet_b200 import _capi
    from admmnet_b200.dataset import FIELDS, generate_dataset, load_split, save_dataset, split_sizes
    assert split_sizes(10000) == (7000, 1500, 1500) and split_sizes(7, 0.7, 0.15) == (4, 1, 2)   # generate_data.py:54-57
    raw = {}
    rng = np.random.default_rng(1)
    for name, n in (("train", 5), ("val", 2), ("test", 1)):
        raw[name] = {k: (np.full((n,), 3, np.int32) if k == "L_true" else
                         rng.standard_normal((n,) if k in ("sigma", "ser") else (n, 3) if k in ("tau", "f", "C_real", "C_imag")
                                             else (n, 100)).astype(np.float32)) for k in FIELDS}
    save_dataset(str(tmp_path), raw, {"Nb": 10, "Nd": 10, "L_max": 3, "snr_range": [5, 25], "total_samples": 8})
    assert json.load(open(tmp_path / "dataset_config.json"))["total_samples"] == 8
    assert (tmp_path / "dataset_info.npz").exists()
    y, b, tau, f, C, L_true, sigma, phi = load_split(str(tmp_path), "val")
    assert np.array_equal(phi.numpy(), raw["val"]["phi_real"] + 1j * raw["val"]["phi_imag"])
    assert np.array_equal(tau.numpy(), raw["val"]["tau"]) and L_true.tolist() == [3, 3]
    if not torch.cuda.is_available():
        with pytest.raises(_capi.AdmmnetError):                 # the writer is a device path: no CPU fallback
            generate_dataset(str(tmp_path / "x"), total_samples=10)


# ------------------------------------------------------------------ train.py loop and metrics
def test_param_rmse_and_detection_match_the_reference_loops():
    """train.py:279-294 (per-sample RMSE over the first L targets) and :409-443 (detection statistics, precision /
    recall / F1), restated literally here as loops, against the vectorised forms."""
    from admmnet_b200.training import detection_counts, detection_scores, param_rmse
    g = torch.Generator().manual_seed(0)
    B, Lmax = 57, 3
    tau_est, tau_true = torch.rand(B, Lmax, generator=g), torch.rand(B, Lmax, generator=g)
    conf = torch.rand(B, Lmax, generator=g)
    L_true = torch.randint(0, Lmax + 1, (B,), generator=g)
    want = []
    for i in range(B):
        L = L_true[i].item()
        if L > 0:
            want.append(torch.sqrt(torch.mean((tau_est[i, :L] - tau_true[i, :L]) ** 2)).item())
    got = param_rmse(tau_est, tau_true, L_true)
    assert got.shape[0] == len(want) and np.allclose(got.numpy(), np.array(want), rtol=1e-6, atol=1e-7)
    tp = fp = fn = 0
    for i in range(B):
        L_i = L_true[i].item()
        det = torch.sum(conf[i] > 0.5).item()
        if L_i > 0 and det > 0:
            tp += min(L_i, det)
        if det > L_i:
            fp += det - L_i
        if L_i > det:
            fn += L_i - det
    assert detection_counts(conf, L_true) == (tp, fp, fn)
    p, r, f1 = detection_scores(tp, fp, fn)
    assert p == tp / (tp + fp) and r == tp / (tp + fn) and abs(f1 - 2 * p * r / (p + r)) < 1e-12
    assert detection_scores(0, 0, 0) == (0, 0, 0)


@pytest.mark.gpu
def test_fit_admmnet_runs_the_train_py_loop_gpu(tmp_path):
    """train.py:160-447 end to end on a small generated dataset: history keys, checkpoint, test metrics."""
    import admmnet_b200
    from admmnet_b200.dataset import generate_dataset, load_split
    from admmnet_b200.training import fit_admmnet
    data = str(tmp_path / "data")
    generate_dataset(data, total_samples=240, seed=5)
    train, val, test = (load_split(data, s) for s in ("train", "val", "test"))
    torch.manual_seed(0)
    model = admmnet_b200.ADMMNet(10, 10, 3, 3).cuda()
    os.makedirs(tmp_path / "ck")
    os.makedirs(tmp_path / "logs")
    cfg = {"batch_size": 64, "epochs": 2, "lr": 1e-3, "weight_decay": 1e-3, "checkpoint_dir": str(tmp_path / "ck"),
           "log_dir": str(tmp_path / "logs")}
    history, result = fit_admmnet(model, train, val, test, cfg, log=lambda *a: None)
    assert all(len(history[k]) == 2 for k in ("train_loss", "val_loss", "tau_rmse", "f_rmse", "lr"))
    assert np.isfinite(history["train_loss"]).all() and history["tau_rmse"][-1] > 0
    assert os.path.exists(tmp_path / "ck" / "best_model.pth")
    assert set(result) == {"test_loss", "tau_rmse", "f_rmse", "precision", "recall", "f1_score", "detection_stats"}
    assert 0 <= result["precision"] <= 1 and 0 <= result["recall"] <= 1
    assert os.path.exists(tmp_path / "logs" / "training_history.json") and os.path.exists(tmp_path / "logs" / "test_result.json")


@pytest.mark.gpu
def test_cuda_graph_train_step_matches_eager_gpu():
    """GraphedTrainStep replays the captured step: same losses and parameters as the eager train_step."""
    from admmnet_b200.autograd import PhiAlignmentLoss, eigh_status
    from admmnet_b200.training import GraphedTrainStep, make_optimizer, train_step
    z, sd, _ = _load()
    K = int(z["K"])
    y, b, s, pt = (torch.from_numpy(z[k]).cuda() for k in ("y", "b", "sigma", "phi_true"))
    crit = PhiAlignmentLoss()
    m1, m2 = _model(sd, K, "cuda"), _model(sd, K, "cuda")
    o1, _ = make_optimizer(m1, 5e-3, 1e-3, capturable=True)
    o2, _ = make_optimizer(m2, 5e-3, 1e-3, capturable=True)
    step = GraphedTrainStep(m2, crit, o2, (y, b, s, pt))
    l1 = [float(train_step(m1, crit, o1, y, b, s, pt)[0]) for _ in range(4)]
    l2 = [float(step(y, b, s, pt)) for _ in range(4)]
    assert eigh_status(y.device) == 0
    assert np.allclose(l1, l2, rtol=2e-3), (l1, l2)
    # AdamW moves every parameter by ~lr per step whatever the gradient's size, so a parameter whose gradient is at
    # rounding level (sign decided by the last bits of the eigen-solver) may legitimately differ by up to
    # 2 * lr * steps between two runs; all others must agree closely.
    lr, steps, loose = 5e-3, 4, 0
    names = [n for n, _ in m1.named_parameters()]
    for (n1, p1), (_, p2) in zip(m1.named_parameters(), m2.named_parameters()):
        diff = float((p1 - p2).abs().max())
        assert diff <= 2 * lr * steps + 1e-6, n1
        if diff > 2e-3 * (float(p1.abs().max()) + 1e-3):
            loose += 1
    assert loose <= max(2, len(names) // 20), loose
