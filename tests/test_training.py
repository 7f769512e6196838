"""Training path: the differentiable graph and its gradients against the reference's own
autograd (tests/golden/train_grads_k4.npz, written by tests/golden/make_golden.py from trainPhi.py's loss/backward),
and the flat gradient all-reduce on two gloo ranks."""
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
    assert abs(float(loss) - float(z["loss"])) < 1e-4 * float(z["loss"])
    assert abs(float(parts["phase_loss"]) - float(z["phase_loss"])) < 1e-4 * float(z["phase_loss"])
    _compare_grads(model, grads, 2e-3)


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
    assert abs(float(loss) - float(z["loss"])) < 1e-4 * float(z["loss"])
    worst = _compare_grads(model, grads, 5e-3)
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
