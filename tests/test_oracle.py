"""CPU tests: pin the oracle restatements (oracle/*.py) to the golden vectors produced by the
reference itself (tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import classic_oracle, net_oracle, peak_oracle
from tests.helpers import GOLDEN, load_net_case, parse_opts, rel_err

# tolerance for "same fp32 algorithm, different op grouping": the reference's own forward moves by
# ~1e-5 (max-norm relative) under 1-ulp perturbations, well inside the 1e-4
# parity bar of BASELINE.json.
NET_TOL = 5e-5


@pytest.mark.parametrize("tag", ["init_k10", "pert_k10", "pert_k5"])
def test_net_oracle_matches_reference(tag):
    z, sd = load_net_case(tag)
    K = int(z["K"])
    y, b, s = (torch.from_numpy(z[k]) for k in ("y", "b", "sigma"))
    taps = []
    phi = net_oracle.forward(sd, y, b, s, 10, 10, K, taps=taps).numpy()
    assert rel_err(phi, z["phi_batch"]).max() < NET_TOL
    # layer-0 phi has no eigendecomposition upstream: bit exact
    assert np.array_equal(taps[0]["phi"].numpy(), z["batch_phi_layers"][0])
    for k in range(K):
        assert rel_err(taps[k]["phi"].numpy(), z["batch_phi_layers"][k]).max() < NET_TOL
    phi1 = net_oracle.forward(sd, y[:1], b[:1], s[:1].reshape(1, 1), 10, 10, K).numpy()
    assert rel_err(phi1, z["phi_single"]).max() < NET_TOL


def test_batch_mean_coupling_is_material():
    """the same signal alone vs in a batch differs by far more than the tolerance."""
    z, _ = load_net_case("init_k10")
    assert rel_err(z["phi_single"][0], z["phi_batch"][0]) > 1e-4


def test_chunked_forward_equals_per_chunk_reference():
    z, sd = load_net_case("init_k10")
    y, b, s = (torch.from_numpy(z[k]) for k in ("y", "b", "sigma"))
    phi = net_oracle.forward(sd, y, b, s, 10, 10, 10, chunk=1).numpy()
    assert rel_err(phi[0], z["phi_single"][0]) < NET_TOL


def test_classic_oracle_matches_reference():
    z = np.load(os.path.join(GOLDEN, "classic.npz"))
    for c in range(len(z["iters"])):
        i = int(z["case_sig"][c])
        opts = dict(rho=float(z["case_rho"][c]), max_iter=int(z["case_max_iter"][c]))
        umi, mi = bool(z["case_use_min_iter"][c]), int(z["case_min_iter"][c])
        n_exec = classic_oracle.executed_iterations(opts["max_iter"], umi, mi)
        assert n_exec == int(z["iters"][c])
        rec = classic_oracle.admm_linear_recursion(z["y"][i][None], z["b"][i][None], opts["rho"], n_exec)[0]
        assert rel_err(rec, z["phi"][c]) < 1e-11
    # the literal loop (slow: SVD per iteration) on two cases
    for c in (0, 12):
        i = int(z["case_sig"][c])
        opts = dict(rho=float(z["case_rho"][c]), max_iter=int(z["case_max_iter"][c]), eta_abs=1e-7, eta_rel=1e-7)
        phi, it = classic_oracle.admm_for_us(z["y"][i], z["b"][i], 10, 10, 1.0, float(z["sigma"][i]), opts,
                                             bool(z["case_use_min_iter"][c]), int(z["case_min_iter"][c]))
        assert it == int(z["iters"][c])
        assert rel_err(phi, z["phi"][c]) < 1e-11


def test_projection_is_a_projection():
    rng = np.random.default_rng(0)
    for _ in range(20):
        t = rng.normal(size=12)
        A = abs(rng.normal()) * 3
        h = classic_oracle.project_linf_sum(t, A)
        assert A * np.abs(h).max() + h.sum() <= 1 + 1e-9
        # optimality: no feasible random point is closer
        for _ in range(50):
            g = h + 0.05 * rng.normal(size=12)
            if A * np.abs(g).max() + g.sum() <= 1:
                assert np.linalg.norm(g - t) >= np.linalg.norm(h - t) - 1e-9
    t = np.full(5, 0.01)
    assert np.array_equal(classic_oracle.project_linf_sum(t, 1.0), t)


def test_local_maxima_plateau_kat():
    """The reference's own plateau example, peakSearchUtils.py:427-436."""
    img = np.array([[1, 1, 1, 2, 3], [1, 5, 5, 4, 3], [2, 5, 5, 4, 2], [3, 4, 4, 3, 1]])
    exp = np.zeros_like(img, dtype=bool)
    exp[1:3, 1:3] = True
    assert np.array_equal(peak_oracle.local_maxima(img), exp)
    assert not peak_oracle.local_maxima(np.ones((4, 4))).any()
    edge = np.array([[3, 1, 1], [1, 1, 1], [1, 1, 2.0]])
    assert np.array_equal(np.argwhere(peak_oracle.local_maxima(edge)), [[0, 0], [2, 2]])


def test_peak_oracle_matches_reference():
    z = np.load(os.path.join(GOLDEN, "peaks.npz"))
    names = sorted({k.split("__")[0] for k in z.files if "__" in k})
    assert len(names) >= 8
    for name in names:
        opts = parse_opts(z[f"{name}__opts"])
        slow = name in ("net0", "classic0_default")   # literal double loop on two cases, separable elsewhere
        surface = peak_oracle.peak_search if slow else (
            lambda phi, X, xb, Y, yb: peak_oracle.peak_search_separable(phi, X[0], xb, Y[:, 0], yb))
        got = peak_oracle.alt_peak_search({"phi": z[f"{name}__phi"], "xbase": 10, "ybase": 10}, opts, surface)
        exp = z[f"{name}__peaks"]
        assert got.shape == exp.shape, name
        assert np.array_equal(got[:, :2], exp[:, :2]), name
        np.testing.assert_allclose(got[:, 2], exp[:, 2], rtol=1e-10, atol=1e-300)
    ax = np.arange(0, 1 - 0.01, 0.01)
    ay = np.arange(-0.5, 0.5 - 0.01, 0.01)
    sep = peak_oracle.peak_search_separable(z["net0__phi"], ax, 10, ay, 10)
    np.testing.assert_allclose(sep, z["surface_net0"], rtol=1e-9, atol=1e-18)


def test_classic_tolerances_of_zero_run_to_max_iter():
    """with eta_abs = eta_rel = 0 the reference's loop (admm.py:95-112) never meets its stopping test and
    runs max_iter iterations; the literal port shows it, the closed-form recursion agrees, and the product's
    iteration rule (admm-net_b200/admm.py::executed_iterations) follows it."""
    import importlib
    from oracle import classic_oracle, signals
    y, b, s, _ = signals.generate(1, seed=4)
    y, b = y[0].astype(np.complex128), b[0].astype(np.complex128)
    opts = {"rho": 1.0, "max_iter": 9, "eta_abs": 0.0, "eta_rel": 0.0}
    phi, it = classic_oracle.admm_for_us(y, b, 10, 10, 0.1, float(s[0]), opts)
    assert it == 9
    rec = classic_oracle.admm_linear_recursion(y[None], b[None], 1.0, 9)[0]
    assert np.abs(phi - rec).max() < 1e-12 * np.abs(rec).max()
    rule = importlib.import_module("admmnet_b200.admm").executed_iterations
    bn = float(np.sqrt(2 * np.sum(np.abs(phi) ** 2) + 1 / 0.1 ** 4))
    assert rule(opts, True, 5, bn, 100) == 9
    assert rule({"max_iter": 9}, True, 5, bn, 100) == 5
    assert rule({"max_iter": 9}, False, 5, bn, 100) == 2
    with pytest.raises(ValueError):
        rule({"max_iter": 9, "eta_abs": 1e-16, "eta_rel": 0.0}, True, 5, bn, 100)
