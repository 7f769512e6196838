"""CPU tests of the host side: the C ABI loads and exports every symbol of include/admmnet_b200.h (no
compute without a GPU), parameter packing, API mirror, error behaviour, and the multi-rank orchestration
(gloo, world_size 2)."""
import ctypes as C
import os
import re
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g
    from admmnet_b200 import _capi
    if not os.path.exists(_capi.LIB_PATH):
        g.build()
    return _capi


def test_library_exports_every_declared_symbol(built):
    hdr = open(os.path.join(ROOT, "include", "admmnet_b200.h")).read()
    declared = set(re.findall(r"^(?:int|double|const char\*)\s+(\w+)\s*\(", hdr, flags=re.M))
    assert declared == set(built.EXPORTS), declared ^ set(built.EXPORTS)
    L = built.lib()
    for name in declared:
        assert getattr(L, name) is not None
    assert L.admmnet_version() >= 100


def test_argument_errors_are_reported_not_crashed(built):
    L = built.lib()
    nb = C.c_size_t()
    assert L.admmnet_forward_workspace_bytes(0, 0, 100, 10, 0, C.byref(nb)) < 0
    assert b"positive" in L.admmnet_last_error()
    assert L.admmnet_forward_workspace_bytes(4, 0, 300, 10, 0, C.byref(nb)) < 0       # n > 256
    assert L.admmnet_forward_workspace_bytes(4, 0, 256, 10, 0, C.byref(nb)) == 0 and nb.value > 0   # Jacobi path sizes
    assert L.admmnet_forward_workspace_bytes(4, 0, 100, 10, 1000, C.byref(nb)) < 0    # rcap not a multiple of 1024
    assert L.admmnet_forward_workspace_bytes(1024, 256, 100, 10, 0, C.byref(nb)) == 0 and nb.value > 0
    small = nb.value
    assert L.admmnet_forward_workspace_bytes(1024, 1024, 100, 10, 0, C.byref(nb)) == 0 and nb.value > small
    assert L.admmnet_forward(None, None, None, 4, 4, 10, 10, 10, None, None, None, 0, 0, None) < 0
    assert L.admm_classic_forward(None, None, 1, 4, 100, 1.0, 5, None, None) < 0
    assert L.peak_search_full(None, 0, 1, 10, 10, None, 1, None, 1, 0., 1., .01, -.5, .5, .01, .1, 1, 8, None, None, 0,
                              None, None, None, None) < 0
    assert L.admmnet_eigh_workspace_bytes(1, 300, 0, C.byref(nb)) < 0


def test_param_packing_matches_layout(built):
    from admmnet_b200 import params
    from tests.helpers import load_net_case
    import torch.nn.functional as F
    z, sd = load_net_case("pert_k10")
    P = params.pack_state_dict(sd, 100, 10)
    assert P.shape == (10, built.lib().admmnet_param_stride(100)) == (10, params.param_stride(100))
    k = 4
    assert P[k, params.P_RHO_PHI] == F.softplus(sd[f"phiLayers.{k}.rho"])
    assert P[k, params.P_KNORM] == torch.tensor(k / 10.0)
    lam = F.softplus(sd[f"gLayers.{k}.lambda_param"])
    assert P[k, params.P_C0] == torch.tensor((1.0 / (lam ** 2 + 1e-8)).item())
    W1 = sd[f"hLayers.{k}.correction_net.0.weight"]
    assert torch.equal(P[k, params.P_HW1T:params.P_HW1T + 6400].reshape(100, 64), W1.t())
    W2 = sd[f"hLayers.{k}.correction_net.2.weight"]
    assert torch.equal(P[k, params.P_HW1T + 6400:params.P_HW1T + 12800].reshape(64, 100), W2.t())
    assert torch.equal(P[k, params.P_ZU1:params.P_ZU1 + 96].reshape(32, 3), sd[f"zLayers.{k}.residual_scale_net.0.weight"])


def test_module_mirror_state_dict_and_seeded_init(built):
    """Same keys, shapes and (same seed) same initial values as the reference module."""
    import admm_net
    from tests.helpers import load_net_case
    z, sd = load_net_case("init_k10")
    torch.manual_seed(0)
    net = admm_net.PhiEstADMMNet(10, 10, 3, 10)
    mine = net.state_dict()
    assert list(mine.keys()) == list(sd.keys())
    for k in sd:
        assert torch.equal(mine[k], sd[k]), k
    assert sum(v.numel() for v in mine.values()) == 132230
    net.load_state_dict(load_net_case("pert_k10")[1])                # reference checkpoints load
    groups = [n for n, _ in net.named_parameters() if any(s in n for s in ("phiLayers", "hLayers", "gLayers", "zLayers"))]
    assert len(groups) == len(list(net.parameters()))                # trainPhi.py:106-111 grouping still matches


def test_admmnet_mirror_accepts_reference_state_dict(built):
    import admm_net
    from admmnet_b200 import params
    z = np.load(os.path.join(ROOT, "tests", "golden", "admmnet_full_k3.npz"))
    sd = {k[4:].replace("__", "."): torch.from_numpy(z[k]) for k in z.files if k.startswith("sd__")}
    net = admm_net.ADMMNet(10, 10, 3, 3)
    assert list(net.state_dict().keys()) == list(sd.keys())
    net.load_state_dict(sd)
    H = params.pack_head(net.state_dict(), 100, 3)
    assert H.numel() == built.lib().admmnet_head_param_count(100, 3) == params.head_param_count(100, 3)
    W1 = sd["peakSearchLayer.feature_extractor.0.weight"]
    assert torch.equal(H[:200 * 128].reshape(200, 128), W1.t())
    # seeded construction consumes the RNG exactly like the reference (same module order)
    torch.manual_seed(0)
    a = admm_net.ADMMNet(4, 5, 2, 2)
    assert a.peakSearchLayer.position_encoder.shape == (20, 2)
    assert float(a.peakSearchLayer.position_encoder[5, 0]) == pytest.approx(1 / 3) and \
        float(a.peakSearchLayer.position_encoder[5, 1]) == pytest.approx(-0.5)


def test_product_path_has_no_cpu_fallback(built):
    import admm_net
    import admm
    from utils import peakSearchUtils
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    net = admm_net.PhiEstADMMNet(10, 10, 3, 2)
    with pytest.raises(built.AdmmnetError):
        net(torch.ones(1, 100, dtype=torch.complex64), torch.ones(1, 100, dtype=torch.complex64), torch.ones(1))
    with pytest.raises(built.AdmmnetError):
        admm.admm_for_us(np.ones(100, complex), np.ones(100, complex), 10, 10, 1.0, 1.0)
    with pytest.raises(built.AdmmnetError):
        peakSearchUtils.alt_peak_search({"phi": np.ones(100, complex), "xbase": 10, "ybase": 10})
    # ... and nothing under the package imports the oracle
    for root, _, files in os.walk(os.path.join(ROOT, "admm-net_b200")):
        for f in files:
            if f.endswith(".py"):
                assert "oracle" not in open(os.path.join(root, f)).read(), f


def test_iteration_count_rule(built):
    from admmnet_b200.admm import executed_iterations
    assert executed_iterations(dict(max_iter=100)) == 5            # main.py:88-95 -> 5 (results/time/time.txt regime)
    assert executed_iterations(None) == 5
    assert executed_iterations(dict(max_iter=3)) == 3
    assert executed_iterations(dict(max_iter=50), use_min_iter=False) == 2
    assert executed_iterations(None, True, 7) == 7
    assert executed_iterations(None, True, 1) == 2


def test_coarse_axes_follow_reference_quirk(built):
    from admmnet_b200.peaksearch import DEFAULT_OPTS, coarse_axes
    ax, ay = coarse_axes({**DEFAULT_OPTS})
    assert len(ax) == 99 and len(ay) == 99
    ax, ay = coarse_axes({**DEFAULT_OPTS, "xstep": 0.04, "ystep": 0.02})
    assert len(ay) == len(np.arange(-0.5, 0.5 - 0.04, 0.02))        # peakSearchUtils.py:106 uses xstep


def test_shard_range():
    from admmnet_b200.sharding import shard_range
    for B, W in [(10, 3), (8, 8), (5, 8), (1048576, 8)]:
        spans = [shard_range(B, r, W) for r in range(W)]
        assert spans[0][0] == 0 and spans[-1][1] == B
        assert all(spans[i][1] == spans[i + 1][0] for i in range(W - 1))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1


# ------------------------------------------------------------------ world_size 2, gloo
class _OracleEngine:
    """Stand-in for sharding.CudaEngine built on the CPU oracle: same interface, lets the all-reduce
    orchestration of run_layers be checked without a GPU."""

    def __init__(self, sd, y, b, s, K):
        from oracle import net_oracle
        self.no, self.sd, self.y, self.b, self.s, self.K = net_oracle, sd, y, b, s, K
        B = y.shape[0]
        self.G = torch.zeros(B, 101, 101)
        self.Z = torch.zeros(B, 101, 101)
        self.rs = torch.zeros(K, dtype=torch.float64)
        self.pending = None

    def count(self):
        return self.y.shape[0]

    def _apply_pending(self):
        if self.pending is not None:
            phi, h, G, k, mean = self.pending
            p = self.no.layer_params(self.sd, k)
            self.Z, _, _ = self.no.z_update(phi, h, G, self.Z, k, p["z"], mean_r=mean)
            self.pending = None

    def layer(self, k):
        self._apply_pending()
        p = self.no.layer_params(self.sd, k)
        phi = self.no.phi_update(self.y, self.b, self.G, self.Z, p["phi"]["rho"])
        h = self.no.h_update(self.G, self.Z, self.s, 100, p["h"])
        G, _, _, _ = self.no.g_update(phi, h, self.Z, p["g"])
        _, r, _ = self.no.z_update(phi, h, G, self.Z, k, p["z"])
        self.G = G
        self.rs[k] = r.double().sum()
        self.cur = (phi, h, G, k)

    def rsum(self, k):
        return self.rs[k:k + 1]

    def set_mean(self, k, count):
        self.pending = self.cur + (torch.tensor(float(self.rs[k] / count), dtype=torch.float32),)

    def final(self):
        self._apply_pending()
        p = self.no.layer_params(self.sd, self.K - 1)
        return self.no.phi_update(self.y, self.b, self.G, self.Z, p["phi"]["rho"])


def _worker(rank, world, port, tag, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from admmnet_b200.sharding import run_layers, shard_range
    from tests.helpers import load_net_case
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    z, sd = load_net_case(tag)
    K = int(z["K"])
    y, b, s = (torch.from_numpy(z[k]) for k in ("y", "b", "sigma"))
    lo, hi = shard_range(y.shape[0], rank, world)
    out = {}
    for scope in ("global", "shard"):
        eng = _OracleEngine(sd, y[lo:hi], b[lo:hi], s[lo:hi], K)
        out[scope] = run_layers(eng, K, scope).numpy()
    q.put((rank, lo, hi, out))
    dist.barrier()
    dist.destroy_process_group()


def test_global_norm_scope_two_ranks_gloo():
    import torch.multiprocessing as mp
    from oracle import net_oracle
    from tests.helpers import load_net_case, rel_err
    tag = "pert_k5"
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, tag, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    z, sd = load_net_case(tag)
    y, b, s = (torch.from_numpy(z[k]) for k in ("y", "b", "sigma"))
    phi_g = np.concatenate([r[3]["global"] for r in res])
    # 'global': two ranks reproduce the reference's whole-batch result
    assert rel_err(phi_g, z["phi_batch"]).max() < 5e-5
    # 'shard': each rank equals the oracle run on its shard alone, and differs from the whole-batch result
    for rank, lo, hi, out in res:
        ref = net_oracle.forward(sd, y[lo:hi], b[lo:hi], s[lo:hi], 10, 10, int(z["K"])).numpy()
        assert rel_err(out["shard"], ref).max() < 5e-5
    assert rel_err(np.concatenate([r[3]["shard"] for r in res]), z["phi_batch"]).max() > 1e-4


def test_arrowhead_algorithm_emulation_is_backward_stable():
    """The layer-0 shortcut's algorithm (fp32 emulation of csrc/arrow_kernels.cu, tests/arrow_emulation.py) on
    adversarial arrowheads: clustered and nearly repeated poles, weak and mixed couplings, a corner element below
    every pole, large pole scale.  Residual and orthogonality stay at fp32 rounding in every case."""
    from tests.arrow_emulation import arrow_eigh
    rng = np.random.default_rng(0)
    f32 = np.float32
    done = 0
    for trial in range(32):
        n = int(rng.choice([3, 7, 20, 40]))
        kind = trial % 8
        h = (rng.standard_normal(n) * 0.1).astype(f32)
        phi = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
        c0 = f32(1.8)
        if kind == 1:
            h = (np.sort(rng.random(n)) * 1e-3).astype(f32)
        elif kind == 2:
            h, phi = (h * f32(0.01)).astype(f32), (phi * 3).astype(np.complex64)
        elif kind == 3:
            phi = (phi * f32(1e-3)).astype(np.complex64)
        elif kind == 4:
            phi[::2] *= f32(1e-4)
        elif kind == 5:
            c0 = f32(rng.standard_normal() * 5)
        elif kind == 6:
            h = (h * f32(100)).astype(f32)
        elif kind == 7:
            h = (np.round(h * 1e3) / 1e3 + np.arange(n) * 1e-7).astype(f32)       # nearly repeated poles
        if len(np.unique(h)) < n:
            continue
        res = arrow_eigh(h, phi, c0)
        assert res is not None, (trial, kind, n)
        lam, U = res
        A = np.zeros((n + 1, n + 1), np.complex128)
        A[:n, :n] = np.diag(h.astype(float))
        A[:n, n], A[n, :n], A[n, n] = phi, np.conj(phi), c0
        sc = np.linalg.norm(A, 2)
        assert np.abs(A @ U - U * lam[None, :]).max() / sc < 2e-6, (trial, kind, n)
        assert np.abs(U.conj().T @ U - np.eye(n + 1)).max() < 3e-6, (trial, kind, n)
        assert np.abs(np.sort(lam) - np.linalg.eigvalsh(A)).max() / sc < 1e-6, (trial, kind, n)
        done += 1
    assert done >= 28


def test_fused_sweep_schedule_equals_sequential_sweeps_for_every_range_pair():
    """k_rotf applies two consecutive QL sweeps in one pass (sweep B one column behind sweep A, out-of-range steps as
    pass-throughs).  Emulation of that schedule (tests/rotf_emulation.py) against the two sweeps applied one after
    the other, for EVERY pair of column ranges of a 9x9 problem — nested, overlapping, disjoint, B above A, B reaching
    column 0 — whether or not the kernel's 85 % rule would pair them."""
    from tests.rotf_emulation import apply_pair, apply_single, would_pair
    d = 9
    rng = np.random.default_rng(0)
    Z0 = rng.standard_normal((5, d))
    n_pairs = n_rule = 0
    for mA in range(1, d):
        for cntA in range(1, mA + 1):
            for mB in range(1, d):
                for cntB in range(1, mB + 1):
                    ang = rng.uniform(-np.pi, np.pi, cntA + cntB)
                    pa = [(np.cos(a), np.sin(a)) for a in ang[:cntA]]
                    pb = [(np.cos(a), np.sin(a)) for a in ang[cntA:]]
                    ref = Z0.copy()
                    apply_single(ref, mA, pa)
                    apply_single(ref, mB, pb)
                    got = Z0.copy()
                    apply_pair(got, mA, pa, mB, pb)
                    assert np.abs(got - ref).max() < 1e-12, (mA, cntA, mB, cntB)
                    n_pairs += 1
                    n_rule += would_pair(mA, cntA, mB, cntB)
    assert n_pairs == 1296 and 0 < n_rule < n_pairs


def test_dc_emulation_tridiagonal_eigensolver_is_accurate_in_fp32():
    """The algorithm of csrc/dc_kernels.cu (fused divide & conquer with every off-diagonal torn, secular roots in
    shifted coordinates, Gu-Eisenstat vectors, xLAED2 deflation) emulated in float32 (tests/dc_emulation.py):
    residual, orthogonality and eigenvalues at rounding level on random, clustered, graded, Wilkinson and degenerate
    tridiagonals - the cases where a fp32 solver without the z re-derivation loses orthogonality."""
    from tests.dc_emulation import check
    rng = np.random.default_rng(0)
    cases = {
        "random": (rng.normal(size=101) * 3, rng.normal(size=100)),
        "wilkinson": (np.abs(np.arange(41) - 20.0), np.ones(40)),
        "graded": (10.0 ** -np.linspace(0, 6, 64), 10.0 ** -np.linspace(0, 6, 64)[:-1] * 0.5),
        "clustered": (np.ones(101) + 1e-5 * rng.normal(size=101), 1e-3 * rng.normal(size=100)),
        "toeplitz": (2 * np.ones(33), -np.ones(32)),
        "some_zero_e": (rng.normal(size=101), rng.normal(size=100) * (rng.random(100) > 0.3)),
        "identity": (np.ones(17), np.zeros(16)),
        "three": (rng.normal(size=3), rng.normal(size=2)),
    }
    for name, (dT, eT) in cases.items():
        res, orth, ev, _ = check(dT, eT)
        assert res < 2e-6 and orth < 2e-6 and ev < 2e-6, (name, res, orth, ev)


def test_dc_closed_form_first_level_equals_the_secular_merge():
    """csrc/dc_kernels.cu solves the first merge level (2 x 2 blocks) by one Jacobi rotation instead of the secular
    machinery (tests/dc_emulation.py::merge2_closed_form): same eigenpairs as the generic merge, including negative and
    zero off-diagonals and equal diagonals, and the full solver agrees with and without the shortcut."""
    from tests.dc_emulation import dc_eigh
    rng = np.random.default_rng(3)
    cases = [(rng.normal(size=2) * 3, rng.normal(size=1)) for _ in range(50)]
    cases += [(np.array([1.0, 1.0]), np.array([0.5])), (np.array([1.0, 1.0]), np.array([-0.5])),
              (np.array([2.0, -1.0]), np.array([0.0])), (np.array([1.0, 1.0 + 1e-6]), np.array([1e-3]))]
    for dT, eT in cases:
        T = np.array([[dT[0], eT[0]], [eT[0], dT[1]]], dtype=np.float64)
        for cf in (True, False):
            lam, Q = dc_eigh(dT, eT, closed_form_level1=cf)
            assert np.abs(T @ Q - Q * lam).max() < 2e-6 * max(1.0, np.abs(T).max()), (dT, eT, cf)
            assert np.abs(Q.T @ Q - np.eye(2)).max() < 5e-7
    dT, eT = rng.normal(size=37) * 2, rng.normal(size=36)
    l1, Q1 = dc_eigh(dT, eT, closed_form_level1=True)
    l0, Q0 = dc_eigh(dT, eT, closed_form_level1=False)
    assert np.abs(np.sort(l1) - np.sort(l0)).max() < 5e-6
