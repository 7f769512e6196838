"""GPU parity tests (run on the B200 box): CUDA path, called through the C ABI (ctypes), against the
golden vectors of the reference and against the CPU oracle on the same seeded inputs."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from tests.helpers import GOLDEN, load_net_case, parse_opts, rel_err

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# BASELINE.json north_star: 1e-4 relative on the recovered phi (max-norm per signal)
PHI_TOL = 1e-4


@pytest.fixture(scope="module")
def pkg():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import admmnet_b200
    admmnet_b200._capi.lib()          # raises if the extension is missing: no silent fallback
    return admmnet_b200


def _eigh(pkg, A, params=None, vecs=True, fn=False):
    from admmnet_b200 import _capi
    L = _capi.lib()
    dev = torch.device("cuda")
    B, d, _ = A.shape
    nb = C.c_size_t()
    _capi.check(L.admmnet_eigh_workspace_bytes(B, d, 0, C.byref(nb)))
    ws = torch.empty(nb.value, dtype=torch.uint8, device=dev)
    ev = torch.empty(B, d, dtype=torch.float32, device=dev)
    U = torch.empty(B, d, d, dtype=torch.complex64, device=dev) if vecs else None
    G = torch.empty(B, d * (d + 1) // 2, dtype=torch.complex64, device=dev) if fn else None
    st = torch.zeros(1, dtype=torch.int32, device=dev)
    Ad = A.to(dev).contiguous()
    _capi.check(L.admmnet_eigh_batched(Ad.data_ptr(), B, d, ev.data_ptr(), U.data_ptr() if vecs else None,
                                       G.data_ptr() if fn else None, params.data_ptr() if params is not None else None,
                                       ws.data_ptr(), nb.value, 0, torch.cuda.current_stream().cuda_stream,
                                       st.data_ptr()))
    torch.cuda.synchronize()
    assert int(st.item()) == 0
    return ev.cpu(), U.cpu() if vecs else None, G.cpu() if fn else None


def _unpack(Gp, d):
    B = Gp.shape[0]
    G = torch.zeros(B, d, d, dtype=Gp.dtype)
    i, j = torch.tril_indices(d, d)
    G[:, i, j] = Gp
    G[:, j, i] = Gp.conj()
    G[:, torch.arange(d), torch.arange(d)] = Gp[:, (torch.arange(d) * (torch.arange(d) + 3)) // 2]
    return G


@pytest.mark.parametrize("d", [3, 4, 17, 65, 101, 104, 105, 128])
def test_eigh_random_hermitian(pkg, d):
    torch.manual_seed(d)
    X = torch.randn(5, d, d, dtype=torch.complex64)
    A = 0.5 * (X + X.transpose(1, 2).conj())
    ev, U, _ = _eigh(pkg, A)
    scale = torch.linalg.matrix_norm(A, ord=2).max().item()
    assert (A @ U - U * ev.unsqueeze(1).to(torch.complex64)).abs().amax() < 2e-5 * scale
    assert (U.transpose(1, 2).conj() @ U - torch.eye(d)).abs().amax() < 1e-5
    w = torch.linalg.eigvalsh(A.to(torch.complex128)).float()
    assert (ev.sort(dim=1)[0] - w).abs().amax() < 1e-5 * scale


@pytest.mark.parametrize("d", [129, 145, 197, 257])
def test_eigh_jacobi_large_orders(pkg, d):
    """Matrix orders above 128 (cfg 4: n = 144, 196, 256) run the cyclic two-sided Jacobi kernel (csrc/big_kernels.cu):
    eigenpairs against torch fp64, and f(A) through the eigenvalue map against the oracle's formula."""
    from oracle import net_oracle
    torch.manual_seed(d)
    X = torch.randn(3, d, d, dtype=torch.complex64) * (3.0 / d ** 0.5)
    A = 0.5 * (X + X.transpose(1, 2).conj())
    A[2] = torch.diag(torch.linspace(-3, 5, d)).to(torch.complex64)          # already diagonal: zero rotations
    net = pkg.PhiEstADMMNet(10, 10, 3, 2)
    from admmnet_b200.params import pack_state_dict
    P = pack_state_dict(net.state_dict(), 100, 2).cuda()
    ev, U, G = _eigh(pkg, A, params=P[0], fn=True)
    scale = torch.linalg.matrix_norm(A, ord=2).max().item()
    assert (A @ U - U * ev.unsqueeze(1).to(torch.complex64)).abs().amax() < 2e-5 * scale
    assert (U.transpose(1, 2).conj() @ U - torch.eye(d)).abs().amax() < 2e-5
    w, Q = torch.linalg.eigh(A.to(torch.complex128))
    assert (ev.sort(dim=1)[0] - w.float()).abs().amax() < 1e-5 * scale
    lp = net_oracle.eig_map(w.float(), net_oracle.layer_params(net.state_dict(), 0)["g"]).to(torch.complex128)
    ref = (Q * lp.unsqueeze(1)) @ Q.transpose(1, 2).conj()
    assert (_unpack(G, d).to(torch.complex128) - ref).abs().amax() < 3e-5 * max(1.0, ref.abs().amax().item())


def test_eigh_structured_cases(pkg):
    d = 101
    torch.manual_seed(11)
    eye = torch.eye(d, dtype=torch.complex64)
    diag = torch.diag(torch.linspace(-3, 5, d)).to(torch.complex64)
    arrow = torch.diag(torch.full((d,), 0.01)).to(torch.complex64)          # layer-0 matrix: arrowhead
    v = torch.randn(d - 1, dtype=torch.complex64)
    arrow[:-1, -1] = v
    arrow[-1, :-1] = v.conj()
    arrow[-1, -1] = 1.8
    rank1 = torch.outer(v.new_ones(d), v.new_ones(d))                       # repeated zero eigenvalue
    A = torch.stack([eye, diag, arrow, rank1, torch.zeros(d, d, dtype=torch.complex64)])
    ev, U, _ = _eigh(pkg, A)
    assert (A @ U - U * ev.unsqueeze(1).to(torch.complex64)).abs().amax() < 2e-4
    assert (U.transpose(1, 2).conj() @ U - torch.eye(d)).abs().amax() < 2e-5      # fp32: a few d*eps (d*eps = 6e-6)
    w = torch.linalg.eigvalsh(A.to(torch.complex128)).float()
    assert (ev.sort(dim=1)[0] - w).abs().amax() < 2e-5 * 101


def _arrow_eigh(h, phi, c0):
    import ctypes as C
    from admmnet_b200 import _capi
    L = _capi.lib()
    B, n = h.shape
    hd, pd, cd = h.cuda().contiguous(), phi.cuda().contiguous(), c0.cuda().contiguous()
    ev = torch.zeros(B, n + 1, dtype=torch.float32, device="cuda")
    U = torch.zeros(B, n + 1, n + 1, dtype=torch.complex64, device="cuda")
    ok = torch.full((B,), -1, dtype=torch.int32, device="cuda")
    _capi.check(L.admmnet_arrow_eigh(hd.data_ptr(), pd.data_ptr(), cd.data_ptr(), B, n, ev.data_ptr(), U.data_ptr(),
                                     ok.data_ptr(), torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    return ev.cpu(), U.cpu(), ok.cpu()


def _arrow_dense(h, phi, c0):
    B, n = h.shape
    A = torch.zeros(B, n + 1, n + 1, dtype=torch.complex64)
    A[:, :n, :n] = torch.diag_embed(h).to(torch.complex64)
    A[:, :n, n] = phi
    A[:, n, :n] = phi.conj()
    A[:, n, n] = c0.to(torch.complex64)
    return A


@pytest.mark.parametrize("n", [2, 5, 33, 100, 127])
def test_arrowhead_shortcut_eigenpairs(pkg, n):
    """Layer-0 shortcut (App. A.2): eigenpairs of [[diag(h), phi],[phi^H, c0]] against torch fp64 eigh."""
    torch.manual_seed(n)
    B = 8
    h = torch.randn(B, n) * 0.1
    phi = torch.randn(B, n, dtype=torch.complex64)
    c0 = torch.full((B,), 1.8)
    h[1] = torch.sort(torch.rand(n))[0] * 1e-3                       # tight cluster of poles
    h[2, : n // 2] = h[2, n // 2: 2 * (n // 2)] + 3e-7             # pairs of nearly equal poles
    phi[3] *= 1e-3                                                  # weak coupling: roots hug the poles
    phi[4, ::2] *= 1e-4                                             # mixed magnitudes
    c0[5] = -7.0                                                    # corner element below every pole
    h[6] *= 50.0
    ev, U, ok = _arrow_eigh(h, phi, c0)
    assert ok.tolist() == [1] * B
    A = _arrow_dense(h, phi, c0)
    scale = torch.linalg.matrix_norm(A, ord=2).view(B, 1, 1)
    assert ((A @ U - U * ev.unsqueeze(1).to(torch.complex64)).abs() / scale).amax() < 4e-6
    assert (U.transpose(1, 2).conj() @ U - torch.eye(n + 1)).abs().amax() < 4e-6
    w = torch.linalg.eigvalsh(A.to(torch.complex128)).float()
    assert ((ev - w).abs() / scale.view(B, 1)).amax() < 1e-6          # ascending order, like torch


def test_arrowhead_shortcut_declines_degenerate_input(pkg):
    n = 100
    torch.manual_seed(0)
    h = torch.randn(4, n) * 0.1
    phi = torch.randn(4, n, dtype=torch.complex64)
    c0 = torch.full((4,), 1.8)
    h[0, 7] = h[0, 3]                    # repeated pole
    phi[1, 11] = 0                       # decoupled row
    h[2, 5] = float("nan")
    ev, U, ok = _arrow_eigh(h, phi, c0)
    assert ok.tolist() == [0, 0, 0, 1]


def test_forward_general_path_at_layer0_when_shortcut_declines(pkg):
    """A signal whose y has a zero entry makes phi_i = 0 at layer 0: k_arrow declines, the dense solver takes it, and
    the result still matches the oracle; the other signals of the batch go through the shortcut."""
    from oracle import net_oracle
    z, sd = load_net_case("pert_k5")
    K = int(z["K"])
    net = pkg.PhiEstADMMNet(10, 10, 3, K).eval()
    net.load_state_dict(sd)
    y, b, s = (torch.from_numpy(z[k]).clone() for k in ("y", "b", "sigma"))
    y[2, 17] = 0
    got = net(y, b, s).detach().numpy()      # grad enabled: the result carries a grad_fn, like the reference's
    ref = net_oracle.forward(sd, y, b, s, 10, 10, K).numpy()
    assert rel_err(got, ref).max() < 1e-4


def test_hermitian_function_matches_oracle(pkg):
    """f(A) = U f(L) U^H with the learned eigenvalue map: basis-invariant, compared with torch fp64."""
    from admmnet_b200.params import pack_state_dict
    from oracle import net_oracle
    z, sd = load_net_case("pert_k10")
    P = pack_state_dict(sd, 100, 10).cuda()
    torch.manual_seed(1)
    d = 101
    X = torch.randn(4, d, d, dtype=torch.complex64) * 0.3
    A = 0.5 * (X + X.transpose(1, 2).conj())
    _, _, Gp = _eigh(pkg, A, params=P[3], vecs=False, fn=True)
    G = _unpack(Gp, d)
    w, U = torch.linalg.eigh(A.to(torch.complex128))
    fw = net_oracle.eig_map(w.float(), net_oracle.layer_params(sd, 3)["g"]).to(torch.complex128)
    Gt = (U * fw.unsqueeze(1)) @ U.transpose(1, 2).conj()
    assert (G.to(torch.complex128) - Gt).abs().amax() / Gt.abs().amax() < 2e-5


@pytest.mark.parametrize("d,B", [(101, 700), (104, 150), (97, 149), (72, 310), (65, 40), (40, 333), (33, 5)])
def test_tensor_core_tail_kernel_sizes_and_persistence(pkg, d, B):
    """k_tail_tc (tcgen05/TMEM back-transformation + rebuild, Z^T by TMA) through the f(A) tap: every block plan
    (4 WY blocks + 4 thread-local reflectors at d = 101; padded last block at d = 65, 40; one block + 8 local at
    d = 33), batches larger than the persistent grid (several signals per CTA, prefetch double buffering) and both
    tridiagonal solvers in front of it (B > 1024 would be plain QL; here divide & conquer)."""
    from admmnet_b200 import _capi
    from admmnet_b200.params import pack_state_dict
    from oracle import net_oracle
    assert _capi.lib().admmnet_tail_tc_smem_bytes(d) > 0, "tensor-core tail kernel not selected"
    z, sd = load_net_case("pert_k10")
    P = pack_state_dict(sd, 100, 10).cuda()
    g = torch.Generator().manual_seed(d * 1000 + B)
    X = torch.randn(B, d, d, dtype=torch.complex64, generator=g) * (3.0 / d ** 0.5)
    A = 0.5 * (X + X.transpose(1, 2).conj())
    _, _, Gp = _eigh(pkg, A, params=P[3], vecs=False, fn=True)
    G = _unpack(Gp, d)
    w, U = torch.linalg.eigh(A.to(torch.complex128))
    fw = net_oracle.eig_map(w.float(), net_oracle.layer_params(sd, 3)["g"]).to(torch.complex128)
    Gt = (U * fw.unsqueeze(1)) @ U.transpose(1, 2).conj()
    err = (G.to(torch.complex128) - Gt).abs().amax(dim=(1, 2)) / Gt.abs().amax(dim=(1, 2))
    assert err.max() < 2e-5, (float(err.max()), int(err.argmax()))


@pytest.mark.parametrize("tag", ["init_k10", "pert_k10", "pert_k5"])
def test_forward_matches_reference_golden(pkg, tag):
    z, sd = load_net_case(tag)
    K = int(z["K"])
    net = pkg.PhiEstADMMNet(10, 10, 3, K).eval()
    net.load_state_dict(sd)
    y, b, s = (torch.from_numpy(z[k]) for k in ("y", "b", "sigma"))
    with torch.no_grad():
        phi = net(y, b, s)
    assert phi.dtype == torch.complex64 and phi.device.type == "cpu" and phi.shape == (7, 100)
    assert rel_err(phi.numpy(), z["phi_batch"]).max() < PHI_TOL
    # the data.npz signal alone with sigma [1,1] (main_for_net.py:93): config 1 of BASELINE.json
    with torch.no_grad():
        phi1 = net(y[:1], b[:1], s[:1].reshape(1, 1))
    assert rel_err(phi1.numpy(), z["phi_single"]).max() < PHI_TOL
    # every prefix depth (phi after layer k) against the reference's per-layer taps
    for kk in range(1, K):
        sub = pkg.PhiEstADMMNet(10, 10, 3, kk)
        sub.load_state_dict({k_: v for k_, v in sd.items() if int(k_.split(".")[1]) < kk})
        with torch.no_grad():
            pk = sub(y, b, s).numpy()
        assert rel_err(pk, z["batch_phi_layers"][kk - 1]).max() < PHI_TOL, kk


def test_forward_vs_oracle_seeded_batch_and_scopes(pkg):
    from oracle import net_oracle, signals
    torch.manual_seed(3)
    net = pkg.PhiEstADMMNet(10, 10, 3, 10).eval()
    y, b, s, _ = signals.generate(96, seed=17)
    yt, bt, st = (torch.from_numpy(a) for a in (y, b, s))
    ref = net_oracle.forward(net.state_dict(), yt, bt, st, 10, 10, 10)
    with torch.no_grad():
        phi = net(yt.cuda(), bt.cuda(), st.cuda())
    assert phi.is_cuda
    assert rel_err(phi.cpu().numpy(), ref.numpy()).max() < PHI_TOL
    # exact whole-batch semantics must not depend on the scratch chunking
    net.chunk = 40
    with torch.no_grad():
        phi_c = net(yt, bt, st)
    assert rel_err(phi_c.numpy(), ref.numpy()).max() < PHI_TOL
    # (one chunk: fused divide & conquer tridiagonal solver, several chunks: the QL pair - two fp32 solvers agree to
    #  the forward's rounding floor, DESIGN.md accuracy budget, not to 1e-5)
    assert rel_err(phi_c.numpy(), phi.cpu().numpy()).max() < 3e-5
    # norm_scope='chunk': independent chunks == the reference run per chunk
    net.norm_scope = "chunk"
    net.chunk = 32
    ref_c = net_oracle.forward(net.state_dict(), yt, bt, st, 10, 10, 10, chunk=32)
    with torch.no_grad():
        phi_s = net(yt, bt, st)
    assert rel_err(phi_s.numpy(), ref_c.numpy()).max() < PHI_TOL
    assert rel_err(ref_c.numpy(), ref.numpy()).max() > PHI_TOL      # the coupling is material


def test_ql_and_divide_and_conquer_paths_agree(pkg):
    """A call of ONE chunk solves the tridiagonal eigenproblem with the fused fp32 divide & conquer kernel (k_dc,
    csrc/dc_kernels.cu), a call of several chunks with the QL pair k_ql + k_rotf (capi.cu::launch_eig_tail).  Same
    batch, same whole-batch norm scope, two scratch chunk sizes -> both paths; they must agree with each other and
    with the oracle."""
    from oracle import net_oracle, signals
    torch.manual_seed(11)
    net = pkg.PhiEstADMMNet(10, 10, 3, 6).eval()
    y, b, s, _ = signals.generate(1280, seed=29)
    yt, bt, st = (torch.from_numpy(a) for a in (y, b, s))
    net.chunk = 4096                                   # one launch of 1280 signals: k_dc
    with torch.no_grad():
        phi_ql = net(yt, bt, st).numpy()
    net.chunk = 320                                    # four launches of 320 signals: k_ql + k_rotf
    with torch.no_grad():
        phi_dc = net(yt, bt, st).numpy()
    assert rel_err(phi_dc, phi_ql).max() < 3e-5         # (the names follow the chunk sizes above the other way round)
    ref = net_oracle.forward(net.state_dict(), yt, bt, st, 10, 10, 6).numpy()
    assert rel_err(phi_ql, ref).max() < PHI_TOL
    assert rel_err(phi_dc, ref).max() < PHI_TOL


def test_benchmarked_path_multichunk_lanes_against_oracle(pkg):
    """the configuration bench.py times — several scratch chunks of MORE than 1024 signals
    (plain-QL path, not divide & conquer), chunk lanes on, the persistent tail kernel on both scratch slots, whole-
    batch norm scope — against the oracle run on the whole batch.  6000 signals = 3 chunks of 2000 over 2 lanes
    (slot 0 twice, slot 1 once), perturbed weights, K = 10.  Several chunks: the QL pair solves the tridiagonal problems."""
    from oracle import net_oracle, signals
    z, sd = load_net_case("pert_k10")
    net = pkg.PhiEstADMMNet(10, 10, 3, 10).eval()
    net.load_state_dict(sd)
    y, b, s, _ = signals.generate(6000, seed=41)
    yt, bt, st = (torch.from_numpy(a) for a in (y, b, s))
    net.chunk = 2000
    with torch.no_grad():
        phi = net(yt.cuda(), bt.cuda(), st.cuda()).cpu().numpy()
    torch.set_num_threads(os.cpu_count() or 1)
    ref = net_oracle.forward(sd, yt, bt, st, 10, 10, 10).numpy()
    err = rel_err(phi, ref)
    assert err.max() < PHI_TOL, (err.max(), int(err.argmax()))
    # ragged last chunk on the other slot: 4500 = 2000 + 2000 + 500
    with torch.no_grad():
        phi2 = net(yt[:4500].cuda(), bt[:4500].cuda(), st[:4500].cuda()).cpu().numpy()
    ref2 = net_oracle.forward(sd, yt[:4500], bt[:4500], st[:4500], 10, 10, 10).numpy()
    assert rel_err(phi2, ref2).max() < PHI_TOL


def test_forward_is_reentrant_across_host_threads_and_streams(pkg):
    """include/admmnet_b200.h: admmnet_forward is re-entrant across streams given distinct workspaces.  Two host
    threads, two models (own workspaces), two CUDA streams, different inputs, run concurrently several times."""
    import threading
    from oracle import net_oracle, signals
    z, sd = load_net_case("pert_k10")
    data, refs, outs, errs = [], [], [None, None], []
    for t in range(2):
        y, b, s, _ = signals.generate(1500, seed=100 + t)
        data.append(tuple(torch.from_numpy(a).cuda() for a in (y, b, s)))
        refs.append(net_oracle.forward(sd, *(torch.from_numpy(a) for a in (y, b, s)), 10, 10, 10).numpy())
    nets = []
    for t in range(2):
        net = pkg.PhiEstADMMNet(10, 10, 3, 10).eval()
        net.load_state_dict(sd)
        net.chunk = 600                     # 3 chunks -> both lanes of each caller stream are busy
        nets.append(net)
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    torch.cuda.synchronize()

    def work(t):
        try:
            with torch.cuda.stream(streams[t]), torch.no_grad():
                for _ in range(3):
                    out = nets[t].forward_device(*data[t])
                streams[t].synchronize()
            outs[t] = out.cpu().numpy()
        except Exception as e:              # pragma: no cover
            errs.append(e)

    th = [threading.Thread(target=work, args=(t,)) for t in range(2)]
    for x in th:
        x.start()
    for x in th:
        x.join()
    assert not errs, errs
    for t in range(2):
        assert rel_err(outs[t], refs[t]).max() < PHI_TOL


def test_forward_other_shapes_and_edges(pkg):
    from oracle import net_oracle, signals
    torch.manual_seed(5)
    # (12,12), (14,14), (16,16): BASELINE.json configs[3]'s n = 144, 196, 256 (matrix order > 128: the Jacobi layer kernel)
    for (M, N, K, B) in [(8, 8, 3, 5), (4, 6, 4, 3), (10, 10, 1, 4), (10, 10, 2, 1), (11, 11, 3, 2), (12, 12, 4, 3),
                         (14, 14, 3, 2), (16, 16, 3, 2), (12, 13, 10, 2)]:
        net = pkg.PhiEstADMMNet(M, N, 3, K).eval()
        y, b, s, _ = signals.generate(B, Nb=M, Nd=N, seed=M * 100 + N)
        yt, bt, st = (torch.from_numpy(a) for a in (y, b, s))
        ref = net_oracle.forward(net.state_dict(), yt, bt, st, M, N, K)
        with torch.no_grad():
            phi = net(yt, bt, st)
        assert rel_err(phi.numpy(), ref.numpy()).max() < PHI_TOL, (M, N, K)
    net = pkg.PhiEstADMMNet(10, 10, 3, 2)
    with pytest.raises(ValueError):
        net(torch.zeros(2, 99, dtype=torch.complex64), torch.zeros(2, 99, dtype=torch.complex64), torch.ones(2))
    with pytest.raises(Exception):
        pkg.PhiEstADMMNet(17, 17, 3, 2)(torch.ones(1, 289, dtype=torch.complex64),
                                       torch.ones(1, 289, dtype=torch.complex64), torch.ones(1))   # n > 256


def test_linearity_free_property_large_batch(pkg):
    """Size-independent property at a bench-sized batch: with norm_scope='chunk' the result for a signal
    depends only on its own chunk, so permuting whole chunks permutes the output bit-exactly, and
    duplicated signals give duplicated outputs."""
    from oracle import signals
    torch.manual_seed(0)
    net = pkg.PhiEstADMMNet(10, 10, 3, 10).eval()
    net.norm_scope, net.chunk = "chunk", 256
    y, b, s, _ = signals.generate(256, seed=1)
    yt = torch.from_numpy(np.tile(y, (8, 1))).cuda()
    bt = torch.from_numpy(np.tile(b, (8, 1))).cuda()
    st = torch.from_numpy(np.tile(s, 8)).cuda()
    with torch.no_grad():
        phi = net(yt, bt, st)
    phi = phi.reshape(8, 256, 100)
    assert torch.isfinite(phi.real).all()
    for c in range(1, 8):
        assert torch.equal(phi[c], phi[0])


def test_classic_matches_reference_golden_and_oracle(pkg):
    from oracle import classic_oracle, signals
    z = np.load(os.path.join(GOLDEN, "classic.npz"))
    for c in range(len(z["iters"])):
        i = int(z["case_sig"][c])
        opts = dict(rho=float(z["case_rho"][c]), max_iter=int(z["case_max_iter"][c]))
        phi, it = pkg.admm_for_us(z["y"][i], z["b"][i], 10, 10, 1.0, float(z["sigma"][i]), opts,
                                  bool(z["case_use_min_iter"][c]), int(z["case_min_iter"][c]))
        assert it == int(z["iters"][c])
        assert phi.dtype == np.complex128 and phi.shape == (100,)
        assert rel_err(phi, z["phi"][c]) < 1e-12
    # complex64 input with even n <= 104 takes the persistent bulk-copy kernel (k_classic_p: tiles of 32 signals, ragged
    # last tile, more tiles than CTAs), everything else the plain kernels
    for n, B, dt in [(100, 1000, torch.complex128), (100, 257, torch.complex64), (256, 33, torch.complex128),
                     (7, 5, torch.complex64), (64, 100, torch.complex64), (104, 33, torch.complex64),
                     (36, 4097, torch.complex64), (100, 31, torch.complex64), (100, 20011, torch.complex64),
                     (2, 3, torch.complex64), (106, 40, torch.complex64)]:
        y, b, _, _ = signals.generate(B, Nb=n, Nd=1, seed=n)
        yt, bt = torch.from_numpy(y).to(dt).cuda(), torch.from_numpy(b).to(dt).cuda()
        out = pkg.admm_for_us_batched(yt, bt, rho=0.7, n_iter=5).cpu().numpy()
        ref = classic_oracle.admm_linear_recursion(yt.cpu().numpy(), bt.cpu().numpy(), 0.7, 5)
        assert rel_err(out, ref).max() < 1e-12


def test_classic_honours_zero_tolerances(pkg):
    """admm.py:95-112 with eta_abs = eta_rel = 0 never stops early: max_iter iterations."""
    from oracle import classic_oracle, signals
    y, b, s, _ = signals.generate(1, seed=4)
    opts = {"rho": 1.0, "max_iter": 9, "eta_abs": 0.0, "eta_rel": 0.0}
    phi, it = pkg.admm_for_us(y[0], b[0], 10, 10, 0.1, float(s[0]), opts)
    ref, it_ref = classic_oracle.admm_for_us(y[0].astype(np.complex128), b[0].astype(np.complex128), 10, 10, 0.1,
                                             float(s[0]), opts)
    assert it == it_ref == 9
    assert np.abs(phi - ref).max() < 1e-12 * np.abs(ref).max()
    phi5, it5 = pkg.admm_for_us(y[0], b[0], 10, 10, 0.1, float(s[0]), {"max_iter": 9})
    assert it5 == 5


def test_peak_search_matches_reference_golden(pkg):
    z = np.load(os.path.join(GOLDEN, "peaks.npz"))
    names = sorted({k.split("__")[0] for k in z.files if "__" in k})
    for name in names:
        opts = parse_opts(z[f"{name}__opts"])
        got = pkg.alt_peak_search({"phi": z[f"{name}__phi"], "xbase": 10, "ybase": 10}, opts)
        exp = z[f"{name}__peaks"]
        assert got.shape == exp.shape and got.dtype == np.float64, name
        assert np.array_equal(got[:, :2], exp[:, :2]), name          # grid positions: bit exact
        np.testing.assert_allclose(got[:, 2], exp[:, 2], rtol=1e-12, atol=1e-300)
    ax = np.arange(0, 1 - 0.01, 0.01)
    ay = np.arange(-0.5, 0.5 - 0.01, 0.01)
    AX, AY = np.meshgrid(ax, ay)
    surf = pkg.peak_search(z["net0__phi"], AX, 10, AY, 10)
    np.testing.assert_allclose(surf, z["surface_net0"], rtol=1e-11, atol=1e-18)
    v = pkg.peak_search_func(z["net0__phi"], 0.3, 10, -0.1, 10)
    from oracle import peak_oracle
    np.testing.assert_allclose(v, peak_oracle.peak_search_func(z["net0__phi"], 0.3, 10, -0.1, 10), rtol=1e-11)


def test_peak_search_batched_vs_oracle(pkg):
    from oracle import peak_oracle
    z, sd = load_net_case("init_k10")
    phis = z["phi_batch"]
    opts = dict(xstep=0.02, ystep=0.02, iter=2)
    r = pkg.alt_peak_search_batched(phis, 10, 10, opts, topl=3, pmax=8)       # pmax too small on purpose: grows
    sep = lambda p, X, xb, Y, yb: peak_oracle.peak_search_separable(p, X[0], xb, Y[:, 0], yb)
    for i in range(len(phis)):
        exp = peak_oracle.alt_peak_search({"phi": phis[i], "xbase": 10, "ybase": 10}, opts, sep)
        P = int(r["count"][i])
        assert P == len(exp)
        got = r["peaks"][i, :P].cpu().numpy()
        assert np.array_equal(got[:, :2], exp[:, :2])
        np.testing.assert_allclose(got[:, 2], exp[:, 2], rtol=1e-11)
        np.testing.assert_array_equal(r["top"][i].cpu().numpy()[:, :2], peak_oracle.top_l(exp, 3)[:, :2])
    # other dictionary sizes (config 4 of BASELINE.json: n in {64..256}, coarse grids 32^2..64^2)
    rng = np.random.default_rng(0)
    for nb, step in [(8, 1 / 32), (12, 1 / 45), (16, 1 / 64), (14, 1 / 32)]:
        phi = (rng.normal(size=nb * nb) + 1j * rng.normal(size=nb * nb)).astype(np.complex64)
        o = dict(xstep=step, ystep=step, iter=1)
        got = pkg.alt_peak_search({"phi": phi, "xbase": nb, "ybase": nb}, o)
        exp = peak_oracle.alt_peak_search({"phi": phi, "xbase": nb, "ybase": nb}, o, sep)
        assert got.shape == exp.shape and np.array_equal(got[:, :2], exp[:, :2]), nb
        np.testing.assert_allclose(got[:, 2], exp[:, 2], rtol=1e-10)


def test_peak_plateau_and_degenerate_images(pkg):
    # phi = 0 -> constant surface -> no maxima ; phi = e_0 -> constant 1 up to rounding (reference self-test shape)
    out = pkg.alt_peak_search({"phi": np.zeros(100, dtype=np.complex64), "xbase": 10, "ybase": 10}, dict(iter=1))
    assert out.shape == (0, 3)
    out = pkg.alt_peak_search({"phi": np.ones(1, dtype=np.complex128), "xbase": 1, "ybase": 1},
                              dict(xstep=0.1, ystep=0.1))
    assert out.shape == (0, 3)                                  # n=1: |phi|^2 everywhere, one big plateau touching the border


def _surface_maxima(pkg, img):
    from admmnet_b200 import _capi
    L = _capi.lib()
    dev = torch.device("cuda")
    img = np.ascontiguousarray(img, dtype=np.float64)
    B, Gy, Gx = img.shape
    surf = torch.from_numpy(img).to(dev)
    ax = torch.arange(Gx, dtype=torch.float64, device=dev)
    ay = torch.arange(Gy, dtype=torch.float64, device=dev)
    pmax = Gx * Gy
    peaks = torch.zeros(B, pmax, 3, dtype=torch.float64, device=dev)
    cnt = torch.zeros(B, dtype=torch.int32, device=dev)
    st = torch.zeros(1, dtype=torch.int32, device=dev)
    _capi.check(L.peak_surface_maxima(surf.data_ptr(), B, ax.data_ptr(), Gx, ay.data_ptr(), Gy, pmax, peaks.data_ptr(),
                                      cnt.data_ptr(), st.data_ptr(), torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    out = []
    for i in range(B):
        m = np.zeros((Gy, Gx), dtype=bool)
        p = peaks[i, :int(cnt[i])].cpu().numpy()
        m[p[:, 1].astype(int), p[:, 0].astype(int)] = True
        out.append((m, p))
    return out


def test_local_maximum_stage_on_arbitrary_surfaces(pkg):
    """the CUDA local-maximum stage fed with surfaces of our choosing — the reference's own
    hand-checkable plateau KAT (peakSearchUtils.py:427-436), tie cases, border plateaus, and random images with many
    exact ties against the oracle's restatement of skimage.morphology.local_maxima."""
    from oracle import peak_oracle
    kat = np.array([[1, 1, 1, 2, 3], [1, 5, 5, 4, 3], [2, 5, 5, 4, 2], [3, 4, 4, 3, 1]], dtype=np.float64)
    want = np.zeros((4, 5), dtype=bool)
    want[1:3, 1:3] = True                                      # the 2x2 plateau of fives, and nothing else
    rng = np.random.default_rng(0)
    cases = [kat,
             np.array([[3, 3, 1, 0, 0], [3, 3, 1, 0, 2], [1, 1, 1, 0, 2], [0, 0, 0, 0, 2]], dtype=np.float64),   # border plateaus
             np.array([[1, 2, 3, 4, 5], [1, 2, 3, 4, 5], [1, 2, 3, 4, 5], [1, 2, 3, 4, 5]], dtype=np.float64),   # ridge on the border
             np.array([[2, 2, 2, 2, 2], [2, 1, 1, 1, 2], [2, 1, 3, 1, 2], [2, 2, 2, 2, 2]], dtype=np.float64),   # ring plateau + centre
             np.array([[1, 1, 1, 1, 1], [1, 2, 2, 2, 1], [1, 2, 3, 2, 1], [1, 1, 1, 1, 1]], dtype=np.float64)]   # plateau below a peak
    got = _surface_maxima(pkg, np.stack(cases))
    assert np.array_equal(got[0][0], want)
    for (m, _), img in zip(got, cases):
        assert np.array_equal(m, peak_oracle.local_maxima(img, connectivity=2)), img
    # row-major discovery order (np.where) and the values
    m, p = got[0]
    assert [tuple(r) for r in p[:, :2].astype(int)] == [(1, 1), (2, 1), (1, 2), (2, 2)] and np.all(p[:, 2] == 5.0)
    # random images drawn from 4 levels: plateaus and ties everywhere
    imgs = rng.integers(0, 4, size=(40, 13, 17)).astype(np.float64)
    for (m, _), img in zip(_surface_maxima(pkg, imgs), imgs):
        assert np.array_equal(m, peak_oracle.local_maxima(img, connectivity=2))


def test_end_to_end_recovers_targets(pkg):
    """Physical sanity on the classical path (main.py:95-120): the three strongest peaks sit on the true (tau,f)."""
    from oracle import signals
    y, b, s, truth = signals.generate(4, seed=123, snr_demod=30.0)
    for i in range(4):
        phi, _ = pkg.admm_for_us(y[i].astype(np.complex128), b[i].astype(np.complex128), 10, 10, 1.0, float(s[i]))
        pk = pkg.alt_peak_search({"phi": phi, "xbase": 10, "ybase": 10}, dict(xstep=0.01, ystep=0.01, iter=3))
        top = sorted(pk, key=lambda p: p[2], reverse=True)[:1]
        amp = np.abs(truth["C"][i])
        j = int(np.argmax(amp))
        assert abs(top[0][0] - truth["tau"][i][j]) < 0.05 and abs(top[0][1] - truth["f"][i][j]) < 0.05


def test_device_signal_generator(pkg):
    """admmnet_generate follows generate_data.py:133-221: structural identities + statistics + determinism."""
    B, Nb, Nd, L = 4096, 10, 10, 3
    y, b, s, truth = pkg.generate_signals(B, Nb, Nd, L, snr_w=20.0, snr_demod=7.0, seed=99, return_truth=True)
    y2, b2, s2 = pkg.generate_signals(64, Nb, Nd, L, snr_w=20.0, snr_demod=7.0, seed=99)
    assert torch.equal(y[:64], y2) and torch.equal(b[:64], b2) and torch.equal(s[:64], s2)     # counter-based RNG
    y3, _, _ = pkg.generate_signals(64, Nb, Nd, L, seed=100)
    assert not torch.equal(y3, y2)
    y, b, s, truth = y.cpu().numpy(), b.cpu().numpy(), s.cpu().numpy().astype(np.float64), truth.cpu().numpy()
    # b on the QPSK constellation with phase offset pi/4 (generate_data.py:210-218)
    np.testing.assert_allclose(np.abs(b), 1.0, atol=1e-6)
    k = (np.angle(b) - np.pi / 4) / (np.pi / 2)
    np.testing.assert_allclose(k, np.round(k), atol=1e-5)
    # sigma = ||e/b|| + 1 with |e|^2 in {0, 2, 4} per symbol; symbol error rate of QPSK at 7 dB is a few per cent
    q = (s - 1.0) ** 2 / 2.0
    np.testing.assert_allclose(q, np.round(q), atol=1e-3)
    assert 0.5 < q.mean() < 8.0
    # parameter ranges (generate_data.py:30-31, 142-144)
    tau, f, C = truth[..., 0], truth[..., 1], truth[..., 2] + 1j * truth[..., 3]
    assert tau.min() >= 0.1 and tau.max() <= 0.9 and f.min() >= -0.4 and f.max() <= 0.4
    assert abs(C.real.std() - 0.7) < 0.03 and abs(C.imag.std() - 0.7) < 0.03 and abs(C.mean()) < 0.05
    # y = (b+e) Psi + w: on error-free symbols the residual y - b Psi is the noise w at SNR 20 dB
    p_idx, q_idx = np.divmod(np.arange(Nb * Nd), Nd)
    Psi = np.einsum("bl,blp,blq->bpq", C, np.exp(2j * np.pi * f[..., None] * np.arange(Nb)),
                    np.exp(-2j * np.pi * tau[..., None] * np.arange(Nd))).reshape(B, -1)
    res = np.abs(y - b * Psi) ** 2
    sig_pow = (np.abs(Psi) ** 2).mean(axis=1)
    ratio = np.median(res, axis=1) / sig_pow                      # median ignores the few symbol errors
    # median of an exponential = ln2 * mean ; mean noise power = signal power / 100
    assert abs(np.mean(ratio) / (np.log(2) * 0.01) - 1.0) < 0.1
    # end to end: the strongest target is recovered from generated data by the classical path
    yq, bq, sq, tq = pkg.generate_signals(4, Nb, Nd, L, snr_demod=30.0, seed=5, return_truth=True)
    for i in range(4):
        phi, _ = pkg.admm_for_us(yq[i].cpu().numpy().astype(np.complex128), bq[i].cpu().numpy().astype(np.complex128),
                                 10, 10, 1.0, float(sq[i]))
        pk = pkg.alt_peak_search({"phi": phi, "xbase": 10, "ybase": 10}, dict(xstep=0.01, ystep=0.01, iter=3))
        top = max(pk, key=lambda r: r[2])
        t = tq[i].cpu().numpy()
        j = int(np.argmax(np.hypot(t[:, 2], t[:, 3])))
        assert abs(top[0] - t[j, 0]) < 0.05 and abs(top[1] - t[j, 1]) < 0.05


def test_admmnet_full_module_matches_reference_golden(pkg):
    """ADMMNet.forward (admm_net.py:791-816): unrolled loop + PeakSearchLayer head, against the reference's outputs."""
    z = np.load(os.path.join(GOLDEN, "admmnet_full_k3.npz"))
    sd = {k[4:].replace("__", "."): torch.from_numpy(z[k]) for k in z.files if k.startswith("sd__")}
    net = pkg.ADMMNet(10, 10, 3, 3).eval()
    net.load_state_dict(sd)
    y, b, s = (torch.from_numpy(z[k]) for k in ("y", "b", "sigma"))
    with torch.no_grad():
        tau, f, conf, phi = net(y, b, s)
    assert tau.shape == (7, 3) and f.shape == (7, 3) and conf.shape == (7, 3) and phi.shape == (7, 100)
    assert rel_err(phi.numpy(), z["phi"]).max() < PHI_TOL
    np.testing.assert_allclose(tau.numpy(), z["tau"], atol=2e-4)
    np.testing.assert_allclose(f.numpy(), z["f"], atol=2e-4)
    np.testing.assert_allclose(conf.numpy(), z["conf"], atol=2e-4)
    # the head alone on the reference's own phi: plain fp32 MLP/attention arithmetic
    t2, f2, c2 = net.head_device(torch.from_numpy(z["phi"]).cuda())
    np.testing.assert_allclose(t2.cpu().numpy(), z["tau"], atol=2e-6, rtol=1e-5)
    np.testing.assert_allclose(f2.cpu().numpy(), z["f"], atol=2e-6, rtol=1e-5)
    np.testing.assert_allclose(c2.cpu().numpy(), z["conf"], atol=2e-6, rtol=1e-5)
    # ragged batch (not a multiple of the 8 signals a CTA takes)
    t3, _, _ = net.head_device(torch.from_numpy(z["phi"][:5]).cuda())
    np.testing.assert_allclose(t3.cpu().numpy(), z["tau"][:5], atol=2e-6, rtol=1e-5)


@pytest.mark.parametrize("switch", ["ADMMNET_ARROW", "ADMMNET_ROTF", "ADMMNET_TAILTC", "ADMMNET_LANES", "ADMMNET_TRD=1",
                                    "ADMMNET_DCK", "ADMMNET_DCK=0,ADMMNET_DC=2", "ADMMNET_DCK=0,ADMMNET_HYB=1",
                                    "ADMMNET_DCK=0,ADMMNET_HYB=2", "ADMMNET_DCK=0,ADMMNET_ROTP=1"])
def test_alternative_kernel_paths_agree(pkg, tmp_path, switch):
    """Every fast path has a plain sibling behind an environment switch (read once per process, hence the
    subprocess): ADMMNET_ARROW=0 dense eigen-solver at layer 0 instead of the arrowhead shortcut, ADMMNET_ROTF=0
    one sweep at a time in the rotation kernel, ADMMNET_TAILTC=0 the SIMT (FFMA2) back-transformation and rebuild
    instead of the tcgen05 kernel, ADMMNET_LANES=0 single stream, ADMMNET_TRD=1 the register-resident
    tridiagonalisation (opt-in) instead of the staged shared-memory one, ADMMNET_DCK=0 the QL pair instead of the fused
    divide & conquer kernel this single-chunk call takes by default (and with ADMMNET_DC=2 the legacy k_merge levels
    on top of it; with ADMMNET_HYB=1|2 the hybrid: QL pair on the 2 | 4 blocks of the torn tridiagonal, top merge
    levels in k_dc; with ADMMNET_ROTP=1 the panel-split rotation kernel, two one-warp CTAs per signal).
    Same inputs, same answer."""
    import subprocess
    import sys
    z, sd = load_net_case("pert_k10")
    net = pkg.PhiEstADMMNet(10, 10, 3, 10).eval()
    net.load_state_dict(sd)
    y, b, s = (torch.from_numpy(z[k]) for k in ("y", "b", "sigma"))
    yy, bb, ss = y.repeat(300, 1), b.repeat(300, 1), s.repeat(300)        # 2100 signals in one chunk
    with torch.no_grad():
        phi = net(yy, bb, ss).numpy()
    out = str(tmp_path / "phi_alt.npy")
    code = (
        "import sys, numpy as np, torch\n"
        f"sys.path.insert(0, {ROOT!r})\n"
        "import admmnet_b200 as pkg\n"
        "from tests.helpers import load_net_case\n"
        "z, sd = load_net_case('pert_k10')\n"
        "net = pkg.PhiEstADMMNet(10, 10, 3, 10).eval(); net.load_state_dict(sd)\n"
        "y, b, s = (torch.from_numpy(z[k]) for k in ('y', 'b', 'sigma'))\n"
        "with torch.no_grad():\n"
        f"    np.save({out!r}, net(y.repeat(300, 1), b.repeat(300, 1), s.repeat(300)).numpy())\n")
    env = dict(os.environ)
    for item in switch.split(","):
        name, _, val = item.partition("=")
        env[name] = val or "0"
    subprocess.run([sys.executable, "-c", code], check=True, env=env, timeout=600)
    phi_alt = np.load(out)
    assert rel_err(phi, phi_alt).max() < 3e-5
    # a batch of 300 copies of the 7 golden signals has the same batch mean as the 7 signals themselves
    assert rel_err(phi_alt[:7], z["phi_batch"]).max() < PHI_TOL
    assert rel_err(phi[:7], z["phi_batch"]).max() < PHI_TOL
