"""GPU unit tests of the tcgen05 / TMEM / TMA layer (csrc/tc.cuh) through the debug tap admmnet_tc_gemm_probe:
one 128 x N x K tf32 UMMA tile per launch, every operand staging variant the tail kernel uses, against numpy."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

A_TMEM, A_NEG, A_MN, B_MN, TMA, SPLIT3, DCOL8 = 1, 2, 4, 8, 16, 32, 64


def _tf32(x):
    """values exactly representable in tf32 (10 mantissa bits)"""
    return (np.ascontiguousarray(x, dtype=np.float32).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def _probe(A, B, D0, flags):
    from admmnet_b200 import _capi
    L = _capi.lib()
    dev = torch.device("cuda")
    N, K = B.shape
    Ad, Bd = torch.from_numpy(A).to(dev), torch.from_numpy(B).to(dev)
    D0d = torch.from_numpy(D0).to(dev) if D0 is not None else None
    out = torch.full((128, N), float("nan"), dtype=torch.float32, device=dev)
    _capi.check(L.admmnet_tc_gemm_probe(Ad.data_ptr(), Bd.data_ptr(), D0d.data_ptr() if D0d is not None else None,
                                        N, K, flags, out.data_ptr(), torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    return out.cpu().numpy()


@pytest.mark.timeout(120)
@pytest.mark.parametrize("flags", [0, A_NEG, A_TMEM, A_TMEM | A_NEG, TMA, A_TMEM | DCOL8],
                         ids=["ss", "ss_neg", "ts", "ts_neg", "tma", "ts_dcol8"])
@pytest.mark.parametrize("N,K", [(112, 32), (104, 32), (208, 16), (48, 64), (64, 24)])
def test_tf32_tile_exact_inputs(flags, N, K):
    rng = np.random.default_rng(N * 100 + K + flags)
    A = _tf32(rng.normal(size=(128, K)))
    B = _tf32(rng.normal(size=(N, K)))
    D0 = rng.normal(size=(128, N)).astype(np.float32)
    sgn = -1.0 if flags & A_NEG else 1.0
    want = D0.astype(np.float64) + sgn * (A.astype(np.float64) @ B.astype(np.float64).T)
    got = _probe(A, B, D0, flags)
    assert np.isfinite(got).all()
    assert np.abs(got - want).max() < 2e-5 * np.abs(want).max()
    # without an initial accumulator the first MMA overwrites D
    got0 = _probe(A, B, None, flags)
    assert np.abs(got0 - (want - D0)).max() < 2e-5 * np.abs(want).max()


@pytest.mark.timeout(120)
@pytest.mark.parametrize("flags", [SPLIT3, SPLIT3 | A_TMEM], ids=["ss", "ts"])
def test_3xtf32_split_reaches_fp32_accuracy(flags):
    """arbitrary fp32 operands through the hi/lo split: error of the order of fp32 rounding, not of tf32 (5e-4)"""
    rng = np.random.default_rng(5)
    N, K = 104, 64
    A = rng.normal(size=(128, K)).astype(np.float32)
    B = rng.normal(size=(N, K)).astype(np.float32)
    want = A.astype(np.float64) @ B.astype(np.float64).T
    got = _probe(A, B, None, flags)
    err = np.abs(got - want).max() / np.abs(want).max()
    single = np.abs(_probe(_tf32(A), _tf32(B), None, flags & A_TMEM) - want).max() / np.abs(want).max()
    assert err < 3e-6, err
    assert single > 20 * err          # the unsplit product really is tf32-class: the split is doing the work
