"""On-device synthetic signals (y, b, sigma) following the reference's recipe, generate_data.py:133-221
(SURVEY.md §8f rank 2).  Thin wrapper over admmnet_generate (csrc/gen_kernels.cu)."""
import torch

from . import _capi


def generate_signals(B, Nb=10, Nd=10, L=3, snr_w=20.0, snr_demod=7.0, seed=1234, device=None, return_truth=False):
    """-> y complex64 [B, Nb*Nd], b complex64 [B, Nb*Nd], sigma float32 [B] (device tensors)
    [, truth float64 [B, L, 4] = (tau, f, Re C, Im C)]."""
    _capi.require_cuda()
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    n = Nb * Nd
    y = torch.empty(B, n, dtype=torch.complex64, device=dev)
    b = torch.empty(B, n, dtype=torch.complex64, device=dev)
    sigma = torch.empty(B, dtype=torch.float32, device=dev)
    truth = torch.empty(B, L, 4, dtype=torch.float64, device=dev) if return_truth else None
    stream = torch.cuda.current_stream(dev).cuda_stream
    _capi.check(_capi.lib().admmnet_generate(y.data_ptr(), b.data_ptr(), sigma.data_ptr(),
                                             truth.data_ptr() if truth is not None else None, B, Nb, Nd, L,
                                             float(snr_w), float(snr_demod), int(seed), stream))
    return (y, b, sigma, truth) if return_truth else (y, b, sigma)
