"""Drop-in mirror of the reference's utils/peakSearchUtils.py API for the hot path:
peak_search_func (9-33), peak_search (37-60), alt_peak_search (63-173), plus batched variants.
All arithmetic runs in csrc/peak_kernels.cu (fp64); this file only marshals options and buffers."""
import ctypes as C

import numpy as np
import torch

from . import _capi

DEFAULT_OPTS = {"xmin": 0, "xmax": 1, "xstep": 0.01, "ymin": -0.5, "ymax": 0.5, "ystep": 0.01,
                "reducefactor": 0.1, "iter": 1}                     # peakSearchUtils.py:84-88
STATUS_PEAK_OVERFLOW = 2


def _dev():
    _capi.require_cuda()
    return torch.device("cuda", torch.cuda.current_device())


def _phi_tensor(phi, dev):
    """-> contiguous [B, n] complex64/complex128 device tensor and is_c128 flag."""
    if isinstance(phi, torch.Tensor):
        t = phi.detach()
    else:
        a = np.asarray(phi)
        if not np.iscomplexobj(a):
            a = a.astype(np.complex128)
        t = torch.from_numpy(np.ascontiguousarray(a))
    if t.dtype not in (torch.complex64, torch.complex128):
        t = t.to(torch.complex128)
    if t.dim() == 1 or (t.dim() == 2 and 1 in t.shape and not isinstance(phi, torch.Tensor)):
        t = t.reshape(1, -1)
    return t.to(dev).contiguous(), int(t.dtype == torch.complex128)


def coarse_axes(so):
    """np.arange grids of peakSearchUtils.py:105-106 (the y bound uses xstep, as the reference does)."""
    ax = np.arange(so["xmin"], so["xmax"] - so["xstep"], so["xstep"])
    ay = np.arange(so["ymin"], so["ymax"] - so["xstep"], so["ystep"])
    return ax, ay


def alt_peak_search_batched(phi, xbase, ybase, opts=None, topl=0, pmax=512, return_surface=False):
    """phi: [B, n] (torch/numpy, complex).  Returns dict with device tensors:
    peaks [B,pmax,3] float64, count [B] int32, top [B,topl,3] (if topl), surface [B,Gy,Gx] (optional)."""
    so = {**DEFAULT_OPTS, **(opts or {})}
    dev = _dev()
    ph, is128 = _phi_tensor(phi, dev)
    B, n = ph.shape
    if n != xbase * ybase:
        raise ValueError("phi length must equal xbase*ybase")
    ax, ay = coarse_axes(so)
    if len(ax) == 0 or len(ay) == 0:                             # peakSearchUtils.py:109-110
        return dict(peaks=torch.zeros(B, pmax, 3, dtype=torch.float64, device=dev),
                    count=torch.zeros(B, dtype=torch.int32, device=dev),
                    top=torch.zeros(B, topl, 3, dtype=torch.float64, device=dev), surface=None)
    axd, ayd = torch.from_numpy(ax).to(dev), torch.from_numpy(ay).to(dev)
    while True:
        peaks = torch.zeros(B, pmax, 3, dtype=torch.float64, device=dev)
        count = torch.zeros(B, dtype=torch.int32, device=dev)
        top = torch.zeros(B, max(topl, 1), 3, dtype=torch.float64, device=dev)
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        surf = torch.empty(B, len(ay), len(ax), dtype=torch.float64, device=dev) if return_surface else None
        stream = torch.cuda.current_stream(dev).cuda_stream
        _capi.check(_capi.lib().peak_search_full(
            ph.data_ptr(), is128, B, int(xbase), int(ybase), axd.data_ptr(), len(ax), ayd.data_ptr(), len(ay),
            float(so["xmin"]), float(so["xmax"]), float(so["xstep"]), float(so["ymin"]), float(so["ymax"]),
            float(so["ystep"]), float(so["reducefactor"]), int(so["iter"]), pmax, peaks.data_ptr(), count.data_ptr(),
            int(topl), top.data_ptr(), surf.data_ptr() if surf is not None else None, status.data_ptr(), stream))
        if int(status.item()) & STATUS_PEAK_OVERFLOW:            # more maxima than pmax: grow and redo
            pmax = int(count.max().item())
            continue
        break
    return dict(peaks=peaks, count=count, top=top[:, :topl], surface=surf)


def alt_peak_search(func_opts, opts=None):
    """Same contract as peakSearchUtils.py:63: ndarray (P,3) float64, rows [x, y, height], unsorted."""
    so = {**DEFAULT_OPTS, **(opts or {})}
    ax, ay = coarse_axes(so)
    if len(ax) == 0 or len(ay) == 0:
        return np.zeros((0, 3))
    phi = np.asarray(func_opts["phi"]).reshape(-1)
    r = alt_peak_search_batched(phi[None], func_opts["xbase"], func_opts["ybase"], opts)
    P = int(r["count"][0].item())
    return r["peaks"][0, :P].cpu().numpy()


def peak_search(phi, X, x_base, Y, y_base):
    """peakSearchUtils.py:37-60: surface at every (X[i,j], Y[i,j]) -> ndarray like X (float64)."""
    dev = _dev()
    ph, is128 = _phi_tensor(np.asarray(phi).reshape(-1), dev)
    X = np.asarray(X, dtype=np.float64)
    Y = np.asarray(Y, dtype=np.float64)
    shape = (Y.shape[0], X.shape[1])
    Xd = torch.from_numpy(np.ascontiguousarray(X[:shape[0], :shape[1]])).to(dev).reshape(-1)
    Yd = torch.from_numpy(np.ascontiguousarray(Y[:shape[0], :shape[1]])).to(dev).reshape(-1)
    out = torch.empty(Xd.numel(), dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    _capi.check(_capi.lib().peak_search_points(ph.data_ptr(), is128, int(x_base), int(y_base), Xd.data_ptr(),
                                               Yd.data_ptr(), Xd.numel(), out.data_ptr(), stream))
    return out.cpu().numpy().reshape(shape)


def peak_search_func(phi, x, x_base, y, y_base):
    """peakSearchUtils.py:9-33 for one point."""
    return peak_search(phi, np.array([[float(x)]]), x_base, np.array([[float(y)]]), y_base)[0, 0]


def top_l(peaks, L):
    """What the callers do with the result (main_for_net.py:119-126): stable sort by height desc, keep L."""
    return np.array(sorted(peaks, key=lambda p: p[2], reverse=True)[:L]).reshape(-1, 3)
