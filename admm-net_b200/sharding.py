"""Multi-GPU sharding of the forward path: one process per GPU, contiguous batch
shards, weights replicated.  The only cross-signal coupling of the reference is the ZLayer batch mean
(admm_net.py:459):

  norm_scope='shard'  : every rank uses the mean of its own shard -> no collective on the data path
                        (throughput mode; parity is against the reference run on that shard);
  norm_scope='global' : exact whole-batch semantics: per active layer ONE all-reduce(sum) of a single
                        fp64 (sum of residual norms) + the signal count, between the layer's kernels
                        and the next layer's dual update.

`run_layers` is engine-agnostic so the orchestration is testable on CPU (gloo) with a stand-in engine;
`CudaEngine` drives the split-phase C ABI (admmnet_layer / _set_mean / _final_phi).
"""
import ctypes as C

import torch
import torch.distributed as dist

from . import _capi


def shard_range(B, rank, world):
    """Contiguous, balanced: the first B % world ranks get one extra signal."""
    base, rem = divmod(B, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def run_layers(engine, num_layers, norm_scope="global", group=None):
    """Drive `engine` through the K-layer forward.  engine must provide:
        layer(k)                 run layer k on the local shard, leaving the local residual-norm sum
        rsum(k) -> tensor[1] f64 view of that sum on the engine's device (all-reduced in place here)
        set_mean(k, count)       finalise mean_k = rsum(k)/count
        count() -> int           local number of signals
        final() -> phi           last layer's phi-update
    """
    if norm_scope not in ("global", "shard"):
        raise ValueError("norm_scope must be 'global' or 'shard'")
    use_dist = norm_scope == "global" and dist.is_available() and dist.is_initialized() and \
        dist.get_world_size(group) > 1
    if hasattr(engine, "begin"):
        engine.begin()
    total = float(engine.count())
    if use_dist:
        cnt = torch.tensor([total], dtype=torch.float64, device=engine.rsum(0).device)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM, group=group)
        total = float(cnt.item())
    for k in range(num_layers - 1):
        engine.layer(k)
        if use_dist:
            dist.all_reduce(engine.rsum(k), op=dist.ReduceOp.SUM, group=group)
        engine.set_mean(k, total)
    return engine.final()


class CudaEngine:
    def __init__(self, model, y, b, sigma):
        self.m = model
        self.y, self.b, self.sigma = y, b, sigma
        self.B = y.shape[0]
        self.dev = y.device
        self.chunk = min(model.chunk, self.B)
        self.ws = model.workspace(self.B, self.chunk, self.dev)
        self.P = model.packed_params(self.dev)
        self.L = _capi.lib()

    def _stream(self):
        return torch.cuda.current_stream(self.dev).cuda_stream

    def count(self):
        return self.B

    def begin(self):
        m, ws = self.m, self.ws
        _capi.check(self.L.admmnet_reset_status(ws.ptr, ws.nbytes, self.B, self.chunk, m.M * m.N, m.num_layers, m.rcap,
                                                self._stream()))

    def layer(self, k):
        # all chunks over the library's chunk lanes, joined back into the current stream, then the residual-norm sum:
        # everything stream-ordered, so the all-reduce that follows needs no host synchronisation
        m, ws = self.m, self.ws
        with torch.cuda.device(self.dev):
            _capi.check(self.L.admmnet_layer(self.y.data_ptr(), self.b.data_ptr(), self.sigma.data_ptr(), self.B,
                                             self.chunk, m.M, m.N, m.num_layers, k, self.P.data_ptr(), ws.ptr,
                                             ws.nbytes, m.rcap, self._stream()))

    def rsum(self, k):
        return self.ws.rsum[k:k + 1]

    def set_mean(self, k, count):
        m, ws = self.m, self.ws
        _capi.check(self.L.admmnet_set_mean(ws.ptr, ws.nbytes, self.B, self.chunk, m.M * m.N, m.num_layers, m.rcap, k,
                                            float(count), self._stream()))

    def final(self):
        m, ws = self.m, self.ws
        out = torch.empty(self.B, m.M * m.N, dtype=torch.complex64, device=self.dev)
        _capi.check(self.L.admmnet_final_phi(self.y.data_ptr(), self.b.data_ptr(), self.B, self.chunk, m.M, m.N,
                                             m.num_layers, self.P.data_ptr(), out.data_ptr(), ws.ptr, ws.nbytes,
                                             m.rcap, self._stream()))
        m._status(ws, self._stream())
        return out


def sharded_forward(model, y_local, b_local, sigma_local, norm_scope="global", group=None):
    """Each rank passes ITS shard (device tensors).  Returns the local phi [B_local, n]."""
    yd, bd, sd, _ = model._prep(y_local, b_local, sigma_local)
    eng = CudaEngine(model, yd, bd, sd)
    return run_layers(eng, model.num_layers, norm_scope, group)
