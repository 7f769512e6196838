#!/usr/bin/env python
"""bench.py — headline benchmark of the hot path (BASELINE.json: signals/s for the K-layer ADMM-Net
forward + peak search at 1/2/4/8 B200, % of roofline, reference CPU path beside it).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N ...            # the reference's algorithm on the host cores

Workload (config.workload): BASELINE.json configs[2] — ADMM-Net K=10, n = 10x10, forward + alt_peak_search
(99x99 coarse grid, iter=3, top-3), a batch of --signals = 1 048 576 synthetic signals of generate_data.py's
recipe, sharded over the N ranks (strong scaling: N=1 processes the whole 1M batch on one B200, 87 GB of
per-signal state).  --per-gpu overrides the shard size (development runs; then the line says "weak").
No collective on the data path (norm_scope='shard': the ZLayer batch mean is taken over each rank's shard);
the exact whole-batch semantics (--norm-scope global: 9 scalar all-reduces per forward) is timed and checked
as extras.global_scope whenever N > 1.  One step = one pass over the batch; timed with CUDA events between
barriers, max over ranks.  After the timed region a random subset of the batch is recomputed with the CPU
oracle (per-layer batch means read back from the device) and compared: the `parity` object of the JSON line.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

K_LAYERS, M, N, TOPL = 10, 10, 10, 3
PEAK_OPTS = dict(xstep=0.01, ystep=0.01, iter=3)          # main_for_net.py:112-116
METRIC = "signals/s, ADMM-Net K=10 forward + peak search"
# SURVEY.md §8d algorithmic work per signal at n=100 (d=101)
FLOP_PER_LAYER = 33.1e6          # eigh 24 d^3 + rebuild 8 d^3 + 0.15 MFLOP elementwise
FLOP_PEAK_SEARCH = 0.86e6        # separable coarse surface
BYTES_PER_SIGNAL = 2404          # y 800 + b 800 + sigma 4 in, phi 800 out
# implementation DRAM traffic of the layer kernels per signal-layer, from the ncu --set full capture in
# profiles/r01_ncu_summary.md (dram__bytes_read+write over 4096 signals): k_head 180 + k_head2 (4 stages) 174 +
# k_ql 56 + k_rot 102 + k_tail 114 KB
TRAFFIC_PER_SIGNAL_LAYER = 626e3      # general layer (k_head .. k_tail_p), ncu dram read+write per signal
TRAFFIC_ARROW_LAYER = 71e3            # layer 0 through k_arrow


def tile_signals(B, seed):
    """generate_data.py-recipe signals (oracle/signals.py is only the INPUT generator here); 4096 unique
    signals tiled to B with a per-copy global phase so no two signals are identical."""
    from oracle import signals
    base = min(B, 4096)
    y, b, s, _ = signals.generate(base, seed=seed)
    reps = (B + base - 1) // base
    if reps > 1:
        rng = np.random.default_rng(seed + 1)
        ph = np.exp(1j * rng.uniform(0, 2 * np.pi, reps)).astype(np.complex64)
        y = (y[None] * ph[:, None, None]).reshape(-1, y.shape[1])[:B]
        b = np.tile(b, (reps, 1))[:B]
        s = np.tile(s, reps)[:B]
    return np.ascontiguousarray(y), np.ascontiguousarray(b), np.ascontiguousarray(s)


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [nm for i, nm in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def cpu_sample(n_fwd, n_peak, threads, pool=None):
    """The reference's algorithm on the host: oracle port of the forward (torch CPU, `threads` intra-op
    threads) on n_fwd signals + literal port of alt_peak_search on n_peak signals (one per worker)."""
    from oracle import net_oracle, peak_oracle, signals
    import admmnet_b200
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    sd = admmnet_b200.PhiEstADMMNet(M, N, 3, K_LAYERS).state_dict()
    y, b, s, _ = signals.generate(n_fwd, seed=99)
    yt, bt, st = (torch.from_numpy(a) for a in (y, b, s))
    t0 = time.perf_counter()
    phi = net_oracle.forward(sd, yt, bt, st, M, N, K_LAYERS).numpy()
    t_f = time.perf_counter() - t0
    jobs = [phi[i % n_fwd] for i in range(n_peak)]
    t0 = time.perf_counter()
    if pool is not None:
        pool.map(_peak_job, jobs)
    else:
        for j in jobs:
            _peak_job(j)
    t_p = time.perf_counter() - t0
    per_signal = t_f / n_fwd + t_p / n_peak
    return 1.0 / per_signal, t_f, t_p


def _peak_job(phi):
    from oracle import peak_oracle
    r = peak_oracle.alt_peak_search({"phi": phi, "xbase": N, "ybase": M}, PEAK_OPTS)
    return peak_oracle.top_l(r, TOPL)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    n_fwd, n_peak = 256, max(cores, 2)
    vals, times = [], []
    with mp.get_context("fork").Pool(cores) as pool:
        for i in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            v, tf, tp = cpu_sample(n_fwd, n_peak, cores, pool)
            if i >= args.warmup:
                vals.append(v)
                times.append(time.perf_counter() - t0)
    v = float(np.mean(vals))
    sample = (f"per step: oracle port of PhiEstADMMNet.forward (torch CPU, {cores} threads, no_grad) on {n_fwd} signals + "
              f"literal port of alt_peak_search (99x99, iter=3) on {n_peak} signals over {cores} processes; "
              "signals/s = 1/(t_fwd/n_fwd + t_peak/n_peak)")
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "signals/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(times)),
            "higher_is_better": True, "scaling": "weak" if args.per_gpu else "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(args),
            "cpu_baseline": {"value": v, "unit": "signals/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "signals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def torch_cuda_baseline(dev, B=512, reps=1):
    """BASELINE.md §3.4 "the GPU bar to beat": the reference's own op sequence (oracle port of
    PhiEstADMMNet.forward: torch.linalg.eigh -> cuSOLVER, bmm -> cuBLAS, ~150 torch ops per layer) with all
    tensors on the B200, no_grad.  /root/reference does not exist on the GPU box, so the op-for-op port is what
    runs; it is pinned to the reference by tests/golden."""
    from oracle import net_oracle
    import admmnet_b200
    torch.manual_seed(0)
    sd = admmnet_b200.PhiEstADMMNet(M, N, 3, K_LAYERS).state_dict()
    y, b, s = tile_signals(B, seed=55)
    yt, bt, st = (torch.from_numpy(a).to(dev) for a in (y, b, s))
    net_oracle.forward(sd, yt[:256], bt[:256], st[:256], M, N, K_LAYERS)        # warm-up (cuSOLVER handles)
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        net_oracle.forward(sd, yt, bt, st, M, N, K_LAYERS)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return {"value": B / (best * 1e-3), "unit": "signals/s (forward only, no peak search)", "batch": B, "ms": best,
            "what": "oracle port of the reference forward as stock PyTorch CUDA ops on this B200 "
                    "(torch.linalg.eigh/cuSOLVER + cuBLAS bmm + elementwise), no_grad, best of %d" % reps}


def run_torch_cuda(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    g = torch_cuda_baseline(dev)
    print(json.dumps({"impl": "torch-cuda", "metric": METRIC + " (forward only)", "value": g["value"],
                      "unit": "signals/s", "n_gpus": 1, "higher_is_better": True, "dtype": "f32", "data": "synthetic",
                      "gpu_baseline": g}))


def parity_check(model, yd, bd, sd_, phi_dev, top_dev, means, n_fwd=64, n_peak=8, seed=7):
    """Recompute a random subset of the batch the timed steps processed with the CPU oracle and compare.
    The ZLayer batch mean couples the signals (admm_net.py:459), so the oracle gets the per-layer means the device
    used (ws.mean, read back) — then each signal's result only depends on its own data.  Peak search: the oracle's
    alt_peak_search (separable surface) on the SAME phi the device search saw, top-L positions must be bit-equal."""
    from oracle import net_oracle, peak_oracle
    g = torch.Generator().manual_seed(seed)
    idx = torch.randperm(yd.shape[0], generator=g)[:n_fwd].sort().values
    idd = idx.to(yd.device)
    yt, bt, st = yd[idd].cpu(), bd[idd].cpu(), sd_[idd].cpu()
    ref = net_oracle.forward(model.state_dict(), yt, bt, st, M, N, K_LAYERS, means=means)
    got = phi_dev[idd].cpu()
    rel = ((got - ref).abs().amax(1) / ref.abs().amax(1))
    top = top_dev[idd[:n_peak]].cpu().numpy()
    same = True
    for i in range(n_peak):
        r = peak_oracle.alt_peak_search({"phi": got[i].numpy(), "xbase": N, "ybase": M}, PEAK_OPTS,
                                        surface=lambda p, X, xb, Y, yb: peak_oracle.peak_search_separable(p, X[0], xb, Y[:, 0], yb))
        t = peak_oracle.top_l(r, TOPL)
        same = same and np.array_equal(np.asarray(t)[:, :2], top[i][:len(t), :2])
    return {"signals_checked": int(n_fwd), "phi_rel_max": float(rel.max()), "phi_rel_median": float(rel.median()),
            "peaks_checked": int(n_peak), "peaks_identical": bool(same), "tolerance": 1e-4,
            "how": "CPU oracle on a random subset of the timed batch with the device's per-layer batch means"}


def run_train(args):
    """BASELINE.json configs[4] (trainPhi.py recipe, SURVEY.md §8d cfg 5): PhiEstADMMNet K=10, PhiAlignmentLoss, AdamW
    + gradient clipping, batch 256 per GPU, labels phi from the classical solver, one flat NCCL gradient all-reduce per
    step.  The step (forward, backward, all-reduce, clip, optimizer) is captured in one CUDA graph and replayed; the
    e2e figure adds the pinned-host -> device copy of every batch and the device -> host read of the loss."""
    import admmnet_b200
    import torch.distributed as dist
    from admmnet_b200.autograd import PhiAlignmentLoss, eigh_status
    from admmnet_b200.training import GraphedTrainStep, make_optimizer, train_step
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    Bt = args.train_batch
    torch.manual_seed(0)
    model = admmnet_b200.PhiEstADMMNet(M, N, 3, K_LAYERS).to(dev)
    opt, _ = make_optimizer(model, 5e-3, 1e-3, capturable=True)
    crit = PhiAlignmentLoss()
    nbatch = 8
    y, b, s = admmnet_b200.generate_signals(Bt * nbatch, M, N, 3, snr_w=20.0, snr_demod=7.0, seed=777 + rank, device=dev)
    pt = admmnet_b200.admm_for_us_batched(y, b, 1.0, 5).to(torch.complex64)
    host = [t.cpu().pin_memory() for t in (y, b, s, pt)]
    ex = tuple(t[:Bt] for t in (y, b, s, pt))
    mode = "cuda-graph"
    try:
        step = GraphedTrainStep(model, crit, opt, ex)
    except Exception as e:                                   # e.g. NCCL capture refused: fall back to eager steps
        mode = f"eager ({type(e).__name__})"
        step = lambda *a: train_step(model, crit, opt, *a, check_status=False)[0]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    dev_batches = [tuple(t[i * Bt:(i + 1) * Bt] for t in (y, b, s, pt)) for i in range(nbatch)]
    losses = []

    def step_dev(i):
        losses.append(step(*dev_batches[i % nbatch]).clone())

    def step_e2e(i):
        bt = tuple(t[(i % nbatch) * Bt:(i % nbatch + 1) * Bt].to(dev, non_blocking=True) for t in host)
        float(step(*bt))                                     # D2H read of the loss every step

    for i in range(args.warmup):
        step_dev(i)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    steps = max(args.steps, 20)
    ms = timed(step_dev, steps)
    clocks = sampler.stop() if rank == 0 else None
    ms_e2e = timed(step_e2e, steps)
    assert eigh_status(dev) == 0
    first, last = float(losses[0]), float(losses[-1])
    if rank == 0:
        nparam = sum(p.numel() for p in model.parameters())
        print(json.dumps({
            "mode": "train", "metric": "training samples/s, trainPhi.py step (PhiEstADMMNet K=10, batch 256 per GPU)",
            "value": world * Bt * steps / (ms * 1e-3), "unit": "samples/s", "n_gpus": world, "steps": steps,
            "warmup": args.warmup, "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"BASELINE.json configs[4]: trainPhi.py step, K={K_LAYERS}, n={M}x{N}, batch {Bt} per GPU, "
                                   "AdamW + clip 1.0, PhiAlignmentLoss, labels from the classical solver",
                       "step_mode": mode, "allreduce_bytes_per_step": 4 * (nparam + 1) if world > 1 else 0},
            "clocks": clocks, "loss_first": first, "loss_last": last,
            "e2e": {"value": world * Bt * steps / (ms_e2e * 1e-3), "unit": "samples/s",
                    "h2d_bytes_per_step": int(sum(t[:Bt].nbytes for t in host)), "d2h_bytes_per_step": 4}}))
    if world > 1:
        dist.destroy_process_group()


def _time_ms(fn, reps):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def measure_tensor_pipe(pkg, dev, tf32_peak_tflops, Bt=16384, d=101):
    """Tensor-pipe roofline of k_tail_tc, timed ALONE (inside the forward its CUDA-event time overlaps the other
    chunk lane's kernels): the f(A) tap (tridiagonalisation, tridiagonal solver, tail) on Bt random Hermitian matrices of
    the benchmark's order with per-kernel events; executed = every tcgen05.mma flop the kernel issues (3xTF32 split
    terms included, admmnet_tail_tc_mma_flops), useful = the 12 d^3 flops of the two contractions it replaces
    (back-transformation 8 d^3 + lower-triangle rebuild 4 d^3)."""
    from admmnet_b200 import _capi
    from admmnet_b200.params import pack_state_dict
    L = _capi.lib()
    mma = float(L.admmnet_tail_tc_mma_flops(d))
    if mma == 0.0 or L.admmnet_tail_tc_smem_bytes(d) <= 0:
        return None
    torch.manual_seed(0)
    net = pkg.PhiEstADMMNet(M, N, 3, K_LAYERS)
    P = pack_state_dict(net.state_dict(), M * N, K_LAYERS).to(dev)
    X = torch.randn(Bt, d, d, dtype=torch.complex64, device=dev) * (3.0 / d ** 0.5)
    A = (0.5 * (X + X.transpose(1, 2).conj())).contiguous()
    del X
    nb = C.c_size_t()
    _capi.check(L.admmnet_eigh_workspace_bytes(Bt, d, 0, C.byref(nb)))
    ws = torch.empty(nb.value, dtype=torch.uint8, device=dev)
    G = torch.empty(Bt, d * (d + 1) // 2, dtype=torch.complex64, device=dev)
    st = torch.zeros(1, dtype=torch.int32, device=dev)

    def run():
        _capi.check(L.admmnet_eigh_batched(A.data_ptr(), Bt, d, None, None, G.data_ptr(), P[3].data_ptr(), ws.data_ptr(),
                                           nb.value, 0, torch.cuda.current_stream().cuda_stream, st.data_ptr()))
    for _ in range(2):
        run()
    torch.cuda.synchronize()
    nk = L.admmnet_profile_kinds()
    reps = 5
    L.admmnet_profile_begin()
    for _ in range(reps):
        run()
    ms = (C.c_double * nk)()
    ln = (C.c_longlong * nk)()
    _capi.check(L.admmnet_profile_end(ms, ln))
    names = [L.admmnet_profile_kind_name(i).decode() for i in range(nk)]
    t_ms = ms[names.index("k_tail")] / reps
    assert int(st.item()) == 0
    executed = mma * Bt / (t_ms * 1e-3) / 1e12
    return {"bound": "tensor", "kernel": "k_tail_tc (tcgen05.mma kind::tf32, 3xTF32 split, X resident in TMEM)",
            "ms_per_launch": t_ms, "signals_per_launch": Bt, "mma_flops_per_signal": mma,
            "achieved": executed, "peak": tf32_peak_tflops, "unit": "TFLOP/s", "frac": executed / tf32_peak_tflops,
            "useful_tflops": 12.0 * d ** 3 * Bt / (t_ms * 1e-3) / 1e12,
            "note": "kernel timed alone on the f(A) tap; achieved counts every issued MMA flop, useful the 12 d^3 of the "
                    "back-transformation and rebuild; peak = TF32 cuBLAS 8192^3 measured in this run"}


def measure_extras(pkg, dev):
    """The other BASELINE.json configs, measured on rank 0 after the headline run (not bench lines of their own):
    configs[0] single-signal latency (what results/time/*.txt of the reference publish), configs[1] classical ADMM
    at batch 64k against the HBM roofline, configs[3] peak-search dictionary sweep."""
    from oracle import signals
    out = {}
    # configs[0]: B = 1, K = 10 / 5, host tensors in and out (test/test_time_net.py:98-100 times model(y,b,sigma))
    y, b, s, _ = signals.generate(1, seed=5)
    yt, bt, st = (torch.from_numpy(a) for a in (y, b, s))
    for K in (10, 5):
        torch.manual_seed(0)
        net = pkg.PhiEstADMMNet(M, N, 3, K).eval()
        net(yt, bt, st)
        ts = []
        for _ in range(20):
            t0 = time.perf_counter()
            net(yt, bt, st)
            ts.append((time.perf_counter() - t0) * 1e3)
        out[f"latency_b1_k{K}_ms"] = float(np.median(ts))
    out["latency_published_ms"] = {"k10_mean": 191.0, "k5_mean": 96.5, "classic_mean": 524.4,
                                   "source": "results/time/time_net.txt, time_net_5.txt, time.txt (hardware not stated)"}
    # configs[1]: classical ADMM, batch 65536, complex64 inputs -> complex128 phi; n_iter = 5 (what admm_for_us
    # executes) and 100 (SURVEY.md §8d cfg 2).  HBM-bound: roofline-style entries against the measured copy bandwidth.
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    Bc = 65536
    y, b, s = tile_signals(Bc, seed=77)
    yd, bd = torch.from_numpy(y).to(dev), torch.from_numpy(b).to(dev)
    o = torch.empty(Bc, M * N, dtype=torch.complex128, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)          # 256 MB > 126 MB L2
    byt = Bc * (800 + 800 + 1600)
    # steady state without any L2 reuse: four distinct (y, b, phi) sets = 840 MB, six times the 126 MB L2, visited round
    # robin back to back.  (The memset flush of the second figure leaves 126 MB of DIRTY lines in L2 whose write-back is
    # charged to the timed launch - it is kept for continuity with round 1.)
    sets = [(yd, bd, o)] + [(yd.clone(), bd.clone(), torch.empty_like(o)) for _ in range(3)]
    for it in (5, 100):
        for ys_, bs_, os_ in sets:
            pkg.admm_for_us_batched(ys_, bs_, 1.0, it, out=os_)
        reps = 40
        ms = _time_ms(lambda: [pkg.admm_for_us_batched(*sets[r % 4][:2], 1.0, it, out=sets[r % 4][2]) for r in range(reps)], 1) / reps
        ts = []
        for _ in range(10):
            flush.zero_()                                                  # L2 flush between timed iterations
            ts.append(_time_ms(lambda: pkg.admm_for_us_batched(yd, bd, 1.0, it, out=o), 1))
        ms_f = float(np.median(ts))
        out[f"classic_64k_iter{it}"] = {
            "signals_per_s": Bc / (ms * 1e-3), "ms": ms, "ms_after_memset_flush": ms_f,
            "roofline": {"bound": "hbm", "achieved": byt / (ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                         "frac": byt / (ms * 1e-3) / 1e9 / hbm_peak, "algorithmic_bytes": byt,
                         "frac_after_memset_flush": byt / (ms_f * 1e-3) / 1e9 / hbm_peak,
                         "peak_source": "measured" if peaks else "fallback"},
            "note": "BASELINE.json configs[1]; c64 y,b in / c128 phi out = 3200 B per signal, 210 MB per launch; "
                    "ms: 40 back-to-back launches over four rotating buffer sets (840 MB working set, no L2 reuse); "
                    "ms_after_memset_flush: single launches, each after a 256 MB memset (median of 10)"
                    + ("; 100 iterations are bound by the fp64 pipe (600 fp64 operations per element), not HBM" if it == 100 else "")}
    del sets
    del flush
    # stock PyTorch CUDA ops of the reference forward on this B200
    out["gpu_baseline"] = torch_cuda_baseline(dev)
    # configs[3]: peak-search steering-dictionary sweep (n x grid), batch 262144 per point
    sweep = []
    rng = np.random.default_rng(3)
    Bp = 262144
    for nb, g in ((8, 32), (10, 45), (12, 64), (14, 64), (16, 64)):
        base = torch.from_numpy((rng.normal(size=(4096, nb * nb)) + 1j * rng.normal(size=(4096, nb * nb))).astype(np.complex64)).to(dev)
        ph = torch.exp(1j * torch.rand(Bp // 4096, 1, 1, device=dev) * 6.2831853).to(torch.complex64)
        phi = (base[None] * ph).reshape(Bp, nb * nb).contiguous()
        opts = dict(xstep=1.0 / g, ystep=1.0 / g, iter=3)
        pkg.alt_peak_search_batched(phi, nb, nb, opts, topl=3, pmax=256)
        ms = _time_ms(lambda: pkg.alt_peak_search_batched(phi, nb, nb, opts, topl=3, pmax=256), 1)
        flop = 8.0 * ((g - 1) * nb * nb + (g - 1) * nb * (g - 1))          # separable coarse surface, SURVEY §8d
        sweep.append({"n": nb * nb, "grid": f"{g - 1}x{g - 1}", "batch": Bp, "signals_per_s": Bp / (ms * 1e-3),
                      "ms": ms, "coarse_surface_gflops": flop * Bp / (ms * 1e-3) / 1e9})
        del phi
    out["peak_search_sweep"] = sweep
    # configs[3], net forward: matrix orders above 128 run the size-agnostic Jacobi layer kernel (k_big_layer) - a
    # correctness path (~40x the flops of the Householder pipeline), timed here so that the sweep has a forward figure
    big = []
    for nb, Bb in ((12, 592), (14, 296), (16, 296)):          # 2 | 1 | 1 signals per persistent CTA (grid 296)
        torch.manual_seed(0)
        netb = pkg.PhiEstADMMNet(nb, nb, 3, K_LAYERS).eval()
        yb_, bb_, sb_ = pkg.generate_signals(Bb, nb, nb, 3, snr_w=20.0, snr_demod=7.0, seed=7, device=dev)
        netb.forward_device(yb_, bb_, sb_)
        ms = _time_ms(lambda: netb.forward_device(yb_, bb_, sb_), 1)
        big.append({"n": nb * nb, "batch": Bb, "K": K_LAYERS, "signals_per_s": Bb / (ms * 1e-3), "ms": ms,
                    "kernel": "k_big_layer (cyclic two-sided Jacobi, L2-resident scratch)"})
        del netb, yb_, bb_, sb_
    out["net_forward_large_n"] = big
    return out


def workload_config(args, world=None):
    world = world or args.gpus
    per = args.per_gpu if args.per_gpu else args.signals // world
    how = (f"{per} signals per GPU per step (--per-gpu override), weak scaling" if args.per_gpu else
           f"{args.signals} signals per step sharded over {world} GPU(s) = {per} per GPU, strong scaling")
    return {"workload": f"BASELINE.json configs[2]: ADMM-Net K={K_LAYERS}, n={M}x{N}, forward + alt_peak_search "
                        f"(99x99, iter=3, top-{TOPL}); {how}",
            "signals_per_step": per * world, "per_gpu_batch": per, "chunk": args.chunk, "norm_scope": args.norm_scope,
            "l2": "working set per step (>10 GB of per-signal state) exceeds the 126 MB L2; no flush needed"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "torch-cuda"])
    ap.add_argument("--signals", type=int, default=1048576, help="batch per step, sharded over the ranks")
    ap.add_argument("--per-gpu", type=int, default=0, help="override: signals per GPU per step (weak scaling)")
    ap.add_argument("--chunk", type=int, default=16384)
    ap.add_argument("--norm-scope", default="shard", choices=["shard", "global"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--mode", default="forward", choices=["forward", "train"],
                    help="forward: the headline metric; train: the trainPhi.py step (BASELINE.json configs[4])")
    ap.add_argument("--train-batch", type=int, default=256)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.impl == "torch-cuda":
        return run_torch_cuda(args)
    if args.mode == "train":
        return run_train(args)

    import admmnet_b200
    from admmnet_b200 import _capi, sharding
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    assert torch.cuda.is_available(), "bench.py needs CUDA devices (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        import datetime
        # a collective that one rank never enters should fail in minutes, not after the default 10-minute watchdog
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=240))
    L = _capi.lib()

    B = args.per_gpu if args.per_gpu else args.signals // world
    torch.manual_seed(0)
    model = admmnet_b200.PhiEstADMMNet(M, N, 3, K_LAYERS).eval()
    model.chunk = args.chunk
    model.check_status = False               # status word is read once after the timed region
    # B unique signals of the generate_data.py recipe, generated on the device (csrc/gen_kernels.cu), seed = 1234 + rank
    yd, bd, sd_ = admmnet_b200.generate_signals(B, M, N, 3, snr_w=20.0, snr_demod=7.0, seed=1234 + rank, device=dev)
    yh, bh, sh = (t.cpu().pin_memory() for t in (yd, bd, sd_))
    phi_h = torch.empty(B, M * N, dtype=torch.complex64).pin_memory()
    top_h = torch.empty(B, TOPL, 3, dtype=torch.float64).pin_memory()
    cnt_h = torch.empty(B, dtype=torch.int32).pin_memory()

    def step_device(y_, b_, s_):
        if args.norm_scope == "global" and world > 1:
            phi = sharding.sharded_forward(model, y_, b_, s_, "global")
        else:
            phi = model.forward_device(y_, b_, s_)
        pk = admmnet_b200.alt_peak_search_batched(phi, N, M, PEAK_OPTS, topl=TOPL, pmax=256)
        return phi, pk

    def step_e2e():
        y_, b_, s_ = yh.to(dev, non_blocking=True), bh.to(dev, non_blocking=True), sh.to(dev, non_blocking=True)
        phi, pk = step_device(y_, b_, s_)
        phi_h.copy_(phi, non_blocking=True)
        top_h.copy_(pk["top"], non_blocking=True)
        cnt_h.copy_(pk["count"], non_blocking=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(args.warmup):
        step_device(yd, bd, sd_)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    nk = L.admmnet_profile_kinds()
    L.admmnet_profile_begin()
    ms_total = timed(lambda: step_device(yd, bd, sd_), args.steps)
    kms = (C.c_double * nk)()
    kln = (C.c_longlong * nk)()
    _capi.check(L.admmnet_profile_end(kms, kln))
    clocks = sampler.stop() if rank == 0 else None
    # status of the last forward (eigen-solver convergence) and of the run as a whole
    ws = model._ws
    st = C.c_int(0)
    _capi.check(L.admmnet_status(ws.ptr, ws.nbytes, ws.B, ws.chunk, ws.n, ws.K, ws.rcap,
                                 torch.cuda.current_stream().cuda_stream, C.byref(st)))
    assert st.value == 0, f"eigen-solver status {st.value}"
    # parity of the benchmarked configuration itself (rank 0): the last timed step's outputs against the CPU oracle
    parity = None
    if not args.no_parity and (rank == 0 or (args.norm_scope == "global" and world > 1)):
        # (with norm_scope=global the step contains collectives: every rank runs it, rank 0 checks its shard)
        phi_last, pk_last = step_device(yd, bd, sd_)
        if rank == 0:
            means = ws.mean[:K_LAYERS - 1].cpu().tolist()
            parity = parity_check(model, yd, bd, sd_, phi_last, pk_last["top"], means)
            assert parity["phi_rel_max"] < 1e-4, parity
            assert parity["peaks_identical"], parity
        del phi_last, pk_last

    # end-to-end through the public API with host buffers
    step_e2e()
    ms_e2e = timed(step_e2e, max(1, min(args.steps, 3)))
    e2e_steps = max(1, min(args.steps, 3))

    # exact whole-batch semantics across the ranks (SURVEY.md §8e 'global'): 9 scalar all-reduces per forward
    global_scope = None
    if world > 1 and not args.no_extras:
        def step_global():
            phi = sharding.sharded_forward(model, yd, bd, sd_, "global")
            return phi, admmnet_b200.alt_peak_search_batched(phi, N, M, PEAK_OPTS, topl=TOPL, pmax=256)
        step_global()
        ms_g = timed(step_global, 1)
        # parity against ONE GPU computing the whole (small) batch: every rank takes 4096 signals of a 4096*world
        # batch generated from one seed; rank 0 also runs all of it alone with norm_scope='batch'
        Bs = 4096
        ya, ba, sa = admmnet_b200.generate_signals(Bs * world, M, N, 3, snr_w=20.0, snr_demod=7.0, seed=4321, device=dev)
        lo = rank * Bs
        phi_g = sharding.sharded_forward(model, ya[lo:lo + Bs], ba[lo:lo + Bs], sa[lo:lo + Bs], "global")
        par = None
        if rank == 0:
            phi_1 = model.forward_device(ya, ba, sa)[:Bs]
            par = float(((phi_g - phi_1).abs().amax(1) / phi_1.abs().amax(1)).max().item())
        global_scope = {"value": world * B / (ms_g * 1e-3), "unit": "signals/s", "ms_per_step": ms_g,
                        "parity_vs_single_gpu": par, "collectives_per_step": K_LAYERS - 1,
                        "note": "norm_scope=global: chunk lanes kept (admmnet_layer), one fp64 NCCL all-reduce per "
                                "active layer, stream-ordered (no host sync); parity = max rel. difference of rank 0's "
                                "shard against one GPU running the whole 4096*N batch"}

    # FP32 FMA peak (roofline denominator), timed alone
    outp = torch.empty(148 * 8 * 256, dtype=torch.float32, device=dev)
    flops = C.c_double()
    stream = torch.cuda.current_stream().cuda_stream
    L.admmnet_fp32_peak_launch(outp.data_ptr(), 148 * 8, 2000, C.byref(flops), stream)
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        L.admmnet_fp32_peak_launch(outp.data_ptr(), 148 * 8, 2000, C.byref(flops), stream)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    fp32_peak_tflops = flops.value / (best * 1e-3) / 1e12

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # TF32 tensor-pipe peak (SURVEY.md §8d asks for both denominators): cuBLAS 8192^3 with TF32 allowed, best of 5
    torch.backends.cuda.matmul.allow_tf32 = True
    am = torch.randn(8192, 8192, device=dev)
    bm = torch.randn(8192, 8192, device=dev)
    torch.matmul(am, bm)
    tf32_best = min(_time_ms(lambda: torch.matmul(am, bm), 1) for _ in range(5))
    tf32_peak_tflops = 2 * 8192 ** 3 / (tf32_best * 1e-3) / 1e12
    torch.backends.cuda.matmul.allow_tf32 = False
    del am, bm
    tensor_pipe = measure_tensor_pipe(admmnet_b200, dev, tf32_peak_tflops)
    extras = None if args.no_extras else measure_extras(admmnet_b200, dev)
    if extras is not None and global_scope is not None:
        extras["global_scope"] = global_scope
    names = [L.admmnet_profile_kind_name(i).decode() for i in range(nk)]
    kern = {names[i]: {"ms_per_step": kms[i] / args.steps, "launches_per_step": kln[i] / args.steps} for i in range(nk) if kln[i]}
    gpu_ms = sum(v["ms_per_step"] for v in kern.values())
    for v in kern.values():
        v["share"] = v["ms_per_step"] / gpu_ms
    eig = ("k_head", "k_head2", "k_ql", "k_rot", "k_merge", "k_dc", "k_tail")
    eig_launch = sum(kern[k]["launches_per_step"] for k in eig if k in kern)
    dom = max(kern, key=lambda k: kern[k]["ms_per_step"])
    ms_step = ms_total / args.steps
    value = world * B / (ms_step * 1e-3)
    # device time of the K-layer forward: the step minus the peak-search launch (which runs alone, after the
    # forward, on the same stream).  The per-kernel event sums of the layer kernels overlap across the chunk
    # lanes, so they give shares, not the forward's duration.
    fwd_ms = ms_step - kern.get("k_peak_search", {"ms_per_step": 0.0})["ms_per_step"]
    achieved = FLOP_PER_LAYER * (K_LAYERS - 1) * B / (fwd_ms * 1e-3) / 1e12
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    line = {
        "metric": METRIC, "value": value, "unit": "signals/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "weak" if args.per_gpu else "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args, world),
        "clocks": clocks,
        "e2e": {"value": world * B / (ms_e2e / e2e_steps * 1e-3), "unit": "signals/s",
                "h2d_bytes_per_step": int(yh.nbytes + bh.nbytes + sh.nbytes),
                "d2h_bytes_per_step": int(phi_h.nbytes + top_h.nbytes + cnt_h.nbytes)},
        "gpu_launches": int(sum(kln[i] for i in range(nk))),
        "parity": parity,
        "peaks": {"fp32_fma_tflops": fp32_peak_tflops, "tf32_tensor_tflops": tf32_peak_tflops,
                  "bf16_tensor_tflops": peaks.get("bf16_tflops"), "hbm_gbs": hbm_peak,
                  "how": "FP32: FFMA micro-kernel in this library; TF32: torch.matmul 8192^3 allow_tf32, best of 5; "
                         "bf16 and HBM: MEASURED_PEAKS.json"},
        "roofline": {
            "bound": "fp32", "kernel": "layer pipeline k_head+k_head2+k_ql+k_rot+k_tail_tc (dominant: %s); the tensor-pipe stage k_tail_tc has its own entry under 'tensor'" % dom,
            "forward_ms_per_step": fwd_ms,
            "achieved": achieved, "peak": fp32_peak_tflops, "unit": "TFLOP/s", "frac": achieved / fp32_peak_tflops,
            "traffic": (TRAFFIC_PER_SIGNAL_LAYER * (K_LAYERS - 2) + TRAFFIC_ARROW_LAYER) * B,
            "traffic_note": "bytes per step of the layer kernels (ncu dram read+write, profiles/r01_ncu_summary.md: layer 0 through k_arrow, 8 general layers); algorithmic "
                            "bytes per step are 2404 B x signals: the path is compute bound, the extra traffic is "
                            "per-layer state (packed Z, G, reflectors, rotation stream) at <5 % of HBM bandwidth",
            "note": "FP32-FMA/shared-memory bound eigen-solver (SURVEY.md §8d): algorithmic 33.1 MFLOP per signal-layer x "
                    "(K-1) layers / device time of the forward (step minus the peak-search launch); peak = FP32 FFMA micro-kernel measured "
                    "live (MEASURED_PEAKS.json has no FP32 figure). launches per step: %d" % eig_launch,
            "hbm": {"achieved_gbs": BYTES_PER_SIGNAL * B / (ms_step * 1e-3) / 1e9, "peak_gbs": hbm_peak,
                    "peak_source": "measured" if peaks else "fallback"},
            "tensor": tensor_pipe,
            "kernels": kern},
    }
    if extras is not None:
        line["extras"] = extras
    if not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        v, tf, tp = cpu_sample(192, 4, cores)
        line["cpu_baseline"] = {"value": v, "unit": "signals/s", "cores": cores, "kind": "port",
                                "sample": f"oracle port of the forward on 192 signals ({tf:.1f} s, {cores} torch threads) + "
                                          f"literal port of alt_peak_search on 4 signals ({tp:.1f} s, 1 thread); "
                                          "signals/s = 1/(t_fwd/192 + t_peak/4)"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
