"""Data-parallel training step for the unrolled net (trainPhi.py:158-179 per batch) with a flat-buffer gradient
all-reduce (NCCL over NVLink on the GPU box, gloo in the CPU tests).  lambda_param's, step_adjust_net and the last
layer's H/G/Z parameters receive no gradient in the reference either: they are all-reduced as zeros,
so no `find_unused_parameters` machinery is needed."""
import torch
import torch.distributed as dist


def allreduce_gradients(model, group=None, weight=1.0):
    """Average .grad over the ranks with ONE all-reduce of a flat fp32 buffer (132 230 floats = 0.53 MB at K=10).
    `weight` = this rank's number of samples in the global batch: the result is the sample-weighted mean
    sum_r w_r g_r / sum_r w_r, i.e. the mean gradient of the global batch even when the shards are ragged; a rank
    whose shard is empty passes weight 0 (its gradients count as zeros) and still takes part in the collective."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    params = [p for p in model.parameters() if p.requires_grad]
    w = torch.tensor([float(weight)], device=params[0].device, dtype=torch.float32)
    flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1).float() * w
                      for p in params] + [w])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat = flat[:-1] / flat[-1].clamp(min=1e-30)
    off = 0
    for p in params:
        nel = p.numel()
        if p.grad is not None:
            p.grad.copy_(flat[off:off + nel].view_as(p))
        elif weight == 0:
            p.grad = flat[off:off + nel].view_as(p).clone()
        off += nel


def train_step(model, criterion, optimizer, y, b, sigma, phi_true, max_norm=1.0, group=None, check_status=True):
    """optimizer.zero_grad -> forward -> loss -> backward -> all-reduce -> clip -> step  (trainPhi.py:165-178).
    A rank with an empty shard (y.shape[0] == 0: tail batch smaller than the world size) skips forward/backward but
    still joins the all-reduce with weight 0 and applies the same optimizer step, so the replicas stay in sync."""
    model.train()
    optimizer.zero_grad()
    nloc = y.shape[0]
    if nloc > 0:
        phi = model(y, b, sigma)
        loss, parts = criterion(phi, phi_true)
        loss.backward()
    else:
        dev = next(model.parameters()).device
        loss, parts = torch.zeros((), device=dev), {}
    allreduce_gradients(model, group, weight=nloc)
    torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=max_norm)
    optimizer.step()
    if check_status and y.is_cuda:
        from .autograd import eigh_status
        if eigh_status(y.device):
            raise RuntimeError("eigen-solver did not converge during the training step")
    return loss.detach(), parts


def make_optimizer(model, lr=5e-3, weight_decay=1e-3, capturable=False, with_head=False):
    """trainPhi.py:100-118: AdamW on the four layer lists at lr/2, cosine warm restarts (T_0=10, T_mult=2).
    with_head=True adds train.py:104-121's second group (every other parameter at the full lr).
    capturable=True keeps the optimizer state on the device so the step can live inside a CUDA graph."""
    admm_params = [p for name, p in model.named_parameters()
                   if any(prefix in name for prefix in ("phiLayers", "hLayers", "gLayers", "zLayers"))]
    groups = [{"params": admm_params, "lr": lr * 0.5}]
    if with_head:
        other = [p for p in model.parameters() if not any(p is q for q in admm_params)]
        if other:
            groups.append({"params": other, "lr": lr})
    optimizer = torch.optim.AdamW(groups, lr=lr, weight_decay=weight_decay, capturable=capturable)
    scheduler = torch.optim.lr_scheduler.CosineAnnealingWarmRestarts(optimizer, T_0=10, T_mult=2, eta_min=1e-6)
    return optimizer, scheduler


class GraphedTrainStep:
    """One training step (trainPhi.py:165-178: zero_grad, forward, loss, backward, gradient all-reduce, clip, AdamW)
    captured ONCE in a CUDA graph and replayed: at batch 256 the step is ~2000 small launches and launch bound when
    run eagerly.  Inputs are copied into static buffers; shapes are fixed (a ragged tail batch runs through the
    eager `train_step`).  The optimizer must be capturable (make_optimizer(..., capturable=True)).  The three eager
    warm-up steps PyTorch's capture needs are run on a snapshot of the model/optimizer state, which is restored
    before the capture, so graph mode takes exactly the same parameter trajectory as eager mode.  The learning rate
    is a capture-time constant (a Python float in the param groups): re-capture after `scheduler.step()` changes it;
    `fit` / `fit_admmnet` therefore use the eager `train_step`."""

    def __init__(self, model, criterion, optimizer, example, max_norm=1.0, group=None):
        import copy
        self.model, self.criterion, self.optimizer, self.max_norm, self.group = model, criterion, optimizer, max_norm, group
        self.static = [t.clone() for t in example]
        snap_m = copy.deepcopy(model.state_dict())
        snap_o = copy.deepcopy(optimizer.state_dict())
        model.train()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                self._body()
        torch.cuda.current_stream().wait_stream(side)
        model.load_state_dict(snap_m)
        self._restore_optimizer(optimizer, snap_o)
        optimizer.zero_grad(set_to_none=True)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = self._body()

    @staticmethod
    def _restore_optimizer(optimizer, snap):
        """Put the optimizer back to the snapshot IN PLACE.  The state tensors the warm-up created (step, exp_avg,
        exp_avg_sq) must stay alive: were they dropped (optimizer.load_state_dict of a fresh snapshot), the captured
        step would lazily re-create and zero them inside the graph, i.e. on every replay."""
        params = [p for g in optimizer.param_groups for p in g["params"]]
        with torch.no_grad():
            for idx, p in enumerate(params):
                st = optimizer.state.get(p)
                if not st:
                    continue
                old = snap["state"].get(idx)
                for k, v in st.items():
                    if torch.is_tensor(v):
                        if old is not None and k in old:
                            v.copy_(old[k])
                        else:
                            v.zero_()
                    elif old is not None and k in old:
                        st[k] = old[k]
        for g, og in zip(optimizer.param_groups, snap["param_groups"]):
            for k, v in og.items():
                if k != "params":
                    g[k] = v

    def _body(self):
        y, b, sigma, phi_true = self.static
        self.optimizer.zero_grad(set_to_none=True)
        phi = self.model(y, b, sigma)
        loss, _ = self.criterion(phi, phi_true)
        loss.backward()
        allreduce_gradients(self.model, self.group, weight=y.shape[0])
        torch.nn.utils.clip_grad_norm_(self.model.parameters(), max_norm=self.max_norm)
        self.optimizer.step()
        return loss.detach()

    def __call__(self, y, b, sigma, phi_true):
        for dst, src in zip(self.static, (y, b, sigma, phi_true)):
            dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.loss


def param_rmse(est, true, L_true):
    """train.py:279-294 without the per-sample Python loop: for every sample with L > 0 true targets
    sqrt(mean((est[:L] - true[:L])^2)); returns the vector of those values (train.py averages them with np.mean)."""
    Lmax = est.shape[1]
    L = L_true.reshape(-1, 1).to(est.device)
    mask = (torch.arange(Lmax, device=est.device).unsqueeze(0) < L).to(est.dtype)
    mse = (mask * (est - true) ** 2).sum(1) / L.clamp(min=1).reshape(-1).to(est.dtype)
    return torch.sqrt(mse)[L.reshape(-1) > 0]


def detection_counts(confidences, L_true, threshold=0.5):
    """train.py:409-421: per sample detected = #(confidence > threshold); TP += min(L, detected) when both are
    positive, FP += max(detected - L, 0), FN += max(L - detected, 0).  Returns (tp, fp, fn) as Python ints."""
    L = L_true.reshape(-1).to(confidences.device).long()
    det = (confidences > threshold).sum(1)
    tp = torch.where((L > 0) & (det > 0), torch.minimum(L, det), torch.zeros_like(L)).sum()
    fp = (det - L).clamp(min=0).sum()
    fn = (L - det).clamp(min=0).sum()
    return int(tp), int(fp), int(fn)


def detection_scores(tp, fp, fn):
    """train.py:437-443: precision, recall, F1 (0 when undefined)."""
    precision = tp / (tp + fp) if (tp + fp) > 0 else 0
    recall = tp / (tp + fn) if (tp + fn) > 0 else 0
    f1 = 2 * precision * recall / (precision + recall) if (precision + recall) > 0 else 0
    return precision, recall, f1


def _admmnet_loss(model, criterion, batch, device):
    y, b, tau, f, L, sigma = (t.to(device, non_blocking=True) for t in batch)
    tau_est, f_est, conf, phi = model(y, b, sigma)
    total, parts = criterion({"tau_est": tau_est, "f_est": f_est, "confidences": conf, "phi_final": phi},
                             {"tau_true": tau, "f_true": f, "L_true": L, "y": y, "b": b})
    return total, (tau_est, f_est, conf, tau, f, L)


def evaluate_admmnet(model, criterion, split, batch_size=256, device="cuda", detection=False):
    """Validation / test pass of train.py (217-294 and 372-434): mean batch loss, tau / f RMSE over the samples with
    targets, and (test) the detection statistics.  `split` is a tuple of dataset.load_split."""
    model.eval()
    sel = (split[0], split[1], split[2], split[3], split[5], split[6])       # y, b, tau, f, L, sigma
    tot, nb, te, fe = 0.0, 0, [], []
    tp = fp = fn = 0
    with torch.no_grad():
        for lo in range(0, sel[0].shape[0], batch_size):
            loss, (tau_est, f_est, conf, tau, f, L) = _admmnet_loss(model, criterion, tuple(t[lo:lo + batch_size] for t in sel), device)
            tot += float(loss)
            nb += 1
            te.append(param_rmse(tau_est, tau, L))
            fe.append(param_rmse(f_est, f, L))
            if detection:
                a, b_, c = detection_counts(conf, L)
                tp, fp, fn = tp + a, fp + b_, fn + c
    te, fe = torch.cat(te), torch.cat(fe)
    out = {"loss": tot / max(nb, 1), "tau_rmse": float(te.mean()) if te.numel() else 0.0,
           "f_rmse": float(fe.mean()) if fe.numel() else 0.0}
    if detection:
        p, r, f1 = detection_scores(tp, fp, fn)
        out.update(precision=p, recall=r, f1_score=f1,
                   detection_stats={"true_positive": tp, "false_positive": fp, "false_negative": fn})
    return out


def fit_admmnet(model, train, val, test, config, device="cuda", group=None, log=print):
    """train.py:160-447 for the full ADMMNet (tau/f/confidence head, BasicANMLoss): epochs of train / validate with tau
    and f RMSE / cosine-restart scheduler step / checkpoint on the best validation loss / early stopping with patience
    10, then the test pass on the best checkpoint with precision, recall and F1 at confidence 0.5.  Data parallel like
    `fit`: every global batch is split over the ranks, one flat gradient all-reduce per step.  Returns
    (history, test_result); with config['log_dir'] both are also written as training_history.json / test_result.json."""
    import json
    import os

    from .autograd import BasicANMLoss
    from .dataset import load_checkpoint, save_checkpoint
    distributed = dist.is_available() and dist.is_initialized()
    rank = dist.get_rank(group) if distributed else 0
    world = dist.get_world_size(group) if distributed else 1
    criterion = BasicANMLoss()
    optimizer, scheduler = make_optimizer(model, config.get("lr", 1e-3), config.get("weight_decay", 1e-3), with_head=True)
    start_epoch, best_val, patience, bad = 0, float("inf"), config.get("patience", 10), 0
    ck_path = os.path.join(config["checkpoint_dir"], "best_model.pth") if config.get("checkpoint_dir") else None
    if ck_path and os.path.exists(ck_path):
        start_epoch, best_val, _ = load_checkpoint(ck_path, model, optimizer, scheduler, map_location=device)
    history = {"train_loss": [], "val_loss": [], "tau_rmse": [], "f_rmse": [], "lr": []}
    sel = lambda t: (t[0], t[1], t[2], t[3], t[5], t[6])                      # y, b, tau, f, L, sigma
    gen = torch.Generator().manual_seed(config.get("seed", 0))
    bs = config.get("batch_size", 256)
    for epoch in range(start_epoch, config.get("epochs", 100)):
        model.train()
        order = torch.randperm(train[0].shape[0], generator=gen)
        tot, nb = 0.0, 0
        for batch in _batches(sel(train), bs, order, rank, world):
            optimizer.zero_grad()
            nloc = batch[0].shape[0]
            if nloc:
                loss, _ = _admmnet_loss(model, criterion, batch, device)
                loss.backward()
                tot += float(loss)
                nb += 1
            allreduce_gradients(model, group, weight=nloc)
            torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=1.0)
            optimizer.step()
        ev = evaluate_admmnet(model, criterion, val, bs, device)
        history["train_loss"].append(tot / max(nb, 1))
        history["val_loss"].append(ev["loss"])
        history["tau_rmse"].append(ev["tau_rmse"])
        history["f_rmse"].append(ev["f_rmse"])
        history["lr"].append(optimizer.param_groups[0]["lr"])
        log(f"epoch {epoch + 1}: train {history['train_loss'][-1]:.6f}  val {ev['loss']:.6f}  "
            f"tau RMSE {ev['tau_rmse']:.6f}  f RMSE {ev['f_rmse']:.6f}  lr {history['lr'][-1]:.2e}")
        scheduler.step()
        if ev["loss"] < best_val:
            best_val, bad = ev["loss"], 0
            if ck_path and rank == 0:
                save_checkpoint(ck_path, epoch, model, optimizer, scheduler, best_val, config, history)
        else:
            bad += 1
        if config.get("log_dir") and rank == 0:
            with open(os.path.join(config["log_dir"], "training_history.json"), "w") as fh:
                json.dump({k: [float(x) for x in v] for k, v in history.items()}, fh, indent=2)
        if bad >= patience:
            break
    if ck_path and os.path.exists(ck_path):
        load_checkpoint(ck_path, model, map_location=device)
    res = evaluate_admmnet(model, criterion, test, bs, device, detection=True)
    test_result = {"test_loss": res["loss"], "tau_rmse": res["tau_rmse"], "f_rmse": res["f_rmse"],
                   "precision": res["precision"], "recall": res["recall"], "f1_score": res["f1_score"],
                   "detection_stats": res["detection_stats"]}
    if config.get("log_dir") and rank == 0:
        with open(os.path.join(config["log_dir"], "test_result.json"), "w") as fh:
            json.dump(test_result, fh, indent=2)
    return history, test_result


def _batches(tensors, batch_size, order, rank, world):
    """Each global batch of `batch_size` samples is split evenly over the ranks (data parallel); the ZLayer batch
    mean (admm_net.py:459) is then taken per rank, i.e. over batch_size/world samples — the statistic
    torch's DistributedDataParallel would also give the reference."""
    from .sharding import shard_range
    n = order.numel()
    for lo in range(0, n, batch_size):
        idx = order[lo:min(lo + batch_size, n)]
        a, b = shard_range(idx.numel(), rank, world)
        yield tuple(t[idx[a:b]] for t in tensors)      # possibly empty: every rank yields the same number of steps


def fit(model, train, val, config, device="cuda", group=None, log=print):
    """trainPhi.py:148-261: epochs of (train, validate, scheduler.step, checkpoint-on-best, early stopping with
    patience 10).  `train`/`val` are the tuples of dataset.load_split; config keys as in trainPhi.py:13-43
    (batch_size, epochs, lr, weight_decay, checkpoint_dir).  Under torch.distributed every rank calls fit with the
    same data and seed; gradients are averaged with one flat all-reduce per step."""
    import os

    from .autograd import PhiAlignmentLoss
    from .dataset import load_checkpoint, save_checkpoint
    distributed = dist.is_available() and dist.is_initialized()
    rank = dist.get_rank(group) if distributed else 0
    world = dist.get_world_size(group) if distributed else 1
    criterion = PhiAlignmentLoss()
    optimizer, scheduler = make_optimizer(model, config.get("lr", 5e-3), config.get("weight_decay", 1e-3))
    start_epoch, best_val, patience, bad = 0, float("inf"), config.get("patience", 10), 0
    ck_path = os.path.join(config["checkpoint_dir"], "best_model.pth") if config.get("checkpoint_dir") else None
    if ck_path and os.path.exists(ck_path):
        start_epoch, best_val, _ = load_checkpoint(ck_path, model, optimizer, scheduler, map_location=device)
    history = {"train_loss": [], "val_loss": [], "tau_rmse": [], "f_rmse": [], "lr": []}
    sel = lambda t: (t[0], t[1], t[6], t[7])                 # y, b, sigma, phi_true
    gen = torch.Generator().manual_seed(config.get("seed", 0))
    bs = config.get("batch_size", 256)
    for epoch in range(start_epoch, config.get("epochs", 1000)):
        order = torch.randperm(train[0].shape[0], generator=gen)
        tot, nb = torch.zeros((), device=device), 0
        for y, b, s, pt in _batches(sel(train), bs, order, rank, world):
            y, b, s, pt = (t.to(device, non_blocking=True) for t in (y, b, s, pt))
            loss, _ = train_step(model, criterion, optimizer, y, b, s, pt, group=group)
            if y.shape[0]:
                tot += loss
                nb += 1
        train_loss = float(tot) / max(nb, 1)
        model.eval()
        vt, vb = torch.zeros((), device=device), 0
        with torch.no_grad():
            for y, b, s, pt in _batches(sel(val), bs, torch.arange(val[0].shape[0]), rank, world):
                if y.shape[0] == 0:
                    continue
                y, b, s, pt = (t.to(device, non_blocking=True) for t in (y, b, s, pt))
                vt += criterion(model(y, b, s), pt)[0]
                vb += 1
        stats = torch.stack([vt, torch.tensor(float(vb), device=device)])
        if distributed:
            dist.all_reduce(stats, group=group)
        val_loss = float(stats[0] / stats[1].clamp(min=1))
        history["train_loss"].append(train_loss)
        history["val_loss"].append(val_loss)
        history["lr"].append(optimizer.param_groups[0]["lr"])
        log(f"epoch {epoch + 1}: train {train_loss:.6f}  val {val_loss:.6f}  lr {history['lr'][-1]:.2e}")
        scheduler.step()
        if val_loss < best_val:
            best_val, bad = val_loss, 0
            if ck_path and rank == 0:
                save_checkpoint(ck_path, epoch, model, optimizer, scheduler, best_val, config, history)
        else:
            bad += 1
        if bad >= patience:
            break
    return history
