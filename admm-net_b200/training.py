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


def train_step(model, criterion, optimizer, y, b, sigma, phi_true, max_norm=1.0, group=None):
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
    return loss.detach(), parts


def make_optimizer(model, lr=5e-3, weight_decay=1e-3):
    """trainPhi.py:100-118: AdamW on the four layer lists at lr/2, cosine warm restarts (T_0=10, T_mult=2)."""
    admm_params = [p for name, p in model.named_parameters()
                   if any(prefix in name for prefix in ("phiLayers", "hLayers", "gLayers", "zLayers"))]
    optimizer = torch.optim.AdamW([{"params": admm_params, "lr": lr * 0.5}], lr=lr, weight_decay=weight_decay)
    scheduler = torch.optim.lr_scheduler.CosineAnnealingWarmRestarts(optimizer, T_0=10, T_mult=2, eta_min=1e-6)
    return optimizer, scheduler


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
