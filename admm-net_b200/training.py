"""Data-parallel training step for the unrolled net (trainPhi.py:158-179 per batch) with a flat-buffer gradient
all-reduce (NCCL over NVLink on the GPU box, gloo in the CPU tests).  lambda_param's, step_adjust_net and the last
layer's H/G/Z parameters receive no gradient in the reference either: they are all-reduced as zeros,
so no `find_unused_parameters` machinery is needed."""
import torch
import torch.distributed as dist


def allreduce_gradients(model, group=None):
    """Average .grad over the ranks with ONE all-reduce of a flat fp32 buffer (132 230 floats = 0.53 MB at K=10)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    params = [p for p in model.parameters() if p.requires_grad]
    flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in params])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat /= dist.get_world_size(group)
    off = 0
    for p in params:
        nel = p.numel()
        if p.grad is not None:
            p.grad.copy_(flat[off:off + nel].view_as(p))
        off += nel


def train_step(model, criterion, optimizer, y, b, sigma, phi_true, max_norm=1.0, group=None):
    """optimizer.zero_grad -> forward -> loss -> backward -> all-reduce -> clip -> step  (trainPhi.py:165-178)."""
    model.train()
    optimizer.zero_grad()
    phi = model(y, b, sigma)
    loss, parts = criterion(phi, phi_true)
    loss.backward()
    allreduce_gradients(model, group)
    torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=max_norm)
    optimizer.step()
    return loss.detach(), parts
