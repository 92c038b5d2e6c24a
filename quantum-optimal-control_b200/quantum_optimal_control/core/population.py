"""Multi-GPU population sweeps (ours; the reference is single-device, main_grape/grape.py:106-109).

Problem instances (random seeds / initial guesses) are independent, so the batch dimension is
sharded across ranks -- one process per GPU -- with NO communication inside the GRAPE iteration.
The only exchange is one all-gather of the per-instance final losses when the caller asks for the
population summary, followed by a broadcast of the winner's pulse from its owner.
Works with any torch.distributed backend (NCCL on GPUs; gloo in the CPU tests).
"""
import numpy as np


def shard_bounds(total, rank, world):
    """Contiguous [lo, hi) slice of ``total`` instances owned by ``rank`` (sizes differ by at most 1)."""
    base, extra = divmod(int(total), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_losses(local_losses, total, group=None):
    """All-gather the per-instance losses of every rank -> tensor [total] in global instance order."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return local_losses.clone()
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = [shard_bounds(total, r, world)[1] - shard_bounds(total, r, world)[0] for r in range(world)]
    width = max(sizes)
    pad = torch.full((width,), float('inf'), dtype=local_losses.dtype, device=local_losses.device)
    pad[:sizes[rank]] = local_losses
    out = torch.empty(world * width, dtype=local_losses.dtype, device=local_losses.device)
    dist.all_gather_into_tensor(out, pad, group=group)           # the ONE collective of a sweep
    return torch.cat([out[r * width:r * width + sizes[r]] for r in range(world)])


def select_best(all_losses, local_uks, local_Uf, total, group=None):
    """Index of the lowest loss and that instance's (uks, U_final), broadcast from its owner."""
    import torch
    import torch.distributed as dist
    best = int(torch.argmin(all_losses).item())
    if not (dist.is_available() and dist.is_initialized()):
        return best, local_uks[best], local_Uf[best]
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    owner = next(r for r in range(world) if shard_bounds(total, r, world)[0] <= best < shard_bounds(total, r, world)[1])
    lo = shard_bounds(total, owner, world)[0]
    dev = all_losses.device
    uks = torch.empty(local_uks.shape[1:], dtype=torch.float64, device=dev)
    Uf = torch.empty(local_Uf.shape[1:], dtype=torch.complex128, device=dev)
    if rank == owner:
        uks.copy_(torch.as_tensor(local_uks[best - lo]))
        Uf.copy_(torch.as_tensor(local_Uf[best - lo]))
    Ufr = torch.view_as_real(Uf)
    dist.broadcast(uks, src=owner, group=group)
    dist.broadcast(Ufr, src=owner, group=group)
    return best, uks.cpu().numpy(), torch.view_as_complex(Ufr).cpu().numpy()


def grape_population(grape_fn, args, initial_guesses, group=None, **kwargs):
    """Run ``Grape`` on this rank's shard of ``initial_guesses`` [B,K,T]; return
    dict(best, loss[B], uks, U_final) identical on every rank."""
    import torch
    import torch.distributed as dist
    total = len(initial_guesses)
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if total < world:
        # an empty shard cannot build an engine (B = 0) and the other ranks would wait in the all-gather forever
        raise ValueError("population of %d instances on %d ranks: every rank needs at least one instance" % (total, world))
    lo, hi = shard_bounds(total, rank, world)
    uks, Uf, losses = grape_fn(*args, initial_guess=np.asarray(initial_guesses)[lo:hi], return_losses=True, **kwargs)
    dev = torch.device('cuda', torch.cuda.current_device()) if torch.cuda.is_available() else torch.device('cpu')
    all_losses = gather_losses(torch.as_tensor(losses, dtype=torch.float64, device=dev), total, group)
    best, buks, bUf = select_best(all_losses, np.asarray(uks), np.asarray(Uf), total, group)
    return dict(best=best, loss=all_losses.cpu().numpy(), uks=buks, U_final=bUf)
