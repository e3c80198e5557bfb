"""Batch sharding across GPUs (one process per GPU, torch.distributed for plumbing).

Problem instances are independent: rank r owns the contiguous slice
``shard_range(batch, r, world)`` and runs the identical persistent sweep on it -- no
data-path collective.  The one exchange is an all-gather of the first-step policy
``[K_0 | k_0]`` (nu x (nx+1) doubles per instance), what a receding-horizon consumer
needs replicated.
"""
from __future__ import annotations


def shard_range(batch, rank, world):
    """Contiguous, balanced partition of `batch` instances: [beg, end) of `rank`."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return rank * batch // world, (rank + 1) * batch // world


def pack_first_step_policy(torch, fb0, ff0, nu, nx, out=None):
    """fb0 [B][nu+nc+nx][nx] (knot 0 of OUT_FB), ff0 [B][nu+nc+nx] -> [B][nu][nx+1]."""
    B = fb0.shape[0]
    if out is None:
        out = torch.empty(B, nu, nx + 1, dtype=fb0.dtype, device=fb0.device)
    out[:, :, :nx] = fb0[:, :nu]
    out[:, :, nx] = ff0[:, :nu]
    return out


def all_gather_policy(torch, dist, pol, world, out=None):
    """All-gather of equally sized per-rank policy blocks -> [world*B][nu][nx+1]."""
    if world == 1:
        return pol
    if out is None:
        out = torch.empty((world * pol.shape[0],) + tuple(pol.shape[1:]), dtype=pol.dtype,
                          device=pol.device)
    dist.all_gather_into_tensor(out, pol.contiguous())
    return out
