"""
sharding -- batch mode over several GPUs (BASELINE configs[3]).

B independent worlds / reward candidates are independent fixed points, so the batch
is split contiguously over the ranks of a `torch.distributed` group and every rank
runs its share with NO collective on the data path; results are gathered once at
the end if the caller wants them on every rank.
"""

import numpy as np


def shard_range(n_items, rank, world):
    """Contiguous, near-equal [begin, end) share of `n_items` for `rank`."""
    base, extra = divmod(n_items, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def _rank_world(group=None):
    import torch.distributed as dist
    if not dist.is_initialized():
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


def gather_rows(local, n_items, group=None):
    """Concatenate every rank's [b_local, ...] tensor into [n_items, ...] on all ranks
    (one all_gather of equal-sized, zero-padded pieces)."""
    import torch
    import torch.distributed as dist
    rank, world = _rank_world(group)
    if world == 1:
        return local
    sizes = [e - b for b, e in (shard_range(n_items, r, world) for r in range(world))]
    m = max(sizes)
    mine = torch.zeros((m,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    mine[:local.shape[0]] = local
    parts = [torch.empty_like(mine) for _ in sizes]
    dist.all_gather(parts, mine.contiguous(), group=group)
    return torch.cat([p[:s] for p, s in zip(parts, sizes)], 0)


def compute_expected_svf_batch_sharded(size, p_slips, p_initial, terminal, rewards, group=None, gather=True,
                                       **kwargs):
    """`maxent.compute_expected_svf_batch` over B grid worlds split across the group's GPUs.

    p_slips [B], rewards [B, S] are the GLOBAL batch (host arrays); every rank builds the tables of
    its own worlds only and runs them on its current CUDA device.  Returns (svf, grad) for the whole
    batch when gather=True, else for the local share; also the local [begin, end)."""
    import _irlb200 as E
    import maxent as M
    rank, world = _rank_world(group)
    B = len(p_slips)
    b0, b1 = shard_range(B, rank, world)
    ef = kwargs.pop("e_features", None)
    if ef is not None and getattr(ef, "ndim", 1) == 2:
        ef = ef[b0:b1]
    if getattr(p_initial, "ndim", 1) == 2:                       # per-world start distributions: this rank's rows
        p_initial = p_initial[b0:b1]
    if b1 == b0:
        # more ranks than worlds: this rank has nothing to run, but still takes part in the gather
        torch = E.require_cuda()
        S = size * size
        d = torch.zeros((0, S), dtype=torch.float64, device=E._dev())
        g = torch.zeros((0, S), dtype=torch.float64, device=E._dev()) if ef is not None else None
    else:
        tabs = E.gridworld_tables(size, np.asarray(p_slips)[b0:b1])
        d, g = M.compute_expected_svf_batch(tabs, p_initial, terminal, np.asarray(rewards)[b0:b1], e_features=ef,
                                            **kwargs)
    if gather:
        d = gather_rows(d, B, group)
        g = gather_rows(g, B, group) if g is not None else None
    return d, g, (b0, b1)
