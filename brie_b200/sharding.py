"""Event sharding across the GPUs of one box.

Per-event parameters (Z_loc, Z_std_log, Wc, per-event intercept / sigma) are independent
when there are no gene features and the intercept is not per-cell -- the same fact the
reference uses to fit events in sequential batches (brie/models/model_wrap.py:241-260).
Each rank therefore owns a contiguous event range aligned to the reference's batch
("convergence group") size and no collective is needed on the data path.  With gene
features / per-cell intercepts the engine all-reduces the shared gradients (engine.py).
"""
import numpy as np


def event_shards(n_events, world, group_size=1):
    """[(start, stop)] per rank: contiguous, covering [0, n_events), boundaries on multiples
    of group_size, sizes differing by at most one group."""
    group_size = max(int(group_size), 1)
    n_groups = -(-n_events // group_size)
    base, extra = divmod(n_groups, world)
    out, g = [], 0
    for r in range(world):
        k = base + (1 if r < extra else 0)
        out.append((min(g * group_size, n_events), min((g + k) * group_size, n_events)))
        g += k
    return out


def gather_event_axis(local, axis, group=None):
    """All-gather numpy arrays that are sharded along `axis` (variable shard sizes)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    parts = [None] * dist.get_world_size(group)
    dist.all_gather_object(parts, local, group=group)
    return np.concatenate(parts, axis=axis)
