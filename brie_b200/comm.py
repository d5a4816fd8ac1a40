"""NCCL communicator of an event-sharded fit, created once per torch.distributed process group.

torch.distributed is the plumbing (rendezvous: it ships rank 0's 128-byte ncclUniqueId to the other
ranks); the data-path all-reduce of the shared per-cell gradients is issued by libbrie_b200.so itself,
in stream order inside `brie_fit_run_steps` (csrc/brie_comm.cu) -- no host round trip per step.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib

_COMMS = {}


class Comm:
    def __init__(self, handle, rank, world):
        self.h, self.rank, self.world = handle, rank, world

    def allreduce_count(self):
        return int(_lib.load().brie_comm_allreduce_count(self.h))


def get(dist_group=None):
    """The library-side communicator spanning `dist_group` (default: WORLD); collective on first use."""
    import torch.distributed as dist
    if dist_group is dist.group.WORLD:
        dist_group = None
    key = id(dist_group) if dist_group is not None else 0
    if key in _COMMS:
        return _COMMS[key]
    lib = _lib.load()
    rank, world = dist.get_rank(dist_group), dist.get_world_size(dist_group)
    buf = np.zeros(128, np.uint8)
    if rank == 0:
        _lib.check(lib.brie_comm_unique_id(buf.ctypes.data))
    dev = torch.device("cuda", torch.cuda.current_device())
    t = torch.from_numpy(buf).to(dev)
    src = dist.get_global_rank(dist_group, 0) if dist_group is not None else 0
    dist.broadcast(t, src=src, group=dist_group)
    buf = t.cpu().numpy()
    h = C.c_void_p()
    _lib.check(lib.brie_comm_create(buf.ctypes.data, rank, world, C.byref(h)))
    _COMMS[key] = Comm(h, rank, world)
    return _COMMS[key]


def destroy_all():
    lib = _lib.load()
    for c in _COMMS.values():
        lib.brie_comm_destroy(c.h)
    _COMMS.clear()
