"""y-slab decomposition of the grid over several GPUs (one process per GPU).

The reference is single-process; this is the host side of the new multi-GPU
path.  Rank r owns ny / nranks interior rows.  Its local arrays
are ordinary Fluids2d arrays -- same halo width in x, same row-major layout --
whose rows are a window of the global haloed array:

    [ghost rows of the south neighbour | owned rows | ghost rows of the north neighbour]

with GHOST = 8 ghost rows at an interface and the usual nh wall-halo rows at a
physical boundary.  All kernels run unchanged on the local arrays; libf2d
refreshes the ghost rows from their owners after every kernel (NCCL send/recv)
and all-reduces the CG / CFL scalars.
"""
import numpy as np

GHOST = 8          # must match Dist::G in csrc/engine.cuh

_comm = None       # (rank, world, nccl unique id bytes)


class Slab:
    """row bookkeeping of one rank"""

    def __init__(self, ny, nh, rank, nranks):
        if ny % nranks:
            raise ValueError(f"ny = {ny} is not divisible by {nranks} ranks")
        self.ny, self.nh, self.rank, self.nranks = ny, nh, rank, nranks
        self.own = ny // nranks
        self.y0 = rank * self.own                    # first owned interior row (global)
        self.gs = GHOST if rank > 0 else 0
        self.gn = GHOST if rank < nranks - 1 else 0
        self.below = self.gs if rank > 0 else nh     # local rows before the owned ones
        self.above = self.gn if rank < nranks - 1 else nh
        self.n2 = self.below + self.own + self.above
        self.ny_ctx = self.n2 - 2 * nh               # what the local context calls ny
        self.row0 = self.y0 + nh - self.below        # global haloed row of local row 0
        if nranks > 1 and self.own < 2 * GHOST:
            raise ValueError(f"{self.own} rows per rank is too few (need >= {2 * GHOST})")

    def window(self):
        """rows of the global haloed array held by this rank (ghosts included)"""
        return slice(self.row0, self.row0 + self.n2)

    def owned_local(self):
        """local rows this rank is the owner of (wall halos belong to the edge ranks)"""
        lo = 0 if self.rank == 0 else self.below
        hi = self.n2 if self.rank == self.nranks - 1 else self.below + self.own
        return slice(lo, hi)

    def owned_global(self):
        s = self.owned_local()
        return slice(self.row0 + s.start, self.row0 + s.stop)

    def scatter(self, a_global):
        """local copy (ghost rows included) of a global haloed array"""
        return a_global[self.window()].copy()

    def gather_into(self, a_global, a_local):
        a_global[self.owned_global()] = a_local[self.owned_local()]


def set_communicator(rank, nranks, unique_id):
    """register the NCCL identity every Mesh created afterwards will join"""
    global _comm
    _comm = (int(rank), int(nranks), bytes(unique_id))


def communicator():
    return _comm


def init_from_torch_distributed():
    """Share an NCCL unique id through an initialised torch.distributed process
    group (any backend) and register it.  Returns (rank, nranks)."""
    import torch.distributed as dist
    from . import _cabi
    rank, world = dist.get_rank(), dist.get_world_size()
    box = [_cabi.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    set_communicator(rank, world, box[0])
    return rank, world


def gather_global(slab, a_local, shape_global):
    """collect a field on every rank (host, torch.distributed) -- for tests and output"""
    import torch
    import torch.distributed as dist
    out = np.zeros(shape_global, dtype=a_local.dtype)
    pieces = [None] * slab.nranks
    dist.all_gather_object(pieces, (slab.owned_global().start, np.ascontiguousarray(a_local[slab.owned_local()])))
    for start, block in pieces:
        out[start:start + block.shape[0]] = block
    return out
