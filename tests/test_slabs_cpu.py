"""Host-side logic of the slab decomposition: row bookkeeping, scatter / gather,
and the gloo plumbing (world_size 2, CPU only)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("ny,nranks", [(64, 1), (64, 2), (256, 4), (2048, 8)])
def test_slab_rows_tile_the_global_array(ny, nranks):
    from fluids2d_b200.slabs import GHOST, Slab
    nh = 3
    covered = np.zeros(ny + 2 * nh, dtype=int)
    for r in range(nranks):
        s = Slab(ny, nh, r, nranks)
        assert s.n2 == s.below + s.own + s.above and s.ny_ctx == s.n2 - 2 * nh
        assert s.below == (nh if r == 0 else GHOST) and s.above == (nh if r == nranks - 1 else GHOST)
        w = s.window()
        assert 0 <= w.start and w.stop <= ny + 2 * nh
        covered[s.owned_global()] += 1
        # the first owned interior row sits at local row `below`
        assert w.start + s.below == s.y0 + nh
    assert np.all(covered == 1)            # every global row has exactly one owner


def test_scatter_gather_roundtrip():
    from fluids2d_b200.slabs import Slab
    rng = np.random.default_rng(0)
    ny, nx, nh, nranks = 96, 20, 3, 3
    a = rng.standard_normal((ny + 2 * nh, nx + 2 * nh))
    out = np.zeros_like(a)
    for r in range(nranks):
        s = Slab(ny, nh, r, nranks)
        loc = s.scatter(a)
        assert loc.shape == (s.n2, nx + 2 * nh)
        s.gather_into(out, loc)
    assert np.array_equal(out, a)


def test_bad_splits_are_rejected():
    from fluids2d_b200.slabs import Slab
    with pytest.raises(ValueError):
        Slab(100, 3, 0, 3)
    with pytest.raises(ValueError):
        Slab(32, 3, 0, 4)


def _gloo_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from fluids2d_b200 import slabs
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ny, nx, nh = 64, 10, 3
    s = slabs.Slab(ny, nh, rank, world)
    g = np.arange((ny + 2 * nh) * (nx + 2 * nh), dtype=np.float64).reshape(ny + 2 * nh, nx + 2 * nh)
    loc = s.scatter(g)
    loc[s.owned_local()] *= 2.0                       # every rank updates the rows it owns
    # ghost rows are deliberately left stale: gather must only take owned rows
    out = slabs.gather_global(s, loc, g.shape)
    ok = bool(np.array_equal(out, 2.0 * g))
    box = [b"id-from-rank-0" if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)            # how the NCCL id travels
    q.put((rank, ok, box[0]))
    dist.barrier()
    dist.destroy_process_group()


def test_gather_over_gloo_world_size_2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29533
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(r[1] for r in res) and all(r[2] == b"id-from-rank-0" for r in res)
