"""CPU (gloo, world_size 2 and 3): the slab partition and the halo exchange of
pyellipticfd_b200.multigpu -- the host-side logic of the N > 1 path."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pyellipticfd_b200.multigpu import slab_partition, halo_exchange


def test_slab_partition_covers_rows():
    for nx, world in ((4095, 1), (4095, 2), (32766, 8), (17, 3)):
        parts = slab_partition(nx, world)
        assert parts[0][0] == 0 and parts[-1][1] == nx
        assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
        sizes = [hi - lo for lo, hi in parts]
        assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, nx, pitch, halo, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = slab_partition(nx, world)[rank]
        nloc = hi - lo
        # global field value = 1000*row + col ; halo rows start as -1
        view = torch.full((nloc + 2 * halo, pitch), -1.0, dtype=torch.float64)
        rows = torch.arange(lo, hi, dtype=torch.float64)[:, None]
        view[halo:halo + nloc] = 1000 * rows + torch.arange(pitch, dtype=torch.float64)[None, :]
        halo_exchange(view, nloc, halo, rank, world)
        ok = True
        glob = lambda r: 1000.0 * r + torch.arange(pitch, dtype=torch.float64)
        for k in range(halo):
            top = lo - halo + k
            bot = hi + k
            exp_top = glob(top) if top >= 0 and rank > 0 else torch.full((pitch,), -1.0, dtype=torch.float64)
            exp_bot = glob(bot) if bot < nx and rank < world - 1 else torch.full((pitch,), -1.0, dtype=torch.float64)
            ok &= bool(torch.equal(view[k], exp_top)) and bool(torch.equal(view[halo + nloc + k], exp_bot))
        # allreduce(max) of a per-rank pair, as the solvers do for (max|F|, max|dU|)
        red = torch.tensor([float(rank), float(world - rank)], dtype=torch.float64)
        dist.all_reduce(red, op=dist.ReduceOp.MAX)
        ok &= red.tolist() == [float(world - 1), float(world)]
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.timeout(120)
def test_halo_exchange_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, 37, 48, 4, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(100)
        assert p.exitcode == 0
    res = dict(q.get(timeout=5) for _ in range(world))
    assert all(res.values()) and len(res) == world


def _gather_worker(rank, world, port, n, q):
    from pyellipticfd_b200.multigpu import gather_rows
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        parts = slab_partition(n, world)
        lo, hi = parts[rank]
        local = torch.arange(lo, hi, dtype=torch.float64) * 2.0
        full = gather_rows(local, [b - a for a, b in parts])
        ok = bool(torch.equal(full, torch.arange(n, dtype=torch.float64) * 2.0))
        loc2 = torch.stack([local, -local], 1)
        full2 = gather_rows(loc2, [b - a for a, b in parts])
        ok &= full2.shape == (n, 2) and bool(torch.equal(full2[:, 1], -full))
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 10), (3, 10), (2, 7)])
@pytest.mark.timeout(120)
def test_point_range_gather_gloo(world, n):
    """Point-cloud sharding: ragged point ranges re-assembled by one padded all-gather."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gather_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(100)
        assert p.exitcode == 0
    res = dict(q.get(timeout=5) for _ in range(world))
    assert all(res.values()) and len(res) == world


def _krylov_worker(rank, world, port, n, q):
    """bicgstab_replicated with a row-sharded matrix (every rank applies its rows, one padded
    all-gather re-assembles the product) against SciPy's bicgstab on the whole matrix."""
    import scipy.sparse as sp
    from scipy.sparse.linalg import bicgstab
    from pyellipticfd_b200.multigpu import bicgstab_replicated, gather_rows
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(0)                       # same matrix on every rank
        # an M-matrix like the policy Jacobians: identity rows + weighted second differences
        rows, cols, vals = [], [], []
        for i in range(n):
            if i % 5 == 0 or i in (0, n - 1):
                rows.append(i); cols.append(i); vals.append(1.0)
            else:
                a, b = rng.uniform(0.5, 2.0, 2)
                j0, j1 = max(i - rng.integers(1, 4), 0), min(i + rng.integers(1, 4), n - 1)
                rows += [i, i, i]; cols += [j0, j1, i]; vals += [-a, -b, a + b]
        A = sp.coo_matrix((vals, (rows, cols)), shape=(n, n)).tocsr()
        b = rng.standard_normal(n)
        parts = slab_partition(n, world)
        lo, hi = parts[rank]
        sizes = [h - l for l, h in parts]
        Aloc = torch.from_numpy(A[lo:hi].toarray())

        def matvec(x):
            return gather_rows(Aloc @ x, sizes)

        dinv = torch.from_numpy(1.0 / A.diagonal())
        x, info, its = bicgstab_replicated(matvec, torch.from_numpy(b), dinv, 1e-10)
        calls = [0]
        xs, infos = bicgstab(A, b, rtol=1e-10, atol=0.0, M=sp.diags(1.0 / A.diagonal()),
                             callback=lambda xk: calls.__setitem__(0, calls[0] + 1))
        ok = info == 0 and infos == 0 and abs(its - calls[0]) <= 2
        ok &= bool(np.abs(x.numpy() - xs).max() < 1e-8 * np.abs(xs).max())
        ok &= bool(np.abs(A.dot(x.numpy()) - b).max() < 1e-8)
        # zero right-hand side, and the breakdown exit of a singular system stays finite
        x0, info0, its0 = bicgstab_replicated(matvec, torch.zeros(n, dtype=torch.float64), dinv, 1e-10)
        ok &= info0 == 0 and its0 == 0 and float(x0.abs().max()) == 0.0
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 60), (3, 47)])
@pytest.mark.timeout(120)
def test_replicated_bicgstab_gloo(world, n):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_krylov_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(100)
        assert p.exitcode == 0
    res = dict(q.get(timeout=5) for _ in range(world))
    assert all(res.values()) and len(res) == world
