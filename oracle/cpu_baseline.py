"""CPU baseline: the reference operator timed on host cores.

TEST / BENCH INFRASTRUCTURE.  ``impl="reference"``: the UNMODIFIED ``pyellipticfd.ddg.d2eigs``
(ddg.py:251-327), imported through ``oracle/refshim.py`` from /root/reference or from the
byte-for-byte copy ``oracle/ref_install.py`` leaves in the git-ignored ``oracle/_ref/`` (that is
what travels to the GPU box) -- the reference's operators accept any object with ``points``,
``pairs``, ``num_interior``, ``num_points``.  ``impl="port"``: ``ddg_oracle.d2eigs``, the NumPy
restatement (gathers + norms + segmented selection, recomputing the geometry on every call
like ddg.py:274-292).  The reference is single-threaded; to use every core the sample is split
into disjoint row chunks, one worker process each (valid because the reference groups pairs by
centre point, ddg.py:283-292).  No product code is imported here.
"""
import multiprocessing as mp
import os
import time
import types

import numpy as np

from . import ddg_oracle


def lattice_chunk(ny, r, table, scaling, b0, row_lo, nrows):
    """Duck-typed grid holding deep rows [row_lo, row_lo+nrows) of an (nx, ny) lattice
    plus the r rows either side they read.  Interior-first numbering: the chunk's deep
    points come first, then the remaining lattice points of the window."""
    H = int(np.abs(table).max())
    lo = row_lo - H
    rows = nrows + 2 * H
    ix, iy = np.meshgrid(np.arange(lo, lo + rows), np.arange(ny), indexing="ij")
    deep = (ix >= row_lo) & (ix < row_lo + nrows) & (iy >= r) & (iy < ny - r)
    order = np.concatenate([np.flatnonzero(deep.ravel()), np.flatnonzero(~deep.ravel())])
    newid = np.empty(order.size, np.int64)
    newid[order] = np.arange(order.size)
    P = np.stack([(ix.ravel() + 1) * scaling[0] + b0[0], (iy.ravel() + 1) * scaling[1] + b0[1]], 1)
    di, dj = ix[deep] - lo, iy[deep]
    I = di * ny + dj
    J0 = (di[:, None] + table[None, :, 0]) * ny + dj[:, None] + table[None, :, 1]
    J1 = (di[:, None] + table[None, :, 2]) * ny + dj[:, None] + table[None, :, 3]
    pairs = np.stack([np.broadcast_to(I[:, None], J0.shape).ravel(), J0.ravel(), J1.ravel()], 1)
    G = types.SimpleNamespace(points=P[order], pairs=newid[pairs], num_interior=int(deep.sum()),
                              num_points=order.size, dim=2)
    return G


def operator(impl):
    """``d2eigs(G, u)`` of the chosen implementation ("reference" | "port")."""
    if impl == "reference":
        from . import refshim
        return refshim.modules().ddg.d2eigs
    return ddg_oracle.d2eigs


def reference_available():
    from . import refshim
    return refshim.available()


def _work(args):
    ny, r, table, scaling, b0, row_lo, nrows, reps, seed, impl = args
    d2eigs = operator(impl)
    G = lattice_chunk(ny, r, np.asarray(table), np.asarray(scaling), np.asarray(b0), row_lo, nrows)
    u = np.random.default_rng(seed).standard_normal(G.num_points)
    d2eigs(G, u)                                  # warm-up (page faults, imports)
    t = time.perf_counter()
    for _ in range(reps):
        d2eigs(G, u)
    dt = (time.perf_counter() - t) / reps
    return G.num_interior, len(G.pairs), dt


def time_d2eigs(ny, r, table, scaling, b0, rows_per_worker=24, workers=None, reps=2, impl="port"):
    """Returns dict(points_per_s, pairs_per_s, cores, sample, seconds)."""
    workers = workers or os.cpu_count() or 1
    H = int(np.abs(np.asarray(table)).max())
    args = [(ny, r, np.asarray(table), tuple(scaling), tuple(b0),
             r + H + w * rows_per_worker, rows_per_worker, reps, w, impl) for w in range(workers)]
    if workers == 1:
        res = [_work(args[0])]
    else:
        with mp.get_context("fork").Pool(workers) as pool:      # run from a CUDA-free process
            res = pool.map(_work, args)
    pts = sum(x[0] for x in res)
    prs = sum(x[1] for x in res)
    wall = max(x[2] for x in res)
    what = ("the unmodified reference's ddg.d2eigs" if impl == "reference"
            else "NumPy port of ddg.d2eigs")
    return dict(points_per_s=pts / wall, pairs_per_s=prs / wall, cores=workers, seconds=wall,
                kind=impl,
                sample="%d workers x %d deep rows x %d cols (%d pairs), %s"
                       % (workers, rows_per_worker, ny - 2 * r, prs, what))


def _work_steps(args):
    """One worker of ``time_d2eigs_steps``: builds its row chunk once, then times every one of the
    ``nsteps`` calls of d2eigs on it."""
    ny, r, table, scaling, b0, row_lo, nrows, nsteps, seed, impl = args
    d2eigs = operator(impl)
    G = lattice_chunk(ny, r, np.asarray(table), np.asarray(scaling), np.asarray(b0), row_lo, nrows)
    u = np.random.default_rng(seed).standard_normal(G.num_points)
    d2eigs(G, u)                                  # page faults, imports
    dts = []
    for _ in range(nsteps):
        t = time.perf_counter()
        d2eigs(G, u)
        dts.append(time.perf_counter() - t)
    return G.num_interior, len(G.pairs), dts


def time_d2eigs_steps(ny, r, table, scaling, b0, workers, steps, warmup, impl, budget_s=150.0,
                      max_rows=32):
    """``warmup + steps`` timed steps for ``bench.py --impl reference``: every step is one call of
    d2eigs per worker process on that worker's chunk of deep rows (all workers at once, one per
    core).  The chunk height is chosen from a calibration call so that the whole run takes about
    ``budget_s`` seconds (1..max_rows rows per worker).  Step time = the slowest worker's."""
    table = np.asarray(table)
    H = int(np.abs(table).max())
    cal = time_d2eigs(ny, r, table, scaling, b0, 4, workers, 1, impl)      # 4 rows per worker
    per_row = cal["seconds"] / 4.0
    rows = int(budget_s / max(steps + warmup, 1) / per_row)
    rows = max(1, min(max_rows, rows))
    args = [(ny, r, table, tuple(scaling), tuple(b0), r + H + w * rows, rows, warmup + steps, w, impl)
            for w in range(workers)]
    if workers == 1:
        res = [_work_steps(args[0])]
    else:
        with mp.get_context("fork").Pool(workers) as pool:
            res = pool.map(_work_steps, args)
    pts = sum(x[0] for x in res)
    prs = sum(x[1] for x in res)
    step_s = [max(x[2][k] for x in res) for k in range(warmup, warmup + steps)]
    what = ("the unmodified reference's ddg.d2eigs" if impl == "reference"
            else "NumPy port of ddg.d2eigs")
    mean = sum(step_s) / len(step_s)
    return dict(points_per_s=pts / mean, pairs_per_s=prs / mean, cores=workers, seconds=mean,
                step_seconds_min=min(step_s), step_seconds_max=max(step_s), steps=steps,
                warmup=warmup, kind=impl,
                sample="%d workers x %d deep rows x %d cols (%d pairs) per step, %s"
                       % (workers, rows, ny - 2 * r, prs, what))


def time_ddi_cached(idx, w, u, nds):
    """The reference's cached ddi.d2eigs path (ddi.py:358-367): Nds SciPy CSR mat-vecs
    (5 nnz per row) followed by an argsort of the (N, Nds) array.  ``idx``/``w`` hold the
    rows of K sample directions ((K, N, 4) each); the mat-vec time is measured on those K
    matrices and scaled to Nds, the argsort is run at full size.  Single-threaded like the
    reference."""
    import scipy.sparse as sp
    K, N, _ = idx.shape
    n = u.shape[0]
    I = np.arange(N)
    mats = []
    for k in range(K):
        rows = np.concatenate([np.repeat(I, 4), I])
        cols = np.concatenate([idx[k].ravel().astype(np.int64), I])
        vals = np.concatenate([w[k].ravel(), -w[k].sum(1)])
        mats.append(sp.coo_matrix((vals, (rows, cols)), shape=(N, n)).tocsr())
    for M in mats:
        M.dot(u)
    t = time.perf_counter()
    cols_ = [M.dot(u) for M in mats]
    t_mv = (time.perf_counter() - t) / K
    d2u = np.empty((N, nds))
    for k in range(nds):
        d2u[:, k] = cols_[k % K] + 1e-9 * k
    t = time.perf_counter()
    arg = np.argsort(d2u)
    lmin, lmax = d2u[I, arg[:, 0]], d2u[I, arg[:, -1]]
    t_sort = time.perf_counter() - t
    return dict(seconds=nds * t_mv + t_sort, t_matvec=t_mv, t_argsort=t_sort,
                sample="%d of %d direction matrices (CSR, %d rows, 5 nnz/row) timed and scaled, "
                       "argsort of the full (%d, %d) array" % (K, nds, N, N, nds))


def main(argv=None):
    """``python -m oracle.cpu_baseline NY R ROWS_PER_WORKER WORKERS REPS [IMPL [STEPS WARMUP]]`` on a unit-square
    dyadic lattice; prints one JSON line.  IMPL = reference | port | auto (reference when the
    tree or its copy under oracle/_ref is there).  bench.py runs this in a fresh (CUDA-free)
    process so that the worker pool can fork safely.  The direction table comes from
    oracle/grid_oracle.py: no product module (and no product .so) is loaded."""
    import json
    import sys
    from . import grid_oracle
    a = (argv or sys.argv[1:])
    ny, r, rows, workers, reps = int(a[0]), int(a[1]), int(a[2]), int(a[3]), int(a[4])
    impl = a[5] if len(a) > 5 else "auto"
    ratio = impl.endswith("+ratio")            # also time port and reference on one core
    impl = impl.replace("+ratio", "")
    if impl == "auto":
        impl = "reference" if reference_available() else "port"
    table = grid_oracle.stencil_table(r)
    h = 1.0 / (ny + 1)
    if len(a) > 7:                             # ... STEPS WARMUP: the --impl reference arm
        out = time_d2eigs_steps(ny, r, table, (h, h), (0.0, 0.0), workers, int(a[6]), int(a[7]),
                                impl, max_rows=rows)
    else:
        out = time_d2eigs(ny, r, table, (h, h), (0.0, 0.0), rows, workers, reps, impl)
    out["D"] = int(table.shape[0])
    if impl == "reference" and ratio:
        # the port on ONE worker's sample beside it: per-core ratio of the two CPU arms
        one_ref = time_d2eigs(ny, r, table, (h, h), (0.0, 0.0), rows, 1, 1, "reference")
        one_port = time_d2eigs(ny, r, table, (h, h), (0.0, 0.0), rows, 1, 1, "port")
        out["port_over_reference_per_core"] = one_port["pairs_per_s"] / one_ref["pairs_per_s"]
        out["reference_pairs_per_s_per_core"] = one_ref["pairs_per_s"]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
