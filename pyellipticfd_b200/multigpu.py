"""Slab sharding of a regular grid over the GPUs of one box.

The reference is single-process; this is the B200 addition for
grids that do not fit, or are too slow on, one GPU.  The lattice is cut into
contiguous slabs of rows (x is the slow axis, rows of ``ny`` values are contiguous:
pointclasses.py:513-515).  Every rank owns one slab plus ``r`` halo rows on each
side in its state buffer; before an operator application the halo rows are
refreshed from the two neighbouring ranks.  On GPUs that is a PUSH over NVLink peer
memory (:class:`HaloChannel`, ``csrc/halo_push.cu``): each rank stores its edge rows
straight into the neighbours' staging buffers and raises a flag, two small kernels on
the rank's own stream, no collective.  (Fallback when peer memory cannot be mapped: one
NCCL all-gather of the edge rows; CPU tensors: gloo send/recv.)  Wall points are
replicated on every rank.  Infinity-norm / dot-product reductions of the solvers are
small NCCL all-reduces.
"""
import os

import numpy as np
import torch
import torch.distributed as dist

from ._context import GridContext, Slab
from .pointclasses import FDRegularGrid


def slab_partition(nx, world):
    """Rows [lo, hi) of every rank: equal slabs, remainder spread over the first ranks."""
    base, rem = divmod(nx, world)
    lo = [r * base + min(r, rem) for r in range(world)]
    hi = [lo[r] + base + (1 if r < rem else 0) for r in range(world)]
    return list(zip(lo, hi))


def halo_exchange(view, nx_local, halo, rank, world, group=None, method="auto"):
    """Refresh the halo rows of ``view`` ((nx_local + 2*halo) x pitch, contiguous rows).

    Sends the first/last ``halo`` owned rows to the previous/next rank and receives
    their counterparts into the halo rows.  Works for CPU tensors (gloo) and CUDA
    tensors (NCCL).  Slabs must hold at least ``halo`` rows."""
    if world == 1 or halo == 0:
        return
    H = halo
    if view.is_cuda and method != "p2p":
        # One small all-gather of every rank's first/last H owned rows (2*H*pitch doubles per
        # rank; 262 kB at ny = 4095, r = 4) instead of four point-to-point operations: a
        # single NCCL kernel over NVLink, no per-peer send/recv matching.
        send = torch.cat([view[H:2 * H], view[nx_local:nx_local + H]])
        got = torch.empty((world,) + tuple(send.shape), dtype=send.dtype, device=send.device)
        dist.all_gather_into_tensor(got, send, group=group)
        if rank > 0:
            view[0:H].copy_(got[rank - 1, H:2 * H])
        if rank < world - 1:
            view[nx_local + H:nx_local + 2 * H].copy_(got[rank + 1, 0:H])
        return
    ops = []
    if rank > 0:
        ops.append(dist.P2POp(dist.isend, view[H:2 * H], rank - 1, group))
        ops.append(dist.P2POp(dist.irecv, view[0:H], rank - 1, group))
    if rank < world - 1:
        ops.append(dist.P2POp(dist.isend, view[nx_local:nx_local + H], rank + 1, group))
        ops.append(dist.P2POp(dist.irecv, view[nx_local + H:nx_local + 2 * H], rank + 1, group))
    for req in dist.batch_isend_irecv(ops):
        req.wait()


def bind_to_gpu_numa_node(device_index):
    """Pin this process (one per GPU) to the CPUs of the NUMA node its GPU hangs off, so that the
    pinned host buffers it allocates afterwards are node-local (first touch) and the ranks do not
    all stream through node 0.  Best effort: returns the node number or None."""
    try:
        prop = torch.cuda.get_device_properties(device_index)
        bdf = "%04x:%02x:%02x.0" % (prop.pci_domain_id, prop.pci_bus_id, prop.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bdf).read())
        if node < 0:
            return None
        cpus = []
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.extend(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:
        pass
    return None


class HaloChannel(object):
    """Neighbour-only halo exchange through peer-mapped staging buffers (torch symmetric
    memory supplies the mapping; the copies and flags are ``pde_halo_push`` /
    ``pde_halo_wait_copy``).  One instance per (rank, halo, pitch)."""

    def __init__(self, halo, pitch, rank, world, device, group=None):
        import torch.distributed._symmetric_memory as symm
        from . import _lib
        self.lib, self.rank, self.world = _lib.lib(), rank, world
        self.halo, self.pitch = int(halo), int(pitch)
        n = int(self.lib.pde_halo_stage_doubles(self.halo, self.pitch))
        self.stage = symm.empty(n, dtype=torch.float64, device=device)
        self.hdl = symm.rendezvous(self.stage, group if group is not None else dist.group.WORLD)
        self.stage.zero_()
        torch.cuda.synchronize(device)
        dist.barrier()                       # nobody pushes before every flag word is zero
        ptrs = self.hdl.buffer_ptrs
        self.lo = int(ptrs[rank - 1]) if rank > 0 else None
        self.hi = int(ptrs[rank + 1]) if rank < world - 1 else None
        self.step = 0

    def exchange(self, view, nx_local):
        """Refresh the halo rows of ``view`` ((nx_local + 2*halo) x pitch); enqueues two kernels
        on the current stream.  Every rank must call it for the same sequence of vectors."""
        from ._lib import check, vp
        if view.stride(0) != self.pitch or view.shape[0] != nx_local + 2 * self.halo:
            raise ValueError("halo channel: unexpected view layout")
        self.step += 1
        st = vp(torch.cuda.current_stream().cuda_stream)
        check(self.lib.pde_halo_push(vp(view.data_ptr()), int(nx_local), self.halo, self.pitch,
                                     vp(self.stage.data_ptr()), vp(self.lo), vp(self.hi),
                                     self.step, st))
        check(self.lib.pde_halo_wait_copy(vp(view.data_ptr()), int(nx_local), self.halo, self.pitch,
                                          vp(self.stage.data_ptr()), int(self.lo is not None),
                                          int(self.hi is not None), self.step, st))


class ShardedGrid(object):
    """One rank's slab of a global FDRegularGrid."""

    def __init__(self, interior_shape, bounds, stencil_radius, rank, world, device, grid=None):
        self.rank, self.world = rank, world
        if grid is None and world > 1:
            from . import _lib
            # one process per GPU: share the host cores among the ranks' grid builders
            _lib.lib().pde_set_host_threads(max(1, (os.cpu_count() or 1) // world))
        self.G = grid if grid is not None else FDRegularGrid(list(interior_shape), bounds,
                                                             stencil_radius, interpolation=False)
        if len(self.G.interior_shape) != 2:
            raise NotImplementedError("slab sharding is implemented for 2-D regular grids")
        nx, ny = self.G.interior_shape
        lo, hi = slab_partition(nx, world)[rank]
        if world > 1 and hi - lo < stencil_radius:
            raise ValueError("slab thinner than the stencil radius")
        halo = stencil_radius if world > 1 else 0
        self.slab = Slab(lo, hi - lo, halo)
        self.ctx = GridContext(self.G, device, slab=self.slab)
        if world > 1 and os.environ.get("PDE_DYNAMIC_SCHEDULE", "0") == "1":
            from ._lib import check
            check(self.ctx.lib.pde_grid_set_dynamic_schedule(self.ctx.h, 1))
        # neighbour-only halo push over NVLink peer memory; PDE_HALO=nccl keeps the all-gather
        self.channel = None
        if (world > 1 and self.ctx.device.type == "cuda" and dist.is_initialized()
                and os.environ.get("PDE_HALO", "push") != "nccl"):
            try:
                self.channel = HaloChannel(halo, self.ctx.pitch, rank, world, self.ctx.device)
            except Exception as e:                      # peer mapping unavailable on this system
                import warnings
                warnings.warn("NVLink halo push unavailable (%s); using the NCCL all-gather" % e)
            ok = torch.tensor([1.0 if self.channel is not None else 0.0], device=self.ctx.device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)   # all ranks take the same path
            if float(ok) == 0.0:
                self.channel = None

    def lattice_view(self, S):
        c = self.ctx
        return S[:c.rows * c.pitch].view(c.rows, c.pitch)

    def exchange_halo(self, S):
        if self.channel is not None:
            self.channel.exchange(self.lattice_view(S), self.ctx.nx_local)
            return
        halo_exchange(self.lattice_view(S), self.ctx.nx_local, self.slab.halo, self.rank,
                      self.world)

    def overlapped(self, S, launch):
        """Run ``launch()`` (stencil launches that read the state ``S``) after refreshing the
        halo rows of ``S``.  With the NVLink push channel that is simply exchange + launch on
        one stream.  NCCL fallback: the all-gather goes to a communication stream, the interior
        rows [r, nx_local - r) are computed meanwhile on the current stream, the 2*r edge rows
        afterwards (both strips in one launch group)."""
        c, H = self.ctx, self.slab.halo
        if self.world == 1 or H == 0:
            return launch()
        if self.channel is not None:
            # push + wait-copy are two small kernels in stream order; the neighbours push at the
            # same point of their streams, so the wait is only the ranks' skew: one launch group
            self.exchange_halo(S)
            return launch()
        if c.nx_local <= 4 * H:
            self.exchange_halo(S)
            return launch()
        if not hasattr(self, "_comm"):
            self._comm = torch.cuda.Stream(device=c.device)
        main = torch.cuda.current_stream()
        self._comm.wait_stream(main)
        with torch.cuda.stream(self._comm):
            self.exchange_halo(S)
        try:
            c.set_row_window(H, c.nx_local - H)
            res = launch()
            main.wait_stream(self._comm)
            c.set_row_windows2(0, H, c.nx_local - H, c.nx_local)     # both edge strips, one launch
            launch()
        finally:
            c.set_row_window()
        return res

    def d2eigs(self, S, out, **kw):
        """Slab-parallel ddg.d2eigs on the state ``S`` (halo refreshed inside); ``out`` is the
        (lmin, lmax, pmin, pmax) tuple of preallocated state-layout outputs (None to skip)."""
        return self.overlapped(S, lambda: self.ctx.d2eigs(S, want_min=out[0] is not None or out[2] is not None,
                                                           want_max=out[1] is not None or out[3] is not None,
                                                           out=out, **kw))

    def d2eigs_host(self, u_host, want_min=True, want_max=True, out=None, chunk_rows=256):
        """Slab-parallel ``ddg.d2eigs`` VALUES for this rank's flat vector (owned rows, then all
        wall points) in pinned host memory, results in pinned host memory.  The slab's edge rows
        are uploaded first and pushed to the neighbours, then H2D copy, stencil kernels and D2H
        copy are pipelined over row chunks (``GridContext.d2eigs_host``).  Returns
        (lambda_min, lambda_max) pinned host tensors over the owned rows (None if not wanted)."""
        hook = self.exchange_halo if self.world > 1 and self.slab.halo else None
        return self.ctx.d2eigs_host(u_host, want_min=want_min, want_max=want_max,
                                    chunk_rows=chunk_rows, halo_hook=hook, out=out)

    def local_flat(self, fn):
        """Evaluate ``fn(points (n,2) tensor) -> values`` on this rank's points:
        owned interior rows first, then all wall points (the rank-local flat vector)."""
        c, G = self.ctx, self.G
        dev = c.device
        hx, hy = G.scaling
        xs = (torch.arange(c.x_offset + 1, c.x_offset + c.nx_local + 1, device=dev,
                           dtype=torch.float64) * hx + G.bounds[0][0])
        ys = torch.arange(1, c.ny + 1, device=dev, dtype=torch.float64) * hy + G.bounds[0][1]
        P = torch.stack([xs[:, None].expand(-1, c.ny), ys[None, :].expand(c.nx_local, -1)],
                        dim=2).reshape(-1, 2)
        W = torch.from_numpy(G.bdry_points).to(dev)
        return torch.cat([fn(P), fn(W)])

    def _step(self, V, g, dt, red):
        """One overlapped Euler step; ``red`` (2 doubles, device) accumulates the maxima."""
        Vn = self.ctx.zeros()
        self.overlapped(V, lambda: self.ctx.ce_euler_step(V, g, dt, out=(Vn, red)))
        return Vn

    # ---- distributed Newton / BiCGStab ------------------------------------------------------
    def _allreduce(self, t, op=None):
        if self.world > 1:
            dist.all_reduce(t, op=op if op is not None else dist.ReduceOp.SUM)
        return t

    def bicgstab(self, Jac, B, rtol, atol=0.0, maxiter=0, check_every=16):
        """Slab-parallel Jacobi-BiCGStab (solvers.py:179-181): the single-GPU passes, with a
        halo exchange of phat / shat before the two operator passes and one SUM all-reduce
        of <= 2 doubles after each of the four reduction passes.  Requires a
        Dirichlet-consistent right-hand side (zero on the replicated wall entries), which
        is what convex_envelope / pucci produce.  Returns (x, info, iterations)."""
        import ctypes as C
        from ._lib import check, vp
        c = self.ctx
        n = c.n_state
        if float(B[c.n_lattice:].abs().max()) != 0.0:
            raise ValueError("distributed solve needs a right-hand side that vanishes on the wall")
        dev = c.device
        work = torch.empty(8 * n, dtype=torch.float64, device=dev)
        x = c.zeros()
        scal = torch.zeros(int(c.lib.pde_bicgstab_scal_bytes()), dtype=torch.uint8, device=dev)
        nblk = c.rows + (c.n_band + c.nb) // 256 + 8 * 256 + 2
        partials = torch.zeros(3 * nblk, dtype=torch.float64, device=dev)
        sums = torch.zeros(3, dtype=torch.float64, device=dev)
        J = Jac._J
        phat, shat = work[4 * n:5 * n], work[6 * n:7 * n]

        def step(k):
            check(c.lib.pde_bicgstab_dist_step(
                c.h, C.byref(J), k, vp(B.data_ptr()), vp(x.data_ptr()), vp(work.data_ptr()),
                vp(scal.data_ptr()), vp(partials.data_ptr()), vp(sums.data_ptr()), float(rtol),
                float(atol), int(maxiter), vp(torch.cuda.current_stream().cuda_stream)))

        step(0)
        self._allreduce(sums)
        step(1)
        out = (C.c_int64 * 3)()
        while True:
            for _ in range(check_every):
                step(2)
                self.exchange_halo(phat)
                step(3)
                self._allreduce(sums)
                step(4)
                step(5)
                self._allreduce(sums)
                step(6)
                self.exchange_halo(shat)
                step(7)
                self._allreduce(sums)
                step(8)
                step(9)
                self._allreduce(sums)
                step(10)
            check(c.lib.pde_bicgstab_poll(vp(scal.data_ptr()), out,
                                          vp(torch.cuda.current_stream().cuda_stream)))
            if out[0]:
                break
        H = self.slab.halo
        if H:
            v = self.lattice_view(x)
            v[:H].zero_()
            v[H + c.nx_local:].zero_()
        return x, int(out[1]), int(out[2])

    def ce_newton(self, g, U0=None, solution_tol=1e-6, operator_tol=1e-6, max_iters=100,
                  krylov_rtol=None, stats=None, check_every=16):
        """convex_envelope.solve(..., fdmethod='grid', solver='newton') on the slabs: the
        loop of solvers.NewtonEulerLS (solvers.py:168-233) with slab-parallel residual,
        Krylov solve and reductions.  A failed solve / decrease test falls back to a bounded
        number of Euler steps.  ``g``, ``U0``: this rank's state tensors."""
        from .linop import PolicyJacobian
        c = self.ctx
        U = g.clone() if U0 is None else U0.clone()
        rho, p = 0.5, 2
        krylov = fallbacks = 0
        diff = float("inf")
        dt = None
        for i in range(int(max_iters) + 1):
            self.exchange_halo(U)
            F, P, red = c.ce_residual(U, g, policy=True)
            fmax = float(self._allreduce(red[:1].clone(), dist.ReduceOp.MAX))
            Jac = PolicyJacobian(c, P, None, cmin=-1.0)
            D, info, kit = self.bicgstab(Jac, -F, min(operator_tol, solution_tol)
                                         if krylov_rtol is None else krylov_rtol,
                                         check_every=check_every)
            krylov += kit
            ok = info <= 0
            if ok:
                self.exchange_halo(D)
                JD = Jac.matvec_state(D)
                # the norms of the decrease test are over OWNED rows: the halo rows of D hold the
                # neighbours' values (they were only needed for J d) and would be counted twice
                H = self.slab.halo
                if H:
                    v = self.lattice_view(D)
                    v[:H].zero_()
                    v[H + c.nx_local:].zero_()
                Un, dmax, gtd, dd = c.newton_update(U, D, F, JD)
                t = torch.tensor([gtd, dd], dtype=torch.float64, device=c.device)
                self._allreduce(t)
                m = self._allreduce(torch.tensor([dmax], dtype=torch.float64, device=c.device),
                                    dist.ReduceOp.MAX)
                gtd, dd, dmax = float(t[0]), float(t[1]), float(m[0])
                if gtd > -rho * np.sqrt(dd) ** p:
                    ok = False
            if ok:
                diff, U = dmax, Un
            else:
                fallbacks += 1
                if dt is None:
                    dt = 0.5 * max(self.G.min_interior_nb_dist, self.G.min_radius) ** 2
                U, diff, _, _ = self.ce_euler(U, g, dt, solution_tol, operator_tol, max_iters=256,
                                              check_every=64)
            if diff < solution_tol and fmax < operator_tol:
                break
        if stats is not None:
            stats.update(newton_iters=i + 1, krylov_iters=krylov, euler_fallbacks=fallbacks,
                         residual=fmax)
        return U, diff, i + 1

    # ---- distributed convex-envelope Euler iteration ------------------------------------
    def ce_euler(self, U, g, dt, solution_tol=1e-6, operator_tol=1e-6, max_iters=None,
                 check_every=16):
        """solvers.euler on the convex-envelope operator, slab-parallel: halo exchange,
        one fused step kernel per rank, MAX all-reduce of (max|F|, max|dU|) every
        ``check_every`` iterations is NOT enough to stop at the reference's iteration --
        so the two maxima of EVERY iteration are kept in a device ring and reduced in one
        all-reduce per batch; the stopping iteration is then found exactly and the state
        recomputed from the last checkpoint if the loop overshot."""
        c = self.ctx
        if max_iters is None:
            max_iters = self.G.num_points
        it = 0
        diff = float("inf")
        fmax = float("inf")
        while it < max_iters:
            n = int(min(check_every, max_iters - it))
            ckpt = U.clone()
            red = torch.zeros((n, 2), dtype=torch.float64, device=c.device)
            V = ckpt.clone()
            for k in range(n):
                V = self._step(V, g, dt, red[k])
            if self.world > 1:
                dist.all_reduce(red, op=dist.ReduceOp.MAX)
            h = red.cpu().numpy()
            ok = (h[:, 1] < solution_tol) & (h[:, 0] < operator_tol)
            if ok.any():
                stop = int(np.flatnonzero(ok)[0]) + 1
                V = ckpt
                scratch = torch.zeros(2, dtype=torch.float64, device=c.device)
                for k in range(stop):          # replay to the exact stopping iterate
                    V = self._step(V, g, dt, scratch)
                return V, float(h[stop - 1, 1]), it + stop, float(h[stop - 1, 0])
            U = V
            it += n
            diff, fmax = float(h[-1, 1]), float(h[-1, 0])
        return U, diff, it, fmax


# ---------------------------------------------------------------------------------------------
# point clouds: independent (point, direction) units, sharded by point range
# ---------------------------------------------------------------------------------------------

def gather_rows(local, sizes, group=None):
    """All-gather of per-rank row blocks of (possibly) different lengths ``sizes`` into one
    tensor, in rank order.  Blocks are padded to the longest one so that a single
    ``all_gather_into_tensor`` (one NCCL kernel over NVLink; gloo on CPU tensors) does it."""
    world = len(sizes)
    if world == 1:
        return local
    m = max(sizes)
    pad = local
    if local.shape[0] < m:
        pad = torch.zeros((m,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        pad[:local.shape[0]] = local
    got = torch.empty((world * m,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(got, pad.contiguous(), group=group)
    if all(s == m for s in sizes):
        return got
    return torch.cat([got[r * m:r * m + sizes[r]] for r in range(world)])


def bicgstab_replicated(matvec, b, dinv, rtol, atol=0.0, maxiter=None):
    """Jacobi-preconditioned BiCGStab with SciPy's statements (scipy/sparse/linalg/_isolve/
    iterative.py::bicgstab, the call of solvers.py:179-181; same order as csrc/krylov.cuh) on
    vectors that are REPLICATED on every rank: only ``matvec`` is distributed (every rank applies
    its rows and the slices are re-assembled by an all-gather), the vector algebra and the dot
    products are repeated on identical data, so all ranks take identical decisions without a
    reduction.  Meant for point clouds, whose vectors are small next to the per-rank ELL table.
    Returns ``(x, info, iterations)``."""
    n = b.numel()
    maxiter = 10 * n if not maxiter else int(maxiter)
    rhotol = float(np.finfo(np.float64).eps) ** 2
    bnrm2 = float(torch.linalg.vector_norm(b))
    x = torch.zeros_like(b)
    if bnrm2 == 0.0:
        return x, 0, 0
    atol = max(float(atol), float(rtol) * bnrm2)
    r = b.clone()
    rtilde = r.clone()
    p = v = None
    rho_prev = omega = alpha = 1.0
    for it in range(maxiter):
        if float(torch.linalg.vector_norm(r)) < atol:
            return x, 0, it
        rho = float(torch.dot(rtilde, r))
        if abs(rho) < rhotol:
            return x, -10, it
        if it > 0:
            if abs(omega) < rhotol:
                return x, -11, it
            beta = (rho / rho_prev) * (alpha / omega)
            p = p - omega * v          # iterative.py: p -= omega*v; p *= beta; p += r
            p = p * beta
            p = p + r
        else:
            p = r.clone()
        phat = dinv * p
        v = matvec(phat)
        rv = float(torch.dot(rtilde, v))
        if rv == 0.0:
            return x, -11, it
        alpha = rho / rv
        r = r - alpha * v              # this is s
        if float(torch.linalg.vector_norm(r)) < atol:
            x = x + alpha * phat
            return x, 0, it + 1
        shat = dinv * r
        t = matvec(shat)
        omega = float(torch.dot(t, r)) / float(torch.dot(t, t))
        x = x + alpha * phat
        x = x + omega * shat
        r = r - omega * t
        rho_prev = rho
    return x, maxiter, maxiter


class _ShardedEllJacobian(object):
    """Rows [lo, hi) of the point-cloud policy Jacobian (convex_envelope.py:48-54,
    pucci.py:136-139) on this rank; vectors replicated; ``matvec_state`` re-assembles the rows
    of all ranks (boundary rows are the identity)."""

    def __init__(self, cloud, kmin, kmax=None, cmin=1.0, cmax=0.0):
        self.cloud, self.ctx = cloud, cloud
        self.kmin, self.kmax, self.cmin, self.cmax = kmin, kmax, float(cmin), float(cmax)
        self.active = kmin >= 0
        self.k0 = torch.where(self.active, kmin, torch.zeros_like(kmin))

    def _rows(self, X):
        c = self.cloud
        y = self.cmin * c.table.apply(self.k0, X)
        if self.kmax is not None:
            y = y + self.cmax * c.table.apply(self.kmax, X)
        return torch.where(self.active, y, X[c.lo:c.hi])

    def matvec_state(self, X, out=None):
        c = self.cloud
        Y = torch.cat([c.gather(self._rows(X)), X[c.n_interior:]])
        if out is not None:
            out.copy_(Y)
            return out
        return Y

    def diagonal(self):
        c = self.cloud
        t = c.table

        def dg(k):      # -(sum of the four weights), the centre entry of the row (ell.cu::ell_diag)
            w = t.w[k.long(), torch.arange(t.N, device=t.device)]
            return -(w[:, 0] + w[:, 1] + w[:, 2] + w[:, 3])

        d = self.cmin * dg(self.k0)
        if self.kmax is not None:
            d = d + self.cmax * dg(self.kmax)
        d = torch.where(self.active, d, torch.ones_like(d))
        return torch.cat([c.gather(d), torch.ones(c.n_points - c.n_interior, dtype=d.dtype,
                                                  device=d.device)])

    def bicgstab_state(self, B, rtol, atol=0.0, maxiter=0, check_every=32):
        return bicgstab_replicated(self.matvec_state, B, 1.0 / self.diagonal(), rtol, atol, maxiter)


class ShardedCloud(object):
    """One rank's share of the ``ddi`` / ``ddf`` operator on a point cloud: the ELL table of
    the interior points ``[lo, hi)`` only (``Nds * 48`` bytes per point stay local), the
    function values ``u`` replicated (16 MB at 2 M points) and re-assembled after an update
    by one all-gather of the ranks' slices.

    ``ce_solve`` / ``pucci_solve`` (round 2) run ``convex_envelope.solve`` / ``pucci.solve`` with
    ``fdmethod='interpolate'|'froese'`` (convex_envelope.py:35-40, pucci.py:60-160) on the
    sharded table: the residual -- all ``Nds`` rows of every point, the expensive part -- is
    evaluated on the rank's points only, the policy Jacobian applies the rank's rows, and the
    Newton / Euler loops are those of the single-GPU path (``solvers.newton_state``) on
    replicated vectors."""

    def __init__(self, G, rank, world, device, scheme="ddi"):
        from . import ddi, ddf, _ell
        mod = {"ddi": ddi, "ddf": ddf}[scheme]
        self.G, self.rank, self.world = G, rank, world
        self.parts = slab_partition(int(G.num_interior), world)
        self.lo, self.hi = self.parts[rank]
        self.sizes = [b - a for a, b in self.parts]
        self.nds = int(mod._num_directions(G))
        self.V = _ell.directions(self.nds)
        self.table = _ell.EllTable(G, self.V, scheme == "ddf", device=device,
                                   rows=(self.lo, self.hi))
        self.device = self.table.device
        self.n_interior = int(G.num_interior)
        self.n_points = self.table.n_points

    def d2eigs(self, u, want_k=False):
        """(lmin, lmax, kmin, kmax) of this rank's points; ``u``: values on ALL points."""
        return self.table.d2eigs(u, want_k=want_k)

    def gather(self, local):
        """Values over all interior points from every rank's slice."""
        return gather_rows(local, self.sizes)

    def d2eigs_all(self, u):
        lmin, lmax, _, _ = self.d2eigs(u)
        return self.gather(lmin), self.gather(lmax)

    # ---- what solvers.newton_state / euler_state ask of a context ------------------------------
    def n_flat(self):
        return self.n_points

    def newton_update(self, U, D, F, JD):
        """solvers.py:188-197 on replicated vectors (identical on every rank)."""
        Un = U + D
        return (Un, float((U - Un).abs().max()), float(torch.dot(F, JD)), float(torch.dot(D, D)))

    def residual(self, W, g, mode, A=(0.0, 0.0), policy=True):
        """Sharded ``pde_ell_residual``: mode 0 convex envelope (convex_envelope.py:44-46), mode 1
        Pucci (pucci.py:48,128,141).  Same row arithmetic and tie rules as the single-GPU kernel
        (the rows come from the same ``k_ell_d2eigs``).  Returns (F, kmin, kmax, max|F|)."""
        N, lo, hi = self.n_interior, self.lo, self.hi
        lmin, lmax, kmin, kmax = self.table.d2eigs(W, want_k=policy)
        f = W[lo:hi] - g[lo:hi]
        if mode == 0:
            nl = -lmin
            active = nl > f
            f = torch.where(active, nl, f)
            if policy:
                kmin = torch.where(active, kmin, torch.full_like(kmin, -1))
            kmax = None
        else:
            f = (A[0] * lmax + A[1] * lmin) - g[lo:hi]
        F = torch.cat([self.gather(f), W[N:] - g[N:]])
        return F, kmin, kmax, F.abs().max()

    def ce_solve(self, g, U0=None, solver="newton", stats=None, **kwargs):
        """convex_envelope.solve(G, g, U0, fdmethod, solver) on the sharded cloud
        (convex_envelope.py:100-123).  Returns (U on all points [device tensor], diff, iterations,
        seconds), identical on every rank."""
        import time
        from . import solvers
        from ._context import to_device
        G, dev = self.G, self.device
        t0 = time.time()
        gs = to_device(g, dev)
        Us = gs.clone() if U0 is None else to_device(U0, dev).clone()
        dt = 1 / 2 * np.max([G.min_interior_nb_dist, G.min_radius]) ** 2      # convex_envelope.py:110
        sol, opt = kwargs.get("solution_tol", 1e-6), kwargs.get("operator_tol", 1e-6)

        def euler_res(U):
            F, _, _, fmax = self.residual(U, gs, 0, policy=False)
            return F, fmax

        if solver == "euler":
            Us, diff, iters = solvers.euler_state(self, Us, euler_res, dt, sol, opt,
                                                  kwargs.get("max_iters"), kwargs.get("timeout"))
            return Us, diff, iters, time.time() - t0
        solvers.drop_unsupported(kwargs)

        def res(U, jacobian):
            F, kmin, _, fmax = self.residual(U, gs, 0, policy=jacobian)
            return F, (_ShardedEllJacobian(self, kmin, None, cmin=-1.0) if jacobian else None), fmax

        def fallback(U, timeout):
            # a wall-clock budget would let the ranks take different numbers of steps: the
            # fallback runs a fixed number of Euler steps instead (solvers.py:166's cap / 100)
            U, diff, _ = solvers.euler_state(self, U, euler_res, dt, sol, opt,
                                             max(self.n_points // 10, 100), None)
            return U, diff

        Us, diff, iters = solvers.newton_state(self, Us, res, fallback, stats=stats, **kwargs)
        return Us, diff, iters, time.time() - t0

    def pucci_solve(self, f, g, A, U0=None, solver="newton", stats=None, **kwargs):
        """pucci.solve on the sharded cloud (pucci.py:105-160), scalar coefficients A."""
        import time
        from . import solvers
        from ._context import to_device
        G, dev, N, n = self.G, self.device, self.n_interior, self.n_points
        t0 = time.time()
        Fv = torch.zeros(n, dtype=torch.float64, device=dev)
        Fv[:N] = to_device(f, dev) if not np.isscalar(f) else float(f)
        if g is not None:
            Fv[N:] = to_device(g, dev) if not np.isscalar(g) else float(g)
        dt = 1 / 2 * np.max([G.min_radius, G.min_interior_nb_dist]) ** 2
        if U0 is None:
            Us = torch.ones(n, dtype=torch.float64, device=dev)
            if g is not None:
                Us[N:] = Fv[N:]
        else:
            Us = to_device(U0, dev).clone()
        sol, opt = kwargs.get("solution_tol", 1e-6), kwargs.get("operator_tol", 1e-6)

        def res(U, jacobian):
            R, kmin, kmax, fmax = self.residual(U, Fv, 1, A, policy=jacobian)
            Jac = (_ShardedEllJacobian(self, kmin, kmax, cmin=float(A[1]), cmax=float(A[0]))
                   if jacobian else None)
            return R, Jac, fmax

        def euler_res(U):
            R, _, _, fmax = self.residual(U, Fv, 1, A, policy=False)
            R[:N].neg_()                      # stable sign for explicit Euler
            return R, fmax

        if solver == "euler":
            Us, diff, iters = solvers.euler_state(self, Us, euler_res, dt, sol, opt,
                                                  kwargs.get("max_iters"), kwargs.get("timeout"))
            return Us, diff, iters, time.time() - t0
        solvers.drop_unsupported(kwargs)

        def fallback(U, timeout):
            U, diff, _ = solvers.euler_state(self, U, euler_res, dt, sol, opt,
                                             max(n // 10, 100), None)
            return U, diff

        Us, diff, iters = solvers.newton_state(self, Us, res, fallback, stats=stats, **kwargs)
        return Us, diff, iters, time.time() - t0
