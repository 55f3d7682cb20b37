// Convex envelope by directional convex-hull sweeps (extension; not in the reference).
//
// The discrete problem of convex_envelope.py:44-46, max(u - g, -lambda_min[u]) = 0 with
// lambda_min the smallest wide-stencil second difference (ddg.py:281-292), has as its solution
// the LARGEST u <= g whose second difference is >= 0 for every pair of every interior point.
// For an antipodal lattice pair (x - v, x + v) "second difference >= 0 at every point of the
// lattice line x + t v" says that the sequence along the line is convex, and the largest
// minorant with that property is the lower convex hull of the sequence (monotone chain, O(L)).
// One sweep = for every stencil direction, hull every lattice line (runs of points that carry
// this pair, with the neighbouring points as fixed ends), then lower the band points to the
// chords of their remaining pairs (the pairs with fractional wall points).  Every operation
// returns the largest minorant under a subset of the constraints, so the iterates decrease
// monotonically from g to the discrete solution -- the same fixed point the reference's Euler
// and Newton solvers approach -- and information travels along whole lines per sweep instead of
// one stencil width per iteration (sweep counts are mesh independent: 13 at 65^2 r=2, 35-37 at
// 129^2 / 257^2 r=4 for the two-cone obstacle).
//
// The line routine is compiled for the device (one thread per lattice line) and for the host
// (OpenMP over lines, device = -1: used by the CPU tests of this file only; the Python package
// calls the device path).  --fmad=false / -ffp-contract=off: identical arithmetic.
#include <cmath>
#include <cstdlib>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "common.cuh"

#define PDE_HD __host__ __device__ __forceinline__

namespace pde {

struct SweepGeom {
    int64_t nx, ny, pitch, wall_base;
    int r;
    const int64_t *ring;                 // wall ids of the integer ring: i=0 | i=nx+1 | j=0 | j=ny+1
    const unsigned long long *flags;     // per lattice offset: bit k = pair k is a lattice pair here
};

PDE_HD bool sw_inside(const SweepGeom &g, int64_t i, int64_t j) {
    return i >= 0 && i <= g.nx + 1 && j >= 0 && j <= g.ny + 1;
}

PDE_HD bool sw_interior(const SweepGeom &g, int64_t i, int64_t j) {
    return i >= 1 && i <= g.nx && j >= 1 && j <= g.ny;
}

// state-layout offset of the padded lattice point (i, j), 0 <= i <= nx+1, 0 <= j <= ny+1
PDE_HD int64_t sw_addr(const SweepGeom &g, int64_t i, int64_t j) {
    if (sw_interior(g, i, j)) return (i - 1) * g.pitch + (j - 1);
    int64_t id;
    if (i == 0) id = g.ring[j];
    else if (i == g.nx + 1) id = g.ring[(g.ny + 2) + j];
    else if (j == 0) id = g.ring[2 * (g.ny + 2) + i];
    else id = g.ring[2 * (g.ny + 2) + (g.nx + 2) + i];
    return g.wall_base + id;
}

// does the interior point (i, j) carry the lattice pair of direction k?
PDE_HD bool sw_constrained(const SweepGeom &g, int k, int64_t i, int64_t j) {
    if (!sw_interior(g, i, j)) return false;
    const bool band = i <= g.r || i > g.nx - g.r || j <= g.r || j > g.ny - g.r;
    if (!band) return true;
    return (g.flags[(i - 1) * g.pitch + (j - 1)] >> k) & 1ull;
}

// Lower convex hull of the points t = lo..hi of the line (i0 + t a, j0 + t b); the two ends
// stay, the points in between are lowered onto the hull.  The stack of hull vertices lives in
// `stk` at the padded positions of the line's own points (lines of one direction are disjoint).
PDE_HD double sw_hull_run(const SweepGeom &g, double *U, int32_t *stk, int64_t i0, int64_t j0,
                          int a, int b, int64_t lo, int64_t hi) {
    const int64_t W = g.ny + 2;
#define SW_STK(s) stk[(i0 + (lo + (s)) * a) * W + (j0 + (lo + (s)) * b)]
#define SW_U(t) U[sw_addr(g, i0 + (t) * a, j0 + (t) * b)]
    int64_t top = 0;
    double y1 = 0, y0 = 0;          // values of the top and second entry
    int64_t t1 = 0, t0 = 0;
    for (int64_t t = lo; t <= hi; ++t) {
        const double y = SW_U(t);
        while (top >= 2) {
            // middle vertex t1 on or above the chord t0 -> t: drop it
            if ((y1 - y0) * (double)(t - t0) >= (y - y0) * (double)(t1 - t0)) {
                --top;
                t1 = t0;
                y1 = y0;
                if (top >= 2) {
                    t0 = SW_STK(top - 2);
                    y0 = SW_U(t0);
                }
            } else {
                break;
            }
        }
        SW_STK(top) = (int32_t)t;
        ++top;
        t0 = t1; y0 = y1;
        t1 = t;  y1 = y;
    }
    double change = 0.0;
    for (int64_t s = 0; s + 1 < top; ++s) {
        const int64_t ta = SW_STK(s), tb = SW_STK(s + 1);
        if (tb == ta + 1) continue;
        const double ya = SW_U(ta), yb = SW_U(tb);
        const double slope_num = yb - ya, len = (double)(tb - ta);
        for (int64_t t = ta + 1; t < tb; ++t) {
            double v = ya + slope_num * ((double)(t - ta) / len);
            const int64_t ad = sw_addr(g, i0 + t * a, j0 + t * b);
            const double old = U[ad];
            if (v < old) {                       // never raise (rounding of the chord)
                if (old - v > change) change = old - v;
                U[ad] = v;
            }
        }
    }
#undef SW_STK
#undef SW_U
    return change;
}

// all runs of the line that starts at the padded point (i0, j0) (its predecessor is outside)
PDE_HD double sw_line(const SweepGeom &g, double *U, int32_t *stk, int64_t i0, int64_t j0, int a,
                      int b, int k) {
    int64_t L = 0;
    for (int64_t i = i0, j = j0; sw_inside(g, i, j); i += a, j += b) ++L;
    double change = 0.0;
    int64_t t = 1;
    while (t < L - 1) {
        if (sw_constrained(g, k, i0 + t * a, j0 + t * b)) {
            const int64_t first = t;
            while (t < L - 1 && sw_constrained(g, k, i0 + t * a, j0 + t * b)) ++t;
            const double c = sw_hull_run(g, U, stk, i0, j0, a, b, first - 1, t);
            if (c > change) change = c;
        } else {
            ++t;
        }
    }
    return change;
}

__global__ void __launch_bounds__(64)
k_sweep_lines(SweepGeom g, double *U, int32_t *stk, int a, int b, int k, double *change) {
    const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t W = g.ny + 2;
    if (idx >= (g.nx + 2) * W) return;
    const int64_t i = idx / W, j = idx % W;
    if (sw_inside(g, i - a, j - b)) return;          // not the first point of its line
    const double c = sw_line(g, U, stk, i, j, a, b, k);
    if (c > 0.0) atomic_max_nonneg(change, c);
}

// band pairs that are not lattice pairs: u_I <- min(u_I, theta u_J0 + (1 - theta) u_J1)
__global__ void k_sweep_relax(int64_t n, const int64_t *__restrict__ rows,
                              const double *__restrict__ theta, double *U, double *change) {
    const int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (q >= n) return;
    const double th = theta[q];
    const double v = __dadd_rn(__dmul_rn(th, U[rows[3 * q + 1]]),
                               __dmul_rn(__dsub_rn(1.0, th), U[rows[3 * q + 2]]));
    unsigned long long *p = reinterpret_cast<unsigned long long *>(U + rows[3 * q]);
    unsigned long long old = *p;
    const double first = __longlong_as_double((long long)old);
    while (v < __longlong_as_double((long long)old)) {
        const unsigned long long assumed = old;
        old = atomicCAS(p, assumed, (unsigned long long)__double_as_longlong(v));
        if (old == assumed) break;
    }
    if (v < first) atomic_max_nonneg(change, first - v);
}

}  // namespace pde

using namespace pde;

extern "C" {

int pde_ce_sweep(int device, int64_t nx, int64_t ny, int64_t pitch, int64_t wall_base,
                 int stencil_radius, int ndir, const int32_t *h_dirs, const int64_t *ring,
                 const uint64_t *flags, int64_t n_rest, const int64_t *rest_rows,
                 const double *rest_theta, double *U, int32_t *scratch, double tol,
                 int64_t max_sweeps, int relax_reps, double *h_out, void *stream) {
    PDE_REQUIRE(nx >= 1 && ny >= 1 && pitch >= ny && ndir >= 1 && ndir <= 64, "bad sweep geometry");
    PDE_REQUIRE(h_dirs && ring && flags && U && scratch && h_out, "null argument");
    PDE_REQUIRE(n_rest == 0 || (rest_rows && rest_theta), "null argument");
    SweepGeom g;
    g.nx = nx; g.ny = ny; g.pitch = pitch; g.wall_base = wall_base; g.r = stencil_radius;
    g.ring = ring;
    g.flags = reinterpret_cast<const unsigned long long *>(flags);
    const int64_t W = ny + 2, npad = (nx + 2) * W;
    double change = 0.0;
    int64_t sweeps = 0;
    if (device < 0) {                                   // host build of the same routine (tests)
        while (sweeps < max_sweeps) {
            change = 0.0;
            for (int k = 0; k < ndir; ++k) {
                const int a = h_dirs[2 * k], b = h_dirs[2 * k + 1];
                double cmax = 0.0;
#pragma omp parallel for schedule(dynamic, 64) reduction(max : cmax)
                for (int64_t idx = 0; idx < npad; ++idx) {
                    const int64_t i = idx / W, j = idx % W;
                    if (sw_inside(g, i - a, j - b)) continue;
                    const double c = sw_line(g, U, scratch, i, j, a, b, k);
                    if (c > cmax) cmax = c;
                }
                if (cmax > change) change = cmax;
            }
            for (int rep = 0; rep < relax_reps; ++rep)
                for (int64_t q = 0; q < n_rest; ++q) {
                    const double th = rest_theta[q];
                    const double v = th * U[rest_rows[3 * q + 1]] + (1.0 - th) * U[rest_rows[3 * q + 2]];
                    double &u = U[rest_rows[3 * q]];
                    if (v < u) {
                        if (u - v > change) change = u - v;
                        u = v;
                    }
                }
            ++sweeps;
            if (change < tol) break;
        }
        h_out[0] = change;
        h_out[1] = (double)sweeps;
        return 0;
    }
    DeviceGuard guard(device);
    cudaStream_t st = (cudaStream_t)stream;
    double *d_change = nullptr;
    PDE_CHECK(cudaMalloc(&d_change, sizeof(double)));
    int rc = 0;
    const unsigned nblk = (unsigned)((npad + 63) / 64);
    while (sweeps < max_sweeps && rc == 0) {
        cudaError_t e = cudaMemsetAsync(d_change, 0, sizeof(double), st);
        for (int k = 0; k < ndir && e == cudaSuccess; ++k) {
            k_sweep_lines<<<nblk, 64, 0, st>>>(g, U, scratch, h_dirs[2 * k], h_dirs[2 * k + 1], k,
                                               d_change);
            count_launch();
            e = cudaGetLastError();
        }
        for (int rep = 0; rep < relax_reps && n_rest > 0 && e == cudaSuccess; ++rep) {
            k_sweep_relax<<<(unsigned)((n_rest + 255) / 256), 256, 0, st>>>(n_rest, rest_rows,
                                                                            rest_theta, U, d_change);
            count_launch();
            e = cudaGetLastError();
        }
        if (e == cudaSuccess) e = cudaMemcpyAsync(&change, d_change, sizeof(double),
                                                  cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) {
            rc = fail_cuda(e, "convex-hull sweep", __FILE__, __LINE__);
            break;
        }
        ++sweeps;
        if (change < tol) break;
    }
    cudaFree(d_change);
    h_out[0] = change;
    h_out[1] = (double)sweeps;
    return rc;
}

}  // extern "C"
