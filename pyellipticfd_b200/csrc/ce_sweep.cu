// Convex envelope by directional convex-hull sweeps (extension; not in the reference).
//
// The discrete problem of convex_envelope.py:44-46, max(u - g, -lambda_min[u]) = 0 with
// lambda_min the smallest wide-stencil second difference (ddg.py:281-292), has as its solution
// the LARGEST u <= g whose second difference is >= 0 for every pair of every interior point.
// For an antipodal lattice pair (x - v, x + v) "second difference >= 0 at every point of the
// lattice line x + t v" says that the sequence along the line is convex, and the largest
// minorant with that property is the lower convex hull of the sequence.  One sweep = for every
// stencil direction, hull every lattice line (points that do not carry this pair are fixed
// vertices), and lower the band points to the chords of their remaining pairs (the pairs with
// fractional wall points).  Every operation returns the largest minorant under a subset of the
// constraints, so the iterates decrease monotonically from g to the discrete solution -- the same
// fixed point the reference's Euler and Newton solvers approach -- and information travels along
// whole lines per sweep instead of one stencil width per iteration.
//
// DEVICE (k_hull_pass): one CTA stages a bundle of WL adjacent lattice lines of one direction in
// shared memory (adjacent lines are adjacent in memory at every step: 32-byte sectors are used
// whole).  Lines that leave the lattice through a side wall are continued at the opposite side
// ("wrapped"), so that every line of a direction has the same number of points and every padded
// lattice point belongs to exactly one line; the pieces of a wrapped line are separated by fixed
// vertices.  The hull of a line is found by divide and conquer on a BITMAP of hull vertices:
//   level 0   one thread per 32-point word: monotone chain, the stack is the word's bit mask;
//   level k   one thread per pair of 32*2^(k-1)-point blocks finds the common lower tangent
//             (bridge) of the two block hulls -- a junction that is already convex costs two
//             cross products; otherwise a binary search for the left tangent point whose
//             predicate is a binary search for the right one, both by POSITION with O(1)
//             predecessor / successor queries on the two-level bitmap -- and clears the vertices
//             the bridge spans;
//   fill      every thread lowers the non-vertices of its word onto the chord of the
//             neighbouring vertices (never raises), records the largest decrease;
// and only changed values are written back.  Any chord between two points of the line lies on or
// above the line's hull, so a tangent misjudged at rounding level costs efficiency, never
// correctness.  The band relaxation is a two-kernel Jacobi step (deterministic).  A batch of
// sweeps is enqueued without host synchronisation; a one-thread kernel after every sweep raises a
// `done` flag when the sweep changed no value by `tol`, which turns the rest of the batch into
// no-ops.
//
// HOST (device = -1): the serial monotone chain per line under OpenMP, the same relaxation --
// used by the CPU tests of this file and as the CPU baseline of bench.py; the Python package
// only calls the device path.  --fmad=false / -ffp-contract=off.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
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
    double ytol;                         // a point less than ytol below a chord counts as on it
    const int64_t *ring;                 // wall ids of the integer ring: i=0 | i=nx+1 | j=0 | j=ny+1
    const unsigned long long *flags;     // per band point (sw_band_slot): bit k = pair k is a lattice pair
};

// compact index of a band point (an interior point within r cells of a wall): the 2r full rows
// next to the i-walls first, then 2r columns of every other row
PDE_HD int64_t sw_band_slot(const SweepGeom &g, int64_t i, int64_t j) {
    const int64_t r = g.r;
    if (g.nx <= 2 * r) return (i - 1) * g.ny + (j - 1);            // every row is a band row
    if (i <= r) return (i - 1) * g.ny + (j - 1);
    if (i > g.nx - r) return (i - (g.nx - r) - 1 + r) * g.ny + (j - 1);
    const int64_t base = 2 * r * g.ny + (i - r - 1) * 2 * r;
    return base + (j <= r ? j - 1 : j - (g.ny - r) - 1 + r);
}

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
    return (g.flags[sw_band_slot(g, i, j)] >> k) & 1ull;
}

// ---------------------------------------------------------------------------------------------
// host: serial monotone chain per line
// ---------------------------------------------------------------------------------------------

// Lower convex hull of the points t = lo..hi of the line (i0 + t a, j0 + t b); the two ends
// stay, the points in between are lowered onto the hull.  The stack of hull vertices lives in
// `stk` at the padded positions of the line's own points (lines of one direction are disjoint).
static double sw_hull_run(const SweepGeom &g, double *U, int32_t *stk, int64_t i0, int64_t j0,
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
            if (((y1 + g.ytol) - y0) * (double)(t - t0) >= (y - y0) * (double)(t1 - t0)) {
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
        const double slope = (yb - ya) / (double)(tb - ta);
        for (int64_t t = ta + 1; t < tb; ++t) {
            double v = ya + slope * (double)(t - ta);
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
static double sw_line(const SweepGeom &g, double *U, int32_t *stk, int64_t i0, int64_t j0, int a,
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

// ---------------------------------------------------------------------------------------------
// band relaxation (pairs that are not lattice pairs): u_I <- min(u_I, theta u_J0 + (1-theta) u_J1)
// Jacobi: phase A evaluates all band points from the old values, phase B commits.
// ---------------------------------------------------------------------------------------------
PDE_HD double sw_relax_value(int64_t b, const int64_t *ptr, const int64_t *center,
                             const int64_t *rj, const double *theta, const double *U) {
    double v = U[center[b]];
    for (int64_t q = ptr[b]; q < ptr[b + 1]; ++q) {
        const double th = theta[q];
#ifdef __CUDA_ARCH__
        const double c = __dadd_rn(__dmul_rn(th, U[rj[2 * q]]),
                                   __dmul_rn(__dsub_rn(1.0, th), U[rj[2 * q + 1]]));
#else
        const double c = th * U[rj[2 * q]] + (1.0 - th) * U[rj[2 * q + 1]];
#endif
        if (c < v) v = c;
    }
    return v;
}

__global__ void k_relax_eval(int64_t n, const int64_t *__restrict__ ptr,
                             const int64_t *__restrict__ center, const int64_t *__restrict__ rj,
                             const double *__restrict__ theta, const double *__restrict__ U,
                             double *__restrict__ tmp, const int *done) {
    if (*done) return;
    const int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (b < n) tmp[b] = sw_relax_value(b, ptr, center, rj, theta, U);
}

__global__ void k_relax_commit(int64_t n, const int64_t *__restrict__ center,
                               const double *__restrict__ tmp, double *U, double *change,
                               const int *done) {
    if (*done) return;
    const int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    double c = 0.0;
    if (b < n) {
        const double v = tmp[b], old = U[center[b]];
        if (v < old) {
            U[center[b]] = v;
            c = old - v;
        }
    }
    c = warp_max(c);
    if ((threadIdx.x & 31) == 0 && c > 0.0) atomic_max_nonneg(change, c);
}

// after a sweep: the next sweep's accumulator is already zero; raise `done` when converged
__global__ void k_sweep_end(const double *change, double tol, int *done, int *sweeps_done) {
    if (*done) return;
    *sweeps_done += 1;
    if (*change < tol) *done = 1;
}

// ---------------------------------------------------------------------------------------------
// bitmap hull of a bundle of lines in shared memory.  The phases are __host__ __device__ so that
// the CPU tests can replay a CTA thread by thread (device = -2) against the serial chain.
// ---------------------------------------------------------------------------------------------
struct HullPass {
    int a, b, k;          // direction (a >= 1, any b), or rows mode; flag bit k
    int rows_mode;        // 1: direction (0, 1): lines are the lattice rows i = 1..nx
    int Ns;               // generic: ny + 2 (wrap modulus); rows: number of lines nx
    int wl_log2;          // lines per CTA = 1 << wl_log2
    int LW;               // 32-point words per line
    int SW;               // summary words per line
    int groups;           // CTAs per start row i0
    int stop_after;       // profiling aid (PDE_SWEEP_STOP_AFTER): leave the kernel after phase n
    int64_t fm_offset;    // this direction's fixed-vertex masks in the mask table (uint32 words)
};

PDE_HD int hp_clz(uint32_t x) {
#ifdef __CUDA_ARCH__
    return __clz((int)x);
#else
    return x ? __builtin_clz(x) : 32;
#endif
}
PDE_HD int hp_ffs(uint32_t x) {
#ifdef __CUDA_ARCH__
    return __ffs((int)x);
#else
    return __builtin_ffs((int)x);
#endif
}
PDE_HD void hp_and(uint32_t *p, uint32_t v) {
#ifdef __CUDA_ARCH__
    atomicAnd(p, v);
#else
    *p &= v;
#endif
}
PDE_HD void hp_or(uint32_t *p, uint32_t v) {
#ifdef __CUDA_ARCH__
    atomicOr(p, v);
#else
    *p |= v;
#endif
}
PDE_HD double hp_mul(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
}
PDE_HD double hp_sub(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dsub_rn(a, b);
#else
    return a - b;
#endif
}
PDE_HD double hp_add(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}

// exact int -> double for 0 <= k < 2^31 without the conversion pipe
PDE_HD double hp_i2d(int k) {
#ifdef __CUDA_ARCH__
    return __dsub_rn(__hiloint2double(0x43300000, k), 4503599627370496.0);
#else
    return (double)k;
#endif
}

// largest set bit in [lo, pos] (-1 if none); reads no mask word below lo's
PDE_HD int bm_pred_in(const uint32_t *M, const uint32_t *S, int pos, int lo) {
    int w = pos >> 5;
    const int lw = lo >> 5;
    const uint32_t x = M[w] & (0xffffffffu >> (31 - (pos & 31)));
    int r;
    if (x) {
        r = (w << 5) + 31 - hp_clz(x);
    } else {
        int sw = w >> 5;
        uint32_t s = S[sw] & ((1u << (w & 31)) - 1u);
        while (!s) {
            if (--sw < (lw >> 5)) return -1;
            s = S[sw];
        }
        w = (sw << 5) + 31 - hp_clz(s);
        if (w < lw) return -1;
        r = (w << 5) + 31 - hp_clz(M[w]);
    }
    return r >= lo ? r : -1;
}

// smallest set bit in [pos, hi] (INT_MAX if none); reads no mask word above hi's
PDE_HD int bm_succ_in(const uint32_t *M, const uint32_t *S, int pos, int hi) {
    int w = pos >> 5;
    const int hw = hi >> 5;
    if (w > hw) return 0x7fffffff;
    const uint32_t x = M[w] & (0xffffffffu << (pos & 31));
    int r;
    if (x) {
        r = (w << 5) + hp_ffs(x) - 1;
    } else {
        int sw = w >> 5;
        uint32_t s = ((w & 31) == 31) ? 0u : (S[sw] & (0xffffffffu << ((w & 31) + 1)));
        while (!s) {
            if (++sw > (hw >> 5)) return 0x7fffffff;
            s = S[sw];
        }
        w = (sw << 5) + hp_ffs(s) - 1;
        if (w > hw) return 0x7fffffff;
        r = (w << 5) + hp_ffs(M[w]) - 1;
    }
    return r <= hi ? r : 0x7fffffff;
}

// largest set bit <= pos (-1 if none)
PDE_HD int bm_pred(const uint32_t *M, const uint32_t *S, int pos) {
    int w = pos >> 5;
    const uint32_t x = M[w] & (0xffffffffu >> (31 - (pos & 31)));
    if (x) return (w << 5) + 31 - hp_clz(x);
    int sw = w >> 5;
    uint32_t s = S[sw] & ((1u << (w & 31)) - 1u);
    while (!s) {
        if (--sw < 0) return -1;
        s = S[sw];
    }
    w = (sw << 5) + 31 - hp_clz(s);
    return (w << 5) + 31 - hp_clz(M[w]);
}

// smallest set bit >= pos (INT_MAX if none)
PDE_HD int bm_succ(const uint32_t *M, const uint32_t *S, int pos, int LW, int SW) {
    int w = pos >> 5;
    if (w >= LW) return 0x7fffffff;
    const uint32_t x = M[w] & (0xffffffffu << (pos & 31));
    if (x) return (w << 5) + hp_ffs(x) - 1;
    int sw = w >> 5;
    uint32_t s = ((w & 31) == 31) ? 0u : (S[sw] & (0xffffffffu << ((w & 31) + 1)));
    while (!s) {
        if (++sw >= SW) return 0x7fffffff;
        s = S[sw];
    }
    w = (sw << 5) + hp_ffs(s) - 1;
    return (w << 5) + hp_ffs(M[w]) - 1;
}

// middle point (t1, y1) on or above the chord (t0, y0) -> (t2, y2) -- or less than tol below
// it: rounding noise on planar pieces of the solution -- : not a hull vertex
PDE_HD bool hp_drop(int t0, double y0, int t1, double y1, int t2, double y2, double tol) {
    return hp_mul(hp_sub(hp_add(y1, tol), y0), hp_i2d(t2 - t0)) >=
           hp_mul(hp_sub(y2, y0), hp_i2d(t1 - t0));
}

// shared-memory index of point t of a line: one padding double per 32-point word, so that the
// 32 word owners of a warp (stride 33 doubles) hit distinct banks
PDE_HD int hp_ix(int t) { return t + (t >> 5); }

struct LineView {
    double *y;
    uint32_t *V, *SV, *F, *SF, *C, *GF;
    int LW, SW;
    double tol;
    PDE_HD double &Y(int t) const { return y[hp_ix(t)]; }
};

// tangent from the point v (left of the chain) to the convex chain of vertices in [lo, hi]:
// the vertex of smallest slope seen from v, the rightmost one on ties
PDE_HD int hp_tangent_right(const LineView &L, int v, int lo, int hi) {
    const double yv = L.Y(v);
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        const int c = bm_pred(L.V, L.SV, mid);
        const int c2 = bm_succ(L.V, L.SV, c + 1, L.LW, L.SW);
        // slope(v -> c2) <= slope(v -> c): c is on or above the chord v -> c2
        if (hp_drop(v, yv, c, L.Y(c), c2, L.Y(c2), L.tol)) lo = c2;
        else hi = c;
    }
    return lo;
}

PDE_HD void hp_clear(const LineView &L, int lo, int hi) {   // inclusive
    if (lo > hi) return;
    for (int w = lo >> 5; w <= (hi >> 5); ++w) {
        uint32_t mask = 0xffffffffu;
        if (w == (lo >> 5)) mask &= 0xffffffffu << (lo & 31);
        if (w == (hi >> 5)) mask &= 0xffffffffu >> (31 - (hi & 31));
        const uint32_t old = L.V[w], nw = old & ~mask;
        if (nw != old) {
            L.V[w] = nw;
            if (!nw) hp_and(&L.SV[w >> 5], ~(1u << (w & 31)));
        }
    }
}

// merge the hulls of the blocks [w-h, w) and [w, w+h) (in words) of one line: find the bridge
// (p, q) -- p the last vertex of the left hull, q the first of the right hull that survive --
// and clear the vertices in between.  First by walking from the junction (every step removes
// one vertex: a convex junction costs two cross products, a few removals a few more), and after
// HP_WALK removals by binary search on the rest: for the left tangent point, with the tangent
// search on the right hull as its predicate.
#define HP_WALK 12
PDE_HD void hp_merge(const LineView &L, int w, int h, int Lp) {
    const int J = w << 5;
    if (J >= Lp) return;
    int aLo = (w - h) << 5;
    int bHi = ((w + h) << 5 < Lp ? (w + h) << 5 : Lp) - 1;
    bool any_fixed = false;                          // summary words the block touches
    for (int sw = (w - h) >> 5; sw <= (bHi >> 10); ++sw) any_fixed = any_fixed || L.SF[sw] != 0u;
    if (any_fixed) {
        int f = bm_pred_in(L.F, L.SF, J - 1, aLo);
        if (f > aLo) aLo = f;
        f = bm_succ_in(L.F, L.SF, J, bHi);
        if (f < bHi) bHi = f;
    }
    int p = J - 1, q = J;
    if (aLo == p && bHi == q) return;
    double yp = L.Y(p), yq = L.Y(q);
    int budget = HP_WALK;
    bool moved = true;
    while (moved && budget > 0) {
        moved = false;
        while (p > aLo && budget > 0) {
            const int pp = bm_pred(L.V, L.SV, p - 1);
            const double ypp = L.Y(pp);
            if (!hp_drop(pp, ypp, p, yp, q, yq, L.tol)) break;
            p = pp; yp = ypp;
            moved = true;
            --budget;
        }
        while (q < bHi && budget > 0) {
            const int qq = bm_succ(L.V, L.SV, q + 1, L.LW, L.SW);
            const double yqq = L.Y(qq);
            if (!hp_drop(p, yp, q, yq, qq, yqq, L.tol)) break;
            q = qq; yq = yqq;
            moved = true;
            --budget;
        }
    }
    if (moved) {                                  // budget spent: search the remaining chains
        int lo = aLo, hi = p;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            const int v = bm_succ(L.V, L.SV, mid, L.LW, L.SW);
            const int pv = bm_pred(L.V, L.SV, v - 1);
            const int qv = hp_tangent_right(L, v, q, bHi);
            if (hp_drop(pv, L.Y(pv), v, L.Y(v), qv, L.Y(qv), L.tol)) hi = pv;
            else lo = v;
        }
        p = lo;
        q = hp_tangent_right(L, p, q, bHi);
    }
    hp_clear(L, p + 1, J - 1);
    hp_clear(L, J, q - 1);
}

PDE_HD int hp_mod(int x, int n) {
    x %= n;
    return x < 0 ? x + n : x;
}

// shared-memory stride of a line in doubles: 33 per word, and = 16 / WL mod 16 so that the WL
// lines a half-warp stages side by side fall into distinct banks
PDE_HD int hull_line_stride(int LW, int wl_log2) {
    const int base = (LW * 33 + 15) / 16 * 16;
    return base + (wl_log2 >= 4 ? 1 : 16 >> wl_log2);
}

// one CTA's view of its bundle
struct HullCta {
    SweepGeom g;
    HullPass P;
    double *U, *ys;
    const uint32_t *FM;   // fixed-vertex masks of this CTA's first line (LW words per line)
    uint32_t *Vs, *Fs, *Cs, *SVs, *SFs, *GFs, *Qs, *Qn;
    int WL, LW, SW, LPS, i0, s_base, nlines, Lp, NJ;
    unsigned deep_nx, deep_ny;       // extent of the deep interior (0 when there is none)
};

PDE_HD void hc_init(HullCta &c, const SweepGeom &g, const HullPass &P, double *U,
                    const uint32_t *fmask, unsigned char *smem, unsigned block) {
    c.g = g; c.P = P; c.U = U;
    c.WL = 1 << P.wl_log2; c.LW = P.LW; c.SW = P.SW;
    c.LPS = hull_line_stride(P.LW, P.wl_log2);
    c.ys = reinterpret_cast<double *>(smem);
    c.Vs = reinterpret_cast<uint32_t *>(c.ys + (size_t)c.WL * c.LPS);
    c.Fs = c.Vs + c.WL * c.LW;
    c.Cs = c.Fs + c.WL * c.LW;
    c.SVs = c.Cs + c.WL * c.LW;
    c.SFs = c.SVs + c.WL * c.SW;
    c.GFs = c.SFs + c.WL * c.SW;                    // per 32-word group: closed by hc_groups
    c.Qs = c.GFs + c.WL * c.SW;                     // queue of the words that need the chain
    c.Qn = c.Qs + c.WL * c.LW;
    c.i0 = P.rows_mode ? 0 : (int)(block / P.groups);
    c.s_base = (int)(block % P.groups) << P.wl_log2;
    c.nlines = c.WL < P.Ns - c.s_base ? c.WL : P.Ns - c.s_base;
    c.Lp = P.rows_mode ? (int)g.ny + 2 : (int)((g.nx + 1 - c.i0) / P.a) + 1;
    c.NJ = (int)g.ny + 2;
    c.deep_nx = g.nx > 2 * g.r ? (unsigned)(g.nx - 2 * g.r) : 0u;
    c.deep_ny = g.ny > 2 * g.r ? (unsigned)(g.ny - 2 * g.r) : 0u;
    // lines are numbered i0 * Ns + s0 (rows mode: s0), LW mask words each
    c.FM = fmask + P.fm_offset + ((int64_t)c.i0 * P.Ns + c.s_base) * P.LW;
}

PDE_HD LineView hc_line(const HullCta &c, int l) {
    LineView L;
    L.y = c.ys + l * c.LPS;
    L.V = c.Vs + l * c.LW;
    L.F = c.Fs + l * c.LW;
    L.C = c.Cs + l * c.LW;
    L.SV = c.SVs + l * c.SW;
    L.SF = c.SFs + l * c.SW;
    L.GF = c.GFs + l * c.SW;
    L.LW = c.LW;
    L.SW = c.SW;
    L.tol = c.g.ytol;
    return L;
}

// Thread <-> word: thread tid owns word (tid & (LWp - 1)...) -- precisely line l = tid / LWp,
// word w = tid % LWp with LWp = 32 SW threads per line, so that the 32 words of a summary word
// belong to one warp: level 0 and the first five merge levels only need __syncwarp().

// stage the bundle and clear the summaries
#define HP_BATCH 8
PDE_HD void hc_load(const HullCta &c, int tid, int nthr) {
    const SweepGeom &g = c.g;
    if (c.P.rows_mode) {
        // a warp takes 32 consecutive points of a line at a time (consecutive addresses)
        const int lane = tid & 31, nwarp = nthr >> 5;
        const int cpl = (c.Lp + 31) >> 5, total = c.nlines * cpl;
        for (int ch = tid >> 5; ch < total; ch += HP_BATCH * nwarp) {
            double v[HP_BATCH];                        // HP_BATCH loads in flight per thread
            int at[HP_BATCH];
#pragma unroll
            for (int u = 0; u < HP_BATCH; ++u) {
                const int cu = ch + u * nwarp;
                const int l = cu / cpl, t = ((cu - l * cpl) << 5) + lane;
                const bool ok = cu < total && t > 0 && t < c.Lp - 1;
                at[u] = ok ? l * c.LPS + hp_ix(t) : -1;
                v[u] = ok ? c.U[(int64_t)(c.s_base + l) * g.pitch + (t - 1)] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < HP_BATCH; ++u)
                if (at[u] >= 0) c.ys[at[u]] = v[u];
        }
        for (int idx = tid; idx < 2 * c.nlines; idx += nthr) {      // the two wall ends
            const int l = idx >> 1, t = (idx & 1) ? c.Lp - 1 : 0;
            c.ys[l * c.LPS + hp_ix(t)] = c.U[sw_addr(g, c.s_base + l + 1, t)];
        }
    } else {
        // nthr is a multiple of WL: a thread keeps its line and advances t by nthr / WL, so
        // that the WL lines of the bundle are read side by side (contiguous in memory).
        // Interior points first (plain strided loads the compiler can keep in flight together),
        // then the few points on the wall ring, whose addresses come from the ring table.
        const int l = tid & (c.WL - 1), dt = nthr >> c.P.wl_log2;
        const int dj = hp_mod(dt * c.P.b, c.NJ);
        const int t_first = tid >> c.P.wl_log2;
        const int di = dt * c.P.a;
        const int64_t doff = (int64_t)di * g.pitch;
        double *yl = c.ys + l * c.LPS;
        if (l < c.nlines) {
            int t = t_first;
            int i = c.i0 + t * c.P.a, j = hp_mod(c.s_base + l + t * c.P.b, c.NJ);
            int64_t off = (int64_t)(i - 1) * g.pitch + (j - 1);
            while (t < c.Lp) {
                double v[HP_BATCH];                    // HP_BATCH loads in flight per thread
                int at[HP_BATCH];
#pragma unroll
                for (int u = 0; u < HP_BATCH; ++u) {
                    const bool ring = i == 0 || i == (int)g.nx + 1 || j == 0 || j == c.NJ - 1;
                    const bool ok = t < c.Lp && !ring;
                    at[u] = ok ? hp_ix(t) : -1;
                    v[u] = ok ? c.U[off] : 0.0;
                    t += dt;
                    i += di;
                    j += dj;
                    off += doff + dj;
                    if (j >= c.NJ) {
                        j -= c.NJ;
                        off -= c.NJ;
                    }
                }
#pragma unroll
                for (int u = 0; u < HP_BATCH; ++u)
                    if (at[u] >= 0) yl[at[u]] = v[u];
            }
            t = t_first;
            i = c.i0 + t * c.P.a;
            j = hp_mod(c.s_base + l + t * c.P.b, c.NJ);
            for (; t < c.Lp; t += dt) {
                const bool ring = i == 0 || i == (int)g.nx + 1 || j == 0 || j == c.NJ - 1;
                if (ring) yl[hp_ix(t)] = c.U[sw_addr(g, i, j)];
                i += di;
                j += dj;
                if (j >= c.NJ) j -= c.NJ;
            }
        }
    }
    for (int idx = tid; idx < c.WL * c.SW; idx += nthr) {
        c.SVs[idx] = 0u;
        c.SFs[idx] = 0u;
        c.GFs[idx] = 0u;
    }
    if (tid == 0) *c.Qn = 0u;
}

// Fixed-vertex mask of word w of line l (bit q: point 32 w + q does not carry this direction's
// lattice pair -- a wall point, a line end, a wrap point, or a band point whose pair is not a
// lattice pair).  Depends on the grid and the direction only: computed once per grid by
// k_hull_forced into the mask table.
PDE_HD uint32_t hc_forced_word(const HullCta &c, int l, int w) {
    const SweepGeom &g = c.g;
    const HullPass &P = c.P;
    const int base = w << 5, NJ = c.NJ;
    if (l >= c.nlines || base >= c.Lp) return 0u;
    const int cnt = 32 < c.Lp - base ? 32 : c.Lp - base;
    int i, j;
    if (P.rows_mode) {
        i = c.s_base + l + 1;
        j = base;
    } else {
        i = c.i0 + base * P.a;
        j = hp_mod(c.s_base + l + base * P.b, NJ);
    }
    uint32_t fm = 0u;
    for (int q = 0; q < cnt; ++q) {
        // deep points (more than r cells from every wall) carry every lattice pair
        const bool deep = (unsigned)(i - g.r - 1) < c.deep_nx && (unsigned)(j - g.r - 1) < c.deep_ny;
        if (!deep) {
            bool forced;
            if (P.rows_mode) {
                forced = (j == 0 || j == NJ - 1);
            } else {
                const int t = base + q;
                forced = (i == 0 || i == (int)g.nx + 1 || j == 0 || j == NJ - 1) || t == 0 ||
                         t == c.Lp - 1 || j - P.b < 0 || j - P.b >= NJ || j + P.b < 0 ||
                         j + P.b >= NJ;
            }
            if (!forced) forced = !((g.flags[sw_band_slot(g, i, j)] >> P.k) & 1ull);
            if (forced) fm |= 1u << q;
        }
        if (P.rows_mode) {
            ++j;
        } else {
            i += P.a;
            j += P.b;
            if (j >= NJ) j -= NJ;
            else if (j < 0) j += NJ;
        }
    }
    return fm;
}

PDE_HD void hc_word_done(const LineView &L, int w, uint32_t m, uint32_t fm) {
    L.V[w] = m;
    L.F[w] = fm;
    L.C[w] = m == 0x80000001u ? 1u : 0u;              // flat word (hc_groups); then junction note
    if (m) hp_or(&L.SV[w >> 5], 1u << (w & 31));
    if (fm) hp_or(&L.SF[w >> 5], 1u << (w & 31));
}

// Level 0, part 1 -- every thread classifies its 32-point word.  Two usual cases after the first
// sweeps are settled at once, without the chain's data-dependent loop: a strictly convex word
// keeps every point (the chain would only compare consecutive triples); a word whose interior
// points all lie on or above the chord first -> last (planar pieces of the solution) keeps its
// two ends.  Every other word goes to the CTA's queue (part 2), so that the chain loops run in
// full warps instead of a few lanes of every warp.
PDE_HD void hc_level0(const HullCta &c, int tid) {
    const int LWp = c.SW << 5;
    const int l = tid / LWp, w = tid - l * LWp;
    if (l >= c.WL || w >= c.LW) return;
    const LineView L = hc_line(c, l);
    if (l >= c.nlines || (w << 5) >= c.Lp) {
        hc_word_done(L, w, 0u, 0u);
        return;
    }
    const uint32_t fm = c.FM[l * c.LW + w];
    if (fm == 0u && ((w + 1) << 5) <= c.Lp) {
        const double tol = c.g.ytol;
        const double *yw = L.y + 33 * w;              // hp_ix(32 w + q) = 33 w + q
        bool convex = true, flat = true;
        double ya = yw[0], yb = yw[1];
        const double yf = ya, dl = hp_sub(yw[31], ya);
#pragma unroll 6
        for (int k = 2; k < 32; ++k) {
            const double yc = yw[k];
            // hp_drop(k-2, ya, k-1, yb, k, yc): (yb + tol - ya) * 2 >= (yc - ya) * 1
            const double d = hp_sub(hp_add(yb, tol), ya);
            convex = convex && !(hp_add(d, d) >= hp_sub(yc, ya));
            // hp_drop(0, yf, k-1, yb, 31, yl)
            flat = flat && hp_mul(hp_sub(hp_add(yb, tol), yf), 31.0) >= hp_mul(dl, hp_i2d(k - 1));
            ya = yb;
            yb = yc;
        }
        if (convex || flat) {
            hc_word_done(L, w, convex ? 0xffffffffu : 0x80000001u, 0u);
            return;
        }
    }
#ifdef __CUDA_ARCH__
    const uint32_t slot = atomicAdd(c.Qn, 1u);
#else
    const uint32_t slot = (*c.Qn)++;
#endif
    c.Qs[slot] = (uint32_t)tid;
}

// Level 0, part 2 -- entry k of the queue: monotone chain within the word.  The stack is the
// bit mask m of the word, its two top entries are cached in (t1, y1), (t0, y0); a fixed vertex
// on top of the stack is never popped.
PDE_HD void hc_chain(const HullCta &c, int k) {
    if (k >= (int)*c.Qn) return;
    const int LWp = c.SW << 5;
    const int owner = (int)c.Qs[k];
    const int l = owner / LWp, w = owner - l * LWp;
    const LineView L = hc_line(c, l);
    const double tol = c.g.ytol;
    const double *yw = L.y + 33 * w;
    const int base = w << 5;
    const int cnt = 32 < c.Lp - base ? 32 : c.Lp - base;
    const uint32_t fm = c.FM[l * c.LW + w];
    uint32_t m = 0u;
    int t1 = -1, t0 = -1;
    double y1 = 0.0, y0 = 0.0;
    for (int q = 0; q < cnt; ++q) {
        const double yq = yw[q];
        while (t0 >= 0 && !((fm >> t1) & 1u)) {
            if (!hp_drop(t0, y0, t1, y1, q, yq, tol)) break;
            m &= ~(1u << t1);
            t1 = t0;
            y1 = y0;
            const uint32_t below = m & ((1u << t1) - 1u);
            if (below) {
                t0 = 31 - hp_clz(below);
                y0 = yw[t0];
            } else {
                t0 = -1;
            }
        }
        m |= 1u << q;
        t0 = t1; y0 = y1;
        t1 = q;  y1 = yq;
    }
    hc_word_done(L, w, m, fm);
}

// Planar pieces of the solution: a full 32-word group (1024 points) whose words are all flat
// and whose word ends all lie on or above the chord first -> last of the group keeps those two
// ends only -- the five warp-local merge levels are skipped.  One lane per group.
PDE_HD void hc_groups(const HullCta &c, int tid) {
    const int LWp = c.SW << 5;
    const int l = tid / LWp, w0 = tid - l * LWp;
    if (l >= c.nlines || (w0 & 31) != 0 || ((w0 + 32) << 5) > c.Lp) return;
    const LineView L = hc_line(c, l);
    if (L.SF[w0 >> 5]) return;
    const int t0 = w0 << 5, t1 = t0 + 1023;
    const double y0 = L.Y(t0), y1 = L.Y(t1);
    for (int k = 0; k < 32; ++k) {
        if (L.C[w0 + k] != 1u) return;
        const double *yw = L.y + 33 * (w0 + k);
        if (k > 0 && !hp_drop(t0, y0, t0 + 32 * k, yw[0], t1, y1, L.tol)) return;
        if (k < 31 && !hp_drop(t0, y0, t0 + 32 * k + 31, yw[31], t1, y1, L.tol)) return;
    }
    for (int k = 1; k < 31; ++k) L.V[w0 + k] = 0u;
    L.V[w0] = 1u;
    L.V[w0 + 31] = 0x80000000u;
    L.SV[w0 >> 5] = 0x80000001u;
    L.GF[w0 >> 5] = 1u;
}

// After level 0 every junction between two words of a group without fixed vertices is tested
// ONCE, by all lanes at the same time: if the last point of word w - 1 and the first point of
// word w are strictly below the chords to their neighbours in those two words, the junction is
// convex.  The note (neighbour positions + verdict) goes to C[w]; the merge of this junction,
// whatever its level, only has to see that the two neighbours are still the same.
#define HP_NOTE(pp, qq, ok) (0x80000000u | ((uint32_t)(ok) << 16) | ((uint32_t)(pp) << 8) | (uint32_t)(qq))
PDE_HD void hc_junctions(const HullCta &c, int tid) {
    const int LWp = c.SW << 5;
    const int l = tid / LWp, w = tid - l * LWp;
    if (l >= c.nlines || w >= c.LW) return;
    const LineView L = hc_line(c, l);
    L.C[w] = 0u;
    if ((w & 31) == 0 || ((w + 1) << 5) > c.Lp) return;
    if (L.SF[w >> 5] || L.GF[w >> 5]) return;
    const uint32_t va = L.V[w - 1] & 0x7fffffffu, vb = L.V[w] & ~1u;
    if (!va || !vb) return;
    const int pp = 31 - hp_clz(va), qq = hp_ffs(vb) - 1;
    const double *ya = L.y + 33 * (w - 1), *yb = L.y + 33 * w;
    const double yp = ya[31], yq = yb[0];
    const bool ok = !hp_drop(pp, ya[pp], 31, yp, 32, yq, L.tol) &&
                    !hp_drop(31, yp, 32, yq, 32 + qq, yb[qq], L.tol);
    L.C[w] = HP_NOTE(pp, qq, ok);
}

// Level h merges the blocks [w - h, w), [w, w + h) at the junction words w = (2k + 1) h.  For
// h <= 16 a junction stays inside one 32-word group = one warp: lane k of the group's warp takes
// junction k (contiguous active lanes).  For h >= 32 the first threads of the line take them.
PDE_HD void hc_merge(const HullCta &c, int tid, int h) {
    const int LWp = c.SW << 5;
    const int l = tid / LWp, x = tid - l * LWp;
    int w;
    if (h <= 16) w = (x & ~31) + (2 * (x & 31) + 1) * h;
    else w = (2 * x + 1) * h;
    if (l >= c.nlines || w >= c.LW || (h <= 16 && (2 * (x & 31) + 1) * h >= 32)) return;
    const LineView L = hc_line(c, l);
    if (h <= 16 && L.GF[w >> 5]) return;             // group closed by hc_groups
    const uint32_t note = L.C[w];
    if (note >> 16 == 0x8001u) {                     // noted convex: are the neighbours the same?
        const uint32_t va = L.V[w - 1], vb = L.V[w];
        if ((va >> 31) && (vb & 1u) && (va & 0x7fffffffu) && (vb & ~1u) &&
            31 - hp_clz(va & 0x7fffffffu) == (int)((note >> 8) & 0xff) &&
            hp_ffs(vb & ~1u) - 1 == (int)(note & 0xff))
            return;
    }
    hp_merge(L, w, h, c.Lp);
}

// lower the non-vertices of a word onto the chord of the neighbouring vertices
PDE_HD double hc_fill(const HullCta &c, int tid) {
    const int LWp = c.SW << 5;
    const int l = tid / LWp, w = tid - l * LWp;
    const int base = w << 5;
    if (l >= c.nlines || w >= c.LW || base >= c.Lp) return 0.0;
    const LineView L = hc_line(c, l);
    const int cnt = 32 < c.Lp - base ? 32 : c.Lp - base;
    const uint32_t valid = cnt == 32 ? 0xffffffffu : ((1u << cnt) - 1u);
    uint32_t nonv = ~L.V[w] & valid, changed = 0u;
    int ta = -1, tb = -1;
    double ya = 0.0, slope = 0.0, cmax = 0.0;
    while (nonv) {
        const int q = hp_ffs(nonv) - 1;
        nonv &= nonv - 1u;
        const int t = base + q;
        if (t > tb) {
            ta = bm_pred(L.V, L.SV, t);
            tb = bm_succ(L.V, L.SV, t, c.LW, c.SW);
            ya = L.Y(ta);
            slope = hp_sub(L.Y(tb), ya) / hp_i2d(tb - ta);
        }
        const double v = hp_add(ya, hp_mul(slope, hp_i2d(t - ta)));
        const double old = L.Y(t);
        if (v < old) {                               // never raise (rounding of the chord)
            L.Y(t) = v;
            changed |= 1u << q;
            if (old - v > cmax) cmax = old - v;
        }
    }
    L.C[w] = changed;
    return cmax;
}

// write the changed values back (same thread <-> point map as hc_load)
PDE_HD void hc_store(const HullCta &c, int tid, int nthr) {
    const SweepGeom &g = c.g;
    if (c.P.rows_mode) {
        const int lane = tid & 31, nwarp = nthr >> 5;
        const int cpl = (c.Lp + 31) >> 5, total = c.nlines * cpl;
        for (int ch = tid >> 5; ch < total; ch += nwarp) {
            const int l = ch / cpl, w = ch - l * cpl;
            if ((c.Cs[l * c.LW + w] >> lane) & 1u)
                c.U[(int64_t)(c.s_base + l) * g.pitch + ((w << 5) + lane - 1)] =
                    c.ys[l * c.LPS + hp_ix((w << 5) + lane)];
        }
    } else {
        const int l = tid & (c.WL - 1), dt = nthr >> c.P.wl_log2;
        const int dj = hp_mod(dt * c.P.b, c.NJ);
        int t = tid >> c.P.wl_log2;
        int j = hp_mod(c.s_base + l + t * c.P.b, c.NJ);
        const int64_t doff = (int64_t)dt * c.P.a * g.pitch;
        int64_t off = (int64_t)(c.i0 + t * c.P.a - 1) * g.pitch + (j - 1);
        const double *yl = c.ys + l * c.LPS;
        const uint32_t *Cl = c.Cs + l * c.LW;
        if (l < c.nlines)
            for (; t < c.Lp; t += dt) {
                if ((Cl[t >> 5] >> (t & 31)) & 1u) c.U[off] = yl[hp_ix(t)];
                j += dj;
                off += doff + dj;
                if (j >= c.NJ) {
                    j -= c.NJ;
                    off -= c.NJ;
                }
            }
    }
}

// one-off per grid: the fixed-vertex masks of every word of every line of one direction
__global__ void k_hull_forced(SweepGeom g, HullPass P, uint32_t *fmask) {
    HullCta c;
    hc_init(c, g, P, nullptr, fmask, nullptr, blockIdx.x);
    const int LWp = P.SW << 5;
    const int l = threadIdx.x / LWp, w = threadIdx.x - l * LWp;
    if (l < c.nlines && w < c.LW) const_cast<uint32_t *>(c.FM)[l * c.LW + w] = hc_forced_word(c, l, w);
}

__global__ void __launch_bounds__(1024, 1)
k_hull_pass(SweepGeom g, HullPass P, double *U, const uint32_t *fmask, double *change,
            const int *done) {
    if (*done) return;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    HullCta c;
    hc_init(c, g, P, U, fmask, smem_raw, blockIdx.x);
    const int tid = threadIdx.x, nthr = blockDim.x;
    hc_load(c, tid, nthr);
    if (P.stop_after == 1) return;
    __syncthreads();
    if (P.stop_after == 11) return;
    hc_level0(c, tid);
    if (P.stop_after == 12) return;
    __syncthreads();
    if (P.stop_after == 13) return;
    for (int k = tid; k < (int)*c.Qn; k += nthr) hc_chain(c, k);
    if (P.stop_after == 2) return;
    __syncthreads();
    hc_groups(c, tid);
    __syncwarp();
    hc_junctions(c, tid);
    if (P.stop_after == 3) return;
    for (int h = 1; h < c.LW && h <= 16; h <<= 1) {      // inside one warp's 32 words
        __syncwarp();
        hc_merge(c, tid, h);
    }
    if (P.stop_after == 4) return;
    for (int h = 32; h < c.LW; h <<= 1) {
        __syncthreads();
        hc_merge(c, tid, h);
    }
    if (P.stop_after == 5) return;
    __syncthreads();
    double cmax = hc_fill(c, tid);
    if (P.stop_after == 6) return;
    __syncthreads();
    hc_store(c, tid, nthr);
    cmax = warp_max(cmax);
    if ((tid & 31) == 0 && cmax > 0.0) atomic_max_nonneg(change, cmax);
}

static void hull_forced_replay(const SweepGeom &g, const HullPass &P, uint32_t *fmask, unsigned grid,
                               unsigned block) {
#pragma omp parallel for schedule(dynamic, 16)
    for (long blk = 0; blk < (long)grid; ++blk) {
        HullCta c;
        hc_init(c, g, P, nullptr, fmask, nullptr, (unsigned)blk);
        const int LWp = P.SW << 5;
        for (int t = 0; t < (int)block; ++t) {
            const int l = t / LWp, w = t - l * LWp;
            if (l < c.nlines && w < c.LW)
                const_cast<uint32_t *>(c.FM)[l * c.LW + w] = hc_forced_word(c, l, w);
        }
    }
}

// host replay of the kernel, one CTA after the other, phase by phase (tests: device = -2)
static double hull_pass_replay(const SweepGeom &g, const HullPass &P, double *U,
                               const uint32_t *fmask, unsigned grid, unsigned block, size_t smem) {
    double change = 0.0;
#pragma omp parallel reduction(max : change)
    {
        std::vector<unsigned char> buf(smem + 16);
#pragma omp for schedule(dynamic, 1)
        for (long blk = 0; blk < (long)grid; ++blk) {
            HullCta c;
            hc_init(c, g, P, U, fmask, buf.data(), (unsigned)blk);
            const int n = (int)block;
            for (int t = 0; t < n; ++t) hc_load(c, t, n);
            for (int t = 0; t < n; ++t) hc_level0(c, t);
            for (int k = 0; k < (int)*c.Qn; ++k) hc_chain(c, k);
            for (int t = 0; t < n; ++t) hc_groups(c, t);
            for (int t = 0; t < n; ++t) hc_junctions(c, t);
            for (int h = 1; h < c.LW; h <<= 1)
                for (int t = 0; t < n; ++t) hc_merge(c, t, h);
            for (int t = 0; t < n; ++t) {
                const double v = hc_fill(c, t);
                if (v > change) change = v;
            }
            for (int t = 0; t < n; ++t) hc_store(c, t, n);
        }
    }
    return change;
}

static size_t hull_smem_bytes(int wl_log2, int LW, int SW) {
    const int WL = 1 << wl_log2;
    return (size_t)WL * hull_line_stride(LW, wl_log2) * sizeof(double) +
           (size_t)(4 * WL * LW + 3 * WL * SW + 1) * sizeof(uint32_t);
}

struct HullPlan {
    HullPass P;
    unsigned grid, block;
    size_t smem;
};

// launch shape of every direction pass; false if a line does not fit one CTA
static bool hull_plan(std::vector<HullPlan> &plan, int64_t nx, int64_t ny, int ndir,
                      const int32_t *h_dirs, size_t budget, size_t max_smem,
                      int64_t *mask_words = nullptr) {
    plan.resize((size_t)ndir);
    int64_t words = 0;
    for (int k = 0; k < ndir; ++k) {
        int a = h_dirs[2 * k], b = h_dirs[2 * k + 1];
        HullPass P;
        P.k = k;
        P.stop_after = 0;
        if (const char *e = getenv("PDE_SWEEP_STOP_AFTER")) P.stop_after = atoi(e);
        int Lp, nstart;
        if (a == 0) {
            if (b != 1 && b != -1) return false;
            P.rows_mode = 1; P.a = 0; P.b = 1;
            P.Ns = (int)nx;
            Lp = (int)ny + 2;
            nstart = 1;
        } else {
            if (a < 0) { a = -a; b = -b; }
            P.rows_mode = 0; P.a = a; P.b = b;
            P.Ns = (int)ny + 2;
            Lp = (int)((nx + 1) / a) + 1;
            nstart = (int)std::min<int64_t>(a, nx + 2);
        }
        P.LW = (Lp + 31) / 32;
        P.SW = (P.LW + 31) / 32;
        int wl = 0;
        while (wl < 5 && hull_smem_bytes(wl + 1, P.LW, P.SW) <= budget &&
               (2 << wl) * P.SW * 32 <= 1024 && (2 << wl) <= P.Ns)
            ++wl;
        if (hull_smem_bytes(wl, P.LW, P.SW) > max_smem || (1 << wl) * P.SW * 32 > 1024)
            return false;
        P.wl_log2 = wl;
        P.groups = (P.Ns + (1 << wl) - 1) >> wl;
        P.fm_offset = words;                       // independent of wl: lines are i0 * Ns + s0
        words += (int64_t)nstart * P.Ns * P.LW;
        plan[k].P = P;
        plan[k].grid = (unsigned)(nstart * P.groups);
        plan[k].block = (unsigned)((1 << wl) * P.SW * 32);
        plan[k].smem = hull_smem_bytes(wl, P.LW, P.SW);
    }
    if (mask_words) *mask_words = words;
    return true;
}

// doubles of `work` in front of the mask table: relaxation scratch, per-sweep changes, flags
static int64_t sweep_work_head(int64_t n_band) { return n_band + 80; }

}  // namespace pde

using namespace pde;

extern "C" {

int64_t pde_ce_sweep_work_bytes(int64_t nx, int64_t ny, int ndir, const int32_t *h_dirs,
                                int64_t n_band) {
    std::vector<HullPlan> plan;
    int64_t words = 0;
    if (nx < 1 || ny < 1 || ndir < 1 || !h_dirs ||
        !hull_plan(plan, nx, ny, ndir, h_dirs, 110 * 1024, 227 * 1024, &words))
        return -1;
    return sweep_work_head(n_band) * (int64_t)sizeof(double) + words * (int64_t)sizeof(uint32_t);
}

static void sweep_geom(SweepGeom &g, int64_t nx, int64_t ny, int64_t pitch, int64_t wall_base,
                       int stencil_radius, const int64_t *ring, const uint64_t *flags,
                       double flat_tol) {
    g.nx = nx; g.ny = ny; g.pitch = pitch; g.wall_base = wall_base; g.r = stencil_radius;
    g.ytol = flat_tol > 0.0 ? flat_tol : 0.0;
    g.ring = ring;
    g.flags = reinterpret_cast<const unsigned long long *>(flags);
}

int pde_ce_sweep_setup(int device, int64_t nx, int64_t ny, int64_t pitch, int64_t wall_base,
                       int stencil_radius, int ndir, const int32_t *h_dirs, const int64_t *ring,
                       const uint64_t *flags, int64_t n_band, double *work, void *stream) {
    PDE_REQUIRE(nx >= 1 && ny >= 1 && pitch >= ny && ndir >= 1 && ndir <= 64, "bad sweep geometry");
    PDE_REQUIRE(h_dirs && ring && flags && work && device >= 0, "null argument");
    SweepGeom g;
    sweep_geom(g, nx, ny, pitch, wall_base, stencil_radius, ring, flags, 0.0);
    std::vector<HullPlan> plan;
    PDE_REQUIRE(hull_plan(plan, nx, ny, ndir, h_dirs, 110 * 1024, 227 * 1024),
                "lattice line too long for the sweep kernel");
    DeviceGuard guard(device);
    uint32_t *fmask = reinterpret_cast<uint32_t *>(work + sweep_work_head(n_band));
    for (int k = 0; k < ndir; ++k) {
        k_hull_forced<<<plan[k].grid, plan[k].block, 0, (cudaStream_t)stream>>>(g, plan[k].P, fmask);
        PDE_LAUNCH_CHECK();
    }
    return 0;
}

int pde_ce_sweep(int device, int64_t nx, int64_t ny, int64_t pitch, int64_t wall_base,
                 int stencil_radius, int ndir, const int32_t *h_dirs, const int64_t *ring,
                 const uint64_t *flags, int64_t n_band, const int64_t *band_ptr,
                 const int64_t *band_center, const int64_t *rest_j, const double *rest_theta,
                 double *U, double *work, double tol, double flat_tol, int64_t max_sweeps,
                 int relax_every, int relax_reps, double *h_out, void *stream) {
    PDE_REQUIRE(nx >= 1 && ny >= 1 && pitch >= ny && ndir >= 1 && ndir <= 64, "bad sweep geometry");
    PDE_REQUIRE(h_dirs && ring && flags && U && h_out, "null argument");
    PDE_REQUIRE(n_band == 0 || (band_ptr && band_center && rest_j && rest_theta), "null argument");
    PDE_REQUIRE(max_sweeps >= 0 && ny + 2 < (1 << 24) && nx + 2 < (1 << 24), "bad sweep arguments");
    SweepGeom g;
    sweep_geom(g, nx, ny, pitch, wall_base, stencil_radius, ring, flags, flat_tol);
    const int64_t W = ny + 2, npad = (nx + 2) * W;
    if (relax_every <= 0 || relax_every > ndir) relax_every = ndir;
    size_t budget = 110 * 1024;                       // two bundles per SM
    if (const char *e = getenv("PDE_SWEEP_SMEM_KB")) budget = (size_t)atoi(e) * 1024;
    std::vector<HullPlan> plan;
    double change = 0.0;
    int64_t sweeps = 0;
    if (device < 0) {                // host builds: -1 serial chain per line, -2 kernel replay
        std::vector<uint32_t> fmask;
        if (device == -2) {
            int64_t words = 0;
            PDE_REQUIRE(hull_plan(plan, nx, ny, ndir, h_dirs, budget, 227 * 1024, &words),
                        "lattice line too long for the sweep kernel");
            fmask.resize((size_t)words);
            for (int k = 0; k < ndir; ++k)
                hull_forced_replay(g, plan[k].P, fmask.data(), plan[k].grid, plan[k].block);
        }
        std::vector<int32_t> stk((size_t)(device == -1 ? npad : 1));
        std::vector<double> tmp((size_t)std::max<int64_t>(n_band, 1));
        auto relax = [&]() {
            double cmax = 0.0;
            for (int rep = 0; rep < relax_reps && n_band > 0; ++rep) {
#pragma omp parallel for schedule(static)
                for (int64_t b = 0; b < n_band; ++b)
                    tmp[b] = sw_relax_value(b, band_ptr, band_center, rest_j, rest_theta, U);
#pragma omp parallel for schedule(static) reduction(max : cmax)
                for (int64_t b = 0; b < n_band; ++b) {
                    double &u = U[band_center[b]];
                    if (tmp[b] < u) {
                        if (u - tmp[b] > cmax) cmax = u - tmp[b];
                        u = tmp[b];
                    }
                }
            }
            return cmax;
        };
        const bool trace = getenv("PDE_SWEEP_TRACE") != nullptr;     // per-direction changes (host builds)
        while (sweeps < max_sweeps) {
            change = 0.0;
            for (int k = 0; k < ndir; ++k) {
                const int a = h_dirs[2 * k], b = h_dirs[2 * k + 1];
                double cmax = 0.0;
                if (device == -2) {
                    cmax = hull_pass_replay(g, plan[k].P, U, fmask.data(), plan[k].grid,
                                            plan[k].block, plan[k].smem);
                } else {
#pragma omp parallel for schedule(dynamic, 64) reduction(max : cmax)
                    for (int64_t idx = 0; idx < npad; ++idx) {
                        const int64_t i = idx / W, j = idx % W;
                        if (sw_inside(g, i - a, j - b)) continue;
                        const double c = sw_line(g, U, stk.data(), i, j, a, b, k);
                        if (c > cmax) cmax = c;
                    }
                }
                if (cmax > change) change = cmax;
                if (trace) fprintf(stderr, "sweep %lld dir %d (%d,%d) change %.3e\n", (long long)sweeps, k, a, b, cmax);
                if ((k + 1) % relax_every == 0 || k == ndir - 1) {
                    double rc_ = relax();
                    if (trace) fprintf(stderr, "sweep %lld relax after dir %d change %.3e\n", (long long)sweeps, k, rc_);
                    change = std::max(change, rc_);
                }
            }
            ++sweeps;
            if (change < tol) break;
        }
        h_out[0] = change;
        h_out[1] = (double)sweeps;
        return 0;
    }

    PDE_REQUIRE(work, "null argument");
    DeviceGuard guard(device);
    cudaStream_t st = (cudaStream_t)stream;
    int max_smem = 0;
    PDE_CHECK(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    PDE_CHECK(cudaFuncSetAttribute(k_hull_pass, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   max_smem));
    PDE_REQUIRE(hull_plan(plan, nx, ny, ndir, h_dirs, std::min(budget, (size_t)max_smem),
                          (size_t)max_smem),
                "lattice line too long for the sweep kernel");

    // work: [n_band tmp | change per sweep (batch_cap + 1) | done, sweeps_done (two ints) | ..
    //        sweep_work_head(n_band) doubles in all | fixed-vertex masks (pde_ce_sweep_setup)]
    constexpr int64_t batch_cap = 64;
    double *d_tmp = work, *d_change = work + n_band;
    int *d_flags = reinterpret_cast<int *>(d_change + batch_cap + 1);
    const uint32_t *fmask = reinterpret_cast<const uint32_t *>(work + sweep_work_head(n_band));
    int rc = 0;
    const unsigned rblk = (unsigned)((n_band + 127) / 128);
    while (sweeps < max_sweeps && rc == 0) {
        const int64_t batch = std::min<int64_t>(batch_cap, max_sweeps - sweeps);
        cudaError_t e = cudaMemsetAsync(d_change, 0, (batch_cap + 2) * sizeof(double), st);
        for (int64_t s = 0; s < batch && e == cudaSuccess; ++s) {
            double *ch = d_change + s;
            for (int k = 0; k < ndir && e == cudaSuccess; ++k) {
                const HullPlan &pl = plan[k];
                k_hull_pass<<<pl.grid, pl.block, pl.smem, st>>>(g, pl.P, U, fmask, ch, d_flags);
                count_launch();
                if (n_band > 0 && ((k + 1) % relax_every == 0 || k == ndir - 1)) {
                    for (int rep = 0; rep < relax_reps; ++rep) {
                        k_relax_eval<<<rblk, 128, 0, st>>>(n_band, band_ptr, band_center, rest_j,
                                                           rest_theta, U, d_tmp, d_flags);
                        k_relax_commit<<<rblk, 128, 0, st>>>(n_band, band_center, d_tmp, U, ch,
                                                             d_flags);
                        count_launch(2);
                    }
                }
                e = cudaGetLastError();
            }
            if (e == cudaSuccess) {
                k_sweep_end<<<1, 1, 0, st>>>(ch, tol, d_flags, d_flags + 1);
                count_launch();
                e = cudaGetLastError();
            }
        }
        double h_change[batch_cap + 2];
        if (e == cudaSuccess)
            e = cudaMemcpyAsync(h_change, d_change, (batch_cap + 2) * sizeof(double),
                                cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) {
            rc = fail_cuda(e, "convex-hull sweep", __FILE__, __LINE__);
            break;
        }
        int h_flags[2];
        memcpy(h_flags, &h_change[batch_cap + 1], sizeof(h_flags));
        const int done_sweeps = h_flags[1];
        if (done_sweeps > 0) change = h_change[done_sweeps - 1];
        sweeps += done_sweeps;
        if (h_flags[0] || done_sweeps < batch) break;
    }
    h_out[0] = change;
    h_out[1] = (double)sweeps;
    return rc;
}

}  // extern "C"
