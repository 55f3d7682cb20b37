// Point-cloud operators: linear-interpolation (ddi) and Froese (ddf) finite differences.
//
// The reference caches one SciPy CSR matrix (5 nnz/row) per search direction
// (ddi.py:355-356, ddf.py:193-194) and evaluates `Nds` SpMVs + a row argsort per call
// (ddi.py:358-367).  Here the Nds matrices are ONE direction-major ELL table
//     idx[k][i][0..3]  int32   neighbour point of interior point i in direction k
//     w  [k][i][0..3]  double  its weight (centre weight = -sum, not stored)
// so that thread i streams 48 contiguous bytes per direction (coalesced over i) and
// gathers u through L2.  Algorithmic traffic 48 B per (point, direction): HBM-bound.
#include "common.cuh"

struct pde_ell {
    int device;
    int64_t n_interior, n_points;
    int nds;
    const int32_t *idx;   // (nds, n_interior, 4), caller-owned device memory
    const double *w;      // (nds, n_interior, 4)
    int64_t row_offset;   // centre point of table row i is point i + row_offset (point-range shards)
    int num_sms;
};

namespace pde {

// d2u[i,k] = sum_j w_j (u[idx_j] - u_i): the first-call formula of ddi.py:283-286 with the
// 2/(hf+hb) factor folded into w (the cached path M.dot(u) of ddi.py:361 differs from it by
// rounding only).  Ties: lowest k for the minimum, highest k for the maximum (what a stable
// ascending sort would give for arg[:,0] / arg[:,-1]; np.argsort's order is unspecified).
template <int UNROLL>
__global__ void __launch_bounds__(256)
k_ell_d2eigs(int64_t n, int64_t off, int nds, const int4 *__restrict__ idx,
             const double2 *__restrict__ w, const double *__restrict__ u, double *__restrict__ lmin,
             double *__restrict__ lmax, int32_t *__restrict__ kmin, int32_t *__restrict__ kmax) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double ui = u[i + off];
    double mn = __longlong_as_double(0x7ff0000000000000LL);
    double mx = __longlong_as_double(0xfff0000000000000LL);
    int an = 0, ax = 0;
    int k = 0;
    for (; k + UNROLL <= nds; k += UNROLL) {
        int4 id[UNROLL];
        double2 wa[UNROLL], wb[UNROLL];
#pragma unroll
        for (int q = 0; q < UNROLL; ++q) {
            const int64_t e = (int64_t)(k + q) * n + i;
            id[q] = __ldcs(&idx[e]);                                 // streamed once
            wa[q] = __ldcs(&w[2 * e]);
            wb[q] = __ldcs(&w[2 * e + 1]);
        }
#pragma unroll
        for (int q = 0; q < UNROLL; ++q) {
            double d = wa[q].x * (u[id[q].x] - ui) + wa[q].y * (u[id[q].y] - ui) +
                       wb[q].x * (u[id[q].z] - ui) + wb[q].y * (u[id[q].w] - ui);
            if (d < mn) { mn = d; an = k + q; }
            if (d >= mx) { mx = d; ax = k + q; }
        }
    }
    for (; k < nds; ++k) {
        const int64_t e = (int64_t)k * n + i;
        int4 id = idx[e];
        double2 wa = w[2 * e], wb = w[2 * e + 1];
        double d = wa.x * (u[id.x] - ui) + wa.y * (u[id.y] - ui) + wb.x * (u[id.z] - ui) +
                   wb.y * (u[id.w] - ui);
        if (d < mn) { mn = d; an = k; }
        if (d >= mx) { mx = d; ax = k; }
    }
    if (lmin) lmin[i] = mn;
    if (lmax) lmax[i] = mx;
    if (kmin) kmin[i] = an;
    if (kmax) kmax[i] = ax;
}

__global__ void k_ell_apply(int64_t n, int64_t off, const int4 *__restrict__ idx,
                            const double2 *__restrict__ w, const int32_t *__restrict__ ksel,
                            const double *__restrict__ x, double *__restrict__ y) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    int k = ksel[i];
    const int64_t e = (int64_t)k * n + i;
    int4 id = idx[e];
    double2 wa = w[2 * e], wb = w[2 * e + 1];
    double xi = x[i + off];
    y[i] = wa.x * (x[id.x] - xi) + wa.y * (x[id.y] - xi) + wb.x * (x[id.z] - xi) +
           wb.y * (x[id.w] - xi);
}

// 2x2 solve X xi = v, the operations of LAPACK gesv behind np.linalg.solve (ddi.py:265, ddf.py:60)
// in the order and with the roundings of the NumPy/OpenBLAS build of this image -- pinned bit for
// bit against np.linalg.solve on the fixture clouds (oracle/make_golden.py::ellpin_case): partial
// pivoting on |a| vs |c| (first maximum wins), multiplier = c * (1/a), Schur complement with a
// separate multiply and subtract, forward and back substitution with fused multiply-adds.
__device__ __forceinline__ void solve2(double a, double b, double c, double d, double v0, double v1,
                                       double &x0, double &x1) {
    // rows: [a b | v0], [c d | v1]
    if (fabs(c) > fabs(a)) {
        double t;
        t = a; a = c; c = t;
        t = b; b = d; d = t;
        t = v0; v0 = v1; v1 = t;
    }
    const double l = __dmul_rn(c, __ddiv_rn(1.0, a));
    const double u22 = __dsub_rn(d, __dmul_rn(l, b));
    const double y1 = __fma_rn(-l, v0, v1);
    x1 = __ddiv_rn(y1, u22);
    x0 = __ddiv_rn(__fma_rn(-b, x1, v0), a);
}

// one thread per (interior point, direction)
__global__ void __launch_bounds__(128)
k_ell_build(int64_t n, int nds, const double *__restrict__ P, const int64_t *__restrict__ simp,
            const int64_t *__restrict__ sptr, const double *__restrict__ V, int v_per_point,
            double eps_h, int scheme, int64_t row_offset, int64_t n_interior_pts,
            int32_t *__restrict__ idx, double *__restrict__ w, int32_t *__restrict__ fail) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int k = blockIdx.y;
    if (i >= n) return;
    const bool froese = scheme == 1;
    const bool first = scheme >= 2;          // first derivative: forward simplex only
    const double vx = v_per_point ? V[2 * i] : V[2 * k];
    const double vy = v_per_point ? V[2 * i + 1] : V[2 * k + 1];
    const int64_t ctr = i + row_offset;      // centre point of this row
    const double px = P[2 * ctr], py = P[2 * ctr + 1];
    bool have_f = false, have_b = false;
    int64_t f1 = 0, f2 = 0, b1 = 0, b2 = 0;
    double xf1 = 0, xf2 = 0, xb1 = 0, xb2 = 0;
    const double lo = froese ? 0.0 : -eps_h, hi = froese ? 0.0 : eps_h;
    if (first) have_b = true;
    for (int64_t s = sptr[i]; s < sptr[i + 1] && !(have_f && have_b); ++s) {
        int64_t s1 = simp[3 * s + 1], s2 = simp[3 * s + 2];
        // boundary domain of ddi.d1 (ddi.py:66-68): only simplices made of interior points
        if (scheme == 3 && (s1 >= n_interior_pts || s2 >= n_interior_pts)) continue;
        double a = P[2 * s1] - px, c = P[2 * s1 + 1] - py;       // first simplex vector
        double b = P[2 * s2] - px, d = P[2 * s2 + 1] - py;       // second
        double x0, x1;
        solve2(a, b, c, d, vx, vy, x0, x1);
        if (!have_f && x0 >= lo && x1 >= lo) {                   // ddi.py:267 / ddf.py:62
            have_f = true; f1 = s1; f2 = s2; xf1 = x0; xf2 = x1;
        }
        if (!have_b && x0 <= hi && x1 <= hi) {                   // ddi.py:268 / ddf.py:63
            have_b = true; b1 = s1; b2 = s2; xb1 = -x0; xb2 = -x1;
        }
    }
    int64_t o = ((int64_t)k * n + i) * 4;
    if (!(have_f && have_b)) {
        atomicAdd(fail, 1);
        idx[o] = idx[o + 1] = idx[o + 2] = idx[o + 3] = (int32_t)ctr;
        w[o] = w[o + 1] = w[o + 2] = w[o + 3] = 0.0;
        return;
    }
    if (first) {
        // ddi.py:72-80: d1u = xi_1 (u[S1]-u_i) + xi_2 (u[S2]-u_i); the row is padded to the
        // 4-wide table format with two zero-weight references to the centre point
        idx[o] = (int32_t)f1; idx[o + 1] = (int32_t)f2; idx[o + 2] = idx[o + 3] = (int32_t)ctr;
        w[o] = xf1; w[o + 1] = xf2; w[o + 2] = w[o + 3] = 0.0;
        return;
    }
    if (!froese) {
        // ddi.py:279-297: hf = 1/sum(xi_f), hb = 1/sum(xi_b); rows scaled by 2/(hf+hb)
        double hf = 1.0 / (xf1 + xf2), hb = 1.0 / (xb1 + xb2);
        double sc = 2.0 / (hf + hb);
        idx[o] = (int32_t)f1; idx[o + 1] = (int32_t)f2; idx[o + 2] = (int32_t)b1; idx[o + 3] = (int32_t)b2;
        w[o] = sc * xf1; w[o + 1] = sc * xf2; w[o + 2] = sc * xb1; w[o + 3] = sc * xb2;
        return;
    }
    // Froese's generalised finite difference, ddf.py:74-116
    const double TWO_PI = 6.283185307179586;
    int64_t S[4] = {f1, f2, b1, b2};
    double h[4], dth[4];
    double vth = atan2(vy, vx);
    for (int j = 0; j < 4; ++j) {
        double X = P[2 * S[j]] - px, Y = P[2 * S[j] + 1] - py;
        h[j] = sqrt(X * X + Y * Y);
        double t = atan2(Y, X) - vth;
        t = fmod(t, TWO_PI);
        if (t < 0) t += TWO_PI;                                   // numpy's % (sign of divisor)
        dth[j] = t;
    }
    // stable sort of 4 angles (np.argsort, ddf.py:92)
    int ord[4] = {0, 1, 2, 3};
    for (int a = 1; a < 4; ++a)
        for (int b = a; b > 0 && dth[ord[b]] < dth[ord[b - 1]]; --b) {
            int t = ord[b]; ord[b] = ord[b - 1]; ord[b - 1] = t;
        }
    double c[4], s[4];
    int64_t Ss[4];
    for (int j = 0; j < 4; ++j) {
        c[j] = h[ord[j]] * cos(dth[ord[j]]);
        s[j] = h[ord[j]] * sin(dth[ord[j]]);
        Ss[j] = S[ord[j]];
    }
    if (c[0] >= 0 && c[1] >= 0 && s[0] >= 0 && s[1] >= 0) {       // ddf.py:104-108 roll by -1
        double c0 = c[0], s0 = s[0];
        int64_t S0 = Ss[0];
        for (int j = 0; j < 3; ++j) { c[j] = c[j + 1]; s[j] = s[j + 1]; Ss[j] = Ss[j + 1]; }
        c[3] = c0; s[3] = s0; Ss[3] = S0;
    }
    double A = c[2] * s[1] - c[1] * s[2], B = c[0] * s[3] - c[3] * s[0];
    double denom = A * (c[0] * c[0] * s[3] - c[3] * c[3] * s[0]) -
                   B * (c[2] * c[2] * s[1] - c[1] * c[1] * s[2]);
    w[o] = 2 * s[3] * A / denom;
    w[o + 1] = 2 * s[2] * B / denom;
    w[o + 2] = -2 * s[1] * B / denom;
    w[o + 3] = -2 * s[0] * A / denom;
    for (int j = 0; j < 4; ++j) idx[o + j] = (int32_t)Ss[j];
}

}  // namespace pde

using namespace pde;

extern "C" {

int pde_ell_create(pde_ell **out, int device, int64_t n_interior, int64_t n_points, int nds,
                   const int32_t *d_idx, const double *d_w) {
    PDE_REQUIRE(out && d_idx && d_w && nds > 0 && n_interior > 0, "bad argument");
    PDE_REQUIRE(n_points < ((int64_t)1 << 31), "point ids must fit int32");
    PDE_REQUIRE((reinterpret_cast<uintptr_t>(d_idx) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(d_w) & 31) == 0, "ELL tables must be 32-byte aligned");
    cudaDeviceProp prop;
    PDE_CHECK(cudaGetDeviceProperties(&prop, device));
    PDE_REQUIRE(prop.major == 10, "pde_b200 requires an sm_100-class (Blackwell) device");
    pde::keep_async_pool(device);
    pde_ell *e = new pde_ell();
    e->device = device;
    e->n_interior = n_interior;
    e->n_points = n_points;
    e->nds = nds;
    e->idx = d_idx;
    e->w = d_w;
    e->row_offset = 0;
    e->num_sms = prop.multiProcessorCount;
    *out = e;
    return 0;
}

int pde_ell_set_row_offset(pde_ell *e, int64_t row_offset) {
    PDE_REQUIRE(e && row_offset >= 0 && row_offset + e->n_interior <= e->n_points, "bad row offset");
    e->row_offset = row_offset;
    return 0;
}

int pde_ell_destroy(pde_ell *e) {
    delete e;
    return 0;
}

int pde_ell_d2eigs(pde_ell *e, const double *d_u, double *d_lmin, double *d_lmax, int32_t *d_kmin,
                   int32_t *d_kmax, void *stream) {
    PDE_REQUIRE(e && d_u, "null argument");
    DeviceGuard guard(e->device);
    int64_t n = e->n_interior;
    unsigned grid = (unsigned)((n + 255) / 256);
    k_ell_d2eigs<4><<<grid, 256, 0, (cudaStream_t)stream>>>(
        n, e->row_offset, e->nds, reinterpret_cast<const int4 *>(e->idx), reinterpret_cast<const double2 *>(e->w),
        d_u, d_lmin, d_lmax, d_kmin, d_kmax);
    PDE_LAUNCH_CHECK();
    return 0;
}

int pde_ell_apply_policy(pde_ell *e, const int32_t *d_k, const double *d_x, double *d_y,
                         void *stream) {
    PDE_REQUIRE(e && d_k && d_x && d_y, "null argument");
    DeviceGuard guard(e->device);
    int64_t n = e->n_interior;
    k_ell_apply<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        n, e->row_offset, reinterpret_cast<const int4 *>(e->idx), reinterpret_cast<const double2 *>(e->w), d_k,
        d_x, d_y);
    PDE_LAUNCH_CHECK();
    return 0;
}

int pde_ell_build(int device, int64_t n_interior, const double *d_points,
                  const int64_t *d_simplices, const int64_t *d_simp_ptr, int nds,
                  const double *d_V, int v_per_point, double eps_over_h, int scheme,
                  int64_t row_offset, int64_t n_interior_pts,
                  int32_t *d_idx, double *d_w, int32_t *d_fail, void *stream) {
    PDE_REQUIRE(d_points && d_simplices && d_simp_ptr && d_V && d_idx && d_w && d_fail,
                "null argument");
    PDE_REQUIRE(nds > 0 && nds < 65536, "bad direction count");
    PDE_REQUIRE(!v_per_point || nds == 1, "per-point directions need nds == 1");
    PDE_REQUIRE(scheme >= 0 && scheme <= 3, "unknown scheme");
    DeviceGuard guard(device);
    dim3 grid((unsigned)((n_interior + 127) / 128), (unsigned)nds);
    k_ell_build<<<grid, 128, 0, (cudaStream_t)stream>>>(n_interior, nds, d_points, d_simplices,
                                                        d_simp_ptr, d_V, v_per_point, eps_over_h, scheme,
                                                        row_offset, n_interior_pts, d_idx, d_w, d_fail);
    PDE_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"

// ===========================================================================================
// Problems and the Newton linear solve on point clouds: convex_envelope / pucci with
// fdmethod='interpolate' | 'froese' (convex_envelope.py:37-54, pucci.py:40-58,127-141) and
// solvers.py:179-181 with the ELL Jacobian.  Vectors are the reference's flat vectors
// (interior points first, then boundary points): no padding, no re-ordering.
// ===========================================================================================
#include "krylov.cuh"

namespace pde {

struct EllJacDev {
    const int32_t *kmin, *kmax;     // selected direction per interior point; kmin < 0 = identity row
    double cmin, cmax;
};

__device__ __forceinline__ double ell_row(const int4 *idx, const double2 *w, int64_t n, int k,
                                          int64_t i, const double *x, double xi) {
    const int64_t e = (int64_t)k * n + i;
    int4 id = idx[e];
    double2 wa = w[2 * e], wb = w[2 * e + 1];
    return wa.x * (x[id.x] - xi) + wa.y * (x[id.y] - xi) + wb.x * (x[id.z] - xi) +
           wb.y * (x[id.w] - xi);
}
__device__ __forceinline__ double ell_diag(const double2 *w, int64_t n, int k, int64_t i) {
    const int64_t e = (int64_t)k * n + i;
    double2 wa = w[2 * e], wb = w[2 * e + 1];
    return -(wa.x + wa.y + wb.x + wb.y);
}
__device__ __forceinline__ double ell_jac_row(const int4 *idx, const double2 *w, int64_t n,
                                              const EllJacDev &J, int64_t i, const double *x) {
    double xi = x[i];
    int k = J.kmin[i];
    if (k < 0) return xi;
    double y = J.cmin * ell_row(idx, w, n, k, i, x, xi);
    if (J.kmax) y += J.cmax * ell_row(idx, w, n, J.kmax[i], i, x, xi);
    return y;
}

// mode 0: y = J x   1: diag (or 1/diag when invert)   2: Av (partial <rt, y>)   3: At (<y,s>,<y,y>)
template <int MODE, int NRED>
__global__ void __launch_bounds__(PASS_THREADS)
k_ell_rows(int64_t n, int64_t npts, const int4 *__restrict__ idx, const double2 *__restrict__ w,
           EllJacDev J, const double *__restrict__ x, const double *__restrict__ aux,
           double *__restrict__ y, int invert, double *partials, BicgScal *scal,
           const StageArgs sa) {
    DD acc[NRED > 0 ? NRED : 1] = {};
    bool skip = scal && (scal->done || (MODE == 3 && scal->half));
    if (!skip) {
        for (int64_t i = blockIdx.x * (int64_t)PASS_THREADS + threadIdx.x; i < npts;
             i += (int64_t)gridDim.x * PASS_THREADS) {
            double v;
            if (MODE == 1) {
                v = 1.0;
                if (i < n && J.kmin[i] >= 0) {
                    v = J.cmin * ell_diag(w, n, J.kmin[i], i);
                    if (J.kmax) v += J.cmax * ell_diag(w, n, J.kmax[i], i);
                }
                if (invert) v = 1.0 / v;
            } else {
                v = i < n ? ell_jac_row(idx, w, n, J, i, x) : x[i];
            }
            y[i] = v;
            if (MODE == 2) dd_mac(acc[0], aux[i], v);
            if (MODE == 3) {
                dd_mac(acc[0], v, aux[i]);
                dd_mac(acc[1], v, v);
            }
        }
    }
    if constexpr (NRED > 0)
        pass_finish(acc, partials + (int64_t)blockIdx.x * pde::DDW * NRED, partials, scal, sa);
}

// residuals: mode 0 convex envelope, mode 1 Pucci
__global__ void __launch_bounds__(256)
k_ell_residual(int64_t n, int64_t npts, int nds, const int4 *__restrict__ idx,
               const double2 *__restrict__ w, const double *__restrict__ W,
               const double *__restrict__ g, double A0, double A1, int mode,
               double *__restrict__ F, int32_t *__restrict__ kmin_out,
               int32_t *__restrict__ kmax_out, double *red_out) {
    double rmax = 0.0;
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < npts) {
        double wi = W[i];
        double f = wi - g[i];
        if (i < n) {
            double mn = __longlong_as_double(0x7ff0000000000000LL);
            double mx = __longlong_as_double(0xfff0000000000000LL);
            int an = 0, ax = 0;
            for (int k = 0; k < nds; ++k) {
                double d = ell_row(idx, w, n, k, i, W, wi);
                if (d < mn) { mn = d; an = k; }
                if (d >= mx) { mx = d; ax = k; }
            }
            if (mode == 0) {                       // convex_envelope.py:44-46
                double nl = -mn;
                bool active = nl > f;
                if (active) f = nl;
                if (kmin_out) kmin_out[i] = active ? an : -1;
            } else {                               // pucci.py:48,128,141: A0*lmax + A1*lmin - f
                f = (A0 * mx + A1 * mn) - g[i];
                if (kmin_out) kmin_out[i] = an;
                if (kmax_out) kmax_out[i] = ax;
            }
        }
        F[i] = f;
        rmax = fabs(f);
    }
    rmax = warp_max(rmax);
    if ((threadIdx.x & 31) == 0 && red_out) atomic_max_nonneg(red_out, rmax);
}

}  // namespace pde

extern "C" {

int pde_ell_residual(pde_ell *e, const double *d_W, const double *d_g, int mode, double A0,
                     double A1, double *d_F, int32_t *d_kmin, int32_t *d_kmax, double *d_red,
                     void *stream) {
    PDE_REQUIRE(e && d_W && d_g && d_F, "null argument");
    PDE_REQUIRE(e->row_offset == 0, "the point-cloud problems need the full table (row_offset 0)");
    DeviceGuard guard(e->device);
    unsigned grid = (unsigned)((e->n_points + 255) / 256);
    k_ell_residual<<<grid, 256, 0, (cudaStream_t)stream>>>(
        e->n_interior, e->n_points, e->nds, reinterpret_cast<const int4 *>(e->idx),
        reinterpret_cast<const double2 *>(e->w), d_W, d_g, A0, A1, mode, d_F, d_kmin, d_kmax, d_red);
    PDE_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"

static pde::EllJacDev ell_jac_dev(const pde_ell_jac *J) {
    pde::EllJacDev d;
    d.kmin = J->d_kmin;
    d.kmax = J->d_kmax;
    d.cmin = J->cmin;
    d.cmax = J->cmax;
    return d;
}

template <int MODE, int NRED>
static int ell_rows(pde_ell *e, const pde::EllJacDev &J, const double *x, const double *aux,
                    double *y, int invert, double *partials, pde::BicgScal *scal,
                    cudaStream_t st, int *nb,
                    pde::StageArgs sa = pde::StageArgs{pde::ST_NONE, 0, 0.0, 0.0}) {
    PDE_REQUIRE(e->row_offset == 0, "the point-cloud Jacobian needs the full table (row_offset 0)");
    int nblk = pde::elem_grid(e->n_points, e->num_sms);
    sa.total = nblk;
    pde::k_ell_rows<MODE, NRED><<<nblk, pde::PASS_THREADS, 0, st>>>(
        e->n_interior, e->n_points, reinterpret_cast<const int4 *>(e->idx),
        reinterpret_cast<const double2 *>(e->w), J, x, aux, y, invert, partials, scal, sa);
    PDE_LAUNCH_CHECK();
    if (nb) *nb = nblk;
    return 0;
}

extern "C" {

int pde_ell_jac_matvec(pde_ell *e, const pde_ell_jac *J, const double *d_x, double *d_y,
                       void *stream) {
    PDE_REQUIRE(e && J && J->d_kmin && d_x && d_y, "null argument");
    DeviceGuard guard(e->device);
    return ell_rows<0, 0>(e, ell_jac_dev(J), d_x, nullptr, d_y, 0, nullptr, nullptr,
                          (cudaStream_t)stream, nullptr);
}

int pde_ell_jac_diagonal(pde_ell *e, const pde_ell_jac *J, double *d_diag, void *stream) {
    PDE_REQUIRE(e && J && J->d_kmin && d_diag, "null argument");
    DeviceGuard guard(e->device);
    return ell_rows<1, 0>(e, ell_jac_dev(J), nullptr, nullptr, d_diag, 0, nullptr, nullptr,
                          (cudaStream_t)stream, nullptr);
}

int pde_ell_bicgstab(pde_ell *e, const pde_ell_jac *J, const double *d_b, double *d_x, double rtol,
                     double atol, int64_t maxiter, int64_t check_every, double *d_work,
                     int64_t *h_info, void *stream) {
    PDE_REQUIRE(e && J && J->d_kmin && d_b && d_x && d_work && h_info, "null argument");
    DeviceGuard guard(e->device);
    struct EllBackend {
        pde_ell *e;
        pde::EllJacDev J;
        int64_t n() const { return e->n_points; }
        int64_t unknowns() const { return e->n_points; }
        int64_t max_row_blocks() const { return pde::elem_grid(e->n_points, e->num_sms); }
        int num_sms() const { return e->num_sms; }
        int dinv(double *out, cudaStream_t st) {
            return ell_rows<1, 0>(e, J, nullptr, nullptr, out, 1, nullptr, nullptr, st, nullptr);
        }
        int Av(const double *src, const double *rtilde, double *dst, double *partials,
               pde::BicgScal *scal, cudaStream_t st, int *nb, const pde::StageArgs &sa) {
            return ell_rows<2, 1>(e, J, src, rtilde, dst, 0, partials, scal, st, nb, sa);
        }
        int At(const double *src, const double *s, double *dst, double *partials,
               pde::BicgScal *scal, cudaStream_t st, int *nb, const pde::StageArgs &sa) {
            return ell_rows<3, 2>(e, J, src, s, dst, 0, partials, scal, st, nb, sa);
        }
        bool fused_st() const { return false; }
        int SAt(const double *, const double *, const double *, double *, double *, double *,
                pde::BicgScal *, cudaStream_t, int *, pde::StageArgs) {
            return pde::fail("no fused s / t pass on point clouds", __FILE__, __LINE__);
        }
    } backend{e, ell_jac_dev(J)};
    return pde::bicgstab_run(backend, d_b, d_x, rtol, atol, maxiter, check_every, d_work, h_info,
                             (cudaStream_t)stream);
}

int pde_ell_newton_update(pde_ell *e, const double *d_U, const double *d_d, const double *d_F,
                          const double *d_Jd, double *d_Unew, double *h_out, void *stream) {
    PDE_REQUIRE(e && d_U && d_d && d_F && d_Jd && d_Unew && h_out, "null argument");
    DeviceGuard guard(e->device);
    return pde::newton_update_launch(e->n_points, e->num_sms, d_U, d_d, d_F, d_Jd, d_Unew, h_out,
                                     (cudaStream_t)stream);
}

}  // extern "C"
