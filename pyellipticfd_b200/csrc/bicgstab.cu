// Matrix-free policy Jacobian (mat-vec, transpose, diagonal) and the
// Jacobi-preconditioned BiCGStab of the Newton / policy-iteration step.
//
// Replaces solvers.py:179-181 of the reference:
//     P = diags(1/Jac.diagonal());  d, info = bicgstab(Jac, -Gu, M=P, tol=...)
// and follows SciPy 1.18.1's `bicgstab` (scipy/sparse/linalg/_isolve/iterative.py)
// statement by statement -- same update order, same stopping and breakdown
// tests -- with every scalar kept on the device so that a whole batch of
// iterations runs without a host round trip.  Dot products are two-stage
// (per-block partials, then one fixed-order pass), hence deterministic.
#include <algorithm>

#include <cstdlib>

#include "jacobian.cuh"
#include "krylov.cuh"
#include "jac_tiled.cuh"
#include "tiles.cuh"

using namespace pde;

namespace pde {

// Lattice rows + band/wall tail.  Op provides deep(o) / band(bi,o) / wall(o) and NRED sums.
template <class Op>
__global__ void __launch_bounds__(PASS_THREADS)
k_rows(const GridDev G, const __grid_constant__ DeepTable tab, Op op, double *partials,
       BicgScal *scal, const int nrb /* row blocks; the rest are tail blocks */,
       const double *partials_base, const StageArgs sa) {
    DD acc[Op::NRED > 0 ? Op::NRED : 1] = {};
    bool skip = scal && (scal->done || (Op::kSkipOnHalf && scal->half));
    if (!skip) {
        op.prepare(scal);
        if ((int)blockIdx.x < nrb) {
            // one block per lattice row: consecutive blocks sweep consecutive rows, so the
            // +-r row gathers of a mat-vec hit L2 (measured faster than persistent row blocks)
            int64_t br = blockIdx.x;
            int64_t lr = br - G.halo_rows;
            if (lr >= 0 && lr < G.nx_local) {
                int64_t gx = lr + G.x_offset;
                if (deep_row(G, gx)) {
                    const int64_t base = br * G.pitch;
                    const int c1 = (int)G.ny - G.radius;
                    for (int c = G.radius + threadIdx.x; c < c1; c += PASS_THREADS)
                        op.deep(tab, G, base + c, acc);
                }
            }
        } else {
            int64_t i = ((int64_t)blockIdx.x - nrb) * PASS_THREADS + threadIdx.x;
            if (i < G.n_band)
                op.band(G, i, acc);
            else if (i < G.n_band + G.nb)
                op.wall(G.rows * G.pitch + (i - G.n_band), acc);
        }
    }
    // `partials`: first slot of THIS launch; `partials_base`: first slot of the pass (the tiled
    // launch that may precede this one on the stream has stored its sums there already)
    if constexpr (Op::NRED > 0)
        pass_finish(acc, partials + (int64_t)blockIdx.x * DDW * Op::NRED, partials_base, scal, sa);
}

// ---- J x ------------------------------------------------------------------------------
struct OpMatvec {
    static constexpr int NRED = 0;
    static constexpr bool kSkipOnHalf = false;
    static constexpr bool kTiled = true;
    JacDev J;
    const double *x;
    double *y;
    const double *source() const { return x; }
    __device__ const JacDev &jac() const { return J; }
    __device__ double aux(int64_t) const { return 0.0; }
    __device__ void tile(int64_t o, double v, double, DD *) { y[o] = v; }
    __device__ void prepare(const BicgScal *) {}
    __device__ void deep(const DeepTable &tab, const GridDev &G, int64_t o, DD *) {
        y[o] = jac_row_deep(tab, J, x, o, G.pitch);
    }
    __device__ void band(const GridDev &G, int64_t bi, DD *) {
        y[G.band_center[bi]] = jac_row_band(G, J, x, bi);
    }
    __device__ void wall(int64_t o, DD *) { y[o] = x[o]; }
};

// y += J^T x by scatter (solvers.py:192 only; not on the Krylov path)
struct OpMatvecT {
    static constexpr int NRED = 0;
    static constexpr bool kSkipOnHalf = false;
    static constexpr bool kTiled = false;
    JacDev J;
    const double *x;
    double *y;
    __device__ void prepare(const BicgScal *) {}
    __device__ void scatter(double xi, int64_t o, int64_t j0, int64_t j1, double w_0, double w_1,
                            double w0, double coef) {
        double A0, A1, Ac;
        row_entries(w_0, w_1, w0, A0, A1, Ac);
        atomicAdd(&y[j0], coef * A0 * xi);
        atomicAdd(&y[j1], coef * A1 * xi);
        atomicAdd(&y[o], coef * Ac * xi);
    }
    __device__ void deep(const DeepTable &tab, const GridDev &G, int64_t o, DD *) {
        unsigned pm = J.pmin[o];
        double xi = x[o];
        if (pm == PDE_POLICY_IDENTITY) {
            atomicAdd(&y[o], xi);
            return;
        }
        scatter(xi, o, o + tab.a0[pm] * G.pitch + tab.b0[pm], o + tab.a1[pm] * G.pitch + tab.b1[pm],
                tab.w_0[pm], tab.w_1[pm], tab.w0[pm], J.cmin ? J.cmin[o] : J.cmin_s);
        if (J.pmax) {
            unsigned px = J.pmax[o];
            scatter(xi, o, o + tab.a0[px] * G.pitch + tab.b0[px],
                    o + tab.a1[px] * G.pitch + tab.b1[px], tab.w_0[px], tab.w_1[px], tab.w0[px],
                    J.cmax ? J.cmax[o] : J.cmax_s);
        }
    }
    __device__ void band(const GridDev &G, int64_t bi, DD *) {
        int64_t o = G.band_center[bi];
        unsigned pm = J.pmin[o];
        double xi = x[o];
        if (pm == PDE_POLICY_IDENTITY) {
            atomicAdd(&y[o], xi);
            return;
        }
        int64_t k = G.band_ptr[bi] + pm;
        scatter(xi, o, G.band_j[2 * k], G.band_j[2 * k + 1], G.band_w[3 * k], G.band_w[3 * k + 1],
                G.band_w[3 * k + 2], J.cmin ? J.cmin[o] : J.cmin_s);
        if (J.pmax) {
            k = G.band_ptr[bi] + J.pmax[o];
            scatter(xi, o, G.band_j[2 * k], G.band_j[2 * k + 1], G.band_w[3 * k],
                    G.band_w[3 * k + 1], G.band_w[3 * k + 2], J.cmax ? J.cmax[o] : J.cmax_s);
        }
    }
    __device__ void wall(int64_t o, DD *) { atomicAdd(&y[o], x[o]); }
};

struct OpDiag {
    static constexpr int NRED = 0;
    static constexpr bool kSkipOnHalf = false;
    static constexpr bool kTiled = false;
    JacDev J;
    double *d;
    int invert;
    __device__ void prepare(const BicgScal *) {}
    __device__ void deep(const DeepTable &tab, const GridDev &, int64_t o, DD *) {
        double v = jac_diag_deep(tab, J, o);
        d[o] = invert ? 1.0 / v : v;
    }
    __device__ void band(const GridDev &G, int64_t bi, DD *) {
        double v = jac_diag_band(G, J, bi);
        d[G.band_center[bi]] = invert ? 1.0 / v : v;
    }
    __device__ void wall(int64_t o, DD *) { d[o] = 1.0; }
};

// ---- BiCGStab passes -------------------------------------------------------------------
// v = J phat,  partial <rtilde, v>          (iterative.py: v = matvec(phat); rv = dot(rtilde, v))
struct OpAv {
    static constexpr int NRED = 1;
    static constexpr bool kSkipOnHalf = false;
    static constexpr bool kTiled = true;
    JacDev J;
    const double *src, *rtilde;
    double *dst;
    const double *source() const { return src; }
    __device__ const JacDev &jac() const { return J; }
    __device__ double aux(int64_t o) const { return rtilde[o]; }
    __device__ void tile(int64_t o, double v, double rt, DD *acc) {
        dst[o] = v;
        dd_mac(acc[0], rt, v);
    }
    __device__ void prepare(const BicgScal *) {}
    __device__ void deep(const DeepTable &tab, const GridDev &G, int64_t o, DD *acc) {
        double v = jac_row_deep(tab, J, src, o, G.pitch);
        dst[o] = v;
        dd_mac(acc[0], rtilde[o], v);
    }
    __device__ void band(const GridDev &G, int64_t bi, DD *acc) {
        int64_t o = G.band_center[bi];
        double v = jac_row_band(G, J, src, bi);
        dst[o] = v;
        dd_mac(acc[0], rtilde[o], v);
    }
    __device__ void wall(int64_t o, DD *acc) {
        double v = src[o];
        dst[o] = v;
        dd_mac(acc[0], rtilde[o], v);
    }
};

// t = J shat, partial <t, s>, <t, t>        (omega = dot(t, s) / dot(t, t))
struct OpAt {
    static constexpr int NRED = 2;
    static constexpr bool kSkipOnHalf = true;
    static constexpr bool kTiled = true;
    JacDev J;
    const double *src, *s;
    double *dst;
    const double *source() const { return src; }
    __device__ const JacDev &jac() const { return J; }
    __device__ double aux(int64_t o) const { return s[o]; }
    __device__ void tile(int64_t o, double t, double so, DD *acc) {
        dst[o] = t;
        dd_mac(acc[0], t, so);
        dd_mac(acc[1], t, t);
    }
    __device__ void prepare(const BicgScal *) {}
    __device__ void one(int64_t o, double t, DD *acc) {
        dst[o] = t;
        dd_mac(acc[0], t, s[o]);
        dd_mac(acc[1], t, t);
    }
    __device__ void deep(const DeepTable &tab, const GridDev &G, int64_t o, DD *acc) {
        one(o, jac_row_deep(tab, J, src, o, G.pitch), acc);
    }
    __device__ void band(const GridDev &G, int64_t bi, DD *acc) {
        one(G.band_center[bi], jac_row_band(G, J, src, bi), acc);
    }
    __device__ void wall(int64_t o, DD *acc) { one(o, src[o], acc); }
};

// band points and wall entries of the fused s / t pass (jac_tiled.cuh::k_jac_fused_st): the
// neighbours' shat = dinv*(r - alpha*v) are recomputed per gathered value
struct OpSAtTail {
    static constexpr int NRED = 3;
    static constexpr bool kSkipOnHalf = false;
    static constexpr bool kTiled = false;
    JacDev J;
    const double *r, *v, *dinv;
    double *s_out, *t_out;
    double alpha;
    __device__ void prepare(const BicgScal *s) { alpha = s->alpha; }
    __device__ double sval(int64_t o) const { return __dsub_rn(r[o], __dmul_rn(alpha, v[o])); }
    __device__ double shat(int64_t o) const { return dinv[o] * sval(o); }
    __device__ void one(int64_t o, double sv, double tv, DD *acc) {
        s_out[o] = sv;
        t_out[o] = tv;
        dd_mac(acc[0], sv, sv);
        dd_mac(acc[1], tv, sv);
        dd_mac(acc[2], tv, tv);
    }
    __device__ void deep(const DeepTable &, const GridDev &, int64_t, DD *) {}   // tiled kernel
    __device__ void band(const GridDev &G, int64_t bi, DD *acc) {
        const int64_t o = G.band_center[bi];
        const double sv = sval(o);
        const double tv = jac_row_band_get(G, J, bi, dinv[o] * sv,
                                           [this](int64_t j) { return shat(j); });
        one(o, sv, tv, acc);
    }
    __device__ void wall(int64_t o, DD *acc) {
        const double sv = sval(o);
        one(o, sv, dinv[o] * sv, acc);
    }
};

template <int H>
static int launch_fused_st_h(pde_grid *g, const JacDev &J, const double *r, const double *v,
                             const double *dinv, double *s_out, double *t_out, double *partials,
                             BicgScal *scal, cudaStream_t st, int *nblocks) {
    constexpr int TX = KF_THREADS / K1_TY * 4;
    DeepRegion dr = deep_region(g, 0, g->nx_local);
    *nblocks = 0;
    if (dr.nrows <= 0 || dr.ncols <= 0) return 0;
    DeepGeom geo;
    geo.pitch = g->pitch;
    geo.row0 = dr.row0;
    geo.col0 = dr.col0;
    geo.nrows = dr.nrows;
    geo.ncols = dr.ncols;
    geo.tiles_x = (dr.nrows + TX - 1) / TX;
    geo.tiles_y = (dr.ncols + K1_TY - 1) / K1_TY;
    geo.split = geo.tiles_x;
    geo.row0b = 0;
    geo.nrowsb = 0;
    CUtensorMap mr, mv, md;
    int rc;
    if ((rc = make_tmap(g, r, K1_TY + 2 * H + 2, TX + 2 * H, &mr))) return rc;
    if ((rc = make_tmap(g, v, K1_TY + 2 * H + 2, TX + 2 * H, &mv))) return rc;
    if ((rc = make_tmap(g, dinv, K1_TY + 2 * H + 2, TX + 2 * H, &md))) return rc;
    constexpr int smem = k_jac_fused_smem_bytes<H>();
    auto kern = k_jac_fused_st<H>;
    static unsigned long long attr_set = 0ull;
    const unsigned long long dev_bit = 1ull << (g->device & 63);
    if (!(attr_set & dev_bit)) {
        PDE_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_set |= dev_bit;
    }
    int ntiles = geo.tiles_x * geo.tiles_y;
    int grid = std::min(ntiles, g->num_sms);
    kern<<<grid, KF_THREADS, smem, st>>>(mr, mv, md, g->tab, geo, J, s_out, t_out, partials, scal);
    PDE_LAUNCH_CHECK();
    *nblocks = grid;
    return 0;
}

// Opt-in (PDE_JAC_FUSED=1, read at every solve): the fused pass is 6 % faster per iteration
// (0.60 vs 0.64 ms at 4097^2) and its first iterate agrees with the separate passes to 4e-16,
// but it regroups the partial sums of <s,s>, <t,s>, <t,t>, and BiCGStab on the policy Jacobians
// amplifies a rounding-level perturbation tenfold per iteration (tools/fused_check.py): the
// iteration COUNT of a solve is a lottery in the grouping (certifying solve at 4097^2: 492
// iterations with the separate passes, 843 fused; tools/certify_counts.py).  The default keeps
// the grouping the committed iteration counts and tests were measured with.
static bool fused_st_enabled() {
    const char *e = getenv("PDE_JAC_FUSED");
    return e && atoi(e) != 0;
}

// deep-interior rows through the TMA-tiled kernel (jac_tiled.cuh); *nblocks = CTAs launched
template <class Op, int H>
static int launch_tiled_h(pde_grid *g, const Op &op, double *partials, BicgScal *scal,
                          cudaStream_t st, int *nblocks) {
    constexpr int P = 8, TX = 2 * P;
    DeepRegion dr = deep_region(g, 0, g->nx_local);
    *nblocks = 0;
    if (dr.nrows <= 0 || dr.ncols <= 0) return 0;
    DeepGeom geo;
    geo.pitch = g->pitch;
    geo.row0 = dr.row0;
    geo.col0 = dr.col0;
    geo.nrows = dr.nrows;
    geo.ncols = dr.ncols;
    geo.tiles_x = (dr.nrows + TX - 1) / TX;
    geo.tiles_y = (dr.ncols + K1_TY - 1) / K1_TY;
    geo.split = geo.tiles_x;
    geo.row0b = 0;
    geo.nrowsb = 0;
    CUtensorMap tmap;
    int rc = make_tmap(g, op.source(), K1_TY + 2 * H + 2, TX + 2 * H, &tmap);
    if (rc) return rc;
    constexpr int smem = k_jac_tiled_smem_bytes<H, P>();
    auto kern = k_jac_tiled<Op, H, P>;
    // the opt-in is per device: one bit per device, per template instantiation
    static unsigned long long attr_set = 0ull;
    const unsigned long long dev_bit = 1ull << (g->device & 63);
    if (!(attr_set & dev_bit)) {
        PDE_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_set |= dev_bit;
    }
    int ntiles = geo.tiles_x * geo.tiles_y;
    int grid = std::min(ntiles, g->num_sms * 3);
    kern<<<grid, K1_THREADS, smem, st>>>(tmap, g->tab, geo, op, partials, scal);
    PDE_LAUNCH_CHECK();
    *nblocks = grid;
    return 0;
}

template <class Op>
static int launch_tiled(pde_grid *g, const Op &op, double *partials, BicgScal *scal,
                        cudaStream_t st, int *nblocks) {
    switch (g->tab.H) {
        case 1: return launch_tiled_h<Op, 1>(g, op, partials, scal, st, nblocks);
        case 2: return launch_tiled_h<Op, 2>(g, op, partials, scal, st, nblocks);
        case 3: return launch_tiled_h<Op, 3>(g, op, partials, scal, st, nblocks);
        case 4: return launch_tiled_h<Op, 4>(g, op, partials, scal, st, nblocks);
        case 5: return launch_tiled_h<Op, 5>(g, op, partials, scal, st, nblocks);
        case 6: return launch_tiled_h<Op, 6>(g, op, partials, scal, st, nblocks);
    }
    return pde::fail("unsupported stencil halo", __FILE__, __LINE__);
}

static bool jac_tiled_enabled() {
    static const bool on = !(getenv("PDE_JAC_TILED") && atoi(getenv("PDE_JAC_TILED")) == 0);
    return on;
}

template <class Op>
static int launch_rows(pde_grid *g, const Op &op, double *partials, BicgScal *scal,
                       cudaStream_t st, int *nblocks_out = nullptr,
                       StageArgs sa = StageArgs{ST_NONE, 0, 0.0, 0.0}) {
    GridDev G = grid_dev(g);
    if constexpr (Op::kTiled) {
        int64_t tb = tail_blocks(g->n_band, g->nb);
        if (g->tab.D > 0 && g->nmid == 0 && g->tab.H >= 1 && g->tab.H <= 6 && jac_tiled_enabled() &&
            tb > 0 && (reinterpret_cast<uintptr_t>(op.source()) & 15) == 0) {      // TMA: 16-byte aligned base
            // deep-interior points: TMA-tiled kernel; band points and wall entries: the tail
            // blocks of the plain kernel, their partial sums appended after the tiled CTAs';
            // the last tail block performs the pass's scalar stage
            int nd = 0;
            int rc = launch_tiled(g, op, partials, scal, st, &nd);
            if (rc) return rc;
            sa.total = nd + (int)tb;
            k_rows<Op><<<(unsigned)tb, PASS_THREADS, 0, st>>>(
                G, g->tab, op, partials ? partials + (int64_t)nd * DDW * Op::NRED : nullptr, scal, 0,
                partials, sa);
            PDE_LAUNCH_CHECK();
            if (nblocks_out) *nblocks_out = nd + (int)tb;
            return 0;
        }
    }
    int nrb = (int)g->rows;
    int64_t nblk = nrb + tail_blocks(g->n_band, g->nb);
    sa.total = (int)nblk;
    k_rows<Op><<<(unsigned)nblk, PASS_THREADS, 0, st>>>(G, g->tab, op, partials, scal, nrb, partials, sa);
    PDE_LAUNCH_CHECK();
    if (nblocks_out) *nblocks_out = (int)nblk;
    return 0;
}

// U_new = U + d, max|U - U_new|, <F, J d>, <d, d> (solvers.py:188-197).  Fixed-order sums: per
// thread over a grid-stride sequence, shuffle tree per warp, warps in order per block, and the
// last block to finish adds the block partials in block order -- the same bits on every run.
__global__ void k_newton_update(int64_t n, const double *U, const double *d, const double *F,
                                const double *Jd, double *Unew, double *out /* 3 */,
                                double *partials /* 3 x gridDim */, unsigned *counter) {
    __shared__ double sh[3][32];
    __shared__ bool last;
    double mx = 0.0, a = 0.0, b = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        double u = U[i], di = d[i];
        double un = __dadd_rn(u, di);
        Unew[i] = un;
        mx = fmax(mx, fabs(__dsub_rn(u, un)));
        a += F[i] * Jd[i];
        b += di * di;
    }
    mx = warp_max(mx);
    a = warp_sum(a);
    b = warp_sum(b);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if (lane == 0) {
        sh[0][wid] = mx;
        sh[1][wid] = a;
        sh[2][wid] = b;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < nw; ++w) {
            mx = fmax(mx, sh[0][w]);
            a += sh[1][w];
            b += sh[2][w];
        }
        partials[3 * blockIdx.x] = mx;
        partials[3 * blockIdx.x + 1] = a;
        partials[3 * blockIdx.x + 2] = b;
        __threadfence();
        last = atomicAdd(counter, 1u) == gridDim.x - 1;
        if (last) {
            __threadfence();
            mx = a = b = 0.0;
            for (unsigned k = 0; k < gridDim.x; ++k) {
                mx = fmax(mx, partials[3 * k]);
                a += partials[3 * k + 1];
                b += partials[3 * k + 2];
            }
            out[0] = mx;
            out[1] = a;
            out[2] = b;
        }
    }
}

}  // namespace pde

extern "C" {

int pde_jac_matvec(pde_grid *g, const pde_jac *J, const double *d_x, double *d_y, int transpose,
                   void *stream) {
    PDE_REQUIRE(g && J && d_x && d_y && J->d_pmin, "null argument");
    DeviceGuard guard(g->device);
    cudaStream_t st = (cudaStream_t)stream;
    if (!transpose) {
        OpMatvec op{jac_dev(J), d_x, d_y};
        return launch_rows(g, op, nullptr, nullptr, st);
    }
    PDE_CHECK(cudaMemsetAsync(d_y, 0, sizeof(double) * pde_grid_state_len(g), st));
    OpMatvecT op{jac_dev(J), d_x, d_y};
    return launch_rows(g, op, nullptr, nullptr, st);
}

int pde_jac_diagonal(pde_grid *g, const pde_jac *J, double *d_diag, void *stream) {
    PDE_REQUIRE(g && J && d_diag && J->d_pmin, "null argument");
    DeviceGuard guard(g->device);
    cudaStream_t st = (cudaStream_t)stream;
    PDE_CHECK(cudaMemsetAsync(d_diag, 0, sizeof(double) * pde_grid_state_len(g), st));
    OpDiag op{jac_dev(J), d_diag, 0};
    return launch_rows(g, op, nullptr, nullptr, st);
}

int pde_ddg_apply_policy(pde_grid *g, const uint16_t *d_policy, const double *d_x, double *d_y,
                         void *stream) {
    pde_jac J{d_policy, nullptr, 1.0, 0.0, nullptr, nullptr};
    return pde_jac_matvec(g, &J, d_x, d_y, 0, stream);
}

int pde_bicgstab(pde_grid *g, const pde_jac *J, const double *d_b, double *d_x, double rtol,
                 double atol, int64_t maxiter, int64_t check_every, double *d_work,
                 int64_t *h_info, void *stream) {
    PDE_REQUIRE(g && J && d_b && d_x && d_work && h_info && J->d_pmin, "null argument");
    DeviceGuard guard(g->device);
    struct GridBackend {
        pde_grid *g;
        JacDev J;
        int64_t n() const { return g->rows * g->pitch + g->nb; }
        int64_t unknowns() const { return g->nx_local * g->ny + g->nb; }
        int64_t max_row_blocks() const { return g->rows + tail_blocks(g->n_band, g->nb); }
        int num_sms() const { return g->num_sms; }
        int dinv(double *out, cudaStream_t st) { return launch_rows(g, OpDiag{J, out, 1}, nullptr, nullptr, st); }
        int Av(const double *src, const double *rtilde, double *dst, double *partials,
               BicgScal *scal, cudaStream_t st, int *nb, const StageArgs &sa) {
            return launch_rows(g, OpAv{J, src, rtilde, dst}, partials, scal, st, nb, sa);
        }
        int At(const double *src, const double *s, double *dst, double *partials,
               BicgScal *scal, cudaStream_t st, int *nb, const StageArgs &sa) {
            return launch_rows(g, OpAt{J, src, s, dst}, partials, scal, st, nb, sa);
        }
        // fused s / t pass: unsharded 2-D lattices with a tiled table and a band / wall tail
        bool fused_st() const {
            return fused_st_enabled() && jac_tiled_enabled() && g->tab.D > 0 && g->nmid == 0 &&
                   g->tab.H >= 1 && g->tab.H <= 6 && g->halo_rows == 0 &&
                   tail_blocks(g->n_band, g->nb) > 0;
        }
        int SAt(const double *r, const double *v, const double *dinv, double *s_out, double *t_out,
                double *partials, BicgScal *scal, cudaStream_t st, int *nb, StageArgs sa) {
            int nd = 0, rc = 1;
            switch (g->tab.H) {
                case 1: rc = launch_fused_st_h<1>(g, J, r, v, dinv, s_out, t_out, partials, scal, st, &nd); break;
                case 2: rc = launch_fused_st_h<2>(g, J, r, v, dinv, s_out, t_out, partials, scal, st, &nd); break;
                case 3: rc = launch_fused_st_h<3>(g, J, r, v, dinv, s_out, t_out, partials, scal, st, &nd); break;
                case 4: rc = launch_fused_st_h<4>(g, J, r, v, dinv, s_out, t_out, partials, scal, st, &nd); break;
                case 5: rc = launch_fused_st_h<5>(g, J, r, v, dinv, s_out, t_out, partials, scal, st, &nd); break;
                case 6: rc = launch_fused_st_h<6>(g, J, r, v, dinv, s_out, t_out, partials, scal, st, &nd); break;
            }
            if (rc) return rc;
            int64_t tb = tail_blocks(g->n_band, g->nb);
            sa.total = nd + (int)tb;
            k_rows<OpSAtTail><<<(unsigned)tb, PASS_THREADS, 0, st>>>(
                grid_dev(g), g->tab, OpSAtTail{J, r, v, dinv, s_out, t_out, 0.0},
                partials + (int64_t)nd * 3 * DDW, scal, 0, partials, sa);
            PDE_LAUNCH_CHECK();
            if (nb) *nb = nd + (int)tb;
            return 0;
        }
    } backend{g, jac_dev(J)};
    return bicgstab_run(backend, d_b, d_x, rtol, atol, maxiter, check_every, d_work, h_info,
                        (cudaStream_t)stream);
}

// ---- slab-parallel BiCGStab: the same passes, one call per step -----------------------------
// The caller (pyellipticfd_b200/multigpu.py) puts the NCCL traffic between the steps: a halo
// exchange of phat / shat before the two operator passes and a SUM all-reduce of d_sums
// between every "pass" step and its "finalize" step.  Dot products need no masking: halo
// rows of r, rtilde, v, t are identically zero, and the replicated wall entries are zero
// whenever the iterate satisfies the Dirichlet data (F_wall = 0), which the caller checks.
int64_t pde_bicgstab_scal_bytes(void) { return (int64_t)sizeof(BicgScal); }

int pde_bicgstab_dist_step(pde_grid *g, const pde_jac *J, int step, const double *d_b, double *d_x,
                           double *d_work, void *d_scal, double *d_partials, double *d_sums,
                           double rtol, double atol, int64_t maxiter, void *stream) {
    PDE_REQUIRE(g && J && d_b && d_x && d_work && d_scal && d_partials && d_sums && J->d_pmin,
                "null argument");
    DeviceGuard guard(g->device);
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t n = g->rows * g->pitch + g->nb;
    const int sms = g->num_sms;
    BicgScal *scal = static_cast<BicgScal *>(d_scal);
    JacDev Jd = jac_dev(J);
    double *r = d_work, *rtilde = r + n, *p = rtilde + n, *v = p + n, *phat = v + n,
           *t = phat + n, *shat = t + n, *dinv = shat + n;   // phat, shat at even multiples of n:
    // 16-byte aligned sources for the TMA-tiled operator passes whatever the parity of n
    int rc = 0, nb = 0;
    auto reduce = [&](int stage, int nred) {
        if (nred == 1) k_stage<1><<<1, PASS_THREADS, 0, st>>>(stage, d_partials, nb, scal, rtol, atol, d_sums, 1);
        else k_stage<2><<<1, PASS_THREADS, 0, st>>>(stage, d_partials, nb, scal, rtol, atol, d_sums, 1);
        pde::count_launch();
    };
    auto finalize = [&](int stage, int nred) {
        if (nred == 1) k_stage<1><<<1, PASS_THREADS, 0, st>>>(stage, d_partials, 0, scal, rtol, atol, d_sums, 2);
        else k_stage<2><<<1, PASS_THREADS, 0, st>>>(stage, d_partials, 0, scal, rtol, atol, d_sums, 2);
        pde::count_launch();
    };
    switch (step) {
        case 0: {   // setup + init pass
            BicgScal h{};
            h.maxiter = maxiter > 0 ? maxiter : 10 * (g->nx_global * g->ny + g->nb);
            PDE_CHECK(cudaMemcpyAsync(scal, &h, sizeof h, cudaMemcpyHostToDevice, st));
            PDE_CHECK(cudaStreamSynchronize(st));     // h lives on this stack frame
            PDE_CHECK(cudaMemsetAsync(dinv, 0, sizeof(double) * n, st));
            if ((rc = launch_rows(g, OpDiag{Jd, dinv, 1}, nullptr, nullptr, st))) return rc;
            PDE_CHECK(cudaMemsetAsync(v, 0, sizeof(double) * n, st));
            PDE_CHECK(cudaMemsetAsync(t, 0, sizeof(double) * n, st));
            PDE_CHECK(cudaMemsetAsync(p, 0, sizeof(double) * n, st));
            PDE_CHECK(cudaMemsetAsync(d_sums, 0, sizeof(double) * 3, st));
            if ((rc = launch_elem(n, sms, OpInit{d_b, d_x, r, rtilde}, d_partials, nullptr, st, &nb))) return rc;
            reduce(ST_INIT, 1);
            break;
        }
        case 1: finalize(ST_INIT, 1); break;
        case 2: rc = launch_elem(n, sms, OpP{r, v, dinv, p, phat, 0, 0, 0}, nullptr, scal, st); break;
        case 3:
            if ((rc = launch_rows(g, OpAv{Jd, phat, rtilde, v}, d_partials, scal, st, &nb))) return rc;
            reduce(ST_V, 1);
            break;
        case 4: finalize(ST_V, 1); break;
        case 5:
            if ((rc = launch_elem(n, sms, OpS{v, dinv, r, shat, 0}, d_partials, scal, st, &nb))) return rc;
            reduce(ST_S, 1);
            break;
        case 6: finalize(ST_S, 1); break;
        case 7:
            if ((rc = launch_rows(g, OpAt{Jd, shat, r, t}, d_partials, scal, st, &nb))) return rc;
            reduce(ST_T, 2);
            break;
        case 8: finalize(ST_T, 2); break;
        case 9:
            if ((rc = launch_elem(n, sms, OpX{phat, shat, t, rtilde, d_x, r, 0, 0, 0}, d_partials, scal, st, &nb))) return rc;
            reduce(ST_X, 2);
            break;
        case 10: finalize(ST_X, 2); break;
        default: return pde::fail("bad step", __FILE__, __LINE__);
    }
    if (rc) return rc;
    PDE_LAUNCH_CHECK();
    return 0;
}

// h_out = [done, info, iterations]; synchronises the stream
int pde_bicgstab_poll(const void *d_scal, int64_t *h_out, void *stream) {
    PDE_REQUIRE(d_scal && h_out, "null argument");
    struct {
        long long iter, info, maxiter;
        int done, half, first, pad;
    } t;
    cudaStream_t st = (cudaStream_t)stream;
    PDE_CHECK(cudaMemcpyAsync(&t, static_cast<const char *>(d_scal) + offsetof(BicgScal, iter), sizeof t,
                              cudaMemcpyDeviceToHost, st));
    PDE_CHECK(cudaStreamSynchronize(st));
    h_out[0] = t.done;
    h_out[1] = t.info;
    h_out[2] = t.iter;
    return 0;
}

int pde_vec_newton_update(pde_grid *g, const double *d_U, const double *d_d, const double *d_F,
                          const double *d_Jd, double *d_Unew, double *h_out, void *stream) {
    PDE_REQUIRE(g && d_U && d_d && d_F && d_Jd && d_Unew && h_out, "null argument");
    DeviceGuard guard(g->device);
    return pde::newton_update_launch(pde_grid_state_len(g), g->num_sms, d_U, d_d, d_F, d_Jd, d_Unew,
                                     h_out, (cudaStream_t)stream);
}

}  // extern "C"

namespace pde {
int newton_update_launch(int64_t n, int num_sms, const double *U, const double *d, const double *F,
                         const double *Jd, double *Unew, double *h_out, cudaStream_t st) {
    const unsigned grid = (unsigned)elem_grid(n, num_sms);
    double *out = nullptr;                       // [3 results | counter word | 3 x grid partials]
    const size_t bytes = (4 + 3 * (size_t)grid) * sizeof(double);
    PDE_CHECK(cudaMallocAsync((void **)&out, bytes, st));
    PDE_CHECK(cudaMemsetAsync(out, 0, 4 * sizeof(double), st));
    k_newton_update<<<grid, 256, 0, st>>>(n, U, d, F, Jd, Unew, out, out + 4,
                                          reinterpret_cast<unsigned *>(out + 3));
    pde::count_launch();
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(h_out, out, 3 * sizeof(double), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFreeAsync(out, st);
    if (e != cudaSuccess) return pde::fail_cuda(e, "newton_update", __FILE__, __LINE__);
    return 0;
}
}  // namespace pde
