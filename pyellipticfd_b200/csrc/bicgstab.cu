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
#include "jacobian.cuh"

using namespace pde;

namespace pde {

constexpr int PASS_THREADS = 256;
constexpr int MAX_RED = 3;

struct BicgScal {
    double bnrm2, atol, rho, rho_prev, alpha, omega, beta, rv, ss, ts, tt, rr, rtr;
    double sums[MAX_RED];
    long long iter, info, maxiter;
    int done, half, first, pad;
};

template <int NRED>
__device__ __forceinline__ void block_store_partials(double (&acc)[NRED], double *partials) {
    __shared__ double sh[NRED][PASS_THREADS / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < NRED; ++j) {
        double v = warp_sum(acc[j]);
        if (lane == 0) sh[j][wid] = v;
    }
    __syncthreads();
    if (threadIdx.x < NRED) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < PASS_THREADS / 32; ++w) s += sh[threadIdx.x][w];
        partials[(int64_t)blockIdx.x * NRED + threadIdx.x] = s;
    }
}

// Lattice rows + band/wall tail.  Op provides deep(o) / band(bi,o) / wall(o) and NRED sums.
template <class Op>
__global__ void __launch_bounds__(PASS_THREADS)
k_rows(const GridDev G, const __grid_constant__ DeepTable tab, Op op, double *partials,
       const BicgScal *scal) {
    double acc[Op::NRED > 0 ? Op::NRED : 1] = {0.0};
    bool skip = scal && (scal->done || (Op::kSkipOnHalf && scal->half));
    if (!skip) {
        op.prepare(scal);
        if ((int64_t)blockIdx.x < G.rows) {
            int64_t br = blockIdx.x;
            int64_t lr = br - G.halo_rows;
            if (lr >= 0 && lr < G.nx_local) {
                int64_t gx = lr + G.x_offset;
                bool deep_row = gx >= G.radius && gx < G.nx_global - G.radius;
                for (int64_t c = threadIdx.x; c < G.ny; c += PASS_THREADS) {
                    bool deep = deep_row && c >= G.radius && c < G.ny - G.radius;
                    if (deep) op.deep(tab, G, br * G.pitch + c, acc);
                }
            }
        } else {
            int64_t i = ((int64_t)blockIdx.x - G.rows) * PASS_THREADS + threadIdx.x;
            if (i < G.n_band)
                op.band(G, i, acc);
            else if (i < G.n_band + G.nb)
                op.wall(G.rows * G.pitch + (i - G.n_band), acc);
        }
    }
    if (Op::NRED > 0) block_store_partials(acc, partials);
}

// ---- J x ------------------------------------------------------------------------------
struct OpMatvec {
    static constexpr int NRED = 0;
    static constexpr bool kSkipOnHalf = false;
    JacDev J;
    const double *x;
    double *y;
    __device__ void prepare(const BicgScal *) {}
    __device__ void deep(const DeepTable &tab, const GridDev &G, int64_t o, double *) {
        y[o] = jac_row_deep(tab, J, x, o, G.pitch);
    }
    __device__ void band(const GridDev &G, int64_t bi, double *) {
        y[G.band_center[bi]] = jac_row_band(G, J, x, bi);
    }
    __device__ void wall(int64_t o, double *) { y[o] = x[o]; }
};

// y += J^T x by scatter (solvers.py:192 only; not on the Krylov path)
struct OpMatvecT {
    static constexpr int NRED = 0;
    static constexpr bool kSkipOnHalf = false;
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
    __device__ void deep(const DeepTable &tab, const GridDev &G, int64_t o, double *) {
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
    __device__ void band(const GridDev &G, int64_t bi, double *) {
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
    __device__ void wall(int64_t o, double *) { atomicAdd(&y[o], x[o]); }
};

struct OpDiag {
    static constexpr int NRED = 0;
    static constexpr bool kSkipOnHalf = false;
    JacDev J;
    double *d;
    int invert;
    __device__ void prepare(const BicgScal *) {}
    __device__ void deep(const DeepTable &tab, const GridDev &, int64_t o, double *) {
        double v = jac_diag_deep(tab, J, o);
        d[o] = invert ? 1.0 / v : v;
    }
    __device__ void band(const GridDev &G, int64_t bi, double *) {
        double v = jac_diag_band(G, J, bi);
        d[G.band_center[bi]] = invert ? 1.0 / v : v;
    }
    __device__ void wall(int64_t o, double *) { d[o] = 1.0; }
};

// ---- BiCGStab passes -------------------------------------------------------------------
// v = J phat,  partial <rtilde, v>          (iterative.py: v = matvec(phat); rv = dot(rtilde, v))
struct OpAv {
    static constexpr int NRED = 1;
    static constexpr bool kSkipOnHalf = false;
    JacDev J;
    const double *src, *rtilde;
    double *dst;
    __device__ void prepare(const BicgScal *) {}
    __device__ void deep(const DeepTable &tab, const GridDev &G, int64_t o, double *acc) {
        double v = jac_row_deep(tab, J, src, o, G.pitch);
        dst[o] = v;
        acc[0] += rtilde[o] * v;
    }
    __device__ void band(const GridDev &G, int64_t bi, double *acc) {
        int64_t o = G.band_center[bi];
        double v = jac_row_band(G, J, src, bi);
        dst[o] = v;
        acc[0] += rtilde[o] * v;
    }
    __device__ void wall(int64_t o, double *acc) {
        double v = src[o];
        dst[o] = v;
        acc[0] += rtilde[o] * v;
    }
};

// t = J shat, partial <t, s>, <t, t>        (omega = dot(t, s) / dot(t, t))
struct OpAt {
    static constexpr int NRED = 2;
    static constexpr bool kSkipOnHalf = true;
    JacDev J;
    const double *src, *s;
    double *dst;
    __device__ void prepare(const BicgScal *) {}
    __device__ void one(int64_t o, double t, double *acc) {
        dst[o] = t;
        acc[0] += t * s[o];
        acc[1] += t * t;
    }
    __device__ void deep(const DeepTable &tab, const GridDev &G, int64_t o, double *acc) {
        one(o, jac_row_deep(tab, J, src, o, G.pitch), acc);
    }
    __device__ void band(const GridDev &G, int64_t bi, double *acc) {
        one(G.band_center[bi], jac_row_band(G, J, src, bi), acc);
    }
    __device__ void wall(int64_t o, double *acc) { one(o, src[o], acc); }
};

// elementwise passes over the whole state buffer (padding stays zero)
template <class Op>
__global__ void __launch_bounds__(PASS_THREADS)
k_elem(int64_t n, Op op, double *partials, const BicgScal *scal) {
    double acc[Op::NRED > 0 ? Op::NRED : 1] = {0.0};
    if (!(scal && scal->done)) {
        op.prepare(scal);
        for (int64_t i = blockIdx.x * (int64_t)PASS_THREADS + threadIdx.x; i < n;
             i += (int64_t)gridDim.x * PASS_THREADS)
            op.at(i, acc);
    }
    if (Op::NRED > 0) block_store_partials(acc, partials);
}

// r = b (x0 = 0), rtilde = r, x = 0; partial <b,b>
struct OpInit {
    static constexpr int NRED = 1;
    const double *b;
    double *x, *r, *rtilde;
    __device__ void prepare(const BicgScal *) {}
    __device__ void at(int64_t i, double *acc) {
        double v = b[i];
        r[i] = v;
        rtilde[i] = v;
        x[i] = 0.0;
        acc[0] += v * v;
    }
};

// p = r + beta*(p - omega*v) in SciPy's order (p -= omega*v; p *= beta; p += r); phat = dinv*p
struct OpP {
    static constexpr int NRED = 0;
    const double *r, *v, *dinv;
    double *p, *phat;
    double beta, omega;
    int first;
    __device__ void prepare(const BicgScal *s) {
        beta = s->beta;
        omega = s->omega;
        first = s->first;
    }
    __device__ void at(int64_t i, double *) {
        double pn;
        if (first) {
            pn = r[i];
        } else {
            pn = __dsub_rn(p[i], __dmul_rn(omega, v[i]));
            pn = __dmul_rn(pn, beta);
            pn = __dadd_rn(pn, r[i]);
        }
        p[i] = pn;
        phat[i] = dinv[i] * pn;
    }
};

// r -= alpha*v (this is s); shat = dinv*s; partial <s,s>
struct OpS {
    static constexpr int NRED = 1;
    const double *v, *dinv;
    double *r, *shat;
    double alpha;
    __device__ void prepare(const BicgScal *s) { alpha = s->alpha; }
    __device__ void at(int64_t i, double *acc) {
        double s = __dsub_rn(r[i], __dmul_rn(alpha, v[i]));
        r[i] = s;
        shat[i] = dinv[i] * s;
        acc[0] += s * s;
    }
};

// x += alpha*phat; x += omega*shat; r -= omega*t; partial <r,r>, <rtilde,r>
struct OpX {
    static constexpr int NRED = 2;
    const double *phat, *shat, *t, *rtilde;
    double *x, *r;
    double alpha, omega;
    int half;
    __device__ void prepare(const BicgScal *s) {
        alpha = s->alpha;
        omega = s->omega;
        half = s->half;
    }
    __device__ void at(int64_t i, double *acc) {
        double xn = __dadd_rn(x[i], __dmul_rn(alpha, phat[i]));
        if (half) {          // converged after the half step: x += alpha*phat only
            x[i] = xn;
            return;
        }
        xn = __dadd_rn(xn, __dmul_rn(omega, shat[i]));
        x[i] = xn;
        double rn = __dsub_rn(r[i], __dmul_rn(omega, t[i]));
        r[i] = rn;
        acc[0] += rn * rn;
        acc[1] += rtilde[i] * rn;
    }
};

// ---- scalar stages ------------------------------------------------------------------------
enum Stage { ST_INIT = 0, ST_V = 1, ST_S = 2, ST_T = 3, ST_X = 4 };

__device__ void top_of_loop(BicgScal *s) {
    // iterative.py: loop head of `for iteration in range(maxiter)`
    const double rhotol = 4.930380657631324e-32;      // eps**2 (iterative.py: rhotol)
    if (s->iter >= s->maxiter) {
        s->done = 1;
        s->info = s->maxiter;
        return;
    }
    if (sqrt(s->rr) < s->atol) {
        s->done = 1;
        s->info = 0;
        return;
    }
    double rho = s->rtr;
    if (fabs(rho) < rhotol) {
        s->done = 1;
        s->info = -10;
        return;
    }
    if (s->iter > 0) {
        if (fabs(s->omega) < rhotol) {
            s->done = 1;
            s->info = -11;
            return;
        }
        s->beta = (rho / s->rho_prev) * (s->alpha / s->omega);
    }
    s->rho = rho;
}

template <int NRED>
__global__ void k_stage(int stage, const double *partials, int nblocks, BicgScal *s, double rtol,
                        double atol) {
    __shared__ double sh[NRED][PASS_THREADS];
    if (s->done) return;
    if (stage == ST_T && s->half) return;
    double acc[NRED];
#pragma unroll
    for (int j = 0; j < NRED; ++j) acc[j] = 0.0;
    for (int b = threadIdx.x; b < nblocks; b += PASS_THREADS)
#pragma unroll
        for (int j = 0; j < NRED; ++j) acc[j] += partials[(int64_t)b * NRED + j];
#pragma unroll
    for (int j = 0; j < NRED; ++j) sh[j][threadIdx.x] = acc[j];
    __syncthreads();
    for (int o = PASS_THREADS / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o)
#pragma unroll
            for (int j = 0; j < NRED; ++j) sh[j][threadIdx.x] += sh[j][threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x != 0) return;
    double s0 = sh[0][0], s1 = NRED > 1 ? sh[NRED > 1 ? 1 : 0][0] : 0.0;
    switch (stage) {
        case ST_INIT:
            s->bnrm2 = sqrt(s0);
            s->atol = fmax(atol, rtol * s->bnrm2);
            s->rr = s0;
            s->rtr = s0;
            s->iter = 0;
            s->first = 1;
            s->half = 0;
            if (s->bnrm2 == 0.0) {
                s->done = 1;
                s->info = 0;
                return;
            }
            top_of_loop(s);
            break;
        case ST_V:
            s->rv = s0;
            if (s0 == 0.0) {
                s->done = 1;
                s->info = -11;
                return;
            }
            s->alpha = s->rho / s0;
            break;
        case ST_S:
            s->ss = s0;
            if (sqrt(s0) < s->atol) s->half = 1;
            break;
        case ST_T:
            s->ts = s0;
            s->tt = s1;
            s->omega = s0 / s1;
            break;
        case ST_X:
            if (s->half) {
                s->done = 1;
                s->info = 0;
                return;
            }
            s->rr = s0;
            s->rtr = s1;
            s->rho_prev = s->rho;
            s->iter += 1;
            s->first = 0;
            top_of_loop(s);
            break;
    }
}

static inline int elem_grid(int64_t n, int num_sms) {
    int64_t b = (n + PASS_THREADS * 4 - 1) / (PASS_THREADS * 4);
    int64_t cap = (int64_t)num_sms * 8;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

template <class Op>
static int launch_rows(pde_grid *g, const Op &op, double *partials, const BicgScal *scal,
                       cudaStream_t st, int *nblocks_out = nullptr) {
    GridDev G = grid_dev(g);
    int64_t nblk = g->rows + tail_blocks(g->n_band, g->nb);
    k_rows<Op><<<(unsigned)nblk, PASS_THREADS, 0, st>>>(G, g->tab, op, partials, scal);
    PDE_LAUNCH_CHECK();
    if (nblocks_out) *nblocks_out = (int)nblk;
    return 0;
}

template <class Op>
static int launch_elem(pde_grid *g, const Op &op, double *partials, const BicgScal *scal,
                       cudaStream_t st, int *nblocks_out = nullptr) {
    int64_t n = g->rows * g->pitch + g->nb;
    int nblk = elem_grid(n, g->num_sms);
    k_elem<Op><<<nblk, PASS_THREADS, 0, st>>>(n, op, partials, scal);
    PDE_LAUNCH_CHECK();
    if (nblocks_out) *nblocks_out = nblk;
    return 0;
}

// fused Newton update helper (solvers.py:189-199):
//   Unew = U + d ; acc: max|U - Unew| (as max), F.(J d), d.d
__global__ void k_newton_update(int64_t n, const double *U, const double *d, const double *F,
                                const double *Jd, double *Unew, double *out) {
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
    if ((threadIdx.x & 31) == 0) {
        atomic_max_nonneg(&out[0], mx);
        atomicAdd(&out[1], a);
        atomicAdd(&out[2], b);
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
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t n = pde_grid_state_len(g);
    const int64_t n_unknowns = g->nx_local * g->ny + g->nb;
    if (maxiter <= 0) maxiter = 10 * n_unknowns;     // iterative.py: maxiter = n*10
    if (check_every < 1) check_every = 32;

    double *r = d_work, *rtilde = r + n, *p = rtilde + n, *v = p + n, *phat = v + n,
           *shat = phat + n, *t = shat + n, *dinv = t + n;
    // scalars + partial sums live behind the 8 vectors?  No: separate small allocations.
    int64_t max_blocks = g->rows + tail_blocks(g->n_band, g->nb) + (int64_t)g->num_sms * 8;
    double *partials = nullptr;
    BicgScal *scal = nullptr;
    PDE_CHECK(cudaMallocAsync((void **)&partials, sizeof(double) * MAX_RED * max_blocks, st));
    PDE_CHECK(cudaMallocAsync((void **)&scal, sizeof(BicgScal), st));
    PDE_CHECK(cudaMemsetAsync(scal, 0, sizeof(BicgScal), st));
    struct Cleanup {
        double *p;
        BicgScal *s;
        cudaStream_t st;
        ~Cleanup() {
            cudaFreeAsync(p, st);
            cudaFreeAsync(s, st);
        }
    } cleanup{partials, scal, st};
    {
        BicgScal h{};
        h.maxiter = maxiter;
        PDE_CHECK(cudaMemcpyAsync(scal, &h, sizeof h, cudaMemcpyHostToDevice, st));
    }
    JacDev Jd = jac_dev(J);
    int rc, nb;
    // dinv = 1/diag(J)   (solvers.py:179); zero on padding
    PDE_CHECK(cudaMemsetAsync(dinv, 0, sizeof(double) * n, st));
    if ((rc = launch_rows(g, OpDiag{Jd, dinv, 1}, nullptr, nullptr, st))) return rc;
    PDE_CHECK(cudaMemsetAsync(v, 0, sizeof(double) * n, st));
    PDE_CHECK(cudaMemsetAsync(t, 0, sizeof(double) * n, st));
    PDE_CHECK(cudaMemsetAsync(p, 0, sizeof(double) * n, st));
    if ((rc = launch_elem(g, OpInit{d_b, d_x, r, rtilde}, partials, nullptr, st, &nb))) return rc;
    k_stage<1><<<1, PASS_THREADS, 0, st>>>(ST_INIT, partials, nb, scal, rtol, atol);
    PDE_LAUNCH_CHECK();

    struct {
        long long iter, info, maxiter;
        int done, half, first, pad;
    } h_tail;
    const size_t tail_off = offsetof(BicgScal, iter);
    int64_t launched = 0;
    while (true) {
        for (int64_t k = 0; k < check_every; ++k, ++launched) {
            if ((rc = launch_elem(g, OpP{r, v, dinv, p, phat, 0, 0, 0}, nullptr, scal, st))) return rc;
            if ((rc = launch_rows(g, OpAv{Jd, phat, rtilde, v}, partials, scal, st, &nb))) return rc;
            k_stage<1><<<1, PASS_THREADS, 0, st>>>(ST_V, partials, nb, scal, rtol, atol);
            PDE_LAUNCH_CHECK();
            if ((rc = launch_elem(g, OpS{v, dinv, r, shat, 0}, partials, scal, st, &nb))) return rc;
            k_stage<1><<<1, PASS_THREADS, 0, st>>>(ST_S, partials, nb, scal, rtol, atol);
            PDE_LAUNCH_CHECK();
            if ((rc = launch_rows(g, OpAt{Jd, shat, r, t}, partials, scal, st, &nb))) return rc;
            k_stage<2><<<1, PASS_THREADS, 0, st>>>(ST_T, partials, nb, scal, rtol, atol);
            PDE_LAUNCH_CHECK();
            if ((rc = launch_elem(g, OpX{phat, shat, t, rtilde, d_x, r, 0, 0, 0}, partials, scal, st,
                                  &nb)))
                return rc;
            k_stage<2><<<1, PASS_THREADS, 0, st>>>(ST_X, partials, nb, scal, rtol, atol);
            PDE_LAUNCH_CHECK();
        }
        PDE_CHECK(cudaMemcpyAsync(&h_tail, reinterpret_cast<char *>(scal) + tail_off,
                                  sizeof h_tail, cudaMemcpyDeviceToHost, st));
        PDE_CHECK(cudaStreamSynchronize(st));
        if (h_tail.done) break;
        if (launched > maxiter + check_every) {
            return pde::fail("bicgstab: device loop failed to terminate", __FILE__, __LINE__);
        }
    }
    h_info[0] = h_tail.info;
    h_info[1] = h_tail.iter;
    return 0;
}

int pde_vec_newton_update(pde_grid *g, const double *d_U, const double *d_d, const double *d_F,
                          const double *d_Jd, double *d_Unew, double *h_out, void *stream) {
    PDE_REQUIRE(g && d_U && d_d && d_F && d_Jd && d_Unew && h_out, "null argument");
    DeviceGuard guard(g->device);
    cudaStream_t st = (cudaStream_t)stream;
    int64_t n = pde_grid_state_len(g);
    double *out = g->d_red + 4;
    PDE_CHECK(cudaMemsetAsync(out, 0, 3 * sizeof(double), st));
    k_newton_update<<<elem_grid(n, g->num_sms), 256, 0, st>>>(n, d_U, d_d, d_F, d_Jd, d_Unew, out);
    PDE_LAUNCH_CHECK();
    PDE_CHECK(cudaMemcpyAsync(h_out, out, 3 * sizeof(double), cudaMemcpyDeviceToHost, st));
    PDE_CHECK(cudaStreamSynchronize(st));
    return 0;
}

}  // extern "C"
