// Backend-independent part of the Jacobi-preconditioned BiCGStab: device-resident
// scalars, the fused vector passes, the one-block scalar stages and the driver loop.
//
// Follows SciPy 1.18.1 scipy/sparse/linalg/_isolve/iterative.py::bicgstab statement by
// statement (the call the reference makes at solvers.py:179-181): x0 = 0,
// atol_eff = max(atol, rtol*||b||), convergence tested at the loop head and after the
// half step, breakdown exits -10 (|rho| < eps^2) and -11 (|omega| < eps^2 or
// <rtilde,v> == 0), maxiter = 10 n, update order p -= omega*v; p *= beta; p += r.
// A backend supplies the two operator passes (v = J phat with <rtilde,v>; t = J shat with
// <t,s>, <t,t>) and 1/diag(J): the regular-grid policy Jacobian (bicgstab.cu) and the
// point-cloud ELL Jacobian (ell.cu).
#pragma once
#include <cstddef>
#include <cstdlib>

#include "common.cuh"

namespace pde {

constexpr int PASS_THREADS = 256;
constexpr int MAX_RED = 3;

struct BicgScal {
    double bnrm2, atol, rho, rho_prev, alpha, omega, beta, rv, ss, ts, tt, rr, rtr;
    double sums[MAX_RED];
    long long iter, info, maxiter;
    int done, half, first;
    unsigned counter;     // blocks of the pass in flight that have stored their partial sums
};

// ---- dot-product accumulators --------------------------------------------------------------
// Plain double sums in a fixed order (per thread over its grid-stride sequence, shuffle tree per
// warp, warps in order per block, blocks in a fixed order by the last block): deterministic for
// a given launch geometry.  BiCGStab's iteration count on the policy Jacobians is chaotic in
// the rounding of these sums: the certifying solve of the 4097^2 lattice took between 320 and
// 1 905 iterations depending only on how the partial sums were grouped (tools/certify_counts.py).
// Accumulating them in two- and three-fold working precision (Dot2 / Dot3, round 2) made the
// count independent of most regroupings, but not smaller on average (360-770), 15 % slower per
// iteration, and one 65^2 solve then stagnated for 10 n iterations -- the plain sums stay.
constexpr int DDW = 1;            // doubles per partial sum
struct DD {
    double hi;
};
__device__ __forceinline__ void dd_mac(DD &s, double a, double b) { s.hi += a * b; }

template <int NRED, int NT = PASS_THREADS>
__device__ __forceinline__ void block_store_partials(DD (&acc)[NRED], double *mine) {
    // `mine`: this block's NRED slots of the pass's partial sums
    __shared__ double sh[NRED][NT / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < NRED; ++j) {
        double v = warp_sum(acc[j].hi);
        if (lane == 0) sh[j][wid] = v;
    }
    __syncthreads();
    if (threadIdx.x < NRED) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < NT / 32; ++w) s += sh[threadIdx.x][w];
        mine[threadIdx.x] = s;
        __threadfence();                  // visible device-wide before this block is counted
    }
}

// ---- scalar stages ------------------------------------------------------------------------
enum Stage { ST_NONE = -1, ST_INIT = 0, ST_V = 1, ST_S = 2, ST_T = 3, ST_X = 4,
             ST_ST = 5 /* fused s / t pass: <s,s>, <t,s>, <t,t> at once */ };

__device__ inline void top_of_loop(BicgScal *s) {
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

// the scalar update that follows the reduction of one pass (one thread)
__device__ inline void stage_update(int stage, BicgScal *s, double s0, double s1, double rtol,
                                    double atol, double s2 = 0.0) {
    switch (stage) {
        case ST_ST:                       // ST_S, then ST_T unless the half step converged
            s->ss = s0;
            if (sqrt(s0) < s->atol) {
                s->half = 1;
            } else {
                s->ts = s1;
                s->tt = s2;
                s->omega = s1 / s2;
            }
            break;
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

// fixed-order sum of the per-block partials by one block of PASS_THREADS threads: thread t adds
// blocks t, t + PASS_THREADS, ..., then a binary tree -- the same bits whoever runs it
template <int NRED>
__device__ __forceinline__ void reduce_partials(const double *partials, int nblocks, double &s0,
                                                double &s1, double &s2) {
    __shared__ double shr[NRED][PASS_THREADS];
    double acc[NRED];
#pragma unroll
    for (int j = 0; j < NRED; ++j) acc[j] = 0.0;
    for (int b = threadIdx.x; b < nblocks; b += PASS_THREADS)
#pragma unroll
        for (int j = 0; j < NRED; ++j) acc[j] += __ldcg(partials + (int64_t)b * NRED + j);   // other blocks' stores: L2
#pragma unroll
    for (int j = 0; j < NRED; ++j) shr[j][threadIdx.x] = acc[j];
    __syncthreads();
    for (int o = PASS_THREADS / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o)
#pragma unroll
            for (int j = 0; j < NRED; ++j) shr[j][threadIdx.x] += shr[j][threadIdx.x + o];
        __syncthreads();
    }
    s0 = shr[0][0];
    s1 = NRED > 1 ? shr[NRED > 1 ? 1 : 0][0] : 0.0;
    s2 = NRED > 2 ? shr[NRED > 2 ? 2 : 0][0] : 0.0;
}

// What the LAST block of a pass does after storing its partial sums: reduce all `total`
// partials (a pass may consist of two launches on one stream; `base` points at the first
// partial of the pass, and only the second launch carries a stage) and update the scalars --
// this replaces a one-block launch after every pass.
struct StageArgs {
    int stage;            // ST_NONE: no in-kernel stage (multi-GPU: an all-reduce sits in between)
    int total;            // partial sums of the whole pass
    double rtol, atol;
};

template <int NRED>
__device__ __forceinline__ void pass_finish(DD (&acc)[NRED], double *mine, const double *base,
                                            BicgScal *scal, const StageArgs &sa) {
    block_store_partials(acc, mine);
    if (sa.stage == ST_NONE || scal == nullptr) return;
    __shared__ bool last;
    __syncthreads();                       // the partial sums of this block are stored
    if (threadIdx.x == 0) {
        __threadfence();
        last = atomicAdd(&scal->counter, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    double s0, s1, s2;
    reduce_partials<NRED>(base, sa.total, s0, s1, s2);
    if (threadIdx.x == 0) {
        scal->counter = 0u;
        if (!scal->done && !(sa.stage == ST_T && scal->half))
            stage_update(sa.stage, scal, s0, s1, sa.rtol, sa.atol, s2);
    }
}

// One-block stage kernel, kept for the slab-parallel path (an all-reduce of the local sums sits
// between the two phases):
//   phase 1: reduce only, local sums -> sums_io        phase 2: update the scalars from sums_io
template <int NRED>
__global__ void k_stage(int stage, const double *partials, int nblocks, BicgScal *s, double rtol,
                        double atol, double *sums_io, int phase) {
    if (s->done) return;
    if (stage == ST_T && s->half) return;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    if (phase == 1) {
        reduce_partials<NRED>(partials, nblocks, s0, s1, s2);
        if (threadIdx.x == 0) {
            sums_io[0] = s0;
            if (NRED > 1) sums_io[1] = s1;
        }
        return;
    }
    if (threadIdx.x != 0) return;
    s0 = sums_io[0];
    if (NRED > 1) s1 = sums_io[1];
    stage_update(stage, s, s0, s1, rtol, atol);
}

// elementwise passes over a whole vector (padding entries stay zero)
template <class Op>
__global__ void __launch_bounds__(PASS_THREADS)
k_elem(int64_t n, Op op, double *partials, BicgScal *scal, const StageArgs sa) {
    DD acc[Op::NRED > 0 ? Op::NRED : 1] = {};
    if (!(scal && scal->done)) {
        op.prepare(scal);
        for (int64_t i = blockIdx.x * (int64_t)PASS_THREADS + threadIdx.x; i < n;
             i += (int64_t)gridDim.x * PASS_THREADS)
            op.at(i, acc);
    }
    if constexpr (Op::NRED > 0)
        pass_finish(acc, partials + (int64_t)blockIdx.x * DDW * Op::NRED, partials, scal, sa);
}

// r = b (x0 = 0), rtilde = r, x = 0; partial <b,b>
struct OpInit {
    static constexpr int NRED = 1;
    const double *b;
    double *x, *r, *rtilde;
    __device__ void prepare(const BicgScal *) {}
    __device__ void at(int64_t i, DD *acc) {
        double v = b[i];
        r[i] = v;
        rtilde[i] = v;
        x[i] = 0.0;
        dd_mac(acc[0], v, v);
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
    __device__ void at(int64_t i, DD *) {
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
    __device__ void at(int64_t i, DD *acc) {
        double s = __dsub_rn(r[i], __dmul_rn(alpha, v[i]));
        r[i] = s;
        shat[i] = dinv[i] * s;
        dd_mac(acc[0], s, s);
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
    __device__ void at(int64_t i, DD *acc) {
        double xn = __dadd_rn(x[i], __dmul_rn(alpha, phat[i]));
        if (half) {          // converged after the half step: x += alpha*phat only
            x[i] = xn;
            return;
        }
        xn = __dadd_rn(xn, __dmul_rn(omega, shat[i]));
        x[i] = xn;
        double rn = __dsub_rn(r[i], __dmul_rn(omega, t[i]));
        r[i] = rn;
        dd_mac(acc[0], rn, rn);
        dd_mac(acc[1], rtilde[i], rn);
    }
};

// OpX after the fused s / t pass (the grid backend): s is kept instead of shat, and shat =
// dinv*s -- the same product OpS stores -- is formed on the fly
struct OpXs {
    static constexpr int NRED = 2;
    const double *phat, *sv, *dinv, *t, *rtilde;
    double *x, *r;
    double alpha, omega;
    int half;
    __device__ void prepare(const BicgScal *s) {
        alpha = s->alpha;
        omega = s->omega;
        half = s->half;
    }
    __device__ void at(int64_t i, DD *acc) {
        double xn = __dadd_rn(x[i], __dmul_rn(alpha, phat[i]));
        if (half) {
            x[i] = xn;
            return;
        }
        const double si = sv[i];
        xn = __dadd_rn(xn, __dmul_rn(omega, __dmul_rn(dinv[i], si)));
        x[i] = xn;
        double rn = __dsub_rn(si, __dmul_rn(omega, t[i]));
        r[i] = rn;
        dd_mac(acc[0], rn, rn);
        dd_mac(acc[1], rtilde[i], rn);
    }
};

static inline int elem_grid(int64_t n, int num_sms) {
    int64_t b = (n + PASS_THREADS * 4 - 1) / (PASS_THREADS * 4);
    int64_t cap = (int64_t)num_sms * 8;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

// stage: the scalar stage the last block of the pass performs (ST_NONE: none)
template <class Op>
static int launch_elem(int64_t n, int num_sms, const Op &op, double *partials, BicgScal *scal,
                       cudaStream_t st, int *nblocks_out = nullptr, int stage = ST_NONE,
                       double rtol = 0.0, double atol = 0.0) {
    // one wave: no more blocks than are resident at once (the 40-register OpX fits 6 blocks per
    // SM; 8 per SM left a second, mostly empty wave -- 57 % achieved occupancy in round 1)
    static int resident = 0;
    if (resident == 0) {
        int occ = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_elem<Op>, PASS_THREADS, 0) != cudaSuccess ||
            occ < 1)
            occ = 4;
        cudaGetLastError();
        resident = occ > 8 ? 8 : occ;
        if (const char *e = getenv("PDE_ELEM_BLOCKS")) resident = atoi(e) > 0 ? atoi(e) : resident;
    }
    int64_t b = (n + PASS_THREADS * 4 - 1) / (PASS_THREADS * 4);
    int64_t cap = (int64_t)num_sms * resident;
    int nblk = (int)(b < 1 ? 1 : (b > cap ? cap : b));
    k_elem<Op><<<nblk, PASS_THREADS, 0, st>>>(n, op, partials, scal, StageArgs{stage, nblk, rtol, atol});
    PDE_LAUNCH_CHECK();
    if (nblocks_out) *nblocks_out = nblk;
    return 0;
}

// fused Newton update (solvers.py:189-199): Unew = U + d; h_out = [max|U - Unew|, F.(J d), d.d]
// (defined in bicgstab.cu; synchronises the stream)
int newton_update_launch(int64_t n, int num_sms, const double *U, const double *d, const double *F,
                         const double *Jd, double *Unew, double *h_out, cudaStream_t st);

// Backend concept:
//   int64_t n() / unknowns() / max_row_blocks(); int num_sms();
//   int dinv(double *out, cudaStream_t);                       out = 1/diag(J), 0 on padding
//   int Av(src, rtilde, dst, partials, scal, st, &nblocks, sa);   dst = J src, partial <rtilde,dst>
//   int At(src, s, dst, partials, scal, st, &nblocks, sa);        dst = J src, partial <dst,s>,<dst,dst>
//   (sa: StageArgs; the last block of the pass's LAST launch performs stage sa.stage)
//   bool fused_st();  int SAt(r, v, dinv, s_out, t_out, partials, scal, st, &nblocks, sa):
//       s = r - alpha*v, t = J (dinv*s), partial <s,s>, <t,s>, <t,t> (optional, grids only)
template <class Backend>
static int bicgstab_run(Backend &B, const double *d_b, double *d_x, double rtol, double atol,
                        int64_t maxiter, int64_t check_every, double *d_work, int64_t *h_info,
                        cudaStream_t st) {
    const int64_t n = B.n();
    const int sms = B.num_sms();
    if (maxiter <= 0) maxiter = 10 * B.unknowns();     // iterative.py: maxiter = n*10
    if (check_every < 1) check_every = 32;
    // TMA sources must be 16-byte aligned whatever the parity of n: they sit at even multiples
    // of n.  Plain passes: phat, shat.  Fused s / t pass (grids): r, v, phat, dinv; slot 7 keeps s.
    const bool fused = B.fused_st();
    double *r = d_work, *rtilde = r + n;
    double *p, *v, *phat, *t, *shat, *dinv;
    if (fused) {
        v = d_work + 2 * n; p = d_work + 3 * n; phat = d_work + 4 * n; t = d_work + 5 * n;
        dinv = d_work + 6 * n; shat = d_work + 7 * n;      // shat: holds s
    } else {
        p = d_work + 2 * n; v = d_work + 3 * n; phat = d_work + 4 * n; t = d_work + 5 * n;
        shat = d_work + 6 * n; dinv = d_work + 7 * n;
    }
    int64_t max_blocks = B.max_row_blocks() + (int64_t)sms * 8;
    double *partials = nullptr;
    BicgScal *scal = nullptr;
    PDE_CHECK(cudaMallocAsync((void **)&partials, sizeof(double) * DDW * MAX_RED * max_blocks, st));
    PDE_CHECK(cudaMallocAsync((void **)&scal, sizeof(BicgScal), st));
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
    int rc, nb;
    // dinv = 1/diag(J)   (solvers.py:179); zero on padding
    PDE_CHECK(cudaMemsetAsync(dinv, 0, sizeof(double) * n, st));
    if ((rc = B.dinv(dinv, st))) return rc;
    PDE_CHECK(cudaMemsetAsync(v, 0, sizeof(double) * n, st));
    PDE_CHECK(cudaMemsetAsync(t, 0, sizeof(double) * n, st));
    PDE_CHECK(cudaMemsetAsync(p, 0, sizeof(double) * n, st));
    // the fused s / t pass writes s at real entries only: its padding entries must read as zero
    if (fused) PDE_CHECK(cudaMemsetAsync(shat, 0, sizeof(double) * n, st));
    if ((rc = launch_elem(n, sms, OpInit{d_b, d_x, r, rtilde}, partials, scal, st, &nb, ST_INIT, rtol,
                          atol)))
        return rc;

    struct {
        long long iter, info, maxiter;
        int done, half, first;
        unsigned counter;
    } h_tail;
    const size_t tail_off = offsetof(BicgScal, iter);
    // one BiCGStab iteration = 5 passes (7 launches on grids: the two operator passes are a
    // tiled launch + a tail launch) whose arguments never change within a solve; the last block
    // of every pass reduces the partial sums and updates the scalars
    auto iteration = [&]() -> int {
        int rc2;
        if ((rc2 = launch_elem(n, sms, OpP{r, v, dinv, p, phat, 0, 0, 0}, nullptr, scal, st))) return rc2;
        if ((rc2 = B.Av(phat, rtilde, v, partials, scal, st, &nb, StageArgs{ST_V, 0, rtol, atol}))) return rc2;
        if (fused) {
            // s = r - alpha*v (-> slot 7), t = J (dinv*s), <s,s>, <t,s>, <t,t> in ONE pass over
            // tiles of r, v, dinv: 5 vector passes instead of the 8 of OpS + OpAt
            if ((rc2 = B.SAt(r, v, dinv, shat, t, partials, scal, st, &nb, StageArgs{ST_ST, 0, rtol, atol})))
                return rc2;
            return launch_elem(n, sms, OpXs{phat, shat, dinv, t, rtilde, d_x, r, 0, 0, 0}, partials, scal,
                               st, &nb, ST_X, rtol, atol);
        }
        if ((rc2 = launch_elem(n, sms, OpS{v, dinv, r, shat, 0}, partials, scal, st, &nb, ST_S, rtol, atol)))
            return rc2;
        if ((rc2 = B.At(shat, r, t, partials, scal, st, &nb, StageArgs{ST_T, 0, rtol, atol}))) return rc2;
        if ((rc2 = launch_elem(n, sms, OpX{phat, shat, t, rtilde, d_x, r, 0, 0, 0}, partials, scal, st,
                               &nb, ST_X, rtol, atol)))
            return rc2;
        return 0;
    };
    // Small and medium grids are launch bound (9 launches of a few microseconds each per
    // iteration): capture ONE iteration into a CUDA graph per solve (instantiating a graph of
    // a whole batch cost ~30 ms, measured) and replay it; the device-side `done` flag turns
    // surplus replays into no-ops.  Large grids gain nothing and skip the capture.
    cudaGraphExec_t exec = nullptr;
    int launches_per_iteration = 0;
    if (B.unknowns() < (int64_t)4 << 20 &&
        cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
        const int64_t l0 = pde::launch_count();
        int rcc = iteration();
        launches_per_iteration = (int)(pde::launch_count() - l0);
        cudaGraph_t graph = nullptr;
        cudaError_t e = cudaStreamEndCapture(st, &graph);
        pde::count_launch(-launches_per_iteration);            // captured, not executed
        if (rcc == 0 && e == cudaSuccess && graph) {
            if (cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess) exec = nullptr;
        }
        if (graph) cudaGraphDestroy(graph);
        if (!exec) cudaGetLastError();                         // fall back to plain launches
    } else {
        cudaGetLastError();
    }
    struct ExecGuard {
        cudaGraphExec_t e;
        ~ExecGuard() {
            if (e) cudaGraphExecDestroy(e);
        }
    } exec_guard{exec};
    int64_t launched = 0;
    while (true) {
        if (exec) {
            for (int64_t k = 0; k < check_every; ++k, ++launched) PDE_CHECK(cudaGraphLaunch(exec, st));
            pde::count_launch((int)(launches_per_iteration * check_every));
        } else {
            for (int64_t k = 0; k < check_every; ++k, ++launched)
                if ((rc = iteration())) return rc;
        }
        PDE_CHECK(cudaMemcpyAsync(&h_tail, reinterpret_cast<char *>(scal) + tail_off,
                                  sizeof h_tail, cudaMemcpyDeviceToHost, st));
        PDE_CHECK(cudaStreamSynchronize(st));
        if (h_tail.done) break;
        if (launched > maxiter + check_every)
            return pde::fail("bicgstab: device loop failed to terminate", __FILE__, __LINE__);
    }
    h_info[0] = h_tail.info;
    h_info[1] = h_tail.iter;
    return 0;
}

}  // namespace pde
