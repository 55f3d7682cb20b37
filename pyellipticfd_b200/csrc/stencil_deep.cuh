// K1: structured wide-stencil min/max kernel for deep-interior lattice points.
//
// Replaces the gather + lexsort of ddg.d2eigs (reference ddg.py:274-292) for
// points whose stencil is translation invariant.  One CTA owns a TX x TY tile of
// points; the tile plus a halo of H cells is staged in shared memory by ONE
// 2-D TMA load (cp.async.bulk.tensor, zero fill outside the lattice), double
// buffered over a persistent tile loop so the next tile's load overlaps the
// current tile's arithmetic.  Lanes run along y (the contiguous axis) so that
// shared-memory reads are conflict free and global stores coalesce; each
// thread owns P consecutive rows of one column and walks the direction table.
//
// Arithmetic: mode 0 reproduces ddg.py:281 operation by operation
// (__dsub_rn/__dmul_rn/__dadd_rn: no FMA contraction, no re-association), so
// values and the selected directions are bit-identical to the reference on
// grids whose weights are position independent.  Tie rule of ddg.py:283-292:
// `<=` keeps the LAST minimal pair, `>` keeps the FIRST maximal pair.
#pragma once
#include "common.cuh"
#include "stencil_ct.cuh"

namespace pde {

constexpr int K1_TY = 128;      // tile columns (y, contiguous)
constexpr int K1_THREADS = 256;

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *tmap, int c_inner,
                                            int c_outer, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(tmap)), "r"(c_inner), "r"(c_outer), "r"(smem_u32(bar))
        : "memory");
}

// Geometry of one launch (all in local buffer coordinates).
struct DeepGeom {
    int64_t pitch;        // doubles per row of every state-layout array
    int row0;             // first buffer row of the deep region handled here
    int col0;             // first column of the deep region
    int nrows, ncols;     // extent of the deep region
    int tiles_x, tiles_y;
    // optional second row range (same columns): tile rows [split, tiles_x) belong to it
    int split, row0b, nrowsb;
};

struct BlockMax2 {   // two running maxima per thread, reduced once per CTA at the end
    double a, b;
};

// ---- epilogues -----------------------------------------------------------------
// Each epilogue sees one evaluated point: state offset `o`, centre value, the two
// extremes and their table rows.
template <bool MIN, bool MAX, bool POL>
struct EpiEigs {    // ddg.d2eigs values + policy (ddg.py:292); d2min / d2max skip the unused half
    double *lmin, *lmax;
    uint16_t *pmin, *pmax;
    static constexpr bool kMin = MIN, kMax = MAX, kRed = false, kPol = POL, kPre = false;
    __device__ __forceinline__ double pre(int64_t) const { return 0.0; }
    __device__ __forceinline__ void operator()(int64_t o, double, double mn, double mx, int kmn,
                                               int kmx, double, BlockMax2 &) const {
        if (lmin) lmin[o] = mn;
        if (lmax) lmax[o] = mx;
        if (pmin) pmin[o] = static_cast<uint16_t>(kmn);
        if (pmax) pmax[o] = static_cast<uint16_t>(kmx);
    }
};

template <bool POL>
struct EpiCeResidual {   // convex_envelope.operator (convex_envelope.py:44-46)
    const double *g;
    double *F;
    uint16_t *policy;
    static constexpr bool kMin = true, kMax = false, kRed = true, kPol = POL, kPre = true;
    // the obstacle value is fetched before the direction loop so that its DRAM latency
    // is hidden behind the arithmetic instead of stalling the tile's epilogue
    __device__ __forceinline__ double pre(int64_t o) const { return g[o]; }
    __device__ __forceinline__ void operator()(int64_t o, double uc, double mn, double, int kmn,
                                               int, double go, BlockMax2 &red) const {
        double f = __dsub_rn(uc, go);
        double nl = -mn;
        bool active = nl > f;            // strict '>' (convex_envelope.py:45)
        if (active) f = nl;
        F[o] = f;
        if (policy) policy[o] = active ? static_cast<uint16_t>(kmn) : (uint16_t)PDE_POLICY_IDENTITY;
        red.a = fmax(red.a, fabs(f));
    }
};

struct EpiCeEuler {   // one explicit Euler step (solvers.py:65-74 on convex_envelope.py:44-46)
    const double *g;
    double *Unew;
    double dt;
    static constexpr bool kMin = true, kMax = false, kRed = true, kPol = false, kPre = true;
    __device__ __forceinline__ double pre(int64_t o) const { return g[o]; }
    __device__ __forceinline__ void operator()(int64_t o, double uc, double mn, double, int, int,
                                               double go, BlockMax2 &red) const {
        double f = __dsub_rn(uc, go);
        double nl = -mn;
        if (nl > f) f = nl;
        double un = __dsub_rn(uc, __dmul_rn(dt, f));
        Unew[o] = un;
        red.a = fmax(red.a, fabs(f));
        red.b = fmax(red.b, fabs(__dsub_rn(uc, un)));
    }
};

template <bool POL>
struct EpiPucciResidual {   // pucci.py:48, :127-141
    const double *f;
    double *F;
    uint16_t *pmin, *pmax;
    double A0, A1;
    const double *dA0, *dA1;
    static constexpr bool kMin = true, kMax = true, kRed = true, kPol = POL, kPre = true;
    __device__ __forceinline__ double pre(int64_t o) const { return f[o]; }
    __device__ __forceinline__ void operator()(int64_t o, double, double mn, double mx, int kmn,
                                               int kmx, double fo, BlockMax2 &red) const {
        double a0 = dA0 ? dA0[o] : A0;
        double a1 = dA1 ? dA1[o] : A1;
        double val = __dadd_rn(__dmul_rn(a0, mx), __dmul_rn(a1, mn));   // A[0]*lmax + A[1]*lmin
        double r = __dsub_rn(val, fo);
        F[o] = r;
        if (pmin) pmin[o] = static_cast<uint16_t>(kmn);
        if (pmax) pmax[o] = static_cast<uint16_t>(kmx);
        red.a = fmax(red.a, fabs(r));
    }
};

// ---- direction-table evaluators -----------------------------------------------------
// Run-time table (any grid whose weights are position independent, e.g. hx != hy): the
// offsets come from the kernel-parameter constant bank.
struct EvalRuntime {
    template <int P, int TYH, int MODE, bool WMIN, bool WMAX, bool WPOL, int ICMP>
    static __device__ __forceinline__ void run(const double *__restrict__ ctr, const DeepTable &tab,
                                               const double (&uc)[P], double (&mn)[P],
                                               double (&mx)[P], int (&kmn)[P], int (&kmx)[P]) {
        const int D = tab.D;
#pragma unroll 2
        for (int k = 0; k < D; ++k) {
            const double *q0 = ctr + tab.a0[k] * TYH + tab.b0[k];
            const double *q1 = ctr + tab.a1[k] * TYH + tab.b1[k];
            const double w_0 = tab.w_0[k], w_1 = tab.w_1[k], w0 = tab.w0[k];
#pragma unroll
            for (int p = 0; p < P; ++p) {
                double u0 = q0[p * TYH], u1 = q1[p * TYH];
                double d;
                if (MODE == 0)
                    d = pair_value_exact(uc[p], u0, u1, w_0, w_1, w0);
                else
                    d = (fma(w_1, u1 - uc[p], w_0 * (u0 - uc[p]))) * w0;   // fewer fp64 ops, not bit-exact
                if (WMIN) {
                    bool le = d <= mn[p];
                    mn[p] = le ? d : mn[p];
                    if (WPOL) kmn[p] = le ? k : kmn[p];
                }
                if (WMAX) {
                    bool gt = d > mx[p];
                    mx[p] = gt ? d : mx[p];
                    if (WPOL) kmx[p] = gt ? k : kmx[p];
                }
            }
        }
    }
};

// ---- the kernel ------------------------------------------------------------------
template <class Eval, int H, int P, int MODE, class Epi, int MINB = 2, int ICMP = 0>
__global__ void __launch_bounds__(K1_THREADS, MINB)
k_deep(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ DeepTable tab,
       const DeepGeom geo, const Epi epi, double *red_out,
       const unsigned long long *done_flag, unsigned long long *sched) {
    constexpr int TX = 2 * P;                 // 256 threads = 128 columns x 2 row groups
    // +2 columns: the TMA box must start on a 16-byte boundary of the row, so when the
    // halo origin is odd the box starts one column early and `ysh` re-centres the reads
    constexpr int TYH = K1_TY + 2 * H + 2;
    constexpr int TXH = TX + 2 * H;
    constexpr int TILE_BYTES = TXH * TYH * 8;
    constexpr int TILE_STRIDE = (TILE_BYTES + 127) / 128 * 128;

    if (done_flag && *done_flag) return;      // device-side early exit of solver loops

    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *buf0 = reinterpret_cast<double *>(smem_raw);
    double *buf1 = reinterpret_cast<double *>(smem_raw + TILE_STRIDE);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + 2 * TILE_STRIDE);

    const int tid = threadIdx.x;
    const int c = tid & (K1_TY - 1);
    const int grp = tid >> 7;
    const int ntiles = geo.tiles_x * geo.tiles_y;
    const int ysh = (geo.col0 - H) & 1;

    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    auto issue = [&](int tile, int slot) {
        int tx = tile / geo.tiles_y, ty = tile - tx * geo.tiles_y;
        int x0 = (tx < geo.split ? geo.row0 + tx * TX : geo.row0b + (tx - geo.split) * TX) - H;
        int y0 = geo.col0 + ty * K1_TY - H - ysh;
        uint64_t *bar = &bars[slot];
        mbar_expect_tx(bar, TILE_BYTES);
        tma_load_2d(slot ? buf1 : buf0, &tmap, y0, x0, bar);
    };

    // Tile schedule.  sched == nullptr: static round robin (tile += gridDim.x).  Otherwise the
    // CTAs draw their next tile from a device counter (sched[0]; sched[1] counts finished CTAs
    // and the last one resets both): CTAs that start late -- e.g. behind the NCCL kernel of an
    // overlapped halo exchange, which cannot share an SM with two 125-register CTAs -- then simply
    // take fewer tiles instead of finishing a full static share late.
    __shared__ int s_next[2];
    BlockMax2 red{0.0, 0.0};
    int tile = blockIdx.x;
    if (tile < ntiles && tid == 0) issue(tile, 0);

    for (int it = 0; tile < ntiles; ++it) {
        const int slot = it & 1;
        int next;
        if (sched) {
            if (tid == 0) {
                next = (int)gridDim.x + (int)atomicAdd(sched, 1ULL);
                s_next[slot] = next;
                if (next < ntiles) issue(next, slot ^ 1);
            }
        } else {
            next = tile + gridDim.x;
            if (next < ntiles && tid == 0) issue(next, slot ^ 1);
        }
        const double *sm = slot ? buf1 : buf0;
        const int tx = tile / geo.tiles_y, ty = tile - tx * geo.tiles_y;
        const bool second = tx >= geo.split;
        const int rbase = second ? geo.row0b : geo.row0;  // first buffer row of this range
        const int rcount = second ? geo.nrowsb : geo.nrows;
        const int lx0 = (second ? tx - geo.split : tx) * TX + grp * P;   // row inside the range
        const int ly = ty * K1_TY + c;
        double pre[Epi::kPre ? P : 1];
        if (Epi::kPre) {      // epilogue operands (obstacle / forcing): loads issued before the wait
#pragma unroll
            for (int p = 0; p < P; ++p) {
                bool in = ly < geo.ncols && lx0 + p < rcount;
                pre[p] = in ? epi.pre((int64_t)(rbase + lx0 + p) * geo.pitch + (geo.col0 + ly)) : 0.0;
            }
        }
        mbar_wait(&bars[slot], (it >> 1) & 1);

        const double *ctr = sm + (H + grp * P) * TYH + (H + c + ysh);

        double uc[P], mn[P], mx[P];
        int kmn[P], kmx[P];
#pragma unroll
        for (int p = 0; p < P; ++p) {
            uc[p] = ctr[p * TYH];
            mn[p] = __longlong_as_double(0x7ff0000000000000LL);    // +inf
            mx[p] = __longlong_as_double(0xfff0000000000000LL);    // -inf
            kmn[p] = 0;
            kmx[p] = 0;
        }
        Eval::template run<P, TYH, MODE, Epi::kMin, Epi::kMax, Epi::kPol, ICMP>(ctr, tab, uc, mn, mx, kmn, kmx);
        if (ly < geo.ncols) {
#pragma unroll
            for (int p = 0; p < P; ++p) {
                if (lx0 + p < rcount) {
                    int64_t o = (int64_t)(rbase + lx0 + p) * geo.pitch + (geo.col0 + ly);
                    epi(o, uc[p], mn[p], mx[p], kmn[p], kmx[p], pre[Epi::kPre ? p : 0], red);
                }
            }
        }
        __syncthreads();     // everyone is done with this slot before it is refilled
        tile = sched ? s_next[slot] : next;
    }
    if (sched && tid == 0) {
        __threadfence();
        if (atomicAdd(sched + 1, 1ULL) == (unsigned long long)gridDim.x - 1) {
            sched[0] = 0ULL;     // last CTA out: re-arm the counters for the next launch
            sched[1] = 0ULL;
            __threadfence();
        }
    }

    if (Epi::kRed) {
        __shared__ double sred[2][K1_THREADS / 32];
        double a = warp_max(red.a), b = warp_max(red.b);
        if ((tid & 31) == 0) {
            sred[0][tid >> 5] = a;
            sred[1][tid >> 5] = b;
        }
        __syncthreads();
        if (tid < 32) {
            a = tid < K1_THREADS / 32 ? sred[0][tid] : 0.0;
            b = tid < K1_THREADS / 32 ? sred[1][tid] : 0.0;
            a = warp_max(a);
            b = warp_max(b);
            if (tid == 0) {
                atomic_max_nonneg(&red_out[0], a);
                atomic_max_nonneg(&red_out[1], b);
            }
        }
    }
}

template <int H, int P>
constexpr int k_deep_smem_bytes() {
    return 2 * ((((2 * P + 2 * H) * (K1_TY + 2 * H + 2) * 8) + 127) / 128 * 128) + 64;
}

}  // namespace pde
