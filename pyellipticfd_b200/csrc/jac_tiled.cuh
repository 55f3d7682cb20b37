// Tiled policy-Jacobian rows for deep-interior points (the operator passes of BiCGStab).
//
// The plain row kernel (bicgstab.cu::k_rows) gathers the two neighbours of a point through
// L1/L2 behind a dependent load chain policy -> direction table -> neighbour; it is latency
// bound (43 % of the DRAM peak, profiles/r01_krylov_full.md) and every attempt to add
// memory-level parallelism lost locality.  Here the source vector is staged like in K1: a
// (2P + 2H) x (128 + 2H + 2) tile per CTA, one TMA load per tile, double buffered over a
// persistent tile loop; the neighbour reads become shared-memory reads, the direction table
// is a shared-memory copy (the policy index differs per thread, and a per-thread index into
// the kernel-parameter constant bank serialises), the policy itself is a coalesced 2-byte
// load.  Same arithmetic as jacobian.cuh::jac_row_deep (difference form in the reference's
// operation order), so rows are bit-identical to the plain kernel; only the grouping of the
// partial sums of the fused dot products changes.
#pragma once
#include "jacobian.cuh"
#include "krylov.cuh"
#include "stencil_deep.cuh"

namespace pde {

struct TileTab {            // shared-memory copy of the direction table, offsets relative to the
    int off0[PDE_MAX_DIRS]; // centre inside a tile row-major with row length TYH
    int off1[PDE_MAX_DIRS];
    double w_0[PDE_MAX_DIRS], w_1[PDE_MAX_DIRS], w0[PDE_MAX_DIRS];
};

__device__ __forceinline__ double jac_row_tile(const TileTab &T, const JacDev &J, int64_t o,
                                               unsigned pm, unsigned px, double xc,
                                               const double *__restrict__ c) {
    if (pm == PDE_POLICY_IDENTITY) return xc;
    double m = pair_value_exact(xc, c[T.off0[pm]], c[T.off1[pm]], T.w_0[pm], T.w_1[pm], T.w0[pm]);
    double cm = J.cmin ? J.cmin[o] : J.cmin_s;
    double y = cm * m;
    if (J.pmax) {
        double mm = pair_value_exact(xc, c[T.off0[px]], c[T.off1[px]], T.w_0[px], T.w_1[px],
                                     T.w0[px]);
        double cx = J.cmax ? J.cmax[o] : J.cmax_s;
        y += cx * mm;
    }
    return y;
}

// Op: NRED, kSkipOnHalf, prepare(scal), jac() -> const JacDev&, aux(o) (the second operand of
// the fused dot product, fetched together with the policies), tile(o, v, aux, acc)
template <class Op, int H, int P>
__global__ void __launch_bounds__(K1_THREADS, 3)
k_jac_tiled(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ DeepTable tab,
            const DeepGeom geo, Op op, double *partials, const BicgScal *scal) {
    constexpr int TX = 2 * P;
    constexpr int TYH = K1_TY + 2 * H + 2;
    constexpr int TXH = TX + 2 * H;
    constexpr int TILE_BYTES = TXH * TYH * 8;
    constexpr int TILE_STRIDE = (TILE_BYTES + 127) / 128 * 128;
    static_assert(K1_THREADS == PASS_THREADS, "block_store_partials assumes PASS_THREADS");

    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *buf0 = reinterpret_cast<double *>(smem_raw);
    double *buf1 = reinterpret_cast<double *>(smem_raw + TILE_STRIDE);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + 2 * TILE_STRIDE);
    __shared__ TileTab T;

    DD acc[Op::NRED > 0 ? Op::NRED : 1] = {};
    const bool skip = scal && (scal->done || (Op::kSkipOnHalf && scal->half));
    if (!skip) {
        op.prepare(scal);
        const int tid = threadIdx.x;
        const int c = tid & (K1_TY - 1);
        const int grp = tid >> 7;
        const int ntiles = geo.tiles_x * geo.tiles_y;
        const int ysh = (geo.col0 - H) & 1;
        if (tid < tab.D) {
            T.off0[tid] = tab.a0[tid] * TYH + tab.b0[tid];
            T.off1[tid] = tab.a1[tid] * TYH + tab.b1[tid];
            T.w_0[tid] = tab.w_0[tid];
            T.w_1[tid] = tab.w_1[tid];
            T.w0[tid] = tab.w0[tid];
        }
        if (tid == 0) {
            mbar_init(&bars[0], 1);
            mbar_init(&bars[1], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();

        auto issue = [&](int tile, int slot) {
            int tx = tile / geo.tiles_y, ty = tile - tx * geo.tiles_y;
            int x0 = geo.row0 + tx * TX - H;
            int y0 = geo.col0 + ty * K1_TY - H - ysh;
            uint64_t *bar = &bars[slot];
            mbar_expect_tx(bar, TILE_BYTES);
            tma_load_2d(slot ? buf1 : buf0, &tmap, y0, x0, bar);
        };

        int tile = blockIdx.x;
        if (tile < ntiles && tid == 0) issue(tile, 0);
        const JacDev &J = op.jac();
        for (int it = 0; tile < ntiles; ++it, tile += gridDim.x) {
            const int slot = it & 1;
            const int next = tile + gridDim.x;
            if (next < ntiles && tid == 0) issue(next, slot ^ 1);
            const double *sm = slot ? buf1 : buf0;
            const int tx = tile / geo.tiles_y, ty = tile - tx * geo.tiles_y;
            const int lx0 = tx * TX + grp * P;
            const int ly = ty * K1_TY + c;
            // global operands of the P points first (policies, dot-product operand): independent
            // loads, issued before the tile is waited for
            unsigned pm[P], px[P];
            double ax[P];
            const bool incol = ly < geo.ncols;
#pragma unroll
            for (int p = 0; p < P; ++p) {
                const bool in = incol && lx0 + p < geo.nrows;
                const int64_t o = (int64_t)(geo.row0 + lx0 + p) * geo.pitch + (geo.col0 + ly);
                pm[p] = in ? (unsigned)J.pmin[o] : (unsigned)PDE_POLICY_IDENTITY;
                px[p] = (in && J.pmax) ? (unsigned)J.pmax[o] : 0u;
                ax[p] = in ? op.aux(o) : 0.0;
            }
            mbar_wait(&bars[slot], (it >> 1) & 1);
            const double *ctr = sm + (H + grp * P) * TYH + (H + c + ysh);
            if (incol) {
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    if (lx0 + p < geo.nrows) {
                        const int64_t o = (int64_t)(geo.row0 + lx0 + p) * geo.pitch + (geo.col0 + ly);
                        const double *cp = ctr + p * TYH;
                        op.tile(o, jac_row_tile(T, J, o, pm[p], px[p], cp[0], cp), ax[p], acc);
                    }
                }
            }
            __syncthreads();     // everyone is done with this slot before it is refilled
        }
    }
    if constexpr (Op::NRED > 0) block_store_partials(acc, partials + (int64_t)blockIdx.x * DDW * Op::NRED);
}

// ---- fused s / t pass of BiCGStab -----------------------------------------------------------
// iterative.py:  r -= alpha*v (this is s);  shat = M s;  t = A shat;  omega = <t,s>/<t,t>
// as ONE pass: tiles of r, v and dinv (three TMA loads per tile on one barrier) give
// shat = dinv*(r - alpha*v) on the tile AND its halo in shared memory (the halo values are
// recomputed by the neighbouring tiles with the same arithmetic: identical bits), the rows of J
// are applied from that tile, and s, t are written once.  5 vector passes (r, v, dinv in; s, t
// out) instead of 8 (OpS: r, v, dinv in, s, shat out; OpAt: shat, s in, t out).  s goes to its
// own vector -- r is read by the neighbouring tiles' halos -- and <s,s> decides the half-step
// exit afterwards (stage ST_ST), so t is also computed on the one iteration that does not use it.
constexpr int KF_THREADS = 512;      // 128 columns x 4 row groups of P = 4 rows: TX = 16 like K1

template <int H>
__global__ void __launch_bounds__(KF_THREADS, 1)
k_jac_fused_st(const __grid_constant__ CUtensorMap tmap_r, const __grid_constant__ CUtensorMap tmap_v,
               const __grid_constant__ CUtensorMap tmap_d, const __grid_constant__ DeepTable tab,
               const DeepGeom geo, const JacDev J, double *__restrict__ s_out,
               double *__restrict__ t_out, double *partials, BicgScal *scal) {
    constexpr int P = 4, NG = KF_THREADS / K1_TY, TX = NG * P;
    constexpr int TYH = K1_TY + 2 * H + 2;
    constexpr int TXH = TX + 2 * H;
    constexpr int TILE_ELEMS = TXH * TYH;
    constexpr int TILE_BYTES = TILE_ELEMS * 8;
    constexpr int TILE_STRIDE = (TILE_BYTES + 127) / 128 * 128;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + 6 * TILE_STRIDE);
    __shared__ TileTab T;
    auto tile_ptr = [&](int slot, int which) {
        return reinterpret_cast<double *>(smem_raw + (size_t)(slot * 3 + which) * TILE_STRIDE);
    };

    DD acc[3] = {};
    if (!scal->done) {
        const double alpha = scal->alpha;
        const int tid = threadIdx.x;
        const int c = tid & (K1_TY - 1);
        const int grp = tid / K1_TY;
        const int ntiles = geo.tiles_x * geo.tiles_y;
        const int ysh = (geo.col0 - H) & 1;
        if (tid < tab.D) {
            T.off0[tid] = tab.a0[tid] * TYH + tab.b0[tid];
            T.off1[tid] = tab.a1[tid] * TYH + tab.b1[tid];
            T.w_0[tid] = tab.w_0[tid];
            T.w_1[tid] = tab.w_1[tid];
            T.w0[tid] = tab.w0[tid];
        }
        if (tid == 0) {
            mbar_init(&bars[0], 1);
            mbar_init(&bars[1], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();

        auto issue = [&](int tile, int slot) {
            int tx = tile / geo.tiles_y, ty = tile - tx * geo.tiles_y;
            int x0 = geo.row0 + tx * TX - H;
            int y0 = geo.col0 + ty * K1_TY - H - ysh;
            uint64_t *bar = &bars[slot];
            mbar_expect_tx(bar, 3 * TILE_BYTES);
            tma_load_2d(tile_ptr(slot, 0), &tmap_r, y0, x0, bar);
            tma_load_2d(tile_ptr(slot, 1), &tmap_v, y0, x0, bar);
            tma_load_2d(tile_ptr(slot, 2), &tmap_d, y0, x0, bar);
        };

        int tile = blockIdx.x;
        if (tile < ntiles && tid == 0) issue(tile, 0);
        for (int it = 0; tile < ntiles; ++it, tile += gridDim.x) {
            const int slot = it & 1;
            const int next = tile + gridDim.x;
            if (next < ntiles && tid == 0) issue(next, slot ^ 1);
            double *rt = tile_ptr(slot, 0);
            const double *vt = tile_ptr(slot, 1), *dt = tile_ptr(slot, 2);
            const int tx = tile / geo.tiles_y, ty = tile - tx * geo.tiles_y;
            const int lx0 = tx * TX + grp * P;
            const int ly = ty * K1_TY + c;
            const bool incol = ly < geo.ncols;
            unsigned pm[P], px[P];
#pragma unroll
            for (int p = 0; p < P; ++p) {        // policies: issued before the tile is waited for
                const bool in = incol && lx0 + p < geo.nrows;
                const int64_t o = (int64_t)(geo.row0 + lx0 + p) * geo.pitch + (geo.col0 + ly);
                pm[p] = in ? (unsigned)J.pmin[o] : (unsigned)PDE_POLICY_IDENTITY;
                px[p] = (in && J.pmax) ? (unsigned)J.pmax[o] : 0u;
            }
            mbar_wait(&bars[slot], (it >> 1) & 1);
            const int ctr = (H + grp * P) * TYH + (H + c + ysh);
            double sown[P];
#pragma unroll
            for (int p = 0; p < P; ++p)          // s at the thread's own points (OpS's expression)
                sown[p] = __dsub_rn(rt[ctr + p * TYH], __dmul_rn(alpha, vt[ctr + p * TYH]));
            __syncthreads();
            for (int e = tid; e < TILE_ELEMS; e += KF_THREADS)   // shat on tile + halo, in place
                rt[e] = dt[e] * __dsub_rn(rt[e], __dmul_rn(alpha, vt[e]));
            __syncthreads();
            if (incol) {
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    if (lx0 + p < geo.nrows) {
                        const int64_t o = (int64_t)(geo.row0 + lx0 + p) * geo.pitch + (geo.col0 + ly);
                        const double *cp = rt + ctr + p * TYH;
                        const double tv = jac_row_tile(T, J, o, pm[p], px[p], cp[0], cp);
                        const double sv = sown[p];
                        s_out[o] = sv;
                        t_out[o] = tv;
                        dd_mac(acc[0], sv, sv);
                        dd_mac(acc[1], tv, sv);
                        dd_mac(acc[2], tv, tv);
                    }
                }
            }
            __syncthreads();     // everyone is done with this slot before it is refilled
        }
    }
    block_store_partials<3, KF_THREADS>(acc, partials + (int64_t)blockIdx.x * 3 * DDW);
}

template <int H>
constexpr int k_jac_fused_smem_bytes() {
    return 6 * ((((KF_THREADS / K1_TY * 4 + 2 * H) * (K1_TY + 2 * H + 2) * 8) + 127) / 128 * 128) + 64;
}

template <int H, int P>
constexpr int k_jac_tiled_smem_bytes() {
    return 2 * ((((2 * P + 2 * H) * (K1_TY + 2 * H + 2) * 8) + 127) / 128 * 128) + 64;
}

}  // namespace pde
