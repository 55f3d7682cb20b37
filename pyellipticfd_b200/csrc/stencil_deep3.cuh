// K1 in three dimensions: structured wide-stencil min/max kernel for the deep-interior points
// of a 3-D lattice (stencils.py:46-61 builds the 3-D stencil; ddg.py:251-292 is dimension
// agnostic).  Replaces the 40-byte explicit pair rows the 3-D grids ran through before: a deep
// point costs 16 B of HBM traffic (read u, write lambda) instead of 40 B per pair (49 pairs at
// r = 2).
//
// The lattice [NX][NY][NZ] is stored as NX*NY state-layout rows of `pitch` doubles (z is the
// contiguous axis), i.e. it IS a 2-D state array whose row index is ix*NY + iy; every
// per-point array and all the "one pass over the state" kernels (pack, band rows, Krylov
// passes) are shared with the 2-D case.  One CTA owns a TX x TY x TZ = P x 4 x 64 brick of
// points; the brick plus a halo of H cells is staged by ONE 3-D TMA load
// (cp.async.bulk.tensor.3d, zero fill outside the lattice), double buffered over a persistent
// brick loop.  Lanes run along z (conflict-free shared-memory reads, coalesced stores); a
// thread owns P consecutive x-planes of one (y, z) column and evaluates the direction table
// from register windows (stencil_ct3.cuh) or from a run-time table.
//
// Arithmetic and tie rule: as K1 (stencil_deep.cuh) -- mode 0 is ddg.py:281 operation by
// operation, `<=` keeps the LAST minimal pair, `>` the FIRST maximal pair.
#pragma once
#include "stencil_deep.cuh"
#include "stencil_ct3.cuh"

namespace pde {

constexpr int K3_TZ = 64;       // brick extent along z (contiguous)
constexpr int K3_TY = 4;        // along y
constexpr int K3_THREADS = K3_TZ * K3_TY;

__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *tmap, int c0, int c1,
                                            int c2, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
        : "memory");
}

struct Deep3Geom {
    int64_t pitch;          // doubles per (x, y) row
    int ny;                 // rows per x-plane
    int x0, y0, z0;         // first deep point on every axis
    int nxd, nyd, nzd;      // extent of the deep region
    int tiles_x, tiles_y, tiles_z;
};

// tile-local offsets of the two neighbours of every table row (run-time-table evaluator)
struct Deep3Offsets {
    int off0[PDE_MAX_DIRS], off1[PDE_MAX_DIRS];
};

struct EvalRuntime3 {
    template <int P, int SX, int SY, int MODE, bool WMIN, bool WMAX, bool WPOL>
    static __device__ __forceinline__ void run(const double *__restrict__ ctr, const DeepTable &tab,
                                               const int *off0, const int *off1,
                                               const double (&uc)[P], double (&mn)[P],
                                               double (&mx)[P], int (&kmn)[P], int (&kmx)[P]) {
        const int D = tab.D;
#pragma unroll 2
        for (int k = 0; k < D; ++k) {
            const double *q0 = ctr + off0[k];
            const double *q1 = ctr + off1[k];
            const double w_0 = tab.w_0[k], w_1 = tab.w_1[k], w0 = tab.w0[k];
#pragma unroll
            for (int p = 0; p < P; ++p) {
                double u0 = q0[p * SX], u1 = q1[p * SX];
                double d;
                if (MODE == 0)
                    d = pair_value_exact(uc[p], u0, u1, w_0, w_1, w0);
                else
                    d = (fma(w_1, u1 - uc[p], w_0 * (u0 - uc[p]))) * w0;
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

template <int H, int P>
struct Deep3Tile {
    static constexpr int TX = P;
    static constexpr int TZH = K3_TZ + 2 * H + 2;     // +2: 16-byte aligned box origin (see K1)
    static constexpr int TYH = K3_TY + 2 * H;
    static constexpr int TXH = TX + 2 * H;
    static constexpr int SY = TZH, SX = TYH * TZH;
    static constexpr int TILE_BYTES = TXH * TYH * TZH * 8;
    static constexpr int TILE_STRIDE = (TILE_BYTES + 127) / 128 * 128;
    static constexpr int SMEM = 2 * TILE_STRIDE + 64;
};

// Band points handled inside the brick loop: after every brick a CTA evaluates its share of the
// band points (one warp per point over its explicit pair rows, as k_band).  In 3-D the band is a
// surface layer two cells thick -- as much work as all deep points at 259^3 -- and a separate
// band kernel cannot overlap with this one (two 125-register CTAs fill the register file); here
// the memory-bound band rows of one CTA can overlap with the fp64-bound brick of its neighbour.
// Opt-in: measured slower than the separate launch (see ddg_kernels.cu::band3_fused).
struct Band3 {
    int64_t n;                       // 0: band points are served by a separate launch
    const int64_t *center, *ptr, *jj;
    const double *ww;
};

template <int MODE, class Epi>
__device__ __forceinline__ void band_point_warp(const Band3 &B, int64_t i, const double *__restrict__ S,
                                                const Epi &epi, BlockMax2 &red) {
    const int lane = threadIdx.x & 31;
    const int64_t o = B.center[i];
    const double uc = S[o];
    double mn = __longlong_as_double(0x7ff0000000000000LL);
    double mx = __longlong_as_double(0xfff0000000000000LL);
    int kmn = -1, kmx = 0x7fffffff;
    const int64_t b = B.ptr[i], e = B.ptr[i + 1];
    for (int64_t k = b + lane; k < e; k += 32) {
        double u0 = S[B.jj[2 * k]], u1 = S[B.jj[2 * k + 1]];
        double w_0 = B.ww[3 * k], w_1 = B.ww[3 * k + 1], w0 = B.ww[3 * k + 2];
        double d;
        if (MODE == 0)
            d = pair_value_exact(uc, u0, u1, w_0, w_1, w0);
        else
            d = (fma(w_1, u1 - uc, w_0 * (u0 - uc))) * w0;
        if (Epi::kMin && d <= mn) {        // rows ascend within a lane: keeps the last
            mn = d;
            kmn = (int)(k - b);
        }
        if (Epi::kMax && d > mx) {         // keeps the first
            mx = d;
            kmx = (int)(k - b);
        }
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {     // tie rule across lanes: min -> highest row, max -> lowest
        double omn = __shfl_xor_sync(0xffffffffu, mn, s);
        int okn = __shfl_xor_sync(0xffffffffu, kmn, s);
        double omx = __shfl_xor_sync(0xffffffffu, mx, s);
        int okx = __shfl_xor_sync(0xffffffffu, kmx, s);
        if (Epi::kMin && (omn < mn || (omn == mn && okn > kmn))) {
            mn = omn;
            kmn = okn;
        }
        if (Epi::kMax && (omx > mx || (omx == mx && okx < kmx))) {
            mx = omx;
            kmx = okx;
        }
    }
    if (lane == 0) epi(o, uc, mn, mx, kmn, kmx, Epi::kPre ? epi.pre(o) : 0.0, red);
}

template <class Eval, int H, int P, int MODE, class Epi>
__global__ void __launch_bounds__(K3_THREADS, 2)
k_deep3(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ DeepTable tab,
        const __grid_constant__ Deep3Offsets offs, const Deep3Geom geo, const Epi epi,
        double *red_out, const unsigned long long *done_flag, const Band3 band,
        const double *__restrict__ S) {
    using T = Deep3Tile<H, P>;
    if (done_flag && *done_flag) return;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *buf0 = reinterpret_cast<double *>(smem_raw);
    double *buf1 = reinterpret_cast<double *>(smem_raw + T::TILE_STRIDE);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + 2 * T::TILE_STRIDE);

    const int tid = threadIdx.x;
    const int c = tid & (K3_TZ - 1);
    const int yy = tid / K3_TZ;
    const int tyz = geo.tiles_y * geo.tiles_z;
    const int ntiles = geo.tiles_x * tyz;
    const int zsh = (geo.z0 - H) & 1;

    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    auto issue = [&](int tile, int slot) {
        int tx = tile / tyz, rem = tile - tx * tyz;
        int ty = rem / geo.tiles_z, tz = rem - ty * geo.tiles_z;
        uint64_t *bar = &bars[slot];
        mbar_expect_tx(bar, T::TILE_BYTES);
        tma_load_3d(slot ? buf1 : buf0, &tmap, geo.z0 + tz * K3_TZ - H - zsh,
                    geo.y0 + ty * K3_TY - H, geo.x0 + tx * T::TX - H, bar);
    };

    BlockMax2 red{0.0, 0.0};
    int tile = blockIdx.x;
    if (tile < ntiles && tid == 0) issue(tile, 0);

    for (int it = 0; tile < ntiles; ++it, tile += gridDim.x) {
        const int slot = it & 1;
        const int next = tile + gridDim.x;
        if (next < ntiles && tid == 0) issue(next, slot ^ 1);
        const double *sm = slot ? buf1 : buf0;
        const int tx = tile / tyz, rem = tile - tx * tyz;
        const int ty = rem / geo.tiles_z, tz = rem - ty * geo.tiles_z;
        const int lx0 = tx * T::TX, ly = ty * K3_TY + yy, lz = tz * K3_TZ + c;
        const bool in_yz = ly < geo.nyd && lz < geo.nzd;
        // state offset of the thread's first point
        const int64_t o0 = ((int64_t)(geo.x0 + lx0) * geo.ny + (geo.y0 + ly)) * geo.pitch + (geo.z0 + lz);
        const int64_t ox = (int64_t)geo.ny * geo.pitch;
        double pre[Epi::kPre ? P : 1];
        if (Epi::kPre) {
#pragma unroll
            for (int p = 0; p < P; ++p)
                pre[p] = (in_yz && lx0 + p < geo.nxd) ? epi.pre(o0 + p * ox) : 0.0;
        }
        mbar_wait(&bars[slot], (it >> 1) & 1);

        const double *ctr = sm + H * T::SX + (H + yy) * T::SY + (H + c + zsh);
        double uc[P], mn[P], mx[P];
        int kmn[P], kmx[P];
#pragma unroll
        for (int p = 0; p < P; ++p) {
            uc[p] = ctr[p * T::SX];
            mn[p] = __longlong_as_double(0x7ff0000000000000LL);
            mx[p] = __longlong_as_double(0xfff0000000000000LL);
            kmn[p] = 0;
            kmx[p] = 0;
        }
        Eval::template run<P, T::SX, T::SY, MODE, Epi::kMin, Epi::kMax, Epi::kPol>(
            ctr, tab, offs.off0, offs.off1, uc, mn, mx, kmn, kmx);
        if (in_yz) {
#pragma unroll
            for (int p = 0; p < P; ++p)
                if (lx0 + p < geo.nxd)
                    epi(o0 + p * ox, uc[p], mn[p], mx[p], kmn[p], kmx[p], pre[Epi::kPre ? p : 0], red);
        }
        __syncthreads();
        if (band.n > 0) {     // this brick's share of the band points, one warp per point
            const int64_t q0 = band.n * tile / ntiles, q1 = band.n * (int64_t)(tile + 1) / ntiles;
            for (int64_t i = q0 + (tid >> 5); i < q1; i += K3_THREADS / 32)
                band_point_warp<MODE>(band, i, S, epi, red);
        }
    }

    if (Epi::kRed) {
        __shared__ double sred[2][K3_THREADS / 32];
        double a = warp_max(red.a), b = warp_max(red.b);
        if ((tid & 31) == 0) {
            sred[0][tid >> 5] = a;
            sred[1][tid >> 5] = b;
        }
        __syncthreads();
        if (tid < 32) {
            a = tid < K3_THREADS / 32 ? sred[0][tid] : 0.0;
            b = tid < K3_THREADS / 32 ? sred[1][tid] : 0.0;
            a = warp_max(a);
            b = warp_max(b);
            if (tid == 0) {
                atomic_max_nonneg(&red_out[0], a);
                atomic_max_nonneg(&red_out[1], b);
            }
        }
    }
}

}  // namespace pde
