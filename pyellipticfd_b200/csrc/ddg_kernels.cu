// ddg operators on the regular grid: structured deep-interior kernel (K1,
// stencil_deep.cuh) + CSR band kernel, with the convex-envelope / Pucci
// epilogues fused in.  C ABI in include/pde_b200.h.
#include <algorithm>
#include <vector>
#include <cstdlib>
#include <type_traits>
#include "stencil_deep.cuh"
#include "stencil_deep3.cuh"
#include "tiles.cuh"

using namespace pde;

namespace pde {

// ---- band kernel: one thread per band point, explicit pair rows -----------------
// Scans the point's rows in the given order with `<=` (min keeps the LAST minimal
// row) and `>` (max keeps the FIRST maximal row): ddg.py:283-292.
template <int MODE, class Epi>
__global__ void __launch_bounds__(128)
k_band(int64_t n_band, const int64_t *__restrict__ center, const int64_t *__restrict__ ptr,
       const int64_t *__restrict__ jj, const double *__restrict__ ww,
       const double *__restrict__ S, const Epi epi, double *red_out,
       const unsigned long long *done_flag) {
    // one WARP per band point: lanes stride over the point's pair rows (coalesced reads of
    // the row tables), then a warp reduction that keeps the tie rule: among equal values
    // the minimum keeps the highest row number, the maximum the lowest.
    if (done_flag && *done_flag) return;
    BlockMax2 red{0.0, 0.0};
    const int lane = threadIdx.x & 31;
    int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (i < n_band) {
        int64_t o = center[i];
        double uc = S[o];
        double mn = __longlong_as_double(0x7ff0000000000000LL);
        double mx = __longlong_as_double(0xfff0000000000000LL);
        int kmn = -1, kmx = 0x7fffffff;
        int64_t b = ptr[i], e = ptr[i + 1];
        for (int64_t k = b + lane; k < e; k += 32) {
            double u0 = S[jj[2 * k]], u1 = S[jj[2 * k + 1]];
            double w_0 = ww[3 * k], w_1 = ww[3 * k + 1], w0 = ww[3 * k + 2];
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
        for (int s = 16; s > 0; s >>= 1) {
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
    if (Epi::kRed) {
        double a = warp_max(red.a), bb = warp_max(red.b);
        if (lane == 0) {
            atomic_max_nonneg(&red_out[0], a);
            atomic_max_nonneg(&red_out[1], bb);
        }
    }
}

// ---- band kernel on translation classes (3-D lattices) ------------------------------------------
// Same evaluation, row order and tie rule as k_band, but the rows of a band point come from its
// class table (pde_grid_set_band_classes): a neighbour is either a lattice point at a fixed state
// offset from the centre or the q-th entry of the point's own list of wall points; the weights
// are the class's.  ~0.2 kB per band point (class number + wall list) instead of ~2.2 kB of rows.
// (One THREAD per band point in class order -- neighbouring threads sharing the class rows -- was
// measured slower: 131^3 0.25 vs 0.13 ms, 259^3 0.82 vs 0.73 ms for the whole d2min application.)
template <int MODE, class Epi>
__global__ void __launch_bounds__(128)
k_band_cls(int64_t n_band, const int64_t *__restrict__ center, const uint16_t *__restrict__ cls_id,
           const int32_t *__restrict__ cls_ptr, const int32_t *__restrict__ cls_ref,
           const double *__restrict__ cls_w, const int32_t *__restrict__ wl_ptr,
           const int32_t *__restrict__ wl, const double *__restrict__ S, const Epi epi,
           double *red_out, const unsigned long long *done_flag) {
    if (done_flag && *done_flag) return;
    BlockMax2 red{0.0, 0.0};
    const int lane = threadIdx.x & 31;
    int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (i < n_band) {
        const int64_t o = center[i];
        const double uc = S[o];
        const int c = cls_id[i];
        const int b = cls_ptr[c], e = cls_ptr[c + 1];
        const int32_t *mywl = wl + wl_ptr[i];
        double mn = __longlong_as_double(0x7ff0000000000000LL);
        double mx = __longlong_as_double(0xfff0000000000000LL);
        int kmn = -1, kmx = 0x7fffffff;
        for (int k = b + lane; k < e; k += 32) {
            const int r0 = cls_ref[2 * k], r1 = cls_ref[2 * k + 1];
            const double u0 = (r0 & 1) ? S[mywl[r0 >> 1]] : S[o + (r0 >> 1)];
            const double u1 = (r1 & 1) ? S[mywl[r1 >> 1]] : S[o + (r1 >> 1)];
            const double w_0 = cls_w[3 * k], w_1 = cls_w[3 * k + 1], w0 = cls_w[3 * k + 2];
            double d;
            if (MODE == 0)
                d = pair_value_exact(uc, u0, u1, w_0, w_1, w0);
            else
                d = (fma(w_1, u1 - uc, w_0 * (u0 - uc))) * w0;
            if (Epi::kMin && d <= mn) {
                mn = d;
                kmn = k - b;
            }
            if (Epi::kMax && d > mx) {
                mx = d;
                kmx = k - b;
            }
        }
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) {
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
    if (Epi::kRed) {
        double a = warp_max(red.a), bb = warp_max(red.b);
        if (lane == 0) {
            atomic_max_nonneg(&red_out[0], a);
            atomic_max_nonneg(&red_out[1], bb);
        }
    }
}

// ---- wall rows of the residual / Euler step -------------------------------------------
// mode 0: F = W - g  (convex_envelope.py:44; pucci.py:130,141), red[0] = max|F|
// mode 1: Unew = U - dt*(U - g), red[0] = max|F|, red[1] = max|U-Unew|
__global__ void k_wall(int64_t nb, int64_t off, const double *__restrict__ W,
                       const double *__restrict__ g, double *__restrict__ out, double dt, int mode,
                       double *red_out, const unsigned long long *done_flag) {
    if (done_flag && *done_flag) return;
    double ra = 0.0, rb = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nb;
         i += (int64_t)gridDim.x * blockDim.x) {
        double u = W[off + i];
        double f = __dsub_rn(u, g[off + i]);
        ra = fmax(ra, fabs(f));
        if (mode == 0) {
            out[off + i] = f;
        } else {
            double un = __dsub_rn(u, __dmul_rn(dt, f));
            out[off + i] = un;
            rb = fmax(rb, fabs(__dsub_rn(u, un)));
        }
    }
    ra = warp_max(ra);
    rb = warp_max(rb);
    if ((threadIdx.x & 31) == 0) {
        atomic_max_nonneg(&red_out[0], ra);
        atomic_max_nonneg(&red_out[1], rb);
    }
}

// ---- host-side launch helpers (deep_region, make_tmap: tiles.cuh) -------------------------
template <class Eval, int H, int P, int MODE, class Epi, int MINB = 2, int ICMP = 0>
static int launch_deep_hp(pde_grid *g, const double *S, const Epi &epi, double *red,
                          const unsigned long long *done, cudaStream_t st) {
    DeepRegion dr = deep_region(g, g->win_lo, g->win_hi);
    DeepRegion dr2 = deep_region(g, g->win2_lo, g->win2_hi);
    if (dr.nrows <= 0) {            // only the second range has deep rows
        dr = dr2;
        dr2.nrows = 0;
    }
    if (dr.nrows <= 0 || dr.ncols <= 0) return 0;
    constexpr int TX = 2 * P;
    DeepGeom geo;
    geo.pitch = g->pitch;
    geo.row0 = dr.row0;
    geo.col0 = dr.col0;
    geo.nrows = dr.nrows;
    geo.ncols = dr.ncols;
    geo.tiles_x = (dr.nrows + TX - 1) / TX;
    geo.tiles_y = (dr.ncols + K1_TY - 1) / K1_TY;
    geo.split = geo.tiles_x;
    geo.row0b = dr2.row0;
    geo.nrowsb = dr2.nrows > 0 ? dr2.nrows : 0;
    if (geo.nrowsb > 0) geo.tiles_x += (dr2.nrows + TX - 1) / TX;
    CUtensorMap tmap;
    int rc = make_tmap(g, S, K1_TY + 2 * H + 2, TX + 2 * H, &tmap);
    if (rc) return rc;
    constexpr int smem = k_deep_smem_bytes<H, P>();
    auto kern = k_deep<Eval, H, P, MODE, Epi, MINB, ICMP>;
    // the opt-in is per device: one bit per device, per template instantiation
    static unsigned long long attr_set = 0ull;
    const unsigned long long dev_bit = 1ull << (g->device & 63);
    if (!(attr_set & dev_bit)) {
        PDE_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_set |= dev_bit;
    }
    int ntiles = geo.tiles_x * geo.tiles_y;
    int grid = std::min(ntiles, g->num_sms * MINB);
    kern<<<grid, K1_THREADS, smem, st>>>(tmap, g->tab, geo, epi, red, done,
                                         g->dyn_sched ? g->d_ctrl + 8 : nullptr);
    PDE_LAUNCH_CHECK();
    return 0;
}

// ---- 3-D lattices: bricks staged by 3-D TMA loads (stencil_deep3.cuh) -----------------------
static bool band3_fused(const pde_grid *g, int ntiles) {
    // opt-in (PDE_K3_FUSE_BAND=1, read per call: a test compares both paths).  Measured SLOWER
    // than the separate band launch -- 259^3: 1.27 vs 0.92 ms, 131^3: 0.30 vs 0.19 ms: at 16 warps
    // per SM the dependent loads of a band point (centre -> row pointer -> offsets -> values) are
    // latency bound and a brick's share of band points takes longer than the brick itself.
    const char *e = getenv("PDE_K3_FUSE_BAND");
    const bool on = e && atoi(e) != 0;
    return on && g->nmid > 0 && g->n_band > 0 && ntiles >= 2 * g->num_sms;
}
static int deep3_tiles(const pde_grid *g) {
    const int r = g->radius, P = 8;
    int nxd = (int)g->nx3 - 2 * r, nyd = (int)g->nmid - 2 * r, nzd = (int)g->ny - 2 * r;
    if (nxd <= 0 || nyd <= 0 || nzd <= 0) return 0;
    return ((nxd + P - 1) / P) * ((nyd + K3_TY - 1) / K3_TY) * ((nzd + K3_TZ - 1) / K3_TZ);
}

template <class Eval, int H, int MODE, class Epi>
static int launch_deep3_h(pde_grid *g, const double *S, const Epi &epi, double *red,
                          const unsigned long long *done, cudaStream_t st) {
    constexpr int P = 8;
    using T = Deep3Tile<H, P>;
    const int r = g->radius;
    Deep3Geom geo;
    geo.pitch = g->pitch;
    geo.ny = (int)g->nmid;
    geo.x0 = geo.y0 = geo.z0 = r;
    geo.nxd = (int)g->nx3 - 2 * r;
    geo.nyd = (int)g->nmid - 2 * r;
    geo.nzd = (int)g->ny - 2 * r;
    if (geo.nxd <= 0 || geo.nyd <= 0 || geo.nzd <= 0) return 0;
    PDE_REQUIRE(g->win_lo == 0 && g->win_hi == g->nx_local, "row windows are not available in 3-D");
    geo.tiles_x = (geo.nxd + T::TX - 1) / T::TX;
    geo.tiles_y = (geo.nyd + K3_TY - 1) / K3_TY;
    geo.tiles_z = (geo.nzd + K3_TZ - 1) / K3_TZ;
    Deep3Offsets offs;
    for (int k = 0; k < g->tab.D; ++k) {
        const int *q = g->nd_abc[k];
        offs.off0[k] = q[0] * T::SX + q[1] * T::SY + q[2];
        offs.off1[k] = q[3] * T::SX + q[4] * T::SY + q[5];
    }
    PDE_REQUIRE((reinterpret_cast<uintptr_t>(S) & 15) == 0, "state buffer must be 16-byte aligned");
    CUtensorMap tmap;
    cuuint64_t gdim[3] = {(cuuint64_t)g->ny, (cuuint64_t)g->nmid, (cuuint64_t)g->nx3};
    cuuint64_t gstr[2] = {(cuuint64_t)g->pitch * 8, (cuuint64_t)g->nmid * g->pitch * 8};
    cuuint32_t box[3] = {(cuuint32_t)T::TZH, (cuuint32_t)T::TYH, (cuuint32_t)T::TXH};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult cr = g->encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double *>(S), gdim,
                            gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) return pde::fail("cuTensorMapEncodeTiled (3-D) failed", __FILE__, __LINE__);
    auto kern = k_deep3<Eval, H, P, MODE, Epi>;
    static unsigned long long attr_set = 0ull;
    const unsigned long long dev_bit = 1ull << (g->device & 63);
    if (!(attr_set & dev_bit)) {
        PDE_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, T::SMEM));
        attr_set |= dev_bit;
    }
    int ntiles = geo.tiles_x * geo.tiles_y * geo.tiles_z;
    static const int ctas = getenv("PDE_K3_CTAS") ? std::max(1, atoi(getenv("PDE_K3_CTAS"))) : 2;
    int grid = std::min(ntiles, g->num_sms * ctas);
    // band points ride along in the brick loop on request (PDE_K3_FUSE_BAND=1) when the lattice
    // has enough bricks to spread them over
    Band3 band{0, g->d_band_center, g->d_band_ptr, g->d_band_j, g->d_band_w};
    if (band3_fused(g, ntiles)) band.n = g->n_band;
    kern<<<grid, K3_THREADS, T::SMEM, st>>>(tmap, g->tab, offs, geo, epi, red, done, band, S);
    PDE_LAUNCH_CHECK();
    return 0;
}

template <int MODE, class Epi>
static int launch_deep3(pde_grid *g, const double *S, const Epi &epi, double *red,
                        const unsigned long long *done, cudaStream_t st) {
    static const bool force_rt = getenv("PDE_K3_RUNTIME") && atoi(getenv("PDE_K3_RUNTIME")) != 0;
    if (!force_rt) {
        if (g->ct3_id == 0) return launch_deep3_h<EvalCt3<0>, 1, MODE, Epi>(g, S, epi, red, done, st);
        if (g->ct3_id == 1) return launch_deep3_h<EvalCt3<1>, 2, MODE, Epi>(g, S, epi, red, done, st);
    }
    if (g->tab.H == 1) return launch_deep3_h<EvalRuntime3, 1, MODE, Epi>(g, S, epi, red, done, st);
    if (g->tab.H == 2) return launch_deep3_h<EvalRuntime3, 2, MODE, Epi>(g, S, epi, red, done, st);
    return pde::fail("unsupported 3-D stencil halo", __FILE__, __LINE__);
}

template <int MODE, class Epi>
static int launch_deep(pde_grid *g, const double *S, const Epi &epi, double *red,
                       const unsigned long long *done, cudaStream_t st) {
    if (g->tab.D == 0) return 0;
    if (g->nmid > 0) return launch_deep3<MODE, Epi>(g, S, epi, red, done, st);
    if constexpr (std::is_same<Epi, EpiEigs<true, false, false>>::value && MODE == 0) {
        // tuning switch for the benchmark kernel (r = 4, d2min values): PDE_K1_VARIANT
        static const int variant = getenv("PDE_K1_VARIANT") ? atoi(getenv("PDE_K1_VARIANT")) : 0;
        if (g->ct_id == 3 && variant == 1)
            return launch_deep_hp<EvalCt<3>, 4, 4, MODE, Epi, 3>(g, S, epi, red, done, st);
        if (g->ct_id == 3 && variant == 2)
            return launch_deep_hp<EvalCt<3>, 4, 4, MODE, Epi, 4>(g, S, epi, red, done, st);
        if (g->ct_id == 3 && variant == 3)
            return launch_deep_hp<EvalCt<3>, 4, 8, MODE, Epi, 1>(g, S, epi, red, done, st);
        if (g->ct_id == 3 && variant == 4)
            return launch_deep_hp<EvalCt<3>, 4, 16, MODE, Epi, 1>(g, S, epi, red, done, st);
        if (g->ct_id == 3 && variant == 5)
            return launch_deep_hp<EvalCt<3>, 4, 6, MODE, Epi, 3>(g, S, epi, red, done, st);
        if (g->ct_id == 3 && variant == 6)       // integer compare-select: all pairs
            return launch_deep_hp<EvalCt<3>, 4, 8, MODE, Epi, 2, 4>(g, S, epi, red, done, st);
        if (g->ct_id == 3 && variant == 7)       // 3 of 4 pairs
            return launch_deep_hp<EvalCt<3>, 4, 8, MODE, Epi, 2, 3>(g, S, epi, red, done, st);
        if (g->ct_id == 3 && variant == 8)       // 2 of 4 pairs
            return launch_deep_hp<EvalCt<3>, 4, 8, MODE, Epi, 2, 2>(g, S, epi, red, done, st);
        if (g->ct_id == 3 && variant == 9)       // all pairs, 3 CTAs / SM of P = 4
            return launch_deep_hp<EvalCt<3>, 4, 4, MODE, Epi, 3, 4>(g, S, epi, red, done, st);
    }
    // compile-time table (square cells): fully unrolled evaluator with register windows
    switch (g->ct_id) {
        case 0: return launch_deep_hp<EvalCt<0>, 1, 8, MODE, Epi>(g, S, epi, red, done, st);
        case 1: return launch_deep_hp<EvalCt<1>, 2, 8, MODE, Epi>(g, S, epi, red, done, st);
        case 2: return launch_deep_hp<EvalCt<2>, 3, 8, MODE, Epi>(g, S, epi, red, done, st);
        case 3: return launch_deep_hp<EvalCt<3>, 4, 8, MODE, Epi>(g, S, epi, red, done, st);
        case 4: return launch_deep_hp<EvalCt<4>, 5, 8, MODE, Epi>(g, S, epi, red, done, st);
        default: break;
    }
    switch (g->tab.H) {      // run-time table (anisotropic cells, user tables)
        case 1: return launch_deep_hp<EvalRuntime, 1, 8, MODE, Epi>(g, S, epi, red, done, st);
        case 2: return launch_deep_hp<EvalRuntime, 2, 8, MODE, Epi>(g, S, epi, red, done, st);
        case 3: return launch_deep_hp<EvalRuntime, 3, 8, MODE, Epi>(g, S, epi, red, done, st);
        case 4: return launch_deep_hp<EvalRuntime, 4, 8, MODE, Epi>(g, S, epi, red, done, st);
        case 5: return launch_deep_hp<EvalRuntime, 5, 8, MODE, Epi>(g, S, epi, red, done, st);
        case 6: return launch_deep_hp<EvalRuntime, 6, 8, MODE, Epi>(g, S, epi, red, done, st);
    }
    return pde::fail("unsupported stencil halo", __FILE__, __LINE__);
}

template <int MODE, class Epi>
static int launch_band(pde_grid *g, const double *S, const Epi &epi, double *red,
                       const unsigned long long *done, cudaStream_t st) {
    if (g->n_band == 0) return 0;
    if (g->n_cls > 0 && g->win_lo == 0 && g->win_hi == g->nx_local && g->win2_hi == g->win2_lo) {
        const char *e = getenv("PDE_BAND_CLASSES");       // read per call: a test compares both
        if (!(e && atoi(e) == 0)) {
            int grid = (int)((g->n_band + 3) / 4);        // one warp per band point
            k_band_cls<MODE, Epi><<<grid, 128, 0, st>>>(g->n_band, g->d_band_center, g->d_cls_id,
                                                         g->d_cls_ptr, g->d_cls_ref, g->d_cls_w,
                                                         g->d_wl_ptr, g->d_wl, S, epi, red, done);
            PDE_LAUNCH_CHECK();
            return 0;
        }
    }
    // band points of the active row windows (band points are sorted by row)
    const int64_t *rb = g->h_band_row, *re = g->h_band_row + g->n_band;
    const int64_t wins[2][2] = {{g->win_lo, g->win_hi}, {g->win2_lo, g->win2_hi}};
    for (int wdx = 0; wdx < 2; ++wdx) {
        int64_t b_lo = std::lower_bound(rb, re, wins[wdx][0]) - rb;
        int64_t b_hi = std::lower_bound(rb, re, wins[wdx][1]) - rb;
        int64_t nbp = b_hi - b_lo;
        if (nbp <= 0) continue;
        int grid = (int)((nbp + 3) / 4);                 // one warp per band point
        k_band<MODE, Epi><<<grid, 128, 0, st>>>(nbp, g->d_band_center + b_lo, g->d_band_ptr + b_lo,
                                                 g->d_band_j, g->d_band_w, S, epi, red, done);
        PDE_LAUNCH_CHECK();
    }
    return 0;
}

template <class Epi>
static int launch_both(pde_grid *g, const double *S, const Epi &epi, int mode, double *red,
                       const unsigned long long *done, cudaStream_t st) {
    // The band kernel (explicit pair rows: latency / bandwidth bound, few registers) is
    // forked onto the context's side stream so that it shares the SMs with the fp64-bound
    // deep kernel instead of running after it; the two write disjoint points.
    int rc;
    if (g->nmid > 0 && g->tab.D > 0 && band3_fused(g, deep3_tiles(g))) {
        // 3-D: the band points are evaluated inside the brick loop of the deep kernel
        if (mode == 0) return launch_deep<0, Epi>(g, S, epi, red, done, st);
        return launch_deep<1, Epi>(g, S, epi, red, done, st);
    }
    const bool fork = g->n_band > 0 && g->tab.D > 0 && g->win_hi - g->win_lo > 64;
    cudaStream_t sb = fork ? g->side : st;
    if (fork) {
        PDE_CHECK(cudaEventRecord(g->ev_fork, st));
        PDE_CHECK(cudaStreamWaitEvent(sb, g->ev_fork, 0));
    }
    // deep first: its persistent CTAs fill the SMs, the band blocks slot into the registers
    // left over and into the tail of the last wave (measured 0.223 ms vs 0.244 ms band-first)
    if (mode == 0) {
        if ((rc = launch_deep<0, Epi>(g, S, epi, red, done, st))) return rc;
        if ((rc = launch_band<0, Epi>(g, S, epi, red, done, sb))) return rc;
    } else {
        if ((rc = launch_deep<1, Epi>(g, S, epi, red, done, st))) return rc;
        if ((rc = launch_band<1, Epi>(g, S, epi, red, done, sb))) return rc;
    }
    if (fork) {
        PDE_CHECK(cudaEventRecord(g->ev_join, sb));
        PDE_CHECK(cudaStreamWaitEvent(st, g->ev_join, 0));
    }
    return 0;
}

static int launch_wall(pde_grid *g, const double *W, const double *gg, double *out, double dt,
                       int mode, double *red, const unsigned long long *done, cudaStream_t st) {
    if (g->nb == 0 || g->win_lo != 0) return 0;      // wall entries go with the first window
    int grid = (int)std::min<int64_t>((g->nb + 255) / 256, 1024);
    k_wall<<<grid, 256, 0, st>>>(g->nb, g->rows * g->pitch, W, gg, out, dt, mode, red, done);
    PDE_LAUNCH_CHECK();
    return 0;
}

// ---- Euler loop control ------------------------------------------------------------------
// ctrl[0] = stop reason (0 running), ctrl[1] = iterations completed,
// red[0] = max|F|, red[1] = max|dU| of the iteration in flight; red[2], red[3] keep the
// values of the last completed iteration.
__global__ void k_euler_finalize(double *red, unsigned long long *ctrl, double sol_tol,
                                 double op_tol, unsigned long long max_iters) {
    if (ctrl[0]) return;
    double fmx = red[0], diff = red[1];
    red[2] = fmx;
    red[3] = diff;
    red[0] = 0.0;
    red[1] = 0.0;
    unsigned long long it = ctrl[1] + 1;
    ctrl[1] = it;
    if (diff < sol_tol && fmx < op_tol)      // solvers.py:84
        ctrl[0] = 1;
    else if (it >= max_iters)                // solvers.py:86
        ctrl[0] = 2;
}

}  // namespace pde

extern "C" {

int pde_ddg_d2eigs(pde_grid *g, const double *d_state, double *d_lmin, double *d_lmax,
                   uint16_t *d_pmin, uint16_t *d_pmax, int mode, void *stream) {
    PDE_REQUIRE(g && d_state, "null argument");
    DeviceGuard guard(g->device);
    cudaStream_t st = (cudaStream_t)stream;
    const bool want_min = d_lmin || d_pmin, want_max = d_lmax || d_pmax;
    const bool pol = d_pmin || d_pmax;
#define PDE_EIGS(MIN, MAX, POL)                                                              \
    return launch_both(g, d_state, EpiEigs<MIN, MAX, POL>{d_lmin, d_lmax, d_pmin, d_pmax}, mode, \
                       nullptr, nullptr, st)
    if (want_min && !want_max) {    // ddg.d2min: the max half of the reference's work is dead
        if (pol) PDE_EIGS(true, false, true);
        PDE_EIGS(true, false, false);
    }
    if (want_max && !want_min) {
        if (pol) PDE_EIGS(false, true, true);
        PDE_EIGS(false, true, false);
    }
    if (pol) PDE_EIGS(true, true, true);
    PDE_EIGS(true, true, false);
#undef PDE_EIGS
}

// layout conversions of a row range on the SMs (the copy engines carry the PCIe traffic)
__global__ void k_rows_in(const double *__restrict__ flat, double *__restrict__ state, int64_t lo,
                          int64_t hi, int64_t ny, int64_t pitch) {
    const int64_t total = (hi - lo) * ny;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = i / ny, c = i - r * ny;
        state[(lo + r) * pitch + c] = flat[(lo + r) * ny + c];
    }
}
__global__ void k_rows_out(const double *__restrict__ state, double *__restrict__ flat, int64_t lo,
                           int64_t hi, int64_t ny, int64_t pitch) {
    const int64_t total = (hi - lo) * ny;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = i / ny, c = i - r * ny;
        flat[(lo + r) * ny + c] = state[(lo + r) * pitch + c];
    }
}
__global__ void k_copy(const double *__restrict__ src, double *__restrict__ dst, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x)
        dst[i] = src[i];
}

// Host-to-host ddg.d2eigs: the whole H2D / stencil / D2H pipeline over row chunks enqueued by
// ONE call (the Python loop spent ~65 us of host time per chunk, which capped the chunk count).
// The PCIe copies are plain linear copies on their own streams (s_in, s_out) to / from flat
// staging buffers; the layout conversions (flat <-> state rows) are device-side 2-D copies on the
// compute stream, between the stencil launches.
int pde_ddg_d2eigs_host(pde_grid *g, const double *h_u, double *h_lmin, double *h_lmax, int mode,
                        int64_t chunk_rows, double *d_S, double *d_lmin, double *d_lmax,
                        double *d_fin, double *d_fout_min, double *d_fout_max, void *s_main,
                        void *s_in_, void *s_out_) {
    PDE_REQUIRE(g && h_u && d_S && d_fin && (h_lmin || h_lmax), "null argument");
    PDE_REQUIRE((!h_lmin || (d_lmin && d_fout_min)) && (!h_lmax || (d_lmax && d_fout_max)),
                "missing staging buffer");
    PDE_REQUIRE(g->halo_rows == 0 && g->nmid == 0, "host pipeline: unsharded 2-D lattices");
    DeviceGuard guard(g->device);
    cudaStream_t main = (cudaStream_t)s_main, s_in = (cudaStream_t)s_in_, s_out = (cudaStream_t)s_out_;
    const int64_t nloc = g->nx_local, ny = g->ny;
    if (chunk_rows < 16) chunk_rows = 16;
    const int nch = (int)((nloc + chunk_rows - 1) / chunk_rows);
    const int need_ev = 2 * nch + 2;
    if (g->n_ev_pool < need_ev) {
        cudaEvent_t *p = new cudaEvent_t[need_ev];
        for (int i = 0; i < g->n_ev_pool; ++i) p[i] = g->ev_pool[i];
        for (int i = g->n_ev_pool; i < need_ev; ++i)
            PDE_CHECK(cudaEventCreateWithFlags(&p[i], cudaEventDisableTiming));
        delete[] g->ev_pool;
        g->ev_pool = p;
        g->n_ev_pool = need_ev;
    }
    cudaEvent_t *ev_in = g->ev_pool, *ev_k = g->ev_pool + nch, ev_a = g->ev_pool[2 * nch],
                ev_b = g->ev_pool[2 * nch + 1];
    auto cgrid = [&](int64_t rows) {
        return (unsigned)std::min<int64_t>((rows * ny + 1023) / 1024, (int64_t)g->num_sms * 4);
    };
    auto lo_of = [&](int k) { return (int64_t)k * chunk_rows; };
    auto hi_of = [&](int k) { return std::min<int64_t>(nloc, (int64_t)(k + 1) * chunk_rows); };
    PDE_CHECK(cudaEventRecord(ev_a, main));                  // earlier work on the compute stream
    PDE_CHECK(cudaStreamWaitEvent(s_in, ev_a, 0));
    PDE_CHECK(cudaStreamWaitEvent(s_out, ev_a, 0));
    // PDE_E2E_TRACE=1: per-chunk completion times of the three streams on stderr (synchronises)
    const bool trace = getenv("PDE_E2E_TRACE") != nullptr;
    std::vector<cudaEvent_t> tev;
    if (trace) {
        tev.resize(3 * (size_t)nch + 1);
        for (auto &e : tev) PDE_CHECK(cudaEventCreate(&e));
        PDE_CHECK(cudaEventRecord(tev[3 * nch], s_in));
    }
    for (int k = 0; k < nch; ++k) {
        const int64_t lo = lo_of(k), hi = hi_of(k);
        PDE_CHECK(cudaMemcpyAsync(d_fin + lo * ny, h_u + lo * ny, (size_t)(hi - lo) * ny * 8,
                                  cudaMemcpyHostToDevice, s_in));
        if (k == 0 && g->nb > 0)
            PDE_CHECK(cudaMemcpyAsync(d_fin + nloc * ny, h_u + nloc * ny, (size_t)g->nb * 8,
                                      cudaMemcpyHostToDevice, s_in));
        PDE_CHECK(cudaEventRecord(ev_in[k], s_in));
        if (trace) PDE_CHECK(cudaEventRecord(tev[k], s_in));
    }
    int converted = 0, rc = 0;
    for (int k = 0; k < nch && !rc; ++k) {
        const int64_t lo = lo_of(k), hi = hi_of(k);
        const int need = std::min(k + 1, nch - 1);           // chunk k reads rows of chunk k + 1
        PDE_CHECK(cudaStreamWaitEvent(main, ev_in[need], 0));
        for (; converted <= need; ++converted) {
            const int64_t clo = lo_of(converted), chi = hi_of(converted);
            k_rows_in<<<cgrid(chi - clo), 256, 0, main>>>(d_fin, d_S, clo, chi, ny, g->pitch);
            PDE_LAUNCH_CHECK();
            if (converted == 0 && g->nb > 0) {
                k_copy<<<64, 256, 0, main>>>(d_fin + nloc * ny, d_S + g->rows * g->pitch, g->nb);
                PDE_LAUNCH_CHECK();
            }
        }
        pde_grid_set_row_window(g, lo, hi);
        rc = pde_ddg_d2eigs(g, d_S, h_lmin ? d_lmin : nullptr, h_lmax ? d_lmax : nullptr, nullptr,
                            nullptr, mode, main);
        if (rc) break;
        if (h_lmin) {
            k_rows_out<<<cgrid(hi - lo), 256, 0, main>>>(d_lmin, d_fout_min, lo, hi, ny, g->pitch);
            PDE_LAUNCH_CHECK();
        }
        if (h_lmax) {
            k_rows_out<<<cgrid(hi - lo), 256, 0, main>>>(d_lmax, d_fout_max, lo, hi, ny, g->pitch);
            PDE_LAUNCH_CHECK();
        }
        PDE_CHECK(cudaEventRecord(ev_k[k], main));
        if (trace) PDE_CHECK(cudaEventRecord(tev[nch + k], main));
        PDE_CHECK(cudaStreamWaitEvent(s_out, ev_k[k], 0));
        if (h_lmin)
            PDE_CHECK(cudaMemcpyAsync(h_lmin + lo * ny, d_fout_min + lo * ny, (size_t)(hi - lo) * ny * 8,
                                      cudaMemcpyDeviceToHost, s_out));
        if (h_lmax)
            PDE_CHECK(cudaMemcpyAsync(h_lmax + lo * ny, d_fout_max + lo * ny, (size_t)(hi - lo) * ny * 8,
                                      cudaMemcpyDeviceToHost, s_out));
        if (trace) PDE_CHECK(cudaEventRecord(tev[2 * nch + k], s_out));
    }
    pde_grid_set_row_window(g, 0, nloc);
    if (rc) return rc;
    PDE_CHECK(cudaEventRecord(ev_b, s_out));
    PDE_CHECK(cudaStreamWaitEvent(main, ev_b, 0));           // results are complete on `main`
    if (trace) {
        PDE_CHECK(cudaStreamSynchronize(main));
        fprintf(stderr, "chunk  h2d_done  kernels_done  d2h_done (ms)\n");
        for (int k = 0; k < nch; ++k) {
            float a = 0, b = 0, c = 0;
            cudaEventElapsedTime(&a, tev[3 * nch], tev[k]);
            cudaEventElapsedTime(&b, tev[3 * nch], tev[nch + k]);
            cudaEventElapsedTime(&c, tev[3 * nch], tev[2 * nch + k]);
            fprintf(stderr, "%3d  %8.3f  %8.3f  %8.3f\n", k, a, b, c);
        }
        for (auto &e : tev) cudaEventDestroy(e);
    }
    return 0;
}

int pde_ce_residual(pde_grid *g, const double *d_W, const double *d_g, double *d_F,
                    uint16_t *d_policy, double *d_red, void *stream) {
    PDE_REQUIRE(g && d_W && d_g && d_F, "null argument");
    DeviceGuard guard(g->device);
    cudaStream_t st = (cudaStream_t)stream;
    double *red = d_red ? d_red : g->d_red + 8;
    int rc = d_policy ? launch_both(g, d_W, EpiCeResidual<true>{d_g, d_F, d_policy}, 0, red, nullptr, st)
                      : launch_both(g, d_W, EpiCeResidual<false>{d_g, d_F, d_policy}, 0, red, nullptr, st);
    if (rc) return rc;
    return launch_wall(g, d_W, d_g, d_F, 0.0, 0, red, nullptr, st);
}

int pde_pucci_residual(pde_grid *g, const double *d_W, const double *d_f, double A0, double A1,
                       const double *d_A0, const double *d_A1, double *d_F, uint16_t *d_pmin,
                       uint16_t *d_pmax, double *d_red, void *stream) {
    PDE_REQUIRE(g && d_W && d_f && d_F, "null argument");
    DeviceGuard guard(g->device);
    cudaStream_t st = (cudaStream_t)stream;
    double *red = d_red ? d_red : g->d_red + 8;
    int rc = (d_pmin || d_pmax)
                 ? launch_both(g, d_W, EpiPucciResidual<true>{d_f, d_F, d_pmin, d_pmax, A0, A1, d_A0, d_A1},
                               0, red, nullptr, st)
                 : launch_both(g, d_W, EpiPucciResidual<false>{d_f, d_F, d_pmin, d_pmax, A0, A1, d_A0, d_A1},
                               0, red, nullptr, st);
    if (rc) return rc;
    return launch_wall(g, d_W, d_f, d_F, 0.0, 0, red, nullptr, st);
}

int pde_ce_euler_step(pde_grid *g, const double *d_U, const double *d_g, double dt,
                      double *d_Unew, double *d_red, void *stream) {
    PDE_REQUIRE(g && d_U && d_g && d_Unew && d_red, "null argument");
    DeviceGuard guard(g->device);
    cudaStream_t st = (cudaStream_t)stream;
    EpiCeEuler epi{d_g, d_Unew, dt};
    int rc = launch_both(g, d_U, epi, 0, d_red, nullptr, st);
    if (rc) return rc;
    return launch_wall(g, d_U, d_g, d_Unew, dt, 1, d_red, nullptr, st);
}

int pde_ce_euler(pde_grid *g, double *d_U0, double *d_U1, const double *d_g, double dt,
                 double solution_tol, double operator_tol, int64_t max_iters, int64_t check_every,
                 double *h_out, void *stream) {
    PDE_REQUIRE(g && d_U0 && d_U1 && d_g && h_out, "null argument");
    PDE_REQUIRE(max_iters >= 1, "max_iters must be >= 1");
    DeviceGuard guard(g->device);
    cudaStream_t st = (cudaStream_t)stream;
    if (check_every < 1) check_every = 64;
    PDE_CHECK(cudaMemsetAsync(g->d_red, 0, 4 * sizeof(double), st));
    PDE_CHECK(cudaMemsetAsync(g->d_ctrl, 0, 2 * sizeof(unsigned long long), st));
    unsigned long long h_ctrl[2] = {0, 0};
    double *buf[2] = {d_U0, d_U1};
    int64_t launched = 0;
    while (true) {
        int64_t n = std::min<int64_t>(check_every, max_iters - launched);
        for (int64_t i = 0; i < n; ++i, ++launched) {
            const double *Uin = buf[launched & 1];
            double *Uout = buf[(launched + 1) & 1];
            EpiCeEuler epi{d_g, Uout, dt};
            int rc = launch_both(g, Uin, epi, 0, g->d_red, g->d_ctrl, st);
            if (rc) return rc;
            if ((rc = launch_wall(g, Uin, d_g, Uout, dt, 1, g->d_red, g->d_ctrl, st))) return rc;
            k_euler_finalize<<<1, 1, 0, st>>>(g->d_red, g->d_ctrl, solution_tol, operator_tol,
                                              (unsigned long long)max_iters);
            PDE_LAUNCH_CHECK();
        }
        PDE_CHECK(cudaMemcpyAsync(h_ctrl, g->d_ctrl, sizeof h_ctrl, cudaMemcpyDeviceToHost, st));
        PDE_CHECK(cudaStreamSynchronize(st));
        if (h_ctrl[0] != 0 || launched >= max_iters) break;
    }
    double h_red[4];
    PDE_CHECK(cudaMemcpyAsync(h_red, g->d_red, sizeof h_red, cudaMemcpyDeviceToHost, st));
    PDE_CHECK(cudaStreamSynchronize(st));
    h_out[0] = (double)h_ctrl[1];
    h_out[1] = (double)(h_ctrl[1] & 1);
    h_out[2] = h_red[3];
    h_out[3] = h_red[2];
    h_out[4] = (double)h_ctrl[0];
    return 0;
}

}  // extern "C"
