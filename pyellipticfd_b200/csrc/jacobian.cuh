// Matrix-free rows of the policy Jacobian.
//
// The reference assembles, every Newton iteration, a SciPy CSR matrix whose
// interior rows are the 3-point finite-difference rows of the selected pair
// (ddg.py:304-323) combined as  Jac = I, Jac[active] = -M_min
// (convex_envelope.py:48-54)  or  Jac[interior] = A0*M_max + A1*M_min
// (pucci.py:50-54, :136-139).  Here a row is (policy index, coefficients): the
// entries are recomputed from the direction table / band weights on the fly.
#pragma once
#include "common.cuh"

namespace pde {

struct JacDev {
    const uint16_t *pmin;
    const uint16_t *pmax;     // may be NULL
    double cmin_s, cmax_s;
    const double *cmin;       // may be NULL -> scalar
    const double *cmax;
};

struct GridDev {
    int64_t pitch, rows, ny, nb;
    int64_t halo_rows, x_offset, nx_global, nx_local;
    int radius;
    int64_t nmid, nx3;          // 3-D lattices: rows = nx3 * nmid (0, 0 in 2-D)
    int64_t n_band;
    const int64_t *band_center, *band_ptr, *band_j;
    const double *band_w;
};

inline GridDev grid_dev(const pde_grid *g) {
    GridDev d;
    d.pitch = g->pitch;
    d.rows = g->rows;
    d.ny = g->ny;
    d.nb = g->nb;
    d.halo_rows = g->halo_rows;
    d.x_offset = g->x_offset;
    d.nx_global = g->nx_global;
    d.nx_local = g->nx_local;
    d.radius = g->radius;
    d.nmid = g->nmid;
    d.nx3 = g->nx3;
    d.n_band = g->n_band;
    d.band_center = g->d_band_center;
    d.band_ptr = g->d_band_ptr;
    d.band_j = g->d_band_j;
    d.band_w = g->d_band_w;
    return d;
}

inline JacDev jac_dev(const pde_jac *J) {
    JacDev d;
    d.pmin = J->d_pmin;
    d.pmax = J->d_pmax;
    d.cmin_s = J->cmin_s;
    d.cmax_s = J->cmax_s;
    d.cmin = J->d_cmin;
    d.cmax = J->d_cmax;
    return d;
}

// 3-point row entries of one pair: (A0, A1, Ac) as the reference stores them
// (ddg.py:313-317: w0*w_0, w0*w_1, w0*(-w_0 + -w_1)).
__device__ __forceinline__ void row_entries(double w_0, double w_1, double w0, double &A0,
                                            double &A1, double &Ac) {
    A0 = __dmul_rn(w0, w_0);
    A1 = __dmul_rn(w0, w_1);
    Ac = __dmul_rn(w0, __dadd_rn(-w_0, -w_1));
}

// is global lattice row gx a row of deep-interior points (columns [radius, ny - radius))?
// 2-D: rows [r, nx - r); 3-D: row = ix * nmid + iy with both ix and iy r away from the walls
__device__ __forceinline__ bool deep_row(const GridDev &G, int64_t gx) {
    if (G.nmid > 0) {
        int64_t ix = gx / G.nmid, iy = gx - ix * G.nmid;
        return ix >= G.radius && ix < G.nx3 - G.radius && iy >= G.radius && iy < G.nmid - G.radius;
    }
    return gx >= G.radius && gx < G.nx_global - G.radius;
}

// (J x)_o for a deep-interior point
__device__ __forceinline__ double jac_row_deep(const DeepTable &tab, const JacDev &J,
                                               const double *__restrict__ x, int64_t o,
                                               int64_t pitch) {
    unsigned pm = J.pmin[o];
    double xc = x[o];
    if (pm == PDE_POLICY_IDENTITY) return xc;
    // difference form w0*((x0-xc)*w_0 + (x1-xc)*w_1): the same row as the reference's
    // stored entries (ddg.py:313-317) up to rounding, with an exactly zero row sum, and
    // bit-identical to the operator value of that pair (ddg.py:281).
    double m = pair_value_exact(xc, x[o + tab.a0[pm] * pitch + tab.b0[pm]],
                                x[o + tab.a1[pm] * pitch + tab.b1[pm]], tab.w_0[pm], tab.w_1[pm],
                                tab.w0[pm]);
    double cm = J.cmin ? J.cmin[o] : J.cmin_s;
    double y = cm * m;
    if (J.pmax) {
        unsigned px = J.pmax[o];
        double mm = pair_value_exact(xc, x[o + tab.a0[px] * pitch + tab.b0[px]],
                                     x[o + tab.a1[px] * pitch + tab.b1[px]], tab.w_0[px],
                                     tab.w_1[px], tab.w0[px]);
        double cx = J.cmax ? J.cmax[o] : J.cmax_s;
        y += cx * mm;
    }
    return y;
}

__device__ __forceinline__ double jac_diag_deep(const DeepTable &tab, const JacDev &J, int64_t o) {
    unsigned pm = J.pmin[o];
    if (pm == PDE_POLICY_IDENTITY) return 1.0;
    double A0, A1, Ac;
    row_entries(tab.w_0[pm], tab.w_1[pm], tab.w0[pm], A0, A1, Ac);
    double cm = J.cmin ? J.cmin[o] : J.cmin_s;
    double y = cm * Ac;
    if (J.pmax) {
        unsigned px = J.pmax[o];
        row_entries(tab.w_0[px], tab.w_1[px], tab.w0[px], A0, A1, Ac);
        double cx = J.cmax ? J.cmax[o] : J.cmax_s;
        y += cx * Ac;
    }
    return y;
}

// band point number bi
__device__ __forceinline__ double jac_row_band(const GridDev &G, const JacDev &J,
                                               const double *__restrict__ x, int64_t bi) {
    int64_t o = G.band_center[bi];
    unsigned pm = J.pmin[o];
    double xc = x[o];
    if (pm == PDE_POLICY_IDENTITY) return xc;
    int64_t k = G.band_ptr[bi] + pm;
    double m = pair_value_exact(xc, x[G.band_j[2 * k]], x[G.band_j[2 * k + 1]], G.band_w[3 * k],
                                G.band_w[3 * k + 1], G.band_w[3 * k + 2]);
    double cm = J.cmin ? J.cmin[o] : J.cmin_s;
    double y = cm * m;
    if (J.pmax) {
        k = G.band_ptr[bi] + J.pmax[o];
        double mm = pair_value_exact(xc, x[G.band_j[2 * k]], x[G.band_j[2 * k + 1]],
                                     G.band_w[3 * k], G.band_w[3 * k + 1], G.band_w[3 * k + 2]);
        double cx = J.cmax ? J.cmax[o] : J.cmax_s;
        y += cx * mm;
    }
    return y;
}

// band row with the source values supplied by `get(offset)` (fused passes recompute them)
template <class Get>
__device__ __forceinline__ double jac_row_band_get(const GridDev &G, const JacDev &J, int64_t bi,
                                                   double xc, Get get) {
    int64_t o = G.band_center[bi];
    unsigned pm = J.pmin[o];
    if (pm == PDE_POLICY_IDENTITY) return xc;
    int64_t k = G.band_ptr[bi] + pm;
    double m = pair_value_exact(xc, get(G.band_j[2 * k]), get(G.band_j[2 * k + 1]), G.band_w[3 * k],
                                G.band_w[3 * k + 1], G.band_w[3 * k + 2]);
    double cm = J.cmin ? J.cmin[o] : J.cmin_s;
    double y = cm * m;
    if (J.pmax) {
        k = G.band_ptr[bi] + J.pmax[o];
        double mm = pair_value_exact(xc, get(G.band_j[2 * k]), get(G.band_j[2 * k + 1]),
                                     G.band_w[3 * k], G.band_w[3 * k + 1], G.band_w[3 * k + 2]);
        double cx = J.cmax ? J.cmax[o] : J.cmax_s;
        y += cx * mm;
    }
    return y;
}

__device__ __forceinline__ double jac_diag_band(const GridDev &G, const JacDev &J, int64_t bi) {
    int64_t o = G.band_center[bi];
    unsigned pm = J.pmin[o];
    if (pm == PDE_POLICY_IDENTITY) return 1.0;
    int64_t k = G.band_ptr[bi] + pm;
    double A0, A1, Ac;
    row_entries(G.band_w[3 * k], G.band_w[3 * k + 1], G.band_w[3 * k + 2], A0, A1, Ac);
    double cm = J.cmin ? J.cmin[o] : J.cmin_s;
    double y = cm * Ac;
    if (J.pmax) {
        k = G.band_ptr[bi] + J.pmax[o];
        row_entries(G.band_w[3 * k], G.band_w[3 * k + 1], G.band_w[3 * k + 2], A0, A1, Ac);
        double cx = J.cmax ? J.cmax[o] : J.cmax_s;
        y += cx * Ac;
    }
    return y;
}

// Work decomposition shared by every "one pass over the state" kernel:
//   blocks [0, rows)             -> one lattice row each (threads stride over columns)
//   blocks [rows, rows + nbw)    -> band points, then wall entries, 256 per block
__host__ __device__ inline int64_t tail_blocks(int64_t n_band, int64_t nb) {
    return (n_band + nb + 255) / 256;
}

}  // namespace pde
