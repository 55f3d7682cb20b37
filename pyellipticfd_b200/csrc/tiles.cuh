// Host-side helpers shared by the tiled kernels (K1 in ddg_kernels.cu, the tiled Jacobian rows in
// bicgstab.cu): the deep-interior region of a row window and the 2-D TMA descriptor of a
// state-layout array.
#pragma once
#include <algorithm>

#include "common.cuh"

namespace pde {

struct DeepRegion {
    int row0, col0, nrows, ncols;   // buffer coordinates
};

inline DeepRegion deep_region(const pde_grid *g, int64_t win_lo, int64_t win_hi) {
    int64_t r = g->radius;
    int64_t gx0 = std::max<int64_t>(r, g->x_offset + win_lo);
    int64_t gx1 = std::min<int64_t>(g->nx_global - r, g->x_offset + win_hi);
    DeepRegion d;
    d.row0 = (int)(gx0 - g->x_offset + g->halo_rows);
    d.nrows = (int)std::max<int64_t>(0, gx1 - gx0);
    d.col0 = (int)r;
    d.ncols = (int)std::max<int64_t>(0, g->ny - 2 * r);
    return d;
}

inline int make_tmap(const pde_grid *g, const double *S, int boxy, int boxx, CUtensorMap *out) {
    PDE_REQUIRE((reinterpret_cast<uintptr_t>(S) & 15) == 0, "state buffer must be 16-byte aligned");
    cuuint64_t gdim[2] = {(cuuint64_t)g->ny, (cuuint64_t)g->rows};
    cuuint64_t gstr[1] = {(cuuint64_t)g->pitch * 8};
    cuuint32_t box[2] = {(cuuint32_t)boxy, (cuuint32_t)boxx};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g->encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double *>(S), gdim,
                           gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        char buf[128];
        snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
        return pde::fail(buf, __FILE__, __LINE__);
    }
    return 0;
}


}  // namespace pde
