// First-derivative operators.
//
// ddg.d1 / ddg.d1n / ddg.d1grad (ddg.py:8-175) work on the neighbour list of a grid or
// point cloud: per centre point, the upwind difference (u[J]-u[I])/|x_J-x_I| of the
// neighbour best aligned with a given direction (d1) or of the neighbour maximising the
// difference quotient itself (d1grad).  The reference does this with a lexsort over all
// neighbour rows and a Python dict lookup; here one thread walks the CSR row of its
// centre point.  Arithmetic follows the reference expression by expression (norm,
// normalisation, dot product, 1/h, product), so values are bit-identical; ties go to the
// first neighbour in row order (np.lexsort is stable, ddg.py:60,149).
//
// pde_rows_apply evaluates row-major 4-wide finite-difference rows
//   y[r] = sum_j w[r,j] * (x[idx[r,j]] - x[r + centre_offset])
// -- the matrix-free form of the 2- and 3-entry Jacobians the reference assembles with
// csr_matrix (ddg.py:73-79,164-171, ddi.py:84-90).
#include "common.cuh"

namespace pde {

__global__ void __launch_bounds__(128)
k_nb_d1(int64_t n_rows, int64_t centre_offset, const int64_t *__restrict__ ptr,
        const int32_t *__restrict__ J, const double *__restrict__ P, const double *__restrict__ u,
        const double *__restrict__ v, double *__restrict__ d1, int32_t *__restrict__ sel,
        double *__restrict__ wsel, double *__restrict__ vout) {
    int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    const int64_t c = r + centre_offset;
    const double px = P[2 * c], py = P[2 * c + 1];
    const double uc = u ? u[c] : 0.0;
    double vx = 0, vy = 0;
    if (v) {
        vx = v[2 * r];
        vy = v[2 * r + 1];
    }
    double best = __longlong_as_double(0xfff0000000000000LL);     // -inf
    int32_t bj = -1;
    double bw = 0, bd = 0, bx = 0, by = 0;
    for (int64_t e = ptr[r]; e < ptr[r + 1]; ++e) {
        int32_t j = J[e];
        double X = P[2 * (int64_t)j] - px, Y = P[2 * (int64_t)j + 1] - py;
        double h = sqrt(X * X + Y * Y);
        double w = 1.0 / h;
        double d = u ? w * (u[j] - uc) : 0.0;
        double key;
        if (v) key = (X / h) * vx + (Y / h) * vy;      // cosine against the direction (ddg.py:59)
        else key = d;                                  // difference quotient (ddg.py:148-149)
        if (key > best) {                              // strict: first maximal neighbour wins
            best = key; bj = j; bw = w; bd = d; bx = X; by = Y;
        }
    }
    if (d1) d1[r] = bd;
    sel[r] = bj;
    wsel[r] = bw;
    if (vout) {                                        // ddg.py:160: v = w * X
        vout[2 * r] = bw * bx;
        vout[2 * r + 1] = bw * by;
    }
}

__global__ void k_rows_apply(int64_t n_rows, int64_t centre_offset, const int4 *__restrict__ idx,
                             const double2 *__restrict__ w, const double *__restrict__ x,
                             double *__restrict__ y) {
    int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    int4 id = idx[r];
    double2 wa = w[2 * r], wb = w[2 * r + 1];
    double xc = x[r + centre_offset];
    y[r] = wa.x * (x[id.x] - xc) + wa.y * (x[id.y] - xc) + wb.x * (x[id.z] - xc) +
           wb.y * (x[id.w] - xc);
}

}  // namespace pde

using namespace pde;

extern "C" {

int pde_nb_d1(int device, int64_t n_rows, int64_t centre_offset, const int64_t *d_ptr,
              const int32_t *d_J, const double *d_points, const double *d_u, const double *d_v,
              double *d_d1, int32_t *d_sel, double *d_w, double *d_vout, void *stream) {
    PDE_REQUIRE(d_ptr && d_J && d_points && d_sel && d_w, "null argument");
    PDE_REQUIRE(d_u || d_v, "need a function (gradient mode) or a direction");
    if (n_rows == 0) return 0;
    DeviceGuard guard(device);
    k_nb_d1<<<(unsigned)((n_rows + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        n_rows, centre_offset, d_ptr, d_J, d_points, d_u, d_v, d_d1, d_sel, d_w, d_vout);
    PDE_LAUNCH_CHECK();
    return 0;
}

int pde_rows_apply(int device, int64_t n_rows, int64_t centre_offset, const int32_t *d_idx,
                   const double *d_w, const double *d_x, double *d_y, void *stream) {
    PDE_REQUIRE(d_idx && d_w && d_x && d_y, "null argument");
    PDE_REQUIRE((reinterpret_cast<uintptr_t>(d_idx) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(d_w) & 15) == 0, "row tables must be 16-byte aligned");
    if (n_rows == 0) return 0;
    DeviceGuard guard(device);
    k_rows_apply<<<(unsigned)((n_rows + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        n_rows, centre_offset, reinterpret_cast<const int4 *>(d_idx),
        reinterpret_cast<const double2 *>(d_w), d_x, d_y);
    PDE_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
