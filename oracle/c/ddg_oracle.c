/*
 * Plain-C oracle for the regular-grid operator -- TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * Restates /root/reference/ddg.py:274-292 for (a) an explicit pair list and (b)
 * the deep interior of a lattice, recomputing every weight from the point
 * coordinates per pair exactly as the reference does on every call:
 *
 *     X = points[J] - points[I];  h = sqrt(Xx*Xx + Xy*Xy);  w = 1/h;
 *     w0 = 2/(h0+h1);  d = w0*((u[J0]-u[I])*w_0 + (u[J1]-u[I])*w_1)
 *
 * with coordinates points = idx*scaling + bounds0 (pointclasses.py:551).
 * Compile with -ffp-contract=off (see Makefile): no FMA, IEEE order.
 * Selection: scan in row order, `<=` for the minimum (last row among ties),
 * `>` for the maximum (first row) == stable lexsort((-d, I)) of ddg.py:283-292.
 */
#include <math.h>
#include <stdint.h>

static inline double pair_value(double xi, double yi, double ui, double x0, double y0, double u0,
                                double x1, double y1, double u1)
{
    double X0 = x0 - xi, Y0 = y0 - yi, X1 = x1 - xi, Y1 = y1 - yi;
    double h0 = sqrt(X0 * X0 + Y0 * Y0), h1 = sqrt(X1 * X1 + Y1 * Y1);
    double w_0 = 1 / h0, w_1 = 1 / h1, w0 = 2 / (h0 + h1);
    double t0 = (u0 - ui) * w_0, t1 = (u1 - ui) * w_1;
    return w0 * (t0 + t1);
}

/* explicit pair rows (I, J0, J1), any order; outputs indexed by interior point */
void ddg_pairs_d2eigs(const double *points, const int64_t *pairs, int64_t npairs, const double *u,
                      int64_t nint, double *lmin, double *lmax, int64_t *rmin, int64_t *rmax)
{
    for (int64_t i = 0; i < nint; ++i) {
        lmin[i] = INFINITY;
        lmax[i] = -INFINITY;
        rmin[i] = rmax[i] = -1;
    }
    for (int64_t k = 0; k < npairs; ++k) {
        int64_t I = pairs[3 * k], J0 = pairs[3 * k + 1], J1 = pairs[3 * k + 2];
        double d = pair_value(points[2 * I], points[2 * I + 1], u[I], points[2 * J0],
                              points[2 * J0 + 1], u[J0], points[2 * J1], points[2 * J1 + 1], u[J1]);
        if (d <= lmin[I]) { lmin[I] = d; rmin[I] = k; }
        if (d > lmax[I]) { lmax[I] = d; rmax[I] = k; }
    }
}

/* deep interior of an nx x ny lattice (interior ids ix*ny+iy, 0-based; lattice
 * coordinate = index+1).  table: D rows (a0,b0,a1,b1).  Rows [row_lo,row_hi) only.
 * Outputs are full-size (nx*ny) arrays; non-deep entries are left untouched. */
void ddg_lattice_d2eigs(int64_t nx, int64_t ny, double sx, double sy, double bx, double by, int r,
                        int D, const int32_t *table, const double *u, int64_t row_lo,
                        int64_t row_hi, double *lmin, double *lmax, int32_t *kmin, int32_t *kmax)
{
    if (row_lo < r) row_lo = r;
    if (row_hi > nx - r) row_hi = nx - r;
    for (int64_t ix = row_lo; ix < row_hi; ++ix) {
        for (int64_t iy = r; iy < ny - r; ++iy) {
            int64_t I = ix * ny + iy;
            double xi = (double)(ix + 1) * sx + bx, yi = (double)(iy + 1) * sy + by;
            double mn = INFINITY, mx = -INFINITY;
            int32_t an = 0, ax = 0;
            for (int k = 0; k < D; ++k) {
                const int32_t *t = table + 4 * k;
                int64_t J0 = (ix + t[0]) * ny + iy + t[1], J1 = (ix + t[2]) * ny + iy + t[3];
                double d = pair_value(xi, yi, u[I], (double)(ix + t[0] + 1) * sx + bx,
                                      (double)(iy + t[1] + 1) * sy + by, u[J0],
                                      (double)(ix + t[2] + 1) * sx + bx,
                                      (double)(iy + t[3] + 1) * sy + by, u[J1]);
                if (d <= mn) { mn = d; an = k; }
                if (d > mx) { mx = d; ax = k; }
            }
            lmin[I] = mn; lmax[I] = mx; kmin[I] = an; kmax[I] = ax;
        }
    }
}
