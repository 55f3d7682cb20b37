/*
 * pde_b200.h -- C ABI of the B200-native (sm_100a) pyellipticfd hot path.
 *
 * The reference (cfinlay/pyellipticfd) is pure Python and has no FFI layer; the
 * boundary it offers is the set of module-level functions ddg/ddi/ddf.d2eigs,
 * convex_envelope/pucci.operator+solve and solvers.euler/NewtonEulerLS.  This
 * header is the C-level contract those functions are re-implemented on: plain
 * pointers and sizes, no torch / Python types.  Every entry point cites the
 * reference code it replaces (paths relative to the reference root).
 *
 * Conventions
 *  - every function returns 0 on success, non-zero on failure; the message of
 *    the last failure on the calling thread is pde_last_error().
 *  - all `double*`, `uint16_t*`, `int32_t*` arguments named d_* are DEVICE
 *    pointers on the context's device; h_* are HOST pointers.
 *  - `stream` is a cudaStream_t passed as void* (0 = legacy default stream).
 *    Calls are asynchronous on that stream unless stated otherwise.
 *
 * State layout ("state vector", length pde_grid_state_len()):
 *    [ rows x pitch doubles : lattice rows (halo rows, owned rows, halo rows),
 *                             y fastest, pitch = ny rounded up to 16, padding = 0 ]
 *    [ nb doubles           : wall (boundary) points in the grid's order ]
 *  EVERY per-point array of this ABI (values, residuals, policies, Krylov
 *  vectors) uses this one layout, so one offset addresses them all.
 *  The reference's flat vector u (pointclasses.py:513-515,540: interior points
 *  first in C order, then boundary points) maps to it by pde_grid_pack/unpack.
 */
#ifndef PDE_B200_H
#define PDE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PDE_B200_ABI_VERSION 1
#define PDE_MAX_DIRS 64          /* antipodal pairs per deep-interior point  */
#define PDE_POLICY_IDENTITY 0xFFFFu /* policy value meaning "identity row"   */

typedef struct pde_grid pde_grid;    /* regular pairs grid (ddg)            */
typedef struct pde_ell pde_ell;      /* point-cloud ELL operator (ddi, ddf) */

int pde_abi_version(void);
const char *pde_last_error(void);
/* number of CUDA kernels this library has launched since load (bench.py's
 * gpu_launches claim is read from here). */
int64_t pde_kernel_launches(void);

/* ------------------------------------------------------------------------
 * Regular-grid context.  Replaces the per-call geometry work of
 * ddg.d2eigs (ddg.py:274-280: X, h, w, w0 recomputed from coordinates on
 * every call) by one device-resident description of the grid:
 *   - deep interior: D antipodal offset pairs (a0,b0,a1,b1) with per-direction
 *     weights (w_0, w_1, w0) -- valid when the spacing makes them position
 *     independent (FDRegularGrid.uniform_weights());
 *   - band: explicit pair rows (I, J0, J1) with per-row weights, grouped by
 *     band point (CSR); I/J are given as offsets into the state layout below
 *     (the host converts the reference's global point ids).
 * h_* arrays are copied; the caller may free them after the call.
 * ---------------------------------------------------------------------- */
int pde_grid_create(pde_grid **out, int device,
                    int64_t nx_global, int64_t ny, int64_t nb, int stencil_radius,
                    /* slab owned by this context: rows [x_offset, x_offset+nx_local)
                     * of the global lattice plus halo_rows extra rows on each side
                     * (single GPU: 0, nx_global, 0) */
                    int64_t x_offset, int64_t nx_local, int64_t halo_rows,
                    int D, const int32_t *h_stencil_pairs /* D x 4 */,
                    const double *h_stencil_w /* D x 3 : w_0, w_1, w0 */,
                    const double *h_stencil_X /* D x 2 : physical vector of (a0,b0) */,
                    int64_t n_band,
                    const int64_t *h_band_center /* n_band : state offset of I */,
                    const int64_t *h_band_ptr /* n_band+1 */,
                    const int64_t *h_band_j /* P x 2 : state offsets of J0, J1 */,
                    const double *h_band_w /* P x 3 : w_0, w_1, w0 */,
                    const double *h_band_X /* P x 2 : points[J0]-points[I] */);
/* 3-D regular pairs grid (the 3-D stencils of stencils.py:46-61 behind the dimension-agnostic
 * ddg.d2eigs, ddg.py:251-292): lattice [nx][ny][nz] = h_shape, stored as nx*ny state-layout rows
 * of nz columns, so every other entry point (pack/unpack, d2eigs, residuals, Euler, Jacobian
 * rows, BiCGStab) applies unchanged.  Deep points: D antipodal offset pairs
 * (a0,b0,c0,a1,b1,c1) with per-direction weights, evaluated by the 3-D TMA kernel (stencil
 * halo 1 or 2, D <= PDE_MAX_DIRS); band points: explicit rows as in pde_grid_create. */
int pde_grid_create_nd(pde_grid **out, int device, int dim /* 3 */,
                       const int64_t *h_shape /* dim */, int64_t nb, int stencil_radius,
                       int D, const int32_t *h_stencil_pairs /* D x 2*dim */,
                       const double *h_stencil_w /* D x 3 */,
                       int64_t n_band, const int64_t *h_band_center, const int64_t *h_band_ptr,
                       const int64_t *h_band_j /* P x 2 */, const double *h_band_w /* P x 3 */);
/* Optional compact form of the band rows for the operator kernels (3-D lattices, where the band
 * is a surface layer as expensive as the deep points): band points with the same distances to
 * the walls share their rows up to translation.  Class c owns rows [h_cls_ptr[c], h_cls_ptr[c+1]);
 * a row is two references -- (state-offset delta << 1) for a lattice neighbour, (q << 1 | 1) for
 * the q-th entry of the band point's own wall list h_wl[h_wl_ptr[i] ..] (state offsets) -- and the
 * weights (w_0, w_1, w0), in the order of the explicit rows (policies keep their meaning).  The
 * explicit rows given to pde_grid_create* stay in use for the Jacobian rows. */
int pde_grid_set_band_classes(pde_grid *g, int n_cls, const int32_t *h_cls_ptr,
                              const int32_t *h_cls_ref /* rows x 2 */,
                              const double *h_cls_w /* rows x 3 */,
                              const uint16_t *h_cls_id /* n_band */,
                              const int32_t *h_wl_ptr /* n_band + 1 */, const int32_t *h_wl);
int pde_grid_destroy(pde_grid *g);
/* Restrict the following stencil launches (pde_ddg_d2eigs, pde_ce_residual,
 * pde_ce_euler_step, pde_pucci_residual) to owned rows [row_lo, row_hi); wall entries
 * are handled by the window that starts at row 0.  (0, nx_local) restores the full
 * slab.  Used to overlap the halo exchange (edge rows last) and host<->device copies
 * (row chunks) with the arithmetic. */
int pde_grid_set_row_window(pde_grid *g, int64_t row_lo, int64_t row_hi);
/* Two disjoint row ranges [lo, hi) and [lo2, hi2) (lo <= hi <= lo2 <= hi2) served by ONE set of
 * launches: the two edge strips of a slab after its halo exchange (multigpu.py), which would
 * otherwise cost a second, almost empty launch group per operator application. */
int pde_grid_set_row_windows2(pde_grid *g, int64_t lo, int64_t hi, int64_t lo2, int64_t hi2);
/* on != 0: the persistent CTAs of the stencil kernel draw their tiles from a device counter
 * instead of a static round robin.  Used by the slab-sharded path, where the NCCL kernel of the
 * overlapped halo exchange delays some CTAs (it cannot share an SM with two 125-register CTAs);
 * launches on one context must then not overlap each other (they never do: one stream). */
int pde_grid_set_dynamic_schedule(pde_grid *g, int on);
int64_t pde_grid_pitch(const pde_grid *g);      /* doubles per lattice row        */
int64_t pde_grid_rows(const pde_grid *g);       /* nx_local + 2*halo_rows         */
int64_t pde_grid_state_len(const pde_grid *g);  /* rows*pitch + nb                */
int64_t pde_grid_policy_len(const pde_grid *g); /* rows*pitch (uint16 entries)    */

/* flat reference vector (N + nb) <-> state layout.  Both on device. */
int pde_grid_pack(pde_grid *g, const double *d_u_flat, double *d_state, void *stream);
int pde_grid_unpack(pde_grid *g, const double *d_state, double *d_u_flat, void *stream);
/* The same conversion as one 2-D DMA copy between HOST memory (h_flat: the reference's
 * flat vector, pinned for asynchronous copies) and the state layout, for owned rows
 * [row_lo, row_hi) only; with_wall also copies the nb wall values.  Together with
 * pde_grid_set_row_window this pipelines H2D / stencil / D2H over row chunks (the
 * state buffer's padding columns must already be zero).  The "host" pointer may also be a
 * DEVICE pointer (unified addressing): the copy is then a device-side 2-D copy, which lets
 * the PCIe transfer itself be a plain linear copy into a flat staging buffer. */
int pde_grid_h2d_rows(pde_grid *g, const double *h_flat, double *d_state,
                      int64_t row_lo, int64_t row_hi, int with_wall, void *stream);
int pde_grid_d2h_rows(pde_grid *g, const double *d_state, double *h_flat_interior,
                      int64_t row_lo, int64_t row_hi, void *stream);
/* interior part of a state-layout array -> flat (N) interior order */
int pde_grid_unpad(pde_grid *g, const double *d_padded, double *d_flat, void *stream);
int pde_grid_unpad_u16(pde_grid *g, const uint16_t *d_padded, int32_t *d_flat, void *stream);

/* ddg.d2eigs (ddg.py:251-327) values + policy.  Any output may be NULL.
 * d_lmin/d_lmax: state layout.  d_pmin/d_pmax (state layout, uint16): selected
 * pair per interior point: deep points -> row of the stencil table, band
 * points -> position inside the point's band_pairs group.  Tie rule as
 * ddg.py:283-292: min = last pair attaining it, max = first.
 * mode 0 = reference arithmetic order (bit-exact), 1 = fast re-associated. */
int pde_ddg_d2eigs(pde_grid *g, const double *d_state,
                   double *d_lmin, double *d_lmax,
                   uint16_t *d_pmin, uint16_t *d_pmax, int mode, void *stream);

/* Selected-pair finite-difference rows applied matrix-free
 * (replaces the CSR assembly ddg.py:304-323):
 *   y[i] = sum over the policy's pair at i of  w0*w_0*x[J0] + w0*w_1*x[J1]
 *          - w0*(w_0+w_1)*x[i]            for interior i;  y = x on the wall. */
int pde_ddg_apply_policy(pde_grid *g, const uint16_t *d_policy,
                         const double *d_x, double *d_y, void *stream);
/* control directions of ddg.py:297-300 (quirk: both scaled by the min pair's
 * 1/h): d_vmin, d_vmax are (N,2) flat. */
int pde_ddg_control(pde_grid *g, const uint16_t *d_pmin, const uint16_t *d_pmax,
                    double *d_vmin, double *d_vmax, void *stream);
/* ------------------------------------------------------------------------
 * Jacobian of a policy, matrix-free.  Replaces convex_envelope.py:48-54
 * (Jac = I; Jac[active] = -M_min[active]) and pucci.py:50-54,136-139
 * (Jac[interior] = A0*M_max + A1*M_min).
 *   interior row i : y = cmax_i * (M_pmax x)_i + cmin_i * (M_pmin x)_i,
 *                    or y = x_i when d_pmin[i] == PDE_POLICY_IDENTITY
 *   wall rows      : y = x
 * cmin/cmax are scalars when d_cmin/d_cmax are NULL (then cmin_s/cmax_s are
 * used), otherwise per-interior-point arrays in padded layout.
 * x, y: state layout.  transpose != 0 applies the transposed operator
 * (solvers.py:192, Jac.transpose().dot). */
typedef struct pde_jac {
    const uint16_t *d_pmin;   /* padded layout, may hold PDE_POLICY_IDENTITY */
    const uint16_t *d_pmax;   /* NULL when cmax == 0 (convex envelope)       */
    double cmin_s, cmax_s;
    const double *d_cmin, *d_cmax;
} pde_jac;
int pde_jac_matvec(pde_grid *g, const pde_jac *J, const double *d_x, double *d_y,
                   int transpose, void *stream);
int pde_jac_diagonal(pde_grid *g, const pde_jac *J, double *d_diag, void *stream);

/* ddg.d2eigs / d2min / d2max (ddg.py:251-339) for a field in PINNED host memory with the results
 * in pinned host memory (the reference's arrays live on the host): H2D copy, stencil kernels and
 * D2H copy pipelined over chunks of `chunk_rows` lattice rows on three streams, enqueued by one
 * call.  h_u: the reference's flat vector (interior points row-major, then the wall points);
 * h_lmin / h_lmax (either may be NULL): values over the interior points.  Staging: d_S, d_lmin,
 * d_lmax state-layout vectors (pde_grid_state_len doubles), d_fin nx*ny + nb doubles,
 * d_fout_* nx*ny doubles.  On return the work is enqueued; the results are complete once
 * s_main is.  Unsharded 2-D lattices. */
int pde_ddg_d2eigs_host(pde_grid *g, const double *h_u, double *h_lmin, double *h_lmax, int mode,
                        int64_t chunk_rows, double *d_S, double *d_lmin, double *d_lmax,
                        double *d_fin, double *d_fout_min, double *d_fout_max, void *s_main,
                        void *s_in, void *s_out);
/* ------------------------------------------------------------------------
 * Convex envelope, convex_envelope.py:8-56 / :58-123.
 * residual: F = W-g everywhere; interior rows with -lambda_min > W-g are
 * replaced by -lambda_min and marked active (policy = selected pair, else
 * PDE_POLICY_IDENTITY).  d_F: state layout.  d_policy may be NULL.
 * d_red (2 doubles, device): [max|F|, unused] accumulated with max. */
int pde_ce_residual(pde_grid *g, const double *d_W, const double *d_g,
                    double *d_F, uint16_t *d_policy, double *d_red, void *stream);

/* Device-resident explicit Euler loop (solvers.py:13-93 specialised to the
 * convex-envelope operator): runs up to max_iters iterations of
 * U <- U - dt*F(U) without host round trips, stopping at the first iteration
 * where max|U_new-U| < solution_tol and max|F(U)| < operator_tol.
 * d_U0/d_U1: two state buffers (ping-pong), d_U0 holds the initial iterate.
 * h_out: [iterations, which buffer holds the result (0/1), diff, max|F|,
 *         stop reason (1 converged, 2 max_iters)] -- synchronises the stream. */
int pde_ce_euler(pde_grid *g, double *d_U0, double *d_U1, const double *d_g,
                 double dt, double solution_tol, double operator_tol,
                 int64_t max_iters, int64_t check_every, double *h_out, void *stream);

/* One explicit Euler step of the same operator (the building block of the
 * multi-GPU loop, where a halo exchange sits between steps):
 * Unew = U - dt*F(U); d_red[0] = max(d_red[0], max|F|), d_red[1] likewise max|U-Unew|.
 * Only owned entries of d_Unew are written. */
int pde_ce_euler_step(pde_grid *g, const double *d_U, const double *d_g, double dt,
                      double *d_Unew, double *d_red, void *stream);

/* Pucci, pucci.py:7-58 / :123-141:  F = A0*lambda_max + A1*lambda_min - f on
 * the interior, W - dirichlet on the wall.  d_f: state layout (interior f,
 * wall g).  A0/A1 scalar, or arrays when d_A0/d_A1 non-NULL (padded). */
int pde_pucci_residual(pde_grid *g, const double *d_W, const double *d_f,
                       double A0, double A1, const double *d_A0, const double *d_A1,
                       double *d_F, uint16_t *d_pmin, uint16_t *d_pmax,
                       double *d_red, void *stream);

/* ------------------------------------------------------------------------
 * Jacobi-preconditioned BiCGStab, matrix-free on a pde_jac.  Replaces
 * solvers.py:179-181 (P = diags(1/Jac.diagonal()); bicgstab(Jac, -Gu, M=P,
 * tol=...)) and follows SciPy 1.18.1 scipy/sparse/linalg/_isolve/iterative.py
 * `bicgstab` step for step: x0 = 0, stop when ||r||_2 < max(atol, rtol*||b||),
 * tested at the top of each iteration and after the half step; breakdown
 * exits -10 (rho) / -11 (omega, <rtilde,v> == 0); maxiter default 10*n.
 * d_b, d_x: state layout.  d_work: 8 * state_len doubles of scratch.
 * h_info[0] = SciPy's `info` (0 ok, >0 = maxiter reached, <0 breakdown),
 * h_info[1] = iterations performed.  Synchronises the stream. */
int pde_bicgstab(pde_grid *g, const pde_jac *J, const double *d_b, double *d_x,
                 double rtol, double atol, int64_t maxiter, int64_t check_every,
                 double *d_work, int64_t *h_info, void *stream);

/* Slab-parallel variant of the same solver, one call per step so that the caller can
 * place the NCCL traffic in between (pyellipticfd_b200/multigpu.py):
 *   step 0 setup + init pass | 1 finalize | 2 p-update | 3 v = J phat (exchange the
 *   halo of phat = d_work + 4n first) | 4 finalize | 5 s-update | 6 finalize |
 *   7 t = J shat (halo of shat = d_work + 6n first) | 8 finalize | 9 x,r-update | 10 finalize.
 * Every "pass" step leaves this rank's partial dot products in d_sums (3 doubles); the
 * caller all-reduces them (SUM) before the following "finalize" step, which updates the
 * device-resident scalars d_scal (pde_bicgstab_scal_bytes() bytes).  d_partials:
 * 3 * (rows + tail blocks + 8*SMs) doubles of scratch.  pde_bicgstab_poll reads
 * [done, info, iterations] (synchronises). */
int64_t pde_bicgstab_scal_bytes(void);
int pde_bicgstab_dist_step(pde_grid *g, const pde_jac *J, int step, const double *d_b,
                           double *d_x, double *d_work, void *d_scal, double *d_partials,
                           double *d_sums, double rtol, double atol, int64_t maxiter,
                           void *stream);
int pde_bicgstab_poll(const void *d_scal, int64_t *h_out, void *stream);

/* small fused vector helpers used by the Newton loop (solvers.py:189-199) */
int pde_vec_newton_update(pde_grid *g, const double *d_U, const double *d_d,
                          const double *d_F, const double *d_Jd, double *d_Unew,
                          double *h_out /* [max|U-Unew|, F.(J d), ||d||^2] */, void *stream);

/* ------------------------------------------------------------------------
 * Point-cloud operators (ddi.py:304-414 cached path, ddf.py:137-252): the Nds
 * cached CSR matrices (5 nnz/row) stored as one direction-major ELL table:
 *   idx (nds, n, 4) int32, w (nds, n, 4) double, centre weight = -sum(w)
 * (n = n_interior; 48 contiguous bytes per (direction, point), coalesced over
 * points).  The tables stay owned by the caller (device memory, 32-byte aligned).
 * ---------------------------------------------------------------------- */
int pde_ell_create(pde_ell **out, int device, int64_t n_interior, int64_t n_points,
                   int nds, const int32_t *d_idx, const double *d_w);
int pde_ell_destroy(pde_ell *e);
/* Point-range shard (multi-GPU): the table holds the rows of the interior points
 * [row_offset, row_offset + n_interior) only; pde_ell_d2eigs / pde_ell_apply_policy then read
 * the centre value u[i + row_offset] and write outputs at the local row i.  The problem /
 * Krylov entry points below need the full table (row_offset 0). */
int pde_ell_set_row_offset(pde_ell *e, int64_t row_offset);
/* d2u[i,k] = sum_j w[k,i,j]*(u[idx[k,i,j]] - u[i]); min/max over k (ddi.py:358-367).
 * Ties: lowest k for the minimum, highest k for the maximum (np.argsort's tie
 * order is unspecified in the reference, ddi.py:363). Outputs may be NULL. */
int pde_ell_d2eigs(pde_ell *e, const double *d_u, double *d_lmin, double *d_lmax,
                   int32_t *d_kmin, int32_t *d_kmax, void *stream);
/* y[i] = row (k[i], i) of the table applied to x (selected-direction rows,
 * replaces the per-row CSR slicing loop ddi.py:369-404). */
int pde_ell_apply_policy(pde_ell *e, const int32_t *d_k, const double *d_x,
                         double *d_y, void *stream);
/* Build the table on the device (ddi.py:253-299): for every row (centre point) and
 * direction V[k], the first simplex (row order) whose cone contains +v and the
 * first containing -v (2x2 solves with partial pivoting), barycentric weights,
 * 2/(hf+hb) scaling.  simplices: (S,3) int64 rows (i, s1, s2) grouped by ascending centre
 * with simp_ptr (n_rows+1).  eps_over_h = 1e-15/spatial_resolution
 * (ddi.py:10,267).  v_per_point != 0: nds must be 1 and d_V is (n_rows,2), one direction
 * per row (ddi.d2 / ddf.d2 / ddi.d1).
 * scheme 0: ddi second derivative; 1: Froese's weights (ddf.py:46-116);
 *        2: ddi FIRST derivative (ddi.py:61-80: forward simplex only, weights = cone
 *           coordinates, row padded with two zero-weight entries);
 *        3: the same on the boundary domain (ddi.py:66-68: only simplices whose vertices
 *           are interior points, id < n_interior_pts).
 * row_offset: centre point of row i is point i + row_offset (0 for the interior domain,
 * num_interior for the boundary domain).
 * *d_fail counts (row, direction) units with no enclosing cone. */
int pde_ell_build(int device, int64_t n_rows, const double *d_points /* (npts,2) */,
                  const int64_t *d_simplices, const int64_t *d_simp_ptr,
                  int nds, const double *d_V, int v_per_point, double eps_over_h,
                  int scheme, int64_t row_offset, int64_t n_interior_pts,
                  int32_t *d_idx, double *d_w, int32_t *d_fail, void *stream);

/* First derivatives on a neighbour list (ddg.d1 ddg.py:8-81, ddg.d1grad ddg.py:110-175).
 * Row r has centre point r + centre_offset and the neighbours d_J[d_ptr[r] .. d_ptr[r+1]).
 * d_v != NULL ((n_rows,2) unit directions): the neighbour with the largest cosine against
 * the direction is selected (ddg.py:59-66).  d_v == NULL: the neighbour with the largest
 * difference quotient (u[J]-u[I])/h (ddg.py:146-153; d_u required).  First maximal
 * neighbour in row order wins.  Outputs per row: d_d1 (value, may be NULL when d_u is NULL),
 * d_sel (selected neighbour id), d_w (1/h), d_vout ((n_rows,2) = w*X, ddg.py:160; may be NULL). */
int pde_nb_d1(int device, int64_t n_rows, int64_t centre_offset, const int64_t *d_ptr,
              const int32_t *d_J, const double *d_points, const double *d_u, const double *d_v,
              double *d_d1, int32_t *d_sel, double *d_w, double *d_vout, void *stream);
/* y[r] = sum_j w[r,j] * (x[idx[r,j]] - x[r + centre_offset]) for row-major 4-wide rows
 * ((n_rows,4) int32 / double, 16-byte aligned): the finite-difference matrices of
 * ddg.py:73-79,164-171 and ddi.py:84-90 applied matrix-free. */
int pde_rows_apply(int device, int64_t n_rows, int64_t centre_offset, const int32_t *d_idx,
                   const double *d_w, const double *d_x, double *d_y, void *stream);

/* Problems and the Newton linear solve on point clouds (fdmethod 'interpolate' /
 * 'froese').  All vectors are the reference's flat vectors (n_points doubles,
 * interior points first).
 * pde_ell_residual mode 0: convex_envelope.py:44-46 (F = W-g; interior rows with
 *   -lambda_min > W-g become -lambda_min; d_kmin = selected direction or -1 for an
 *   identity row).  mode 1: pucci.py:48,127-141 (interior A0*lambda_max + A1*lambda_min
 *   - f, boundary W - dirichlet; d_g holds f then the Dirichlet data).
 * d_red[0] accumulates max|F|. */
int pde_ell_residual(pde_ell *e, const double *d_W, const double *d_g, int mode,
                     double A0, double A1, double *d_F, int32_t *d_kmin, int32_t *d_kmax,
                     double *d_red, void *stream);
/* Jacobian of a direction policy: interior row i = cmin*row(kmin[i]) (+ cmax*row(kmax[i])),
 * identity when kmin[i] < 0; boundary rows identity (convex_envelope.py:48-54,
 * pucci.py:50-54,136-139 with the rows of ddi.py:369-404). */
typedef struct pde_ell_jac {
    const int32_t *d_kmin;
    const int32_t *d_kmax;   /* NULL when cmax == 0 */
    double cmin, cmax;
} pde_ell_jac;
int pde_ell_jac_matvec(pde_ell *e, const pde_ell_jac *J, const double *d_x, double *d_y,
                       void *stream);
int pde_ell_jac_diagonal(pde_ell *e, const pde_ell_jac *J, double *d_diag, void *stream);
/* same contract as pde_bicgstab (solvers.py:179-181); d_work: 8 * n_points doubles */
int pde_ell_bicgstab(pde_ell *e, const pde_ell_jac *J, const double *d_b, double *d_x,
                     double rtol, double atol, int64_t maxiter, int64_t check_every,
                     double *d_work, int64_t *h_info, void *stream);
int pde_ell_newton_update(pde_ell *e, const double *d_U, const double *d_d, const double *d_F,
                          const double *d_Jd, double *d_Unew, double *h_out, void *stream);

/* ------------------------------------------------------------------------
 * Slab halo exchange over NVLink peer memory (no counterpart in the single-process
 * reference).  Every rank owns a staging buffer of pde_halo_stage_doubles(halo, pitch) doubles
 * in peer-mapped (symmetric) memory, zero-initialised; `peer_*_stage` are the neighbours' staging
 * buffers as mapped in THIS process (NULL at the ends of the slab chain).  `d_view` = the lattice
 * part of a state vector ((nx_local + 2 halo) rows x pitch, 16-byte aligned).  `step` = exchange
 * number, 1, 2, 3, ... the same on all ranks.  Both calls only enqueue one small kernel on
 * `stream`: push = remote stores of the first / last `halo` owned rows + release of the
 * neighbours' flag words; wait_copy = acquire of this rank's flag words, then staged rows ->
 * halo rows. */
int64_t pde_halo_stage_doubles(int64_t halo, int64_t pitch);
int pde_halo_push(const double *d_view, int64_t nx_local, int64_t halo, int64_t pitch,
                  double *my_stage, double *peer_lo_stage, double *peer_hi_stage, uint64_t step,
                  void *stream);
int pde_halo_wait_copy(double *d_view, int64_t nx_local, int64_t halo, int64_t pitch,
                       const double *my_stage, int has_lo, int has_hi, uint64_t step,
                       void *stream);

/* ------------------------------------------------------------------------
 * EXTENSION (no counterpart in the reference): the convex envelope of convex_envelope.py:44-46
 * on a structured 2-D pairs grid by directional convex-hull sweeps.  The discrete solution is
 * the largest u <= g with a non-negative second difference for every pair (ddg.py:281) of every
 * interior point; one sweep hulls every lattice line of every stencil direction (runs of points
 * that carry the lattice pair, neighbours as fixed ends) and lowers the band points to the
 * chords of their other pairs.  Same fixed point as the reference's Euler / Newton solvers.
 *   U          state layout (lattice rows x pitch, wall points from `wall_base`); in: g, out: u
 *   h_dirs     (ndir, 2) first offset (a, b) of every deep pair (HOST pointer)
 *   ring       wall-point number of the integer ring points, [i=0 | i=nx+1] (ny+2 each, by j)
 *              then [j=0 | j=ny+1] (nx+2 each, by i)
 *   flags      per lattice offset, bit k set = the point carries the lattice pair of direction k
 *              (only read at points within stencil_radius cells of a wall)
 *   band_ptr   (n_band + 1) CSR over the band points that own pairs which are NOT lattice pairs,
 *   band_center (n_band) their state offsets, rest_j (n_rest, 2) state offsets (J0, J1) of those
 *              pairs, rest_theta (n_rest) = w_0 / (w_0 + w_1): u_I <- min(u_I, theta u_J0 +
 *              (1 - theta) u_J1), a Jacobi step repeated relax_reps times after every
 *              relax_every direction passes (0: once per sweep)
 *   flat_tol   a point less than flat_tol below a chord counts as on it (rounding noise on the
 *              planar pieces of the solution; a few ulps of max|g|; 0: exact comparisons)
 *   work       device scratch of pde_ce_sweep_work_bytes(...) bytes (device path only): relaxation
 *              and convergence words, then the table of fixed-vertex masks of every 32-point
 *              word of every lattice line of every direction, which pde_ce_sweep_setup fills
 *              ONCE per grid (it depends on the grid only) before the first pde_ce_sweep
 *   h_out      [last sweep's largest decrease, sweeps done]; stops when the decrease < tol
 * device >= 0: every pointer but h_dirs / h_out is a DEVICE pointer; up to 64 sweeps are enqueued
 * per host synchronisation.  device = -1: host pointers, the serial monotone chain per line under
 * OpenMP (used by the CPU tests and as the CPU baseline of bench.py). */
int64_t pde_ce_sweep_work_bytes(int64_t nx, int64_t ny, int ndir, const int32_t *h_dirs,
                                int64_t n_band);
int pde_ce_sweep_setup(int device, int64_t nx, int64_t ny, int64_t pitch, int64_t wall_base,
                       int stencil_radius, int ndir, const int32_t *h_dirs, const int64_t *ring,
                       const uint64_t *flags, int64_t n_band, double *work, void *stream);
int pde_ce_sweep(int device, int64_t nx, int64_t ny, int64_t pitch, int64_t wall_base,
                 int stencil_radius, int ndir, const int32_t *h_dirs, const int64_t *ring,
                 const uint64_t *flags, int64_t n_band, const int64_t *band_ptr,
                 const int64_t *band_center, const int64_t *rest_j, const double *rest_theta,
                 double *U, double *work, double tol, double flat_tol, int64_t max_sweeps,
                 int relax_every, int relax_reps, double *h_out, void *stream);

/* ------------------------------------------------------------------------
 * Host-side pre-processing (no GPU needed): pair rows of the band points of a
 * regular grid.  Replaces, for the points within `stencil_radius` cells of a wall,
 * pointclasses.py:421-471 (neighbour search), :37-103 (colinear pruning, tol 1e-3)
 * and :12-35 (antipodal pairing) of FDRegularGrid(interpolation=False).
 * wall: (nb,2) unscaled wall-point coordinates (pointclasses.py:523-538);
 * band_ids: interior point ids (ascending).  Outputs are malloc'd: *out_ptr
 * (n_band+1) and *out_pairs (P x 3 rows (I, J0, J1) with global point ids);
 * release them with pde_free.  Returns 2 with "Point i has no pairs"
 * (pointclasses.py:26-27) when a point ends up without a pair. */
int pde_band_build(int64_t nx, int64_t ny, int stencil_radius, double sx, double sy,
                   double bx, double by, int64_t nb, const double *h_wall,
                   int64_t n_band, const int64_t *h_band_ids, double colinear_tol,
                   int64_t **out_ptr, int64_t **out_pairs, double *out_min_nb_dist,
                   /* optional (may be NULL): the pruned neighbour lists themselves, CSR over the
                    * band points (what ddg.d1 / d1grad walk, ddg.py:44-47) */
                   int64_t **out_nb_ptr, int64_t **out_nb);
/* The same pre-processing for lattices of any dimension (dim = 2, 3 or 4; used for the 3-D grids
 * of stencils.py:46-61 `create_nd`): shape, scaling, bounds0 have `dim` entries, h_wall is
 * (nb, dim), band ids are C-order lattice numbers (last axis fastest, pointclasses.py:513-515).
 * Same outputs and error behaviour as pde_band_build. */
int pde_band_build_nd(int dim, const int64_t *shape, int stencil_radius, const double *scaling,
                      const double *bounds0, int64_t nb, const double *h_wall,
                      int64_t n_band, const int64_t *h_band_ids, double colinear_tol,
                      int64_t **out_ptr, int64_t **out_pairs, double *out_min_nb_dist,
                      int64_t **out_nb_ptr, int64_t **out_nb);
void pde_free(void *p);
/* Number of host threads of the pre-processing routines (pde_band_build, host pde_cloud_build);
 * n <= 0 only queries.  torchrun exports OMP_NUM_THREADS=1, which would build an N-GPU job's
 * grid on one core per rank. */
int pde_set_host_threads(int n);

/* Stencils of a triangulated point cloud: replaces, per centre point, the O(N^2)
 * pre-processing of FDTriMesh.__init__ -- neighbours within `depth` edges of the
 * triangulation graph (pointclasses.py:373-382, sum of adjacency-matrix powers), the radius
 * mask 0 < D <= max_radius and (D >= min_radius or boundary neighbour) (:385-392), colinear
 * pruning with prefer="max" (:37-103) and the 2-D simplices (ConvexHull of the unit stencil
 * vectors = consecutive neighbours in angular order, :105-121).
 * device >= 0: every pointer is a DEVICE pointer and the work runs on that GPU (one thread per
 * point); device = -1: HOST pointers, the same routine under OpenMP (constructor usable
 * without a GPU; also the executable specification the device build is tested against).
 * adj_ptr/adj: CSR of the triangulation graph (both directions of every edge).
 * Outputs, rows of `kmax` entries per point: out_count[i] = number of kept neighbours K_i
 * (negative: 1 BFS, 2 candidate, 3 row capacity exceeded -- retry with larger capacities),
 * out_nb[i, :K_i] the neighbours by ascending id, out_ring[i, :K_i] the same neighbours in
 * angular order (simplices of i: (i, ring[k], ring[(k+1) % K_i])).
 * out_status[1..3] count the points that overflowed. */
int pde_cloud_build(int device, int64_t n_points, int64_t n_interior, const double *points,
                    const int64_t *adj_ptr, const int32_t *adj, int depth, double min_radius,
                    double max_radius, double colinear_tol, int kmax, int bfs_capacity,
                    int32_t *out_count, int32_t *out_nb, int32_t *out_ring, int32_t *out_status,
                    void *stream);

/* fp64 FMA-chain microbenchmark: returns achieved DFMA/s of the device in
 * *h_dfma_per_s (synchronous).  Used by bench.py to evidence which roof
 * (HBM or the fp64 pipe) binds the stencil kernel. */
int pde_fp64_peak(int device, double *h_dfma_per_s);
/* Arithmetic-only probes with K1's register blocking (8 independent rows per thread, 2 CTAs of
 * 256 threads per SM): kind 1 DADD, 2 DMUL, 3 compare-select (DSETP + 2 FSEL) -- units/s = thread
 * instructions per second; kind 4 = the exact pair formula of ddg.py:281 plus the min update on
 * register operands (units/s = pairs per second): the speed of light of the reference's
 * operation order, which K1 is compared against in bench.py; kind 5 = independent DSETP (the
 * predicate feeds an integer select + add: no fp64 select, no dependency between compares). */
int pde_fp64_probe(int device, int kind, double *h_units_per_s);

#ifdef __cplusplus
}
#endif
#endif /* PDE_B200_H */
