// Context management, layout conversion and error plumbing for the C ABI
// declared in include/pde_b200.h.
#include <atomic>
#include <cstring>
#include <vector>

#include "common.cuh"
#include "stencil_ct.cuh"
#include "stencil_ct3.cuh"

namespace pde {

static thread_local std::string g_last_error;
static std::atomic<int64_t> g_launches{0};

void set_error(const std::string &msg) { g_last_error = msg; }
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
int64_t launch_count() { return g_launches.load(); }

void keep_async_pool(int device) {
    static std::atomic<unsigned long long> done{0ull};
    const unsigned long long bit = 1ull << (device & 63);
    if (done.load() & bit) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        uint64_t thr = UINT64_MAX;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    cudaGetLastError();
    done.fetch_or(bit);
}

int fail(const char *what, const char *file, int line) {
    char buf[512];
    snprintf(buf, sizeof buf, "%s (%s:%d)", what, file, line);
    g_last_error = buf;
    return 1;
}
int fail_cuda(cudaError_t e, const char *what, const char *file, int line) {
    char buf[768];
    snprintf(buf, sizeof buf, "CUDA error %d (%s) in %s (%s:%d)", (int)e, cudaGetErrorString(e),
             what, file, line);
    g_last_error = buf;
    return (int)e ? (int)e : 1;
}

// ---- layout kernels ------------------------------------------------------------
__global__ void k_pack(const double *__restrict__ flat, double *__restrict__ state, int64_t nloc,
                       int64_t ny, int64_t pitch, int64_t halo_rows, int64_t flat_row0,
                       int64_t n_interior_global, int64_t nb, int64_t rows) {
    // one thread per state entry; padding and halo rows are zeroed
    int64_t total = rows * pitch + nb;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        double v = 0.0;
        if (i >= rows * pitch) {
            v = flat[n_interior_global + (i - rows * pitch)];
        } else {
            int64_t r = i / pitch, cidx = i - r * pitch;
            int64_t lr = r - halo_rows;
            if (cidx < ny && lr >= 0 && lr < nloc) v = flat[(flat_row0 + lr) * ny + cidx];
        }
        state[i] = v;
    }
}

__global__ void k_unpack(const double *__restrict__ state, double *__restrict__ flat, int64_t nloc,
                         int64_t ny, int64_t pitch, int64_t halo_rows, int64_t flat_row0,
                         int64_t n_interior_global, int64_t nb, int64_t rows, int with_wall) {
    int64_t total = nloc * ny + (with_wall ? nb : 0);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        if (i < nloc * ny) {
            int64_t r = i / ny, cidx = i - r * ny;
            flat[(flat_row0 + r) * ny + cidx] = state[(r + halo_rows) * pitch + cidx];
        } else {
            int64_t j = i - nloc * ny;
            flat[n_interior_global + j] = state[rows * pitch + j];
        }
    }
}

__global__ void k_unpad_u16(const uint16_t *__restrict__ padded, int32_t *__restrict__ flat,
                            int64_t nloc, int64_t ny, int64_t pitch, int64_t halo_rows,
                            int64_t flat_row0) {
    int64_t total = nloc * ny;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = i / ny, cidx = i - r * ny;
        flat[(flat_row0 + r) * ny + cidx] = padded[(r + halo_rows) * pitch + cidx];
    }
}

template <class T>
static int upload(T **dst, const T *src, size_t n) {
    *dst = nullptr;
    if (n == 0) n = 1;      // keep pointers valid
    PDE_CHECK(cudaMalloc((void **)dst, n * sizeof(T)));
    if (src) PDE_CHECK(cudaMemcpy(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice));
    return 0;
}

}  // namespace pde

using namespace pde;

extern "C" {

int pde_abi_version(void) { return PDE_B200_ABI_VERSION; }
const char *pde_last_error(void) { return pde::g_last_error.c_str(); }
int64_t pde_kernel_launches(void) { return pde::g_launches.load(); }

int pde_grid_create(pde_grid **out, int device, int64_t nx_global, int64_t ny, int64_t nb,
                    int stencil_radius, int64_t x_offset, int64_t nx_local, int64_t halo_rows,
                    int D, const int32_t *h_stencil_pairs, const double *h_stencil_w,
                    const double *h_stencil_X, int64_t n_band, const int64_t *h_band_center,
                    const int64_t *h_band_ptr, const int64_t *h_band_j, const double *h_band_w,
                    const double *h_band_X) {
    PDE_REQUIRE(out != nullptr, "out is NULL");
    PDE_REQUIRE(D >= 0 && D <= PDE_MAX_DIRS, "D out of range");
    PDE_REQUIRE(nx_local >= 0 && ny > 0 && x_offset >= 0 && x_offset + nx_local <= nx_global,
                "bad slab extents");
    PDE_REQUIRE(halo_rows >= 0, "bad halo_rows");
    int ndev = 0;
    PDE_CHECK(cudaGetDeviceCount(&ndev));
    PDE_REQUIRE(device >= 0 && device < ndev, "no such CUDA device");
    DeviceGuard guard(device);
    pde::keep_async_pool(device);

    pde_grid *g = new pde_grid();
    memset(g, 0, sizeof(*g));
    g->device = device;
    g->nx_global = nx_global;
    g->ny = ny;
    g->x_offset = x_offset;
    g->nx_local = nx_local;
    g->halo_rows = halo_rows;
    g->rows = nx_local + 2 * halo_rows;
    g->pitch = (ny + 15) / 16 * 16;
    g->nb = nb;
    g->radius = stencil_radius;
    g->tab.D = D;
    int H = 0;
    for (int k = 0; k < D; ++k) {
        g->tab.a0[k] = h_stencil_pairs[4 * k + 0];
        g->tab.b0[k] = h_stencil_pairs[4 * k + 1];
        g->tab.a1[k] = h_stencil_pairs[4 * k + 2];
        g->tab.b1[k] = h_stencil_pairs[4 * k + 3];
        for (int j = 0; j < 4; ++j) H = max(H, abs(h_stencil_pairs[4 * k + j]));
        g->tab.w_0[k] = h_stencil_w[3 * k + 0];
        g->tab.w_1[k] = h_stencil_w[3 * k + 1];
        g->tab.w0[k] = h_stencil_w[3 * k + 2];
        g->stencil_X[k][0] = h_stencil_X ? h_stencil_X[2 * k + 0] : 0.0;
        g->stencil_X[k][1] = h_stencil_X ? h_stencil_X[2 * k + 1] : 0.0;
    }
    g->tab.H = H;
    g->ct_id = -1;
    g->ct3_id = -1;
    for (int t = 0; t < pde::CT_NUM_TABLES && g->ct_id < 0; ++t) {
        if (pde::CT_TABLE_D[t] != D) continue;
        bool same = true;
        for (int k = 0; k < D && same; ++k)
            for (int j = 0; j < 4; ++j) same = same && pde::CT_TABLES[t][k][j] == h_stencil_pairs[4 * k + j];
        if (same) g->ct_id = t;
    }
    if (H > 6 || H > stencil_radius) {
        delete g;
        return pde::fail("stencil halo larger than supported (6) or than the radius", __FILE__,
                         __LINE__);
    }
    g->n_band = n_band;
    g->win_lo = 0;
    g->win_hi = nx_local;
    g->win2_lo = g->win2_hi = 0;
    g->dyn_sched = 0;
    g->h_band_row = new int64_t[n_band > 0 ? n_band : 1];
    for (int64_t i = 0; i < n_band; ++i) g->h_band_row[i] = h_band_center[i] / g->pitch - halo_rows;
    g->n_band_pairs = n_band ? h_band_ptr[n_band] : 0;
    int rc = 0;
    rc |= upload(&g->d_band_center, h_band_center, (size_t)n_band);
    rc |= upload(&g->d_band_ptr, h_band_ptr, (size_t)n_band + 1);
    rc |= upload(&g->d_band_j, h_band_j, (size_t)g->n_band_pairs * 2);
    rc |= upload(&g->d_band_w, h_band_w, (size_t)g->n_band_pairs * 3);
    rc |= upload(&g->d_band_X, h_band_X, (size_t)g->n_band_pairs * 2);
    rc |= upload<double>(&g->d_red, nullptr, 16);
    rc |= upload<unsigned long long>(&g->d_ctrl, nullptr, 16);
    if (rc) {
        pde_grid_destroy(g);
        return rc;
    }
    PDE_CHECK(cudaMemset(g->d_red, 0, 16 * sizeof(double)));
    PDE_CHECK(cudaMemset(g->d_ctrl, 0, 16 * sizeof(unsigned long long)));
    PDE_CHECK(cudaStreamCreateWithFlags(&g->side, cudaStreamNonBlocking));
    PDE_CHECK(cudaEventCreateWithFlags(&g->ev_fork, cudaEventDisableTiming));
    PDE_CHECK(cudaEventCreateWithFlags(&g->ev_join, cudaEventDisableTiming));

    cudaDeviceProp prop;
    PDE_CHECK(cudaGetDeviceProperties(&prop, device));
    g->num_sms = prop.multiProcessorCount;
    if (prop.major != 10) {
        pde_grid_destroy(g);
        return pde::fail("pde_b200 requires an sm_100-class (Blackwell) device", __FILE__, __LINE__);
    }
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    PDE_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (qres != cudaDriverEntryPointSuccess || !fn) {
        pde_grid_destroy(g);
        return pde::fail("cuTensorMapEncodeTiled not available", __FILE__, __LINE__);
    }
    g->encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
    *out = g;
    return 0;
}

// 3-D lattice [nx][ny][nz] (stencils.py:46-61): stored as nx*ny state-layout rows of nz
// columns, so that every per-point array, the band rows and the Krylov passes are those of the
// 2-D case; deep points are served by the 3-D TMA kernel (stencil_deep3.cuh).
int pde_grid_create_nd(pde_grid **out, int device, int dim, const int64_t *h_shape, int64_t nb,
                       int stencil_radius, int D, const int32_t *h_stencil_pairs /* D x 6 */,
                       const double *h_stencil_w, int64_t n_band, const int64_t *h_band_center,
                       const int64_t *h_band_ptr, const int64_t *h_band_j, const double *h_band_w) {
    PDE_REQUIRE(out && h_shape, "null argument");
    PDE_REQUIRE(dim == 3, "structured n-D lattices: dim = 3 only");
    PDE_REQUIRE(D >= 1 && D <= PDE_MAX_DIRS, "D out of range");
    const int64_t nx = h_shape[0], ny = h_shape[1], nz = h_shape[2];
    PDE_REQUIRE(nx > 0 && ny > 0 && nz > 0 && nx * ny < (1ll << 31), "bad lattice shape");
    int H = 0;
    for (int k = 0; k < 6 * D; ++k) H = max(H, abs(h_stencil_pairs[k]));
    PDE_REQUIRE(H >= 1 && H <= 2 && H <= stencil_radius, "3-D stencil halo must be 1 or 2");
    // band rows carry 2-D control vectors in the 2-D ABI; none in 3-D (host side, ddg.py)
    std::vector<double> zeroX((size_t)(n_band ? h_band_ptr[n_band] : 0) * 2 + 2, 0.0);
    int rc = pde_grid_create(out, device, nx * ny, nz, nb, stencil_radius, 0, nx * ny, 0, 0, nullptr,
                             nullptr, nullptr, n_band, h_band_center, h_band_ptr, h_band_j, h_band_w,
                             zeroX.data());
    if (rc) return rc;
    pde_grid *g = *out;
    g->nmid = ny;
    g->nx3 = nx;
    g->tab.D = D;
    g->tab.H = H;
    for (int k = 0; k < D; ++k) {
        const int32_t *q = h_stencil_pairs + 6 * k;
        for (int j = 0; j < 6; ++j) g->nd_abc[k][j] = q[j];
        g->tab.a0[k] = q[0] * (int)ny + q[1];      // row offset in the flattened (x, y) index
        g->tab.b0[k] = q[2];
        g->tab.a1[k] = q[3] * (int)ny + q[4];
        g->tab.b1[k] = q[5];
        g->tab.w_0[k] = h_stencil_w[3 * k + 0];
        g->tab.w_1[k] = h_stencil_w[3 * k + 1];
        g->tab.w0[k] = h_stencil_w[3 * k + 2];
    }
    for (int t = 0; t < pde::CT3_NUM_TABLES && g->ct3_id < 0; ++t) {
        if (pde::CT3_TABLE_D[t] != D) continue;
        bool same = true;
        for (int k = 0; k < D && same; ++k)
            for (int j = 0; j < 6; ++j) same = same && pde::CT3_TABLES[t][k][j] == h_stencil_pairs[6 * k + j];
        if (same) g->ct3_id = t;
    }
    return 0;
}

int pde_grid_set_band_classes(pde_grid *g, int n_cls, const int32_t *h_cls_ptr,
                              const int32_t *h_cls_ref, const double *h_cls_w,
                              const uint16_t *h_cls_id, const int32_t *h_wl_ptr,
                              const int32_t *h_wl) {
    PDE_REQUIRE(g && n_cls >= 1 && n_cls < 65536 && h_cls_ptr && h_cls_ref && h_cls_w && h_cls_id &&
                    h_wl_ptr && h_wl,
                "bad band-class tables");
    PDE_REQUIRE(g->n_cls == 0, "band classes already set");
    DeviceGuard guard(g->device);
    const size_t rows = (size_t)h_cls_ptr[n_cls], nw = (size_t)h_wl_ptr[g->n_band];
    int rc = 0;
    rc |= upload(&g->d_cls_ptr, h_cls_ptr, (size_t)n_cls + 1);
    rc |= upload(&g->d_cls_ref, h_cls_ref, rows * 2);
    rc |= upload(&g->d_cls_w, h_cls_w, rows * 3);
    rc |= upload(&g->d_cls_id, h_cls_id, (size_t)g->n_band);
    rc |= upload(&g->d_wl_ptr, h_wl_ptr, (size_t)g->n_band + 1);
    rc |= upload(&g->d_wl, h_wl, nw);
    if (rc) return rc;
    g->n_cls = n_cls;
    return 0;
}

int pde_grid_destroy(pde_grid *g) {
    if (!g) return 0;
    DeviceGuard guard(g->device);
    cudaFree(g->d_cls_ptr);
    cudaFree(g->d_cls_ref);
    cudaFree(g->d_cls_w);
    cudaFree(g->d_cls_id);
    cudaFree(g->d_wl_ptr);
    cudaFree(g->d_wl);
    cudaFree(g->d_band_center);
    cudaFree(g->d_band_ptr);
    cudaFree(g->d_band_j);
    cudaFree(g->d_band_w);
    cudaFree(g->d_band_X);
    cudaFree(g->d_red);
    cudaFree(g->d_ctrl);
    delete[] g->h_band_row;
    for (int i = 0; i < g->n_ev_pool; ++i) cudaEventDestroy(g->ev_pool[i]);
    delete[] g->ev_pool;
    if (g->side) cudaStreamDestroy(g->side);
    if (g->ev_fork) cudaEventDestroy(g->ev_fork);
    if (g->ev_join) cudaEventDestroy(g->ev_join);
    delete g;
    return 0;
}

int pde_grid_set_row_window(pde_grid *g, int64_t row_lo, int64_t row_hi) {
    PDE_REQUIRE(g, "null argument");
    if (row_lo < 0) row_lo = 0;
    if (row_hi > g->nx_local) row_hi = g->nx_local;
    if (row_hi < row_lo) row_hi = row_lo;
    g->win_lo = row_lo;
    g->win_hi = row_hi;
    g->win2_lo = g->win2_hi = 0;
    return 0;
}

int pde_grid_set_dynamic_schedule(pde_grid *g, int on) {
    PDE_REQUIRE(g, "null argument");
    g->dyn_sched = on ? 1 : 0;
    return 0;
}

int pde_grid_set_row_windows2(pde_grid *g, int64_t lo, int64_t hi, int64_t lo2, int64_t hi2) {
    PDE_REQUIRE(g, "null argument");
    PDE_REQUIRE(lo >= 0 && lo <= hi && hi <= lo2 && lo2 <= hi2 && hi2 <= g->nx_local,
                "bad row ranges (need lo <= hi <= lo2 <= hi2)");
    g->win_lo = lo;
    g->win_hi = hi;
    g->win2_lo = lo2;
    g->win2_hi = hi2;
    return 0;
}

int64_t pde_grid_pitch(const pde_grid *g) { return g->pitch; }
int64_t pde_grid_rows(const pde_grid *g) { return g->rows; }
int64_t pde_grid_state_len(const pde_grid *g) { return g->rows * g->pitch + g->nb; }
int64_t pde_grid_policy_len(const pde_grid *g) { return g->rows * g->pitch; }

static inline int grid_for(int64_t n, int num_sms) {
    int64_t b = (n + 255) / 256;
    int64_t cap = (int64_t)num_sms * 16;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

int pde_grid_pack(pde_grid *g, const double *d_u_flat, double *d_state, void *stream) {
    DeviceGuard guard(g->device);
    int64_t total = pde_grid_state_len(g);
    k_pack<<<grid_for(total, g->num_sms), 256, 0, (cudaStream_t)stream>>>(
        d_u_flat, d_state, g->nx_local, g->ny, g->pitch, g->halo_rows, 0,
        g->nx_local * g->ny, g->nb, g->rows);
    PDE_LAUNCH_CHECK();
    return 0;
}

int pde_grid_unpack(pde_grid *g, const double *d_state, double *d_u_flat, void *stream) {
    DeviceGuard guard(g->device);
    int64_t total = g->nx_local * g->ny + g->nb;
    k_unpack<<<grid_for(total, g->num_sms), 256, 0, (cudaStream_t)stream>>>(
        d_state, d_u_flat, g->nx_local, g->ny, g->pitch, g->halo_rows, 0,
        g->nx_local * g->ny, g->nb, g->rows, 1);
    PDE_LAUNCH_CHECK();
    return 0;
}

int pde_grid_unpad(pde_grid *g, const double *d_padded, double *d_flat, void *stream) {
    DeviceGuard guard(g->device);
    int64_t total = g->nx_local * g->ny;
    k_unpack<<<grid_for(total, g->num_sms), 256, 0, (cudaStream_t)stream>>>(
        d_padded, d_flat, g->nx_local, g->ny, g->pitch, g->halo_rows, 0,
        g->nx_local * g->ny, g->nb, g->rows, 0);
    PDE_LAUNCH_CHECK();
    return 0;
}

int pde_grid_unpad_u16(pde_grid *g, const uint16_t *d_padded, int32_t *d_flat, void *stream) {
    DeviceGuard guard(g->device);
    int64_t total = g->nx_local * g->ny;
    k_unpad_u16<<<grid_for(total, g->num_sms), 256, 0, (cudaStream_t)stream>>>(
        d_padded, d_flat, g->nx_local, g->ny, g->pitch, g->halo_rows, 0);
    PDE_LAUNCH_CHECK();
    return 0;
}

// Host <-> state-layout copies of a range of owned rows, without a pack/unpack kernel:
// one 2-D DMA copy converts between the reference's dense rows (ny doubles) and the
// padded rows (pitch doubles).  Used to pipeline H2D / stencil / D2H over row chunks.
int pde_grid_h2d_rows(pde_grid *g, const double *h_flat, double *d_state, int64_t row_lo,
                      int64_t row_hi, int with_wall, void *stream) {
    PDE_REQUIRE(g && h_flat && d_state, "null argument");
    PDE_REQUIRE(row_lo >= 0 && row_hi <= g->nx_local && row_lo <= row_hi, "bad row range");
    DeviceGuard guard(g->device);
    cudaStream_t st = (cudaStream_t)stream;
    if (row_hi > row_lo)
        PDE_CHECK(cudaMemcpy2DAsync(d_state + (g->halo_rows + row_lo) * g->pitch, g->pitch * 8,
                                    h_flat + row_lo * g->ny, g->ny * 8, g->ny * 8,
                                    (size_t)(row_hi - row_lo), cudaMemcpyDefault, st));
    if (with_wall && g->nb > 0)
        PDE_CHECK(cudaMemcpyAsync(d_state + g->rows * g->pitch, h_flat + g->nx_local * g->ny,
                                  g->nb * 8, cudaMemcpyDefault, st));
    return 0;
}

int pde_grid_d2h_rows(pde_grid *g, const double *d_state, double *h_flat, int64_t row_lo,
                      int64_t row_hi, void *stream) {
    PDE_REQUIRE(g && h_flat && d_state, "null argument");
    PDE_REQUIRE(row_lo >= 0 && row_hi <= g->nx_local && row_lo <= row_hi, "bad row range");
    DeviceGuard guard(g->device);
    if (row_hi > row_lo)
        PDE_CHECK(cudaMemcpy2DAsync(h_flat + row_lo * g->ny, g->ny * 8,
                                    d_state + (g->halo_rows + row_lo) * g->pitch, g->pitch * 8,
                                    g->ny * 8, (size_t)(row_hi - row_lo), cudaMemcpyDefault,
                                    (cudaStream_t)stream));
    return 0;
}

}  // extern "C"
