// Control directions of ddg.d2eigs and the fp64-pipe microbenchmark.
#include "jacobian.cuh"

using namespace pde;

namespace pde {

// ddg.py:297-300:  v_min = X[min pair, 0] * w[min pair, 0]
//                  v_max = X[max pair, 0] * w[MIN pair, 0]     (quirk Q1, replicated)
__global__ void k_control(const GridDev G, const __grid_constant__ DeepTable tab,
                          const double *__restrict__ sX /* D x 2 */,
                          const double *__restrict__ bandX, const uint16_t *__restrict__ pmin,
                          const uint16_t *__restrict__ pmax, double *__restrict__ vmin,
                          double *__restrict__ vmax) {
    if ((int64_t)blockIdx.x < G.rows) {
        int64_t br = blockIdx.x;
        int64_t lr = br - G.halo_rows;
        if (lr < 0 || lr >= G.nx_local) return;
        int64_t gx = lr + G.x_offset;
        if (!deep_row(G, gx)) return;
        for (int64_t c = threadIdx.x; c < G.ny; c += blockDim.x) {
            if (c < G.radius || c >= G.ny - G.radius) continue;
            int64_t o = br * G.pitch + c;
            int64_t f = lr * G.ny + c;
            unsigned a = pmin[o], b = pmax[o];
            double what = tab.w_0[a];
            vmin[2 * f] = sX[2 * a] * what;
            vmin[2 * f + 1] = sX[2 * a + 1] * what;
            vmax[2 * f] = sX[2 * b] * what;
            vmax[2 * f + 1] = sX[2 * b + 1] * what;
        }
    } else {
        int64_t i = ((int64_t)blockIdx.x - G.rows) * blockDim.x + threadIdx.x;
        if (i >= G.n_band) return;
        int64_t o = G.band_center[i];
        int64_t br = o / G.pitch, c = o - br * G.pitch;
        int64_t f = (br - G.halo_rows) * G.ny + c;
        int64_t ka = G.band_ptr[i] + pmin[o], kb = G.band_ptr[i] + pmax[o];
        double what = G.band_w[3 * ka];
        vmin[2 * f] = bandX[2 * ka] * what;
        vmin[2 * f + 1] = bandX[2 * ka + 1] * what;
        vmax[2 * f] = bandX[2 * kb] * what;
        vmax[2 * f + 1] = bandX[2 * kb + 1] * what;
    }
}

__global__ void k_dfma_chain(double *out, int iters) {
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3;
    double a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 0.999999, c = 1e-7;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

// Arithmetic-only probes of the fp64 pipe, 8 independent rows per thread like K1:
//   kind 1: DADD chain        2: DMUL chain        3: compare-select (DSETP + 2 FSEL)
//   kind 4: the exact pair formula of ddg.py:281 + the min update (7 fp64-pipe
//           instructions per pair) on register operands perturbed by one integer op each,
//           i.e. K1's inner loop without shared-memory traffic: the arithmetic speed of
//           light of the reference's operation order.
//   kind 5: INDEPENDENT fp64 compares: DSETP on two register operands (one perturbed by an
//           integer op), the predicate consumed by an integer select + add -- no fp64 select,
//           no dependency between compares: the issue rate of DSETP alone.
template <int KIND>
__global__ void k_fp64_probe(double *out, int iters) {
    double a[8], m[8];
    int cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int p = 0; p < 8; ++p) {
        a[p] = 1.0 + threadIdx.x * 1e-6 + p * 1e-3;
        m[p] = 1e300;
    }
    const double w = 0.70710678118654752, w0 = 1.4142135623730951, c = 1.0000001;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int p = 0; p < 8; ++p) {
            if (KIND == 1) a[p] = __dadd_rn(a[p], c);
            if (KIND == 2) a[p] = __dmul_rn(a[p], c);
            if (KIND == 3) {
                double d = __hiloint2double(__double2hiint(a[p]), __double2loint(a[p]) ^ i);
                m[p] = d <= m[p] ? d : m[p];
            }
            if (KIND == 5) {
                double d = __hiloint2double(__double2hiint(a[p]), __double2loint(a[p]) ^ i);
                int q;
                asm volatile("{ .reg .pred t; setp.lt.f64 t, %1, %2; selp.s32 %0, 1, 0, t; }"
                             : "=r"(q) : "d"(d), "d"(a[(p + 1) & 7]));
                cnt[p] += q;
            }
            if (KIND == 4) {
                int hi = __double2hiint(a[p]), lo = __double2loint(a[p]);
                double u0 = __hiloint2double(hi, lo ^ i), u1 = __hiloint2double(hi, lo + i);
                double d = pair_value_exact(a[p], u0, u1, w, w, w0);
                m[p] = d <= m[p] ? d : m[p];
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int p = 0; p < 8; ++p) s += a[p] + m[p] + cnt[p];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace pde

extern "C" {

int pde_ddg_control(pde_grid *g, const uint16_t *d_pmin, const uint16_t *d_pmax, double *d_vmin,
                    double *d_vmax, void *stream) {
    PDE_REQUIRE(g && d_pmin && d_pmax && d_vmin && d_vmax, "null argument");
    PDE_REQUIRE(g->nmid == 0, "pde_ddg_control serves 2-D lattices");
    DeviceGuard guard(g->device);
    cudaStream_t st = (cudaStream_t)stream;
    double *sX = nullptr;
    PDE_CHECK(cudaMallocAsync((void **)&sX, sizeof(double) * 2 * PDE_MAX_DIRS, st));
    PDE_CHECK(cudaMemcpyAsync(sX, g->stencil_X, sizeof(double) * 2 * PDE_MAX_DIRS,
                              cudaMemcpyHostToDevice, st));
    GridDev G = grid_dev(g);
    int64_t nblk = g->rows + (g->n_band + 255) / 256;
    k_control<<<(unsigned)nblk, 256, 0, st>>>(G, g->tab, sX, g->d_band_X, d_pmin, d_pmax, d_vmin,
                                              d_vmax);
    pde::count_launch();
    cudaError_t e = cudaGetLastError();
    cudaFreeAsync(sX, st);
    if (e != cudaSuccess) return pde::fail_cuda(e, "k_control", __FILE__, __LINE__);
    return 0;
}

int pde_fp64_peak(int device, double *h_dfma_per_s) {
    PDE_REQUIRE(h_dfma_per_s, "null argument");
    DeviceGuard guard(device);
    cudaDeviceProp prop;
    PDE_CHECK(cudaGetDeviceProperties(&prop, device));
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 15;
    double *out = nullptr;
    PDE_CHECK(cudaMalloc((void **)&out, sizeof(double) * blocks * threads));
    cudaEvent_t e0, e1;
    PDE_CHECK(cudaEventCreate(&e0));
    PDE_CHECK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        PDE_CHECK(cudaEventRecord(e0));
        k_dfma_chain<<<blocks, threads>>>(out, iters);
        pde::count_launch();
        PDE_CHECK(cudaEventRecord(e1));
        PDE_CHECK(cudaEventSynchronize(e1));
        float ms = 0;
        PDE_CHECK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    *h_dfma_per_s = (double)blocks * threads * 8.0 * iters / (best * 1e-3);
    return 0;
}

int pde_fp64_probe(int device, int kind, double *h_units_per_s) {
    PDE_REQUIRE(h_units_per_s && kind >= 1 && kind <= 5, "bad argument");
    DeviceGuard guard(device);
    cudaDeviceProp prop;
    PDE_CHECK(cudaGetDeviceProperties(&prop, device));
    const int blocks = prop.multiProcessorCount * 2, threads = 256, iters = 1 << 13;
    double *out = nullptr;
    PDE_CHECK(cudaMalloc((void **)&out, sizeof(double) * blocks * threads));
    cudaEvent_t e0, e1;
    PDE_CHECK(cudaEventCreate(&e0));
    PDE_CHECK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        PDE_CHECK(cudaEventRecord(e0));
        switch (kind) {
            case 1: k_fp64_probe<1><<<blocks, threads>>>(out, iters); break;
            case 2: k_fp64_probe<2><<<blocks, threads>>>(out, iters); break;
            case 3: k_fp64_probe<3><<<blocks, threads>>>(out, iters); break;
            case 5: k_fp64_probe<5><<<blocks, threads>>>(out, iters); break;
            default: k_fp64_probe<4><<<blocks, threads>>>(out, iters); break;
        }
        pde::count_launch();
        PDE_CHECK(cudaEventRecord(e1));
        PDE_CHECK(cudaEventSynchronize(e1));
        float ms = 0;
        PDE_CHECK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    *h_units_per_s = (double)blocks * threads * 8.0 * iters / (best * 1e-3);
    return 0;
}

}  // extern "C"
