// Shared helpers for the pde_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>
#include <cudaTypedefs.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/pde_b200.h"

namespace pde {

// ---- error plumbing ---------------------------------------------------------
void set_error(const std::string &msg);
int fail(const char *what, const char *file, int line);
int fail_cuda(cudaError_t e, const char *what, const char *file, int line);
void count_launch(int n = 1);
int64_t launch_count();
// keep the stream-ordered allocator's memory across synchronisations (the default release
// threshold of 0 hands every freed block back to the driver at the next sync, which made each
// cudaMallocAsync of the solver loops a real allocation: 12 ms per Newton update, measured)
void keep_async_pool(int device);

#define PDE_CHECK(call)                                                        \
    do {                                                                       \
        cudaError_t _e = (call);                                               \
        if (_e != cudaSuccess) return pde::fail_cuda(_e, #call, __FILE__, __LINE__); \
    } while (0)
#define PDE_REQUIRE(cond, msg)                                                 \
    do {                                                                       \
        if (!(cond)) return pde::fail(msg, __FILE__, __LINE__);                \
    } while (0)
#define PDE_LAUNCH_CHECK()                                                     \
    do {                                                                       \
        pde::count_launch();                                                   \
        cudaError_t _e = cudaGetLastError();                                   \
        if (_e != cudaSuccess) return pde::fail_cuda(_e, "kernel launch", __FILE__, __LINE__); \
    } while (0)

struct DeviceGuard {
    int prev;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (dev != prev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// ---- the deep-interior direction table, passed by value as a kernel param ----
struct DeepTable {
    int D;        // antipodal pairs per point
    int H;        // halo = max |offset|
    int a0[PDE_MAX_DIRS], b0[PDE_MAX_DIRS], a1[PDE_MAX_DIRS], b1[PDE_MAX_DIRS];
    double w_0[PDE_MAX_DIRS], w_1[PDE_MAX_DIRS], w0[PDE_MAX_DIRS];
};

}  // namespace pde

// Grid context (opaque in the C ABI).
struct pde_grid {
    int device;
    int64_t nx_global, ny;       // interior lattice
    int64_t x_offset, nx_local;  // rows owned by this context (slab)
    int64_t halo_rows;           // extra rows stored above/below the owned rows
    int64_t rows;                // nx_local + 2*halo_rows
    int64_t pitch;               // doubles per row (multiple of 16)
    int64_t nb;                  // wall points
    int radius;                  // stencil radius r: deep = [r, n-r) on both axes
    pde::DeepTable tab;
    int ct_id;                   // index of the matching compile-time table, or -1
    // 3-D lattices (pde_grid_create_nd): [nx3][nmid][ny] stored as nx3*nmid rows; the table's
    // row offsets a0/a1 are flattened (a*nmid + b), nd_abc keeps the (a,b,c | a',b',c') triples
    int64_t nmid, nx3;           // 0, 0 for 2-D lattices
    int ct3_id;                  // matching compile-time 3-D table, or -1
    int nd_abc[PDE_MAX_DIRS][6];
    double stencil_X[PDE_MAX_DIRS][2];  // physical vector of (a0,b0)
    // band (CSR over band points owned by this context)
    int64_t n_band, n_band_pairs;
    int64_t win_lo, win_hi;      // owned rows [win_lo, win_hi) the stencil launches act on
    int64_t win2_lo, win2_hi;    // optional second range handled by the same launches (empty: 0, 0)
    int dyn_sched;               // stencil CTAs draw tiles from a device counter (d_ctrl[8..9])
    int64_t *h_band_row;         // host: owned-row number of every band point (ascending)
    int64_t *d_band_center;      // n_band: state offset of the centre
    int64_t *d_band_pcenter;     // n_band: padded-layout offset (policy / per-point outputs)
    int64_t *d_band_ptr;         // n_band+1
    int64_t *d_band_j;           // P x 2 state offsets of J0, J1
    double *d_band_w;            // P x 3
    double *d_band_X;            // P x 2 physical vector of J0 - I
    // scratch for reductions / control words
    double *d_red;               // 16 doubles
    unsigned long long *d_ctrl;  // 16 words
    cudaStream_t side;           // band kernel runs here, concurrently with the deep kernel
    cudaEvent_t ev_fork, ev_join;
    cudaEvent_t *ev_pool;        // events of the host-to-host pipeline (pde_ddg_d2eigs_host)
    int n_ev_pool;
    // band translation classes (pde_grid_set_band_classes): the rows of a band point as a class
    // table (relative lattice offsets / positions in the point's wall list + weights) instead of
    // 40-byte explicit rows, for the operator kernels of 3-D lattices
    int n_cls;
    int32_t *d_cls_ptr;          // n_cls + 1
    int32_t *d_cls_ref;          // rows x 2: (delta << 1) lattice | (position << 1 | 1) wall list
    double *d_cls_w;             // rows x 3
    uint16_t *d_cls_id;          // n_band
    int32_t *d_wl_ptr;           // n_band + 1
    int32_t *d_wl;               // state offsets of the wall points a band point refers to
    PFN_cuTensorMapEncodeTiled_v12000 encode;
    int num_sms;
};

// ---- device helpers ---------------------------------------------------------------
namespace pde {

__device__ __forceinline__ void atomic_max_nonneg(double *addr, double v) {
    // valid for non-negative finite doubles: IEEE order == unsigned integer order
    atomicMax(reinterpret_cast<unsigned long long *>(addr),
              static_cast<unsigned long long>(__double_as_longlong(v)));
}

__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Reference operation order of ddg.py:281 -- no contraction, no re-association.
__device__ __forceinline__ double pair_value_exact(double uc, double u0, double u1,
                                                   double w_0, double w_1, double w0) {
    double t0 = __dmul_rn(__dsub_rn(u0, uc), w_0);
    double t1 = __dmul_rn(__dsub_rn(u1, uc), w_1);
    return __dmul_rn(w0, __dadd_rn(t0, t1));
}

}  // namespace pde
