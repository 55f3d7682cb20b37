// Halo exchange of a row slab over NVLink peer memory (multigpu.py).
//
// Every rank owns a small STAGING buffer that its two slab neighbours can write to directly
// (peer-mapped memory: torch symmetric memory hands out the peers' device pointers).  One
// exchange is two tiny kernels on the rank's own stream, no collective and no stream fork:
//
//   k_halo_push        copies the rank's first / last H owned rows straight into the neighbours'
//                      staging slots (remote stores over NVLink), fences, and the last block
//                      publishes the exchange number in the neighbours' flag words;
//   k_halo_wait_copy   waits until both flag words show this exchange number (the neighbours
//                      push at the same point of their own stream, so the wait is the ranks'
//                      skew) and copies the staged rows into the halo rows of the vector.
//
// Staging slots alternate with the parity of the exchange number.  A rank can only be one
// exchange ahead of a neighbour (its next wait needs the neighbour's next push, which the
// neighbour issues after its own wait-copy of this exchange), so a slot is never overwritten
// while it is still being read.
//
// Layout of a rank's staging buffer (doubles), identical on all ranks:
//   [slot 0: from-lower (H x pitch) | from-upper (H x pitch)] [slot 1: same] [8 control words]
//   control: 0 = flag written by the lower neighbour, 1 = flag written by the upper neighbour,
//            2 = local block counter of k_halo_push
#include "common.cuh"

namespace pde {

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// rows_lo: my rows [H, 2H) -> lower neighbour's from-upper slot; rows_hi: my rows
// [nx_local, nx_local + H) -> upper neighbour's from-lower slot
__global__ void k_halo_push(const double2 *__restrict__ rows_lo, const double2 *__restrict__ rows_hi,
                            double2 *peer_lo_slot, double2 *peer_hi_slot, int64_t n2,
                            unsigned long long *peer_lo_flag, unsigned long long *peer_hi_flag,
                            unsigned *counter, unsigned long long step) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n2) {
        if (peer_lo_slot) peer_lo_slot[i] = rows_lo[i];
        if (peer_hi_slot) peer_hi_slot[i] = rows_hi[i];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned done = atomicAdd(counter, 1u);
        if (done == gridDim.x - 1) {               // every block's stores are fenced: publish
            *counter = 0u;
            __threadfence_system();
            if (peer_lo_flag) st_release_sys(peer_lo_flag, step);
            if (peer_hi_flag) st_release_sys(peer_hi_flag, step);
        }
    }
}

__global__ void k_halo_wait_copy(double2 *halo_lo, double2 *halo_hi, const double2 *from_lo,
                                 const double2 *from_hi, int64_t n2,
                                 const unsigned long long *flag_lo, const unsigned long long *flag_hi,
                                 unsigned long long step) {
    if (threadIdx.x == 0) {
        if (halo_lo) while (ld_acquire_sys(flag_lo) < step) __nanosleep(64);
        if (halo_hi) while (ld_acquire_sys(flag_hi) < step) __nanosleep(64);
    }
    __syncthreads();
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n2) {
        if (halo_lo) halo_lo[i] = from_lo[i];
        if (halo_hi) halo_hi[i] = from_hi[i];
    }
}

}  // namespace pde

using namespace pde;

extern "C" {

int64_t pde_halo_stage_doubles(int64_t halo, int64_t pitch) { return 4 * halo * pitch + 8; }

int pde_halo_push(const double *d_view, int64_t nx_local, int64_t halo, int64_t pitch,
                  double *my_stage, double *peer_lo_stage, double *peer_hi_stage,
                  uint64_t step, void *stream) {
    PDE_REQUIRE(d_view && my_stage && halo > 0 && nx_local >= halo && pitch % 2 == 0 && step > 0,
                "bad halo push arguments");
    const int64_t n = halo * pitch, slot = (int64_t)(step & 1) * 2 * n, ctl = 4 * n;
    PDE_REQUIRE((reinterpret_cast<uintptr_t>(d_view) & 15) == 0, "halo view must be 16-byte aligned");
    double *lo_slot = peer_lo_stage ? peer_lo_stage + slot + n : nullptr;   // its from-upper
    double *hi_slot = peer_hi_stage ? peer_hi_stage + slot : nullptr;       // its from-lower
    auto flag = [&](double *stage, int k) {
        return stage ? reinterpret_cast<unsigned long long *>(stage + ctl) + k : nullptr;
    };
    const unsigned blocks = (unsigned)((n / 2 + 255) / 256);
    k_halo_push<<<blocks, 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const double2 *>(d_view + halo * pitch),
        reinterpret_cast<const double2 *>(d_view + nx_local * pitch),
        reinterpret_cast<double2 *>(lo_slot), reinterpret_cast<double2 *>(hi_slot), n / 2,
        flag(peer_lo_stage, 1), flag(peer_hi_stage, 0),
        reinterpret_cast<unsigned *>(flag(my_stage, 2)), step);
    PDE_LAUNCH_CHECK();
    return 0;
}

int pde_halo_wait_copy(double *d_view, int64_t nx_local, int64_t halo, int64_t pitch,
                       const double *my_stage, int has_lo, int has_hi, uint64_t step,
                       void *stream) {
    PDE_REQUIRE(d_view && my_stage && halo > 0 && pitch % 2 == 0 && step > 0,
                "bad halo wait arguments");
    if (!has_lo && !has_hi) return 0;
    const int64_t n = halo * pitch, slot = (int64_t)(step & 1) * 2 * n, ctl = 4 * n;
    const unsigned long long *flags = reinterpret_cast<const unsigned long long *>(my_stage + ctl);
    const unsigned blocks = (unsigned)((n / 2 + 255) / 256);
    k_halo_wait_copy<<<blocks, 256, 0, (cudaStream_t)stream>>>(
        has_lo ? reinterpret_cast<double2 *>(d_view) : nullptr,
        has_hi ? reinterpret_cast<double2 *>(d_view + (nx_local + halo) * pitch) : nullptr,
        reinterpret_cast<const double2 *>(my_stage + slot),
        reinterpret_cast<const double2 *>(my_stage + slot + n), n / 2, flags, flags + 1, step);
    PDE_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
