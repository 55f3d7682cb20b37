// Stencils of a triangulated point cloud (FDTriMesh pre-processing).
//
// Replaces, per centre point, the reference's O(N^2) pipeline of FDTriMesh.__init__:
//   pointclasses.py:373-382  neighbours = points within `depth` edges of the
//                            triangulation graph (sum of adjacency-matrix powers)
//   pointclasses.py:385-392  kept when  0 < D <= max_radius  and
//                            (D >= min_radius  or  the neighbour is a boundary point)
//   pointclasses.py:37-103   colinear pruning, prefer="max": of two neighbours whose
//                            cosine distance is < tol the NEARER one is discarded (first
//                            on ties); boundary points are re-admitted unless they lost
//                            a boundary/interior double
//   pointclasses.py:105-121  simplices = ConvexHull of the unit stencil vectors; in 2-D
//                            every unit vector is a hull vertex and the hull edges join
//                            angular neighbours (cyclically), so the simplices are the
//                            consecutive pairs of the neighbours sorted by angle.
// One routine, `cloud_point`, does all of it for one centre point; it is compiled for
// the device (one thread per point, scratch in global memory, grid-stride loop) and for
// the host (OpenMP), operation by operation the same arithmetic (no FMA contraction),
// so the host build is the executable specification the device build is tested against
// and both are tested against the reference's neighbour / simplex sets.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "common.cuh"

namespace pde {

struct CloudParams {
    int64_t n, nint;
    const double *pts;        // (n, 2)
    const int64_t *adj_ptr;   // (n + 1)
    const int32_t *adj;       // neighbours in the triangulation graph
    int depth;
    double rmin, rmax, tol;
    int kmax;                 // capacity of the per-point output rows
    int qcap, hcap;           // BFS queue / hash capacities (hcap a power of two)
    int ccap;                 // candidate capacity
};

// Per-worker scratch (device: one slice of a global buffer per thread).
struct CloudScratch {
    unsigned long long *hash;   // hcap slots: (stamp << 32) | node
    int32_t *queue;             // qcap
    int32_t *cj;                // ccap candidate ids
    double *cvx, *cvy, *cd;     // ccap stencil vectors and lengths
    double *cang;               // ccap angles
    unsigned char *flag;        // ccap: bit0 discarded, bit1 lost a mixed double, bit2 boundary
};

__host__ __device__ inline size_t cloud_scratch_bytes(const CloudParams &P) {
    return (size_t)P.hcap * 8 + (size_t)P.qcap * 4 + (size_t)P.ccap * (4 + 8 * 4 + 1) + 64;
}

__host__ __device__ inline CloudScratch cloud_scratch_at(unsigned char *base, const CloudParams &P) {
    CloudScratch s;
    s.hash = reinterpret_cast<unsigned long long *>(base);
    base += (size_t)P.hcap * 8;
    s.cvx = reinterpret_cast<double *>(base);
    base += (size_t)P.ccap * 8;
    s.cvy = reinterpret_cast<double *>(base);
    base += (size_t)P.ccap * 8;
    s.cd = reinterpret_cast<double *>(base);
    base += (size_t)P.ccap * 8;
    s.cang = reinterpret_cast<double *>(base);
    base += (size_t)P.ccap * 8;
    s.queue = reinterpret_cast<int32_t *>(base);
    base += (size_t)P.qcap * 4;
    s.cj = reinterpret_cast<int32_t *>(base);
    base += (size_t)P.ccap * 4;
    s.flag = base;
    return s;
}

// returns 0 ok, 1 BFS overflow, 2 candidate overflow, 3 output overflow
__host__ __device__ inline int cloud_point(const CloudParams &P, int64_t i, unsigned stamp,
                                           CloudScratch &s, int32_t *out_count, int32_t *out_nb,
                                           int32_t *out_ring) {
    const unsigned long long tag = (unsigned long long)stamp << 32;
    const unsigned hmask = (unsigned)P.hcap - 1u;
    // ---- breadth-first search to `depth` edges (pointclasses.py:373-382) ----
    auto insert = [&](int32_t j) -> bool {       // true when j was not yet visited
        unsigned h = ((unsigned)j * 2654435761u) & hmask;
        for (;;) {
            unsigned long long v = s.hash[h];
            if ((v >> 32) != stamp) {
                s.hash[h] = tag | (unsigned)j;
                return true;
            }
            if ((int32_t)(v & 0xffffffffu) == j) return false;
            h = (h + 1) & hmask;
        }
    };
    int qn = 0;
    insert((int32_t)i);
    s.queue[qn++] = (int32_t)i;
    int lo = 0, hi = 1;
    for (int d = 0; d < P.depth && lo < hi; ++d) {
        for (int q = lo; q < hi; ++q) {
            int32_t a = s.queue[q];
            for (int64_t e = P.adj_ptr[a]; e < P.adj_ptr[a + 1]; ++e) {
                int32_t j = P.adj[e];
                if (insert(j)) {
                    if (qn >= P.qcap || 2 * qn >= P.hcap) return 1;
                    s.queue[qn++] = j;
                }
            }
        }
        lo = hi;
        hi = qn;
    }
    // ---- radius mask (pointclasses.py:385-392) ----
    const double xi = P.pts[2 * i], yi = P.pts[2 * i + 1];
    int K = 0;
    for (int q = 1; q < qn; ++q) {
        int32_t j = s.queue[q];
        double vx = P.pts[2 * (int64_t)j] - xi, vy = P.pts[2 * (int64_t)j + 1] - yi;
        double dd = sqrt(vx * vx + vy * vy);
        bool isb = j >= P.nint;
        if (dd <= P.rmax && (dd >= P.rmin || isb) && dd > 0.0) {
            if (K >= P.ccap) return 2;
            s.cj[K] = j;
            s.cvx[K] = vx;
            s.cvy[K] = vy;
            s.cd[K] = dd;
            s.flag[K] = isb ? 4 : 0;
            ++K;
        }
    }
    // order by point id (the reference walks its candidates in index order; this only
    // matters for exact distance ties) -- shell sort, all arrays in step
    for (int gap = K / 2; gap > 0; gap /= 2) {
        for (int a = gap; a < K; ++a) {
            int32_t j = s.cj[a];
            double vx = s.cvx[a], vy = s.cvy[a], dd = s.cd[a];
            unsigned char f = s.flag[a];
            int b = a;
            while (b >= gap && s.cj[b - gap] > j) {
                s.cj[b] = s.cj[b - gap];
                s.cvx[b] = s.cvx[b - gap];
                s.cvy[b] = s.cvy[b - gap];
                s.cd[b] = s.cd[b - gap];
                s.flag[b] = s.flag[b - gap];
                b -= gap;
            }
            s.cj[b] = j;
            s.cvx[b] = vx;
            s.cvy[b] = vy;
            s.cd[b] = dd;
            s.flag[b] = f;
        }
    }
    // ---- colinear pruning, prefer = "max" (pointclasses.py:37-103) ----
    for (int a = 0; a < K; ++a) {
        const double ax = s.cvx[a], ay = s.cvy[a], ad = s.cd[a];
        const bool ab = (s.flag[a] & 4) != 0;
        for (int b = a + 1; b < K; ++b) {
            // scipy pdist(..., 'cosine'): 1 - u.v / (|u| |v|), cosine clipped to [-1, 1]
            double c = (ax * s.cvx[b] + ay * s.cvy[b]) / (ad * s.cd[b]);
            if (fabs(c) > 1.0) c = c < 0 ? -1.0 : 1.0;
            if (1.0 - c < P.tol) {
                int loser = (ad <= s.cd[b]) ? a : b;        // np.argmin: first on ties
                bool mixed = ab != ((s.flag[b] & 4) != 0);
                s.flag[loser] |= mixed ? 3 : 1;
            }
        }
    }
    int Kk = 0;
    for (int a = 0; a < K; ++a) {
        unsigned char f = s.flag[a];
        bool keep = !(f & 1) || ((f & 4) && !(f & 2));
        if (keep) {
            if (Kk >= P.kmax) return 3;
            s.cj[Kk] = s.cj[a];
            s.cang[Kk] = atan2(s.cvy[a], s.cvx[a]);
            out_nb[Kk] = s.cj[a];
            ++Kk;
        }
    }
    *out_count = Kk;
    // ---- angular order -> ring of hull edges (pointclasses.py:105-121) ----
    for (int a = 0; a < Kk; ++a) s.queue[a] = a;
    for (int gap = Kk / 2; gap > 0; gap /= 2) {
        for (int a = gap; a < Kk; ++a) {
            int32_t k = s.queue[a];
            double ang = s.cang[k];
            int b = a;
            while (b >= gap && s.cang[s.queue[b - gap]] > ang) {
                s.queue[b] = s.queue[b - gap];
                b -= gap;
            }
            s.queue[b] = k;
        }
    }
    for (int a = 0; a < Kk; ++a) out_ring[a] = s.cj[s.queue[a]];
    return 0;
}

__global__ void k_cloud_build(const CloudParams P, unsigned char *scratch, size_t stride,
                              int32_t *count, int32_t *nb, int32_t *ring, int32_t *status) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    CloudScratch s = cloud_scratch_at(scratch + stride * t, P);
    unsigned stamp = 0;
    for (int64_t i = t; i < P.n; i += nthreads) {
        ++stamp;
        int rc = cloud_point(P, i, stamp, s, count + i, nb + i * P.kmax, ring + i * P.kmax);
        if (rc) {
            count[i] = -rc;
            atomicAdd(&status[rc], 1);
        }
    }
}

}  // namespace pde

using namespace pde;

extern "C" {

int pde_cloud_build(int device, int64_t n_points, int64_t n_interior, const double *points,
                    const int64_t *adj_ptr, const int32_t *adj, int depth, double min_radius,
                    double max_radius, double colinear_tol, int kmax, int bfs_capacity,
                    int32_t *out_count, int32_t *out_nb, int32_t *out_ring, int32_t *out_status,
                    void *stream) {
    PDE_REQUIRE(points && adj_ptr && adj && out_count && out_nb && out_ring && out_status,
                "null argument");
    PDE_REQUIRE(n_points < (int64_t)1 << 31, "point ids must fit in 32 bits");
    PDE_REQUIRE(depth >= 1 && kmax >= 3 && bfs_capacity >= 16, "bad depth / capacity");
    PDE_REQUIRE(kmax <= bfs_capacity, "kmax must not exceed the BFS capacity");
    CloudParams P;
    P.n = n_points;
    P.nint = n_interior;
    P.pts = points;
    P.adj_ptr = adj_ptr;
    P.adj = adj;
    P.depth = depth;
    P.rmin = min_radius;
    P.rmax = max_radius;
    P.tol = colinear_tol;
    P.kmax = kmax;
    P.qcap = bfs_capacity;
    int h = 64;
    while (h < 4 * bfs_capacity) h *= 2;
    P.hcap = h;
    P.ccap = bfs_capacity;
    const size_t stride = (cloud_scratch_bytes(P) + 63) / 64 * 64;

    if (device < 0) {     // host build (OpenMP): same routine, host pointers
        int nthreads = 1;
#ifdef _OPENMP
        nthreads = omp_get_max_threads();
#endif
        std::vector<unsigned char> scratch(stride * (size_t)nthreads, 0);
        int st[4] = {0, 0, 0, 0};
#pragma omp parallel
        {
            int tid = 0;
#ifdef _OPENMP
            tid = omp_get_thread_num();
#endif
            CloudScratch s = cloud_scratch_at(scratch.data() + stride * (size_t)tid, P);
            unsigned stamp = 0;
#pragma omp for schedule(dynamic, 64)
            for (int64_t i = 0; i < n_points; ++i) {
                ++stamp;
                int rc = cloud_point(P, i, stamp, s, out_count + i, out_nb + i * kmax,
                                     out_ring + i * kmax);
                if (rc) {
                    out_count[i] = -rc;
#pragma omp atomic
                    st[rc] += 1;
                }
            }
        }
        for (int k = 0; k < 4; ++k) out_status[k] = st[k];
        return 0;
    }

    DeviceGuard guard(device);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    cudaDeviceProp prop;
    PDE_CHECK(cudaGetDeviceProperties(&prop, device));
    const int threads = 64;
    int64_t blocks = (int64_t)prop.multiProcessorCount * 8;
    blocks = std::min<int64_t>(blocks, (n_points + threads - 1) / threads);
    // bound the scratch (one slice per resident thread) to 8 GiB
    blocks = std::min<int64_t>(blocks, (int64_t)(((size_t)8 << 30) / (stride * threads)));
    blocks = std::max<int64_t>(blocks, 1);
    unsigned char *scratch = nullptr;
    const size_t bytes = stride * (size_t)(blocks * threads);
    PDE_CHECK(cudaMallocAsync(&scratch, bytes, st));
    PDE_CHECK(cudaMemsetAsync(scratch, 0, bytes, st));
    PDE_CHECK(cudaMemsetAsync(out_status, 0, 4 * sizeof(int32_t), st));
    k_cloud_build<<<(unsigned)blocks, threads, 0, st>>>(P, scratch, stride, out_count, out_nb,
                                                       out_ring, out_status);
    PDE_LAUNCH_CHECK();
    PDE_CHECK(cudaFreeAsync(scratch, st));
    return 0;
}

}  // extern "C"
