// Host-side construction of the band stencils of a regular pairs grid.
//
// Replaces, for points within `r` cells of a wall, the reference's O(N^2) pipeline
// (pointclasses.py:421-471 joggled Delaunay + sparse matrix powers, :37-103 colinear
// pruning, :12-35 antipodal pairing).  Membership rules are the reference's:
//   neighbours : l-inf distance in (0, r] on the unscaled coordinates (:443-447)
//   pruning    : cosine distance < tol -> drop the farther one (first on ties); wall
//                points are re-admitted unless they lost in a wall/interior double
//   pairing    : normalised vectors with cos < -1 + tol, upper triangular order
// The arithmetic (scaled vectors, norms, cosines) is evaluated operation by operation
// like the NumPy expressions of pyellipticfd_b200/pointclasses.py::_prune_and_pair,
// which is the executable specification this file is tested against.  Pure host code
// (OpenMP over band points); compiled with -ffp-contract=off.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "common.cuh"

namespace {

struct WallList {
    int axis;             // 0: wall x = val, 1: wall y = val
    double val;
    std::vector<int64_t> ids;   // wall point numbers, sorted along the wall
    std::vector<double> t;      // coordinate along the wall
};

}  // namespace

extern "C" {

void pde_free(void *p) { free(p); }

int pde_set_host_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
#else
    (void)n;
    return 1;
#endif
}

int pde_band_build(int64_t nx, int64_t ny, int r, double sx, double sy, double bx, double by,
                   int64_t nb, const double *wall /* nb x 2 unscaled */, int64_t n_band,
                   const int64_t *band_ids, double tol, int64_t **out_ptr, int64_t **out_pairs,
                   double *out_min_nb_dist, int64_t **out_nb_ptr, int64_t **out_nb) {
    PDE_REQUIRE(out_ptr && out_pairs && out_min_nb_dist && (n_band == 0 || band_ids), "null argument");
    const int64_t N = nx * ny;
    // four wall lists, as in the Python reference implementation
    WallList lists[4];
    const double vals[4] = {0.0, (double)(nx + 1), 0.0, (double)(ny + 1)};
    for (int w = 0; w < 4; ++w) {
        WallList &L = lists[w];
        L.axis = w < 2 ? 0 : 1;
        L.val = vals[w];
        std::vector<std::pair<double, int64_t>> tmp;
        for (int64_t i = 0; i < nb; ++i) {
            double x = wall[2 * i], y = wall[2 * i + 1];
            bool on = (L.axis == 0 ? x : y) == L.val;
            if (L.axis == 1 && (x == 0.0 || x == (double)(nx + 1))) on = false;   // corners once
            if (on) tmp.emplace_back(L.axis == 0 ? y : x, i);
        }
        std::stable_sort(tmp.begin(), tmp.end(),
                         [](const std::pair<double, int64_t> &a, const std::pair<double, int64_t> &b) {
                             return a.first < b.first;
                         });
        for (auto &p : tmp) {
            L.t.push_back(p.first);
            L.ids.push_back(p.second);
        }
    }

    std::vector<std::vector<int64_t>> rows(n_band);   // per point: flattened (J0, J1)
    const bool want_nb = out_nb_ptr && out_nb;
    std::vector<std::vector<int64_t>> nbs(want_nb ? n_band : 0);   // per point: kept neighbours
    std::vector<double> min_nb(n_band, INFINITY);
    int64_t bad_point = -1;

#pragma omp parallel
    {
        std::vector<double> PX, PY, VX, VY, D, SX, SY;
        std::vector<int64_t> ID;
        std::vector<char> isb, disc, discm, keep;
#pragma omp for schedule(dynamic, 16)
        for (int64_t p = 0; p < n_band; ++p) {
            const int64_t id = band_ids[p];
            const double cx = (double)(id / ny + 1), cy = (double)(id % ny + 1);
            PX.clear(); PY.clear(); ID.clear(); isb.clear();
            for (int a = -r; a <= r; ++a)
                for (int b = -r; b <= r; ++b) {
                    if (a == 0 && b == 0) continue;
                    double lx = cx + a, ly = cy + b;
                    if (lx >= 1 && lx <= (double)nx && ly >= 1 && ly <= (double)ny) {
                        PX.push_back(lx); PY.push_back(ly);
                        ID.push_back((int64_t)(lx - 1) * ny + (int64_t)(ly - 1));
                        isb.push_back(0);
                    }
                }
            for (int w = 0; w < 4; ++w) {
                const WallList &L = lists[w];
                double c_al = L.axis == 0 ? cy : cx, c_ac = L.axis == 0 ? cx : cy;
                if (!(std::fabs(c_ac - L.val) <= r) || L.ids.empty()) continue;
                auto lo = std::lower_bound(L.t.begin(), L.t.end(), c_al - r - 1e-9);
                auto hi = std::upper_bound(L.t.begin(), L.t.end(), c_al + r + 1e-9);
                for (auto it = lo; it != hi; ++it) {
                    int64_t wid = L.ids[it - L.t.begin()];
                    double wx = wall[2 * wid], wy = wall[2 * wid + 1];
                    double dinf = std::max(std::fabs(wx - cx), std::fabs(wy - cy));
                    if (dinf <= r && dinf > 0) {
                        PX.push_back(wx); PY.push_back(wy);
                        ID.push_back(N + wid);
                        isb.push_back(1);
                    }
                }
            }
            const int K = (int)PX.size();
            VX.resize(K); VY.resize(K); D.resize(K); SX.resize(K); SY.resize(K);
            disc.assign(K, 0); discm.assign(K, 0); keep.assign(K, 0);
            const double pcx = cx * sx + bx, pcy = cy * sy + by;
            for (int k = 0; k < K; ++k) {
                VX[k] = (PX[k] * sx + bx) - pcx;
                VY[k] = (PY[k] * sy + by) - pcy;
                D[k] = std::sqrt(VX[k] * VX[k] + VY[k] * VY[k]);
            }
            for (int k = 0; k < K; ++k)
                for (int l = k + 1; l < K; ++l) {
                    double dot = VX[k] * VX[l] + VY[k] * VY[l];
                    double cosv = dot / (D[k] * D[l]);
                    cosv = cosv < -1.0 ? -1.0 : (cosv > 1.0 ? 1.0 : cosv);
                    if ((1.0 - cosv) < tol) {
                        int loser = D[k] >= D[l] ? k : l;
                        disc[loser] = 1;
                        if (isb[k] != isb[l]) discm[loser] = 1;
                    }
                }
            double mnb = INFINITY;
            for (int k = 0; k < K; ++k) {
                keep[k] = !disc[k] || (isb[k] && !discm[k]);
                SX[k] = VX[k] / D[k];
                SY[k] = VY[k] / D[k];
                if (keep[k]) mnb = std::min(mnb, D[k]);
            }
            min_nb[p] = mnb;
            if (want_nb)
                for (int k = 0; k < K; ++k)
                    if (keep[k]) nbs[p].push_back(ID[k]);
            std::vector<int64_t> &out = rows[p];
            for (int k = 0; k < K; ++k) {
                if (!keep[k]) continue;
                for (int l = k + 1; l < K; ++l) {
                    if (!keep[l]) continue;
                    double cs = SX[k] * SX[l] + SY[k] * SY[l];
                    if (cs < -1 + tol) {
                        out.push_back(ID[k]);
                        out.push_back(ID[l]);
                    }
                }
            }
            if (out.empty()) {
#pragma omp critical
                if (bad_point < 0 || id < bad_point) bad_point = id;
            }
        }
    }
    if (bad_point >= 0) {
        char buf[96];
        snprintf(buf, sizeof buf, "Point %lld has no pairs", (long long)bad_point);
        pde::set_error(buf);
        return 2;
    }
    int64_t *ptr = (int64_t *)malloc(sizeof(int64_t) * (n_band + 1));
    PDE_REQUIRE(ptr, "out of memory");
    ptr[0] = 0;
    for (int64_t p = 0; p < n_band; ++p) ptr[p + 1] = ptr[p] + (int64_t)rows[p].size() / 2;
    int64_t P = ptr[n_band];
    int64_t *pairs = (int64_t *)malloc(sizeof(int64_t) * 3 * (P > 0 ? P : 1));
    if (!pairs) {
        free(ptr);
        return pde::fail("out of memory", __FILE__, __LINE__);
    }
#pragma omp parallel for schedule(static)
    for (int64_t p = 0; p < n_band; ++p) {
        int64_t o = ptr[p];
        for (size_t k = 0; k < rows[p].size() / 2; ++k) {
            pairs[3 * (o + k)] = band_ids[p];
            pairs[3 * (o + k) + 1] = rows[p][2 * k];
            pairs[3 * (o + k) + 2] = rows[p][2 * k + 1];
        }
    }
    if (want_nb) {
        int64_t *nptr = (int64_t *)malloc(sizeof(int64_t) * (n_band + 1));
        nptr[0] = 0;
        for (int64_t p = 0; p < n_band; ++p) nptr[p + 1] = nptr[p] + (int64_t)nbs[p].size();
        int64_t *nb_out = (int64_t *)malloc(sizeof(int64_t) * (nptr[n_band] > 0 ? nptr[n_band] : 1));
        for (int64_t p = 0; p < n_band; ++p)
            std::copy(nbs[p].begin(), nbs[p].end(), nb_out + nptr[p]);
        *out_nb_ptr = nptr;
        *out_nb = nb_out;
    }
    double m = INFINITY;
    for (int64_t p = 0; p < n_band; ++p) m = std::min(m, min_nb[p]);
    *out_ptr = ptr;
    *out_pairs = pairs;
    *out_min_nb_dist = m;
    return 0;
}

}  // extern "C"
