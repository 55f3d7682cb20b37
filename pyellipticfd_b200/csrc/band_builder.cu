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
#include <unordered_map>
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

struct ClassEntry {                  // result of the stencil rules for one translation class
    std::vector<double> V;           // candidate vectors of the representative (K x dim)
    std::vector<char> isb;           // candidate is a wall point
    std::vector<int> kept, pk, pl;   // kept candidates; paired candidates (k < l)
};

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
            // unscaled wall coordinates carry the rounding of pointclasses.py:538
            // (4000.9999999999995 for 4001): compare with a tolerance far below the 1/r spacing
            bool on = std::fabs((L.axis == 0 ? x : y) - L.val) < 1e-6;
            if (L.axis == 1 && (std::fabs(x) < 1e-6 || std::fabs(x - (double)(nx + 1)) < 1e-6))
                on = false;                                                        // corners once
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
        std::unordered_map<int64_t, ClassEntry> cache;       // translation classes, per thread
        const double vscale = std::min(std::fabs(sx), std::fabs(sy));
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
            VX.resize(K); VY.resize(K); D.resize(K);
            const double pcx = cx * sx + bx, pcy = cy * sy + by;
            for (int k = 0; k < K; ++k) {
                VX[k] = (PX[k] * sx + bx) - pcx;
                VY[k] = (PY[k] * sy + by) - pcy;
                D[k] = std::sqrt(VX[k] * VX[k] + VY[k] * VY[k]);
            }
            // translation classes (see pde_band_build_nd): same wall distances, same candidate
            // vectors -> the kept / paired candidate numbers of the class representative
            const int64_t i1 = id / ny + 1, j1 = id % ny + 1;
            int64_t sig = std::min<int64_t>(i1 - 1, r + 2);
            sig = sig * (r + 3) + std::min<int64_t>(nx - i1, r + 2);
            sig = sig * (r + 3) + std::min<int64_t>(j1 - 1, r + 2);
            sig = sig * (r + 3) + std::min<int64_t>(ny - j1, r + 2);
            ClassEntry *ce = nullptr;
            {
                auto it = cache.find(sig);
                if (it != cache.end() && (int)it->second.isb.size() == K) {
                    ClassEntry &e = it->second;
                    bool same = true;
                    for (int k = 0; k < K && same; ++k)
                        same = e.isb[k] == isb[k] && std::fabs(e.V[2 * k] - VX[k]) <= 1e-9 * vscale &&
                               std::fabs(e.V[2 * k + 1] - VY[k]) <= 1e-9 * vscale;
                    if (same) ce = &e;
                }
            }
            if (!ce) {
                SX.resize(K); SY.resize(K);
                disc.assign(K, 0); discm.assign(K, 0); keep.assign(K, 0);
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
                ClassEntry e;
                e.V.resize((size_t)2 * K);
                e.isb = isb;
                for (int k = 0; k < K; ++k) {
                    e.V[2 * k] = VX[k];
                    e.V[2 * k + 1] = VY[k];
                    keep[k] = !disc[k] || (isb[k] && !discm[k]);
                    SX[k] = VX[k] / D[k];
                    SY[k] = VY[k] / D[k];
                    if (keep[k]) e.kept.push_back(k);
                }
                for (int k = 0; k < K; ++k) {
                    if (!keep[k]) continue;
                    for (int l = k + 1; l < K; ++l) {
                        if (!keep[l]) continue;
                        double cs = SX[k] * SX[l] + SY[k] * SY[l];
                        if (cs < -1 + tol) {
                            e.pk.push_back(k);
                            e.pl.push_back(l);
                        }
                    }
                }
                ce = &(cache[sig] = std::move(e));
            }
            double mnb = INFINITY;
            for (int k : ce->kept) mnb = std::min(mnb, D[k]);
            min_nb[p] = mnb;
            if (want_nb)
                for (int k : ce->kept) nbs[p].push_back(ID[k]);
            std::vector<int64_t> &out = rows[p];
            out.reserve(2 * ce->pk.size());
            for (size_t q = 0; q < ce->pk.size(); ++q) {
                out.push_back(ID[ce->pk[q]]);
                out.push_back(ID[ce->pl[q]]);
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


// ---------------------------------------------------------------------------------------------
// n-D (dim >= 2; used for 3-D) version of pde_band_build: the same three rules on lattices of any
// dimension.  Wall points are bucketed by integer cell (floor of the unscaled coordinates) in a
// sorted key array; a band point visits the cells of its l-inf box that touch a wall.  Lattice
// candidates are enumerated in C order of the offset (last axis fastest), then the wall candidates
// in ascending wall-point number: the row order of a point's group is this library's own (the
// reference's is an artefact of its joggled Delaunay triangulation and only matters at exact ties).
int pde_band_build_nd(int dim, const int64_t *shape, int r, const double *scaling,
                      const double *bounds0, int64_t nb, const double *wall /* nb x dim */,
                      int64_t n_band, const int64_t *band_ids, double tol, int64_t **out_ptr,
                      int64_t **out_pairs, double *out_min_nb_dist, int64_t **out_nb_ptr,
                      int64_t **out_nb) {
    PDE_REQUIRE(dim >= 2 && dim <= 4, "pde_band_build_nd: dim must be 2, 3 or 4");
    PDE_REQUIRE(shape && scaling && bounds0 && out_ptr && out_pairs && out_min_nb_dist &&
                    (n_band == 0 || band_ids) && (nb == 0 || wall),
                "null argument");
    int64_t N = 1, cstride[4], lstride[4];
    for (int d = 0; d < dim; ++d) N *= shape[d];
    lstride[dim - 1] = 1;
    cstride[dim - 1] = 1;
    for (int d = dim - 2; d >= 0; --d) {
        lstride[d] = lstride[d + 1] * shape[d + 1];
        cstride[d] = cstride[d + 1] * (shape[d + 1] + 2);
    }
    double vscale = INFINITY;
    for (int d = 0; d < dim; ++d) vscale = std::min(vscale, std::fabs(scaling[d]));
    // wall points sorted by cell key
    std::vector<std::pair<int64_t, int64_t>> cells(nb);
    for (int64_t i = 0; i < nb; ++i) {
        int64_t key = 0;
        for (int d = 0; d < dim; ++d) {
            int64_t c = (int64_t)std::floor(wall[dim * i + d]);
            c = c < 0 ? 0 : (c > shape[d] + 1 ? shape[d] + 1 : c);
            key += c * cstride[d];
        }
        cells[i] = {key, i};
    }
    std::sort(cells.begin(), cells.end());

    std::vector<std::vector<int64_t>> rows(n_band);
    const bool want_nb = out_nb_ptr && out_nb;
    std::vector<std::vector<int64_t>> nbs(want_nb ? n_band : 0);
    std::vector<double> min_nb(n_band, INFINITY);
    int64_t bad_point = -1;

#pragma omp parallel
    {
        std::vector<double> V, D, S;      // K x dim, K, K x dim
        std::vector<int64_t> ID, wids;
        std::vector<char> isb, disc, discm, keep;
        std::unordered_map<int64_t, ClassEntry> cache;       // translation classes, per thread
#pragma omp for schedule(dynamic, 8)
        for (int64_t p = 0; p < n_band; ++p) {
            const int64_t id = band_ids[p];
            double c[4], pc[4];
            int64_t ci[4];
            {
                int64_t rem = id;
                for (int d = 0; d < dim; ++d) {
                    ci[d] = rem / lstride[d] + 1;
                    rem %= lstride[d];
                    c[d] = (double)ci[d];
                    pc[d] = c[d] * scaling[d] + bounds0[d];
                }
            }
            V.clear(); ID.clear(); isb.clear();
            // lattice candidates: odometer over the offset box
            int o[4];
            for (int d = 0; d < dim; ++d) o[d] = -r;
            for (;;) {
                bool zero = true, inside = true;
                int64_t lid = 0;
                for (int d = 0; d < dim; ++d) {
                    zero &= o[d] == 0;
                    int64_t x = ci[d] + o[d];
                    inside &= x >= 1 && x <= shape[d];
                    lid += (x - 1) * lstride[d];
                }
                if (!zero && inside) {
                    for (int d = 0; d < dim; ++d)
                        V.push_back(((c[d] + o[d]) * scaling[d] + bounds0[d]) - pc[d]);
                    ID.push_back(lid);
                    isb.push_back(0);
                }
                int d = dim - 1;
                while (d >= 0 && ++o[d] > r) o[d--] = -r;
                if (d < 0) break;
            }
            // wall candidates: cells [c-r-1, c+r] per axis that lie on a wall
            wids.clear();
            int64_t lo[4], hi[4], cc[4];
            bool near_wall = false;
            for (int d = 0; d < dim; ++d) {
                lo[d] = std::max<int64_t>(0, ci[d] - r - 1);
                hi[d] = std::min<int64_t>(shape[d] + 1, ci[d] + r);
                near_wall |= lo[d] == 0 || hi[d] >= shape[d];
                cc[d] = lo[d];
            }
            if (near_wall && nb > 0)
                for (;;) {
                    bool on_wall = false;
                    int64_t key = 0;
                    for (int d = 0; d < dim; ++d) {
                        on_wall |= cc[d] == 0 || cc[d] >= shape[d];   // cell n too: (n+1)-eps after rounding
                        key += cc[d] * cstride[d];
                    }
                    if (on_wall) {
                        auto it = std::lower_bound(cells.begin(), cells.end(),
                                                   std::make_pair(key, (int64_t)-1));
                        for (; it != cells.end() && it->first == key; ++it) {
                            const double *w = wall + dim * it->second;
                            double dinf = 0;
                            for (int d = 0; d < dim; ++d) dinf = std::max(dinf, std::fabs(w[d] - c[d]));
                            if (dinf <= r && dinf > 0) wids.push_back(it->second);
                        }
                    }
                    int d = dim - 1;
                    while (d >= 0 && ++cc[d] > hi[d]) { cc[d] = lo[d]; --d; }
                    if (d < 0) break;
                }
            std::sort(wids.begin(), wids.end());
            for (int64_t wid : wids) {
                const double *w = wall + dim * wid;
                for (int d = 0; d < dim; ++d) V.push_back((w[d] * scaling[d] + bounds0[d]) - pc[d]);
                ID.push_back(N + wid);
                isb.push_back(1);
            }
            const int K = (int)ID.size();
            D.resize(K);
            for (int k = 0; k < K; ++k) {
                double s = 0;
                for (int d = 0; d < dim; ++d) s += V[(size_t)k * dim + d] * V[(size_t)k * dim + d];
                D[k] = std::sqrt(s);
            }
            // Translation classes: band points with the same distances to the walls (capped at
            // r + 2 per side) see the same candidate vectors, hence the same pruning and pairing.
            // The O(K^2) rules run once per class and thread; every other point of the class
            // re-uses the kept / paired candidate NUMBERS after checking that its candidate
            // vectors equal the class representative's to 1e-9 of the cell size (inexact
            // lattices differ in the last bits; decisions have a tolerance of 1e-3).
            int64_t sig = 0;
            for (int d = 0; d < dim; ++d) {
                const int64_t a = std::min<int64_t>(ci[d] - 1, r + 2);
                const int64_t b = std::min<int64_t>(shape[d] - ci[d], r + 2);
                sig = (sig * (r + 3) + a) * (r + 3) + b;
            }
            ClassEntry *ce = nullptr;
            {
                auto it = cache.find(sig);
                if (it != cache.end() && (int)it->second.isb.size() == K) {
                    ClassEntry &e = it->second;
                    bool same = true;
                    for (int k = 0; k < K && same; ++k) same = e.isb[k] == isb[k];
                    for (size_t q = 0; q < (size_t)K * dim && same; ++q)
                        same = std::fabs(e.V[q] - V[q]) <= 1e-9 * vscale;
                    if (same) ce = &e;
                }
            }
            if (!ce) {
                S.resize((size_t)K * dim);
                disc.assign(K, 0); discm.assign(K, 0); keep.assign(K, 0);
                for (int k = 0; k < K; ++k)
                    for (int l = k + 1; l < K; ++l) {
                        double dot = 0;
                        for (int d = 0; d < dim; ++d) dot += V[(size_t)k * dim + d] * V[(size_t)l * dim + d];
                        double cosv = dot / (D[k] * D[l]);
                        cosv = cosv < -1.0 ? -1.0 : (cosv > 1.0 ? 1.0 : cosv);
                        if ((1.0 - cosv) < tol) {
                            int loser = D[k] >= D[l] ? k : l;
                            disc[loser] = 1;
                            if (isb[k] != isb[l]) discm[loser] = 1;
                        }
                    }
                ClassEntry e;
                e.V = V;
                e.isb = isb;
                for (int k = 0; k < K; ++k) {
                    keep[k] = !disc[k] || (isb[k] && !discm[k]);
                    for (int d = 0; d < dim; ++d) S[(size_t)k * dim + d] = V[(size_t)k * dim + d] / D[k];
                    if (keep[k]) e.kept.push_back(k);
                }
                for (int k = 0; k < K; ++k) {
                    if (!keep[k]) continue;
                    for (int l = k + 1; l < K; ++l) {
                        if (!keep[l]) continue;
                        double cs = 0;
                        for (int d = 0; d < dim; ++d) cs += S[(size_t)k * dim + d] * S[(size_t)l * dim + d];
                        if (cs < -1 + tol) {
                            e.pk.push_back(k);
                            e.pl.push_back(l);
                        }
                    }
                }
                ce = &(cache[sig] = std::move(e));
            }
            double mnb = INFINITY;
            for (int k : ce->kept) mnb = std::min(mnb, D[k]);
            min_nb[p] = mnb;
            if (want_nb)
                for (int k : ce->kept) nbs[p].push_back(ID[k]);
            std::vector<int64_t> &out = rows[p];
            out.reserve(2 * ce->pk.size());
            for (size_t q = 0; q < ce->pk.size(); ++q) {
                out.push_back(ID[ce->pk[q]]);
                out.push_back(ID[ce->pl[q]]);
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
    const int64_t P = ptr[n_band];
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
        int64_t *nb_out = nullptr;
        if (nptr) {
            nptr[0] = 0;
            for (int64_t p = 0; p < n_band; ++p) nptr[p + 1] = nptr[p] + (int64_t)nbs[p].size();
            nb_out = (int64_t *)malloc(sizeof(int64_t) * (nptr[n_band] > 0 ? nptr[n_band] : 1));
        }
        if (!nptr || !nb_out) {
            free(ptr); free(pairs); free(nptr); free(nb_out);
            return pde::fail("out of memory", __FILE__, __LINE__);
        }
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
