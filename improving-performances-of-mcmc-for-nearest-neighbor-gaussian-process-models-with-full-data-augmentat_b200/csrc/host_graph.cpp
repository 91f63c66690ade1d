// host_graph.cpp -- host-side set-up utilities of libnngp_b200.so (C ABI: include/nngp_b200.h, nngp_host_*).
//
// These replace init-time code of the reference (GpGp::find_ordered_nn, GpGp::order_maxmin, the crossprod() moral graph
// and Coloring.R), which runs once per model on the host in the reference too.  They are NOT a CPU fallback of the GPU
// hot path.  Semantics (and the reference lines they replace) are documented in the header.
#include <omp.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <queue>
#include <vector>

#include "../../include/nngp_b200.h"
#include "nngp_internal.h"

namespace {

struct Cand {
    double d2;
    int idx;
    bool operator<(const Cand &o) const { return d2 < o.d2 || (d2 == o.d2 && idx < o.idx); }
};

// Uniform grid over (up to) the first three coordinate dimensions of the first `cnt` points.
struct Grid {
    int gd;              // gridded dims (<= 3)
    double lo[3], h[3];  // origin and cell edge per gridded dim
    int nc[3];           // cells per gridded dim
    double hmin;         // smallest edge among non-collapsed dims (inf if all collapsed)
    std::vector<int> start, pts;  // CSR: points of each cell, ascending index

    int cell_coord(double x, int k) const {
        int c = (int)std::floor((x - lo[k]) / h[k]);
        return c < 0 ? 0 : (c >= nc[k] ? nc[k] - 1 : c);
    }
    void build(const double *P /* row-major cnt x d */, int cnt, int d, double target_per_cell) {
        gd = std::min(d, 3);
        double ext[3] = {0, 0, 0}, hi[3];
        for (int k = 0; k < gd; k++) { lo[k] = INFINITY; hi[k] = -INFINITY; }
        for (int i = 0; i < cnt; i++)
            for (int k = 0; k < gd; k++) {
                double x = P[(size_t)i * d + k];
                if (x < lo[k]) lo[k] = x;
                if (x > hi[k]) hi[k] = x;
            }
        for (int k = 0; k < gd; k++) ext[k] = hi[k] - lo[k];
        // choose a common edge so that (non-degenerate) cells are roughly cubic and hold ~target points; dims whose
        // extent is below the edge collapse to a single cell.  Iterate because collapsing changes the effective dim.
        double ncell_target = std::max(1.0, cnt / target_per_cell);
        bool active[3] = {true, true, true};
        double edge = 1.0;
        for (int it = 0; it < 4; it++) {
            double vol = 1.0; int na = 0;
            for (int k = 0; k < gd; k++) if (active[k] && ext[k] > 0) { vol *= ext[k]; na++; } else active[k] = false;
            if (na == 0) { edge = 1.0; break; }
            edge = std::pow(vol / ncell_target, 1.0 / na);
            bool changed = false;
            for (int k = 0; k < gd; k++) if (active[k] && ext[k] < edge) { active[k] = false; changed = true; }
            if (!changed) break;
        }
        hmin = INFINITY;
        size_t total = 1;
        for (int k = 0; k < gd; k++) {
            if (active[k]) {
                nc[k] = std::max(1, (int)std::ceil(ext[k] / edge));
                if (nc[k] > 4096) nc[k] = 4096;
                h[k] = ext[k] / nc[k];
                if (!(h[k] > 0)) { h[k] = 1.0; nc[k] = 1; }
                else hmin = std::min(hmin, h[k]);
            } else { nc[k] = 1; h[k] = (ext[k] > 0 ? ext[k] : 1.0) * 1.0000001; }
            total *= (size_t)nc[k];
        }
        for (int k = gd; k < 3; k++) { nc[k] = 1; lo[k] = 0; h[k] = 1; }
        start.assign(total + 1, 0);
        std::vector<int> cell(cnt);
        for (int i = 0; i < cnt; i++) {
            int c[3] = {0, 0, 0};
            for (int k = 0; k < gd; k++) c[k] = cell_coord(P[(size_t)i * d + k], k);
            cell[i] = (c[2] * nc[1] + c[1]) * nc[0] + c[0];
            start[cell[i] + 1]++;
        }
        for (size_t c = 0; c < total; c++) start[c + 1] += start[c];
        pts.resize(cnt);
        std::vector<int> pos(start.begin(), start.end() - 1);
        for (int i = 0; i < cnt; i++) pts[pos[cell[i]]++] = i;  // ascending index inside each cell
    }
};

inline double dist2(const double *P, int d, int a, int b) {
    double s = 0.0;
    for (int k = 0; k < d; k++) {
        double t = P[(size_t)a * d + k] - P[(size_t)b * d + k];
        s += t * t;
    }
    return s;
}

// PREV: m nearest among points with index < i; otherwise m nearest among all points but i itself.
// Result ascending by (d2, idx) in heap[0..found)
template <bool PREV>
int query_knn(const Grid &g, const double *P, int d, int i, int m, Cand *heap /* size m */) {
    int found = 0;
    int ci[3] = {0, 0, 0};
    for (int k = 0; k < g.gd; k++) ci[k] = g.cell_coord(P[(size_t)i * d + k], k);
    int maxr = 0;
    for (int k = 0; k < g.gd; k++) maxr = std::max(maxr, std::max(ci[k], g.nc[k] - 1 - ci[k]));
    for (int r = 0; r <= maxr; r++) {
        int z0 = std::max(0, ci[2] - r), z1 = std::min(g.nc[2] - 1, ci[2] + r);
        int y0 = std::max(0, ci[1] - r), y1 = std::min(g.nc[1] - 1, ci[1] + r);
        int x0 = std::max(0, ci[0] - r), x1 = std::min(g.nc[0] - 1, ci[0] + r);
        for (int z = z0; z <= z1; z++)
            for (int y = y0; y <= y1; y++) {
                bool edge_zy = (std::abs(z - ci[2]) == r) || (std::abs(y - ci[1]) == r);
                for (int x = x0; x <= x1; x++) {
                    if (!edge_zy && std::abs(x - ci[0]) != r) {  // interior of the ring: jump to the far side
                        if (x < ci[0] + r) { x = ci[0] + r - 1; }
                        continue;
                    }
                    int c = (z * g.nc[1] + y) * g.nc[0] + x;
                    for (int q = g.start[c]; q < g.start[c + 1]; q++) {
                        int p = g.pts[q];
                        if (PREV) { if (p >= i) break; }  // ascending inside the cell
                        else if (p == i) continue;
                        Cand cd{dist2(P, d, i, p), p};
                        if (found < m) {
                            heap[found++] = cd;
                            std::push_heap(heap, heap + found);
                        } else if (cd < heap[0]) {
                            std::pop_heap(heap, heap + found);
                            heap[found - 1] = cd;
                            std::push_heap(heap, heap + found);
                        }
                    }
                }
            }
        if (found == m && std::isfinite(g.hmin)) {
            double bound = r * g.hmin * (1.0 - 1e-12);
            if (heap[0].d2 < bound * bound) break;
        }
    }
    std::sort_heap(heap, heap + found);
    return found;
}

inline int query_prev(const Grid &g, const double *P, int d, int i, int m, Cand *heap) { return query_knn<true>(g, P, d, i, m, heap); }

// GpGp's coordinate jitter: locs + matrix(ee * 1e-4 * rnorm(n * d), n, d), ee = the smallest column standard deviation
// (stats::sd, n - 1).  Column-major in, column-major out; advances R's stream by n * d normals.
std::vector<double> gpgp_jitter(const double *locs_cm, int n, int d, nngp::RStream &rs) {
    double ee = INFINITY;
    for (int k = 0; k < d; k++) {
        const double *c = locs_cm + (size_t)n * k;
        double mean = 0.0;
        for (int i = 0; i < n; i++) mean += c[i];
        mean /= n;
        double ss = 0.0;
        for (int i = 0; i < n; i++) ss += (c[i] - mean) * (c[i] - mean);
        ee = std::min(ee, std::sqrt(ss / (n - 1)));
    }
    std::vector<double> out((size_t)n * d);
    for (size_t t = 0; t < out.size(); t++) out[t] = locs_cm[t] + ee * 1e-4 * rs.norm_rand();
    return out;
}

}  // namespace

namespace nngp {

void find_ordered_nn(const double *locs_cm, int n, int d, int m, int *NNarray) {
    std::vector<double> P((size_t)n * d);
    for (int i = 0; i < n; i++)
        for (int k = 0; k < d; k++) P[(size_t)i * d + k] = locs_cm[(size_t)i + (size_t)n * k];
    auto put = [&](int i, const Cand *c, int found) {
        NNarray[(size_t)i] = i + 1;
        for (int j = 1; j <= m; j++) NNarray[(size_t)i + (size_t)n * j] = (j <= found) ? c[j - 1].idx + 1 : NNGP_NA_INT;
    };
    // brute force head
    int head = std::min(n, std::max(4 * m + 8, 256));
    {
        std::vector<Cand> c(head);
        for (int i = 0; i < head; i++) {
            for (int p = 0; p < i; p++) c[p] = Cand{dist2(P.data(), d, i, p), p};
            int take = std::min(i, m);
            std::partial_sort(c.begin(), c.begin() + take, c.begin() + i);
            put(i, c.data(), take);
        }
    }
    // doubling generations: sites [s, e) query a grid over sites [0, e), filtering index < i
    for (int s = head; s < n;) {
        int e = (int)std::min<int64_t>((int64_t)s * 2, n);
        Grid g;
        g.build(P.data(), e, d, 3.0);
#pragma omp parallel
        {
            std::vector<Cand> heap(m);
#pragma omp for schedule(dynamic, 256)
            for (int i = s; i < e; i++) {
                int found = query_prev(g, P.data(), d, i, m, heap.data());
                put(i, heap.data(), found);
            }
        }
        s = e;
    }
}

// GpGp::find_ordered_nn(locs, m) as the reference calls it (Scripts/mcmc_nngp_initialize.R:93): the search above on
// coordinates jittered with R's stream
void find_ordered_nn_gpgp(const double *locs_cm, int n, int d, int m, RStream &rs, int *NNarray) {
    if (n < 2) { find_ordered_nn(locs_cm, n, d, m, NNarray); return; }
    std::vector<double> x = gpgp_jitter(locs_cm, n, d, rs);
    find_ordered_nn(x.data(), n, d, m, NNarray);
}

// GpGp::order_maxmin(locs, lonlat) (Scripts/mcmc_nngp_initialize.R:29), the reference's default reordering, regenerated
// on R's random stream.  Published algorithm: jitter; k = round(sqrt(n)); a random start permutation sample(n) written
// into the first half of a list of positions; positions j = 2 .. 2n are visited once, and the index found at j is moved to
// the end of the list whenever one of its round(min(k, n / (j - nmoved + 1))) nearest neighbours sits at an earlier
// position (a moved index is visited again later, with fewer neighbours); the ordering is what remains, front to back.
// GpGp takes the neighbours from one FNN::get.knn(locs, k) table (n x k); here they are queried when a position is visited,
// which needs the same neighbours (the first nneigh of the k nearest ARE the nneigh nearest) and no n x sqrt(n) table.
// With set.seed(1) on the vignette's toy locations this reproduces the ordering the vignette prints (Vignette.md:406-419,
// :322-328; tests/test_abi_cpu.py).  The lon/lat branch (coordinates mapped to the unit sphere after the jitter) follows
// the same published source but no reference output pins it.
void order_maxmin_gpgp(const double *locs_cm, int n, int d, bool lonlat, RStream &rs, int *order) {
    if (n < 2) { if (n == 1) order[0] = 1; return; }
    std::vector<double> x = gpgp_jitter(locs_cm, n, d, rs);
    int dd = d;
    if (lonlat) {  // (lon, lat[, time]) -> (x, y, z[, time])
        dd = d + 1;
        std::vector<double> y((size_t)n * dd);
        const double kPi = 3.14159265358979323846;
        for (int i = 0; i < n; i++) {
            double lonrad = x[i] * 2 * kPi / 360, latrad = (x[(size_t)n + i] + 90) * 2 * kPi / 360;
            y[i] = std::sin(latrad) * std::cos(lonrad);
            y[(size_t)n + i] = std::sin(latrad) * std::sin(lonrad);
            y[(size_t)2 * n + i] = std::cos(latrad);
            for (int k = 2; k < d; k++) y[(size_t)(k + 1) * n + i] = x[(size_t)k * n + i];
        }
        x.swap(y);
    }
    // The pass below touches each visited point's neighbours at random; with the points stored cell by cell (index q = position
    // in the grid's cell-major list, old index = cell_major[q]) those reads stay inside a few cache lines.  Distances, hence
    // the neighbour sets, do not depend on the numbering.
    std::vector<int> cell_major;
    std::vector<double> P((size_t)n * dd);
    {
        std::vector<double> P0((size_t)n * dd);
        for (int i = 0; i < n; i++)
            for (int k = 0; k < dd; k++) P0[(size_t)i * dd + k] = x[(size_t)i + (size_t)n * k];
        Grid g0;
        g0.build(P0.data(), n, dd, 3.0);
        cell_major.swap(g0.pts);
        for (int q = 0; q < n; q++)
            for (int k = 0; k < dd; k++) P[(size_t)q * dd + k] = P0[(size_t)cell_major[q] * dd + k];
    }
    std::vector<int> new_of_old(n);
    for (int q = 0; q < n; q++) new_of_old[cell_major[q]] = q;
    const int k = std::min((int)std::nearbyint(std::sqrt((double)n)), n - 1);   // R's round(): halves to even
    Grid g;
    g.build(P.data(), n, dd, 3.0);
    std::vector<int> iip((size_t)3 * n + 2, -1);      // index_in_position (0-based NEW index, -1 = NA); grows to < 3n entries
    std::vector<int> poi(n);                          // position_of_index, 1-based positions
    rs.sample_int(n, n, iip.data());
    for (int t = 0; t < n; t++) { iip[t] = new_of_old[iip[t] - 1]; poi[iip[t]] = t + 1; }
    // The pass is sequential in j, but the neighbour queries are not: the entries at positions j .. min(j + B - 1, curlen) are
    // already final (indices are only ever appended behind curlen), and the number of neighbours a visit needs never grows
    // (j - nmoved counts the visits that kept their index).  So a batch of positions is queried in parallel with the batch's
    // first -- largest -- neighbour count, and the sequential pass then reads the first nneigh(j) of each sorted list.
    int curlen = n, nmoved = 0;
    auto nneigh_at = [&](int j) {
        const double lim = (double)n / (double)(j - nmoved + 1);
        const int v = (int)std::nearbyint(std::min((double)k, lim));
        return v < 1 ? 1 : v;                         // R: NNall[i, 1:0] selects column 1
    };
    std::vector<Cand> cand;
    std::vector<int> found;
    for (int j0 = 2; j0 <= 2 * n;) {
        const int kmax = nneigh_at(j0);
        const int B = std::max(1, std::min(std::min(8192, (1 << 22) / kmax), std::min(curlen, 2 * n) - j0 + 1));
        cand.resize((size_t)B * kmax);
        found.assign(B, 0);
#pragma omp parallel for schedule(dynamic, 16)
        for (int b = 0; b < B; b++) {
            const int v = iip[j0 - 1 + b];
            if (v >= 0) found[b] = query_knn<false>(g, P.data(), dd, v, kmax, cand.data() + (size_t)b * kmax);
        }
        for (int b = 0; b < B; b++) {
            const int j = j0 + b;
            const int v = iip[j - 1];
            if (v < 0) continue;                      // an emptied position: R's min(NA, na.rm = TRUE) is Inf
            const int nneigh = std::min(nneigh_at(j), found[b]);
            const Cand *c = cand.data() + (size_t)b * kmax;
            bool earlier = false;
            for (int q = 0; q < nneigh && !earlier; q++) earlier = poi[c[q].idx] < j;
            if (earlier) {
                nmoved++;
                curlen++;
                poi[v] = curlen;
                iip[curlen - 1] = v;
                iip[j - 1] = -1;
            }
        }
        j0 += B;
    }
    int o = 0;
    for (int t = 0; t < curlen && o < n; t++) if (iip[t] >= 0) order[o++] = cell_major[iip[t]] + 1;
}

// children lists: for site s the rows r (0-based) that contain s, r ascending (includes r == s)
void build_csc(const int *NNarray, int n, int m, std::vector<int64_t> &ptr, std::vector<int> &rows, std::vector<int> *slots_opt) {
    // Counting sort by site over the n x (m+1) table, in parallel: counts and fill positions are claimed with atomics, which
    // leaves the entries of a column in arbitrary order; every column is then sorted by row (a row holds a site at most once, so
    // the result is the unique ascending-row order -- the same arrays as a sequential fill in row order).
    const bool prof = std::getenv("NNGP_PROFILE_COLORING") != nullptr;
    double t0 = omp_get_wtime();
    ptr.assign((size_t)n + 1, 0);
    int64_t *cnt = ptr.data();
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; i++)
        for (int j = 0; j <= m; j++) {
            int v = NNarray[(size_t)i + (size_t)n * j];
            if (v != NNGP_NA_INT) {
#pragma omp atomic
                cnt[v]++;
            }
        }
    if (prof) { fprintf(stderr, "[csc] count %.3f\n", omp_get_wtime() - t0); t0 = omp_get_wtime(); }
    for (int s = 0; s < n; s++) ptr[s + 1] += ptr[s];
    rows.resize(ptr[n]);
    if (slots_opt) slots_opt->resize(ptr[n]);
    int *slots = slots_opt ? slots_opt->data() : nullptr;   // the slot of each entry is optional (the colouring needs rows only)
    std::vector<int64_t> pos(ptr.begin(), ptr.end() - 1);
    int64_t *posp = pos.data();
    if (prof) { fprintf(stderr, "[csc] alloc %.3f\n", omp_get_wtime() - t0); t0 = omp_get_wtime(); }
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; i++)
        for (int j = 0; j <= m; j++) {
            int v = NNarray[(size_t)i + (size_t)n * j];
            if (v != NNGP_NA_INT) {
                int64_t q;
#pragma omp atomic capture
                q = posp[v - 1]++;
                rows[q] = i;
                if (slots) slots[q] = j;
            }
        }
    if (prof) { fprintf(stderr, "[csc] fill %.3f\n", omp_get_wtime() - t0); t0 = omp_get_wtime(); }
#pragma omp parallel
    {
        std::vector<int64_t> key;
#pragma omp for schedule(dynamic, 2048)
        for (int s = 0; s < n; s++) {
            const int64_t e0 = ptr[s], e1 = ptr[s + 1];
            if (e1 - e0 < 2) continue;
            if (!slots) { std::sort(rows.begin() + e0, rows.begin() + e1); continue; }   // cheap: ~ m + 1 ints
            key.resize((size_t)(e1 - e0));
            for (int64_t e = e0; e < e1; e++) key[(size_t)(e - e0)] = ((int64_t)rows[e] << 8) | (int64_t)slots[e];   // m + 1 <= 32 slots
            std::sort(key.begin(), key.end());
            for (int64_t e = e0; e < e1; e++) { rows[e] = (int)(key[(size_t)(e - e0)] >> 8); slots[e] = (int)(key[(size_t)(e - e0)] & 255); }
        }
    }
}

// Exact first-fit colouring in index order, in parallel.  colour(i) = smallest colour not used by a moral neighbour j < i, so
// the sites of a block [b0, b1) of consecutive indices can all scan their neighbourhoods at once: colours of neighbours
// j < b0 are final and go into a 64-bit "forbidden" mask, neighbours inside the block with b0 <= j < i are only recorded.
// A cheap sequential pass over the block then resolves the recorded dependencies in index order.  The result is identical
// to the sequential loop (Coloring.R:2-20) by construction; the expensive part -- ~ (m+1)^2 random reads per site -- is
// spread over the host cores.  Returns 0 when a colour >= 63 shows up (caller falls back to the sequential loop).
static int greedy_coloring_blocked(const int *nn_rm, int n, int M, const std::vector<int64_t> &ptr, const std::vector<int> &rows, int *coloring) {
    const int B = 8192, CAP = 96;
    const char *lim_env = std::getenv("NNGP_COLORING_MASK_COLOURS");   // testing hook: pretend the mask is narrower than 63 colours
    const int limit = lim_env ? std::max(1, std::min(63, std::atoi(lim_env))) : 63;
    std::vector<uint64_t> mask(B);
    std::vector<int> ndep(B), dep((size_t)B * CAP);
    int K = 0;
    double tp = 0, ts = 0;
    long long rescans = 0, rescan_entries = 0;
    for (int b0 = 0; b0 < n;) {
        const int b1 = std::min(n, b0 + std::max(32, std::min(B, b0 / 8)));
        double t0 = omp_get_wtime();
#pragma omp parallel for schedule(dynamic, 64)
        for (int i = b0; i < b1; i++) {
            uint64_t mk = 0;
            int nd = 0;
            for (int64_t k = ptr[i]; k < ptr[i + 1]; k++) {
                const int *row = nn_rm + (size_t)rows[k] * M;
                for (int j = 0; j < M; j++) {
                    const int v = row[j];
                    if (v == NNGP_NA_INT) continue;
                    const int s = v - 1;
                    if (s >= i) continue;
                    if (s < b0) mk |= (uint64_t)1 << coloring[s];
                    else if (nd < CAP) dep[(size_t)(i - b0) * CAP + nd++] = s;
                    else nd = CAP + 1;                 // too many in-block neighbours: rescan in the sequential pass
                }
            }
            mask[i - b0] = mk;
            ndep[i - b0] = nd;
        }
        tp += omp_get_wtime() - t0; t0 = omp_get_wtime();
        for (int i = b0; i < b1; i++) {
            uint64_t mk = mask[i - b0];
            const int nd = ndep[i - b0];
            if (nd <= CAP) {
                for (int k = 0; k < nd; k++) mk |= (uint64_t)1 << coloring[dep[(size_t)(i - b0) * CAP + k]];
            } else {
                rescans++; rescan_entries += ptr[i + 1] - ptr[i];
                for (int64_t k = ptr[i]; k < ptr[i + 1]; k++) {
                    const int *row = nn_rm + (size_t)rows[k] * M;
                    for (int j = 0; j < M; j++) {
                        const int v = row[j];
                        if (v != NNGP_NA_INT && v - 1 < i && v - 1 >= b0) mk |= (uint64_t)1 << coloring[v - 1];
                    }
                }
            }
            mk |= 1;                                    // bit 0 is not a colour
            const int c = __builtin_ctzll(~mk);
            if (c >= limit) return 0;
            coloring[i] = c;
            if (c > K) K = c;
        }
        ts += omp_get_wtime() - t0;
        b0 = b1;
    }
    if (std::getenv("NNGP_PROFILE_COLORING")) fprintf(stderr, "[coloring] parallel scan %.3f s, sequential resolve %.3f s, %lld rescans of %lld rows\n", tp, ts, rescans, rescan_entries);
    return K;
}

int greedy_coloring(const int *NNarray, int n, int m, int *coloring) {
    const bool prof = std::getenv("NNGP_PROFILE_COLORING") != nullptr;
    double t0 = omp_get_wtime();
    std::vector<int64_t> ptr; std::vector<int> rows;
    build_csc(NNarray, n, m, ptr, rows, nullptr);
    if (prof) { fprintf(stderr, "[coloring] build_csc %.3f s\n", omp_get_wtime() - t0); t0 = omp_get_wtime(); }
    std::vector<int> stamp(64, -1);
    int K = 0;
    for (int i = 0; i < n; i++) coloring[i] = 0;
    // row-major copy of the neighbour table: the loop below reads whole rows at random, and in R's column-major layout the
    // m+1 members of a row sit n ints apart (one cache miss each instead of one per row)
    const int M = m + 1;
    std::vector<int> nn_rm((size_t)n * M);
#pragma omp parallel for schedule(static)
    for (int r = 0; r < n; r++)
        for (int j = 0; j < M; j++) nn_rm[(size_t)r * M + j] = NNarray[(size_t)r + (size_t)n * j];
    if (prof) { fprintf(stderr, "[coloring] row-major copy %.3f s\n", omp_get_wtime() - t0); t0 = omp_get_wtime(); }
    if (std::getenv("NNGP_COLORING_SEQUENTIAL") == nullptr) {
        K = greedy_coloring_blocked(nn_rm.data(), n, M, ptr, rows, coloring);
        if (prof) fprintf(stderr, "[coloring] blocked first-fit %.3f s\n", omp_get_wtime() - t0);
        if (K > 0 || n == 0) return K;
        for (int i = 0; i < n; i++) coloring[i] = 0;    // 63 colours or more: the mask is too narrow, use the sequential loop
    }
    for (int i = 0; i < n; i++) {
        // moral neighbours of i = members of every row that contains i (Scripts/mcmc_nngp_initialize.R:103)
        for (int64_t k = ptr[i]; k < ptr[i + 1]; k++) {
            const int *row = nn_rm.data() + (size_t)rows[k] * M;
            for (int j = 0; j <= m; j++) {
                int v = row[j];
                if (v == NNGP_NA_INT) continue;
                int c = coloring[v - 1];
                if (c > 0) {
                    if (c >= (int)stamp.size()) stamp.resize(2 * c + 2, -1);
                    stamp[c] = i;
                }
            }
        }
        int c = 1;
        while (c < (int)stamp.size() && stamp[c] == i) c++;   // Coloring.R:16  match(0, incompatibilities[i,])
        if (c >= (int)stamp.size()) stamp.resize(2 * c + 2, -1);
        coloring[i] = c;
        if (c > K) K = c;
    }
    return K;
}

// exact farthest-point ordering: start from the site closest to the centroid, then repeatedly take the site whose
// distance to the already-ordered set is largest (ties: lower index).
void order_maxmin(const double *locs_cm, int n, int d, int *order) {
    std::vector<double> P((size_t)n * d);
    std::vector<double> cen(d, 0.0);
    for (int i = 0; i < n; i++)
        for (int k = 0; k < d; k++) { P[(size_t)i * d + k] = locs_cm[(size_t)i + (size_t)n * k]; cen[k] += P[(size_t)i * d + k]; }
    for (int k = 0; k < d; k++) cen[k] /= std::max(n, 1);
    if (n == 0) return;
    int first = 0; double best = INFINITY;
    for (int i = 0; i < n; i++) {
        double s = 0; for (int k = 0; k < d; k++) { double t = P[(size_t)i * d + k] - cen[k]; s += t * t; }
        if (s < best) { best = s; first = i; }
    }
    Grid g; g.build(P.data(), n, d, 2.0);
    // everything the selection loop touches is stored in CELL ORDER (position q of g.pts): scanning a cell then reads
    // contiguous coordinates / distances / flags instead of one cache miss per point (10.2 -> ~3 s at n = 1M)
    std::vector<double> Ps((size_t)n * d), dist(n, INFINITY);
    std::vector<char> done(n, 0);
    std::vector<int> pos_of(n);
    for (int q = 0; q < n; q++) {
        const int s = g.pts[q];
        pos_of[s] = q;
        for (int k = 0; k < d; k++) Ps[(size_t)q * d + k] = P[(size_t)s * d + k];
    }
    // Indexed max-heap with one node per unselected site (key = (distance to the selected set, lower original index
    // first)); a relaxation lowers a key in place (sift-down) instead of pushing a fresh entry -- the lazy heap of the first
    // version grew to ~14 M mostly stale entries at n = 1M and its cache misses were most of the run time.
    struct Node { double d2; int idx; int pos; };
    auto before = [](const Node &a, const Node &b) { return a.d2 > b.d2 || (a.d2 == b.d2 && a.idx < b.idx); };   // a leaves the heap first
    std::vector<Node> heap;
    std::vector<int> where(n, -1);
    int hn = 0;
    auto sift_down = [&](int h) {
        Node t = heap[h];
        for (;;) {
            int c = 2 * h + 1;
            if (c >= hn) break;
            if (c + 1 < hn && before(heap[c + 1], heap[c])) c++;
            if (!before(heap[c], t)) break;
            heap[h] = heap[c]; where[heap[h].pos] = h;
            h = c;
        }
        heap[h] = t; where[t.pos] = h;
    };
    bool heap_ready = false;
    auto relax_around = [&](int pq, double r2) {
        // every unselected q with |q-p|^2 < dist[q] has dist[q] <= r2, hence lies within radius sqrt(r2) of p
        double r = std::sqrt(r2);
        int c0[3] = {0, 0, 0}, c1[3] = {0, 0, 0};
        for (int k = 0; k < g.gd; k++) {
            double x = Ps[(size_t)pq * d + k];
            if (std::isfinite(r)) { c0[k] = g.cell_coord(x - r, k); c1[k] = g.cell_coord(x + r, k); }
            else { c0[k] = 0; c1[k] = g.nc[k] - 1; }
        }
        for (int z = c0[2]; z <= c1[2]; z++)
            for (int y = c0[1]; y <= c1[1]; y++) {
                // the cells x = c0[0] .. c1[0] of one grid line are adjacent in memory: one contiguous run of positions
                const int ca = (z * g.nc[1] + y) * g.nc[0] + c0[0], cb = (z * g.nc[1] + y) * g.nc[0] + c1[0];
                for (int q = g.start[ca]; q < g.start[cb + 1]; q++) {
                    if (done[q]) continue;
                    double dd = dist2(Ps.data(), d, q, pq);
                    if (dd < dist[q]) {
                        dist[q] = dd;
                        if (heap_ready) { heap[where[q]].d2 = dd; sift_down(where[q]); }
                    }
                }
            }
    };
    int cnt = 0;
    done[pos_of[first]] = 1; order[cnt++] = first + 1;
    relax_around(pos_of[first], INFINITY);             // every other site now has a finite distance
    heap.reserve(n);
    for (int q = 0; q < n; q++)
        if (!done[q]) { heap.push_back(Node{dist[q], g.pts[q], q}); }
    hn = (int)heap.size();
    for (int h = 0; h < hn; h++) where[heap[h].pos] = h;
    for (int h = hn / 2 - 1; h >= 0; h--) sift_down(h);
    heap_ready = true;
    while (hn > 0) {
        const Node t = heap[0];
        hn--;
        if (hn > 0) { heap[0] = heap[hn]; sift_down(0); }
        where[t.pos] = -1;
        done[t.pos] = 1; order[cnt++] = t.idx + 1;
        relax_around(t.pos, t.d2);
    }
}

// depth of every row in the solve DAG: 0 for rows without parents, else 1 + max over parents
int solve_levels(const int *NNarray, int n, int m, std::vector<int> &level) {
    level.assign(n, 0);
    int depth = 0;
    for (int i = 0; i < n; i++) {
        int l = 0;
        for (int j = 1; j <= m; j++) {
            int v = NNarray[(size_t)i + (size_t)n * j];
            if (v != NNGP_NA_INT) l = std::max(l, level[v - 1] + 1);
        }
        level[i] = l;
        depth = std::max(depth, l + 1);
    }
    return depth;
}

}  // namespace nngp

extern "C" {

void nngp_host_find_ordered_nn(const double *locs, const int *n, const int *d, const int *m, int *NNarray, int *status) {
    if (!locs || !n || !d || !m || !NNarray || *n < 0 || *d < 1 || *m < 0) { nngp::set_error("nngp_host_find_ordered_nn: bad argument"); if (status) *status = NNGP_ERR_ARG; return; }
    nngp::find_ordered_nn(locs, *n, *d, *m, NNarray);
    *status = NNGP_OK;
}

void nngp_host_greedy_coloring(const int *NNarray, const int *n, const int *m, int *coloring, int *n_colors, int *status) {
    if (!NNarray || !n || !m || !coloring || *n < 0 || *m < 0) { nngp::set_error("nngp_host_greedy_coloring: bad argument"); if (status) *status = NNGP_ERR_ARG; return; }
    int K = nngp::greedy_coloring(NNarray, *n, *m, coloring);
    if (n_colors) *n_colors = K;
    *status = NNGP_OK;
}

// first-fit colouring straight from an adjacency structure in compressed-column form (0-based row ids, diagonal allowed), i.e.
// the slots @p / @i of the dgCMatrix the reference hands to naive_greedy_coloring (Scripts/Coloring.R:2-20, called at
// Scripts/mcmc_nngp_initialize.R:110): same colours 1..K as the R loop, O(n + nnz) instead of its (n+1) x maxdeg double scratch
void nngp_host_greedy_coloring_adj(const int *adj_p, const int *adj_i, const int *n, int *coloring, int *n_colors, int *status) {
    if (!adj_p || !adj_i || !n || !coloring || *n < 0) { nngp::set_error("nngp_host_greedy_coloring_adj: bad argument"); if (status) *status = NNGP_ERR_ARG; return; }
    const int nn = *n;
    int K = 0;
    std::vector<int> mark;   // mark[c] = last node whose neighbourhood contained colour c
    for (int i = 0; i < nn; i++) {
        if (adj_p[i + 1] < adj_p[i]) { nngp::set_error("nngp_host_greedy_coloring_adj: adj_p is not non-decreasing"); if (status) *status = NNGP_ERR_ARG; return; }
        coloring[i] = 0;
    }
    // Coloring.R marks the neighbours of i (self included) as incompatible with cols[i] AFTER colouring i, so node i sees the
    // colours of its lower-indexed neighbours only: first colour not used by an already coloured neighbour
    for (int i = 0; i < nn; i++) {
        for (int e = adj_p[i]; e < adj_p[i + 1]; e++) {
            const int j = adj_i[e];
            if (j < 0 || j >= nn) { nngp::set_error("nngp_host_greedy_coloring_adj: adj_i[%d] = %d out of range", e, j); if (status) *status = NNGP_ERR_ARG; return; }
            const int c = coloring[j];
            if (j != i && c > 0) { if ((int)mark.size() <= c) mark.resize(c + 1, -1); mark[c] = i; }
        }
        int c = 1;
        while (c < (int)mark.size() && mark[c] == i) c++;
        coloring[i] = c;
        K = std::max(K, c);
    }
    if (n_colors) *n_colors = K;
    if (status) *status = NNGP_OK;
}

// number of OpenMP threads of the host set-up utilities (a launcher such as torchrun exports OMP_NUM_THREADS=1, which the OpenMP
// runtime reads once, when it is first loaded -- possibly long before this library is); n <= 0 = all processors
void nngp_host_set_num_threads(const int *n, int *status) {
    if (!n) { nngp::set_error("nngp_host_set_num_threads: null argument"); if (status) *status = NNGP_ERR_ARG; return; }
    omp_set_num_threads(*n > 0 ? *n : omp_get_num_procs());
    if (status) *status = NNGP_OK;
}

void nngp_host_order_maxmin_gpgp(const double *locs, const int *n, const int *d, const int *lonlat, int *rstate, int *order, int *status) {
    if (!locs || !n || !d || !lonlat || !rstate || !order || *n < 0 || *d < 1 || (*lonlat && *d < 2)) { nngp::set_error("nngp_host_order_maxmin_gpgp: bad argument"); if (status) *status = NNGP_ERR_ARG; return; }
    nngp::RStream rs;
    rs.load(rstate);
    nngp::order_maxmin_gpgp(locs, *n, *d, *lonlat != 0, rs, order);
    rs.store(rstate);
    *status = NNGP_OK;
}

void nngp_host_find_ordered_nn_gpgp(const double *locs, const int *n, const int *d, const int *m, int *rstate, int *NNarray, int *status) {
    if (!locs || !n || !d || !m || !rstate || !NNarray || *n < 0 || *d < 1 || *m < 0) { nngp::set_error("nngp_host_find_ordered_nn_gpgp: bad argument"); if (status) *status = NNGP_ERR_ARG; return; }
    nngp::RStream rs;
    rs.load(rstate);
    nngp::find_ordered_nn_gpgp(locs, *n, *d, *m, rs, NNarray);
    rs.store(rstate);
    *status = NNGP_OK;
}

void nngp_host_order_maxmin(const double *locs, const int *n, const int *d, int *order, int *status) {
    if (!locs || !n || !d || !order || *n < 0 || *d < 1) { nngp::set_error("nngp_host_order_maxmin: bad argument"); if (status) *status = NNGP_ERR_ARG; return; }
    nngp::order_maxmin(locs, *n, *d, order);
    *status = NNGP_OK;
}

}  // extern "C"
