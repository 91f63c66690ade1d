// host_shard.cpp -- host-side set-up of a spatially sharded field (BASELINE.json config 4): the block owner of
// every site and, per rank, the local site set, local NNarray and per-(colour, peer) halo lists.  O(n (m+1)) per rank: two passes
// over NNarray and a handful of n-vectors (the numpy prototype this replaces looped over colours x peers x n).
//
// Rank g OWNS the sites of one spatial block.  To sweep them it needs, locally,
//   * the factor rows of every row that contains an owned site (its own rows + "ghost rows": children that live elsewhere),
//   * the field value of every site appearing in those rows ("ghost sites" = moral-graph neighbours across the cut), kept
//     current by a per-colour halo exchange of boundary values.
// Local numbering preserves the global order, so the local factor stays lower triangular and the local NNarray obeys the same
// invariants as a global one (Scripts/mcmc_nngp_initialize.R:93-101).
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <mutex>
#include <numeric>
#include <vector>

#include "../../include/nngp_b200.h"
#include "nngp_internal.h"

namespace nngp {

// recursive coordinate bisection into n_parts blocks of (almost) equal counts, cutting the longer of the first two axes;
// ties are broken by site index, so every rank computes the same owners from the same coordinates
static void bisect(const double *locs_cm, int n, int d, int *idx, int count, int lo, int parts, int *owner) {
    if (parts == 1) {
        for (int k = 0; k < count; k++) owner[idx[k]] = lo;
        return;
    }
    const int left = parts / 2;
    int axis = 0;
    double best = -1.0;
    for (int k = 0; k < std::min(d, 2); k++) {
        double mn = 1e300, mx = -1e300;
        const double *x = locs_cm + (size_t)n * k;
        for (int t = 0; t < count; t++) { const double v = x[idx[t]]; mn = std::min(mn, v); mx = std::max(mx, v); }
        if (mx - mn > best) { best = mx - mn; axis = k; }
    }
    const double *x = locs_cm + (size_t)n * axis;
    const int cut = (int)(((long long)count * left) / parts);
    std::nth_element(idx, idx + cut, idx + count, [&](int a, int b) { return x[a] < x[b] || (x[a] == x[b] && a < b); });
    bisect(locs_cm, n, d, idx, cut, lo, left, owner);
    bisect(locs_cm, n, d, idx + cut, count - cut, lo + left, parts - left, owner);
}

struct ShardPlan {
    int n_local = 0, n_obs_local = 0, n_owned = 0, K = 0, W = 0, d = 0, M = 0;
    std::vector<double> locs;
    std::vector<int> nn, coloring, owned, global_id, global_zpos, global_level, obs_index, locs_match, send_site, send_ptr, recv_site, recv_ptr;
};

static std::mutex g_plan_mu;
static std::vector<ShardPlan *> g_plans;

static ShardPlan *build_plan(const double *locs, const int *NNarray, const int *coloring, int n, int d, int m, int n_obs,
                             const int *locs_match, const int *owner, int rank, int W) {
    const int M = m + 1;
    ShardPlan *P = new ShardPlan();
    P->W = W; P->d = d; P->M = M;
    int K = 0;
    for (int i = 0; i < n; i++) K = std::max(K, coloring[i]);
    P->K = K;
    // mask[s] bit h: site s appears in a row that contains a site owned by h  (=> s is local to h)
    std::vector<uint8_t> mask(n, 0), rowmask(n, 0);
    for (int i = 0; i < n; i++) {
        uint8_t rm = 0;
        for (int j = 0; j < M; j++) {
            const int v = NNarray[(size_t)i + (size_t)n * j];
            if (v != NNGP_NA_INT) rm |= (uint8_t)(1u << owner[v - 1]);
        }
        rowmask[i] = rm;
        for (int j = 0; j < M; j++) {
            const int v = NNarray[(size_t)i + (size_t)n * j];
            if (v != NNGP_NA_INT) mask[v - 1] |= rm;
        }
    }
    const uint8_t me = (uint8_t)(1u << rank);
    std::vector<int> g2l(n, -1);
    int nl = 0;
    for (int s = 0; s < n; s++) if (mask[s] & me) g2l[s] = nl++;
    P->n_local = nl;
    P->global_id.resize(nl); P->coloring.resize(nl); P->owned.resize(nl); P->global_zpos.resize(nl); P->global_level.resize(nl);
    // depth of every row in the WHOLE field's solve DAG: all ranks order their owned rows by it (sharded triangular solve)
    std::vector<int> level;
    solve_levels(NNarray, n, m, level);
    P->locs.resize((size_t)nl * d);
    P->nn.assign((size_t)nl * M, NNGP_NA_INT);
    // position of every site in the reference's rnorm() hand-out order: colour 1..K, ascending index inside a colour
    std::vector<int> cnext(K + 2, 0);
    for (int s = 0; s < n; s++) cnext[coloring[s] + 1]++;
    for (int c = 1; c <= K + 1; c++) cnext[c] += cnext[c - 1];   // cnext[c] = first position of colour c (1-based colours)
    for (int s = 0; s < n; s++) {
        const int zp = cnext[coloring[s]]++;
        const int l = g2l[s];
        if (l < 0) continue;
        P->global_zpos[l] = zp;
        P->global_level[l] = level[s];
        P->global_id[l] = s;
        P->coloring[l] = coloring[s];
        P->owned[l] = owner[s] == rank ? 1 : 0;
        P->n_owned += P->owned[l];
        for (int k = 0; k < d; k++) P->locs[(size_t)l + (size_t)nl * k] = locs[(size_t)s + (size_t)n * k];
        P->nn[l] = l + 1;
        if (rowmask[s] & me) {   // a row this rank needs: all its parents are local by construction
            for (int j = 1; j < M; j++) {
                const int v = NNarray[(size_t)s + (size_t)n * j];
                if (v != NNGP_NA_INT) P->nn[(size_t)l + (size_t)nl * j] = g2l[v - 1] + 1;
            }
        }   // else: a ghost site whose own row is not needed keeps a trivial self-only row
    }
    // observations of owned sites only (every observation is counted by exactly one rank)
    for (int o = 0; o < n_obs; o++) {
        const int s = locs_match[o] - 1;
        if (owner[s] == rank) { P->obs_index.push_back(o); P->locs_match.push_back(g2l[s] + 1); }
    }
    P->n_obs_local = (int)P->obs_index.size();
    // halo lists per (colour, peer), ascending global id on both sides: what this rank sends to h for colour c is what h expects
    const size_t nb = (size_t)K * W;
    P->send_ptr.assign(nb + 1, 0);
    P->recv_ptr.assign(nb + 1, 0);
    for (int l = 0; l < nl; l++) {
        const int s = P->global_id[l], c = coloring[s] - 1;
        if (owner[s] == rank) {
            const uint8_t others = mask[s] & (uint8_t)~me;
            for (int h = 0; h < W; h++) if (others & (1u << h)) P->send_ptr[(size_t)c * W + h + 1]++;
        } else {
            P->recv_ptr[(size_t)c * W + owner[s] + 1]++;
        }
    }
    for (size_t k = 0; k < nb; k++) { P->send_ptr[k + 1] += P->send_ptr[k]; P->recv_ptr[k + 1] += P->recv_ptr[k]; }
    P->send_site.resize(P->send_ptr[nb]);
    P->recv_site.resize(P->recv_ptr[nb]);
    std::vector<int> spos(P->send_ptr.begin(), P->send_ptr.end() - 1), rpos(P->recv_ptr.begin(), P->recv_ptr.end() - 1);
    for (int l = 0; l < nl; l++) {
        const int s = P->global_id[l], c = coloring[s] - 1;
        if (owner[s] == rank) {
            const uint8_t others = mask[s] & (uint8_t)~me;
            for (int h = 0; h < W; h++) if (others & (1u << h)) P->send_site[spos[(size_t)c * W + h]++] = l + 1;
        } else {
            P->recv_site[rpos[(size_t)c * W + owner[s]]++] = l + 1;
        }
    }
    return P;
}

}  // namespace nngp

extern "C" {

void nngp_host_spatial_blocks(const double *locs, const int *n, const int *d, const int *n_parts, int *owner, int *status) {
    if (!locs || !n || !d || !n_parts || !owner || *n < 0 || *d < 1 || *n_parts < 1) { nngp::set_error("nngp_host_spatial_blocks: bad argument"); if (status) *status = NNGP_ERR_ARG; return; }
    std::vector<int> idx(*n);
    std::iota(idx.begin(), idx.end(), 0);
    nngp::bisect(locs, *n, *d, idx.data(), *n, 0, *n_parts, owner);
    if (status) *status = NNGP_OK;
}

void nngp_host_shard_plan_build(const double *locs, const int *NNarray, const int *coloring, const int *n, const int *d, const int *m,
                                const int *n_obs, const int *locs_match, const int *owner, const int *rank, const int *world,
                                int *plan_id, int *sizes6, int *status) {
    if (!locs || !NNarray || !coloring || !n || !d || !m || !n_obs || (!locs_match && *n_obs > 0) || !owner || !rank || !world || !plan_id || !sizes6 ||
        *n < 1 || *d < 1 || *m < 1 || *world < 1 || *world > 8 || *rank < 0 || *rank >= *world) {
        nngp::set_error("nngp_host_shard_plan_build: bad argument (1 <= world <= 8)");
        if (status) *status = NNGP_ERR_ARG;
        return;
    }
    for (int i = 0; i < *n; i++)
        if (owner[i] < 0 || owner[i] >= *world || coloring[i] < 1) {
            nngp::set_error("nngp_host_shard_plan_build: owner[%d] = %d / coloring = %d out of range", i, owner[i], coloring[i]);
            if (status) *status = NNGP_ERR_ARG;
            return;
        }
    for (int o = 0; o < *n_obs; o++)
        if (locs_match[o] < 1 || locs_match[o] > *n) {
            nngp::set_error("nngp_host_shard_plan_build: locs_match[%d] = %d out of range", o + 1, locs_match[o]);
            if (status) *status = NNGP_ERR_ARG;
            return;
        }
    nngp::ShardPlan *P = nullptr;
    try {
        P = nngp::build_plan(locs, NNarray, coloring, *n, *d, *m, *n_obs, locs_match, owner, *rank, *world);
    } catch (const std::bad_alloc &) {
        nngp::set_error("nngp_host_shard_plan_build: host allocation failed");
        if (status) *status = NNGP_ERR_ALLOC;
        return;
    }
    sizes6[0] = P->n_local; sizes6[1] = P->n_obs_local; sizes6[2] = (int)P->send_site.size(); sizes6[3] = (int)P->recv_site.size();
    sizes6[4] = P->K; sizes6[5] = P->n_owned;
    std::lock_guard<std::mutex> lk(nngp::g_plan_mu);
    int id = -1;
    for (size_t k = 0; k < nngp::g_plans.size(); k++) if (!nngp::g_plans[k]) { id = (int)k; break; }
    if (id < 0) { nngp::g_plans.push_back(nullptr); id = (int)nngp::g_plans.size() - 1; }
    nngp::g_plans[id] = P;
    *plan_id = id;
    if (status) *status = NNGP_OK;
}

void nngp_host_shard_plan_get(const int *plan_id, double *locs, int *NNarray, int *coloring, int *owned, int *global_id, int *global_zpos,
                              int *global_level, int *obs_index, int *locs_match, int *send_site, int *send_ptr, int *recv_site, int *recv_ptr, int *status) {
    nngp::ShardPlan *P = nullptr;
    {
        std::lock_guard<std::mutex> lk(nngp::g_plan_mu);
        if (plan_id && *plan_id >= 0 && *plan_id < (int)nngp::g_plans.size()) { P = nngp::g_plans[*plan_id]; nngp::g_plans[*plan_id] = nullptr; }
    }
    if (!P) { nngp::set_error("nngp_host_shard_plan_get: unknown plan id"); if (status) *status = NNGP_ERR_ARG; return; }
    auto put = [](auto *dst, const auto &v) { if (dst && !v.empty()) std::memcpy(dst, v.data(), v.size() * sizeof(v[0])); };
    put(locs, P->locs); put(NNarray, P->nn); put(coloring, P->coloring); put(owned, P->owned); put(global_id, P->global_id);
    put(global_zpos, P->global_zpos); put(global_level, P->global_level); put(obs_index, P->obs_index); put(locs_match, P->locs_match); put(send_site, P->send_site);
    put(send_ptr, P->send_ptr); put(recv_site, P->recv_site); put(recv_ptr, P->recv_ptr);
    delete P;
    if (status) *status = NNGP_OK;
}

}  // extern "C"
