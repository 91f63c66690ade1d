/*
 * oracle/gpgp_order.c -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).
 *
 * CPU restatement of the two GpGp routines that fix the reference's ordering and neighbour table, INCLUDING their use of
 * R's random stream (Scripts/mcmc_nngp_initialize.R:29 GpGp::order_maxmin, :93 GpGp::find_ordered_nn).  GpGp is a CRAN
 * dependency that is absent from /root/reference (version unpinned by the reference; the vignette was rendered
 * 2021-06-14, i.e. GpGp 0.3.x / 0.4.0); what follows restates its published algorithm:
 *
 *   order_maxmin(locs):  jitter the coordinates by 1e-4 * (smallest column sd) * rnorm(n*d);  all-points kNN with
 *     k = round(sqrt(n)) (FNN::get.knn: self excluded, increasing distance);  start from sample(n);  walk positions
 *     j = 2 .. 2n of a list twice as long, and move the index at position j to the end of the list whenever one of its
 *     round(min(k, n/(j - nmoved + 1))) nearest neighbours sits at an earlier position;  the ordering is what is left.
 *   find_ordered_nn(locs, m):  the same jitter (fresh rnorm(n*d)), then for every i the m nearest PREVIOUS points of
 *     the jittered coordinates, self first, by increasing distance.
 *
 * PINNED against the reference's own printed values (tests/test_vignette_pin.py): with set.seed(1) the restatement
 * reproduces all 100 printed entries of the vignette's ordering (Vignette.md:406-419), locs_match[1:100] (:322-328, which
 * needs the whole permutation), the NNarray head (:221-227), and -- through the random-stream position it leaves behind
 * -- the initial states (:472-523) and every Gelman-Rubin-Brooks value the vignette prints after that.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "nngp_oracle.h"

/* locs + matrix(ee * 1e-4 * rnorm(n*d), n, d), ee = min over columns of sd(column) (n-1 denominator, as stats::sd) */
static double *jittered(const double *locs, int n, int d)
{
    double ee = INFINITY;
    for (int k = 0; k < d; k++) {
        const double *c = locs + (size_t)n * k;
        double mean = 0.0;
        for (int i = 0; i < n; i++) mean += c[i];
        mean /= n;
        double ss = 0.0;
        for (int i = 0; i < n; i++) ss += (c[i] - mean) * (c[i] - mean);
        double sd = sqrt(ss / (n - 1));
        if (sd < ee) ee = sd;
    }
    double *out = (double *)malloc(sizeof(double) * (size_t)n * d);
    for (size_t t = 0; t < (size_t)n * d; t++) out[t] = locs[t] + ee * 1e-4 * r_norm_rand(); /* column-major fill */
    return out;
}

/* R >= 4.0.0 round(x): to nearest, halves to even */
static int r_round(double x) { return (int)nearbyint(x); }

/* GpGp::order_maxmin(locs, lonlat = FALSE).  locs n x d column-major; order_out: n 1-based indices. */
void oracle_order_maxmin_gpgp(const double *locs, int n, int d, int *order_out)
{
    if (n < 2) { if (n == 1) order_out[0] = 1; return; }
    double *x = jittered(locs, n, d);
    int k = r_round(sqrt((double)n));
    if (k > n - 1) k = n - 1;
    /* FNN::get.knn(x, k)$nn.index: k nearest other points by increasing distance (exact; brute force here) */
    int *NNall = (int *)malloc(sizeof(int) * (size_t)n * (k > 0 ? k : 1));
    double *bd = (double *)malloc(sizeof(double) * (size_t)(k + 1));
    int *bi = (int *)malloc(sizeof(int) * (size_t)(k + 1));
    for (int i = 0; i < n; i++) {
        int cnt = 0;
        for (int p = 0; p < n; p++) {
            if (p == i) continue;
            double s = 0.0;
            for (int c = 0; c < d; c++) {
                double t = x[i + (size_t)n * c] - x[p + (size_t)n * c];
                s += t * t;
            }
            if (cnt == k && s >= bd[cnt - 1]) continue;
            int q = cnt < k ? cnt : k - 1;
            while (q > 0 && bd[q - 1] > s) { bd[q] = bd[q - 1]; bi[q] = bi[q - 1]; q--; }
            bd[q] = s;
            bi[q] = p;
            if (cnt < k) cnt++;
        }
        for (int q = 0; q < k; q++) NNall[(size_t)i * k + q] = bi[q];
    }
    free(bd);
    free(bi);
    free(x);
    /* index_in_position <- c(sample(n), rep(NA, n)) (R grows it on assignment past the end); 0-based, -1 = NA */
    size_t cap = (size_t)4 * n + 4;
    int *iip = (int *)malloc(sizeof(int) * cap);
    for (size_t t = 0; t < cap; t++) iip[t] = -1;
    int *perm = (int *)malloc(sizeof(int) * (size_t)n);
    r_sample_perm(n, perm);
    int *poi = (int *)malloc(sizeof(int) * (size_t)n); /* position_of_index, 1-based positions */
    for (int t = 0; t < n; t++) { iip[t] = perm[t] - 1; poi[perm[t] - 1] = t + 1; }
    free(perm);
    int curlen = n, nmoved = 0;
    for (int j = 2; j <= 2 * n; j++) {
        int v = iip[j - 1];
        if (v < 0) continue; /* NA row: min(NA, na.rm = TRUE) = Inf, never < j */
        double lim = (double)n / (double)(j - nmoved + 1);
        int nneigh = r_round(lim < (double)k ? lim : (double)k);
        if (nneigh < 1) nneigh = 1; /* R: NNall[i, 1:0] is column 1 */
        int first = 2 * n + 1 + 2 * n;
        for (int q = 0; q < nneigh && q < k; q++) {
            int pp = poi[NNall[(size_t)v * k + q]];
            if (pp < first) first = pp;
        }
        if (first < j) {
            nmoved++;
            curlen++;
            poi[v] = curlen;
            iip[curlen - 1] = v;
            iip[j - 1] = -1;
        }
    }
    int o = 0;
    for (int t = 0; t < curlen && o < n; t++) if (iip[t] >= 0) order_out[o++] = iip[t] + 1;
    free(iip);
    free(poi);
    free(NNall);
}

/* GpGp::find_ordered_nn(locs, m): jitter (consumes rnorm(n*d)), then the exact ordered search of oracle_find_ordered_nn */
void oracle_find_ordered_nn_gpgp(const double *locs, int n, int d, int m, int *NNarray)
{
    double *x = jittered(locs, n, d);
    oracle_find_ordered_nn(x, n, d, m, NNarray);
    free(x);
}
