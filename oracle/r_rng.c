/*
 * oracle/r_rng.c -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).
 *
 * CPU restatement of the random-number stream the reference consumes.  The reference is plain R and calls
 * set.seed() / runif() / rnorm() / sample() (Scripts/mcmc_nngp_update_Gaussian.R:36,75,113,133,165,189,223,273,285,288;
 * Scripts/mcmc_nngp_initialize.R:17,30,154,189,193,208; Vignette.rmd:26-43).  R's default generators are
 * Mersenne-Twister (uniforms), "Inversion" (normals) and "Rejection" (sample, R >= 3.6).  R itself is a third-party
 * dependency that is absent from /root/reference and from this image; the algorithms below are the published ones
 * (Matsumoto & Nishimura MT19937; Wichura AS241 PPND16 for qnorm) with R's seeding/scrambling convention, and they
 * are pinned against the printed values of the rendered vignette (Vignette.md:136-142 runif, :193-198 rnorm) by
 * tests/test_oracle_golden.py.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>
#include "nngp_oracle.h"

#define MT_N 624
#define MT_M 397

struct r_rng {
    uint32_t mt[MT_N];
    int mti;
};

static struct r_rng g_rng;

/* set.seed(seed): 50 rounds of LCG scrambling, then 625 LCG outputs fill i_seed; i_seed[0] (= mti) is forced to N. */
void r_set_seed(uint32_t seed)
{
    for (int j = 0; j < 50; j++) seed = 69069u * seed + 1u;
    uint32_t dummy0;
    seed = 69069u * seed + 1u;
    dummy0 = seed; /* i_seed[0], overwritten by FixupSeeds -> mti = N */
    (void)dummy0;
    for (int j = 0; j < MT_N; j++) {
        seed = 69069u * seed + 1u;
        g_rng.mt[j] = seed;
    }
    g_rng.mti = MT_N;
}

static uint32_t mt_genrand(void)
{
    static const uint32_t mag01[2] = {0x0u, 0x9908b0dfu};
    uint32_t y;
    uint32_t *mt = g_rng.mt;
    if (g_rng.mti >= MT_N) {
        int kk;
        for (kk = 0; kk < MT_N - MT_M; kk++) {
            y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
            mt[kk] = mt[kk + MT_M] ^ (y >> 1) ^ mag01[y & 0x1];
        }
        for (; kk < MT_N - 1; kk++) {
            y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
            mt[kk] = mt[kk + (MT_M - MT_N)] ^ (y >> 1) ^ mag01[y & 0x1];
        }
        y = (mt[MT_N - 1] & 0x80000000u) | (mt[0] & 0x7fffffffu);
        mt[MT_N - 1] = mt[MT_M - 1] ^ (y >> 1) ^ mag01[y & 0x1];
        g_rng.mti = 0;
    }
    y = mt[g_rng.mti++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}

/* unif_rand(): [0,1) real with 32-bit resolution, then nudged into the open interval. */
double r_unif_rand(void)
{
    const double i2_32m1 = 2.328306437080797e-10; /* 1/(2^32 - 1) */
    double v = mt_genrand() * 2.3283064365386963e-10; /* 2^-32 */
    if (v <= 0.0) return 0.5 * i2_32m1;
    if ((1.0 - v) <= 0.0) return 1.0 - 0.5 * i2_32m1;
    return v;
}

/* qnorm(p, 0, 1, lower, !log): Wichura (1988) algorithm AS241, PPND16. */
double r_qnorm(double p)
{
    double q, r, val;
    if (isnan(p)) return p;
    if (p <= 0.0) return -INFINITY;
    if (p >= 1.0) return INFINITY;
    q = p - 0.5;
    if (fabs(q) <= 0.425) {
        r = .180625 - q * q;
        val = q *
              (((((((r * 2509.0809287301226727 + 33430.575583588128105) * r + 67265.770927008700853) * r +
                   45921.953931549871457) * r + 13731.693765509461125) * r + 1971.5909503065514427) * r +
                133.14166789178437745) * r + 3.387132872796366608) /
              (((((((r * 5226.495278852545925 + 28729.085735721942674) * r + 39307.89580009271061) * r +
                   21213.794301586595867) * r + 5394.1960214247511077) * r + 687.1870074920579083) * r +
                42.313330701600911252) * r + 1.);
        return val;
    }
    r = (q < 0) ? p : 1.0 - p;
    r = sqrt(-log(r));
    if (r <= 5.) {
        r += -1.6;
        val = (((((((r * 7.7454501427834140764e-4 + .0227238449892691845833) * r + .24178072517745061177) * r +
                  1.27045825245236838258) * r + 3.64784832476320460504) * r + 5.7694972214606914055) * r +
               4.6303378461565452959) * r + 1.42343711074968357734) /
              (((((((r * 1.05075007164441684324e-9 + 5.475938084995344946e-4) * r + .0151986665636164571966) * r +
                   .14810397642748007459) * r + .68976733498510000455) * r + 1.6763848301838038494) * r +
                2.05319162663775882187) * r + 1.);
    } else {
        r += -5.;
        val = (((((((r * 2.01033439929228813265e-7 + 2.71155556874348757815e-5) * r + .0012426609473880784386) * r +
                  .026532189526576123093) * r + .29656057182850489123) * r + 1.7848265399172913358) * r +
               5.4637849111641143699) * r + 6.6579046435011037772) /
              (((((((r * 2.04426310338993978564e-15 + 1.4215117583164458887e-7) * r + 1.8463183175100546818e-5) * r +
                   7.868691311456132591e-4) * r + .0148753612908506148525) * r + .13692988092273580531) * r +
                .59983220655588793769) * r + 1.);
    }
    if (q < 0.0) val = -val;
    return val;
}

/* norm_rand() with normal.kind = "Inversion": two uniforms give a 59-bit-resolution probability. */
double r_norm_rand(void)
{
    const double BIG = 134217728.0; /* 2^27 */
    double u = r_unif_rand();
    u = (double)(int)(BIG * u) + r_unif_rand();
    return r_qnorm(u / BIG);
}

void r_runif(int n, double *out)
{
    for (int i = 0; i < n; i++) {
        double u;
        do { u = r_unif_rand(); } while (u <= 0.0 || u >= 1.0);
        out[i] = u;
    }
}

void r_rnorm(int n, double mean, double sd, double *out)
{
    for (int i = 0; i < n; i++) out[i] = mean + sd * r_norm_rand();
}

/* R >= 3.6 sample.kind = "Rejection": draw ceil(log2(dn)) random bits, reject values >= dn. */
static double r_rbits(int bits)
{
    int64_t v = 0;
    for (int n = 0; n <= bits; n += 16) {
        int v1 = (int)floor(r_unif_rand() * 65536);
        v = 65536 * v + v1;
    }
    if (bits < 64) v &= ((((int64_t)1) << bits) - 1);
    return (double)v;
}

double r_unif_index(double dn)
{
    if (dn <= 0) return 0.0;
    int bits = (int)ceil(log2(dn));
    double dv;
    do { dv = r_rbits(bits); } while (dn <= dv);
    return dv;
}

/* sample(seq(n)) -- a full permutation without replacement, 1-based (Scripts/mcmc_nngp_initialize.R:30). */
void r_sample_perm(int n, int *out)
{
    int *x = out; /* in place: out holds the shrinking pool at the front, results are emitted to a temp */
    int *pool = (int *)__builtin_malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    for (int i = 0; i < n; i++) pool[i] = i;
    int nn = n;
    for (int i = 0; i < n; i++) {
        int j = (int)r_unif_index((double)nn);
        x[i] = pool[j] + 1;
        pool[j] = pool[--nn];
    }
    __builtin_free(pool);
}

/* snapshot / restore so that a caller can interleave several logical streams in tests */
void r_rng_get_state(uint32_t *state625)
{
    memcpy(state625, g_rng.mt, sizeof(uint32_t) * MT_N);
    state625[MT_N] = (uint32_t)g_rng.mti;
}

void r_rng_set_state(const uint32_t *state625)
{
    memcpy(g_rng.mt, state625, sizeof(uint32_t) * MT_N);
    g_rng.mti = (int)state625[MT_N];
}

/* sample.int(n, size) without replacement, 1-based (R >= 3.6 "Rejection"): the first `size` draws of the shuffle that
 * r_sample_perm completes.  Scripts/mcmc_nngp_initialize.R:154 draws sample(x, 1) = x[sample.int(length(x), 1)]. */
void r_sample_int(int n, int size, int *out)
{
    int *pool = (int *)__builtin_malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    for (int i = 0; i < n; i++) pool[i] = i;
    int nn = n;
    for (int i = 0; i < size && i < n; i++) {
        int j = (int)r_unif_index((double)nn);
        out[i] = pool[j] + 1;
        pool[j] = pool[--nn];
    }
    __builtin_free(pool);
}

/* rbeta(1, aa, bb) (Scripts/mcmc_nngp_initialize.R:193-194 with shape1 = shape2 = 10): R's nmath/rbeta.c is Cheng (1978)
 * -- algorithm BB, the branch for min(aa, bb) > 1 -- with the constants as R prints them (1.3862944, 2.609438, ...).
 * Pinned by the initial log_scale / log_noise_variance the vignette prints (Vignette.md:486-499). */
double r_rbeta(double aa, double bb)
{
    const double expmax = 1024 * 0.693147180559945309417232121458; /* DBL_MAX_EXP * M_LN2 */
    if (isnan(aa) || isnan(bb) || aa < 0. || bb < 0.) return NAN;
    if (isinf(aa) && isinf(bb)) return 0.5;
    if (aa == 0. && bb == 0.) return (r_unif_rand() < 0.5) ? 0. : 1.;
    if (isinf(aa) || bb == 0.) return 1.0;
    if (isinf(bb) || aa == 0.) return 0.0;
    double a = fmin(aa, bb), b = fmax(aa, bb), alpha = a + b;
    double r, s, t, u1, u2, v, w, z;
#define V_W_FROM_U1_BET(AA)                                   \
    v = beta * log(u1 / (1.0 - u1));                          \
    if (v <= expmax) { w = AA * exp(v); if (isinf(w)) w = 1.7976931348623157e308; } \
    else w = 1.7976931348623157e308
    if (a <= 1.0) return NAN; /* algorithm BC: never reached by the reference (shape1 = shape2 = 10), not restated */
    { /* algorithm BB */
        double beta = sqrt((alpha - 2.0) / (2.0 * a * b - alpha));
        double gamma = a + 1.0 / beta;
        do {
            u1 = r_unif_rand();
            u2 = r_unif_rand();
            V_W_FROM_U1_BET(a);
            z = u1 * u1 * u2;
            r = gamma * v - 1.3862944;
            s = a + r - w;
            if (s + 2.609438 >= 5.0 * z) break;
            t = log(z);
            if (s > t) break;
        } while (r + alpha * log(alpha / (b + w)) < t);
        return (aa != a) ? b / (b + w) : w / (b + w);
    }
#undef V_W_FROM_U1_BET
}
