// r_stream.cpp -- R-compatible random-number stream for the host side of nngp_chain_run.
//
// The reference seeds each chain with set.seed(iter_start + i) (Scripts/mcmc_nngp_update_Gaussian.R:36) and then draws
// rnorm()/runif() in a fixed order.  To be a drop-in whose scalar draws (proposal innovations, accept/reject uniforms,
// beta_0, noise-variance steps) follow the same stream, the shim carries R's default generators: Mersenne-Twister
// MT19937 with R's LCG seed scrambling, and normal.kind = "Inversion" (two uniforms -> 59-bit probability -> Wichura's
// AS241 quantile).  Published algorithms; nothing here comes from /root/reference (which contains no RNG code).
#include <algorithm>
#include <cmath>
#include "../../include/nngp_b200.h"
#include "nngp_internal.h"

namespace nngp {

void RStream::set_seed(uint32_t seed) {
    for (int j = 0; j < 50; j++) seed = 69069u * seed + 1u;
    seed = 69069u * seed + 1u;  // would be i_seed[0]; R overwrites it with mti = 624
    for (int j = 0; j < 624; j++) { seed = 69069u * seed + 1u; mt_[j] = seed; }
    mti_ = 624;
}

uint32_t RStream::genrand() {
    const int N = 624, M = 397;
    if (mti_ >= N) {
        if (mti_ == N + 1) set_seed(4357u);
        int kk = 0;
        for (; kk < N - M; kk++) {
            uint32_t y = (mt_[kk] & 0x80000000u) | (mt_[kk + 1] & 0x7fffffffu);
            mt_[kk] = mt_[kk + M] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        for (; kk < N - 1; kk++) {
            uint32_t y = (mt_[kk] & 0x80000000u) | (mt_[kk + 1] & 0x7fffffffu);
            mt_[kk] = mt_[kk + (M - N)] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        uint32_t y = (mt_[N - 1] & 0x80000000u) | (mt_[0] & 0x7fffffffu);
        mt_[N - 1] = mt_[M - 1] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        mti_ = 0;
    }
    uint32_t y = mt_[mti_++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}

double RStream::unif_rand() {
    const double i2_32m1 = 2.328306437080797e-10;
    double v = genrand() * 2.3283064365386963e-10;
    if (v <= 0.0) return 0.5 * i2_32m1;
    if (1.0 - v <= 0.0) return 1.0 - 0.5 * i2_32m1;
    return v;
}

static double qnorm_as241(double p) {
    if (p <= 0.0) return -INFINITY;
    if (p >= 1.0) return INFINITY;
    double q = p - 0.5, r, val;
    if (std::fabs(q) <= 0.425) {
        r = .180625 - q * q;
        return q *
               (((((((r * 2509.0809287301226727 + 33430.575583588128105) * r + 67265.770927008700853) * r +
                    45921.953931549871457) * r + 13731.693765509461125) * r + 1971.5909503065514427) * r +
                 133.14166789178437745) * r + 3.387132872796366608) /
               (((((((r * 5226.495278852545925 + 28729.085735721942674) * r + 39307.89580009271061) * r +
                    21213.794301586595867) * r + 5394.1960214247511077) * r + 687.1870074920579083) * r +
                 42.313330701600911252) * r + 1.);
    }
    r = (q < 0) ? p : 1.0 - p;
    r = std::sqrt(-std::log(r));
    if (r <= 5.) {
        r -= 1.6;
        val = (((((((r * 7.7454501427834140764e-4 + .0227238449892691845833) * r + .24178072517745061177) * r +
                  1.27045825245236838258) * r + 3.64784832476320460504) * r + 5.7694972214606914055) * r +
               4.6303378461565452959) * r + 1.42343711074968357734) /
              (((((((r * 1.05075007164441684324e-9 + 5.475938084995344946e-4) * r + .0151986665636164571966) * r +
                   .14810397642748007459) * r + .68976733498510000455) * r + 1.6763848301838038494) * r +
                2.05319162663775882187) * r + 1.);
    } else {
        r -= 5.;
        val = (((((((r * 2.01033439929228813265e-7 + 2.71155556874348757815e-5) * r + .0012426609473880784386) * r +
                  .026532189526576123093) * r + .29656057182850489123) * r + 1.7848265399172913358) * r +
               5.4637849111641143699) * r + 6.6579046435011037772) /
              (((((((r * 2.04426310338993978564e-15 + 1.4215117583164458887e-7) * r + 1.8463183175100546818e-5) * r +
                   7.868691311456132591e-4) * r + .0148753612908506148525) * r + .13692988092273580531) * r +
                .59983220655588793769) * r + 1.);
    }
    return q < 0.0 ? -val : val;
}

double RStream::norm_rand() {
    const double BIG = 134217728.0;
    double u = unif_rand();
    u = (double)(int)(BIG * u) + unif_rand();
    return qnorm_as241(u / BIG);
}

void RStream::rnorm(double *out, int64_t n) {
    for (int64_t i = 0; i < n; i++) out[i] = norm_rand();
}

// R >= 3.6, sample.kind = "Rejection": ceil(log2(dn)) random bits, 16 per uniform, redrawn while >= dn
double RStream::unif_index(double dn) {
    if (dn <= 0) return 0.0;
    const int bits = (int)std::ceil(std::log2(dn));
    double dv;
    do {
        int64_t v = 0;
        for (int b = 0; b <= bits; b += 16) v = 65536 * v + (int)std::floor(unif_rand() * 65536);
        if (bits < 64) v &= (((int64_t)1) << bits) - 1;
        dv = (double)v;
    } while (dn <= dv);
    return dv;
}

// sample.int(n, size) without replacement: draw j uniform on the remaining pool, emit pool[j], move the last entry into slot j
void RStream::sample_int(int n, int size, int *out) {
    std::vector<int> pool((size_t)std::max(n, 1));
    for (int i = 0; i < n; i++) pool[i] = i;
    int left = n;
    for (int i = 0; i < size && i < n; i++) {
        int j = (int)unif_index((double)left);
        out[i] = pool[j] + 1;
        pool[j] = pool[--left];
    }
}

// rbeta(1, aa, bb) for min(aa, bb) > 1: Cheng (1978) algorithm BB with the constants of R's nmath/rbeta.c
// (Scripts/mcmc_nngp_initialize.R:193-194 draws rbeta(1, 10, 10))
double RStream::rbeta(double aa, double bb) {
    const double a = std::min(aa, bb), b = std::max(aa, bb), alpha = a + b;
    if (!(a > 1.0) || !std::isfinite(b)) return NAN;
    const double beta = std::sqrt((alpha - 2.0) / (2.0 * a * b - alpha)), gamma = a + 1.0 / beta;
    const double expmax = 1024 * 0.693147180559945309417232121458, dmax = 1.7976931348623157e308;
    double r, s, t, v, w, z;
    do {
        const double u1 = unif_rand(), u2 = unif_rand();
        v = beta * std::log(u1 / (1.0 - u1));
        if (v <= expmax) { w = a * std::exp(v); if (!std::isfinite(w)) w = dmax; } else w = dmax;
        z = u1 * u1 * u2;
        r = gamma * v - 1.3862944;
        s = a + r - w;
        if (s + 2.609438 >= 5.0 * z) break;
        t = std::log(z);
        if (s > t) break;
    } while (r + alpha * std::log(alpha / (b + w)) < t);
    return (aa != a) ? b / (b + w) : w / (b + w);
}

void RStream::load(const int *st) {
    mti_ = st[0];
    for (int j = 0; j < 624; j++) mt_[j] = (uint32_t)st[1 + j];
}

void RStream::store(int *st) const {
    st[0] = mti_;
    for (int j = 0; j < 624; j++) st[1 + j] = (int)mt_[j];
}

}  // namespace nngp

// ---- C ABI: R's random stream for hosts that are not R (include/nngp_b200.h, "R-compatible random stream") ----
extern "C" {

#define RS_CHECK(cond, name)                                                                              \
    if (!(cond)) { nngp::set_error(name ": bad argument"); if (status) *status = NNGP_ERR_ARG; return; }

void nngp_rng_set_seed(const int *seed, int *rstate, int *status) {
    RS_CHECK(seed && rstate, "nngp_rng_set_seed");
    nngp::RStream rs;
    rs.set_seed((uint32_t)*seed);
    rs.store(rstate);
    if (status) *status = NNGP_OK;
}

void nngp_rng_runif(int *rstate, const int *n, double *out, int *status) {
    RS_CHECK(rstate && n && out && *n >= 0, "nngp_rng_runif");
    nngp::RStream rs;
    rs.load(rstate);
    for (int i = 0; i < *n; i++) out[i] = rs.unif_rand();
    rs.store(rstate);
    if (status) *status = NNGP_OK;
}

void nngp_rng_rnorm(int *rstate, const int *n, double *out, int *status) {
    RS_CHECK(rstate && n && out && *n >= 0, "nngp_rng_rnorm");
    nngp::RStream rs;
    rs.load(rstate);
    rs.rnorm(out, *n);
    rs.store(rstate);
    if (status) *status = NNGP_OK;
}

void nngp_rng_sample_int(int *rstate, const int *n, const int *size, int *out, int *status) {
    RS_CHECK(rstate && n && size && out && *n >= 0 && *size >= 0 && *size <= *n, "nngp_rng_sample_int");
    nngp::RStream rs;
    rs.load(rstate);
    rs.sample_int(*n, *size, out);
    rs.store(rstate);
    if (status) *status = NNGP_OK;
}

void nngp_rng_rbeta(int *rstate, const int *n, const double *shape1, const double *shape2, double *out, int *status) {
    RS_CHECK(rstate && n && shape1 && shape2 && out && *n >= 0, "nngp_rng_rbeta");
    if (!(std::min(*shape1, *shape2) > 1.0) || !std::isfinite(*shape1) || !std::isfinite(*shape2)) {
        nngp::set_error("nngp_rng_rbeta: only shape1, shape2 > 1 (the reference draws rbeta(1, 10, 10))");
        if (status) *status = NNGP_ERR_ARG;
        return;
    }
    nngp::RStream rs;
    rs.load(rstate);
    for (int i = 0; i < *n; i++) out[i] = rs.rbeta(*shape1, *shape2);
    rs.store(rstate);
    if (status) *status = NNGP_OK;
}

}  // extern "C"
