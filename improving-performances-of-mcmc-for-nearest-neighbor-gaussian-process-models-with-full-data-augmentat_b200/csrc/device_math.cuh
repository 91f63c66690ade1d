// device_math.cuh -- small device-side building blocks: block reductions, Philox4x32-10, modified Bessel K of real order.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace nngp {

// ---------------------------------------------------------------------------------------------------------------
// warp-shuffle block reduction of NV doubles; result valid in thread 0.  blockDim.x must be a multiple of 32, <= 1024.
// ---------------------------------------------------------------------------------------------------------------
template <int NV>
__device__ __forceinline__ void block_reduce_sum(double (&v)[NV]) {
    __shared__ double sh[NV][32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int k = 0; k < NV; k++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_down_sync(0xffffffffu, v[k], o);
        if (lane == 0) sh[k][wid] = v[k];
    }
    __syncthreads();
    if (wid == 0) {
#pragma unroll
        for (int k = 0; k < NV; k++) {
            double t = (lane < nw) ? sh[k][lane] : 0.0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
            v[k] = t;
        }
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011).  Counter = (site id, sweep counter, 0, 0), key = 64-bit seed.
// ---------------------------------------------------------------------------------------------------------------
struct u32x4 { uint32_t x, y, z, w; };

__device__ __forceinline__ u32x4 philox4x32_10(u32x4 c, uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; r++) {
        uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
        uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
        u32x4 n;
        n.x = hi1 ^ c.y ^ k0;
        n.y = lo1;
        n.z = hi0 ^ c.w ^ k1;
        n.w = lo0;
        c = n;
        k0 += W0;
        k1 += W1;
    }
    return c;
}

// one standard normal per (site, sweep): Box-Muller on two 53-bit uniforms in (0,1)
__device__ __forceinline__ double philox_normal(uint32_t site, uint32_t sweep_lo, uint32_t sweep_hi, uint32_t k0, uint32_t k1) {
    u32x4 c{site, sweep_lo, sweep_hi, 0x4e4e4750u};
    u32x4 r = philox4x32_10(c, k0, k1);
    uint64_t a = ((uint64_t)r.x << 32) | r.y, b = ((uint64_t)r.z << 32) | r.w;
    double u1 = ((double)(a >> 11) + 0.5) * 1.1102230246251565e-16;  // 2^-53
    double u2 = ((double)(b >> 11) + 0.5) * 1.1102230246251565e-16;
    double s, c2;
    sincospi(2.0 * u2, &s, &c2);
    return sqrt(-2.0 * log(u1)) * c2;
}

// ---------------------------------------------------------------------------------------------------------------
// K_nu(x), real order 0 <= nu < ~2.5, x > 0: Temme's series for x < 2, Steed's continued fraction (CF2) beyond, upward
// recurrence from |mu| <= 1/2.  GpGp's Matern kernels call boost::math::cyl_bessel_k; CUDA has no real-order
// cyl_bessel_k.  gam1/gam2 come from the Taylor series of 1/Gamma (Abramowitz & Stegun 6.1.34), which avoids the
// cancellation of (1/Gamma(1-mu) - 1/Gamma(1+mu)) / (2 mu) near mu = 0.  Checked against std::cyl_bessel_k (oracle) and
// scipy.special.kv to ~1e-13 relative.
// ---------------------------------------------------------------------------------------------------------------
// scaled = true returns K_nu(x) * exp(x) (no underflow for large x)
__device__ __noinline__ double bessel_k_real(double nu, double x, bool scaled = false) {
    const double EPS = 1.0e-16, PI = 3.14159265358979323846;
    const int MAXIT = 10000;
    const int nl = (int)(nu + 0.5);
    const double xmu = nu - nl, xmu2 = xmu * xmu, xi = 1.0 / x, xi2 = 2.0 * xi;
    double rkmu, rk1;
    if (x < 2.0) {
        // odd-index coefficients c2,c4,... (gam1) and even-index c1,c3,... (gam2) of 1/Gamma(z) = sum c_k z^k
        const double ce[13] = {0.5772156649015329, -0.0420026350340952, -0.0421977345555443, 0.0072189432466630,
                               -0.0002152416741149, -0.0000201348547807, 0.0000011330272320, 0.0000000061160950,
                               -0.0000000011812746, 0.0000000000077823, 0.0000000000005100, -0.0000000000000054,
                               0.0000000000000001};
        const double co[13] = {1.0, -0.6558780715202538, 0.1665386113822915, -0.0096219715278770, -0.0011651675918591,
                               0.0001280502823882, -0.0000012504934821, -0.0000002056338417, 0.0000000050020075,
                               0.0000000001043427, -0.0000000000036968, -0.0000000000000206, 0.0000000000000014};
        double g1 = 0.0, g2 = 0.0;
#pragma unroll
        for (int k = 12; k >= 0; k--) { g1 = g1 * xmu2 + ce[k]; g2 = g2 * xmu2 + co[k]; }
        const double gam1 = -g1, gam2 = g2;
        const double gampl = gam2 - xmu * gam1, gammi = gam2 + xmu * gam1;  // 1/Gamma(1+mu), 1/Gamma(1-mu)
        const double x2 = 0.5 * x, pimu = PI * xmu;
        const double fact = (fabs(pimu) < EPS) ? 1.0 : pimu / sin(pimu);
        double d = -log(x2), e = xmu * d;
        const double fact2 = (fabs(e) < EPS) ? 1.0 : sinh(e) / e;
        double ff = fact * (gam1 * cosh(e) + gam2 * fact2 * d);
        double sum = ff;
        e = exp(e);
        double p = 0.5 * e / gampl, q = 0.5 / (e * gammi), c = 1.0;
        d = x2 * x2;
        double sum1 = p;
        for (int i = 1; i <= MAXIT; i++) {
            ff = (i * ff + p + q) / (i * (double)i - xmu2);
            c *= (d / i);
            p /= (i - xmu);
            q /= (i + xmu);
            const double del = c * ff;
            sum += del;
            sum1 += c * (p - i * ff);
            if (fabs(del) < fabs(sum) * EPS) break;
        }
        rkmu = sum;
        rk1 = sum1 * xi2;
    } else {
        double b = 2.0 * (1.0 + x), d = 1.0 / b, h = d, delh = d, q1 = 0.0, q2 = 1.0;
        const double a1 = 0.25 - xmu2;
        double q = a1, c = a1, a = -a1, s = 1.0 + q * delh;
        for (int i = 2; i <= MAXIT; i++) {
            a -= 2 * (i - 1);
            c = -a * c / i;
            const double qnew = (q1 - b * q2) / a;
            q1 = q2;
            q2 = qnew;
            q += c * qnew;
            b += 2.0;
            d = 1.0 / (b + a * d);
            delh = (b * d - 1.0) * delh;
            h += delh;
            const double dels = q * delh;
            s += dels;
            if (fabs(dels / s) < EPS) break;
        }
        h = a1 * h;
        rkmu = sqrt(PI / (2.0 * x)) * (scaled ? 1.0 : exp(-x)) / s;
        rk1 = rkmu * (xmu + x + 0.5 - h) * xi;
    }
    if (scaled && x < 2.0) {
        const double ex = exp(x);
        rkmu *= ex;
        rk1 *= ex;
    }
    for (int i = 1; i <= nl; i++) {
        const double t = (xmu + i) * xi2 * rk1 + rkmu;
        rkmu = rk1;
        rk1 = t;
    }
    return rkmu;
}

}  // namespace nngp
