// kernels.cuh -- sm_100a kernels of the NNGP hot path.
//
// Two internal numberings (both invisible at the ABI):
//   STORAGE order   -- rows / sites of nn, linv, tl, field, r and every n-vector.  NNGP_LAYOUT_MORTON: pure Z-curve order of
//                      the coordinates, so that the 8-byte gathers of a row's parents (log-lik, SpMV, factor build) and of a
//                      site's children (sweep: r) fall into few 32-byte sectors; NNGP_LAYOUT_COLOR[_MORTON]: colour-major.
//   PROCESSING order -- sites of the Gibbs sweep: colour-major (a colour class is a contiguous range), storage order inside
//                      a colour.  The CSC (transpose) arrays, the tiles and the per-site constants pd / nobs / S / gid /
//                      zpos are laid out in processing order; psite[p] is the storage id of processing site p.
// Device layout:
//   nn    int32  [M][ld]   neighbour table, slot-major ("column-major" as R stores NNarray): nn[j*ld + q]; -1 = NA
//   linv  double [M][ld]   compressed factor, same shape: linv[j*ld + q]     (slot 0 = 1/sqrt(F_q), slot j = -B_qj/sqrt(F_q))
//   tl    double [n][DT]   coordinates after the covariance family's transformation (range-scaled / unit sphere)
//   CSC (transpose) of the factor pattern, used by the Gibbs sweep:
//     colptr int32 [n+1], crow int32 [nnz] (row of each entry), csrc int32 [nnz] (position of the entry in linv),
//     valT double [nnz] (values gathered from the CURRENT factor), pd double [n] (precision_diag)
// Slot-major storage makes every thread-per-row access of nn / linv a perfectly coalesced 128 B / 256 B warp transaction;
// the only irregular accesses are the gathers of 8-byte operands (coordinates, field, residual), which stay in L2.
#pragma once
#include "device_math.cuh"

namespace nngp {

struct CovConst {
    double variance, nugget, smooth, normcon;
    double range[4];  // per-dimension divisors (after the family's mapping)
    int covfun;       // NNGP_* id
    int d;            // raw coordinate dimension
    int dt;           // transformed dimension
    const double *mtab;   // Matern families: piecewise-polynomial table of the kernel (nullptr = evaluate K_nu directly)
    // second copy of the factor in the order in which the triangular solve walks the rows (DAG level order): row q also goes to
    // position lpos[q] of linv_lvl ([m+1][nsl]); the solve then reads its rows with coalesced loads instead of one 32-byte
    // sector per 8-byte value (ncu round 1: 950 MB of DRAM traffic for 228 MB of data).  nullptr = no copy.
    double *linv_lvl;
    const int *lpos;
    int nsl;
};

// ---------------------------------------------------------------------------------------------------------------
// Matern kernel table.  Inside one factor build the smoothness nu is fixed, so the kernel is a fixed smooth function of the scaled
// distance; evaluating K_nu directly (Temme series / continued fraction, ~2.7 k FP64 instructions per call, 55 calls per row at
// m = 10) made the Matern factor build 26x slower than the exponential one.  The table is indexed by the SQUARED scaled distance
// s = x^2 -- what the factor kernels have at hand, so no square root is taken per pair -- and holds, for every binary octave
// [2^e, 2^(e+1)) of s, e in [MT_EMIN, MT_EMAX), MT_S sub-intervals (uniform in the mantissa) with a degree-7 Newton interpolant on
// Chebyshev nodes of   g(s) = normcon x^nu K_nu(x)            for s < 4  (the kernel value itself: no exp per pair either), and of
//                      g(s) = normcon x^nu K_nu(x) e^x        for s >= 4 (multiplied by exp(-x) at look-up; rare among neighbours).
// The segment and the local coordinate come straight from the exponent / mantissa bits of s.  Interpolation error < 1e-13
// relative (checked through the factor parity tests against std::cyl_bessel_k).  Outside the table range the exact routine is used.
// ---------------------------------------------------------------------------------------------------------------
#define MT_EMIN (-80)
#define MT_EMAX 20
#define MT_ESCALED 2   /* octaves of s from here on hold the e^x-scaled kernel */
#define MT_S 16
#define MT_SEGS ((MT_EMAX - MT_EMIN) * MT_S)

__device__ __forceinline__ double mt_node(int k) {   // Chebyshev nodes on [0, 1]
    const double nodes[8] = {0.99039264020161522, 0.91573480615127262, 0.77778511650980109, 0.59754516100806417,
                             0.40245483899193590, 0.22221488349019886, 0.08426519384872738, 0.00960735979838478};
    return nodes[k];
}

__global__ void __launch_bounds__(128) matern_table_kernel(double *__restrict__ tab, double nu, double normcon) {
    __shared__ double f[128];
    const int seg = blockIdx.x * 16 + (threadIdx.x >> 3), k = threadIdx.x & 7;
    if (seg < MT_SEGS) {
        const int e = MT_EMIN + seg / MT_S, msub = seg % MT_S;
        const double x = sqrt(ldexp(1.0 + ((double)msub + mt_node(k)) / MT_S, e));
        f[threadIdx.x] = normcon * pow(x, nu) * bessel_k_real(nu, x, e >= MT_ESCALED);
    }
    __syncthreads();
    if (seg < MT_SEGS && k == 0) {
        double c[8];
        for (int j = 0; j < 8; j++) c[j] = f[threadIdx.x + j];
        for (int j = 1; j < 8; j++)
            for (int i = 7; i >= j; i--) c[i] = (c[i] - c[i - 1]) / (mt_node(i) - mt_node(i - j));
        for (int j = 0; j < 8; j++) tab[(size_t)seg * 8 + j] = c[j];
    }
}

// ---------------------------------------------------------------------------------------------------------------
// exp(-sqrt(d2)) for the register-resident factor kernel.  SASS of the first version (cuobjdump, m = 10): 3.5 k
// instructions per row of which only 1.4 k were FP64 math -- 870 UMOV + 730 (I)MAD/MOV re-materialised the 64-bit
// polynomial coefficients of libdevice's exp() for each of the 55 pairs (an FP64 immediate cannot be encoded in DFMA), and
// every sqrt() carried a slow-path CALL.  Here the coefficients are constant-bank operands of the DFMAs (no moves), the
// argument is known to be <= 0 and finite (no special cases), a 32-entry table of 2^(j/32) shortens the polynomial to
// degree 6, and the distance comes from the MUFU seed + one Goldschmidt step + one Newton correction (no slow path).
// ---------------------------------------------------------------------------------------------------------------
// exp(x) = 2^n * 2^(j/32) * P(r),  x = (32 n + j) ln2/32 + r,  |r| <= ln2/64: degree-6 Taylor (remainder 4e-18),
// table of 2^(j/32) correctly rounded (L1-resident 256 bytes), the power of two is added into the table entry's exponent
__device__ const double g_exp2_tab32[32] = {
    0x1.0000000000000p+0, 0x1.059b0d3158574p+0, 0x1.0b5586cf9890fp+0, 0x1.11301d0125b51p+0, 0x1.172b83c7d517bp+0, 0x1.1d4873168b9aap+0,
    0x1.2387a6e756238p+0, 0x1.29e9df51fdee1p+0, 0x1.306fe0a31b715p+0, 0x1.371a7373aa9cbp+0, 0x1.3dea64c123422p+0, 0x1.44e086061892dp+0,
    0x1.4bfdad5362a27p+0, 0x1.5342b569d4f82p+0, 0x1.5ab07dd485429p+0, 0x1.6247eb03a5585p+0, 0x1.6a09e667f3bcdp+0, 0x1.71f75e8ec5f74p+0,
    0x1.7a11473eb0187p+0, 0x1.82589994cce13p+0, 0x1.8ace5422aa0dbp+0, 0x1.93737b0cdc5e5p+0, 0x1.9c49182a3f090p+0, 0x1.a5503b23e255dp+0,
    0x1.ae89f995ad3adp+0, 0x1.b7f76f2fb5e47p+0, 0x1.c199bdd85529cp+0, 0x1.cb720dcef9069p+0, 0x1.d5818dcfba487p+0, 0x1.dfc97337b9b5fp+0,
    0x1.ea4afa2a490dap+0, 0x1.f50765b6e4540p+0};
__constant__ double c_exp_red[4] = {0x1.71547652b82fep+5 /* 32 / ln2 */, 6755399441055744.0 /* 1.5 * 2^52 */,
                                    0x1.62e42fee00000p-6 /* ln2 / 32, high part (21 trailing zero bits) */, 0x1.a39ef35793c76p-38 /* low part */};
__constant__ double c_exp_taylor[4] = {1.0 / 6, 1.0 / 24, 1.0 / 120, 1.0 / 720};

// sqrt(d2), d2 >= 0: MUFU.RSQ64H seed (>= 20 bits), one coupled Goldschmidt step (g -> sqrt, h -> 1/(2 sqrt)), one Newton
// correction; no special-case slow path (libdevice's sqrt()/rsqrt() carry a CALL each).  <= 1 ulp; a coincident pair
// (d2 = 0) comes out as 1e-150, whose covariance is exactly the variance.
__device__ __forceinline__ double fast_sqrt_nonneg(double d2) {
    const double a = fmax(d2, 1e-300);
    double r;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
    double g = a * r, h = 0.5 * r;
    double e = fma(-g, h, 0.5);
    g = fma(g, e, g);
    h = fma(h, e, h);
    e = fma(-g, g, a);
    return fma(e, h, g);
}

// exp(x) for -745 < x <= 0 (returns 0 below -708: the result would be subnormal)
__device__ __forceinline__ double fast_exp_nonpos(double x) {
    const double kd = fma(x, c_exp_red[0], c_exp_red[1]);
    const int ki = __double2loint(kd);
    const double nd = kd - c_exp_red[1];
    double r = fma(nd, -c_exp_red[2], x);
    r = fma(nd, -c_exp_red[3], r);
    double p = fma(c_exp_taylor[3], r, c_exp_taylor[2]);
    p = fma(p, r, c_exp_taylor[1]);
    p = fma(p, r, c_exp_taylor[0]);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    const double t = __ldg(g_exp2_tab32 + (ki & 31));
    const double scale = __hiloint2double(__double2hiint(t) + ((ki >> 5) << 20), __double2loint(t));
    return x < -708.0 ? 0.0 : p * scale;
}

// s = squared scaled distance (> 0)
__device__ __forceinline__ double matern_from_table(const double *__restrict__ tab, double s, double &out_ok) {
    const long long bits = __double_as_longlong(s);
    const int e = (int)((bits >> 52) & 0x7ff) - 1023;
    if (e < MT_EMIN || e >= MT_EMAX) { out_ok = 0.0; return 0.0; }
    const int seg = (e - MT_EMIN) * MT_S + (int)((bits >> 48) & 0xF);
    // position inside the segment, exactly: y = 1.mantissa, yb = y with the low 48 mantissa bits cleared, u = 16 (y - yb)
    const long long mant = (bits & 0x000FFFFFFFFFFFFFll) | 0x3FF0000000000000ll;
    const double u = (__longlong_as_double(mant) - __longlong_as_double(mant & (long long)0xFFFF000000000000ull)) * 16.0;
    const double2 *c2 = reinterpret_cast<const double2 *>(tab + (size_t)seg * 8);   // 64-byte aligned: four 128-bit loads
    const double2 c01 = c2[0], c23 = c2[1], c45 = c2[2], c67 = c2[3];
    double p = c67.y;
    p = p * (u - mt_node(6)) + c67.x;
    p = p * (u - mt_node(5)) + c45.y;
    p = p * (u - mt_node(4)) + c45.x;
    p = p * (u - mt_node(3)) + c23.y;
    p = p * (u - mt_node(2)) + c23.x;
    p = p * (u - mt_node(1)) + c01.y;
    p = p * (u - mt_node(0)) + c01.x;
    out_ok = 1.0;
    return e >= MT_ESCALED ? p * fast_exp_nonpos(-fast_sqrt_nonneg(s)) : p;
}

// parameters of the sweep that change between launches live in device memory so that the captured graph is static
struct SweepParams {
    double beta0, e_ls, e_ln;  // exp(-log_scale), exp(-log_noise_variance)
    unsigned long long sweep_counter;
    unsigned long long z_offset;  // offset of the current sweep's normals inside zbuf (supplied mode)
    unsigned int key0, key1;
    int rng_mode;
};

// ---------------------------------------------------------------------------------------------------------------
// permutation / utility kernels
// ---------------------------------------------------------------------------------------------------------------
__global__ void gather_f64_kernel(double *__restrict__ dst, const double *__restrict__ src, const int *__restrict__ map, int n) {
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) dst[t] = src[map[t]];
}

// dst (n x M, column-major with leading dim n, reference order) <- linv (slot-major, ld, internal order)
__global__ void linv_to_host_order_kernel(double *__restrict__ dst, const double *__restrict__ linv, const int *__restrict__ g2i,
                                          int n, int ld, int M) {
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
        const int q = g2i[t];
        for (int j = 0; j < M; j++) dst[(size_t)j * n + t] = linv[(size_t)j * ld + q];
    }
}

// ---------------------------------------------------------------------------------------------------------------
// device-side record store (records$field of update_Gaussian.R:56,311): stored samples are written
// straight into the final R layout (n_rec x n, column-major, reference site order) in HBM and leave the device in one copy
// at the end of the cycle, instead of one download + strided host scatter per stored sample.
// ---------------------------------------------------------------------------------------------------------------
__global__ void record_field_kernel(double *__restrict__ rec, const double *__restrict__ field, const int *__restrict__ g2i,
                                    int row, int n_rec, int n) {
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x)
        rec[(size_t)row + (size_t)n_rec * t] = field[g2i[t]];
}

// Posterior summary of the stored samples of every site, on the device (get_summary, Scripts/mcmc_nngp_estimate.R:1-6 applied
// to records$field as in :88-94): mean, quantiles 2.5 / 50 / 97.5 % (R's default type 7), sd (n-1), after subtracting a
// per-sample offset (beta_0 of the same iteration).  One thread per site; its samples are contiguous in the record store.
__global__ void __launch_bounds__(128) records_summary_kernel(const double *__restrict__ rec, int n_rec_total, int row0, int n_rows,
                                                              const double *__restrict__ offsets, int n, double *__restrict__ out,
                                                              double *__restrict__ scratch) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    double *v = scratch + (size_t)t * n_rows;   // per-site workspace
    double mean = 0.0;
    for (int k = 0; k < n_rows; k++) {
        const double x = rec[(size_t)(row0 + k) + (size_t)n_rec_total * t] - (offsets ? offsets[k] : 0.0);
        mean += x;
        int j = k;                               // insertion sort (n_rows is a few hundred at most)
        while (j > 0 && v[j - 1] > x) { v[j] = v[j - 1]; j--; }
        v[j] = x;
    }
    mean /= n_rows;
    double ss = 0.0;
    for (int k = 0; k < n_rows; k++) { const double d = v[k] - mean; ss += d * d; }
    auto quantile = [&](double p) {              // type 7: h = (n-1) p, linear interpolation between order statistics
        const double h = (n_rows - 1) * p;
        const int lo = (int)floor(h);
        const int hi = min(lo + 1, n_rows - 1);
        return v[lo] + (h - lo) * (v[hi] - v[lo]);
    };
    out[t] = mean;
    out[(size_t)n + t] = quantile(0.025);
    out[(size_t)2 * n + t] = quantile(0.5);
    out[(size_t)3 * n + t] = quantile(0.975);
    out[(size_t)4 * n + t] = n_rows > 1 ? sqrt(ss / (n_rows - 1)) : 0.0;
}

__global__ void fill_f64_kernel(double *__restrict__ dst, double v, size_t n) {
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x) dst[t] = v;
}

// ---------------------------------------------------------------------------------------------------------------
// coordinate transformation (once per factor build): divide by the range(s); lon/lat -> unit sphere for *_sphere
// ---------------------------------------------------------------------------------------------------------------
__global__ void transform_locs_kernel(const double *__restrict__ locs /* [n][d] */, double *__restrict__ tl /* [n][dt] */, int n,
                                      CovConst cc) {
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < n; q += gridDim.x * blockDim.x) {
        const double *p = locs + (size_t)q * cc.d;
        double *o = tl + (size_t)q * cc.dt;
        if (cc.covfun == 1 || cc.covfun == 5) {
            const double lon = p[0] * 3.14159265358979323846 / 180.0, lat = p[1] * 3.14159265358979323846 / 180.0;
            o[0] = cos(lat) * cos(lon) / cc.range[0];
            o[1] = cos(lat) * sin(lon) / cc.range[0];
            o[2] = sin(lat) / cc.range[0];
        } else {
            for (int k = 0; k < cc.d; k++) o[k] = p[k] / cc.range[k];
        }
    }
}

// the shared-memory copy of the Matern table (octaves 2^-16 .. 2^6 of the scaled distance, coefficient-major: 22.5 KB)
#define MTS_E0 (-26)
#define MTS_E1 6
#define MTS_SEGS ((MTS_E1 - MTS_E0) * MT_S)

__device__ __forceinline__ double matern_from_smem(const double *__restrict__ tab_s, double s, double &out_ok) {
    const long long bits = __double_as_longlong(s);
    const int e = (int)((bits >> 52) & 0x7ff) - 1023;
    if (e < MTS_E0 || e >= MTS_E1) { out_ok = 0.0; return 0.0; }   // the caller falls back to matern_slow (global table / exact)
    const int seg = (e - MTS_E0) * MT_S + (int)((bits >> 48) & 0xF);
    const long long mant = (bits & 0x000FFFFFFFFFFFFFll) | 0x3FF0000000000000ll;
    const double u = (__longlong_as_double(mant) - __longlong_as_double(mant & (long long)0xFFFF000000000000ull)) * 16.0;
    // layout [4][MTS_SEGS] of coefficient PAIRS: four 128-bit loads per look-up.  ncu with eight 64-bit loads from an [8][MTS_SEGS]
    // layout: 5.6 wavefronts per load (the lanes of a warp sit in ~30 different segments), the L1 data pipe 72 % busy and the
    // top limiter of the Matern build; pairs halve the number of loads at ~7 wavefronts each
    const double2 *t2 = reinterpret_cast<const double2 *>(tab_s);
    const double2 c67 = t2[3 * MTS_SEGS + seg], c45 = t2[2 * MTS_SEGS + seg], c23 = t2[1 * MTS_SEGS + seg], c01 = t2[seg];
    double p = c67.y;
    p = p * (u - mt_node(6)) + c67.x;
    p = p * (u - mt_node(5)) + c45.y;
    p = p * (u - mt_node(4)) + c45.x;
    p = p * (u - mt_node(3)) + c23.y;
    p = p * (u - mt_node(2)) + c23.x;
    p = p * (u - mt_node(1)) + c01.y;
    p = p * (u - mt_node(0)) + c01.x;
    out_ok = 1.0;
    return e >= MT_ESCALED ? p * fast_exp_nonpos(-fast_sqrt_nonneg(s)) : p;
}

// Everything that is not a hit in the shared-memory table.  NOT inlined: the register-resident kernels evaluate the kernel at 55
// (m = 10) to 210 (m = 20) unrolled sites, and with the global-table lookup, pow() and the Bessel routine inlined at each of them the
// Matern build of m = 10 was 20.8 k SASS instructions against 4.7 k for the exponential one -- 330 KB of code, far beyond the
// instruction cache, which is what made it 3x slower (cuobjdump; profiles/r02_matern_factor.txt).
__device__ __noinline__ double matern_slow(const CovConst &cc, double d2) {
    if (d2 == 0.0) return cc.variance;
    if (cc.mtab) {
        double ok;
        const double v = matern_from_table(cc.mtab, d2, ok);
        if (ok != 0.0) return v;
    }
    const double dist = sqrt(d2);
    return cc.normcon * pow(dist, cc.smooth) * bessel_k_real(cc.smooth, dist);
}

template <bool MATERN>
__device__ __forceinline__ double kernel_value_fast(const CovConst &cc, double d2, const double *__restrict__ tab_s = nullptr) {
    if (!MATERN) return cc.variance * fast_exp_nonpos(-fast_sqrt_nonneg(d2));
    if (tab_s) {   // d2 == 0 (coincident sites) misses the table like any other out-of-range argument
        double ok;
        const double v = matern_from_smem(tab_s, d2, ok);
        if (ok != 0.0) return v;
    }
    return matern_slow(cc, d2);
}

template <bool MATERN>
__device__ __forceinline__ double kernel_value(const CovConst &cc, double dist) {
    if (!MATERN) return cc.variance * exp(-dist);
    if (dist == 0.0) return cc.variance;
    if (cc.mtab) {
        double ok;
        const double v = matern_from_table(cc.mtab, dist * dist, ok);
        if (ok != 0.0) return v;
    }
    return cc.normcon * pow(dist, cc.smooth) * bessel_k_real(cc.smooth, dist);
}

// One row of the factor with the triangle in local memory: any M <= MCAP, any DT <= 4, any number of valid neighbours.
// Deliberately not inlined: the register-resident kernel calls it for the (at most m) rows that have fewer than m
// neighbours, so that those rows ride along with the main launch instead of costing a serial 40 us launch of their own.
template <int MCAP, bool MATERN>
__device__ __noinline__ void factor_row_generic(int q, const int *__restrict__ nn, const double *__restrict__ tl,
                                                double *__restrict__ linv, int ld, int M, const CovConst &cc, int *__restrict__ n_bad) {
    const int DT = cc.dt;
    int idx[MCAP];
    int bsize = 0;
    for (int j = 0; j < M; j++) {
        const int v = nn[(size_t)j * ld + q];
        if (v >= 0) idx[bsize++] = v;  // valid slots are a prefix of the row
    }
    double L[MCAP * (MCAP + 1) / 2];
    double x[MCAP];
    bool ok = true;
    for (int a = 0; a < bsize; a++) {
        const double *pa = tl + (size_t)idx[bsize - 1 - a] * DT;
        for (int b = 0; b <= a; b++) {
            double s;
            if (a == b) {
                s = cc.variance + cc.nugget;
            } else {
                const double *pb = tl + (size_t)idx[bsize - 1 - b] * DT;
                double d2 = 0.0;
                for (int c = 0; c < DT; c++) {
                    const double u = pa[c] - pb[c];
                    d2 += u * u;
                }
                s = kernel_value<MATERN>(cc, sqrt(d2));
            }
            for (int k = 0; k < b; k++) s -= L[a * (a + 1) / 2 + k] * L[b * (b + 1) / 2 + k];
            if (a == b) {
                ok = ok && (s > 0.0);
                L[a * (a + 1) / 2 + a] = sqrt(s);
            } else {
                L[a * (a + 1) / 2 + b] = s / L[b * (b + 1) / 2 + b];
            }
        }
    }
    for (int a = bsize - 1; a >= 0; a--) {
        double s = (a == bsize - 1) ? 1.0 : 0.0;
        for (int k = a + 1; k < bsize; k++) s -= L[k * (k + 1) / 2 + a] * x[k];
        x[a] = s / L[a * (a + 1) / 2 + a];
    }
    for (int j = 0; j < M; j++) linv[(size_t)j * ld + q] = (j < bsize) ? x[bsize - 1 - j] : 0.0;
    if (cc.linv_lvl) {
        const int t = cc.lpos[q];
        if (t >= 0)
            for (int j = 0; j < M; j++) cc.linv_lvl[(size_t)j * cc.nsl + t] = (j < bsize) ? x[bsize - 1 - j] : 0.0;
    }
    if (!ok) atomicAdd(n_bad, 1);
}

// ---------------------------------------------------------------------------------------------------------------
// Vecchia factor build, register-resident: one thread per row, M = m+1 and DT compile-time, the (M x M) lower triangle
// lives in registers, all loops fully unrolled.  Rows with fewer than m neighbours (at most m of them) take the generic
// local-memory path (factor_row_generic) inside the same launch (PARTIAL; M <= 24).  FP64-pipe bound: ~M(M-1)/2 exp + sqrt, M^3/6 FMA, M(M+1)/2 div.
// Operation order = oracle_vecchia_linv (oracle/nngp_oracle.c): neighbours farthest-first, self last; row-by-row Cholesky;
// back-substitution for the last row of L^-1.
// ---------------------------------------------------------------------------------------------------------------
template <int M, int DT, bool MATERN, int MINB = 3, bool FAST = true, bool PARTIAL = true>
__global__ void __launch_bounds__(128, MINB) vecchia_factor_reg_kernel(const int *__restrict__ nn, const double *__restrict__ tl,
                                                                 double *__restrict__ linv, int n, int ld, CovConst cc,
                                                                 int *__restrict__ n_bad) {
    // Matern: the interpolation table moves to shared memory, coefficient-major.  The lanes of a warp evaluate the same pair (a, b) of
    // different rows -- similar distances, a handful of table segments -- but each global 64-byte segment read cost up to 32 L1
    // wavefronts per 128-bit load (ncu, m = 20: 340 M sectors of table gathers, 1.65 of the 2.55 ms per 500 k rows)
    __shared__ __align__(16) double tab_s[MATERN ? 8 * MTS_SEGS : 2];
    const bool smem_tab = MATERN && FAST && cc.mtab != nullptr;
    if (smem_tab) {
        for (int i = threadIdx.x; i < 8 * MTS_SEGS; i += blockDim.x) {   // tab_s[((k / 2) * MTS_SEGS + segment) * 2 + (k & 1)] = coefficient k
            const int kk = i / (2 * MTS_SEGS), sg = (i >> 1) % MTS_SEGS, k = 2 * kk + (i & 1);
            tab_s[i] = cc.mtab[(size_t)(sg + (MTS_E0 - MT_EMIN) * MT_S) * 8 + k];
        }
        __syncthreads();
    }
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    int idx[M];
    bool full = true;
#pragma unroll
    for (int j = 0; j < M; j++) {
        idx[j] = nn[(size_t)j * ld + q];
        full = full && (idx[j] >= 0);
    }
    if (!full) {   // one of the first m rows: generic path, concurrently with the rest of this launch
        if (PARTIAL) factor_row_generic<24, MATERN>(q, nn, tl, linv, ld, M, cc, n_bad);
        return;
    }
    double p[M][DT];
#pragma unroll
    for (int k = 0; k < M; k++) {
        const double *src = tl + (size_t)idx[M - 1 - k] * DT;
        if (DT == 2) {
            const double2 v = *reinterpret_cast<const double2 *>(src);
            p[k][0] = v.x;
            p[k][1] = v.y;
        } else {
#pragma unroll
            for (int c = 0; c < DT; c++) p[k][c] = src[c];
        }
    }
    // ncu (profiles/r01_factor_full.txt): FP64 pipe 41 % busy, ~2.6 k FP64 instructions per row of which the 66 divisions
    // and 11 square roots are a third.  One reciprocal square root per pivot replaces them: inv[a] = rsqrt(pivot),
    // L[a][a] = pivot * inv[a], L[a][b] = s * inv[b], x[a] = s * inv[a]  (<= 2 ulp per operation away from the oracle's
    // sqrt / divide; the parity tests bound the effect on a factor row at 1e-10).
    // Matern: the M(M-1)/2 kernel values come from a ROLLED loop (coordinates and values dynamically indexed, i.e. in local memory,
    // L1-resident) instead of M(M-1)/2 unrolled copies of the table lookup: 120 instructions per site made the kernel 11 k (m = 10)
    // to 35 k (m = 20) instructions long, several times the instruction cache.
    constexpr bool ROLLCOV = MATERN && FAST && (M > 11);
    double Kv[ROLLCOV ? M * (M - 1) / 2 : 1];
    if (ROLLCOV) {
        const double *tabp = smem_tab ? tab_s : nullptr;
        int e = 0;
#pragma unroll 1
        for (int a = 1; a < M; a++) {
#pragma unroll 1
            for (int b = 0; b < a; b++) {
                double d2 = 0.0;
#pragma unroll
                for (int c = 0; c < DT; c++) {
                    const double t = p[a][c] - p[b][c];
                    d2 += t * t;
                }
                Kv[e++] = kernel_value_fast<MATERN>(cc, d2, tabp);
            }
        }
    }
    double L[M * (M + 1) / 2];
    double inv[M];
    bool ok = true;
#pragma unroll
    for (int a = 0; a < M; a++) {
#pragma unroll
        for (int b = 0; b <= a; b++) {
            double s;
            if (a == b) {
                s = cc.variance + cc.nugget;
            } else if (ROLLCOV) {
                s = Kv[a * (a - 1) / 2 + b];
            } else {
                double d2 = 0.0;
#pragma unroll
                for (int c = 0; c < DT; c++) {
                    const double t = p[a][c] - p[b][c];
                    d2 += t * t;
                }
                s = FAST ? kernel_value_fast<MATERN>(cc, d2, smem_tab ? tab_s : nullptr) : kernel_value<MATERN>(cc, sqrt(d2));
            }
#pragma unroll
            for (int k = 0; k < b; k++) s -= L[a * (a + 1) / 2 + k] * L[b * (b + 1) / 2 + k];
            if (a == b) {
                ok = ok && (s > 0.0);
                inv[a] = rsqrt(s);
                L[a * (a + 1) / 2 + a] = s * inv[a];
            } else {
                L[a * (a + 1) / 2 + b] = s * inv[b];
            }
        }
    }
    double x[M];
#pragma unroll
    for (int a = M - 1; a >= 0; a--) {
        double s = (a == M - 1) ? 1.0 : 0.0;
#pragma unroll
        for (int k = a + 1; k < M; k++) s -= L[k * (k + 1) / 2 + a] * x[k];
        x[a] = s * inv[a];
    }
#pragma unroll
    for (int j = 0; j < M; j++) linv[(size_t)j * ld + q] = x[M - 1 - j];
    if (cc.linv_lvl) {
        const int t = cc.lpos[q];
        if (t >= 0) {
#pragma unroll
            for (int j = 0; j < M; j++) cc.linv_lvl[(size_t)j * cc.nsl + t] = x[M - 1 - j];
        }
    }
    if (!ok) atomicAdd(n_bad, 1);
}

// ---------------------------------------------------------------------------------------------------------------
// Vecchia factor build, one WARP per row (m = 20: BASELINE config 4).  The 231-entry triangle of a 21 x 21 block does not fit the
// registers of one thread: ncu of the thread-per-row kernel at M = 21 (profiles/r02_matern_m20_factor_before.txt) shows 214
// registers, 8 warps per SM, 23.5 M local-memory load instructions (316 M sectors, half of them L1 misses) and, for the Matern
// family, 340 M sectors of divergent table gathers -- 2.5 ms per 500 k rows against 0.9 ms exponential.  Here lane r owns point r
// and row r of the triangle (M registers):
//   * covariances: the M(M-1)/2 pairs are dealt round-robin to the 32 lanes (coordinates by shuffle), staged through a per-warp
//     shared tile with an odd leading dimension, and read back as rows;
//   * right-looking Cholesky, fully unrolled: step k broadcasts the pivot (one rsqrt per step, as in the thread-per-row kernel)
//     and column k by shuffle; entry (a, b) receives its updates in the order k = 0 .. b-1, the same sequence of FMAs as the
//     row-by-row form of the oracle;
//   * last row of L^-1 = [-b', 1] / sqrt(F) with L_nn' b = u: the factor is transposed through the shared tile so that lane a
//     holds column a, then M-1 steps of (broadcast b_k, one FMA);
//   * the Matern interpolation table (octaves 2^-16 .. 2^6 of the scaled distance: 22.5 KB, coefficient-major so that the lanes
//     of a warp hit different banks) lives in shared memory; outside that range the global table / the exact routine is used;
//   * neighbour ids come in and factor rows go out through shared staging of 32 consecutive rows, so that global accesses are
//     coalesced although a warp works on one row.
// Rows with fewer than m neighbours (at most m of them) take factor_row_generic on lane 0.
// ---------------------------------------------------------------------------------------------------------------
template <int M, int DT, bool MATERN>
__global__ void __launch_bounds__(256, 2) vecchia_factor_warp_kernel(const int *__restrict__ nn, const double *__restrict__ tl,
                                                                     double *__restrict__ linv, int n, int ld, CovConst cc,
                                                                     int *__restrict__ n_bad) {
    constexpr int WARPS = 8, RPB = 32;          // 32 consecutive rows per block round, 4 per warp
    constexpr int LDT = M | 1;                  // odd leading dimension of the per-warp tile
    constexpr int E = M * (M - 1) / 2;          // off-diagonal pairs
    constexpr int ROUNDS = (E + 31) / 32;
    extern __shared__ __align__(16) double smem[];
    double *tab_s = smem;                                                   // [8][MTS_SEGS]   (MATERN only)
    double *tiles = smem + (MATERN ? 8 * MTS_SEGS : 0);                     // [WARPS][M * LDT]
    double *sout = tiles + WARPS * M * LDT;                                 // [M][RPB]
    int *snn = reinterpret_cast<int *>(sout + M * RPB);                     // [M][RPB]
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    double *S = tiles + w * (M * LDT);
    const bool use_tab = MATERN && cc.mtab != nullptr;
    if (use_tab) {
        for (int i = tid; i < 8 * MTS_SEGS; i += 256) {
            const int kk = i / (2 * MTS_SEGS), sg = (i >> 1) % MTS_SEGS, k = 2 * kk + (i & 1);
            tab_s[i] = cc.mtab[(size_t)(sg + (MTS_E0 - MT_EMIN) * MT_S) * 8 + k];
        }
    }
    // the pair (a, b), b < a, that this lane evaluates in each round: the same for every row
    int pa_[ROUNDS], pb_[ROUNDS];
#pragma unroll
    for (int t = 0; t < ROUNDS; t++) {
        const int e = t * 32 + lane;
        int a = (int)((1.0f + sqrtf(8.0f * (float)e + 1.0f)) * 0.5f);
        if (a * (a - 1) / 2 > e) a--;
        if (a * (a + 1) / 2 <= e) a++;
        pa_[t] = (e < E) ? a : 0;
        pb_[t] = (e < E) ? e - a * (a - 1) / 2 : 0;
    }
    const double diag = cc.variance + cc.nugget;
    for (int base = blockIdx.x * RPB; base < n; base += gridDim.x * RPB) {
        __syncthreads();   // the previous round's output has been written, its neighbour ids are no longer needed
        for (int i = tid; i < M * RPB; i += 256) {
            const int j = i / RPB, q = base + (i % RPB);
            snn[i] = q < n ? nn[(size_t)j * ld + q] : -1;
        }
        __syncthreads();
        for (int rl = w; rl < RPB; rl += WARPS) {
            const int q = base + rl;
            if (q >= n) break;
            const int id = lane < M ? snn[(M - 1 - lane) * RPB + rl] : 0;   // point `lane` of the block: neighbours farthest first, self last
            if (!__all_sync(0xffffffffu, id >= 0)) {   // one of the first m rows of the ordering
                if (lane == 0) {
                    factor_row_generic<24, MATERN>(q, nn, tl, linv, ld, M, cc, n_bad);
                    for (int j = 0; j < M; j++) sout[j * RPB + rl] = linv[(size_t)j * ld + q];
                }
                __syncwarp();
                continue;
            }
            double p[DT];
            if (lane < M) {
                const double *src = tl + (size_t)id * DT;
                if (DT == 2) {
                    const double2 v = *reinterpret_cast<const double2 *>(src);
                    p[0] = v.x;
                    p[1] = v.y;
                } else {
#pragma unroll
                    for (int c = 0; c < DT; c++) p[c] = src[c];
                }
            } else {
#pragma unroll
                for (int c = 0; c < DT; c++) p[c] = 0.0;
            }
            // ---- covariances ----
#pragma unroll
            for (int t = 0; t < ROUNDS; t++) {
                double d2 = 0.0;
#pragma unroll
                for (int c = 0; c < DT; c++) {
                    const double u = __shfl_sync(0xffffffffu, p[c], pa_[t]) - __shfl_sync(0xffffffffu, p[c], pb_[t]);
                    d2 += u * u;
                }
                if (t * 32 + lane < E) {
                    double v;
                    if (!MATERN) {
                        v = cc.variance * fast_exp_nonpos(-fast_sqrt_nonneg(d2));
                    } else {
                        v = kernel_value_fast<true>(cc, d2, use_tab ? tab_s : nullptr);
                    }
                    S[pa_[t] * LDT + pb_[t]] = v;
                }
            }
            __syncwarp();
            double A[M];
#pragma unroll
            for (int c = 0; c < M; c++) A[c] = (c < lane && lane < M) ? S[lane * LDT + c] : (c == lane ? diag : 0.0);
            // ---- right-looking Cholesky: after step k, A[k] of lane r >= k is L[r][k] ----
            bool ok = true;
            double myinv = 0.0;
#pragma unroll
            for (int k = 0; k < M; k++) {
                const double dk = __shfl_sync(0xffffffffu, A[k], k);
                ok = ok && (dk > 0.0);
                const double ik = rsqrt(dk);
                myinv = (lane == k) ? ik : myinv;
                A[k] = (lane == k ? dk : A[k]) * ik;
#pragma unroll
                for (int c = 0; c < M; c++) {   // constant trip count (c > k folds at compile time): a triangular bound was left rolled, A[] in local memory
                    if (c > k) {
                        const double lc = __shfl_sync(0xffffffffu, A[k], c);
                        A[c] = (lane >= c) ? fma(-A[k], lc, A[c]) : A[c];
                    }
                }
            }
            // ---- transpose through the tile: lane a gets column a, C[k] = L[k][a] (reusing A) ----
            __syncwarp();
            if (lane < M) {
#pragma unroll
                for (int c = 0; c < M; c++)
                    if (c <= lane) S[lane * LDT + c] = A[c];
            }
            __syncwarp();
#pragma unroll
            for (int k = 0; k < M; k++) A[k] = (k > lane && lane < M) ? S[k * LDT + lane] : 0.0;
            __syncwarp();
            // ---- L_nn' b = u (u = row M-1 of L), then x = [-b, 1] / sqrt(F) ----
            double tacc = A[M - 1], mine = 0.0;
#pragma unroll
            for (int kk = 0; kk < M - 1; kk++) {
                const int k = M - 2 - kk;
                const double bk = __shfl_sync(0xffffffffu, tacc * myinv, k);
                mine = (lane == k) ? bk : mine;
                tacc = (lane < k) ? fma(-A[k], bk, tacc) : tacc;
            }
            const double ilast = __shfl_sync(0xffffffffu, myinv, M - 1);
            if (lane < M) sout[(M - 1 - lane) * RPB + rl] = (lane == M - 1) ? ilast : -mine * ilast;
            if (!ok && lane == 0) atomicAdd(n_bad, 1);
        }
        __syncthreads();
        for (int i = tid; i < M * RPB; i += 256) {
            const int j = i / RPB, q = base + (i % RPB);
            if (q < n) {
                const double v = sout[i];
                linv[(size_t)j * ld + q] = v;
                if (cc.linv_lvl) {
                    const int t = cc.lpos[q];
                    if (t >= 0) cc.linv_lvl[(size_t)j * cc.nsl + t] = v;
                }
            }
        }
    }
}

// Generic factor build: any M <= MCAP, any DT <= 4, rows listed in `rows` (or all rows when rows == nullptr); handles
// partial rows (fewer than m neighbours).  Triangle in local memory (factor_row_generic).
template <int MCAP, bool MATERN>
__global__ void __launch_bounds__(128) vecchia_factor_generic_kernel(const int *__restrict__ nn, const double *__restrict__ tl,
                                                                     double *__restrict__ linv, const int *__restrict__ rows,
                                                                     int n_rows, int ld, int M, CovConst cc,
                                                                     int *__restrict__ n_bad) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_rows) return;
    factor_row_generic<MCAP, MATERN>(rows ? rows[t] : t, nn, tl, linv, ld, M, cc, n_bad);
}

// ---------------------------------------------------------------------------------------------------------------
// row kernels: u_q = sum_j linv[j,q] * (v[nn[j,q]] - shift)
// ---------------------------------------------------------------------------------------------------------------
template <int MT>
__device__ __forceinline__ double row_dot(const int *__restrict__ nn, const double *__restrict__ linv, const double *__restrict__ v,
                                          double shift, int q, int ld, int M) {
    double u = 0.0;
    if (MT > 0) {
        int idx[MT > 0 ? MT : 1];
        double a[MT > 0 ? MT : 1];
#pragma unroll
        for (int j = 0; j < MT; j++) {
            idx[j] = nn[(size_t)j * ld + q];
            a[j] = linv[(size_t)j * ld + q];
        }
#pragma unroll
        for (int j = 0; j < MT; j++)
            if (idx[j] >= 0) u += a[j] * (v[idx[j]] - shift);
    } else {
        for (int j = 0; j < M; j++) {
            const int id = nn[(size_t)j * ld + q];
            if (id >= 0) u += linv[(size_t)j * ld + q] * (v[id] - shift);
        }
    }
    return u;
}

// Vecchia log-likelihood partial sums: partials[b] = (sum log linv[0,q], sum u_q^2) over the rows of block b
// (ll_compressed_sparse_chol, Scripts/mcmc_nngp_update_Gaussian.R:8-12; GpGp::Linv_mult fused with the reductions)
template <int MT>
__global__ void __launch_bounds__(256) loglik_partial_kernel(const int *__restrict__ nn, const double *__restrict__ linv,
                                                             const double *__restrict__ field, double shift, int n, int ld, int M,
                                                             const unsigned char *__restrict__ row_mask,
                                                             double2 *__restrict__ partials) {
    double acc[2] = {0.0, 0.0};
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < n; q += gridDim.x * blockDim.x) {
        if (row_mask && !row_mask[q]) continue;   // sharded field: ghost rows are summed by the rank that owns them
        const double u = row_dot<MT>(nn, linv, field, shift, q, ld, M);
        acc[0] += log(linv[q]);
        acc[1] += u * u;
    }
    block_reduce_sum<2>(acc);
    if (threadIdx.x == 0) partials[blockIdx.x] = make_double2(acc[0], acc[1]);
}

// ---------------------------------------------------------------------------------------------------------------
// TMA-staged variant of the log-likelihood pass.  The slot-major tables make a tile of TR rows M contiguous segments of
// nn (TR*4 B each) and M of linv (TR*8 B each); one elected thread fetches them with cp.async.bulk (the TMA engine's 1-D
// bulk copy, SASS UBLKCP) into a STAGES-deep shared-memory ring, completion is tracked by an mbarrier per stage
// (complete_tx::bytes), and the 256 consumer threads -- one row each -- read their indices / coefficients from shared
// memory and gather the field from L2.  The plain kernel keeps ~22 coalesced loads per thread in flight behind the
// gathers and reached 49 % of DRAM peak in ncu (latency bound); here the streaming part runs ahead of the consumers
// with no register cost, so DRAM stays busy while the gathers of the previous tile resolve.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned int smem_u32(const void *p) { return (unsigned int)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long *bar, unsigned int bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst_smem, const void *src_gmem, unsigned int bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned int parity) {
    unsigned int done, spins = 0;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (!done && ++spins > (1u << 26)) __trap();   // a lost completion must surface as an error, never as a hung device
    } while (!done);
}

template <int MT, int STAGES>
__global__ void __launch_bounds__(256) loglik_tma_kernel(const int *__restrict__ nn, const double *__restrict__ linv,
                                                         const double *__restrict__ field, double shift, int n, int ld,
                                                         const unsigned char *__restrict__ row_mask,
                                                         double2 *__restrict__ partials) {
    constexpr int TR = 256;
    extern __shared__ __align__(128) unsigned char tma_smem[];
    double *lin_s = reinterpret_cast<double *>(tma_smem);                                   // [STAGES][MT][TR]
    int *nn_s = reinterpret_cast<int *>(tma_smem + (size_t)STAGES * MT * TR * 8);           // [STAGES][MT][TR]
    unsigned long long *full = reinterpret_cast<unsigned long long *>(tma_smem + (size_t)STAGES * MT * TR * 12);
    const int tid = threadIdx.x;
    const int n_tiles = (n + TR - 1) / TR;
    const int my_tiles = ((int)blockIdx.x < n_tiles) ? (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

    auto issue = [&](int i) {   // elected thread: fetch this CTA's i-th tile into stage i % STAGES
        const int stage = i % STAGES;
        const int row0 = ((int)blockIdx.x + i * (int)gridDim.x) * TR;
        const int rows = min(TR, ld - row0);   // ld is a multiple of 32: every segment is a multiple of 128 B
        mbar_arrive_expect_tx(full + stage, (unsigned int)(MT * rows * 12));
#pragma unroll
        for (int j = 0; j < MT; j++) {
            tma_bulk_g2s(lin_s + ((size_t)stage * MT + j) * TR, linv + (size_t)j * ld + row0, (unsigned int)(rows * 8), full + stage);
            tma_bulk_g2s(nn_s + ((size_t)stage * MT + j) * TR, nn + (size_t)j * ld + row0, (unsigned int)(rows * 4), full + stage);
        }
    };
    if (tid == 0) {
        for (int s = 0; s < STAGES; s++) mbar_init(full + s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0)
        for (int i = 0; i < STAGES && i < my_tiles; i++) issue(i);

    double acc[2] = {0.0, 0.0};
    for (int i = 0; i < my_tiles; i++) {
        const int stage = i % STAGES;
        mbar_wait(full + stage, (unsigned int)((i / STAGES) & 1));
        const int q = ((int)blockIdx.x + i * (int)gridDim.x) * TR + tid;
        if (q < n && (!row_mask || row_mask[q])) {
            int idx[MT];
            double a[MT];
#pragma unroll
            for (int j = 0; j < MT; j++) {
                idx[j] = nn_s[((size_t)stage * MT + j) * TR + tid];
                a[j] = lin_s[((size_t)stage * MT + j) * TR + tid];
            }
            double u = 0.0;
#pragma unroll
            for (int j = 0; j < MT; j++)
                if (idx[j] >= 0) u += a[j] * (field[idx[j]] - shift);
            acc[0] += log(a[0]);
            acc[1] += u * u;
        }
        __syncthreads();                                   // every consumer is done with this stage ...
        if (tid == 0 && i + STAGES < my_tiles) issue(i + STAGES);   // ... so the TMA engine may refill it
    }
    block_reduce_sum<2>(acc);
    if (tid == 0) partials[blockIdx.x] = make_double2(acc[0], acc[1]);
}

// generic final reduction of NV-vectors of per-block partials: out[k] = sum_b partials[b*NV + k]
template <int NV>
__global__ void __launch_bounds__(1024) reduce_partials_kernel(const double *__restrict__ partials, int n_blocks, double *__restrict__ out) {
    double acc[NV];
#pragma unroll
    for (int k = 0; k < NV; k++) acc[k] = 0.0;
    for (int b = threadIdx.x; b < n_blocks; b += blockDim.x) {
#pragma unroll
        for (int k = 0; k < NV; k++) acc[k] += partials[(size_t)b * NV + k];
    }
    block_reduce_sum<NV>(acc);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < NV; k++) out[k] = acc[k];
    }
}

// out_q = u_q (sparse_chol %*% (v - shift))
template <int MT>
__global__ void __launch_bounds__(256) spmv_rows_kernel(const int *__restrict__ nn, const double *__restrict__ linv,
                                                        const double *__restrict__ v, double shift, int n, int ld, int M,
                                                        double *__restrict__ out) {
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < n; q += gridDim.x * blockDim.x)
        out[q] = row_dot<MT>(nn, linv, v, shift, q, ld, M);
}

// two fused products for the beta_0 update: partial sums of (v.v, u.v) with v = L^-1 1, u = L^-1 field
// (Scripts/mcmc_nngp_update_Gaussian.R:221-222)
template <int MT>
__global__ void __launch_bounds__(256) beta0_partial_kernel(const int *__restrict__ nn, const double *__restrict__ linv,
                                                            const double *__restrict__ field, int n, int ld, int M,
                                                            const unsigned char *__restrict__ row_mask,
                                                            double2 *__restrict__ partials) {
    double acc[2] = {0.0, 0.0};
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < n; q += gridDim.x * blockDim.x) {
        if (row_mask && !row_mask[q]) continue;
        double u = 0.0, v = 0.0;
        for (int j = 0; j < (MT > 0 ? MT : M); j++) {
            const int id = nn[(size_t)j * ld + q];
            if (id >= 0) {
                const double a = linv[(size_t)j * ld + q];
                u += a * field[id];
                v += a * 1.0;
            }
        }
        acc[0] += v * v;
        acc[1] += u * v;
    }
    block_reduce_sum<2>(acc);
    if (threadIdx.x == 0) partials[blockIdx.x] = make_double2(acc[0], acc[1]);
}

// t(sparse_chol) %*% u through the transpose map (no atomics)
__global__ void __launch_bounds__(256) sptmv_kernel(const int *__restrict__ colptr, const int *__restrict__ crow,
                                                    const int *__restrict__ csrc, const int *__restrict__ psite,
                                                    const double *__restrict__ linv, const double *__restrict__ u, int n,
                                                    double *__restrict__ out) {
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < n; q += gridDim.x * blockDim.x) {
        double s = 0.0;
        for (int k = colptr[q]; k < colptr[q + 1]; k++) s += linv[csrc[k]] * u[crow[k]];
        out[psite[q]] = s;   // CSC columns are in processing order, vectors in storage order
    }
}

// dst[map[t]] = src[t]
__global__ void scatter_f64_kernel(double *__restrict__ dst, const double *__restrict__ src, const int *__restrict__ map, int n) {
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) dst[map[t]] = src[t];
}

// accept branch: gather the CSC values of the new current factor and its precision_diag in one pass
// (replaces Matrix::sparseMatrix assembly + the (x^2) %*% indicator product, update_Gaussian.R:73-74,141-142,196-197)
__global__ void __launch_bounds__(256) transpose_values_kernel(const int *__restrict__ colptr, const int *__restrict__ csrc,
                                                               const double *__restrict__ linv, int q0, int q1,
                                                               double *__restrict__ valT, double *__restrict__ pd) {
    for (int q = q0 + blockIdx.x * blockDim.x + threadIdx.x; q < q1; q += gridDim.x * blockDim.x) {
        double s = 0.0;
        for (int k = colptr[q]; k < colptr[q + 1]; k++) {
            const double v = linv[csrc[k]];
            valT[k] = v;
            s += v * v;
        }
        pd[q] = s;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// level-scheduled sparse triangular solve: rows of one level (or, single-block variant, of a run of narrow levels)
//   x_q = (b_q - sum_{j>=1} linv[j,q] x[nn[j,q]]) / linv[0,q]
// optional fused epilogue: y_q = shift + scale * x_q   (initial field draw / ancillary proposal)
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void sptrsv_row(const int *__restrict__ nn, const double *__restrict__ linv, const double *__restrict__ b,
                                           double *__restrict__ x, double *__restrict__ y, double shift, double scale, int q, int ld,
                                           int M) {
    double s = b[q];
    for (int j = 1; j < M; j++) {
        const int id = nn[(size_t)j * ld + q];
        if (id >= 0) s -= linv[(size_t)j * ld + q] * x[id];
    }
    const double xv = s / linv[q];
    x[q] = xv;
    if (y) y[q] = shift + scale * xv;
}

__global__ void __launch_bounds__(256) sptrsv_level_kernel(const int *__restrict__ nn, const double *__restrict__ linv,
                                                           const int *__restrict__ lvl_rows, int lo, int hi,
                                                           const double *__restrict__ b, double *__restrict__ x,
                                                           double *__restrict__ y, double shift, double scale, int ld, int M) {
    const int t = lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (t < hi) sptrsv_row(nn, linv, b, x, y, shift, scale, lvl_rows[t], ld, M);
}

// one CTA walks levels [l0, l1) with a block barrier between them (the head and tail of the DAG are narrow)
__global__ void __launch_bounds__(1024) sptrsv_multilevel_kernel(const int *__restrict__ nn, const double *__restrict__ linv,
                                                                 const int *__restrict__ lvl_rows, const int *__restrict__ lvl_ptr,
                                                                 int l0, int l1, const double *__restrict__ b,
                                                                 double *__restrict__ x, double *__restrict__ y, double shift,
                                                                 double scale, int ld, int M) {
    for (int l = l0; l < l1; l++) {
        const int lo = lvl_ptr[l], hi = lvl_ptr[l + 1];
        for (int t = lo + threadIdx.x; t < hi; t += blockDim.x) sptrsv_row(nn, linv, b, x, y, shift, scale, lvl_rows[t], ld, M);
        __syncthreads();
    }
}

// Synchronisation-free variant (the production path): ONE launch for the whole DAG.  Threads are laid out in level order
// (each level padded to a multiple of 32 so that a warp never holds both a row and one of its ancestors); the solution
// vector doubles as the ready flag: it is pre-filled with a NaN payload that arithmetic never produces, and a thread polls
// its parents in L2 (ld.relaxed.gpu) until their value has been published (st.relaxed.gpu).  Dependencies only point to
// threads with a smaller logical index, and logical CTA indices are handed out by an atomic ticket at CTA start, so every
// CTA a thread can wait on has already started: no deadlock even when the grid exceeds residency.  All index / coefficient
// loads of all levels are in flight at once; only the chain of x dependencies (DAG depth ~200 L2 round trips) serialises.
#define NNGP_SOLVE_SENTINEL 0xFFF8DEADBEEF0001ull

// conditional simulation at new sites: rows of the observed sites are known (x = (field - beta_0)/sd), rows of the new sites
// are marked pending for the sync-free solve, which then only walks the new rows
__global__ void predict_prepare_kernel(unsigned long long *__restrict__ x, const double *__restrict__ known, const int *__restrict__ i2g,
                                       int n0, int n) {
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < n; q += gridDim.x * blockDim.x)
        x[q] = (i2g[q] >= n0) ? NNGP_SOLVE_SENTINEL : (unsigned long long)__double_as_longlong(known[q]);
}

__global__ void fill_u64_kernel(unsigned long long *__restrict__ dst, unsigned long long v, int n) {
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) dst[t] = v;
}

struct PeerTable { double *area[8]; };   // a sharded field: the peers' areas (own entry = own area), see below

__device__ __forceinline__ unsigned long long ld_relaxed_sys_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_sys_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_gpu_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_gpu_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// SHARD: the rows are the OWNED rows of one spatial block in the order of the whole field's DAG levels; a parent that is a ghost
// site is awaited exactly like a local one -- its owner stores the solution value straight into this rank's x over NVLink (the
// 8-byte value is its own ready flag) -- and every owned boundary value is stored into the x of the peers that ghost the site.
struct ShardSolve {
    PeerTable peers;
    const int *sxptr;             // [n_local + 1] by storage id: destinations of an owned boundary site's solution value
    const int2 *sxdst;            // (peer, storage id of the site on that peer)
    unsigned int x_off[8];        // offset (doubles) of the current solve's x buffer inside every peer's area
};

// LVL: nn / linv are the level-ordered copies ([m+1][n_slots], position t of the row list) instead of the storage-ordered tables.
template <int MT, bool SHARD = false, bool LVL = false>
__global__ void __launch_bounds__(256) sptrsv_syncfree_kernel(const int *__restrict__ nn, const double *__restrict__ linv,
                                                              const int *__restrict__ rows_padded, int n_slots,
                                                              const double *__restrict__ b, unsigned long long *x,
                                                              double *__restrict__ y, double shift, double scale, int ld, int M,
                                                              int *ticket, int *err, unsigned int sleep_ns,
                                                              const __grid_constant__ ShardSolve ss, int slot0 = 0,
                                                              unsigned long long *tline = nullptr) {
    // The grid is a sliding window over the level-ordered rows: each CTA repeatedly takes the next chunk of 256 slots.  A
    // bounded window (gridDim.x * 256 rows, a few DAG levels wide) keeps the number of polling threads -- and the L2 traffic
    // they generate -- small; with one thread per row for the whole DAG in flight ncu showed 1.1 GB of DRAM reads and 3 ms.
    // Three chunks are in flight per CTA: (A) ticket + row id of chunk i+2, (B) the index / coefficient loads of chunk
    // i+1, (C) the dependency polls of chunk i -- so the only latency left on a CTA's critical path is the polling itself.
    // Deadlock freedom: tickets are handed out in order and every CTA works through the chunks it holds in increasing
    // order, so the lowest unfinished chunk is always being processed and only depends on finished chunks (or on earlier
    // levels inside itself, which other warps of the same CTA handle).
    __shared__ int s_chunk;
    constexpr int MC = MT > 0 ? MT : 32;
    const int Mr = MT > 0 ? MT : M;
    // slot0 > 0: the walk starts at that slot of the row list (the rows before it are already solved)
    const int n_chunks = (n_slots - slot0 + (int)blockDim.x - 1) / (int)blockDim.x;
    auto take = [&]() -> int {
        __syncthreads();
        if (threadIdx.x == 0) s_chunk = atomicAdd(ticket, 1);
        __syncthreads();
        return s_chunk;
    };
    auto row_of = [&](int chunk) -> int {
        const int t = slot0 + chunk * (int)blockDim.x + (int)threadIdx.x;
        return (chunk < n_chunks && t < n_slots) ? rows_padded[t] : -1;
    };
    int idx[MC], idx1[MC];
    double a[MC], a1[MC];
    double bq = 0.0, bq1 = 0.0;
    auto load_row = [&](int q, int chunk, int (&id)[MC], double (&av)[MC], double &bv) {
        if (q < 0) return;
        const size_t at = LVL ? (size_t)slot0 + (size_t)chunk * blockDim.x + threadIdx.x : (size_t)q;   // LVL: ld is the row list's padded length
#pragma unroll
        for (int j = 0; j < MC; j++) {
            if (j < Mr) {
                id[j] = nn[(size_t)j * ld + at];
                av[j] = linv[(size_t)j * ld + at];
            }
        }
        bv = b[q];
    };
    int c0 = take();
    int q0 = row_of(c0);
    load_row(q0, c0, idx, a, bq);
    int c1 = take();
    int q1 = row_of(c1);
    while (c0 < n_chunks) {
        const int c2 = (c1 < n_chunks) ? take() : n_chunks;   // (A) ticket of chunk i+2 ...
        const int q2 = row_of(c2);                            //     ... and its row id (in flight)
        load_row(q1, c1, idx1, a1, bq1);                      // (B) loads of chunk i+1 (row id arrived last iteration)
        if (q0 >= 0) {                                        // (C) chunk i
            double s = bq;
            // Poll ALL still-pending parents in every round: the loads of one round are independent and overlap, so a
            // round costs one L2 round trip however many parents are outstanding (polling them one after the other cost
            // up to m serial round trips per DAG level: 7 us per level measured, against a 0.38 us flag hop).
            unsigned long long bits[MC];
#pragma unroll
            for (int j = 1; j < MC; j++) bits[j] = (j < Mr && idx[j] >= 0) ? NNGP_SOLVE_SENTINEL : 0ull;
            const double inv_diag = 1.0 / a[0];
            unsigned int spins = 0;
            bool pending = true;
            while (pending) {
#pragma unroll
                for (int j = 1; j < MC; j++)
                    if (bits[j] == NNGP_SOLVE_SENTINEL) bits[j] = SHARD ? ld_relaxed_sys_u64(x + idx[j]) : ld_relaxed_gpu_u64(x + idx[j]);
                pending = false;
#pragma unroll
                for (int j = 1; j < MC; j++) pending = pending || (bits[j] == NNGP_SOLVE_SENTINEL);
                if (pending) {
                    if (++spins > (1u << 22)) {   // never hang the device on a corrupted structure (or a dead peer): report the row
                        if (atomicExch(err, 1) == 0 && SHARD) {
                            int jp = 1;
                            for (int j = 1; j < MC; j++) if (bits[j] == NNGP_SOLVE_SENTINEL) { jp = j; break; }
                            err[2] = q0; err[3] = idx[jp]; err[4] = c0;
                        }
                        break;
                    }
                    if (sleep_ns) __nanosleep(sleep_ns);
                }
            }
            // the arithmetic after the last parent has arrived is the per-level critical path: four interleaved partial sums in a
            // fixed association (deterministic) instead of one chain of m dependent FMAs, and the reciprocal of the diagonal was
            // taken while the polls were in flight
            double acc[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
            for (int j = 1; j < MC; j++)
                if (j < Mr && idx[j] >= 0) acc[j & 3] += a[j] * __longlong_as_double((long long)bits[j]);
            s -= (acc[0] + acc[1]) + (acc[2] + acc[3]);
            const double xv = s * inv_diag;
            if (SHARD) {
                st_relaxed_sys_u64(x + q0, (unsigned long long)__double_as_longlong(xv));
                for (int k = ss.sxptr[q0]; k < ss.sxptr[q0 + 1]; k++) {   // the peers that hold this site as a ghost wait for it
                    const int2 d = ss.sxdst[k];
                    st_relaxed_sys_u64(reinterpret_cast<unsigned long long *>(ss.peers.area[d.x] + ss.x_off[d.x]) + d.y, (unsigned long long)__double_as_longlong(xv));
                }
            } else {
                st_relaxed_gpu_u64(x + q0, (unsigned long long)__double_as_longlong(xv));
            }
            if (y) y[q0] = shift + scale * xv;
        }
        if (tline && (threadIdx.x & 31) == 0) {   // development aid (nngp_solve_timeline): when was this chunk finished
            unsigned long long tnow;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tnow));
            atomicMax(tline + c0, tnow);
        }
        // rotate the pipeline
        c0 = c1; q0 = q1; bq = bq1;
#pragma unroll
        for (int j = 0; j < MC; j++) { idx[j] = idx1[j]; a[j] = a1[j]; }
        c1 = c2; q1 = q2;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// observation-side kernels (gathers through locs_match)
// ---------------------------------------------------------------------------------------------------------------
// S_q = sum of y_minus_xb over the observations located at site q   (residuals_sum_matrix %*% ., update_Gaussian.R:90,260)
__global__ void __launch_bounds__(256) site_obs_sum_kernel(const int *__restrict__ optr, const int *__restrict__ oidx,
                                                           const double *__restrict__ ymx, int n, double *__restrict__ S) {
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < n; q += gridDim.x * blockDim.x) {
        double s = 0.0;
        for (int k = optr[q]; k < optr[q + 1]; k++) s += ymx[oidx[k]];
        S[q] = s;
    }
}

// partial sums of ( (ymx - fnew[lm])^2 , (ymx - f[lm])^2 ); fnew may alias f for a plain SSR
__global__ void __launch_bounds__(256) obs_sq_partial_kernel(const int *__restrict__ lm, const double *__restrict__ ymx,
                                                             const double *__restrict__ fnew, const double *__restrict__ f,
                                                             int n_obs, double2 *__restrict__ partials) {
    double acc[2] = {0.0, 0.0};
    for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < n_obs; o += gridDim.x * blockDim.x) {
        const int s = lm[o];
        const double y = ymx[o];
        const double en = y - fnew[s], eo = y - f[s];
        acc[0] += en * en;
        acc[1] += eo * eo;
    }
    block_reduce_sum<2>(acc);
    if (threadIdx.x == 0) partials[blockIdx.x] = make_double2(acc[0], acc[1]);
}

// ---------------------------------------------------------------------------------------------------------------
// chromatic Gibbs sweep, one launch per colour: sites [q0, q1) are one colour class (contiguous in the internal
// numbering).  Residual-maintained form: r = L^-1 (field - beta0) is kept up to date, so one site costs
// one pass over its CSC column for the conditional mean and one to patch r.  Same-colour sites never share a row
// (they would be moral neighbours), so the r updates inside a launch are conflict-free: no atomics.
// (Scripts/mcmc_nngp_update_Gaussian.R:261-274)
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double sweep_normal(const SweepParams &sp, const double *__restrict__ zbuf, const int *__restrict__ zpos,
                                               const int *__restrict__ gid, int q) {
    if (sp.rng_mode == 0) return zbuf[sp.z_offset + (unsigned long long)zpos[q]];
    return philox_normal((uint32_t)gid[q], (uint32_t)sp.sweep_counter, (uint32_t)(sp.sweep_counter >> 32), sp.key0, sp.key1);
}

// r-independent part of one site's update, computed while the tile is staged (i.e. off the critical path):
//   f_new = c0 - c1 * a,  a = sum_k valT[k] r[crow[k]]   with
//   c1 = e_ls / prec,  c0 = beta0 + (Qss w_old e_ls + e_ln resid) / prec + z / sqrt(prec)
// (algebraically identical to update_Gaussian.R:264-273; differs from the literal order of operations by FP64 rounding)
struct SiteConst { double c0, c1, f_old; };

__device__ __forceinline__ SiteConst site_const(const SweepParams &sp, double f_old, double Qss, double no, double Sq, double z) {
    SiteConst sc;
    const double w_old = f_old - sp.beta0;
    const double prec = sp.e_ls * Qss + sp.e_ln * no;
    const double inv = 1.0 / prec;
    const double resid = Sq - no * sp.beta0;
    sc.c1 = sp.e_ls * inv;
    sc.c0 = sp.beta0 + (Qss * w_old * sp.e_ls + sp.e_ln * resid) * inv + z * sqrt(inv);
    sc.f_old = f_old;
    return sc;
}

__global__ void __launch_bounds__(256) gibbs_color_kernel(const int *__restrict__ colptr, const int *__restrict__ crow,
                                                          const double *__restrict__ valT, const double *__restrict__ pd,
                                                          const double *__restrict__ nobs, const double *__restrict__ S,
                                                          const int *__restrict__ zpos, const int *__restrict__ gid,
                                                          const int *__restrict__ psite,
                                                          const double *__restrict__ zbuf, const SweepParams *__restrict__ spp,
                                                          double *__restrict__ field, double *__restrict__ r, int q0, int q1) {
    const int q = q0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= q1) return;
    const SweepParams sp = *spp;
    const int k0 = colptr[q], k1 = colptr[q + 1];
    const int sq = psite[q];
    const double f_old = field[sq];
    const double w_old = f_old - sp.beta0;
    double a = 0.0;
    for (int k = k0; k < k1; k++) a += valT[k] * r[crow[k]];
    const double Qss = pd[q], no = nobs[q];
    const double prec = sp.e_ls * Qss + sp.e_ln * no;
    const double t = a - Qss * w_old;
    const double resid = S[q] - no * sp.beta0;
    const double mean = sp.beta0 - (1.0 / prec) * (t * sp.e_ls - sp.e_ln * resid);
    const double z = sweep_normal(sp, zbuf, zpos, gid, q);
    const double f_new = mean + z / sqrt(prec);
    const double delta = (f_new - sp.beta0) - w_old;
    for (int k = k0; k < k1; k++) r[crow[k]] += valT[k] * delta;
    field[sq] = f_new;
}

// ---------------------------------------------------------------------------------------------------------------
// Tile kernel with a BLOCKED segmented reduction (production path since the L1 data-pipe analysis in profiles/):
// gibbs_tile_kernel is bound by L1 data-pipe wavefronts, and more than half of them (920 of 1620 per tile, ncu
// l1tex__data_pipe_lsu_wavefronts_mem_shared vs _mem_lgds) are SHARED-memory traffic of the per-site segment sums: one thread
// per site walks its column with a stride of ~11 doubles (2-way bank conflicts on top of the 2 wavefronts a 64-bit request
// needs) and then writes delta once per ENTRY.  Here
//   * products go to shared memory once, padded by one double per 8 (position e + e/8);
//   * thread t reads back the 8 CONTIGUOUS products [8t, 8t+8) -- stride 9 doubles, conflict-free -- and reduces them by
//     runs of equal site (the local site id of every entry is a byte stream, cloc, laid out per tile and padded to the tile
//     capacity so that the 8 ids of a thread are one aligned 64-bit load).  A run that starts inside the thread belongs to
//     a site whose column starts there: its sum goes to sstart[site] (single writer).  A run that continues the previous
//     thread's last run goes to shead[t];
//   * the site's owner adds sstart[site] and the shead[] of the threads its column runs through (1.4 on average): a fixed
//     order, so the result does not depend on scheduling;
//   * delta is written once per SITE; the scatter phase looks it up by the entry's local site id (a broadcast read).
// Shared-memory wavefronts per tile: ~270 instead of ~920.
// ---------------------------------------------------------------------------------------------------------------
#define NNGP_PADPOS(e) ((e) + ((e) >> 3))

// reduces the padded product array by runs; must be called by all THREADS threads between two __syncthreads()
template <int THREADS>
__device__ __forceinline__ void blocked_run_sums(const double *__restrict__ sprod, unsigned long long sid, unsigned int prev_last,
                                                 double *__restrict__ sstart, double *__restrict__ shead) {
    const int t = threadIdx.x;
    unsigned int cur = (unsigned int)(sid & 0xffull);
    bool cont = (cur == prev_last);
    double acc = 0.0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const unsigned int sj = (unsigned int)((sid >> (8 * j)) & 0xffull);
        if (sj != cur) {
            if (cur != 255u) { if (cont) shead[t] = acc; else sstart[cur] = acc; }
            cur = sj;
            cont = false;
            acc = 0.0;
        }
        acc += sprod[9 * t + j];
    }
    if (cur != 255u) { if (cont) shead[t] = acc; else sstart[cur] = acc; }
}

__device__ __forceinline__ double blocked_site_sum(const double *__restrict__ sstart, const double *__restrict__ shead, int site,
                                                   int k0, int k1) {
    double a = sstart[site];
    const int t1 = (k1 - 1) >> 3;
    for (int t = (k0 >> 3) + 1; t <= t1; t++) a += shead[t];
    return a;
}

// one-touch streams (factor values, row ids, local ids): do not allocate in L1 and mark the L2 line evict-first, so that the
// re-used vectors (r, field: 8n bytes each) are what stays resident in the 126 MB L2 while ~13 bytes per entry stream through
__device__ __forceinline__ unsigned long long l2_evict_first_policy() {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ unsigned long long l2_evict_last_policy() {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ double ld_keep_f64(const double *p, unsigned long long pol) {
    double v;
    asm volatile("ld.global.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol) : "memory");
    return v;
}
__device__ __forceinline__ void st_keep_f64(double *p, double v, unsigned long long pol) {
    asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(p), "d"(v), "l"(pol) : "memory");
}
template <int HINT> __device__ __forceinline__ double ld_stream_f64(const double *p, unsigned long long pol) {
    if (!HINT) return *p;
    double v;
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
    return v;
}
template <int HINT> __device__ __forceinline__ int ld_stream_s32(const int *p, unsigned long long pol) {
    if (!HINT) return *p;
    int v;
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
    return v;
}
template <int HINT> __device__ __forceinline__ unsigned int ld_stream_u8(const unsigned char *p, unsigned long long pol) {
    if (!HINT) return *p;
    unsigned int v;
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.u8 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
    return v;
}

__global__ void int_to_f64_kernel(const int *__restrict__ src, double *__restrict__ dst) {
    if (blockIdx.x == 0 && threadIdx.x == 0) dst[0] = (double)src[0];
}

// sharded solve, after every rank has finished: the ghost sites' solution values (stored here by their owners) -> y
__global__ void shard_solve_ghosts_kernel(const unsigned long long *__restrict__ x, const unsigned char *__restrict__ owned, int n,
                                          double *__restrict__ xout, double *__restrict__ y, double shift, double scale) {
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < n; q += gridDim.x * blockDim.x) {
        const double xv = __longlong_as_double((long long)x[q]);
        if (xout) xout[q] = xv;
        if (y && !owned[q]) y[q] = shift + scale * xv;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Sharded field: peers' receive areas, mapped through CUDA IPC (one process per GPU) or addressed directly
// (several contexts in one process).  Area layout (doubles; the header is the same on every rank):
//   [ 16 halo flags (u64) | 16 reduction flags | 2 x 32 reduction slots | 16-double header (u64: parity stride, ...) |
//     receive values, parity 0 | receive values, parity 1 ]
// ---------------------------------------------------------------------------------------------------------------

__device__ __forceinline__ void st_release_sys_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// A ghost slot that holds this bit pattern (a NaN payload no computation produces) is EMPTY.  The value itself is the message:
// the owner stores the 8-byte field value straight into the slot (single-copy atomic), the receiver polls the slot until it is
// not empty, consumes it and empties it again.  No fence, no flag, no counter: one NVLink hop per value (1.2 us measured, against
// 2.9 us for payload + system fence + flag), and a ghost site can be applied as soon as ITS value has landed.
#define NNGP_HALO_EMPTY 0xFFF8DEADBEEF0002ull


// per-context constants of a sharded sweep
struct ShardConst {
    PeerTable peers;
    const int *bptr;              // [n_owned + 1] by processing id: destinations of a boundary site's value (empty for interior sites)
    const int2 *bdst;             // (peer, offset inside the peer's receive values of one parity)
    unsigned long long *tl;       // development aid (nngp_shard_timeline): [K][6] %globaltimer stamps per colour, or nullptr
    const int4 *ginfo;            // ghost sites in receive order: (storage id, first entry of the local column, its end, processing id)
    unsigned long long *state;    // [0] sweeps completed since the peers were connected; [1..4] timeout report
    int *err;
    int world, rank, K;
    unsigned int val_off;         // in doubles; the same on every rank
    unsigned int peer_stride[8];  // distance between the two parities of peer h's receive values (its own number of ghosts, padded)
};
// per-colour launch parameters of a sharded sweep
struct ShardColour {
    int n_tiles, n_btiles;        // tiles of the colour on this rank; the first n_btiles hold its boundary sites
    int g0, g1;                   // this colour's ghost sites: [g0, g1) of the receive order
    int col;
    int ghost_first;              // 1: the ghost CTAs lead the grid (default); 0: they close it (comparison)
};

__device__ __forceinline__ unsigned long long nngp_globaltimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// stamps per colour: [0] first tile past its wait (min), [1] last boundary push (max), [2] first ghost value seen (min), [3] last ghost
// value seen (max), [4] last ghost CTA done (max), [5] last tile done (max)
#define NNGP_TL_MIN(sc, col, k) do { if ((sc).tl) atomicMin((sc).tl + (size_t)(col) * 6 + (k), nngp_globaltimer()); } while (0)
#define NNGP_TL_MAX(sc, col, k) do { if ((sc).tl) atomicMax((sc).tl + (size_t)(col) * 6 + (k), nngp_globaltimer()); } while (0)

// Ghost CTAs of the sweep kernel (blockIdx.x >= n_tiles), one warp per ghost site: load the site's local column (static), wait
// until the owner's value has landed in the site's slot, then replace the ghost value and patch r along the column.  Ghost sites of
// colour c never share a row with owned sites of colour c (the colouring is proper), so this runs concurrently with the tiles.
template <bool PDL, bool TL>
__device__ __forceinline__ void shard_ghost_apply(const ShardConst &sc, const ShardColour &cl, const int *__restrict__ colptr,
                                                  const int *__restrict__ crow, const double *__restrict__ valT,
                                                  const int *__restrict__ psite, double *field, double *r) {
    const unsigned long long epoch = *reinterpret_cast<volatile unsigned long long *>(sc.state);
    unsigned long long *slots = reinterpret_cast<unsigned long long *>(sc.peers.area[sc.rank] + sc.val_off + (size_t)(epoch & 1ull) * sc.peer_stride[sc.rank]);
    const int warps = (int)(blockDim.x >> 5), lane = threadIdx.x & 31;
    const int n_gcta = (int)gridDim.x - cl.n_tiles;
    bool first = true;
    const int gcta = cl.ghost_first ? (int)blockIdx.x : (int)blockIdx.x - cl.n_tiles;
    for (int k = cl.g0 + gcta * warps + (int)(threadIdx.x >> 5); k < cl.g1; k += n_gcta * warps) {
        const int4 gi = sc.ginfo[k];
        const int sq = gi.x, e0 = gi.y, e1 = gi.z;
        const int e = e0 + lane;
        int row = -1;
        double val = 0.0;
        if (e < e1) { row = crow[e]; val = valT[e]; }      // the first 32 entries of the column are in registers before the value arrives
        if (PDL && first) {   // r / field of the previous colour are complete; only then may the NEXT colour's kernel become resident
            asm volatile("griddepcontrol.wait;" ::: "memory");
            asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
        }
        first = false;
        // nothing of this colour touches the ghost site's old value or the r entries along its column (same-colour sites never share a
        // row): fetch them now, so that once the owner's value lands only arithmetic and stores remain on the boundary cycle
        const double f_old = field[sq];
        const double r_old = row >= 0 ? r[row] : 0.0;
        unsigned long long bits = 0ull;
        if (lane == 0) {
            unsigned int spins = 0;
            while ((bits = ld_relaxed_sys_u64(slots + k)) == NNGP_HALO_EMPTY) {
                if (++spins > (1u << 24)) {   // the owner died or fell out of step: report what was awaited instead of hanging the box
                    if (atomicExch(sc.err, 2) != 2) { sc.state[1] = (unsigned long long)cl.col; sc.state[2] = (unsigned long long)k; sc.state[3] = epoch; }
                    bits = 0ull;
                    break;
                }
            }
            st_relaxed_sys_u64(slots + k, NNGP_HALO_EMPTY);   // empty again for the sweep after next (same parity)
            if (TL) { NNGP_TL_MIN(sc, cl.col, 2); NNGP_TL_MAX(sc, cl.col, 3); }
        }
        bits = __shfl_sync(0xffffffffu, bits, 0);
        const double f_new = __longlong_as_double((long long)bits);
        const double delta = f_new - f_old;
        if (lane == 0) field[sq] = f_new;
        if (row >= 0) r[row] = r_old + val * delta;
        for (int e2 = e + 32; e2 < e1; e2 += 32) r[crow[e2]] += valT[e2] * delta;
        if (TL && lane == 0) NNGP_TL_MAX(sc, cl.col, 4);
    }
}

// a boundary site's new value goes straight into the ghost slots of the peers that hold it (NVLink peer stores)
__device__ __forceinline__ void shard_push_site(const ShardConst &sc, unsigned long long epoch, int q, double f_new) {
    const size_t parity = (size_t)(epoch & 1ull);
    for (int k = sc.bptr[q]; k < sc.bptr[q + 1]; k++) {
        const int2 d = sc.bdst[k];
        st_relaxed_sys_u64(reinterpret_cast<unsigned long long *>(sc.peers.area[d.x] + sc.val_off + parity * sc.peer_stride[d.x] + d.y),
                           (unsigned long long)__double_as_longlong(f_new));
    }
}

// SHARD: one spatial block of a larger field with the peer-to-peer transport.  The grid is [boundary tiles | interior tiles |
// ghost CTAs]: boundary tiles come first and store their sites' new values directly into the peers' ghost slots, so the NVLink
// hop overlaps the interior tiles; the ghost CTAs at the end of the grid wait for the values of this colour's ghost sites and
// apply them.  One launch per colour, same PDL chain as the unsharded sweep.
// LATE: griddepcontrol.launch_dependents is issued AFTER this kernel's own griddepcontrol.wait instead of at its top.  Triggering at
// the top lets the whole rest of the sweep become resident early -- colour c+1's CTAs trigger colour c+2 before they wait, and so on --
// and those waiting CTAs hold SM slots.  That is harmless for one stream per GPU (a kernel's predecessor is always fully resident),
// but several streams on one GPU (shards of one field on one device, chains sharing a device) then starve, or -- when they wait for
// each other, as shards do -- deadlock.  With LATE at most two kernels per stream are resident: the running colour and the next
// colour's prologue.  SHARD kernels are always LATE.
template <int THREADS, bool PDL, int MINB, int HINT = 0, bool SHARD = false, bool LATE = SHARD, bool TL = false>
__global__ void __launch_bounds__(THREADS, MINB) gibbs_tile2_kernel(const int4 *__restrict__ tiles, int tile_base,
                                                              const int *__restrict__ colptr, const int *__restrict__ crow,
                                                              const unsigned char *__restrict__ cloc,
                                                              const double *__restrict__ valT, const double *__restrict__ pd,
                                                              const double *__restrict__ nobs, const double *__restrict__ S,
                                                              const int *__restrict__ zpos, const int *__restrict__ gid,
                                                              const int *__restrict__ psite, const double *__restrict__ zbuf,
                                                              const SweepParams *__restrict__ spp, double *__restrict__ field,
                                                              double *__restrict__ r, const __grid_constant__ ShardConst sc, const __grid_constant__ ShardColour cl) {
    constexpr int EPT = 8;
    constexpr int ECAP = THREADS * EPT;
    __shared__ double sprod[ECAP + ECAP / 8];
    __shared__ double sstart[THREADS], shead[THREADS];
    __shared__ double sbc[2];
    const int tid = threadIdx.x;
    if (PDL && !LATE) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    // SHARD: the ghost CTAs lead the grid (CTAs are dispatched in index order: they are resident, with their columns loaded,
    // when the peers' values land -- at the end of the grid they only started after a full wave of tiles had drained)
    const int n_gcta = SHARD ? (int)gridDim.x - cl.n_tiles : 0;
    if (SHARD && (cl.ghost_first ? (int)blockIdx.x < n_gcta : (int)blockIdx.x >= cl.n_tiles)) {
        shard_ghost_apply<PDL, TL>(sc, cl, colptr, crow, valT, psite, field, r);
        return;
    }
    const int bx = (SHARD && cl.ghost_first) ? (int)blockIdx.x - n_gcta : (int)blockIdx.x;
    const bool btile = SHARD && bx < cl.n_btiles;
    const unsigned long long epoch = btile ? *reinterpret_cast<volatile unsigned long long *>(sc.state) : 0ull;
    const int4 tile = tiles[bx];
    const int s0 = tile.x, s1 = tile.y, e0 = tile.z, e1 = tile.w;
    const SweepParams sp = *spp;
    if (e1 - e0 > ECAP) {   // a single site whose column does not fit the tile: whole-CTA reduction
        const int q = s0;
        if (PDL) asm volatile("griddepcontrol.wait;" ::: "memory");
        if (PDL && LATE) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
        double acc[1] = {0.0};
        for (int e = e0 + tid; e < e1; e += THREADS) acc[0] += valT[e] * r[crow[e]];
        block_reduce_sum<1>(acc);
        if (tid == 0) {
            const int sq = psite[q];
            const double w_old = field[sq] - sp.beta0;
            const double Qss = pd[q], no = nobs[q];
            const double prec = sp.e_ls * Qss + sp.e_ln * no;
            const double t = acc[0] - Qss * w_old;
            const double resid = S[q] - no * sp.beta0;
            const double mean = sp.beta0 - (1.0 / prec) * (t * sp.e_ls - sp.e_ln * resid);
            const double f_new = mean + sweep_normal(sp, zbuf, zpos, gid, q) / sqrt(prec);
            sbc[0] = (f_new - sp.beta0) - w_old;
            field[sq] = f_new;
            if (btile) shard_push_site(sc, epoch, q, f_new);
        }
        __syncthreads();
        const double delta = sbc[0];
        for (int e = e0 + tid; e < e1; e += THREADS) r[crow[e]] += valT[e] * delta;
        return;
    }
    const unsigned char *tloc = cloc + (size_t)(tile_base + bx) * ECAP;   // this tile's local site ids (255 = padding)
    const unsigned long long pol = HINT ? l2_evict_first_policy() : 0ull;
    const unsigned long long keep = HINT >= 2 ? l2_evict_last_policy() : 0ull;
    double val[EPT], rr[EPT];
    int row[EPT];
    unsigned int loc[EPT];
#pragma unroll
    for (int k = 0; k < EPT; k++) {
        const int e = e0 + k * THREADS + tid;
        loc[k] = ld_stream_u8<(HINT > 0)>(tloc + k * THREADS + tid, pol);
        if (e < e1) {
            val[k] = ld_stream_f64<(HINT > 0)>(valT + e, pol);
            row[k] = ld_stream_s32<(HINT > 0)>(crow + e, pol);
        } else {
            val[k] = 0.0;
            row[k] = -1;
        }
    }
    const unsigned long long sid = reinterpret_cast<const unsigned long long *>(tloc)[tid];
    const unsigned int prev_last = tid > 0 ? (unsigned int)tloc[8 * tid - 1] : 255u;
    int k0 = 0, k1 = 0, sq = 0;
    SiteConst scn{0.0, 0.0, 0.0};
    if (!PDL) {
#pragma unroll
        for (int k = 0; k < EPT; k++)
            if (row[k] >= 0) rr[k] = r[row[k]];
    }
    int kb0 = 0, kb1 = 0;
    int2 dst0 = make_int2(0, 0);
    if (tid < s1 - s0) {
        const int q = s0 + tid;
        k0 = colptr[q] - e0;
        k1 = colptr[q + 1] - e0;
        sq = psite[q];
        scn = site_const(sp, field[sq], pd[q], nobs[q], S[q], sweep_normal(sp, zbuf, zpos, gid, q));
        if (btile) {   // where the new value goes: known before the wait, so that the push is one store once the value exists
            kb0 = sc.bptr[q];
            kb1 = sc.bptr[q + 1];
            if (kb1 > kb0) dst0 = sc.bdst[kb0];
        }
    }
    if (PDL) {
        asm volatile("griddepcontrol.wait;" ::: "memory");
        if (LATE) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
        if (SHARD && TL && tid == 0) NNGP_TL_MIN(sc, cl.col, 0);
#pragma unroll
        for (int k = 0; k < EPT; k++)
            if (row[k] >= 0) rr[k] = (HINT >= 2) ? ld_keep_f64(r + row[k], keep) : r[row[k]];
    }
#pragma unroll
    for (int k = 0; k < EPT; k++) sprod[NNGP_PADPOS(k * THREADS + tid)] = (row[k] >= 0) ? val[k] * rr[k] : 0.0;
    __syncthreads();
    blocked_run_sums<THREADS>(sprod, sid, prev_last, sstart, shead);
    __syncthreads();
    if (tid < s1 - s0) {
        const double a = blocked_site_sum(sstart, shead, tid, k0, k1);
        const double f_new = scn.c0 - scn.c1 * a;
        sstart[tid] = f_new - scn.f_old;   // delta, once per site (sstart[tid] was read by this thread only)
        if (btile && kb1 > kb0) {
            const size_t parity = (size_t)(epoch & 1ull);
            st_relaxed_sys_u64(reinterpret_cast<unsigned long long *>(sc.peers.area[dst0.x] + sc.val_off + parity * sc.peer_stride[dst0.x] + dst0.y),
                               (unsigned long long)__double_as_longlong(f_new));
            for (int k = kb0 + 1; k < kb1; k++) {   // a corner site: further peers
                const int2 d = sc.bdst[k];
                st_relaxed_sys_u64(reinterpret_cast<unsigned long long *>(sc.peers.area[d.x] + sc.val_off + parity * sc.peer_stride[d.x] + d.y),
                                   (unsigned long long)__double_as_longlong(f_new));
            }
            if (TL) NNGP_TL_MAX(sc, cl.col, 1);
        }
        field[sq] = f_new;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < EPT; k++)
        if (row[k] >= 0) {
            const double rn = rr[k] + val[k] * sstart[loc[k]];
            if (HINT >= 2) st_keep_f64(r + row[k], rn, keep); else r[row[k]] = rn;
        }
    if (SHARD && TL && tid == 0) NNGP_TL_MAX(sc, cl.col, 5);
}

// transposition + precision_diag with the same blocked reduction (see gibbs_tile2_kernel); csrc / linv are one-touch streams
template <int THREADS>
__global__ void __launch_bounds__(THREADS) transpose_tile2_kernel(const int4 *__restrict__ tiles, const int *__restrict__ colptr,
                                                                  const int *__restrict__ csrc, const unsigned char *__restrict__ cloc,
                                                                  const double *__restrict__ linv, double *__restrict__ valT,
                                                                  double *__restrict__ pd) {
    constexpr int EPT = 8;
    constexpr int ECAP = THREADS * EPT;
    __shared__ double ssq[ECAP + ECAP / 8];
    __shared__ double sstart[THREADS], shead[THREADS];
    const int tid = threadIdx.x;
    const int4 tile = tiles[blockIdx.x];
    const int s0 = tile.x, s1 = tile.y, e0 = tile.z, e1 = tile.w;
    if (e1 - e0 > ECAP) {   // single site with an oversize column
        double acc[1] = {0.0};
        for (int e = e0 + tid; e < e1; e += THREADS) {
            const double v = linv[csrc[e]];
            valT[e] = v;
            acc[0] += v * v;
        }
        block_reduce_sum<1>(acc);
        if (tid == 0) pd[s0] = acc[0];
        return;
    }
    const unsigned long long pol = l2_evict_first_policy();
    const unsigned char *tloc = cloc + (size_t)blockIdx.x * ECAP;
    int src[EPT];
#pragma unroll
    for (int k = 0; k < EPT; k++) {
        const int e = e0 + k * THREADS + tid;
        src[k] = (e < e1) ? ld_stream_s32<1>(csrc + e, pol) : -1;
    }
    const unsigned long long sid = reinterpret_cast<const unsigned long long *>(tloc)[tid];
    const unsigned int prev_last = tid > 0 ? (unsigned int)tloc[8 * tid - 1] : 255u;
    int k0 = 0, k1 = 0;
    if (tid < s1 - s0) { k0 = colptr[s0 + tid] - e0; k1 = colptr[s0 + tid + 1] - e0; }
    double v[EPT];
#pragma unroll
    for (int k = 0; k < EPT; k++)
        if (src[k] >= 0) v[k] = linv[src[k]];
#pragma unroll
    for (int k = 0; k < EPT; k++) {
        if (src[k] >= 0) valT[e0 + k * THREADS + tid] = v[k];
        ssq[NNGP_PADPOS(k * THREADS + tid)] = (src[k] >= 0) ? v[k] * v[k] : 0.0;
    }
    __syncthreads();
    blocked_run_sums<THREADS>(ssq, sid, prev_last, sstart, shead);
    __syncthreads();
    if (tid < s1 - s0) pd[s0 + tid] = blocked_site_sum(sstart, shead, tid, k0, k1);
}

// ---------------------------------------------------------------------------------------------------------------
// halo exchange of a sharded field, one colour at a time: pack the new values of the owned boundary sites of
// the colour, (NCCL send/recv between the two kernels), then apply what arrived to the local ghost copies: the ghost value
// is replaced and r is patched along the ghost site's local column exactly as an owned update would have done.
// ---------------------------------------------------------------------------------------------------------------
__global__ void halo_pack_kernel(const int *__restrict__ send_storage, const double *__restrict__ field, int k0, int k1,
                                 double *__restrict__ sendbuf) {
    const int k = k0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (k < k1) sendbuf[k] = field[send_storage[k]];
}

__global__ void halo_apply_kernel(const int *__restrict__ recv_proc, const double *__restrict__ recvbuf, int k0, int k1,
                                  const int *__restrict__ colptr, const int *__restrict__ crow, const double *__restrict__ valT,
                                  const int *__restrict__ psite, double *__restrict__ field, double *__restrict__ r) {
    const int k = k0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= k1) return;
    const int p = recv_proc[k];
    const int sq = psite[p];
    const double f_new = recvbuf[k];
    const double delta = f_new - field[sq];
    field[sq] = f_new;
    for (int e = colptr[p]; e < colptr[p + 1]; e++) r[crow[e]] += valT[e] * delta;   // same-colour sites never share a row
}

// all-reduce (sum) of `count` <= 4 scalars over the ranks, one launch: lane h pushes this rank's partials into peer h's slots
// and raises its flag there, then waits for peer h's flag; lane 0 sums the slots in rank order (deterministic, identical on
// every rank).  Two slot sets alternate with the epoch: a fast rank may already push reduction e+1 while a peer still sums e.
__global__ void allreduce_p2p_kernel(PeerTable peers, const double *__restrict__ partial, int count, int world, int rank,
                                     size_t slot_off, size_t flag_off, unsigned long long epoch, int *err, double *out) {
    const int h = threadIdx.x;
    if (h < world) {
        double *slots = peers.area[h] + slot_off + (size_t)(epoch & 1ull) * 32 + (size_t)rank * 4;
        for (int k = 0; k < count; k++) slots[k] = partial[k];
        __threadfence_system();
        st_release_sys_u64(reinterpret_cast<unsigned long long *>(peers.area[h] + flag_off) + rank, epoch);
        const unsigned long long *flags = reinterpret_cast<const unsigned long long *>(peers.area[rank] + flag_off);
        unsigned int spins = 0;
        while (ld_acquire_sys_u64(flags + h) < epoch) {
            if (++spins > (1u << 24)) { atomicExch(err, 2); break; }
        }
    }
    __syncwarp();
    if (threadIdx.x == 0) {
        const double *area = peers.area[rank];
        for (int k = 0; k < count; k++) {
            double s = 0.0;
            for (int g = 0; g < world; g++) s += __ldcv(area + slot_off + (size_t)(epoch & 1ull) * 32 + (size_t)g * 4 + k);
            out[k] = s;
        }
    }
}

// closes a sweep: next Philox counter / next block of supplied normals; a sharded field also counts the sweep for its flags
__global__ void advance_sweep_kernel(SweepParams *spp, unsigned long long n, unsigned long long *shard_state) {
    if (threadIdx.x == 0) {
        spp->sweep_counter += 1ull;
        spp->z_offset += n;
        if (shard_state) shard_state[0] += 1ull;
    }
}

// FP64 peak micro-benchmark ("a measured FP64 peak from a micro-benchmark", the denominator of the factor
// kernel's roofline fraction): 8 independent DFMA chains per thread, no memory traffic; 2 flops per DFMA.
__global__ void __launch_bounds__(256) fp64_peak_kernel(double *out, int iters, double a, double b) {
    double x0 = threadIdx.x * 1e-3, x1 = x0 + 1.0, x2 = x0 + 2.0, x3 = x0 + 3.0, x4 = x0 + 4.0, x5 = x0 + 5.0, x6 = x0 + 6.0, x7 = x0 + 7.0;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        }
    }
    const double s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (s == 123.456) out[0] = s;   // never true: keeps the chains alive
}

// ---------------------------------------------------------------------------------------------------------------
// regression coefficients (Scripts/mcmc_nngp_update_Gaussian.R:226-250): the only dense algebra on the path.  X$X
// (n_obs x p, column-major) and the site-level design cbind(1, X$X[hctam_scol_1, X$locs]) stay resident in HBM; every
// product with them is a streaming pass (HBM-bound, FP64), the (p+1)-dimensional solves are host scalars.
// ---------------------------------------------------------------------------------------------------------------
// out_o = observed_field_o - field[locs_match_o] + beta_0      (:229, left factor of the crossprod)
__global__ void __launch_bounds__(256) obs_resid_kernel(const int *__restrict__ lm, const double *__restrict__ y,
                                                        const double *__restrict__ field, double beta0, int n_obs,
                                                        double *__restrict__ out) {
    for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < n_obs; o += gridDim.x * blockDim.x) out[o] = y[o] - field[lm[o]] + beta0;
}

// ymx_o = observed_field_o - sum_k X[o,k] beta_k  (= observed_field - mu + beta_0, :249 and its uses :129,260,281);
// X is the block of columns 1..p of cbind(1, X$X)
__global__ void __launch_bounds__(256) obs_minus_xb_kernel(const double *__restrict__ X, const double *__restrict__ beta, int p,
                                                           const double *__restrict__ y, int n_obs, double *__restrict__ ymx) {
    for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < n_obs; o += gridDim.x * blockDim.x) {
        double s = 0.0;
        for (int k = 0; k < p; k++) s += X[(size_t)k * n_obs + o] * __ldg(beta + k);
        ymx[o] = y[o] - s;
    }
}

// out_q = in_q + sign * sum_{l=1..q_cols-1} Xl[q,l] coef_l    (:240 other_field, :245 back; column 0 of Xl is the intercept)
__global__ void __launch_bounds__(256) site_design_axpy_kernel(const double *__restrict__ Xl, const double *__restrict__ coef, int q_cols,
                                                               int n, double sign, const double *__restrict__ in, double *__restrict__ out) {
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < n; q += gridDim.x * blockDim.x) {
        double s = 0.0;
        for (int l = 1; l < q_cols; l++) s += Xl[(size_t)l * n + q] * __ldg(coef + l);
        out[q] = in[q] + sign * s;
    }
}

// v_q = (v_q - sub) + add      (:232 field - beta_0 + innovation[1], evaluated in R's order)
__global__ void __launch_bounds__(256) shift_kernel(double *__restrict__ v, int n, double sub, double add) {
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < n; q += gridDim.x * blockDim.x) v[q] = (v[q] - sub) + add;
}

// C = A^T B for tall column-major A (rows x qa) and B (rows x qb): crossprod() of :78,81,229,241.
// grid (ceil(qa/16), ceil(qb/16), chunks); a CTA owns a 16 x 16 tile of C over one chunk of rows, staged 64 rows at a time
// through shared memory (row-contiguous, i.e. coalesced, loads); per-chunk partials are summed in chunk order by
// atb_reduce_kernel, so the result does not depend on the launch geometry's scheduling (no atomics).
#define NNGP_ATB_ROWS 64
__global__ void __launch_bounds__(256) atb_partial_kernel(const double *__restrict__ A, int qa, const double *__restrict__ B, int qb,
                                                          int rows, int rows_per_chunk, double *__restrict__ part) {
    __shared__ double As[16][NNGP_ATB_ROWS + 1], Bs[16][NNGP_ATB_ROWS + 1];
    const int ta = threadIdx.x & 15, tb = threadIdx.x >> 4;
    const int ja0 = blockIdx.x * 16, jb0 = blockIdx.y * 16;
    const int r_begin = blockIdx.z * rows_per_chunk, r_end = min(rows, r_begin + rows_per_chunk);
    double acc = 0.0;
    for (int r0 = r_begin; r0 < r_end; r0 += NNGP_ATB_ROWS) {
        for (int e = threadIdx.x; e < 16 * NNGP_ATB_ROWS; e += 256) {
            const int col = e / NNGP_ATB_ROWS, rr = e % NNGP_ATB_ROWS, r = r0 + rr;
            As[col][rr] = (r < r_end && ja0 + col < qa) ? A[(size_t)(ja0 + col) * rows + r] : 0.0;
            Bs[col][rr] = (r < r_end && jb0 + col < qb) ? B[(size_t)(jb0 + col) * rows + r] : 0.0;
        }
        __syncthreads();
#pragma unroll 8
        for (int rr = 0; rr < NNGP_ATB_ROWS; rr++) acc += As[ta][rr] * Bs[tb][rr];
        __syncthreads();
    }
    if (ja0 + ta < qa && jb0 + tb < qb) part[(size_t)blockIdx.z * qa * qb + (size_t)(jb0 + tb) * qa + (ja0 + ta)] = acc;
}

// C[e] = sum over chunks (ascending) of part[chunk][e]; C is qa x qb column-major
__global__ void __launch_bounds__(256) atb_reduce_kernel(const double *__restrict__ part, int n_chunks, int qaqb, double *__restrict__ C) {
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < qaqb; e += gridDim.x * blockDim.x) {
        double s = 0.0;
        for (int k = 0; k < n_chunks; k++) s += part[(size_t)k * qaqb + e];
        C[e] = s;
    }
}


}  // namespace nngp
