// nngp_internal.h -- declarations shared by the host-only (host_graph.cpp, r_stream.cpp) and CUDA (nngp_b200.cu)
// translation units of libnngp_b200.so.  Nothing here is part of the C ABI (that is include/nngp_b200.h).
#ifndef NNGP_INTERNAL_H
#define NNGP_INTERNAL_H
#include <cstdint>
#include <vector>

namespace nngp {

void set_error(const char *fmt, ...);

// host_graph.cpp
void find_ordered_nn(const double *locs_cm, int n, int d, int m, int *NNarray);
void build_csc(const int *NNarray, int n, int m, std::vector<int64_t> &ptr, std::vector<int> &rows, std::vector<int> *slots_opt);
int greedy_coloring(const int *NNarray, int n, int m, int *coloring);
void order_maxmin(const double *locs_cm, int n, int d, int *order);
class RStream;
void order_maxmin_gpgp(const double *locs_cm, int n, int d, bool lonlat, RStream &rs, int *order);
void find_ordered_nn_gpgp(const double *locs_cm, int n, int d, int m, RStream &rs, int *NNarray);
int solve_levels(const int *NNarray, int n, int m, std::vector<int> &level);

// r_stream.cpp: the random-number stream R hands to the reference sampler (Mersenne-Twister, inversion normals), so that
// nngp_chain_run can consume draws in exactly the order Scripts/mcmc_nngp_update_Gaussian.R does -- and, for a host that is
// not R (the Python mirror), the draws of Scripts/mcmc_nngp_initialize.R: sample() (R >= 3.6 "Rejection") and rbeta().
class RStream {
public:
    void set_seed(uint32_t seed);
    double unif_rand();
    double norm_rand();
    void rnorm(double *out, int64_t n);
    double unif_index(double dn);                    // R_unif_index: uniform on 0 .. dn-1 by rejection on ceil(log2(dn)) bits
    void sample_int(int n, int size, int *out);      // sample.int(n, size), without replacement, 1-based
    double rbeta(double aa, double bb);              // Cheng's BB (min(aa, bb) > 1); NaN otherwise
    void load(const int *state625);                  // state625[0] = mti, [1..624] = mt: R's .Random.seed[2:626]
    void store(int *state625) const;
private:
    uint32_t genrand();
    uint32_t mt_[624];
    int mti_ = 625;
};

}  // namespace nngp
#endif
