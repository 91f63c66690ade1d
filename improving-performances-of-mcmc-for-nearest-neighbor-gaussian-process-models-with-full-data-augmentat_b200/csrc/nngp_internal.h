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
void build_csc(const int *NNarray, int n, int m, std::vector<int64_t> &ptr, std::vector<int> &rows, std::vector<int> &slots);
int greedy_coloring(const int *NNarray, int n, int m, int *coloring);
void order_maxmin(const double *locs_cm, int n, int d, int *order);
int solve_levels(const int *NNarray, int n, int m, std::vector<int> &level);

// r_stream.cpp: the random-number stream R hands to the reference sampler (Mersenne-Twister, inversion normals), so that
// nngp_chain_run can consume draws in exactly the order Scripts/mcmc_nngp_update_Gaussian.R does.
class RStream {
public:
    void set_seed(uint32_t seed);
    double unif_rand();
    double norm_rand();
    void rnorm(double *out, int64_t n);
private:
    uint32_t genrand();
    uint32_t mt_[624];
    int mti_ = 625;
};

}  // namespace nngp
#endif
