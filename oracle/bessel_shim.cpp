// oracle/bessel_shim.cpp -- TEST INFRASTRUCTURE ONLY.
// GpGp's Matern kernels call boost::math::cyl_bessel_k (not vendored in /root/reference); the C++17 standard library's
// std::cyl_bessel_k computes the same function and is what the oracle uses as an implementation-independent check of the
// device-side Temme evaluation.
#include <cmath>
extern "C" double oracle_bessel_k(double nu, double x) { return std::cyl_bessel_k(nu, x); }
