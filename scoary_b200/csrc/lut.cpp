// lut.cpp -- host-side log-factorial table for the Fisher kernel.
//
// log(k!) for k = 0..n as double-double (hi + lo), computed in binary128 so
// that hi+lo carries ~1e-32 relative error.  A plain double table has an
// absolute error of ~1e-11 at n = 1e4 (ulp of 8e4), which would eat the
// 1e-10 relative budget against scipy.stats.fisher_exact (the arithmetic the
// reference calls at scoary/methods.py:854).  Compiled by g++ (not nvcc)
// because of __float128.
#include <quadmath.h>
#include <stdint.h>

extern "C" void sb_build_logfact_dd(int32_t n, double *hi_lo /* [n+1][2] */)
{
    __float128 acc = 0;
    for (int32_t k = 0; k <= n; ++k) {
        if (k > 1) acc += logq((__float128)k);
        double hi = (double)acc;
        double lo = (double)(acc - (__float128)hi);
        hi_lo[2 * k] = hi;
        hi_lo[2 * k + 1] = lo;
    }
}
