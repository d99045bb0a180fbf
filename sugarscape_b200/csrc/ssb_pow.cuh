// ssb_pow.cuh -- bit-exact restatement of glibc 2.39's double pow as CPython calls it on an
// x86-64 host with FMA3 (the __pow_fma variant of sysdeps/ieee754/dbl-64/e_pow.c, algorithm by
// Szabolcs Nagy, ARM optimized-routines).
//
// Why bit-exact: the reference ranks candidate cells by welfare = s**a * p**b (reference
// agent.py:1189,1200, CPython float ** -> libm pow).  With integer cell resources and small
// integer metabolisms, candidate cells are often MATHEMATICALLY tied (4*9 = 6*6); which one wins
// is decided by pow's last-bit rounding, so a merely accurate device pow (CUDA's is 2 ulp)
// diverges from the reference trajectory after a few thousand decisions.  The sequence below
// follows the compiled glibc code operation for operation, including which multiply-adds are
// fused; tests/test_pow_exact.py checks it against the host libm on millions of inputs.
//
// Plain C so that gcc can compile it for that test; nvcc compiles it for the device.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define SSB_POW_FN __device__ __forceinline__
#define SSB_POW_CORE_FN __device__ __noinline__
#define SSB_POW_TABLE static __device__ const
#else
#define SSB_POW_FN static inline
#define SSB_POW_CORE_FN static inline
#define SSB_POW_TABLE static const
#endif
#include "ssb_pow_tables.cuh"

#if defined(__CUDA_ARCH__)
#define SSB_AS_DOUBLE(u) __longlong_as_double((long long)(u))
#define SSB_AS_U64(d) ((uint64_t)__double_as_longlong(d))
#define SSB_FMA(a, b, c) __fma_rn((a), (b), (c))
#define SSB_TAB(t, i) (t[i])
#else
SSB_POW_FN double ssb_as_double_(uint64_t u) { double d; memcpy(&d, &u, 8); return d; }
SSB_POW_FN uint64_t ssb_as_u64_(double d) { uint64_t u; memcpy(&u, &d, 8); return u; }
#define SSB_AS_DOUBLE(u) ssb_as_double_(u)
#define SSB_AS_U64(d) ssb_as_u64_(d)
#define SSB_FMA(a, b, c) fma((a), (b), (c))
#define SSB_TAB(t, i) (t[i])
#endif

// Valid for finite x > 0 (normal), 2^-65 <= |y| < 2^63 and |y*log(x)| in [2^-54, 512): every
// welfare evaluation of the simulation.  Returns 0 in *ok otherwise (callers fall back).
SSB_POW_CORE_FN double ssb_glibc_pow_core(double x, double y, int *ok) {
    const uint64_t ix = SSB_AS_U64(x);
    *ok = 1;
    // ---- log_inline: log(x) = k ln2 + log(c) + log1p(z/c - 1), as hi + lo ----
    const uint64_t tmp = ix - 0x3fe6955500000000ull;
    const int i = (int)((tmp >> 45) & 127);
    const int k = (int)((int64_t)tmp >> 52);
    const uint64_t iz = ix - (tmp & 0xfff0000000000000ull);
    const double z = SSB_AS_DOUBLE(iz);
    const double kd = (double)k;
    const double invc = SSB_AS_DOUBLE(SSB_TAB(ssb_pow_log_tab, 3 * i));
    const double logc = SSB_AS_DOUBLE(SSB_TAB(ssb_pow_log_tab, 3 * i + 1));
    const double logctail = SSB_AS_DOUBLE(SSB_TAB(ssb_pow_log_tab, 3 * i + 2));
    const double A0 = SSB_AS_DOUBLE(SSB_TAB(ssb_pow_log_poly, 0)), A1 = SSB_AS_DOUBLE(SSB_TAB(ssb_pow_log_poly, 1));
    const double A2 = SSB_AS_DOUBLE(SSB_TAB(ssb_pow_log_poly, 2)), A3 = SSB_AS_DOUBLE(SSB_TAB(ssb_pow_log_poly, 3));
    const double A4 = SSB_AS_DOUBLE(SSB_TAB(ssb_pow_log_poly, 4)), A5 = SSB_AS_DOUBLE(SSB_TAB(ssb_pow_log_poly, 5));
    const double A6 = SSB_AS_DOUBLE(SSB_TAB(ssb_pow_log_poly, 6));
    const double t1 = SSB_FMA(kd, SSB_AS_DOUBLE(SSB_POW_LN2HI), logc);
    const double lo1 = SSB_FMA(kd, SSB_AS_DOUBLE(SSB_POW_LN2LO), logctail);
    const double r = SSB_FMA(z, invc, -1.0);
    const double ar = r * A0;
    const double p12 = SSB_FMA(r, A2, A1);
    const double p34 = SSB_FMA(r, A4, A3);
    const double t2 = r + t1;
    const double lo2 = (t1 - t2) + r;
    const double ar2 = r * ar;
    const double ar3 = r * ar2;
    const double lo3 = SSB_FMA(ar, r, -ar2);
    const double hi = t2 + ar2;
    const double p56 = SSB_FMA(r, A6, A5);
    const double lo4 = (t2 - hi) + ar2;
    const double q = SSB_FMA(ar2, SSB_FMA(p56, ar2, p34), p12);
    double lo = lo1 + lo2;
    lo = lo + lo3;
    lo = lo + lo4;
    lo = SSB_FMA(ar3, q, lo);
    const double lhi = hi + lo;
    const double llo = (hi - lhi) + lo;
    // ---- y * log(x) as ehi + elo ----
    const double ehi = y * lhi;
    const double elo = SSB_FMA(y, llo, SSB_FMA(lhi, y, -ehi));
    // ---- exp_inline(ehi, elo) ----
    const uint32_t abstop = (uint32_t)(SSB_AS_U64(ehi) >> 52) & 0x7ff;
    if (abstop - 0x3c9u > 0x3eu) { *ok = 0; return 0.0; }
    const double shift = SSB_AS_DOUBLE(SSB_EXP_SHIFT);
    double kd2 = SSB_FMA(ehi, SSB_AS_DOUBLE(SSB_EXP_INVLN2N), shift);
    const uint64_t ki = SSB_AS_U64(kd2);
    kd2 = kd2 - shift;
    double rr = SSB_FMA(kd2, SSB_AS_DOUBLE(SSB_EXP_NEGLN2HIN), ehi);
    rr = SSB_FMA(kd2, SSB_AS_DOUBLE(SSB_EXP_NEGLN2LON), rr);
    rr = elo + rr;
    const int idx = 2 * (int)(ki & 127);
    const double tail = SSB_AS_DOUBLE(SSB_TAB(ssb_exp_tab, idx));
    const uint64_t sbits = SSB_TAB(ssb_exp_tab, idx + 1) + (ki << 45);
    const double c23 = SSB_FMA(rr, SSB_AS_DOUBLE(SSB_EXP_C3), SSB_AS_DOUBLE(SSB_EXP_C2));
    const double tr = rr + tail;
    const double r2 = rr * rr;
    const double c45 = SSB_FMA(rr, SSB_AS_DOUBLE(SSB_EXP_C5), SSB_AS_DOUBLE(SSB_EXP_C4));
    double t = SSB_FMA(c23, r2, tr);
    t = SSB_FMA(c45, r2 * r2, t);
    const double scale = SSB_AS_DOUBLE(sbits);
    return SSB_FMA(t, scale, scale);
}

// CPython float.__pow__ for the operand range of the simulation (x >= 0 finite, y >= 0 finite).
// *exact is cleared when the argument falls outside the restated path.
SSB_POW_FN double ssb_py_pow(double x, double y, int *exact) {
    *exact = 1;
    if (y == 0.0) return 1.0;            // float_pow: anything ** 0 == 1
    if (x == 1.0) return 1.0;
    if (y == 1.0) return x;              // exact in glibc as well (result is representable)
    if (x == 0.0) return 0.0;            // y > 0
    const uint64_t ix = SSB_AS_U64(x), iy = SSB_AS_U64(y);
    const uint32_t topx = (uint32_t)(ix >> 52), topy = (uint32_t)(iy >> 52) & 0x7ff;
    if (topx - 1u > 0x7fdu || topy - 0x3beu > 0x7fu) { *exact = 0; return pow(x, y); }
    int ok;
    const double v = ssb_glibc_pow_core(x, y, &ok);
    if (!ok) { *exact = 0; return pow(x, y); }
    return v;
}
