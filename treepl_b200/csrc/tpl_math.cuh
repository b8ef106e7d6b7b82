// fp64 device math for the PL kernels (sm_100a).
//
// The PL objective needs one log and one reciprocal per branch (SURVEY.md §7 "hard parts"):
// on B200 the fp64 pipe (64 DFMA/clk/SM) sits right at the HBM ridge for this kernel, so the
// libdevice log (~30 fp64 instructions) and the IEEE division (~12 + a slow path) are replaced
// by table/Newton forms of ~9 and 4 fp64 instructions whose error (<= ~2 ulp) is far inside the
// parity gates (objective 1e-10, gradient 1e-8 relative).  Inputs outside the "ordinary" range
// take the exact libdevice path, so infinities, zeros, subnormals (durations forced to DBL_MIN,
// pl_calc_parallel.cpp:905) behave exactly as IEEE arithmetic does in the reference.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tpl {

constexpr int LOG_TABLE_BITS = 8;
constexpr int LOG_TABLE_N = 1 << LOG_TABLE_BITS;      // 256 entries x (invc, logc) = 4 KB
constexpr double DBLMIN = 2.2250738585072014e-308;    // numeric_limits<double>::min()

// x in [2^-1000, 2^1000]: exponent arithmetic below cannot overflow and x is normal
__device__ __forceinline__ bool ordinary(double x) {
    return (x >= 9.3326361850321888e-302) && (x <= 1.0715086071862673e+301);
}

// 1/x for ordinary x: MUFU.RCP64H seed (rel. error <= 2^-23, low word zero) + two Newton steps.
__device__ __forceinline__ double fast_rcp(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    return y;
}

// natural log for ordinary x > 0.
//   x = 2^k * m, m in [OFF, 2*OFF) with OFF ~ 0.7071 (so that log x is not formed by cancellation
//   around x = 1); i = top 8 mantissa bits of (bits(x) - bits(OFF)); z = m*invc[i] - 1, |z| < 2^-9;
//   log x = k ln2 + logc[i] + (z - z^2/2 + z^3/3 - z^4/4 + z^5/5), truncation < 2^-54/6 absolute.
// tab = shared-memory copy of the (invc, logc) table built on the host in long double.
__device__ __forceinline__ double fast_log(double x, const double2* __restrict__ tab) {
    const uint64_t OFF = 0x3fe6a09e00000000ull;
    uint64_t ix = (uint64_t)__double_as_longlong(x);
    uint64_t tmp = ix - OFF;
    int i = (int)((tmp >> (52 - LOG_TABLE_BITS)) & (LOG_TABLE_N - 1));
    int k = (int)((int64_t)tmp >> 52);
    uint64_t iz = ix - (tmp & 0xfff0000000000000ull);
    double m = __longlong_as_double((long long)iz);
    double2 t = tab[i];
    double z = fma(m, t.x, -1.0);
    // k -> double without the slow I2F.F64: 2^52+2^51 magic
    double kd = __longlong_as_double((long long)(0x4338000000000000ull + (uint64_t)(int64_t)k)) - 6755399441055744.0;
    double q = fma(z, 0.2, -0.25);
    q = fma(z, q, 0.33333333333333333);
    q = fma(z, q, -0.5);
    double z2 = z * z;
    double hi = fma(kd, 0.69314718055994530942, t.y);
    double r = fma(z2, q, z);
    return hi + r;
}

}  // namespace tpl
