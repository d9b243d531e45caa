// ste_fastmath.cuh - fp64 elementary functions written for the FP64 pipe of sm_100a.
//
// Why not the CUDA math library: its sincos/atan2/asin/sqrt/div inline ~850 SASS instructions per
// process-model evaluation, of which only ~185 are FP64 arithmetic - the rest are UMOV pairs
// materialising polynomial coefficients, integer fix-ups and slow-path calls (profiles/, round 1).
// At the occupancy these kernels reach, instruction ISSUE bound the kernel before the FP64 pipe
// did.  The versions below keep every coefficient in __constant__ memory, so DFMA reads it as a
// constant-bank operand (no UMOV), are branch-free on the arguments the filter produces, and fall
// back to the library only for huge or non-finite arguments.
//
// Accuracy: <= 2 ulp on the fast paths (tests/test_gpu_parity.py::test_fastmath_accuracy compares
// against the CUDA library over the argument ranges the filter uses).  Coefficients are the
// classical fdlibm minimax sets for sin/cos on [-pi/4, pi/4] and atan on [-7/16, 7/16].
#pragma once
#include <math.h>
#include <stdint.h>
#if defined(__CUDACC__)
#include <cuda_runtime.h>
#define STE_DEV __host__ __device__ __forceinline__
#define STE_HD __host__ __device__
#define STE_COLD __host__ __device__ __noinline__
#else
#define STE_DEV inline
#define STE_HD
#define STE_COLD __attribute__((noinline)) inline
#endif

#if defined(__CUDACC__)
#define STE_CONST __constant__
#else
#include <string.h>
#define STE_CONST static const
#endif

// streaming (evict-first) global accesses for write-once / read-once per-step state
#if defined(__CUDA_ARCH__)
#define STE_STORE_STREAM(p, v) __stcs((p), (v))
#define STE_LOAD_STREAM(p) __ldcs(p)
#else
#define STE_STORE_STREAM(p, v) (*(p) = (v))
#define STE_LOAD_STREAM(p) (*(p))
#endif

namespace ste {

// ---- coefficient tables (constant bank on the device) ---------------------------------------- //
STE_CONST double kSinC[6] = {-1.66666666666666324348e-01, 8.33333333332248946124e-03, -1.98412698298579493134e-04,
                             2.75573137070700676789e-06,  -2.50507602534068634195e-08, 1.58969099521155010221e-10};
STE_CONST double kCosC[6] = {4.16666666666666019037e-02,  -1.38888888888741095749e-03, 2.48015872894767294178e-05,
                             -2.75573143513906633035e-07, 2.08757232129817482790e-09,  -1.13596475577881948265e-11};
STE_CONST double kAtanC[11] = {3.33333333333329318027e-01,  -1.99999999998764832476e-01, 1.42857142725034663711e-01,
                               -1.11111104054623557880e-01, 9.09088713343650656196e-02,  -7.69187620504482999495e-02,
                               6.66107313738753120669e-02,  -5.83357013379057348645e-02, 4.97687799461593236017e-02,
                               -3.65315727442169155270e-02, 1.62858201153657823623e-02};
// pi/2 in three parts (hi + mid + lo), 2/pi, and the round-to-nearest magic number 1.5 * 2^52
STE_CONST double kTrigK[6] = {1.5707963267948966e+00, 6.1232339957367574e-17, 8.4784276603688985e-32,
                              6.36619772367581382433e-01, 6755399441055744.0, 105615.0};

constexpr double kPi = 3.14159265358979323846;
constexpr double kPiHalf = 1.57079632679489661923;
constexpr double kPiQuarter = 0.78539816339744830962;
constexpr double kTanPiEighth = 0.41421356237309504880;

// ---- bit access and MUFU seeds (the only architecture-specific pieces) -------------------------- //
// Host builds (tools/host_emul, nvcc's host pass) emulate the seeds' ~20-bit accuracy so that the
// Newton / polynomial code below is exercised bit-for-bit the same way on the CPU sandbox.
STE_DEV uint64_t f64_bits(double v) {
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(v);
#else
    uint64_t u;
    memcpy(&u, &v, 8);
    return u;
#endif
}
STE_DEV double f64_from_bits(uint64_t u) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)u);
#else
    double v;
    memcpy(&v, &u, 8);
    return v;
#endif
}
STE_DEV double seed_rcp(double x) {
#if defined(__CUDA_ARCH__)
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    return y;
#else
    const double xt = f64_from_bits(f64_bits(x) & 0xFFFFFFFF00000000ull);
    return f64_from_bits(f64_bits(1.0 / xt) & 0xFFFFFFFF00000000ull);
#endif
}
STE_DEV double seed_rsqrt(double x) {
#if defined(__CUDA_ARCH__)
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    return y;
#else
    const double xt = f64_from_bits(f64_bits(x) & 0xFFFFFFFF00000000ull);
    return f64_from_bits(f64_bits(1.0 / sqrt(xt)) & 0xFFFFFFFF00000000ull);
#endif
}

// ---- reciprocal, reciprocal square root, square root, division ------------------------------- //
// MUFU seed (~20 bits) + Newton steps in the FP64 pipe; no denormal / special-case slow paths:
// callers guarantee finite, normal, non-zero arguments (or guard the result themselves).
STE_DEV double fast_rcp(double x) {
    double y = seed_rcp(x);
    double e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    return y;
}

STE_DEV double fast_rsqrt(double x) {
    double y = seed_rsqrt(x);
    // y <- y (1 + e/2 + 3 e^2 / 8), e = 1 - x y^2 : cubic convergence, 2^-20 -> 2^-60
    const double t = x * y;
    const double e = fma(-t, y, 1.0);
    const double p = fma(0.375, e, 0.5);
    return fma(y * e, p, y);
}

// sqrt for x >= 0 (x == 0 -> 0)
STE_DEV double fast_sqrt(double x) {
    const double y = fast_rsqrt(x);
    double s = x * y;
    const double r = fma(-s, s, x);   // residual of the rounded product
    s = fma(0.5 * y, r, s);
    return (x > 0.0) ? s : 0.0;
}

STE_DEV double fast_div(double num, double den) {
    const double r = fast_rcp(den);
    const double q = num * r;
    const double rem = fma(-den, q, num);
    return fma(rem, r, q);
}

// ---- sin and cos together --------------------------------------------------------------------- //
// Branch-free; valid for |x| <= kSinCosMaxArg (three-term Cody-Waite reduction).  Callers check
// the range once per filter step and take the library ("cold") path otherwise.
constexpr double kSinCosMaxArg = 105615.0;

STE_DEV void fast_sincos(double x, double *sn, double *cs) {
    // n = rint(x * 2/pi) through the magic-number trick; the low word of t holds n (two's complement)
    const double t = fma(x, kTrigK[3], kTrigK[4]);
    const uint32_t q = (uint32_t)f64_bits(t);
    const double n = t - kTrigK[4];
    double r = fma(-n, kTrigK[0], x);
    r = fma(-n, kTrigK[1], r);
    r = fma(-n, kTrigK[2], r);
    const double z = r * r;
    // sin(r) = r + r z (S1 + z (S2 + ... )) ; cos(r) = 1 - z/2 + z^2 (C1 + z (C2 + ...))
    double ps = fma(z, kSinC[5], kSinC[4]);
    double pc = fma(z, kCosC[5], kCosC[4]);
    ps = fma(z, ps, kSinC[3]);
    pc = fma(z, pc, kCosC[3]);
    ps = fma(z, ps, kSinC[2]);
    pc = fma(z, pc, kCosC[2]);
    ps = fma(z, ps, kSinC[1]);
    pc = fma(z, pc, kCosC[1]);
    ps = fma(z, ps, kSinC[0]);
    pc = fma(z, pc, kCosC[0]);
    const double s = fma(r * z, ps, r);
    const double c = fma(z * z, pc, fma(-0.5, z, 1.0));
    // quadrant: q & 1 swaps, bit 1 of q / (q + 1) flips the signs (integer pipe)
    const bool swap = q & 1u;
    const double so = swap ? c : s;
    const double co = swap ? s : c;
    const uint64_t fs = (uint64_t)(q & 2u) << 62, fc = (uint64_t)((q + 1u) & 2u) << 62;
    *sn = f64_from_bits(f64_bits(so) ^ fs);
    *cs = f64_from_bits(f64_bits(co) ^ fc);
}

// sin and cos of a SMALL angle, |x| <= kSmallAngle: no range reduction, series cut after the terms
// that still matter at that size (next omitted terms: 2.3e-17 relative for sin, 1.7e-18 for cos).
// The sigma-point offsets of a converged filter are a few degrees at most.
constexpr double kSmallAngle = 0.125;

STE_DEV void small_sincos(double x, double *sn, double *cs) {
    const double z = x * x;
    double ps = fma(z, kSinC[3], kSinC[2]);
    double pc = fma(z, kCosC[4], kCosC[3]);
    ps = fma(z, ps, kSinC[1]);
    pc = fma(z, pc, kCosC[2]);
    ps = fma(z, ps, kSinC[0]);
    pc = fma(z, pc, kCosC[1]);
    pc = fma(z, pc, kCosC[0]);
    *sn = fma(x * z, ps, x);
    *cs = fma(z * z, pc, fma(-0.5, z, 1.0));
}

// ---- atan2 ------------------------------------------------------------------------------------- //
// One division: with mn = min(|y|,|x|), mx = max(|y|,|x|) the argument is reduced to
// q = mn/mx (mn <= tan(pi/8) mx) or q = (mn - mx)/(mn + mx) (then atan = pi/4 + atan q), |q| <= 0.4143.
// Branch-free for finite arguments; (0, 0) -> 0 (the sign conventions of atan2(+-0, -0) are not
// reproduced, the filter never needs them); non-finite input gives NaN.
// X_NONNEG: the caller guarantees x >= 0 (latitude from (up, horizontal)).
template <bool X_NONNEG>
STE_DEV double fast_atan2(double y, double x) {
    const double ay = fabs(y), ax = X_NONNEG ? x : fabs(x);
    const bool swap = ay > ax;
    const double mx = swap ? ay : ax, mn = swap ? ax : ay;
    const bool big = mn > kTanPiEighth * mx;
    const double num = big ? mn - mx : mn;
    double den = big ? mn + mx : mx;
    den = (f64_bits(mx) << 1) != 0 ? den : 1.0;          // (0, 0): 0 / 1
    const double q = fast_div(num, den);
    const double z = q * q, w = z * z;
    double s1 = fma(w, kAtanC[10], kAtanC[8]);
    double s2 = fma(w, kAtanC[9], kAtanC[7]);
    s1 = fma(w, s1, kAtanC[6]);
    s2 = fma(w, s2, kAtanC[5]);
    s1 = fma(w, s1, kAtanC[4]);
    s2 = fma(w, s2, kAtanC[3]);
    s1 = fma(w, s1, kAtanC[2]);
    s2 = fma(w, s2, kAtanC[1]);
    s1 = fma(w, s1, kAtanC[0]);
    const double p = fma(-q, fma(z, s1, w * s2), q);   // atan(q) = q - q (z s1 + w s2)
    // angle of (mx, mn) in [0, pi/4], then undo the reflections; r = off + sgn * p
    double off = big ? kPiQuarter : 0.0;
    double sg = 1.0;
    if (swap) { off = kPiHalf - off; sg = -sg; }
    if (!X_NONNEG) {
        if ((int64_t)f64_bits(x) < 0) { off = kPi - off; sg = -sg; }
    }
    const double r = fma(sg, p, off);
    return f64_from_bits(f64_bits(r) | (f64_bits(y) & 0x8000000000000000ull));   // copysign(r >= 0, y)
}

}  // namespace ste
