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

// A streaming load the compiler may not sink below a later branch (volatile asm): used where a
// bandwidth-bound loop must have ALL its loads in flight before the first dependent decision.
#if defined(__CUDA_ARCH__)
static __device__ __forceinline__ double ste_load_stream_pinned(const double *p) {
    double v;
    asm volatile("ld.global.cs.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
#define STE_LOAD_STREAM_PINNED(p) ste_load_stream_pinned(p)
#else
#define STE_LOAD_STREAM_PINNED(p) (*(p))
#endif

#define STE_PRAGMA_(x) _Pragma(#x)
#define STE_UNROLL(n) STE_PRAGMA_(unroll n)

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

// ---- lock-step ("vector in a thread") forms ------------------------------------------------------ //
// Every routine below evaluates N independent arguments stage by stage.  Written this way the N
// dependency chains sit next to each other in program order and ptxas issues them interleaved:
// a DFMA has ~9 cycles of latency and a 2-cycle issue interval, so one chain alone leaves the FP64
// pipe idle three quarters of the time (profiles/: the scalar forms ran chain after chain).
#define STE_LANES _Pragma("unroll") for (int l = 0; l < N; ++l)

// reciprocal, reciprocal square root, square root, division: MUFU seed (~20 bits) + Newton steps in
// the FP64 pipe; no denormal / special-case slow paths: callers guarantee finite, normal, non-zero
// arguments (or guard the result themselves).
// NEWTON = 2: full precision (2^-20 -> 2^-40 -> 2^-80).  NEWTON = 1: 2^-40, for callers whose own correction
// step or whose error budget makes the second step redundant (fast_div_v; first-order terms of a series).
template <int N, int NEWTON = 2>
STE_DEV void fast_rcp_v(const double (&x)[N], double (&y)[N]) {
    double e[N];
    STE_LANES y[l] = seed_rcp(x[l]);
#pragma unroll
    for (int it = 0; it < NEWTON; ++it) {
        STE_LANES e[l] = fma(-x[l], y[l], 1.0);
        STE_LANES y[l] = fma(y[l], e[l], y[l]);
    }
}

// y <- y (1 + e/2 + 3 e^2 / 8), e = 1 - x y^2 : cubic convergence, 2^-20 -> 2^-60
template <int N>
STE_DEV void fast_rsqrt_v(const double (&x)[N], double (&y)[N]) {
    double t[N], e[N], p[N];
    STE_LANES y[l] = seed_rsqrt(x[l]);
    STE_LANES t[l] = x[l] * y[l];
    STE_LANES e[l] = fma(-t[l], y[l], 1.0);
    STE_LANES p[l] = fma(0.375, e[l], 0.5);
    STE_LANES t[l] = y[l] * e[l];
    STE_LANES y[l] = fma(t[l], p[l], y[l]);
}

// sqrt for x >= 0 (x == 0 -> 0)
template <int N>
STE_DEV void fast_sqrt_v(const double (&x)[N], double (&s)[N]) {
    double y[N], r[N];
    fast_rsqrt_v<N>(x, y);
    STE_LANES s[l] = x[l] * y[l];
    STE_LANES r[l] = fma(-s[l], s[l], x[l]);   // residual of the rounded product
    STE_LANES y[l] = 0.5 * y[l];
    STE_LANES s[l] = fma(y[l], r[l], s[l]);
    STE_LANES s[l] = (x[l] > 0.0) ? s[l] : 0.0;
}

template <int N>
STE_DEV void fast_div_v(const double (&num)[N], const double (&den)[N], double (&q)[N]) {
    // r to 2^-40 is enough: q0 = num r is off by 2^-40 relative, the residual num - den q0 is exact (fma), and the
    // correction rem r carries the 2^-40 of r on a term that is itself 2^-40 of q: 2^-80 before the final rounding
    double r[N], rem[N];
    fast_rcp_v<N, 1>(den, r);
    STE_LANES q[l] = num[l] * r[l];
    STE_LANES rem[l] = fma(-den[l], q[l], num[l]);
    STE_LANES q[l] = fma(rem[l], r[l], q[l]);
}

STE_DEV double fast_rcp(double x) {
    const double xv[1] = {x};
    double y[1];
    fast_rcp_v<1>(xv, y);
    return y[0];
}
STE_DEV double fast_rsqrt(double x) {
    const double xv[1] = {x};
    double y[1];
    fast_rsqrt_v<1>(xv, y);
    return y[0];
}
STE_DEV double fast_sqrt(double x) {
    const double xv[1] = {x};
    double y[1];
    fast_sqrt_v<1>(xv, y);
    return y[0];
}
STE_DEV double fast_div(double num, double den) {
    const double n[1] = {num}, d[1] = {den};
    double q[1];
    fast_div_v<1>(n, d, q);
    return q[0];
}

// ---- sin and cos together --------------------------------------------------------------------- //
// Branch-free; valid for |x| <= kSinCosMaxArg (three-term Cody-Waite reduction).  Callers check
// the range once per filter step and take the library ("cold") path otherwise.
constexpr double kSinCosMaxArg = 105615.0;

template <int N>
STE_DEV void fast_sincos_v(const double (&x)[N], double (&sn)[N], double (&cs)[N]) {
    double t[N], n[N], r[N], z[N], ps[N], pc[N], s[N], c[N];
    uint32_t q[N];
    // n = rint(x * 2/pi) through the magic-number trick; the low word of t holds n (two's complement)
    STE_LANES t[l] = fma(x[l], kTrigK[3], kTrigK[4]);
    STE_LANES q[l] = (uint32_t)f64_bits(t[l]);
    STE_LANES n[l] = t[l] - kTrigK[4];
    STE_LANES r[l] = fma(-n[l], kTrigK[0], x[l]);
    STE_LANES r[l] = fma(-n[l], kTrigK[1], r[l]);
    STE_LANES r[l] = fma(-n[l], kTrigK[2], r[l]);
    STE_LANES z[l] = r[l] * r[l];
    // sin(r) = r + r z (S1 + z (S2 + ... )) ; cos(r) = 1 - z/2 + z^2 (C1 + z (C2 + ...))
    STE_LANES { ps[l] = fma(z[l], kSinC[5], kSinC[4]); pc[l] = fma(z[l], kCosC[5], kCosC[4]); }
    STE_LANES { ps[l] = fma(z[l], ps[l], kSinC[3]); pc[l] = fma(z[l], pc[l], kCosC[3]); }
    STE_LANES { ps[l] = fma(z[l], ps[l], kSinC[2]); pc[l] = fma(z[l], pc[l], kCosC[2]); }
    STE_LANES { ps[l] = fma(z[l], ps[l], kSinC[1]); pc[l] = fma(z[l], pc[l], kCosC[1]); }
    STE_LANES { ps[l] = fma(z[l], ps[l], kSinC[0]); pc[l] = fma(z[l], pc[l], kCosC[0]); }
    STE_LANES { s[l] = fma(r[l] * z[l], ps[l], r[l]); c[l] = fma(z[l] * z[l], pc[l], fma(-0.5, z[l], 1.0)); }
    // quadrant: q & 1 swaps, bit 1 of q / (q + 1) flips the signs (integer pipe)
    STE_LANES {
        const bool swap = q[l] & 1u;
        const double so = swap ? c[l] : s[l];
        const double co = swap ? s[l] : c[l];
        const uint64_t fs = (uint64_t)(q[l] & 2u) << 62, fc = (uint64_t)((q[l] + 1u) & 2u) << 62;
        sn[l] = f64_from_bits(f64_bits(so) ^ fs);
        cs[l] = f64_from_bits(f64_bits(co) ^ fc);
    }
}

// sin and cos of SMALL angles, |x| <= kSmallAngle: no range reduction, series cut after the terms
// that still matter at that size (next omitted terms: 2.3e-17 relative for sin, 1.7e-18 for cos).
// The sigma-point offsets of a converged filter are a few degrees at most.
constexpr double kSmallAngle = 0.125;

// STE_POLY_LITERALS: the coefficients of the short series as literals instead of constant-table reads (experiment:
// where the compiler keeps a table value in a vector register, the DFMA that uses it reads three register pairs and
// issues every 3 cycles instead of 2; a literal is materialised in a uniform register, which is not a register-file read)
#ifdef STE_POLY_LITERALS
#define STE_SINC(i) ((i) == 0 ? -1.66666666666666324348e-01 : (i) == 1 ? 8.33333333332248946124e-03 : (i) == 2 ? -1.98412698298579493134e-04 : 2.75573137070700676789e-06)
#define STE_COSC(i) ((i) == 0 ? 4.16666666666666019037e-02 : (i) == 1 ? -1.38888888888741095749e-03 : (i) == 2 ? 2.48015872894767294178e-05 : (i) == 3 ? -2.75573143513906633035e-07 : 2.08757232129817482790e-09)
#else
#define STE_SINC(i) kSinC[i]
#define STE_COSC(i) kCosC[i]
#endif
template <int N>
STE_DEV void small_sincos_v(const double (&x)[N], double (&sn)[N], double (&cs)[N]) {
    double z[N], ps[N], pc[N];
    STE_LANES z[l] = x[l] * x[l];
    // cosine: the z^6 term (coefficient 2.1e-9) is below 4.2e-20 for |x| <= 0.125 and is left out
    STE_LANES { ps[l] = fma(z[l], STE_SINC(3), STE_SINC(2)); pc[l] = fma(z[l], STE_COSC(3), STE_COSC(2)); }
    STE_LANES { ps[l] = fma(z[l], ps[l], STE_SINC(1)); pc[l] = fma(z[l], pc[l], STE_COSC(1)); }
    STE_LANES { ps[l] = fma(z[l], ps[l], STE_SINC(0)); pc[l] = fma(z[l], pc[l], STE_COSC(0)); }
    STE_LANES { sn[l] = fma(x[l] * z[l], ps[l], x[l]); cs[l] = fma(z[l] * z[l], pc[l], fma(-0.5, z[l], 1.0)); }
}

STE_DEV void fast_sincos(double x, double *sn, double *cs) {
    const double xv[1] = {x};
    double s[1], c[1];
    fast_sincos_v<1>(xv, s, c);
    *sn = s[0];
    *cs = c[0];
}

// ---- atan2 ------------------------------------------------------------------------------------- //
// One division: with mn = min(|y|,|x|), mx = max(|y|,|x|) the argument is reduced to
// q = mn/mx (mn <= tan(pi/8) mx) or q = (mn - mx)/(mn + mx) (then atan = pi/4 + atan q), |q| <= 0.4143.
// Branch-free for finite arguments; (0, 0) -> 0 (the sign conventions of atan2(+-0, -0) are not
// reproduced, the filter never needs them); non-finite input gives NaN.
// Lanes l >= UNIT_FROM promise x >= 0 and x^2 + y^2 = 1 (a latitude from (up, hypot(east, north))): no reflection for
// x < 0 and no (0, 0) guard there - same values, fewer selects.
template <int N, int UNIT_FROM = N>
STE_DEV void fast_atan2_v(const double (&y)[N], const double (&x)[N], double (&out)[N]) {
    double mx[N], mn[N], num[N], den[N], q[N], z[N], w[N], s1[N], s2[N], p[N], off[N], sg[N];
    bool swap[N], big[N];
    STE_LANES {
        const double ay = fabs(y[l]), ax = fabs(x[l]);
        swap[l] = ay > ax;
        mx[l] = swap[l] ? ay : ax;
        mn[l] = swap[l] ? ax : ay;
    }
    double bigf[N];
    STE_LANES {
        // second reduction, (mn - mx) / (mn + mx) above tan(pi/8), as multiply-adds by a 0 / 1 flag: the same values as
        // selecting between mn - mx and mn (a zero product adds exactly nothing), one select per lane instead of six
        big[l] = mn[l] > kTanPiEighth * mx[l];
        bigf[l] = big[l] ? 1.0 : 0.0;
        num[l] = fma(-bigf[l], mx[l], mn[l]);
        const double d = fma(bigf[l], mn[l], mx[l]);
        den[l] = (l >= UNIT_FROM || (f64_bits(mx[l]) << 1) != 0) ? d : 1.0;          // (0, 0): 0 / 1
    }
    fast_div_v<N>(num, den, q);
    STE_LANES z[l] = q[l] * q[l];
    STE_LANES w[l] = z[l] * z[l];
    STE_LANES { s1[l] = fma(w[l], kAtanC[10], kAtanC[8]); s2[l] = fma(w[l], kAtanC[9], kAtanC[7]); }
    STE_LANES { s1[l] = fma(w[l], s1[l], kAtanC[6]); s2[l] = fma(w[l], s2[l], kAtanC[5]); }
    STE_LANES { s1[l] = fma(w[l], s1[l], kAtanC[4]); s2[l] = fma(w[l], s2[l], kAtanC[3]); }
    STE_LANES { s1[l] = fma(w[l], s1[l], kAtanC[2]); s2[l] = fma(w[l], s2[l], kAtanC[1]); }
    STE_LANES s1[l] = fma(w[l], s1[l], kAtanC[0]);
    STE_LANES p[l] = fma(-q[l], fma(z[l], s1[l], w[l] * s2[l]), q[l]);   // atan(q) = q - q (z s1 + w s2)
    // angle of (mx, mn) in [0, pi/4], then undo the reflections: r = off + sg * p
    STE_LANES {
        off[l] = bigf[l] * kPiQuarter;
        sg[l] = 1.0;
        if (swap[l]) { off[l] = kPiHalf - off[l]; sg[l] = -1.0; }
        if (l < UNIT_FROM && (int64_t)f64_bits(x[l]) < 0) { off[l] = kPi - off[l]; sg[l] = -sg[l]; }
    }
    STE_LANES {
        const double r = fma(sg[l], p[l], off[l]);
        out[l] = f64_from_bits(f64_bits(r) | (f64_bits(y[l]) & 0x8000000000000000ull));   // copysign(r >= 0, y)
    }
}

// ---- small-argument inverse trigonometry: Maclaurin series, no reduction, no selects ----------- //
// Used by the small-displacement tier of the geodetic step (a ship moves a few km per predict):
//   atan q       for |q| <= 0.0645  (first neglected term q^14/15   < 2e-18 relative)
//   asin s       for |s| <= 2^-6    (first neglected term 0.022 s^10 < 2e-20 relative)
//   sqrt(1 + e)  for 0 <= e <= 0.0042 (first neglected term 0.016 e^7 < 1e-18)
// The callers establish the ranges once per filter step (step_is_small).
STE_CONST double kAtanS[6] = {-1.0 / 3.0, 1.0 / 5.0, -1.0 / 7.0, 1.0 / 9.0, -1.0 / 11.0, 1.0 / 13.0};
STE_CONST double kAsinS[4] = {1.0 / 6.0, 3.0 / 40.0, 15.0 / 336.0, 105.0 / 3456.0};
STE_CONST double kSqrt1pS[6] = {0.5, -0.125, 0.0625, -5.0 / 128.0, 7.0 / 256.0, -21.0 / 1024.0};
constexpr double kSmallAsinMax = 0.015625;      // 2^-6 rad: 99.5 km on the sphere
constexpr double kSmallLatMaxDeg = 75.0;        // with it |tan(dlon)| <= sin(2^-6) / cos(75.9 deg) / cos(dlon) = 0.0643

template <int N>
STE_DEV void small_atan_v(const double (&q)[N], const double (&z)[N], double (&out)[N]) {   // z = q*q
    double p[N];
    STE_LANES p[l] = fma(z[l], kAtanS[5], kAtanS[4]);
    STE_LANES p[l] = fma(z[l], p[l], kAtanS[3]);
    STE_LANES p[l] = fma(z[l], p[l], kAtanS[2]);
    STE_LANES p[l] = fma(z[l], p[l], kAtanS[1]);
    STE_LANES p[l] = fma(z[l], p[l], kAtanS[0]);
    STE_LANES out[l] = fma(q[l], z[l] * p[l], q[l]);
}
template <int N>
STE_DEV void small_asin_v(const double (&s)[N], double (&out)[N]) {
    double z[N], p[N];
    STE_LANES z[l] = s[l] * s[l];
    STE_LANES p[l] = fma(z[l], kAsinS[3], kAsinS[2]);
    STE_LANES p[l] = fma(z[l], p[l], kAsinS[1]);
    STE_LANES p[l] = fma(z[l], p[l], kAsinS[0]);
    STE_LANES out[l] = fma(s[l], z[l] * p[l], s[l]);
}
// a * sqrt(1 + e)
template <int N>
STE_DEV void small_hypot_scale_v(const double (&a)[N], const double (&e)[N], double (&out)[N]) {
    double p[N];
    STE_LANES p[l] = fma(e[l], kSqrt1pS[5], kSqrt1pS[4]);
    STE_LANES p[l] = fma(e[l], p[l], kSqrt1pS[3]);
    STE_LANES p[l] = fma(e[l], p[l], kSqrt1pS[2]);
    STE_LANES p[l] = fma(e[l], p[l], kSqrt1pS[1]);
    STE_LANES p[l] = fma(e[l], p[l], kSqrt1pS[0]);
    STE_LANES out[l] = fma(a[l], e[l] * p[l], a[l]);
}

// a * (sqrt(1 + e) - 1)
// (without the e^6 term: it is 0.041 e^5 <= 5.4e-14 of the excess, 1.1e-16 absolute at the tier's limit)
template <int N>
STE_DEV void small_hypot_excess_v(const double (&a)[N], const double (&e)[N], double (&out)[N]) {
    double p[N];
    STE_LANES p[l] = fma(e[l], kSqrt1pS[4], kSqrt1pS[3]);
    STE_LANES p[l] = fma(e[l], p[l], kSqrt1pS[2]);
    STE_LANES p[l] = fma(e[l], p[l], kSqrt1pS[1]);
    STE_LANES p[l] = fma(e[l], p[l], kSqrt1pS[0]);
    STE_LANES out[l] = a[l] * (e[l] * p[l]);
}

template <bool X_NONNEG>
STE_DEV double fast_atan2(double y, double x) {
    const double yv[1] = {y}, xv[1] = {x};
    double o[1];
    fast_atan2_v<1>(yv, xv, o);
    return o[0];
}

}  // namespace ste
