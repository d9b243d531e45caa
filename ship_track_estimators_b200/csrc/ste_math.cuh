// ste_math.cuh - register-resident fp64 building blocks of the UKF/URTSS kernels (sm_100a).
//
// Everything here is per-thread scalar code on fully unrolled fixed-size arrays, so that the
// compiler keeps all matrices in registers (no local memory).  Symmetric 4x4 matrices are held
// as their 10 upper-triangle entries in the order
//     [00 01 02 03 11 12 13 22 23 33]            (SYM(i,j), i <= j)
//
// Reference arithmetic being restated (file:line under /root/reference/src/track_estimators/):
//   sqrtm(3P)            kalman_filters/unscented.py:95-97   -> sqrt_psd4   (principal root)
//   np.linalg.pinv(S)    kalman_filters/unscented.py:243,333 -> pinv_sym4   (rcond 1e-15)
//   geodetic_dynamics    kalman_filters/non_linear_process.py:54-78 -> geodetic_step
//   x % 360, wrap        kalman_filters/unscented.py:250,257,340,346 -> py_mod360, wrap180
#pragma once
// A plain C++ build of these headers exists only for the developer-side numerical sandbox
// (tools/host_emul); the shipped library is always built by nvcc for sm_100a.
#include "ste_fastmath.cuh"

namespace ste {

constexpr double kEarthRadiusKm = 6378.137;           // constants.py:1
constexpr double kDegToRad = 0.017453292519943295;    // numpy radians(): x * (pi / 180)
constexpr double kRadToDeg = 57.29577951308232;       // numpy degrees(): x * (180 / pi)
constexpr double kW0 = 1.0 - 4.0 / 3.0;               // unscented.py:125  (n = 4)
constexpr double kWi = (1.0 - kW0) / 8.0;             // unscented.py:132
constexpr double kSigmaScale = 4.0 / (1.0 - kW0);     // unscented.py:95   (= 3)
constexpr double kPinvRcond = 1e-15;                  // numpy.linalg.pinv default rcond

STE_HD constexpr int SYM(int i, int j) {
    return i <= j ? (i * 4 - (i * (i - 1)) / 2 + (j - i)) : (j * 4 - (j * (j - 1)) / 2 + (i - j));
}

// Python / numpy floored modulo by 360 (result in [0, 360], sign of the divisor).
STE_DEV double py_mod360(double a) {
    double r = fmod(a, 360.0);
    if (r != 0.0) {
        if (r < 0.0) r += 360.0;
    } else {
        r = 0.0;  // copysign(0, 360)
    }
    return r;
}
STE_DEV double wrap180(double a) { return py_mod360(a + 180.0) - 180.0; }

// ------------------------------------------------------------------------------------------ //
// Cyclic Jacobi eigen-decomposition of a symmetric 4x4:  A = V diag(w) V^T.
// On return a[SYM(i,i)] hold the eigenvalues, V (row-major 4x4) the eigenvectors as columns.
// Rotations on an exactly-zero off-diagonal are the identity, so structurally block-diagonal
// inputs (S = H P H^T + R with zero rows/columns) keep their exact zeros and unit eigenvectors.
// ------------------------------------------------------------------------------------------ //
template <int P_, int Q_>
STE_DEV void jacobi_rotate(double (&a)[10], double (&V)[16]) {
    const double apq = a[SYM(P_, Q_)];
    const double app = a[SYM(P_, P_)];
    const double aqq = a[SYM(Q_, Q_)];
    // t = tan(rotation angle), the smaller root of t^2 + 2 theta t - 1 = 0, theta = (aqq-app)/(2apq)
    const double d = aqq - app;
    const double b = apq + apq;
    const double h = fast_sqrt(fma(d, d, b * b));
    double t = fast_div(d >= 0.0 ? b : -b, fabs(d) + h);   // sign(d) * b / (|d| + sqrt(d^2 + b^2))
    t = (apq == 0.0) ? 0.0 : t;                             // also covers d == b == 0 (0/0 -> NaN)
    const double c = fast_rsqrt(fma(t, t, 1.0));
    const double s = t * c;
    a[SYM(P_, P_)] = fma(-t, apq, app);
    a[SYM(Q_, Q_)] = fma(t, apq, aqq);
    a[SYM(P_, Q_)] = 0.0;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        if (r != P_ && r != Q_) {
            const double arp = a[SYM(r, P_)], arq = a[SYM(r, Q_)];
            a[SYM(r, P_)] = fma(c, arp, -s * arq);
            a[SYM(r, Q_)] = fma(s, arp, c * arq);
        }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const double vrp = V[r * 4 + P_], vrq = V[r * 4 + Q_];
        V[r * 4 + P_] = fma(c, vrp, -s * vrq);
        V[r * 4 + Q_] = fma(s, vrp, c * vrq);
    }
}

STE_DEV void jacobi_eig4(double (&a)[10], double (&V)[16]) {
#pragma unroll
    for (int i = 0; i < 16; ++i) V[i] = (i % 5 == 0) ? 1.0 : 0.0;
    // Quadratic convergence: 4x4 matrices reach off ~ 1e-17 |A| in 4-6 sweeps.
    // A pair (p,q) is converged when a_pq is negligible against sqrt(a_pp a_qq) (keeps the small
    // eigen-directions of badly scaled covariances accurate) or against the whole matrix.
    for (int sweep = 0; sweep < 12; ++sweep) {
        const double dia = a[0] * a[0] + a[4] * a[4] + a[7] * a[7] + a[9] * a[9];
        const double floor2 = 1e-40 * dia;
        bool more = false;
#pragma unroll
        for (int p = 0; p < 3; ++p)
#pragma unroll
            for (int q = p + 1; q < 4; ++q) {
                const double o2 = a[SYM(p, q)] * a[SYM(p, q)];
                more |= (o2 > 1e-33 * fabs(a[SYM(p, p)] * a[SYM(q, q)])) && (o2 > floor2);
            }
        if (!more) break;   // also leaves on NaN
        jacobi_rotate<0, 1>(a, V);
        jacobi_rotate<2, 3>(a, V);
        jacobi_rotate<0, 2>(a, V);
        jacobi_rotate<1, 3>(a, V);
        jacobi_rotate<0, 3>(a, V);
        jacobi_rotate<1, 2>(a, V);
    }
}

// out = V diag(f) V^T (symmetric, 10 entries)
STE_DEV void sym_from_eig(const double (&V)[16], const double (&f)[4], double (&out)[10]) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = i; j < 4; ++j) {
            double acc = 0.0;
#pragma unroll
            for (int k = 0; k < 4; ++k) acc = fma(V[i * 4 + k] * f[k], V[j * 4 + k], acc);
            out[SYM(i, j)] = acc;
        }
}

// M = Re sqrtm(scale * A): principal square root with negative eigenvalues contributing 0
// (scipy returns a complex root there and numpy's float assignment drops the imaginary part,
// unscented.py:104-105).  Returns true if a negative eigenvalue was clamped.
STE_DEV bool sqrt_psd4(const double (&A)[10], double scale, double (&M)[10]) {
    double a[10], V[16], f[4];
#pragma unroll
    for (int i = 0; i < 10; ++i) a[i] = A[i] * scale;
    jacobi_eig4(a, V);
    const double w[4] = {a[SYM(0, 0)], a[SYM(1, 1)], a[SYM(2, 2)], a[SYM(3, 3)]};
    const double wmax = fmax(fmax(fabs(w[0]), fabs(w[1])), fmax(fabs(w[2]), fabs(w[3])));
    bool clamped = false;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        clamped |= (w[k] < -1e-13 * wmax);
        f[k] = fast_sqrt(fmax(w[k], 0.0));
    }
    sym_from_eig(V, f, M);
    return clamped;
}

// Moore-Penrose inverse of a symmetric 4x4 with numpy's cutoff: singular values (|eigenvalues|)
// <= rcond * max are treated as zero.  Returns the number of dropped eigenvalues.
STE_DEV int pinv_sym4(const double (&A)[10], double (&Ainv)[10]) {
    double a[10], V[16], f[4];
#pragma unroll
    for (int i = 0; i < 10; ++i) a[i] = A[i];
    jacobi_eig4(a, V);
    const double w[4] = {a[SYM(0, 0)], a[SYM(1, 1)], a[SYM(2, 2)], a[SYM(3, 3)]};
    const double wmax = fmax(fmax(fabs(w[0]), fabs(w[1])), fmax(fabs(w[2]), fabs(w[3])));
    const double cut = kPinvRcond * wmax;
    int dropped = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const bool keep = fabs(w[k]) > cut;
        f[k] = keep ? fast_rcp(w[k]) : 0.0;
        dropped += keep ? 0 : 1;
    }
    sym_from_eig(V, f, Ainv);
    return dropped;
}

// Pseudo-inverse of a symmetric 2x2 [a b; b c] with the same cutoff rule (one Jacobi rotation
// diagonalises a 2x2 exactly).  `floor_sv` is an extra singular value taking part in the
// max (0 when the 2x2 is the only non-zero block of the full matrix).
STE_DEV int pinv_sym2(double a, double b, double c, double (&inv)[3]) {
    const double d = c - a;
    const double bb = b + b;
    const double h = fast_sqrt(fma(d, d, bb * bb));
    double t = fast_div(d >= 0.0 ? bb : -bb, fabs(d) + h);
    t = (b == 0.0) ? 0.0 : t;
    const double cs = fast_rsqrt(fma(t, t, 1.0));
    const double sn = t * cs;
    const double w0 = fma(-t, b, a), w1 = fma(t, b, c);
    const double cut = kPinvRcond * fmax(fabs(w0), fabs(w1));
    const bool k0 = fabs(w0) > cut, k1 = fabs(w1) > cut;
    const double f0 = k0 ? fast_rcp(w0) : 0.0, f1 = k1 ? fast_rcp(w1) : 0.0;
    // eigenvectors: v0 = (c, -s), v1 = (s, c)
    inv[0] = fma(cs * f0, cs, sn * f1 * sn);
    inv[1] = fma(-cs * f0, sn, sn * f1 * cs);
    inv[2] = fma(sn * f0, sn, cs * f1 * cs);
    return (k0 ? 0 : 1) + (k1 ? 0 : 1);
}

// ------------------------------------------------------------------------------------------ //
// Process model: great-circle propagation on the sphere (non_linear_process.py:54-78).
// dtR = dt / R_earth is hoisted by the caller (shared by the 9 sigma points of a step).
//
// The reference evaluates lat' = asin(s), s = sin(phi) cos(delta) + cos(phi) sin(delta) cos(alpha).
// (east, north, s) is the unit vector of the new position in the frame of the old meridian, so
// cos(lat') = hypot(east, north) and lat' = atan2(s, hypot(east, north)): the same angle, well
// conditioned up to the poles, and it reuses the one-division atan2 (no separate asin code).
// ------------------------------------------------------------------------------------------ //
STE_DEV void geodetic_step(const double (&x)[4], double dt, double dtR, double sog_rate,
                           double cog_rate, double (&y)[4]) {
    const double lam = x[0] * kDegToRad;
    const double phi = x[1] * kDegToRad;
    const double u = x[2];
    const double alpha = x[3] * kDegToRad;
    double sphi, cphi, sal, cal, sd, cd;
    fast_sincos(phi, &sphi, &cphi);
    fast_sincos(alpha, &sal, &cal);
    fast_sincos(u * dtR, &sd, &cd);
    const double east = sd * sal;
    const double sdca = sd * cal;
    const double north = fma(cphi, cd, -sphi * sdca);
    const double up = fma(sphi, cd, cphi * sdca);
    const double horiz = fast_sqrt(fma(east, east, north * north));
    y[0] = (lam + fast_atan2(east, north)) * kRadToDeg;
    y[1] = fast_atan2(up, horiz) * kRadToDeg;
    y[2] = fma(sog_rate, dt, u);
    y[3] = fma(cog_rate, dt, alpha * kRadToDeg);
}

}  // namespace ste
