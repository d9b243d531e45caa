// ste_filter.cuh - per-track predict / update / smoother-step device functions (fp64, n = 4).
//
// One thread owns one track: the 4-vector mean and the 10 unique entries of the symmetric
// covariance live in registers across the whole time loop.  Citations are to
// /root/reference/src/track_estimators/kalman_filters/unscented.py unless stated.
#pragma once
#include "ste_math.cuh"
#include "../../include/ste_ukf.h"

namespace ste {

// Shared (per launch) matrices, read through the constant bank of the kernel parameters.
struct Model {
    const double *H;   // [16] row-major
    const double *Q;   // [16]
    const double *R;   // [16]
};

// Per-thread scratch slots outside the register file (shared memory on the device, a plain
// array in the host sandbox).  Slot k of this thread lives at base[k * stride].  It holds the
// square-root factor so that the sigma-point loop can stay ROLLED (one instance of the
// trigonometric code in the instruction cache) and index its columns at run time.
struct Scratch {
    double *base;
    int stride;
    STE_DEV double &at(int slot) const { return base[(long)slot * stride]; }
};
constexpr int kScratchRoot = 0;   // 16 slots: M[r][c] at kScratchRoot + c * 4 + r (column-major)
constexpr int kScratchSlots = 16;

STE_DEV void stash_root(const Scratch &sc, const double (&M)[10]) {
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int r = 0; r < 4; ++r) sc.at(kScratchRoot + c * 4 + r) = M[SYM(r, c)];
}

// The rolled sigma-point loop visits the points in the order x, x+m_0, x-m_0, x+m_1, ... so that
// the sincos of an offset column is computed once and reused by its mirror point.
//   j = 0: centre;  j = 2c+1: x + m_c;  j = 2c+2: x - m_c
// sigma_ref_index maps j to the reference's column index (unscented.py:100-105: X[:, 1+c] = x + m_c,
// X[:, 5+c] = x - m_c).
STE_DEV int sigma_ref_index(int j) { return j == 0 ? 0 : ((j & 1) ? 1 + (j >> 1) : 4 + (j >> 1)); }

// Propagate sigma point j of (x, root in scratch).  `base` / `off` persist across iterations:
// base = trig of the centre angles, off = trig of the current offset column.
// o receives X_j - x, y the propagated point.
STE_DEV void propagate_sigma(const Scratch &sc, int j, const double (&x)[4], double dt, double dtR, double sog_rate,
                             double cog_rate, AngleTrig &base, AngleTrig &off, double (&o)[4], double (&xi)[4],
                             double (&y)[4]) {
    const bool plus = (j & 1) || j == 0;
    const int c = (j - 1) >> 1;
    double m[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) m[r] = (j == 0) ? x[r] : sc.at(kScratchRoot + c * 4 + r);
    if (plus) off = angle_trig(m[1], m[3], m[2], dtR);   // centre angles at j = 0, offset column otherwise
    // j = 0: base is still the zero angle (sin 0, cos 1), so base (+) off == off exactly and becomes
    // the centre trig; afterwards base stays fixed and off cycles through the columns.
    const AngleTrig cur = angle_add(base, off, plus);
    if (j == 0) base = cur;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        o[r] = (j == 0) ? 0.0 : (plus ? m[r] : -m[r]);
        xi[r] = x[r] + o[r];
    }
    geodetic_finish(xi, cur, dt, sog_rate, cog_rate, y);
}

// ------------------------------------------------------------------------------------------ //
// predict (:144-207).  Unscented transform of (x, P) through the geodetic model.
//
// Moments are accumulated in one pass about the propagated centre point c = f(x):
//   d_i = f(X_i) - c,  mu = Wi * sum d_i,   mean = c + mu            (W0 + 8 Wi = 1, d_0 = 0)
//   P'  = sum_i W_i (f(X_i) - mean - e)(...)^T + Q = Wi * sum d_i d_i^T - mu mu^T + e e^T + Q
// with e the additive noise the reference adds to the mean BEFORE forming deviations (:198-205).
// This needs no storage for the 9 propagated points.
// ------------------------------------------------------------------------------------------ //
STE_DEV void ukf_predict(double (&x)[4], double (&P)[10], const double *Q, double dt,
                         double sog_rate, double cog_rate, const double (&e)[4],
                         int &status, const Scratch &sc, double *sig_prior, double *sig_post, int64_t ld) {
    {
        double M[10];
        if (sqrt_psd4(P, kSigmaScale, M)) status |= STE_STATUS_INDEFINITE;
        stash_root(sc, M);
    }
    const double dtR = dt / kEarthRadiusKm;
    double c[4] = {0.0, 0.0, 0.0, 0.0};
    double s1[4] = {0.0, 0.0, 0.0, 0.0};
    double s2[10];
#pragma unroll
    for (int k = 0; k < 10; ++k) s2[k] = 0.0;
    AngleTrig base = {0.0, 1.0, 0.0, 1.0, 0.0, 1.0}, off = base;
#pragma unroll 1
    for (int j = 0; j < 9; ++j) {
        double o[4], xi[4], yi[4];
        propagate_sigma(sc, j, x, dt, dtR, sog_rate, cog_rate, base, off, o, xi, yi);
        if (sig_prior) {
            const int jr = sigma_ref_index(j);
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                sig_prior[(r * 9 + jr) * ld] = xi[r];
                sig_post[(r * 9 + jr) * ld] = yi[r];
            }
        }
        if (j == 0) {
#pragma unroll
            for (int r = 0; r < 4; ++r) c[r] = yi[r];
        }
        double d[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            d[r] = yi[r] - c[r];
            s1[r] += d[r];
        }
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int q = r; q < 4; ++q) s2[SYM(r, q)] = fma(d[r], d[q], s2[SYM(r, q)]);
    }
    double mu[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        mu[r] = kWi * s1[r];
        x[r] = c[r] + (mu[r] + e[r]);
    }
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int q = r; q < 4; ++q)
            P[SYM(r, q)] = fma(kWi, s2[SYM(r, q)], fma(-mu[r], mu[q], fma(e[r], e[q], Q[r * 4 + q])));
}

// ------------------------------------------------------------------------------------------ //
// Robustification (:353-511, dead code in the reference, enabled for BASELINE config 4).
// gamma = |y^T S^+ y| with y = z - x (NOT z - Hx, no angle wrap, :420-426); while gamma > chi:
// lambda += (gamma - chi) / (y^T S^+ R S^+ y); R *= lambda (compounding); zero noise.
// `rs` is the accumulated scale of R.  GENERIC: any H, R.
// ------------------------------------------------------------------------------------------ //
STE_DEV void innovation_cov(const double (&P)[10], const double *H, const double *R,
                                               double rs, double (&HP)[16], double (&S)[10]) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            double acc = 0.0;
#pragma unroll
            for (int k = 0; k < 4; ++k) acc = fma(H[i * 4 + k], P[SYM(k, j)], acc);
            HP[i * 4 + j] = acc;
        }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = i; j < 4; ++j) {
            double acc = rs * R[i * 4 + j];
#pragma unroll
            for (int k = 0; k < 4; ++k) acc = fma(HP[i * 4 + k], H[j * 4 + k], acc);
            S[SYM(i, j)] = acc;
        }
}

STE_DEV double quad_form(const double (&A)[10], const double (&y)[4]) {
    double acc = 0.0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        double row = 0.0;
#pragma unroll
        for (int j = 0; j < 4; ++j) row = fma(A[SYM(i, j)], y[j], row);
        acc = fma(y[i], row, acc);
    }
    return acc;
}

// update (:209-265), generic H / R.  z is consumed (noise added in place as the reference does).
template <bool GATING>
STE_DEV void ukf_update_generic(double (&x)[4], double (&P)[10], const Model &m,
                                                   double (&z)[4], const double *unit_noise /*[4] or null*/,
                                                   double chi, int max_iter, int &status, int &gate_it,
                                                   double &gate_lam, double &rs) {
    const double *H = m.H, *R = m.R;
    double HP[16], S[10], Sinv[10];
    rs = 1.0;
    innovation_cov(P, H, R, rs, HP, S);
    pinv_sym4(S, Sinv);
    gate_it = 0;
    gate_lam = 1.0;
    if (GATING) {
        double yg[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) yg[i] = z[i] - x[i];
        double gamma = fabs(quad_form(Sinv, yg));
        while (gamma > chi) {
            if (gate_it >= max_iter) {
                status |= STE_STATUS_GATE_CAP;
                break;
            }
            // v = S^+ y ; den = v^T (rs R) v
            double v[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                double acc = 0.0;
#pragma unroll
                for (int j = 0; j < 4; ++j) acc = fma(Sinv[SYM(i, j)], yg[j], acc);
                v[i] = acc;
            }
            double den = 0.0;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                double row = 0.0;
#pragma unroll
                for (int j = 0; j < 4; ++j) row = fma(R[i * 4 + j], v[j], row);
                den = fma(v[i], row, den);
            }
            den *= rs;
            gate_lam = gate_lam + (gamma - chi) / den;
            rs *= gate_lam;
            innovation_cov(P, H, R, rs, HP, S);
            pinv_sym4(S, Sinv);
            gamma = fabs(quad_form(Sinv, yg));
            ++gate_it;
        }
    }
    if (unit_noise) {
#pragma unroll
        for (int i = 0; i < 4; ++i) z[i] = fma(unit_noise[i], sqrt(rs * R[i * 4 + i]), z[i]);
    }
    // K = P H^T S^+  with P H^T = (H P)^T
    double K[16];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            double acc = 0.0;
#pragma unroll
            for (int k = 0; k < 4; ++k) acc = fma(HP[k * 4 + i], Sinv[SYM(k, j)], acc);
            K[i * 4 + j] = acc;
        }
    double y[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        double acc = z[i];
#pragma unroll
        for (int k = 0; k < 4; ++k) acc = fma(-H[i * 4 + k], x[k], acc);
        y[i] = acc;
    }
    y[3] = wrap180(y[3]);  // :250
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        double acc = x[i];
#pragma unroll
        for (int k = 0; k < 4; ++k) acc = fma(K[i * 4 + k], y[k], acc);
        x[i] = acc;
    }
    x[3] = py_mod360(x[3]);  // :257
    // Joseph form (:260-265): (I - K H) P (I - K H)^T + K R K^T
    double A[16], AP[16], KR[16];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            double acc = (i == j) ? 1.0 : 0.0;
#pragma unroll
            for (int k = 0; k < 4; ++k) acc = fma(-K[i * 4 + k], H[k * 4 + j], acc);
            A[i * 4 + j] = acc;
        }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            double acc = 0.0, acc2 = 0.0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                acc = fma(A[i * 4 + k], P[SYM(k, j)], acc);
                acc2 = fma(K[i * 4 + k], R[k * 4 + j], acc2);
            }
            AP[i * 4 + j] = acc;
            KR[i * 4 + j] = acc2 * rs;
        }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = i; j < 4; ++j) {
            double acc = 0.0;
#pragma unroll
            for (int k = 0; k < 4; ++k) acc = fma(AP[i * 4 + k], A[j * 4 + k], fma(KR[i * 4 + k], K[j * 4 + k], acc));
            P[SYM(i, j)] = acc;
        }
}

// update for the position-only measurement H = diag(1,1,0,0) with R zero outside its leading
// 2x2 block (every example and the CLI input.json of the reference).  S = H P H^T + R is then
// block-diagonal with an exactly-zero trailing block, pinv(S) = [pinv(S22) 0; 0 0], and K has two
// non-zero columns; the arithmetic below is the generic path with the zeros removed.
template <bool GATING>
STE_DEV void ukf_update_position(double (&x)[4], double (&P)[10], const Model &m,
                                                    double (&z)[4], const double *unit_noise, double chi,
                                                    int max_iter, int &status, int &gate_it,
                                                    double &gate_lam, double &rs) {
    const double r00 = m.R[0], r01 = m.R[1], r11 = m.R[5];
    rs = 1.0;
    double Si[3];
    pinv_sym2(fma(rs, r00, P[SYM(0, 0)]), fma(rs, r01, P[SYM(0, 1)]), fma(rs, r11, P[SYM(1, 1)]), Si);
    gate_it = 0;
    gate_lam = 1.0;
    if (GATING) {
        const double y0 = z[0] - x[0], y1 = z[1] - x[1];
        double gamma = fabs(fma(y0, fma(Si[0], y0, Si[1] * y1), y1 * fma(Si[1], y0, Si[2] * y1)));
        while (gamma > chi) {
            if (gate_it >= max_iter) {
                status |= STE_STATUS_GATE_CAP;
                break;
            }
            const double v0 = fma(Si[0], y0, Si[1] * y1), v1 = fma(Si[1], y0, Si[2] * y1);
            const double den = rs * fma(v0, fma(r00, v0, r01 * v1), v1 * fma(r01, v0, r11 * v1));
            gate_lam = gate_lam + (gamma - chi) / den;
            rs *= gate_lam;
            pinv_sym2(fma(rs, r00, P[SYM(0, 0)]), fma(rs, r01, P[SYM(0, 1)]), fma(rs, r11, P[SYM(1, 1)]), Si);
            gamma = fabs(fma(y0, fma(Si[0], y0, Si[1] * y1), y1 * fma(Si[1], y0, Si[2] * y1)));
            ++gate_it;
        }
    }
    if (unit_noise) {
        z[0] = fma(unit_noise[0], sqrt(rs * r00), z[0]);
        z[1] = fma(unit_noise[1], sqrt(rs * r11), z[1]);
    }
    double K0[4], K1[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        K0[r] = fma(P[SYM(r, 0)], Si[0], P[SYM(r, 1)] * Si[1]);
        K1[r] = fma(P[SYM(r, 0)], Si[1], P[SYM(r, 1)] * Si[2]);
    }
    const double y0 = z[0] - x[0], y1 = z[1] - x[1];
#pragma unroll
    for (int r = 0; r < 4; ++r) x[r] = fma(K0[r], y0, fma(K1[r], y1, x[r]));
    x[3] = py_mod360(x[3]);
    // AP = P - K P[0:2, :]
    double AP[16];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
            AP[i * 4 + j] = fma(-K0[i], P[SYM(0, j)], fma(-K1[i], P[SYM(1, j)], P[SYM(i, j)]));
    // K (rs R22) K^T
    double KR0[4], KR1[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        KR0[r] = rs * fma(K0[r], r00, K1[r] * r01);
        KR1[r] = rs * fma(K0[r], r01, K1[r] * r11);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = i; j < 4; ++j) {
            double acc = fma(-AP[i * 4 + 0], K0[j], fma(-AP[i * 4 + 1], K1[j], AP[i * 4 + j]));
            P[SYM(i, j)] = fma(KR0[i], K0[j], fma(KR1[i], K1[j], acc));
        }
}

// ------------------------------------------------------------------------------------------ //
// One URTSS backward iteration (rts_step, :297-349), in two phases so that the long sigma-point
// loop runs with few live registers:
//
//  phase 1  urtss_moments   sigma points of the filtered state (xf, Pf) -> weighted sums
//      x_b - xf = sum W_i d_i,            d_i = f(X_i) - xf
//      P_b      = sum W_i d_i d_i^T + Q   (about the FILTERED mean: reference quirk, :324-325)
//      Delta_c  = d(x + m_c) - d(x - m_c) (c = 0..3, parked in scratch)
//  phase 2  urtss_gain      cross covariance, gain and the smoothed state
//      D = sum W_i (X_i - xf)(f(X_i) - x_b)^T = Wi sum_c m_c Delta_c^T
//          (X_0 - xf = 0 and sum W_i (X_i - xf) = 0, so neither x_b nor the centre point enters)
//      K = D pinv(P_b) = Wi sum_c m_c (pinv(P_b) Delta_c)^T
//      xs = xf + K wrap(xs - x_b);   Ps = Pf + K (Ps - P_b) K^T
//
// The smoothed state of step+1 (xs, Ps) is only touched in phase 2 and lives in scratch between
// steps; Pf is re-read by the caller for phase 2 instead of being held across phase 1.
// ------------------------------------------------------------------------------------------ //
constexpr int kScratchXs = 16;      // 4 slots
constexpr int kScratchPs = 20;      // 10 slots
constexpr int kScratchDelta = 30;   // 16 slots: Delta_c[r] at kScratchDelta + c * 4 + r
constexpr int kScratchSlotsBwd = 46;

STE_DEV void urtss_moments(const double (&xf)[4], const double (&Pf)[10], const double *Q, double dt,
                           double sog_rate, double cog_rate, double (&s1)[4], double (&Pb)[10], int &status,
                           const Scratch &sc) {
    {
        double M[10];
        if (sqrt_psd4(Pf, kSigmaScale, M)) status |= STE_STATUS_INDEFINITE;
        stash_root(sc, M);
    }
    const double dtR = dt / kEarthRadiusKm;
#pragma unroll
    for (int r = 0; r < 4; ++r) s1[r] = 0.0;
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int q = r; q < 4; ++q) Pb[SYM(r, q)] = Q[r * 4 + q];
    AngleTrig base = {0.0, 1.0, 0.0, 1.0, 0.0, 1.0}, off = base;
    double dplus[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll 1
    for (int j = 0; j < 9; ++j) {
        double o[4], xi[4], d[4];
        propagate_sigma(sc, j, xf, dt, dtR, sog_rate, cog_rate, base, off, o, xi, d);
        const double w = (j == 0) ? kW0 : kWi;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            d[r] -= xf[r];
            s1[r] = fma(w, d[r], s1[r]);
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const double wd = w * d[r];
#pragma unroll
            for (int q = r; q < 4; ++q) Pb[SYM(r, q)] = fma(wd, d[q], Pb[SYM(r, q)]);
        }
        if (j & 1) {
#pragma unroll
            for (int r = 0; r < 4; ++r) dplus[r] = d[r];
        } else if (j > 0) {
            const int c = (j - 1) >> 1;
#pragma unroll
            for (int r = 0; r < 4; ++r) sc.at(kScratchDelta + c * 4 + r) = dplus[r] - d[r];
        }
    }
}

STE_DEV void urtss_gain(const double (&xf)[4], const double (&Pf)[10], const double (&s1)[4], const double (&Pb)[10],
                        const double (&e)[4], double (&xs)[4], double (&Ps)[10], int &status, const Scratch &sc) {
    double Pbinv[10];
    if (pinv_spd4(Pb, Pbinv) > 0) status |= STE_STATUS_RANK_DEFICIENT;
    double K[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) K[k] = 0.0;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        double dl[4], v[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) dl[r] = sc.at(kScratchDelta + c * 4 + r);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            double acc = 0.0;
#pragma unroll
            for (int k = 0; k < 4; ++k) acc = fma(Pbinv[SYM(j, k)], dl[k], acc);
            v[j] = acc;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const double wm = kWi * sc.at(kScratchRoot + c * 4 + i);
#pragma unroll
            for (int j = 0; j < 4; ++j) K[i * 4 + j] = fma(wm, v[j], K[i * 4 + j]);
        }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) xs[r] = sc.at(kScratchXs + r);
#pragma unroll
    for (int k = 0; k < 10; ++k) Ps[k] = sc.at(kScratchPs + k);
    double y[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) y[r] = xs[r] - (xf[r] + (s1[r] + e[r]));
    y[3] = wrap180(y[3]);  // :340
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        double acc = xf[i];
#pragma unroll
        for (int k = 0; k < 4; ++k) acc = fma(K[i * 4 + k], y[k], acc);
        xs[i] = acc;
    }
    xs[3] = py_mod360(xs[3]);  // :346
    // Ps <- Pf + K (Ps - Pb) K^T
    double KG[16];
#pragma unroll
    for (int k = 0; k < 10; ++k) Ps[k] -= Pb[k];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            double acc = 0.0;
#pragma unroll
            for (int k = 0; k < 4; ++k) acc = fma(K[i * 4 + k], Ps[SYM(k, j)], acc);
            KG[i * 4 + j] = acc;
        }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = i; j < 4; ++j) {
            double acc = Pf[SYM(i, j)];
#pragma unroll
            for (int k = 0; k < 4; ++k) acc = fma(KG[i * 4 + k], K[j * 4 + k], acc);
            Ps[SYM(i, j)] = acc;
        }
#pragma unroll
    for (int r = 0; r < 4; ++r) sc.at(kScratchXs + r) = xs[r];
#pragma unroll
    for (int k = 0; k < 10; ++k) sc.at(kScratchPs + k) = Ps[k];
}

}  // namespace ste
