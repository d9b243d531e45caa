// Developer-side numerical sandbox: compiles the per-track device code (csrc/ste_*.cuh) for the
// HOST with g++ so that algorithmic changes can be checked against the golden fixtures without
// a GPU.  NOT part of the product: the package never loads this library, tests and bench never
// use it, and its results are not bit-identical to the GPU (different libm).
#include <cstdint>
#include "../../ship_track_estimators_b200/csrc/ste_tracks.cuh"

using namespace ste;

extern "C" int emul_forward(const SteProblem *prob, const SteInputs *in, SteOutputs *out) {
    KernelArgs a;
    a.prob = *prob; a.in = *in; a.out = *out;
    bool pos = !(prob->flags & STE_FLAG_FORCE_GENERIC);
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            const double h = (i == j && i < 2) ? 1.0 : 0.0;
            if (prob->H[i * 4 + j] != h) pos = false;
            if ((i >= 2 || j >= 2) && prob->R[i * 4 + j] != 0.0) pos = false;
        }
    const bool gating = prob->flags & STE_FLAG_GATING;
    double scratch[kScratchSlotsFwd];
    const Scratch sc{scratch, 1};
    for (int t = 0; t < prob->n_tracks; ++t) {
        if (pos) { if (gating) forward_track<true, true>(a, t, sc); else forward_track<true, false>(a, t, sc); }
        else     { if (gating) forward_track<false, true>(a, t, sc); else forward_track<false, false>(a, t, sc); }
    }
    return 0;
}

extern "C" int emul_backward(const SteProblem *prob, const SteInputs *in, SteOutputs *out) {
    KernelArgs a;
    a.prob = *prob; a.in = *in; a.out = *out;
    double scratch[kScratchSlotsBwd];
    const Scratch sc{scratch, 1};
    for (int t = 0; t < prob->n_tracks; ++t) backward_track(a, t, sc);
    return 0;
}

// geodetic step of n states through the generic branch-free tier and the small-displacement tier
extern "C" void emul_geodetic_tiers(int n, const double *x, const double *dt, const double *sr, const double *cr,
                                    double *y_generic, double *y_small) {
    for (int i = 0; i < n; ++i) {
        double xi[4], a[4], b[4];
        for (int r = 0; r < 4; ++r) xi[r] = x[i * 4 + r];
        const double dtR = dt[i] / kEarthRadiusKm;
        const AngleTrig t = angle_trig<false>(xi[1], xi[3], xi[2], dtR);
        geodetic_finish<false, false>(xi, t, dt[i], sr[i], cr[i], a);
        geodetic_finish<false, true>(xi, t, dt[i], sr[i], cr[i], b);
        for (int r = 0; r < 4; ++r) { y_generic[i * 4 + r] = a[r]; y_small[i * 4 + r] = b[r]; }
    }
}

#if defined(STE_EMUL_STATS)
extern "C" void emul_sweep_log(unsigned char *buf, long long cap) { ste::ste_emul_sweep_log = buf; ste::ste_emul_sweep_log_cap = cap; ste::ste_emul_sweep_log_n = 0; }
extern "C" long long emul_sweep_log_count() { return ste::ste_emul_sweep_log_n; }
extern "C" void emul_sweep_hist(long long *out) { for (int i = 0; i < 8; ++i) { out[i] = ste::ste_emul_sweep_hist[i]; ste::ste_emul_sweep_hist[i] = 0; } }
#endif
