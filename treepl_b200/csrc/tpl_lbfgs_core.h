// Per-vector ("lane") logic of the device-resident batched L-BFGS optimiser.
//
// The reference minimises ONE problem at a time on the host: optimize_plcp_tnc / optimize_plcp_nlopt
// (reference src/optimize_tnc.cpp:107-172, src/optimize_nlopt.cpp:116-238) driven by optimize_full
// (src/utils.cpp:671-763), one callback round trip per trial point.  Here B independent problems
// (CV folds, smoothing-grid points, restarts) advance together: every trial point of every vector is
// evaluated by ONE launch of the PL kernels, and each vector is its own small state machine
// (INIT -> SEARCH -> DONE) whose state lives in HBM — no host decision per evaluation.
//
// L-BFGS is kept in the "vector-free" form: the two-loop recursion runs on the (2m+1) x (2m+1) Gram matrix of
// the basis {S_0..S_{m-1}, Y_0..Y_{m-1}, Gp} and yields coefficients delta; the direction is one streaming
// pass D = -free * sum_j delta_j b_j.  Per accepted step the history is read twice (dots, direction) instead of
// the ~10 m vector passes of the textbook two-loop.
//
// This header is host/device neutral: the CUDA kernels (tpl_lbfgs.cuh) call it per lane, and
// tests/lbfgs_sim.cpp compiles the very same functions with g++ to exercise the state machine on the CPU
// against the oracle (test infrastructure; the product path is the CUDA one).
#ifndef TPL_LBFGS_CORE_H
#define TPL_LBFGS_CORE_H

#include <math.h>

#if defined(__CUDACC__)
#define TPL_HD __host__ __device__ __forceinline__
#else
#define TPL_HD inline
#endif

namespace tplopt {

constexpr double OPT_LARGE = 1e15;          // infeasible-point sentinel of the objective (pl_calc_parallel.cpp:29)

enum { ST_INIT = 0, ST_SEARCH = 1, ST_DONE = 2 };
enum { DEC_SKIP = 0, DEC_ACCEPT = 1, DEC_REJECT = 2, DEC_INIT = 3, DEC_INIT_BAD = 4 };
// why a vector stopped
enum { FLAG_RUNNING = 0, FLAG_CONV_GRAD = 1, FLAG_CONV_F = 2, FLAG_STALLED = 3, FLAG_BAD_START = 4, FLAG_NONFINITE = 5 };

struct OptConst {
    int P, B, nr;            // parameters, vectors, leading rate rows (optimised as ln r)
    int max_ls, patience;
    double c1, ftol, gtol;
};

// accumulator layout produced by the update pass for history size HM:
//   a[6k+0..5] = s.S_k, s.Y_k, y.S_k, y.Y_k, gp.S_k, gp.Y_k   (slot k; the slot being written holds the new s, y)
//   a[6HM] = gp.gp, a[6HM+1] = max|gp|
template <int HM>
struct Layout {
    static constexpr int NB = 2 * HM + 1;    // basis size
    static constexpr int IG = 2 * HM;        // index of Gp in the basis
    static constexpr int NACC = 6 * HM + 2;
    static constexpr int A_GG = 6 * HM;
    static constexpr int A_GMAX = 6 * HM + 1;
};

TPL_HD bool opt_finite(double v) { return fabs(v) <= 1.79769313486231570e308; }   // false for NaN and inf

// Armijo decision of one vector for the trial value Ft (identical in the streaming pass and in the lane pass)
TPL_HD int decide(int st, double Ft, double F0, double t, double gd, double c1) {
    if (st == ST_DONE) return DEC_SKIP;
    const bool bad = !opt_finite(Ft) || !(Ft < OPT_LARGE);
    if (st == ST_INIT) return bad ? DEC_INIT_BAD : DEC_INIT;
    if (!bad && Ft <= F0 + c1 * t * gd) return DEC_ACCEPT;
    return DEC_REJECT;
}

// Transformed variables.  Row p of vector b is optimised as z = w / sc with w = ln r (rate rows) or w = h (date rows)
// and sc > 0 a per-element scale (a diagonal preconditioner: sc ~ 1 / sqrt(d2f/dw2), see opt_scale_kernel).
TPL_HD double to_param(bool rate_row, double z, double sc) {
    const double w = z * sc;
    return rate_row ? exp(w) : w;
}
// chain rule: df/dz = df/dx * dx/dw * sc, dx/dw = x for rate rows, 1 for date rows
TPL_HD double chain_grad(bool rate_row, double gx, double x, double sc) {
    return (rate_row ? gx * x : gx) * sc;
}
// trial point z + t d, clamped into the (open) box lo <= z <= hi (bounds already divided by the row's scale)
TPL_HD double trial_point(double z, double t, double d, double lo, double hi) {
    return fmin(fmax(fma(t, d, z), lo), hi);
}
// a variable on a bound whose gradient pushes outward is frozen (projected gradient)
TPL_HD bool is_free(double z, double g, double lo, double hi) {
    return !((z <= lo && g > 0.0) || (z >= hi && g < 0.0));
}

// ratio test of one branch: the step length at which the duration zp - zi of z + t d reaches zero (1e300: never)
TPL_HD double branch_step_limit(double zi, double di, double zp, double dp) {
    const double shrink = di - dp;
    return shrink > 0.0 ? (zp - zi) / shrink : 1e300;
}

// fraction-to-the-boundary rule for the box: the step length at which z + t d reaches its bound (1e300: never).  A
// variable already ON the bound is left to the clamp of trial_point.  Date bounds are never active at an optimum (a
// calibrated node carries the barrier term, every other node's bound is implied by its neighbours' durations), and a
// step clamped onto them can close a chain of durations to exactly zero, where the objective is finite (d == 0 ->
// DBL_MIN, pl_calc_parallel.cpp:905) and the iteration stays trapped.
TPL_HD double bound_step_limit(double z, double d, double lo, double hi) {
    if (d < 0.0 && z > lo) return (z - lo) / (-d);
    if (d > 0.0 && z < hi) return (hi - z) / d;
    return 1e300;
}

struct LaneIO {
    // scalars of this vector
    int *st, *ls, *head, *stall, *flag, *newdir, *iters, *nfail;
    unsigned* valid;
    double *F0, *t, *gd, *gamma, *gmax, *tlog;   // tlog: sum of log10(accepted step lengths), a diagnostic
    double* rho;   int rho_s;       // rho[k * rho_s], k < HM
    double* delta; int delta_s;     // delta[j * delta_s], j < 2 HM + 1
    double* M;     int M_s;         // Gram matrix M[(i * NB + j) * M_s]
    const double* acc; int acc_s;   // reduced accumulators acc[a * acc_s]
};

// One optimiser step of one vector after its trial value Ft is known.  Returns true when the vector
// finished in THIS call.
template <int HM>
TPL_HD bool lane_step(const LaneIO& L, double Ft, const OptConst& o) {
    typedef Layout<HM> Y;
    constexpr int NB = Y::NB, IG = Y::IG;
#define TPL_M(i, j) L.M[((i) * NB + (j)) * L.M_s]
#define TPL_A(a) L.acc[(a) * L.acc_s]
    const int dec = decide(*L.st, Ft, *L.F0, *L.t, *L.gd, o.c1);
    if (dec == DEC_SKIP) return false;
    *L.newdir = 0;
    if (dec == DEC_INIT_BAD) {
        *L.st = ST_DONE;
        *L.flag = FLAG_BAD_START;
        return true;
    }
    double gg, gmax;
    bool done = false;
    if (dec == DEC_REJECT) {
        const int ls = *L.ls + 1;
        if (ls < o.max_ls) {          // backtrack
            *L.ls = ls;
            *L.t *= 0.5;
            return false;
        }
        // line search failed: drop the history, count a stall, restart along -gamma * Gp
        *L.valid = 0u;
        *L.nfail += 1;
        const int stl = *L.stall + 1;
        *L.stall = stl;
        if (stl >= o.patience) { *L.flag = FLAG_STALLED; done = true; }
        gg = TPL_M(IG, IG);
        gmax = *L.gmax;
    } else {
        gg = TPL_A(Y::A_GG);
        gmax = TPL_A(Y::A_GMAX);
        if (!opt_finite(gg)) {
            *L.st = ST_DONE;
            *L.flag = FLAG_NONFINITE;
            return true;
        }
        if (dec == DEC_ACCEPT) {
            const int h = *L.head;
            for (int k = 0; k < HM; k++) {
                const double sS = TPL_A(6 * k + 0), sY = TPL_A(6 * k + 1), yS = TPL_A(6 * k + 2), yY = TPL_A(6 * k + 3);
                TPL_M(h, k) = sS;            TPL_M(k, h) = sS;
                TPL_M(h, HM + k) = sY;       TPL_M(HM + k, h) = sY;
                TPL_M(HM + h, k) = yS;       TPL_M(k, HM + h) = yS;
                TPL_M(HM + h, HM + k) = yY;  TPL_M(HM + k, HM + h) = yY;
            }
            const double ss = TPL_A(6 * h + 0), sy = TPL_A(6 * h + 1), yy = TPL_A(6 * h + 3);
            const bool good = sy > 1e-10 * sqrt(ss * yy + 1e-300);
            unsigned v = *L.valid;
            v = good ? (v | (1u << h)) : (v & ~(1u << h));
            *L.valid = v;
            L.rho[h * L.rho_s] = good ? 1.0 / sy : 0.0;
            if (good) *L.gamma = sy / yy;
            *L.head = (h + 1) % HM;
            const double F0 = *L.F0;
            const double rel = fabs(F0 - Ft) / fmax(fabs(F0), 1.0);
            // a first-trial step that the ratio test cut short (t < 1 without backtracking) says nothing about
            // convergence: the boundary, not the objective, limited the decrease
            const bool cut_short = *L.ls == 0 && *L.t < 1.0;
            const int stl = (rel <= o.ftol) ? *L.stall + (cut_short ? 0 : 1) : 0;
            *L.stall = stl;
            *L.iters += 1;
            *L.tlog += log10(*L.t);
            if (stl >= o.patience) { *L.flag = FLAG_CONV_F; done = true; }
        } else {
            *L.gamma = 1.0;
        }
        for (int k = 0; k < HM; k++) {
            const double gS = TPL_A(6 * k + 4), gY = TPL_A(6 * k + 5);
            TPL_M(IG, k) = gS;       TPL_M(k, IG) = gS;
            TPL_M(IG, HM + k) = gY;  TPL_M(HM + k, IG) = gY;
        }
        TPL_M(IG, IG) = gg;
        *L.F0 = Ft;
        *L.gmax = gmax;
        if (gmax <= o.gtol * fmax(fabs(Ft), 1.0)) { *L.flag = FLAG_CONV_GRAD; done = true; }
    }
    if (done) {
        *L.st = ST_DONE;
        return true;
    }
    // ---- two-loop recursion on basis coefficients: q = sum_j dl[j] b_j, starts as Gp ----
    double dl[NB], al[HM];
    for (int j = 0; j < NB; j++) dl[j] = 0.0;
    dl[IG] = 1.0;
    const unsigned valid = *L.valid;
    const int head = *L.head;
    auto dotrow = [&](int r) {                     // b_r . q over the valid part of the basis
        double acc = dl[IG] * TPL_M(r, IG);
        for (int k = 0; k < HM; k++)
            if ((valid >> k) & 1u) acc += dl[k] * TPL_M(r, k) + dl[HM + k] * TPL_M(r, HM + k);
        return acc;
    };
    for (int j = 0; j < HM; j++) {                 // newest -> oldest
        const int k = (head - 1 - j + 2 * HM) % HM;
        al[k] = 0.0;
        if (!((valid >> k) & 1u)) continue;
        const double a = L.rho[k * L.rho_s] * dotrow(k);
        al[k] = a;
        dl[HM + k] -= a;
    }
    const double gamma = *L.gamma;
    for (int j = 0; j < NB; j++) dl[j] *= gamma;
    for (int j = HM - 1; j >= 0; j--) {            // oldest -> newest
        const int k = (head - 1 - j + 2 * HM) % HM;
        if (!((valid >> k) & 1u)) continue;
        const double b = L.rho[k * L.rho_s] * dotrow(HM + k);
        dl[k] += al[k] - b;
    }
    double gd = -dotrow(IG);
    if (dec == DEC_INIT || !(gd < 0.0) || !opt_finite(gd)) {
        // first step, or not a descent direction: scaled steepest descent
        for (int j = 0; j < NB; j++) dl[j] = 0.0;
        dl[IG] = 0.1 / fmax(gmax, 1e-300);
        gd = -dl[IG] * gg;
    }
    for (int j = 0; j < NB; j++) L.delta[j * L.delta_s] = dl[j];
    *L.gd = gd;
    *L.t = 1.0;
    *L.ls = 0;
    *L.newdir = 1;
    *L.st = ST_SEARCH;
    return false;
#undef TPL_M
#undef TPL_A
}

}  // namespace tplopt
#endif
