// TEST INFRASTRUCTURE: CPU emulation of the optimiser kernels of treepl_b200/csrc/tpl_lbfgs.cuh.
// The per-vector state machine, the Gram-matrix two-loop recursion and the element formulas are the SAME
// host/device functions the CUDA kernels call (tpl_lbfgs_core.h); only the thread grid is replaced by loops.
// tests/test_lbfgs_sim.py drives it with the CPU oracle as the evaluator to check the optimiser's logic
// without a GPU.  Not part of the product (the product path is the CUDA one and has no CPU fallback).
#include <cmath>
#include <cstring>
#include <vector>

#include "../treepl_b200/csrc/tpl_lbfgs_core.h"

using namespace tplopt;

typedef void (*eval_fn)(const double* X, double* F, double* G, void* ctx);   // X, G: [P][B]; F: [B]

template <int HM>
static int run(int P, int B, int nr, double* X, double* Fout, const double* lo, const double* hi, const double* SC, int max_steps,
               int patience, int max_ls, double ftol, double gtol, double c1, eval_fn ev, void* ctx, int* status, int* iters,
               int n, const int* parent, const int* dslot, const double* hfix, double boundary_frac) {
    typedef Layout<HM> L;
    constexpr int NB = L::NB, NACC = L::NACC;
    const size_t PB = (size_t)P * B;
    OptConst o;
    o.P = P; o.B = B; o.nr = nr; o.max_ls = max_ls; o.patience = patience; o.c1 = c1; o.ftol = ftol; o.gtol = gtol;
    std::vector<double> Z(PB), Gz(PB, 0.0), D(PB, 0.0), Gx(PB), S(PB * HM, 0.0), Y(PB * HM, 0.0), Ft(B);
    std::vector<int> st(B, ST_INIT), ls(B, 0), head(B, 0), stall(B, 0), flag(B, 0), newdir(B, 0), it(B, 0), nfail(B, 0);
    std::vector<unsigned> valid(B, 0u);
    std::vector<double> F0(B, 0.0), t(B, 0.0), gd(B, 0.0), gamma(B, 0.0), gmax(B, 0.0), tlog(B, 0.0);
    std::vector<double> rho((size_t)HM * B, 0.0), delta((size_t)NB * B, 0.0), M((size_t)NB * NB * B, 0.0), acc((size_t)NACC * B, 0.0);
    // opt_init_kernel
    for (int p = 0; p < P; p++)
        for (int b = 0; b < B; b++) {
            const size_t i = (size_t)p * B + b;
            const double w = p < nr ? log(X[i]) : X[i];
            Z[i] = fmin(fmax(w / SC[i], lo[p] / SC[i]), hi[p] / SC[i]);
        }
    auto trial_kernel = [&]() {
        for (int b = 0; b < B; b++) {
            if (st[b] == ST_DONE) continue;
            if (newdir[b]) {
                // opt_direction_kernel
                for (int p = 0; p < P; p++) {
                    const size_t i = (size_t)p * B + b;
                    double q = delta[(size_t)L::IG * B + b] * Gz[i];
                    for (int k = 0; k < HM; k++)
                        if ((valid[b] >> k) & 1u) {
                            q = fma(delta[(size_t)k * B + b], S[(size_t)k * PB + i], q);
                            q = fma(delta[(size_t)(HM + k) * B + b], Y[(size_t)k * PB + i], q);
                        }
                    D[i] = is_free(Z[i], Gz[i], lo[p] / SC[i], hi[p] / SC[i]) ? -q : 0.0;
                }
                // opt_stepmax_kernel
                if (boundary_frac > 0.0) {
                    double tmin = 1e300;
                    for (int i = 1; i < n; i++) {
                        const int pa = parent[i], ri = dslot[i], rp = dslot[pa];
                        // in date units: h = z sc, dh/dt = d sc
                        const double zi = ri >= 0 ? Z[(size_t)ri * B + b] * SC[(size_t)ri * B + b] : hfix[i];
                        const double di = ri >= 0 ? D[(size_t)ri * B + b] * SC[(size_t)ri * B + b] : 0.0;
                        const double zp = rp >= 0 ? Z[(size_t)rp * B + b] * SC[(size_t)rp * B + b] : hfix[pa];
                        const double dp = rp >= 0 ? D[(size_t)rp * B + b] * SC[(size_t)rp * B + b] : 0.0;
                        tmin = fmin(tmin, branch_step_limit(zi, di, zp, dp));
                    }
                    t[b] = fmin(t[b], boundary_frac * tmin);
                }
            }
            // opt_trial_kernel
            for (int p = 0; p < P; p++) {
                const size_t i = (size_t)p * B + b;
                const double zt = trial_point(Z[i], t[b], D[i], lo[p] / SC[i], hi[p] / SC[i]);
                X[i] = to_param(p < nr, zt, SC[i]);
            }
        }
    };
    trial_kernel();
    int ndone = 0, steps = 0;
    while (steps < max_steps && ndone < B) {
        ev(X, Ft.data(), Gx.data(), ctx);
        steps++;
        // opt_update_kernel
        for (int b = 0; b < B; b++) {
            const int dec = decide(st[b], Ft[b], F0[b], t[b], gd[b], o.c1);
            if (dec != DEC_ACCEPT && dec != DEC_INIT) continue;
            const bool al = dec == DEC_ACCEPT;
            const int h = head[b];
            double a[NACC];
            for (int k = 0; k < NACC; k++) a[k] = 0.0;
            for (int p = 0; p < P; p++) {
                const size_t i = (size_t)p * B + b;
                const double gzn = chain_grad(p < nr, Gx[i], X[i], SC[i]);
                double z = Z[i], s = 0.0, y = 0.0;
                if (al) {
                    const double zt = trial_point(z, t[b], D[i], lo[p] / SC[i], hi[p] / SC[i]);
                    s = zt - z;
                    y = gzn - Gz[i];
                    z = zt;
                    Z[i] = zt;
                    S[(size_t)h * PB + i] = s;
                    Y[(size_t)h * PB + i] = y;
                }
                Gz[i] = gzn;
                const double gp = is_free(z, gzn, lo[p] / SC[i], hi[p] / SC[i]) ? gzn : 0.0;
                a[L::A_GG] = fma(gp, gp, a[L::A_GG]);
                a[L::A_GMAX] = fmax(a[L::A_GMAX], fabs(gp));
                if (al)
                    for (int k = 0; k < HM; k++) {
                        const bool cur = k == h;
                        if (cur || ((valid[b] >> k) & 1u)) {
                            const double Sk = cur ? s : S[(size_t)k * PB + i], Yk = cur ? y : Y[(size_t)k * PB + i];
                            a[6 * k + 0] = fma(s, Sk, a[6 * k + 0]);
                            a[6 * k + 1] = fma(s, Yk, a[6 * k + 1]);
                            a[6 * k + 2] = fma(y, Sk, a[6 * k + 2]);
                            a[6 * k + 3] = fma(y, Yk, a[6 * k + 3]);
                            a[6 * k + 4] = fma(gp, Sk, a[6 * k + 4]);
                            a[6 * k + 5] = fma(gp, Yk, a[6 * k + 5]);
                        }
                    }
            }
            for (int k = 0; k < NACC; k++) acc[(size_t)k * B + b] = a[k];
        }
        // opt_lane_kernel
        for (int b = 0; b < B; b++) {
            LaneIO io;
            io.st = &st[b]; io.ls = &ls[b]; io.head = &head[b]; io.stall = &stall[b]; io.flag = &flag[b];
            io.newdir = &newdir[b]; io.iters = &it[b]; io.nfail = &nfail[b]; io.valid = &valid[b];
            io.F0 = &F0[b]; io.t = &t[b]; io.gd = &gd[b]; io.gamma = &gamma[b]; io.gmax = &gmax[b]; io.tlog = &tlog[b];
            io.rho = &rho[b]; io.rho_s = B;
            io.delta = &delta[b]; io.delta_s = B;
            io.M = &M[b]; io.M_s = B;
            io.acc = &acc[b]; io.acc_s = B;
            if (lane_step<HM>(io, Ft[b], o)) ndone++;
        }
        trial_kernel();
    }
    // opt_final_kernel
    for (int p = 0; p < P; p++)
        for (int b = 0; b < B; b++) {
            const size_t i = (size_t)p * B + b;
            X[i] = to_param(p < nr, Z[i], SC[i]);
        }
    for (int b = 0; b < B; b++) {
        Fout[b] = F0[b];
        if (status) status[b] = flag[b];
        if (iters) iters[b] = it[b];
    }
    return steps;
}

extern "C" int lbfgs_sim(int history, int P, int B, int nr, double* X, double* F, const double* lo, const double* hi, const double* SC,
                         int max_steps, int patience, int max_ls, double ftol, double gtol, double c1, eval_fn ev, void* ctx,
                         int* status, int* iters, int n, const int* parent, const int* dslot, const double* hfix, double boundary_frac) {
#define TPL_RUN(H) run<H>(P, B, nr, X, F, lo, hi, SC, max_steps, patience, max_ls, ftol, gtol, c1, ev, ctx, status, iters, n, parent, dslot, hfix, boundary_frac)
    if (history == 4) return TPL_RUN(4);
    if (history == 8) return TPL_RUN(8);
    if (history == 16) return TPL_RUN(16);
    return -1;
}
