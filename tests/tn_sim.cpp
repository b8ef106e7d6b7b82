// TEST INFRASTRUCTURE: CPU emulation of the truncated-Newton kernels of treepl_b200/csrc/tpl_tn.cuh.
// The per-vector state machine and the node-level Hessian are the SAME host/device functions the CUDA kernels call
// (tpl_tn_core.h); only the thread grid is replaced by loops and the objective comes from a callback (the CPU oracle in
// tests/test_tn_sim.py).  Also exports the Hessian-vector product alone so that it can be checked against finite
// differences of the oracle's gradient.  Not part of the product.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../treepl_b200/csrc/tpl_tn_core.h"

using namespace tpltn;
using tplopt::chain_grad;
using tplopt::is_free;
using tplopt::to_param;
using tplopt::trial_point;
using tplopt::branch_step_limit;

typedef void (*eval_fn)(const double* X, double* F, double* G, void* ctx);   // X, G: [P][B]; F: [B]

struct Tree {
    int n;
    const int *parent, *child_off, *children, *rslot, *dslot;
    const double *c, *hfix, *rfix, *pmin, *pmax;
    const unsigned char* mask;      // [B][n] per-vector CV masks or null
};

struct HostView {
    const Tree& t;
    int B, b;
    const double *X, *Vv, *Gx, *sc_;
    double* W;
    double lam, pb;
    bool log_pen = false;
    double mu = 0.0;
    const double* DGv = nullptr;
    double dg(int row) const { return DGv[(size_t)row * B + b]; }
    // tree preconditioner (tpl_tn_core.h::tree_*_node)
    const double* Minv_ = nullptr;
    const double* Rhs = nullptr;
    double* TW = nullptr;
    bool fr(int row) const { return Minv_[(size_t)row * B + b] > 0.0; }
    double rhs(int row) const { return Rhs[(size_t)row * B + b]; }
    double outv(int row) const { return W[(size_t)row * B + b]; }
    double tw(int q, int i) const { return TW[((size_t)q * t.n + i) * B + b]; }
    void tw_put(int q, int i, double v) { TW[((size_t)q * t.n + i) * B + b] = v; }
    int parent(int i) const { return t.parent[i]; }
    int child_begin(int i) const { return t.child_off[i]; }
    int child_end(int i) const { return t.child_off[i + 1]; }
    int child(int j) const { return t.children[j]; }
    int rslot(int i) const { return t.rslot[i]; }
    int dslot(int i) const { return t.dslot[i]; }
    double c(int i) const { return t.c[i]; }
    double hfix(int i) const { return t.hfix[i]; }
    double rfix(int i) const { return t.rfix[i]; }
    double pmin(int i) const { return t.pmin[i]; }
    double pmax(int i) const { return t.pmax[i]; }
    bool masked(int i) const { return t.mask && t.mask[(size_t)b * t.n + i]; }
    double sc(int row) const { return sc_[row]; }
    double x(int row) const { return X[(size_t)row * B + b]; }
    double v(int row) const { return Vv[(size_t)row * B + b]; }
    double gx(int row) const { return Gx[(size_t)row * B + b]; }
    void put(int row, double val) { W[(size_t)row * B + b] = val; }
};

extern "C" void tn_sim_hessvec(int P, int B, int n, const int* parent, const int* child_off, const int* children, const int* rslot,
                               const int* dslot, const double* c, const double* hfix, const double* rfix, const double* pmin,
                               const double* pmax, const unsigned char* mask, const double* sc, const double* lam, double pb,
                               const double* X, const double* Gx, const double* V, double* W, double* DG, int log_pen) {
    Tree t = {n, parent, child_off, children, rslot, dslot, c, hfix, rfix, pmin, pmax, mask};
    for (int b = 0; b < B; b++) {
        HostView hv = {t, B, b, X, V, Gx, sc, W, lam[b], pb, log_pen != 0};
        HostView dv = {t, B, b, X, V, Gx, sc, DG, lam[b], pb, log_pen != 0};
        for (int i = 0; i < n; i++) {
            double vw, vv;
            hess_node(hv, i, vw, vv);
            diag_node(dv, i);
        }
    }
}

// M^-1 rhs for vector b: factorise (when `factor`), forward and back substitution.  Children have larger preorder
// numbers than their parents, so descending node order is a valid elimination order (the kernels go by height levels).
static void tree_solve(const Tree& t, int B, int b, const double* X, const double* Gx, const double* sc, double lam, double pb,
                       bool log_pen, const double* Minv, const double* Rhs, double* Out, double* TW, bool factor, double mu = 0.0) {
    HostView v = {t, B, b, X, nullptr, Gx, sc, Out, lam, pb, log_pen, mu, nullptr, Minv, Rhs, TW};
    for (int i = t.n - 1; i >= 0; i--) {
        if (factor) tree_factor_node(v, i);
        tree_up_node(v, i);
    }
    for (int i = 0; i < t.n; i++) tree_down_node(v, i);
}

// test entry: Out = M^-1 Rhs with every row whose Free flag is > 0 a variable
extern "C" void tn_sim_treesolve(int P, int B, int n, const int* parent, const int* child_off, const int* children, const int* rslot,
                                 const int* dslot, const double* c, const double* hfix, const double* rfix, const double* pmin,
                                 const double* pmax, const unsigned char* mask, const double* sc, const double* lam, double pb,
                                 const double* X, const double* Gx, const double* Free, const double* Rhs, double* Out, int log_pen) {
    Tree t = {n, parent, child_off, children, rslot, dslot, c, hfix, rfix, pmin, pmax, mask};
    std::vector<double> TW((size_t)TW_PLANES * n * B, 0.0);
    for (int b = 0; b < B; b++) tree_solve(t, B, b, X, Gx, sc, lam[b], pb, log_pen != 0, Free, Rhs, Out, TW.data(), true);
}

extern "C" int tn_sim(int P, int B, int nr, double* X, double* Fout, const double* lo, const double* hi, const double* sc,
                      int max_steps, int max_cg, int cg_per_step, int patience, int max_ls, double ftol, double gtol, double c1, double boundary_frac,
                      eval_fn ev, void* ctx, int* status, int* iters, int* ncg_out,
                      int n, const int* parent, const int* child_off, const int* children, const int* rslot, const int* dslot,
                      const double* c, const double* hfix, const double* rfix, const double* pmin, const double* pmax,
                      const unsigned char* mask, const double* lam, double pb, int precond, int log_pen) {
    Tree t = {n, parent, child_off, children, rslot, dslot, c, hfix, rfix, pmin, pmax, mask};
    const size_t PB = (size_t)P * B;
    const bool lp = log_pen != 0;
    std::vector<double> TW(precond ? (size_t)TW_PLANES * n * B : 0, 0.0);
    TnConst o;
    o.P = P; o.B = B; o.nr = nr; o.max_ls = max_ls; o.max_cg = max_cg; o.patience = patience; o.c1 = c1; o.ftol = ftol; o.gtol = gtol;
    o.boundary_frac = boundary_frac;
    o.fnoise = 0.0;
    for (int i = 0; i < n; i++) o.fnoise += 1e-14 * c[i];
    std::vector<double> Z(PB), Gz(PB, 0.0), S(PB, 0.0), Gx(PB, 0.0), R(PB, 0.0), ZZ(PB, 0.0), Pv(PB, 0.0), Hp(PB, 0.0), Minv(PB, 0.0), DG(PB, 0.0), Ft(B);
    std::vector<int> st(B, TS_INIT), ls(B, 0), cgit(B, 0), flag(B, 0), stall(B, 0), it(B, 0), ncg(B, 0), act(B, ACT_TRIAL);
    std::vector<double> F0(B, 0.0), tt(B, 0.0), gd(B, 0.0), rz(B, 0.0), rz0(B, 0.0), tol2(B, 0.0), alpha(B, 0.0), beta(B, 0.0), gmax(B, 0.0), tlog(B, 0.0), mu(B, 0.0);
    auto lane = [&](int b) {
        Lane L;
        L.st = &st[b]; L.ls = &ls[b]; L.cgit = &cgit[b]; L.flag = &flag[b]; L.stall = &stall[b]; L.iters = &it[b]; L.ncg = &ncg[b]; L.act = &act[b];
        L.F0 = &F0[b]; L.t = &tt[b]; L.gd = &gd[b]; L.rz = &rz[b]; L.rz0 = &rz0[b]; L.tol2 = &tol2[b]; L.alpha = &alpha[b]; L.beta = &beta[b];
        L.gmax = &gmax[b]; L.tlog = &tlog[b]; L.mu = &mu[b];
        return L;
    };
    // tn_init_kernel
    for (int p = 0; p < P; p++)
        for (int b = 0; b < B; b++) {
            const size_t i = (size_t)p * B + b;
            const double w = p < nr ? log(X[i]) : X[i];
            Z[i] = fmin(fmax(w / sc[p], lo[p]), hi[p]);
        }
    auto trial_kernel = [&]() {
        for (int b = 0; b < B; b++) {
            if (st[b] == TS_DONE || !(act[b] & ACT_TRIAL)) continue;
            for (int p = 0; p < P; p++) {
                const size_t i = (size_t)p * B + b;
                if (act[b] & ACT_SD) S[i] = -Gz[i] * Minv[i];
                X[i] = to_param(p < nr, trial_point(Z[i], tt[b], S[i], lo[p], hi[p]), sc[p]);
            }
        }
    };
    trial_kernel();
    int ndone = 0, steps = 0;
    while (steps < max_steps && ndone < B) {
        ev(X, Ft.data(), Gx.data(), ctx);
        steps++;
        for (int b = 0; b < B; b++) {
            const int dec = decide(st[b], Ft[b], F0[b], tt[b], gd[b], o.c1, f_allowance(F0[b], o.ftol, o.fnoise));
            double red[3] = {0.0, 0.0, 0.0};
            const double mu_eff = dec == DEC_ACCEPT ? mu_after_accept(mu[b], tt[b], ls[b]) : mu[b];
            if (dec == DEC_ACCEPT || dec == DEC_INIT) {
                // tn_diag_kernel
                HostView dv = {t, B, b, X, Pv.data(), Gx.data(), sc, DG.data(), lam[b], pb, lp};
                for (int i = 0; i < n; i++) diag_node(dv, i);
                // tn_setup_kernel
                for (int p = 0; p < P; p++) {
                    const size_t i = (size_t)p * B + b;
                    double z = Z[i];
                    if (dec == DEC_ACCEPT) z = trial_point(z, tt[b], S[i], lo[p], hi[p]);
                    Z[i] = z;
                    const double gz = chain_grad(p < nr, Gx[i], X[i], sc[p]);
                    Gz[i] = gz;
                    const bool fr = is_free(z, gz, lo[p], hi[p]);
                    const double mi = precond_inverse(DG[i], fr, mu_eff);
                    Minv[i] = mi;
                    const double gp = mi > 0.0 ? gz : 0.0;
                    R[i] = -gp;
                    ZZ[i] = -gp * mi;
                    Pv[i] = ZZ[i];
                    S[i] = 0.0;
                    red[0] = fma(gp * mi, gp, red[0]);
                    red[1] = fmax(red[1], fabs(gp));
                    red[2] = fma(gp, gp, red[2]);
                }
                if (precond) {           // tn_tree_factor / up / down kernels: zz = M^-1 r, p = zz, rz0 = r.zz
                    tree_solve(t, B, b, X, Gx.data(), sc, lam[b], pb, lp, Minv.data(), R.data(), ZZ.data(), TW.data(), true, mu_eff);
                    red[0] = 0.0;
                    for (int p = 0; p < P; p++) {
                        const size_t i = (size_t)p * B + b;
                        Pv[i] = ZZ[i];
                        red[0] = fma(R[i], ZZ[i], red[0]);
                    }
                }
            }
            if (getenv("TN_SIM_VERBOSE") && dec != DEC_SKIP)
                fprintf(stderr, "step %d lane %d dec %d Ft %.10g F0 %.10g t %.3g ls %d cgit %d gd %.3g rz0 %.3g gmax %.3g mu %.3g\n", steps, b, dec, Ft[b], F0[b], tt[b], ls[b], cgit[b], gd[b], red[0], red[1], mu_eff);
            if (phase_setup(lane(b), Ft[b], red, o)) ndone++;
        }
        for (int sub = 0; sub < cg_per_step; sub++) {
        // tn_hv_kernel
        for (int b = 0; b < B; b++) {
            if (st[b] != TS_CG || (act[b] & ACT_FIN)) continue;
            HostView hv = {t, B, b, X, Pv.data(), Gx.data(), sc, Hp.data(), lam[b], pb, lp, mu[b], DG.data()};
            double red[2] = {0.0, 0.0};
            for (int i = 0; i < n; i++) {
                double vw, vv;
                hess_node(hv, i, vw, vv);
                red[0] += vw;
                red[1] += vv;
            }
            // rows frozen by the preconditioner carry no search direction; their Hp rows are never used
            phase_hv(lane(b), red, o);
        }
        // tn_update_kernel
        for (int b = 0; b < B; b++) {
            if (st[b] != TS_CG || !(act[b] & ACT_CG)) continue;
            double red[1] = {0.0};
            for (int p = 0; p < P; p++) {
                const size_t i = (size_t)p * B + b;
                S[i] = fma(alpha[b], Pv[i], S[i]);
                const double r = fma(-alpha[b], Hp[i], R[i]);
                R[i] = r;
                const double zz = r * Minv[i];
                ZZ[i] = zz;
                red[0] = fma(r, zz, red[0]);
            }
            if (precond) {
                tree_solve(t, B, b, X, Gx.data(), sc, lam[b], pb, lp, Minv.data(), R.data(), ZZ.data(), TW.data(), false, mu[b]);
                red[0] = 0.0;
                for (int p = 0; p < P; p++) red[0] = fma(R[(size_t)p * B + b], ZZ[(size_t)p * B + b], red[0]);
            }
            phase_update(lane(b), red, o);
        }
        // tn_dir_kernel
        for (int b = 0; b < B; b++) {
            if (st[b] != TS_CG || !(act[b] & ACT_DIR)) continue;
            for (int p = 0; p < P; p++) {
                const size_t i = (size_t)p * B + b;
                Pv[i] = fma(beta[b], Pv[i], ZZ[i]);
            }
        }
        }
        // tn_gd_kernel + tn_stepmax_kernel
        for (int b = 0; b < B; b++) {
            if (st[b] != TS_CG || !(act[b] & ACT_FIN)) continue;
            double g = 0.0;
            for (int p = 0; p < P; p++) {
                const size_t i = (size_t)p * B + b;
                if (act[b] & ACT_TAKEP) S[i] = Pv[i];
                g = fma(Gz[i], S[i], g);
            }
            double tmin = 1e300;
            for (int i = 1; i < n; i++) {
                const int pa = parent[i], ri = dslot[i], rp = dslot[pa];
                const double zi = ri >= 0 ? Z[(size_t)ri * B + b] * sc[ri] : hfix[i];
                const double di = ri >= 0 ? S[(size_t)ri * B + b] * sc[ri] : 0.0;
                const double zp = rp >= 0 ? Z[(size_t)rp * B + b] * sc[rp] : hfix[pa];
                const double dp = rp >= 0 ? S[(size_t)rp * B + b] * sc[rp] : 0.0;
                tmin = fmin(tmin, branch_step_limit(zi, di, zp, dp));
            }
            for (int i = 0; i < n; i++) {
                const int ri = dslot[i];
                if (ri >= 0) tmin = fmin(tmin, tplopt::bound_step_limit(Z[(size_t)ri * B + b], S[(size_t)ri * B + b], lo[ri], hi[ri]));
            }
            if (phase_finish(lane(b), g, tmin, o)) ndone++;
        }
        trial_kernel();
    }
    for (int p = 0; p < P; p++)
        for (int b = 0; b < B; b++) {
            const size_t i = (size_t)p * B + b;
            X[i] = to_param(p < nr, Z[i], sc[p]);
        }
    for (int b = 0; b < B; b++) {
        Fout[b] = F0[b];
        if (status) status[b] = flag[b];
        if (iters) iters[b] = it[b];
        if (ncg_out) ncg_out[b] = ncg[b];
    }
    return steps;
}
