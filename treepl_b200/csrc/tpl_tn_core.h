// Batched truncated-Newton optimiser for the PL objective: per-vector logic and the node-level Hessian
// (second optimiser).
//
// The reference's workhorse is TNC (reference src/tnc.c:483-827: truncated Newton, the Newton system solved by
// preconditioned CG with FINITE-DIFFERENCE Hessian-vector products, i.e. one extra objective+gradient callback per CG
// iteration) called through optimize_plcp_tnc (src/optimize_tnc.cpp:107-172).  Here the Hessian-vector product is
// ANALYTIC: every term of the PL objective couples a node with its parent and its children only, so
// H v is one more "pull" kernel with the gradient kernel's access pattern, exact to rounding, and B independent
// problems run their own Newton / CG / line-search state machines side by side in HBM.
//
// Objective (non-log penalty, AD flavour incl. the calibration barrier), x = (r, h):
//   f = sum_i [ -c_i ln(r_i d_i) + r_i d_i + lf_i ] + lambda sum_{p(i) != root} (r_i - r_p(i))^2
//       + lambda (sum_R r^2 - (sum_R r)^2 / m) / m + pb sum [ 1/(h - pmin) + 1/(pmax - h) ],   d_i = h_p(i) - h_i
// Second derivatives per branch i (p = parent):  d2/dr2 = c/r^2,  d2/dd2 = c/d^2,  d2/dr dd = 1, d = h_p - h_i;
// roughness: 2 lambda [[1,-1],[-1,1]] on (r_i, r_p); root children R: (2 lambda / m)(I - 11^T/m); barrier: 2 pb / q^3.
// The optimiser's variables are z = w / sc with w = ln r (rate rows) or w = h (date rows):
//   H_z = J^T H_x J + diag(g_x x''),  J = diag(sc r | sc),  x'' = sc^2 r | 0.
//
// Host/device neutral: the CUDA kernels (tpl_tn.cuh) and the CPU emulation (tests/tn_sim.cpp) compile these same
// functions; tests/test_tn_sim.py checks hess_node against finite differences of the oracle's gradient.
#ifndef TPL_TN_CORE_H
#define TPL_TN_CORE_H

#include "tpl_lbfgs_core.h"

namespace tpltn {

using tplopt::opt_finite;
using tplopt::OPT_LARGE;

enum { TS_INIT = 0, TS_CG = 1, TS_LS = 2, TS_DONE = 3 };
enum { DEC_SKIP = 0, DEC_ACCEPT = 1, DEC_REJECT = 2, DEC_INIT = 3, DEC_INIT_BAD = 4 };
// per-step activity of a vector (what the streaming kernels of this step do for it)
enum { ACT_CG = 1, ACT_DIR = 2, ACT_FIN = 4, ACT_TRIAL = 8, ACT_TAKEP = 16, ACT_SD = 32 };
// why a vector stopped: same codes as the L-BFGS driver
using tplopt::FLAG_CONV_GRAD;
using tplopt::FLAG_CONV_F;
using tplopt::FLAG_STALLED;
using tplopt::FLAG_BAD_START;
using tplopt::FLAG_NONFINITE;

struct TnConst {
    int P, B, nr;
    int max_ls, max_cg, patience;
    double c1, ftol, gtol, boundary_frac;
    double fnoise;          // absolute evaluation noise of the objective: ~1e-14 sum_i c_i (terms of size c ln(rd) cancel)
};

// Armijo test with a noise allowance: the objective is a sum of ~N cancelling terms of size c ln(r d), so near the optimum
// its evaluation noise (~1e-16 sum c |ln(rd)|, far above ulp(f)) exceeds the decrease a converged Newton step predicts;
// without the allowance such steps are halved max_ls times on noise.  Accepted steps within the allowance count towards
// `patience` (phase_setup), which then stops the vector.
TPL_HD double f_allowance(double F0, double ftol, double fnoise) { return ftol * fmax(fabs(F0), 1.0) + fnoise; }
TPL_HD int decide(int st, double Ft, double F0, double t, double gd, double c1, double allowance) {
    if (st == TS_DONE || st == TS_CG) return DEC_SKIP;
    const bool bad = !opt_finite(Ft) || !(Ft < OPT_LARGE);
    if (st == TS_INIT) return bad ? DEC_INIT_BAD : DEC_INIT;
    if (!bad && Ft <= F0 + c1 * t * gd + allowance) return DEC_ACCEPT;
    return DEC_REJECT;
}

// Levenberg-Marquardt damping of the Newton system, (H + mu |diag H|) s = -g, per vector: raised while steps have to be
// cut (ratio test, backtracking), lowered towards the pure Newton step while unit steps succeed.  Evaluated with the
// step that has just been accepted, BEFORE the factorisation at the new point (the kernels of that step call this
// function themselves; phase_setup stores the value).  The thresholds come from a sweep over the example trees and a
// 20,000-tip synthetic tree (smoothing 1e4 ... 1e-4): 2.3 x fewer Newton steps in total than without damping, and little
// sensitivity to the exact values.
TPL_HD double mu_after_accept(double mu, double t, int ls) {
    if (ls == 0 && t >= 0.7) return mu * 0.1 < 1e-8 ? 0.0 : mu * 0.1;
    if (t < 0.1) return fmin(fmax(4.0 * mu, 1e-3), 1e8);
    return mu;
}

struct Lane {
    int *st, *ls, *cgit, *flag, *stall, *iters, *ncg, *act;
    double *F0, *t, *gd, *rz, *rz0, *tol2, *alpha, *beta, *gmax, *tlog, *mu;
};

// after the evaluation of the trial point and the set-up pass (red = {r.M^-1 r, max|gp|, gp.gp} of the new point)
TPL_HD bool phase_setup(const Lane& L, double Ft, const double* red, const TnConst& o) {
    const int dec = decide(*L.st, Ft, *L.F0, *L.t, *L.gd, o.c1, f_allowance(*L.F0, o.ftol, o.fnoise));
    if (dec == DEC_SKIP) return false;
    *L.act = 0;
    if (dec == DEC_INIT_BAD) { *L.st = TS_DONE; *L.flag = FLAG_BAD_START; return true; }
    if (dec == DEC_REJECT) {
        const int ls = *L.ls + 1;
        if (ls < o.max_ls) { *L.ls = ls; *L.t *= 0.5; *L.act = ACT_TRIAL; return false; }
        if (*L.cgit >= 0) {
            // the truncated-Newton direction found no decrease: one more line search along the preconditioned
            // steepest-descent direction (cgit = -1 marks the retry) before the vector is declared stalled
            *L.cgit = -1;
            *L.ls = 0;
            *L.t = 1.0;
            *L.gd = -*L.rz0;
            *L.act = ACT_TRIAL | ACT_SD;
            return false;
        }
        *L.st = TS_DONE;
        *L.flag = FLAG_STALLED;
        return true;
    }
    const double rz0 = red[0], gmax = red[1], gg = red[2];
    if (!opt_finite(gg) || !opt_finite(rz0)) { *L.st = TS_DONE; *L.flag = FLAG_NONFINITE; return true; }
    bool done = false;
    if (dec == DEC_ACCEPT) {
        const double F0 = *L.F0;
        const bool cut_short = *L.ls == 0 && *L.t < 1.0;        // the ratio test, not the objective, limited the step
        const int stl = (fabs(F0 - Ft) <= f_allowance(F0, o.ftol, o.fnoise)) ? *L.stall + (cut_short ? 0 : 1) : 0;
        *L.stall = stl;
        *L.iters += 1;
        *L.tlog += log10(*L.t);
        *L.mu = mu_after_accept(*L.mu, *L.t, *L.ls);
        if (stl >= o.patience) { *L.flag = FLAG_CONV_F; done = true; }
    }
    *L.F0 = Ft;
    *L.gmax = gmax;
    if (gmax <= o.gtol * fmax(fabs(Ft), 1.0) || !(rz0 > 0.0)) { *L.flag = FLAG_CONV_GRAD; done = true; }
    if (done) { *L.st = TS_DONE; return true; }
    // start CG on H s = -gp: forcing sequence eta = min(0.1, sqrt(|gp|)) on the preconditioned residual
    const double eta = fmin(0.1, sqrt(sqrt(gg)));
    *L.st = TS_CG;
    *L.cgit = 0;
    *L.rz = rz0;
    *L.rz0 = rz0;
    *L.tol2 = eta * eta * rz0;
    return false;
}

// after H p: red = {p.Hp, p.p}
TPL_HD void phase_hv(const Lane& L, const double* red, const TnConst& o) {
    if (*L.st != TS_CG || (*L.act & ACT_FIN)) return;
    const double pHp = red[0], pp = red[1];
    if (!(pHp > 1e-14 * pp)) {                 // negative / vanishing curvature (or NaN): truncate
        *L.act = ACT_FIN | (*L.cgit == 0 ? ACT_TAKEP : 0);
        return;
    }
    *L.alpha = *L.rz / pHp;
    *L.cgit += 1;
    *L.ncg += 1;
    *L.act = ACT_CG;
}

// after s += alpha p, r -= alpha Hp, zz = M^-1 r: red = {r.zz}
TPL_HD void phase_update(const Lane& L, const double* red, const TnConst& o) {
    if (!(*L.act & ACT_CG)) return;
    const double rzn = red[0];
    if (!(rzn > *L.tol2) || *L.cgit >= o.max_cg) { *L.act = ACT_FIN; return; }     // also catches NaN
    *L.beta = rzn / *L.rz;
    *L.rz = rzn;
    *L.act = ACT_DIR;
}

// after the CG solve: gd = gp.s, tmin = largest step that keeps every duration positive
// Returns true when the vector stops here: the predicted decrease -gd of the (damped) Newton step is below the resolution
// of the objective (a hundredth of ftol |f|) — without this test a converged vector whose largest gradient entry sits
// just above gtol |f| keeps running line searches on rounding noise until `patience` runs out.
TPL_HD bool phase_finish(const Lane& L, double gd, double tmin, const TnConst& o) {
    if (!(*L.act & ACT_FIN)) return false;
    if (gd <= 0.0 && -gd <= 1e-2 * f_allowance(*L.F0, o.ftol, o.fnoise)) {
        *L.st = TS_DONE;
        *L.flag = FLAG_CONV_F;
        *L.act = 0;
        return true;
    }
    *L.st = TS_LS;
    *L.ls = 0;
    if (!(gd < 0.0)) {                          // not a descent direction: preconditioned steepest descent
        *L.gd = -*L.rz0;
        *L.t = 1.0;
        *L.act = ACT_TRIAL | ACT_SD;
        return false;
    }
    *L.gd = gd;
    *L.t = o.boundary_frac > 0.0 ? fmin(1.0, o.boundary_frac * tmin) : 1.0;
    *L.act = ACT_TRIAL;
    return false;
}

// ---- node-level Hessian --------------------------------------------------------------------------------------------
// View: accessor of ONE vector's operands (device: strided global arrays; host: plain arrays).  Required members:
//   int n; int parent(i), child_begin(i), child_end(i), child(j), rslot(i), dslot(i);
//   double c(i), hfix(i), rfix(i), pmin(i), pmax(i), sc(row), x(row), v(row), gx(row), lam, pb; bool masked(i), log_pen;
//   double mu (Levenberg-Marquardt damping of this vector), dg(row) (diagonal of H_z, as written by diag_node).
// Writes W(row) through view.put(row, value).

template <class View>
TPL_HD void node_values(const View& V, int i, bool& rfree, double& r, double& h, double& vr, double& vh) {
    const int rs = V.rslot(i), ds = V.dslot(i);
    rfree = rs >= 0 && !V.masked(i);
    r = rs >= 0 ? V.x(rs) : V.rfix(i);        // a masked node keeps the rate of its X row (it is not a variable)
    h = ds >= 0 ? V.x(ds) : V.hfix(i);
    vr = rfree ? V.sc(rs) * r * V.v(rs) : 0.0;
    vh = ds >= 0 ? V.sc(ds) * V.v(ds) : 0.0;
}

TPL_HD double safe_duration(double d) { return d > 2.2250738585072014e-308 ? d : 2.2250738585072014e-308; }

// Reciprocal for the Hessian entries (c / r^2, c / d^2, 1 / det).  The node-indexed kernels are bound by the latency of one
// warp's instruction chain per node, and an IEEE fp64 division is ~25 dependent instructions: on the device an ordinary
// operand (2^-400 .. 2^400) takes the MUFU.RCP64H seed + two Newton steps (5 instructions, <= 2 ulp - the preconditioner
// and the CG products tolerate far more), anything else the correctly rounded reciprocal.  Host: 1.0 / x.
TPL_HD double tn_rcp(double x) {
#ifdef __CUDA_ARCH__
    const unsigned e = ((unsigned)__double2hiint(x)) >> 20;
    if ((e - 623u) < 800u) {
        double y;
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
        double t = fma(-x, y, 1.0);
        y = fma(y, t, y);
        t = fma(-x, y, 1.0);
        return fma(y, t, y);
    }
    return __drcp_rn(x);
#else
    return 1.0 / x;
#endif
}

// W = H_z v restricted to the rows of node i; returns v.W over those rows in vw and v.v in vv
template <class View>
TPL_HD void hess_node(View& V, int i, double& vw, double& vv) {
    bool rfree;
    double r, h, vr, vh;
    node_values(V, i, rfree, r, h, vr, vh);
    const int rs = V.rslot(i), ds = V.dslot(i);
    const bool mi = V.masked(i);
    const double lam2 = 2.0 * V.lam;
    double w_r = 0.0, w_h = 0.0;
    if (i != 0 && !mi) {
        const int p = V.parent(i);
        bool pfree;
        double rp, hp, vrp, vhp;
        node_values(V, p, pfree, rp, hp, vrp, vhp);
        const double d = safe_duration(hp - h);
        const double c = V.c(i);
        const double dv = vhp - vh;
        w_r = (c * tn_rcp(r * r)) * vr + dv;
        w_h = -((c * tn_rcp(d * d)) * dv + vr);
        if (p != 0) {
            w_r += lam2 * (vr - vrp);
        } else {                                             // child of the root: variance term over ALL root children
            double sum = 0.0, m = 0.0, sy = 0.0;
            for (int j = V.child_begin(0); j < V.child_end(0); j++) {
                bool f2;
                double r2, h2, vr2, vh2;
                node_values(V, V.child(j), f2, r2, h2, vr2, vh2);
                if (V.log_pen) { sum += vr2 / r2; sy += log(r2); }      // y = ln r: dy = dr / r
                else sum += vr2;
                m += 1.0;
            }
            if (V.log_pen)                                    // pl_calc_parallel.cpp:866-879 with y = ln r
                w_r += (lam2 / m) * ((vr / r - sum / m) / r - (log(r) - sy / m) * vr / (r * r));
            else
                w_r += (lam2 / m) * (vr - sum / m);
        }
    }
    for (int j = V.child_begin(i); j < V.child_end(i); j++) {
        const int ch = V.child(j);
        if (V.masked(ch)) continue;
        bool cfree;
        double rc, hc, vrc, vhc;
        node_values(V, ch, cfree, rc, hc, vrc, vhc);
        if (ds >= 0) {
            const double dc = safe_duration(h - hc);
            w_h += (V.c(ch) * tn_rcp(dc * dc)) * (vh - vhc) + vrc;
        }
        if (rfree && i != 0 && !mi) w_r -= lam2 * (vrc - vr);
    }
    if (ds >= 0) {
        double bar = 0.0;
        const double pmn = V.pmin(i), pmx = V.pmax(i);
        if (pmn != -1.0) { const double q = h - pmn; if (q > 0.0) bar += 2.0 / (q * q * q); }
        if (pmx != -1.0) { const double q = pmx - h; if (q > 0.0) bar += 2.0 / (q * q * q); }
        w_h += V.pb * bar * vh;
    }
    vw = 0.0;
    vv = 0.0;
    if (rs >= 0) {
        double W = 0.0;
        if (rfree) {
            const double sc = V.sc(rs), v = V.v(rs);
            W = sc * r * w_r + V.gx(rs) * sc * sc * r * v;
            if (V.mu > 0.0) W += V.mu * fabs(V.dg(rs)) * v;
            vw += v * W;
            vv += v * v;
        }
        V.put(rs, W);
    }
    if (ds >= 0) {
        const double v = V.v(ds);
        double W = V.sc(ds) * w_h;
        if (V.mu > 0.0) W += V.mu * fabs(V.dg(ds)) * v;
        vw += v * W;
        vv += v * v;
        V.put(ds, W);
    }
}

template <class View>
TPL_HD double root_child_curvature(const View& V, int i, double r);

// diagonal of H_z for the rows of node i
template <class View>
TPL_HD void diag_node(View& V, int i) {
    bool rfree;
    double r, h, vr, vh;
    node_values(V, i, rfree, r, h, vr, vh);
    const int rs = V.rslot(i), ds = V.dslot(i);
    const bool mi = V.masked(i);
    const double lam2 = 2.0 * V.lam;
    double H_r = 0.0, H_h = 0.0;
    if (i != 0 && !mi) {
        const int p = V.parent(i);
        const int pds = V.dslot(p);
        const double hp = pds >= 0 ? V.x(pds) : V.hfix(p);
        const double d = safe_duration(hp - h);
        const double c = V.c(i);
        H_r = c * tn_rcp(r * r);
        H_h = c * tn_rcp(d * d);
        if (p != 0) H_r += lam2;
        else H_r += root_child_curvature(V, i, r);
    }
    for (int j = V.child_begin(i); j < V.child_end(i); j++) {
        const int ch = V.child(j);
        if (V.masked(ch)) continue;
        if (ds >= 0) {
            const int cds = V.dslot(ch);
            const double hc = cds >= 0 ? V.x(cds) : V.hfix(ch);
            const double dc = safe_duration(h - hc);
            H_h += V.c(ch) * tn_rcp(dc * dc);
        }
        if (rfree && i != 0 && !mi) H_r += lam2;
    }
    if (ds >= 0) {
        const double pmn = V.pmin(i), pmx = V.pmax(i);
        if (pmn != -1.0) { const double q = h - pmn; if (q > 0.0) H_h += V.pb * 2.0 / (q * q * q); }
        if (pmx != -1.0) { const double q = pmx - h; if (q > 0.0) H_h += V.pb * 2.0 / (q * q * q); }
    }
    if (rs >= 0) {
        double dg = 0.0;
        if (rfree) {
            const double sc = V.sc(rs);
            dg = sc * sc * r * r * H_r + V.gx(rs) * sc * sc * r;
        }
        V.put(rs, dg);
    }
    if (ds >= 0) {
        const double sc = V.sc(ds);
        V.put(ds, sc * sc * H_h);
    }
}

// ---- tree-structured preconditioner -----------------------------------------------------------------------------------
// With one 2x2 block (z_u, z_h) per node, H_z is BLOCK-TREE structured: every term of the objective couples a node with
// its parent only (branch likelihood: r_i, h_i, h_p; roughness: r_i, r_p), except the variance term over the children of
// the root, which couples siblings.  M = H_z without those sibling couplings is therefore factorised EXACTLY, without
// fill-in, by eliminating children before parents (block LDL^T over the tree, 2x2 pivots):
//     S_i = A_i - sum_ch C_ch^T S_ch^-1 C_ch          (A_i diagonal block, C_i coupling of node i with its parent)
//     up:   y_i = rhs_i - sum_ch K_ch^T y_ch,  K_i = S_i^-1 C_i
//     down: x_i = S_i^-1 y_i - K_i x_parent
// A pivot that is not safely positive definite (far from the optimum H_z can be indefinite) falls back to the absolute
// diagonal for that variable, so M is SPD by construction and CG stays valid; near the optimum M = H_z up to the sibling
// couplings of the root's children (a rank-2 perturbation for a binary root), and CG needs a handful of iterations
// whatever the smoothing value.  The reference's TNC preconditions with a diagonal BFGS update (src/tnc.c:829-905), which
// is what makes its solves at large smoothing x large trees slow; this replaces the Jacobi preconditioner of round 1.
// Node-indexed work planes TW[q][node] per vector:
enum { TW_S = 0 /* s11 s12 s22: inverse pivot */, TW_K = 3 /* k11 k12 k21 k22: S^-1 C */, TW_T = 7 /* t11 t12 t22: C^T K */,
       TW_Y = 10 /* y_u y_h */, TW_PLANES = 12 };
// Additional View members: bool fr(row) (row is a free variable of this solve), double rhs(row), double outv(row),
// double tw(q, i), void tw_put(q, i, value), bool log_pen.

template <class View>
TPL_HD void node_rh(const View& V, int i, double& r, double& h) {
    const int rs = V.rslot(i), ds = V.dslot(i);
    r = rs >= 0 ? V.x(rs) : V.rfix(i);
    h = ds >= 0 ? V.x(ds) : V.hfix(i);
}

// second derivative of the root-children variance term with respect to the rate of root child i (x space), and its
// sibling part; plain penalty: (lam2/m)(1 - 1/m).  Log penalty (pl_calc_parallel.cpp:866-879, y = ln r):
// d2/dr_i2 = (lam2/m)[(1 - 1/m) - (y_i - ybar)] / r_i^2
template <class View>
TPL_HD double root_child_curvature(const View& V, int i, double r) {
    const int j0 = V.child_begin(0), j1 = V.child_end(0);
    const double m = (double)(j1 - j0);
    const double lam2 = 2.0 * V.lam;
    if (!V.log_pen) return (lam2 / m) * (1.0 - 1.0 / m);
    double sy = 0.0;
    for (int j = j0; j < j1; j++) {
        double r2, h2;
        node_rh(V, V.child(j), r2, h2);
        sy += log(r2);
    }
    return (lam2 / m) * ((1.0 - 1.0 / m) - (log(r) - sy / m)) / (r * r);
}

// Factorisation of node i in two parts.
//   tree_local_node: everything that depends on the point X only - the node's own diagonal block of H_z (branch
//     likelihood, roughness, root-children variance, barrier) and its coupling with the parent's (u, h), in z space.
//     No dependence on other nodes' factors: the kernels run it for ALL nodes in one parallel launch.  Results go to
//     the node's own S / K planes (the pivot step overwrites them):
//       TW_L + 0..6 = a0 (signed), d0 (signed), b0, cuu, cuh, chh, flags (1: rate variable free, 2: date variable free)
//   tree_pivot_node: the elimination step proper (children done): Schur complement from the children's T, pivot
//     safeguards, inverse pivot S, K = S^-1 C, T = C^T K.  This is the part on the leaf-to-root critical path, and it is
//     short: ~100 arithmetic instructions, one reciprocal, no square root on the positive-definite path.
enum { TW_L = 0 };

template <class View>
TPL_HD void tree_local_node(View& V, int i) {
    const int rs = V.rslot(i), ds = V.dslot(i);
    const bool mi = V.masked(i);
    const bool fu = rs >= 0 && !mi && V.fr(rs);
    const bool fh = ds >= 0 && V.fr(ds);
    const double lam2 = 2.0 * V.lam;
    double r, h;
    node_rh(V, i, r, h);
    double Hrr = 0.0, Hhh = 0.0, Hrh = 0.0;
    double cuu = 0.0, cuh = 0.0, chh = 0.0;                   // z-space coupling with the parent's (u, h)
    const double su = rs >= 0 ? V.sc(rs) * r : 0.0, sh = ds >= 0 ? V.sc(ds) : 0.0;
    if (i != 0 && !mi) {
        const int p = V.parent(i);
        const double c = V.c(i);
        Hrr = c * tn_rcp(r * r);
        Hrh = -1.0;
        const int prs = V.rslot(p), pds = V.dslot(p);
        const bool pm = V.masked(p);
        if (p != 0) Hrr += lam2;
        else Hrr += root_child_curvature(V, i, r);
        const bool pu = prs >= 0 && !pm && V.fr(prs);
        const bool ph = pds >= 0 && V.fr(pds);
        if (fu && pu && p != 0) cuu = -lam2 * su * (V.sc(prs) * V.x(prs));
        if (fu && ph) cuh = su * V.sc(pds);
        if (ds >= 0) {                                         // only a node with a date variable needs its own duration
            const double hp = pds >= 0 ? V.x(pds) : V.hfix(p);
            const double d = safe_duration(hp - h);
            const double cdd = c * tn_rcp(d * d);
            Hhh = cdd;
            if (fh && ph) chh = -cdd * sh * V.sc(pds);
        }
    }
    for (int j = V.child_begin(i); j < V.child_end(i); j++) {
        const int ch = V.child(j);
        if (V.masked(ch)) continue;
        if (ds >= 0) {
            const int cds = V.dslot(ch);
            const double hc = cds >= 0 ? V.x(cds) : V.hfix(ch);
            const double dc = safe_duration(h - hc);
            Hhh += V.c(ch) * tn_rcp(dc * dc);
        }
        if (rs >= 0 && i != 0 && !mi) Hrr += lam2;
    }
    if (ds >= 0) {
        const double pmn = V.pmin(i), pmx = V.pmax(i);
        if (pmn != -1.0) { const double q = h - pmn; if (q > 0.0) Hhh += V.pb * 2.0 * tn_rcp(q * q * q); }
        if (pmx != -1.0) { const double q = pmx - h; if (q > 0.0) Hhh += V.pb * 2.0 * tn_rcp(q * q * q); }
    }
    V.tw_put(TW_L + 0, i, fu ? su * su * Hrr + V.gx(rs) * V.sc(rs) * su : 0.0);
    V.tw_put(TW_L + 1, i, fh ? sh * sh * Hhh : 0.0);
    V.tw_put(TW_L + 2, i, (fu && fh) ? su * sh * Hrh : 0.0);
    V.tw_put(TW_L + 3, i, cuu);
    V.tw_put(TW_L + 4, i, cuh);
    V.tw_put(TW_L + 5, i, chh);
    V.tw_put(TW_L + 6, i, (fu ? 1.0 : 0.0) + (fh ? 2.0 : 0.0));
}

template <class View>
TPL_HD void tree_pivot_node(View& V, int i) {
    const double a0s = V.tw(TW_L + 0, i), d0s = V.tw(TW_L + 1, i), b0 = V.tw(TW_L + 2, i);
    const double cuu = V.tw(TW_L + 3, i), cuh = V.tw(TW_L + 4, i), chh = V.tw(TW_L + 5, i);
    const double flags = V.tw(TW_L + 6, i);
    const bool fu = flags == 1.0 || flags == 3.0, fh = flags >= 2.0;
    double a = 0.0, b = 0.0, dd = 0.0;
    for (int j = V.child_begin(i); j < V.child_end(i); j++) {
        const int ch = V.child(j);
        a -= V.tw(TW_T + 0, ch);
        b -= V.tw(TW_T + 1, ch);
        dd -= V.tw(TW_T + 2, ch);
    }
    double a0 = 1.0, d0 = 1.0;                                 // the node's own diagonal (what Jacobi would use)
    if (fu) { a0 = a0s; a += a0; a0 = fabs(a0); a += V.mu * a0; } else { a = 1.0; b = 0.0; }
    if (fh) { d0 = d0s; dd += d0; d0 = fabs(d0); dd += V.mu * d0; } else { dd = 1.0; b = 0.0; }
    if (fu && fh) b += b0;
    // safeguards: every pivot positive definite (M is SPD whatever H_z is).  An indefinite pivot is replaced by its
    // absolute value |S| = Q |L| Q^T (eigenvalues mirrored, floored relative to the node's own diagonal): the curvature
    // MAGNITUDE along a negative-curvature direction still sets the step length there, and M = H_z wherever H_z is
    // positive definite.
    double s11 = 0.0, s12 = 0.0, s22 = 0.0;
    if (fu && fh) {
        const double mean = 0.5 * (a + dd), dif = 0.5 * (a - dd);
        const double q2 = dif * dif + b * b;                     // (l1 - l2)^2 / 4
        const double fl = 1e-10 * fmax(a0, d0);
        // smaller eigenvalue l2 = mean - sqrt(q2) > fl  <=>  mean - fl > 0 and (mean - fl)^2 > q2: no square root here
        const double mf = mean - fl;
        const bool pd = mf > 0.0 && mf * mf > q2 && opt_finite(mean) && opt_finite(q2);
        if (!pd) {
            const double rad = sqrt(q2);
            const double l1 = mean + rad, l2 = mean - rad;       // l1 >= l2
            if (!opt_finite(l1) || !opt_finite(l2)) {            // NaN / inf: Jacobi for this node
                a = a0 > 1e-300 ? a0 : 1.0; dd = d0 > 1e-300 ? d0 : 1.0; b = 0.0;
            } else if (!(l2 > fl)) {
                const double m1 = fmax(fabs(l1), fl), m2 = fmax(fabs(l2), fl);
                // eigenvector of l1: (b, l1 - a) or (l1 - dd, b); pick the better conditioned one
                double vx, vy;
                if (fabs(l1 - a) > fabs(l1 - dd)) { vx = b; vy = l1 - a; } else { vx = l1 - dd; vy = b; }
                const double nrm = sqrt(vx * vx + vy * vy);
                if (nrm > 0.0) { vx /= nrm; vy /= nrm; } else { vx = 1.0; vy = 0.0; }
                a = m1 * vx * vx + m2 * vy * vy;
                dd = m1 * vy * vy + m2 * vx * vx;
                b = (m1 - m2) * vx * vy;
            }
        }
        const double idet = tn_rcp(a * dd - b * b);
        s11 = dd * idet; s12 = -b * idet; s22 = a * idet;
    } else if (fu) {
        if (!(a > 1e-10 * a0)) a = fmax(fabs(a), a0 > 1e-300 ? 1e-10 * a0 : 1.0);
        if (!opt_finite(a)) a = a0 > 1e-300 && opt_finite(a0) ? a0 : 1.0;
        s11 = tn_rcp(a);
    } else if (fh) {
        if (!(dd > 1e-10 * d0)) dd = fmax(fabs(dd), d0 > 1e-300 ? 1e-10 * d0 : 1.0);
        if (!opt_finite(dd)) dd = d0 > 1e-300 && opt_finite(d0) ? d0 : 1.0;
        s22 = tn_rcp(dd);
    }
    V.tw_put(TW_S + 0, i, s11);
    V.tw_put(TW_S + 1, i, s12);
    V.tw_put(TW_S + 2, i, s22);
    const double k11 = s11 * cuu, k12 = s11 * cuh + s12 * chh, k21 = s12 * cuu, k22 = s12 * cuh + s22 * chh;
    V.tw_put(TW_K + 0, i, k11);
    V.tw_put(TW_K + 1, i, k12);
    V.tw_put(TW_K + 2, i, k21);
    V.tw_put(TW_K + 3, i, k22);
    V.tw_put(TW_T + 0, i, cuu * k11);
    V.tw_put(TW_T + 1, i, cuu * k12);
    V.tw_put(TW_T + 2, i, cuh * k12 + chh * k22);
}

// both parts in one call (the CPU emulation and small tests)
template <class View>
TPL_HD void tree_factor_node(View& V, int i) {
    tree_local_node(V, i);
    tree_pivot_node(V, i);
}

// forward substitution at node i (children done): y_i = rhs_i - sum_ch K_ch^T y_ch
template <class View>
TPL_HD void tree_up_node(View& V, int i) {
    const int rs = V.rslot(i), ds = V.dslot(i);
    double yu = (rs >= 0 && V.fr(rs) && !V.masked(i)) ? V.rhs(rs) : 0.0;
    double yh = (ds >= 0 && V.fr(ds)) ? V.rhs(ds) : 0.0;
    for (int j = V.child_begin(i); j < V.child_end(i); j++) {
        const int ch = V.child(j);
        const double cu = V.tw(TW_Y + 0, ch), chh = V.tw(TW_Y + 1, ch);
        yu -= V.tw(TW_K + 0, ch) * cu + V.tw(TW_K + 2, ch) * chh;
        yh -= V.tw(TW_K + 1, ch) * cu + V.tw(TW_K + 3, ch) * chh;
    }
    V.tw_put(TW_Y + 0, i, yu);
    V.tw_put(TW_Y + 1, i, yh);
}

// back substitution at node i (parent done): x_i = S_i^-1 y_i - K_i x_parent, written to the rows of node i
template <class View>
TPL_HD void tree_down_node(View& V, int i) {
    const int rs = V.rslot(i), ds = V.dslot(i);
    const double yu = V.tw(TW_Y + 0, i), yh = V.tw(TW_Y + 1, i);
    const double s12 = V.tw(TW_S + 1, i);
    double xu = V.tw(TW_S + 0, i) * yu + s12 * yh;
    double xh = s12 * yu + V.tw(TW_S + 2, i) * yh;
    if (i != 0) {
        const int p = V.parent(i);
        const int prs = V.rslot(p), pds = V.dslot(p);
        const double pu = prs >= 0 ? V.outv(prs) : 0.0, ph = pds >= 0 ? V.outv(pds) : 0.0;
        xu -= V.tw(TW_K + 0, i) * pu + V.tw(TW_K + 1, i) * ph;
        xh -= V.tw(TW_K + 2, i) * pu + V.tw(TW_K + 3, i) * ph;
    }
    if (rs >= 0) V.put(rs, xu);
    if (ds >= 0) V.put(ds, xh);
}

// inverse of the (absolute) diagonal; 0 freezes the variable (on a bound with the gradient pushing outward,
// or a masked row)
TPL_HD double precond_inverse(double dg, bool free_, double mu = 0.0) {
    if (!free_) return 0.0;
    const double a = fabs(dg) * (1.0 + mu);
    return a > 1e-30 ? 1.0 / a : 0.0;
}

}  // namespace tpltn
#endif
