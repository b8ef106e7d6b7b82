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
};

TPL_HD int decide(int st, double Ft, double F0, double t, double gd, double c1) {
    if (st == TS_DONE || st == TS_CG) return DEC_SKIP;
    const bool bad = !opt_finite(Ft) || !(Ft < OPT_LARGE);
    if (st == TS_INIT) return bad ? DEC_INIT_BAD : DEC_INIT;
    if (!bad && Ft <= F0 + c1 * t * gd) return DEC_ACCEPT;
    return DEC_REJECT;
}

struct Lane {
    int *st, *ls, *cgit, *flag, *stall, *iters, *ncg, *act;
    double *F0, *t, *gd, *rz, *rz0, *tol2, *alpha, *beta, *gmax, *tlog;
};

// after the evaluation of the trial point and the set-up pass (red = {r.M^-1 r, max|gp|, gp.gp} of the new point)
TPL_HD bool phase_setup(const Lane& L, double Ft, const double* red, const TnConst& o) {
    const int dec = decide(*L.st, Ft, *L.F0, *L.t, *L.gd, o.c1);
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
        const double rel = fabs(F0 - Ft) / fmax(fabs(F0), 1.0);
        const bool cut_short = *L.ls == 0 && *L.t < 1.0;        // the ratio test, not the objective, limited the step
        const int stl = (rel <= o.ftol) ? *L.stall + (cut_short ? 0 : 1) : 0;
        *L.stall = stl;
        *L.iters += 1;
        *L.tlog += log10(*L.t);
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
TPL_HD void phase_finish(const Lane& L, double gd, double tmin, const TnConst& o) {
    if (!(*L.act & ACT_FIN)) return;
    *L.st = TS_LS;
    *L.ls = 0;
    if (!(gd < 0.0)) {                          // not a descent direction: preconditioned steepest descent
        *L.gd = -*L.rz0;
        *L.t = 1.0;
        *L.act = ACT_TRIAL | ACT_SD;
        return;
    }
    *L.gd = gd;
    *L.t = o.boundary_frac > 0.0 ? fmin(1.0, o.boundary_frac * tmin) : 1.0;
    *L.act = ACT_TRIAL;
}

// ---- node-level Hessian --------------------------------------------------------------------------------------------
// View: accessor of ONE vector's operands (device: strided global arrays; host: plain arrays).  Required members:
//   int n; int parent(i), child_begin(i), child_end(i), child(j), rslot(i), dslot(i);
//   double c(i), hfix(i), rfix(i), pmin(i), pmax(i), sc(row), x(row), v(row), gx(row), lam, pb; bool masked(i);
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
        w_r = (c / (r * r)) * vr + dv;
        w_h = -((c / (d * d)) * dv + vr);
        if (p != 0) {
            w_r += lam2 * (vr - vrp);
        } else {                                             // child of the root: variance term over ALL root children
            double sum = 0.0, m = 0.0;
            for (int j = V.child_begin(0); j < V.child_end(0); j++) {
                bool f2;
                double r2, h2, vr2, vh2;
                node_values(V, V.child(j), f2, r2, h2, vr2, vh2);
                sum += vr2;
                m += 1.0;
            }
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
            w_h += (V.c(ch) / (dc * dc)) * (vh - vhc) + vrc;
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
            vw += v * W;
            vv += v * v;
        }
        V.put(rs, W);
    }
    if (ds >= 0) {
        const double v = V.v(ds);
        const double W = V.sc(ds) * w_h;
        vw += v * W;
        vv += v * v;
        V.put(ds, W);
    }
}

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
        H_r = c / (r * r);
        H_h = c / (d * d);
        if (p != 0) H_r += lam2;
        else {
            const double m = (double)(V.child_end(0) - V.child_begin(0));
            H_r += (lam2 / m) * (1.0 - 1.0 / m);
        }
    }
    for (int j = V.child_begin(i); j < V.child_end(i); j++) {
        const int ch = V.child(j);
        if (V.masked(ch)) continue;
        if (ds >= 0) {
            const int cds = V.dslot(ch);
            const double hc = cds >= 0 ? V.x(cds) : V.hfix(ch);
            const double dc = safe_duration(h - hc);
            H_h += V.c(ch) / (dc * dc);
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

// inverse of the (absolute) diagonal; 0 freezes the variable (on a bound with the gradient pushing outward,
// or a masked row)
TPL_HD double precond_inverse(double dg, bool free_) {
    if (!free_) return 0.0;
    const double a = fabs(dg);
    return a > 1e-30 ? 1.0 / a : 0.0;
}

}  // namespace tpltn
#endif
