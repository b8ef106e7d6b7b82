// Batched simulated annealing with O(degree) delta evaluation.
//
// The reference anneals ONE chain on the host and pays a full O(N) calc_pl per proposal
// (reference src/siman_calc_par.cpp:80-229: pick one parameter — rates and dates with probability 1/2 each, :48-77 —
// scale it by (1 + u), u uniform in (-step, step) (:146-149, :241), evaluate, keep an improvement, keep a worse point with
// probability exp((-cur - test) / T) (:162-185, the reference's formula as written), cool T every 100 iterations and adapt the
// two step sizes to a 15 % acceptance rate (:187-200), stop on temperature, iteration count or 2 x 5000 non-improving
// iterations (:211-227)).  A proposal touches one rate or one date, and every term of the objective couples a node with
// its parent and children only, so the change of the objective is a handful of terms: here every chain is one GPU thread
// that walks its own proposal sequence with a counter-free xorshift generator and updates the running sums.
//
// Objective = pl_calc_parallel::calc_pl (pl_calc_parallel.cpp:73-130), PL layout:
//   F = (LL + pb * TP) + lambda * (SU + (ss - s^2/m)/m)
//   LL = sum_i -(c ln(r d) - r d - lf)         calc_log_like :821-849       (unmasked branches)
//   SU = sum_{parent != root} (r_i - r_p)^2     calc_roughness_penalty :852-881 ; s, ss, m over ALL children of the root
//   TP = barrier sum with the accumulator reset on inf (calc_penalty :975-998): recomputed over the (few) calibrated nodes
//   infeasible (date outside [min,max], negative duration, negative parameter) -> the proposal is discarded without
//   counting an iteration (:155-159).
// Parity is statistical (the reference draws from libc rand()); what IS exact and tested is the delta evaluation: the
// running objective equals a full evaluation at every step (tests/test_anneal_sim.py against the oracle's calc_pl).
#ifndef TPL_ANNEAL_CORE_H
#define TPL_ANNEAL_CORE_H

#include "tpl_lbfgs_core.h"

namespace tplsa {

constexpr double SA_LARGE = 10e+13;          // siman_calc_par.cpp:18 (the annealer's own threshold)
constexpr double SA_DBL_MIN = 2.2250738585072014e-308;

struct SaConst {
    int n, P, nr, ndates;       // nodes, parameters, rate rows, free-date rows
    int maxit;
    double cool, stop_temp, pb;
    int log_pen;
};

struct Chain {
    unsigned long long rng;
    double f, LL, SU, TP, s, ss;        // running objective and its pieces
    double temperature, step_rt, step_dt;
    int iteration, same_value, converged;
    long long rate_trials, rate_accepts, date_trials, date_accepts, discarded;
};

TPL_HD double sa_uniform(unsigned long long& st) {            // xorshift64*, 53-bit mantissa, in [0, 1)
    st ^= st >> 12;
    st ^= st << 25;
    st ^= st >> 27;
    const unsigned long long r = st * 2685821657736338717ull;
    return (double)(r >> 11) * (1.0 / 9007199254740992.0);
}

// one branch of calc_log_like (:830-845); bad is set when the reference would return LARGE
TPL_HD double sa_branch_ll(double r, double d, double c, double lf, bool& bad) {
    const double x = r * d;
    if (x > 0.0) return -(c * log(x) - x - lf);
    if (x == 0.0 && c > 0.0) bad = true;
    return 0.0;
}

// set_durations' rule for one branch (:898-907)
TPL_HD double sa_duration(double hp, double h, bool& bad) {
    double d = hp - h;
    if (d < 0.0) bad = true;
    if (d == 0.0 || d != d) d = SA_DBL_MIN;
    return d;
}

// View: one chain's operands.  Members: parent(i), child_begin(i), child_end(i), child(j), rslot(i), dslot(i), date_node(k),
// c(i), lf(i), hfix(i), rfix(i), hmin(i), hmax(i), masked(i), x(row), setx(row, v), lam, nbar, bar_node(k), bar_pmin(k), bar_pmax(k)
template <class View>
TPL_HD double sa_rate(const View& V, int i) {
    const int rs = V.rslot(i);
    return rs >= 0 ? V.x(rs) : V.rfix(i);     // a masked node keeps the rate of its X row (never proposed)
}
template <class View>
TPL_HD double sa_date(const View& V, int i) {
    const int ds = V.dslot(i);
    return ds >= 0 ? V.x(ds) : V.hfix(i);
}

// calc_penalty (:975-998) over the calibrated nodes in ascending node order, accumulator reset on inf
template <class View>
TPL_HD double sa_barrier(const View& V) {
    double tp = 0.0;
    for (int k = 0; k < V.nbar; k++) {
        const double h = sa_date(V, V.bar_node(k));
        const double pmn = V.bar_pmin(k), pmx = V.bar_pmax(k);
        if (pmn != -1.0) { tp += 1.0 / (h - pmn); if (tp > 1.79769313486231570e308 || tp < -1.79769313486231570e308) tp = 0.0; }
        if (pmx != -1.0) { tp += 1.0 / (pmx - h); if (tp > 1.79769313486231570e308 || tp < -1.79769313486231570e308) tp = 0.0; }
    }
    return tp;
}

template <class View>
TPL_HD double sa_total(const View& V, const Chain& ch, const SaConst& o) {
    const double m = (double)(V.child_end(0) - V.child_begin(0));
    return (ch.LL + o.pb * ch.TP) + V.lam * (ch.SU + (ch.ss - ch.s * ch.s / m) / m);
}

// full evaluation of the running sums (start of a chain); returns false when the start is infeasible
template <class View>
TPL_HD bool sa_init(const View& V, Chain& ch, const SaConst& o) {
    bool bad = false;
    double LL = 0.0, SU = 0.0, s = 0.0, ss = 0.0;
    for (int i = 0; i < o.n; i++) {
        const double h = sa_date(V, i);
        if (h < V.hmin(i) || h > V.hmax(i)) bad = true;
        if (i == 0) continue;
        const int p = V.parent(i);
        const double r = sa_rate(V, i);
        if (r < 0.0 || h < 0.0) bad = true;
        const double d = sa_duration(sa_date(V, p), h, bad);
        if (p == 0) {
            const double y = o.log_pen ? log(r) : r;
            s += y;
            ss += y * y;
        }
        if (V.masked(i)) continue;
        LL += sa_branch_ll(r, d, V.c(i), V.lf(i), bad);
        if (p != 0) {
            const double q = r - sa_rate(V, p);
            SU += q * q;
        }
    }
    ch.LL = LL; ch.SU = SU; ch.s = s; ch.ss = ss;
    ch.TP = sa_barrier(V);
    ch.f = sa_total(V, ch, o);
    return !bad;
}

// One proposal.  Returns true when it counted as an iteration (feasible proposal).
template <class View>
TPL_HD bool sa_step(View& V, Chain& ch, const SaConst& o) {
    // ---- choose the parameter: rates and dates with probability 1/2 each, uniform inside the class (:48-77, :98-121) ----
    const bool is_rate = (sa_uniform(ch.rng) < 0.5) || o.ndates == 0;
    int node;
    if (is_rate) {
        for (;;) {                                        // masked tips have no rate parameter
            node = 1 + (int)(sa_uniform(ch.rng) * (o.n - 1));
            if (node > o.n - 1) node = o.n - 1;
            if (V.rslot(node) >= 0 && !V.masked(node)) break;
        }
        ch.rate_trials++;
    } else {
        int k = (int)(sa_uniform(ch.rng) * o.ndates);
        if (k > o.ndates - 1) k = o.ndates - 1;
        node = V.date_node(k);
        ch.date_trials++;
    }
    const double step_size = is_rate ? ch.step_rt : ch.step_dt;
    const double step = 2.0 * step_size * sa_uniform(ch.rng) - step_size;        // :146-147
    bool bad = false;
    double dLL = 0.0, dSU = 0.0, ns = ch.s, nss = ch.ss, nTP = ch.TP;
    int row;
    double oldv, newv;
    if (is_rate) {
        row = V.rslot(node);
        oldv = V.x(row);
        newv = (1.0 + step) * oldv;                                              // make_step_plcp :241
        if (newv < 0.0) bad = true;                                              // calc_pl :74-78
        const int p = V.parent(node);
        const double d = sa_duration(sa_date(V, p), sa_date(V, node), bad);
        bool b2 = false;
        dLL = sa_branch_ll(newv, d, V.c(node), V.lf(node), bad) - sa_branch_ll(oldv, d, V.c(node), V.lf(node), b2);
        if (p != 0) {
            const double rp = sa_rate(V, p);
            dSU += (newv - rp) * (newv - rp) - (oldv - rp) * (oldv - rp);
        } else {
            const double yo = o.log_pen ? log(oldv) : oldv, yn = o.log_pen ? log(newv) : newv;
            ns += yn - yo;
            nss += yn * yn - yo * yo;
        }
        for (int j = V.child_begin(node); j < V.child_end(node); j++) {
            const int c = V.child(j);
            if (V.masked(c)) continue;
            const double rc = sa_rate(V, c);
            dSU += (rc - newv) * (rc - newv) - (rc - oldv) * (rc - oldv);
        }
    } else {
        row = V.dslot(node);
        oldv = V.x(row);
        newv = (1.0 + step) * oldv;
        if (newv < 0.0 || newv < V.hmin(node) || newv > V.hmax(node)) bad = true;   // calc_pl :74-78, set_durations :887-896
        if (node != 0) {
            const int p = V.parent(node);
            const double hp = sa_date(V, p);
            bool b2 = false;
            const double dn = sa_duration(hp, newv, bad), dold = sa_duration(hp, oldv, b2);
            if (!V.masked(node)) {
                const double r = sa_rate(V, node);
                dLL += sa_branch_ll(r, dn, V.c(node), V.lf(node), bad) - sa_branch_ll(r, dold, V.c(node), V.lf(node), b2);
            }
        }
        for (int j = V.child_begin(node); j < V.child_end(node); j++) {
            const int c = V.child(j);
            const double hc = sa_date(V, c);
            bool b2 = false;
            const double dn = sa_duration(newv, hc, bad), dold = sa_duration(oldv, hc, b2);      // a masked tip still constrains
            if (V.masked(c)) continue;
            const double rc = sa_rate(V, c);
            dLL += sa_branch_ll(rc, dn, V.c(c), V.lf(c), bad) - sa_branch_ll(rc, dold, V.c(c), V.lf(c), b2);
        }
    }
    double test = SA_LARGE;
    if (!bad) {
        if (!is_rate) {
            bool calibrated = false;
            for (int k = 0; k < V.nbar; k++) calibrated = calibrated || V.bar_node(k) == node;
            if (calibrated) {
                V.setx(row, newv);
                nTP = sa_barrier(V);
                V.setx(row, oldv);
            }
        }
        Chain t2 = ch;
        t2.LL = ch.LL + dLL; t2.SU = ch.SU + dSU; t2.s = ns; t2.ss = nss; t2.TP = nTP;
        test = sa_total(V, t2, o);
        if (!(test == test)) test = SA_LARGE;                 // NaN compares false everywhere in the reference: never kept
    }
    if (test >= SA_LARGE) {                                   // :155-159: discarded, no iteration counted
        ch.discarded++;
        ch.converged = 0;
        return false;
    }
    const double cur = ch.f;
    const double sa = exp((-cur - test) / ch.temperature);   // :162, as written in the reference
    bool accept = false;
    if (test < cur) {
        ch.same_value = (cur - test) > 0.01 ? 0 : ch.same_value + 1;
        accept = true;
        if (is_rate) ch.rate_accepts++; else ch.date_accepts++;
    } else if (sa_uniform(ch.rng) < sa) {
        accept = true;
    } else {
        ch.converged = 0;
    }
    if (accept) {
        V.setx(row, newv);
        ch.LL += dLL; ch.SU += dSU; ch.s = ns; ch.ss = nss; ch.TP = nTP;
        ch.f = test;
    }
    ch.iteration++;
    if (ch.iteration % 100 == 0) {
        ch.temperature *= o.cool;
        ch.step_rt *= ((double)ch.rate_accepts / (double)(ch.rate_trials > 0 ? ch.rate_trials : 1)) < 0.15 ? 0.99 : 1.01;
        ch.step_dt *= ((double)ch.date_accepts / (double)(ch.date_trials > 0 ? ch.date_trials : 1)) < 0.15 ? 0.99 : 1.01;
    }
    return true;
}

// the chain's stopping rule (:211-227); call after a counted iteration
TPL_HD bool sa_finished(Chain& ch, const SaConst& o) {
    if (ch.same_value >= 5000) {
        if (!ch.converged) { ch.converged = 1; ch.same_value = 0; }
        if ((ch.converged && ch.same_value >= 5000) || ch.iteration > o.maxit) return true;
    }
    return ch.iteration > o.maxit || !(ch.temperature > o.stop_temp);
}

}  // namespace tplsa
#endif
