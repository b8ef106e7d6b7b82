// pl_calc_b200.h — C++ host-side stand-in for the reference's `class pl_calc_parallel`
// (blackrim/treePL src/pl_calc_parallel.h:17-82) over the C ABI of libtreepl_b200.so.
//
// Same member names, argument meaning and error behaviour as the reference class, so the reference's
// optimiser adapters (src/optimize_tnc.cpp, src/optimize_nlopt.cpp), annealer (src/siman_calc_par.cpp)
// and main() compile against it by replacing ONE include.  Header only; every
// evaluation runs on the GPU — there is no CPU fallback, construction failures throw.
//
// Differences a maintainer should know (all documented in INTEGRATION.md):
//  * `rates`, `dates`, `durations` are read-back accessors (methods), because the state lives with the
//    engine; main() reads them after optimisation (main.cpp:679-681, 749, 865-878) -> add "()" there.
//  * calc_gradient() differentiates at the state left by the preceding calc_pl(), exactly like the
//    reference (which ignores its `params` argument apart from params[0] in LF mode).
#pragma once
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/treepl_b200.h"

class pl_calc_b200 {
public:
    // ---- public data members of the reference that adapters read (optimize_tnc.cpp:110-139) ----
    int numparams = 0;
    int numnodes = 0;
    bool lf = false;
    std::vector<int> freeparams;
    const std::vector<int>* free = nullptr;
    const std::vector<double>* min = nullptr;
    const std::vector<double>* max = nullptr;
    const std::vector<double>* pen_min = nullptr;
    const std::vector<double>* pen_max = nullptr;
    const std::vector<std::vector<int> >* children_vec = nullptr;
    double minrate = 1e-10;
    double ftol = 1e-7;
    bool paramverbose = false;      // kept for source compatibility; the GPU path never writes paramverbose files

    pl_calc_b200() {}
    pl_calc_b200(const pl_calc_b200&) = delete;
    pl_calc_b200& operator=(const pl_calc_b200&) = delete;
    ~pl_calc_b200() { delete_ad_arrays(); }

    // pl_calc_parallel::setup_starting_bits (pl_calc_parallel.cpp:34-71).  The reference installs pen_min,
    // pen_max and children_vec by assignment right after this call (main.cpp:488-490); here the engine is
    // created lazily at the first use, when those three members have been assigned.
    void setup_starting_bits(const std::vector<int>* parents_nds, const std::vector<int>* child_cnts,
                             const std::vector<int>* vfree, const std::vector<double>* char_dur,
                             const std::vector<double>* log_fact_ch, const std::vector<double>* vmn,
                             const std::vector<double>* vmx, const std::vector<double>* start_dates,
                             const std::vector<double>* start_rates, const std::vector<double>* start_durations) {
        delete_ad_arrays();
        parents_ = parents_nds; child_counts_ = child_cnts; free = vfree; c_ = char_dur; lfc_ = log_fact_ch;
        min = vmn; max = vmx; sd_ = start_dates; sr_ = start_rates; sdur_ = start_durations;
        numnodes = (int)parents_nds->size();
        smoothing_ = 1; log_pen_ = false; penalty_boundary_ = 0.0025;
    }
    void set_log_pen(bool v) { log_pen_ = v; if (e_) check(tpl_set_log_pen(e_, v)); }
    void set_smoothing(double s) { smoothing_ = s; if (e_) check(tpl_set_smoothing(e_, s)); }   // reference: plp.smoothing = s
    double get_smoothing() const { return smoothing_; }
    void set_cv_nodes(std::vector<int>& cv) { ensure(); check(tpl_set_cv_nodes(e_, cv.data())); cv_ = cv; }
    std::vector<int>* get_cv_nodes() { return &cv_; }

    // pl_calc_parallel::set_freeparams (pl_calc_parallel.cpp:920-955)
    void set_freeparams(int nump, bool ilf, std::vector<int>* freep, std::vector<double>* params) {
        ensure();
        freeparams = *freep;
        numparams = nump;
        lf = ilf;
        params->assign(nump, 0.0);
        int k = check(tpl_set_freeparams(e_, nump, ilf ? 1 : 0, freep->data(), params->data()));
        params->resize(k);
    }

    // pl_calc_parallel::calc_pl (pl_calc_parallel.cpp:73-130)
    double calc_pl(std::vector<double>& params) { ensure(); return ret(tpl_calc_pl(e_, params.data())); }
    // pl_calc_parallel::calc_gradient (:657-662)
    void calc_gradient(std::vector<double>& params, std::vector<double>* g) {
        ensure();
        check(tpl_calc_gradient(e_, params.data(), g->data()));
    }
    // pl_calc_parallel::calc_function_gradient (:132-138) and its ADOL-C twin (:622-655)
    double calc_function_gradient(std::vector<double>* params, std::vector<double>* g) {
        ensure();
        return ret(tpl_calc_function_gradient(e_, params->data(), g->data()));
    }
    double calc_pl_function_gradient_adolc(std::vector<double>* params, std::vector<double>* g) {
        return calc_function_gradient(params, g);
    }
    double calc_lf_function_gradient_adolc(std::vector<double>* params, std::vector<double>* g) {
        return calc_function_gradient(params, g);
    }
    // one round trip for function_plcp (optimize_tnc.cpp:75-105): f and the analytic gradient together
    double calc_pl_and_gradient(const double* x, double* g) { ensure(); return ret(tpl_calc_pl_grad(e_, x, g, TPL_GRAD_ANALYTIC)); }

    // read-back of the members main() reads (main.cpp:679-681,749,865-878)
    std::vector<double> rates() { return get(tpl_get_rates); }
    std::vector<double> dates() { return get(tpl_get_dates); }
    std::vector<double> durations() { return get(tpl_get_durations); }

    // batched evaluation (folds / smoothing grid / restarts), X[p*B+b]
    void calc_pl_grad_batch(int B, const double* X, double* F, double* G, int mode) {
        ensure();
        check(tpl_calc_pl_grad_batch(e_, B, X, F, G, mode));
    }
    void batch_configure(int B, const double* lane_smoothing, const int* mask_off, const int* mask_nodes) {
        ensure();
        check(tpl_batch_configure(e_, B, lane_smoothing, mask_off, mask_nodes));
    }

    // Device-resident batched optimiser: B independent problems (folds, smoothing values, restarts) minimised in one
    // call instead of B x optimize_full / optimize_plcp_tnc round trips (utils.cpp:671-763, optimize_tnc.cpp:107-172).
    // X [P][B] in/out, F [B] out; lower/upper as the reference adapters build them (optimize_tnc.cpp:116-147).
    tpl_opt_result optimize_batch(int B, double* X, double* F, const double* lower, const double* upper, int mode,
                                  const tpl_opt_options* opts = nullptr, int* status = nullptr, int* iters = nullptr) {
        ensure();
        tpl_opt_result res;
        check(tpl_optimize_batch(e_, B, X, F, lower, upper, mode, opts, status, iters, &res));
        return res;
    }

    // B annealing chains at once (siman_calc_par::optimize, siman_calc_par.cpp:80-229): X [P][B] in/out, F [B] out
    void anneal_batch(int B, double* X, double* F, const tpl_anneal_options* opts = nullptr, int* status = nullptr,
                      int* iterations = nullptr) {
        ensure();
        check(tpl_anneal_batch(e_, B, X, F, opts, status, iterations));
    }

    void delete_ad_arrays() { if (e_) { tpl_destroy(e_); e_ = nullptr; } }     // main.cpp:798
    tpl_engine* handle() { ensure(); return e_; }

private:
    tpl_engine* e_ = nullptr;
    const std::vector<int>* parents_ = nullptr;
    const std::vector<int>* child_counts_ = nullptr;
    const std::vector<double>*c_ = nullptr, *lfc_ = nullptr, *sd_ = nullptr, *sr_ = nullptr, *sdur_ = nullptr;
    std::vector<int> cv_;
    double smoothing_ = 1, penalty_boundary_ = 0.0025;
    bool log_pen_ = false;

    static int check(int rc) {
        if (rc < 0) throw std::runtime_error(std::string("treepl_b200: ") + tpl_last_error());
        return rc;
    }
    static double ret(double f) {
        if (f != f && tpl_last_error()[0]) throw std::runtime_error(std::string("treepl_b200: ") + tpl_last_error());
        return f;
    }
    std::vector<double> get(int (*fn)(tpl_engine*, double*)) {
        ensure();
        std::vector<double> v(numnodes);
        check(fn(e_, v.data()));
        return v;
    }
    void ensure() {
        if (e_) return;
        if (!parents_ || !children_vec || !pen_min || !pen_max)
            throw std::runtime_error("treepl_b200: setup_starting_bits / children_vec / pen_min / pen_max not set");
        std::vector<int> flat;
        for (const auto& ch : *children_vec) flat.insert(flat.end(), ch.begin(), ch.end());
        if (flat.empty()) flat.push_back(0);
        check(tpl_create(&e_, -1, numnodes, parents_->data(), child_counts_->data(), flat.data(), free->data(), c_->data(),
                         lfc_->data(), min->data(), max->data(), pen_min->data(), pen_max->data(), sd_->data(), sr_->data(),
                         sdur_->data()));
        check(tpl_set_smoothing(e_, smoothing_));
        check(tpl_set_log_pen(e_, log_pen_ ? 1 : 0));
        check(tpl_set_penalty_boundary(e_, penalty_boundary_));
    }
};
