// pl_calc_b200.h — C++ host-side stand-in for the reference's `class pl_calc_parallel`
// (blackrim/treePL src/pl_calc_parallel.h:17-82) over the C ABI of libtreepl_b200.so.
//
// Same member names, argument meaning and error behaviour as the reference class, INCLUDING its public data
// members, so the reference's optimiser adapters (src/optimize_tnc.cpp, src/optimize_nlopt.cpp), annealer
// (src/siman_calc_par.cpp), optimize_full (src/utils.cpp) and main() compile against it UNMODIFIED: put
// treepl_b200/host/shim/ (a one-line "pl_calc_parallel.h" that typedefs this class) in front of the reference's
// own include directory.  tests/dropin/ does exactly that with the reference sources where they lie and runs the
// reference's TNC / NLopt drivers and its whole CLI on the GPU engine.
//
// Header only; every evaluation runs on the GPU — there is no CPU fallback, construction failures throw.
//
// How the reference's public STATE members are kept source compatible although the state lives with the engine:
//  * `rates`, `dates`, `durations` are read-only views (`member_vector`): size(), [], at(), begin()/end() and an
//    implicit conversion to `const std::vector<double>&`, refreshed lazily from the engine after every call that
//    changes its state (main.cpp:664-681, 749, 865-878; optimize_tnc.cpp:121; siman_calc_par.cpp:65-71, 401-403).
//  * `smoothing` is a `double`-like member whose assignment reaches the engine (main.cpp:490, 673, 818, 821).
//  * calc_gradient() differentiates at the state left by the preceding calc_pl(), exactly like the
//    reference (which ignores its `params` argument apart from params[0] in LF mode).
#pragma once
#include <cstddef>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/treepl_b200.h"

class pl_calc_b200 {
public:
    // read-only view of an engine-owned state vector (reference: `vector<double> rates, dates, durations`)
    class member_vector {
    public:
        typedef std::vector<double>::const_iterator const_iterator;
        size_t size() const { return (size_t)o_->numnodes; }
        double operator[](size_t i) const { return sync()[i]; }
        double at(size_t i) const { return sync().at(i); }
        const_iterator begin() const { return sync().begin(); }
        const_iterator end() const { return sync().end(); }
        operator const std::vector<double>&() const { return sync(); }
    private:
        friend class pl_calc_b200;
        pl_calc_b200* o_ = nullptr;
        int (*get_)(tpl_engine*, double*) = nullptr;
        mutable std::vector<double> cache_;
        mutable unsigned long seen_ = 0;
        const std::vector<double>& sync() const {
            if (seen_ != o_->version_ || cache_.size() != (size_t)o_->numnodes) {
                o_->ensure();
                cache_.assign(o_->numnodes, 0.0);
                check(get_(o_->e_, cache_.data()));
                seen_ = o_->version_;
            }
            return cache_;
        }
    };
    // reference: `double smoothing` (public, assigned by main())
    class smoothing_member {
    public:
        operator double() const { return v_; }
        smoothing_member& operator=(double s) { v_ = s; if (o_->e_) check(tpl_set_smoothing(o_->e_, s)); return *this; }
        smoothing_member& operator=(const smoothing_member& s) { return *this = (double)s; }
    private:
        friend class pl_calc_b200;
        pl_calc_b200* o_ = nullptr;
        double v_ = 1;
    };

    // ---- public data members of the reference (pl_calc_parallel.h:57-81) ----
    int numparams = 0;
    bool cvnodeset = false;
    member_vector dates, rates, durations;
    std::vector<int> freeparams;
    const std::vector<int>* free = nullptr;
    const std::vector<double>* min = nullptr;
    const std::vector<double>* max = nullptr;
    const std::vector<double>* pen_min = nullptr;
    const std::vector<double>* pen_max = nullptr;
    const std::vector<int>* parents_nds_ints = nullptr;
    const std::vector<int>* child_counts = nullptr;
    const std::vector<double>* char_durations = nullptr;
    const std::vector<double>* log_fact_char_durations = nullptr;
    const std::vector<std::vector<int> >* children_vec = nullptr;
    int numnodes = 0;
    int numsites = 0;
    smoothing_member smoothing;
    bool lf = false;
    bool paramverbose = false;      // kept for source compatibility; the GPU path never writes paramverbose files
    double minrate = 1e-10;
    double ftol = 1e-7;
    bool isConstrained = false;
    double penalty_boundary = 0.0025;
    // diagnostic hook (not in the reference): when set, every objective value handed back to a caller is appended
    std::vector<double>* trace = nullptr;

    pl_calc_b200() {
        rates.o_ = dates.o_ = durations.o_ = this;
        rates.get_ = tpl_get_rates; dates.get_ = tpl_get_dates; durations.get_ = tpl_get_durations;
        smoothing.o_ = this;
    }
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
        parents_nds_ints = parents_nds; child_counts = child_cnts; free = vfree; char_durations = char_dur;
        log_fact_char_durations = log_fact_ch;
        min = vmn; max = vmx; sd_ = start_dates; sr_ = start_rates; sdur_ = start_durations;
        numnodes = (int)parents_nds->size();
        cv_.assign(numnodes, 0);                 // :57-59
        cvnodeset = false;
        smoothing.v_ = 1;                        // :61
        log_pen_ = false;
        penalty_boundary = 0.0025;               // :69
        version_++;
    }
    void set_log_pen(bool v) { log_pen_ = v; if (e_) check(tpl_set_log_pen(e_, v)); }          // :1000
    void set_smoothing(double s) { smoothing = s; }
    double get_smoothing() const { return smoothing; }
    void set_cv_nodes(std::vector<int>& cv) {                                                    // :959-963
        ensure();
        check(tpl_set_cv_nodes(e_, cv.data()));
        cv_ = cv;
        cvnodeset = true;
        version_++;
    }
    std::vector<int>* get_cv_nodes() { return &cv_; }                                            // :956
    int calc_isum(std::vector<int>& inc) { int s = 0; for (size_t i = 0; i < inc.size(); i++) s += inc[i]; return s; }   // :913

    // pl_calc_parallel::set_freeparams (pl_calc_parallel.cpp:920-955)
    void set_freeparams(int nump, bool ilf, std::vector<int>* freep, std::vector<double>* params) {
        ensure();
        freeparams = *freep;
        numparams = nump;
        lf = ilf;
        params->assign(nump, 0.0);
        int k = check(tpl_set_freeparams(e_, nump, ilf ? 1 : 0, freep->data(), params->data()));
        params->resize(k);
        version_++;
    }

    // pl_calc_parallel::calc_pl (pl_calc_parallel.cpp:73-130)
    double calc_pl(std::vector<double>& params) {
        sync_members();
        version_++;
        return traced(ret(tpl_calc_pl(e_, params.data())));
    }
    // pl_calc_parallel::calc_gradient (:657-662)
    void calc_gradient(std::vector<double>& params, std::vector<double>* g) {
        sync_members();
        check(tpl_calc_gradient(e_, params.data(), g->data()));
    }
    void calc_lf_gradient(std::vector<double>& params, std::vector<double>* g) { calc_gradient(params, g); }
    void calc_pl_gradient(std::vector<double>& params, std::vector<double>* g) { calc_gradient(params, g); }
    // pl_calc_parallel::calc_function_gradient (:132-138) and its ADOL-C twins (:622-655, :224-333)
    double calc_function_gradient(std::vector<double>* params, std::vector<double>* g) {
        sync_members();
        return traced(ret(tpl_calc_function_gradient(e_, params->data(), g->data())));
    }
    double calc_pl_function_gradient(std::vector<double>* params, std::vector<double>* g) { return calc_function_gradient(params, g); }
    double calc_lf_function_gradient(std::vector<double>* params, std::vector<double>* g) { return calc_function_gradient(params, g); }
    double calc_pl_function_gradient_adolc(std::vector<double>* params, std::vector<double>* g) { return calc_function_gradient(params, g); }
    double calc_lf_function_gradient_adolc(std::vector<double>* params, std::vector<double>* g) { return calc_function_gradient(params, g); }
    // one round trip for function_plcp (optimize_tnc.cpp:75-105): f and the analytic gradient together
    double calc_pl_and_gradient(const double* x, double* g) {
        sync_members();
        version_++;
        return traced(ret(tpl_calc_pl_grad(e_, x, g, TPL_GRAD_ANALYTIC)));
    }

    // batched evaluation (folds / smoothing grid / restarts), X[p*B+b]
    void calc_pl_grad_batch(int B, const double* X, double* F, double* G, int mode) {
        sync_members();
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
        sync_members();
        tpl_opt_result res;
        check(tpl_optimize_batch(e_, B, X, F, lower, upper, mode, opts, status, iters, &res));
        return res;
    }

    // The fold loop of a CV run (main.cpp:689-801): vector b starts from column start_col[b] of Xstart [P][K] (the full-data
    // optimum of its smoothing value, main.cpp:679-681) and only the rows its chi-square needs come back:
    // probe_out[j][b] = X[probe_rows[j][b]][b] (main.cpp:737-793).  Truncated Newton (opts->method = 1) only.
    tpl_opt_result optimize_batch_cols(int B, int K, const double* Xstart, const int* start_col, double* F, const double* lower,
                                       const double* upper, int mode, const tpl_opt_options* opts, int nprobe, const int* probe_rows,
                                       double* probe_out, double* Xout = nullptr, int* status = nullptr, int* iters = nullptr) {
        sync_members();
        tpl_opt_result res;
        check(tpl_optimize_batch_cols(e_, B, K, Xstart, start_col, F, lower, upper, mode, opts, status, iters, &res, nprobe, probe_rows,
                                      probe_out, Xout));
        return res;
    }

    // B annealing chains at once (siman_calc_par::optimize, siman_calc_par.cpp:80-229): X [P][B] in/out, F [B] out
    void anneal_batch(int B, double* X, double* F, const tpl_anneal_options* opts = nullptr, int* status = nullptr,
                      int* iterations = nullptr) {
        sync_members();
        check(tpl_anneal_batch(e_, B, X, F, opts, status, iterations));
    }

    void delete_ad_arrays() { if (e_) { tpl_destroy(e_); e_ = nullptr; version_++; } }     // main.cpp:798
    tpl_engine* handle() { ensure(); return e_; }

private:
    tpl_engine* e_ = nullptr;
    const std::vector<double>*sd_ = nullptr, *sr_ = nullptr, *sdur_ = nullptr;
    std::vector<int> cv_;
    bool log_pen_ = false;
    double pb_sent_ = 0.0025;
    unsigned long version_ = 1;     // bumped by every call that may change the engine's rates / dates / durations

    static int check(int rc) {
        if (rc < 0) throw std::runtime_error(std::string("treepl_b200: ") + tpl_last_error());
        return rc;
    }
    static double ret(double f) {
        if (f != f && tpl_last_error()[0]) throw std::runtime_error(std::string("treepl_b200: ") + tpl_last_error());
        return f;
    }
    double traced(double f) { if (trace) trace->push_back(f); return f; }
    // public scalars a caller may have assigned since the last call
    void sync_members() {
        ensure();
        if (penalty_boundary != pb_sent_) { check(tpl_set_penalty_boundary(e_, penalty_boundary)); pb_sent_ = penalty_boundary; }
    }
    void ensure() {
        if (e_) return;
        if (!parents_nds_ints || !children_vec || !pen_min || !pen_max)
            throw std::runtime_error("treepl_b200: setup_starting_bits / children_vec / pen_min / pen_max not set");
        std::vector<int> flat;
        for (const auto& ch : *children_vec) flat.insert(flat.end(), ch.begin(), ch.end());
        if (flat.empty()) flat.push_back(0);
        check(tpl_create(&e_, -1, numnodes, parents_nds_ints->data(), child_counts->data(), flat.data(), free->data(),
                         char_durations->data(), log_fact_char_durations->data(), min->data(), max->data(), pen_min->data(),
                         pen_max->data(), sd_->data(), sr_->data(), sdur_->data()));
        check(tpl_set_smoothing(e_, smoothing));
        check(tpl_set_log_pen(e_, log_pen_ ? 1 : 0));
        check(tpl_set_penalty_boundary(e_, penalty_boundary));
        pb_sent_ = penalty_boundary;
    }
};
