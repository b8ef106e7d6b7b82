// TEST INFRASTRUCTURE ONLY — an ABI-conformance double: the evaluation subset of the C ABI of include/treepl_b200.h
// implemented on top of the UNMODIFIED reference class (`pl_calc_parallel`, /root/reference/src/pl_calc_parallel.{h,cpp},
// objects built by oracle/Makefile).  It lets the SAME host code — the C++ mirror treepl_b200/host/pl_calc_b200.h plus the
// reference's own optimiser adapters compiled against it (tests/dropin/Makefile) — run once over the GPU engine
// (libtreepl_b200.so) and once over the reference's CPU arithmetic (this file), so that the two runs can be diffed call
// by call (tests/test_dropin*.py).  Never linked into or loaded by the product package.
#include <cmath>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#include "pl_calc_parallel.h"                      // the REFERENCE header (this TU is compiled without the shim)
#include "../../include/treepl_b200.h"

using namespace std;

struct tpl_engine {
    vector<int> parents, child_counts, free_;
    vector<double> c, lf, mn, mx, pmn, pmx, sd, sr, sdur;
    vector<vector<int> > children;
    vector<int> cv;
    pl_calc_parallel plp;
    int numparams = 0;
    bool have_fp = false, called = false;
    vector<double> lastx;
};

static thread_local string g_err;
static int fail(int code, const char* msg) { g_err = msg; return code; }

extern "C" {

const char* tpl_last_error(void) { return g_err.c_str(); }

int tpl_create(tpl_engine** out, int, int n, const int* parents, const int* child_counts, const int* children,
               const int* free_, const double* c, const double* lf, const double* mn, const double* mx,
               const double* pen_min, const double* pen_max, const double* sd, const double* sr, const double* sdur) {
    tpl_engine* e = new tpl_engine();
    e->parents.assign(parents, parents + n); e->child_counts.assign(child_counts, child_counts + n);
    e->free_.assign(free_, free_ + n); e->c.assign(c, c + n); e->lf.assign(lf, lf + n);
    e->mn.assign(mn, mn + n); e->mx.assign(mx, mx + n); e->pmn.assign(pen_min, pen_min + n); e->pmx.assign(pen_max, pen_max + n);
    e->sd.assign(sd, sd + n); e->sr.assign(sr, sr + n); e->sdur.assign(sdur, sdur + n);
    e->children.resize(n);
    int k = 0;
    for (int i = 0; i < n; i++) { e->children[i].assign(children + k, children + k + child_counts[i]); k += child_counts[i]; }
    e->plp.setup_starting_bits(&e->parents, &e->child_counts, &e->free_, &e->c, &e->lf, &e->mn, &e->mx, &e->sd, &e->sr, &e->sdur);
    e->plp.pen_min = &e->pmn; e->plp.pen_max = &e->pmx; e->plp.children_vec = &e->children;   // main.cpp:488-490
    e->plp.paramverbose = false;
    *out = e;
    return TPL_OK;
}
void tpl_destroy(tpl_engine* e) { if (e) { e->plp.delete_ad_arrays(); delete e; } }
int tpl_set_smoothing(tpl_engine* e, double s) { e->plp.smoothing = s; return TPL_OK; }
int tpl_set_log_pen(tpl_engine* e, int lp) { e->plp.set_log_pen(lp != 0); return TPL_OK; }
int tpl_set_penalty_boundary(tpl_engine* e, double pb) { e->plp.penalty_boundary = pb; return TPL_OK; }
int tpl_set_cv_nodes(tpl_engine* e, const int* cv) {
    if (cv) e->cv.assign(cv, cv + e->parents.size()); else e->cv.assign(e->parents.size(), 0);
    e->plp.set_cv_nodes(e->cv);
    return TPL_OK;
}
int tpl_set_state(tpl_engine* e, const double* rates, const double* dates) {
    for (size_t i = 0; i < e->parents.size(); i++) { e->plp.rates[i] = rates[i]; e->plp.dates[i] = dates[i]; }
    return TPL_OK;
}
int tpl_set_freeparams(tpl_engine* e, int nump, int lf, const int* freep, double* params_out) {
    vector<int> fp(freep, freep + 2 * e->parents.size());
    vector<double> x;
    e->plp.set_freeparams(nump, lf != 0, &fp, &x);
    e->numparams = nump;
    e->have_fp = true;
    if (params_out) for (size_t i = 0; i < x.size() && (int)i < nump; i++) params_out[i] = x[i];
    return (int)x.size();
}
int tpl_numnodes(const tpl_engine* e) { return (int)e->parents.size(); }
int tpl_numparams(const tpl_engine* e) { return e->numparams; }
double tpl_calc_pl(tpl_engine* e, const double* x) {
    if (!e->have_fp) { fail(TPL_ERR_STATE, "set_freeparams has not been called"); return numeric_limits<double>::quiet_NaN(); }
    e->lastx.assign(x, x + e->numparams);
    e->called = true;
    return e->plp.calc_pl(e->lastx);
}
int tpl_calc_gradient(tpl_engine* e, const double*, double* g) {
    if (!e->called) return fail(TPL_ERR_STATE, "calc_gradient needs a preceding calc_pl");
    vector<double> gv(e->numparams, 0.0);
    e->plp.calc_gradient(e->lastx, &gv);
    memcpy(g, gv.data(), sizeof(double) * e->numparams);
    return TPL_OK;
}
double tpl_calc_function_gradient(tpl_engine* e, const double* x, double* g) {
    vector<double> xv(x, x + e->numparams), gv(g, g + e->numparams);      // g untouched on the early return
    const double f = e->plp.calc_function_gradient(&xv, &gv);
    memcpy(g, gv.data(), sizeof(double) * e->numparams);
    return f;
}
double tpl_calc_pl_grad(tpl_engine* e, const double* x, double* g, int mode) {
    if (mode == TPL_GRAD_AD) return tpl_calc_function_gradient(e, x, g);
    const double f = tpl_calc_pl(e, x);
    if (g) tpl_calc_gradient(e, x, g);
    return f;
}
int tpl_get_rates(tpl_engine* e, double* out) { memcpy(out, e->plp.rates.data(), sizeof(double) * e->parents.size()); return TPL_OK; }
int tpl_get_dates(tpl_engine* e, double* out) { memcpy(out, e->plp.dates.data(), sizeof(double) * e->parents.size()); return TPL_OK; }
int tpl_get_durations(tpl_engine* e, double* out) { memcpy(out, e->plp.durations.data(), sizeof(double) * e->parents.size()); return TPL_OK; }

// batched entry points: not part of the reference; the double refuses them loudly
int tpl_batch_configure(tpl_engine*, int, const double*, const int*, const int*) { return fail(TPL_ERR_STATE, "CPU conformance double: no batched path"); }
int tpl_calc_pl_grad_batch(tpl_engine*, int, const double*, double*, double*, int) { return fail(TPL_ERR_STATE, "CPU conformance double: no batched path"); }
int tpl_optimize_batch(tpl_engine*, int, double*, double*, const double*, const double*, int, const tpl_opt_options*, int*, int*, tpl_opt_result*) { return fail(TPL_ERR_STATE, "CPU conformance double: no batched path"); }
int tpl_release_cached_memory(void) { return TPL_OK; }
int tpl_optimize_batch_cols(tpl_engine*, int, int, const double*, const int*, double*, const double*, const double*, int, const tpl_opt_options*, int*, int*, tpl_opt_result*, int, const int*, double*, double*) { return fail(TPL_ERR_STATE, "CPU conformance double: no batched path"); }
int tpl_anneal_batch(tpl_engine*, int, double*, double*, const tpl_anneal_options*, int*, int*) { return fail(TPL_ERR_STATE, "CPU conformance double: no batched path"); }

}  // extern "C"
