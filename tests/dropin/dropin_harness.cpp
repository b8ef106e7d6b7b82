// TEST INFRASTRUCTURE ONLY.  Drives the reference's OWN optimiser adapters — optimize_plcp_tnc, optimize_plcp_tnc_ad_parallel
// (src/optimize_tnc.cpp:107-231), optimize_plcp_nlopt, optimize_plcp_nlopt_ad_parallel (src/optimize_nlopt.cpp:116-341) and
// optimize_full (src/utils.cpp:671-763), all compiled UNMODIFIED from /root/reference/src — on an engine object that is
// the C++ mirror treepl_b200/host/pl_calc_b200.h (this TU is compiled with treepl_b200/host/shim first on the include
// path, so `pl_calc_parallel` IS the mirror).  The library this file lands in is linked either against
// libtreepl_b200.so (GPU) or against tests/_build/libtpl_over_ref.so (the reference's CPU arithmetic behind the same C ABI).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "pl_calc_parallel.h"      // the shim -> pl_calc_b200
#include "optimize_tnc.h"
#include "optimize_nlopt.h"
#include "optim_options.h"
#include "utils.h"

namespace {
struct CoutCapture {
    std::streambuf* old;
    std::ostringstream os;
    CoutCapture() { old = std::cout.rdbuf(os.rdbuf()); }
    ~CoutCapture() { std::cout.rdbuf(old); }
};
}  // namespace

extern "C" {

// which: 0 optimize_plcp_tnc(ad=false)   1 optimize_plcp_tnc(ad=true)   2 optimize_plcp_tnc_ad_parallel
//        10+k optimize_plcp_nlopt(whichone=k, ad=false)   20+k optimize_plcp_nlopt(whichone=k, ad=true)
//        30+k optimize_plcp_nlopt_ad_parallel(whichone=k)
//        100 optimize_full(cv=false) with the default OptimOptions (srand(seed) first, as main.cpp:287-293)
// x: in = ignored when use_x0 == 0 (the start is what set_freeparams emits, as in main()), out = final vector.
// trace: objective value of every engine call the adapter made (capacity trace_cap); *ntrace = number of calls.
// Returns the adapter's return code (TNC / NLopt rc), -1000 on an exception (message in log).
int dropin_optimize(int n, const int* parents, const int* child_counts, const int* children, const int* free_,
                    const double* c, const double* lf, const double* mn, const double* mx, const double* pen_min,
                    const double* pen_max, const double* sd, const double* sr, const double* sdur,
                    int islf, double smoothing, int log_pen, const int* cvnodes, int which, int numiter, double ftol,
                    double xtol, int seed, int use_x0, double* x, int* numparams_out, double* f_out, double* trace,
                    int trace_cap, int* ntrace, double* rates_out, double* dates_out, char* log, int logcap) {
    std::vector<int> vpar(parents, parents + n), vcc(child_counts, child_counts + n), vfree(free_, free_ + n);
    std::vector<double> vc(c, c + n), vlf(lf, lf + n), vmn(mn, mn + n), vmx(mx, mx + n), vpmn(pen_min, pen_min + n),
        vpmx(pen_max, pen_max + n), vsd(sd, sd + n), vsr(sr, sr + n), vsdur(sdur, sdur + n);
    std::vector<std::vector<int> > kids(n);
    int k = 0;
    for (int i = 0; i < n; i++) { kids[i].assign(children + k, children + k + child_counts[i]); k += child_counts[i]; }
    std::vector<double> tr;
    CoutCapture cap;
    int rc = 0;
    try {
        pl_calc_parallel plp;                                               /* main.cpp:482-495 */
        plp.setup_starting_bits(&vpar, &vcc, &vfree, &vc, &vlf, &vmn, &vmx, &vsd, &vsr, &vsdur);
        plp.set_log_pen(log_pen != 0);
        plp.minrate = 0.0;
        plp.pen_min = &vpmn;
        plp.pen_max = &vpmx;
        plp.children_vec = &kids;
        plp.paramverbose = false;
        plp.smoothing = smoothing;
        std::vector<int> cv;
        if (cvnodes) { cv.assign(cvnodes, cvnodes + n); plp.set_cv_nodes(cv); }          /* main.cpp:711 */
        std::vector<int> fp;
        int np = generate_param_order_vector(&fp, islf != 0, cvnodes ? &cv : NULL, &vfree);   /* utils.cpp:295 */
        std::vector<double> params;
        plp.set_freeparams(np, islf != 0, &fp, &params);
        if (use_x0) params.assign(x, x + np);
        *numparams_out = np;
        plp.trace = &tr;
        plp.calc_pl(params);                                                 /* main.cpp:505 / 676 */
        if (which == 100) {
            srand(seed);
            OptimOptions oo;
            optimize_full(plp, &params, &oo, false);
        } else {
            std::vector<double> x2(params);
            if (which == 0) rc = optimize_plcp_tnc(x2.data(), &plp, numiter, false, ftol);
            else if (which == 1) rc = optimize_plcp_tnc(x2.data(), &plp, numiter, true, ftol);
            else if (which == 2) rc = optimize_plcp_tnc_ad_parallel(x2.data(), &plp, ftol);
            else if (which >= 10 && which < 20) rc = optimize_plcp_nlopt(x2.data(), &plp, numiter, which - 10, false, false, ftol, xtol);
            else if (which >= 20 && which < 30) rc = optimize_plcp_nlopt(x2.data(), &plp, numiter, which - 20, true, false, ftol, xtol);
            else if (which >= 30 && which < 40) rc = optimize_plcp_nlopt_ad_parallel(x2.data(), &plp, numiter, which - 30, false, ftol, xtol);
            else throw std::runtime_error("bad `which`");
            params = x2;
        }
        plp.trace = NULL;
        *f_out = plp.calc_pl(params);                                        /* main.cpp:678: "after opt calc" */
        for (int i = 0; i < np; i++) x[i] = params[i];
        if (rates_out) for (int i = 0; i < n; i++) rates_out[i] = plp.rates[i];          /* main.cpp:679-681 */
        if (dates_out) { std::vector<double> d(plp.dates); for (int i = 0; i < n; i++) dates_out[i] = d[i]; }
    } catch (const std::exception& ex) {
        snprintf(log, logcap, "EXCEPTION %s", ex.what());
        return -1000;
    }
    *ntrace = (int)tr.size();
    for (int i = 0; i < (int)tr.size() && i < trace_cap; i++) trace[i] = tr[i];
    if (log && logcap > 0) { strncpy(log, cap.os.str().c_str(), logcap - 1); log[logcap - 1] = 0; }
    return rc;
}

}  // extern "C"
