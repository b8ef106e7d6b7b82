/* TEST INFRASTRUCTURE ONLY — never linked into, imported by or called from the
 * product path (treepl_b200/).  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load the library this
 * file builds (oracle/_ref/libtreepl_ref.so).
 *
 * What it is: a thin extern "C" harness around the UNMODIFIED reference objects
 * (compiled where they lie under /root/reference/src by oracle/Makefile).  It
 * replays the setup sequence of the reference's main() (main.cpp:318-497, the
 * recipe of SURVEY.md Appendix B) and exposes the reference's own
 *   pl_calc_parallel::calc_pl                  (pl_calc_parallel.cpp:73)
 *   pl_calc_parallel::calc_gradient            (pl_calc_parallel.cpp:657)
 *   pl_calc_parallel::calc_function_gradient   (pl_calc_parallel.cpp:132, RAD tape)
 *   generate_param_order_vector / set_freeparams / set_cv_nodes
 *   optimize_full (utils.cpp:671: annealer + TNC rounds) inside a replay of main()'s
 *   Langley-Fitch fit and cross-validation loop (ref_cv below)
 * so that tests can (1) pin the C restatement in oracle/pl_oracle.c and the
 * golden vectors under tests/golden/, and (2) time the reference's CPU path on
 * the GPU box's host cores (bench.py cpu_baseline.kind == "reference").
 *
 * NLopt 2.4.2 and ADOL-C 2.6.3 are built from the reference's own deps/ tarballs (oracle/build_deps.sh), so the
 * reference's optimize_nlopt.cpp and its ADOL-C taping paths are the real ones:
 *   pl_calc_parallel::calc_pl_function_gradient_adolc   (pl_calc_parallel.cpp:622)
 *   pl_calc_parallel::calc_lf_function_gradient_adolc   (pl_calc_parallel.cpp:224)
 */
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <limits>
#include <map>
#include <sstream>
#include <streambuf>
#include <string>
#include <vector>
#include <chrono>
#include <omp.h>
#include <adolc/adolc_openmp.h>

#include "node.h"
#include "tree.h"
#include "tree_reader.h"
#include "utils.h"
#include "pl_calc_parallel.h"
#include "optimize_nlopt.h"
#include "optimize_tnc.h"
#include "optim_options.h"

using namespace std;

namespace {

struct NullBuf : public streambuf { int overflow(int c) override { return c; } };

/* the reference chats on cout; silence it while we are inside reference code */
struct CoutMute {
    streambuf* old; NullBuf nb;
    CoutMute() { old = cout.rdbuf(&nb); }
    ~CoutMute() { cout.rdbuf(old); }
};

struct RefProblem {
    Tree* tree = nullptr;
    vector<int> free, parents, child_counts;
    vector<double> c, lf, vmin, vmax, penmin, penmax, sdates, srates, sdur;
    vector<vector<int> > children;
    vector<string> names;   /* by node number; "" for unnamed */
    pl_calc_parallel plp;
    vector<int> freeparams;
    vector<int> cv;
    bool log_pen = false;
    int numparams = 0;
    bool islf = false;

    void wire() {
        plp.setup_starting_bits(&parents, &child_counts, &free, &c, &lf, &vmin, &vmax,
                                &sdates, &srates, &sdur);          /* main.cpp:483-485 */
        plp.set_log_pen(log_pen);                                    /* main.cpp:486 */
        plp.minrate = 0.0;
        plp.pen_min = &penmin;                                       /* main.cpp:488-490 */
        plp.pen_max = &penmax;
        plp.children_vec = &children;
        plp.paramverbose = false;
    }
};

void wire_clone(RefProblem* src, pl_calc_parallel& q) {
    q.setup_starting_bits(&src->parents, &src->child_counts, &src->free, &src->c, &src->lf,
                          &src->vmin, &src->vmax, &src->sdates, &src->srates, &src->sdur);
    q.set_log_pen(src->log_pen);
    q.minrate = 0.0;
    q.pen_min = &src->penmin;
    q.pen_max = &src->penmax;
    q.children_vec = &src->children;
    q.paramverbose = false;
    q.smoothing = src->plp.smoothing;
    if (!src->cv.empty()) q.set_cv_nodes(src->cv);
    vector<int> fp(src->freeparams);
    vector<double> x;
    q.set_freeparams(src->numparams, src->islf, &fp, &x);
}

}  // namespace

extern "C" {

/* Build a problem exactly as the reference's main() does (main.cpp:318-490).
 * Calibrations: cal_tips[k] = space-separated tip names whose MRCA is calibrated. */
void* ref_from_newick(const char* newick, int numsites, int collapse, int set1,
                      int ncal, const char** cal_tips, const int* has_min, const double* cmin,
                      const int* has_max, const double* cmax, int seed, int log_pen) {
    CoutMute mute;
    srand(seed);                                                    /* main.cpp:287-293 */
    RefProblem* p = new RefProblem();
    p->log_pen = log_pen != 0;
    TreeReader tr;
    Tree* tree = tr.readTree(string(newick));                       /* main.cpp:322 */
    p->tree = tree;
    process_initial_branch_lengths(tree, collapse != 0, set1 != 0, numsites);   /* :326 */
    tree->setHeightFromTipToNodes();                                /* :327 */
    map<Node*, double> mins, maxs;
    /* mins first, then maxs, as main.cpp:336-387 */
    for (int k = 0; k < ncal; k++) {
        if (!has_min[k]) continue;
        vector<string> tips; Tokenize(string(cal_tips[k]), tips, " ");
        Node* m = tree->getMRCA(tips);
        if (!m) { delete p; return nullptr; }
        mins[m] = cmin[k]; m->min = cmin[k]; m->minb = true; m->pen_minb = true; m->pen_min = cmin[k];
    }
    for (int k = 0; k < ncal; k++) {
        if (!has_max[k]) continue;
        vector<string> tips; Tokenize(string(cal_tips[k]), tips, " ");
        Node* m = tree->getMRCA(tips);
        if (!m) { delete p; return nullptr; }
        maxs[m] = cmax[k]; m->max = cmax[k]; m->maxb = true;
        if (m->minb == true) {
            if (m->min != m->max) { m->pen_maxb = true; m->pen_max = cmax[k]; }
            else { m->pen_minb = false; m->pen_max = 0; }
        } else { m->pen_maxb = true; m->pen_max = cmax[k]; }
    }
    apply_node_label_preorder(tree);                                /* main.cpp:450 */
    calc_char_durations(tree, numsites);                            /* :452 */
    set_mins_maxs(tree, &mins, &maxs);                              /* :454 */
    setup_date_constraints(tree, &mins, &maxs);                     /* :456 */
    extract_tree_info(tree, &p->free, &p->parents, &p->child_counts, &p->c, &p->lf,
                      &p->vmin, &p->vmax, &p->children, &p->penmin, &p->penmax);   /* :467 */
    set_node_order(tree);                                           /* :472 */
    get_feasible_start_dates(tree, &p->sdates, &p->srates, &p->sdur);   /* :474 */
    double start_rate = get_start_rate(tree, &p->sdur);             /* :475 */
    p->srates[1] = start_rate;                                      /* :477 */
    int n = (int)p->parents.size();
    p->names.assign(n, "");
    for (int i = 0; i < tree->getNodeCount(); i++) {
        Node* nd = tree->getNode(i);
        p->names[nd->getNumber()] = nd->getName();
    }
    p->wire();
    return p;
}

/* Build from flat arrays (bypasses the reference's O(N^2) tree setup). */
void* ref_from_flat(int n, const int* parents, const int* child_counts, const int* children_flat,
                    const int* free, const double* c, const double* lf, const double* vmin,
                    const double* vmax, const double* penmin, const double* penmax,
                    const double* sdates, const double* srates, const double* sdur, int log_pen) {
    RefProblem* p = new RefProblem();
    p->log_pen = log_pen != 0;
    p->parents.assign(parents, parents + n);
    p->child_counts.assign(child_counts, child_counts + n);
    p->free.assign(free, free + n);
    p->c.assign(c, c + n); p->lf.assign(lf, lf + n);
    p->vmin.assign(vmin, vmin + n); p->vmax.assign(vmax, vmax + n);
    p->penmin.assign(penmin, penmin + n); p->penmax.assign(penmax, penmax + n);
    p->sdates.assign(sdates, sdates + n); p->srates.assign(srates, srates + n);
    p->sdur.assign(sdur, sdur + n);
    p->children.resize(n);
    int k = 0;
    for (int i = 0; i < n; i++) {
        p->children[i].assign(children_flat + k, children_flat + k + child_counts[i]);
        k += child_counts[i];
    }
    p->names.assign(n, "");
    p->wire();
    return p;
}

void ref_free(void* h) {
    RefProblem* p = (RefProblem*)h;
    if (!p) return;
    p->plp.delete_ad_arrays();
    delete p->tree;
    delete p;
}

int ref_numnodes(void* h) { return (int)((RefProblem*)h)->parents.size(); }

/* which: 0 parents, 1 child_counts, 2 free, 3 children (flattened in node order), 4 freeparams(2N) */
int ref_get_ints(void* h, int which, int* out) {
    RefProblem* p = (RefProblem*)h;
    const vector<int>* v = nullptr;
    vector<int> flat;
    switch (which) {
        case 0: v = &p->parents; break;
        case 1: v = &p->child_counts; break;
        case 2: v = &p->free; break;
        case 3: for (auto& ch : p->children) for (int x : ch) flat.push_back(x); v = &flat; break;
        case 4: v = &p->freeparams; break;
        default: return -1;
    }
    if (out) memcpy(out, v->data(), v->size() * sizeof(int));
    return (int)v->size();
}

/* which: 0 c, 1 lf, 2 min, 3 max, 4 pen_min, 5 pen_max, 6 start_dates, 7 start_rates,
 *        8 start_durations, 9 plp.rates, 10 plp.dates, 11 plp.durations */
int ref_get_doubles(void* h, int which, double* out) {
    RefProblem* p = (RefProblem*)h;
    const vector<double>* v = nullptr;
    switch (which) {
        case 0: v = &p->c; break;       case 1: v = &p->lf; break;
        case 2: v = &p->vmin; break;    case 3: v = &p->vmax; break;
        case 4: v = &p->penmin; break;  case 5: v = &p->penmax; break;
        case 6: v = &p->sdates; break;  case 7: v = &p->srates; break;
        case 8: v = &p->sdur; break;    case 9: v = &p->plp.rates; break;
        case 10: v = &p->plp.dates; break; case 11: v = &p->plp.durations; break;
        default: return -1;
    }
    if (out) memcpy(out, v->data(), v->size() * sizeof(double));
    return (int)v->size();
}

/* node name by node number (tips have names); returns length */
int ref_get_name(void* h, int node, char* out, int cap) {
    RefProblem* p = (RefProblem*)h;
    const string& s = p->names[node];
    if (out && cap > 0) { strncpy(out, s.c_str(), cap - 1); out[cap - 1] = 0; }
    return (int)s.size();
}

void ref_set_smoothing(void* h, double s) { ((RefProblem*)h)->plp.smoothing = s; }
void ref_set_log_pen(void* h, int lp) { RefProblem* p = (RefProblem*)h; p->log_pen = lp != 0; p->plp.set_log_pen(lp != 0); }
void ref_set_penalty_boundary(void* h, double pb) { ((RefProblem*)h)->plp.penalty_boundary = pb; }

/* overwrite the engine's current rates/dates (what the CV loop does through
 * setup_starting_bits with the snapshot vectors, main.cpp:701-704) */
void ref_set_state(void* h, const double* rates, const double* dates) {
    RefProblem* p = (RefProblem*)h;
    int n = (int)p->parents.size();
    for (int i = 0; i < n; i++) { p->plp.rates[i] = rates[i]; p->plp.dates[i] = dates[i]; }
}

/* cv == NULL clears the mask */
void ref_set_cv(void* h, const int* cv) {
    RefProblem* p = (RefProblem*)h;
    int n = (int)p->parents.size();
    if (cv) p->cv.assign(cv, cv + n); else p->cv.assign(n, 0);
    p->plp.set_cv_nodes(p->cv);                                     /* main.cpp:711 */
}

/* generate_param_order_vector (utils.cpp:295) + set_freeparams (pl_calc_parallel.cpp:920).
 * x_out may be NULL; returns numparams. */
int ref_set_freeparams(void* h, int lf, int use_cv, double* x_out) {
    RefProblem* p = (RefProblem*)h;
    p->freeparams.clear();
    vector<int>* cvp = (use_cv && !p->cv.empty()) ? &p->cv : nullptr;
    int np = generate_param_order_vector(&p->freeparams, lf != 0, cvp, &p->free);
    vector<double> x;
    vector<int> fp(p->freeparams);
    p->plp.set_freeparams(np, lf != 0, &fp, &x);
    p->numparams = np; p->islf = lf != 0;
    if (x_out) memcpy(x_out, x.data(), x.size() * sizeof(double));
    return np;
}

double ref_calc_pl(void* h, const double* x) {
    RefProblem* p = (RefProblem*)h;
    vector<double> v(x, x + p->numparams);
    return p->plp.calc_pl(v);
}

/* analytic gradient; the reference requires calc_pl(x) to have run first (it reads
 * member state), so this harness does both, as function_plcp does (optimize_tnc.cpp:83-93) */
double ref_calc_pl_and_gradient(void* h, const double* x, double* g) {
    RefProblem* p = (RefProblem*)h;
    vector<double> v(x, x + p->numparams), gv(p->numparams, 0.0);
    double f = p->plp.calc_pl(v);
    p->plp.calc_gradient(v, &gv);
    memcpy(g, gv.data(), gv.size() * sizeof(double));
    return f;
}

/* RAD reverse-mode f and gradient (pl_calc_parallel.cpp:132). g is pre-filled with NaN so
 * an early `return LARGE` (which leaves g untouched in the reference) is visible. */
double ref_calc_function_gradient(void* h, const double* x, double* g) {
    RefProblem* p = (RefProblem*)h;
    CoutMute mute;
    vector<double> v(x, x + p->numparams), gv(p->numparams, numeric_limits<double>::quiet_NaN());
    double f = p->plp.calc_function_gradient(&v, &gv);
    memcpy(g, gv.data(), gv.size() * sizeof(double));
    return f;
}

/* ADOL-C taped f and gradient: calc_pl_function_gradient_adolc (pl_calc_parallel.cpp:622-655) in PL mode,
 * calc_lf_function_gradient_adolc (:224-333) in LF mode — what function_plcp_ad_parallel / function_plcp_nlopt_ad_parallel
 * call (optimize_tnc.cpp:58-73, optimize_nlopt.cpp:88-105).  g pre-filled with NaN as above. */
double ref_calc_function_gradient_adolc(void* h, const double* x, double* g) {
    RefProblem* p = (RefProblem*)h;
    CoutMute mute;
    vector<double> v(x, x + p->numparams), gv(p->numparams, numeric_limits<double>::quiet_NaN());
    double f = p->islf ? p->plp.calc_lf_function_gradient_adolc(&v, &gv) : p->plp.calc_pl_function_gradient_adolc(&v, &gv);
    memcpy(g, gv.data(), gv.size() * sizeof(double));
    return f;
}

/* Wall seconds for `reps` calls of calc_pl (+ calc_gradient if with_grad) per thread on
 * `nthreads` independent pl_calc_parallel instances (one per OpenMP thread, as the
 * reference's fold loop does, main.cpp:689-710).  Returns seconds; total evals = reps*nthreads. */
double ref_time_evals(void* h, const double* x, int reps, int nthreads, int with_grad) {
    RefProblem* p = (RefProblem*)h;
    if (nthreads < 1) nthreads = 1;
    vector<pl_calc_parallel> inst(nthreads);
    for (int t = 0; t < nthreads; t++) wire_clone(p, inst[t]);
    double sink = 0.0;
    auto run = [&](int r) {
#pragma omp parallel for num_threads(nthreads) reduction(+ : sink) schedule(static, 1)
        for (int t = 0; t < nthreads; t++) {
            vector<double> v(x, x + p->numparams), gv(p->numparams, 0.0);
            for (int k = 0; k < r; k++) {
                v[0] = x[0] * (1.0 + 1e-12 * k);     /* defeat hoisting; stays feasible */
                double f = inst[t].calc_pl(v);
                if (with_grad) { inst[t].calc_gradient(v, &gv); f += gv[k % p->numparams]; }
                sink += f;
            }
        }
    };
    run(1);   /* warm-up */
    auto t0 = chrono::steady_clock::now();
    run(reps);
    auto t1 = chrono::steady_clock::now();
    for (int t = 0; t < nthreads; t++) inst[t].delete_ad_arrays();
    if (sink == 12345.678) fprintf(stderr, "sink\n");
    return chrono::duration<double>(t1 - t0).count();
}

/* The reference's cross-validation run (main.cpp:497-822) replayed around the reference's OWN optimiser
 * stack: optimize_full (utils.cpp:671-763) = siman_calc_par annealing + optimize_best (TNC, tnc.c) rounds.
 *   1. Langley-Fitch fit of the full data (main.cpp:499-578)
 *   2. for each smoothing value (main.cpp:668-682): PL fit of the full data, warm-started from the previous state
 *   3. for each fold (main.cpp:689-801, OpenMP over folds with `nthreads`): prune the fold's tips (CV mask),
 *      re-optimise from the full-data optimum, chi-square of the pruned branches (LOOCV :737-757, random :769-793)
 * groups: CSR lists of tip node numbers per fold.  OptimOptions are the constructor defaults (optim_options.cpp)
 * with the TNC optimiser (opt = optad = 0; NLopt is not built here, see the stubs above).
 * The caller seeds libc rand() through ref_from_newick(seed) exactly as main() does; with nthreads = 1 the run
 * is then reproducible.  Returns 0, fills chisq[nlam], fullpl[nlam], seconds[3] = {LF, full fits, folds}. */
int ref_cv(void* h, int nlam, const double* lams, int ngroups, const int* goff, const int* gnodes, int randomcv,
           int nthreads, double* chisq_out, double* fullpl_out, double* seconds) {
    RefProblem* p = (RefProblem*)h;
    CoutMute mute;
    OptimOptions oopt;
    oopt.nthreads = nthreads < 1 ? 1 : nthreads;
    pl_calc_parallel& plp = p->plp;
    auto now = [] { return chrono::steady_clock::now(); };
    auto secs = [](chrono::steady_clock::time_point a, chrono::steady_clock::time_point b) {
        return chrono::duration<double>(b - a).count();
    };
    auto t0 = now();
    vector<int> freeparams;
    vector<double> params;
    int numparams = generate_param_order_vector(&freeparams, true, NULL, &p->free);
    plp.set_freeparams(numparams, true, &freeparams, &params);
    double init = plp.calc_pl(params);
    if (std::isnan(init) || std::isinf(init) || init == 1e15) return -1;
    optimize_full(plp, &params, &oopt, false);
    plp.calc_pl(params);
    auto t1 = now();
    vector<double> strates(p->srates.size(), params[0]), stdates(plp.dates), stdur(plp.durations);
    double tfull = 0.0, tfold = 0.0;
    const int n = (int)p->parents.size();
    for (int li = 0; li < nlam; li++) {
        auto a = now();
        freeparams.clear();
        numparams = generate_param_order_vector(&freeparams, false, NULL, &p->free);
        plp.smoothing = lams[li];
        plp.set_freeparams(numparams, false, &freeparams, &params);
        plp.calc_pl(params);
        optimize_full(plp, &params, &oopt, false);
        fullpl_out[li] = plp.calc_pl(params);
        stdates = plp.dates;
        stdur = plp.durations;
        strates = plp.rates;
        auto b = now();
        tfull += secs(a, b);
        double chisq = 0.0;
#pragma omp parallel for ADOLC_OPENMP_NC reduction(+ : chisq) num_threads(oopt.nthreads)
        for (int g = 0; g < ngroups; g++) {
            vector<int> cvnodes(n, 0);
            for (int j = goff[g]; j < goff[g + 1]; j++) cvnodes[gnodes[j]] = 1;
            pl_calc_parallel q;
            q.setup_starting_bits(&p->parents, &p->child_counts, &p->free, &p->c, &p->lf, &p->vmin, &p->vmax,
                                  &stdates, &strates, &stdur);
            q.set_log_pen(p->log_pen);
            q.children_vec = &p->children;
            q.pen_min = &p->penmin;
            q.pen_max = &p->penmax;
            q.smoothing = lams[li];
            q.paramverbose = false;
            q.set_cv_nodes(cvnodes);
            vector<int> fpcv;
            int npcv = generate_param_order_vector(&fpcv, false, &cvnodes, &p->free);
            vector<double> xcv;
            q.set_freeparams(npcv, false, &fpcv, &xcv);
            q.calc_pl(xcv);
            optimize_full(q, &xcv, &oopt, true);
            q.calc_pl(xcv);
            double tch = 0.0;
            for (int j = goff[g]; j < goff[g + 1]; j++) {
                const int tip = gnodes[j];
                const int par = p->parents[tip];
                double ratees = 0.0;
                if (par == 0) {
                    for (int v = 0; v < n; v++)
                        if (p->parents[v] == par && v != tip) ratees = q.rates[v];
                } else {
                    ratees = q.rates[par];
                }
                const double d = q.durations[tip];
                const double expe = ratees * d;
                const double len = p->c[tip];
                const double sq = (len - expe) * (len - expe);
                if (!(d < 1e-100 || expe < 1e-100)) tch += sq / expe;
            }
            chisq += randomcv ? tch / (float)(goff[g + 1] - goff[g]) : tch;
            q.delete_ad_arrays();
        }
        chisq_out[li] = chisq;
        tfold += secs(b, now());
    }
    seconds[0] = secs(t0, t1);
    seconds[1] = tfull;
    seconds[2] = tfold;
    return 0;
}

/* The reference's optimiser adapters on the reference's own class — the CPU side of the drop-in comparison
 * (tests/test_dropin*.py; same `which` codes as tests/dropin/dropin_harness.cpp):
 *   0 optimize_plcp_tnc(ad=false)  1 optimize_plcp_tnc(ad=true)  2 optimize_plcp_tnc_ad_parallel      (optimize_tnc.cpp:107-231)
 *   10+k optimize_plcp_nlopt(whichone=k, ad=false)  20+k (ad=true)  30+k optimize_plcp_nlopt_ad_parallel (optimize_nlopt.cpp:116-341)
 *   100 optimize_full(cv=false), default OptimOptions, srand(seed) first                               (utils.cpp:671-763)
 * Starts from x (numparams doubles, in/out) after a calc_pl(x) as main() does; *f_out = calc_pl(final x).
 * The adapters' chatter ("nfeval: N rc: ...") is returned in log. */
int ref_optimize(void* h, int which, int numiter, double ftol, double xtol, int seed, double* x, double* f_out,
                 char* log, int logcap) {
    RefProblem* p = (RefProblem*)h;
    ostringstream os;
    streambuf* old = cout.rdbuf(os.rdbuf());
    pl_calc_parallel& plp = p->plp;
    vector<double> params(x, x + p->numparams);
    plp.calc_pl(params);
    int rc = 0;
    if (which == 100) {
        srand(seed);
        OptimOptions oo;
        optimize_full(plp, &params, &oo, false);
    } else {
        vector<double> x2(params);
        if (which == 0) rc = optimize_plcp_tnc(x2.data(), &plp, numiter, false, ftol);
        else if (which == 1) rc = optimize_plcp_tnc(x2.data(), &plp, numiter, true, ftol);
        else if (which == 2) rc = optimize_plcp_tnc_ad_parallel(x2.data(), &plp, ftol);
        else if (which >= 10 && which < 20) rc = optimize_plcp_nlopt(x2.data(), &plp, numiter, which - 10, false, false, ftol, xtol);
        else if (which >= 20 && which < 30) rc = optimize_plcp_nlopt(x2.data(), &plp, numiter, which - 20, true, false, ftol, xtol);
        else if (which >= 30 && which < 40) rc = optimize_plcp_nlopt_ad_parallel(x2.data(), &plp, numiter, which - 30, false, ftol, xtol);
        else rc = -999;
        params = x2;
    }
    *f_out = plp.calc_pl(params);
    memcpy(x, params.data(), params.size() * sizeof(double));
    cout.rdbuf(old);
    if (log && logcap > 0) { strncpy(log, os.str().c_str(), logcap - 1); log[logcap - 1] = 0; }
    return rc;
}

/* Multi-start on the host cores: nstarts independent optimisations with the reference's optimize_full (utils.cpp:671-763,
 * default OptimOptions), one pl_calc_parallel per OpenMP thread exactly as the reference runs its CV folds
 * (main.cpp:689-710, same ADOLC_OPENMP_NC clause).  Start k = X0[k*numparams ...].  -> f_out[k] = calc_pl(final x of start k); returns seconds. */
double ref_multistart(void* h, int nstarts, const double* X0, int nthreads, int seed, double* f_out) {
    RefProblem* p = (RefProblem*)h;
    CoutMute mute;
    srand(seed);
    const int np = p->numparams;
    auto t0 = chrono::steady_clock::now();
#pragma omp parallel for ADOLC_OPENMP_NC schedule(dynamic, 1) num_threads(nthreads < 1 ? 1 : nthreads)
    for (int k = 0; k < nstarts; k++) {
        pl_calc_parallel q;
        wire_clone(p, q);
        vector<double> x(X0 + (size_t)k * np, X0 + (size_t)(k + 1) * np);
        q.calc_pl(x);
        OptimOptions oo;
        /* cv = true selects the thread-parallel ADOL-C adapters (optimize_best_parallel), the only gradient path the
         * reference itself runs inside an OpenMP region (main.cpp:689-722); the RAD tape of cv = false is global state */
        optimize_full(q, &x, &oo, true);
        f_out[k] = q.calc_pl(x);
        q.delete_ad_arrays();
    }
    return chrono::duration<double>(chrono::steady_clock::now() - t0).count();
}

int ref_hw_threads(void) { return omp_get_num_procs(); }

}  /* extern "C" */
