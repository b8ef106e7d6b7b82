// TEST INFRASTRUCTURE: CPU emulation of the batched annealer (treepl_b200/csrc/tpl_anneal.cuh): the chain logic and the
// delta evaluation are the SAME host/device functions the CUDA kernel calls (tpl_anneal_core.h); one loop per chain replaces
// one thread per chain.  tests/test_anneal_sim.py steps the chains and compares the running objective with the oracle's full
// calc_pl after every batch of proposals.  Not part of the product.
#include <cmath>
#include <vector>

#include "../treepl_b200/csrc/tpl_anneal_core.h"

using namespace tplsa;

struct Tree {
    int n;
    const int *parent, *child_off, *children, *rslot, *dslot, *date_node;
    const double *c, *lf, *hfix, *rfix, *hmin, *hmax;
    const unsigned char* mask;      // [B][n] or null
    int nbar;
    const int* bar_node;
    const double *bar_pmin, *bar_pmax;
};

struct HostView {
    const Tree& t;
    int B, b;
    double* X;
    double lam;
    int nbar;
    int parent(int i) const { return t.parent[i]; }
    int child_begin(int i) const { return t.child_off[i]; }
    int child_end(int i) const { return t.child_off[i + 1]; }
    int child(int j) const { return t.children[j]; }
    int rslot(int i) const { return t.rslot[i]; }
    int dslot(int i) const { return t.dslot[i]; }
    int date_node(int k) const { return t.date_node[k]; }
    double c(int i) const { return t.c[i]; }
    double lf(int i) const { return t.lf[i]; }
    double hfix(int i) const { return t.hfix[i]; }
    double rfix(int i) const { return t.rfix[i]; }
    double hmin(int i) const { return t.hmin[i]; }
    double hmax(int i) const { return t.hmax[i]; }
    bool masked(int i) const { return t.mask && t.mask[(size_t)b * t.n + i]; }
    double x(int row) const { return X[(size_t)row * B + b]; }
    void setx(int row, double v) { X[(size_t)row * B + b] = v; }
    int bar_node(int k) const { return t.bar_node[k]; }
    double bar_pmin(int k) const { return t.bar_pmin[k]; }
    double bar_pmax(int k) const { return t.bar_pmax[k]; }
};

struct Sim {
    Tree t;
    SaConst o;
    int B;
    double* X;
    std::vector<double> lam;
    std::vector<Chain> ch;
    std::vector<int> done;
};

extern "C" void* sa_sim_create(int P, int B, int nr, int n, const int* parent, const int* child_off, const int* children, const int* rslot,
                               const int* dslot, const int* date_node, int ndates, const double* c, const double* lf, const double* hfix,
                               const double* rfix, const double* hmin, const double* hmax, const unsigned char* mask, int nbar,
                               const int* bar_node, const double* bar_pmin, const double* bar_pmax, const double* lam, double pb, int log_pen,
                               double* X, double temp, double cool, double rt_step, double dt_step, int maxit, double stop_temp,
                               unsigned long long seed, int* ok) {
    Sim* s = new Sim();
    s->t = Tree{n, parent, child_off, children, rslot, dslot, date_node, c, lf, hfix, rfix, hmin, hmax, mask, nbar, bar_node, bar_pmin, bar_pmax};
    s->o.n = n; s->o.P = P; s->o.nr = nr; s->o.ndates = ndates; s->o.maxit = maxit; s->o.cool = cool; s->o.stop_temp = stop_temp; s->o.pb = pb;
    s->o.log_pen = log_pen;
    s->B = B;
    s->X = X;
    s->lam.assign(lam, lam + B);
    s->ch.resize(B);
    s->done.assign(B, 0);
    for (int b = 0; b < B; b++) {
        Chain& ch = s->ch[b];
        ch = Chain();
        ch.rng = seed * 0x9E3779B97F4A7C15ull + (unsigned long long)(b + 1) * 0xD1B54A32D192ED03ull;
        if (ch.rng == 0) ch.rng = 1;
        ch.temperature = temp; ch.step_rt = rt_step; ch.step_dt = dt_step;
        HostView V = {s->t, B, b, X, s->lam[b], nbar};
        ok[b] = sa_init(V, ch, s->o) ? 1 : 0;
        if (!ok[b]) s->done[b] = 1;
    }
    return s;
}

// up to `proposals` proposals per chain; returns the number of chains still running
extern "C" int sa_sim_steps(void* h, int proposals) {
    Sim* s = (Sim*)h;
    int running = 0;
    for (int b = 0; b < s->B; b++) {
        if (s->done[b]) continue;
        HostView V = {s->t, s->B, b, s->X, s->lam[b], s->t.nbar};
        Chain& ch = s->ch[b];
        for (int k = 0; k < proposals; k++) {
            if (sa_step(V, ch, s->o) && sa_finished(ch, s->o)) { s->done[b] = 1; break; }
        }
        if (!s->done[b]) running++;
    }
    return running;
}

// out: f, temperature, step_rt, step_dt, iteration, discarded, rate_accepts, rate_trials, date_accepts, date_trials
extern "C" void sa_sim_state(void* h, int b, double* out) {
    const Chain& ch = ((Sim*)h)->ch[b];
    out[0] = ch.f; out[1] = ch.temperature; out[2] = ch.step_rt; out[3] = ch.step_dt; out[4] = ch.iteration; out[5] = (double)ch.discarded;
    out[6] = (double)ch.rate_accepts; out[7] = (double)ch.rate_trials; out[8] = (double)ch.date_accepts; out[9] = (double)ch.date_trials;
}

extern "C" void sa_sim_free(void* h) { delete (Sim*)h; }
