// Batched simulated annealing (sm_100a): one chain per thread, O(degree) delta evaluation per proposal.
// Chain logic and what it replaces in the reference (siman_calc_par::optimize, src/siman_calc_par.cpp:80-229):
// tpl_anneal_core.h.  Proposals of a chain are inherently sequential, so the parallelism is across chains (folds, smoothing
// values, restarts); a chain touches ~10 scattered doubles per proposal, i.e. the kernel is L2-latency bound and 1,024 chains
// cost about as much wall time as one.
#pragma once
#include "tpl_anneal_core.h"
#include "tpl_kernels.cuh"

namespace tplsa {

struct SaArgs {
    SaConst o;
    tpl::DevTree t;
    const int* date_node;                    // [ndates] node of the k-th free-date row
    const int* bar_node;                     // [nbar] node of the k-th calibration barrier row
    const double *hmin, *hmax;               // [n] date bounds (set_durations :887-896)
    const unsigned long long* lanemask;      // per-vector CV masks (EvalArgs layout) or null
    int mask_stride;
    const double* lane_lambda;               // [B] or null
    double lambda;
    double* X;                               // [P][B]
    Chain* chains;                           // [B]
    int* done;                               // [B] 1 = finished, 2 = infeasible start
    int* ndone;
    int B;
};

struct SaView {
    const SaArgs& a;
    int b;
    double lam;
    int nbar;
    __device__ int parent(int i) const { return a.t.parent[i]; }
    __device__ int child_begin(int i) const { return a.t.child_off[i]; }
    __device__ int child_end(int i) const { return a.t.child_off[i + 1]; }
    __device__ int child(int j) const { return a.t.children[j]; }
    __device__ int rslot(int i) const { return a.t.rslot[i]; }
    __device__ int dslot(int i) const { return a.t.dslot[i]; }
    __device__ int date_node(int k) const { return a.date_node[k]; }
    __device__ double c(int i) const { return a.t.c[i]; }
    __device__ double lf(int i) const { return a.t.lf[i]; }
    __device__ double hfix(int i) const { return a.t.hfix[i]; }
    __device__ double rfix(int i) const { return a.t.rfix[i]; }
    __device__ double hmin(int i) const { return a.hmin[i]; }
    __device__ double hmax(int i) const { return a.hmax[i]; }
    __device__ bool masked(int i) const {
        if (!a.lanemask) return false;
        return (a.lanemask[(size_t)(b >> 6) * a.mask_stride + i] >> (b & 63)) & 1ull;
    }
    __device__ double x(int row) const { return a.X[(size_t)row * a.B + b]; }
    __device__ void setx(int row, double v) { a.X[(size_t)row * a.B + b] = v; }
    __device__ int bar_node(int k) const { return a.bar_node[k]; }
    __device__ double bar_pmin(int k) const { return a.t.bar_pmin[k]; }
    __device__ double bar_pmax(int k) const { return a.t.bar_pmax[k]; }
};

__global__ void __launch_bounds__(32) sa_init_kernel(const __grid_constant__ SaArgs a, double temperature, double step_rt, double step_dt,
                                                      unsigned long long seed) {
    const int b = blockIdx.x * 32 + threadIdx.x;
    if (b >= a.B) return;
    Chain ch = Chain();
    ch.rng = seed * 0x9E3779B97F4A7C15ull + (unsigned long long)(b + 1) * 0xD1B54A32D192ED03ull;
    if (ch.rng == 0) ch.rng = 1;
    ch.temperature = temperature;
    ch.step_rt = step_rt;
    ch.step_dt = step_dt;
    SaView V = {a, b, a.lane_lambda ? a.lane_lambda[b] : a.lambda, a.t.nbar};
    const bool ok = sa_init(V, ch, a.o);
    a.chains[b] = ch;
    a.done[b] = ok ? 0 : 2;
    if (!ok) atomicAdd(a.ndone, 1);
}

// up to `proposals` proposals per chain and launch (bounded launches keep the kernel far below any watchdog)
__global__ void __launch_bounds__(32) sa_run_kernel(const __grid_constant__ SaArgs a, int proposals) {
    const int b = blockIdx.x * 32 + threadIdx.x;
    if (b >= a.B) return;
    if (a.done[b]) return;
    Chain ch = a.chains[b];
    SaView V = {a, b, a.lane_lambda ? a.lane_lambda[b] : a.lambda, a.t.nbar};
    bool fin = false;
    for (int k = 0; k < proposals && !fin; k++)
        if (sa_step(V, ch, a.o)) fin = sa_finished(ch, a.o);
    a.chains[b] = ch;
    if (fin) {
        a.done[b] = 1;
        atomicAdd(a.ndone, 1);
    }
}

}  // namespace tplsa
