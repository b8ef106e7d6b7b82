// Streaming kernels of the batched truncated-Newton optimiser (sm_100a).  Method, math and what it replaces in the
// reference (TNC, src/tnc.c:483-827 behind src/optimize_tnc.cpp:107-172): tpl_tn_core.h.
//
// One optimiser step = one PL evaluation of the trial points (for the vectors in their line search) + `cg_per_step`
// CG iterations (analytic Hessian-vector product + vector updates, for the vectors inside their CG solve).  Vectors advance
// independently; each kernel skips the 32-vector groups that have nothing to do in it.  Every per-vector reduction
// (p.Hp, r.M^-1 r, gp.s, the ratio test) is finished by the LAST block of the vector group to arrive (fixed-order
// sums over the chunk partials, so the result does not depend on which block that is), and that block also runs the
// vector's state machine (tpl_tn_core.h::phase_*): no separate reduction or bookkeeping launches.
//
//   tn_diag_kernel    node-indexed   diagonal of H_z at a newly accepted point (Jacobi preconditioner)
//   tn_setup_kernel   row-indexed    accept the step: Z, Gz, M^-1, r = -gp, zz = M^-1 r, p = zz, s = 0      -> phase_setup
//   tn_hv_kernel      node-indexed   Hp = H_z p (pull: node, parent, children)                               -> phase_hv
//   tn_update_kernel  row-indexed    s += alpha p, r -= alpha Hp, zz = M^-1 r                                 -> phase_update
//   tn_dir_kernel     row-indexed    p = zz + beta p
//   tn_gd_kernel      row-indexed    gd = Gz.s  (s = p when CG stopped at its first iteration)
//   tn_stepmax_kernel node-indexed   largest step keeping all durations positive                             -> phase_finish
//   tn_trial_kernel   row-indexed    X = to_param(z + t s)
// Tree preconditioner (tpl_tn_core.h::tree_*_node; TnArgs::precond): zz = M^-1 r by one leaf-to-root and one root-to-leaf
// sweep over HEIGHT levels (a node's children sit on lower levels; level sizes never increase): one launch per wide
// level, and ONE block per 32-vector group for all the narrow top levels (block-level barriers between levels):
//   tn_tree_local_kernel         FACTOR only, all nodes at once: the node's own block of H_z and its coupling with the parent;
//                                for the childless nodes (level 0) also their pivot step and forward substitution
//   tn_tree_up_kernel<FACTOR>    one wide level: [pivot step: S_i, K_i, T_i] + forward substitution
//   tn_tree_top_kernel<FACTOR>   narrow levels: up to the root, then back down
//   tn_tree_down_kernel<FACTOR>  one wide level: back substitution -> ZZ rows
//   tn_pc_finish_kernel<FACTOR>  row-indexed    [p = zz] and r.zz                                        -> phase_setup / phase_update
// The node-indexed kernels are latency bound (one warp = one node's dependent chain): the sweeps read host-built index
// records through "record views" that fetch all operands of a node side by side, tn_hv_kernel has a hand-specialised fast
// path on its own records, and with 16 or fewer vectors a warp carries several nodes at once (node_lane).  DESIGN.md 8.1.
#pragma once
#include <cooperative_groups.h>

#include "tpl_kernels.cuh"
#include "tpl_lbfgs.cuh"
#include "tpl_tn_core.h"

namespace tpltn {

using tplopt::OPT_WY;

struct TnArgs {
    TnConst o;
    tpl::DevTree t;
    const double *pmin_n, *pmax_n;           // [n] node-indexed calibration penalties (-1 = none)
    const unsigned long long* lanemask;      // per-vector CV masks (EvalArgs layout) or null
    int mask_stride;
    const double* lane_lambda;               // [B] or null
    double lambda, pb;
    const double *sc, *lo, *hi;              // [P]
    double *Z, *Gz, *S, *X, *R, *ZZ, *Pv, *Hp, *Minv, *DG;   // [P][B]
    const double* Gx;                        // [P][B] gradient of the evaluated point, original parameters
    const double* Ft;                        // [B]
    int *st, *ls, *cgit, *flag, *stall, *iters, *ncg, *act;
    double *F0, *t_, *gd, *rz, *rz0, *tol2, *alpha, *beta, *gmax, *tlog, *gdtmp, *mu;
    int pack_shift;                          // node-indexed kernels at B <= 16: a warp carries 32 >> pack_shift nodes x (1 << pack_shift) vectors
    int row_chunks, rows_per_chunk;          // row-indexed kernels
    int node_chunks, nodes_per_chunk;        // node-indexed kernels
    int diag_chunks, diag_nodes_per_chunk;   // tn_diag_kernel and tn_stepmax_kernel need a third / half of the Hessian-vector kernel's registers:
    int tm_chunks, tm_nodes_per_chunk;       // their grids are sized by their OWN occupancy (these kernels are bound by warps in flight)
    double *part_setup, *part_hv, *part_upd, *part_gd, *part_tm;
    int *cnt_setup, *cnt_hv, *cnt_upd, *cnt_gd, *cnt_tm;      // [groups] arrival counters, zero between launches
    int* ndone;
    // tree preconditioner
    int precond;                             // 0 Jacobi, 1 block-tree factorisation
    int log_pen;
    double* TW;                              // [TW_PLANES][n][B] node-indexed factors and sweep scratch
    const int* lev_nodes;                    // [n] node numbers grouped by height level (tips first)
    const int* lev_off;                      // [nlev + 1]
    const int* lev_rec;                      // [n][TREE_REC_INTS] index record of node lev_nodes[k] (tpl_tn.cuh: record views)
    const int* hv_rec;                       // [n][HV_REC_INTS] index record of node i for hess_node_fast
    int hv_pipe;                             // host side: tn_hv_pipe_kernel instead of tn_hv_kernel
    const double* hv_const;                  // [n][HV_CONST] per-node constants of tn_hv_pipe_kernel (scales, c, calibration penalties, fixed dates)
    int nlev, nlev_wide;                     // levels [0, nlev_wide): one launch each; [nlev_wide, nlev): tn_tree_top_kernel
    int coop_levels;                         // host side: the wide levels of a sweep run as ONE cooperative launch (tn_tree_levels_kernel)
    double *gmax_tmp, *gg_tmp;               // [B] reductions of the set-up pass kept until the solve has produced r.zz
    double* part_pc;
    int* cnt_pc;
};

__device__ __forceinline__ Lane lane_of(const TnArgs& a, int b) {
    Lane L;
    L.st = a.st + b; L.ls = a.ls + b; L.cgit = a.cgit + b; L.flag = a.flag + b; L.stall = a.stall + b; L.iters = a.iters + b;
    L.ncg = a.ncg + b; L.act = a.act + b;
    L.F0 = a.F0 + b; L.t = a.t_ + b; L.gd = a.gd + b; L.rz = a.rz + b; L.rz0 = a.rz0 + b; L.tol2 = a.tol2 + b; L.alpha = a.alpha + b;
    L.beta = a.beta + b; L.gmax = a.gmax + b; L.tlog = a.tlog + b; L.mu = a.mu + b;
    return L;
}

// Per-vector reduction across the blocks (blockIdx.y) of a 32-vector group.  v[k]: this thread's partial; bit k of MAXM /
// MINM selects max / min instead of sum.  Returns true in the last block to arrive, with the group totals in v[] of the
// ty == 0 threads.  Deterministic: partials are combined in chunk order, then in ty order.
// pack_shift < 5 (node-indexed kernels at B <= 16, see node_lane below): the 32 >> pack_shift sub-warps of a warp hold
// partials of the SAME vectors (different nodes); they are folded in sub-warp order, and the totals land in sub-warp 0.
template <int K, unsigned MAXM, unsigned MINM>
__device__ __forceinline__ bool group_reduce(double (&v)[K], double* part, int* counter, int B, int b, int pack_shift = 5) {
    __shared__ double red[OPT_WY][K][32];
    __shared__ int s_last;
    const int tx = threadIdx.x, ty = threadIdx.y;
    auto comb = [](int k, double x, double y) {
        if ((MAXM >> k) & 1u) return fmax(x, y);
        if ((MINM >> k) & 1u) return fmin(x, y);
        return x + y;
    };
#pragma unroll
    for (int k = 0; k < K; k++) red[ty][k][tx] = v[k];
    __syncthreads();
    const int lpw = 1 << pack_shift, npw = 32 >> pack_shift, sub = tx >> pack_shift;
    if (ty == 0 && sub == 0 && b < B) {
#pragma unroll
        for (int k = 0; k < K; k++) {
            double x = red[0][k][tx];
            for (int y = 0; y < OPT_WY; y++)
                for (int q = (y == 0 ? 1 : 0); q < npw; q++) x = comb(k, x, red[y][k][q * lpw + tx]);
            part[((size_t)blockIdx.y * K + k) * B + b] = x;
        }
    }
    __threadfence();
    __syncthreads();
    if (tx == 0 && ty == 0) s_last = atomicAdd(counter + blockIdx.x, 1) == (int)gridDim.y - 1;
    __syncthreads();
    if (!s_last) return false;
    __threadfence();
    const int nch = gridDim.y;
#pragma unroll
    for (int k = 0; k < K; k++) {
        double x = ((MAXM >> k) & 1u) ? 0.0 : (((MINM >> k) & 1u) ? 1e300 : 0.0);
        if (b < B)
            for (int c = ty; c < nch; c += OPT_WY) x = comb(k, x, __ldcg(part + ((size_t)c * K + k) * B + b));
        red[ty][k][tx] = x;
    }
    __syncthreads();
    if (ty == 0) {
#pragma unroll
        for (int k = 0; k < K; k++) {
            double x = red[0][k][tx];
            for (int y = 1; y < OPT_WY; y++) x = comb(k, x, red[y][k][tx]);
            v[k] = x;
        }
    }
    if (tx == 0 && ty == 0) counter[blockIdx.x] = 0;        // ready for the next launch
    return true;
}

// Thread mapping of the optimiser kernels.  Normally a warp is one node (or row) x 32 vectors (tx = vector).  With 16 or fewer
// vectors most of those lanes would idle - and these kernels are bound by the latency of one node's chain per warp - so the
// warp is cut into sub-warps of LPW = 1 << pack_shift lanes (LPW >= B, one vector group only) that work on 32 / LPW different
// nodes at once: `b` = the vector, `sub` = which of the warp's nodes, `npw` = nodes per warp.
struct NodeLane {
    int b, sub, npw;
};
__device__ __forceinline__ NodeLane node_lane(const TnArgs& a) {
    const int sh = a.pack_shift;                 // 5: no packing
    NodeLane l;
    l.b = blockIdx.x * 32 + (threadIdx.x & ((1 << sh) - 1));
    l.sub = threadIdx.x >> sh;
    l.npw = 32 >> sh;
    return l;
}

struct DevView {
    const TnArgs& a;
    int B, b;
    const double* Vv;
    double* W;
    double lam, pb;
    bool log_pen;
    double mu;                                   // Levenberg-Marquardt damping of this vector
    // element offset in a [P][B] plane: 32-bit (tn_prepare refuses P * B >= 2^32), one IMAD instead of a 64-bit multiply-add chain
    __device__ unsigned ix(int row) const { return (unsigned)row * (unsigned)B + (unsigned)b; }
    __device__ double dg(int row) const { return a.DG[ix(row)]; }
    __device__ bool fr(int row) const { return a.Minv[ix(row)] > 0.0; }       // free variable of this solve
    __device__ double rhs(int row) const { return a.R[ix(row)]; }
    // the sweeps read what OTHER blocks wrote earlier in the same (cooperative) launch: past L1
    __device__ double outv(int row) const { return __ldcg(W + ix(row)); }
    __device__ double tw(int q, int i) const { return __ldcg(a.TW + (size_t)q * ((size_t)a.t.n * B) + ix(i)); }
    __device__ void tw_put(int q, int i, double val) { a.TW[(size_t)q * ((size_t)a.t.n * B) + ix(i)] = val; }
    __device__ int parent(int i) const { return a.t.parent[i]; }
    __device__ int child_begin(int i) const { return a.t.child_off[i]; }
    __device__ int child_end(int i) const { return a.t.child_off[i + 1]; }
    __device__ int child(int j) const { return a.t.children[j]; }
    __device__ int rslot(int i) const { return a.t.rslot[i]; }
    __device__ int dslot(int i) const { return a.t.dslot[i]; }
    __device__ double c(int i) const { return a.t.c[i]; }
    __device__ double hfix(int i) const { return a.t.hfix[i]; }
    __device__ double rfix(int i) const { return a.t.rfix[i]; }
    __device__ double pmin(int i) const { return a.pmin_n[i]; }
    __device__ double pmax(int i) const { return a.pmax_n[i]; }
    __device__ bool masked(int i) const {
        if (!a.lanemask) return false;
        return (a.lanemask[(size_t)(b >> 6) * a.mask_stride + i] >> (b & 63)) & 1ull;
    }
    __device__ double sc(int row) const { return a.sc[row]; }
    __device__ double x(int row) const { return a.X[ix(row)]; }
    __device__ double v(int row) const { return Vv[ix(row)]; }
    __device__ double gx(int row) const { return a.Gx[ix(row)]; }
    __device__ void put(int row, double val) { W[ix(row)] = val; }
};

__device__ __forceinline__ int lane_decision(const TnArgs& a, int b) {
    return decide(a.st[b], a.Ft[b], a.F0[b], a.t_[b], a.gd[b], a.o.c1, f_allowance(a.F0[b], a.o.ftol, a.o.fnoise));
}

// the damping the factorisation at a newly accepted point uses: the value phase_setup is about to store
__device__ __forceinline__ double lane_mu_eff(const TnArgs& a, int b, int dec) {
    return dec == DEC_ACCEPT ? mu_after_accept(a.mu[b], a.t_[b], a.ls[b]) : a.mu[b];
}

// Z = clamp(w / sc), S = 0
__global__ void __launch_bounds__(32 * OPT_WY) tn_init_kernel(TnArgs a) {
    const NodeLane nl = node_lane(a);           // B <= 16: a warp carries several rows (as the node-indexed kernels carry nodes)
    const int b = nl.b;
    if (b >= a.o.B) return;
    if (blockIdx.y == 0 && threadIdx.y == 0 && nl.sub == 0) a.act[b] = ACT_TRIAL;      // the first trial point is the start itself (t = 0)
    const int p0 = blockIdx.y * a.rows_per_chunk, p1 = min(a.o.P, p0 + a.rows_per_chunk);
    for (int p = p0 + threadIdx.y * nl.npw + nl.sub; p < p1; p += OPT_WY * nl.npw) {
        const size_t i = (size_t)p * a.o.B + b;
        const double x = a.X[i];
        const double w = p < a.o.nr ? log(x) : x;
        a.Z[i] = fmin(fmax(w / a.sc[p], a.lo[p]), a.hi[p]);
        a.S[i] = 0.0;
    }
}

__global__ void __launch_bounds__(32 * OPT_WY) tn_trial_kernel(TnArgs a) {
    const NodeLane nl = node_lane(a);
    const int b = nl.b;
    const int B = a.o.B, P = a.o.P, nr = a.o.nr;
    if (b >= B) return;
    const int act = a.act[b];
    if (a.st[b] == TS_DONE || !(act & ACT_TRIAL)) return;
    const bool sd = (act & ACT_SD) != 0;
    const double t = a.t_[b];
    const int p0 = blockIdx.y * a.rows_per_chunk, p1 = min(P, p0 + a.rows_per_chunk);
    for (int p = p0 + threadIdx.y * nl.npw + nl.sub; p < p1; p += OPT_WY * nl.npw) {
        const size_t i = (size_t)p * B + b;
        double s;
        if (sd) {
            s = -a.Gz[i] * a.Minv[i];
            a.S[i] = s;
        } else {
            s = a.S[i];
        }
        a.X[i] = tplopt::to_param(p < nr, tplopt::trial_point(a.Z[i], t, s, a.lo[p], a.hi[p]), a.sc[p]);
    }
}

// ---- record views: the tree sweeps are LATENCY bound (one warp = one node x 32 vectors, a chain of dependent loads
// node -> slots -> parent -> parent's slots -> rows -> children -> children's factors: ~12 round trips to L2 / HBM, ~10 us per
// node at batch sizes whose work planes exceed L2).  The host therefore flattens everything index-like into one 48-byte
// record per node, stored in level order (tn_prepare), and a view loads the record (round trip 1), then ALL operands the
// node function will ask for, unconditionally and side by side (round trip 2); the node functions of tpl_tn_core.h run
// unchanged on top and find their operands in registers.  Nodes with more than two children fall through to the generic
// accessors for the extra children.
constexpr int TREE_REC_INTS = 12;            // i rs ds p | prs pds coff nch | ch0 ch1 cds0 cds1

struct TreeRec {
    int i, rs, ds, p, prs, pds, coff, nch, ch0, ch1, cds0, cds1;
};
__device__ __forceinline__ TreeRec tree_rec_load(const int* recs, int k) {
    const int4* r = reinterpret_cast<const int4*>(recs + (size_t)k * TREE_REC_INTS);
    const int4 q0 = __ldg(r), q1 = __ldg(r + 1), q2 = __ldg(r + 2);
    TreeRec t = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w};
    return t;
}

struct RecBase : DevView {
    TreeRec t;
    __device__ RecBase(const DevView& v, const TreeRec& rec) : DevView(v), t(rec) {}
    __device__ int parent(int n) const { return n == t.i ? t.p : DevView::parent(n); }
    __device__ int child_begin(int n) const { return n == t.i ? t.coff : DevView::child_begin(n); }
    __device__ int child_end(int n) const { return n == t.i ? t.coff + t.nch : DevView::child_end(n); }
    __device__ int child(int j) const {
        const int q = j - t.coff;
        return (q == 0 && t.nch > 0) ? t.ch0 : ((q == 1 && t.nch > 1) ? t.ch1 : DevView::child(j));
    }
    __device__ int rslot(int n) const { return n == t.i ? t.rs : (n == t.p ? t.prs : DevView::rslot(n)); }
    __device__ int dslot(int n) const {
        return n == t.i ? t.ds : (n == t.p ? t.pds : (n == t.ch0 ? t.cds0 : (n == t.ch1 ? t.cds1 : DevView::dslot(n))));
    }
    __device__ unsigned long long mask_word(int n) const { return a.lanemask[(size_t)(b >> 6) * a.mask_stride + n]; }
};


// diag_node for the ordinary node (same records and the same idea as hess_node_fast below: every row requested before the
// first is used); same arithmetic, in the same order, as diag_node
__device__ __forceinline__ void diag_node_fast(const TnArgs& a, const int4 q0, const int4 q1, const int4 q2, int B, int b, double lam) {
    const int i = q0.x, rs = q0.y, ds = q0.z, pds = q1.y, nch = q1.z, p = q0.w;
    const int ch[2] = {q2.x, q2.y}, cds[2] = {q2.z, q2.w};
    const size_t Bs = (size_t)B;
    bool mi = false, mc[2] = {false, false};
    if (a.lanemask) {
        const unsigned long long* mw = a.lanemask + (size_t)(b >> 6) * a.mask_stride;
        const unsigned long long wi = mw[i], w0 = nch ? mw[ch[0]] : 0ull, w1 = nch ? mw[ch[1]] : 0ull;
        mi = (wi >> (b & 63)) & 1ull; mc[0] = (w0 >> (b & 63)) & 1ull; mc[1] = (w1 >> (b & 63)) & 1ull;
    }
    const double r = a.X[(unsigned)rs * (unsigned)B + (unsigned)b];
    const double h = ds >= 0 ? a.X[(unsigned)ds * (unsigned)B + (unsigned)b] : a.t.hfix[i];
    const double hp = pds >= 0 ? a.X[(unsigned)pds * (unsigned)B + (unsigned)b] : a.t.hfix[p];
    double hc[2] = {0.0, 0.0}, cc[2] = {0.0, 0.0};
    if (nch && ds >= 0) {
#pragma unroll
        for (int k = 0; k < 2; k++) {
            hc[k] = cds[k] >= 0 ? a.X[(unsigned)cds[k] * (unsigned)B + (unsigned)b] : a.t.hfix[ch[k]];
            cc[k] = a.t.c[ch[k]];
        }
    }
    const double c = a.t.c[i];
    const double gx = a.Gx[(unsigned)rs * (unsigned)B + (unsigned)b];
    const double sc_rs = a.sc[rs], sc_ds = ds >= 0 ? a.sc[ds] : 0.0;
    const double pmn = ds >= 0 ? a.pmin_n[i] : -1.0, pmx = ds >= 0 ? a.pmax_n[i] : -1.0;
    const bool rfree = !mi;
    const double lam2 = 2.0 * lam;
    double H_r = 0.0, H_h = 0.0;
    if (!mi) {
        const double d = safe_duration(hp - h);
        H_r = c * tn_rcp(r * r);
        H_h = c * tn_rcp(d * d);
        H_r += lam2;
    }
    if (nch) {
#pragma unroll
        for (int k = 0; k < 2; k++) {
            if (mc[k]) continue;
            if (ds >= 0) {
                const double dc = safe_duration(h - hc[k]);
                H_h += cc[k] * tn_rcp(dc * dc);
            }
            if (rfree) H_r += lam2;
        }
    }
    if (ds >= 0) {
        if (pmn != -1.0) { const double q = h - pmn; if (q > 0.0) H_h += a.pb * 2.0 / (q * q * q); }
        if (pmx != -1.0) { const double q = pmx - h; if (q > 0.0) H_h += a.pb * 2.0 / (q * q * q); }
    }
    a.DG[(unsigned)rs * (unsigned)B + (unsigned)b] = rfree ? sc_rs * sc_rs * r * r * H_r + gx * sc_rs * sc_rs * r : 0.0;
    if (ds >= 0) a.DG[(unsigned)ds * (unsigned)B + (unsigned)b] = sc_ds * sc_ds * H_h;
}

__global__ void __launch_bounds__(32 * OPT_WY) tn_diag_kernel(TnArgs a) {
    const NodeLane nl = node_lane(a);
    const int b = nl.b;
    const int B = a.o.B;
    if (b >= B) return;
    const int dec = lane_decision(a, b);
    if (dec != DEC_ACCEPT && dec != DEC_INIT) return;
    DevView V = {a, B, b, a.Pv, a.DG, a.lane_lambda ? a.lane_lambda[b] : a.lambda, a.pb, a.log_pen != 0, 0.0};
    const int i0 = blockIdx.y * a.diag_nodes_per_chunk, i1 = min(a.t.n, i0 + a.diag_nodes_per_chunk);
    const int4* recs = reinterpret_cast<const int4*>(a.hv_rec);
    for (int i = i0 + threadIdx.y * nl.npw + nl.sub; i < i1; i += OPT_WY * nl.npw) {
        const int4 q1 = __ldg(recs + 4 * (size_t)i + 1);
        if (q1.w) diag_node_fast(a, __ldg(recs + 4 * (size_t)i), q1, __ldg(recs + 4 * (size_t)i + 2), B, b, V.lam);
        else diag_node(V, i);
    }
}

__global__ void __launch_bounds__(32 * OPT_WY) tn_setup_kernel(TnArgs a) {
    const int ty = threadIdx.y;
    const NodeLane nl = node_lane(a);
    const int b = nl.b;
    const int B = a.o.B, P = a.o.P, nr = a.o.nr;
    int dec = DEC_SKIP;
    double t = 0.0;
    if (b < B) {
        dec = lane_decision(a, b);
        t = a.t_[b];
    }
    if (!__syncthreads_or(dec != DEC_SKIP)) return;          // uniform over blockIdx.y
    double v[3] = {0.0, 0.0, 0.0};
    if (dec == DEC_ACCEPT || dec == DEC_INIT) {
        const double mu = lane_mu_eff(a, b, dec);
        const int p0 = blockIdx.y * a.rows_per_chunk, p1 = min(P, p0 + a.rows_per_chunk);
        for (int p = p0 + ty * nl.npw + nl.sub; p < p1; p += OPT_WY * nl.npw) {
            const size_t i = (size_t)p * B + b;
            const double lo = a.lo[p], hi = a.hi[p];
            double z = a.Z[i];
            if (dec == DEC_ACCEPT) z = tplopt::trial_point(z, t, a.S[i], lo, hi);
            const double gz = tplopt::chain_grad(p < nr, a.Gx[i], a.X[i], a.sc[p]);
            const double mi = precond_inverse(a.DG[i], tplopt::is_free(z, gz, lo, hi), mu);
            const double gp = mi > 0.0 ? gz : 0.0;
            const double zz = -gp * mi;
            a.Z[i] = z;
            a.Gz[i] = gz;
            a.Minv[i] = mi;
            a.R[i] = -gp;
            a.ZZ[i] = zz;
            a.Pv[i] = zz;
            a.S[i] = 0.0;
            v[0] = fma(gp * mi, gp, v[0]);
            v[1] = fmax(v[1], fabs(gp));
            v[2] = fma(gp, gp, v[2]);
        }
    }
    if (!group_reduce<3, 2u, 0u>(v, a.part_setup, a.cnt_setup, B, b, a.pack_shift)) return;
    if (ty == 0 && nl.sub == 0 && b < B) {
        if (a.precond && (dec == DEC_ACCEPT || dec == DEC_INIT)) {
            // r.M^-1 r comes from the tree solve (tn_pc_finish_kernel runs phase_setup); the decision inputs stay untouched
            a.gmax_tmp[b] = v[1];
            a.gg_tmp[b] = v[2];
        } else if (phase_setup(lane_of(a, b), a.Ft[b], v, a.o)) {
            atomicAdd(a.ndone, 1);
        }
    }
}

// hess_node for the ordinary node - not the root, not a child of the root, no or two children (99.9 % of a tree) - written
// against a 64-byte host-built index record (node order) so that ALL rows the node needs (its own, its parent's, its
// children's X and P rows) are requested side by side: one round trip to HBM per node instead of the three to four that
// the accessor chain of the generic function serialises (own rows -> parent -> children), which is what bound this
// kernel (long-scoreboard stalls, 0.26-0.47 of the HBM rate).  Same arithmetic, in the same order, as hess_node.
constexpr int HV_REC_INTS = 16;              // i rs ds p | prs pds nch fast | ch0 ch1 cds0 cds1 | crs0 crs1 - -

// everything hess_node reads for an ordinary node and one vector
struct HvOperands {
    double r, v_rs, h, v_ds, rp, v_prs, hp, v_pds;
    double rc[2], hc[2], v_crs[2], v_cds[2], cc[2], sc_crs[2], sc_cds[2];
    double sc_rs, sc_ds, sc_prs, sc_pds, c, gx, dg_rs, dg_ds, pmn, pmx;
    bool mi, mp, mc[2];
};

// hess_node on loaded operands: same arithmetic, in the same order.  rs / ds: the node's own rows (ds < 0: no date row);
// pds, cds: only their sign is used.  -> Hp rows written, vw = v.Hv and vv = v.v contributions returned.
// Element offsets row * B + b are 32-bit here (tn_prepare refuses P * B >= 2^32: eleven such planes do not fit any GPU).
__device__ __forceinline__ void hess_node_fast_math(const TnArgs& a, const HvOperands& o, int rs, int ds, int pds, int nch, int cds0, int cds1,
                                                    unsigned Bs, unsigned b, double lam, double mu, double& vw, double& vv) {
    const int cds[2] = {cds0, cds1};
    const bool damp = mu > 0.0;
    const bool rfree = !o.mi;
    const double lam2 = 2.0 * lam;
    const double vr = rfree ? o.sc_rs * o.r * o.v_rs : 0.0;
    const double vh = ds >= 0 ? o.sc_ds * o.v_ds : 0.0;
    double w_r = 0.0, w_h = 0.0;
    if (!o.mi) {
        const double vrp = !o.mp ? o.sc_prs * o.rp * o.v_prs : 0.0;
        const double vhp = pds >= 0 ? o.sc_pds * o.v_pds : 0.0;
        const double d = safe_duration(o.hp - o.h);
        const double dv = vhp - vh;
        w_r = (o.c * tn_rcp(o.r * o.r)) * vr + dv;
        w_h = -((o.c * tn_rcp(d * d)) * dv + vr);
        w_r += lam2 * (vr - vrp);
    }
    if (nch) {
#pragma unroll
        for (int k = 0; k < 2; k++) {
            if (o.mc[k]) continue;
            const double vrc = o.sc_crs[k] * o.rc[k] * o.v_crs[k];
            const double vhc = cds[k] >= 0 ? o.sc_cds[k] * o.v_cds[k] : 0.0;
            if (ds >= 0) {
                const double dc = safe_duration(o.h - o.hc[k]);
                w_h += (o.cc[k] * tn_rcp(dc * dc)) * (vh - vhc) + vrc;
            }
            if (rfree) w_r -= lam2 * (vrc - vr);
        }
    }
    if (ds >= 0) {
        double bar = 0.0;
        if (o.pmn != -1.0) { const double q = o.h - o.pmn; if (q > 0.0) bar += 2.0 / (q * q * q); }
        if (o.pmx != -1.0) { const double q = o.pmx - o.h; if (q > 0.0) bar += 2.0 / (q * q * q); }
        w_h += a.pb * bar * vh;
    }
    vw = 0.0;
    vv = 0.0;
    {
        double W = 0.0;
        if (rfree) {
            W = o.sc_rs * o.r * w_r + o.gx * o.sc_rs * o.sc_rs * o.r * o.v_rs;
            if (damp) W += mu * fabs(o.dg_rs) * o.v_rs;
            vw += o.v_rs * W;
            vv += o.v_rs * o.v_rs;
        }
        a.Hp[(unsigned)rs * Bs + b] = W;
    }
    if (ds >= 0) {
        double W = o.sc_ds * w_h;
        if (damp) W += mu * fabs(o.dg_ds) * o.v_ds;
        vw += o.v_ds * W;
        vv += o.v_ds * o.v_ds;
        a.Hp[(unsigned)ds * Bs + b] = W;
    }
}

__device__ __forceinline__ void hess_node_fast(const TnArgs& a, const int4 q0, const int4 q1, const int4 q2, const int4 q3, int B, int b,
                                               double lam, double mu, double& vw, double& vv) {
    const int i = q0.x, rs = q0.y, ds = q0.z, p = q0.w, prs = q1.x, pds = q1.y, nch = q1.z;
    const int ch[2] = {q2.x, q2.y}, cds[2] = {q2.z, q2.w}, crs[2] = {q3.x, q3.y};
    const unsigned Bs = (unsigned)B, bu = (unsigned)b;
    const double* __restrict__ X = a.X;
    const double* __restrict__ Pv = a.Pv;
    HvOperands o;
    // ---- every load first ----
    o.mi = false; o.mp = false; o.mc[0] = o.mc[1] = false;
    if (a.lanemask) {
        const unsigned long long* mw = a.lanemask + (size_t)(b >> 6) * a.mask_stride;
        const unsigned long long wi = mw[i], wp = mw[p], w0 = nch ? mw[ch[0]] : 0ull, w1 = nch ? mw[ch[1]] : 0ull;
        o.mi = (wi >> (b & 63)) & 1ull; o.mp = (wp >> (b & 63)) & 1ull; o.mc[0] = (w0 >> (b & 63)) & 1ull; o.mc[1] = (w1 >> (b & 63)) & 1ull;
    }
    o.r = X[(unsigned)rs * Bs + bu]; o.v_rs = Pv[(unsigned)rs * Bs + bu];
    o.h = ds >= 0 ? X[(unsigned)ds * Bs + bu] : a.t.hfix[i];
    o.v_ds = ds >= 0 ? Pv[(unsigned)ds * Bs + bu] : 0.0;
    o.rp = X[(unsigned)prs * Bs + bu]; o.v_prs = Pv[(unsigned)prs * Bs + bu];
    o.hp = pds >= 0 ? X[(unsigned)pds * Bs + bu] : a.t.hfix[p];
    o.v_pds = pds >= 0 ? Pv[(unsigned)pds * Bs + bu] : 0.0;
#pragma unroll
    for (int k = 0; k < 2; k++) { o.rc[k] = o.hc[k] = o.v_crs[k] = o.v_cds[k] = o.cc[k] = o.sc_crs[k] = o.sc_cds[k] = 0.0; }
    if (nch) {
#pragma unroll
        for (int k = 0; k < 2; k++) {
            o.rc[k] = X[(unsigned)crs[k] * Bs + bu];
            o.v_crs[k] = Pv[(unsigned)crs[k] * Bs + bu];
            o.hc[k] = cds[k] >= 0 ? X[(unsigned)cds[k] * Bs + bu] : a.t.hfix[ch[k]];
            o.v_cds[k] = cds[k] >= 0 ? Pv[(unsigned)cds[k] * Bs + bu] : 0.0;
            o.cc[k] = a.t.c[ch[k]];
            o.sc_crs[k] = a.sc[crs[k]];
            o.sc_cds[k] = cds[k] >= 0 ? a.sc[cds[k]] : 0.0;
        }
    }
    o.sc_rs = a.sc[rs]; o.sc_ds = ds >= 0 ? a.sc[ds] : 0.0; o.sc_prs = a.sc[prs]; o.sc_pds = pds >= 0 ? a.sc[pds] : 0.0;
    o.c = a.t.c[i];
    o.gx = a.Gx[(unsigned)rs * Bs + bu];
    const bool damp = mu > 0.0;
    o.dg_rs = damp ? a.DG[(unsigned)rs * Bs + bu] : 0.0; o.dg_ds = (damp && ds >= 0) ? a.DG[(unsigned)ds * Bs + bu] : 0.0;
    o.pmn = ds >= 0 ? a.pmin_n[i] : -1.0; o.pmx = ds >= 0 ? a.pmax_n[i] : -1.0;
    hess_node_fast_math(a, o, rs, ds, pds, nch, cds[0], cds[1], Bs, bu, lam, mu, vw, vv);
}

__global__ void __launch_bounds__(32 * OPT_WY) tn_hv_kernel(TnArgs a) {
    const int ty = threadIdx.y;
    const NodeLane nl = node_lane(a);
    const int b = nl.b;
    const int B = a.o.B;
    // a vector whose CG solve ended in an earlier sub-iteration of this step waits for the line-search set-up
    const bool active = b < B && a.st[b] == TS_CG && !(a.act[b] & ACT_FIN);
    if (!__syncthreads_or(active)) return;
    double v[2] = {0.0, 0.0};
    if (active) {
        DevView V = {a, B, b, a.Pv, a.Hp, a.lane_lambda ? a.lane_lambda[b] : a.lambda, a.pb, a.log_pen != 0, a.mu[b]};
        const int i0 = blockIdx.y * a.nodes_per_chunk, i1 = min(a.t.n, i0 + a.nodes_per_chunk);
        // the record of the NEXT node is requested before this node's rows: its round trip overlaps this node's work instead of
        // heading the next node's chain (per node: record -> rows -> arithmetic; the kernel's rate is warps in flight / that latency)
        const int4* recs = reinterpret_cast<const int4*>(a.hv_rec);
        const int istep = OPT_WY * nl.npw;
        int i = i0 + ty * nl.npw + nl.sub;
        int4 n0, n1, n2, n3;
        if (i < i1) { n0 = __ldg(recs + 4 * (unsigned)i); n1 = __ldg(recs + 4 * (unsigned)i + 1); n2 = __ldg(recs + 4 * (unsigned)i + 2); n3 = __ldg(recs + 4 * (unsigned)i + 3); }
        for (; i < i1; i += istep) {
            const int4 q0 = n0, q1 = n1, q2 = n2, q3 = n3;
            const int inext = i + istep;
            if (inext < i1) { n0 = __ldg(recs + 4 * (unsigned)inext); n1 = __ldg(recs + 4 * (unsigned)inext + 1); n2 = __ldg(recs + 4 * (unsigned)inext + 2); n3 = __ldg(recs + 4 * (unsigned)inext + 3); }
            double vw, vv;
            if (q1.w) hess_node_fast(a, q0, q1, q2, q3, B, b, V.lam, V.mu, vw, vv);
            else hess_node(V, i, vw, vv);
            v[0] += vw;
            v[1] += vv;
        }
    }
    if (!group_reduce<2, 0u, 0u>(v, a.part_hv, a.cnt_hv, B, b, a.pack_shift)) return;
    if (ty == 0 && nl.sub == 0 && active) phase_hv(lane_of(a, b), v, a.o);
}

// ---- tn_hv_pipe_kernel: the same product with the operand rows staged ASYNCHRONOUSLY in shared memory ----
// tn_hv_kernel keeps one node per warp in flight: record -> ~19 row loads held in registers -> ~100 dependent fp64 instructions,
// 16 warps per SM at 118 registers; its rate is warps in flight / latency of that chain (0.31-0.33 of the HBM rate at 64 vectors
// on the 100k-tip tree).  Here every warp runs its own software pipeline with cp.async: while it computes node k from shared
// memory, the operand rows of nodes k+1 and k+2 and the records of node k+3 are on the wire - three times the bytes in flight
// per SM without a register to hold them.  A lane copies, and later reads, only its own vector's 8 bytes of each row, so the row
// ring needs no barrier at all; the per-node records (64-byte index record, 128-byte constant record built by the host, four
// CV-mask words) are copied by the first lanes and read by all after cp.async.wait_group + __syncwarp.
//   ring per warp: 3 row slots x 16 rows x 32 lanes x 8 B + 4 record slots x 224 B = 13,184 B; 8 warps = 105,472 B per block,
//   two blocks per SM.  Groups are committed in the order R(k+3), W(k+2) per iteration, so "wait_group 2" at the top of
//   iteration k has R(k+2) and W(k) complete and leaves W(k+1), R(k+3) (and then W(k+2)) pending.
// Nodes outside the fast path (root, its children, polytomies) take hess_node on global memory as before.
constexpr int HV_CONST = 16;                 // sc_rs, sc_ds|hfix_i, sc_prs, sc_pds|hfix_p, sc_crs0, sc_cds0|hfix_c0, sc_crs1, sc_cds1|hfix_c1, c, cc0, cc1, pmn, pmx
constexpr int HVP_ROWS = 16, HVP_WSLOTS = 3, HVP_RSLOTS = 4;
constexpr int HVP_REC_BYTES = HV_REC_INTS * 4 + HV_CONST * 8 + 32;       // 224
constexpr int HVP_WARP_BYTES = HVP_WSLOTS * HVP_ROWS * 32 * 8 + HVP_RSLOTS * HVP_REC_BYTES;
constexpr int HVP_SMEM = OPT_WY * HVP_WARP_BYTES;

__device__ __forceinline__ void cp_async8(void* dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// One fast-path node from the staged rows.  The host marks a node fast only in two shapes - a tip (no children, no date row) or an
// internal node with two children and a date row - and tn_hv_pipe_kernel's issue step fills the slot of every row that does not
// exist (the fixed date of a tip child or of a calibrated parent, with a zero direction entry), so nothing here depends on the
// record at run time: the class is a template parameter and hess_node_fast_math folds to straight-line code.
template <bool INTERNAL, bool MASK>
__device__ __forceinline__ void hvp_node(const TnArgs& a, const unsigned char* rec, const double* w, unsigned Bs, unsigned bu, int b,
                                         double lam, double mu, double& vw, double& vv) {
    const int rs = reinterpret_cast<const int*>(rec)[1], ds = reinterpret_cast<const int*>(rec)[2];
    const double2* c2 = reinterpret_cast<const double2*>(rec + HV_REC_INTS * 4);
    HvOperands o;
    o.mi = o.mp = o.mc[0] = o.mc[1] = false;
    if (MASK) {
        const unsigned long long* mk = reinterpret_cast<const unsigned long long*>(rec + HV_REC_INTS * 4 + HV_CONST * 8);
        o.mi = (mk[0] >> (b & 63)) & 1ull; o.mp = (mk[1] >> (b & 63)) & 1ull;
        if (INTERNAL) { o.mc[0] = (mk[2] >> (b & 63)) & 1ull; o.mc[1] = (mk[3] >> (b & 63)) & 1ull; }
    }
    const double2 s0 = c2[0], s1 = c2[1], k4 = c2[4];            // (sc_rs, sc_ds | hfix) (sc_prs, sc_pds | hfix) (c, cc0)
    o.r = w[0 * 32]; o.v_rs = w[1 * 32]; o.h = w[2 * 32]; o.v_ds = INTERNAL ? w[3 * 32] : 0.0;
    o.rp = w[4 * 32]; o.v_prs = w[5 * 32]; o.hp = w[6 * 32]; o.v_pds = w[7 * 32];
    o.sc_rs = s0.x; o.sc_ds = INTERNAL ? s0.y : 0.0; o.sc_prs = s1.x; o.sc_pds = s1.y; o.c = k4.x;
#pragma unroll
    for (int q = 0; q < 2; q++) { o.rc[q] = o.hc[q] = o.v_crs[q] = o.v_cds[q] = o.cc[q] = o.sc_crs[q] = o.sc_cds[q] = 0.0; }
    o.pmn = o.pmx = -1.0;
    if (INTERNAL) {
        const double2 s2 = c2[2], s3 = c2[3], k5 = c2[5], k6 = c2[6];   // (sc_crs0, sc_cds0 | hfix) (sc_crs1, sc_cds1 | hfix) (cc1, pmn) (pmx, -)
        o.rc[0] = w[8 * 32]; o.v_crs[0] = w[9 * 32]; o.hc[0] = w[10 * 32]; o.v_cds[0] = w[11 * 32];
        o.rc[1] = w[12 * 32]; o.v_crs[1] = w[13 * 32]; o.hc[1] = w[14 * 32]; o.v_cds[1] = w[15 * 32];
        o.sc_crs[0] = s2.x; o.sc_cds[0] = s2.y; o.sc_crs[1] = s3.x; o.sc_cds[1] = s3.y;
        o.cc[0] = k4.y; o.cc[1] = k5.x; o.pmn = k5.y; o.pmx = k6.x;
    }
    o.gx = a.Gx[(unsigned)rs * Bs + bu];
    const bool damp = mu > 0.0;
    o.dg_rs = damp ? a.DG[(unsigned)rs * Bs + bu] : 0.0;
    o.dg_ds = (INTERNAL && damp) ? a.DG[(unsigned)ds * Bs + bu] : 0.0;
    hess_node_fast_math(a, o, rs, INTERNAL ? ds : -1, 0, INTERNAL ? 2 : 0, 0, 0, Bs, bu, lam, mu, vw, vv);
}

template <bool MASK>
__global__ void __launch_bounds__(32 * OPT_WY, 2) tn_hv_pipe_kernel(TnArgs a) {
    extern __shared__ __align__(16) unsigned char hvp_smem[];
    const int ty = threadIdx.y, tx = threadIdx.x;
    const int B = a.o.B;
    const int b = blockIdx.x * 32 + tx;
    const bool active = b < B && a.st[b] == TS_CG && !(a.act[b] & ACT_FIN);
    if (!__syncthreads_or(active)) return;
    double v[2] = {0.0, 0.0};
    unsigned char* wbase = hvp_smem + (size_t)ty * HVP_WARP_BYTES;
    double* wrows = reinterpret_cast<double*>(wbase);                                  // [HVP_WSLOTS][HVP_ROWS][32]
    unsigned char* rrecs = wbase + HVP_WSLOTS * HVP_ROWS * 32 * 8;                      // [HVP_RSLOTS][HVP_REC_BYTES]
    const unsigned Bs = (unsigned)B, bu = (unsigned)b;
    const double* __restrict__ X = a.X;
    const double* __restrict__ Pv = a.Pv;
    const int i0 = blockIdx.y * a.nodes_per_chunk, i1 = min(a.t.n, i0 + a.nodes_per_chunk);
    const int m = i0 + ty < i1 ? (i1 - i0 - ty + OPT_WY - 1) / OPT_WY : 0;               // this warp's nodes: i0 + ty + OPT_WY k
    const unsigned long long* mw = MASK ? a.lanemask + (size_t)(b >> 6) * a.mask_stride : nullptr;   // (b >> 6 is warp uniform)
    const double lam = active ? (a.lane_lambda ? a.lane_lambda[b] : a.lambda) : 0.0, mu = active ? a.mu[b] : 0.0;
    DevView V = {a, B, b, a.Pv, a.Hp, lam, a.pb, a.log_pen != 0, mu};

    auto issue_rec = [&](int k) {             // R(k): index record (lanes 0-3) and constant record (lanes 4-11), 16 bytes each
        if (k < m) {
            const int i = i0 + ty + OPT_WY * k;
            unsigned char* dst = rrecs + (k % HVP_RSLOTS) * HVP_REC_BYTES;
            if (tx < 4) cp_async16(dst + 16 * tx, reinterpret_cast<const unsigned char*>(a.hv_rec + (size_t)i * HV_REC_INTS) + 16 * tx);
            else if (tx < 12) cp_async16(dst + 16 * tx, reinterpret_cast<const unsigned char*>(a.hv_const + (size_t)i * HV_CONST) + 16 * (tx - 4));
        }
        cp_async_commit();
    };
    // W(k): the operand rows of node k (its records are in shared memory): X and P entries of the rows rs ds prs pds crs0 cds0 crs1 cds1
    // land in ring rows 2q (X) and 2q + 1 (P); a row that does not exist gets its fixed date and a zero direction entry instead
    auto issue_rows = [&](int k) {
        if (k < m) {
            unsigned char* rec = rrecs + (k % HVP_RSLOTS) * HVP_REC_BYTES;
            const int4 q0 = *reinterpret_cast<const int4*>(rec), q1 = *reinterpret_cast<const int4*>(rec + 16);
            if (q1.w) {
                const int4 q2 = *reinterpret_cast<const int4*>(rec + 32);
                if (active) {
                    const double* cst = reinterpret_cast<const double*>(rec + HV_REC_INTS * 4);
                    double* dst = wrows + (k % HVP_WSLOTS) * (HVP_ROWS * 32) + tx;
                    auto row = [&](int q, int r, double fixed) {
                        if (r >= 0) {
                            const unsigned off = (unsigned)r * Bs + bu;
                            cp_async8(dst + (2 * q) * 32, X + off);
                            cp_async8(dst + (2 * q + 1) * 32, Pv + off);
                        } else {
                            dst[(2 * q) * 32] = fixed;
                            dst[(2 * q + 1) * 32] = 0.0;
                        }
                    };
                    row(0, q0.y, 0.0);
                    row(2, q1.x, 0.0);
                    row(3, q1.y, cst[3]);
                    if (q1.z) {
                        const int2 q3 = *reinterpret_cast<const int2*>(rec + 48);
                        row(1, q0.z, 0.0);
                        row(4, q3.x, 0.0);
                        row(5, q2.z, cst[5]);
                        row(6, q3.y, 0.0);
                        row(7, q2.w, cst[7]);
                    } else {
                        dst[2 * 32] = cst[1];                 // the tip's date
                    }
                }
                if (MASK && tx < 4) {
                    const int nd = tx == 0 ? q0.x : (tx == 1 ? q0.w : (tx == 2 ? q2.x : q2.y));   // i p ch0 ch1
                    if (tx < 2 || q1.z) cp_async8(rec + HV_REC_INTS * 4 + HV_CONST * 8 + 8 * tx, mw + nd);
                }
            }
        }
        cp_async_commit();
    };

    issue_rec(0); issue_rec(1); issue_rec(2);
    cp_async_wait<2>(); __syncwarp();
    issue_rows(0);
    cp_async_wait<2>(); __syncwarp();
    issue_rows(1);
    for (int k = 0; k < m; k++) {
        __syncwarp();                                         // every lane is done with node k-1 (its record slot is reused below)
        issue_rec(k + 3);
        cp_async_wait<2>();
        __syncwarp();                                         // R(k+2), W(k) visible to all lanes
        issue_rows(k + 2);
        if (!active) continue;
        const unsigned char* rec = rrecs + (k % HVP_RSLOTS) * HVP_REC_BYTES;
        const int fast = reinterpret_cast<const int*>(rec)[7], nch = reinterpret_cast<const int*>(rec)[6];
        const double* w = wrows + (k % HVP_WSLOTS) * (HVP_ROWS * 32) + tx;
        double vw, vv;
        if (fast) {
            if (nch) hvp_node<true, MASK>(a, rec, w, Bs, bu, b, lam, mu, vw, vv);
            else hvp_node<false, MASK>(a, rec, w, Bs, bu, b, lam, mu, vw, vv);
        } else {
            hess_node(V, reinterpret_cast<const int*>(rec)[0], vw, vv);
        }
        v[0] += vw;
        v[1] += vv;
    }
    cp_async_wait<0>();
    if (!group_reduce<2, 0u, 0u>(v, a.part_hv, a.cnt_hv, B, b, 5)) return;
    if (ty == 0 && active) phase_hv(lane_of(a, b), v, a.o);
}

__global__ void __launch_bounds__(32 * OPT_WY) tn_update_kernel(TnArgs a) {
    const int ty = threadIdx.y;
    const NodeLane nl = node_lane(a);
    const int b = nl.b;
    const int B = a.o.B, P = a.o.P;
    const bool active = b < B && a.st[b] == TS_CG && (a.act[b] & ACT_CG);
    if (!__syncthreads_or(active)) return;
    double v[1] = {0.0};
    if (active) {
        const double alpha = a.alpha[b];
        const int p0 = blockIdx.y * a.rows_per_chunk, p1 = min(P, p0 + a.rows_per_chunk);
        for (int p = p0 + ty * nl.npw + nl.sub; p < p1; p += OPT_WY * nl.npw) {
            const size_t i = (size_t)p * B + b;
            const double pv = a.Pv[i];
            a.S[i] = fma(alpha, pv, a.S[i]);
            const double r = fma(-alpha, a.Hp[i], a.R[i]);
            a.R[i] = r;
            const double zz = r * a.Minv[i];
            a.ZZ[i] = zz;
            v[0] = fma(r, zz, v[0]);
        }
    }
    if (a.precond) return;                       // zz = M^-1 r and r.zz follow from the tree sweeps
    if (!group_reduce<1, 0u, 0u>(v, a.part_upd, a.cnt_upd, B, b, a.pack_shift)) return;
    if (ty == 0 && nl.sub == 0 && active) phase_update(lane_of(a, b), v, a.o);
}

// ---- tree preconditioner sweeps ------------------------------------------------------------------------------------------
constexpr int TREE_TOP_WY = 16;              // warps of the block that walks the narrow top levels
constexpr int TREE_NODES_PER_BLOCK = 4 * OPT_WY;

// FACTOR: the sweep that follows an accepted step (factorise at the new point, solve for the gradient);
// otherwise the solve inside a CG iteration (same vectors as tn_update_kernel)
template <bool FACTOR>
__device__ __forceinline__ bool tree_lane_active(const TnArgs& a, int b) {
    if (b >= a.o.B) return false;
    if (FACTOR) {
        const int dec = lane_decision(a, b);
        return dec == DEC_ACCEPT || dec == DEC_INIT;
    }
    return a.st[b] == TS_CG && (a.act[b] & ACT_CG);
}

template <bool FACTOR>
__device__ __forceinline__ double tree_lane_mu(const TnArgs& a, int b) {
    return FACTOR ? lane_mu_eff(a, b, lane_decision(a, b)) : a.mu[b];
}

// the point-only part of the factorisation (tree_local_node): every node independently
struct LocalView : RecBase {
    bool m_i, m_p, m_c0, m_c1;
    double r_i, h_i, x_prs, h_p, h_c0, h_c1;
    double f_rs, f_ds, f_prs, f_pds, sc_rs, sc_ds, sc_prs, sc_pds, gx_rs, c_i, c_c0, c_c1, pmn, pmx;
    __device__ LocalView(const DevView& v, const TreeRec& rec) : RecBase(v, rec) {
        const size_t Bs = (size_t)B;
        const bool hp = t.p >= 0, h0 = t.nch > 0, h1 = t.nch > 1;
        m_i = m_p = m_c0 = m_c1 = false;
        if (a.lanemask) {
            const unsigned long long w_i = mask_word(t.i), w_p = hp ? mask_word(t.p) : 0ull;
            const unsigned long long w_0 = h0 ? mask_word(t.ch0) : 0ull, w_1 = h1 ? mask_word(t.ch1) : 0ull;
            m_i = (w_i >> (b & 63)) & 1ull; m_p = (w_p >> (b & 63)) & 1ull; m_c0 = (w_0 >> (b & 63)) & 1ull; m_c1 = (w_1 >> (b & 63)) & 1ull;
        }
        r_i = t.rs >= 0 ? a.X[(unsigned)t.rs * (unsigned)B + (unsigned)b] : a.t.rfix[t.i];
        h_i = t.ds >= 0 ? a.X[(unsigned)t.ds * (unsigned)B + (unsigned)b] : a.t.hfix[t.i];
        x_prs = t.prs >= 0 ? a.X[(unsigned)t.prs * (unsigned)B + (unsigned)b] : 0.0;
        h_p = t.pds >= 0 ? a.X[(unsigned)t.pds * (unsigned)B + (unsigned)b] : (hp ? a.t.hfix[t.p] : 0.0);
        h_c0 = t.cds0 >= 0 ? a.X[(unsigned)t.cds0 * (unsigned)B + (unsigned)b] : (h0 ? a.t.hfix[t.ch0] : 0.0);
        h_c1 = t.cds1 >= 0 ? a.X[(unsigned)t.cds1 * (unsigned)B + (unsigned)b] : (h1 ? a.t.hfix[t.ch1] : 0.0);
        f_rs = t.rs >= 0 ? a.Minv[(unsigned)t.rs * (unsigned)B + (unsigned)b] : 0.0;
        f_ds = t.ds >= 0 ? a.Minv[(unsigned)t.ds * (unsigned)B + (unsigned)b] : 0.0;
        f_prs = t.prs >= 0 ? a.Minv[(unsigned)t.prs * (unsigned)B + (unsigned)b] : 0.0;
        f_pds = t.pds >= 0 ? a.Minv[(unsigned)t.pds * (unsigned)B + (unsigned)b] : 0.0;
        gx_rs = t.rs >= 0 ? a.Gx[(unsigned)t.rs * (unsigned)B + (unsigned)b] : 0.0;
        sc_rs = t.rs >= 0 ? a.sc[t.rs] : 0.0;
        sc_ds = t.ds >= 0 ? a.sc[t.ds] : 0.0;
        sc_prs = t.prs >= 0 ? a.sc[t.prs] : 0.0;
        sc_pds = t.pds >= 0 ? a.sc[t.pds] : 0.0;
        c_i = a.t.c[t.i];
        c_c0 = h0 ? a.t.c[t.ch0] : 0.0;
        c_c1 = h1 ? a.t.c[t.ch1] : 0.0;
        pmn = a.pmin_n[t.i];
        pmx = a.pmax_n[t.i];
    }
    __device__ bool masked(int n) const {
        if (!a.lanemask) return false;
        return n == t.i ? m_i : (n == t.p ? m_p : (n == t.ch0 ? m_c0 : (n == t.ch1 ? m_c1 : DevView::masked(n))));
    }
    __device__ double x(int row) const {
        return row == t.rs ? r_i : (row == t.ds ? h_i : (row == t.prs ? x_prs : (row == t.pds ? h_p : (row == t.cds0 ? h_c0 : (row == t.cds1 ? h_c1 : DevView::x(row))))));
    }
    __device__ double rfix(int n) const { return (n == t.i && t.rs < 0) ? r_i : DevView::rfix(n); }
    __device__ double hfix(int n) const {
        if (n == t.i && t.ds < 0) return h_i;
        if (n == t.p && t.pds < 0) return h_p;
        if (n == t.ch0 && t.cds0 < 0) return h_c0;
        if (n == t.ch1 && t.cds1 < 0) return h_c1;
        return DevView::hfix(n);
    }
    __device__ bool fr(int row) const {
        return (row == t.rs ? f_rs : (row == t.ds ? f_ds : (row == t.prs ? f_prs : (row == t.pds ? f_pds : a.Minv[(unsigned)row * (unsigned)B + (unsigned)b])))) > 0.0;
    }
    __device__ double sc(int row) const {
        return row == t.rs ? sc_rs : (row == t.ds ? sc_ds : (row == t.prs ? sc_prs : (row == t.pds ? sc_pds : DevView::sc(row))));
    }
    __device__ double gx(int row) const { return row == t.rs ? gx_rs : DevView::gx(row); }
    __device__ double c(int n) const { return n == t.i ? c_i : (n == t.ch0 ? c_c0 : (n == t.ch1 ? c_c1 : DevView::c(n))); }
    __device__ double pmin(int n) const { return n == t.i ? pmn : DevView::pmin(n); }
    __device__ double pmax(int n) const { return n == t.i ? pmx : DevView::pmax(n); }
};

// forward substitution (the solve inside a CG iteration), optionally preceded by the pivot step of the factorisation
// (PIVOT: the node's own local block and the children's T are loaded with everything else)
template <bool PIVOT>
struct UpView : RecBase {
    bool m_i;
    double f_rs, f_ds, rhs_rs, rhs_ds, k0[4], k1[4], y0[2], y1[2];
    double loc[PIVOT ? 7 : 1], t0[PIVOT ? 3 : 1], t1[PIVOT ? 3 : 1];
    __device__ UpView(const DevView& v, const TreeRec& rec) : RecBase(v, rec) {
        const size_t Bs = (size_t)B, n = (size_t)a.t.n;
        const bool h0 = t.nch > 0, h1 = t.nch > 1;
        m_i = a.lanemask ? ((mask_word(t.i) >> (b & 63)) & 1ull) : false;
        f_rs = t.rs >= 0 ? a.Minv[(unsigned)t.rs * (unsigned)B + (unsigned)b] : 0.0;
        f_ds = t.ds >= 0 ? a.Minv[(unsigned)t.ds * (unsigned)B + (unsigned)b] : 0.0;
        rhs_rs = t.rs >= 0 ? a.R[(unsigned)t.rs * (unsigned)B + (unsigned)b] : 0.0;
        rhs_ds = t.ds >= 0 ? a.R[(unsigned)t.ds * (unsigned)B + (unsigned)b] : 0.0;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            k0[q] = h0 ? __ldcg(a.TW + (size_t)(TW_K + q) * (n * Bs) + ((unsigned)t.ch0 * (unsigned)B + (unsigned)b)) : 0.0;
            k1[q] = h1 ? __ldcg(a.TW + (size_t)(TW_K + q) * (n * Bs) + ((unsigned)t.ch1 * (unsigned)B + (unsigned)b)) : 0.0;
        }
#pragma unroll
        for (int q = 0; q < 2; q++) {
            y0[q] = h0 ? __ldcg(a.TW + (size_t)(TW_Y + q) * (n * Bs) + ((unsigned)t.ch0 * (unsigned)B + (unsigned)b)) : 0.0;
            y1[q] = h1 ? __ldcg(a.TW + (size_t)(TW_Y + q) * (n * Bs) + ((unsigned)t.ch1 * (unsigned)B + (unsigned)b)) : 0.0;
        }
        if (PIVOT) {
#pragma unroll
            for (int q = 0; q < 7; q++) loc[q] = __ldcg(a.TW + (size_t)(TW_L + q) * (n * Bs) + ((unsigned)t.i * (unsigned)B + (unsigned)b));
#pragma unroll
            for (int q = 0; q < 3; q++) {
                t0[q] = h0 ? __ldcg(a.TW + (size_t)(TW_T + q) * (n * Bs) + ((unsigned)t.ch0 * (unsigned)B + (unsigned)b)) : 0.0;
                t1[q] = h1 ? __ldcg(a.TW + (size_t)(TW_T + q) * (n * Bs) + ((unsigned)t.ch1 * (unsigned)B + (unsigned)b)) : 0.0;
            }
        }
    }
    __device__ bool masked(int n) const { return n == t.i ? m_i : DevView::masked(n); }
    __device__ bool fr(int row) const { return (row == t.rs ? f_rs : (row == t.ds ? f_ds : a.Minv[(unsigned)row * (unsigned)B + (unsigned)b])) > 0.0; }
    __device__ double rhs(int row) const { return row == t.rs ? rhs_rs : (row == t.ds ? rhs_ds : DevView::rhs(row)); }
    __device__ double tw(int q, int n) const {
        if (PIVOT && n == t.i && q < TW_L + 7) return loc[q - TW_L];          // tree_pivot_node reads its local block first
        if (q >= TW_K && q < TW_T) {
            if (n == t.ch0) return k0[q - TW_K];
            if (n == t.ch1) return k1[q - TW_K];
        }
        if (PIVOT && q >= TW_T && q < TW_Y) {
            if (n == t.ch0) return t0[q - TW_T];
            if (n == t.ch1) return t1[q - TW_T];
        }
        if (q >= TW_Y) {
            if (n == t.ch0) return y0[q - TW_Y];
            if (n == t.ch1) return y1[q - TW_Y];
        }
        return DevView::tw(q, n);
    }
};

// back substitution
struct DownView : RecBase {
    double sS[3], sK[4], sY[2], o_prs, o_pds;
    __device__ DownView(const DevView& v, const TreeRec& rec) : RecBase(v, rec) {
        const size_t Bs = (size_t)B, n = (size_t)a.t.n;
#pragma unroll
        for (int q = 0; q < 3; q++) sS[q] = __ldcg(a.TW + (size_t)(TW_S + q) * (n * Bs) + ((unsigned)t.i * (unsigned)B + (unsigned)b));
#pragma unroll
        for (int q = 0; q < 4; q++) sK[q] = __ldcg(a.TW + (size_t)(TW_K + q) * (n * Bs) + ((unsigned)t.i * (unsigned)B + (unsigned)b));
#pragma unroll
        for (int q = 0; q < 2; q++) sY[q] = __ldcg(a.TW + (size_t)(TW_Y + q) * (n * Bs) + ((unsigned)t.i * (unsigned)B + (unsigned)b));
        o_prs = t.prs >= 0 ? __ldcg(W + ((unsigned)t.prs * (unsigned)B + (unsigned)b)) : 0.0;
        o_pds = t.pds >= 0 ? __ldcg(W + ((unsigned)t.pds * (unsigned)B + (unsigned)b)) : 0.0;
    }
    __device__ double outv(int row) const { return row == t.prs ? o_prs : (row == t.pds ? o_pds : DevView::outv(row)); }
    __device__ double tw(int q, int n) const {
        if (n == t.i) {
            if (q < TW_K) return sS[q - TW_S];
            if (q < TW_T) return sK[q - TW_K];
            if (q >= TW_Y) return sY[q - TW_Y];
        }
        return DevView::tw(q, n);
    }
};

template <bool FACTOR>
__device__ __forceinline__ void tree_up_at(const TnArgs& a, const DevView& V, int k) {
    const TreeRec rec = tree_rec_load(a.lev_rec, k);
    UpView<FACTOR> U(V, rec);
    if (FACTOR) tree_pivot_node(U, rec.i);
    tree_up_node(U, rec.i);
}
__device__ __forceinline__ void tree_down_at(const TnArgs& a, const DevView& V, int k) {
    const TreeRec rec = tree_rec_load(a.lev_rec, k);
    DownView D(V, rec);
    tree_down_node(D, rec.i);
}

// the point-only part of the factorisation for every node (no dependence between nodes): one launch before the sweeps.
// REC: through the record view (few vector groups: the launch is latency bound); otherwise through the plain accessors,
// which cost fewer instructions and registers when there are enough warps to hide the load chains (measured: the views
// halve the throughput of the node-indexed kernels at thousands of vectors)
// (the plain variant at four blocks per SM - 64 registers, no spills: this kernel is bound by its dependent loads, long scoreboard
// 69 % of the stall samples; -2...3 % per optimisation.  Forcing the factorising sweeps to three blocks per SM spills and loses.)
template <bool REC>
__global__ void __launch_bounds__(32 * OPT_WY, REC ? 2 : 4) tn_tree_local_kernel(TnArgs a) {
    const NodeLane nl = node_lane(a);
    const int b = nl.b;
    if (!tree_lane_active<true>(a, b)) return;
    DevView V = {a, a.o.B, b, a.Pv, a.ZZ, a.lane_lambda ? a.lane_lambda[b] : a.lambda, a.pb, a.log_pen != 0, tree_lane_mu<true>(a, b)};
    const int k0 = blockIdx.y * TREE_NODES_PER_BLOCK, k1 = min(a.t.n, k0 + TREE_NODES_PER_BLOCK);
    // A node WITHOUT children (height level 0: the tips, half of the tree) depends on nobody: its pivot step and its forward
    // substitution are finished right here, and the factorising sweep starts at level 1 (one launch less, and the largest).
    for (int k = k0 + threadIdx.y * nl.npw + nl.sub; k < k1; k += OPT_WY * nl.npw) {
        if (REC) {
            const TreeRec rec = tree_rec_load(a.lev_rec, k);
            LocalView L(V, rec);
            tree_local_node(L, rec.i);
            if (rec.nch == 0) {
                tree_pivot_node(L, rec.i);
                tree_up_node(L, rec.i);
            }
        } else {
            tree_local_node(V, k);
            if (V.child_begin(k) == V.child_end(k)) {
                tree_pivot_node(V, k);
                tree_up_node(V, k);
            }
        }
    }
}

template <bool FACTOR>
__global__ void __launch_bounds__(32 * OPT_WY) tn_tree_up_kernel(TnArgs a, int level) {
    const NodeLane nl = node_lane(a);
    const int b = nl.b;
    if (!tree_lane_active<FACTOR>(a, b)) return;
    DevView V = {a, a.o.B, b, a.Pv, a.ZZ, a.lane_lambda ? a.lane_lambda[b] : a.lambda, a.pb, a.log_pen != 0, tree_lane_mu<FACTOR>(a, b)};
    const int k0 = a.lev_off[level] + blockIdx.y * TREE_NODES_PER_BLOCK;
    const int k1 = min(a.lev_off[level + 1], k0 + TREE_NODES_PER_BLOCK);
    for (int k = k0 + threadIdx.y * nl.npw + nl.sub; k < k1; k += OPT_WY * nl.npw) tree_up_at<FACTOR>(a, V, k);
}

template <bool FACTOR>
__global__ void __launch_bounds__(32 * OPT_WY) tn_tree_down_kernel(TnArgs a, int level) {
    const NodeLane nl = node_lane(a);
    const int b = nl.b;
    if (!tree_lane_active<FACTOR>(a, b)) return;
    DevView V = {a, a.o.B, b, a.Pv, a.ZZ, a.lane_lambda ? a.lane_lambda[b] : a.lambda, a.pb, a.log_pen != 0, tree_lane_mu<FACTOR>(a, b)};
    const int k0 = a.lev_off[level] + blockIdx.y * TREE_NODES_PER_BLOCK;
    const int k1 = min(a.lev_off[level + 1], k0 + TREE_NODES_PER_BLOCK);
    for (int k = k0 + threadIdx.y * nl.npw + nl.sub; k < k1; k += OPT_WY * nl.npw) tree_down_at(a, V, k);
}

// SEVERAL wide levels in ONE cooperative launch, a grid-wide barrier between levels (2.7 us on a B200 with 1,184 blocks)
// instead of a kernel boundary (10-12 us): the sweeps of a latency-bound optimiser step are 130 launches otherwise.
// Work item = (32-vector group, warp-slot of the level's node list); items are dealt round-robin to all warps of the grid.
// DOWN: levels l_last - 1 ... l_first, back substitution; otherwise l_first ... l_last - 1, [pivot +] forward substitution.
template <bool FACTOR, bool DOWN>
__global__ void __launch_bounds__(32 * OPT_WY) tn_tree_levels_kernel(TnArgs a, int l_first, int l_last, int groups) {
    cooperative_groups::grid_group grid = cooperative_groups::this_grid();
    const int sh = a.pack_shift, npw = 32 >> sh, sub = threadIdx.x >> sh, bl = threadIdx.x & ((1 << sh) - 1);
    const long long warp0 = (long long)blockIdx.x * OPT_WY + threadIdx.y, nwarps = (long long)gridDim.x * OPT_WY;
    for (int s = 0; s < l_last - l_first; s++) {
        const int level = DOWN ? l_last - 1 - s : l_first + s;
        const int k0 = a.lev_off[level], k1 = a.lev_off[level + 1];
        const long long items = (long long)((k1 - k0 + npw - 1) / npw) * groups;
        for (long long it = warp0; it < items; it += nwarps) {
            const int g = (int)(it % groups), k = k0 + (int)(it / groups) * npw + sub;
            const int b = g * 32 + bl;
            if (k < k1 && tree_lane_active<FACTOR>(a, b)) {
                DevView V = {a, a.o.B, b, a.Pv, a.ZZ, a.lane_lambda ? a.lane_lambda[b] : a.lambda, a.pb, a.log_pen != 0, tree_lane_mu<FACTOR>(a, b)};
                if (DOWN) tree_down_at(a, V, k);
                else tree_up_at<FACTOR>(a, V, k);
            }
        }
        if (s + 1 < l_last - l_first) grid.sync();
    }
}

// the narrow levels [nlev_wide, nlev): up to the root and back down inside one block per vector group
template <bool FACTOR>
__global__ void __launch_bounds__(32 * TREE_TOP_WY) tn_tree_top_kernel(TnArgs a) {
    const NodeLane nl = node_lane(a);
    const int b = nl.b;
    const bool active = tree_lane_active<FACTOR>(a, b);
    if (!__syncthreads_or(active)) return;
    DevView V = {a, a.o.B, b, a.Pv, a.ZZ, a.lane_lambda ? a.lane_lambda[min(b, a.o.B - 1)] : a.lambda, a.pb, a.log_pen != 0,
                 active ? tree_lane_mu<FACTOR>(a, b) : 0.0};
    const int first = threadIdx.y * nl.npw + nl.sub, stride = TREE_TOP_WY * nl.npw;
    for (int level = max(a.nlev_wide, FACTOR ? 1 : 0); level < a.nlev; level++) {    // level 0 of a factorising sweep: tn_tree_local_kernel
        if (active)
            for (int k = a.lev_off[level] + first; k < a.lev_off[level + 1]; k += stride) tree_up_at<FACTOR>(a, V, k);
        __syncthreads();
    }
    for (int level = a.nlev - 1; level >= a.nlev_wide; level--) {
        if (active)
            for (int k = a.lev_off[level] + first; k < a.lev_off[level + 1]; k += stride) tree_down_at(a, V, k);
        __syncthreads();
    }
}

// after the sweeps: [p = zz,] r.zz -> the state machine that the Jacobi path runs inside tn_setup / tn_update
template <bool FACTOR>
__global__ void __launch_bounds__(32 * OPT_WY) tn_pc_finish_kernel(TnArgs a) {
    const int ty = threadIdx.y;
    const NodeLane nl = node_lane(a);
    const int b = nl.b;
    const int B = a.o.B, P = a.o.P;
    const bool active = tree_lane_active<FACTOR>(a, b);
    if (!__syncthreads_or(active)) return;
    double v[1] = {0.0};
    if (active) {
        const int p0 = blockIdx.y * a.rows_per_chunk, p1 = min(P, p0 + a.rows_per_chunk);
        for (int p = p0 + ty * nl.npw + nl.sub; p < p1; p += OPT_WY * nl.npw) {
            const size_t i = (size_t)p * B + b;
            const double zz = a.ZZ[i];
            if (FACTOR) a.Pv[i] = zz;
            v[0] = fma(a.R[i], zz, v[0]);
        }
    }
    if (!group_reduce<1, 0u, 0u>(v, a.part_pc, a.cnt_pc, B, b, a.pack_shift)) return;
    if (ty == 0 && nl.sub == 0 && active) {
        if (FACTOR) {
            const double red[3] = {v[0], a.gmax_tmp[b], a.gg_tmp[b]};
            if (phase_setup(lane_of(a, b), a.Ft[b], red, a.o)) atomicAdd(a.ndone, 1);
        } else {
            phase_update(lane_of(a, b), v, a.o);
        }
    }
}

__global__ void __launch_bounds__(32 * OPT_WY) tn_dir_kernel(TnArgs a) {
    const NodeLane nl = node_lane(a);
    const int b = nl.b;
    const int B = a.o.B, P = a.o.P;
    if (b >= B) return;
    if (a.st[b] != TS_CG || !(a.act[b] & ACT_DIR)) return;
    const double beta = a.beta[b];
    const int p0 = blockIdx.y * a.rows_per_chunk, p1 = min(P, p0 + a.rows_per_chunk);
    for (int p = p0 + threadIdx.y * nl.npw + nl.sub; p < p1; p += OPT_WY * nl.npw) {
        const size_t i = (size_t)p * B + b;
        a.Pv[i] = fma(beta, a.Pv[i], a.ZZ[i]);
    }
}

__global__ void __launch_bounds__(32 * OPT_WY) tn_gd_kernel(TnArgs a) {
    const int ty = threadIdx.y;
    const NodeLane nl = node_lane(a);
    const int b = nl.b;
    const int B = a.o.B, P = a.o.P;
    int act = 0;
    if (b < B && a.st[b] == TS_CG) act = a.act[b];
    const bool active = (act & ACT_FIN) != 0;
    if (!__syncthreads_or(active)) return;
    double v[1] = {0.0};
    if (active) {
        const bool takep = (act & ACT_TAKEP) != 0;
        const int p0 = blockIdx.y * a.rows_per_chunk, p1 = min(P, p0 + a.rows_per_chunk);
        for (int p = p0 + ty * nl.npw + nl.sub; p < p1; p += OPT_WY * nl.npw) {
            const size_t i = (size_t)p * B + b;
            double s;
            if (takep) {
                s = a.Pv[i];
                a.S[i] = s;
            } else {
                s = a.S[i];
            }
            v[0] = fma(a.Gz[i], s, v[0]);
        }
    }
    if (!group_reduce<1, 0u, 0u>(v, a.part_gd, a.cnt_gd, B, b, a.pack_shift)) return;
    if (ty == 0 && nl.sub == 0 && active) a.gdtmp[b] = v[0];
}

__global__ void __launch_bounds__(32 * OPT_WY) tn_stepmax_kernel(TnArgs a) {
    const int ty = threadIdx.y;
    const NodeLane nl = node_lane(a);
    const int b = nl.b;
    const int B = a.o.B;
    const bool active = b < B && a.st[b] == TS_CG && (a.act[b] & ACT_FIN);
    if (!__syncthreads_or(active)) return;
    double v[1] = {1e300};
    if (active) {
        const int i0 = max(1, blockIdx.y * a.tm_nodes_per_chunk), i1 = min(a.t.n, (blockIdx.y + 1) * a.tm_nodes_per_chunk);
        for (int i = i0 + ty * nl.npw + nl.sub; i < i1; i += OPT_WY * nl.npw) {
            const int pa = a.t.parent[i];
            const int ri = a.t.dslot[i], rp = a.t.dslot[pa];
            double zi, di = 0.0, zp, dp = 0.0;
            if (ri >= 0) { const double sc = a.sc[ri]; zi = a.Z[(unsigned)ri * (unsigned)B + (unsigned)b] * sc; di = a.S[(unsigned)ri * (unsigned)B + (unsigned)b] * sc; } else zi = a.t.hfix[i];
            if (rp >= 0) { const double sc = a.sc[rp]; zp = a.Z[(unsigned)rp * (unsigned)B + (unsigned)b] * sc; dp = a.S[(unsigned)rp * (unsigned)B + (unsigned)b] * sc; } else zp = a.t.hfix[pa];
            v[0] = fmin(v[0], tplopt::branch_step_limit(zi, di, zp, dp));
            if (ri >= 0) v[0] = fmin(v[0], tplopt::bound_step_limit(a.Z[(unsigned)ri * (unsigned)B + (unsigned)b], a.S[(unsigned)ri * (unsigned)B + (unsigned)b], a.lo[ri], a.hi[ri]));
        }
        if (blockIdx.y == 0 && ty == 0 && nl.sub == 0) {            // the root's own date (no branch above it)
            const int r0 = a.t.dslot[0];
            if (r0 >= 0) v[0] = fmin(v[0], tplopt::bound_step_limit(a.Z[(unsigned)r0 * (unsigned)B + (unsigned)b], a.S[(unsigned)r0 * (unsigned)B + (unsigned)b], a.lo[r0], a.hi[r0]));
        }
    }
    if (!group_reduce<1, 0u, 1u>(v, a.part_tm, a.cnt_tm, B, b, a.pack_shift)) return;
    if (ty == 0 && nl.sub == 0 && active)
        if (phase_finish(lane_of(a, b), a.gdtmp[b], v[0], a.o)) atomicAdd(a.ndone, 1);
}

__global__ void __launch_bounds__(32 * OPT_WY) tn_final_kernel(TnArgs a, double* F) {
    const NodeLane nl = node_lane(a);
    const int b = nl.b;
    if (b >= a.o.B) return;
    if (blockIdx.y == 0 && threadIdx.y == 0 && nl.sub == 0) F[b] = a.F0[b];
    const int p0 = blockIdx.y * a.rows_per_chunk, p1 = min(a.o.P, p0 + a.rows_per_chunk);
    for (int p = p0 + threadIdx.y * nl.npw + nl.sub; p < p1; p += OPT_WY * nl.npw) {
        const size_t i = (size_t)p * a.o.B + b;
        a.X[i] = tplopt::to_param(p < a.o.nr, a.Z[i], a.sc[p]);
    }
}

}  // namespace tpltn
