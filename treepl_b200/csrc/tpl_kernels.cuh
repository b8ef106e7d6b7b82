// PL objective + gradient kernels (sm_100a), generation 1: "pull" formulation.
//
// Math (reference src/pl_calc_parallel.cpp):
//   branch i>=1, p = parent[i]:  d_i = h_p - h_i,  x = r_i d_i
//     objective   += -(c_i ln x - x - lf_i)                        calc_log_like        :821-849
//     roughness   += (r_i - r_p)^2           (p != root)           calc_roughness_penalty :852-881
//   rate adjoint   d_i - c_i/r_i + 2L(r_i - r_p)[p!=root] - 2L sum_ch (r_ch - r_i)       :757-811
//   date adjoint   -(r_i - c_i/d_i) + sum_ch (r_ch - c_ch/d_ch)                          :727-751
// Every term is 1-hop local (node, parent, children), so there is no level sweep: each
// (node, vector) pair is owned by exactly one thread which PULLS its parent's and children's
// rows, recomputes the children's 1/(r d), and writes its own two gradient rows.  No atomics,
// no inter-thread exchange => bitwise deterministic.  The root-children variance term and the
// calibration barrier (sparse, order dependent: the reference resets the accumulator on inf,
// :983-991) are applied per vector by the finalize kernel, which also reduces the per-tile
// objective partials in fixed order.
//
// Layout: X[p*B + b], G[p*B + b] (batch innermost).  Thread block = LX lanes x NY node rows;
// with LX = 32 a warp reads one 256-byte row per parameter (fully coalesced) and all topology
// loads are warp-uniform broadcasts.  With B = 1 (the optimiser callbacks) LX = 1 and a warp
// covers 32 consecutive nodes, whose parameter rows are consecutive in the reference layout.
#pragma once
#include "tpl_math.cuh"

namespace tpl {

constexpr double LARGE = 1e15;

struct DevTree {
    int n;
    const int* parent;            // [n]
    const int* rslot;             // [n] row of X with the node's rate, -1 = none
    const int* dslot;             // [n] row of X with the node's date, -1 = fixed
    const int* child_off;         // [n+1]
    const int* children;          // [n-1] Newick order
    const unsigned char* cv;      // [n] engine-level CV mask (pl_calc_parallel::cvnodes)
    const double* c;              // [n] char_durations
    const double* lf;             // [n] log_fact_char_durations
    const double* hfix;           // [n] member dates[]  (used when dslot < 0)
    const double* rfix;           // [n] member rates[]  (used when rslot < 0)
    // sparse date-bound rows (set_durations :887-896).  Only nodes whose [min,max] is NOT implied by
    // a neighbour are listed: h_i <= h_parent <= max_parent <= max_i makes an inherited max redundant
    // once every duration is checked to be >= 0, and likewise an inherited min through a child.
    int nbnd;
    const int* bnd_ds;            // [nbnd] date row of the node (always free)
    const double* bnd_min;        // [nbnd]
    const double* bnd_max;        // [nbnd]
    // sparse calibration barrier rows, ascending node order (calc_penalty :975-998)
    int nbar;
    const int* bar_ds;            // [nbar] date row of the node, -1 = fixed date
    const double* bar_hfix;       // [nbar] the fixed date when bar_ds < 0
    const double* bar_pmin;       // [nbar] -1 = none
    const double* bar_pmax;       // [nbar] -1 = none
    double sum_c;                 // sum of char_durations over ALL nodes (calc_lf_gradient :675)
    double dur0;                  // member durations[0] after set_durations' fix-up
};

struct EvalArgs {
    DevTree t;
    const double* X;
    double* G;
    double* F;
    int B;
    const unsigned long long* lanemask;   // [nchunks][mask_stride] word (b/64, node): bit b%64 = node masked for vector b
    int nchunks;
    int mask_stride;                      // >= n (padded to whole tiles)
    // layout of the objective partials: 0 = [ntiles][B] (generations 1, 2);
    // 1 = generation 3: entry ((cta*2 + slot)*p_nwarps + warp) holds 64 vectors of the chunk that CTA
    //     was working on
    int part_mode;
    int p_ncta, p_ntiles, p_nchunks, p_nwarps;
    const double* lane_lambda;            // [B] or null
    double lambda;
    double pb;                            // penalty_boundary
    double* part_ll;                      // [ntiles][B]
    double* part_su;                      // [ntiles][B]
    double* part_gr;                      // [ntiles][B]  (LF rate-gradient partials)
    int* part_bad;                        // [ntiles][B]
    int ntiles;
    int tile_nodes;
    const double2* logtab;                // [LOG_TABLE_N]
};

__device__ __forceinline__ bool in_range400(double v) {      // 2^-400 <= v <= 2^400, v > 0, finite
    unsigned e = ((unsigned)__double2hiint(v)) >> 20;
    return (e - 623u) < 800u;
}

__device__ __forceinline__ double log_any(double x, const double2* tab) {
    return in_range400(x) ? fast_log(x, tab) : log(x);
}

// MASKMODE 0: no CV mask anywhere (the common case); 1: engine-level cvnodes; 2: per-vector bit masks
template <int MASKMODE>
__device__ __forceinline__ bool is_masked(const EvalArgs& a, int node, int lane) {
    if (MASKMODE == 2) {
        unsigned long long w = a.lanemask[(size_t)(lane >> 6) * a.mask_stride + node];
        return (w >> (lane & 63)) & 1ull;
    } else if (MASKMODE == 1) {
        return a.t.cv[node] != 0;
    } else {
        return false;
    }
}

// Per-branch quantities.  dr = d(-loglike_i)/d r_i, delta = d(-loglike_i)/d d_i.
struct Branch {
    double ll, dr, delta;
};

// MODE 0: calc_pl / calc_pl_gradient semantics.  MODE 1: the AD expression (tape order).
// d has already been through the duration rule; dconst = duration replaced by the DBL_MIN constant.
template <int MODE, bool LF, bool WANT_LL>
__device__ __forceinline__ Branch branch_terms(double r, double d, double c, double lf, bool dconst,
                                               const double2* tab) {
    Branch b;
    if (in_range400(r) && in_range400(d)) {
        double xx = r * d;
        double inv = fast_rcp(xx);
        double ac = c * inv;
        b.dr = fma(-ac, d, d);            // d - c/r
        b.delta = fma(-ac, r, r);         // r - c/d
        if (WANT_LL) b.ll = fma(-c, fast_log(xx, tab), xx) + lf;
        else b.ll = 0.0;
    } else if (MODE == 0) {
        double xx = r * d;
        b.ll = 0.0;
        if (xx > 0.0) b.ll = -(c * log(xx) - xx - lf);                    // :834-835
        else if (xx == 0.0) b.ll = (c > 0.0) ? LARGE : 0.0;             // :836-840
        b.dr = -(c / r - d);                                             // :764
        b.delta = (c == 0.0) ? r : (-c / d + r);                         // :730-734
    } else {
        double xx = r * d;
        if (LF || xx > 0.0) {                                            // :414-417 ; the LF tape has no x>0 guard (:206-210)
            b.ll = -(c * log(xx) - xx - lf);
            double au = (1.0 / xx) * (-c);                               // reverse sweep, tape order
            b.dr = d * au + d;
            b.delta = dconst ? 0.0 : (r * au + r);
        } else {                                                         // :418-424
            b.ll = (c > 0.0) ? LARGE : 0.0;
            b.dr = 0.0;
            b.delta = 0.0;
        }
    }
    return b;
}

// duration rule of set_durations (:898-907) / the AD paths (:396-402)
template <int MODE, bool LF>
__device__ __forceinline__ double duration_rule(double hp, double h, int& bad, bool& dconst) {
    double d = hp - h;
    dconst = false;
    if (d < 0.0) bad = 1;
    if (MODE == 0) {
        if (d == 0.0 || d != d) d = DBLMIN;
    } else if (!LF) {
        if (d == 0.0) { d = DBLMIN; dconst = true; }
    }
    return d;
}

template <int MODE, bool LF, bool LOGPEN, bool WANT_G, int MASKMODE>
__global__ void __launch_bounds__(256) pl_main_kernel(const __grid_constant__ EvalArgs a) {
    __shared__ double2 s_tab[LOG_TABLE_N];
    __shared__ double s_red[3][256];
    __shared__ int s_bad[256];
    const int LX = blockDim.x, NY = blockDim.y;
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int tid = ty * LX + tx;
    for (int k = tid; k < LOG_TABLE_N; k += LX * NY) s_tab[k] = a.logtab[k];
    __syncthreads();

    const DevTree& t = a.t;
    const int B = a.B;
    const int lane = blockIdx.y * LX + tx;
    const bool active = lane < B;
    const int n0 = blockIdx.x * a.tile_nodes;
    const int n1 = min(t.n, n0 + a.tile_nodes);
    double acc_ll = 0.0, acc_su = 0.0, acc_gr = 0.0;
    int bad = 0;

    if (active) {
        const double lam = a.lane_lambda ? a.lane_lambda[lane] : a.lambda;
        const double lam2 = 2.0 * lam;
        const double* __restrict__ X = a.X;
        for (int i = n0 + ty; i < n1; i += NY) {
            const int rs = t.rslot[i], ds = t.dslot[i];
            const bool mi = is_masked<MASKMODE>(a, i, lane);
            // calc_pl scatters x into rates[i] whenever a slot exists, masked or not (:82-89).  In the per-vector form a
            // masked node's X row is the rate the node KEEPS (the member rate of the reference's fold engine, i.e. the
            // full-data optimum, main.cpp:681,702-704): it is read like any other rate but it is not a variable.
            const bool rfree = (rs >= 0) && !(MASKMODE == 2 && mi);
            const double r = (rs >= 0) ? X[(size_t)rs * B + lane] : t.rfix[i];
            const double h = (ds >= 0) ? X[(size_t)ds * B + lane] : t.hfix[i];
            if (MODE == 0) {                                             // calc_pl :74-78
                if (rfree && r < 0.0) bad = 1;
                if (ds >= 0 && h < 0.0) bad = 1;
            }
            double g_r = 0.0, g_h = 0.0;
            double lr = 0.0;                                             // ln r_i (analytic log_pen only)
            if (i != 0) {
                const int p = t.parent[i];
                const int prs = t.rslot[p], pds = t.dslot[p];
                const double hp = (pds >= 0) ? X[(size_t)pds * B + lane] : t.hfix[p];
                bool dconst;
                const double d = duration_rule<MODE, LF>(hp, h, bad, dconst);
                const double c = t.c[i];
                if (LF && MODE == 0) acc_gr += d;                        // sumtimes, all nodes (:676)
                if (!mi || (MODE == 0 && WANT_G && ds >= 0)) {
                    Branch b = branch_terms<MODE, LF, true>(r, d, c, t.lf[i], dconst, s_tab);
                    if (!mi) {
                        acc_ll += b.ll;
                        if (LF) { if (MODE == 1) acc_gr += b.dr; }
                        else g_r = b.dr;
                        g_h = -b.delta;
                    } else {
                        g_h = -b.delta;                                  // analytic: own term regardless of cv (:728-735)
                    }
                }
                if (!LF && !mi) {
                    if (p != 0) {
                        const double rp = (prs >= 0) ? X[(size_t)prs * B + lane] : t.rfix[p];
                        const double q = r - rp;
                        acc_su = fma(q, q, acc_su);                      // :862-864
                        if (WANT_G) {
                            if (LOGPEN && MODE == 0) {
                                lr = log_any(r, s_tab);
                                g_r += lam2 * (lr - log_any(rp, s_tab)) / r;     // :795-796
                            } else {
                                g_r = fma(lam2, q, g_r);                 // :798
                            }
                        }
                    } else if (WANT_G && LOGPEN && MODE == 0) {
                        lr = log_any(r, s_tab);
                    }
                }
            }
            if (WANT_G) {
                const int cb = t.child_off[i], ce = t.child_off[i + 1];
                double sum_delta = 0.0, sum_q = 0.0;
                const bool need_q = !LF && rfree && i != 0;
                if (ds >= 0 || need_q) {
                    for (int j = cb; j < ce; j++) {
                        const int ch = t.children[j];
                        if (is_masked<MASKMODE>(a, ch, lane)) continue;  // :738, :784
                        const int crs = t.rslot[ch], cds = t.dslot[ch];
                        const double rc = (crs >= 0) ? X[(size_t)crs * B + lane] : t.rfix[ch];
                        if (ds >= 0) {
                            const double hc = (cds >= 0) ? X[(size_t)cds * B + lane] : t.hfix[ch];
                            bool dconst;
                            int dummy = 0;
                            const double dc = duration_rule<MODE, LF>(h, hc, dummy, dconst);
                            Branch b = branch_terms<MODE, LF, false>(rc, dc, t.c[ch], 0.0, dconst, s_tab);
                            sum_delta += b.delta;
                        }
                        if (need_q) {
                            if (LOGPEN && MODE == 0) sum_q += (log_any(rc, s_tab) - lr) / r;   // :786, :802
                            else sum_q += rc - r;
                        }
                    }
                }
                if (!LF && rs >= 0) a.G[(size_t)rs * B + lane] = (rfree && !mi) ? fma(-lam2, sum_q, g_r) : 0.0;
                if (ds >= 0) a.G[(size_t)ds * B + lane] = g_h + sum_delta;
            }
        }
    }
    // fixed-order reduction over the NY node rows of the block
    s_red[0][tid] = acc_ll;
    s_red[1][tid] = acc_su;
    s_red[2][tid] = acc_gr;
    s_bad[tid] = bad;
    __syncthreads();
    if (ty == 0 && active) {
        double s0 = 0.0, s1 = 0.0, s2 = 0.0;
        int bb = 0;
        for (int y = 0; y < NY; y++) {
            s0 += s_red[0][y * LX + tx];
            s1 += s_red[1][y * LX + tx];
            s2 += s_red[2][y * LX + tx];
            bb |= s_bad[y * LX + tx];
        }
        const size_t o = (size_t)blockIdx.x * B + lane;
        a.part_ll[o] = s0;
        if (!LF) a.part_su[o] = s1;
        if (LF) a.part_gr[o] = s2;
        a.part_bad[o] = bb;
    }
}

// Finalize: per vector, reduce the tile partials in fixed order (32 segments of consecutive tiles,
// then the 32 segment sums in order), check the sparse date bounds, add the root-children variance
// term (:866-879) and the calibration barrier (:975-998 / :381-406), write F, and apply the
// matching sparse gradient corrections to rows the main kernel has already written.
constexpr int FIN_SEGS = 32;
template <int MODE, bool LF, bool LOGPEN, bool WANT_G, int MASKMODE>
__global__ void __launch_bounds__(32 * FIN_SEGS) pl_finalize_kernel(const __grid_constant__ EvalArgs a) {
    __shared__ double s_sum[3][FIN_SEGS][32];
    __shared__ int s_bad[FIN_SEGS][32];
    const DevTree& t = a.t;
    const int B = a.B;
    const int tx = threadIdx.x, seg = threadIdx.y;
    const int lane = blockIdx.x * 32 + tx;
    const bool active = lane < B;
    // ---- sparse rows, spread over the 32 segment threads of each vector ----
    // The barrier sum is order dependent in the reference (accumulator reset on inf, :983-991), so the
    // segment threads only FETCH: the dependent  slot -> X row  loads run 32-wide, the terms land in
    // shared memory, and one thread per vector then accumulates them in ascending node order.
    constexpr int FIN_BAR = 48;                                          // barrier rows staged in shared memory
    constexpr int FIN_RC = 8;                                            // root children staged in shared memory
    __shared__ double s_bh[FIN_BAR][32];                                 // date of barrier row k
    __shared__ double s_pmn[FIN_BAR], s_pmx[FIN_BAR];
    __shared__ double s_rc[FIN_RC][32];                                  // rate of root child j (as the objective sees it)
    __shared__ int s_rcs[FIN_RC];                                        // its rate row (-1: none)
    __shared__ unsigned char s_rcm[FIN_RC][32];                          // masked for this vector
    const double* __restrict__ X = a.X;
    int bnd_bad = 0;
    if (active) {
        for (int k = seg; k < t.nbnd; k += FIN_SEGS) {                   // set_durations :887-896
            const double h = X[(size_t)t.bnd_ds[k] * B + lane];          // fixed dates are checked on the host
            if (h < t.bnd_min[k] || h > t.bnd_max[k]) bnd_bad = 1;
        }
        if (!LF) {
            for (int k = seg; k < t.nbar && k < FIN_BAR; k += FIN_SEGS) {
                const int ds = t.bar_ds[k];
                s_bh[k][tx] = (ds >= 0) ? X[(size_t)ds * B + lane] : t.bar_hfix[k];
                if (tx == 0) { s_pmn[k] = t.bar_pmin[k]; s_pmx[k] = t.bar_pmax[k]; }
            }
            const int nrc = t.child_off[1] - t.child_off[0];
            for (int j = seg; j < nrc && j < FIN_RC; j += FIN_SEGS) {
                const int ch = t.children[t.child_off[0] + j];
                const bool mc = is_masked<MASKMODE>(a, ch, lane);
                const int crs = t.rslot[ch];
                s_rc[j][tx] = (crs >= 0) ? X[(size_t)crs * B + lane] : t.rfix[ch];
                s_rcm[j][tx] = mc ? 1 : 0;
                if (tx == 0) s_rcs[j] = crs;
            }
        }
    }

    // (everything above reads X only; the objective partials and gradient rows of the main kernel are read from here on)
    if (a.part_mode == 1) {
        // generation-3 partials: CTA c owned work items [c W / G, (c+1) W / G) of the chunk-major list
        // (chunk k = items [k ntiles, (k+1) ntiles)); slot 1 if the CTA had started in the previous chunk
        const long long W = (long long)a.p_nchunks * a.p_ntiles, G = a.p_ncta;
        const int k = lane >> 6, l = lane & 63;
        const long long c0 = (long long)k * a.p_ntiles, c1 = c0 + a.p_ntiles;
        const int len = (a.p_ncta + FIN_SEGS - 1) / FIN_SEGS;
        const int g0 = seg * len, g1 = min(a.p_ncta, g0 + len);
        double ll = 0.0, su = 0.0;
        int bad = 0;
        const bool small = (G + 1) * W < 0x7fffffffLL;        // 32-bit divisions (the usual case) cost a fifth of 64-bit ones
        if (active)
            for (int c = g0; c < g1; c++) {
                long long lo, hi;
                if (small) {
                    lo = (unsigned)(c * (int)W) / (unsigned)G;
                    hi = (unsigned)((c + 1) * (int)W) / (unsigned)G;
                } else {
                    lo = c * W / G;
                    hi = (c + 1) * W / G;
                }
                if (lo < c1 && hi > c0 && lo < hi) {
                    const size_t o = ((size_t)c * 2 + (lo < c0 ? 1 : 0)) * a.p_nwarps * 64 + l;
                    for (int wv = 0; wv < a.p_nwarps; wv++) {
                        ll += a.part_ll[o + (size_t)wv * 64];
                        su += a.part_su[o + (size_t)wv * 64];
                        bad |= a.part_bad[o + (size_t)wv * 64];
                    }
                }
            }
        s_sum[0][seg][tx] = ll;
        s_sum[1][seg][tx] = su;
        s_sum[2][seg][tx] = 0.0;
        s_bad[seg][tx] = bad;
    } else {
        const int len = (a.ntiles + FIN_SEGS - 1) / FIN_SEGS;
        const int k0 = seg * len, k1 = min(a.ntiles, k0 + len);
        double ll = 0.0, su = 0.0, gr = 0.0;
        int bad = 0;
        if (active) {
            // four independent chains so the loads of consecutive tiles overlap; fixed grouping
            double l4[4] = {0, 0, 0, 0}, s4[4] = {0, 0, 0, 0}, g4[4] = {0, 0, 0, 0};
            int k = k0;
            for (; k + 4 <= k1; k += 4) {
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const size_t o = (size_t)(k + u) * B + lane;
                    l4[u] += a.part_ll[o];
                    if (!LF) s4[u] += a.part_su[o];
                    if (LF) g4[u] += a.part_gr[o];
                    bad |= a.part_bad[o];
                }
            }
            for (; k < k1; k++) {
                const size_t o = (size_t)k * B + lane;
                l4[0] += a.part_ll[o];
                if (!LF) s4[0] += a.part_su[o];
                if (LF) g4[0] += a.part_gr[o];
                bad |= a.part_bad[o];
            }
            ll = (l4[0] + l4[1]) + (l4[2] + l4[3]);
            su = (s4[0] + s4[1]) + (s4[2] + s4[3]);
            gr = (g4[0] + g4[1]) + (g4[2] + g4[3]);
        }
        s_sum[0][seg][tx] = ll;
        s_sum[1][seg][tx] = su;
        s_sum[2][seg][tx] = gr;
        s_bad[seg][tx] = bad;
    }
    s_bad[seg][tx] |= bnd_bad;
    __syncthreads();
    __shared__ int s_dead[32];                                           // AD: barrier sum was inf/NaN -> no barrier gradient
    if (seg == 0 && active) {
        double ll = 0.0, su = 0.0, gr = 0.0;
        int bad = 0;
        for (int k = 0; k < FIN_SEGS; k++) {
            ll += s_sum[0][k][tx];
            su += s_sum[1][k][tx];
            gr += s_sum[2][k][tx];
            bad |= s_bad[k][tx];
        }
        double f;
        if (LF) {
            f = ll;
            if (WANT_G) {
                if (MODE == 0) {
                    const double rate = X[lane];                             // params[0]  (:671)
                    a.G[lane] = -(t.sum_c / rate - (gr + t.dur0));           // :678
                } else {
                    a.G[lane] = gr;
                }
            }
        } else {
            const double lam = a.lane_lambda ? a.lane_lambda[lane] : a.lambda;
            // ---- root children block ----
            const int cb = t.child_off[0], ce = t.child_off[1];
            double s = 0.0, ss = 0.0, m = 0.0;          // all children (objective; AD gradient)
            double su_s = 0.0, su_m = 0.0;              // unmasked children (analytic gradient :767-778)
            for (int j = cb; j < ce; j++) {
                double rc;
                bool mc;
                if (j - cb < FIN_RC) { rc = s_rc[j - cb][tx]; mc = s_rcm[j - cb][tx] != 0; }
                else {
                    const int ch = t.children[j];
                    mc = is_masked<MASKMODE>(a, ch, lane);
                    const int crs = t.rslot[ch];
                    rc = (crs >= 0) ? X[(size_t)crs * B + lane] : t.rfix[ch];
                }
                const double y = LOGPEN ? log(rc) : rc;
                s += y;
                ss += y * y;
                m += 1.0;
                if (!mc) { su_s += y; su_m += 1.0; }
            }
            const double rp = su + (ss - s * s / m) / m;                     // :878
            // ---- barrier ----
            double tpen = 0.0;
            for (int k = 0; k < t.nbar; k++) {
                double h;
                double pmn, pmx;
                if (k < FIN_BAR) { h = s_bh[k][tx]; pmn = s_pmn[k]; pmx = s_pmx[k]; }
                else {
                    const int ds = t.bar_ds[k];
                    h = (ds >= 0) ? X[(size_t)ds * B + lane] : t.bar_hfix[k];
                    pmn = t.bar_pmin[k]; pmx = t.bar_pmax[k];
                }
                if (MODE == 0) {
                    if (pmn != -1.0) { tpen += 1.0 / (h - pmn); if (isinf(tpen)) tpen = 0.0; }    // :981-985
                    if (pmx != -1.0) { tpen += 1.0 / (pmx - h); if (isinf(tpen)) tpen = 0.0; }    // :988-992
                } else {
                    if (pmn != -1.0) { const double q = h - pmn; if (q > 0.0) tpen += 1.0 / q; }  // :381-384
                    if (pmx != -1.0) { const double q = pmx - h; if (q > 0.0) tpen += 1.0 / q; }  // :391-394
                }
            }
            bool bar_dead = false;
            if (MODE == 1 && (isinf(tpen) || isnan(tpen))) { tpen = 0.0; bar_dead = true; }      // :405-406
            s_dead[tx] = bar_dead ? 1 : 0;
            if (MODE == 0) f = (ll + a.pb * tpen) + lam * rp;                // :121-123
            else f = (a.pb * tpen + ll) + lam * rp;                          // :407, :463
            if (WANT_G) {
                for (int j = cb; j < ce; j++) {
                    int crs;
                    bool mc;
                    double rc;
                    if (j - cb < FIN_RC) { crs = s_rcs[j - cb]; mc = s_rcm[j - cb][tx] != 0; rc = s_rc[j - cb][tx]; }
                    else {
                        const int ch = t.children[j];
                        mc = is_masked<MASKMODE>(a, ch, lane);
                        crs = t.rslot[ch];
                        rc = (crs >= 0 && !mc) ? X[(size_t)crs * B + lane] : 0.0;
                    }
                    if (crs < 0 || mc) continue;
                    double add;
                    if (MODE == 0) {
                        const double mean = su_s / su_m;
                        if (LOGPEN) add = (2.0 * lam / rc) * (log(rc) - mean) / su_m;            // :780
                        else add = (2.0 * lam) * (rc - mean) / su_m;                             // :782
                    } else {
                        const double y = LOGPEN ? log(rc) : rc;
                        add = lam * (2.0 * y - 2.0 * s / m) / m;
                        if (LOGPEN) add /= rc;
                    }
                    a.G[(size_t)crs * B + lane] += add;
                }
            }
        }
        a.F[lane] = bad ? LARGE : f;
    }
    if (!LF && MODE == 1 && WANT_G) {
        // barrier gradient (AD expression only): one row per (segment thread, barrier row)
        __syncthreads();
        if (active && !s_dead[tx])
            for (int k = seg; k < t.nbar; k += FIN_SEGS) {
                const int ds = t.bar_ds[k];
                if (ds < 0) continue;
                const double h = (k < FIN_BAR) ? s_bh[k][tx] : X[(size_t)ds * B + lane];
                const double pmn = t.bar_pmin[k], pmx = t.bar_pmax[k];
                double gb = 0.0;
                if (pmn != -1.0) { const double q = h - pmn; if (q > 0.0) gb -= 1.0 / (q * q); }
                if (pmx != -1.0) { const double q = pmx - h; if (q > 0.0) gb += 1.0 / (q * q); }
                a.G[(size_t)ds * B + lane] += a.pb * gb;
            }
    }
}

// test hook: the fast math on its ordinary range (tests/test_engine_gpu.py)
__global__ void debug_math_kernel(int n, const double* x, double* lo, double* rc, const double2* tab) {
    __shared__ double2 s_tab[LOG_TABLE_N];
    for (int k = threadIdx.x; k < LOG_TABLE_N; k += blockDim.x) s_tab[k] = tab[k];
    __syncthreads();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        lo[i] = log_any(x[i], s_tab);
        rc[i] = in_range400(x[i]) ? fast_rcp(x[i]) : 1.0 / x[i];
    }
}

}  // namespace tpl
