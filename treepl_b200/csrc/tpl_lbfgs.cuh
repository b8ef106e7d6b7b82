// Streaming kernels of the device-resident batched L-BFGS optimiser (sm_100a).  See tpl_lbfgs_core.h for the
// method and for what it replaces in the reference (optimize_plcp_tnc / optimize_plcp_nlopt / optimize_full).
//
// All [P][B] operands keep the engine's layout (batch innermost): thread (tx, ty) of a block owns vector
// b = blockIdx.x*32 + tx and every WY-th row of the block's row chunk, so a warp reads one 256-byte row per
// operand (coalesced) and per-vector control data is one load per thread.  All HBM-bound:
//   opt_update_kernel : per accepted vector reads Z, D, X, Gx, Gz and the 2m history rows once, writes Z, Gz and the
//                       new (s, y) pair, and accumulates the 6m+2 dot products the Gram matrix needs  ((2m+9) * 8 B / element)
//   opt_lane_kernel   : per vector: fixed-order reduction of the chunk partials, Gram update, two-loop recursion
//                       on coefficients, Armijo / convergence state machine (tpl_lbfgs_core.h)
//   opt_direction_kernel : D = -free * sum_j delta_j b_j  (reads 2m+2 rows) for vectors with a new direction
//   opt_stepmax_kernel   : ratio test over the tree's branches: a new direction's first step stays inside the
//                          feasible region (all durations > 0) instead of discovering the boundary by backtracking
//   opt_trial_kernel     : the trial point X = to_param(z + t D) for every running vector
// Deterministic: static row -> thread mapping, fixed-order reductions, no floating-point atomics.
#pragma once
#include "tpl_lbfgs_core.h"

namespace tplopt {

struct OptArgs {
    OptConst o;
    double *Z, *Gz, *D, *X;          // [P][B]
    const double* Gx;                // [P][B] gradient of the trial point in the ORIGINAL parameters (PL kernels)
    double *S, *Y;                   // [HM][P][B]
    const double* sc;                // [P] per-row scale: the optimiser's variable is z = w / sc, w = ln r | h
    const double *lo, *hi;           // [P] bounds of z (bounds of w divided by sc)
    int *st, *ls, *head, *stall, *flag, *newdir, *iters, *nfail;
    unsigned* valid;
    double *F0, *t, *gd, *gamma, *gmax, *tlog;
    const double* Ft;                // [B] objective of the trial point
    double *rho, *delta, *M, *part;  // [HM][B], [NB][B], [NB*NB][B], [nchunks][NACC][B]
    int nchunks, rows_per_chunk;
    int dir_chunks, dir_rows;        // row chunking of the direction kernel (one full wave of its own occupancy)
    int light_chunks, light_rows;    // finer row chunking for the light streaming kernels (init, trial, final)
    double* part_red;                // [NACC][B] reduced partials (opt_reduce_kernel)
    const double* lane_part;         // what the lane kernel reduces: part (few chunks) or part_red
    int lane_nchunks;
    int* ndone;
    // ratio test (largest step that keeps every branch duration positive)
    int n;                           // nodes
    const int *parent, *dslot;       // [n] parent node, date row of the node (-1 = fixed date)
    const double* hfix;              // [n] fixed dates
    double* tpart;                   // [tchunks][B] per-chunk minima of the ratio test
    int* tcount;                     // [ceil(B/32)] arrival counters of the ratio test (zero between launches)
    int tchunks, nodes_per_chunk;
    double boundary_frac;            // fraction of the distance to the boundary a first trial step may take
};

constexpr int OPT_WY = 8;            // row sub-lanes per block (block = 32 x 8)
constexpr int OPT_LX = 16;           // vectors per block of the lane kernel
constexpr int OPT_NY = 32;

// Z = clamp(transform^-1(X0)), D = 0
__global__ void __launch_bounds__(32 * OPT_WY) opt_init_kernel(OptArgs a) {
    const int b = blockIdx.x * 32 + threadIdx.x;
    if (b >= a.o.B) return;
    const int p0 = blockIdx.y * a.light_rows, p1 = min(a.o.P, p0 + a.light_rows);
    for (int p = p0 + threadIdx.y; p < p1; p += OPT_WY) {
        const size_t i = (size_t)p * a.o.B + b;
        const double x = a.X[i];
        const double w = p < a.o.nr ? log(x) : x;
        a.Z[i] = fmin(fmax(w / a.sc[p], a.lo[p]), a.hi[p]);
        a.D[i] = 0.0;
    }
}

// Block = 32 vectors x HM threads.  Rows are processed in tiles of HM rows, two phases per tile:
//   phase 1  thread (tx, r) owns ROW r of the tile: new gradient in the transformed variables, accepted point,
//            the new (s, y) pair (written to slot h), projected gradient gp; s, y, gp go to shared memory
//   phase 2  thread (tx, k) owns history SLOT k: streams S_k, Y_k of the tile's rows (2 HM independent loads in
//            flight) and accumulates its six dot products against the tile's s, y, gp
// so a thread carries 6 accumulators instead of 6 HM + 2 (no register-bound occupancy), and the per-chunk
// partials of slot k are complete in one thread (fixed order, no cross-thread reduction).
template <int HM>
__global__ void __launch_bounds__(32 * HM) opt_update_kernel(OptArgs a) {
    typedef Layout<HM> L;
    constexpr int NACC = L::NACC;
    __shared__ double sh_s[2][HM][32], sh_y[2][HM][32], sh_g[2][HM][32];
    const int tx = threadIdx.x, k = threadIdx.y;
    const int b = blockIdx.x * 32 + tx;
    const int B = a.o.B, P = a.o.P, nr = a.o.nr;
    const size_t PB = (size_t)P * B;
    int dec = DEC_SKIP, h = 0;
    unsigned valid = 0;
    double t = 0.0;
    if (b < B) {
        t = a.t[b];
        dec = decide(a.st[b], a.Ft[b], a.F0[b], t, a.gd[b], a.o.c1);
        h = a.head[b];
        valid = a.valid[b];
    }
    const bool acc_lane = dec == DEC_ACCEPT;
    const bool active = acc_lane || dec == DEC_INIT;
    if (!__syncthreads_or(active)) return;          // no vector of this block moved: partials stay unused
    const bool cur = k == h;
    const bool my_slot = acc_lane && (cur || ((valid >> k) & 1u));
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0, a4 = 0.0, a5 = 0.0, gg = 0.0, gmx = 0.0;
    const int p0 = blockIdx.y * a.rows_per_chunk, p1 = min(P, p0 + a.rows_per_chunk);
    const double* Sk = a.S + (size_t)k * PB + b;
    const double* Yk = a.Y + (size_t)k * PB + b;
    int buf = 0;
    // phase-1 operands of the NEXT tile are fetched before phase 2 of the current one (software pipeline)
    double nx = 0.0, ngx = 0.0, nz = 0.0, nd = 0.0, ngz = 0.0;
    auto fetch = [&](int p) {
        if (active && p < p1) {
            const size_t i = (size_t)p * B + b;
            nx = a.X[i];
            ngx = a.Gx[i];
            nz = a.Z[i];
            if (acc_lane) {
                nd = a.D[i];
                ngz = a.Gz[i];
            }
        }
    };
    fetch(p0 + k);
    for (int r0 = p0; r0 < p1; r0 += HM, buf ^= 1) {
        {   // ---- phase 1: row r0 + k ----
            const int p = r0 + k;
            double s = 0.0, y = 0.0, gp = 0.0;
            if (active && p < p1) {
                const size_t i = (size_t)p * B + b;
                const double lo = a.lo[p], hi = a.hi[p], sc = a.sc[p];
                double z = nz;
                const double gzn = chain_grad(p < nr, ngx, nx, sc);
                if (acc_lane) {
                    const double zt = trial_point(z, t, nd, lo, hi);
                    s = zt - z;
                    y = gzn - ngz;
                    z = zt;
                    a.Z[i] = zt;
                    a.S[(size_t)h * PB + i] = s;
                    a.Y[(size_t)h * PB + i] = y;
                }
                a.Gz[i] = gzn;
                gp = is_free(z, gzn, lo, hi) ? gzn : 0.0;
                gg = fma(gp, gp, gg);
                gmx = fmax(gmx, fabs(gp));
            }
            sh_s[buf][k][tx] = s;
            sh_y[buf][k][tx] = y;
            sh_g[buf][k][tx] = gp;
        }
        fetch(r0 + HM + k);
        __syncthreads();
        if (my_slot) {   // ---- phase 2: slot k against the tile's rows ----
            double Sv[HM], Yv[HM];
#pragma unroll
            for (int r = 0; r < HM; r++) {
                Sv[r] = 0.0;
                Yv[r] = 0.0;
                if (!cur && r0 + r < p1) {
                    Sv[r] = Sk[(size_t)(r0 + r) * B];
                    Yv[r] = Yk[(size_t)(r0 + r) * B];
                }
            }
#pragma unroll
            for (int r = 0; r < HM; r++) {
                const double s = sh_s[buf][r][tx], y = sh_y[buf][r][tx], gp = sh_g[buf][r][tx];
                const double sv = cur ? s : Sv[r], yv = cur ? y : Yv[r];
                a0 = fma(s, sv, a0);
                a1 = fma(s, yv, a1);
                a2 = fma(y, sv, a2);
                a3 = fma(y, yv, a3);
                a4 = fma(gp, sv, a4);
                a5 = fma(gp, yv, a5);
            }
        }
        // double buffering: the next tile's phase 1 writes the other buffer, and its barrier orders this
        // tile's phase-2 reads before the buffer is written again
    }
    __syncthreads();
    double* part = a.part + (size_t)blockIdx.y * NACC * B;
    if (b < B) {
        part[(size_t)(6 * k + 0) * B + b] = a0;
        part[(size_t)(6 * k + 1) * B + b] = a1;
        part[(size_t)(6 * k + 2) * B + b] = a2;
        part[(size_t)(6 * k + 3) * B + b] = a3;
        part[(size_t)(6 * k + 4) * B + b] = a4;
        part[(size_t)(6 * k + 5) * B + b] = a5;
    }
    sh_s[0][k][tx] = gg;
    sh_y[0][k][tx] = gmx;
    __syncthreads();
    if (k == 0 && b < B) {
        for (int r = 1; r < HM; r++) {
            gg += sh_s[0][r][tx];
            gmx = fmax(gmx, sh_y[0][r][tx]);
        }
        part[(size_t)L::A_GG * B + b] = gg;
        part[(size_t)L::A_GMAX * B + b] = gmx;
    }
}

// fixed-order reduction of the chunk partials, one block per (32 vectors, accumulator): used when there are many chunks
template <int HM>
__global__ void __launch_bounds__(32 * OPT_WY) opt_reduce_kernel(OptArgs a) {
    typedef Layout<HM> L;
    constexpr int NACC = L::NACC;
    __shared__ double red[OPT_WY][32];
    const int tx = threadIdx.x, ty = threadIdx.y, k = blockIdx.y;
    const int b = blockIdx.x * 32 + tx;
    const int B = a.o.B;
    double v = 0.0;
    if (b < B) {
        if (k == L::A_GMAX) {
            for (int c = ty; c < a.nchunks; c += OPT_WY) v = fmax(v, a.part[((size_t)c * NACC + k) * B + b]);
        } else {
            for (int c = ty; c < a.nchunks; c += OPT_WY) v += a.part[((size_t)c * NACC + k) * B + b];
        }
    }
    red[ty][tx] = v;
    __syncthreads();
    if (ty == 0 && b < B) {
        if (k == L::A_GMAX) {
            for (int y = 1; y < OPT_WY; y++) v = fmax(v, red[y][tx]);
        } else {
            for (int y = 1; y < OPT_WY; y++) v += red[y][tx];
        }
        a.part_red[(size_t)k * B + b] = v;
    }
}

template <int HM>
__global__ void __launch_bounds__(OPT_LX * OPT_NY) opt_lane_kernel(OptArgs a) {
    typedef Layout<HM> L;
    constexpr int NACC = L::NACC, NB = L::NB;
    extern __shared__ double sm[];
    double* sacc = sm;                         // [NACC][LX]
    double* sM = sm + NACC * OPT_LX;           // [NB*NB][LX]
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int b = blockIdx.x * OPT_LX + tx;
    const int B = a.o.B;
    if (b < B) {
        for (int k = ty; k < NACC; k += OPT_NY) {
            double v = 0.0;
            if (k == L::A_GMAX) {
                for (int c = 0; c < a.lane_nchunks; c++) v = fmax(v, a.lane_part[((size_t)c * NACC + k) * B + b]);
            } else {
                for (int c = 0; c < a.lane_nchunks; c++) v += a.lane_part[((size_t)c * NACC + k) * B + b];
            }
            sacc[k * OPT_LX + tx] = v;
        }
        for (int k = ty; k < NB * NB; k += OPT_NY) sM[k * OPT_LX + tx] = a.M[(size_t)k * B + b];
    }
    __syncthreads();
    if (ty == 0 && b < B) {
        LaneIO io;
        io.st = a.st + b; io.ls = a.ls + b; io.head = a.head + b; io.stall = a.stall + b; io.flag = a.flag + b;
        io.newdir = a.newdir + b; io.iters = a.iters + b; io.nfail = a.nfail + b; io.valid = a.valid + b;
        io.F0 = a.F0 + b; io.t = a.t + b; io.gd = a.gd + b; io.gamma = a.gamma + b; io.gmax = a.gmax + b; io.tlog = a.tlog + b;
        io.rho = a.rho + b; io.rho_s = B;
        io.delta = a.delta + b; io.delta_s = B;
        io.M = sM + tx; io.M_s = OPT_LX;
        io.acc = sacc + tx; io.acc_s = OPT_LX;
        if (lane_step<HM>(io, a.Ft[b], a.o)) atomicAdd(a.ndone, 1);
    }
    __syncthreads();
    if (b < B)
        for (int k = ty; k < NB * NB; k += OPT_NY) a.M[(size_t)k * B + b] = sM[k * OPT_LX + tx];
}

// D = -free * sum_j delta_j b_j for the vectors that got a new direction
template <int HM>
__global__ void __launch_bounds__(32 * OPT_WY) opt_direction_kernel(OptArgs a) {
    typedef Layout<HM> L;
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int b = blockIdx.x * 32 + tx;
    const int B = a.o.B, P = a.o.P;
    if (b >= B) return;
    if (a.st[b] == ST_DONE || a.newdir[b] == 0) return;
    const size_t PB = (size_t)P * B;
    const unsigned valid = a.valid[b];
    double dl[L::NB];
#pragma unroll
    for (int j = 0; j < L::NB; j++) dl[j] = a.delta[(size_t)j * B + b];
    const int p0 = blockIdx.y * a.dir_rows, p1 = min(P, p0 + a.dir_rows);
    // two rows per iteration: 2 (2 m + 2) independent loads in flight per thread
    for (int p = p0 + ty; p < p1; p += 2 * OPT_WY) {
        const int pb = p + OPT_WY;
        const bool two = pb < p1;
        const size_t i = (size_t)p * B + b, j = two ? (size_t)pb * B + b : i;
        const double z0 = a.Z[i], g0 = a.Gz[i], z1 = a.Z[j], g1 = a.Gz[j];
        double q0 = dl[L::IG] * g0, q1 = dl[L::IG] * g1;
        double s0[HM], y0[HM], s1[HM], y1[HM];
#pragma unroll
        for (int k = 0; k < HM; k++) {
            s0[k] = y0[k] = s1[k] = y1[k] = 0.0;
            if ((valid >> k) & 1u) {
                s0[k] = a.S[(size_t)k * PB + i];
                y0[k] = a.Y[(size_t)k * PB + i];
                s1[k] = a.S[(size_t)k * PB + j];
                y1[k] = a.Y[(size_t)k * PB + j];
            }
        }
#pragma unroll
        for (int k = 0; k < HM; k++) {
            q0 = fma(dl[k], s0[k], q0);
            q0 = fma(dl[HM + k], y0[k], q0);
            q1 = fma(dl[k], s1[k], q1);
            q1 = fma(dl[HM + k], y1[k], q1);
        }
        a.D[i] = is_free(z0, g0, a.lo[p], a.hi[p]) ? -q0 : 0.0;
        if (two) a.D[j] = is_free(z1, g1, a.lo[pb], a.hi[pb]) ? -q1 : 0.0;
    }
}

// Ratio test for vectors with a new direction: the largest t for which every branch duration
// h_parent - h_node of z + t D stays positive (the objective is +1e15 beyond it, set_durations
// pl_calc_parallel.cpp:883-910).  Per-chunk minima; the LAST block of a 32-vector group to finish combines them
// (min is order independent, so the result is deterministic) and shortens the unit step to boundary_frac of
// the distance to the boundary.
__global__ void __launch_bounds__(32 * OPT_WY) opt_stepmax_kernel(OptArgs a) {
    __shared__ double red[OPT_WY][32];
    __shared__ int s_last;
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int b = blockIdx.x * 32 + tx;
    const int B = a.o.B;
    const bool active = b < B && a.st[b] != ST_DONE && a.newdir[b] != 0;
    if (!__syncthreads_or(active)) return;          // uniform over blockIdx.y: the whole group skips
    double tmin = 1e300;
    if (active) {
        const int i0 = max(1, blockIdx.y * a.nodes_per_chunk), i1 = min(a.n, (blockIdx.y + 1) * a.nodes_per_chunk);
        for (int i = i0 + ty; i < i1; i += OPT_WY) {
            const int pa = a.parent[i];
            const int ri = a.dslot[i], rp = a.dslot[pa];
            double zi, di = 0.0, zp, dp = 0.0;
            // in date units: h = z sc, dh/dt = d sc
            if (ri >= 0) { const double sc = a.sc[ri]; zi = a.Z[(size_t)ri * B + b] * sc; di = a.D[(size_t)ri * B + b] * sc; } else zi = a.hfix[i];
            if (rp >= 0) { const double sc = a.sc[rp]; zp = a.Z[(size_t)rp * B + b] * sc; dp = a.D[(size_t)rp * B + b] * sc; } else zp = a.hfix[pa];
            tmin = fmin(tmin, branch_step_limit(zi, di, zp, dp));
        }
    }
    red[ty][tx] = tmin;
    __syncthreads();
    if (ty == 0 && b < B) {
        for (int y = 1; y < OPT_WY; y++) tmin = fmin(tmin, red[y][tx]);
        a.tpart[(size_t)blockIdx.y * B + b] = tmin;
    }
    __threadfence();
    __syncthreads();
    if (tx == 0 && ty == 0) s_last = atomicAdd(a.tcount + blockIdx.x, 1) == (int)gridDim.y - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    tmin = 1e300;
    if (active)
        for (int c = ty; c < a.tchunks; c += OPT_WY) tmin = fmin(tmin, __ldcg(a.tpart + (size_t)c * B + b));
    red[ty][tx] = tmin;
    __syncthreads();
    if (ty == 0 && active) {
        for (int y = 1; y < OPT_WY; y++) tmin = fmin(tmin, red[y][tx]);
        a.t[b] = fmin(a.t[b], a.boundary_frac * tmin);
    }
    if (tx == 0 && ty == 0) a.tcount[blockIdx.x] = 0;       // ready for the next launch
}

// trial point X = to_param(z + t D) for every running vector
__global__ void __launch_bounds__(32 * OPT_WY) opt_trial_kernel(OptArgs a) {
    const int b = blockIdx.x * 32 + threadIdx.x;
    const int B = a.o.B, P = a.o.P, nr = a.o.nr;
    if (b >= B) return;
    if (a.st[b] == ST_DONE) return;
    const double t = a.t[b];
    const int p0 = blockIdx.y * a.light_rows, p1 = min(P, p0 + a.light_rows);
    for (int p = p0 + threadIdx.y; p < p1; p += OPT_WY) {
        const size_t i = (size_t)p * B + b;
        const double sc = a.sc[p];
        const double zt = trial_point(a.Z[i], t, a.D[i], a.lo[p], a.hi[p]);
        a.X[i] = to_param(p < nr, zt, sc);
    }
}

// X = transform(Z) of the accepted point, F = F0
__global__ void __launch_bounds__(32 * OPT_WY) opt_final_kernel(OptArgs a, double* F) {
    const int b = blockIdx.x * 32 + threadIdx.x;
    if (b >= a.o.B) return;
    if (blockIdx.y == 0 && threadIdx.y == 0) F[b] = a.F0[b];
    const int p0 = blockIdx.y * a.light_rows, p1 = min(a.o.P, p0 + a.light_rows);
    for (int p = p0 + threadIdx.y; p < p1; p += OPT_WY) {
        const size_t i = (size_t)p * a.o.B + b;
        const double z = a.Z[i];
        a.X[i] = to_param(p < a.o.nr, z, a.sc[p]);
    }
}

}  // namespace tplopt
