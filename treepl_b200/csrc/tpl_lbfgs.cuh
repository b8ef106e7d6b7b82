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
//   opt_trial_kernel  : D = -free * sum_j delta_j b_j  (reads 2m+2 rows) for vectors with a new direction,
//                       then the trial point X = (exp(z + t D) | ds (z + t D)) for every running vector
// Deterministic: static row -> thread mapping, fixed-order reductions, no floating-point atomics.
#pragma once
#include "tpl_lbfgs_core.h"

namespace tplopt {

struct OptArgs {
    OptConst o;
    double *Z, *Gz, *D, *X;          // [P][B]
    const double* Gx;                // [P][B] gradient of the trial point in the ORIGINAL parameters (PL kernels)
    double *S, *Y;                   // [HM][P][B]
    const double *lo, *hi;           // [P] bounds in the transformed variables
    int *st, *ls, *head, *stall, *flag, *newdir, *iters, *nfail;
    unsigned* valid;
    double *F0, *t, *gd, *gamma, *gmax;
    const double* Ft;                // [B] objective of the trial point
    double *rho, *delta, *M, *part;  // [HM][B], [NB][B], [NB*NB][B], [nchunks][NACC][B]
    int nchunks, rows_per_chunk;
    int* ndone;
};

constexpr int OPT_WY = 8;            // row sub-lanes per block (block = 32 x 8)
constexpr int OPT_LX = 16;           // vectors per block of the lane kernel
constexpr int OPT_NY = 32;

// Z = clamp(transform^-1(X0)), D = 0
__global__ void __launch_bounds__(32 * OPT_WY) opt_init_kernel(OptArgs a) {
    const int b = blockIdx.x * 32 + threadIdx.x;
    if (b >= a.o.B) return;
    const int p0 = blockIdx.y * a.rows_per_chunk, p1 = min(a.o.P, p0 + a.rows_per_chunk);
    for (int p = p0 + threadIdx.y; p < p1; p += OPT_WY) {
        const size_t i = (size_t)p * a.o.B + b;
        const double x = a.X[i];
        const double z = p < a.o.nr ? log(x) : x / a.o.ds;
        a.Z[i] = fmin(fmax(z, a.lo[p]), a.hi[p]);
        a.D[i] = 0.0;
    }
}

template <int HM>
__global__ void __launch_bounds__(32 * OPT_WY) opt_update_kernel(OptArgs a) {
    typedef Layout<HM> L;
    constexpr int NACC = L::NACC;
    __shared__ double red[8][OPT_WY][32];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int b = blockIdx.x * 32 + tx;
    const int B = a.o.B, P = a.o.P, nr = a.o.nr;
    const double ds = a.o.ds;
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
    double acc[NACC];
#pragma unroll
    for (int k = 0; k < NACC; k++) acc[k] = 0.0;
    if (active) {
        const int p0 = blockIdx.y * a.rows_per_chunk, p1 = min(P, p0 + a.rows_per_chunk);
        for (int p = p0 + ty; p < p1; p += OPT_WY) {
            const size_t i = (size_t)p * B + b;
            const double lo = a.lo[p], hi = a.hi[p];
            const double x = a.X[i], gx = a.Gx[i];
            const double gzn = p < nr ? gx * x : gx * ds;          // chain rule: d/du = r d/dr, d/dz = ds d/dh
            double z = a.Z[i];
            double s = 0.0, y = 0.0;
            if (acc_lane) {
                const double zt = trial_point(z, t, a.D[i], lo, hi);
                s = zt - z;
                y = gzn - a.Gz[i];
                z = zt;
                a.Z[i] = zt;
                a.S[(size_t)h * PB + i] = s;
                a.Y[(size_t)h * PB + i] = y;
            }
            a.Gz[i] = gzn;
            const double gp = is_free(z, gzn, lo, hi) ? gzn : 0.0;
            acc[L::A_GG] = fma(gp, gp, acc[L::A_GG]);
            acc[L::A_GMAX] = fmax(acc[L::A_GMAX], fabs(gp));
            if (acc_lane) {
#pragma unroll
                for (int k = 0; k < HM; k++) {
                    const bool cur = k == h;
                    if (cur || ((valid >> k) & 1u)) {
                        const double Sk = cur ? s : a.S[(size_t)k * PB + i];
                        const double Yk = cur ? y : a.Y[(size_t)k * PB + i];
                        acc[6 * k + 0] = fma(s, Sk, acc[6 * k + 0]);
                        acc[6 * k + 1] = fma(s, Yk, acc[6 * k + 1]);
                        acc[6 * k + 2] = fma(y, Sk, acc[6 * k + 2]);
                        acc[6 * k + 3] = fma(y, Yk, acc[6 * k + 3]);
                        acc[6 * k + 4] = fma(gp, Sk, acc[6 * k + 4]);
                        acc[6 * k + 5] = fma(gp, Yk, acc[6 * k + 5]);
                    }
                }
            }
        }
    }
    // fixed-order block reduction over ty, 8 accumulators per round
    double* part = a.part + (size_t)blockIdx.y * NACC * B;
#pragma unroll
    for (int g0 = 0; g0 < NACC; g0 += 8) {
#pragma unroll
        for (int j = 0; j < 8; j++)
            if (g0 + j < NACC) red[j][ty][tx] = acc[g0 + j];
        __syncthreads();
        if (g0 + ty < NACC && b < B) {
            double v = red[ty][0][tx];
            if (g0 + ty == L::A_GMAX) {
                for (int y = 1; y < OPT_WY; y++) v = fmax(v, red[ty][y][tx]);
            } else {
                for (int y = 1; y < OPT_WY; y++) v += red[ty][y][tx];
            }
            part[(size_t)(g0 + ty) * B + b] = v;
        }
        __syncthreads();
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
                for (int c = 0; c < a.nchunks; c++) v = fmax(v, a.part[((size_t)c * NACC + k) * B + b]);
            } else {
                for (int c = 0; c < a.nchunks; c++) v += a.part[((size_t)c * NACC + k) * B + b];
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
        io.F0 = a.F0 + b; io.t = a.t + b; io.gd = a.gd + b; io.gamma = a.gamma + b; io.gmax = a.gmax + b;
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

template <int HM>
__global__ void __launch_bounds__(32 * OPT_WY) opt_trial_kernel(OptArgs a) {
    typedef Layout<HM> L;
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int b = blockIdx.x * 32 + tx;
    const int B = a.o.B, P = a.o.P, nr = a.o.nr;
    if (b >= B) return;
    if (a.st[b] == ST_DONE) return;
    const double ds = a.o.ds;
    const size_t PB = (size_t)P * B;
    const bool nd = a.newdir[b] != 0;
    const double t = a.t[b];
    const unsigned valid = a.valid[b];
    double dl[L::NB];
    if (nd) {
#pragma unroll
        for (int j = 0; j < L::NB; j++) dl[j] = a.delta[(size_t)j * B + b];
    }
    const int p0 = blockIdx.y * a.rows_per_chunk, p1 = min(P, p0 + a.rows_per_chunk);
    for (int p = p0 + ty; p < p1; p += OPT_WY) {
        const size_t i = (size_t)p * B + b;
        const double lo = a.lo[p], hi = a.hi[p];
        const double z = a.Z[i];
        double d;
        if (nd) {
            const double gz = a.Gz[i];
            double q = dl[L::IG] * gz;
#pragma unroll
            for (int k = 0; k < HM; k++)
                if ((valid >> k) & 1u) {
                    q = fma(dl[k], a.S[(size_t)k * PB + i], q);
                    q = fma(dl[HM + k], a.Y[(size_t)k * PB + i], q);
                }
            d = is_free(z, gz, lo, hi) ? -q : 0.0;
            a.D[i] = d;
        } else {
            d = a.D[i];
        }
        const double zt = trial_point(z, t, d, lo, hi);
        a.X[i] = p < nr ? exp(zt) : zt * ds;
    }
}

// X = transform(Z) of the accepted point, F = F0
__global__ void __launch_bounds__(32 * OPT_WY) opt_final_kernel(OptArgs a, double* F) {
    const int b = blockIdx.x * 32 + threadIdx.x;
    if (b >= a.o.B) return;
    if (blockIdx.y == 0 && threadIdx.y == 0) F[b] = a.F0[b];
    const int p0 = blockIdx.y * a.rows_per_chunk, p1 = min(a.o.P, p0 + a.rows_per_chunk);
    for (int p = p0 + threadIdx.y; p < p1; p += OPT_WY) {
        const size_t i = (size_t)p * a.o.B + b;
        const double z = a.Z[i];
        a.X[i] = p < a.o.nr ? exp(z) : z * a.o.ds;
    }
}

}  // namespace tplopt
