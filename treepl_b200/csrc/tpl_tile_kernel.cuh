// PL objective + gradient, generations 2 and 3 (sm_100a): shared-memory staged node tiles.
//
// Why (profiles/r01a_pull_kernel_ncu.md): the generation-1 kernel spends 304 warp instructions per
// 32-lane node row, ~42 of them fp64 — it is bound by address arithmetic, topology loads and
// memory latency, not by HBM (21 %) or the fp64 pipe (16 %).  These kernels remove that overhead:
//
//  * tile = 64 consecutive nodes (reference preorder) x 64 vectors.  In the reference parameter
//    layout (generate_param_order_vector, utils.cpp:295-340) the rate rows of consecutive nodes are
//    consecutive rows of X, and so are their free-date rows: a tile's operands are two dense row
//    ranges, fetched with one cp.async.bulk (TMA, 512 B) per row into shared memory against an
//    mbarrier, together with the tile's packed topology records.
//  * preorder locality: a node's parent is usually a few rows above it and its first child is the
//    next row, so ~90 % of the rows find parent and children inside the tile (host-classified
//    "fast" rows: straight-line code, LDS.128 operands, 32-bit addressing).  The other rows take a
//    general path that reads halo operands from global memory (L2: the neighbouring tile has just
//    pulled them) and reproduces every corner of the reference semantics.
//  * each thread owns TWO vectors (one LDS.128 / STG.128 per operand row), so topology decoding,
//    branches and address arithmetic are amortised over 64 vectors per warp row.
//  * gradient rows are written exactly once by their owner straight from registers (coalesced
//    512 B per warp); objective partials are reduced in a fixed order: bitwise deterministic.
//
// Generation 2 (pl_tile_kernel): one CTA per tile, load -> barrier -> compute -> barrier.
//   profiles/r01d: 46 % of its stall samples are barrier waits (load phase + warp imbalance).
// Generation 3 (pl_persist_kernel): one persistent CTA per SM, 24 consumer warps + 1 producer
//   warp, 3-stage ring of tiles with full/empty mbarriers.  Consumers never wait for each other:
//   a warp that finishes its rows of tile k moves on to tile k+1 while the producer refills the
//   stage that all warps have released.  Objective accumulators stay in registers across tiles.
#pragma once
#include "tpl_kernels.cuh"

namespace tpl {

constexpr int TILE_NODES = 64;
constexpr int TILE_LANES = 64;

// packed topology record (one LDS.128 / LDG.128)
//   x parent
//   y rslot : >= 0 row of X ; -1 rate = rfix[node]
//   z dslot : >= 0 row of X ; -1 date is exactly 0 (tips) ; -2 date = hfix[node]
//   w kind  : >= 0 FAST binary node, children (node+1, w), parent and children in the same tile
//             -1   FAST tip, parent in the same tile
//             -2   general row (children through the CSR list)
struct TileInfo {
    int rs0, nrs, ds0, nds;
};

struct TileArgs {
    const int4* recs;          // [ntiles*64] (padded)
    const double2* clf;        // [ntiles*64] (char_duration, log_fact)
    const TileInfo* tiles;     // [ntiles]
    int hrows;                 // max date rows of any tile (shared-memory sizing)
    int ntiles;
    int nchunks;
    int ncta;                  // generation 3: persistent CTAs
};

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// TMA 1-D bulk copy global -> shared, completion counted in bytes on the mbarrier
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ double2 ldg2(const double* p) { return __ldg((const double2*)p); }

// One staged tile as the row code sees it.
struct TileView {
    const double* sR;                 // [64][64] rate rows
    const double* sH;                 // [hrows][64] date rows
    const int4* sRec;                 // [64]
    const double2* sCL;               // [64]
    const unsigned long long* sMask;  // [64]
    int n0, tn;
    TileInfo ti;
};

// Rows row0, row0+row_step, ... of the tile, for this thread's two vectors (lane, lane+1).
template <int MODE, bool LOGPEN, bool WANT_G, bool MASK>
__device__ __forceinline__ void tile_rows(const EvalArgs& a, const TileArgs& ta, const TileView& v,
                                          const double2* sTab, const double* sZero, int tx, int lane, int chunk,
                                          const double (&lam2)[2], int row0, int row_step,
                                          double (&acc_ll)[2], double (&acc_su)[2], int& bad) {
    constexpr int LW = TILE_LANES;
    const double* sR = v.sR;
    const double* sH = v.sH;
    const int4* sRec = v.sRec;
    const double2* sCL = v.sCL;
    const unsigned long long* sMask = v.sMask;
    const int n0 = v.n0, tn = v.tn, n1 = v.n0 + v.tn;
    const TileInfo ti = v.ti;
    const unsigned B = (unsigned)a.B;
    const double* __restrict__ X = a.X;
    const int shl = 2 * tx;
    {
        // operand fetch: in-tile rows from shared memory, halo rows from global (L2)
        auto rate_of = [&](int node, const int4& rc, bool intile, double (&out)[2]) {
            if (rc.y >= 0) {
                double2 v = intile ? *reinterpret_cast<const double2*>(sR + (rc.y - ti.rs0) * LW + 2 * tx)
                                   : ldg2(X + ((unsigned)rc.y * B + (unsigned)lane));
                out[0] = v.x; out[1] = v.y;
            } else {
                out[0] = out[1] = a.t.rfix[node];
            }
        };
        auto date_of = [&](int node, const int4& rc, bool intile, double (&out)[2]) {
            if (rc.z >= 0) {
                double2 v = intile ? *reinterpret_cast<const double2*>(sH + (rc.z - ti.ds0) * LW + 2 * tx)
                                   : ldg2(X + ((unsigned)rc.z * B + (unsigned)lane));
                out[0] = v.x; out[1] = v.y;
            } else if (rc.z == -1) {
                out[0] = out[1] = 0.0;
            } else {
                out[0] = out[1] = a.t.hfix[node];
            }
        };
        auto mask_of = [&](int node, bool intile) -> unsigned {
            if (!MASK) return 0u;
            unsigned long long mw = intile ? sMask[node - n0] : a.lanemask[(size_t)chunk * a.mask_stride + node];
            return (unsigned)(mw >> shl) & 3u;
        };

        // ---- fast path: a tip or a binary internal node whose parent and children are ordinary free
        // parameters AND lie in this tile (host-classified: rec.w != -2; ~90 % of the rows), and whose
        // operands are all in the ordinary range.  Straight-line code, shared-memory operands only
        // (32-bit addressing), tips' dates from a shared zero row, CV masks as predicates.  Nothing is
        // committed unless every operand passed the range test; otherwise the general path runs.
        constexpr bool FASTABLE = !(LOGPEN && MODE == 0);
        auto fast_row = [&](int j, const int4& rec, const double2& cl, unsigned mi) -> bool {
            const bool tip = rec.w == -1;
            const double2 r2 = *reinterpret_cast<const double2*>(sR + (rec.y - ti.rs0) * LW + 2 * tx);
            const double* ph = tip ? (sZero + 2 * tx) : (sH + (rec.z - ti.ds0) * LW + 2 * tx);
            const double2 h2 = *reinterpret_cast<const double2*>(ph);
            const int4 prec = sRec[rec.x - n0];
            const double2 hp2 = *reinterpret_cast<const double2*>(sH + (prec.z - ti.ds0) * LW + 2 * tx);
            const double2 rp2 = *reinterpret_cast<const double2*>(sR + (prec.y - ti.rs0) * LW + 2 * tx);
            const double r[2] = {r2.x, r2.y}, h[2] = {h2.x, h2.y};
            const double hp[2] = {hp2.x, hp2.y}, rp[2] = {rp2.x, rp2.y};
            bool ok = true;
            double ll[2], g_r[2], g_h[2], q[2];
#pragma unroll
            for (int l = 0; l < 2; l++) {
                const double d = hp[l] - h[l];
                ok = ok && in_range400(r[l]) && in_range400(d);
                if (MODE == 0 && !tip) ok = ok && (__double2hiint(h[l]) >= 0);       // h < 0 -> LARGE (:74-78)
                const double xx = r[l] * d;
                const double ac = cl.x * fast_rcp(xx);
                g_r[l] = fma(-ac, d, d);                                             // d - c/r
                g_h[l] = -fma(-ac, r[l], r[l]);                                      // -(r - c/d)
                ll[l] = fma(-cl.x, fast_log(xx, sTab), xx) + cl.y;
                q[l] = r[l] - rp[l];
                if (WANT_G) g_r[l] = fma(lam2[l], q[l], g_r[l]);
            }
            if (WANT_G && !tip) {
                const int j2 = rec.w - n0;
                const int4 cr1 = sRec[j + 1];
                const int4 cr2 = sRec[j2];
                const double cca[2] = {sCL[j + 1].x, sCL[j2].x};
                const double2 rc1 = *reinterpret_cast<const double2*>(sR + (cr1.y - ti.rs0) * LW + 2 * tx);
                const double2 rc2 = *reinterpret_cast<const double2*>(sR + (cr2.y - ti.rs0) * LW + 2 * tx);
                const double* ph1 = cr1.z < 0 ? (sZero + 2 * tx) : (sH + (cr1.z - ti.ds0) * LW + 2 * tx);
                const double* ph2 = cr2.z < 0 ? (sZero + 2 * tx) : (sH + (cr2.z - ti.ds0) * LW + 2 * tx);
                const double2 hc1 = *reinterpret_cast<const double2*>(ph1);
                const double2 hc2 = *reinterpret_cast<const double2*>(ph2);
                unsigned mca[2] = {0u, 0u};
                if (MASK) { mca[0] = (unsigned)(sMask[j + 1] >> shl) & 3u; mca[1] = (unsigned)(sMask[j2] >> shl) & 3u; }
                const double rca[2][2] = {{rc1.x, rc1.y}, {rc2.x, rc2.y}};
                const double hca[2][2] = {{hc1.x, hc1.y}, {hc2.x, hc2.y}};
#pragma unroll
                for (int k = 0; k < 2; k++) {
#pragma unroll
                    for (int l = 0; l < 2; l++) {
                        const double dc = h[l] - hca[k][l];
                        ok = ok && in_range400(rca[k][l]) && in_range400(dc);
                        const double acc = cca[k] * fast_rcp(rca[k][l] * dc);
                        const double dlt = fma(-acc, rca[k][l], rca[k][l]);         // r_ch - c_ch/d_ch
                        const double dq = rca[k][l] - r[l];
                        if (!MASK || !((mca[k] >> l) & 1u)) {
                            g_h[l] += dlt;
                            g_r[l] = fma(-lam2[l], dq, g_r[l]);
                        }
                    }
                }
            }
            if (!ok) return false;
#pragma unroll
            for (int l = 0; l < 2; l++) {
                if (!MASK || !((mi >> l) & 1u)) {
                    acc_ll[l] += ll[l];
                    acc_su[l] = fma(q[l], q[l], acc_su[l]);
                } else {
                    g_r[l] = 0.0;
                }
            }
            if (WANT_G) {
                *reinterpret_cast<double2*>(a.G + ((unsigned)rec.y * B + (unsigned)lane)) = make_double2(g_r[0], g_r[1]);
                if (!tip) *reinterpret_cast<double2*>(a.G + ((unsigned)rec.z * B + (unsigned)lane)) = make_double2(g_h[0], g_h[1]);
            }
            return true;
        };

        // static row assignment (warp w takes rows w, w+8, ...): which warp sums which rows must not
        // depend on timing, or the objective would not be bitwise reproducible
        for (int j = row0; j < tn; j += row_step) {
            const int i = n0 + j;
            const int4 rec = sRec[j];
            const double2 cl = sCL[j];
            const unsigned mi = mask_of(i, true);
            if (FASTABLE && rec.w != -2 && (rec.w == -1 || mi == 0u)) {
                if (fast_row(j, rec, cl, mi)) continue;
            }
            double r[2], h[2];
            rate_of(i, rec, true, r);
            date_of(i, rec, true, h);
            if (MODE == 0) {                             // calc_pl :74-78
#pragma unroll
                for (int l = 0; l < 2; l++) {
                    if (rec.y >= 0 && !((mi >> l) & 1u) && r[l] < 0.0) bad |= 1 << l;
                    if (rec.z >= 0 && h[l] < 0.0) bad |= 1 << l;
                }
            }
            double g_r[2] = {0.0, 0.0}, g_h[2] = {0.0, 0.0}, lr[2] = {0.0, 0.0};
            if (i != 0) {
                const int p = rec.x;
                const bool pin = p >= n0;
                const int4 prec = pin ? sRec[p - n0] : __ldg(ta.recs + p);
                double hp[2];
                date_of(p, prec, pin, hp);
                double d[2];
                bool dconst[2];
#pragma unroll
                for (int l = 0; l < 2; l++) {
                    int b1 = 0;
                    d[l] = duration_rule<MODE, false>(hp[l], h[l], b1, dconst[l]);
                    bad |= b1 << l;
                }
                double rp[2] = {0.0, 0.0};
                if (p != 0) {
                    rate_of(p, prec, pin, rp);
                }
#pragma unroll
                for (int l = 0; l < 2; l++) {
                    const bool ml = (mi >> l) & 1u;
                    if (!ml || (MODE == 0 && WANT_G && rec.z >= 0)) {
                        Branch b = branch_terms<MODE, false, true>(r[l], d[l], cl.x, cl.y, dconst[l], sTab);
                        g_h[l] = -b.delta;               // analytic: own term regardless of cv (:728-735)
                        if (!ml) { acc_ll[l] += b.ll; g_r[l] = b.dr; }
                        else if (MODE == 1) g_h[l] = 0.0;
                    }
                    if (!ml) {
                        if (p != 0) {
                            const double q = r[l] - rp[l];
                            acc_su[l] = fma(q, q, acc_su[l]);                                // :862-864
                            if (WANT_G) {
                                if (LOGPEN && MODE == 0) {
                                    lr[l] = log_any(r[l], sTab);
                                    g_r[l] += lam2[l] * (lr[l] - log_any(rp[l], sTab)) / r[l];   // :795-796
                                } else {
                                    g_r[l] = fma(lam2[l], q, g_r[l]);                          // :798
                                }
                            }
                        } else if (WANT_G && LOGPEN && MODE == 0) {
                            lr[l] = log_any(r[l], sTab);
                        }
                    }
                }
            }
            if (WANT_G) {
                double sum_delta[2] = {0.0, 0.0}, sum_q[2] = {0.0, 0.0};
                const bool need_q = (rec.y >= 0) && (i != 0);
                const bool need_d = rec.z >= 0;
                auto child = [&](int ch) {
                    const bool cin = ch < n1;
                    const int4 crec = cin ? sRec[ch - n0] : __ldg(ta.recs + ch);
                    const unsigned mc = mask_of(ch, cin);
                    if (MASK && mc == 3u) return;
                    double rc[2];
                    rate_of(ch, crec, cin, rc);
                    if (need_d) {
                        const double cc = cin ? sCL[ch - n0].x : __ldg(ta.clf + ch).x;
                        double hc[2];
                        date_of(ch, crec, cin, hc);
#pragma unroll
                        for (int l = 0; l < 2; l++) {
                            if (MASK && ((mc >> l) & 1u)) continue;                          // :738
                            bool dc_const;
                            int dummy = 0;
                            const double dc = duration_rule<MODE, false>(h[l], hc[l], dummy, dc_const);
                            Branch b = branch_terms<MODE, false, false>(rc[l], dc, cc, 0.0, dc_const, sTab);
                            sum_delta[l] += b.delta;
                        }
                    }
                    if (need_q) {
#pragma unroll
                        for (int l = 0; l < 2; l++) {
                            if (MASK && ((mc >> l) & 1u)) continue;                          // :784
                            if (LOGPEN && MODE == 0) sum_q[l] += (log_any(rc[l], sTab) - lr[l]) / r[l];   // :786, :802
                            else sum_q[l] += rc[l] - r[l];
                        }
                    }
                };
                if (need_q || need_d) {
                    if (rec.w >= 0) {
                        child(i + 1);
                        child(rec.w);
                    } else if (rec.w == -2) {
                        for (int k = a.t.child_off[i]; k < a.t.child_off[i + 1]; k++) child(a.t.children[k]);
                    }
                }
                if (rec.y >= 0) {
                    double2 o;
                    o.x = ((mi & 1u) ? 0.0 : fma(-lam2[0], sum_q[0], g_r[0]));
                    o.y = ((mi & 2u) ? 0.0 : fma(-lam2[1], sum_q[1], g_r[1]));
                    *reinterpret_cast<double2*>(a.G + ((unsigned)rec.y * B + (unsigned)lane)) = o;
                }
                if (rec.z >= 0) {
                    double2 o;
                    o.x = g_h[0] + sum_delta[0];
                    o.y = g_h[1] + sum_delta[1];
                    *reinterpret_cast<double2*>(a.G + ((unsigned)rec.z * B + (unsigned)lane)) = o;
                }
            }
        }
    }
}

inline size_t tile_kernel_smem(int hrows) {
    return (size_t)TILE_NODES * TILE_LANES * 8 + (size_t)hrows * TILE_LANES * 8 + TILE_NODES * (16 + 16 + 8) +
           LOG_TABLE_N * 16 + TILE_LANES * 8 + 32;
}

// ---------------------------------------------------------------------------------------------
// generation 2: one CTA (8 warps) per (tile, 64-vector chunk)
// ---------------------------------------------------------------------------------------------
template <int MODE, bool LOGPEN, bool WANT_G, bool MASK>
__global__ void __launch_bounds__(256, 3) pl_tile_kernel(const __grid_constant__ EvalArgs a, const __grid_constant__ TileArgs ta) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int T = TILE_NODES, LW = TILE_LANES;
    double* sR = reinterpret_cast<double*>(smem_raw);                      // [T][LW]
    double* sH = sR + T * LW;                                              // [hrows][LW]
    int4* sRec = reinterpret_cast<int4*>(sH + (size_t)ta.hrows * LW);      // [T]
    double2* sCL = reinterpret_cast<double2*>(sRec + T);                   // [T]
    unsigned long long* sMask = reinterpret_cast<unsigned long long*>(sCL + T);   // [T]
    double2* sTab = reinterpret_cast<double2*>(sMask + T);                 // [LOG_TABLE_N]
    double* sZero = reinterpret_cast<double*>(sTab + LOG_TABLE_N);         // [LW] zeros: the date row of every tip
    uint64_t* mbar = reinterpret_cast<uint64_t*>(sZero + LW);

    const int tid = threadIdx.x;
    const int w = tid >> 5, tx = tid & 31;
    const int tile = blockIdx.x, chunk = blockIdx.y;
    const unsigned B = (unsigned)a.B;
    const int lane0 = chunk * LW;
    const int nl = min(LW, a.B - lane0);                 // even
    const unsigned rowbytes = (unsigned)nl * 8u;
    const int n0 = tile * T;
    const int tn = min(T, a.t.n - n0);
    const TileInfo ti = ta.tiles[tile];
    const double* __restrict__ X = a.X;

    if (tid == 0) {
        mbar_init(mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (w == 0) {
        if (tx == 0) mbar_expect_tx(mbar, (unsigned)(ti.nrs + ti.nds) * rowbytes);
        __syncwarp();
        for (int k = tx; k < ti.nrs; k += 32)
            bulk_g2s(sR + k * LW, X + ((size_t)(ti.rs0 + k) * B + lane0), rowbytes, mbar);
        for (int k = tx; k < ti.nds; k += 32)
            bulk_g2s(sH + k * LW, X + ((size_t)(ti.ds0 + k) * B + lane0), rowbytes, mbar);
    }
    for (int j = tid; j < tn; j += 256) {
        sRec[j] = ta.recs[n0 + j];
        sCL[j] = ta.clf[n0 + j];
        if (MASK) sMask[j] = a.lanemask[(size_t)chunk * a.mask_stride + n0 + j];
    }
    sTab[tid] = a.logtab[tid];
    if (tid < LW) sZero[tid] = 0.0;
    __syncthreads();
    mbar_wait(mbar, 0);

    const int lane = lane0 + 2 * tx;                     // first of this thread's two vectors
    const bool active = 2 * tx < nl;
    double acc_ll[2] = {0.0, 0.0}, acc_su[2] = {0.0, 0.0};
    int bad = 0;                                         // bit l = vector l infeasible
    if (active) {
        double lam[2];
        if (a.lane_lambda) { double2 t2 = ldg2(a.lane_lambda + lane); lam[0] = t2.x; lam[1] = t2.y; }
        else { lam[0] = lam[1] = a.lambda; }
        const double lam2[2] = {2.0 * lam[0], 2.0 * lam[1]};
        TileView v;
        v.sR = sR; v.sH = sH; v.sRec = sRec; v.sCL = sCL; v.sMask = sMask; v.n0 = n0; v.tn = tn; v.ti = ti;
        // static row assignment (warp w takes rows w, w+8, ...): which warp sums which rows must not
        // depend on timing, or the objective would not be bitwise reproducible
        tile_rows<MODE, LOGPEN, WANT_G, MASK>(a, ta, v, sTab, sZero, tx, lane, chunk, lam2, w, 8, acc_ll, acc_su, bad);
    }
    // fixed-order reduction over the 8 warps; the operand tile is dead, reuse it as scratch
    __syncthreads();
    double* sRed = sR;                                   // [8][2][LW]
    int* sBad = reinterpret_cast<int*>(sR + 8 * 2 * LW); // [8][32]
    sRed[(w * 2 + 0) * LW + 2 * tx] = acc_ll[0];
    sRed[(w * 2 + 0) * LW + 2 * tx + 1] = acc_ll[1];
    sRed[(w * 2 + 1) * LW + 2 * tx] = acc_su[0];
    sRed[(w * 2 + 1) * LW + 2 * tx + 1] = acc_su[1];
    sBad[w * 32 + tx] = bad;
    __syncthreads();
    if (w == 0 && active) {
        double s0[2] = {0.0, 0.0}, s1[2] = {0.0, 0.0};
        int bb = 0;
        for (int y = 0; y < 8; y++) {
            s0[0] += sRed[(y * 2 + 0) * LW + 2 * tx];
            s0[1] += sRed[(y * 2 + 0) * LW + 2 * tx + 1];
            s1[0] += sRed[(y * 2 + 1) * LW + 2 * tx];
            s1[1] += sRed[(y * 2 + 1) * LW + 2 * tx + 1];
            bb |= sBad[y * 32 + tx];
        }
        const size_t o = (size_t)tile * B + lane;
        *reinterpret_cast<double2*>(a.part_ll + o) = make_double2(s0[0], s0[1]);
        *reinterpret_cast<double2*>(a.part_su + o) = make_double2(s1[0], s1[1]);
        a.part_bad[o] = bb & 1;
        a.part_bad[o + 1] = (bb >> 1) & 1;
    }
}

// ---------------------------------------------------------------------------------------------
// generation 3: persistent, warp-specialised, 3-stage ring
// ---------------------------------------------------------------------------------------------
constexpr int P_CONSUMER_WARPS = 20;
constexpr int P_THREADS = (P_CONSUMER_WARPS + 1) * 32;      // + 1 producer warp
constexpr int P_STAGES = 3;

__host__ __device__ inline size_t persist_stage_bytes(int hrows) {
    return (size_t)TILE_NODES * TILE_LANES * 8 + (size_t)hrows * TILE_LANES * 8 + TILE_NODES * (16 + 16 + 8);
}
inline size_t persist_kernel_smem(int hrows) {
    return P_STAGES * persist_stage_bytes(hrows) + LOG_TABLE_N * 16 + TILE_LANES * 8 + 2 * P_STAGES * 8 + 64;
}
// work item k (0 <= k < nchunks*ntiles) = (chunk k / ntiles, tile k % ntiles); CTA c owns [lo(c), lo(c+1))
__host__ __device__ inline long long persist_lo(long long c, long long W, long long G) { return c * W / G; }

template <int MODE, bool LOGPEN, bool WANT_G, bool MASK>
__global__ void __launch_bounds__(P_THREADS, 1) pl_persist_kernel(const __grid_constant__ EvalArgs a, const __grid_constant__ TileArgs ta) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int T = TILE_NODES, LW = TILE_LANES, S = P_STAGES, NCW = P_CONSUMER_WARPS;
    const size_t stage_bytes = persist_stage_bytes(ta.hrows);
    double2* sTab = reinterpret_cast<double2*>(smem_raw + S * stage_bytes);        // [LOG_TABLE_N]
    double* sZero = reinterpret_cast<double*>(sTab + LOG_TABLE_N);                 // [LW]
    uint64_t* full = reinterpret_cast<uint64_t*>(sZero + LW);                      // [S]
    uint64_t* empty = full + S;                                                    // [S]
    auto stage_R = [&](int s) { return reinterpret_cast<double*>(smem_raw + s * stage_bytes); };
    auto stage_H = [&](int s) { return stage_R(s) + T * LW; };
    auto stage_Rec = [&](int s) { return reinterpret_cast<int4*>(stage_H(s) + (size_t)ta.hrows * LW); };
    auto stage_CL = [&](int s) { return reinterpret_cast<double2*>(stage_Rec(s) + T); };
    auto stage_Mask = [&](int s) { return reinterpret_cast<unsigned long long*>(stage_CL(s) + T); };

    const int tid = threadIdx.x;
    const int w = tid >> 5, tx = tid & 31;
    const unsigned B = (unsigned)a.B;
    const long long W = (long long)ta.nchunks * ta.ntiles;
    const long long G = gridDim.x;
    const long long k_lo = persist_lo(blockIdx.x, W, G), k_hi = persist_lo(blockIdx.x + 1, W, G);
    const double* __restrict__ X = a.X;

    if (tid == 0) {
        for (int s = 0; s < S; s++) { mbar_init(full + s, 1); mbar_init(empty + s, NCW); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < LOG_TABLE_N) sTab[tid] = a.logtab[tid];
    if (tid < LW) sZero[tid] = 0.0;
    __syncthreads();

    if (w == NCW) {
        // ---------------- producer warp ----------------
        for (long long k = k_lo; k < k_hi; k++) {
            const int it = (int)(k - k_lo);
            const int s = it % S;
            if (it >= S) mbar_wait(empty + s, ((it / S) - 1) & 1);           // all consumer warps released the stage
            const int chunk = (int)(k / ta.ntiles), tile = (int)(k % ta.ntiles);
            const int lane0 = chunk * LW;
            const unsigned rowbytes = (unsigned)min(LW, a.B - lane0) * 8u;
            const TileInfo ti = ta.tiles[tile];
            const int n0 = tile * T;
            if (tx == 0)
                mbar_expect_tx(full + s, (unsigned)(ti.nrs + ti.nds) * rowbytes + T * 32u + (MASK ? T * 8u : 0u));
            __syncwarp();
            double* dR = stage_R(s);
            double* dH = stage_H(s);
            for (int r = tx; r < ti.nrs; r += 32)
                bulk_g2s(dR + r * LW, X + ((size_t)(ti.rs0 + r) * B + lane0), rowbytes, full + s);
            for (int r = tx; r < ti.nds; r += 32)
                bulk_g2s(dH + r * LW, X + ((size_t)(ti.ds0 + r) * B + lane0), rowbytes, full + s);
            if (tx == 0) bulk_g2s(stage_Rec(s), ta.recs + n0, T * 16u, full + s);
            if (tx == 1) bulk_g2s(stage_CL(s), ta.clf + n0, T * 16u, full + s);
            if (MASK && tx == 2) bulk_g2s(stage_Mask(s), a.lanemask + ((size_t)chunk * a.mask_stride + n0), T * 8u, full + s);
        }
        return;
    }

    // ---------------- consumer warps ----------------
    double acc_ll[2] = {0.0, 0.0}, acc_su[2] = {0.0, 0.0};
    int bad = 0;
    int cur_chunk = (k_lo < k_hi) ? (int)(k_lo / ta.ntiles) : 0;
    int slot = 0;
    double lam2[2] = {0.0, 0.0};
    auto load_lambda = [&](int chunk) {
        const int lane = chunk * LW + 2 * tx;
        if (2 * tx < min(LW, a.B - chunk * LW)) {
            if (a.lane_lambda) { double2 t2 = ldg2(a.lane_lambda + lane); lam2[0] = 2.0 * t2.x; lam2[1] = 2.0 * t2.y; }
            else { lam2[0] = lam2[1] = 2.0 * a.lambda; }
        }
    };
    load_lambda(cur_chunk);

    // per-warp partials, entry ((cta*2 + slot)*NCW + warp), 64 vectors each; slot 0 = first chunk this
    // CTA touches.  No CTA-level reduction (it would need a barrier and scratch): the finalize kernel
    // sums the entries in fixed (cta, warp) order.
    auto write_partials = [&](int slot_id) {
        const size_t o = (((size_t)blockIdx.x * 2 + slot_id) * NCW + w) * LW + 2 * tx;
        *reinterpret_cast<double2*>(a.part_ll + o) = make_double2(acc_ll[0], acc_ll[1]);
        *reinterpret_cast<double2*>(a.part_su + o) = make_double2(acc_su[0], acc_su[1]);
        a.part_bad[o] = bad & 1;
        a.part_bad[o + 1] = (bad >> 1) & 1;
    };

    for (long long k = k_lo; k < k_hi; k++) {
        const int it = (int)(k - k_lo);
        const int s = it % S;
        const int chunk = (int)(k / ta.ntiles), tile = (int)(k % ta.ntiles);
        if (chunk != cur_chunk) {                       // at most once per CTA
            write_partials(slot);
            slot = 1;
            acc_ll[0] = acc_ll[1] = acc_su[0] = acc_su[1] = 0.0;
            bad = 0;
            cur_chunk = chunk;
            load_lambda(chunk);
        }
        mbar_wait(full + s, (it / S) & 1);
        const int lane0 = chunk * LW;
        const int nl = min(LW, a.B - lane0);
        if (2 * tx < nl) {
            TileView v;
            v.sR = stage_R(s); v.sH = stage_H(s); v.sRec = stage_Rec(s); v.sCL = stage_CL(s); v.sMask = stage_Mask(s);
            v.n0 = tile * T; v.tn = min(T, a.t.n - tile * T); v.ti = ta.tiles[tile];
            tile_rows<MODE, LOGPEN, WANT_G, MASK>(a, ta, v, sTab, sZero, tx, lane0 + 2 * tx, chunk, lam2, w, NCW,
                                                   acc_ll, acc_su, bad);
        }
        __syncwarp();
        if (tx == 0) mbar_arrive(empty + s);             // this warp is done with the stage
    }
    if (k_lo < k_hi) write_partials(slot);
}

}  // namespace tpl
