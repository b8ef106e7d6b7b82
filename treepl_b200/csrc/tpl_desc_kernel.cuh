// PL objective + gradient, generation 4 (sm_100a): persistent ring kernel with per-tile operand
// descriptors and TMA-prefetched halo rows.
//
// Why (profiles/r01e_persist_ring_ncu.md): in generation 3 the consumer warps are latency bound.
// Each row walks a dependent chain  record -> parent record -> operand row  through shared memory,
// and the ~10 % of rows with a parent or child outside the tile read it from global memory (three
// dependent L2 round trips); those slow rows hold the ring stage that all warps share, so the
// producer starves (stall_long_sb 35 %, spin on the full barrier).
//
// Here the HOST resolves every operand of every node into an index local to the node's tile:
//   rows   0 .. 63            the tile's rate rows        (dense in X, one cp.async.bulk each)
//   rows  64 .. 64+hrows-1    the tile's free-date rows   (dense in X)
//   rows  .. +DESC_HALO-1     halo rows: rate / date rows of parents and children that live in
//                             other tiles, listed per tile and fetched by the producer like any other
//   last row                  zeros (the date of every tip)
// so a row needs ONE descriptor (48 B) and then all its operand loads are independent LDS.128s with
// 32-bit addresses — no chained lookups, no global loads, no in-tile/halo branches.  Rows that do not
// fit the scheme (root, children of the root, fixed calibrations, polytomies, halo overflow, operands
// outside the ordinary range) take general_row_global(), the exact reference-semantics path.
#pragma once
#include <cuda.h>      // CUtensorMap (types only; the encoder is fetched through cudaGetDriverEntryPoint)

#include "tpl_tile_kernel.cuh"

namespace tpl {

constexpr int DESC_HALO = 24;                 // halo rows per tile (p94 of the 100k-tip Yule tree; overflow -> general rows)
constexpr int DESC_INTS = 16;                 // 64-byte descriptor = 4 x int4
constexpr int DESC_DBOX = 8;                  // rows per TMA box of the date range (rate range: one 64-row box)
// consumer warps: 15 + the producer = 16 warps = 4 per SM sub-partition, whose 16,384 registers then allow 128 per thread
// (the kernel uses 118).  A 17th warp would put 5 on one sub-partition and cap every thread at 96 registers (spills).
constexpr int DESC_NCW = 15;

// descriptor words (int4 x 4):
//   q0 = { idx_lo, idx_hi, rslot, dslot }   idx bytes: r_own h_own r_par h_par | r_c1 h_c1 r_c2 h_c2
//        rslot == -2  -> general row ; dslot < 0 -> tip
//   q1 = { c, lf }  (double2)
//   q2 = { c_child1, c_child2 }  (double2)
//   q3 = { jc1, jc2, node_c1, node_c2 }   children's index inside the tile (-1: halo) and node numbers (CV masks)
struct DescTile {
    int rs0, nrs, ds0, nds, nh, pad0, pad1, pad2;
};

struct DescArgs {
    const int4* desc;          // [ntiles*64*4]
    const DescTile* tiles;     // [ntiles]
    const int* halo_rows;      // [ntiles*DESC_HALO] X row ids
    int hrows, ntiles, nchunks, ncta;
};

// hrows is rounded up to whole date boxes by the host
__host__ __device__ inline int desc_stage_rows(int hrows) { return TILE_NODES + hrows + DESC_HALO + 1; }
__host__ __device__ inline size_t desc_stage_bytes(int hrows) {
    return (size_t)desc_stage_rows(hrows) * TILE_LANES * 8 + TILE_NODES * 64 + TILE_NODES * 8;
}
// + the CTA-level reduction scratch of the objective partials: [NCW][64 vectors][ll, su] doubles and 64 "bad" flags
inline size_t desc_kernel_smem(int hrows, int stages) {
    return stages * desc_stage_bytes(hrows) + LOG_TABLE_N * 16 + 2 * 8 * 8 + 64 + (size_t)DESC_NCW * TILE_LANES * 16 + TILE_LANES * 4;
}

// One row entirely through global memory (L2): root, children of the root, fixed calibrations and their
// neighbours, polytomies, rows whose halo did not fit, and any row with an operand outside the ordinary
// range.  Reproduces every corner of the reference semantics; ~1 % of the rows.
// TMA 2-D tiled load: box (64 columns x R rows) of the tensor map at (col, row) -> shared memory
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int col, int row, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(col), "r"(row), "r"(smem_u32(bar))
        : "memory");
}

struct RowAcc {
    double ll0, ll1, su0, su1;
    int bad;
};
template <int MODE, bool LOGPEN, bool WANT_G, bool MASK>
__device__ __noinline__ RowAcc general_row_global(const EvalArgs& a, const TileArgs& ta, const double2* sTab, int i, int tx,
                                                  int lane, int chunk, double lam2a, double lam2b) {
    const unsigned B = (unsigned)a.B;
    const double* __restrict__ X = a.X;
    const int shl = 2 * tx;
    const double lam2[2] = {lam2a, lam2b};
    double acc_ll[2] = {0.0, 0.0}, acc_su[2] = {0.0, 0.0};
    int bad = 0;
    auto rate_of = [&](int node, const int4& rc, double (&out)[2]) {
        if (rc.y >= 0) {
            double2 v = ldg2(X + ((unsigned)rc.y * B + (unsigned)lane));
            out[0] = v.x; out[1] = v.y;
        } else {
            out[0] = out[1] = a.t.rfix[node];
        }
    };
    auto date_of = [&](int node, const int4& rc, double (&out)[2]) {
        if (rc.z >= 0) {
            double2 v = ldg2(X + ((unsigned)rc.z * B + (unsigned)lane));
            out[0] = v.x; out[1] = v.y;
        } else if (rc.z == -1) {
            out[0] = out[1] = 0.0;
        } else {
            out[0] = out[1] = a.t.hfix[node];
        }
    };
    auto mask_of = [&](int node) -> unsigned {
        if (!MASK) return 0u;
        unsigned long long mw = a.lanemask[(size_t)chunk * a.mask_stride + node];
        return (unsigned)(mw >> shl) & 3u;
    };
    const int4 rec = __ldg(ta.recs + i);
    const double2 cl = __ldg(ta.clf + i);
    const unsigned mi = mask_of(i);
    {
            double r[2], h[2];
            rate_of(i, rec, r);
            date_of(i, rec, h);
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
                                const int4 prec = __ldg(ta.recs + p);
                double hp[2];
                date_of(p, prec, hp);
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
                    rate_of(p, prec, rp);
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
                                        const int4 crec = __ldg(ta.recs + ch);
                    const unsigned mc = mask_of(ch);
                    if (MASK && mc == 3u) return;
                    double rc[2];
                    rate_of(ch, crec, rc);
                    if (need_d) {
                        const double cc = __ldg(ta.clf + ch).x;
                        double hc[2];
                        date_of(ch, crec, hc);
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
    RowAcc out;
    out.ll0 = acc_ll[0]; out.ll1 = acc_ll[1]; out.su0 = acc_su[0]; out.su1 = acc_su[1]; out.bad = bad;
    return out;
}


template <int MODE, bool LOGPEN, bool WANT_G, bool MASK, int S, int NCW>
__global__ void __launch_bounds__((NCW + 1) * 32, 1)
pl_desc_kernel(const __grid_constant__ EvalArgs a, const __grid_constant__ TileArgs ta, const __grid_constant__ DescArgs da, const __grid_constant__ CUtensorMap tmapR,
               const __grid_constant__ CUtensorMap tmapD) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int T = TILE_NODES, LW = TILE_LANES;
    const size_t stage_bytes = desc_stage_bytes(da.hrows);
    const int nrows = desc_stage_rows(da.hrows);
    double2* sTab = reinterpret_cast<double2*>(smem_raw + S * stage_bytes);        // [LOG_TABLE_N]
    uint64_t* full = reinterpret_cast<uint64_t*>(sTab + LOG_TABLE_N);              // [S]
    uint64_t* empty = full + S;                                                    // [S]
    double2* sRed = reinterpret_cast<double2*>(full + 16);                         // [NCW][LW] (ll, su) of one warp and vector
    int* sBad = reinterpret_cast<int*>(sRed + NCW * LW);                           // [LW]
    auto stage_rows = [&](int s) { return reinterpret_cast<double*>(smem_raw + s * stage_bytes); };
    auto stage_desc = [&](int s) { return reinterpret_cast<int4*>(stage_rows(s) + (size_t)nrows * LW); };
    auto stage_mask = [&](int s) { return reinterpret_cast<unsigned long long*>(stage_desc(s) + T * 4); };

    const int tid = threadIdx.x;
    const int w = tid >> 5, tx = tid & 31;
    const unsigned B = (unsigned)a.B;
    const long long W = (long long)da.nchunks * da.ntiles;
    const long long G = gridDim.x;
    const long long k_lo = persist_lo(blockIdx.x, W, G), k_hi = persist_lo(blockIdx.x + 1, W, G);
    const double* __restrict__ X = a.X;

    if (tid == 0) {
        for (int s = 0; s < S; s++) { mbar_init(full + s, 1); mbar_init(empty + s, NCW); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < LOG_TABLE_N) sTab[tid] = a.logtab[tid];
    if (tid < LW) sBad[tid] = 0;
    for (int s = 0; s < S; s++)                       // the zero row of every stage (never overwritten)
        if (tid < LW) stage_rows(s)[(size_t)(nrows - 1) * LW + tid] = 0.0;
    __syncthreads();

    if (w == NCW) {
        // ---------------- producer warp ----------------
        for (long long k = k_lo; k < k_hi; k++) {
            const int it = (int)(k - k_lo);
            const int s = it % S;
            const int chunk = (int)(k / da.ntiles), tile = (int)(k % da.ntiles);
            const int lane0 = chunk * LW;
            const unsigned rowbytes = (unsigned)min(LW, a.B - lane0) * 8u;
            // tile metadata first (two dependent L2 round trips), THEN wait for the stage: the loads overlap the wait
            const DescTile ti = da.tiles[tile];
            const int my_halo_row = (tx < ti.nh) ? da.halo_rows[tile * DESC_HALO + tx] : 0;
            if (it >= S) mbar_wait(empty + s, ((it / S) - 1) & 1);           // all consumer warps released the stage
            // Dense row ranges go through 2-D tensor maps of X ([P rows][B columns], fp64): ONE request for
            // the 64-row rate box and one per 8-row date box.  (Per-row 1-D copies made the single
            // producer warp the bottleneck: UBLKCP is a uniform-datapath instruction, so a per-lane loop
            // serialises into ~130 issue sequences per tile — profiles/r01f.)  A box always transfers
            // its full 64 x rows x 8 bytes; columns beyond B and rows beyond P arrive as zeros.
            const int ndbox = (ti.nds + DESC_DBOX - 1) / DESC_DBOX;
            if (tx == 0)
                mbar_expect_tx(full + s, (unsigned)(T * LW * 8) + (unsigned)ndbox * (DESC_DBOX * LW * 8) +
                                             (unsigned)ti.nh * rowbytes + T * 64u + (MASK ? T * 8u : 0u));
            __syncwarp();
            double* rows = stage_rows(s);
            if (tx == 0) tma_load_2d(rows, &tmapR, lane0, ti.rs0, full + s);
            if (tx >= 1 && tx <= ndbox)
                tma_load_2d(rows + (T + (tx - 1) * DESC_DBOX) * LW, &tmapD, lane0, ti.ds0 + (tx - 1) * DESC_DBOX, full + s);
            if (tx < ti.nh)
                bulk_g2s(rows + (T + da.hrows + tx) * LW, X + ((size_t)my_halo_row * B + lane0), rowbytes, full + s);
            if (tx == 31) bulk_g2s(stage_desc(s), da.desc + (size_t)tile * T * 4, T * 64u, full + s);
            if (MASK && tx == 30)
                bulk_g2s(stage_mask(s), a.lanemask + ((size_t)chunk * a.mask_stride + (size_t)tile * T), T * 8u, full + s);
        }
        return;
    }

    // ---------------- consumer warps ----------------
    double acc[4] = {0.0, 0.0, 0.0, 0.0};            // ll0 ll1 su0 su1
    int bad = 0;
    int cur_chunk = (k_lo < k_hi) ? (int)(k_lo / da.ntiles) : 0;
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
    // The objective partials of the CTA's consumer warps are summed HERE (fixed warp order) so that the finalize kernel
    // reads one row per CTA and slot instead of NCW rows: its chain of dependent partial loads was 9 % of the step.
    // Named barrier 1 = the consumer warps only (the producer warp has its own control flow).
    auto write_partials = [&](int slot_id) {
        sRed[w * LW + 2 * tx] = make_double2(acc[0], acc[2]);
        sRed[w * LW + 2 * tx + 1] = make_double2(acc[1], acc[3]);
        if (bad & 1) atomicOr(sBad + 2 * tx, 1);             // order independent
        if (bad & 2) atomicOr(sBad + 2 * tx + 1, 1);
        asm volatile("bar.sync 1, %0;" ::"r"(NCW * 32) : "memory");
        if (w < 2) {
            const int l = w * 32 + tx;
            double ll = 0.0, su = 0.0;
#pragma unroll
            for (int v = 0; v < NCW; v++) {
                const double2 t2 = sRed[v * LW + l];
                ll += t2.x;
                su += t2.y;
            }
            const size_t o = ((size_t)blockIdx.x * 2 + slot_id) * LW + l;
            a.part_ll[o] = ll;
            a.part_su[o] = su;
            a.part_bad[o] = sBad[l];
            sBad[l] = 0;
        }
        asm volatile("bar.sync 1, %0;" ::"r"(NCW * 32) : "memory");     // scratch may be rewritten (second slot)
    };
    constexpr bool FASTABLE = !(LOGPEN && MODE == 0);

    for (long long k = k_lo; k < k_hi; k++) {
        const int it = (int)(k - k_lo);
        const int s = it % S;
        const int chunk = (int)(k / da.ntiles), tile = (int)(k % da.ntiles);
        if (chunk != cur_chunk) {                       // at most once per CTA
            write_partials(slot);
            slot = 1;
            acc[0] = acc[1] = acc[2] = acc[3] = 0.0;
            bad = 0;
            cur_chunk = chunk;
            load_lambda(chunk);
        }
        mbar_wait(full + s, (it / S) & 1);
        const int lane0 = chunk * LW;
        const int nl = min(LW, a.B - lane0);
        if (2 * tx < nl) {
            const int lane = lane0 + 2 * tx;
            const double* rows = stage_rows(s) + 2 * tx;                 // this thread's two vectors
            const int4* sDesc = stage_desc(s);
            const unsigned long long* sMask = stage_mask(s);
            const int n0 = tile * T;
            const int tn = min(T, a.t.n - n0);
            // 64 rows over 15 warps leave 4 warps with a fifth row; the extra rows rotate from tile to tile so that no warp
            // is systematically the slowest (the ring releases a stage only when its slowest warp is done)
            const int rot = (it * (T % NCW)) % NCW;
            for (int j = (w + NCW - rot) % NCW; j < tn; j += NCW) {
                const int4 q0 = sDesc[4 * j];
                bool done = false;
                if (FASTABLE && q0.z != -2) {
                    const bool tip = q0.w < 0;
                    unsigned mi = 0u;
                    if (MASK) mi = (unsigned)(sMask[j] >> (2 * tx)) & 3u;
                    if (!MASK || tip || mi == 0u) {
                        const double2 cl = *reinterpret_cast<const double2*>(sDesc + 4 * j + 1);
                        const unsigned lo = (unsigned)q0.x, hi = (unsigned)q0.y;
                        const double2 r2 = *reinterpret_cast<const double2*>(rows + (lo & 0xffu) * LW);
                        const double2 h2 = *reinterpret_cast<const double2*>(rows + ((lo >> 8) & 0xffu) * LW);
                        const double2 rp2 = *reinterpret_cast<const double2*>(rows + ((lo >> 16) & 0xffu) * LW);
                        const double2 hp2 = *reinterpret_cast<const double2*>(rows + (lo >> 24) * LW);
                        const double r[2] = {r2.x, r2.y}, h[2] = {h2.x, h2.y};
                        const double hp[2] = {hp2.x, hp2.y}, rp[2] = {rp2.x, rp2.y};
                        bool ok = true;
                        double ll[2], g_r[2], g_h[2], q[2];
#pragma unroll
                        for (int l = 0; l < 2; l++) {
                            const double d = hp[l] - h[l];
                            const double xx = r[l] * d;
                            // r > 0, d > 0, both finite and r d ordinary <=> r's sign bit clear and r d in the ordinary range
                            // (one exponent test + one sign test instead of two exponent tests)
                            ok = ok && in_range400(xx) && (__double2hiint(r[l]) >= 0);
                            if (MODE == 0 && !tip) ok = ok && (__double2hiint(h[l]) >= 0);   // h < 0 -> LARGE (:74-78)
                            const double ac = cl.x * fast_rcp(xx);
                            g_r[l] = fma(-ac, d, d);                                         // d - c/r
                            g_h[l] = -fma(-ac, r[l], r[l]);                                  // -(r - c/d)
                            ll[l] = fma(-cl.x, fast_log(xx, sTab), xx) + cl.y;
                            q[l] = r[l] - rp[l];
                            if (WANT_G) g_r[l] = fma(lam2[l], q[l], g_r[l]);
                        }
                        if (WANT_G && !tip) {
                            const double2 cc = *reinterpret_cast<const double2*>(sDesc + 4 * j + 2);
                            const double2 rc1 = *reinterpret_cast<const double2*>(rows + (hi & 0xffu) * LW);
                            const double2 hc1 = *reinterpret_cast<const double2*>(rows + ((hi >> 8) & 0xffu) * LW);
                            const double2 rc2 = *reinterpret_cast<const double2*>(rows + ((hi >> 16) & 0xffu) * LW);
                            const double2 hc2 = *reinterpret_cast<const double2*>(rows + (hi >> 24) * LW);
                            unsigned mca[2] = {0u, 0u};
                            if (MASK) {
                                const int4 q3 = sDesc[4 * j + 3];
                                const unsigned long long m1 = q3.x >= 0 ? sMask[q3.x] : a.lanemask[(size_t)chunk * a.mask_stride + q3.z];
                                const unsigned long long m2 = q3.y >= 0 ? sMask[q3.y] : a.lanemask[(size_t)chunk * a.mask_stride + q3.w];
                                mca[0] = (unsigned)(m1 >> (2 * tx)) & 3u;
                                mca[1] = (unsigned)(m2 >> (2 * tx)) & 3u;
                            }
                            const double rca[2][2] = {{rc1.x, rc1.y}, {rc2.x, rc2.y}};
                            const double hca[2][2] = {{hc1.x, hc1.y}, {hc2.x, hc2.y}};
                            const double cca[2] = {cc.x, cc.y};
#pragma unroll
                            for (int kk = 0; kk < 2; kk++) {
#pragma unroll
                                for (int l = 0; l < 2; l++) {
                                    const double dc = h[l] - hca[kk][l];
                                    const double xc = rca[kk][l] * dc;
                                    ok = ok && in_range400(xc) && (__double2hiint(rca[kk][l]) >= 0);
                                    const double acn = cca[kk] * fast_rcp(xc);
                                    const double dlt = fma(-acn, rca[kk][l], rca[kk][l]);     // r_ch - c_ch/d_ch
                                    const double dq = rca[kk][l] - r[l];
                                    if (!MASK || !((mca[kk] >> l) & 1u)) {
                                        g_h[l] += dlt;
                                        g_r[l] = fma(-lam2[l], dq, g_r[l]);
                                    }
                                }
                            }
                        }
                        if (ok) {
#pragma unroll
                            for (int l = 0; l < 2; l++) {
                                if (!MASK || !((mi >> l) & 1u)) {
                                    acc[l] += ll[l];
                                    acc[2 + l] = fma(q[l], q[l], acc[2 + l]);
                                } else {
                                    g_r[l] = 0.0;
                                }
                            }
                            if (WANT_G) {
                                *reinterpret_cast<double2*>(a.G + ((unsigned)q0.z * B + (unsigned)lane)) = make_double2(g_r[0], g_r[1]);
                                if (!tip)
                                    *reinterpret_cast<double2*>(a.G + ((unsigned)q0.w * B + (unsigned)lane)) = make_double2(g_h[0], g_h[1]);
                            }
                            done = true;
                        }
                    }
                }
                if (!done) {
                    const RowAcc ra = general_row_global<MODE, LOGPEN, WANT_G, MASK>(a, ta, sTab, n0 + j, tx, lane, chunk,
                                                                                  lam2[0], lam2[1]);
                    acc[0] += ra.ll0; acc[1] += ra.ll1; acc[2] += ra.su0; acc[3] += ra.su1;
                    bad |= ra.bad;
                }
            }
        }
        __syncwarp();
        if (tx == 0) mbar_arrive(empty + s);             // this warp is done with the stage
    }
    if (k_lo < k_hi) write_partials(slot);
}

}  // namespace tpl
