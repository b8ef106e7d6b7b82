// C ABI of the B200 PL engine (include/treepl_b200.h): handle, device-resident tree,
// slot maps, staging, launches.  Host-side bookkeeping mirrors pl_calc_parallel's members
// (reference src/pl_calc_parallel.h:17-82); all arithmetic of the hot path runs in the
// sm_100a kernels of tpl_kernels.cuh.  There is no CPU evaluation path in this file.
#include "../../include/treepl_b200.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

#include "tpl_kernels.cuh"
#include "tpl_tile_kernel.cuh"
#include "tpl_desc_kernel.cuh"
#include "tpl_lbfgs.cuh"
#include "tpl_tn.cuh"
#include "tpl_anneal.cuh"

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t _e = (call);                                                                   \
        if (_e != cudaSuccess)                                                                     \
            return fail(TPL_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(_e));         \
    } while (0)

// Device memory blocks outlive the engine that allocated them: a CV run or a multi-start job builds one engine per phase and
// per call, each with tens of GB of work space, and cudaMalloc / cudaFree of blocks that size cost 0.1-0.6 s per engine
// (more when several ranks of a box allocate at once: the driver serialises them).  Freed blocks are kept per device and
// handed to the next request of a similar size; tpl_release_cached_memory() returns them to the driver, and so does an
// allocation failure before it is reported.
struct DevPool {
    struct Blk { void* p; size_t bytes; int dev; };
    std::mutex m;
    std::vector<Blk> blocks;
    void* take(size_t bytes, int dev, size_t* got) {
        std::lock_guard<std::mutex> g(m);
        int best = -1;
        for (int k = 0; k < (int)blocks.size(); k++) {
            const Blk& b = blocks[k];
            if (b.dev != dev || b.bytes < bytes || b.bytes > std::max(2 * bytes, bytes + (size_t)(1 << 20))) continue;
            if (best < 0 || b.bytes < blocks[best].bytes) best = k;
        }
        if (best < 0) return nullptr;
        void* p = blocks[best].p;
        *got = blocks[best].bytes;
        blocks.erase(blocks.begin() + best);
        return p;
    }
    void give(void* p, size_t bytes, int dev) {
        std::lock_guard<std::mutex> g(m);
        blocks.push_back({p, bytes, dev});
    }
    void trim(int dev) {                       // dev < 0: every device
        std::lock_guard<std::mutex> g(m);
        int cur = 0;
        cudaGetDevice(&cur);
        for (size_t k = 0; k < blocks.size();) {
            if (dev >= 0 && blocks[k].dev != dev) { k++; continue; }
            cudaSetDevice(blocks[k].dev);
            cudaFree(blocks[k].p);
            blocks.erase(blocks.begin() + k);
        }
        cudaSetDevice(cur);
    }
};
DevPool& dev_pool() { static DevPool* pool = new DevPool(); return *pool; }    // never destroyed: no cudaFree after the runtime is gone

template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;            // elements the caller may use (the block may be larger)
    size_t bytes = 0;          // size of the block as allocated
    int dev = 0;
    void release() {
        if (p) dev_pool().give(p, bytes, dev);
        p = nullptr;
        cap = bytes = 0;
    }
    ~DevBuf() { release(); }
    cudaError_t ensure(size_t n) {
        if (n <= cap) return cudaSuccess;
        release();
        cudaGetDevice(&dev);
        const size_t want = n * sizeof(T);
        size_t got = 0;
        void* q = dev_pool().take(want, dev, &got);
        if (!q) {
            cudaError_t e = cudaMalloc(&q, want);
            if (e != cudaSuccess) {            // give the cached blocks back and try once more
                cudaGetLastError();
                dev_pool().trim(dev);
                e = cudaMalloc(&q, want);
                if (e != cudaSuccess) return e;
            }
            got = want;
        }
        p = (T*)q;
        bytes = got;
        cap = got / sizeof(T);
        return cudaSuccess;
    }
    cudaError_t upload(const std::vector<T>& v, cudaStream_t s) {
        cudaError_t e = ensure(v.size() ? v.size() : 1);
        if (e != cudaSuccess) return e;
        if (v.empty()) return cudaSuccess;
        return cudaMemcpyAsync(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, s);
    }
};

template <class T>
struct PinBuf {
    T* p = nullptr;
    size_t cap = 0;
    ~PinBuf() { if (p) cudaFreeHost(p); }
    cudaError_t ensure(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMallocHost((void**)&p, n * sizeof(T));
        if (e == cudaSuccess) cap = n;
        return e;
    }
};

// (invc, logc) table for tpl::fast_log, built in long double (see tpl_math.cuh)
void build_log_table(std::vector<double2>& tab) {
    tab.resize(tpl::LOG_TABLE_N);
    const unsigned long long OFF = 0x3fe6a09e00000000ull;
    for (int i = 0; i < tpl::LOG_TABLE_N; i++) {
        unsigned long long bits = OFF + ((unsigned long long)i << (52 - tpl::LOG_TABLE_BITS)) +
                                  (1ull << (51 - tpl::LOG_TABLE_BITS));
        // the interval that straddles 1.0 crosses an exponent boundary: centre it in VALUE space
        unsigned long long lo = OFF + ((unsigned long long)i << (52 - tpl::LOG_TABLE_BITS));
        unsigned long long hi = lo + (1ull << (52 - tpl::LOG_TABLE_BITS)) - 1;
        double dlo, dhi, dc;
        memcpy(&dlo, &lo, 8);
        memcpy(&dhi, &hi, 8);
        memcpy(&dc, &bits, 8);
        if ((lo >> 52) != (hi >> 52)) dc = 0.5 * (dlo + dhi);
        double invc = (double)(1.0L / (long double)dc);
        double logc = (double)(-logl((long double)invc));
        tab[i].x = invc;
        tab[i].y = logc;
    }
}

}  // namespace

struct tpl_engine {
    int device = 0;
    int sm = 0;
    int num_sms = 148;
    size_t max_smem = 227 * 1024;
    int n = 0;
    cudaStream_t stream = nullptr;
    // ---- host mirrors of pl_calc_parallel's inputs / members ----
    std::vector<int> parent, child_off, children, free_;
    std::vector<double> c, lf, hmin, hmax, pen_min, pen_max;
    std::vector<double> rates, dates, durations;     // member state (base values)
    std::vector<int> cv;                             // cvnodes
    std::vector<int> freeparams;                     // 2n
    int numparams = 0;
    bool lf_mode = false;
    bool have_fp = false;
    double smoothing = 1.0;
    double pb = 0.0025;
    bool log_pen = false;
    // last calc_pl vector = contents of the pinned staging buffer h_X (the member state is the
    // scatter of it over the base member values; materialised lazily, only when someone reads it)
    bool last_x_valid = false;
    // calc_pl -> calc_gradient pairs (every TNC / NLopt callback, optimize_tnc.cpp:75-105): once the pattern has been seen,
    // calc_pl computes the analytic gradient in the same launch and calc_gradient returns it from the pinned buffer; a second
    // calc_pl in a row (the annealer, siman_calc_par.cpp:154) switches the speculation off again
    bool spec_grad = false, g_cached = false;
    int pl_since_grad = 0;
    double g_lambda = 0.0, g_pb = 0.0;
    bool g_logpen = false;
    // ---- device tree ----
    DevBuf<int> d_parent, d_rslot, d_dslot, d_child_off, d_children, d_bar_ds;
    DevBuf<double> d_bar_hfix;
    DevBuf<unsigned char> d_cv;
    DevBuf<double> d_c, d_lf, d_hfix, d_rfix, d_bar_pmin, d_bar_pmax, d_bnd_min, d_bnd_max;
    DevBuf<int> d_bnd_ds;
    DevBuf<double2> d_logtab;
    int nbnd = 0;
    // generation-2 (tile kernel) operands
    DevBuf<int4> d_recs;
    DevBuf<double2> d_clf;
    DevBuf<tpl::TileInfo> d_tiles;
    int tile_hrows = 0;
    int ntiles2 = 0;
    bool gen2_ok = false;             // slot map is dense per tile (reference PL layouts are)
    // generation-4 operands (per-tile descriptors + halo row lists)
    DevBuf<int4> d_desc;
    DevBuf<tpl::DescTile> d_dtiles;
    DevBuf<int> d_halo;
    bool gen4_ok = false;
    double gen4_fast_frac = 0.0;
    bool cv_any = false;
    int force_gen = 0;                // test hook: 1 = always generation 1, 2 = generation 2 when legal
    int nbar = 0;
    double sum_c = 0.0;
    bool static_bad = false;          // a FIXED date violates its own bounds (set_durations :887-896)
    bool topo_dirty = true, slots_dirty = true, state_dirty = true;
    // ---- evaluation buffers ----
    DevBuf<double> d_X, d_G, d_F, d_lane_lambda;
    // objective partials: slot 0 for the plain paths, slots 1..3 for the pipelined host path (whose
    // sub-batches run on separate streams and may overlap)
    struct PartSet {
        DevBuf<double> ll, su, gr;
        DevBuf<int> bad;
    };
    PartSet parts[4];
    // pipelined host path (tpl_calc_pl_grad_batch): 3 slots of (stream, X, G, F) for 64-vector sub-batches
    cudaStream_t pstream[3] = {nullptr, nullptr, nullptr};
    DevBuf<double> p_X[3], p_G[3], p_F[3];
    DevBuf<unsigned long long> d_lanemask;
    PinBuf<double> h_X, h_G, h_F;
    // batch configuration
    int cfg_B = 0;
    bool cfg_lambda = false, cfg_mask = false;
    int cfg_nchunks = 0;
    int last_launches = 0;
    int last_gen = 0;
    // device-resident optimiser work space (tpl_optimize_batch)
    struct OptWork {
        DevBuf<double> Z, Gz, D, X, Gx, S, Y, lo, hi, sc, lane_d, rho, delta, M, part, part_red, tpart, Ft, Fout;
        DevBuf<int> lane_i;
        PinBuf<int> h_done;
        // truncated Newton
        DevBuf<double> R, ZZ, Pv, Hp, Minv, DG, pmin_n, pmax_n, tn_part, tn_lane_d;
        DevBuf<int> tn_lane_i, tn_lev_nodes, tn_lev_off, tn_lev_rec, tn_hv_rec, tn_cols, tn_probe_rows;
        DevBuf<double> tn_hv_const;
        DevBuf<double> tn_tw;
        std::vector<int> tn_lev_off_host;
        // annealer
        DevBuf<int> sa_date_node, sa_bar_node, sa_done;
        DevBuf<double> sa_hmin, sa_hmax;
        DevBuf<unsigned char> sa_chains;
    } opt;
    // generation-4 launch state that does not change between launches on the same buffers: the two tensor maps of X
    // (keyed by pointer and shape) and the opt-in shared-memory size already granted to each kernel instantiation
    struct MapCache {
        const double* X = nullptr;
        int P = 0, B = 0;
        CUtensorMap mapR, mapD;
    } maps[4];
    int maps_next = 0;
    std::vector<std::pair<const void*, size_t> > smem_granted;
    // optional per-launch device timing (bench.py roofline): events around the main kernel
    bool timing = false;
    std::vector<cudaEvent_t> ev;      // triples: before main, after main, after finalize
    size_t ev_used = 0;
};

namespace {

void materialize_state(tpl_engine* e) {
    // scatter the last calc_pl vector into the member rates/dates (pl_calc_parallel.cpp:82-99)
    if (!e->last_x_valid || !e->have_fp) return;
    const int n = e->n;
    for (int i = 0; i < n; i++)
        if (e->freeparams[i] != -1) e->rates[i] = e->h_X.p[e->freeparams[i]];
    for (int i = 0; i < n; i++)
        if (e->freeparams[n + i] != -1) e->dates[i] = e->h_X.p[e->freeparams[n + i]];
    // set_durations (:883-910), without the early outs
    for (int i = 0; i < n; i++) {
        if (i != 0) e->durations[i] = e->dates[e->parent[i]] - e->dates[i];
        if (e->durations[i] == 0 || std::isnan(e->durations[i]))
            e->durations[i] = std::numeric_limits<double>::min();
    }
    e->last_x_valid = false;
    // no device refresh needed: hfix/rfix are only READ for slots that are not free, and those
    // member values are not touched by the scatter; a new slot map re-uploads them (slots_dirty)
}

int sync_device_tree(tpl_engine* e) {
    cudaStream_t s = e->stream;
    const int n = e->n;
    if (e->topo_dirty) {
        CK(e->d_parent.upload(e->parent, s));
        CK(e->d_child_off.upload(e->child_off, s));
        CK(e->d_children.upload(e->children, s));
        CK(e->d_c.upload(e->c, s));
        CK(e->d_lf.upload(e->lf, s));
        const int npad = ((n + tpl::TILE_NODES - 1) / tpl::TILE_NODES) * tpl::TILE_NODES;
        std::vector<double2> clf(npad, make_double2(0.0, 0.0));     // padded: the persistent kernel copies whole tiles
        for (int i = 0; i < n; i++) clf[i] = make_double2(e->c[i], e->lf[i]);
        CK(e->d_clf.upload(clf, s));
        e->sum_c = 0.0;
        for (int i = 0; i < n; i++) e->sum_c += e->c[i];
        std::vector<double2> tab;
        build_log_table(tab);
        CK(e->d_logtab.upload(tab, s));
        CK(cudaStreamSynchronize(s));   // the host vectors above are temporaries
        e->topo_dirty = false;
    }
    if (e->slots_dirty) {
        std::vector<int> rs(n, -1), ds(n, -1);
        std::vector<unsigned char> cvb(n, 0);
        if (e->have_fp)
            for (int i = 0; i < n; i++) {
                rs[i] = e->freeparams[i];
                ds[i] = e->freeparams[n + i];
            }
        for (int i = 0; i < n; i++) cvb[i] = e->cv[i] ? 1 : 0;
        CK(e->d_rslot.upload(rs, s));
        CK(e->d_dslot.upload(ds, s));
        CK(e->d_cv.upload(cvb, s));
        CK(cudaStreamSynchronize(s));
        e->slots_dirty = false;
        e->state_dirty = true;   // static_bad depends on which dates are fixed
    }
    if (e->state_dirty) {
        CK(e->d_hfix.upload(e->dates, s));
        CK(e->d_rfix.upload(e->rates, s));
        e->static_bad = false;
        for (int i = 0; i < n; i++) {
            bool fixed = !e->have_fp || e->freeparams[n + i] == -1;
            if (fixed && (e->dates[i] < e->hmin[i] || e->dates[i] > e->hmax[i])) e->static_bad = true;
        }
        // ---- sparse calibration barrier rows, ascending node order (calc_penalty :975-998) ----
        {
            std::vector<int> bds;
            std::vector<double> bhf, pmin, pmax;
            for (int i = 0; i < n; i++)
                if (e->pen_min[i] != -1 || e->pen_max[i] != -1) {
                    bds.push_back(e->have_fp ? e->freeparams[n + i] : -1);
                    bhf.push_back(e->dates[i]);
                    pmin.push_back(e->pen_min[i]);
                    pmax.push_back(e->pen_max[i]);
                }
            e->nbar = (int)bds.size();
            CK(e->d_bar_ds.upload(bds, s));
            CK(e->d_bar_hfix.upload(bhf, s));
            CK(e->d_bar_pmin.upload(pmin, s));
            CK(e->d_bar_pmax.upload(pmax, s));
        }
        // ---- sparse bound rows: free dates whose [min,max] is not implied by a neighbour ----
        std::vector<int> bn;
        std::vector<double> bmin, bmax;
        for (int i = 0; i < n; i++) {
            if (!e->have_fp || e->freeparams[n + i] == -1) continue;
            const int p = e->parent[i];
            bool max_implied = (p >= 0) && (e->hmax[i] >= e->hmax[p]);
            bool min_implied = false;
            for (int j = e->child_off[i]; j < e->child_off[i + 1]; j++)
                if (e->hmin[e->children[j]] >= e->hmin[i]) { min_implied = true; break; }
            if (!(max_implied && min_implied)) {
                bn.push_back(e->freeparams[n + i]);
                bmin.push_back(e->hmin[i]);
                bmax.push_back(e->hmax[i]);
            }
        }
        e->nbnd = (int)bn.size();
        CK(e->d_bnd_ds.upload(bn, s));
        CK(e->d_bnd_min.upload(bmin, s));
        CK(e->d_bnd_max.upload(bmax, s));
        // ---- generation-2 operands: packed records, tile table ----
        e->cv_any = false;
        for (int i = 0; i < n; i++) if (e->cv[i]) e->cv_any = true;
        const int npad = ((n + tpl::TILE_NODES - 1) / tpl::TILE_NODES) * tpl::TILE_NODES;
        std::vector<int4> recs(npad, make_int4(-1, -1, -1, -2));
        bool ok = e->have_fp && !e->lf_mode;
        int next_r = -1, next_d = -1;
        for (int i = 0; i < n && e->have_fp; i++) {
            int4 r;
            r.x = e->parent[i];
            const int rs = e->freeparams[i], ds = e->freeparams[n + i];
            r.y = rs >= 0 ? rs : -1;
            r.z = ds >= 0 ? ds : (e->dates[i] == 0.0 ? -1 : -2);
            const int nc = e->child_off[i + 1] - e->child_off[i];
            // rec.w: >= 0 second child of a FAST binary node (first child is i+1); -1 FAST tip;
            // -2 everything else (general path, children through the CSR list)
            auto plain = [&](int v) { return e->freeparams[v] >= 0; };                     // has a rate row
            auto dfree = [&](int v) { return e->freeparams[n + v] >= 0; };
            auto dzero = [&](int v) { return e->freeparams[n + v] < 0 && e->dates[v] == 0.0; };
            const int p = e->parent[i];
            const int t0 = (i / tpl::TILE_NODES) * tpl::TILE_NODES, t1 = t0 + tpl::TILE_NODES;
            auto local = [&](int v) { return v >= t0 && v < t1; };                         // same tile
            const bool par_ok = p > 0 && plain(p) && dfree(p) && local(p);
            r.w = -2;
            if (par_ok && plain(i)) {
                if (nc == 0 && dzero(i)) r.w = -1;
                else if (nc == 2 && dfree(i)) {
                    const int a0 = e->children[e->child_off[i]], a1 = e->children[e->child_off[i] + 1];
                    const int other = (a0 == i + 1) ? a1 : a0;
                    if ((a0 == i + 1 || a1 == i + 1) && a0 != a1 && local(a0) && local(a1) && plain(a0) && plain(a1) &&
                        (dfree(a0) || dzero(a0)) && (dfree(a1) || dzero(a1)))
                        r.w = other;
                }
            }
            recs[i] = r;
            if (rs >= 0) { if (next_r >= 0 && rs != next_r) ok = false; next_r = rs + 1; }
            if (ds >= 0) { if (next_d >= 0 && ds != next_d) ok = false; next_d = ds + 1; }
        }
        const int T = tpl::TILE_NODES;
        const int nt = (n + T - 1) / T;
        std::vector<tpl::TileInfo> tiles(nt);
        int hrows = 0;
        for (int t = 0; t < nt && e->have_fp; t++) {
            tpl::TileInfo ti = {0, 0, 0, 0};
            bool fr = true, fd = true;
            for (int i = t * T; i < std::min(n, (t + 1) * T); i++) {
                const int rs = e->freeparams[i], ds = e->freeparams[n + i];
                if (rs >= 0) { if (fr) { ti.rs0 = rs; fr = false; } ti.nrs = rs - ti.rs0 + 1; }
                if (ds >= 0) { if (fd) { ti.ds0 = ds; fd = false; } ti.nds = ds - ti.ds0 + 1; }
            }
            if (ti.nrs > T || ti.nds > T || ti.nrs < 0 || ti.nds < 0) ok = false;
            hrows = std::max(hrows, ti.nds);
            tiles[t] = ti;
        }
        e->gen2_ok = ok;
        e->ntiles2 = nt;
        e->tile_hrows = ((std::max(hrows, 1) + tpl::DESC_DBOX - 1) / tpl::DESC_DBOX) * tpl::DESC_DBOX;   // whole date boxes
        CK(e->d_recs.upload(recs, s));
        CK(e->d_tiles.upload(tiles, s));
        // ---- generation-4 operands: every operand of every node as a tile-local row index ----
        {
            const int HM = tpl::DESC_HALO;
            const int hr = e->tile_hrows;
            const int ZERO = T + hr + HM;                     // local index of the all-zero row
            std::vector<int4> desc((size_t)nt * T * 4, make_int4(0, 0, -2, -1));   // default: general row
            std::vector<tpl::DescTile> dt(nt);
            std::vector<int> halo((size_t)nt * HM, 0);
            const bool usable = ok && ZERO <= 255;
            long long nfast = 0;
            const std::vector<int>& fp = e->freeparams;
            for (int t = 0; t < nt && usable; t++) {
                const tpl::TileInfo ti = tiles[t];
                const int t0 = t * T, t1 = std::min(n, t0 + T);
                int nh = 0;
                int* hl = &halo[(size_t)t * HM];
                auto halo_row = [&](int xrow) -> int {         // local index of a row that lives in another tile
                    for (int k = 0; k < nh; k++) if (hl[k] == xrow) return T + hr + k;
                    if (nh >= HM) return -1;
                    hl[nh] = xrow;
                    return T + hr + nh++;
                };
                auto loc_r = [&](int v) -> int {
                    const int rs = fp[v];
                    if (rs < 0) return -1;
                    return (v >= t0 && v < t1) ? rs - ti.rs0 : halo_row(rs);
                };
                auto loc_h = [&](int v) -> int {
                    const int ds = fp[n + v];
                    if (ds >= 0) return (v >= t0 && v < t1) ? T + ds - ti.ds0 : halo_row(ds);
                    return e->dates[v] == 0.0 ? ZERO : -1;
                };
                for (int i = t0; i < t1; i++) {
                    const int p = e->parent[i];
                    if (i == 0 || p <= 0) continue;                                   // root, children of the root
                    const int rs = fp[i], ds = fp[n + i];
                    if (rs < 0 || fp[p] < 0 || fp[n + p] < 0) continue;
                    const int nc = e->child_off[i + 1] - e->child_off[i];
                    if (!(nc == 0 || nc == 2)) continue;
                    if (nc == 0 && !(ds < 0 && e->dates[i] == 0.0)) continue;
                    if (nc == 2 && ds < 0) continue;
                    const int save = nh;
                    int idx[8] = {rs - ti.rs0, nc == 0 ? ZERO : T + ds - ti.ds0, loc_r(p), loc_h(p), ZERO, ZERO, ZERO, ZERO};
                    int c1 = -1, c2 = -1;
                    bool good = idx[2] >= 0 && idx[3] >= 0;
                    if (good && nc == 2) {
                        const int a0 = e->children[e->child_off[i]], a1 = e->children[e->child_off[i] + 1];
                        c1 = (a0 == i + 1) ? a0 : a1;
                        c2 = (a0 == i + 1) ? a1 : a0;
                        if (c1 != i + 1 || c1 == c2) good = false;
                        if (good) {
                            idx[4] = loc_r(c1); idx[5] = loc_h(c1); idx[6] = loc_r(c2); idx[7] = loc_h(c2);
                            for (int k = 4; k < 8; k++) if (idx[k] < 0) good = false;
                        }
                    }
                    if (!good) { nh = save; continue; }
                    int4* q = &desc[((size_t)t * T + (i - t0)) * 4];
                    q[0].x = idx[0] | (idx[1] << 8) | (idx[2] << 16) | (int)((unsigned)idx[3] << 24);
                    q[0].y = idx[4] | (idx[5] << 8) | (idx[6] << 16) | (int)((unsigned)idx[7] << 24);
                    q[0].z = rs;
                    q[0].w = nc == 2 ? ds : -1;
                    double2 cl = make_double2(e->c[i], e->lf[i]);
                    double2 cc = make_double2(nc == 2 ? e->c[c1] : 0.0, nc == 2 ? e->c[c2] : 0.0);
                    memcpy(&q[1], &cl, 16);
                    memcpy(&q[2], &cc, 16);
                    q[3] = make_int4((c1 >= t0 && c1 < t1) ? c1 - t0 : -1, (c2 >= t0 && c2 < t1) ? c2 - t0 : -1,
                                     c1 >= 0 ? c1 : 0, c2 >= 0 ? c2 : 0);
                    nfast++;
                }
                tpl::DescTile d = {ti.rs0, ti.nrs, ti.ds0, ti.nds, nh, 0, 0, 0};
                dt[t] = d;
            }
            e->gen4_ok = usable;
            e->gen4_fast_frac = n > 0 ? (double)nfast / n : 0.0;
            CK(e->d_desc.upload(desc, s));
            CK(e->d_dtiles.upload(dt, s));
            CK(e->d_halo.upload(halo, s));
        }
        CK(cudaStreamSynchronize(s));
        e->state_dirty = false;
    }
    return TPL_OK;
}

tpl::DevTree dev_tree(tpl_engine* e) {
    tpl::DevTree t;
    t.n = e->n;
    t.parent = e->d_parent.p;
    t.rslot = e->d_rslot.p;
    t.dslot = e->d_dslot.p;
    t.child_off = e->d_child_off.p;
    t.children = e->d_children.p;
    t.cv = e->d_cv.p;
    t.c = e->d_c.p;
    t.lf = e->d_lf.p;
    t.hfix = e->d_hfix.p;
    t.rfix = e->d_rfix.p;
    t.nbnd = e->nbnd;
    t.bnd_ds = e->d_bnd_ds.p;
    t.bnd_min = e->d_bnd_min.p;
    t.bnd_max = e->d_bnd_max.p;
    t.nbar = e->nbar;
    t.bar_ds = e->d_bar_ds.p;
    t.bar_hfix = e->d_bar_hfix.p;
    t.bar_pmin = e->d_bar_pmin.p;
    t.bar_pmax = e->d_bar_pmax.p;
    t.sum_c = e->sum_c;
    double d0 = e->durations[0];
    if (d0 == 0 || std::isnan(d0)) d0 = std::numeric_limits<double>::min();
    t.dur0 = d0;
    return t;
}

constexpr int G4_NCW = tpl::DESC_NCW;  // consumer warps of the generation-4 kernel

// cuTensorMapEncodeTiled through the runtime (no link against libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encoder() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}
// X viewed as [P rows][B columns] fp64, box = 64 columns x box_rows rows
bool make_x_map(CUtensorMap* m, const double* X, int P, int B, int box_rows) {
    EncodeTiledFn enc = get_encoder();
    if (!enc) return false;
    cuuint64_t dims[2] = {(cuuint64_t)B, (cuuint64_t)P};
    cuuint64_t strides[1] = {(cuuint64_t)B * 8};
    cuuint32_t box[2] = {(cuuint32_t)tpl::TILE_LANES, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)X, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

struct LaunchCfg {
    tpl_engine* eng;
    int gen;
    int stages = 3;
    tpl::DescArgs da;
    CUtensorMap tmapR, tmapD;
    bool lf, lp, want_g;
    int maskmode;          // 0 none, 1 engine cv, 2 per-vector
    bool gen2;
    dim3 grid, block;
    size_t smem;
    tpl::TileArgs ta;
    cudaStream_t s;
    cudaEvent_t* ev;
};

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (engine, kernel instantiation, size), not per launch
template <class K>
cudaError_t grant_smem(tpl_engine* e, K kern, size_t smem) {
    const void* key = (const void*)kern;
    for (auto& kv : e->smem_granted)
        if (kv.first == key) {
            if (kv.second >= smem) return cudaSuccess;
            cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (err == cudaSuccess) kv.second = smem;
            return err;
        }
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err == cudaSuccess) e->smem_granted.push_back(std::make_pair(key, smem));
    return err;
}

template <int MODE, bool LF, bool LOGPEN, bool WANT_G, int MASKMODE>
cudaError_t launch_pair(const tpl::EvalArgs& a, const LaunchCfg& c) {
    if (c.ev) cudaEventRecord(c.ev[0], c.s);
    if (c.gen2) {
        if constexpr (!LF && MASKMODE != 1) {
            if (c.gen == 4) {
                // LOGPEN only matters to the analytic gradient, which never runs here (see desc_legal)
                cudaError_t e = cudaSuccess;
                if (c.stages == 3) {
                    auto kern = tpl::pl_desc_kernel<MODE, false, WANT_G, MASKMODE == 2, 3, G4_NCW>;
                    e = grant_smem(c.eng, kern, c.smem);
                    if (e != cudaSuccess) return e;
                    kern<<<c.grid, c.block, c.smem, c.s>>>(a, c.ta, c.da, c.tmapR, c.tmapD);
                } else {
                    auto kern = tpl::pl_desc_kernel<MODE, false, WANT_G, MASKMODE == 2, 2, G4_NCW>;
                    e = grant_smem(c.eng, kern, c.smem);
                    if (e != cudaSuccess) return e;
                    kern<<<c.grid, c.block, c.smem, c.s>>>(a, c.ta, c.da, c.tmapR, c.tmapD);
                }
            } else if (c.gen == 3) {
                auto kern = tpl::pl_persist_kernel<MODE, LOGPEN, WANT_G, MASKMODE == 2>;
                cudaError_t e = grant_smem(c.eng, kern, c.smem);
                if (e != cudaSuccess) return e;
                kern<<<c.grid, c.block, c.smem, c.s>>>(a, c.ta);
            } else {
                auto kern = tpl::pl_tile_kernel<MODE, LOGPEN, WANT_G, MASKMODE == 2>;
                cudaError_t e = grant_smem(c.eng, kern, c.smem);
                if (e != cudaSuccess) return e;
                kern<<<c.grid, c.block, c.smem, c.s>>>(a, c.ta);
            }
        }
    } else {
        tpl::pl_main_kernel<MODE, LF, LOGPEN, WANT_G, MASKMODE><<<c.grid, c.block, 0, c.s>>>(a);
    }
    if (c.ev) cudaEventRecord(c.ev[1], c.s);
    tpl::pl_finalize_kernel<MODE, LF, LOGPEN, WANT_G, MASKMODE>
        <<<(a.B + 31) / 32, dim3(32, tpl::FIN_SEGS), 0, c.s>>>(a);
    if (c.ev) cudaEventRecord(c.ev[2], c.s);
    return cudaGetLastError();
}

template <int MODE, bool LF, bool LOGPEN, bool WANT_G>
cudaError_t launch_m(const tpl::EvalArgs& a, const LaunchCfg& c) {
    if (c.maskmode == 2) return launch_pair<MODE, LF, LOGPEN, WANT_G, 2>(a, c);
    if (c.maskmode == 1) return launch_pair<MODE, LF, LOGPEN, WANT_G, 1>(a, c);
    return launch_pair<MODE, LF, LOGPEN, WANT_G, 0>(a, c);
}
template <int MODE, bool LF, bool LOGPEN>
cudaError_t launch_g(const tpl::EvalArgs& a, const LaunchCfg& c) {
    return c.want_g ? launch_m<MODE, LF, LOGPEN, true>(a, c) : launch_m<MODE, LF, LOGPEN, false>(a, c);
}
template <int MODE>
cudaError_t launch_lf(const tpl::EvalArgs& a, const LaunchCfg& c) {
    if (c.lf) return launch_g<MODE, true, false>(a, c);    // log_pen has no effect on the LF objective
    return c.lp ? launch_g<MODE, false, true>(a, c) : launch_g<MODE, false, false>(a, c);
}

// Evaluate B vectors resident on the device.  Enqueues 2 kernels on stream s.
// lane_off / cfg_B: this launch evaluates vectors [lane_off, lane_off + B) of a configured batch of cfg_B
// vectors (per-vector smoothing and masks are addressed relative to lane_off, a multiple of 64).
int eval_device(tpl_engine* e, int B, const double* dX, double* dF, double* dG, int mode, bool use_cfg,
                cudaStream_t s, int lane_off = 0, int cfg_B = -1, int part_slot = 0) {
    if (cfg_B < 0) cfg_B = B;
    if (!e->have_fp) return fail(TPL_ERR_STATE, "set_freeparams has not been called");
    if (mode != TPL_GRAD_ANALYTIC && mode != TPL_GRAD_AD) return fail(TPL_ERR_ARG, "bad gradient mode");
    if (B < 1) return fail(TPL_ERR_ARG, "B < 1");
    int rc = sync_device_tree(e);
    if (rc != TPL_OK) return rc;
    const int n = e->n;
    const bool lanemask = use_cfg && e->cfg_mask && e->cfg_B == cfg_B;
    const int maskmode = lanemask ? 2 : (e->cv_any ? 1 : 0);
    // generations 2/3 (tile kernels): PL, dense slot map, even B, 32-bit element offsets
    const bool tile_legal = e->gen2_ok && !e->lf_mode && maskmode != 1 && (B % 2 == 0) &&
                            (double)e->numparams * (double)B < 2.0e9;
    const int nchunks2 = (B + tpl::TILE_LANES - 1) / tpl::TILE_LANES;
    const long long W = (long long)nchunks2 * e->ntiles2;
    const int G = (int)std::min<long long>(e->num_sms, W);
    // generation 3 (persistent ring) additionally needs every CTA's item range to touch <= 2 chunks
    const bool persist_legal = tile_legal && nchunks2 <= G && tpl::persist_kernel_smem(e->tile_hrows) <= e->max_smem;
    // generation 4 (descriptors + prefetched halo): as 3, but the analytic log_pen gradient has no fast path
    int g4_stages = 0;
    if (tpl::desc_kernel_smem(e->tile_hrows, 3) <= e->max_smem) g4_stages = 3;
    else if (tpl::desc_kernel_smem(e->tile_hrows, 2) <= e->max_smem) g4_stages = 2;
    const bool desc_legal = tile_legal && e->gen4_ok && nchunks2 <= G && g4_stages > 0 &&
                            !(e->log_pen && mode == TPL_GRAD_ANALYTIC);
    int gen = 1;
    if (tile_legal && B >= 32) gen = desc_legal ? 4 : (persist_legal ? 3 : 2);
    if (e->force_gen == 1) gen = 1;
    if (e->force_gen == 2 && tile_legal) gen = 2;
    if (e->force_gen == 3 && persist_legal) gen = 3;
    if (e->force_gen == 4 && desc_legal) gen = 4;
    const bool gen2 = gen >= 2;
    LaunchCfg c;
    c.eng = e;
    c.gen = gen;
    int ntiles, tile;
    size_t npart;
    int nwarps = tpl::P_CONSUMER_WARPS;
    if (gen == 4) {
        tile = tpl::TILE_NODES;
        ntiles = e->ntiles2;
        nwarps = 1;                    // the kernel reduces its consumer warps' partials itself: one row per CTA and slot
        c.grid = dim3(G);
        c.block = dim3((G4_NCW + 1) * 32);
        c.smem = tpl::desc_kernel_smem(e->tile_hrows, g4_stages);
        c.stages = g4_stages;
        npart = (size_t)G * 2 * tpl::TILE_LANES;
    } else if (gen == 3) {
        tile = tpl::TILE_NODES;
        ntiles = e->ntiles2;
        c.grid = dim3(G);
        c.block = dim3(tpl::P_THREADS);
        c.smem = tpl::persist_kernel_smem(e->tile_hrows);
        npart = (size_t)G * 2 * tpl::P_CONSUMER_WARPS * tpl::TILE_LANES;
    } else if (gen == 2) {
        tile = tpl::TILE_NODES;
        ntiles = e->ntiles2;
        c.grid = dim3(ntiles, nchunks2);
        c.block = dim3(256);
        c.smem = tpl::tile_kernel_smem(e->tile_hrows);
        npart = (size_t)ntiles * B;
    } else {
        // block shape: LX lanes x NY node rows, 256 threads
        int LX = 1;
        while (LX < 32 && LX < B) LX <<= 1;
        const int NY = 256 / LX;
        tile = (LX == 1) ? 1024 : 256;
        ntiles = (n + tile - 1) / tile;
        c.grid = dim3(ntiles, (B + LX - 1) / LX);
        c.block = dim3(LX, NY);
        c.smem = 0;
        npart = (size_t)ntiles * B;
    }
    tpl_engine::PartSet& ps = e->parts[part_slot];
    CK(ps.ll.ensure(npart));
    CK(ps.su.ensure(npart));
    CK(ps.gr.ensure(npart));
    CK(ps.bad.ensure(npart));
    tpl::EvalArgs a;
    a.t = dev_tree(e);
    a.X = dX;
    a.G = dG;
    a.F = dF;
    a.B = B;
    a.lanemask = lanemask ? e->d_lanemask.p + (size_t)(lane_off / 64) * (size_t)(e->ntiles2 * tpl::TILE_NODES) : nullptr;
    a.nchunks = e->cfg_nchunks;
    a.mask_stride = e->ntiles2 * tpl::TILE_NODES;
    a.part_mode = gen >= 3 ? 1 : 0;
    a.p_ncta = G;
    a.p_ntiles = e->ntiles2;
    a.p_nchunks = nchunks2;
    a.p_nwarps = nwarps;
    a.lane_lambda = (use_cfg && e->cfg_lambda && e->cfg_B == cfg_B) ? e->d_lane_lambda.p + lane_off : nullptr;
    a.lambda = e->smoothing;
    a.pb = e->pb;
    a.part_ll = ps.ll.p;
    a.part_su = ps.su.p;
    a.part_gr = ps.gr.p;
    a.part_bad = ps.bad.p;
    a.ntiles = ntiles;
    a.tile_nodes = tile;
    a.logtab = e->d_logtab.p;
    c.lf = e->lf_mode;
    c.lp = e->log_pen;
    c.want_g = dG != nullptr;
    c.maskmode = maskmode;
    c.gen2 = gen2;
    c.ta.recs = e->d_recs.p;
    c.ta.clf = e->d_clf.p;
    c.ta.tiles = e->d_tiles.p;
    c.ta.hrows = e->tile_hrows;
    c.ta.ntiles = e->ntiles2;
    c.ta.nchunks = nchunks2;
    c.ta.ncta = G;
    c.da.desc = e->d_desc.p;
    c.da.tiles = e->d_dtiles.p;
    c.da.halo_rows = e->d_halo.p;
    c.da.hrows = e->tile_hrows;
    c.da.ntiles = e->ntiles2;
    c.da.nchunks = nchunks2;
    c.da.ncta = G;
    if (gen == 4) {
        tpl_engine::MapCache* mc = nullptr;
        for (auto& m : e->maps)
            if (m.X == dX && m.P == e->numparams && m.B == B) mc = &m;
        if (!mc) {
            mc = &e->maps[e->maps_next];
            e->maps_next = (e->maps_next + 1) % 4;
            mc->X = nullptr;
            if (!make_x_map(&mc->mapR, dX, e->numparams, B, tpl::TILE_NODES) ||
                !make_x_map(&mc->mapD, dX, e->numparams, B, tpl::DESC_DBOX))
                return fail(TPL_ERR_CUDA, "cuTensorMapEncodeTiled failed");
            mc->X = dX; mc->P = e->numparams; mc->B = B;
        }
        c.tmapR = mc->mapR;
        c.tmapD = mc->mapD;
    }
    c.s = s;
    c.ev = nullptr;
    if (e->timing && e->ev_used + 3 <= e->ev.size()) { c.ev = &e->ev[e->ev_used]; e->ev_used += 3; }
    CK(mode == TPL_GRAD_ANALYTIC ? launch_lf<0>(a, c) : launch_lf<1>(a, c));
    e->last_gen = gen;
    e->last_launches = 2;
    return TPL_OK;
}

// single point: host x -> device, evaluate, bring f (and g) back
double eval_point(tpl_engine* e, const double* x, double* g, int mode, bool keep_state, bool spec_g = false) {
    const double nan = std::numeric_limits<double>::quiet_NaN();
    g_err.clear();      // a NaN RESULT is legal (the reference propagates NaN); callers tell it from a failure by tpl_last_error()
    if (!e || !x) { fail(TPL_ERR_ARG, "null argument"); return nan; }
    if (!e->have_fp) { fail(TPL_ERR_STATE, "set_freeparams has not been called"); return nan; }
    if (cudaSetDevice(e->device) != cudaSuccess) { fail(TPL_ERR_CUDA, "cudaSetDevice failed"); return nan; }
    e->g_cached = false;
    const bool want_g = g != nullptr || spec_g;
    const int P = e->numparams;
    if (!keep_state && e->last_x_valid && x != e->h_X.p) materialize_state(e);   // h_X is about to be overwritten
    if (e->h_X.ensure(P) != cudaSuccess || e->h_G.ensure(P) != cudaSuccess || e->h_F.ensure(1) != cudaSuccess ||
        e->d_X.ensure(P) != cudaSuccess || e->d_G.ensure(P) != cudaSuccess || e->d_F.ensure(1) != cudaSuccess) {
        fail(TPL_ERR_CUDA, "allocation failed");
        return nan;
    }
    if (x != e->h_X.p) memcpy(e->h_X.p, x, sizeof(double) * P);
    cudaStream_t s = e->stream;
    if (cudaMemcpyAsync(e->d_X.p, e->h_X.p, sizeof(double) * P, cudaMemcpyHostToDevice, s) != cudaSuccess) {
        fail(TPL_ERR_CUDA, "H2D copy failed");
        return nan;
    }
    if (eval_device(e, 1, e->d_X.p, e->d_F.p, want_g ? e->d_G.p : nullptr, mode, false, s) != TPL_OK) return nan;
    cudaMemcpyAsync(e->h_F.p, e->d_F.p, sizeof(double), cudaMemcpyDeviceToHost, s);
    if (want_g) cudaMemcpyAsync(e->h_G.p, e->d_G.p, sizeof(double) * P, cudaMemcpyDeviceToHost, s);
    cudaError_t err = cudaStreamSynchronize(s);
    if (err != cudaSuccess) { fail(TPL_ERR_CUDA, std::string("evaluation failed: ") + cudaGetErrorString(err)); return nan; }
    double f = e->h_F.p[0];
    if (e->static_bad) f = TPL_LARGE;
    if (keep_state) e->last_x_valid = true;
    if (want_g && mode == TPL_GRAD_ANALYTIC && keep_state) {      // h_G = analytic gradient at the state calc_pl just left
        e->g_cached = true;
        e->g_lambda = e->smoothing; e->g_pb = e->pb; e->g_logpen = e->log_pen;
    }
    if (g) {
        const bool infeasible = (f == TPL_LARGE);
        if (mode == TPL_GRAD_AD && infeasible) {
            // the reference returns before touching g (:377,:387,:398)
        } else {
            memcpy(g, e->h_G.p, sizeof(double) * P);
            if (mode == TPL_GRAD_ANALYTIC && e->lf_mode && std::isnan(g[0]))
                for (int i = 0; i < P; i++) g[i] = 1000000;               // calc_lf_gradient :680-687
        }
    }
    return f;
}


#include "tpl_optimize_host.inl"   // run_optimizer, tn_prepare, optimize_tn, tpl_optimize_batch, tpl_debug_hessvec


extern "C" {

const char* tpl_last_error(void) { return g_err.c_str(); }

int tpl_create(tpl_engine** out, int device, int numnodes, const int* parents, const int* child_counts,
               const int* children, const int* free_, const double* c, const double* lf, const double* mn,
               const double* mx, const double* pen_min, const double* pen_max, const double* start_dates,
               const double* start_rates, const double* start_durations) {
    if (!out || numnodes < 1 || !parents || !child_counts || !free_ || !c || !lf || !mn || !mx || !pen_min ||
        !pen_max || !start_dates || !start_rates || !start_durations || (numnodes > 1 && !children))
        return fail(TPL_ERR_ARG, "tpl_create: null or empty argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1)
        return fail(TPL_ERR_CUDA, "no CUDA device: the PL engine has no CPU fallback");
    if (device < 0) CK(cudaGetDevice(&device));
    if (device >= ndev) return fail(TPL_ERR_ARG, "device ordinal out of range");
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return fail(TPL_ERR_CUDA, "device is not sm_100 class; this library carries sm_100a code only");
    tpl_engine* e = new tpl_engine();
    e->device = device;
    e->sm = prop.major * 100 + prop.minor * 10;
    e->num_sms = prop.multiProcessorCount;
    e->max_smem = prop.sharedMemPerBlockOptin;
    e->n = numnodes;
    const int n = numnodes;
    e->parent.assign(parents, parents + n);
    e->child_off.assign(n + 1, 0);
    for (int i = 0; i < n; i++) e->child_off[i + 1] = e->child_off[i] + child_counts[i];
    if (e->child_off[n] != n - 1) { delete e; return fail(TPL_ERR_ARG, "sum(child_counts) != numnodes-1"); }
    e->children.assign(children, children + (n - 1));
    for (int i = 0; i < n; i++)
        for (int j = e->child_off[i]; j < e->child_off[i + 1]; j++) {
            int ch = e->children[j];
            if (ch < 1 || ch >= n || e->parent[ch] != i) { delete e; return fail(TPL_ERR_ARG, "children / parents disagree"); }
        }
    e->free_.assign(free_, free_ + n);
    e->c.assign(c, c + n);
    e->lf.assign(lf, lf + n);
    e->hmin.assign(mn, mn + n);
    e->hmax.assign(mx, mx + n);
    e->pen_min.assign(pen_min, pen_min + n);
    e->pen_max.assign(pen_max, pen_max + n);
    e->dates.assign(start_dates, start_dates + n);
    e->rates.assign(start_rates, start_rates + n);
    e->durations.assign(start_durations, start_durations + n);
    e->cv.assign(n, 0);
    cudaError_t err = cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking);
    if (err != cudaSuccess) { delete e; return fail(TPL_ERR_CUDA, "cudaStreamCreate failed"); }
    *out = e;
    return TPL_OK;
}

void tpl_destroy(tpl_engine* e) {
    if (!e) return;
    cudaSetDevice(e->device);
    if (e->stream) { cudaStreamSynchronize(e->stream); cudaStreamDestroy(e->stream); }
    for (int k = 0; k < 3; k++) if (e->pstream[k]) { cudaStreamSynchronize(e->pstream[k]); cudaStreamDestroy(e->pstream[k]); }
    for (auto& v : e->ev) cudaEventDestroy(v);
    cudaDeviceSynchronize();            // the engine's blocks go back to the process-wide cache: nothing may still be using them
    delete e;
}

int tpl_release_cached_memory(void) {
    dev_pool().trim(-1);
    return TPL_OK;
}

int tpl_set_smoothing(tpl_engine* e, double s) { if (!e) return fail(TPL_ERR_ARG, "null"); e->smoothing = s; return TPL_OK; }
int tpl_set_log_pen(tpl_engine* e, int lp) { if (!e) return fail(TPL_ERR_ARG, "null"); e->log_pen = lp != 0; return TPL_OK; }
int tpl_set_penalty_boundary(tpl_engine* e, double pb) { if (!e) return fail(TPL_ERR_ARG, "null"); e->pb = pb; return TPL_OK; }

int tpl_set_cv_nodes(tpl_engine* e, const int* cv) {
    if (!e) return fail(TPL_ERR_ARG, "null");
    if (cv) e->cv.assign(cv, cv + e->n); else e->cv.assign(e->n, 0);
    e->slots_dirty = true;
    return TPL_OK;
}

int tpl_set_state(tpl_engine* e, const double* rates, const double* dates) {
    if (!e || !rates || !dates) return fail(TPL_ERR_ARG, "null");
    e->last_x_valid = false;
    e->rates.assign(rates, rates + e->n);
    e->dates.assign(dates, dates + e->n);
    e->state_dirty = true;
    return TPL_OK;
}

int tpl_generate_param_order_vector(int n, const int* free_, int lf, const int* cv, int* fp) {
    if (n < 1 || !free_ || !fp) return fail(TPL_ERR_ARG, "null");
    int icount = 0;                                   /* utils.cpp:295-340 */
    for (int i = 0; i < n; i++) {
        if (i == 0) fp[i] = -1;
        else if (cv) {
            if (cv[i] == 0) fp[i] = icount++; else fp[i] = -1;
        } else {
            fp[i] = icount;
            if (!lf) icount++;
        }
    }
    if (lf) icount++;
    for (int i = 0; i < n; i++) fp[n + i] = (free_[i] == 1) ? icount++ : -1;
    return icount;
}

int tpl_set_freeparams(tpl_engine* e, int nump, int lf, const int* freep, double* params_out) {
    if (!e || !freep || nump < 1) return fail(TPL_ERR_ARG, "null or empty");
    const int n = e->n;
    for (int i = 0; i < 2 * n; i++)
        if (freep[i] < -1 || freep[i] >= nump) return fail(TPL_ERR_ARG, "freeparams entry out of range");
    materialize_state(e);                              // x0 comes from the CURRENT member state
    e->freeparams.assign(freep, freep + 2 * n);
    e->numparams = nump;
    e->lf_mode = lf != 0;
    e->have_fp = true;
    e->slots_dirty = true;
    int k = 0;                                         /* pl_calc_parallel.cpp:930-951 */
    if (params_out) {
        for (int i = 0; i < n; i++)
            if (e->freeparams[i] != -1) {
                if (k < nump) params_out[k] = e->rates[i];
                k++;
                if (e->lf_mode) break;
            }
        for (int i = 0; i < n; i++)
            if (e->freeparams[n + i] != -1) {
                if (k < nump) params_out[k] = e->dates[i];
                k++;
            }
    }
    return k;
}

int tpl_numnodes(const tpl_engine* e) { return e ? e->n : TPL_ERR_ARG; }
int tpl_numparams(const tpl_engine* e) { return e ? e->numparams : TPL_ERR_ARG; }
int tpl_device_sm(const tpl_engine* e) { return e ? e->sm : TPL_ERR_ARG; }
int tpl_last_launch_count(const tpl_engine* e) { return e ? e->last_launches : 0; }

double tpl_calc_pl(tpl_engine* e, const double* x) {
    if (e && ++e->pl_since_grad > 1) e->spec_grad = false;           // calc_pl twice in a row: nobody is asking for gradients
    return eval_point(e, x, nullptr, TPL_GRAD_ANALYTIC, true, e && e->spec_grad);
}

int tpl_calc_gradient(tpl_engine* e, const double* x, double* g) {
    if (!e || !g) return fail(TPL_ERR_ARG, "null");
    if (!e->last_x_valid) return fail(TPL_ERR_STATE, "calc_gradient needs a preceding calc_pl (it differentiates at the engine state)");
    (void)x;
    e->spec_grad = true;
    e->pl_since_grad = 0;
    if (e->g_cached && !e->topo_dirty && !e->slots_dirty && !e->state_dirty && e->g_lambda == e->smoothing && e->g_pb == e->pb &&
        e->g_logpen == e->log_pen) {
        // the preceding calc_pl already computed it (same launch, same kernels as the path below)
        const int P = e->numparams;
        memcpy(g, e->h_G.p, sizeof(double) * P);
        if (e->lf_mode && std::isnan(g[0]))
            for (int i = 0; i < P; i++) g[i] = 1000000;               // calc_lf_gradient :680-687
        return TPL_OK;
    }
    double f = eval_point(e, e->h_X.p, g, TPL_GRAD_ANALYTIC, true);
    return std::isnan(f) && !g_err.empty() ? TPL_ERR_CUDA : TPL_OK;
}

double tpl_calc_function_gradient(tpl_engine* e, const double* x, double* g) {
    return eval_point(e, x, g, TPL_GRAD_AD, false);
}

double tpl_calc_pl_grad(tpl_engine* e, const double* x, double* g, int mode) {
    return eval_point(e, x, g, mode, mode == TPL_GRAD_ANALYTIC);
}

int tpl_get_rates(tpl_engine* e, double* out) {
    if (!e || !out) return fail(TPL_ERR_ARG, "null");
    materialize_state(e);
    memcpy(out, e->rates.data(), sizeof(double) * e->n);
    return TPL_OK;
}
int tpl_get_dates(tpl_engine* e, double* out) {
    if (!e || !out) return fail(TPL_ERR_ARG, "null");
    materialize_state(e);
    memcpy(out, e->dates.data(), sizeof(double) * e->n);
    return TPL_OK;
}
int tpl_get_durations(tpl_engine* e, double* out) {
    if (!e || !out) return fail(TPL_ERR_ARG, "null");
    materialize_state(e);
    memcpy(out, e->durations.data(), sizeof(double) * e->n);
    return TPL_OK;
}

int tpl_batch_configure(tpl_engine* e, int B, const double* lane_smoothing, const int* mask_off, const int* mask_nodes) {
    if (!e || B < 1) return fail(TPL_ERR_ARG, "null or B < 1");
    // validate everything first: a failed call leaves the previous configuration intact
    if (mask_off) {
        if (mask_off[0] != 0) return fail(TPL_ERR_ARG, "mask_off[0] != 0");
        for (int b = 0; b < B; b++)
            if (mask_off[b + 1] < mask_off[b]) return fail(TPL_ERR_ARG, "mask_off is not non-decreasing");
        if (!mask_nodes && mask_off[B] > 0) return fail(TPL_ERR_ARG, "mask_nodes is null");
        for (int k = 0; k < mask_off[B]; k++)
            if (mask_nodes[k] < 0 || mask_nodes[k] >= e->n) return fail(TPL_ERR_ARG, "mask node out of range");
    }
    CK(cudaSetDevice(e->device));
    if (lane_smoothing) {
        std::vector<double> v(lane_smoothing, lane_smoothing + B);
        CK(e->d_lane_lambda.upload(v, e->stream));
        CK(cudaStreamSynchronize(e->stream));
    }
    int nch = e->cfg_nchunks;
    if (mask_off) {
        nch = (B + 63) / 64;
        const size_t npad = (size_t)((e->n + tpl::TILE_NODES - 1) / tpl::TILE_NODES) * tpl::TILE_NODES;
        std::vector<unsigned long long> bits(npad * nch, 0ull);
        for (int b = 0; b < B; b++)
            for (int k = mask_off[b]; k < mask_off[b + 1]; k++)
                bits[(size_t)(b >> 6) * npad + mask_nodes[k]] |= 1ull << (b & 63);
        CK(e->d_lanemask.upload(bits, e->stream));
        CK(cudaStreamSynchronize(e->stream));
    }
    e->cfg_B = B;
    e->cfg_lambda = lane_smoothing != nullptr;
    e->cfg_mask = mask_off != nullptr;
    e->cfg_nchunks = nch;
    return TPL_OK;
}

int tpl_calc_pl_grad_batch_dev(tpl_engine* e, int B, const double* dX, double* dF, double* dG, int mode, void* stream) {
    if (!e || !dX || !dF) return fail(TPL_ERR_ARG, "null");
    CK(cudaSetDevice(e->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : e->stream;
    return eval_device(e, B, dX, dF, dG, mode, true, s);
}

// Would eval_device pick the descriptor kernel (generation 4) for this batch?  (same conditions as there)
static bool batch_takes_desc_kernel(const tpl_engine* e, int B, int mode) {
    if (e->force_gen != 0 && e->force_gen != 4) return false;
    const bool lanemask = e->cfg_mask && e->cfg_B == B;
    const int maskmode = lanemask ? 2 : (e->cv_any ? 1 : 0);
    const bool tile_legal = e->gen2_ok && !e->lf_mode && maskmode != 1 && (B % 2 == 0) && (double)e->numparams * (double)B < 2.0e9;
    const int nchunks2 = (B + tpl::TILE_LANES - 1) / tpl::TILE_LANES;
    const long long W = (long long)nchunks2 * e->ntiles2;
    const int G = (int)std::min<long long>(e->num_sms, W);
    const bool stages = tpl::desc_kernel_smem(e->tile_hrows, 2) <= e->max_smem;
    return tile_legal && B >= 32 && e->gen4_ok && nchunks2 <= G && stages && !(e->log_pen && mode == TPL_GRAD_ANALYTIC);
}

// The device's address of a PINNED (cudaMallocHost / cudaHostRegister / tpl_host_alloc) host buffer, or null for pageable memory
static void* mapped_host_pointer(const void* p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return at.type == cudaMemoryTypeHost ? at.devicePointer : nullptr;
}

int tpl_calc_pl_grad_batch(tpl_engine* e, int B, const double* X, double* F, double* G, int mode) {
    if (!e || !X || !F || B < 1) return fail(TPL_ERR_ARG, "null or B < 1");
    if (!e->have_fp) return fail(TPL_ERR_STATE, "set_freeparams has not been called");
    CK(cudaSetDevice(e->device));
    const int P = e->numparams;
    // Zero-copy for batches too small to pipeline (B < 128): when X and G are pinned host buffers and the batch takes the descriptor
    // kernel - which reads every X row once (TMA) and writes every G row once - the kernels work ON the host buffers through their
    // device mapping, so reads and writes cross PCIe at the same time instead of one after the other: 12.4 k against 9.5 k evals/s at
    // B = 64 on the 100k-tip tree.  From 128 vectors on the staged pipeline below is faster (B = 256: 16.5 k against 15.3 k; the
    // kernel's reads of host memory - 512-byte row segments, halo rows a second time - reach ~40 GB/s, the copy engine 47-55).
    // Results are bitwise those of the device-buffer entry.  tpl_debug_force_kernel(e, 5) switches this path off (tests).
    if (G && B >= 32 && B < 128 && batch_takes_desc_kernel(e, B, mode)) {
        double* mX = static_cast<double*>(mapped_host_pointer(X));
        double* mG = static_cast<double*>(mapped_host_pointer(G));
        if (mX && mG) {
            CK(e->d_F.ensure(B));
            cudaStream_t s = e->stream;
            int rc = eval_device(e, B, mX, e->d_F.p, mG, mode, true, s);
            if (rc != TPL_OK) return rc;
            CK(cudaMemcpyAsync(F, e->d_F.p, B * sizeof(double), cudaMemcpyDeviceToHost, s));
            CK(cudaStreamSynchronize(s));
            if (e->static_bad) for (int b = 0; b < B; b++) F[b] = TPL_LARGE;
            return TPL_OK;
        }
    }
    if (B >= 128 && B % 2 == 0 && !e->lf_mode) {
        // Pipelined: 64-vector sub-batches on three streams, so that the upload of sub-batch k+1, the kernels
        // of sub-batch k and the download of sub-batch k-1 overlap (PCIe is full duplex; a single H2D -> kernel
        // -> D2H sequence leaves each direction idle half of the time).  X and G are [P][B] on the host, so a
        // sub-batch is a 2-D copy: P rows of 512 bytes with pitch 8 B.
        // 64-vector sub-batches (512-byte rows of the 2-D copies) measured best: 32 shortens the pipeline's fill and
        // drain but its 256-byte rows copy slower (14.4 k vs 16.5 k evals/s), 128 lengthens fill and drain (14.2 k).
        // Per-vector CV masks are addressed in 64-vector words, so masked batches never go below 64.
        // (Half-size first and last pieces, to shorten the fill and drain, lose for the same reason: B = 256 15.8 k against 16.7 k.)
        // A ragged tail shorter than 32 vectors would run the generic kernel: the last two pieces share the remainder instead.
        const int SB = 64;
        std::vector<std::pair<int, int>> pieces;       // (first vector, vectors)
        {
            int off = 0;
            while (B - off > 2 * SB) { pieces.emplace_back(off, SB); off += SB; }
            const int rest = B - off;                  // 2 ... 128
            const bool masked = e->cfg_mask && e->cfg_B == B;          // mask words cover 64 vectors: pieces start at multiples of 64
            if (rest > SB && rest < SB + SB / 2 && !masked && e->force_gen != 5) {
                const int first = ((rest / 2) + 1) & ~1;
                pieces.emplace_back(off, first);
                pieces.emplace_back(off + first, rest - first);
            } else {
                pieces.emplace_back(off, std::min(SB, rest));
                if (rest > SB) pieces.emplace_back(off + SB, rest - SB);
            }
        }
        const int nsub = (int)pieces.size();
        for (int k = 0; k < 3; k++) {
            if (!e->pstream[k]) CK(cudaStreamCreateWithFlags(&e->pstream[k], cudaStreamNonBlocking));
            CK(e->p_X[k].ensure((size_t)P * SB));
            CK(e->p_F[k].ensure(SB));
            if (G) CK(e->p_G[k].ensure((size_t)P * SB));
        }
        CK(cudaStreamSynchronize(e->stream));          // pending tree uploads of the engine stream
        int rc0 = sync_device_tree(e);
        if (rc0 != TPL_OK) return rc0;
        CK(cudaStreamSynchronize(e->stream));
        for (int sidx = 0; sidx < nsub; sidx++) {
            const int k = sidx % 3;
            const int off = pieces[sidx].first;
            const int nl = pieces[sidx].second;
            cudaStream_t st = e->pstream[k];
            CK(cudaMemcpy2DAsync(e->p_X[k].p, (size_t)nl * 8, X + off, (size_t)B * 8, (size_t)nl * 8, P,
                                 cudaMemcpyHostToDevice, st));
            int rc = eval_device(e, nl, e->p_X[k].p, e->p_F[k].p, G ? e->p_G[k].p : nullptr, mode, true, st, off, B, 1 + k);
            if (rc != TPL_OK) return rc;
            CK(cudaMemcpyAsync(F + off, e->p_F[k].p, (size_t)nl * 8, cudaMemcpyDeviceToHost, st));
            if (G)
                CK(cudaMemcpy2DAsync(G + off, (size_t)B * 8, e->p_G[k].p, (size_t)nl * 8, (size_t)nl * 8, P,
                                     cudaMemcpyDeviceToHost, st));
        }
        for (int k = 0; k < 3; k++) CK(cudaStreamSynchronize(e->pstream[k]));
        e->last_launches = 2 * nsub;
        if (e->static_bad) for (int b = 0; b < B; b++) F[b] = TPL_LARGE;
        return TPL_OK;
    }
    const size_t PB = (size_t)P * B;
    CK(e->d_X.ensure(PB));
    CK(e->d_F.ensure(B));
    if (G) CK(e->d_G.ensure(PB));
    cudaStream_t s = e->stream;
    CK(cudaMemcpyAsync(e->d_X.p, X, PB * sizeof(double), cudaMemcpyHostToDevice, s));
    int rc = eval_device(e, B, e->d_X.p, e->d_F.p, G ? e->d_G.p : nullptr, mode, true, s);
    if (rc != TPL_OK) return rc;
    CK(cudaMemcpyAsync(F, e->d_F.p, B * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (G) CK(cudaMemcpyAsync(G, e->d_G.p, PB * sizeof(double), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (e->static_bad) for (int b = 0; b < B; b++) F[b] = TPL_LARGE;
    return TPL_OK;
}

int tpl_debug_force_kernel(tpl_engine* e, int gen) {
    if (!e || gen < 0 || gen > 5) return fail(TPL_ERR_ARG, "bad argument");
    e->force_gen = gen;
    return TPL_OK;
}
int tpl_last_kernel_generation(const tpl_engine* e) { return e ? e->last_gen : TPL_ERR_ARG; }

int tpl_set_timing(tpl_engine* e, int on) {
    if (!e) return fail(TPL_ERR_ARG, "null");
    CK(cudaSetDevice(e->device));
    e->timing = on != 0;
    e->ev_used = 0;
    if (e->timing && e->ev.empty()) {
        e->ev.resize(3 * 4096);
        for (auto& v : e->ev) CK(cudaEventCreate(&v));
    }
    return TPL_OK;
}

int tpl_get_timing(tpl_engine* e, double* main_ms, double* finalize_ms) {
    if (!e || !main_ms || !finalize_ms) return fail(TPL_ERR_ARG, "null");
    CK(cudaSetDevice(e->device));
    double m = 0.0, f = 0.0;
    int cnt = 0;
    for (size_t k = 0; k + 3 <= e->ev_used; k += 3) {
        CK(cudaEventSynchronize(e->ev[k + 2]));
        float a = 0.f, b = 0.f;
        CK(cudaEventElapsedTime(&a, e->ev[k], e->ev[k + 1]));
        CK(cudaEventElapsedTime(&b, e->ev[k + 1], e->ev[k + 2]));
        m += a; f += b; cnt++;
    }
    *main_ms = m;
    *finalize_ms = f;
    e->ev_used = 0;
    return cnt;
}

int tpl_debug_math(int n, const double* x, double* log_out, double* rcp_out) {
    if (n < 1 || !x || !log_out || !rcp_out) return fail(TPL_ERR_ARG, "null");
    double *dx = nullptr, *dl = nullptr, *dr = nullptr;
    double2* dt = nullptr;
    std::vector<double2> tab;
    build_log_table(tab);
    CK(cudaMalloc((void**)&dx, n * sizeof(double)));
    CK(cudaMalloc((void**)&dl, n * sizeof(double)));
    CK(cudaMalloc((void**)&dr, n * sizeof(double)));
    CK(cudaMalloc((void**)&dt, tab.size() * sizeof(double2)));
    CK(cudaMemcpy(dx, x, n * sizeof(double), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dt, tab.data(), tab.size() * sizeof(double2), cudaMemcpyHostToDevice));
    tpl::debug_math_kernel<<<(n + 255) / 256, 256>>>(n, dx, dl, dr, dt);
    CK(cudaGetLastError());
    CK(cudaMemcpy(log_out, dl, n * sizeof(double), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(rcp_out, dr, n * sizeof(double), cudaMemcpyDeviceToHost));
    cudaFree(dx); cudaFree(dl); cudaFree(dr); cudaFree(dt);
    return TPL_OK;
}

void* tpl_host_alloc(unsigned long long bytes) {
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes) != cudaSuccess) { fail(TPL_ERR_CUDA, "cudaMallocHost failed"); return nullptr; }
    return p;
}
void tpl_host_free(void* p) { if (p) cudaFreeHost(p); }

}  // extern "C"
