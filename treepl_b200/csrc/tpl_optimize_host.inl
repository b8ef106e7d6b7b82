// Host side of the device-resident batched optimisers (tpl_optimize_batch): work-space set-up, kernel sequences,
// CUDA-graph replay, result read-back.  Part of the tpl_engine.cu translation unit (it uses the engine's internals:
// eval_device, sync_device_tree, DevBuf, fail / CK); included there INSIDE the anonymous namespace, which this file closes before its extern "C" entry points.
// Kernels: tpl_lbfgs.cuh (method 0), tpl_tn.cuh (method 1).
#pragma once
// start / stop events of one optimisation, released on every exit path
struct EventPair {
    cudaEvent_t a = nullptr, b = nullptr;
    cudaError_t create() {
        cudaError_t e = cudaEventCreate(&a);
        return e != cudaSuccess ? e : cudaEventCreate(&b);
    }
    ~EventPair() {
        if (a) cudaEventDestroy(a);
        if (b) cudaEventDestroy(b);
    }
};

// the replayed step graph and the engine's timing flag (no event records inside a capture): released / restored on
// every exit path, CK early returns included
struct GraphGuard {
    tpl_engine* e;
    bool was_timing;
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t gexec = nullptr;
    explicit GraphGuard(tpl_engine* eng) : e(eng), was_timing(eng->timing) { e->timing = false; }
    ~GraphGuard() {
        if (gexec) cudaGraphExecDestroy(gexec);
        if (graph) cudaGraphDestroy(graph);
        e->timing = was_timing;
    }
};

// ---- device-resident batched L-BFGS (tpl_lbfgs_core.h / tpl_lbfgs.cuh) --------------------------------------

template <int HM>
int run_optimizer(tpl_engine* e, tplopt::OptArgs a, int B, int mode, int max_steps, int check_every, bool use_graph,
                  tpl_opt_result* res) {
    using namespace tplopt;
    typedef Layout<HM> L;
    cudaStream_t s = e->stream;
    const dim3 blk(32, OPT_WY);
    const dim3 grd((B + 31) / 32, a.nchunks);
    const dim3 lblk(OPT_LX, OPT_NY);
    const dim3 lgrd((B + OPT_LX - 1) / OPT_LX);
    const size_t lsmem = (size_t)(L::NACC + L::NB * L::NB) * OPT_LX * sizeof(double);
    CK(cudaFuncSetAttribute(opt_lane_kernel<HM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lsmem));
    int launches = 0;
    const dim3 tgrd((B + 31) / 32, a.tchunks);
    const dim3 fgrd((B + 31) / 32, a.light_chunks);
    opt_init_kernel<<<fgrd, blk, 0, s>>>(a);
    opt_trial_kernel<<<fgrd, blk, 0, s>>>(a);
    launches += 2;
    CK(cudaGetLastError());
    auto one_step = [&]() -> int {
        int rc = eval_device(e, B, a.X, const_cast<double*>(a.Ft), const_cast<double*>(a.Gx), mode, true, s);
        if (rc != TPL_OK) return rc;
        opt_update_kernel<HM><<<grd, dim3(32, HM), 0, s>>>(a);
        if (a.lane_part == a.part_red) opt_reduce_kernel<HM><<<dim3((B + 31) / 32, L::NACC), blk, 0, s>>>(a);
        opt_lane_kernel<HM><<<lgrd, lblk, lsmem, s>>>(a);
        opt_direction_kernel<HM><<<dim3((B + 31) / 32, a.dir_chunks), blk, 0, s>>>(a);
        if (a.boundary_frac > 0.0) opt_stepmax_kernel<<<tgrd, blk, 0, s>>>(a);
        opt_trial_kernel<<<fgrd, blk, 0, s>>>(a);
        CK(cudaGetLastError());
        return TPL_OK;
    };
    const int per_step = (a.boundary_frac > 0.0 ? 7 : 6) + (a.lane_part == a.part_red ? 1 : 0);
    int steps = 0;
    int rc = one_step();               // first step outside any capture: sizes the evaluation work space
    if (rc != TPL_OK) return rc;
    steps++;
    launches += per_step;
    GraphGuard gg(e);
    if (use_graph) {
        CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
        for (int k = 0; k < check_every && rc == TPL_OK; k++) rc = one_step();
        cudaError_t ce = cudaStreamEndCapture(s, &gg.graph);
        if (rc != TPL_OK || ce != cudaSuccess)
            return rc != TPL_OK ? rc : fail(TPL_ERR_CUDA, std::string("graph capture failed: ") + cudaGetErrorString(ce));
        CK(cudaGraphInstantiate(&gg.gexec, gg.graph, 0));
    }
    int ndone = 0;
    while (steps < max_steps) {
        // the last chunk is clamped to max_steps (plain launches; a graph replays whole chunks only)
        const int chunk = std::min(check_every, max_steps - steps);
        if (use_graph && chunk == check_every) {
            CK(cudaGraphLaunch(gg.gexec, s));
        } else {
            for (int k = 0; k < chunk && rc == TPL_OK; k++) rc = one_step();
            if (rc != TPL_OK) return rc;
        }
        steps += chunk;
        launches += per_step * chunk;
        CK(cudaMemcpyAsync(e->opt.h_done.p, a.ndone, sizeof(int), cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        ndone = e->opt.h_done.p[0];
        if (ndone >= B) break;
    }
    opt_final_kernel<<<fgrd, blk, 0, s>>>(a, e->opt.Fout.p);
    launches++;
    CK(cudaGetLastError());
    res->steps = steps;
    res->n_done = ndone;
    res->launches = launches;
    return TPL_OK;
}

// ---- batched truncated Newton (tpl_tn_core.h / tpl_tn.cuh) ---------------------------------------------------------

// tpl_optimize_batch_cols: the B start vectors are columns of a small host matrix and only a few rows per vector return
struct ColsIO {
    int K;                       // columns of Xstart
    const double* Xstart;        // host [P][K]
    const int* start_col;        // host [B]
    int nprobe;
    const int* probe_rows;       // host [nprobe][B], -1 = none
    double* probe_out;           // host [nprobe][B]
    double* Xout;                // host [P][B] or null
};

__global__ void cols_expand_kernel(int P, int B, int K, const double* __restrict__ Xs, const int* __restrict__ col, double* __restrict__ X) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int c = col[b];
    for (int p = blockIdx.y; p < P; p += gridDim.y) X[(size_t)p * B + b] = Xs[(size_t)p * K + c];
}

__global__ void cols_probe_kernel(int B, int nprobe, const double* __restrict__ X, const int* __restrict__ rows, double* __restrict__ out) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    for (int j = blockIdx.y; j < nprobe; j += gridDim.y) {
        const int r = rows[(size_t)j * B + b];
        out[(size_t)j * B + b] = r >= 0 ? X[(size_t)r * B + b] : 0.0;
    }
}

// index record of node i for the kernels' record views (tpl_tn.cuh::TreeRec)
static void tn_fill_record(const tpl_engine* e, int i, int* r) {
    const int n = e->n;
    const int pa = i == 0 ? -1 : e->parent[i];
    const int c0 = e->child_off[i], nc = e->child_off[i + 1] - c0;
    r[0] = i; r[1] = e->freeparams[i]; r[2] = e->freeparams[n + i]; r[3] = pa;
    r[4] = pa >= 0 ? e->freeparams[pa] : -1; r[5] = pa >= 0 ? e->freeparams[n + pa] : -1; r[6] = c0; r[7] = nc;
    for (int j = 0; j < 2 && j < nc; j++) {
        const int ch = e->children[c0 + j];
        r[8 + j] = ch; r[10 + j] = e->freeparams[n + ch];
    }
}

// work space, tree operands and chunking of the truncated-Newton kernels; uploads lo / hi / sc and X
// tn_hv_pipe_kernel needs 105 KB of dynamic shared memory per block: opt in once per DEVICE (function attributes belong to the
// device's context; a process may hold engines on several GPUs)
bool tn_hv_pipe_ready() {
    static std::mutex mu;
    static int ok[64];                              // 0 unknown, 1 ready, -1 not available
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) { cudaGetLastError(); return false; }
    std::lock_guard<std::mutex> lk(mu);
    if (ok[dev] == 0) {
        const bool set = cudaFuncSetAttribute(tpltn::tn_hv_pipe_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tpltn::HVP_SMEM) == cudaSuccess &&
                         cudaFuncSetAttribute(tpltn::tn_hv_pipe_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tpltn::HVP_SMEM) == cudaSuccess;
        if (!set) cudaGetLastError();
        ok[dev] = set ? 1 : -1;
    }
    return ok[dev] == 1;
}
// Hp = H p: the software-pipelined kernel for full 32-vector groups, the register-staged one for packed lanes (B <= 16)
void tn_launch_hv(const tpltn::TnArgs& a, dim3 ngrd, dim3 blk, cudaStream_t s) {
    if (a.hv_pipe && a.lanemask) tpltn::tn_hv_pipe_kernel<true><<<ngrd, blk, tpltn::HVP_SMEM, s>>>(a);
    else if (a.hv_pipe) tpltn::tn_hv_pipe_kernel<false><<<ngrd, blk, tpltn::HVP_SMEM, s>>>(a);
    else tpltn::tn_hv_kernel<<<ngrd, blk, 0, s>>>(a);
}

int tn_prepare(tpl_engine* e, int B, const double* X, const std::vector<double>& lo, const std::vector<double>& hi,
               const std::vector<double>& sc, int nr, const tpl_opt_options& o, tpltn::TnArgs& a, const ColsIO* io = nullptr) {
    using namespace tpltn;
    if (e->lf_mode) return fail(TPL_ERR_ARG, "truncated Newton needs the PL layout (one rate per branch)");
    if ((double)e->numparams * (double)B >= 4.0e9) return fail(TPL_ERR_ARG, "truncated Newton: P * B must stay below 2^32 elements (32-bit element offsets in the node kernels)");
    if (e->cv_any) return fail(TPL_ERR_ARG, "truncated Newton takes CV masks per vector (tpl_batch_configure), not set_cv_nodes");
    const int P = e->numparams, n = e->n;
    const size_t PB = (size_t)P * B;
    tpl_engine::OptWork& w = e->opt;
    cudaStream_t s = e->stream;
    const int groups = (B + 31) / 32;
    const int WY = tplopt::OPT_WY;
    auto one_wave = [&](int occ, int total, int align, int* nch, int* per) {
        int c = std::max(1, (e->num_sms * std::max(occ, 1)) / groups);
        c = std::min(c, (total + align - 1) / align);
        int r = (total + c - 1) / c;
        r = ((r + align - 1) / align) * align;
        *per = r;
        *nch = (total + r - 1) / r;
    };
    int occ_row = 1, occ_node = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_row, tn_update_kernel, 32 * WY, 0);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_node, tn_hv_kernel, 32 * WY, 0);
    one_wave(occ_row, P, WY, &a.row_chunks, &a.rows_per_chunk);
    one_wave(occ_node, n, WY, &a.node_chunks, &a.nodes_per_chunk);
    int occ_diag = 1, occ_tm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_diag, tn_diag_kernel, 32 * WY, 0);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_tm, tn_stepmax_kernel, 32 * WY, 0);
    one_wave(occ_diag, n, WY, &a.diag_chunks, &a.diag_nodes_per_chunk);
    one_wave(occ_tm, n, WY, &a.tm_chunks, &a.tm_nodes_per_chunk);
    CK(w.Z.ensure(PB)); CK(w.Gz.ensure(PB)); CK(w.D.ensure(PB)); CK(w.X.ensure(PB)); CK(w.Gx.ensure(PB));
    CK(w.R.ensure(PB)); CK(w.ZZ.ensure(PB)); CK(w.Pv.ensure(PB)); CK(w.Hp.ensure(PB)); CK(w.Minv.ensure(PB)); CK(w.DG.ensure(PB));
    CK(w.lo.ensure(P)); CK(w.hi.ensure(P)); CK(w.sc.ensure(P)); CK(w.Ft.ensure(B)); CK(w.Fout.ensure(B)); CK(w.h_done.ensure(1));
    const size_t npart = (size_t)(3 + 1 + 1 + 1) * a.row_chunks * B + (size_t)2 * a.node_chunks * B + (size_t)a.tm_chunks * B;
    CK(w.tn_part.ensure(npart));
    CK(w.tn_lane_d.ensure((size_t)14 * B));
    CK(w.tn_lane_i.ensure((size_t)8 * B + 1 + 6 * (size_t)groups));
    CK(w.pmin_n.upload(e->pen_min, s));
    CK(w.pmax_n.upload(e->pen_max, s));
    CK(cudaMemcpyAsync(w.lo.p, lo.data(), sizeof(double) * P, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(w.hi.p, hi.data(), sizeof(double) * P, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(w.sc.p, sc.data(), sizeof(double) * P, cudaMemcpyHostToDevice, s));
    if (io) {
        // start vectors = columns of a [P][K] matrix: K columns cross the bus, the batch is expanded on the device
        // (w.Hp as the staging buffer: it is zeroed below, after the expansion)
        if (io->K < 1 || (size_t)io->K > (size_t)B) return fail(TPL_ERR_ARG, "start columns: need 1 <= K <= B");
        for (int b = 0; b < B; b++)
            if (io->start_col[b] < 0 || io->start_col[b] >= io->K) return fail(TPL_ERR_ARG, "start_col out of range");
        CK(w.tn_cols.ensure((size_t)B));
        CK(cudaMemcpyAsync(w.Hp.p, io->Xstart, sizeof(double) * (size_t)P * io->K, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(w.tn_cols.p, io->start_col, sizeof(int) * B, cudaMemcpyHostToDevice, s));
        cols_expand_kernel<<<dim3((B + 127) / 128, std::min(P, 1024)), 128, 0, s>>>(P, B, io->K, w.Hp.p, w.tn_cols.p, w.X.p);
        CK(cudaGetLastError());
    } else {
        CK(cudaMemcpyAsync(w.X.p, X, sizeof(double) * PB, cudaMemcpyHostToDevice, s));
    }
    CK(cudaMemsetAsync(w.tn_lane_d.p, 0, sizeof(double) * 14 * B, s));
    CK(cudaMemsetAsync(w.tn_lane_i.p, 0, sizeof(int) * (8 * (size_t)B + 1 + 6 * (size_t)groups), s));     // st = TS_INIT, t = 0
    for (DevBuf<double>* buf : {&w.Gz, &w.D, &w.R, &w.ZZ, &w.Pv, &w.Hp, &w.Minv, &w.DG, &w.Gx})
        CK(cudaMemsetAsync(buf->p, 0, sizeof(double) * PB, s));
    CK(cudaStreamSynchronize(s));       // pen_min / pen_max uploads read host vectors of the engine; lo / hi / sc are the caller's
    a.o.P = P; a.o.B = B; a.o.nr = nr;
    a.pack_shift = 5;                               // B <= 16: sub-warps of the next power of two >= B lanes (tpl_tn.cuh::node_lane)
    if (B <= 16) { a.pack_shift = 0; while ((1 << a.pack_shift) < B) a.pack_shift++; }
    a.o.max_ls = o.max_ls > 0 ? o.max_ls : 40;
    a.o.max_cg = o.max_cg > 0 ? o.max_cg : 500;     // a cap of 100 truncates stiff solves (large smoothing x large rates) into crawling
    a.o.patience = o.patience > 0 ? o.patience : 3;
    a.o.c1 = o.armijo > 0 ? o.armijo : 1e-4;
    a.o.ftol = o.ftol > 0 ? o.ftol : 1e-13;
    a.o.gtol = o.gtol > 0 ? o.gtol : 1e-8;
    a.o.boundary_frac = o.boundary_frac < 0 ? 0.0 : (o.boundary_frac > 0 ? std::min(o.boundary_frac, 0.999999) : 0.9);
    a.o.fnoise = 1e-14 * e->sum_c;
    a.t = dev_tree(e);
    a.pmin_n = w.pmin_n.p; a.pmax_n = w.pmax_n.p;
    const bool lanemask = e->cfg_mask && e->cfg_B == B;
    a.lanemask = lanemask ? e->d_lanemask.p : nullptr;
    a.mask_stride = e->ntiles2 * tpl::TILE_NODES;
    a.lane_lambda = (e->cfg_lambda && e->cfg_B == B) ? e->d_lane_lambda.p : nullptr;
    a.lambda = e->smoothing;
    a.pb = e->pb;
    a.sc = w.sc.p; a.lo = w.lo.p; a.hi = w.hi.p;
    a.Z = w.Z.p; a.Gz = w.Gz.p; a.S = w.D.p; a.X = w.X.p; a.R = w.R.p; a.ZZ = w.ZZ.p; a.Pv = w.Pv.p; a.Hp = w.Hp.p;
    a.Minv = w.Minv.p; a.DG = w.DG.p; a.Gx = w.Gx.p; a.Ft = w.Ft.p;
    int* li = w.tn_lane_i.p;
    const size_t Bs = (size_t)B;
    a.st = li; a.ls = li + Bs; a.cgit = li + 2 * Bs; a.flag = li + 3 * Bs; a.stall = li + 4 * Bs; a.iters = li + 5 * Bs;
    a.ncg = li + 6 * Bs; a.act = li + 7 * Bs; a.ndone = li + 8 * Bs;
    int* cnt = li + 8 * Bs + 1;
    a.cnt_setup = cnt; a.cnt_hv = cnt + groups; a.cnt_upd = cnt + 2 * groups; a.cnt_gd = cnt + 3 * groups; a.cnt_tm = cnt + 4 * groups;
    a.cnt_pc = cnt + 5 * groups;
    double* ld = w.tn_lane_d.p;
    a.F0 = ld; a.t_ = ld + Bs; a.gd = ld + 2 * Bs; a.rz = ld + 3 * Bs; a.rz0 = ld + 4 * Bs; a.tol2 = ld + 5 * Bs; a.alpha = ld + 6 * Bs;
    a.beta = ld + 7 * Bs; a.gmax = ld + 8 * Bs; a.tlog = ld + 9 * Bs; a.gdtmp = ld + 10 * Bs;
    a.gmax_tmp = ld + 11 * Bs; a.gg_tmp = ld + 12 * Bs; a.mu = ld + 13 * Bs;
    double* pp = w.tn_part.p;
    a.part_setup = pp; pp += (size_t)3 * a.row_chunks * B;
    a.part_upd = pp; pp += (size_t)a.row_chunks * B;
    a.part_gd = pp; pp += (size_t)a.row_chunks * B;
    a.part_hv = pp; pp += (size_t)2 * a.node_chunks * B;
    a.part_tm = pp; pp += (size_t)a.tm_chunks * B;
    a.part_pc = pp;
    {   // index records (node order) of the Hessian-vector kernel's fast path: every non-root node whose parent is not the root,
        // with no or exactly two children (rates exist for all of them in the PL layout)
        std::vector<int> hrec((size_t)n * tpltn::HV_REC_INTS, -1);
        for (int i = 0; i < n; i++) {
            int* r = hrec.data() + (size_t)i * tpltn::HV_REC_INTS;
            const int pa = i == 0 ? -1 : e->parent[i];
            const int c0 = e->child_off[i], nc = e->child_off[i + 1] - c0;
            r[0] = i; r[1] = e->freeparams[i]; r[2] = e->freeparams[n + i]; r[3] = pa;
            r[4] = pa >= 0 ? e->freeparams[pa] : -1; r[5] = pa >= 0 ? e->freeparams[n + pa] : -1; r[6] = nc;
            // two shapes only: a tip without a date row, or an internal node with two children and a date row
            bool fast = i != 0 && pa != 0 && ((nc == 0 && r[2] < 0) || (nc == 2 && r[2] >= 0)) && r[1] >= 0 && r[4] >= 0;
            for (int j = 0; j < 2 && j < nc; j++) {
                const int ch = e->children[c0 + j];
                r[8 + j] = ch; r[10 + j] = e->freeparams[n + ch]; r[12 + j] = e->freeparams[ch];
                if (e->freeparams[ch] < 0) fast = false;
            }
            r[7] = fast ? 1 : 0;
        }
        CK(w.tn_hv_rec.upload(hrec, s));
        // constant record of the same nodes for tn_hv_pipe_kernel (one 128-byte fetch instead of ~13 dependent broadcast loads):
        // scales of the eight rows (a fixed date in the place of the scale of a date row that does not exist), c of the node and
        // of its children, calibration penalties
        a.hv_const = nullptr;
        a.hv_pipe = 0;
        if (o.hv_pipe >= 0 && a.pack_shift == 5 && tn_hv_pipe_ready()) {
            std::vector<double> hc((size_t)n * tpltn::HV_CONST, 0.0);
            for (int i = 0; i < n; i++) {
                const int* r = hrec.data() + (size_t)i * tpltn::HV_REC_INTS;
                if (!r[7]) continue;
                double* q = hc.data() + (size_t)i * tpltn::HV_CONST;
                const int pa = r[3];
                q[0] = sc[r[1]];
                q[1] = r[2] >= 0 ? sc[r[2]] : e->dates[i];
                q[2] = sc[r[4]];
                q[3] = r[5] >= 0 ? sc[r[5]] : e->dates[pa];
                for (int j = 0; j < 2 && j < r[6]; j++) {
                    q[4 + 2 * j] = sc[r[12 + j]];
                    q[5 + 2 * j] = r[10 + j] >= 0 ? sc[r[10 + j]] : e->dates[r[8 + j]];
                    q[9 + j] = e->c[r[8 + j]];
                }
                q[8] = e->c[i];
                q[11] = e->pen_min[i];
                q[12] = e->pen_max[i];
            }
            CK(w.tn_hv_const.upload(hc, s));
            a.hv_const = w.tn_hv_const.p;
            a.hv_pipe = 1;
        }
        CK(cudaStreamSynchronize(s));
        a.hv_rec = w.tn_hv_rec.p;
    }
    // ---- tree preconditioner: height levels (children before parents), node lists per level, work planes ----
    a.precond = o.precond == 1 ? 0 : 1;
    a.log_pen = e->log_pen ? 1 : 0;
    a.TW = nullptr; a.lev_nodes = nullptr; a.lev_off = nullptr; a.lev_rec = nullptr; a.nlev = a.nlev_wide = 0;
    if (a.precond) {
        std::vector<int> height(n, 0);
        int hmax = 0;
        for (int i = n - 1; i >= 0; i--) {              // preorder numbering: a child's number exceeds its parent's
            int h = 0;
            for (int j = e->child_off[i]; j < e->child_off[i + 1]; j++) h = std::max(h, height[e->children[j]] + 1);
            height[i] = h;
            hmax = std::max(hmax, h);
        }
        std::vector<int> off(hmax + 2, 0), nodes(n);
        for (int i = 0; i < n; i++) off[height[i] + 1]++;
        for (int l = 0; l <= hmax; l++) off[l + 1] += off[l];
        std::vector<int> fill(off.begin(), off.end() - 1);
        for (int i = 0; i < n; i++) nodes[fill[height[i]]++] = i;
        a.nlev = hmax + 1;
        a.nlev_wide = 0;
        // a level is worth its own launch when it fills the top kernel's block several times over: 16 warps x (nodes per warp)
        const int wide_min = o.tree_wide_min > 0 ? o.tree_wide_min : 4 * tpltn::TREE_TOP_WY * (32 >> a.pack_shift);
        while (a.nlev_wide < a.nlev && off[a.nlev_wide + 1] - off[a.nlev_wide] >= wide_min) a.nlev_wide++;
        // measured (scratch A/B on one B200): the single launch wins 5-15 % where the sweeps are latency bound
        // (16+ vectors in up to ~6 groups) and loses 1-7 % at 2-5 vectors and at 32+ groups, where a level fills the device anyway
        a.coop_levels = o.tree_coop < 0 ? 0 : (o.tree_coop > 0 ? 1 : (B >= 16 && (B + 31) / 32 <= 6));
        // one index record per node, in level order: what the sweeps' record views load first (tpl_tn.cuh)
        std::vector<int> recs((size_t)n * tpltn::TREE_REC_INTS, -1);
        for (int k = 0; k < n; k++) tn_fill_record(e, nodes[k], recs.data() + (size_t)k * tpltn::TREE_REC_INTS);
        CK(w.tn_lev_nodes.upload(nodes, s));
        CK(w.tn_lev_off.upload(off, s));
        CK(w.tn_lev_rec.upload(recs, s));
        CK(w.tn_tw.ensure((size_t)tpltn::TW_PLANES * n * B));
        CK(cudaStreamSynchronize(s));
        w.tn_lev_off_host = off;
        a.TW = w.tn_tw.p; a.lev_nodes = w.tn_lev_nodes.p; a.lev_off = w.tn_lev_off.p; a.lev_rec = w.tn_lev_rec.p;
    }
    return TPL_OK;
}

// wide levels [l_first, l_last) of one sweep: ONE cooperative launch with grid barriers between the levels when the device
// supports it (and there is more than one level), else one launch per level
template <bool FACTOR, bool DOWN>
int tn_launch_levels(const tpltn::TnArgs& a, const std::vector<int>& loff, int groups, int l_first, int l_last, cudaStream_t s) {
    using namespace tpltn;
    if (l_last <= l_first) return 0;
    const dim3 blk(32, tplopt::OPT_WY);
    // per device (a process may hold engines on several GPUs): blocks of this kernel that fit the device at once, and whether it
    // launches cooperatively at all; 0 = not asked yet
    static int resident_of[64], coop_of[64];
    int dev = 0;
    cudaGetDevice(&dev);
    dev = std::min(std::max(dev, 0), 63);
    if (coop_of[dev] == 0) {
        int sms = 0, occ = 0, sup = 0;
        cudaDeviceGetAttribute(&sup, cudaDevAttrCooperativeLaunch, dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, tn_tree_levels_kernel<FACTOR, DOWN>, 32 * tplopt::OPT_WY, 0);
        resident_of[dev] = occ * sms;
        coop_of[dev] = (sup && resident_of[dev] >= 1) ? 1 : -1;
    }
    const int resident = resident_of[dev];
    int& coop = coop_of[dev];
    if (coop > 0 && a.coop_levels && l_last - l_first >= 2) {
        const int npw = 32 >> a.pack_shift;
        long long most = 1;
        for (int l = l_first; l < l_last; l++) most = std::max(most, (long long)((loff[l + 1] - loff[l] + npw - 1) / npw) * groups);
        const int blocks = (int)std::min<long long>(resident, (most + tplopt::OPT_WY - 1) / tplopt::OPT_WY);
        TnArgs aa = a;
        int lf = l_first, ll = l_last, gr = groups;
        void* args[] = {&aa, &lf, &ll, &gr};
        if (cudaLaunchCooperativeKernel((void*)tn_tree_levels_kernel<FACTOR, DOWN>, dim3(blocks), blk, args, 0, s) == cudaSuccess) return 1;
        cudaGetLastError();
        coop = -1;                                  // not launchable here: per-level launches from now on
    }
    for (int q = 0; q < l_last - l_first; q++) {
        const int l = DOWN ? l_last - 1 - q : l_first + q;
        const dim3 g(groups, (loff[l + 1] - loff[l] + TREE_NODES_PER_BLOCK - 1) / TREE_NODES_PER_BLOCK);
        if (DOWN) tn_tree_down_kernel<FACTOR><<<g, blk, 0, s>>>(a, l);
        else tn_tree_up_kernel<FACTOR><<<g, blk, 0, s>>>(a, l);
    }
    return l_last - l_first;
}

// zz = M^-1 r through the block-tree factorisation: the wide levels of a sweep in one cooperative launch (or one launch each), the
// narrow top in one block per group.  -> number of launches
int tn_launch_tree_solve(const tpltn::TnArgs& a, const std::vector<int>& loff, int groups, cudaStream_t s, bool factor, bool finish) {
    using namespace tpltn;
    const dim3 blk(32, tplopt::OPT_WY), tblk(32, TREE_TOP_WY);
    int nl = 0;
    if (factor) {
        const dim3 lg(groups, (a.t.n + TREE_NODES_PER_BLOCK - 1) / TREE_NODES_PER_BLOCK);
        // record view only for a single vector group: from two groups on the plain accessors (72 registers, three blocks per SM) are
        // as fast or faster since the element offsets are 32-bit (100k tips / 64 vectors -1.8 %, 128 restarts -2.2 %, one group +0.8 %)
        if (groups < 2) tn_tree_local_kernel<true><<<lg, blk, 0, s>>>(a); else tn_tree_local_kernel<false><<<lg, blk, 0, s>>>(a);
        nl++;
    }
    // a factorising sweep finishes level 0 (the tips) in the local kernel
    nl += factor ? tn_launch_levels<true, false>(a, loff, groups, 1, a.nlev_wide, s) : tn_launch_levels<false, false>(a, loff, groups, 0, a.nlev_wide, s);
    if (a.nlev > a.nlev_wide) {
        if (factor) tn_tree_top_kernel<true><<<groups, tblk, 0, s>>>(a); else tn_tree_top_kernel<false><<<groups, tblk, 0, s>>>(a);
        nl++;
    }
    nl += factor ? tn_launch_levels<true, true>(a, loff, groups, 0, a.nlev_wide, s) : tn_launch_levels<false, true>(a, loff, groups, 0, a.nlev_wide, s);
    if (finish) {
        const dim3 rgrd(groups, a.row_chunks);
        if (factor) tn_pc_finish_kernel<true><<<rgrd, blk, 0, s>>>(a); else tn_pc_finish_kernel<false><<<rgrd, blk, 0, s>>>(a);
        nl++;
    }
    return nl;
}

int optimize_tn(tpl_engine* e, int B, double* X, double* F, const std::vector<double>& lo, const std::vector<double>& hi,
                const std::vector<double>& sc, int nr, int mode, const tpl_opt_options& o, int* status, int* iters,
                tpl_opt_result* result, const ColsIO* io = nullptr) {
    using namespace tpltn;
    if (mode != TPL_GRAD_AD) return fail(TPL_ERR_ARG, "truncated Newton differentiates the AD objective (TPL_GRAD_AD)");
    TnArgs a;
    int prc = tn_prepare(e, B, X, lo, hi, sc, nr, o, a, io);
    if (prc != TPL_OK) return prc;
    const int P = e->numparams;
    const size_t PB = (size_t)P * B;
    tpl_engine::OptWork& w = e->opt;
    cudaStream_t s = e->stream;
    const int groups = (B + 31) / 32;
    const int WY = tplopt::OPT_WY;
    const int max_steps = o.max_steps > 0 ? o.max_steps : 20000;
    const int check_every = o.check_every > 0 ? o.check_every : 32;
    // CG iterations per evaluation: most vectors are inside a CG solve most of the time, and only the vectors in a
    // line search need the objective; 4 CG iterations per step make the (then mostly idle) evaluation a ~5 % overhead
    // (with the tree preconditioner a solve takes one to three iterations: 2 per step)
    const int cg_per_step = o.cg_per_step > 0 ? std::min(o.cg_per_step, 64) : (a.precond ? 2 : 4);
    const dim3 blk(32, WY);
    const dim3 rgrd(groups, a.row_chunks), ngrd(groups, a.node_chunks);
    EventPair evp;
    CK(evp.create());
    cudaEvent_t ev0 = evp.a, ev1 = evp.b;
    CK(cudaEventRecord(ev0, s));
    int launches = 0;
    tn_init_kernel<<<rgrd, blk, 0, s>>>(a);
    tn_trial_kernel<<<rgrd, blk, 0, s>>>(a);
    launches += 2;
    CK(cudaGetLastError());
    int step_launches = 0;                                  // kernels of one optimiser step, counted while the first one is enqueued
    auto tree_solve = [&](bool factor) { step_launches += tn_launch_tree_solve(a, w.tn_lev_off_host, groups, s, factor, true); };
    auto one_step = [&]() -> int {
        int rc = eval_device(e, B, a.X, const_cast<double*>(a.Ft), const_cast<double*>(a.Gx), mode, true, s);
        if (rc != TPL_OK) return rc;
        tn_diag_kernel<<<dim3(groups, a.diag_chunks), blk, 0, s>>>(a);
        tn_setup_kernel<<<rgrd, blk, 0, s>>>(a);
        if (a.precond) tree_solve(true);
        for (int k = 0; k < cg_per_step; k++) {
            tn_launch_hv(a, ngrd, blk, s);
            tn_update_kernel<<<rgrd, blk, 0, s>>>(a);
            if (a.precond) tree_solve(false);
            tn_dir_kernel<<<rgrd, blk, 0, s>>>(a);
        }
        tn_gd_kernel<<<rgrd, blk, 0, s>>>(a);
        tn_stepmax_kernel<<<dim3(groups, a.tm_chunks), blk, 0, s>>>(a);
        tn_trial_kernel<<<rgrd, blk, 0, s>>>(a);
        CK(cudaGetLastError());
        return TPL_OK;
    };
    int steps = 0;
    int rc = one_step();               // first step outside any capture: sizes the evaluation work space
    if (rc != TPL_OK) return rc;
    const int per_step = 5 + e->last_launches + 3 * cg_per_step + step_launches;
    steps++;
    launches += per_step;
    GraphGuard gg(e);
    if (o.use_graph) {
        CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
        for (int k = 0; k < check_every && rc == TPL_OK; k++) rc = one_step();
        cudaError_t ce = cudaStreamEndCapture(s, &gg.graph);
        if (rc != TPL_OK || ce != cudaSuccess)
            return rc != TPL_OK ? rc : fail(TPL_ERR_CUDA, std::string("graph capture failed: ") + cudaGetErrorString(ce));
        CK(cudaGraphInstantiate(&gg.gexec, gg.graph, 0));
    }
    int ndone = 0;
    while (steps < max_steps) {
        // the last chunk is clamped to max_steps (plain launches; a graph replays whole chunks only)
        const int chunk = std::min(check_every, max_steps - steps);
        if (o.use_graph && chunk == check_every) {
            CK(cudaGraphLaunch(gg.gexec, s));
        } else {
            for (int k = 0; k < chunk && rc == TPL_OK; k++) rc = one_step();
            if (rc != TPL_OK) return rc;
        }
        steps += chunk;
        launches += per_step * chunk;
        CK(cudaMemcpyAsync(w.h_done.p, a.ndone, sizeof(int), cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        ndone = w.h_done.p[0];
        if (ndone >= B) break;
    }
    tn_final_kernel<<<rgrd, blk, 0, s>>>(a, w.Fout.p);
    launches++;
    CK(cudaGetLastError());
    CK(cudaEventRecord(ev1, s));
    if (io) {
        if (io->nprobe > 0) {          // only the probed rows of every vector return to the host (w.Hp as the gather buffer)
            const size_t np = (size_t)io->nprobe * B;
            if (np > PB) return fail(TPL_ERR_ARG, "more probes than parameters");
            CK(w.tn_probe_rows.ensure(np));
            CK(cudaMemcpyAsync(w.tn_probe_rows.p, io->probe_rows, sizeof(int) * np, cudaMemcpyHostToDevice, s));
            cols_probe_kernel<<<dim3((B + 127) / 128, std::min(io->nprobe, 1024)), 128, 0, s>>>(B, io->nprobe, w.X.p, w.tn_probe_rows.p, w.Hp.p);
            CK(cudaGetLastError());
            CK(cudaMemcpyAsync(io->probe_out, w.Hp.p, sizeof(double) * np, cudaMemcpyDeviceToHost, s));
        }
        if (io->Xout) CK(cudaMemcpyAsync(io->Xout, w.X.p, sizeof(double) * PB, cudaMemcpyDeviceToHost, s));
    } else {
        CK(cudaMemcpyAsync(X, w.X.p, sizeof(double) * PB, cudaMemcpyDeviceToHost, s));
    }
    CK(cudaMemcpyAsync(F, w.Fout.p, sizeof(double) * B, cudaMemcpyDeviceToHost, s));
    if (status) CK(cudaMemcpyAsync(status, a.flag, sizeof(int) * B, cudaMemcpyDeviceToHost, s));
    std::vector<double> tlog(B);
    std::vector<int> nit(B), ncg(B);
    CK(cudaMemcpyAsync(tlog.data(), a.tlog, sizeof(double) * B, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(nit.data(), a.iters, sizeof(int) * B, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(ncg.data(), a.ncg, sizeof(int) * B, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ev0, ev1);
    tpl_opt_result res;
    memset(&res, 0, sizeof(res));
    double ts = 0.0;
    long long ni = 0, nc = 0;
    for (int b = 0; b < B; b++) { ts += tlog[b]; ni += nit[b]; nc += ncg[b]; if (iters) iters[b] = nit[b]; }
    res.steps = steps;
    res.n_done = ndone;
    res.launches = launches;
    res.cg_iterations = (int)std::min<long long>(nc, 2147483647LL);
    res.device_ms = ms;
    res.mean_log10_step = ni > 0 ? ts / (double)ni : 0.0;
    e->last_launches = launches;
    if (result) *result = res;
    return TPL_OK;
}

}  // namespace

static int optimize_batch_impl(tpl_engine* e, int B, double* X, double* F, const double* lower, const double* upper,
                               int mode, const tpl_opt_options* opts, int* status, int* iters, tpl_opt_result* result, const ColsIO* io);

extern "C" int tpl_optimize_batch(tpl_engine* e, int B, double* X, double* F, const double* lower, const double* upper,
                                  int mode, const tpl_opt_options* opts, int* status, int* iters, tpl_opt_result* result) {
    if (!X) return fail(TPL_ERR_ARG, "tpl_optimize_batch: null X");
    return optimize_batch_impl(e, B, X, F, lower, upper, mode, opts, status, iters, result, nullptr);
}

extern "C" int tpl_optimize_batch_cols(tpl_engine* e, int B, int K, const double* Xstart, const int* start_col, double* F,
                                       const double* lower, const double* upper, int mode, const tpl_opt_options* opts, int* status,
                                       int* iters, tpl_opt_result* result, int nprobe, const int* probe_rows, double* probe_out,
                                       double* Xout) {
    if (!Xstart || !start_col || nprobe < 0 || (nprobe > 0 && (!probe_rows || !probe_out)))
        return fail(TPL_ERR_ARG, "tpl_optimize_batch_cols: null argument");
    if (!opts || opts->method != 1) return fail(TPL_ERR_ARG, "tpl_optimize_batch_cols: method 1 (truncated Newton) only");
    if (e && e->have_fp)
        for (size_t k = 0; k < (size_t)nprobe * (size_t)std::max(B, 0); k++)
            if (probe_rows[k] >= e->numparams) return fail(TPL_ERR_ARG, "probe row out of range");
    const ColsIO io = {K, Xstart, start_col, nprobe, probe_rows, probe_out, Xout};
    return optimize_batch_impl(e, B, nullptr, F, lower, upper, mode, opts, status, iters, result, &io);
}

static int optimize_batch_impl(tpl_engine* e, int B, double* X, double* F, const double* lower, const double* upper,
                               int mode, const tpl_opt_options* opts, int* status, int* iters, tpl_opt_result* result, const ColsIO* io) {
    using namespace tplopt;
    if (!e || !F || !lower || !upper || B < 1) return fail(TPL_ERR_ARG, "tpl_optimize_batch: null or B < 1");
    if (!e->have_fp) return fail(TPL_ERR_STATE, "set_freeparams has not been called");
    CK(cudaSetDevice(e->device));
    int rc = sync_device_tree(e);
    if (rc != TPL_OK) return rc;
    if (e->static_bad) return fail(TPL_ERR_STATE, "a fixed date violates its own bounds: every point is infeasible");
    tpl_opt_options o;
    memset(&o, 0, sizeof(o));
    if (opts) o = *opts;
    const int HM = o.history ? o.history : 8;
    if (HM != 4 && HM != 8 && HM != 16) return fail(TPL_ERR_ARG, "history must be 4, 8 or 16");
    const int max_steps = o.max_steps > 0 ? o.max_steps : 20000;
    const int check_every = o.check_every > 0 ? o.check_every : 32;
    const int P = e->numparams, n = e->n;
    int nr = 0;
    for (int i = 0; i < n; i++) nr = std::max(nr, e->freeparams[i] + 1);
    const double ds = o.date_scale > 0 ? o.date_scale : 1.0;
    // bounds in the transformed variables.  Date bounds are kept OPEN by a relative margin: exactly ON a penalised
    // bound the reference's objectives drop the infinite barrier term (calc_penalty resets, the AD paths skip it,
    // SURVEY.md 8 a13), so a step clamped onto the bound would sit in a spurious discontinuous minimum.
    std::vector<double> lo(P), hi(P), sc(P);
    for (int p = 0; p < P; p++) {
        sc[p] = p < nr ? 1.0 : ds;
        if (p < nr) {
            lo[p] = std::log(std::max(lower[p], 1e-300)) / sc[p];
            hi[p] = std::log(upper[p]) / sc[p];
        } else {
            double dlo = lower[p], dhi = upper[p];
            const double ref = dhi < 1e19 ? dhi : dlo;
            const double margin = 1e-9 * std::max(1.0, std::fabs(ref));
            dlo += margin;
            if (dhi < 1e19) dhi -= margin;
            lo[p] = dlo / sc[p];
            hi[p] = dhi / sc[p];
        }
        if (!(lo[p] <= hi[p])) return fail(TPL_ERR_ARG, "empty bound interval");
    }
    if (o.method == 1) return optimize_tn(e, B, X, F, lo, hi, sc, nr, mode, o, status, iters, result, io);
    if (io) return fail(TPL_ERR_ARG, "start columns / probes need method 1");
    if (o.method != 0) return fail(TPL_ERR_ARG, "method must be 0 (L-BFGS) or 1 (truncated Newton)");
    const size_t PB = (size_t)P * B;
    tpl_engine::OptWork& w = e->opt;
    const int NB = 2 * HM + 1, NACC = 6 * HM + 2;
    // Row chunking.  The two heavy kernels run as exactly ONE wave of equally loaded blocks (blocks = resident
    // slots of the kernel's own occupancy): with a few hundred long blocks a partial second wave would cost up to
    // half of the kernel's time (ncu: 1.14 waves -> 68 % of the HBM rate, profiles/r01j).
    const int groups = (B + 31) / 32;
    int occ_upd = 1, occ_dir = 1;
    if (HM == 4) {
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_upd, tplopt::opt_update_kernel<4>, 32 * 4, 0);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_dir, tplopt::opt_direction_kernel<4>, 32 * OPT_WY, 0);
    } else if (HM == 8) {
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_upd, tplopt::opt_update_kernel<8>, 32 * 8, 0);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_dir, tplopt::opt_direction_kernel<8>, 32 * OPT_WY, 0);
    } else {
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_upd, tplopt::opt_update_kernel<16>, 32 * 16, 0);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_dir, tplopt::opt_direction_kernel<16>, 32 * OPT_WY, 0);
    }
    auto one_wave = [&](int occ, int align, int* nch, int* rws) {
        int c = std::max(1, (e->num_sms * std::max(occ, 1)) / groups);
        c = std::min(c, (P + align - 1) / align);
        int r = (P + c - 1) / c;
        r = ((r + align - 1) / align) * align;
        *rws = r;
        *nch = (P + r - 1) / r;
    };
    int nchunks, rows, dir_chunks, dir_rows;
    one_wave(occ_upd, 16, &nchunks, &rows);          // whole row tiles of the update kernel for every history size
    one_wave(occ_dir, 2 * OPT_WY, &dir_chunks, &dir_rows);
    CK(w.Z.ensure(PB)); CK(w.Gz.ensure(PB)); CK(w.D.ensure(PB)); CK(w.X.ensure(PB)); CK(w.Gx.ensure(PB));
    CK(w.S.ensure(PB * HM)); CK(w.Y.ensure(PB * HM));
    CK(w.lo.ensure(P)); CK(w.hi.ensure(P)); CK(w.sc.ensure(P));
    CK(w.lane_d.ensure((size_t)6 * B)); CK(w.lane_i.ensure((size_t)9 * B + 1 + groups));
    CK(w.rho.ensure((size_t)HM * B)); CK(w.delta.ensure((size_t)NB * B)); CK(w.M.ensure((size_t)NB * NB * B));
    CK(w.part.ensure((size_t)nchunks * NACC * B)); CK(w.Ft.ensure(B)); CK(w.Fout.ensure(B));
    CK(w.h_done.ensure(1));
    cudaStream_t s = e->stream;
    CK(cudaMemcpyAsync(w.lo.p, lo.data(), sizeof(double) * P, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(w.hi.p, hi.data(), sizeof(double) * P, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(w.sc.p, sc.data(), sizeof(double) * P, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(w.X.p, X, sizeof(double) * PB, cudaMemcpyHostToDevice, s));
    CK(cudaMemsetAsync(w.lane_d.p, 0, sizeof(double) * 6 * B, s));
    CK(cudaMemsetAsync(w.lane_i.p, 0, sizeof(int) * (9 * (size_t)B + 1 + groups), s));      // st = ST_INIT = 0, t = 0
    CK(cudaMemsetAsync(w.rho.p, 0, sizeof(double) * HM * B, s));
    CK(cudaMemsetAsync(w.delta.p, 0, sizeof(double) * NB * B, s));
    CK(cudaMemsetAsync(w.M.p, 0, sizeof(double) * NB * NB * B, s));
    CK(cudaMemsetAsync(w.S.p, 0, sizeof(double) * PB * HM, s));
    CK(cudaMemsetAsync(w.Y.p, 0, sizeof(double) * PB * HM, s));
    CK(cudaMemsetAsync(w.Gz.p, 0, sizeof(double) * PB, s));
    OptArgs a;
    a.o.P = P; a.o.B = B; a.o.nr = nr;
    a.o.max_ls = o.max_ls > 0 ? o.max_ls : 40;
    a.o.patience = o.patience > 0 ? o.patience : 5;
    a.o.c1 = o.armijo > 0 ? o.armijo : 1e-4;
    a.o.ftol = o.ftol > 0 ? o.ftol : 1e-12;
    a.o.gtol = o.gtol > 0 ? o.gtol : 1e-8;
    a.Z = w.Z.p; a.Gz = w.Gz.p; a.D = w.D.p; a.X = w.X.p; a.Gx = w.Gx.p; a.S = w.S.p; a.Y = w.Y.p;
    a.lo = w.lo.p; a.hi = w.hi.p; a.sc = w.sc.p;
    int* li = w.lane_i.p;
    a.st = li; a.ls = li + B; a.head = li + 2 * (size_t)B; a.stall = li + 3 * (size_t)B; a.flag = li + 4 * (size_t)B;
    a.newdir = li + 5 * (size_t)B; a.iters = li + 6 * (size_t)B; a.nfail = li + 7 * (size_t)B;
    a.valid = (unsigned*)(li + 8 * (size_t)B);
    a.ndone = li + 9 * (size_t)B;
    double* ld = w.lane_d.p;
    a.F0 = ld; a.t = ld + B; a.gd = ld + 2 * (size_t)B; a.gamma = ld + 3 * (size_t)B; a.gmax = ld + 4 * (size_t)B; a.tlog = ld + 5 * (size_t)B;
    a.Ft = w.Ft.p;
    a.rho = w.rho.p; a.delta = w.delta.p; a.M = w.M.p; a.part = w.part.p;
    a.nchunks = nchunks; a.rows_per_chunk = rows;
    a.dir_chunks = dir_chunks; a.dir_rows = dir_rows;
    CK(w.part_red.ensure((size_t)NACC * B));
    a.part_red = w.part_red.p;
    const bool two_level = nchunks > 32;            // many chunks: reduce them with a grid of blocks first
    a.lane_part = two_level ? w.part_red.p : w.part.p;
    a.lane_nchunks = two_level ? 1 : nchunks;
    a.tcount = li + 9 * (size_t)B + 1;
    a.n = n; a.parent = e->d_parent.p; a.dslot = e->d_dslot.p; a.hfix = e->d_hfix.p;
    {   // light kernels carry ~40 registers: give the machine ~16 blocks per SM
        int lc = (16 * e->num_sms + groups - 1) / groups;
        lc = std::max(1, std::min(lc, (P + 4 * OPT_WY - 1) / (4 * OPT_WY)));
        int lr = (P + lc - 1) / lc;
        lr = ((lr + OPT_WY - 1) / OPT_WY) * OPT_WY;
        a.light_rows = lr;
        a.light_chunks = (P + lr - 1) / lr;
    }
    int tchunks = std::max(1, std::min(4 * nchunks, (n + 8 * OPT_WY - 1) / (8 * OPT_WY)));
    int npc = (n + tchunks - 1) / tchunks;
    npc = ((npc + OPT_WY - 1) / OPT_WY) * OPT_WY;
    tchunks = (n + npc - 1) / npc;
    a.tchunks = tchunks; a.nodes_per_chunk = npc;
    CK(w.tpart.ensure((size_t)tchunks * B));
    a.tpart = w.tpart.p;
    a.boundary_frac = o.boundary_frac < 0 ? 0.0 : (o.boundary_frac > 0 ? std::min(o.boundary_frac, 0.999999) : 0.9);
    tpl_opt_result res;
    memset(&res, 0, sizeof(res));
    EventPair evp;
    CK(evp.create());
    cudaEvent_t ev0 = evp.a, ev1 = evp.b;
    CK(cudaEventRecord(ev0, s));
    const bool graph = o.use_graph != 0;
    if (HM == 4) rc = run_optimizer<4>(e, a, B, mode, max_steps, check_every, graph, &res);
    else if (HM == 8) rc = run_optimizer<8>(e, a, B, mode, max_steps, check_every, graph, &res);
    else rc = run_optimizer<16>(e, a, B, mode, max_steps, check_every, graph, &res);
    if (rc != TPL_OK) return rc;
    CK(cudaEventRecord(ev1, s));
    CK(cudaMemcpyAsync(X, w.X.p, sizeof(double) * PB, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(F, w.Fout.p, sizeof(double) * B, cudaMemcpyDeviceToHost, s));
    if (status) CK(cudaMemcpyAsync(status, a.flag, sizeof(int) * B, cudaMemcpyDeviceToHost, s));
    if (iters) CK(cudaMemcpyAsync(iters, a.iters, sizeof(int) * B, cudaMemcpyDeviceToHost, s));
    std::vector<double> tlog(B);
    std::vector<int> nit(B);
    CK(cudaMemcpyAsync(tlog.data(), a.tlog, sizeof(double) * B, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(nit.data(), a.iters, sizeof(int) * B, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    {
        double ts = 0.0;
        long long ni = 0;
        for (int b = 0; b < B; b++) { ts += tlog[b]; ni += nit[b]; }
        res.mean_log10_step = ni > 0 ? ts / (double)ni : 0.0;
    }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ev0, ev1);
    res.device_ms = ms;
    e->last_launches = res.launches;
    if (result) *result = res;
    return TPL_OK;
}


extern "C" int tpl_debug_hessvec(tpl_engine* e, int B, const double* X, const double* V, const double* sc_rows, double* W, double* DG) {
    using namespace tpltn;
    if (!e || !X || !V || !sc_rows || !W || !DG || B < 1) return fail(TPL_ERR_ARG, "null or B < 1");
    if (!e->have_fp) return fail(TPL_ERR_STATE, "set_freeparams has not been called");
    CK(cudaSetDevice(e->device));
    int rc = sync_device_tree(e);
    if (rc != TPL_OK) return rc;
    const int P = e->numparams;
    const size_t PB = (size_t)P * B;
    int nr = 0;
    for (int i = 0; i < e->n; i++) nr = std::max(nr, e->freeparams[i] + 1);
    std::vector<double> lo(P, -1e300), hi(P, 1e300), sc(sc_rows, sc_rows + P);
    tpl_opt_options o;
    memset(&o, 0, sizeof(o));
    o.precond = 1;                      // no tree work space for a single product
    TnArgs a;
    rc = tn_prepare(e, B, X, lo, hi, sc, nr, o, a);
    if (rc != TPL_OK) return rc;
    cudaStream_t s = e->stream;
    tpl_engine::OptWork& w = e->opt;
    rc = eval_device(e, B, a.X, const_cast<double*>(a.Ft), const_cast<double*>(a.Gx), TPL_GRAD_AD, true, s);
    if (rc != TPL_OK) return rc;
    const dim3 blk(32, tplopt::OPT_WY);
    const dim3 ngrd((B + 31) / 32, a.node_chunks);
    CK(cudaMemcpyAsync(w.Pv.p, V, sizeof(double) * PB, cudaMemcpyHostToDevice, s));
    tn_diag_kernel<<<dim3((B + 31) / 32, a.diag_chunks), blk, 0, s>>>(a);                     // st = TS_INIT and a finite objective: every vector "accepts"
    std::vector<int> cg(B, TS_CG);
    CK(cudaMemcpyAsync(a.st, cg.data(), sizeof(int) * B, cudaMemcpyHostToDevice, s));
    tn_launch_hv(a, ngrd, blk, s);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(W, w.Hp.p, sizeof(double) * PB, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(DG, w.DG.p, sizeof(double) * PB, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return TPL_OK;
}

// test hook: Out = M^-1 Rhs with the block-tree factorisation of the Hessian at X; rows with Free > 0 are the variables
extern "C" int tpl_debug_treesolve(tpl_engine* e, int B, const double* X, const double* Free, const double* Rhs, const double* sc_rows,
                                   double* Out, const double* Rhs2, double* Out2) {
    using namespace tpltn;
    if (!e || !X || !Free || !Rhs || !sc_rows || !Out || B < 1) return fail(TPL_ERR_ARG, "null or B < 1");
    if (!e->have_fp) return fail(TPL_ERR_STATE, "set_freeparams has not been called");
    CK(cudaSetDevice(e->device));
    int rc = sync_device_tree(e);
    if (rc != TPL_OK) return rc;
    const int P = e->numparams;
    const size_t PB = (size_t)P * B;
    int nr = 0;
    for (int i = 0; i < e->n; i++) nr = std::max(nr, e->freeparams[i] + 1);
    std::vector<double> lo(P, -1e300), hi(P, 1e300), sc(sc_rows, sc_rows + P);
    tpl_opt_options o;
    memset(&o, 0, sizeof(o));
    TnArgs a;
    rc = tn_prepare(e, B, X, lo, hi, sc, nr, o, a);
    if (rc != TPL_OK) return rc;
    cudaStream_t s = e->stream;
    tpl_engine::OptWork& w = e->opt;
    rc = eval_device(e, B, a.X, const_cast<double*>(a.Ft), const_cast<double*>(a.Gx), TPL_GRAD_AD, true, s);
    if (rc != TPL_OK) return rc;
    CK(cudaMemcpyAsync(w.Minv.p, Free, sizeof(double) * PB, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(w.R.p, Rhs, sizeof(double) * PB, cudaMemcpyHostToDevice, s));
    tn_launch_tree_solve(a, w.tn_lev_off_host, (B + 31) / 32, s, true, false);   // st = TS_INIT, finite objective: every vector active
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(Out, w.ZZ.p, sizeof(double) * PB, cudaMemcpyDeviceToHost, s));
    if (Rhs2 && Out2) {
        // the solve-only sweeps of a CG iteration (FACTOR = false) with the factors just computed: every vector in TS_CG / ACT_CG
        std::vector<int> st(B, TS_CG), act(B, ACT_CG);
        CK(cudaMemcpyAsync(a.st, st.data(), sizeof(int) * B, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(a.act, act.data(), sizeof(int) * B, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(w.R.p, Rhs2, sizeof(double) * PB, cudaMemcpyHostToDevice, s));
        tn_launch_tree_solve(a, w.tn_lev_off_host, (B + 31) / 32, s, false, false);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(Out2, w.ZZ.p, sizeof(double) * PB, cudaMemcpyDeviceToHost, s));
    }
    CK(cudaStreamSynchronize(s));
    return TPL_OK;
}

// ---- batched simulated annealing (tpl_anneal_core.h / tpl_anneal.cuh) -----------------------------------------------

extern "C" int tpl_anneal_batch(tpl_engine* e, int B, double* X, double* F, const tpl_anneal_options* opts, int* status, int* iterations) {
    using namespace tplsa;
    if (!e || !X || !F || B < 1) return fail(TPL_ERR_ARG, "tpl_anneal_batch: null or B < 1");
    if (!e->have_fp) return fail(TPL_ERR_STATE, "set_freeparams has not been called");
    CK(cudaSetDevice(e->device));
    int rc = sync_device_tree(e);
    if (rc != TPL_OK) return rc;
    if (e->lf_mode) return fail(TPL_ERR_ARG, "the batched annealer needs the PL layout (one rate per branch)");
    if (e->cv_any) return fail(TPL_ERR_ARG, "the batched annealer takes CV masks per vector (tpl_batch_configure), not set_cv_nodes");
    if (e->static_bad) return fail(TPL_ERR_STATE, "a fixed date violates its own bounds: every point is infeasible");
    tpl_anneal_options o;
    memset(&o, 0, sizeof(o));
    if (opts) o = *opts;
    const int P = e->numparams, n = e->n;
    const size_t PB = (size_t)P * B;
    int nr = 0;
    for (int i = 0; i < n; i++) nr = std::max(nr, e->freeparams[i] + 1);
    std::vector<int> date_node, bar_node;
    for (int i = 0; i < n; i++) if (e->freeparams[n + i] != -1) date_node.push_back(i);
    for (int i = 0; i < n; i++) if (e->pen_min[i] != -1 || e->pen_max[i] != -1) bar_node.push_back(i);
    tpl_engine::OptWork& w = e->opt;
    cudaStream_t s = e->stream;
    CK(w.X.ensure(PB));
    CK(w.sa_date_node.upload(date_node, s));
    CK(w.sa_bar_node.upload(bar_node, s));
    CK(w.sa_hmin.upload(e->hmin, s));
    CK(w.sa_hmax.upload(e->hmax, s));
    CK(w.sa_chains.ensure((size_t)B * sizeof(Chain)));
    CK(w.sa_done.ensure((size_t)B + 1));
    CK(w.h_done.ensure(1));
    CK(cudaMemcpyAsync(w.X.p, X, sizeof(double) * PB, cudaMemcpyHostToDevice, s));
    CK(cudaMemsetAsync(w.sa_done.p, 0, sizeof(int) * ((size_t)B + 1), s));
    CK(cudaStreamSynchronize(s));                   // the uploads above read temporaries
    SaArgs a;
    a.o.n = n; a.o.P = P; a.o.nr = nr; a.o.ndates = (int)date_node.size();
    a.o.maxit = o.max_iterations > 0 ? o.max_iterations : 15000;            // lfsimaniter / plsimaniter (optim_options.cpp)
    a.o.cool = o.cool > 0 ? o.cool : 0.999;                                  // utils.cpp:689
    a.o.stop_temp = o.stop_temperature > 0 ? o.stop_temperature : 0.001;
    a.o.pb = e->pb;
    a.o.log_pen = e->log_pen ? 1 : 0;
    a.t = dev_tree(e);
    a.date_node = w.sa_date_node.p;
    a.bar_node = w.sa_bar_node.p;
    a.hmin = w.sa_hmin.p;
    a.hmax = w.sa_hmax.p;
    const bool lanemask = e->cfg_mask && e->cfg_B == B;
    a.lanemask = lanemask ? e->d_lanemask.p : nullptr;
    a.mask_stride = e->ntiles2 * tpl::TILE_NODES;
    a.lane_lambda = (e->cfg_lambda && e->cfg_B == B) ? e->d_lane_lambda.p : nullptr;
    a.lambda = e->smoothing;
    a.X = w.X.p;
    a.chains = reinterpret_cast<Chain*>(w.sa_chains.p);
    a.done = w.sa_done.p;
    a.ndone = w.sa_done.p + B;
    a.B = B;
    const int grid = (B + 31) / 32;
    sa_init_kernel<<<grid, 32, 0, s>>>(a, o.temperature > 0 ? o.temperature : 50.0, o.rate_step > 0 ? o.rate_step : 0.07,
                                       o.date_step > 0 ? o.date_step : 0.25, o.seed ? o.seed : 1ull);
    CK(cudaGetLastError());
    int launches = 1;
    // a chain needs maxit counted iterations plus its discarded proposals; the cap only guards against a chain whose every
    // proposal is infeasible
    const long long cap = 40LL * a.o.maxit + 100000;
    long long issued = 0;
    int ndone = 0;
    while (ndone < B && issued < cap) {
        sa_run_kernel<<<grid, 32, 0, s>>>(a, 2048);
        issued += 2048;
        launches++;
        CK(cudaMemcpyAsync(w.h_done.p, a.ndone, sizeof(int), cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        ndone = w.h_done.p[0];
    }
    std::vector<Chain> chains(B);
    std::vector<int> done(B);
    CK(cudaMemcpyAsync(chains.data(), a.chains, sizeof(Chain) * B, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(done.data(), a.done, sizeof(int) * B, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(X, w.X.p, sizeof(double) * PB, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    for (int b = 0; b < B; b++) {
        F[b] = done[b] == 2 ? TPL_LARGE : chains[b].f;
        if (status) status[b] = done[b] == 2 ? 4 : (done[b] == 1 ? 1 : 0);
        if (iterations) iterations[b] = chains[b].iteration;
    }
    e->last_launches = launches;
    return TPL_OK;
}
