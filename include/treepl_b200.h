/* treepl_b200.h — C ABI of the B200-native penalized-likelihood (PL) engine.
 *
 * Drop-in boundary for ONE hot path of blackrim/treePL: the objective and gradient that
 * `class pl_calc_parallel` (reference src/pl_calc_parallel.h:17-82) serves to the TNC / NLopt
 * callbacks (src/optimize_tnc.cpp:33-105, src/optimize_nlopt.cpp:22-105), to the annealer
 * (src/siman_calc_par.cpp:154) and to the CV loop (src/main.cpp:689-801).
 * Plain pointers and sizes only; every entry point cites the reference interface it replaces.
 * All node-indexed arrays use the REFERENCE node numbering (apply_node_label_preorder,
 * src/utils.cpp:167-180; root = 0) and the reference parameter layout
 * (generate_param_order_vector, src/utils.cpp:295-340).
 *
 * There is NO CPU fallback: every compute entry point runs sm_100a kernels and fails
 * (returns TPL_ERR_* / NaN, see tpl_last_error) when no CUDA device is usable.
 *
 * Threading: calls on DIFFERENT handles may run concurrently (each handle owns its CUDA
 * stream and staging buffers, as each OpenMP fold owns its pl_calc_parallel, main.cpp:701);
 * calls on the same handle must be serialised by the caller, as in the reference.
 */
#ifndef TREEPL_B200_H
#define TREEPL_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define TPL_LARGE 1e+15              /* infeasible-point sentinel, pl_calc_parallel.cpp:29 */

#define TPL_OK 0
#define TPL_ERR_CUDA (-1)            /* CUDA runtime / no device */
#define TPL_ERR_ARG (-2)             /* bad argument */
#define TPL_ERR_STATE (-3)           /* call order (e.g. gradient before set_freeparams) */

/* gradient flavours: the reference has two mutually inconsistent gradients */
#define TPL_GRAD_ANALYTIC 0          /* calc_gradient: calc_pl_gradient / calc_lf_gradient  (:657-818) */
#define TPL_GRAD_AD 1                /* calc_function_gradient: exact adjoint of the RAD / ADOL-C tapes (:132-655) */

typedef struct tpl_engine tpl_engine;

/* ---- lifetime ------------------------------------------------------------------------------ */

/* pl_calc_parallel::setup_starting_bits (pl_calc_parallel.cpp:34-71) plus the member
 * assignments main() makes right after it (main.cpp:486-490: pen_min, pen_max, children_vec).
 * `children` is children_vec flattened in node order (node 0's children, node 1's, ...), each
 * list in Newick order; its length is sum(child_counts) = numnodes-1.  All arrays are COPIED
 * (the reference borrows them; copying removes the lifetime coupling).  smoothing starts at 1,
 * penalty_boundary at 0.0025, log_pen off, no CV mask — as in the reference.
 * `device` = CUDA device ordinal (-1: current device).                                        */
int tpl_create(tpl_engine** out, int device, int numnodes,
               const int* parents_nds_ints, const int* child_counts, const int* children,
               const int* free_, const double* char_durations, const double* log_fact_char_durations,
               const double* min_, const double* max_, const double* pen_min, const double* pen_max,
               const double* start_dates, const double* start_rates, const double* start_durations);

/* pl_calc_parallel::delete_ad_arrays + destructor (main.cpp:798) */
void tpl_destroy(tpl_engine* e);

/* Device memory of destroyed engines is kept by the library for the next engine of the process (a CV run builds one
 * engine per phase, each with tens of GB of work space; cudaMalloc / cudaFree of that size cost tenths of a second).
 * This returns every cached block to the driver.  The cache is also emptied before an allocation failure is reported. */
int tpl_release_cached_memory(void);

/* message of the last failing call on this thread.  The evaluation entry points that return a double (tpl_calc_pl,
 * tpl_calc_function_gradient, tpl_calc_pl_grad) clear it on entry: NaN with an EMPTY message is a NaN objective (the
 * reference propagates NaN to its wrappers), NaN with a message is a failure. */
const char* tpl_last_error(void);

/* ---- configuration (public data members / setters of pl_calc_parallel) ---------------------- */
int tpl_set_smoothing(tpl_engine* e, double smoothing);            /* plp.smoothing = ...     main.cpp:490,669 */
int tpl_set_log_pen(tpl_engine* e, int log_pen);                   /* set_log_pen             pl_calc_parallel.cpp:1000 */
int tpl_set_penalty_boundary(tpl_engine* e, double pb);            /* penalty_boundary member pl_calc_parallel.cpp:69 */
int tpl_set_cv_nodes(tpl_engine* e, const int* cvnodes);           /* set_cv_nodes (:959); NULL clears the mask */
/* overwrite member rates/dates, as the CV loop does by re-running setup_starting_bits on the
 * snapshot vectors (main.cpp:701-704) */
int tpl_set_state(tpl_engine* e, const double* rates, const double* dates);

/* generate_param_order_vector (utils.cpp:295-340), host helper: fills freeparams[2*numnodes],
 * returns numparams.  cvnodes may be NULL.                                                     */
int tpl_generate_param_order_vector(int numnodes, const int* free_, int lf, const int* cvnodes, int* freeparams);

/* set_freeparams (pl_calc_parallel.cpp:920-955): installs the 2N slot map and the LF/PL switch,
 * and emits the initial parameter vector from the current member rates/dates into params_out
 * (capacity >= numparams; may be NULL).  Returns the number of values emitted or TPL_ERR_*.     */
int tpl_set_freeparams(tpl_engine* e, int numparams, int lf, const int* freeparams, double* params_out);

int tpl_numnodes(const tpl_engine* e);
int tpl_numparams(const tpl_engine* e);

/* ---- single-point evaluation: what the optimiser callbacks call ------------------------------ */

/* pl_calc_parallel::calc_pl (pl_calc_parallel.cpp:73-130).  x = host pointer, numparams doubles.
 * Returns the objective, TPL_LARGE for an infeasible point, NaN on engine failure.  Like the
 * reference it makes x the engine's current state (rates/dates/durations) for later readers.   */
double tpl_calc_pl(tpl_engine* e, const double* x);

/* pl_calc_parallel::calc_gradient (:657-662).  Like the reference it differentiates at the state
 * left by the preceding tpl_calc_pl (it does not re-read x, except that it must be the same
 * vector); g = host pointer, numparams doubles.                                                 */
int tpl_calc_gradient(tpl_engine* e, const double* x, double* g);

/* pl_calc_parallel::calc_function_gradient (:132-138) == calc_pl_function_gradient_adolc (:622):
 * objective of the AD expression and its exact gradient.  Does not change the engine state.
 * An infeasible point returns TPL_LARGE and leaves g untouched, as the reference does.          */
double tpl_calc_function_gradient(tpl_engine* e, const double* x, double* g);

/* Fused form of function_plcp (optimize_tnc.cpp:75-105): one upload, one launch sequence, one
 * download.  mode = TPL_GRAD_ANALYTIC gives calc_pl + calc_gradient, TPL_GRAD_AD gives
 * calc_function_gradient.  g may be NULL (objective only, the annealer's call).                 */
double tpl_calc_pl_grad(tpl_engine* e, const double* x, double* g, int mode);

/* read-back of the public members main() reads after an optimisation (main.cpp:679-681,749,865-878) */
int tpl_get_rates(tpl_engine* e, double* out);
int tpl_get_dates(tpl_engine* e, double* out);
int tpl_get_durations(tpl_engine* e, double* out);

/* ---- batched evaluation: B independent parameter vectors per launch ---------------------------
 * The batch dimension is innermost: X[p*B + b] is parameter p of vector b, G likewise, F[b].
 * Vectors use the FULL parameter layout of the installed slot map (set_freeparams); vector b may
 * additionally carry its own smoothing value and its own CV mask (folds, main.cpp:693-696).  A masked node's
 * rate row is not a variable (its gradient row is written as 0, optimisers and the annealer never move it); its
 * VALUE is the rate the pruned branch keeps — what the reference's per-fold engine holds in its member `rates`
 * (the full-data optimum, main.cpp:681,702-704), which still enters the root-children variance term when the
 * pruned tip hangs off the root (pl_calc_parallel.cpp:866-878).
 * This is the data-parallel form of the fold loop (main.cpp:689), of the lambda grid
 * (main.cpp:668) and of multi-start optimisation (utils.cpp:475-553).                           */

/* lane_smoothing: B doubles or NULL (engine smoothing for all).  mask_off/mask_nodes: CSR list
 * of masked node numbers per vector (mask_off has B+1 entries) or NULL/NULL for no masks.
 * The configuration stays installed until the next call.                                       */
int tpl_batch_configure(tpl_engine* e, int B, const double* lane_smoothing,
                        const int* mask_off, const int* mask_nodes);

/* host buffers: X[P*B] in, F[B] and G[P*B] out (G may be NULL: objective only).  Any host memory works; with 128 or more vectors
 * the batch moves as pipelined 64-vector pieces (upload, kernels and download of neighbouring pieces overlap), and PINNED buffers
 * (tpl_host_alloc, cudaMallocHost, cudaHostRegister) of 32 ... 126 vectors are evaluated in place over PCIe, without staging copies. */
int tpl_calc_pl_grad_batch(tpl_engine* e, int B, const double* X, double* F, double* G, int mode);

/* device buffers, asynchronous on `cuda_stream` (a cudaStream_t, NULL = the engine's stream).
 * No host synchronisation is performed.                                                        */
int tpl_calc_pl_grad_batch_dev(tpl_engine* e, int B, const double* dX, double* dF, double* dG,
                               int mode, void* cuda_stream);

/* ---- device-resident batched optimiser --------------------------------------------------------
 * Replaces, for B independent problems at once, the host optimiser round trips of the reference:
 * optimize_plcp_tnc (optimize_tnc.cpp:107-172) / optimize_plcp_nlopt (optimize_nlopt.cpp:116-238) as
 * driven by optimize_full (utils.cpp:671-763) inside the smoothing-grid / CV-fold loops of main()
 * (main.cpp:668-682, 689-801).  Projected L-BFGS with Armijo backtracking; rates are optimised as
 * ln r, dates as h / date_scale; every vector keeps its own line-search and convergence state in HBM
 * and every trial point of all B vectors is ONE launch of the PL kernels (no host round trip per
 * evaluation; the host only polls a done counter every `check_every` steps).
 * X [P][B] host in/out (start must be feasible), F [B] host out.  lower/upper [P]: the bounds the
 * reference adapters build (optimize_tnc.cpp:116-147).  Uses the batch configuration installed by
 * tpl_batch_configure when it was made for the same B (per-vector smoothing and CV masks).
 * status [B] (may be NULL): 1 gradient converged, 2 objective converged, 3 stalled, 4 infeasible start,
 * 5 non-finite gradient, 0 step cap reached.  iters [B] (may be NULL): accepted steps per vector.      */
typedef struct tpl_opt_options {
    int history;        /* L-BFGS pairs kept: 4, 8 or 16 (0 = default 8) */
    int max_steps;      /* cap on batched objective+gradient evaluations (0 = default 20000) */
    int patience;       /* stop a vector after this many consecutive relative decreases <= ftol (0 = 5) */
    int max_ls;         /* halvings per line search (0 = 40) */
    int check_every;    /* host polls the device-side done counter every this many steps (0 = 32) */
    int use_graph;      /* 1: replay `check_every` steps as one CUDA graph launch */
    int method;         /* 0: L-BFGS (any layout); 1: truncated Newton, CG on the ANALYTIC Hessian-vector product — what the
                         * reference's TNC approximates with finite differences (tnc.c:483-827).  PL layout, AD gradient,
                         * plain or log penalty. */
    int max_cg;         /* method 1: cap on CG iterations per Newton step (0 = 500; the forcing sequence usually stops far earlier) */
    int cg_per_step;    /* method 1: CG iterations run between two batched evaluations (0 = 2 with the tree preconditioner, 4 with Jacobi) */
    int precond;        /* method 1: preconditioner of the CG solve.  0 / 2: exact block-tree factorisation of the Hessian
                         * (children eliminated before parents, 2x2 pivots; a handful of CG iterations per Newton step for
                         * every smoothing value); 1: Jacobi (round-1 behaviour; TNC itself uses a diagonal, tnc.c:829-905) */
    double ftol;        /* 0 = 1e-12 */
    double gtol;        /* 0 = 1e-8; on max |projected gradient| relative to max(|f|, 1) */
    double armijo;      /* 0 = 1e-4 */
    double date_scale;  /* 0 = 1 */
    double boundary_frac; /* a new direction's first step covers at most this fraction of the distance to the
                           * nearest zero-duration boundary (ratio test over the branches); 0 = 0.9, < 0 = off */
    int tree_wide_min;  /* method 1, tree preconditioner: a height level with at least this many nodes gets its own launch per
                         * sweep; the narrower levels above are walked inside one block per 32-vector group (0 = 64) */
    int tree_coop;      /* method 1, tree preconditioner: the wide levels of a sweep run as ONE cooperative launch with grid
                         * barriers between the levels (0 = where it was measured to pay: 16+ vectors in at most 6 groups of 32; 1 = always where
                         * the device supports it; -1 = one launch per level) */
    int hv_pipe;        /* method 1: Hessian-vector product through the software-pipelined kernel (operand rows staged in shared
                         * memory by cp.async, three nodes per warp in flight) for batches of more than 16 vectors (0 = yes,
                         * -1 = the register-staged kernel everywhere) */
    int reserved;
} tpl_opt_options;

typedef struct tpl_opt_result {
    int steps;          /* batched evaluations performed */
    int n_done;         /* vectors that stopped by themselves */
    int launches;       /* kernels launched */
    int cg_iterations;  /* method 1: Hessian-vector products used, summed over the vectors */
    double device_ms;   /* CUDA-event time of the whole optimisation on the engine stream */
    double mean_log10_step; /* diagnostic: mean log10 of the accepted step lengths over all vectors (0 = unit steps) */
} tpl_opt_result;

int tpl_optimize_batch(tpl_engine* e, int B, double* X, double* F, const double* lower, const double* upper,
                       int mode, const tpl_opt_options* opts, int* status, int* iters, tpl_opt_result* result);

/* The same optimisation for a cross-validation job: vector b starts from column start_col[b] of the host matrix
 * Xstart [P][K] (K <= B: e.g. the full-data optimum of the vector's smoothing value - what every fold of the reference's
 * CV loop starts from, main.cpp:681,702-704), so K columns instead of B cross the bus and the batch is expanded on the
 * device; and only `nprobe` rows of every optimised vector come back: probe_out[j][b] = X[probe_rows[j][b]][b]
 * (host [nprobe][B] each, probe_rows = -1 -> 0) - the rows the fold's chi-square needs (main.cpp:737-793).  Xout
 * (host [P][B]) may be NULL.  Method 1 only; everything else as tpl_optimize_batch. */
int tpl_optimize_batch_cols(tpl_engine* e, int B, int K, const double* Xstart, const int* start_col, double* F,
                            const double* lower, const double* upper, int mode, const tpl_opt_options* opts, int* status,
                            int* iters, tpl_opt_result* result, int nprobe, const int* probe_rows, double* probe_out,
                            double* Xout);

/* ---- batched simulated annealing ------------------------------------------------------------------
 * siman_calc_par::optimize (siman_calc_par.cpp:80-229) for B chains at once, one GPU thread per chain: the reference's
 * proposal (one parameter scaled by 1 + u, rates and dates with probability 1/2 each), acceptance rule, cooling, step-size
 * adaptation and stopping rules, with the objective updated by an O(degree) delta evaluation instead of a full calc_pl per
 * proposal.  Objective = calc_pl (TPL_GRAD_ANALYTIC flavour), PL layout; per-vector smoothing and CV masks of the installed
 * batch configuration apply.  The random stream is a per-chain xorshift generator, not libc rand(): trajectories are
 * reproducible for a given seed but are not the reference's.
 * X [P][B] host in/out, F [B] out; status [B] (may be NULL): 1 finished, 4 infeasible start; iterations [B] (may be NULL). */
typedef struct tpl_anneal_options {
    double temperature;        /* 0 = 50     (set_pl arguments of optimize_full, utils.cpp:689) */
    double cool;               /* 0 = 0.999 */
    double rate_step;          /* 0 = 0.07 */
    double date_step;          /* 0 = 0.25 */
    double stop_temperature;   /* 0 = 0.001 */
    int max_iterations;        /* 0 = 15000 (lfsimaniter, optim_options.cpp) */
    unsigned long long seed;   /* 0 = 1 */
} tpl_anneal_options;

int tpl_anneal_batch(tpl_engine* e, int B, double* X, double* F, const tpl_anneal_options* opts, int* status, int* iterations);

/* number of kernels the last evaluation call launched (for bench accounting) */
int tpl_last_launch_count(const tpl_engine* e);

/* pinned host memory for X/F/G staging */
void* tpl_host_alloc(unsigned long long bytes);
void tpl_host_free(void* p);

/* Device timing for roofline accounting (bench.py): while on, every evaluation records CUDA
 * events around its main kernel and its finalize kernel on the launching stream (up to 4096
 * evaluations).  tpl_get_timing waits for them, returns the number of evaluations and the summed
 * kernel milliseconds, and resets the log.                                                     */
int tpl_set_timing(tpl_engine* e, int on);
int tpl_get_timing(tpl_engine* e, double* main_ms, double* finalize_ms);

/* test hooks: pin the kernel generation (0 = automatic, 1 = generic "pull" kernel, 2 = shared-memory
 * tile kernel whenever it is legal; 5 = automatic generation, but the host-buffer pipeline of tpl_calc_pl_grad_batch in
 * uniform 64-vector pieces), and report which one the last evaluation used              */
int tpl_debug_force_kernel(tpl_engine* e, int generation);
int tpl_last_kernel_generation(const tpl_engine* e);

/* test hook: W = H_z V and DG = diag(H_z) at X through the truncated-Newton kernels (z = w / sc_rows[p], w = ln r | h);
 * X, V, W, DG: host [P][B].  Uses the installed batch configuration (per-vector smoothing and masks) when made for B. */
int tpl_debug_hessvec(tpl_engine* e, int B, const double* X, const double* V, const double* sc_rows, double* W, double* DG);

/* test hook: Out = M^-1 Rhs, M = the truncated-Newton method's block-tree factorisation of H_z at X (tpl_opt_options.precond);
 * rows with Free > 0 are the variables of the solve, the others come out as 0.  X, Free, Rhs, Out: host [P][B].
 * Rhs2 / Out2 (both or neither, may be NULL): a second right-hand side solved with the SAME factors through the
 * solve-only sweeps a CG iteration runs. */
int tpl_debug_treesolve(tpl_engine* e, int B, const double* X, const double* Free, const double* Rhs, const double* sc_rows,
                        double* Out, const double* Rhs2, double* Out2);

/* test hook: evaluates the kernels' log and reciprocal on n host values (tests only) */
int tpl_debug_math(int n, const double* x, double* log_out, double* rcp_out);

/* library / device info: returns sm version * 10 (e.g. 1000) of the engine's device, <0 on error */
int tpl_device_sm(const tpl_engine* e);

#ifdef __cplusplus
}
#endif
#endif /* TREEPL_B200_H */
