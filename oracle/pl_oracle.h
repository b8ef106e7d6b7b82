/* TEST INFRASTRUCTURE ONLY — see pl_oracle.c.  Plain-C restatement of the reference's
 * penalized-likelihood objective and gradients (src/pl_calc_parallel.cpp). */
#ifndef TREEPL_B200_PL_ORACLE_H
#define TREEPL_B200_PL_ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif

#define PLO_LARGE 1e+15            /* pl_calc_parallel.cpp:29 */

typedef struct {
    int n;                          /* numnodes, root = 0 */
    const int* parent;              /* [n], parent[0] = -1                      (parents_nds_ints) */
    const int* child_off;           /* [n+1] CSR offsets into children                              */
    const int* children;            /* [n-1] child node numbers, Newick order   (children_vec)     */
    const int* free_;               /* [n]                                      (free)             */
    const double* c;                /* [n] char_durations                                          */
    const double* lf;               /* [n] log_fact_char_durations                                 */
    const double* min_;             /* [n]                                                         */
    const double* max_;             /* [n]                                                         */
    const double* pen_min;          /* [n], -1 = none                                              */
    const double* pen_max;          /* [n], -1 = none                                              */
    const int* freeparams;          /* [2n] slot maps (generate_param_order_vector)                */
    const int* cv;                  /* [n] 0/1 mask (cvnodes)                                      */
    int numparams;
    int lf_mode;                    /* pl_calc_parallel::lf                                        */
    int log_pen;
    double smoothing;
    double penalty_boundary;        /* 0.0025, pl_calc_parallel.cpp:69                             */
    double* rates;                  /* [n] member state, mutated by plo_calc_pl                    */
    double* dates;                  /* [n]                                                         */
    double* durations;              /* [n]                                                         */
} plo_problem;

double plo_calc_pl(plo_problem* p, const double* x);
void plo_calc_gradient(plo_problem* p, const double* x, double* g);
double plo_calc_function_gradient(plo_problem* p, const double* x, double* g);
int plo_generate_param_order_vector(int n, const int* free_, int lf, const int* cv /*or NULL*/, int* freeparams /*[2n]*/);
int plo_set_freeparams_emit(const plo_problem* p, double* x0);

#ifdef __cplusplus
}
#endif
#endif
