/*
 * b2s.h -- C ABI of the B200-native dense-tableau two-phase simplex hot path.
 *
 * This is the drop-in boundary for the solve path of rik1599/SimplexOnCuda.  The reference has
 * no C ABI of its own (its API is C++-linkage: include/problem.h, include/solver.h,
 * include/twoPhaseMethod.h, include/reduction.cuh, include/gaussian.cuh, include/tabular.cuh);
 * every entry point below names the reference interface (file:line under the reference tree)
 * it replaces.  The C++ headers with the reference's own names (include/compat/) are thin shims
 * over these functions.
 *
 * Conventions
 *   - plain pointers and sizes only; host pointers unless the name says `_device`;
 *   - every function returns a b2s error code (B2S_OK == 0); nothing calls exit() (the
 *     reference prints and exits on any CUDA error, src/error.cu:5-18);
 *   - solver statuses are the reference's (include/twoPhaseMethod.h:5-8) plus two new ones;
 *   - the tableau is "variable-major" exactly like the reference (include/tabular.cuh:5-30):
 *     device row r holds tableau column r (row 0 = RHS b, rows 1..n structural variables, then
 *     m slack rows, then m artificial rows), m contiguous values per row; reduced costs live in
 *     a separate vector whose element 0 is the objective value.
 */
#ifndef B2S_H
#define B2S_H

#include <stddef.h> /* size_t */

#ifdef __cplusplus
extern "C" {
#endif

/* ---- error codes (return values) ---------------------------------------------------------- */
#define B2S_OK 0
#define B2S_ERR_ARG 1    /* bad argument                                  */
#define B2S_ERR_STATE 2  /* call out of order (e.g. iterate before build) */
#define B2S_ERR_CUDA 3   /* CUDA runtime error, see b2s_last_error()      */
#define B2S_ERR_NOMEM 4  /* device or host allocation failed              */
#define B2S_ERR_NCCL 5   /* NCCL error (sharded solves)                   */
#define B2S_ERR_NOGPU 6  /* no CUDA device: there is no CPU fallback      */
#define B2S_ERR_PEER 7   /* sharded solve: a peer rank stopped publishing (bounded wait expired); the handle must be
                            reloaded before it can solve again                                                      */

/* ---- solver statuses (include/twoPhaseMethod.h:5-8, src/solver.cu:77) -------------------- */
#define B2S_FEASIBLE 0
#define B2S_INFEASIBLE (-1)
#define B2S_UNBOUNDED (-2)
#define B2S_DEGENERATE (-3)
#define B2S_ITER_LIMIT (-4) /* new: pivot budget exhausted (the reference has no cap) */
#define B2S_RUNNING (-10)   /* NOT_ENDED in src/solver.cu:77 */

/* ---- options ----------------------------------------------------------------------------------- */
#define B2S_F64 0
#define B2S_F32 1

#define B2S_RULE_REFERENCE 0 /* epsilon-tournament argmin, reference tie order (src/reduction.cu:10-104) */
#define B2S_RULE_LOWEST 1    /* exact Dantzig minimum, lowest index among ties                           */
#define B2S_RULE_BLAND 2     /* Bland's anti-cycling rule                                                */

#define B2S_RAND_GLIBC 0 /* srand/rand of glibc: what the reference yields when built on Linux */
#define B2S_RAND_MSVC 1  /* MSVC rand(): what the reference's published runs used (Windows)    */

typedef struct b2s_options {
    int device;           /* CUDA device ordinal (reference: always 0, main.cu:126)                 */
    int dtype;            /* B2S_F64 (reference TYPE, include/macro.h:6) or B2S_F32                  */
    int pivot_rule;       /* B2S_RULE_*                                                               */
    int fold_artificials; /* 1: do not store the artificial rows (they are bitwise copies of the
                             slack rows during phase 1); 0: reference layout with 1+n+2m rows        */
    int skip_zero_rows;   /* 1 (default): rows whose pivot-constraint entry a_pr is exactly 0 are neither read nor
                             written by the rank-1 update -- fma(s_i, 0, x) == x (src/solver.cu:43), so results are
                             value-identical; 0: stream every stored row (the nominal 2*R*m*sizeof bytes per pivot) */
    int use_graph;        /* 1: replay each batch of pivots as a CUDA graph                          */
    int batch;            /* pivots enqueued between two host status polls; 0 = choose from size     */
    long long max_pivots; /* total pivot cap for b2s_solve_two_phase; <= 0 = none (as the reference) */
    long long trace_capacity; /* (q,p) pairs kept on the device for b2s_copy_trace; 0 = default 1<<20 */
    int update_variant;   /* rank-1 update kernel variant (default 8: 256-bit accesses, 8 rows in flight,
                             device-wide ticket scheduler over 8-row tiles); others exist for tuning  */
    int persistent;       /* 1: run each batch of pivots as ONE persistent cooperative kernel with device-wide
                             barriers between the phases; 0: three launches per pivot (CUDA graph);
                             2 (default): the loop kernel for tableaux below 32 MB on one GPU, launches otherwise */
    int relative_infeasibility; /* 0 (default): the reference's absolute phase-1 test cost[0] <= -1e-9
                             (src/twoPhaseMethod.cu:265-268); 1: tolerance relative to the magnitude the phase-1
                             objective started from (the absolute test mis-declares large feasible LPs infeasible) */
    int lookahead;        /* 2 (default) / 1: ONE launch per pivot -- helper CTAs of the streaming update select the NEXT pivot
                             (entering column, ratio test, pivot-constraint gather; sharded: both exchanges over NVLink peer
                             memory) under the stream, replacing the reference's serial chain src/solver.cu:86-105;
                             0: three launches per pivot (ratio, gather, update).  Results are bit-identical.            */
    int fp64_polish;      /* fp32 solves only (the reference has none): 1 (default) = after phase 2, x_B = B^-1 b and the objective
                             c_B.x_B are recomputed in fp64 for the final basis by iterative refinement against the fp64 problem
                             data, with the fp32 tableau's slack block as approximate inverse; 0 = report the fp32 tableau's own
                             values.  Single-GPU solves.                                                                     */
    int drive_out_artificials; /* 0 (default): the reference's behaviour -- DEGENERATE when an artificial variable is still basic after
                             a feasible phase 1 (src/twoPhaseMethod.cu:206-223, :274-282); 1: pivot such artificials out (lowest-index
                             structural/slack column with |a| >= 1e-9 in that constraint; a constraint with none is redundant and
                             keeps its artificial at zero) and carry on into phase 2.  Single-GPU solves.                    */
    int reserved[2];
} b2s_options;

typedef struct b2s_stats {
    long long pivots_phase1;
    long long pivots_phase2;
    unsigned long long trace_hash; /* FNV-1a over the (q,p) int32 pairs of every pivot            */
    double seconds_total;          /* host wall time of the call                                   */
    double seconds_load;           /* host->device transfer + tableau build                        */
    double seconds_phase1;         /* device time, price-out + pivots of phase 1                   */
    double seconds_phase2;
    long long rows_streamed;       /* sum over pivots of tableau rows actually read+written        */
    long long rows_total;          /* sum over pivots of stored tableau rows                       */
    long long reserved[4];
} b2s_stats;

typedef struct b2s_solver b2s_solver; /* opaque */

void b2s_default_options(b2s_options *opt);
int b2s_create(const b2s_options *opt, b2s_solver **out);
void b2s_destroy(b2s_solver *s);
const char *b2s_last_error(const b2s_solver *s); /* s may be NULL: last error of the calling thread */
int b2s_device_count(void);

/* ---- problem input ------------------------------------------------------------------------- */
/* Replaces the host->device path of fillTableu (src/twoPhaseMethod.cu:145-200) for a problem_t
 * (include/problem.h:10-26): A is variable-major, A[j*m+i] = coefficient of variable j in
 * constraint i; b[m]; c[n].  The arrays are pageable host memory and are only read. */
int b2s_load_problem_host(b2s_solver *s, int n, int m, const double *A, const double *b, const double *c);

/* Replaces generateRandomProblem (src/problem.cu:49-126) + generator kernels (src/generator.cu:9-32):
 * the LP is generated on the device straight into the tableau, bit-identical to what the reference
 * generates for the same three kernel seeds (b, c, A in that order).  64-bit stream offsets, so
 * n*m may exceed 2^31 (the reference overflows int there, src/generator.cu:15). */
int b2s_generate_problem_device(b2s_solver *s, int n, int m, const unsigned seeds[3], double lo, double hi);
/* srand(seed); rand() x3 of src/problem.cu:63-67 for either C library. */
void b2s_seed_triplet(unsigned seed, int rand_flavour, unsigned out[3]);
/* Copy the current problem (as loaded or generated) back to host arrays; any pointer may be NULL. */
int b2s_copy_problem(b2s_solver *s, double *A, double *b, double *c);

/* ---- whole solve: twoPhaseMethod (src/twoPhaseMethod.cu:385-435) --------------------------- */
/* Returns a b2s error code; *status receives FEASIBLE/INFEASIBLE/UNBOUNDED/DEGENERATE (or
 * ITER_LIMIT).  x[n], *objective and basis[m] (0-based variable ids, src/solver.cu:105) are
 * written when *status == B2S_FEASIBLE, like the reference writes solution/optimalValue;
 * basis and stats may be NULL. */
int b2s_solve_two_phase(b2s_solver *s, int *status, double *x, double *objective, int *basis, b2s_stats *stats);

/* ---- stepping API (parity tests, benchmarks) --------------------------------------------------- */
int b2s_build_phase1(b2s_solver *s);  /* fillTableu, src/twoPhaseMethod.cu:145-200                  */
int b2s_price_out(b2s_solver *s);     /* updateObjectiveFunction, src/gaussian.cu:132-162           */
int b2s_select_entering(b2s_solver *s); /* minElement(costs+1,...), src/solver.cu:86 (first pivot)  */
/* Up to max_pivots iterations of src/solver.cu:78-126 (max_pivots < 0: until the phase ends).
 * *status: B2S_RUNNING when the budget ran out first, else FEASIBLE (phase optimal) / UNBOUNDED. */
int b2s_iterate(b2s_solver *s, long long max_pivots, int *status, long long *pivots_done);
int b2s_phase1_verdict(b2s_solver *s, int *status); /* src/twoPhaseMethod.cu:258-282             */
int b2s_switch_phase2(b2s_solver *s);               /* src/twoPhaseMethod.cu:285-318: drop artificials, load -c;
                                                       follow with b2s_price_out + b2s_select_entering        */
int b2s_extract_solution(b2s_solver *s, double *x, double *objective); /* :370-383               */

/* ---- introspection ------------------------------------------------------------------------ */
/* rows_active = reference `tabular->rows` (1+n+2m in phase 1, 1+n+m in phase 2); rows_stored =
 * rows resident in HBM (differs when artificials are folded); ld = row pitch in elements. */
int b2s_get_dims(const b2s_solver *s, int *n, int *m, long long *rows_active, long long *rows_stored, long long *ld);
/* Dense copy of the active tableau in the reference's unfolded layout: rows_active x m. */
int b2s_copy_tableau(b2s_solver *s, double *out);
int b2s_copy_costs(b2s_solver *s, double *out); /* rows_active values, [0] = objective */
int b2s_copy_basis(b2s_solver *s, int *out);    /* m values */
int b2s_copy_trace(b2s_solver *s, int *qp_pairs, long long capacity, long long *length, unsigned long long *hash);
int b2s_get_stats(b2s_solver *s, b2s_stats *stats);

/* ---- kernel-level hooks (micro-benchmarks, unit parity) ------------------------------------ */
/* minElement(vec, N, &idx) of src/reduction.cu:82-104 on a host vector: returns value and index. */
int b2s_tournament(b2s_solver *s, const double *vec, long long n, double *value, int *index);
/* Time `launches` launches of the fused rank-1 update (+cost update + next-pivot partial argmin)
 * on the current tableau with a synthetic but representative pivot (dense s and pivot-row
 * vectors); the tableau contents are destroyed.  ms_each[launches] receives per-launch CUDA
 * event times measured on the solver's stream; flush_l2 != 0 rewrites a >L2 buffer between
 * launches.  *bytes_per_launch = 2 * rows_stored * m * sizeof(real). */
int b2s_bench_update(b2s_solver *s, int launches, int flush_l2, float *ms_each, double *bytes_per_launch);

/* ---- caller-owned tableau: the reference's tabular_t level (include/tabular.cuh:5-30) ------------ */
/* Point the solver at a device tableau the caller allocated in the reference layout (pitched rows,
 * rows_active x cols, separate cost vector of rows_active entries; no folding).  This is what
 * `int solve(tabular_t*, int* base)` (include/solver.h:26) and `updateObjectiveFunction(tabular_t*,
 * int*)` (include/gaussian.cuh:5) are implemented with: attach, b2s_set_basis, then
 * b2s_price_out / b2s_select_entering + b2s_iterate, then b2s_copy_basis.  fp64 solvers only. */
int b2s_attach_tableau_device(b2s_solver *s, double *table, size_t pitch_bytes, int rows_active, int cols,
                              double *costs, int n_vars);
int b2s_set_basis(b2s_solver *s, const int *basis_host); /* m entries */
/* minElement(g_vet,size,&idx) (include/reduction.cuh:12), minElement(knownTerms,rowPivot,size,&idx)
 * (:14) and isLessOrEqualThanZero (:23) on DEVICE vectors, reference semantics bit for bit. */
int b2s_min_element_device(b2s_solver *s, const double *dvec, long long n, double *value, unsigned *index);
int b2s_ratio_min_device(b2s_solver *s, const double *known_terms, const double *column, long long n, double *value,
                         unsigned *index);
int b2s_max_le_zero_device(b2s_solver *s, const double *dvec, long long n, int *result);

/* `count` real pivots of the current phase launched one kernel at a time with CUDA events between
 * the three launches of each pivot (ratio test / gather+normalise / fused rank-1 update); the
 * arrays receive per-pivot kernel times in ms.  This is how bench.py measures the update kernel's
 * share and its achieved HBM bandwidth on live data. */
int b2s_profile_pivots(b2s_solver *s, int count, float *ms_ratio, float *ms_gather, float *ms_update,
                       long long *pivots_done);

/* How the pivot loop of the current problem runs: *launches_per_pivot (1 = look-ahead kernel, 3 / 4 = separate ratio, gather,
 * (svec,) update launches, 0 = persistent cooperative loop kernel), *lookahead, *persistent (booleans). */
int b2s_get_loop_info(b2s_solver *s, int *launches_per_pivot, int *lookahead, int *persistent);

/* Look-ahead kernel only: `count` real pivots launched one at a time; kernel_ms[count] = CUDA-event time of each launch,
 * stage_us[count*6] = microseconds from kernel start (globaltimer, helper CTA 0 of this rank) to: RHS row done, next entering
 * variable known, next leaving constraint known (sharded: after exchange 1), next pivot constraint complete on this rank
 * (sharded: after exchange 2), proposal complete, pivot committed (last CTA left).  Works on sharded solvers: every rank calls
 * it with the same count. */
int b2s_profile_lookahead(b2s_solver *s, int count, float *kernel_ms, double *stage_us, long long *pivots_done);

/* ---- sharded solves (constraint slabs over the GPUs of one box) ------------------------------ */
#define B2S_NCCL_ID_BYTES 128
int b2s_dist_unique_id(char id[B2S_NCCL_ID_BYTES]); /* rank 0 creates, the host layer broadcasts */
/* One process per GPU; call after b2s_create and before loading/generating the problem.  Rank r
 * owns constraints [r*m/world, (r+1)*m/world) of every tableau row; m/world must be a multiple of 512. */
int b2s_dist_init(b2s_solver *s, int rank, int world, const char id[B2S_NCCL_ID_BYTES]);

/* The same without NCCL: the ranks are bootstrapped through the caller's own all-gather (MPI, torch.distributed/gloo, sockets
 * ...).  `allgather(user, send, recv, bytes)` must deliver every rank's `bytes` bytes, in rank order, into recv (world * bytes,
 * host memory) on every rank and return 0; the library calls it, collectively and in the same order on all ranks, for the
 * CUDA-IPC handles of the peer-memory arenas, for barriers, for the price-out chain and for the solution vector.  The per-pivot
 * exchanges are peer-memory kernels (NVLink between GPUs; plain device memory when several ranks share one GPU, which is how
 * the driver's single-GPU test tier runs the sharded kernels).  At most 8 ranks. */
typedef int (*b2s_allgather_fn)(void *user, const void *send, void *recv, size_t bytes);
int b2s_dist_init_host(b2s_solver *s, int rank, int world, b2s_allgather_fn allgather, void *user);

#ifdef __cplusplus
}
#endif
#endif /* B2S_H */
