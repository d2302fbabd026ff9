/*
 * oracle/serial_tableau.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Serial host restatement (plain C) of the dense-tableau two-phase simplex that
 * rik1599/SimplexOnCuda runs on the GPU.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load this; the product path
 * (simplexoncuda_b200/csrc) never links or calls it.
 *
 * Parity status: PINNED.  The restatement reproduces the reference's published per-phase
 * pivot counts (data/measures/<gpu>/benchmark_<n>_<m>.txt, lines-1 per phase) on ALL 36 published
 * instances plus the false-INFEASIBLE 37th (tests/golden/oracle_results.json) when the three
 * kernel seeds are derived with MSVC rand() -- see tests/test_oracle_golden.py -- and it is
 * cross-checked on the GPU box against the unmodified reference build in oracle/_ref
 * (tests/test_reference_parity.py).
 *
 * Every function cites the reference file:line whose arithmetic it restates.  Build with
 * -ffp-contract=off: a fused multiply-add is used only where the reference's SASS has one
 * (explicit fma() calls below).
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_FEASIBLE 0
#define ORC_INFEASIBLE (-1)
#define ORC_UNBOUNDED (-2)
#define ORC_DEGENERATE (-3)
#define ORC_ITER_LIMIT (-4) /* not a reference status: the reference has no iteration cap */
#define ORC_CONTINUE (-10)

#define RULE_REFERENCE 0 /* epsilon tournament, reference tie order           */
#define RULE_LOWEST 1    /* exact Dantzig minimum, lowest index among ties     */
#define RULE_BLAND 2     /* Bland: first improving column, lowest basic var    */

/* include/macro.h:28-42 -- absolute-epsilon three-way compare. */
static inline int cmp3(double x, double y)
{
    if (fabs(x - y) < 1e-9)
        return 0;
    return x < y ? -1 : 1;
}

/* ------------------------------------------------------------------------------------------
 * Tournament argmin.  src/reduction.cu:10-22 (warp tree), :24-49 (block tree), :51-80 (grid
 * stride scan + per-block winner), :82-104 (two launches: ceil(N/512) blocks of 512 threads,
 * then one block of 1024 threads when there is more than one block).
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    double v;
    int i;
} cand_t;

/* src/reduction.cu:10-22: offsets 16,8,4,2,1; lane L reads lane L+off (its own value when
 * L+off >= 32, which is what shfl_down returns out of range) and takes it only when strictly
 * smaller by >= eps.  All lanes step together from the previous level's values. */
static void warp_tree(cand_t w[32])
{
    for (int off = 16; off > 0; off >>= 1) {
        cand_t nxt[32];
        for (int l = 0; l < 32; ++l) {
            cand_t other = (l + off < 32) ? w[l + off] : w[l];
            nxt[l] = (cmp3(other.v, w[l].v) < 0) ? other : w[l];
        }
        memcpy(w, nxt, sizeof(nxt));
    }
}

/* src/reduction.cu:24-49: per-warp winners go to shared slots; warp 0 reloads slot[lane] for
 * lane < nthreads/32 (others hold (DBL_MAX,-1)) and runs the same tree. */
static cand_t block_tree(const cand_t *thr, int nthreads)
{
    cand_t slots[32];
    int nw = nthreads / 32;
    for (int w = 0; w < nw; ++w) {
        cand_t lanes[32];
        memcpy(lanes, thr + 32 * w, sizeof(lanes));
        warp_tree(lanes);
        slots[w] = lanes[0];
    }
    cand_t lanes[32];
    for (int l = 0; l < 32; ++l) {
        if (l < nw)
            lanes[l] = slots[l];
        else {
            lanes[l].v = DBL_MAX;
            lanes[l].i = -1;
        }
    }
    warp_tree(lanes);
    return lanes[0];
}

/* src/reduction.cu:82-104.  Returns the tournament value, *idx the winning position (or -1). */
static double tournament(const double *vec, long N, int *idx)
{
    long G = (N + 511) / 512;
    if (G > 1024)
        G = 1024;
    if (G < 1)
        G = 1;
    cand_t *stage = (cand_t *)malloc(sizeof(cand_t) * (size_t)G);
    cand_t thr[1024];
    for (long b = 0; b < G; ++b) {
        for (int t = 0; t < 512; ++t) {
            cand_t c = {DBL_MAX, -1};
            for (long i = b * 512 + t; i < N; i += 512 * G) { /* :56-71 */
                if (cmp3(vec[i], c.v) < 0) {
                    c.v = vec[i];
                    c.i = (int)i;
                }
            }
            thr[t] = c;
        }
        stage[b] = block_tree(thr, 512);
    }
    cand_t win = stage[0];
    if (G > 1) { /* :92-93: second launch, 1 block x 1024 threads over the G block winners */
        for (int t = 0; t < 1024; ++t) {
            cand_t c = {DBL_MAX, -1};
            for (long i = t; i < G; i += 1024) {
                if (cmp3(stage[i].v, c.v) < 0)
                    c = stage[i];
            }
            thr[t] = c;
        }
        win = block_tree(thr, 1024);
    }
    free(stage);
    *idx = win.i;
    return win.v;
}

/* ------------------------------------------------------------------------------------------
 * Solver state.  include/tabular.cuh:5-30, src/tabular.cu:25-39: device row r holds one tableau
 * column (row 0 = RHS b, rows 1..n structural, then m slack rows, then m artificial rows), m
 * contiguous doubles each; the reduced costs live in a separate vector whose element 0 is the
 * objective value.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int n, m;
    long R;  /* active rows: 1+n+2m in phase 1, 1+n+m in phase 2 */
    long R1; /* allocated rows */
    double *T;
    double *cost;
    int *base;
    double *col, *rowp, *s, *ratio; /* scratch */
    const double *A, *b, *c;
    int rule;
    int threads;
    long pivots[2];
    int phase; /* 0 before build, 1, 2 */
    int *trace;
    long trace_cap, trace_len;
    uint64_t hash;
    int last_q, last_p;
    double last_cq;
    int relative_infeasibility; /* opt-in, not reference behaviour: see orc_phase1_verdict */
    int drive_out;              /* opt-in, not reference behaviour: see orc_drive_out_artificials */
    double cost0_phase1_start;  /* cost[0] right after the phase-1 price-out = -(sum of |b_i|) */
} orc_t;

orc_t *orc_create(int n, int m, const double *A, const double *b, const double *c, int rule, int threads)
{
    orc_t *o = (orc_t *)calloc(1, sizeof(orc_t));
    o->n = n;
    o->m = m;
    o->R1 = 1 + (long)n + 2 * (long)m;
    o->R = o->R1;
    o->T = (double *)malloc(sizeof(double) * (size_t)o->R1 * (size_t)m);
    o->cost = (double *)malloc(sizeof(double) * (size_t)o->R1);
    o->base = (int *)malloc(sizeof(int) * (size_t)m);
    o->col = (double *)malloc(sizeof(double) * (size_t)m);
    o->s = (double *)malloc(sizeof(double) * (size_t)m);
    o->ratio = (double *)malloc(sizeof(double) * (size_t)m);
    o->rowp = (double *)malloc(sizeof(double) * (size_t)o->R1);
    o->A = A;
    o->b = b;
    o->c = c;
    o->rule = rule;
    o->threads = threads > 0 ? threads : 1;
    o->hash = 1469598103934665603ULL;
    o->last_q = o->last_p = -1;
    return o;
}

void orc_destroy(orc_t *o)
{
    if (!o)
        return;
    free(o->T);
    free(o->cost);
    free(o->base);
    free(o->col);
    free(o->s);
    free(o->ratio);
    free(o->rowp);
    free(o);
}

void orc_set_trace(orc_t *o, int *buf, long cap)
{
    o->trace = buf;
    o->trace_cap = cap;
    o->trace_len = 0;
}

/* src/twoPhaseMethod.cu:145-200 (fillTableu) and :86-111 (checkColumns/negateColumn).  The
 * reference relies on fresh allocations being zero for the off-diagonal identity entries. */
void orc_build_phase1(orc_t *o)
{
    const int n = o->n, m = o->m;
    const long R = o->R1;
    o->R = R;
    o->phase = 1;
    memset(o->T, 0, sizeof(double) * (size_t)R * (size_t)m);
    for (long r = 0; r < R; ++r)
        o->cost[r] = (r > (long)n + m) ? 1.0 : 0.0; /* :150-156 */
    for (int j = 0; j < n; ++j)                     /* :159-169 rows 1..n <- A (variable-major) */
        memcpy(o->T + (size_t)(1 + j) * m, o->A + (size_t)j * m, sizeof(double) * (size_t)m);
    for (int i = 0; i < m; ++i) {
        o->T[(size_t)(1 + n + i) * m + i] = 1.0;     /* :30-39 slack identity      */
        o->T[(size_t)(1 + n + m + i) * m + i] = 1.0; /* artificial identity        */
        o->T[i] = o->b[i];                           /* :178-184 row 0 <- b        */
        o->base[i] = n + m + i;                      /* :44-52, :187-190           */
    }
    for (int i = 0; i < m; ++i) /* :100-111: negate every constraint whose RHS <= -1e-9, all rows */
        if (cmp3(o->T[i], 0.0) < 0)
            for (long r = 0; r < R; ++r)
                o->T[(size_t)r * m + i] = -o->T[(size_t)r * m + i];
}

/* src/gaussian.cu:119-127 (coef), :13-20 + :98-107 (pair terms, fp64 atomicAdd of -term), launch
 * :142-158 (block 32x32, grid.x = 1 => x = lane + 64k).  nvcc contracts a*b + c*d into
 * DMUL(c,d) then DFMA(a,b,.) (checked in the sm_100 SASS of the reference build).  The atomic
 * order is unspecified in the reference; this restatement fixes it to ascending x. */
void orc_priceout(orc_t *o)
{
    const int m = o->m;
    double *coef = o->s;
    for (int i = 0; i < m; ++i)
        coef[i] = (1 + (long)o->base[i] < o->R) ? o->cost[1 + o->base[i]] : 0.0; /* (a redundant constraint's artificial, drive-out mode) */
    for (long y = 0; y < o->R; ++y) {
        const double *row = o->T + (size_t)y * m;
        double acc = o->cost[y];
        for (int k0 = 0; k0 < m; k0 += 64) {
            for (int l = 0; l < 32; ++l) {
                int x = k0 + l;
                if (x >= m)
                    break;
                double term;
                if (x + 32 < m)
                    term = fma(row[x], coef[x], row[x + 32] * coef[x + 32]);
                else
                    term = row[x] * coef[x];
                acc = acc + (-term);
            }
        }
        o->cost[y] = acc;
    }
    if (o->phase == 1)
        o->cost0_phase1_start = o->cost[0];
}

void orc_set_relative_infeasibility(orc_t *o, int on) { o->relative_infeasibility = on; }

static void trace_push(orc_t *o, int q, int p)
{
    if (o->trace && o->trace_len < o->trace_cap) {
        o->trace[2 * o->trace_len] = q;
        o->trace[2 * o->trace_len + 1] = p;
    }
    o->trace_len++;
    uint32_t w[2] = {(uint32_t)q, (uint32_t)p};
    const unsigned char *bytes = (const unsigned char *)w;
    for (int k = 0; k < 8; ++k) {
        o->hash ^= bytes[k];
        o->hash *= 1099511628211ULL;
    }
}

/* The pivot itself once (q, p) is chosen: src/solver.cu:105 (basis), :24-32/:62-66 (gather), :34-46 (rank-1 update),
 * :48-56 (cost update).  o->col holds the entering column. */
static void apply_pivot(orc_t *o, int q, int p, double cq)
{
    const int m = o->m;
    const long R = o->R;
    o->base[p] = q; /* src/solver.cu:105 */

    /* --- gather pivot constraint (raw), src/solver.cu:24-32, :62-66 -------------------------- */
    for (long r = 0; r < R; ++r)
        o->rowp[r] = o->T[(size_t)r * m + p];
    const double piv = o->col[p];
    for (int i = 0; i < m; ++i)
        o->s[i] = (-o->col[i]) / piv; /* :43 quotient rounded before the FMA */
    const double sc = (-cq) / piv;    /* :54 */

    /* --- rank-1 update :34-46 and cost update :48-56 --------------------------------------- */
    const double *s = o->s;
    double *T = o->T;
    const double *rowp = o->rowp;
#ifdef _OPENMP
#pragma omp parallel for num_threads(o->threads) schedule(static)
#endif
    for (long r = 0; r < R; ++r) {
        double *row = T + (size_t)r * m;
        const double a = rowp[r];
        const double keep = row[p] / piv;
        for (int i = 0; i < m; ++i)
            row[i] = fma(s[i], a, row[i]);
        row[p] = keep;
    }
    for (long r = 0; r < R; ++r)
        o->cost[r] = fma(sc, rowp[r], o->cost[r]);

    o->last_q = q;
    o->last_p = p;
    o->last_cq = cq;
    o->pivots[o->phase == 2 ? 1 : 0]++;
    trace_push(o, q, p);
}

/* One iteration.  src/solver.cu:78-126 (select, unbounded test, ratio test, base update) and
 * :58-75 / :24-56 (gather, rank-1 update, cost update). */
int orc_pivot(orc_t *o)
{
    const int m = o->m;
    const long R = o->R;
    int q = -1, p = -1;
    double cq;

    /* --- entering column: src/solver.cu:86-88 ------------------------------------------- */
    if (o->rule == RULE_REFERENCE) {
        cq = tournament(o->cost + 1, R - 1, &q);
    } else if (o->rule == RULE_LOWEST) {
        cq = DBL_MAX;
        for (long j = 0; j < R - 1; ++j)
            if (o->cost[1 + j] < cq) {
                cq = o->cost[1 + j];
                q = (int)j;
            }
    } else {
        cq = 0.0;
        for (long j = 0; j < R - 1; ++j)
            if (cmp3(o->cost[1 + j], 0.0) < 0) {
                cq = o->cost[1 + j];
                q = (int)j;
                break;
            }
    }
    if (!(cmp3(cq, 0.0) < 0) || q < 0)
        return ORC_FEASIBLE; /* :119-125 optimal for this phase */

    /* --- entering column copy :90-94, unbounded test src/reduction.cu:186-201 ------------ */
    const double *qrow = o->T + (size_t)(1 + q) * m;
    double mx = DBL_MIN;
    for (int i = 0; i < m; ++i) {
        o->col[i] = qrow[i];
        mx = fmax(mx, qrow[i]);
    }
    if (cmp3(mx, 0.0) <= 0)
        return ORC_UNBOUNDED;

    /* --- ratio test: src/reduction.cu:106-140 --------------------------------------------- */
    for (int i = 0; i < m; ++i)
        o->ratio[i] = (cmp3(o->col[i], 0.0) > 0) ? o->T[i] / o->col[i] : DBL_MAX;
    if (o->rule == RULE_REFERENCE) {
        tournament(o->ratio, m, &p);
    } else if (o->rule == RULE_LOWEST) {
        double best = DBL_MAX;
        for (int i = 0; i < m; ++i)
            if (o->ratio[i] < best) {
                best = o->ratio[i];
                p = i;
            }
    } else {
        double best = DBL_MAX;
        for (int i = 0; i < m; ++i)
            if (o->ratio[i] < best)
                best = o->ratio[i];
        int bestvar = INT32_MAX;
        for (int i = 0; i < m; ++i)
            if (o->ratio[i] == best && best < DBL_MAX && o->base[i] < bestvar) {
                bestvar = o->base[i];
                p = i;
            }
    }
    if (p < 0)
        return ORC_UNBOUNDED; /* cannot happen after the max test; defensive */
    apply_pivot(o, q, p, cq);
    return ORC_CONTINUE;
}

/* Run pivots until the phase ends or `budget` pivots were made (budget < 0: no cap). */
int orc_iterate(orc_t *o, long budget)
{
    int st = ORC_CONTINUE;
    while (budget != 0 && (st = orc_pivot(o)) == ORC_CONTINUE)
        if (budget > 0)
            --budget;
    return st;
}

/* Beyond the reference (opt-in): where the reference stops with DEGENERATE because an artificial variable is
 * still basic after a feasible phase 1 (src/twoPhaseMethod.cu:206-223, :274-282), pivot it out.  Constraints in ascending
 * order; for constraint i with an artificial basic variable take the LOWEST-index structural or slack variable j whose entry
 * in that constraint is non-zero by the reference's own threshold (|a_ij| >= 1e-9) and pivot on (j, i) -- a degenerate pivot,
 * the artificial sits at level 0.  A constraint with no such entry is redundant: its artificial stays basic at zero and can
 * never be chosen again (its row is zero on every column that survives into phase 2).  Returns the number of pivots made. */
long orc_drive_out_artificials(orc_t *o)
{
    const int n = o->n, m = o->m;
    long made = 0;
    for (int i = 0; i < m; ++i) {
        if (o->base[i] < n + m)
            continue;
        int q = -1;
        for (int j = 0; j < n + m; ++j)
            if (cmp3(fabs(o->T[(size_t)(1 + j) * m + i]), 0.0) > 0) {
                q = j;
                break;
            }
        if (q < 0)
            continue; /* redundant constraint */
        const double *qrow = o->T + (size_t)(1 + q) * m;
        for (int k = 0; k < m; ++k)
            o->col[k] = qrow[k];
        apply_pivot(o, q, i, o->cost[1 + q]);
        ++made;
    }
    return made;
}
void orc_set_drive_out(orc_t *o, int on) { o->drive_out = on; }

/* src/twoPhaseMethod.cu:258-282 (:265-268 infeasible test on cost[0], :206-223 degeneracy). */
int orc_phase1_verdict(const orc_t *o)
{
    /* Reference: absolute test cost[0] <= -1e-9 (:265-268).  Opt-in alternative (beyond the reference, SURVEY
     * 8(f)-4): the residual of a feasible phase 1 scales with the magnitude the objective started from, so the
     * tolerance is taken relative to it: infeasible iff cost[0] < -1e-9 * max(1, |cost[0] at phase-1 start|). */
    if (o->relative_infeasibility) {
        const double scale = fmax(1.0, fabs(o->cost0_phase1_start));
        if (o->cost[0] < -1e-9 * scale)
            return ORC_INFEASIBLE;
    } else if (cmp3(o->cost[0], 0.0) < 0)
        return ORC_INFEASIBLE;
    const int lo = o->n + o->m, hi = o->n + 2 * o->m;
    for (int i = 0; i < o->m; ++i)
        if (o->base[i] >= lo && o->base[i] < hi) {
            if (!o->drive_out)
                return ORC_DEGENERATE;
            /* opt-in mode, after orc_drive_out_artificials: an artificial may only remain in a redundant constraint */
            for (int j = 0; j < lo; ++j)
                if (cmp3(fabs(o->T[(size_t)(1 + j) * o->m + i]), 0.0) > 0)
                    return ORC_DEGENERATE;
        }
    return ORC_FEASIBLE;
}

/* src/twoPhaseMethod.cu:285-337: drop the trailing artificial rows, reload -c, keep cost[0]. */
void orc_switch_phase2(orc_t *o)
{
    const int n = o->n, m = o->m;
    o->R = 1 + (long)n + m;
    o->phase = 2;
    for (int i = 0; i < m; ++i)
        o->cost[1 + n + i] = 0.0;
    for (int j = 0; j < n; ++j)
        o->cost[1 + j] = -o->c[j];
}

/* src/twoPhaseMethod.cu:370-383. */
void orc_extract(const orc_t *o, double *x, double *obj)
{
    *obj = o->cost[0];
    for (int j = 0; j < o->n; ++j)
        x[j] = 0.0;
    for (int i = 0; i < o->m; ++i)
        if (o->base[i] < o->n)
            x[o->base[i]] = o->T[i];
}

/* src/twoPhaseMethod.cu:385-435 end to end.  max_pivots < 0: uncapped (as the reference). */
int orc_two_phase(orc_t *o, long max_pivots, double *x, double *obj)
{
    orc_build_phase1(o);
    orc_priceout(o);
    int st = orc_iterate(o, max_pivots);
    if (st == ORC_CONTINUE)
        return ORC_ITER_LIMIT;
    st = orc_phase1_verdict(o); /* the phase-1 solve status is ignored: :258 */
    if (st == ORC_DEGENERATE && o->drive_out) { /* opt-in: pivot the basic artificials out instead of stopping */
        orc_drive_out_artificials(o);
        st = orc_phase1_verdict(o);
    }
    if (st != ORC_FEASIBLE)
        return st;
    orc_switch_phase2(o);
    orc_priceout(o);
    long left = max_pivots < 0 ? -1 : max_pivots - o->pivots[0];
    st = orc_iterate(o, left);
    if (st == ORC_CONTINUE)
        return ORC_ITER_LIMIT;
    if (st != ORC_FEASIBLE)
        return st;
    orc_extract(o, x, obj);
    return ORC_FEASIBLE;
}

/* ---- accessors for the tests ------------------------------------------------------------- */
long orc_rows(const orc_t *o) { return o->R; }
long orc_pivots(const orc_t *o, int phase) { return o->pivots[phase == 2 ? 1 : 0]; }
long orc_trace_len(const orc_t *o) { return o->trace_len; }
uint64_t orc_hash(const orc_t *o) { return o->hash; }
const double *orc_tableau(const orc_t *o) { return o->T; }
const double *orc_costs(const orc_t *o) { return o->cost; }
const int *orc_basis(const orc_t *o) { return o->base; }
int orc_last_q(const orc_t *o) { return o->last_q; }
int orc_last_p(const orc_t *o) { return o->last_p; }
double orc_tournament(const double *v, long n, int *idx) { return tournament(v, n, idx); }
int orc_compare(double x, double y) { return cmp3(x, y); }

/* ------------------------------------------------------------------------------------------
 * Instance generator.  src/problem.cu:49-126 (three kernel seeds from srand/rand), src/generator.cu
 * :9-32 (XORWOW stream positions), cuRAND XORWOW (curand_kernel.h: state init from the seed,
 * xorshift step + Weyl counter 362437) and curand_uniform (curand_uniform.h:69-72).
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    uint32_t v[5], d;
} xorwow_t;

static void xorwow_seed(xorwow_t *s, uint64_t seed)
{
    uint32_t s0 = ((uint32_t)seed) ^ 0xaad26b49u;
    uint32_t s1 = (uint32_t)(seed >> 32) ^ 0xf7dcefddu;
    uint32_t t0 = 1099087573u * s0;
    uint32_t t1 = 2591861531u * s1;
    s->d = 6615241u + t1 + t0;
    s->v[0] = 123456789u + t0;
    s->v[1] = 362436069u ^ t0;
    s->v[2] = 521288629u + t1;
    s->v[3] = 88675123u ^ t1;
    s->v[4] = 5783321u + t0;
}

static inline uint32_t xorwow_next(xorwow_t *s)
{
    uint32_t t = s->v[0] ^ (s->v[0] >> 2);
    s->v[0] = s->v[1];
    s->v[1] = s->v[2];
    s->v[2] = s->v[3];
    s->v[3] = s->v[4];
    s->v[4] = (s->v[4] ^ (s->v[4] << 4)) ^ (t ^ (t << 1));
    s->d += 362437u;
    return s->v[4] + s->d;
}

/* curand_uniform then `u * (max - min) + min` in double, contracted by nvcc to one DFMA
 * (src/generator.cu:18, :30). */
static inline double draw_value(uint32_t x, double lo, double hi)
{
    float u = fmaf((float)x, 2.3283064e-10f, 2.3283064e-10f / 2.0f);
    return fma((double)u, hi - lo, lo);
}

/* flavour 0: glibc rand() (what the reference does when built on Linux: srand(seed); rand() x3)
 * flavour 1: MSVC rand()  (the published measurements were taken on Windows). */
void orc_seed_triplet(unsigned seed, int flavour, unsigned out[3])
{
    if (flavour == 1) {
        uint32_t st = seed;
        for (int k = 0; k < 3; ++k) {
            st = st * 214013u + 2531011u;
            out[k] = (st >> 16) & 0x7fff;
        }
        return;
    }
    /* glibc TYPE_3 additive feedback generator (r[i] = r[i-3] + r[i-31], output >> 1). */
    int32_t r[34 + 310 + 3];
    r[0] = (int32_t)(seed ? seed : 1);
    for (int i = 1; i < 31; ++i) {
        int64_t hi = r[i - 1] / 127773, lo = r[i - 1] % 127773;
        int64_t w = 16807 * lo - 2836 * hi;
        if (w < 0)
            w += 2147483647;
        r[i] = (int32_t)w;
    }
    for (int i = 31; i < 34; ++i)
        r[i] = r[i - 31];
    for (int i = 34; i < 344 + 3; ++i)
        r[i] = (int32_t)((uint32_t)r[i - 31] + (uint32_t)r[i - 3]);
    for (int k = 0; k < 3; ++k)
        out[k] = ((uint32_t)r[344 + k]) >> 1;
}

/* A is variable-major (A[j*m+i]); element (var j, cons i) is output #(i*n+j) of the stream
 * seeded with seeds[2]; b[i] / c[j] are output #i / #j of the streams seeded with seeds[0] /
 * seeds[1] (src/problem.cu:63-110: seedOne -> b, seedTwo -> c, seedThree -> A). */
void orc_generate(int n, int m, const unsigned seeds[3], double lo, double hi, double *A, double *b, double *c)
{
    xorwow_t s;
    xorwow_seed(&s, seeds[0]);
    for (int i = 0; i < m; ++i)
        b[i] = draw_value(xorwow_next(&s), lo, hi);
    xorwow_seed(&s, seeds[1]);
    for (int j = 0; j < n; ++j)
        c[j] = draw_value(xorwow_next(&s), lo, hi);
    xorwow_seed(&s, seeds[2]);
    for (int i = 0; i < m; ++i)
        for (int j = 0; j < n; ++j)
            A[(size_t)j * m + i] = draw_value(xorwow_next(&s), lo, hi);
}

/* Raw XORWOW outputs [offset, offset+count) of the stream for `seed` -- used to check the GPU
 * generator's jump-ahead. */
void orc_xorwow_outputs(uint64_t seed, uint64_t offset, long count, uint32_t *out)
{
    xorwow_t s;
    xorwow_seed(&s, seed);
    for (uint64_t k = 0; k < offset; ++k)
        xorwow_next(&s);
    for (long k = 0; k < count; ++k)
        out[k] = xorwow_next(&s);
}

/* ------------------------------------------------------------------------------------------
 * CLI: serial_tableau <n> <m> <seed> <lo> <hi> <flavour> <rule> <threads> <max_pivots>
 * Prints one line: status pivots1 pivots2 objective hash seconds_total seconds_in_pivot_loops.  Used by bench.py's
 * cpu_baseline leg (bounded pivot budget) and by hand.
 * ---------------------------------------------------------------------------------------- */
#ifdef ORC_MAIN
#include <time.h>
static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}
int main(int argc, char **argv)
{
    if (argc < 10) {
        fprintf(stderr, "usage: %s n m seed lo hi flavour rule threads max_pivots\n", argv[0]);
        return 2;
    }
    int n = atoi(argv[1]), m = atoi(argv[2]);
    unsigned seed = (unsigned)strtoul(argv[3], 0, 10);
    double lo = atof(argv[4]), hi = atof(argv[5]);
    int flavour = atoi(argv[6]), rule = atoi(argv[7]), threads = atoi(argv[8]);
    long maxp = atol(argv[9]);
    unsigned seeds[3];
    orc_seed_triplet(seed, flavour, seeds);
    double *A = malloc(sizeof(double) * (size_t)n * m), *b = malloc(sizeof(double) * m), *c = malloc(sizeof(double) * n);
    double *x = malloc(sizeof(double) * n), obj = 0;
    orc_generate(n, m, seeds, lo, hi, A, b, c);
    orc_t *o = orc_create(n, m, A, b, c, rule, threads);
    double t0 = now_s(), loop = 0.0, t;
    /* same sequence as orc_two_phase, with the pivot loops timed on their own */
    orc_build_phase1(o);
    orc_priceout(o);
    t = now_s();
    int st = orc_iterate(o, maxp);
    loop += now_s() - t;
    if (st == ORC_CONTINUE) {
        st = ORC_ITER_LIMIT;
    } else if ((st = orc_phase1_verdict(o)) == ORC_FEASIBLE) {
        orc_switch_phase2(o);
        orc_priceout(o);
        t = now_s();
        st = orc_iterate(o, maxp < 0 ? -1 : maxp - o->pivots[0]);
        loop += now_s() - t;
        if (st == ORC_CONTINUE)
            st = ORC_ITER_LIMIT;
        else if (st == ORC_FEASIBLE)
            orc_extract(o, x, &obj);
    }
    double t1 = now_s();
    printf("%d %ld %ld %.17g %llu %.6f %.6f\n", st, o->pivots[0], o->pivots[1], st == 0 ? obj : o->cost[0],
           (unsigned long long)o->hash, t1 - t0, loop);
    return 0;
}
#endif

/* ---- stage-level access to the tournament, for tests of the sharded protocol ---------------
 * orc_stage1_block: winner of reference stage-1 block `b` when the grid has G blocks
 * (src/reduction.cu:51-80).  orc_stage2: the second launch (:92-93) over G (value,index) slots. */
#ifndef ORC_MAIN
double orc_stage1_block(const double *vec, long N, long G, long b, int *idx)
{
    cand_t thr[512];
    for (int t = 0; t < 512; ++t) {
        cand_t c = {DBL_MAX, -1};
        for (long i = b * 512 + t; i < N; i += 512 * G)
            if (cmp3(vec[i], c.v) < 0) {
                c.v = vec[i];
                c.i = (int)i;
            }
        thr[t] = c;
    }
    cand_t w = block_tree(thr, 512);
    *idx = w.i;
    return w.v;
}

double orc_stage2(const double *slot_v, const int *slot_i, long G, int *idx)
{
    cand_t thr[1024];
    for (int t = 0; t < 1024; ++t) {
        cand_t c = {DBL_MAX, -1};
        for (long i = t; i < G; i += 1024)
            if (cmp3(slot_v[i], c.v) < 0) {
                c.v = slot_v[i];
                c.i = slot_i[i];
            }
        thr[t] = c;
    }
    cand_t w = block_tree(thr, 1024);
    *idx = w.i;
    return w.v;
}
#endif
