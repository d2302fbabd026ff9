/* examples/solve_file.c -- a plain C99 client of the C ABI (include/b2s.h): read an LP in the reference's text
 * format (n m / c[n] / m rows "a_i1 .. a_in b_i", README.MD:46-53 of the reference), solve it on GPU 0, print the
 * status, the optimum and the basis.
 *   gcc -std=c99 -Iinclude examples/solve_file.c -Lsimplexoncuda_b200/lib -lb2s -Wl,-rpath,$PWD/simplexoncuda_b200/lib
 */
#include <stdio.h>
#include <stdlib.h>

#include "b2s.h"

int main(int argc, char **argv)
{
    int n, m, i, j, status = B2S_RUNNING, rc;
    double *A, *b, *c, *x, objective = 0.0;
    int *basis;
    b2s_options opt;
    b2s_solver *solver = NULL;
    b2s_stats stats;
    FILE *f;

    if (argc < 2 || !(f = fopen(argv[1], "r"))) {
        fprintf(stderr, "usage: %s <lp.txt>\n", argv[0]);
        return 2;
    }
    if (fscanf(f, "%d %d", &n, &m) != 2) return 2;
    A = (double *)malloc(sizeof(double) * (size_t)n * (size_t)m); /* variable-major: A[j*m + i] */
    b = (double *)malloc(sizeof(double) * (size_t)m);
    c = (double *)malloc(sizeof(double) * (size_t)n);
    x = (double *)malloc(sizeof(double) * (size_t)n);
    basis = (int *)malloc(sizeof(int) * (size_t)m);
    for (j = 0; j < n; ++j)
        if (fscanf(f, "%lf", &c[j]) != 1) return 2;
    for (i = 0; i < m; ++i) {
        for (j = 0; j < n; ++j)
            if (fscanf(f, "%lf", &A[(size_t)j * m + i]) != 1) return 2;
        if (fscanf(f, "%lf", &b[i]) != 1) return 2;
    }
    fclose(f);

    b2s_default_options(&opt);
    if ((rc = b2s_create(&opt, &solver)) != B2S_OK) {
        fprintf(stderr, "b2s_create: %s\n", b2s_last_error(NULL));
        return 1;
    }
    if ((rc = b2s_load_problem_host(solver, n, m, A, b, c)) != B2S_OK ||
        (rc = b2s_solve_two_phase(solver, &status, x, &objective, basis, &stats)) != B2S_OK) {
        fprintf(stderr, "b2s error %d: %s\n", rc, b2s_last_error(solver));
        return 1;
    }
    printf("status %d  pivots %lld+%lld\n", status, stats.pivots_phase1, stats.pivots_phase2);
    if (status == B2S_FEASIBLE) {
        printf("optimal value %.17g\nx =", objective);
        for (j = 0; j < n; ++j) printf(" %.17g", x[j]);
        printf("\nbasis =");
        for (i = 0; i < m; ++i) printf(" %d", basis[i]);
        printf("\n");
    }
    b2s_destroy(solver);
    free(A); free(b); free(c); free(x); free(basis);
    return 0;
}
