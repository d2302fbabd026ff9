// oracle/ref_shim.cu -- TEST INFRASTRUCTURE.  C-linkage entry points around the UNMODIFIED
// reference (rik1599/SimplexOnCuda), compiled from the sources where they lie under
// /root/reference by oracle/build_ref.sh into oracle/_ref/libsimplex_ref.so.  No reference source
// is copied into this repository.
//
// The reference offers no pivot trace and no per-phase pivot count.  To observe them without
// editing its files, this translation unit compiles the reference's src/solver.cu *textually*
// (#include from the reference tree) with its two `solve` overloads renamed, and provides the
// public `int solve(tabular_t*, int*)` itself: the same 6-line driver loop as
// src/solver.cu:128-149 around the reference's own per-iteration function, plus bookkeeping
// (iteration count, wall time, optional basis diff to recover (q,p)).  Everything else
// (twoPhaseMethod, tableau build, reductions, price-out, update kernels) is the reference's code,
// linked unmodified.
#include <chrono>
#include <cstring>
#include <vector>

#define solve ref_stock_solve
#include "src/solver.cu"  // resolved through -I/root/reference
#undef solve

#include "problem.h"
#include "twoPhaseMethod.h"

namespace {
struct Recorder {
    bool trace_on = false;
    std::vector<int> trace;       // q,p pairs
    std::vector<int> basis;       // basis after the last solve() call
    long long pivots[2] = {0, 0};
    double loop_seconds[2] = {0, 0};
    int calls = 0;
} g_rec;
}  // namespace

// Replacement for src/solver.cu:128-149 (same resources, same loop).
int solve(tabular_t* tabular, int* base)
{
    TYPE *rowPivot, *colPivot;
    HANDLE_ERROR(cudaMalloc((void**)&rowPivot, BYTE_SIZE(tabular->cols)));
    HANDLE_ERROR(cudaMalloc((void**)&colPivot, BYTE_SIZE(tabular->rows)));
    cudaStream_t streams[2];
    for (size_t i = 0; i < 2; i++) HANDLE_ERROR(cudaStreamCreate(streams + i));

    const int m = tabular->cols;
    const int phase = g_rec.calls < 2 ? g_rec.calls : 1;
    std::vector<int> prev;
    if (g_rec.trace_on) prev.assign(base, base + m);
    HANDLE_ERROR(cudaDeviceSynchronize());
    const auto t0 = std::chrono::steady_clock::now();
    int status;
    while ((status = ref_stock_solve(tabular, base, rowPivot, colPivot, streams)) == NOT_ENDED) {
        g_rec.pivots[phase]++;
        if (g_rec.trace_on) {
            int p = -1;
            for (int i = 0; i < m; ++i)
                if (base[i] != prev[i]) {
                    p = i;
                    break;
                }
            g_rec.trace.push_back(p >= 0 ? base[p] : -1);
            g_rec.trace.push_back(p);
            if (p >= 0) prev[p] = base[p];
        }
    }
    HANDLE_ERROR(cudaDeviceSynchronize());
    g_rec.loop_seconds[phase] += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    g_rec.basis.assign(base, base + m);
    g_rec.calls++;

    HANDLE_ERROR(cudaFree(rowPivot));
    HANDLE_ERROR(cudaFree(colPivot));
    for (size_t i = 0; i < 2; i++) HANDLE_ERROR(cudaStreamDestroy(streams[i]));
    return status;
}

extern "C" {

// main.cu:117-133
int ref_setup_device(void)
{
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, 0) != cudaSuccess) return -1;
    cudaSetDeviceFlags(cudaDeviceMapHost);
    return 0;
}

// twoPhaseMethod on caller-supplied arrays (problem_t layout, include/problem.h:10-26).
// seconds[0] = wall time of twoPhaseMethod, seconds[1..2] = time inside the pivot loops.
int ref_two_phase(int n, int m, const double* A, const double* b, const double* c, double* x, double* obj,
                  int* basis_out, int trace_on, int* trace_out, long long trace_cap, long long* pivots_out,
                  double* seconds)
{
    problem_t P;
    P.vars = n;
    P.constraints = m;
    P.constraintsMatrix = (TYPE*)malloc(sizeof(TYPE) * (size_t)n * m);
    P.knownTermsVector = (TYPE*)malloc(sizeof(TYPE) * m);
    P.objectiveFunction = (TYPE*)malloc(sizeof(TYPE) * n);
    memcpy(P.constraintsMatrix, A, sizeof(TYPE) * (size_t)n * m);
    memcpy(P.knownTermsVector, b, sizeof(TYPE) * m);
    memcpy(P.objectiveFunction, c, sizeof(TYPE) * n);
    g_rec = Recorder();
    g_rec.trace_on = trace_on != 0;
    TYPE* sol = (TYPE*)malloc(sizeof(TYPE) * n);
    TYPE opt = 0;
    const auto t0 = std::chrono::steady_clock::now();
    const int status = twoPhaseMethod(&P, sol, &opt);
    cudaDeviceSynchronize();
    const double total = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (status == FEASIBLE) {
        if (x) memcpy(x, sol, sizeof(TYPE) * n);
        if (obj) *obj = opt;
    }
    if (basis_out && (int)g_rec.basis.size() == m) memcpy(basis_out, g_rec.basis.data(), sizeof(int) * m);
    if (pivots_out) {
        pivots_out[0] = g_rec.pivots[0];
        pivots_out[1] = g_rec.pivots[1];
    }
    if (trace_out) {
        const long long cnt = std::min<long long>((long long)g_rec.trace.size() / 2, trace_cap);
        memcpy(trace_out, g_rec.trace.data(), sizeof(int) * 2 * (size_t)cnt);
    }
    if (seconds) {
        seconds[0] = total;
        seconds[1] = g_rec.loop_seconds[0];
        seconds[2] = g_rec.loop_seconds[1];
    }
    free(sol);
    freeProblem(&P);
    return status;
}

// generateRandomProblem (src/problem.cu:49-126) -> caller arrays.  Seeds come from the C library's
// rand() exactly as in the reference.
int ref_generate(int n, int m, unsigned seed, int lo, int hi, double* A, double* b, double* c)
{
    problem_t* P = generateRandomProblem(n, m, seed, lo, hi);
    memcpy(A, P->constraintsMatrix, sizeof(TYPE) * (size_t)n * m);
    memcpy(b, P->knownTermsVector, sizeof(TYPE) * m);
    memcpy(c, P->objectiveFunction, sizeof(TYPE) * n);
    freeProblem(P);
    free(P);
    return 0;
}

}  // extern "C"
