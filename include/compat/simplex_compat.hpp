// simplex_compat.hpp -- source-level drop-in for the public headers of rik1599/SimplexOnCuda.
//
// A program written against the reference's headers (its own main.cu is the test: see
// tests/test_dropin_main.py) compiles unchanged with -Iinclude/compat and links against
// libb2s_compat.so + libb2s.so.  Every declaration below has the name, argument meaning and
// status/ownership convention of the reference declaration it cites; the definitions
// (simplexoncuda_b200/csrc/compat.cu) are thin shims over the C ABI in include/b2s.h.
//
// The reference headers are one file each (problem.h, tabular.cuh, solver.h, twoPhaseMethod.h,
// reduction.cuh, gaussian.cuh, generator.cuh, macro.h, error.cuh, chrono.cuh); here each of those
// names is a one-line forwarder to this file.
#pragma once

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <cuda_runtime.h>

// ---- macro.h --------------------------------------------------------------------------------
#define TYPE double                          // reference include/macro.h:6
#define TYPE_SIZE sizeof(TYPE)               // :8
#define BYTE_SIZE(count) count * TYPE_SIZE   // :9 (unparenthesised on purpose: same expansion)

#if defined(__CUDACC__)
#define B2S_HD __host__ __device__
#else
#define B2S_HD
#endif

// pitched addressing helpers, reference include/macro.h:11-19
B2S_HD inline TYPE* ROW(TYPE* base, int row, size_t pitch) { return (TYPE*)((char*)base + (size_t)row * pitch); }
B2S_HD inline TYPE* INDEX(TYPE* base, int row, int col, size_t pitch) { return ROW(base, row, pitch) + col; }

// three-way comparison with absolute tolerance, reference include/macro.h:28-42
B2S_HD inline int compare(double x, double y = 0.0, double epsilon = 1e-9)
{
    const double d = x - y;
    if ((d < 0 ? -d : d) < epsilon) return 0;
    return x < y ? -1 : 1;
}

// reference include/macro.h:44-53: fopen or terminate
inline FILE* openFile(const char* path, const char* mode)
{
    FILE* f = fopen(path, mode);
    if (!f) {
        fprintf(stderr, "Cannot open file!\n");
        exit(-1);
    }
    return f;
}

// ---- error.cuh --------------------------------------------------------------------------------
void HandleError(cudaError_t err, const char* file, int line);   // reference src/error.cu:5-12
void checkKernelError(const char* file, int line);               // reference src/error.cu:14-18
#define HANDLE_ERROR(err) (HandleError(err, __FILE__, __LINE__))
#define HANDLE_KERNEL_ERROR() (checkKernelError(__FILE__, __LINE__))

// ---- problem.h ---------------------------------------------------------------------------------
// max c.x  s.t.  A x <= b, x >= 0.  reference include/problem.h:10-26 (field order kept: the struct
// is shared by value with user code).
typedef struct {
    TYPE* constraintsMatrix;   // A, variable-major: A[j * constraints + i]
    TYPE* knownTermsVector;    // b[constraints]
    TYPE* objectiveFunction;   // c[vars]
    int vars;
    int constraints;
} problem_t;

problem_t* readProblemFromFile(FILE* file);         // reference include/problem.h:37, src/problem.cu:20-47
problem_t* readRandomProblemFromFile(FILE* file);   // :45, src/problem.cu:128-139
problem_t* generateRandomProblem(int nVars, int nConstraints, unsigned int seed, int minGenerator = -100,
                                 int maxGenerator = 100);            // :54, src/problem.cu:49-126
void printProblemToStream(FILE* Stream, problem_t* problem);         // :67, src/problem.cu:141-181
void freeProblem(problem_t* problem);                                // :73, src/problem.cu:183-188

// ---- tabular.cuh --------------------------------------------------------------------------------
// reference include/tabular.cuh:5-30
typedef struct {
    problem_t* problem;
    TYPE* table;              // device, rows x cols, row pitch `pitch` bytes; device row r = tableau column r
    TYPE* knownTermsVector;   // = table (device row 0, the RHS)
    TYPE* constraintsMatrix;  // = device row 1
    TYPE* costsVector;        // device, rows entries, [0] = objective value
    size_t pitch;
    int rows;
    int cols;
} tabular_t;

tabular_t* newTabular(problem_t* problem);                                    // reference include/tabular.cuh:37
void printTableauToStream(FILE* Stream, tabular_t* tabular, int* base);       // :45
void freeTabular(tabular_t* tabular);                                         // :51

// ---- solver.h / twoPhaseMethod.h -------------------------------------------------------------------
int solve(tabular_t* tabular, int* base);                                     // reference include/solver.h:26

#define INFEASIBLE -1   // reference include/twoPhaseMethod.h:5-8
#define UNBOUNDED -2
#define DEGENERATE -3
#define FEASIBLE 0

int twoPhaseMethod(problem_t* problem, TYPE* solution, TYPE* optimalValue);   // reference include/twoPhaseMethod.h:19

#ifdef TIMER
void enableBenchmarkMode();    // reference include/twoPhaseMethod.h:21-24
void disableBenchmarkMode();
// TIMER is a compile-time switch of the reference (per-step CSV, src/chrono.cu); the shim library is built
// once, so a program compiled with -D TIMER announces it at start-up and gets the same CSV files.
void b2s_compat_timer_build();
namespace {
struct B2sTimerAnnounce {
    B2sTimerAnnounce() { b2s_compat_timer_build(); }
} b2s_timer_announce_;
}  // namespace
#endif

// ---- reduction.cuh / gaussian.cuh ---------------------------------------------------------------------
TYPE minElement(TYPE* g_vet, unsigned int size, unsigned int* outIndex);                        // reference include/reduction.cuh:12
TYPE minElement(TYPE* knownTerms, TYPE* rowPivot, unsigned int size, unsigned int* outIndex);    // :14
bool isLessOrEqualThanZero(TYPE* g_vet, unsigned int size);                                      // :23
void updateObjectiveFunction(tabular_t* tabular, int* base);                                     // reference include/gaussian.cuh:5

// ---- generator.cuh ------------------------------------------------------------------------------------
// reference include/generator.cuh:17,30: generate on `dst` asynchronously, return the (malloc'd) stream.
cudaStream_t* generateVectorInParallelAsync(TYPE* dst, int size, unsigned int seed, double minimum, double maximum);
cudaStream_t* generateMatrixInParallelAsync(TYPE* dst, int width, int height, unsigned int seed, double minimum,
                                            double maximum);

// ---- chrono.cuh -----------------------------------------------------------------------------------------
void initCsv();                                        // reference include/chrono.cuh:6-13
void initCsvBenchmark(int vars, int constraints);
void start(tabular_t* tabular, const char* operation);
void stop();
void closeCsv();
