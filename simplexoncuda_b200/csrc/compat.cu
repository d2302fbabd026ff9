// compat.cu -- definitions behind include/compat/simplex_compat.hpp: the reference's C++ entry points
// (same names, arguments, statuses, ownership) as thin shims over the C ABI of include/b2s.h.
// Built into simplexoncuda_b200/lib/libb2s_compat.so; links against libb2s.so.
#include "../../include/compat/simplex_compat.hpp"
#include "../../include/b2s.h"

#include <time.h>
#include <vector>

#include "b2s_generator.cuh"

namespace {

// One process-wide solver handle, created on first use (the reference is single-device,
// single-threaded and not re-entrant either: file-scope state in src/chrono.cu:4-6,
// src/twoPhaseMethod.cu:13).
b2s_solver* handle()
{
    static b2s_solver* h = nullptr;
    if (!h) {
        b2s_options opt;
        b2s_default_options(&opt);
        if (const char* e = getenv("B2S_PIVOT_RULE")) opt.pivot_rule = atoi(e);
        if (b2s_create(&opt, &h) != B2S_OK) {
            printf("%s\n", b2s_last_error(nullptr));  // the reference prints and exits on any CUDA error
            exit(EXIT_FAILURE);
        }
    }
    return h;
}

void must(int rc)
{
    if (rc != B2S_OK) {
        printf("%s\n", b2s_last_error(handle()));
        exit(EXIT_FAILURE);
    }
}

problem_t* alloc_problem(int nVars, int nConstraints)  // src/problem.cu:7-18
{
    problem_t* p = (problem_t*)malloc(sizeof(problem_t));
    p->vars = nVars;
    p->constraints = nConstraints;
    p->objectiveFunction = (TYPE*)malloc(sizeof(TYPE) * (size_t)nVars);
    p->constraintsMatrix = (TYPE*)malloc(sizeof(TYPE) * (size_t)nVars * (size_t)nConstraints);
    p->knownTermsVector = (TYPE*)malloc(sizeof(TYPE) * (size_t)nConstraints);
    return p;
}

bool g_benchmark = false;

}  // namespace

static FILE* g_csv = nullptr;  // TIMER CSV (chrono.cuh section below)

// ---- error.cuh ------------------------------------------------------------------------------------
void HandleError(cudaError_t err, const char* file, int line)
{
    if (err != cudaSuccess) {
        printf("%s in %s at line %d\n", cudaGetErrorString(err), file, line);
        exit(EXIT_FAILURE);
    }
}
void checkKernelError(const char* file, int line)
{
    cudaDeviceSynchronize();
    HandleError(cudaGetLastError(), file, line);
}

// ---- problem.h -------------------------------------------------------------------------------------
problem_t* readProblemFromFile(FILE* file)
{
    int n = 0, m = 0;
    if (fscanf(file, "%d %d", &n, &m) != 2) return alloc_problem(0, 0);
    problem_t* p = alloc_problem(n, m);
    for (int j = 0; j < n; ++j)
        if (fscanf(file, "%lf", &p->objectiveFunction[j]) != 1) break;
    for (int i = 0; i < m; ++i) {  // one text line per constraint: a_i1 .. a_in b_i ; stored variable-major
        for (int j = 0; j < n; ++j)
            if (fscanf(file, "%lf", &p->constraintsMatrix[(size_t)j * m + i]) != 1) break;
        if (fscanf(file, "%lf", &p->knownTermsVector[i]) != 1) break;
    }
    return p;
}

problem_t* generateRandomProblem(int nVars, int nConstraints, unsigned int seed, int minGenerator, int maxGenerator)
{
    // seeds exactly as the reference derives them on the platform it is built on: srand(seed); rand() x3
    // (B2S_RAND_FLAVOUR=1 selects the MSVC sequence of the published measurements instead)
    unsigned seeds[3];
    int flavour = B2S_RAND_GLIBC;
    if (const char* e = getenv("B2S_RAND_FLAVOUR")) flavour = atoi(e);
    b2s_seed_triplet(seed, flavour, seeds);
    problem_t* p = alloc_problem(nVars, nConstraints);
    must(b2s_generate_problem_device(handle(), nVars, nConstraints, seeds, (double)minGenerator, (double)maxGenerator));
    must(b2s_copy_problem(handle(), p->constraintsMatrix, p->knownTermsVector, p->objectiveFunction));
    return p;
}

problem_t* readRandomProblemFromFile(FILE* file)
{
    int n = 0, m = 0, lo = 0, hi = 0;
    unsigned seed = 0;
    if (fscanf(file, "%d %d %u %d %d", &n, &m, &seed, &lo, &hi) != 5) return alloc_problem(0, 0);
    return generateRandomProblem(n, m, seed, lo, hi);
}

void printProblemToStream(FILE* Stream, problem_t* problem)
{
    const int n = problem->vars, m = problem->constraints;
    fprintf(Stream, "max ");
    for (int j = 0; j < n; ++j) {
        const double v = problem->objectiveFunction[j];
        fprintf(Stream, "%s %.2lf X%d ", v >= 0 ? "+" : "-", v < 0 ? -v : v, j + 1);
    }
    fprintf(Stream, "\nsubject to \n");
    for (int i = 0; i < m; ++i) {
        for (int j = 0; j < n; ++j) {
            const double v = problem->constraintsMatrix[(size_t)j * m + i];
            fprintf(Stream, "%s %.2lf X%d ", v >= 0 ? "+" : "-", v < 0 ? -v : v, j + 1);
        }
        fprintf(Stream, "<= %.2lf\n", problem->knownTermsVector[i]);
    }
}

void freeProblem(problem_t* problem)
{
    free(problem->constraintsMatrix);
    free(problem->knownTermsVector);
    free(problem->objectiveFunction);
}

// ---- twoPhaseMethod.h --------------------------------------------------------------------------------
static bool g_timer_build = false;  // a translation unit of the program was compiled with -D TIMER
void b2s_compat_timer_build() { g_timer_build = true; }
static bool timer_on() { return g_timer_build || g_benchmark || getenv("B2S_TIMER") != nullptr; }

// Pivot loop of one phase.  With the TIMER CSV enabled every iteration is timed on its own with CUDA events
// (b2s_profile_pivots) and logged as one `solve` line, like the reference's -D TIMER build does around each
// call of its per-iteration function (src/solver.cu:84-124); the last line is the terminating optimality check.
static int run_phase(b2s_solver* h, tabular_t* shape)
{
    int st = B2S_RUNNING;
    long long done = 0;
    if (!timer_on()) {
        must(b2s_iterate(h, -1, &st, &done));
        return st;
    }
    const int chunk = 256;
    std::vector<float> a(chunk), b(chunk), c(chunk);
    while (true) {
        must(b2s_profile_pivots(h, chunk, a.data(), b.data(), c.data(), &done));
        for (long long k = 0; k < done; ++k)
            fprintf(g_csv, "%d,%d,solve,%f\n", shape->rows, shape->cols, (double)(a[k] + b[k] + c[k]) * 1000.0);
        if (done < chunk) break;
    }
    start(shape, "solve");  // the iteration that finds no entering column (or an unbounded one)
    must(b2s_iterate(h, 0, &st, &done));
    stop();
    must(b2s_iterate(h, -1, &st, &done));  // state is final already: returns the phase status
    return st;
}

int twoPhaseMethod(problem_t* problem, TYPE* solution, TYPE* optimalValue)
{
    b2s_solver* h = handle();
    const bool timed = timer_on();
    tabular_t shape;  // only rows/cols are used, by the CSV lines
    memset(&shape, 0, sizeof(shape));
    shape.cols = problem->constraints;
    shape.rows = 1 + problem->vars + 2 * problem->constraints;
    if (timed) {
        if (g_benchmark)
            initCsvBenchmark(problem->vars, problem->constraints);
        else
            initCsv();
    }
    // the reference's progress lines (src/twoPhaseMethod.cu:228,243,257,296,333,347)
    printf("Phase 1: Filling Tableau\n");
    if (timed) start(&shape, "fillTableau");
    must(b2s_load_problem_host(h, problem->vars, problem->constraints, problem->constraintsMatrix,
                               problem->knownTermsVector, problem->objectiveFunction));
    must(b2s_build_phase1(h));
    if (timed) stop();
    printf("Phase 1: Resetting out-of-base variables\n");
    if (timed) start(&shape, "gauss1");
    must(b2s_price_out(h));
    if (timed) stop();
    printf("Phase 1: Solving auxiliary problem\n");
    must(b2s_select_entering(h));
    run_phase(h, &shape);  // the phase-1 status is ignored by the reference too (src/twoPhaseMethod.cu:258)
    int verdict = FEASIBLE;
    if (timed) start(&shape, "checkDegeneracy");
    must(b2s_phase1_verdict(h, &verdict));
    if (timed) stop();
    int result = verdict;
    if (verdict == FEASIBLE) {
        shape.rows -= shape.cols;  // src/twoPhaseMethod.cu:288
        printf("Phase 2: Filling costs vector with the original one\n");
        if (timed) start(&shape, "costsVector");
        must(b2s_switch_phase2(h));
        if (timed) stop();
        printf("Phase 2: Resetting out-of-base variables\n");
        if (timed) start(&shape, "gauss2");
        must(b2s_price_out(h));
        if (timed) stop();
        printf("Phase 2: Solving original problem\n");
        must(b2s_select_entering(h));
        result = run_phase(h, &shape);
        if (result == FEASIBLE) {
            if (timed) start(&shape, "solution");
            must(b2s_extract_solution(h, solution, optimalValue));
            if (timed) stop();
        }
    }
    if (timed) closeCsv();
    return result;
}

void enableBenchmarkMode() { g_benchmark = true; }
void disableBenchmarkMode() { g_benchmark = false; }

// ---- tabular.cuh ---------------------------------------------------------------------------------------
tabular_t* newTabular(problem_t* problem)  // src/tabular.cu:25-39 (zero-filled here; the reference does not)
{
    tabular_t* t = (tabular_t*)malloc(sizeof(tabular_t));
    t->problem = problem;
    t->cols = problem->constraints;
    t->rows = 1 + problem->vars + 2 * problem->constraints;
    HANDLE_ERROR(cudaMallocPitch((void**)&t->table, &t->pitch, sizeof(TYPE) * (size_t)t->cols, (size_t)t->rows));
    HANDLE_ERROR(cudaMemset2D(t->table, t->pitch, 0, t->pitch, (size_t)t->rows));
    HANDLE_ERROR(cudaDeviceSynchronize());
    HANDLE_ERROR(cudaMalloc((void**)&t->costsVector, sizeof(TYPE) * (size_t)t->rows));
    t->knownTermsVector = t->table;
    t->constraintsMatrix = ROW(t->table, 1, t->pitch);
    return t;
}

void freeTabular(tabular_t* tabular)
{
    if (tabular->table) {
        HANDLE_ERROR(cudaFree(tabular->table));
        HANDLE_ERROR(cudaFree(tabular->costsVector));
    }
    free(tabular);
}

void printTableauToStream(FILE* Stream, tabular_t* tabular, int* base)  // src/tabular.cu:41-98
{
    if (!tabular->table) return;
    std::vector<TYPE> tab((size_t)tabular->rows * tabular->cols), costs((size_t)tabular->rows);
    HANDLE_ERROR(cudaMemcpy2D(tab.data(), sizeof(TYPE) * tabular->cols, tabular->table, tabular->pitch,
                              sizeof(TYPE) * tabular->cols, (size_t)tabular->rows, cudaMemcpyDeviceToHost));
    HANDLE_ERROR(cudaMemcpy(costs.data(), tabular->costsVector, sizeof(TYPE) * tabular->rows, cudaMemcpyDeviceToHost));
    fprintf(Stream, "\n--------------- Tabular --------------\n");
    for (int r = 0; r < tabular->rows; ++r) {
        for (int c = 0; c < tabular->cols; ++c) fprintf(Stream, "%.2lf\t", tab[(size_t)r * tabular->cols + c]);
        fprintf(Stream, "\t|\t %.11lf\n", costs[r]);
        if (r == 0) fprintf(Stream, "\n");
    }
    fprintf(Stream, "Base\n");
    for (int c = 0; c < tabular->cols; ++c) fprintf(Stream, "%d\t", base[c]);
}

// ---- solver.h / gaussian.cuh / reduction.cuh on a caller-owned tabular_t --------------------------------
static void attach(tabular_t* t, int* base)
{
    // the caller filled the tableau on its own streams; the solver works on a private non-blocking stream
    HANDLE_ERROR(cudaDeviceSynchronize());
    must(b2s_attach_tableau_device(handle(), t->table, t->pitch, t->rows, t->cols, t->costsVector, t->problem->vars));
    must(b2s_set_basis(handle(), base));
}

int solve(tabular_t* tabular, int* base)  // src/solver.cu:128-149: pivot until optimal (FEASIBLE) or UNBOUNDED
{
    attach(tabular, base);
    must(b2s_select_entering(handle()));
    int st = B2S_RUNNING;
    long long done = 0;
    must(b2s_iterate(handle(), -1, &st, &done));
    must(b2s_copy_basis(handle(), base));  // base[p] = q updates (src/solver.cu:105)
    return st;
}

void updateObjectiveFunction(tabular_t* tabular, int* base)  // src/gaussian.cu:132-162
{
    attach(tabular, base);
    must(b2s_price_out(handle()));
    cudaDeviceSynchronize();
}

TYPE minElement(TYPE* g_vet, unsigned int size, unsigned int* outIndex)
{
    double v = 0;
    HANDLE_ERROR(cudaDeviceSynchronize());
    must(b2s_min_element_device(handle(), g_vet, size, &v, outIndex));
    return v;
}

TYPE minElement(TYPE* knownTerms, TYPE* rowPivot, unsigned int size, unsigned int* outIndex)
{
    double v = 0;
    HANDLE_ERROR(cudaDeviceSynchronize());
    must(b2s_ratio_min_device(handle(), knownTerms, rowPivot, size, &v, outIndex));
    return v;
}

bool isLessOrEqualThanZero(TYPE* g_vet, unsigned int size)
{
    int r = 0;
    HANDLE_ERROR(cudaDeviceSynchronize());
    must(b2s_max_le_zero_device(handle(), g_vet, size, &r));
    return r != 0;
}

// ---- generator.cuh ---------------------------------------------------------------------------------------
static const uint32_t* jump_tables()
{
    static uint32_t* dev = nullptr;
    if (!dev) {
        std::vector<uint32_t> host;
        b2s::xorwow_build_jump_tables(host);
        HANDLE_ERROR(cudaMalloc(&dev, host.size() * sizeof(uint32_t)));
        HANDLE_ERROR(cudaMemcpy(dev, host.data(), host.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
        HANDLE_ERROR(cudaDeviceSynchronize());  // consumers run on other (possibly non-blocking) streams
    }
    return dev;
}

// src/generator.cu:36-44: fill a device-accessible vector on a fresh stream, return the stream.
cudaStream_t* generateVectorInParallelAsync(TYPE* dst, int size, unsigned int seed, double minimum, double maximum)
{
    const uint32_t* tables = jump_tables();
    cudaStream_t* stream = (cudaStream_t*)malloc(sizeof(cudaStream_t));
    HANDLE_ERROR(cudaStreamCreate(stream));
    const long long threads = ((long long)size + b2s::kVecRun - 1) / b2s::kVecRun;
    b2s::generate_vector_kernel<double><<<(unsigned)((threads + 255) / 256), 256, 0, *stream>>>(
        dst, (long long)size, 0ll, seed, minimum, maximum - minimum, tables);
    return stream;
}

// src/generator.cu:46-77: generate the variable-major matrix (width = constraints, height = variables) on
// the device and copy it to the HOST buffer dst on the returned stream.
cudaStream_t* generateMatrixInParallelAsync(TYPE* dst, int width, int height, unsigned int seed, double minimum,
                                            double maximum)
{
    const uint32_t* tables = jump_tables();
    cudaStream_t* stream = (cudaStream_t*)malloc(sizeof(cudaStream_t));
    HANDLE_ERROR(cudaStreamCreate(stream));
    TYPE* dev = nullptr;
    size_t pitch = 0;
    HANDLE_ERROR(cudaMallocPitch((void**)&dev, &pitch, sizeof(TYPE) * (size_t)width, (size_t)height));
    dim3 grid((unsigned)((width + 255) / 256), (unsigned)((height + b2s::kMatRun - 1) / b2s::kMatRun));
    b2s::generate_matrix_kernel<double><<<grid, 256, 0, *stream>>>(dev, (long long)(pitch / sizeof(TYPE)), 0ll, height, width, 0,
                                                                 seed, minimum, maximum - minimum, tables);
    HANDLE_ERROR(cudaMemcpy2DAsync(dst, sizeof(TYPE) * (size_t)width, dev, pitch, sizeof(TYPE) * (size_t)width, (size_t)height,
                                   cudaMemcpyDeviceToHost, *stream));
    HANDLE_ERROR(cudaFreeAsync(dev, *stream));
    return stream;
}

// ---- chrono.cuh: the reference's TIMER CSV (src/chrono.cu:8-56), same file naming and line format ----------
static cudaEvent_t g_ev0, g_ev1;

static void open_csv(const char* name)
{
    g_csv = openFile(name, "w");
    fprintf(g_csv, "vars,contraints,operation,elapsed_time\n");
    HANDLE_ERROR(cudaEventCreate(&g_ev0));
    HANDLE_ERROR(cudaEventCreate(&g_ev1));
}
void initCsv()
{
    time_t now = time(NULL);
    char stamp[20], name[64];
    strftime(stamp, sizeof(stamp), "%Y%m%d%H%M%S.%d", localtime(&now));
    snprintf(name, sizeof(name), "..\\data\\measures\\times_%s.txt", stamp);
    open_csv(name);
}
void initCsvBenchmark(int vars, int constraints)
{
    char name[64];
    snprintf(name, sizeof(name), "..\\data\\measures\\benchmark_%d_%d.txt", vars, constraints);
    open_csv(name);
}
void start(tabular_t* tabular, const char* operation)
{
    fprintf(g_csv, "%d,%d,%s,", tabular->rows, tabular->cols, operation);
    cudaDeviceSynchronize();
    cudaEventRecord(g_ev0, 0);
}
void stop()
{
    cudaDeviceSynchronize();
    cudaEventRecord(g_ev1, 0);
    cudaEventSynchronize(g_ev1);
    float ms = 0;
    cudaEventElapsedTime(&ms, g_ev0, g_ev1);
    fprintf(g_csv, "%f\n", ms * 1000);
}
void closeCsv()
{
    if (g_csv) fclose(g_csv);
    g_csv = nullptr;
    cudaEventDestroy(g_ev0);
    cudaEventDestroy(g_ev1);
}
