// tests/cpp/tabular_level.cu -- exercises the tabular_t level of the drop-in API exactly the way the
// reference's own twoPhaseMethod.cu uses it: newTabular -> (caller fills the tableau) ->
// updateObjectiveFunction(tabular, base) -> solve(tabular, base), plus the reduction.cuh primitives on device
// vectors.  Prints one JSON object; tests/test_tabular_level.py compares it with the serial oracle.
//   usage: test_tabular_level <lp.txt>      (text LP format of src/problem.cu:20-47)
#include <vector>

#include "problem.h"
#include "reduction.cuh"
#include "gaussian.cuh"
#include "solver.h"
#include "tabular.cuh"
#include "twoPhaseMethod.h"

int main(int argc, char** argv)
{
    if (argc < 2) return 2;
    FILE* f = openFile(argv[1], "r");
    problem_t* p = readProblemFromFile(f);
    fclose(f);
    const int n = p->vars, m = p->constraints;
    tabular_t* t = newTabular(p);
    const int R = t->rows;  // 1 + n + 2m

    // phase-1 tableau in the reference layout (src/twoPhaseMethod.cu:145-200), built on the host
    std::vector<double> T((size_t)R * m, 0.0), cost(R, 0.0);
    std::vector<int> base(m);
    for (int i = 0; i < m; ++i) T[i] = p->knownTermsVector[i];
    for (int j = 0; j < n; ++j)
        for (int i = 0; i < m; ++i) T[(size_t)(1 + j) * m + i] = p->constraintsMatrix[(size_t)j * m + i];
    for (int i = 0; i < m; ++i) {
        T[(size_t)(1 + n + i) * m + i] = 1.0;
        T[(size_t)(1 + n + m + i) * m + i] = 1.0;
        base[i] = n + m + i;
    }
    for (int i = 0; i < m; ++i)
        if (compare(T[i]) < 0)
            for (int r = 0; r < R; ++r) T[(size_t)r * m + i] = -T[(size_t)r * m + i];
    for (int r = n + m + 1; r < R; ++r) cost[r] = 1.0;
    HANDLE_ERROR(cudaMemcpy2D(t->table, t->pitch, T.data(), sizeof(double) * m, sizeof(double) * m, R, cudaMemcpyHostToDevice));
    HANDLE_ERROR(cudaMemcpy(t->costsVector, cost.data(), sizeof(double) * R, cudaMemcpyHostToDevice));

    updateObjectiveFunction(t, base.data());
    HANDLE_ERROR(cudaMemcpy(cost.data(), t->costsVector, sizeof(double) * R, cudaMemcpyDeviceToHost));
    printf("{\"priced_cost0\": %.17g, ", cost[0]);

    // the reduction.cuh primitives on the priced cost vector and on the RHS / first structural column
    unsigned idx = 0;
    const double mn = minElement(t->costsVector + 1, (unsigned)(R - 1), &idx);
    printf("\"min_cost\": %.17g, \"min_cost_index\": %u, ", mn, idx);
    unsigned ridx = 0;
    const double rmin = minElement(t->knownTermsVector, t->constraintsMatrix, (unsigned)m, &ridx);
    printf("\"ratio_min\": %.17g, \"ratio_index\": %u, ", rmin, ridx);
    printf("\"col1_nonpositive\": %d, ", isLessOrEqualThanZero(t->constraintsMatrix, (unsigned)m) ? 1 : 0);

    const int st = solve(t, base.data());
    HANDLE_ERROR(cudaMemcpy(cost.data(), t->costsVector, sizeof(double) * R, cudaMemcpyDeviceToHost));
    printf("\"status\": %d, \"cost0\": %.17g, \"base\": [", st, cost[0]);
    for (int i = 0; i < m; ++i) printf("%s%d", i ? ", " : "", base[i]);
    printf("]}\n");
    freeTabular(t);
    freeProblem(p);
    return 0;
}
