#!/bin/bash
# Build the C-ABI shared library in-tree:  simplexoncuda_b200/lib/libb2s.so  (sm_100a only).
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/../lib"
mkdir -p "$OUT"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
NCCL_INC="${NCCL_INC:-$(python - <<'PY'
import os, sys
try:
    import nvidia.nccl as n
    print(os.path.join(list(n.__path__)[0], "include"))
except Exception:
    print("/usr/include")
PY
)}"
NCCL_LIB="${NCCL_LIB:-$(python - <<'PY'
import os
try:
    import nvidia.nccl as n
    print(os.path.join(list(n.__path__)[0], "lib"))
except Exception:
    print("/usr/lib/x86_64-linux-gnu")
PY
)}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false
       -Xcompiler -fPIC,-O2,-ffp-contract=off,-Wall -shared -DB2S_WITH_NCCL -I"$NCCL_INC")
"$NVCC" "${FLAGS[@]}" ${B2S_PTXAS_V:+-Xptxas -v} -o "$OUT/libb2s.so" "$HERE/b2s_solver.cu" \
    -L"$NCCL_LIB" -l:libnccl.so.2 -Xlinker -rpath,"$NCCL_LIB"
"$NVCC" -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -Xcompiler -fPIC,-O2,-ffp-contract=off \
    -shared -o "$OUT/libb2s_compat.so" "$HERE/compat.cu" -L"$OUT" -lb2s -Xlinker -rpath,'$ORIGIN'
# C++ client of the tabular_t-level drop-in API (used by tests/test_tabular_level.py on the GPU box)
"$NVCC" -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I"$HERE/../../include/compat" \
    -o "$OUT/test_tabular_level" "$HERE/../../tests/cpp/tabular_level.cu" -L"$OUT" -lb2s_compat -lb2s -Xlinker -rpath,'$ORIGIN'
# plain C99 client of the C ABI (proves the header is C, used by tests/test_tabular_level.py)
GCC=/usr/bin/gcc; [ -x "$GCC" ] || GCC=gcc
"$GCC" -std=c99 -pedantic -Wall -Wextra -I"$HERE/../../include" "$HERE/../../examples/solve_file.c" \
    -L"$OUT" -lb2s -Wl,-rpath,'$ORIGIN' -o "$OUT/example_solve_file"
echo "built $OUT/libb2s.so $OUT/libb2s_compat.so $OUT/test_tabular_level $OUT/example_solve_file"
