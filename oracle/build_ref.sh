#!/bin/bash
# oracle/build_ref.sh -- compile the UNMODIFIED reference (rik1599/SimplexOnCuda) from the sources
# where they lie (read-only /root/reference) into oracle/_ref/ (git-ignored, travels to the GPU
# box with gpurun).  Outputs only; no reference source is copied.
#
#   oracle/_ref/SimplexOnCuda_ref     the stock program (main.cu + src/*.cu), flags of compile.sh
#                                     (-rdc=true -D TIMER) retargeted from sm_60 to sm_100
#   oracle/_ref/libsimplex_ref.so     the same sources (minus main.cu / chrono.cu, no TIMER) plus
#                                     oracle/ref_shim.cu: C entry points + pivot bookkeeping
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${REFERENCE_DIR:-/root/reference}"
OUT="$HERE/_ref"
if [ ! -d "$REF/src" ]; then
    echo "reference tree $REF not present: keeping prebuilt $OUT (if any)"; exit 0
fi
mkdir -p "$OUT"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
SRCS="src/error.cu src/problem.cu src/tabular.cu src/twoPhaseMethod.cu src/generator.cu src/reduction.cu src/gaussian.cu"
cd "$REF"
"$NVCC" -rdc=true -o "$OUT/SimplexOnCuda_ref" -Iinclude/ -arch=sm_100 -D TIMER -w \
    main.cu $SRCS src/solver.cu src/chrono.cu
"$NVCC" -rdc=true -shared -Xcompiler -fPIC -o "$OUT/libsimplex_ref.so" -Iinclude/ -I"$REF" -arch=sm_100 -w \
    $SRCS "$HERE/ref_shim.cu"
# Drop-in proof: the reference's own, unmodified main.cu compiled against OUR headers (include/compat)
# and linked against OUR libraries instead of the reference's sources.
LIBDIR="$HERE/../simplexoncuda_b200/lib"
if [ -f "$LIBDIR/libb2s_compat.so" ]; then
    "$NVCC" -arch=sm_100 -w -D TIMER -I"$HERE/../include/compat" -o "$OUT/SimplexOnCuda_dropin" main.cu \
        -L"$LIBDIR" -lb2s_compat -lb2s -Xlinker -rpath,"$LIBDIR"
fi
echo "built $OUT/SimplexOnCuda_ref, $OUT/libsimplex_ref.so, $OUT/SimplexOnCuda_dropin"
