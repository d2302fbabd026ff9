// b2s_generator.cuh -- on-device synthetic LP generator.
//
// Produces, for the same three kernel seeds, exactly the instance the reference's generator
// builds (src/generator.cu:9-32 with cuRAND's XORWOW; src/problem.cu:82-110 for which seed feeds
// which array): element k of a stream is `curand_uniform` of XORWOW output #k after
// curand_init(seed, 0, 0), mapped by u*(max-min)+min.  For the matrix, element (variable j,
// constraint i) is stream position i*n+j.  Unlike the reference nothing goes through the host:
// values are written straight into the tableau rows, stream positions are 64-bit (the
// reference's `idX * rows` is an int and overflows beyond 2^31 elements), and the position of a
// thread's first element is reached with a GF(2) jump instead of cuRAND's skipahead.
//
// XORWOW (Marsaglia 2003, as parameterised by cuRAND): five 32-bit xorshift words v[0..4] plus a
// Weyl counter d += 362437; output = v[4] + d.  The xorshift part is linear over GF(2), so a jump
// of k steps is the 160x160 bit matrix L^k applied to v; the table holds L^(2^b), b = 0..kJumpBits-1,
// each as 160 column images of 5 words.
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>

namespace b2s {

constexpr int kJumpBits = 48;
constexpr int kStateBits = 160;
constexpr int kJumpWords = kStateBits * 5;  // words per matrix

struct Xorwow {
    uint32_t v[5];
    uint32_t d;
};

__host__ __device__ inline void xorwow_seed(Xorwow& s, unsigned long long seed)
{
    // state scrambling of curand_init for XORWOW (curand_kernel.h, _curand_init_scratch)
    const uint32_t s0 = ((uint32_t)seed) ^ 0xaad26b49u;
    const uint32_t s1 = (uint32_t)(seed >> 32) ^ 0xf7dcefddu;
    const uint32_t t0 = 1099087573u * s0;
    const uint32_t t1 = 2591861531u * s1;
    s.d = 6615241u + t1 + t0;
    s.v[0] = 123456789u + t0;
    s.v[1] = 362436069u ^ t0;
    s.v[2] = 521288629u + t1;
    s.v[3] = 88675123u ^ t1;
    s.v[4] = 5783321u + t0;
}

__host__ __device__ inline void xorwow_step_linear(uint32_t v[5])
{
    const uint32_t t = v[0] ^ (v[0] >> 2);
    v[0] = v[1];
    v[1] = v[2];
    v[2] = v[3];
    v[3] = v[4];
    v[4] = (v[4] ^ (v[4] << 4)) ^ (t ^ (t << 1));
}

__host__ __device__ inline uint32_t xorwow_next(Xorwow& s)
{
    xorwow_step_linear(s.v);
    s.d += 362437u;
    return s.v[4] + s.d;
}

// Host: build L^(2^b) tables.  Matrix layout: col[i*5 + w] = word w of the image of basis bit i.
inline void xorwow_build_jump_tables(std::vector<uint32_t>& tables)
{
    tables.assign((size_t)kJumpBits * kJumpWords, 0u);
    uint32_t* M0 = tables.data();
    for (int i = 0; i < kStateBits; ++i) {
        uint32_t v[5] = {0, 0, 0, 0, 0};
        v[i / 32] = 1u << (i % 32);
        xorwow_step_linear(v);
        for (int w = 0; w < 5; ++w) M0[i * 5 + w] = v[w];
    }
    for (int b = 1; b < kJumpBits; ++b) {
        const uint32_t* A = tables.data() + (size_t)(b - 1) * kJumpWords;
        uint32_t* C = tables.data() + (size_t)b * kJumpWords;
        // C = A * A : image of basis bit i under C = A applied to (A's image of bit i)
        for (int i = 0; i < kStateBits; ++i) {
            uint32_t acc[5] = {0, 0, 0, 0, 0};
            for (int j = 0; j < kStateBits; ++j)
                if ((A[i * 5 + j / 32] >> (j % 32)) & 1u)
                    for (int w = 0; w < 5; ++w) acc[w] ^= A[j * 5 + w];
            for (int w = 0; w < 5; ++w) C[i * 5 + w] = acc[w];
        }
    }
}

__host__ __device__ inline void xorwow_jump(Xorwow& s, unsigned long long k, const uint32_t* __restrict__ tables)
{
    s.d += 362437u * (uint32_t)k;
    for (int b = 0; k != 0 && b < kJumpBits; ++b, k >>= 1) {
        if (!(k & 1ull)) continue;
        const uint32_t* M = tables + (size_t)b * kJumpWords;
        uint32_t acc[5] = {0, 0, 0, 0, 0};
        for (int w = 0; w < 5; ++w) {
            uint32_t bits = s.v[w];
            while (bits) {
#ifdef __CUDA_ARCH__
                const int bit = __ffs(bits) - 1;
#else
                const int bit = __builtin_ctz(bits);
#endif
                bits &= bits - 1;
                const uint32_t* colp = M + (w * 32 + bit) * 5;
                for (int x = 0; x < 5; ++x) acc[x] ^= colp[x];
            }
        }
        for (int w = 0; w < 5; ++w) s.v[w] = acc[w];
    }
}

// curand_uniform (curand_uniform.h:69-72, one FFMA in the reference's SASS) followed by
// u*(max-min)+min contracted to one DFMA (src/generator.cu:18,30).
__device__ __forceinline__ double xorwow_value(uint32_t x, double lo, double span)
{
    const float u = __fmaf_rn((float)x, 2.3283064e-10f, 2.3283064e-10f / 2.0f);
    return __fma_rn((double)u, span, lo);
}

// Vector: local element k holds stream position id0 + k (id0 = first constraint of a sharded slab).
// One thread per kVecRun consecutive elements.
constexpr int kVecRun = 64;
template <typename real>
__global__ void __launch_bounds__(256) generate_vector_kernel(real* out, long long count, long long id0, unsigned seed,
                                                              double lo, double span, const uint32_t* __restrict__ tables)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long first = t * kVecRun;
    if (first >= count) return;
    Xorwow s;
    xorwow_seed(s, seed);
    xorwow_jump(s, (unsigned long long)(id0 + first), tables);
    const long long last = min(count, first + kVecRun);
    for (long long k = first; k < last; ++k) out[k] = (real)xorwow_value(xorwow_next(s), lo, span);
}

// Matrix: thread (constraint i, run r) produces variables j in [r*kMatRun, (r+1)*kMatRun) of
// constraint i = stream positions i*n + j, and writes T[(row0 + j)*ld + (i - col0)].  Consecutive
// threads own consecutive constraints, so every store instruction is coalesced along a row.
constexpr int kMatRun = 128;
template <typename real>
__global__ void __launch_bounds__(256) generate_matrix_kernel(real* T, long long ld, long long row0, int n, int m_loc,
                                                              int col0, unsigned seed, double lo, double span,
                                                              const uint32_t* __restrict__ tables)
{
    const int li = blockIdx.x * blockDim.x + threadIdx.x;
    if (li >= m_loc) return;
    const long long j0 = (long long)blockIdx.y * kMatRun;
    if (j0 >= n) return;
    const long long i = (long long)col0 + li;
    Xorwow s;
    xorwow_seed(s, seed);
    xorwow_jump(s, (unsigned long long)(i * (long long)n + j0), tables);
    const long long j1 = min((long long)n, j0 + kMatRun);
    real* dst = T + (row0 + j0) * ld + li;
    for (long long j = j0; j < j1; ++j, dst += ld) *dst = (real)xorwow_value(xorwow_next(s), lo, span);
}

}  // namespace b2s
