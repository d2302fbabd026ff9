// b2s_bulk.cuh -- the rank-1 update with the tableau moved by the bulk-copy (TMA) engine instead of
// per-thread vector loads: update_variant 14.
//
// Same arithmetic and the same fused cost update / entering tournament as update_kernel; only the data
// movement differs.  Each CTA runs a ring of kBulkStages shared-memory stages of 2 tableau rows x 2048
// columns (32 KB).  One elected thread feeds the ring with 1-D bulk copies global -> shared
// (cp.async.bulk ... mbarrier::complete_tx::bytes, SASS UBLKCP) up to kBulkStages-2 tiles ahead; all 512
// threads wait on the stage's mbarrier, apply T = fma(s, a_p, T) in shared memory, and the elected thread
// writes the stage back with bulk copies shared -> global (bulk_group), recycling a stage once its
// write-back has finished READING shared memory (cp.async.bulk.wait_group.read).  Tiles come from the same
// device-wide ticket counter.  Measured against the register-streaming variant in
// profiles/r01_scaling_and_loop_modes.md; kept as an alternative, not the default.
#pragma once
#include "b2s_kernels.cuh"

namespace b2s {

constexpr int kBulkStages = 7;  // 7 x 32 KB = 224 KB of the 227 KB a CTA may use
constexpr int kBulkRows = 2;
constexpr int kBulkCols = 2048;  // elements of 8 bytes; fp32 uses the same byte width (4096 elements)

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// bounded wait on an mbarrier phase: false after ~2 s (a broken pipeline must not hang the GPU)
__device__ __forceinline__ bool mbar_wait(unsigned long long* bar, unsigned parity)
{
    const long long t0 = clock64();
    unsigned done = 0;
    while (true) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (done) return true;
        if (clock64() - t0 > 4000000000ll) return false;
    }
}
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gmem_src, unsigned bytes, unsigned long long* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_store(void* gmem_dst, const void* smem_src, unsigned bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read()
{
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_arrive(unsigned long long* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <typename real>
__global__ void __launch_bounds__(kSelBlock, 1) update_bulk_kernel(PivotParams<real> P)
{
    extern __shared__ __align__(128) unsigned char bulk_smem[];
    __shared__ TreeSmem<real> sm;
    __shared__ int s_flag;
    __shared__ __align__(8) unsigned long long full[kBulkStages];  // loads of the stage have landed
    __shared__ long long s_tile[kBulkStages];                      // tile held by each stage (-1: none)
    __shared__ int s_err;
    constexpr int EPR = kBulkCols * 8 / (int)sizeof(real);  // elements per row segment (16 KB)
    constexpr int EPT = 32 / (int)sizeof(real);             // elements per thread per row (32 bytes)
    constexpr unsigned kStageBytes = kBulkRows * kBulkCols * 8;
    constexpr int S = kBulkStages;
    DevState* st = P.st;
    if (!__ldcg(&st->live)) return;

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(&full[s], 1);
            s_tile[s] = -1;
        }
        s_err = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_async_smem();
    }
    __syncthreads();

    if (blockIdx.x < P.Gc) cost_select_blocks<real, true>(P, P.rowp, (real)__ldcg(&st->sc), sm, &s_flag);

    const long long nchunks = (P.ld + EPR - 1) / EPR;
    const long long nrb = (P.Rs + kBulkRows - 1) / kBulkRows;
    const long long ntiles = nrb * nchunks;
    auto stage_ptr = [&](int stage, int r) { return bulk_smem + (size_t)stage * kStageBytes + (size_t)r * kBulkCols * 8; };

    // elected thread: start the loads of `tile` into `stage`; an exhausted ticket still completes the
    // stage's mbarrier phase so that the consumers' acquire also publishes s_tile[stage] = -1
    auto issue = [&](int stage, long long tile) {
        if (tile >= ntiles) {
            s_tile[stage] = -1;
            mbar_arrive(&full[stage]);
            return;
        }
        const long long rb = tile / nchunks, chunk = tile % nchunks;
        const long long c0 = chunk * EPR;
        const unsigned row_bytes = (unsigned)(min((long long)EPR, P.ld - c0) * (long long)sizeof(real));
        const int rows = (int)min((long long)kBulkRows, P.Rs - rb * kBulkRows);
        s_tile[stage] = tile;
        mbar_expect_tx(&full[stage], row_bytes * (unsigned)rows);
        for (int r = 0; r < rows; ++r)
            bulk_load(stage_ptr(stage, r), P.T + (rb * kBulkRows + r) * P.ld + c0, row_bytes, &full[stage]);
    };

    // the first S-1 tiles of a CTA are static (blockIdx + k*grid); later ones come from the ticket counter,
    // of which the elected thread keeps one in flight so the atomic's L2 round trip overlaps a stage
    long long ticket = 0;
    if (threadIdx.x == 0) {
        for (int s = 0; s < S - 1; ++s) issue(s, (long long)blockIdx.x + (long long)s * gridDim.x);
        ticket = (long long)atomicAdd(&st->tile_ticket, 1u) + (long long)(S - 1) * gridDim.x;
    }

    const int tx = threadIdx.x;  // 512 threads x 32 B = one 16 KB row segment
    unsigned full_par = 0;       // parity of each stage's mbarrier
    long long cur_chunk = -1;
    real sreg[EPT];
    for (int it = 0;; ++it) {
        const int stage = it % S;
        if (!mbar_wait(&full[stage], (full_par >> stage) & 1u)) s_err = 1;
        full_par ^= 1u << stage;
        const long long tile = s_tile[stage];
        if (tile < 0) break;  // the ring ran dry (uniform: every thread reads the same slot after the same acquire)
        const long long rb = tile / nchunks, chunk = tile % nchunks;
        const long long c = chunk * EPR + (long long)tx * EPT;
        if (chunk != cur_chunk) {
            cur_chunk = chunk;
#pragma unroll
            for (int e = 0; e < EPT; ++e) sreg[e] = (c + e < P.ld) ? __ldg(P.s + c + e) : (real)0;
        }
        if (c < P.ld) {
#pragma unroll
            for (int r = 0; r < kBulkRows; ++r) {
                const long long row = rb * kBulkRows + r;
                if (row < P.Rs) {
                    const real a = __ldg(P.rowp + row);
                    real* v = reinterpret_cast<real*>(stage_ptr(stage, r)) + (size_t)tx * EPT;
                    PackView<real, 32> pk;
                    pk.p = *reinterpret_cast<const Pack<32>*>(v);
#pragma unroll
                    for (int e = 0; e < EPT; ++e) pk.e[e] = fma_r(sreg[e], a, pk.e[e]);
                    *reinterpret_cast<Pack<32>*>(v) = pk.p;
                }
            }
        }
        fence_async_smem();  // generic-proxy writes above -> visible to the bulk store below
        __syncthreads();     // the only block-wide barrier per stage
        if (*(volatile int*)&s_err) break;
        if (threadIdx.x == 0) {
            const long long c0 = chunk * EPR;
            const unsigned row_bytes = (unsigned)(min((long long)EPR, P.ld - c0) * (long long)sizeof(real));
            const int rows = (int)min((long long)kBulkRows, P.Rs - rb * kBulkRows);
            for (int r = 0; r < rows; ++r) bulk_store(P.T + (rb * kBulkRows + r) * P.ld + c0, stage_ptr(stage, r), row_bytes);
            bulk_commit();
            // the stage written back one iteration ago has been read out by now: refill it
            bulk_wait_read<1>();
            issue((it + S - 1) % S, ticket);
            ticket = (long long)atomicAdd(&st->tile_ticket, 1u) + (long long)(S - 1) * gridDim.x;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        bulk_wait_all();
        if (s_err) {
            st->status = -98;  // bulk pipeline timeout
            st->live = 0;
        }
        __threadfence();
        const unsigned fin = atomicAdd(&st->tile_done, 1u);
        if (fin == gridDim.x - 1) {
            st->tile_ticket = 0;
            st->tile_done = 0;
            __threadfence();
        }
    }
}

}  // namespace b2s
