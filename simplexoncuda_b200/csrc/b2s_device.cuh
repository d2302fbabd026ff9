// b2s_device.cuh -- device-side building blocks of the per-pivot hot path (sm_100a).
//
// What the reference does per simplex iteration (src/solver.cu:78-126) with 7-10 kernels, 5
// cudaMalloc/cudaFree pairs, 10 blocking copies and 6 device syncs is done here by three
// stream-ordered launches that never return to the host:
//
//   ratio_kernel   : entering column snapshot, unbounded test, ratio-test tournament  (A3-A6, A8)
//   gather_kernel  : pivot-constraint gather + in-place normalisation, s = -a_q/pivot  (A7, A9)
//   update_kernel  : rank-1 update T += s (x) a_p streamed through HBM exactly once, fused with
//                    the cost-vector update and the next pivot's entering-column tournament
//                                                                                (A9, A10, A1, A2)
//
// All arithmetic that decides the pivot path is spelled with explicit rounding intrinsics so
// that no compiler contraction can change a bit (build with -fmad=false as a second guard).
#pragma once
#include <cfloat>
#include <cstdint>
#include <cuda_runtime.h>

namespace b2s {

struct LaState;

constexpr int kRunning = -10;
constexpr int kFeasible = 0;
constexpr int kUnbounded = -2;

constexpr int kRuleReference = 0;
constexpr int kRuleLowest = 1;
constexpr int kRuleBland = 2;

constexpr int kSelBlock = 512;   // reference stage-1 block size (src/reduction.cu:6)
constexpr int kMaxSlots = 1024;  // reference stage-1 grid cap  (src/reduction.cu:7)
constexpr int kMaxPeers = 8;     // GPUs of one box
constexpr int kLaMaxHelpers = 16;  // helper CTAs of the look-ahead chain (b2s_lookahead.cuh)

// Programmatic dependent launch (sm_90+): a kernel launched with the programmatic-stream-serialization
// attribute may start while its predecessor drains; it must call pdl_wait() before touching anything the
// predecessor wrote.  pdl_trigger() lets the successor's CTAs be scheduled as soon as SM resources free up.
// Both are no-ops for ordinary launches.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// Device-resident loop state: the host only polls it between batches of pivots.
struct DevState {
    int status;       // kRunning until a phase ends (kFeasible = optimal, kUnbounded)
    int live;         // 1 while the current iteration's pivot (q,p) is valid for gather/update
    int q;            // entering variable (0-based; device row 1+q)
    int p;            // leaving constraint (global column index)
    double cq;        // tournament value of the entering reduced cost (src/solver.cu:86)
    double piv;       // a_pq
    double sc;        // (-cq)/piv (src/solver.cu:54)
    long long pivots; // pivots started since build
    long long limit;  // stop when pivots == limit
    unsigned long long hash; // FNV-1a over (q,p)
    unsigned int ticket_ratio;
    unsigned int ticket_cost;
    long long rows_streamed; // statistics
    long long rows_total;
    int any_negated;
    unsigned int tile_ticket;  // dynamic tile scheduler of the update kernel
    unsigned int tile_done;
    unsigned int ticket_gather;  // P2P sharding: CTAs of the owner's gather kernel that have published
    unsigned int bar_count;      // grid barrier of the persistent loop kernel (zeroed by the host per launch)
    unsigned int pad2[3];
};

template <typename real>
struct Limits;
template <>
struct Limits<double> {
    __host__ __device__ static constexpr double big() { return DBL_MAX; }   // src/reduction.cu:54,112
    __host__ __device__ static constexpr double tiny() { return DBL_MIN; }  // src/reduction.cu:171
};
template <>
struct Limits<float> {
    __host__ __device__ static constexpr float big() { return FLT_MAX; }
    __host__ __device__ static constexpr float tiny() { return FLT_MIN; }
};

// include/macro.h:28-42: |x-y| < 1e-9 -> 0, x<y -> -1, else +1.  Arguments are promoted to double
// exactly as the reference's signature does.
__device__ __forceinline__ int cmp3(double x, double y)
{
    const double d = fabs(__dsub_rn(x, y));
    if (d < 1e-9) return 0;
    return x < y ? -1 : 1;
}
__device__ __forceinline__ bool eps_less(double a, double b) { return cmp3(a, b) < 0; }

__device__ __forceinline__ double fma_r(double a, double b, double c) { return __fma_rn(a, b, c); }
__device__ __forceinline__ float fma_r(float a, float b, float c) { return __fmaf_rn(a, b, c); }
__device__ __forceinline__ double div_r(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ float div_r(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ double mul_r(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float mul_r(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double add_r(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float add_r(float a, float b) { return __fadd_rn(a, b); }

// A tournament candidate: value, position, and a tie-break key (position, or the basic variable
// of the row for Bland's leaving rule).
template <typename real>
struct Cand {
    real v;
    int i;
    int k;
};

// "other beats mine".  kRuleReference is the reference comparator (src/reduction.cu:16, :63):
// strictly smaller by at least 1e-9, so ties keep the incumbent.
template <typename real>
__device__ __forceinline__ bool beats(int rule, const Cand<real>& o, const Cand<real>& me)
{
    if (rule == kRuleReference) return eps_less((double)o.v, (double)me.v);
    if (rule == kRuleLowest) return o.v < me.v || (o.v == me.v && (unsigned)o.k < (unsigned)me.k);
    return (unsigned)o.k < (unsigned)me.k;  // Bland entering: lowest improving index
}

// src/reduction.cu:10-22.  shfl_down past the end of the warp returns the caller's own value,
// which compares as a tie and leaves the lane unchanged -- same as the reference.
template <typename real>
__device__ __forceinline__ void warp_tree(int rule, Cand<real>& c)
{
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        Cand<real> o;
        o.v = __shfl_down_sync(0xffffffffu, c.v, off);
        o.i = __shfl_down_sync(0xffffffffu, c.i, off);
        o.k = __shfl_down_sync(0xffffffffu, c.k, off);
        if (beats(rule, o, c)) c = o;
    }
}

template <typename real>
struct TreeSmem {
    real v[32];
    int i[32];
    int k[32];
};

// src/reduction.cu:24-49 for a 512-thread block.  Result valid in thread 0.
template <typename real>
__device__ __forceinline__ void block_tree_512(int rule, Cand<real>& c, TreeSmem<real>& sm)
{
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    warp_tree(rule, c);
    if (lane == 0) {
        sm.v[w] = c.v;
        sm.i[w] = c.i;
        sm.k[w] = c.k;
    }
    __syncthreads();
    if (w == 0) {
        if (lane < kSelBlock / 32) {
            c.v = sm.v[lane];
            c.i = sm.i[lane];
            c.k = sm.k[lane];
        } else {
            c.v = Limits<real>::big();
            c.i = -1;
            c.k = -1;
        }
        warp_tree(rule, c);
    }
    __syncthreads();
}

// Second launch of src/reduction.cu:92-93 (one block of 1024 threads over the G <= 1024 block
// winners), executed by a 512-thread CTA: each thread plays virtual threads t and t+512, i.e.
// each warp plays virtual warps w and w+16.  Result valid in thread 0.
template <typename real>
__device__ __forceinline__ void stage2_1024(int rule, const real* slot_v, const int* slot_i, const int* slot_k, int G,
                                            Cand<real>& out, TreeSmem<real>& sm)
{
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int vt = threadIdx.x + kSelBlock * h;
        Cand<real> c;
        c.v = Limits<real>::big();
        c.i = -1;
        c.k = -1;
        if (vt < G) {
            Cand<real> o;
            o.v = __ldcg(slot_v + vt);
            o.i = __ldcg(slot_i + vt);
            o.k = __ldcg(slot_k + vt);
            if (beats(rule, o, c)) c = o;  // src/reduction.cu:60-71 against (DBL_MAX,-1)
        }
        warp_tree(rule, c);
        if (lane == 0) {
            sm.v[w + 16 * h] = c.v;
            sm.i[w + 16 * h] = c.i;
            sm.k[w + 16 * h] = c.k;
        }
    }
    __syncthreads();
    if (w == 0) {
        out.v = sm.v[lane];
        out.i = sm.i[lane];
        out.k = sm.k[lane];
        warp_tree(rule, out);
    }
    __syncthreads();
}

// Block-wide fmax (src/reduction.cu:143-167); exact and order independent.
template <typename real>
__device__ __forceinline__ real block_max_512(real v, real* sm32)
{
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, off));
    if (lane == 0) sm32[w] = v;
    __syncthreads();
    v = lane < kSelBlock / 32 ? sm32[lane] : Limits<real>::tiny();
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, off));
    __syncthreads();
    return v;
}

// Everything the per-pivot kernels need; passed by value.
template <typename real>
struct PivotParams {
    real* T;              // stored tableau, Rs x ld (local constraint slab when sharded)
    long long ld;         // row pitch in elements (multiple of 64)
    int n;                // structural variables
    int m;                // constraints (global)
    int m_loc;            // constraints in this rank's slab (== m on one GPU)
    int col0;             // global index of the slab's first constraint
    long long Rs;         // stored rows
    long long Rc;         // active reference rows == length of the cost vector in this phase
    long long fold_from;  // cost index >= fold_from lives in stored row (index - m); LLONG_MAX = no folding
    real* cost;           // Rc entries, [0] = objective value
    int* base;            // m entries, basic variable of each constraint
    real* col;            // m_loc: snapshot of the entering column a_.q
    real* s;              // ld: -a_iq / pivot (0 at the pivot column and in the padding)
    real* rowp;           // Rs: raw pivot-constraint entries a_p.
    real* rslot_v;        // ratio stage-1 block winners (global slot index = column block)
    int* rslot_i;
    int* rslot_k;
    real* rslot_max;      // per-block max of the entering column
    real* cslot_v;        // cost stage-1 block winners
    int* cslot_i;
    int* cslot_k;
    DevState* st;
    int2* trace;
    long long trace_cap;
    int rule;
    int skip_zero;
    int Gm;       // ratio stage-1 blocks over the global constraint range
    int Gm_loc0;  // first global block owned by this rank
    int Gm_loc;   // blocks owned by this rank
    int Gc;       // cost stage-1 blocks
    // peer-memory sharding (b2s_p2p.cuh): arena base of every rank as mapped in this process
    int rank;
    int world;
    long long arena_rows;          // capacity of one arena rowp buffer (elements)
    unsigned char* peers[kMaxPeers];
    // update-kernel tiling
    int serpentine; // alternate the sweep direction of the update every pivot (L2 reuse across pivots)
    int log2_tpr;   // log2(threads per tableau row)
    int nchunks;    // column chunks per row
    int tile_groups;  // unrolled row groups per tile
    long long ntiles;
    // look-ahead pivot kernel (b2s_lookahead.cuh)
    LaState* la;
    unsigned* tile_rec;     // per tile: pivot number of the last update that completed it
    real* col2;             // 2 x ld: entering column of the prepared pivot, by pivot parity
    real* s2;               // 2 x ld: -a_q / pivot
    real* rowp2;            // 2 x rowp_stride: raw pivot constraint (one GPU; sharded solves use the arenas)
    int* rowlist;           // 2 x rowp_stride: stored rows the update streams (ascending), by pivot parity
    real* rowval;           // 2 x rowp_stride: their pivot-constraint entries a_pr
    int* rowpos;            // 2 x rowp_stride: row -> position in rowlist, -1 when the row is not streamed
    long long rowp_stride;
    long long wait_cycles;  // bound of every device-side wait (clock64 ticks)
    int helpers;            // CTAs of the update kernel that run the look-ahead chain before they stream
    int la_u;               // 256-bit loads each streaming thread keeps in flight (8, or 4: half the queueing delay for the chain)
    int fault_rank;         // fault injection (tests): this rank stops publishing from pivot fault_pivot on
    long long fault_pivot;
};

template <typename real>
__device__ __forceinline__ long long stored_row(const PivotParams<real>& P, long long cost_index)
{
    return cost_index >= P.fold_from ? cost_index - P.m : cost_index;
}

// ---- peer-memory arena of a sharded solve (see b2s_p2p.cuh) -----------------------------------
template <typename real>
struct ArenaHeader {
    real slot_v[2][kMaxSlots];
    real slot_max[2][kMaxSlots];
    int slot_i[2][kMaxSlots];
    int slot_k[2][kMaxSlots];
    unsigned long long flag_slots[2][kMaxPeers];
    unsigned long long flag_rowp[2];
    unsigned long long pad[6];
    // look-ahead chain (b2s_lookahead.cuh): one flag per (rank, helper) / per owner helper, so nobody waits for a "last" CTA
    unsigned long long la_flag_slots[2][kMaxPeers][kLaMaxHelpers];
    unsigned long long la_flag_rowp[2][kLaMaxHelpers];
};

template <typename real>
__host__ __device__ inline size_t arena_bytes(long long rows)
{
    return sizeof(ArenaHeader<real>) + 2 * sizeof(real) * (size_t)rows;
}
template <typename real>
__device__ __forceinline__ ArenaHeader<real>* arena_of(const PivotParams<real>& P, int r)
{
    return reinterpret_cast<ArenaHeader<real>*>(P.peers[r]);
}
template <typename real>
__device__ __forceinline__ real* arena_rowp(const PivotParams<real>& P, int r, int parity)
{
    return reinterpret_cast<real*>(P.peers[r] + sizeof(ArenaHeader<real>)) + (size_t)parity * P.arena_rows;
}


}  // namespace b2s
