// b2s_p2p.cuh -- the sharded pivot's two exchanges done by the compute kernels themselves over
// NVLink peer memory (no NCCL call, no host, graph-capturable).
//
// Every rank owns an "arena" (cudaMalloc'd, exported with cudaIpcGetMemHandle, mapped by all peers):
//
//   slots  [2][1024] x {value, max, index, key}   ratio-test stage-1 block winners of ALL ranks
//   flag_slots[2][kMaxPeers]                       "rank r has published its winners of pivot #seq"
//   flag_rowp [2]                                  "the owner has published pivot constraint #seq"
//   rowp   [2][arena_rows]                         raw pivot constraint a_p.
//
// Writers store straight into every peer's arena (st.global on mapped peer addresses), fence at
// system scope, then release a sequence number; readers poll their OWN memory.  Buffers are
// indexed by pivot parity: a rank can run at most one pivot ahead of the slowest peer (it cannot
// finish the ratio test of pivot k+1 before every peer has published its winners for k+1, which a
// peer does only after it is done with pivot k), so two generations suffice.
//
//   exchange 1 (ratio_p2p_kernel)  replaces ncclAllGather of the block winners
//   exchange 2 (gather_p2p_kernel / svec_p2p_kernel) replaces the integer-sum ncclAllReduce: the owner
//             of constraint p writes the raw pivot constraint into every arena while it normalises
//             its column -- the transfer rides on the gather it has to do anyway.
#pragma once
#include "b2s_kernels.cuh"

namespace b2s {

constexpr int kStatusPeerTimeout = -99;  // a peer did not publish in time: surfaced as a CUDA-side error status

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// Poll a flag in local memory until it reaches seq.  Bounded: ~2 s at the B200's clock, then give up
// (the caller turns that into kStatusPeerTimeout) so a dead peer cannot hang the GPU.
__device__ __forceinline__ bool wait_flag(const unsigned long long* flag, unsigned long long seq, long long cycles = 4000000000ll)
{
    const long long t0 = clock64();
    while (ld_acquire_sys(flag) < seq) {
        if (clock64() - t0 > cycles) return false;
        __nanosleep(64);
    }
    return true;
}

// ---- exchange 1 --------------------------------------------------------------------------------
template <typename real>
__global__ void __launch_bounds__(kSelBlock) ratio_p2p_kernel(PivotParams<real> P)
{
    __shared__ TreeSmem<real> sm;
    __shared__ real smax[32];
    __shared__ int s_flag;
    __shared__ Cand<real> s_win;
    __shared__ real s_mx;
    DevState* st = P.st;
    const int status = __ldcg(&st->status);
    const long long pivots = __ldcg(&st->pivots), limit = __ldcg(&st->limit);
    if (status != kRunning || pivots >= limit) {
        if (blockIdx.x == 0 && threadIdx.x == 0) st->live = 0;
        return;
    }
    const unsigned long long seq = (unsigned long long)pivots + 1ull;
    const int par = (int)(seq & 1ull);
    const int q = __ldcg(&st->q);
    const real* qrow = P.T + stored_row(P, 1 + (long long)q) * P.ld;
    const real* brow = P.T;
    const int rule = P.rule;
    const int tree_rule = (rule == kRuleReference) ? kRuleReference : kRuleLowest;

    const int gb = P.Gm_loc0 + blockIdx.x;
    Cand<real> c;
    c.v = Limits<real>::big();
    c.i = -1;
    c.k = -1;
    real mx = Limits<real>::tiny();
    {
        // one element per thread: prepare() refuses sharded problems with more than kSelBlock * kMaxSlots constraints, so the
        // Gm_loc CTAs launched here cover the slab exactly once (b2s_solver.cu, prepare)
        const long long gi = (long long)gb * kSelBlock + threadIdx.x;
        if (gi < P.m) {
            const long long li = gi - P.col0;
            const real a = qrow[li];
            const real bb = brow[li];
            P.col[li] = a;
            mx = fmax(mx, a);
            Cand<real> o;
            o.v = (cmp3((double)a, 0.0) > 0) ? div_r(bb, a) : Limits<real>::big();
            o.i = (int)gi;
            o.k = (rule == kRuleBland) ? ((o.v < Limits<real>::big()) ? P.base[gi] : -1) : (int)gi;
            if (beats(tree_rule, o, c)) c = o;
        }
    }
    mx = block_max_512(mx, smax);
    block_tree_512(tree_rule, c, sm);
    if (threadIdx.x == 0) {
        s_win = c;
        s_mx = mx;
    }
    __syncthreads();
    // publish this block's winner into every rank's arena (own included)
    if (threadIdx.x < P.world) {
        ArenaHeader<real>* a = arena_of(P, threadIdx.x);
        a->slot_v[par][gb] = s_win.v;
        a->slot_max[par][gb] = s_mx;
        a->slot_i[par][gb] = s_win.i;
        a->slot_k[par][gb] = s_win.k;
        __threadfence_system();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned t = atomicAdd(&st->ticket_ratio, 1u);
        s_flag = (t == (unsigned)P.Gm_loc - 1u);
    }
    __syncthreads();
    if (!s_flag) return;
    // last CTA of this rank: all local winners are out -> raise our flag everywhere, wait for everyone's
    __threadfence_system();
    const bool mute = P.fault_rank == P.rank && (long long)seq >= P.fault_pivot;   // fault injection (tests): this rank goes silent
    if (threadIdx.x < P.world && !mute) st_release_sys(&arena_of(P, threadIdx.x)->flag_slots[par][P.rank], seq);
    int ok = 1;
    if (threadIdx.x < P.world) ok = wait_flag(&arena_of(P, P.rank)->flag_slots[par][threadIdx.x], seq, P.wait_cycles) ? 1 : 0;
    ok = __syncthreads_and(ok);
    if (!ok) {
        if (threadIdx.x == 0) {
            st->ticket_ratio = 0;
            st->status = kStatusPeerTimeout;
            st->live = 0;
        }
        return;
    }
    __threadfence_system();
    const ArenaHeader<real>* mine = arena_of(P, P.rank);
    ratio_finish(P, mine->slot_v[par], mine->slot_i[par], mine->slot_k[par], mine->slot_max[par], sm, smax);
}

// ---- exchange 2, owner side ----------------------------------------------------------------------
template <typename real>
__global__ void __launch_bounds__(256) gather_p2p_kernel(PivotParams<real> P)
{
    __shared__ int s_flag;
    DevState* st = P.st;
    if (!__ldcg(&st->live)) return;
    const int p = __ldcg(&st->p);
    const int lp = p - P.col0;
    if (lp < 0 || lp >= P.m_loc) return;  // not the owner: svec_p2p_kernel waits for the owner's flag
    const unsigned long long seq = (unsigned long long)__ldcg(&st->pivots);  // already counts this pivot
    const int par = (int)(seq & 1ull);
    const real piv = P.col[lp];
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < P.Rs) {
        real* e = P.T + t * P.ld + lp;
        const real a = *e;
        *e = div_r(a, piv);
        for (int w = 0; w < P.world; ++w) arena_rowp(P, w, par)[t] = a;
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned done = atomicAdd(&st->ticket_gather, 1u);
        s_flag = (done == gridDim.x - 1u);
    }
    __syncthreads();
    if (!s_flag) return;
    __threadfence_system();
    if (threadIdx.x < P.world) st_release_sys(&arena_of(P, threadIdx.x)->flag_rowp[par], seq);
    if (threadIdx.x == 0) st->ticket_gather = 0;
}

// ---- exchange 2, every rank: wait for the pivot constraint, localise it, build s ---------------------
template <typename real>
__global__ void __launch_bounds__(256) svec_p2p_kernel(PivotParams<real> P)
{
    __shared__ int s_ok;
    DevState* st = P.st;
    if (!__ldcg(&st->live)) return;
    const unsigned long long seq = (unsigned long long)__ldcg(&st->pivots);
    const int par = (int)(seq & 1ull);
    if (threadIdx.x == 0) s_ok = wait_flag(&arena_of(P, P.rank)->flag_rowp[par], seq, P.wait_cycles) ? 1 : 0;
    __syncthreads();
    if (!s_ok) {
        if (threadIdx.x == 0) st->status = kStatusPeerTimeout;  // update_kernel still sees live: make it stop too
        if (threadIdx.x == 0) st->live = 0;
        return;
    }
    const real* src = arena_rowp(P, P.rank, par);
    const int lp = __ldcg(&st->p) - P.col0;
    const real piv = __ldcg(src + stored_row(P, 1 + (long long)__ldcg(&st->q)));
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long r = t; r < P.Rs; r += stride) P.rowp[r] = __ldcg(src + r);
    if (t < P.ld) {
        real sv = 0;
        if (t < P.m_loc && t != lp) sv = div_r(-P.col[t], piv);
        P.s[t] = sv;
    }
    if (t == 0) {
        st->piv = (double)piv;
        st->sc = (double)div_r((real)(-st->cq), piv);
    }
}

}  // namespace b2s
