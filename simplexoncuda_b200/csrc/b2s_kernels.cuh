// b2s_kernels.cuh -- the kernels of the two-phase dense-tableau simplex (sm_100a).
#pragma once
#include "b2s_device.cuh"

namespace b2s {

// ---------------------------------------------------------------------------------------------
// 128/256-bit global accesses with cache hints.  The tableau is streamed once per pivot and is
// far larger than L2 at the sizes that matter, so variants with streaming hints exist.
// ---------------------------------------------------------------------------------------------
template <int VB>
struct Pack;
template <>
struct alignas(16) Pack<16> {
    unsigned long long w[2];
};
template <>
struct alignas(32) Pack<32> {
    unsigned long long w[4];
};

template <int HINT>
__device__ __forceinline__ Pack<16> ld_pack(const Pack<16>* p)
{
    Pack<16> r;
    if (HINT == 0)
        asm volatile("ld.global.v2.b64 {%0,%1}, [%2];" : "=l"(r.w[0]), "=l"(r.w[1]) : "l"(p));
    else if (HINT == 1)
        asm volatile("ld.global.cs.v2.b64 {%0,%1}, [%2];" : "=l"(r.w[0]), "=l"(r.w[1]) : "l"(p));
    else if (HINT == 3)
        asm volatile("ld.global.cg.v2.b64 {%0,%1}, [%2];" : "=l"(r.w[0]), "=l"(r.w[1]) : "l"(p));
    else
        asm volatile("ld.global.L1::no_allocate.v2.b64 {%0,%1}, [%2];" : "=l"(r.w[0]), "=l"(r.w[1]) : "l"(p));
    return r;
}
template <int HINT>
__device__ __forceinline__ void st_pack(Pack<16>* p, const Pack<16>& r)
{
    if (HINT == 1)
        asm volatile("st.global.cs.v2.b64 [%0], {%1,%2};" ::"l"(p), "l"(r.w[0]), "l"(r.w[1]) : "memory");
    else
        asm volatile("st.global.v2.b64 [%0], {%1,%2};" ::"l"(p), "l"(r.w[0]), "l"(r.w[1]) : "memory");
}
template <int HINT>
__device__ __forceinline__ Pack<32> ld_pack(const Pack<32>* p)
{
    Pack<32> r;
    if (HINT == 0)
        asm volatile("ld.global.v4.b64 {%0,%1,%2,%3}, [%4];"
                     : "=l"(r.w[0]), "=l"(r.w[1]), "=l"(r.w[2]), "=l"(r.w[3])
                     : "l"(p));
    else if (HINT == 1)
        asm volatile("ld.global.cs.v4.b64 {%0,%1,%2,%3}, [%4];"
                     : "=l"(r.w[0]), "=l"(r.w[1]), "=l"(r.w[2]), "=l"(r.w[3])
                     : "l"(p));
    else if (HINT == 3)
        asm volatile("ld.global.cg.v4.b64 {%0,%1,%2,%3}, [%4];"
                     : "=l"(r.w[0]), "=l"(r.w[1]), "=l"(r.w[2]), "=l"(r.w[3])
                     : "l"(p));
    else
        asm volatile("ld.global.L1::no_allocate.v4.b64 {%0,%1,%2,%3}, [%4];"
                     : "=l"(r.w[0]), "=l"(r.w[1]), "=l"(r.w[2]), "=l"(r.w[3])
                     : "l"(p));
    return r;
}
template <int HINT>
__device__ __forceinline__ void st_pack(Pack<32>* p, const Pack<32>& r)
{
    if (HINT == 1)
        asm volatile("st.global.cs.v4.b64 [%0], {%1,%2,%3,%4};" ::"l"(p), "l"(r.w[0]), "l"(r.w[1]), "l"(r.w[2]),
                     "l"(r.w[3])
                     : "memory");
    else
        asm volatile("st.global.v4.b64 [%0], {%1,%2,%3,%4};" ::"l"(p), "l"(r.w[0]), "l"(r.w[1]), "l"(r.w[2]),
                     "l"(r.w[3])
                     : "memory");
}

template <typename real, int VB>
union PackView {
    Pack<VB> p;
    real e[VB / sizeof(real)];
    __device__ PackView() {}
};

// ---------------------------------------------------------------------------------------------
// Entering-column tournament over the cost vector (src/reduction.cu:51-104 applied to
// costsVector+1, src/solver.cu:86), optionally fused with the cost update of src/solver.cu:48-56.
// Executed by the first CTAs of the update kernel (kUpdate = true) or on its own for the first
// pivot of a phase.  Block b plays reference stage-1 block b; the CTA that draws the last ticket
// plays the stage-2 block and publishes (cq, q) and the optimality verdict (src/solver.cu:87-88).
// ---------------------------------------------------------------------------------------------
// Read-only vectors are fetched through the non-coherent path in the per-pivot kernels (they were
// written by an earlier launch); inside the persistent loop kernel they change between grid barriers
// and must come from L2 (COH).
template <bool COH, typename X>
__device__ __forceinline__ X ld_vec(const X* p)
{
    return COH ? __ldcg(p) : __ldg(p);
}

template <typename real, bool kUpdate, bool COH = false>
__device__ __forceinline__ void cost_select_blocks(const PivotParams<real>& P, const real* rowp, real sc,
                                                   TreeSmem<real>& sm, int* s_flag)
{
    const long long Nc = P.Rc - 1;
    const int rule = P.rule;
    for (int b = blockIdx.x; b < P.Gc; b += gridDim.x) {
        Cand<real> c;
        c.v = Limits<real>::big();
        c.i = -1;
        c.k = -1;
        for (long long i = (long long)b * kSelBlock + threadIdx.x; i < Nc; i += (long long)kSelBlock * P.Gc) {
            const long long j = 1 + i;
            real v = COH ? __ldcg(P.cost + j) : P.cost[j];
            if (kUpdate) {
                v = fma_r(sc, ld_vec<COH>(rowp + stored_row(P, j)), v);  // src/solver.cu:54
                P.cost[j] = v;
            }
            Cand<real> o;
            o.v = v;
            o.i = (int)i;
            o.k = (rule == kRuleBland) ? (cmp3((double)v, 0.0) < 0 ? (int)i : -1) : (int)i;
            if (beats(rule, o, c)) c = o;
        }
        if (kUpdate && b == 0 && threadIdx.x == 0)
            P.cost[0] = fma_r(sc, ld_vec<COH>(rowp), COH ? __ldcg(P.cost) : P.cost[0]);  // objective
        block_tree_512(rule, c, sm);
        if (threadIdx.x == 0) {
            P.cslot_v[b] = c.v;
            P.cslot_i[b] = c.i;
            P.cslot_k[b] = c.k;
            __threadfence();
            const unsigned t = atomicAdd(&P.st->ticket_cost, 1u);
            *s_flag = (t == (unsigned)P.Gc - 1u);
        }
        __syncthreads();
        const bool last = *s_flag != 0;
        __syncthreads();
        if (last) {
            __threadfence();
            Cand<real> w;
            if (P.Gc > 1) {
                stage2_1024(rule, P.cslot_v, P.cslot_i, P.cslot_k, P.Gc, w, sm);
            } else {
                w = c;  // thread 0 holds the single block's winner
            }
            if (threadIdx.x == 0) {
                DevState* st = P.st;
                st->q = w.i;
                st->cq = (double)w.v;
                st->ticket_cost = 0;
                if (!(w.i >= 0 && cmp3((double)w.v, 0.0) < 0)) st->status = kFeasible;  // optimal for this phase
                __threadfence();
            }
        }
    }
}

template <typename real>
__global__ void __launch_bounds__(kSelBlock) select_kernel(PivotParams<real> P)
{
    __shared__ TreeSmem<real> sm;
    __shared__ int s_flag;
    cost_select_blocks<real, false>(P, P.rowp, (real)0, sm, &s_flag);
}

// ---------------------------------------------------------------------------------------------
// ratio_kernel: one CTA per reference stage-1 block of 512 constraints.
//   col[i]   = T[1+q][i]                                  (src/solver.cu:90-94)
//   max_i col[i] < 1e-9  -> UNBOUNDED                      (src/reduction.cu:186-201)
//   ratio_i  = col[i] >= 1e-9 ? b_i / col[i] : DBL_MAX     (src/reduction.cu:106-114)
//   p        = tournament(ratio)                           (src/reduction.cu:116-140)
//   base[p]  = q                                           (src/solver.cu:105)
// When the tableau is sharded the stage-2 part runs in ratio_finish_kernel after the all-gather.
// ---------------------------------------------------------------------------------------------
template <typename real>
__device__ __forceinline__ void ratio_finish(const PivotParams<real>& P, const real* slot_v, const int* slot_i,
                                             const int* slot_k, const real* slot_max, TreeSmem<real>& sm, real* smax)
{
    // global max of the entering column
    real mx = Limits<real>::tiny();
    for (int b = threadIdx.x; b < P.Gm; b += kSelBlock) mx = fmax(mx, __ldcg(slot_max + b));
    mx = block_max_512(mx, smax);
    const int tree_rule = (P.rule == kRuleReference) ? kRuleReference : kRuleLowest;
    Cand<real> w;
    if (P.Gm > 1) {
        stage2_1024(tree_rule, slot_v, slot_i, slot_k, P.Gm, w, sm);
    } else {
        w.v = __ldcg(slot_v);
        w.i = __ldcg(slot_i);
        w.k = __ldcg(slot_k);
    }
    if (threadIdx.x == 0) {
        DevState* st = P.st;
        st->ticket_ratio = 0;
        if (cmp3((double)mx, 0.0) <= 0 || w.i < 0) {
            st->status = kUnbounded;
            st->live = 0;
        } else {
            const int p = w.i, q = st->q;
            st->p = p;
            P.base[p] = q;
            const long long k = st->pivots;
            if (k < P.trace_cap) P.trace[k] = make_int2(q, p);
            unsigned long long h = st->hash;
            const unsigned int words[2] = {(unsigned)q, (unsigned)p};
#pragma unroll
            for (int wd = 0; wd < 2; ++wd)
#pragma unroll
                for (int by = 0; by < 4; ++by) {
                    h ^= (words[wd] >> (8 * by)) & 0xffu;
                    h *= 1099511628211ULL;
                }
            st->hash = h;
            st->pivots = k + 1;
            st->live = 1;
        }
        __threadfence();
    }
}

template <typename real, bool kSharded>
__global__ void __launch_bounds__(kSelBlock) ratio_kernel(PivotParams<real> P)
{
    __shared__ TreeSmem<real> sm;
    __shared__ real smax[32];
    __shared__ int s_flag;
    pdl_wait();
    pdl_trigger();
    DevState* st = P.st;
    const int status = __ldcg(&st->status);
    const long long pivots = __ldcg(&st->pivots), limit = __ldcg(&st->limit);
    if (status != kRunning || pivots >= limit) {
        if (blockIdx.x == 0 && threadIdx.x == 0) st->live = 0;
        return;
    }
    const int q = __ldcg(&st->q);
    const real* qrow = P.T + stored_row(P, 1 + (long long)q) * P.ld;
    const real* brow = P.T;
    const int rule = P.rule;
    const int tree_rule = (rule == kRuleReference) ? kRuleReference : kRuleLowest;

    const int gb = P.Gm_loc0 + blockIdx.x;  // global stage-1 block id
    Cand<real> c;
    c.v = Limits<real>::big();
    c.i = -1;
    c.k = -1;
    real mx = Limits<real>::tiny();
    for (long long gi = (long long)gb * kSelBlock + threadIdx.x; gi < P.m; gi += (long long)kSelBlock * P.Gm) {
        const long long li = gi - P.col0;  // local column
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
    mx = block_max_512(mx, smax);
    block_tree_512(tree_rule, c, sm);
    if (threadIdx.x == 0) {
        P.rslot_v[gb] = c.v;
        P.rslot_i[gb] = c.i;
        P.rslot_k[gb] = c.k;
        P.rslot_max[gb] = mx;
    }
    if (kSharded) return;  // stage 2 after the all-gather
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned t = atomicAdd(&st->ticket_ratio, 1u);
        s_flag = (t == (unsigned)P.Gm - 1u);
    }
    __syncthreads();
    if (!s_flag) return;
    __threadfence();
    ratio_finish(P, P.rslot_v, P.rslot_i, P.rslot_k, P.rslot_max, sm, smax);
}

template <typename real>
__global__ void __launch_bounds__(kSelBlock) ratio_finish_kernel(PivotParams<real> P)
{
    __shared__ TreeSmem<real> sm;
    __shared__ real smax[32];
    DevState* st = P.st;
    const int status = __ldcg(&st->status);
    const long long pivots = __ldcg(&st->pivots), limit = __ldcg(&st->limit);
    if (status != kRunning || pivots >= limit) return;  // ratio_kernel already cleared `live`
    ratio_finish(P, P.rslot_v, P.rslot_i, P.rslot_k, P.rslot_max, sm, smax);
}

// ---------------------------------------------------------------------------------------------
// gather_kernel: rowp[r] = T[r][p] (raw, src/solver.cu:24-32), pivot column normalised in place
// T[r][p] = rowp[r]/pivot (src/solver.cu:43, the `col == colPivotIndex` arm), and
// s[i] = (-col[i])/pivot with s[p] = 0 so that the streaming kernel is branch free:
// fma(0, rowp[r], T[r][p]) returns the already normalised entry bit for bit.
// ---------------------------------------------------------------------------------------------
template <typename real, bool kSharded>
__global__ void __launch_bounds__(256) gather_kernel(PivotParams<real> P)
{
    pdl_wait();
    pdl_trigger();
    DevState* st = P.st;
    if (!__ldcg(&st->live)) return;
    const int p = __ldcg(&st->p);
    const int lp = p - P.col0;  // single GPU: col0 == 0
    const bool owner = lp >= 0 && lp < P.m_loc;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (kSharded) {
        // Only the rank that owns constraint p holds the pivot-constraint entries; the others
        // contribute all-zero bit patterns to the integer sum all-reduce that follows, which
        // therefore delivers the owner's bits unchanged to every rank.
        if (t < P.Rs) {
            real a = 0;
            if (owner) {
                real* e = P.T + t * P.ld + lp;
                a = *e;
                *e = div_r(a, P.col[lp]);
            }
            P.rowp[t] = a;
        }
        return;
    }
    const real piv = P.col[lp];
    int nz = 0;
    if (t < P.Rs) {
        real* e = P.T + t * P.ld + lp;
        const real a = *e;
        P.rowp[t] = a;
        *e = div_r(a, piv);
        nz = (a != (real)0);
    }
    if (t < P.ld) {
        real sv = 0;
        if (t < P.m_loc && t != lp) sv = div_r(-P.col[t], piv);
        P.s[t] = sv;
    }
    if (t == 0) {
        st->piv = (double)piv;
        st->sc = (double)div_r((real)(-st->cq), piv);
    }
    if (P.skip_zero) {
        nz = __syncthreads_count(nz);
        if (threadIdx.x == 0 && nz) atomicAdd((unsigned long long*)&st->rows_streamed, (unsigned long long)nz);
    }
}

// Sharded solves: after the all-reduce every rank holds the raw pivot constraint; the pivot is
// its entry in the entering variable's row (a_pq = T[1+q][p]).
template <typename real>
__global__ void __launch_bounds__(256) svec_kernel(PivotParams<real> P)
{
    DevState* st = P.st;
    if (!__ldcg(&st->live)) return;
    const int lp = __ldcg(&st->p) - P.col0;
    const real piv = P.rowp[stored_row(P, 1 + (long long)__ldcg(&st->q))];
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
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

// ---------------------------------------------------------------------------------------------
// stream_tiles: the tile loop of the rank-1 update, shared by update_kernel (one launch per pivot)
// and pivot_loop_kernel (persistent).  `svec` holds s = -a_q/pivot; when it is null the thread
// derives its s values from the entering-column snapshot (`colv`), the pivot and the pivot column
// index lp -- the same division, so the same bits.
// ---------------------------------------------------------------------------------------------
template <typename real, int VB, int U, int HINT, bool SKIP, bool DYN, bool COH>
__device__ __forceinline__ void stream_tiles(const PivotParams<real>& P, const real* rowp, const real* svec,
                                             const real* colv, real piv, long long lp, long long* s_next,
                                             bool reverse = false)
{
    constexpr int EPT = VB / (int)sizeof(real);
    const int tx = threadIdx.x & ((1 << P.log2_tpr) - 1);
    const int ty = threadIdx.x >> P.log2_tpr;
    const int rpp = kSelBlock >> P.log2_tpr;           // rows per pass
    const long long tile_rows = (long long)rpp * U * P.tile_groups;
    const long long chunk_cols = (long long)EPT << P.log2_tpr;
    int cur_chunk = -1;
    real sreg[EPT];
    // Tiles are handed out either statically (tile += gridDim) or, DYN, by a device-wide ticket
    // counter fetched one tile ahead, so SMs that stream faster simply take more tiles and the
    // whole grid walks the tableau as one compact address window.
    long long tile = blockIdx.x;
    while (tile < P.ntiles) {
        if (DYN && threadIdx.x == 0)
            *s_next = (long long)atomicAdd(&P.st->tile_ticket, 1u) + gridDim.x;
        // Serpentine sweep: odd pivots walk the tableau backwards, so the tiles the previous pivot touched last
        // -- still sitting (dirty) in the 126 MB L2 -- are the first ones this pivot reads and rewrites, and
        // never travel to HBM in between.  (Explicit L2 evict_first / evict_last hints on top of this were
        // measured and changed nothing: profiles/r01_scaling_and_loop_modes.md.)
        const long long tmap = reverse ? (P.ntiles - 1 - tile) : tile;
        const int chunk = (int)(tmap % P.nchunks);
        const long long rb = tmap / P.nchunks;
        const long long c = chunk * chunk_cols + (long long)tx * EPT;
        if (c < P.ld) {
            const long long r0 = rb * tile_rows + ty;
            for (int g = 0; g < P.tile_groups; ++g) {
                real a[U];
                Pack<VB>* ptr[U];
                PackView<real, VB> v[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const long long r = r0 + (long long)(g * U + u) * rpp;
                    a[u] = (r < P.Rs) ? ld_vec<COH>(rowp + r) : (real)0;
                    ptr[u] = reinterpret_cast<Pack<VB>*>(P.T + r * P.ld + c);
                    if (!SKIP && !(r < P.Rs)) ptr[u] = nullptr;
                    if (SKIP && a[u] == (real)0) ptr[u] = nullptr;
                }
#pragma unroll
                for (int u = 0; u < U; ++u)
                    if (ptr[u]) v[u].p = ld_pack<HINT>(ptr[u]);
                // the thread's s values are (re)computed while the tile's loads are in flight
                if (chunk != cur_chunk) {
                    cur_chunk = chunk;
#pragma unroll
                    for (int e = 0; e < EPT; ++e) {
                        if (svec) {
                            sreg[e] = ld_vec<COH>(svec + c + e);
                        } else {
                            const long long ci = c + e;
                            sreg[e] = (ci < P.m_loc && ci != lp) ? div_r(-ld_vec<COH>(colv + ci), piv) : (real)0;
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < U; ++u)
                    if (ptr[u]) {
#pragma unroll
                        for (int e = 0; e < EPT; ++e) v[u].e[e] = fma_r(sreg[e], a[u], v[u].e[e]);
                        st_pack<HINT>(ptr[u], v[u].p);
                    }
            }
        }
        if (DYN) {
            __syncthreads();
            tile = *s_next;
            __syncthreads();
        } else {
            tile += gridDim.x;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// update_kernel: the HBM-bound kernel.  T[r][i] = fma(s[i], rowp[r], T[r][i]) over the whole
// stored tableau, read once and written once (2*Rs*ld*sizeof(real) bytes), persistent CTAs of
// 512 threads.  Each thread owns VB bytes of consecutive columns (its s values stay in
// registers) and walks rows in unrolled groups of U independent 128/256-bit loads.  The first Gc
// CTAs first update the cost vector and run the next pivot's entering-column tournament.
// ---------------------------------------------------------------------------------------------
template <typename real, int VB, int U, int HINT, bool SKIP, bool DYN>
__global__ void __launch_bounds__(kSelBlock, (VB * U <= 128) ? 2 : 1) update_kernel(PivotParams<real> P)
{
    __shared__ TreeSmem<real> sm;
    __shared__ int s_flag;
    __shared__ long long s_next;
    pdl_wait();
    pdl_trigger();  // the next pivot's ratio CTAs may take the SMs this grid frees while it drains
    if (!__ldcg(&P.st->live)) return;

    const real* rowp = P.rowp;
    if (blockIdx.x < P.Gc) cost_select_blocks<real, true>(P, rowp, (real)__ldcg(&P.st->sc), sm, &s_flag);

    const bool reverse = P.serpentine && (__ldcg(&P.st->pivots) & 1);
    stream_tiles<real, VB, U, HINT, SKIP, DYN, false>(P, rowp, P.s, nullptr, (real)0, -1, &s_next, reverse);
    if (DYN) {
        // the last CTA to leave re-arms the ticket counters for the next launch
        if (threadIdx.x == 0) {
            __threadfence();
            const unsigned done = atomicAdd(&P.st->tile_done, 1u);
            if (done == gridDim.x - 1) {
                P.st->tile_ticket = 0;
                P.st->tile_done = 0;
                __threadfence();
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Price-out (src/gaussian.cu:98-162): cost[y] -= sum_x T[y][x]*coef[x] with the reference's pair
// terms fma(T[y][x],coef[x], T[y][x+32]*coef[x+32]), x = lane + 64k, accumulated into cost[y]
// one term at a time.  The reference's fp64 atomicAdd order is unspecified; here the order is
// fixed to ascending x (one warp per cost entry, lane 0 owns the running sum) so that results are
// reproducible and equal to the oracle's.
// ---------------------------------------------------------------------------------------------
template <typename real>
__global__ void __launch_bounds__(256) coef_kernel(PivotParams<real> P, real* coef)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < P.m_loc) {
        const long long j = 1 + (long long)P.base[P.col0 + i];   // src/gaussian.cu:119-127
        coef[i] = j < P.Rc ? P.cost[j] : (real)0;                 // (drive-out mode: a redundant constraint's artificial in phase 2)
    }
}

template <typename real>
__global__ void __launch_bounds__(256) priceout_kernel(PivotParams<real> P, const real* __restrict__ coef)
{
    const int lane = threadIdx.x & 31;
    const long long y = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (y >= P.Rc) return;
    const real* row = P.T + stored_row(P, y) * P.ld;
    real acc = P.cost[y];
    const int m = P.m_loc;
    for (int k0 = 0; k0 < m; k0 += 64) {
        const int x = k0 + lane;
        real term = 0;
        bool valid = x < m;
        if (valid) {
            if (x + 32 < m)
                term = fma_r(row[x], coef[x], mul_r(row[x + 32], coef[x + 32]));
            else
                term = mul_r(row[x], coef[x]);
        }
        const int cnt = min(32, m - k0);
        for (int l = 0; l < cnt; ++l) {
            const real tl = __shfl_sync(0xffffffffu, term, l);
            acc = add_r(acc, -tl);
        }
    }
    if (lane == 0) P.cost[y] = acc;
}

// ---------------------------------------------------------------------------------------------
// Tableau build (src/twoPhaseMethod.cu:145-200).  Rows 1..n already hold A and row 0 holds b
// (copied or generated in place); slack rows were zero-filled.
// ---------------------------------------------------------------------------------------------
template <typename real>
__global__ void __launch_bounds__(256) build_misc_kernel(PivotParams<real> P, int* neg, int store_artificials)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int n = P.n, m = P.m;
    if (t < P.m_loc) {
        const long long gi = P.col0 + t;
        const int ng = cmp3((double)P.T[t], 0.0) < 0;  // :100-111 checkColumns on the RHS
        neg[t] = ng;
        if (ng) P.st->any_negated = 1;
        const real one = ng ? (real)-1 : (real)1;  // identity entry, negated with its constraint (:86-98)
        P.T[(1 + (long long)n + gi) * P.ld + t] = one;
        if (store_artificials) P.T[(1 + (long long)n + m + gi) * P.ld + t] = one;
    }
    if (t < m) P.base[t] = n + m + (int)t;                                            // :44-52
    if (t < P.Rc) P.cost[t] = (t > (long long)n + m) ? (real)1 : (real)0;             // :150-156
}

template <typename real>
__global__ void __launch_bounds__(256) negate_kernel(PivotParams<real> P, const int* __restrict__ neg)
{
    if (!__ldcg(&P.st->any_negated)) return;
    const long long rows = 1 + (long long)P.n;  // RHS row + structural rows; identities were written negated
    const long long total = rows * P.m_loc;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        const long long r = t / P.m_loc, i = t % P.m_loc;
        if (neg[i]) {
            real* e = P.T + r * P.ld + i;
            *e = -*e;
        }
    }
}

// Phase switch (src/twoPhaseMethod.cu:306-318): slack costs = 0, structural costs = -c; cost[0] kept.
template <typename real>
__global__ void __launch_bounds__(256) phase2_costs_kernel(PivotParams<real> P, const real* __restrict__ c)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < P.n)
        P.cost[1 + t] = -c[t];
    else if (t < (long long)P.n + P.m)
        P.cost[1 + t] = (real)0;
}

// Phase-1 verdict (src/twoPhaseMethod.cu:265-268 and :206-223) -> out[0] infeasible, out[1] #artificials in basis.
template <typename real>
__global__ void __launch_bounds__(256) verdict_kernel(PivotParams<real> P, int* out)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t == 0 && cmp3((double)P.cost[0], 0.0) < 0) out[0] = 1;
    if (t < P.m) {
        const int v = P.base[t];
        if (v >= P.n + P.m && v < P.n + 2 * P.m) atomicAdd(out + 1, 1);
    }
}

// Solution (src/twoPhaseMethod.cu:116-128): x[base[i]] = b_i for structural basics (x pre-zeroed).
template <typename real>
__global__ void __launch_bounds__(256) solution_kernel(PivotParams<real> P, double* x)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < P.m_loc) {
        const int v = P.base[P.col0 + t];
        if (v < P.n) x[v] = (double)P.T[t];
    }
}

// Expand the stored tableau into the reference's unfolded rows_active x m layout (tests).
template <typename real>
__global__ void __launch_bounds__(256) export_kernel(PivotParams<real> P, double* out)
{
    const long long total = P.Rc * P.m_loc;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        const long long r = t / P.m_loc, i = t % P.m_loc;
        out[t] = (double)P.T[stored_row(P, r) * P.ld + i];
    }
}

// Synthetic dense pivot for b2s_bench_update: s and rowp dense and small so values stay finite.
template <typename real>
__global__ void __launch_bounds__(256) bench_fill_kernel(PivotParams<real> P)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < P.ld) P.s[t] = (t < P.m_loc) ? (real)(1e-3 * (double)((t * 2654435761u) % 1000u) / 1000.0 - 5e-4) : (real)0;
    if (t < P.Rs) P.rowp[t] = (real)(1e-3 * (double)((t * 40503u) % 1000u) / 1000.0 + 1e-4);
    if (t == 0) {
        P.st->live = 1;
        P.st->sc = 0.0;
        P.st->status = kRunning;
    }
}

// b2s_bench_update: advance the pivot counter between launches so the sweep direction alternates as in a solve
__global__ void bench_bump_kernel(DevState* st) { st->pivots += 1; }

// Stand-alone vector primitives behind the reference's reduction.cuh entry points.
template <typename real>
__global__ void __launch_bounds__(256) ratio_vector_kernel(const real* __restrict__ known, const real* __restrict__ column,
                                                          long long n, real* out)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (cmp3((double)column[i], 0.0) > 0) ? div_r(known[i], column[i]) : Limits<real>::big();  // src/reduction.cu:106-114
}
template <typename real>
__global__ void __launch_bounds__(kSelBlock) max_vector_kernel(const real* __restrict__ v, long long n, real* out)
{
    __shared__ real smax[32];
    real mx = Limits<real>::tiny();  // src/reduction.cu:171
    for (long long i = threadIdx.x; i < n; i += kSelBlock) mx = fmax(mx, v[i]);
    mx = block_max_512(mx, smax);
    if (threadIdx.x == 0) *out = mx;
}

// Loop-state (re)initialisation: fresh = 1 at build (counters and hash restart), 0 at the phase switch.
__global__ void state_reset_kernel(DevState* st, int fresh)
{
    st->status = kRunning;
    st->live = 0;
    st->ticket_ratio = 0;
    st->ticket_cost = 0;
    st->tile_ticket = 0;
    st->tile_done = 0;
    if (fresh) {
        st->pivots = 0;
        st->hash = 1469598103934665603ULL;
        st->rows_streamed = 0;
    }
    st->limit = st->pivots;
}

// fp64 staging <-> working precision (used only when real == float)
template <typename real>
__global__ void __launch_bounds__(256) convert_rows(real* dst, long long dst_ld, const double* src, long long src_ld,
                                                    long long rows, int cols)
{
    const long long total = rows * cols;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        const long long r = t / cols, c = t % cols;
        dst[r * dst_ld + c] = (real)src[r * src_ld + c];
    }
}
template <typename real>
__global__ void __launch_bounds__(256) widen_rows(double* dst, long long dst_ld, const real* src, long long src_ld,
                                                  long long rows, int cols)
{
    const long long total = rows * cols;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        const long long r = t / cols, c = t % cols;
        dst[r * dst_ld + c] = (double)src[r * src_ld + c];
    }
}

// ---------------------------------------------------------------------------------------------
// Driving artificial variables out of the basis (opt-in, beyond the reference: it stops with DEGENERATE at
// src/twoPhaseMethod.cu:206-223, :274-282).  find: lowest-index structural/slack variable with a non-zero entry
// (|a| >= 1e-9, the reference's own threshold) in constraint i; setup: make (q, i) the current pivot for the ordinary
// gather + update kernels.
// ---------------------------------------------------------------------------------------------
template <typename real>
__global__ void __launch_bounds__(256) driveout_find_kernel(PivotParams<real> P, int i, int* out)
{
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < (long long)P.n + P.m && cmp3(fabs((double)P.T[(1 + j) * P.ld + i]), 0.0) > 0) atomicMin(out, (int)j);
}
template <typename real>
__global__ void __launch_bounds__(256) driveout_setup_kernel(PivotParams<real> P, int q, int p)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < P.m_loc) P.col[t] = P.T[(1 + (long long)q) * P.ld + t];   // entering column snapshot (src/solver.cu:90-94)
    if (t == 0) {
        DevState* st = P.st;
        st->q = q;
        st->cq = (double)P.cost[1 + q];
        st->p = p;
        P.base[p] = q;
        const long long k = st->pivots;
        if (k < P.trace_cap) P.trace[k] = make_int2(q, p);
        unsigned long long h = st->hash;
        const unsigned int words[2] = {(unsigned)q, (unsigned)p};
        for (int wd = 0; wd < 2; ++wd)
            for (int by = 0; by < 4; ++by) {
                h ^= (words[wd] >> (8 * by)) & 0xffu;
                h *= 1099511628211ULL;
            }
        st->hash = h;
        st->pivots = k + 1;
        st->limit = k + 1;
        st->status = kRunning;
        st->live = 1;
    }
}
// In drive-out mode an artificial may stay basic in a redundant constraint: out[1] counts only the non-redundant ones.
template <typename real>
__global__ void __launch_bounds__(256) driveout_verdict_kernel(PivotParams<real> P, int* out)
{
    const int i = blockIdx.x;   // one block per constraint
    const int v = P.base[i];
    if (!(v >= P.n + P.m && v < P.n + 2 * P.m)) return;
    int any = 0;
    for (long long j = threadIdx.x; j < (long long)P.n + P.m; j += blockDim.x)
        any |= cmp3(fabs((double)P.T[(1 + j) * P.ld + i]), 0.0) > 0;
    any = __syncthreads_or(any);
    if (threadIdx.x == 0 && any) atomicAdd(out + 1, 1);
}

// ---------------------------------------------------------------------------------------------
// fp64 polish of an fp32 solve (no reference counterpart: the reference is fp64 only, include/macro.h:6).
// The final fp32 tableau holds an approximate inverse of the optimal basis in its slack rows
// (T[1+n+k][i] ~ (B^-1)[i][k] for the ORIGINAL system [A | I]); iterative refinement against the fp64 problem data
//     r = b - B x_B,   x_B += B^-1_fp32 r
// recovers x_B = B^-1 b (and with it the objective c_B . x_B) to fp64 accuracy whenever the fp32 inverse is good to
// better than 100 %.  orig = [ b (m) | A, variable-major (n x m) | c (n) ] in fp64.
// ---------------------------------------------------------------------------------------------
template <typename real>
__global__ void __launch_bounds__(256) polish_init_kernel(PivotParams<real> P, double* xB)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < P.m) xB[i] = (double)P.T[i];   // row 0 of the fp32 tableau: its own estimate of x_B
}
// r[k] = b[k] - sum_i B[k][i] * xB[i];   B[:, i] = A[v_i][:] for a structural basic variable v_i, e_(v_i - n) for a slack
template <typename real>
__global__ void __launch_bounds__(256) polish_residual_kernel(PivotParams<real> P, const double* __restrict__ orig,
                                                             const double* __restrict__ xB, double* r)
{
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= P.m) return;
    const double* A = orig + P.m;
    double acc = orig[k];
    for (int i = 0; i < P.m; ++i) {
        const int v = P.base[i];
        if (v < P.n)
            acc = __fma_rn(-A[(long long)v * P.m + k], xB[i], acc);
        else if (v - P.n == k)
            acc -= xB[i];
    }
    r[k] = acc;
}
// xB[i] += sum_k T[1+n+k][i] * r[k]; also reports max |dx| and max |x| (as ordered-int bit patterns via atomicMax)
template <typename real>
__global__ void __launch_bounds__(256) polish_correct_kernel(PivotParams<real> P, const double* __restrict__ r, double* xB,
                                                            unsigned long long* norms)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.m) return;
    const real* S0 = P.T + (1 + (long long)P.n) * P.ld + i;
    double acc = 0.0;
    for (int k = 0; k < P.m; ++k) acc = __fma_rn((double)S0[(long long)k * P.ld], r[k], acc);
    const double x = xB[i] + acc;
    xB[i] = x;
    atomicMax(norms + 0, (unsigned long long)__double_as_longlong(fabs(acc)));   // non-negative doubles order like integers
    atomicMax(norms + 1, (unsigned long long)__double_as_longlong(fabs(x)));
}
// x[v_i] = xB[i] for structural basics; objective = sum c[v_i] * xB[i] in ascending i (one block, deterministic)
template <typename real>
__global__ void __launch_bounds__(kSelBlock) polish_finish_kernel(PivotParams<real> P, const double* __restrict__ orig,
                                                                 const double* __restrict__ xB, double* x, double* obj)
{
    __shared__ double part[kSelBlock];
    const double* c = orig + P.m + (long long)P.n * P.m;
    double acc = 0.0;
    for (int i = threadIdx.x; i < P.m; i += kSelBlock) {
        const int v = P.base[i];
        if (v < P.n) {
            x[v] = xB[i];
            acc = __fma_rn(c[v], xB[i], acc);
        }
    }
    part[threadIdx.x] = acc;
    __syncthreads();
    for (int off = kSelBlock / 2; off > 0; off >>= 1) {
        if ((int)threadIdx.x < off) part[threadIdx.x] += part[threadIdx.x + off];
        __syncthreads();
    }
    if (threadIdx.x == 0) *obj = part[0];
}

}  // namespace b2s
