// b2s_persistent.cuh -- the whole pivot loop as ONE persistent cooperative kernel.
//
// pivot_loop_kernel runs up to `batch` simplex iterations per launch with one CTA per SM; the
// three per-pivot launches of b2s_kernels.cuh become phases separated by device-wide barriers, so
// the fixed cost of a pivot is three barrier round trips (~1 us each) instead of three kernel
// launches.  That is what the launch/latency-bound regime (tableaux that live in the 126 MB L2,
// BASELINE.json config "1024x2048") and the sharded solve (slabs get small as GPUs are added) need;
// in the HBM-bound regime it is worth the last percent.  Arithmetic and tournament trees are the
// same device functions as in the per-pivot kernels, so results are bit-identical.
//
//   phase A   ratio-test stage 1 per 512-constraint block        | sharded: winners written into every
//             -- barrier --                                       | rank's arena, flags over NVLink
//   phase B   every CTA replays stage 2 (<= 1024 slots) -> p      |
//   phase C   gather raw pivot constraint, normalise its column  | sharded: only the owner; it writes the
//             -- barrier --                                       | vector into every arena, then a flag
//   phase D   cost update + next entering tournament (first CTAs), rank-1 update over ticketed tiles
//             -- barrier --
//
// Vectors that change between barriers are read with ld.global.cg (L2) and tableau tiles with
// 256-bit ld.global.cg, so no SM can see a stale L1 line; the barrier itself is release/acquire at
// gpu scope.
#pragma once
#include "b2s_p2p.cuh"

namespace b2s {

__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// All CTAs of the (co-resident) grid meet here.  `epoch` counts the barriers this CTA has passed.
// Bounded like every other wait in this file: if a CTA left the loop early (peer timeout) the others
// give up after ~2 s instead of hanging the GPU; returns false in that case.
__device__ __forceinline__ bool grid_barrier(unsigned* counter, unsigned& epoch, int* s_ok)
{
    __syncthreads();
    epoch += 1;
    if (threadIdx.x == 0) {
        const unsigned target = epoch * gridDim.x;
        __threadfence();
        atomicAdd(counter, 1u);
        const long long t0 = clock64();
        int ok = 1;
        while (ld_acquire_gpu(counter) < target) {
            if (clock64() - t0 > 4000000000ll) {
                ok = 0;
                break;
            }
        }
        __threadfence();
        *s_ok = ok;
    }
    __syncthreads();
    return *s_ok != 0;
}

// Phase D's streaming loop is kept out of line so that it gets the same register allocation as the
// stand-alone update_kernel (8 x 256-bit loads in flight need 64 data registers of the 128 available).
template <typename real, int VB, int U, bool SKIP>
__device__ __noinline__ void stream_phase(const PivotParams<real>& P, const real* rowp, long long* s_next, bool reverse)
{
    stream_tiles<real, VB, U, 3, SKIP, true, true>(P, rowp, P.s, nullptr, (real)0, -1, s_next, reverse);
}

template <typename real, int VB, int U, bool SKIP>
__global__ void __launch_bounds__(kSelBlock, 1) pivot_loop_kernel(const __grid_constant__ PivotParams<real> P, int batch)
{
    __shared__ TreeSmem<real> sm;
    __shared__ real smax[32];
    __shared__ int s_flag;
    __shared__ long long s_next;
    __shared__ Cand<real> s_win;
    __shared__ real s_mx;
    __shared__ int s_ok;

    DevState* st = P.st;
    const bool sharded = P.world > 1;
    const int rule = P.rule;
    const int tree_rule = (rule == kRuleReference) ? kRuleReference : kRuleLowest;
    const long long gtid = (long long)blockIdx.x * kSelBlock + threadIdx.x;
    const long long gstride = (long long)gridDim.x * kSelBlock;
    unsigned epoch = 0;

    for (int it = 0; it < batch; ++it) {
        // state written before the previous barrier: identical in every CTA
        const int status = __ldcg(&st->status);
        const long long pivots = __ldcg(&st->pivots), limit = __ldcg(&st->limit);
        if (status != kRunning || pivots >= limit) break;
        const unsigned long long seq = (unsigned long long)pivots + 1ull;
        const int par = (int)(seq & 1ull);
        const int q = __ldcg(&st->q);
        const real cq = (real)__ldcg(&st->cq);
        const real* qrow = P.T + stored_row(P, 1 + (long long)q) * P.ld;
        const real* brow = P.T;

        // ---- phase A: entering column snapshot, block max, ratio stage 1 -------------------------
        for (int b = blockIdx.x; b < P.Gm_loc; b += gridDim.x) {
            const int gb = P.Gm_loc0 + b;
            Cand<real> c;
            c.v = Limits<real>::big();
            c.i = -1;
            c.k = -1;
            real mx = Limits<real>::tiny();
            for (long long gi = (long long)gb * kSelBlock + threadIdx.x; gi < P.m; gi += (long long)kSelBlock * P.Gm) {
                const long long li = gi - P.col0;
                const real a = __ldcg(qrow + li);
                const real bb = __ldcg(brow + li);
                P.col[li] = a;
                mx = fmax(mx, a);
                Cand<real> o;
                o.v = (cmp3((double)a, 0.0) > 0) ? div_r(bb, a) : Limits<real>::big();
                o.i = (int)gi;
                o.k = (rule == kRuleBland) ? ((o.v < Limits<real>::big()) ? __ldcg(P.base + gi) : -1) : (int)gi;
                if (beats(tree_rule, o, c)) c = o;
            }
            mx = block_max_512(mx, smax);
            block_tree_512(tree_rule, c, sm);
            if (threadIdx.x == 0) {
                s_win = c;
                s_mx = mx;
            }
            __syncthreads();
            if (!sharded) {
                if (threadIdx.x == 0) {
                    P.rslot_v[gb] = s_win.v;
                    P.rslot_max[gb] = s_mx;
                    P.rslot_i[gb] = s_win.i;
                    P.rslot_k[gb] = s_win.k;
                }
            } else if (threadIdx.x < P.world) {
                ArenaHeader<real>* a = arena_of(P, threadIdx.x);
                a->slot_v[par][gb] = s_win.v;
                a->slot_max[par][gb] = s_mx;
                a->slot_i[par][gb] = s_win.i;
                a->slot_k[par][gb] = s_win.k;
                __threadfence_system();
            }
            __syncthreads();
        }
        if (!grid_barrier(&st->bar_count, epoch, &s_ok)) {
            if (threadIdx.x == 0) {  // a CTA never arrived: make the failure visible to the host instead of looking like progress
                st->status = kStatusPeerTimeout;
                st->live = 0;
                __threadfence();
            }
            break;
        }
        const real *slot_v = P.rslot_v, *slot_max = P.rslot_max;
        const int *slot_i = P.rslot_i, *slot_k = P.rslot_k;
        if (sharded) {
            // this rank's winners are all out: tell the peers, then wait for theirs (local polling)
            if (blockIdx.x == 0 && threadIdx.x < P.world) {
                __threadfence_system();
                st_release_sys(&arena_of(P, threadIdx.x)->flag_slots[par][P.rank], seq);
            }
            int ok = 1;
            if (threadIdx.x < P.world) ok = wait_flag(&arena_of(P, P.rank)->flag_slots[par][threadIdx.x], seq, P.wait_cycles) ? 1 : 0;
            ok = __syncthreads_and(ok);
            if (!ok) {
                if (blockIdx.x == 0 && threadIdx.x == 0) {
                    st->status = kStatusPeerTimeout;
                    st->live = 0;
                }
                break;  // every CTA polls the same flags; a CTA that did see them stops at the next barrier's status check
            }
            __threadfence_system();
            const ArenaHeader<real>* mine = arena_of(P, P.rank);
            slot_v = mine->slot_v[par];
            slot_max = mine->slot_max[par];
            slot_i = mine->slot_i[par];
            slot_k = mine->slot_k[par];
        }

        // ---- phase B: stage 2, replayed by every CTA -------------------------------------------------
        real mx = Limits<real>::tiny();
        for (int b = threadIdx.x; b < P.Gm; b += kSelBlock) mx = fmax(mx, __ldcg(slot_max + b));
        mx = block_max_512(mx, smax);
        Cand<real> w;
        if (P.Gm > 1) {
            stage2_1024(tree_rule, slot_v, slot_i, slot_k, P.Gm, w, sm);
        } else {
            w.v = __ldcg(slot_v);
            w.i = __ldcg(slot_i);
            w.k = __ldcg(slot_k);
        }
        if (threadIdx.x == 0) {
            s_win = w;
            s_mx = mx;
        }
        __syncthreads();
        const int p = s_win.i;
        const bool unbounded = cmp3((double)s_mx, 0.0) <= 0 || p < 0;
        __syncthreads();
        if (unbounded) {
            if (blockIdx.x == 0 && threadIdx.x == 0) {
                st->status = kUnbounded;
                st->live = 0;
            }
            break;  // uniform: every CTA computed the same verdict
        }
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            st->p = p;
            P.base[p] = q;  // src/solver.cu:105
            if (pivots < P.trace_cap) P.trace[pivots] = make_int2(q, p);
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
            st->pivots = pivots + 1;
            st->live = 1;
            st->tile_ticket = 0;  // every CTA left phase D of the previous pivot before the last barrier
        }

        // ---- phase C: gather the raw pivot constraint, normalise the pivot column in place, build s ----
        const long long lp = (long long)p - P.col0;
        const bool owner = lp >= 0 && lp < P.m_loc;
        real piv = 0;
        if (owner) {
            piv = __ldcg(P.col + lp);
            long long nz = 0;
            for (long long r = gtid; r < P.Rs; r += gstride) {
                real* e = P.T + r * P.ld + lp;
                const real a = __ldcg(e);
                *e = div_r(a, piv);
                if (!sharded) {
                    P.rowp[r] = a;
                } else {
                    for (int wk = 0; wk < P.world; ++wk) arena_rowp(P, wk, par)[r] = a;
                }
                nz += (a != (real)0);
            }
            if (sharded) __threadfence_system();
            if (P.skip_zero && nz) atomicAdd((unsigned long long*)&st->rows_streamed, (unsigned long long)nz);
        }
        if (!sharded) {  // one GPU: the pivot is local, s can be built before the barrier
            for (long long i = gtid; i < P.ld; i += gstride)
                P.s[i] = (i < P.m_loc && i != lp) ? div_r(-__ldcg(P.col + i), piv) : (real)0;
        }
        if (!grid_barrier(&st->bar_count, epoch, &s_ok)) {
            if (threadIdx.x == 0) {  // a CTA never arrived: make the failure visible to the host instead of looking like progress
                st->status = kStatusPeerTimeout;
                st->live = 0;
                __threadfence();
            }
            break;
        }
        const real* rowp = P.rowp;
        if (sharded) {
            if (owner && blockIdx.x == 0 && threadIdx.x < P.world) {
                __threadfence_system();
                st_release_sys(&arena_of(P, threadIdx.x)->flag_rowp[par], seq);
            }
            if (threadIdx.x == 0) s_ok = wait_flag(&arena_of(P, P.rank)->flag_rowp[par], seq, P.wait_cycles) ? 1 : 0;
            __syncthreads();
            if (!s_ok) {
                if (blockIdx.x == 0 && threadIdx.x == 0) {
                    st->status = kStatusPeerTimeout;
                    st->live = 0;
                }
                break;
            }
            __threadfence_system();
            rowp = arena_rowp(P, P.rank, par);
            piv = __ldcg(rowp + stored_row(P, 1 + (long long)q));  // a_pq = T[1+q][p]
            for (long long i = gtid; i < P.ld; i += gstride)
                P.s[i] = (i < P.m_loc && i != lp) ? div_r(-__ldcg(P.col + i), piv) : (real)0;
            if (!grid_barrier(&st->bar_count, epoch, &s_ok)) {
            if (threadIdx.x == 0) {  // a CTA never arrived: make the failure visible to the host instead of looking like progress
                st->status = kStatusPeerTimeout;
                st->live = 0;
                __threadfence();
            }
            break;
        }
        }
        const real sc = div_r(-cq, piv);  // src/solver.cu:54
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            st->piv = (double)piv;
            st->sc = (double)sc;
        }

        // ---- phase D: cost update + next entering tournament, then the rank-1 update ----------------
        if (blockIdx.x < P.Gc) cost_select_blocks<real, true, true>(P, rowp, sc, sm, &s_flag);
        stream_phase<real, VB, U, SKIP>(P, rowp, &s_next, P.serpentine && (seq & 1ull));
        if (!grid_barrier(&st->bar_count, epoch, &s_ok)) {
            if (threadIdx.x == 0) {  // a CTA never arrived: make the failure visible to the host instead of looking like progress
                st->status = kStatusPeerTimeout;
                st->live = 0;
                __threadfence();
            }
            break;
        }
    }
}

}  // namespace b2s
