// b2s_lookahead.cuh -- ONE kernel per simplex pivot: the streaming rank-1 update of pivot k and, hidden
// under it, the complete selection of pivot k+1 ("look-ahead selection").
//
// The reference runs the dependent chain select -> copy column -> max test -> ratio test -> gather ->
// update strictly in sequence (src/solver.cu:86-105).  Only the update touches O(R*m) data; every other
// step needs O(R+m) values that can be formed from the OLD tableau plus the rank-1 term of the running
// update -- by the same FMA the streaming loop applies, hence the same bits:
//
//   cost'      = fma(sc, rowp, cost)                      -> q'   (needs nothing from the tableau)
//   b'_i       = fma(s_i, rowp[0],    T[0][i])            \  ratio test of pivot k+1 -> p'
//   col'_i     = fma(s_i, rowp[1+q'], T[1+q'][i])         /
//   rowp'[r]   = fma(s_p', rowp[r],   T[r][p'])              raw pivot constraint of pivot k+1
//   s'_i       = (-col'_i) / col'_p'
//
// update_la_kernel: CTAs 0..H-1 ("helpers") run that chain while all other CTAs stream tiles; then the
// helpers stream too.  When the tableau is sharded the two exchanges of the chain (stage-1 winners to
// everybody, pivot constraint from its owner to everybody) go over NVLink peer memory from inside the same
// kernel, i.e. they also overlap the streaming.  The next launch finds a complete "proposal" (q, p, pivot,
// s, col, rowp) for its pivot in device memory and starts streaming at once: no ratio / gather launches, no
// exposed exchange.
//
// Hazard protocol (the streaming CTAs overwrite the tableau in place while the helpers read it):
//   * row 0 (RHS) is never touched by the streaming CTAs; the helpers update it.
//   * Tiles are handed out by ONE 64-bit atomic word: [ticket count | published row | published column].
//     A helper publishes the entering variable's row (later: the next pivot column) by adding it into that
//     word; the value the atomic returns is the number of tiles claimed before the publication.  A CTA that
//     claims a tile learns, from the same atomic that gives it the tile, which publications preceded its
//     claim -- single-location coherence order, no fences.  It then leaves the published row / column
//     untouched ("held old"); the helpers compute those elements themselves from the old values.
//   * Tiles claimed BEFORE a publication are updated in full; the helpers wait for the tile's completion
//     record and read the new values.
//   * The next pivot column may thus stay stale in the tableau.  It is never read: during the next pivot
//     its raw values ARE the gathered vector rowp', and the streaming loop overwrites the column with
//     rowp'[r] / pivot (src/solver.cu:43, `col == colPivotIndex`).  la_flush_kernel writes rowp' back when
//     the host wants to look at the tableau.
//
// Row list.  The chain also compacts the rows the next update has to stream into a list (row index + pivot-
// constraint entry a_pr): every stored row except row 0, or -- skip_zero_rows -- only those with a_pr != 0
// (fma(s_i, 0, x) == x, src/solver.cu:43).  Tiles are cut from that list, so every thread keeps 8 independent
// 256-bit loads in flight whatever the sparsity of the pivot constraint.
//
// Arithmetic, tournament trees and tie order are the device functions of b2s_device.cuh, so results are
// bit-identical to the three-launch path and to the oracle.
#pragma once
#include "b2s_p2p.cuh"

namespace b2s {

constexpr int kLaMaxHelpers = 16;
constexpr int kStatusInternal = -97;  // the loop found no proposal for its pivot (cannot happen; surfaced as an error)

// Ticket word layout.
constexpr int kTicketBits = 21;                                   // tiles claimed (<= 2^20 tiles + one overshoot per CTA)
constexpr int kRowBits = 19;                                      // published stored row + 1 (<= 196609 rows at 65536^2)
constexpr unsigned long long kTicketMask = (1ull << kTicketBits) - 1ull;
constexpr unsigned long long kRowMask = (1ull << kRowBits) - 1ull;
constexpr int kColShift = kTicketBits + kRowBits;                 // published local column + 1 (<= 2^23)
constexpr unsigned kNoColumn = 0x7fffffu;                         // "no column will be published during this pivot"

// The complete selection of one pivot.  Two generations, indexed by the parity of the pivot number.
struct Proposal {
    int q;            // entering variable (0-based, device row 1+q)
    int p;            // leaving constraint (global column index)
    double cq;        // tournament value of the entering reduced cost
    double piv;       // a_pq
    double sc;        // (-cq)/piv
    int status_next;  // kRunning, or how the phase ends INSTEAD of this pivot: kFeasible (no entering column) / kUnbounded
    unsigned q_seq;   // == pivot number once q, cq, status_next are valid
    unsigned p_seq;   // == pivot number once p is valid (or status_next == kUnbounded)
    unsigned rowp_seq;   // one GPU: == pivot number once rowp' is complete (sharded: the arena's flag_rowp)
    unsigned ready_seq;  // == pivot number once everything is in place: the pivot may execute
    unsigned row_pub_seq, col_pub_seq;  // publication hand-shake between the helpers
    unsigned c_row, c_col;              // tiles claimed before the row / column publication
    long long nz;     // length of the row list: rows (other than row 0) the update of this pivot streams
};

struct LaState {
    Proposal prop[2];
    unsigned long long word;   // ticket word of the running update
    unsigned tile_done;        // CTAs that have left the streaming loop
    unsigned ticket_cost, ticket_ratio, ticket_gather, ticket_s;
    unsigned pad0;
    unsigned cnt_seq[kLaMaxHelpers];  // row-list compaction: helper h has published cnt[h] for pivot cnt_seq[h]
    int cnt[kLaMaxHelpers];
    unsigned long long stamps[8];  // %globaltimer at the chain's milestones of the last pivot (profiling)
};

// ---- small helpers ---------------------------------------------------------------------------------
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u32(unsigned* p, unsigned v)
{
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// Block-uniform bounded wait until *flag == want (thread 0 polls).  Returns false on timeout.
__device__ __forceinline__ bool la_wait_u32(const unsigned* flag, unsigned want, long long cycles, int* s_ok)
{
    if (threadIdx.x == 0) {
        const long long t0 = clock64();
        int ok = 1;
        while (ld_acquire_u32(flag) != want) {
            if (clock64() - t0 > cycles) {
                ok = 0;
                break;
            }
        }
        *s_ok = ok;
    }
    __syncthreads();
    const bool ok = *s_ok != 0;
    __syncthreads();
    return ok;
}
__device__ __forceinline__ bool wait_flag_cycles(const unsigned long long* flag, unsigned long long seq, long long cycles)
{
    const long long t0 = clock64();
    while (ld_acquire_sys(flag) < seq) {
        if (clock64() - t0 > cycles) return false;
        __nanosleep(32);
    }
    return true;
}

template <typename real>
__device__ __forceinline__ real* la_rowp(const PivotParams<real>& P, int par)
{
    return P.world > 1 ? arena_rowp(P, P.rank, par) : P.rowp2 + (size_t)par * P.rowp_stride;
}

// Sweep position of tile (row block rb, column chunk) -- the inverse of the streaming loop's mapping.
template <typename real>
__device__ __forceinline__ long long la_ticket_of(const PivotParams<real>& P, long long rb, int chunk, bool reverse)
{
    const long long tmap = rb * P.nchunks + chunk;
    return reverse ? (P.ntiles - 1 - tmap) : tmap;
}

// Double-buffered staging of a tile's list entries (tile_rows <= 512 >> 0 ... bounded by kLaMaxTileRows) and ticket words.
constexpr int kLaMaxTileRows = 512;
template <typename real>
struct LaTileSmem {
    unsigned long long word[2];
    int row[2][kLaMaxTileRows];
    real val[2][kLaMaxTileRows];
};

struct LaShared {
    unsigned long long next_word;
    int ok;
    int flag;
    int i0, i1;
    double d0, d1;
};

// ---------------------------------------------------------------------------------------------
// Stage "cost": cost update of the running pivot (src/solver.cu:48-56) + entering tournament of the next
// (src/reduction.cu:51-104 over costsVector+1).  Block b plays reference stage-1 block b; the CTA that
// draws the last ticket plays stage 2 and publishes q' (or "optimal") into the next proposal.
// ---------------------------------------------------------------------------------------------
template <typename real>
__device__ __noinline__ void la_cost_blocks(const PivotParams<real>& P, LaState* la, Proposal* nxt, unsigned tseq,
                                            const real* rowp, real sc, TreeSmem<real>& sm, LaShared& sh)
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
            real v = P.cost[j];
            v = fma_r(sc, __ldg(rowp + stored_row(P, j)), v);  // src/solver.cu:54
            P.cost[j] = v;
            Cand<real> o;
            o.v = v;
            o.i = (int)i;
            o.k = (rule == kRuleBland) ? (cmp3((double)v, 0.0) < 0 ? (int)i : -1) : (int)i;
            if (beats(rule, o, c)) c = o;
        }
        if (b == 0 && threadIdx.x == 0) P.cost[0] = fma_r(sc, __ldg(rowp), P.cost[0]);  // objective value
        block_tree_512(rule, c, sm);
        if (threadIdx.x == 0) {
            P.cslot_v[b] = c.v;
            P.cslot_i[b] = c.i;
            P.cslot_k[b] = c.k;
            __threadfence();
            const unsigned t = atomicAdd(&la->ticket_cost, 1u);
            sh.flag = (t == (unsigned)P.Gc - 1u);
        }
        __syncthreads();
        const bool last = sh.flag != 0;
        __syncthreads();
        if (last) {
            __threadfence();
            Cand<real> w;
            if (P.Gc > 1)
                stage2_1024(rule, P.cslot_v, P.cslot_i, P.cslot_k, P.Gc, w, sm);
            else
                w = c;
            if (threadIdx.x == 0) {
                nxt->q = w.i;
                nxt->cq = (double)w.v;
                nxt->status_next = (w.i >= 0 && cmp3((double)w.v, 0.0) < 0) ? kRunning : kFeasible;  // src/solver.cu:87-88
                la->ticket_cost = 0;
                __threadfence();
                st_release_u32(&nxt->q_seq, tseq);
            }
        }
    }
}

// Stage 2 of the ratio test over all Gm stage-1 slots + the unbounded test (src/reduction.cu:186-201).
// Result valid in thread 0: w (winner), mx (max of the entering column).
template <typename real>
__device__ __forceinline__ void la_ratio_stage2(const PivotParams<real>& P, const real* slot_v, const int* slot_i,
                                                const int* slot_k, const real* slot_max, TreeSmem<real>& sm, real* smax,
                                                Cand<real>& w, real& mx)
{
    mx = Limits<real>::tiny();
    for (int b = threadIdx.x; b < P.Gm; b += kSelBlock) mx = fmax(mx, __ldcg(slot_max + b));
    mx = block_max_512(mx, smax);
    const int tree_rule = (P.rule == kRuleReference) ? kRuleReference : kRuleLowest;
    if (P.Gm > 1) {
        stage2_1024(tree_rule, slot_v, slot_i, slot_k, P.Gm, w, sm);
    } else {
        w.v = __ldcg(slot_v);
        w.i = __ldcg(slot_i);
        w.k = __ldcg(slot_k);
    }
}

// ---------------------------------------------------------------------------------------------
// The helpers' chain.  LIVE = true: called from inside update_la_kernel for pivot `seq` (current proposal:
// q, p, piv; vectors svec / rowp and the row list of the running update) and builds the proposal of pivot
// seq+1 while the other CTAs stream.  LIVE = false: the tableau is quiescent (prologue kernel: first pivot of
// a phase or of an iterate() call); q' comes from the select kernel (st->q / st->cq), every element is final.
// ---------------------------------------------------------------------------------------------
template <typename real, bool LIVE>
__device__ __noinline__ void la_chain(const PivotParams<real>& P, LaState* la, unsigned seq, int h, int H,
                                      const real* rowp, const real* svec, real piv, long long lp, int p_cur, int q_cur,
                                      bool reverse, long long ntiles, TreeSmem<real>& sm, real* smax, LaShared& sh)
{
    DevState* st = P.st;
    const unsigned tseq = seq + 1u;     // the pivot this chain prepares
    const int par = (int)(seq & 1u), tpar = par ^ 1;
    Proposal* nxt = &la->prop[tpar];
    const bool sharded = P.world > 1;
    const int rule = P.rule;
    const int tree_rule = (rule == kRuleReference) ? kRuleReference : kRuleLowest;
    const long long cyc = P.wait_cycles;
    const long long base = (long long)gridDim.x - H;   // tiles claimed implicitly at launch (LIVE only)
    real* colN = P.col2 + (size_t)tpar * P.ld;
    real* sN = P.s2 + (size_t)tpar * P.ld;
    real* rowpN = la_rowp(P, tpar);
    const int* posC = P.rowpos + (size_t)par * P.rowp_stride;   // row -> position in the running update's row list
    const int rpp = kSelBlock >> P.log2_tpr;
    const long long tile_rows = (long long)rpp * 8;
    const long long chunk_cols = (long long)(32 / (int)sizeof(real)) << P.log2_tpr;
    const real a0 = LIVE ? __ldg(rowp) : (real)0;   // rowp[0] = b_p of the running pivot
    const bool own_cur = LIVE && lp >= 0 && lp < P.m_loc;

    if (LIVE) {
        // ---- stage 0: row 0 (RHS) of the running update -- it is not in the row list ---------------------------
        for (int bl = h; bl < P.Gm_loc; bl += H) {
            const long long li = (long long)bl * kSelBlock + threadIdx.x;
            if (li < P.m_loc) {
                const real x = __ldcg(P.T + li);
                P.T[li] = (li == lp) ? div_r(a0, piv) : fma_r(__ldg(svec + li), a0, x);
            }
        }
        // rows the list leaves out (a_pr == 0) are unchanged by this update, pivot column included: a_pr / pivot = a_pr.
        // The tableau may still hold a held-old value there (see the header): write the true entry.
        if (own_cur && P.skip_zero) {
            for (long long r = (long long)h * kSelBlock + threadIdx.x; r < P.Rs; r += (long long)H * kSelBlock)
                if (r != 0 && __ldg(posC + r) < 0) P.T[r * P.ld + lp] = __ldg(rowp + r);
        }
        if (h == 0 && threadIdx.x == 0) la->stamps[1] = globaltimer();
        // ---- stage Q: the entering variable of the next pivot ------------------------------------------------
        if (!la_wait_u32(&nxt->q_seq, tseq, cyc, &sh.ok)) {
            if (h == 0 && threadIdx.x == 0) {
                nxt->status_next = kStatusPeerTimeout;
                atomicAdd(&la->word, (unsigned long long)kNoColumn << kColShift);
            }
            return;
        }
        __threadfence();
        if (__ldcg(&nxt->status_next) != kRunning) {   // optimal after this pivot: nothing to prepare
            if (h == 0 && threadIdx.x == 0) atomicAdd(&la->word, (unsigned long long)kNoColumn << kColShift);
            return;
        }
    } else if (h == 0 && threadIdx.x == 0) {
        nxt->q = __ldcg(&st->q);
        nxt->cq = __ldcg(&st->cq);
        nxt->status_next = kRunning;
    }
    const int qn = LIVE ? __ldcg(&nxt->q) : __ldcg(&st->q);
    const real cqn = (real)(LIVE ? __ldcg(&nxt->cq) : __ldcg(&st->cq));
    const long long rq = stored_row(P, 1 + (long long)qn);
    if (h == 0 && threadIdx.x == 0) la->stamps[2] = globaltimer();

    // ---- stage R: entering column + RHS after the running update, ratio-test stage 1 ----------------------
    long long c_row = 0;
    long long rbq = -1;   // row block (of the running update's list) that holds row rq; -1: no tile touches it
    if (LIVE) {
        if (h == 0 && threadIdx.x == 0) {
            const unsigned long long old = atomicAdd(&la->word, (unsigned long long)(rq + 1) << kTicketBits);
            nxt->c_row = (unsigned)(old & kTicketMask);
            __threadfence();
            st_release_u32(&nxt->row_pub_seq, tseq);
        }
        if (!la_wait_u32(&nxt->row_pub_seq, tseq, cyc, &sh.ok)) return;
        c_row = (long long)__ldcg(&nxt->c_row);
        const int posq = __ldg(posC + rq);
        if (posq >= 0) rbq = posq / tile_rows;
    }
    const real aq = LIVE ? __ldg(rowp + rq) : (real)0;   // rowp[1+q'] of the running pivot
    for (int bl = h; bl < P.Gm_loc; bl += H) {
        const int gb = P.Gm_loc0 + bl;
        const long long li = (long long)bl * kSelBlock + threadIdx.x;
        bool old_vals = false;
        if (LIVE) {
            old_vals = true;
            if (rbq >= 0) {
                const int chunk = (int)(((long long)bl * kSelBlock) / chunk_cols);
                const long long tmap = rbq * P.nchunks + chunk;
                const long long t = reverse ? (ntiles - 1 - tmap) : tmap;
                old_vals = t >= c_row + base;  // claimed after the publication: the claimer leaves row rq alone
                if (!old_vals) {               // claimed before: wait until that tile is complete, then read the new values
                    if (!la_wait_u32(P.tile_rec + tmap, seq, cyc, &sh.ok)) {
                        if (threadIdx.x == 0) nxt->status_next = kStatusPeerTimeout;
                        return;
                    }
                }
            }
        }
        Cand<real> c;
        c.v = Limits<real>::big();
        c.i = -1;
        c.k = -1;
        real mx = Limits<real>::tiny();
        if (li < P.m_loc) {
            real* e = P.T + rq * P.ld + li;
            real a = __ldcg(e);
            if (old_vals) {
                a = (li == lp) ? div_r(aq, piv) : fma_r(__ldg(svec + li), aq, a);
                *e = a;
            }
            const real bb = __ldcg(P.T + li);
            colN[li] = a;
            mx = fmax(mx, a);   // src/reduction.cu:143-184, identity DBL_MIN
            const long long gi = P.col0 + li;
            Cand<real> o;
            o.v = (cmp3((double)a, 0.0) > 0) ? div_r(bb, a) : Limits<real>::big();   // src/reduction.cu:106-114
            o.i = (int)gi;
            if (rule == kRuleBland) {
                const int bv = (LIVE && gi == p_cur) ? q_cur : P.base[gi];   // base[p] = q of the running pivot is committed at its end
                o.k = (o.v < Limits<real>::big()) ? bv : -1;
            } else {
                o.k = (int)gi;
            }
            if (beats(tree_rule, o, c)) c = o;
        }
        mx = block_max_512(mx, smax);
        block_tree_512(tree_rule, c, sm);
        if (!sharded) {
            if (threadIdx.x == 0) {
                P.rslot_v[gb] = c.v;
                P.rslot_max[gb] = mx;
                P.rslot_i[gb] = c.i;
                P.rslot_k[gb] = c.k;
            }
        } else {
            if (threadIdx.x == 0) {
                sh.d0 = (double)c.v;
                sh.d1 = (double)mx;
                sh.i0 = c.i;
                sh.i1 = c.k;
            }
            __syncthreads();
            if (threadIdx.x < P.world) {
                ArenaHeader<real>* a = arena_of(P, threadIdx.x);
                a->slot_v[tpar][gb] = (real)sh.d0;
                a->slot_max[tpar][gb] = (real)sh.d1;
                a->slot_i[tpar][gb] = sh.i0;
                a->slot_k[tpar][gb] = sh.i1;
                __threadfence_system();
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        if (sharded) __threadfence_system(); else __threadfence();
        const unsigned t = atomicAdd(&la->ticket_ratio, 1u);
        sh.flag = (t == (unsigned)H - 1u);
    }
    __syncthreads();
    const bool last_ratio = sh.flag != 0;
    __syncthreads();
    if (last_ratio) {
        __threadfence();
        const real *slot_v = P.rslot_v, *slot_max = P.rslot_max;
        const int *slot_i = P.rslot_i, *slot_k = P.rslot_k;
        bool ok = true;
        if (sharded) {
            // all local winners are out: raise our flag everywhere, then wait for everybody's (local polling)
            __threadfence_system();
            const bool mute = P.fault_rank == P.rank && (long long)tseq >= P.fault_pivot;   // fault injection (tests)
            if (threadIdx.x < P.world && !mute)
                st_release_sys(&arena_of(P, threadIdx.x)->flag_slots[tpar][P.rank], (unsigned long long)tseq);
            int okk = 1;
            if (threadIdx.x < P.world)
                okk = wait_flag_cycles(&arena_of(P, P.rank)->flag_slots[tpar][threadIdx.x], (unsigned long long)tseq, cyc) ? 1 : 0;
            ok = __syncthreads_and(okk) != 0;
            __threadfence_system();
            const ArenaHeader<real>* mine = arena_of(P, P.rank);
            slot_v = mine->slot_v[tpar];
            slot_max = mine->slot_max[tpar];
            slot_i = mine->slot_i[tpar];
            slot_k = mine->slot_k[tpar];
        }
        Cand<real> w;
        w.v = Limits<real>::big();
        w.i = -1;
        w.k = -1;
        real mx = Limits<real>::tiny();
        if (ok) la_ratio_stage2(P, slot_v, slot_i, slot_k, slot_max, sm, smax, w, mx);
        if (threadIdx.x == 0) {
            la->ticket_ratio = 0;
            if (!ok) {
                nxt->status_next = kStatusPeerTimeout;
                nxt->p = -1;
            } else if (cmp3((double)mx, 0.0) <= 0 || w.i < 0) {
                nxt->status_next = kUnbounded;   // src/solver.cu:98-99
                nxt->p = -1;
            } else {
                nxt->p = w.i;
            }
            __threadfence();
            st_release_u32(&nxt->p_seq, tseq);
        }
    }
    // ---- stage G: the owner of constraint p' gathers the raw pivot constraint after the running update ----
    if (!la_wait_u32(&nxt->p_seq, tseq, cyc, &sh.ok)) return;
    __threadfence();
    if (h == 0 && threadIdx.x == 0) la->stamps[3] = globaltimer();
    const int pn = __ldcg(&nxt->p);
    if (pn < 0) {   // unbounded (or a peer went silent): the phase ends instead of pivot tseq
        if (LIVE && h == 0 && threadIdx.x == 0) atomicAdd(&la->word, (unsigned long long)kNoColumn << kColShift);
        return;
    }
    const long long lpn = (long long)pn - P.col0;
    const bool owner = lpn >= 0 && lpn < P.m_loc;
    const bool same_col = LIVE && owner && lpn == lp;   // the same constraint leaves twice in a row
    long long c_col = 0;
    if (LIVE) {
        if (h == 0 && threadIdx.x == 0) {
            const unsigned long long field = (owner && !same_col) ? (unsigned long long)(lpn + 1) : (unsigned long long)kNoColumn;
            const unsigned long long old = atomicAdd(&la->word, field << kColShift);
            nxt->c_col = (unsigned)(old & kTicketMask);
            __threadfence();
            st_release_u32(&nxt->col_pub_seq, tseq);
        }
        if (owner && !same_col) {
            if (!la_wait_u32(&nxt->col_pub_seq, tseq, cyc, &sh.ok)) return;
            c_col = (long long)__ldcg(&nxt->c_col);
        }
    }
    if (owner) {
        const int chunk = (int)(lpn / chunk_cols);
        const real sp = (LIVE && !same_col) ? __ldg(svec + lpn) : (real)0;
        int bad = 0;
        for (long long r = (long long)h * kSelBlock + threadIdx.x; r < P.Rs; r += (long long)H * kSelBlock) {
            real v;
            if (same_col) {
                v = div_r(__ldg(rowp + r), piv);   // T'[r][p] = a_pr / pivot (src/solver.cu:43)
            } else {
                const real* e = P.T + r * P.ld + lpn;
                const int pos = (LIVE && r != 0 && r != rq) ? __ldg(posC + r) : -2;
                if (pos >= 0) {
                    const long long tmap = (pos / tile_rows) * P.nchunks + chunk;
                    const long long t = reverse ? (ntiles - 1 - tmap) : tmap;
                    if (t >= c_col + base) {
                        v = fma_r(sp, __ldg(rowp + r), __ldcg(e));   // held old by its claimer: apply the running update here
                    } else {
                        const unsigned* rec = P.tile_rec + tmap;
                        const long long t0 = clock64();
                        while (ld_acquire_u32(rec) != seq) {
                            if (clock64() - t0 > cyc) {
                                bad = 1;
                                break;
                            }
                        }
                        v = __ldcg(e);   // updated in full by a tile claimed before the publication
                    }
                } else if (pos == -1) {
                    v = fma_r(sp, __ldg(rowp + r), __ldcg(e));       // not in the list (a_pr == 0): no tile touches it
                } else {
                    v = __ldcg(e);       // rows 0 and 1+q' were finished by stages 0 / R; quiescent tableau: final
                }
            }
            if (!sharded) {
                rowpN[r] = v;
            } else {
                for (int w = 0; w < P.world; ++w) arena_rowp(P, w, tpar)[r] = v;
            }
        }
        bad = __syncthreads_or(bad);
        if (threadIdx.x == 0) {
            if (bad) nxt->status_next = kStatusPeerTimeout;
            if (sharded) __threadfence_system(); else __threadfence();
            const unsigned t = atomicAdd(&la->ticket_gather, 1u);
            sh.flag = (t == (unsigned)H - 1u);
        }
        __syncthreads();
        const bool last_g = sh.flag != 0;
        __syncthreads();
        if (last_g) {
            if (sharded) {
                __threadfence_system();
                if (threadIdx.x < P.world) st_release_sys(&arena_of(P, threadIdx.x)->flag_rowp[tpar], (unsigned long long)tseq);
            }
            if (threadIdx.x == 0) {
                la->ticket_gather = 0;
                __threadfence();
                st_release_u32(&nxt->rowp_seq, tseq);
            }
        }
    }
    // ---- stage S: s' = -col'/pivot' for the local slab, sc', and the row list of the next update -----------
    if (sharded) {
        if (threadIdx.x == 0) sh.ok = wait_flag_cycles(&arena_of(P, P.rank)->flag_rowp[tpar], (unsigned long long)tseq, cyc) ? 1 : 0;
        __syncthreads();
        const bool ok = sh.ok != 0;
        __syncthreads();
        if (!ok) {
            if (h == 0 && threadIdx.x == 0) nxt->status_next = kStatusPeerTimeout;
            return;
        }
        __threadfence_system();
    } else {
        if (!la_wait_u32(&nxt->rowp_seq, tseq, cyc, &sh.ok)) return;
        __threadfence();
    }
    if (h == 0 && threadIdx.x == 0) la->stamps[4] = globaltimer();
    const real pivn = __ldcg(rowpN + rq);   // a_pq = T[1+q'][p']
    for (long long i = (long long)h * kSelBlock + threadIdx.x; i < P.ld; i += (long long)H * kSelBlock)
        sN[i] = (i < P.m_loc && i != lpn) ? div_r(-__ldcg(colN + i), pivn) : (real)0;
    {
        // Row list: helper h compacts the contiguous slice [r_lo, r_hi) of stored rows.  Pass 1 counts, the counts are
        // exchanged through LaState, pass 2 writes (row, a_pr) pairs in ascending row order.
        int* listN = P.rowlist + (size_t)tpar * P.rowp_stride;
        real* valN = P.rowval + (size_t)tpar * P.rowp_stride;
        int* posN = P.rowpos + (size_t)tpar * P.rowp_stride;
        const long long slice = (((P.Rs + H - 1) / H) + kSelBlock - 1) / kSelBlock * kSelBlock;
        const long long r_lo = (long long)h * slice, r_hi = (r_lo + slice < P.Rs) ? r_lo + slice : P.Rs;
        const bool skip = P.skip_zero != 0;
        int mine = 0;
        for (long long c0 = r_lo; c0 < r_hi; c0 += kSelBlock) {
            const long long r = c0 + threadIdx.x;
            const int live = (r < r_hi && r != 0 && (!skip || __ldcg(rowpN + r) != (real)0)) ? 1 : 0;
            mine += __syncthreads_count(live);
        }
        if (threadIdx.x == 0) {
            la->cnt[h] = mine;
            __threadfence();
            st_release_u32(&la->cnt_seq[h], tseq);
        }
        int okk = 1;
        if (threadIdx.x < H) {
            const long long t0 = clock64();
            while (ld_acquire_u32(&la->cnt_seq[threadIdx.x]) != tseq) {
                if (clock64() - t0 > cyc) {
                    okk = 0;
                    break;
                }
            }
        }
        if (!__syncthreads_and(okk)) {
            if (threadIdx.x == 0) nxt->status_next = kStatusPeerTimeout;
            return;
        }
        __threadfence();
        int offset = 0, total = 0;
        for (int k = 0; k < H; ++k) {
            const int ck = __ldcg(&la->cnt[k]);
            if (k < h) offset += ck;
            total += ck;
        }
        int* s_w = reinterpret_cast<int*>(&sm);   // 16 warp totals (the tree buffers are idle here)
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
        for (long long c0 = r_lo; c0 < r_hi; c0 += kSelBlock) {
            const long long r = c0 + threadIdx.x;
            real v = (real)0;
            if (r < r_hi) v = __ldcg(rowpN + r);
            const int live = (r < r_hi && r != 0 && (!skip || v != (real)0)) ? 1 : 0;
            const unsigned bal = __ballot_sync(0xffffffffu, live);
            if (lane == 0) s_w[wid] = __popc(bal);
            __syncthreads();
            int before = 0, chunk_total = 0;
#pragma unroll
            for (int k = 0; k < kSelBlock / 32; ++k) {
                const int wk = s_w[k];
                if (k < wid) before += wk;
                chunk_total += wk;
            }
            const int pos = offset + before + __popc(bal & ((1u << lane) - 1u));
            if (r < r_hi) {
                if (live) {
                    listN[pos] = (int)r;
                    valN[pos] = v;
                    posN[r] = pos;
                } else {
                    posN[r] = -1;
                }
            }
            offset += chunk_total;
            __syncthreads();
        }
        if (h == 0 && threadIdx.x == 0) {
            nxt->nz = total;
            nxt->piv = (double)pivn;
            nxt->sc = (double)div_r(-cqn, pivn);   // src/solver.cu:54
        }
    }
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned t = atomicAdd(&la->ticket_s, 1u);
        if (t == (unsigned)H - 1u) {
            la->ticket_s = 0;
            __threadfence();
            st_release_u32(&nxt->ready_seq, tseq);
            la->stamps[5] = globaltimer();
        }
    }
}

// ---------------------------------------------------------------------------------------------
// update_la_kernel -- see the header of this file.  256-bit accesses, 8 rows in flight per thread, tiles of
// (512 >> log2_tpr) * 8 list rows x one column chunk, handed out by the ticket word.
// ---------------------------------------------------------------------------------------------
template <typename real>
__global__ void __launch_bounds__(kSelBlock, 1) update_la_kernel(const __grid_constant__ PivotParams<real> P)
{
    constexpr int VB = 32, U = 8;
    constexpr int EPT = VB / (int)sizeof(real);
    __shared__ TreeSmem<real> sm;
    __shared__ real smax[32];
    __shared__ LaShared sh;
    __shared__ LaTileSmem<real> ts;

    DevState* st = P.st;
    LaState* la = P.la;
    const int status = __ldcg(&st->status);
    const long long pivots = __ldcg(&st->pivots), limit = __ldcg(&st->limit);
    if (status != kRunning || pivots >= limit) return;
    const unsigned seq = (unsigned)(pivots + 1);
    const int par = (int)(seq & 1u);
    const Proposal* cur = &la->prop[par];
    if (__ldcg(&cur->ready_seq) != seq) {
        if (blockIdx.x == 0 && threadIdx.x == 0) st->status = kStatusInternal;
        return;
    }
    const int H = P.helpers;
    const bool helper = (int)blockIdx.x < H;
    const int p = __ldcg(&cur->p), q = __ldcg(&cur->q);
    const int lp = p - P.col0;   // outside [0, m_loc) when another rank owns the pivot column
    const real* rowp = la_rowp(P, par);
    const real* svec = P.s2 + (size_t)par * P.ld;
    const bool reverse = P.serpentine && (seq & 1u);
    const int rpp = kSelBlock >> P.log2_tpr;
    const int tile_rows = rpp * U;
    const int nlive = (int)__ldcg(&cur->nz);
    const int ntiles = ((nlive + tile_rows - 1) / tile_rows) * P.nchunks;
    if (blockIdx.x == 0 && threadIdx.x == 0) la->stamps[0] = globaltimer();

    if ((int)blockIdx.x < P.Gc)
        la_cost_blocks<real>(P, la, &la->prop[par ^ 1], seq + 1u, rowp, (real)__ldcg(&cur->sc), sm, sh);
    if (helper)
        la_chain<real, true>(P, la, seq, (int)blockIdx.x, H, rowp, svec, (real)__ldcg(&cur->piv), (long long)lp, p, q, reverse,
                             (long long)ntiles, sm, smax, sh);

    // ---- streaming ---------------------------------------------------------------------------------------
    // One barrier per tile.  While the tile's 256-bit loads are in flight, warp 0 receives the next ticket word and
    // stages the next tile's list entries (row index, a_pr) in shared memory, so no thread ever waits for a dependent
    // global load before it can issue its tile loads.
    {
        const int* rlist = P.rowlist + (size_t)par * P.rowp_stride;
        const real* rval = P.rowval + (size_t)par * P.rowp_stride;
        const int tx = threadIdx.x & ((1 << P.log2_tpr) - 1);
        const int ty = threadIdx.x >> P.log2_tpr;
        const int chunk_cols = EPT << P.log2_tpr;
        const int base = (int)gridDim.x - H;
        int buf = 0;
        if (threadIdx.x == 0) {
            // helpers have no implicit tile: their first claim is an ordinary one (and sees both publications)
            ts.word[0] = helper ? atomicAdd(&la->word, 1ull) : 0ull;
        }
        __syncthreads();
        int tile, skip_row, skip_col;
        {
            const unsigned long long w = ts.word[0];
            tile = helper ? (int)(w & kTicketMask) + base : (int)blockIdx.x - H;
            skip_row = helper ? (int)((w >> kTicketBits) & kRowMask) - 1 : -1;
            skip_col = helper ? (int)(w >> kColShift) - 1 : -1;
        }
        if (tile < ntiles) {
            const int tmap0 = reverse ? (ntiles - 1 - tile) : tile;
            const int rb0 = tmap0 / P.nchunks;
            for (int e = threadIdx.x; e < tile_rows; e += kSelBlock) {
                const int k = rb0 * tile_rows + e;
                ts.row[0][e] = (k < nlive) ? __ldg(rlist + k) : -1;
                ts.val[0][e] = (k < nlive) ? __ldg(rval + k) : (real)0;
            }
        }
        __syncthreads();
        int rec_pending = -1;   // thread 0: tile whose completion record is still to be written
        int cur_chunk = -1;
        real sreg[EPT];
        while (tile < ntiles) {
            unsigned long long wnext = 0ull;
            if (threadIdx.x == 0) wnext = atomicAdd(&la->word, 1ull);
            const int tmap = reverse ? (ntiles - 1 - tile) : tile;
            const int chunk = tmap % P.nchunks;
            const int c = chunk * chunk_cols + tx * EPT;
            real a[U];
            int row[U];   // stored row of list entry ty + u*rpp of this tile; -1: nothing to do
            PackView<real, VB> v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int r = ts.row[buf][ty + u * rpp];
                a[u] = ts.val[buf][ty + u * rpp];
                row[u] = (r == skip_row || c >= P.ld) ? -1 : r;
            }
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (row[u] >= 0) v[u].p = ld_pack<0>(reinterpret_cast<const Pack<VB>*>(P.T + (long long)row[u] * P.ld + c));
            if (chunk != cur_chunk && c < P.ld) {
                cur_chunk = chunk;
#pragma unroll
                for (int e = 0; e < EPT; ++e) sreg[e] = __ldg(svec + c + e);
            }
            if (threadIdx.x < 32) {
                // warp 0: the record of the previous tile goes out, the next tile's list entries come in
                if (threadIdx.x == 0 && rec_pending >= 0) st_release_u32(P.tile_rec + rec_pending, seq);
                wnext = __shfl_sync(0xffffffffu, wnext, 0);
                const int ntile = (int)(wnext & kTicketMask) + base;
                if (threadIdx.x == 0) ts.word[buf ^ 1] = wnext;
                if (ntile < ntiles) {
                    const int ntmap = reverse ? (ntiles - 1 - ntile) : ntile;
                    const int nrb = ntmap / P.nchunks;
                    for (int e = threadIdx.x; e < tile_rows; e += 32) {
                        const int k = nrb * tile_rows + e;
                        ts.row[buf ^ 1][e] = (k < nlive) ? __ldg(rlist + k) : -1;
                        ts.val[buf ^ 1][e] = (k < nlive) ? __ldg(rval + k) : (real)0;
                    }
                }
            }
            const int he = skip_col - c;   // lane of a column published before this tile was claimed: held old
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (row[u] >= 0) {
#pragma unroll
                    for (int e = 0; e < EPT; ++e) {
                        const real se = (e == he) ? (real)0 : sreg[e];
                        v[u].e[e] = fma_r(se, a[u], v[u].e[e]);
                    }
                    st_pack<0>(reinterpret_cast<Pack<VB>*>(P.T + (long long)row[u] * P.ld + c), v[u].p);
                }
            if (lp >= c && lp < c + EPT) {
                // the thread that owns the pivot column overwrites its entries with a_pr / pivot (src/solver.cu:43)
                const real pv = (real)__ldcg(&cur->piv);
#pragma unroll 1
                for (int u = 0; u < U; ++u) {
                    const int r = ts.row[buf][ty + u * rpp];
                    if (r >= 0 && r != skip_row) P.T[(long long)r * P.ld + lp] = div_r(ts.val[buf][ty + u * rpp], pv);
                }
            }
            // tiles claimed after both publications are never waited for: no record needed
            rec_pending = (skip_col >= 0) ? -1 : tmap;
            __syncthreads();
            buf ^= 1;
            const unsigned long long w = ts.word[buf];
            tile = (int)(w & kTicketMask) + base;
            skip_row = (int)((w >> kTicketBits) & kRowMask) - 1;
            skip_col = (int)(w >> kColShift) - 1;
        }
        if (threadIdx.x == 0 && rec_pending >= 0) st_release_u32(P.tile_rec + rec_pending, seq);
    }

    // ---- the last CTA to leave commits the pivot and re-arms the ticket word -------------------------------
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned done = atomicAdd(&la->tile_done, 1u);
        if (done == gridDim.x - 1) {
            __threadfence();
            const Proposal* nxt = &la->prop[par ^ 1];
            P.base[p] = q;   // src/solver.cu:105
            if (pivots < P.trace_cap) P.trace[pivots] = make_int2(q, p);
            unsigned long long hsh = st->hash;
            const unsigned int words[2] = {(unsigned)q, (unsigned)p};
#pragma unroll
            for (int wd = 0; wd < 2; ++wd)
#pragma unroll
                for (int by = 0; by < 4; ++by) {
                    hsh ^= (words[wd] >> (8 * by)) & 0xffu;
                    hsh *= 1099511628211ULL;
                }
            st->hash = hsh;
            st->q = q;
            st->p = p;
            st->rows_streamed += (long long)nlive + 1;
            int next_status = kStatusPeerTimeout;   // an incomplete chain means somebody stopped publishing
            if (__ldcg(&nxt->q_seq) == seq + 1u) {
                const int sn = __ldcg(&nxt->status_next);
                if (sn != kRunning)
                    next_status = sn;
                else if (__ldcg(&nxt->ready_seq) == seq + 1u)
                    next_status = kRunning;
            }
            st->status = next_status;
            st->pivots = pivots + 1;
            la->word = 0ull;
            la->tile_done = 0u;
            la->stamps[6] = globaltimer();
            __threadfence();
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Prologue: the same chain on a quiescent tableau (first pivot of a phase / of an iterate() call).  H CTAs.
// Does nothing when the proposal of the next pivot already exists (left by the previous update_la_kernel).
// ---------------------------------------------------------------------------------------------
template <typename real>
__global__ void __launch_bounds__(kSelBlock, 1) la_prologue_kernel(const __grid_constant__ PivotParams<real> P)
{
    __shared__ TreeSmem<real> sm;
    __shared__ real smax[32];
    __shared__ LaShared sh;
    DevState* st = P.st;
    LaState* la = P.la;
    const int status = __ldcg(&st->status);
    const long long pivots = __ldcg(&st->pivots), limit = __ldcg(&st->limit);
    if (status != kRunning || pivots >= limit) return;
    const unsigned seq = (unsigned)pivots;   // the chain prepares pivot seq + 1
    if (__ldcg(&la->prop[(seq + 1u) & 1u].ready_seq) == seq + 1u) return;
    la_chain<real, false>(P, la, seq, (int)blockIdx.x, (int)gridDim.x, nullptr, nullptr, (real)0, -1, -1, -1, false, 0, sm, smax, sh);
}

// The prologue's verdicts (unbounded at the first pivot, a silent peer) have no running pivot to ride on.
template <typename real>
__global__ void la_prologue_commit_kernel(PivotParams<real> P)
{
    DevState* st = P.st;
    if (__ldcg(&st->status) != kRunning || __ldcg(&st->pivots) >= __ldcg(&st->limit)) return;
    const unsigned tseq = (unsigned)(__ldcg(&st->pivots) + 1);
    const Proposal* nxt = &P.la->prop[tseq & 1u];
    if (__ldcg(&nxt->ready_seq) == tseq) return;
    const int sn = __ldcg(&nxt->status_next);
    st->status = (__ldcg(&nxt->p_seq) == tseq && sn != kRunning) ? sn : kStatusPeerTimeout;
}

// Write the prepared pivot constraint back into its (possibly held-old) tableau column, so that the host --
// price-out, phase switch, solution, parity tests -- sees the tableau the reference would hold.
template <typename real>
__global__ void __launch_bounds__(256) la_flush_kernel(PivotParams<real> P)
{
    DevState* st = P.st;
    const unsigned tseq = (unsigned)(__ldcg(&st->pivots) + 1);
    const Proposal* nxt = &P.la->prop[tseq & 1u];
    if (__ldcg(&nxt->ready_seq) != tseq) return;
    const long long lpn = (long long)__ldcg(&nxt->p) - P.col0;
    if (lpn < 0 || lpn >= P.m_loc) return;
    const real* rowpN = la_rowp(P, (int)(tseq & 1u));
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < P.Rs; r += (long long)gridDim.x * blockDim.x)
        P.T[r * P.ld + lpn] = rowpN[r];
}

}  // namespace b2s
