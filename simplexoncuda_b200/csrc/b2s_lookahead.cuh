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
//   * row 0 (RHS) is not in the row list; the helpers update it.
//   * Tiles are handed out by ONE 64-bit atomic word: [ticket count | published row | published column | quiet bit].
//     The entering variable's row and the next pivot column are published by the helpers by or-ing them into that
//     word (every helper ors the same field in -- idempotent); the value the atomic returns is the number of tiles
//     claimed before that helper's publication.  A CTA that claims a tile learns, from the same atomic that gives it the
//     tile, which publications preceded its claim -- single-location coherence order, no fences.  It then leaves the
//     published row untouched and holds the published column old; the helpers compute those elements themselves
//     from the old values.
//   * Tiles claimed before a helper's publication may or may not have seen an earlier helper's: the helper waits for
//     the tile's completion record (pivot number + "left the row alone" / "held the column" bits) and reads the new
//     values or applies the update itself accordingly.  Once every helper has classified its rows the
//     quiet bit stops the record traffic.
//   * The next pivot column may thus stay stale in the tableau.  It is never read: during the next pivot its raw
//     values ARE the gathered vector rowp', and the streaming loop overwrites the column with rowp'[r] / pivot
//     (src/solver.cu:43, `col == colPivotIndex`); rows that update skips (a_pr == 0) get the true entry from the
//     helpers (stage 0).  la_flush_kernel writes rowp' back when the host wants to look at the tableau.
//   * Nothing of pivot k+1 is committed by pivot k: basis, trace, hash, counters and status change only when the last
//     CTA of a launch leaves (la_finalize + commit), so a pivot budget stops exactly and an unused proposal stays valid.
//
// Hand-overs.  Under a saturated memory system a dependent global access costs 2-4 us, so the chain avoids them: each
// stage issues all its loads at once; the cost CTAs and the helpers hand over through per-block / per-helper flags that
// everybody polls (no "last CTA" stage); stage 2 of both tournaments is replayed by every helper; the column publication is an idempotent
// atomicOr whose return value is each helper's own ticket snapshot; helper h gathers (owner rank) or receives (other
// ranks, from the owner's helper h) the contiguous slice of the pivot constraint it later compacts into the row list.
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

constexpr int kStatusInternal = -97;  // the loop found no proposal for its pivot (cannot happen; surfaced as an error)

// Ticket word layout.
constexpr int kTicketBits = 21;                                   // tiles claimed (<= 2^20 tiles + one overshoot per CTA)
constexpr int kRowBits = 19;                                      // published stored row + 1 (<= 196609 rows at 65536^2)
constexpr unsigned long long kTicketMask = (1ull << kTicketBits) - 1ull;
constexpr unsigned long long kRowMask = (1ull << kRowBits) - 1ull;
constexpr int kColShift = kTicketBits + kRowBits;                 // published local column + 1 (<= 2^23)
constexpr unsigned kNoColumn = 0x7ffffeu;                         // "no column will be published during this pivot"
constexpr unsigned long long kColMask = (1ull << 23) - 1ull;
constexpr unsigned long long kQuietBit = 1ull << 63;              // the helpers are done classifying: stop writing tile records
constexpr unsigned kRecHeld = 0x80000000u;                        // tile record: the claimer held the published column old
constexpr unsigned kRecRowHeld = 0x40000000u;                     // tile record: the claimer left the published row alone
constexpr unsigned kRecSeqMask = 0x3fffffffu;                     // tile record: pivot number (mod 2^30)

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
    unsigned pad1, pad2, pad3, pad4;
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
    unsigned slot_seq[kLaMaxHelpers]; // one GPU: helper h has written its ratio-test block winners for pivot slot_seq[h]
    unsigned done_seq[kLaMaxHelpers]; // helper h has finished the chain for pivot done_seq[h]
    unsigned cost_seq[kMaxSlots];     // cost stage-1 block b has written its winner for pivot cost_seq[b]
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
__device__ __noinline__ bool la_wait_u32(const unsigned* flag, unsigned want, long long cycles, int* s_ok)
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
__device__ __noinline__ bool wait_flag_cycles(const unsigned long long* flag, unsigned long long seq, long long cycles)
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

__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned* p)
{
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// Per-thread bounded wait for a tile's completion record (pivot number, possibly with the held-column bit); returns it.
__device__ __noinline__ unsigned la_poll_rec(const unsigned* rec, unsigned seq, long long cycles)
{
    const long long t0 = clock64();
    unsigned v;
    while (((v = ld_acquire_u32(rec)) & kRecSeqMask) != (seq & kRecSeqMask)) {
        if (clock64() - t0 > cycles) break;
    }
    return v;
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

// Publications carried by a ticket word: row to leave alone, column to hold old (-1: none), and whether the tile's completion
// record is still wanted.
struct LaClaim {
    int tile;       // ticket count (add the number of implicitly claimed tiles)
    int skip_row;
    int skip_col;
    bool record;
};
__device__ __forceinline__ LaClaim la_decode(unsigned long long w)
{
    LaClaim c;
    c.tile = (int)(w & kTicketMask);
    c.skip_row = (int)((w >> kTicketBits) & kRowMask) - 1;
    const unsigned col = (unsigned)((w >> kColShift) & kColMask);
    c.skip_col = (col == 0u || col == kNoColumn) ? -1 : (int)col - 1;
    c.record = !(w & kQuietBit) && col != kNoColumn;
    return c;
}

// Loads of data other SMs wrote earlier.  Per-launch kernel (COH = false): written by an EARLIER launch, so the read-only
// path is fine.  Persistent variant (COH = true): several pivots run inside one launch, L1 is never invalidated in between,
// so everything another SM may have written comes from L2.
template <bool COH, typename X>
__device__ __forceinline__ X la_ld(const X* p)
{
    return COH ? __ldcg(p) : __ldg(p);
}
template <bool COH, typename X>
__device__ __forceinline__ X la_ldrw(const X* p)   // read-write data of this kernel (the cost vector)
{
    return COH ? __ldcg(p) : *p;
}
__device__ __forceinline__ long long ld_acquire_s64(const long long* p)
{
    long long v;
    asm volatile("ld.acquire.gpu.global.s64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_s64(long long* p, long long v)
{
    asm volatile("st.release.gpu.global.s64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

struct LaShared {
    unsigned long long next_word;
    int ok;
    int flag;
    int i0, i1;
    double d0, d1;
};

// ---------------------------------------------------------------------------------------------
// The chain runs once per pivot on a handful of SMs whose instruction caches are cold, while the memory system
// is saturated by the streaming CTAs: every instruction-cache miss costs microseconds.  Its code is therefore
// kept small -- divisions, tournament trees and the per-block ratio test are out-of-line functions shared by all
// call sites instead of being inlined a dozen times (216 KB -> tens of KB of SASS; profiles/r02_lookahead.md).
// ---------------------------------------------------------------------------------------------
template <typename real>
__device__ __noinline__ real la_div(real a, real b)
{
    return div_r(a, b);
}
template <typename real>
__device__ __noinline__ void la_tree512(int rule, Cand<real>& c, TreeSmem<real>& sm)
{
    block_tree_512(rule, c, sm);
}
template <typename real>
__device__ __noinline__ void la_stage2(int rule, const real* slot_v, const int* slot_i, const int* slot_k, int G, Cand<real>& w,
                                       TreeSmem<real>& sm)
{
    if (G > 1) {
        stage2_1024(rule, slot_v, slot_i, slot_k, G, w, sm);
    } else {
        w.v = __ldcg(slot_v);
        w.i = __ldcg(slot_i);
        w.k = __ldcg(slot_k);
    }
}
template <typename real>
__device__ __noinline__ real la_max512(real v, real* smax)
{
    return block_max_512(v, smax);
}

// ---------------------------------------------------------------------------------------------
// Stage "cost": cost update of the running pivot (src/solver.cu:48-56) + entering tournament of the next
// (src/reduction.cu:51-104 over costsVector+1).  Block b plays reference stage-1 block b and raises flag b; stage 2 is
// replayed by every helper (la_chain, stage Q).  Runs on the CTAs first_cta .. first_cta+ncta-1.
// ---------------------------------------------------------------------------------------------
template <typename real, bool COH>
__device__ __noinline__ void la_cost_blocks(const PivotParams<real>& P, LaState* la, unsigned tseq, const real* rowp, real sc,
                                            int first_cta, int ncta, TreeSmem<real>& sm)
{
    const long long Nc = P.Rc - 1;
    const int rule = P.rule;
    for (int b = (int)blockIdx.x - first_cta; b < P.Gc; b += ncta) {
        Cand<real> c;
        c.v = Limits<real>::big();
        c.i = -1;
        c.k = -1;
        for (long long i = (long long)b * kSelBlock + threadIdx.x; i < Nc; i += (long long)kSelBlock * P.Gc) {
            const long long j = 1 + i;
            real v = la_ldrw<COH>(P.cost + j);
            v = fma_r(sc, la_ld<COH>(rowp + stored_row(P, j)), v);  // src/solver.cu:54
            P.cost[j] = v;
            Cand<real> o;
            o.v = v;
            o.i = (int)i;
            o.k = (rule == kRuleBland) ? (cmp3((double)v, 0.0) < 0 ? (int)i : -1) : (int)i;
            if (beats(rule, o, c)) c = o;
        }
        if (b == 0 && threadIdx.x == 0) P.cost[0] = fma_r(sc, la_ld<COH>(rowp), la_ldrw<COH>(P.cost));  // objective value
        la_tree512(rule, c, sm);
        if (threadIdx.x == 0) {
            // block winner + its flag: every helper polls the flags and replays stage 2 itself (no "last CTA" stage)
            P.cslot_v[b] = c.v;
            P.cslot_i[b] = c.i;
            P.cslot_k[b] = c.k;
            st_release_u32(&la->cost_seq[b], tseq);
        }
        __syncthreads();
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
    mx = la_max512(mx, smax);
    const int tree_rule = (P.rule == kRuleReference) ? kRuleReference : kRuleLowest;
    la_stage2(tree_rule, slot_v, slot_i, slot_k, P.Gm, w, sm);
}

// One reference stage-1 block of the ratio test (src/reduction.cu:106-140) on values already in registers: entering
// column entry a (after the running update), RHS entry bb.  Writes col'[li]; block winner and block max valid in thread 0.
template <typename real>
__device__ __noinline__ void la_ratio_block(const PivotParams<real>& P, real a, real bb, long long li, int bvar, real* colN,
                                            Cand<real>& c, real& mx, TreeSmem<real>& sm, real* smax)
{
    const int rule = P.rule;
    const int tree_rule = (rule == kRuleReference) ? kRuleReference : kRuleLowest;
    c.v = Limits<real>::big();
    c.i = -1;
    c.k = -1;
    mx = Limits<real>::tiny();
    if (li < P.m_loc) {
        colN[li] = a;
        mx = fmax(mx, a);   // src/reduction.cu:143-184, identity DBL_MIN
        const long long gi = P.col0 + li;
        Cand<real> o;
        o.v = (cmp3((double)a, 0.0) > 0) ? div_r(bb, a) : Limits<real>::big();   // src/reduction.cu:106-114
        o.i = (int)gi;
        o.k = (rule == kRuleBland) ? ((o.v < Limits<real>::big()) ? bvar : -1) : (int)gi;
        if (beats(tree_rule, o, c)) c = o;
    }
    mx = block_max_512(mx, smax);
    block_tree_512(tree_rule, c, sm);
}

// ---------------------------------------------------------------------------------------------
// The helpers' chain.  LIVE = true: called from inside update_la_kernel for pivot `seq` (current proposal:
// q, p, piv; vectors svec / rowp and the row list of the running update) and builds the proposal of pivot
// seq+1 while the other CTAs stream.  LIVE = false: the tableau is quiescent (prologue kernel: first pivot of
// a phase or of an iterate() call); q' comes from the select kernel (st->q / st->cq), every element is final.
//
// The memory system is saturated by the streaming CTAs while this runs, so a dependent global access costs
// 2-3 us (measured: profiles/r02_la_stage_profile_v1_*.jsonl).  Every stage therefore issues all of its loads
// at once (batches of kLaRB per thread) and the stages hand over through as few flags as possible.
// ---------------------------------------------------------------------------------------------
constexpr int kLaBB = 2;   // ratio-test blocks per batch and helper
constexpr int kLaRB = 4;   // pivot-constraint rows per batch and thread

// Bounded wait until the `count` flags flags[0 .. count) (stride in elements) all equal `want`; thread i polls flag i.
__device__ __noinline__ bool la_wait_many_u32(const unsigned* flags, int count, unsigned want, long long cycles)
{
    int ok = 1;
    if ((int)threadIdx.x < count) {
        const long long t0 = clock64();
        while (ld_acquire_u32(flags + threadIdx.x) != want) {
            if (clock64() - t0 > cycles) {
                ok = 0;
                break;
            }
        }
    }
    return __syncthreads_and(ok) != 0;
}
// The same for any number of flags (each thread polls flags tid, tid + 512, ...).
__device__ __noinline__ bool la_wait_all_u32(const unsigned* flags, int count, unsigned want, long long cycles)
{
    int ok = 1;
    const long long t0 = clock64();
    for (int i = threadIdx.x; i < count && ok; i += kSelBlock) {
        while (ld_acquire_u32(flags + i) != want) {
            if (clock64() - t0 > cycles) {
                ok = 0;
                break;
            }
        }
    }
    return __syncthreads_and(ok) != 0;
}
// Block-uniform bounded wait for a tile's completion record; returns the record (0 on timeout).
__device__ __noinline__ unsigned la_wait_rec(const unsigned* rec, unsigned seq, long long cycles, LaShared& sh)
{
    if (threadIdx.x == 0) {
        const unsigned v = la_poll_rec(rec, seq, cycles);
        sh.i1 = (int)(((v & kRecSeqMask) == (seq & kRecSeqMask)) ? v : 0u);
    }
    __syncthreads();
    const unsigned v = (unsigned)sh.i1;
    __syncthreads();
    return v;
}
__device__ __noinline__ bool la_wait_many_sys(const unsigned long long* flags, int count, unsigned long long want, long long cycles)
{
    int ok = 1;
    if ((int)threadIdx.x < count) ok = wait_flag_cycles(flags + threadIdx.x, want, cycles) ? 1 : 0;
    return __syncthreads_and(ok) != 0;
}

template <typename real, bool LIVE, bool COH>
__device__ __noinline__ void la_chain(const PivotParams<real>& P, LaState* la, unsigned seq, int h, int H,
                                      const real* rowp, const real* svec, real piv, long long lp, int p_cur, int q_cur,
                                      bool reverse, long long ntiles, TreeSmem<real>& sm, real* smax, LaShared& sh)
{
    __shared__ int s_scan[kLaRB][kSelBlock / 32];
    DevState* st = P.st;
    const unsigned tseq = seq + 1u;     // the pivot this chain prepares
    const int par = (int)(seq & 1u), tpar = par ^ 1;
    Proposal* nxt = &la->prop[tpar];
    const bool sharded = P.world > 1;
    const int rule = P.rule;
    const long long cyc = P.wait_cycles;
    const long long base = (long long)gridDim.x - H;   // tiles claimed implicitly at launch (LIVE only)
    real* colN = P.col2 + (size_t)tpar * P.ld;
    real* sN = P.s2 + (size_t)tpar * P.ld;
    real* rowpN = la_rowp(P, tpar);
    const int* posC = P.rowpos + (size_t)par * P.rowp_stride;   // row -> position in the running update's row list
    const int rpp = kSelBlock >> P.log2_tpr;
    const long long tile_rows = (long long)rpp * P.la_u;
    const long long chunk_cols = (long long)(32 / (int)sizeof(real)) << P.log2_tpr;
    const real a0 = LIVE ? la_ld<COH>(rowp) : (real)0;   // rowp[0] = b_p of the running pivot
    const bool own_cur = LIVE && lp >= 0 && lp < P.m_loc;
    // Helper h owns the contiguous slice [r_lo, r_hi) of stored rows for the pivot-constraint gather and the row list.
    const long long slice = (((P.Rs + H - 1) / H) + kSelBlock - 1) / kSelBlock * kSelBlock;
    const long long r_lo = (long long)h * slice, r_hi = (r_lo + slice < P.Rs) ? r_lo + slice : P.Rs;
    const int nchunk = r_hi > r_lo ? (int)((r_hi - r_lo + kSelBlock - 1) / kSelBlock) : 0;
    const int nbatch = (nchunk + kLaRB - 1) / kLaRB;
    // what the gather will need from the running update for the first batch of its slice: fetched now, used much later
    int pos0[kLaRB];
    real ak0[kLaRB];
#pragma unroll
    for (int k = 0; k < kLaRB; ++k) {
        const long long r = r_lo + (long long)k * kSelBlock + threadIdx.x;
        pos0[k] = (LIVE && r < r_hi) ? la_ld<COH>(posC + r) : -2;
        ak0[k] = (LIVE && r < r_hi) ? la_ld<COH>(rowp + r) : (real)0;
    }

    if (LIVE) {
        // ---- stage 0: row 0 (RHS) of the running update -- it is not in the row list ---------------------------
        for (int bl0 = h; bl0 < P.Gm_loc; bl0 += H * kLaBB) {
            real x[kLaBB], sv[kLaBB];
#pragma unroll
            for (int k = 0; k < kLaBB; ++k) {
                const long long li = (long long)(bl0 + k * H) * kSelBlock + threadIdx.x;
                const bool ok = bl0 + k * H < P.Gm_loc && li < P.m_loc;
                x[k] = ok ? __ldcg(P.T + li) : (real)0;
                sv[k] = ok ? la_ld<COH>(svec + li) : (real)0;
            }
#pragma unroll
            for (int k = 0; k < kLaBB; ++k) {
                const long long li = (long long)(bl0 + k * H) * kSelBlock + threadIdx.x;
                if (bl0 + k * H < P.Gm_loc && li < P.m_loc) P.T[li] = (li == lp) ? la_div(a0, piv) : fma_r(sv[k], a0, x[k]);
            }
        }
        // rows the list leaves out (a_pr == 0) are unchanged by this update, pivot column included: a_pr / pivot = a_pr.
        // The tableau may still hold a held-old value there (see the header): write the true entry.
        if (own_cur && P.skip_zero) {
            for (int bt = 0; bt < nbatch; ++bt) {
#pragma unroll
                for (int k = 0; k < kLaRB; ++k) {
                    const long long r = r_lo + ((long long)bt * kLaRB + k) * kSelBlock + threadIdx.x;
                    if (r < r_hi && r != 0) {
                        const int pos = bt == 0 ? pos0[k] : la_ld<COH>(posC + r);
                        if (pos < 0) P.T[r * P.ld + lp] = bt == 0 ? ak0[k] : la_ld<COH>(rowp + r);
                    }
                }
            }
        }
        if (h == 0 && threadIdx.x == 0) la->stamps[1] = globaltimer();
        // ---- stage Q: every helper replays stage 2 of the entering tournament over the cost CTAs' block winners ----------
        if (!la_wait_all_u32(la->cost_seq, P.Gc, tseq, cyc)) {
            if (threadIdx.x == 0) atomicOr(&la->word, (unsigned long long)kNoColumn << kColShift);
            return;   // (q_seq stays behind: la_finalize reports the silent stage)
        }
    }
    int qn;
    real cqn;
    long long c_row = 0, rbq = -1;   // rbq: row block of the running list that holds row 1+q'; -1: no tile touches it
    real aq = (real)0;               // rowp[1+q'] of the running pivot
    if (LIVE) {
        Cand<real> wq;
        la_stage2(rule, P.cslot_v, P.cslot_i, P.cslot_k, P.Gc, wq, sm);
        if (threadIdx.x == 0) {
            sh.i0 = wq.i;
            sh.d0 = (double)wq.v;
        }
        __syncthreads();
        qn = sh.i0;
        cqn = (real)sh.d0;
        __syncthreads();
        const bool go = qn >= 0 && cmp3((double)cqn, 0.0) < 0;   // src/solver.cu:87-88
        if (h == 0 && threadIdx.x == 0) {
            nxt->q = qn;
            nxt->cq = (double)cqn;
            nxt->status_next = go ? kRunning : kFeasible;
        }
        if (!go) {   // optimal after this pivot: nothing to prepare, no column will be published
            if (threadIdx.x == 0) {
                atomicOr(&la->word, (unsigned long long)kNoColumn << kColShift);
                if (h == 0) st_release_u32(&nxt->q_seq, tseq);
            }
            return;
        }
        // Publish the entering variable's row: tiles claimed from now on leave it to the helpers.  Every helper ORs the same
        // field in and keeps the ticket count ITS atomic returned (as for the column below); meanwhile the two lookups fly.
        const long long rq0 = stored_row(P, 1 + (long long)qn);
        const int posq = la_ld<COH>(posC + rq0);
        aq = la_ld<COH>(rowp + rq0);
        if (threadIdx.x == 0) {
            const unsigned long long old = atomicOr(&la->word, (unsigned long long)(rq0 + 1) << kTicketBits);
            sh.next_word = old & kTicketMask;
            if (h == 0) st_release_u32(&nxt->q_seq, tseq);
        }
        __syncthreads();
        c_row = (long long)sh.next_word;
        __syncthreads();
        if (posq >= 0) rbq = posq / tile_rows;
    } else {
        if (h == 0 && threadIdx.x == 0) {
            nxt->q = __ldcg(&st->q);
            nxt->cq = __ldcg(&st->cq);
            nxt->status_next = kRunning;
        }
        qn = __ldcg(&st->q);
        cqn = (real)__ldcg(&st->cq);
    }
    const long long rq = stored_row(P, 1 + (long long)qn);
    if (h == 0 && threadIdx.x == 0) la->stamps[2] = globaltimer();

    // ---- stage R: entering column + RHS after the running update, ratio-test stage 1 ----------------------
    for (int bl0 = h; bl0 < P.Gm_loc; bl0 += H * kLaBB) {
        bool old_vals[kLaBB];
        real av[kLaBB], bv[kLaBB], sv[kLaBB];
#pragma unroll
        for (int k = 0; k < kLaBB; ++k) {
            const int bl = bl0 + k * H;
            old_vals[k] = false;
            if (LIVE && bl < P.Gm_loc) {
                old_vals[k] = true;
                if (rbq >= 0) {
                    const int chunk = (int)(((long long)bl * kSelBlock) / chunk_cols);
                    const long long tmap = rbq * P.nchunks + chunk;
                    const long long t = reverse ? (ntiles - 1 - tmap) : tmap;
                    old_vals[k] = t >= c_row + base;  // claimed after my snapshot: the claimer certainly leaves row rq alone
                    if (!old_vals[k]) {               // claimed before: wait until that tile is complete; its record says what it did
                        const unsigned rec = la_wait_rec(P.tile_rec + tmap, seq, cyc, sh);
                        if (rec == 0u) {
                            if (threadIdx.x == 0) nxt->status_next = kStatusPeerTimeout;
                            return;
                        }
                        old_vals[k] = (rec & kRecRowHeld) != 0u;
                    }
                }
            }
        }
#pragma unroll
        for (int k = 0; k < kLaBB; ++k) {
            const long long li = (long long)(bl0 + k * H) * kSelBlock + threadIdx.x;
            const bool ok = bl0 + k * H < P.Gm_loc && li < P.m_loc;
            av[k] = ok ? __ldcg(P.T + rq * P.ld + li) : (real)0;
            bv[k] = ok ? __ldcg(P.T + li) : (real)0;
            sv[k] = (ok && old_vals[k]) ? la_ld<COH>(svec + li) : (real)0;
        }
#pragma unroll
        for (int k = 0; k < kLaBB; ++k) {   // (unrolled: av/bv/sv live in registers; the body is a call)
            const int bl = bl0 + k * H;
            if (bl >= P.Gm_loc) break;
            const int gb = P.Gm_loc0 + bl;
            const long long li = (long long)bl * kSelBlock + threadIdx.x;
            real a = av[k];
            int bvar = -1;
            if (li < P.m_loc) {
                if (old_vals[k]) {
                    a = (li == lp) ? la_div(aq, piv) : fma_r(sv[k], aq, a);
                    P.T[rq * P.ld + li] = a;
                }
                if (rule == kRuleBland) {
                    const long long gi = P.col0 + li;
                    bvar = (LIVE && gi == p_cur) ? q_cur : la_ld<COH>(P.base + gi);   // base[p] = q of the running pivot is committed at its end
                }
            }
            Cand<real> c;
            real mx;
            la_ratio_block(P, a, bv[k], li, bvar, colN, c, mx, sm, smax);
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
                    ArenaHeader<real>* a2 = arena_of(P, threadIdx.x);
                    a2->slot_v[tpar][gb] = (real)sh.d0;
                    a2->slot_max[tpar][gb] = (real)sh.d1;
                    a2->slot_i[tpar][gb] = sh.i0;
                    a2->slot_k[tpar][gb] = sh.i1;
                }
            }
            __syncthreads();
        }
    }
    // My block winners are out: raise my flag (everywhere, when sharded), wait for every helper's (of every rank), then
    // EVERY helper replays stage 2 itself -- no "last CTA", no second hand-over.
    const real *slot_v = P.rslot_v, *slot_max = P.rslot_max;
    const int *slot_i = P.rslot_i, *slot_k = P.rslot_k;
    bool ok_x = true;
    if (!sharded) {
        if (threadIdx.x == 0) st_release_u32(&la->slot_seq[h], tseq);
        ok_x = la_wait_many_u32(la->slot_seq, H, tseq, cyc);
    } else {
        const bool mute = P.fault_rank == P.rank && (long long)tseq >= P.fault_pivot;   // fault injection (tests)
        // (thread r wrote this helper's winners into rank r's arena itself: its release store orders them, no extra fence)
        if (threadIdx.x < P.world && !mute)
            st_release_sys(&arena_of(P, threadIdx.x)->la_flag_slots[tpar][P.rank][h], (unsigned long long)tseq);
        // flags of rank r, helper k: la_flag_slots[tpar][r][k]; ranks >= world never publish, so poll rank by rank
        const ArenaHeader<real>* mine = arena_of(P, P.rank);
        int okk = 1;
        if ((int)threadIdx.x < P.world * H) {
            const int r = threadIdx.x / H, k = threadIdx.x % H;
            okk = wait_flag_cycles(&mine->la_flag_slots[tpar][r][k], (unsigned long long)tseq, cyc) ? 1 : 0;
        }
        ok_x = __syncthreads_and(okk) != 0;
        slot_v = mine->slot_v[tpar];
        slot_max = mine->slot_max[tpar];
        slot_i = mine->slot_i[tpar];
        slot_k = mine->slot_k[tpar];
    }
    Cand<real> w;
    w.v = Limits<real>::big();
    w.i = -1;
    w.k = -1;
    real mxall = Limits<real>::tiny();
    if (ok_x) la_ratio_stage2(P, slot_v, slot_i, slot_k, slot_max, sm, smax, w, mxall);
    if (threadIdx.x == 0) {
        int pnn = -1, snn = kRunning;
        if (!ok_x)
            snn = kStatusPeerTimeout;
        else if (cmp3((double)mxall, 0.0) <= 0 || w.i < 0)
            snn = kUnbounded;   // src/solver.cu:98-99
        else
            pnn = w.i;
        sh.i0 = pnn;
        if (h == 0) {
            nxt->p = pnn;
            if (snn != kRunning) nxt->status_next = snn;
        }
    }
    __syncthreads();
    const int pn = sh.i0;
    __syncthreads();
    if (h == 0 && threadIdx.x == 0) la->stamps[3] = globaltimer();
    const long long lpn = (long long)pn - P.col0;
    const bool owner = pn >= 0 && lpn >= 0 && lpn < P.m_loc;
    const bool same_col = LIVE && owner && lpn == lp;   // the same constraint leaves twice in a row
    long long c_col = 0;
    if (LIVE) {
        // Publish the next pivot column (its owner rank only, and not when the same constraint leaves twice in a row): tiles
        // claimed from now on hold that column old.  Every helper ORs the same field in and keeps the ticket count ITS atomic
        // returned: tiles at or beyond that count certainly saw the column; earlier ones say in their record what they did.
        if (threadIdx.x == 0) {
            const unsigned long long field = (owner && !same_col) ? (unsigned long long)(lpn + 1) : (unsigned long long)kNoColumn;
            const unsigned long long old = atomicOr(&la->word, field << kColShift);
            sh.next_word = old & kTicketMask;
        }
        __syncthreads();
        c_col = (long long)sh.next_word;
        __syncthreads();
    }
    if (pn < 0) return;   // unbounded (or a peer went silent): the phase ends instead of pivot tseq

    // ---- stage G: the owner of constraint p' gathers the raw pivot constraint after the running update ----
    const bool skip = P.skip_zero != 0;
    real v0[kLaRB];   // values of the first batch stay in registers for the list pass
#pragma unroll
    for (int k = 0; k < kLaRB; ++k) v0[k] = (real)0;
    int mine_cnt = 0;
    if (!owner) {   // sharded, another rank owns p': wait for its helper h, which delivers exactly my slice
        if (threadIdx.x == 0)
            sh.ok = wait_flag_cycles(&arena_of(P, P.rank)->la_flag_rowp[tpar][h], (unsigned long long)tseq, cyc) ? 1 : 0;
        __syncthreads();
        const bool ok = sh.ok != 0;
        __syncthreads();
        if (!ok) {
            if (h == 0 && threadIdx.x == 0) nxt->status_next = kStatusPeerTimeout;
            return;
        }
    }
    {
        const int chunk = owner ? (int)(lpn / chunk_cols) : 0;
        const real sp = (LIVE && owner && !same_col) ? la_ld<COH>(svec + lpn) : (real)0;
        int bad = 0;
        for (int bt = 0; bt < nbatch; ++bt) {
            real v[kLaRB];
            const long long rb0 = r_lo + (long long)bt * kLaRB * kSelBlock + threadIdx.x;
            if (!owner) {
#pragma unroll
                for (int k = 0; k < kLaRB; ++k) {
                    const long long r = rb0 + (long long)k * kSelBlock;
                    v[k] = (r < r_hi) ? __ldcg(rowpN + r) : (real)0;
                }
            } else if (same_col) {
#pragma unroll
                for (int k = 0; k < kLaRB; ++k) {
                    const long long r = rb0 + (long long)k * kSelBlock;
                    v[k] = (r < r_hi) ? la_div(bt == 0 ? ak0[k] : la_ld<COH>(rowp + r), piv) : (real)0;   // T'[r][p] = a_pr / pivot (src/solver.cu:43)
                }
            } else {
                int pos[kLaRB];
                real ak[kLaRB];
                unsigned rec[kLaRB];
#pragma unroll
                for (int k = 0; k < kLaRB; ++k) {
                    const long long r = rb0 + (long long)k * kSelBlock;
                    pos[k] = bt == 0 ? pos0[k] : ((LIVE && r < r_hi) ? la_ld<COH>(posC + r) : -2);
                    ak[k] = bt == 0 ? ak0[k] : ((LIVE && r < r_hi) ? la_ld<COH>(rowp + r) : (real)0);
                    if (!LIVE || r == 0 || r == rq || r >= r_hi) pos[k] = -2;
                }
                // -2: final in the tableau (rows 0 and 1+q' were finished by stages 0 / R; quiescent tableau)
                // -1: apply the running update here: not in the running list (a_pr == 0, no tile touches it), or held old by
                //     its claimer
                // >= 0 after classification: updated in full by its tile, read the new value once that tile is complete
                // All records are fetched at once; only tiles still in flight are polled.
#pragma unroll
                for (int k = 0; k < kLaRB; ++k) {
                    rec[k] = 0u;
                    if (pos[k] >= 0) {
                        const long long tmap = (pos[k] / tile_rows) * P.nchunks + chunk;
                        const long long t = reverse ? (ntiles - 1 - tmap) : tmap;
                        if (t >= c_col + base)
                            pos[k] = -1;
                        else
                            rec[k] = ld_relaxed_u32(P.tile_rec + tmap);
                    }
                }
#pragma unroll
                for (int k = 0; k < kLaRB; ++k) {
                    if (pos[k] >= 0) {
                        if ((rec[k] & kRecSeqMask) != (seq & kRecSeqMask)) {
                            const long long tmap = (pos[k] / tile_rows) * P.nchunks + chunk;
                            rec[k] = la_poll_rec(P.tile_rec + tmap, seq, cyc);
                            if ((rec[k] & kRecSeqMask) != (seq & kRecSeqMask)) bad = 1;
                        }
                        pos[k] = (rec[k] & kRecHeld) ? -1 : -2;
                    }
                }
                __threadfence();   // (acquire side of the relaxed record loads)
#pragma unroll
                for (int k = 0; k < kLaRB; ++k) {
                    const long long r = rb0 + (long long)k * kSelBlock;
                    v[k] = (r < r_hi) ? __ldcg(P.T + r * P.ld + lpn) : (real)0;
                }
#pragma unroll
                for (int k = 0; k < kLaRB; ++k)
                    if (pos[k] == -1) v[k] = fma_r(sp, ak[k], v[k]);
            }
            if (owner) {
#pragma unroll
                for (int k = 0; k < kLaRB; ++k) {
                    const long long r = rb0 + (long long)k * kSelBlock;
                    if (r < r_hi) {
                        if (!sharded) {
                            rowpN[r] = v[k];
                        } else {
#pragma unroll 1
                            for (int wr = 0; wr < P.world; ++wr) arena_rowp(P, wr, tpar)[r] = v[k];
                        }
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < kLaRB; ++k) {
                const long long r = rb0 + (long long)k * kSelBlock;
                const int live = (r < r_hi && r != 0 && (!skip || v[k] != (real)0)) ? 1 : 0;
                mine_cnt += __syncthreads_count(live);
            }
            if (bt == 0) {
#pragma unroll
                for (int k = 0; k < kLaRB; ++k) v0[k] = v[k];
            }
        }
        bad = __syncthreads_or(bad);
        if (sharded && owner) {
            // my slice is in every arena: tell helper h of every rank
            // (release is cumulative over the stores the barrier above ordered before it)
            if (threadIdx.x < P.world) st_release_sys(&arena_of(P, threadIdx.x)->la_flag_rowp[tpar][h], (unsigned long long)tseq);
        }
        if (threadIdx.x == 0) {
            if (bad) nxt->status_next = kStatusPeerTimeout;
            la->cnt[h] = mine_cnt;
            st_release_u32(&la->cnt_seq[h], tseq);
        }
    }
    // ---- stage S: everybody's counts -> list offsets; s' = -col'/pivot' for the local slab; the row list -------------
    if (!la_wait_many_u32(la->cnt_seq, H, tseq, cyc)) {
        if (threadIdx.x == 0) nxt->status_next = kStatusPeerTimeout;
        return;
    }
    if (h == 0 && threadIdx.x == 0) la->stamps[4] = globaltimer();
    // (LIVE) every helper has classified its rows: the streaming CTAs may stop writing tile records
    if (LIVE && h == 0 && threadIdx.x == 0) atomicOr(&la->word, kQuietBit);
    int cnts = 0;
    if ((int)threadIdx.x < H) cnts = __ldcg(&la->cnt[threadIdx.x]);
    const real pivn = __ldcg(rowpN + rq);   // a_pq = T[1+q'][p']
    real cv[kLaBB];
#pragma unroll
    for (int k = 0; k < kLaBB; ++k) {
        const long long i = (long long)h * kSelBlock + threadIdx.x + (long long)k * H * kSelBlock;
        cv[k] = (i < P.m_loc) ? __ldcg(colN + i) : (real)0;
    }
    int offset = 0, total = 0;
    if (threadIdx.x < 32) {   // H <= 16 counts live in the first lanes of warp 0
#pragma unroll
        for (int k = 0; k < kLaMaxHelpers; ++k) {
            const int ck = __shfl_sync(0xffffffffu, cnts, k);
            if (k < H) {
                if (k < h) offset += ck;
                total += ck;
            }
        }
        if (threadIdx.x == 0) {
            sh.i0 = offset;
            sh.i1 = total;
        }
    }
    __syncthreads();
    offset = sh.i0;
    total = sh.i1;
    __syncthreads();
    for (long long i0 = (long long)h * kSelBlock + threadIdx.x; i0 < P.ld; i0 += (long long)H * kSelBlock * kLaBB) {
        const bool first = i0 == (long long)h * kSelBlock + threadIdx.x;
#pragma unroll
        for (int k = 0; k < kLaBB; ++k) {
            const long long i = i0 + (long long)k * H * kSelBlock;
            if (i < P.ld) {
                const real cvk = first ? cv[k] : ((i < P.m_loc) ? __ldcg(colN + i) : (real)0);
                sN[i] = (i < P.m_loc && i != lpn) ? la_div(-cvk, pivn) : (real)0;
            }
        }
    }
    {
        int* listN = P.rowlist + (size_t)tpar * P.rowp_stride;
        real* valN = P.rowval + (size_t)tpar * P.rowp_stride;
        int* posN = P.rowpos + (size_t)tpar * P.rowp_stride;
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
        for (int bt = 0; bt < nbatch; ++bt) {
            real v[kLaRB];
            const long long rb0 = r_lo + (long long)bt * kLaRB * kSelBlock + threadIdx.x;
            if (bt == 0) {
#pragma unroll
                for (int k = 0; k < kLaRB; ++k) v[k] = v0[k];
            } else {
#pragma unroll
                for (int k = 0; k < kLaRB; ++k) {
                    const long long r = rb0 + (long long)k * kSelBlock;
                    v[k] = (r < r_hi) ? __ldcg(rowpN + r) : (real)0;
                }
            }
            unsigned bal[kLaRB];
#pragma unroll
            for (int k = 0; k < kLaRB; ++k) {
                const long long r = rb0 + (long long)k * kSelBlock;
                const int live = (r < r_hi && r != 0 && (!skip || v[k] != (real)0)) ? 1 : 0;
                bal[k] = __ballot_sync(0xffffffffu, live);
                if (lane == 0) s_scan[k][wid] = __popc(bal[k]);
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < kLaRB; ++k) {
                const long long r = rb0 + (long long)k * kSelBlock;
                int before = 0, chunk_total = 0;
#pragma unroll
                for (int wq = 0; wq < kSelBlock / 32; ++wq) {
                    const int c = s_scan[k][wq];
                    if (wq < wid) before += c;
                    chunk_total += c;
                }
                if (r < r_hi) {
                    if ((bal[k] >> lane) & 1u) {
                        const int pos = offset + before + __popc(bal[k] & ((1u << lane) - 1u));
                        listN[pos] = (int)r;
                        valN[pos] = v[k];
                        posN[r] = pos;
                    } else {
                        posN[r] = -1;
                    }
                }
                offset += chunk_total;
            }
            __syncthreads();
        }
        if (h == 0 && threadIdx.x == 0) {
            nxt->nz = total;
            nxt->piv = (double)pivn;
            nxt->sc = (double)la_div(-cqn, pivn);   // src/solver.cu:54
        }
    }
    // done: the proposal is complete once every helper has said so (checked by the CTA that commits the pivot)
    __syncthreads();
    if (threadIdx.x == 0) {
        st_release_u32(&la->done_seq[h], tseq);
        if (h == 0) la->stamps[5] = globaltimer();
    }
}

// After the chain: the verdict for pivot tseq.  kRunning (and the proposal is marked ready) iff every helper finished;
// the phase-ending status the chain found (kFeasible / kUnbounded); kStatusPeerTimeout when somebody stopped publishing.
__device__ __forceinline__ int la_finalize(LaState* la, Proposal* nxt, unsigned tseq, int H, bool live)
{
    const unsigned qs = __ldcg(&nxt->q_seq);
    const int sn = __ldcg(&nxt->status_next);
    unsigned ds[kLaMaxHelpers];
#pragma unroll
    for (int k = 0; k < kLaMaxHelpers; ++k) ds[k] = __ldcg(&la->done_seq[k]);   // all in flight together
    if (live && qs != tseq) return kStatusPeerTimeout;
    if (sn != kRunning) return sn;
    bool all = true;
#pragma unroll
    for (int k = 0; k < kLaMaxHelpers; ++k) all = all && (k >= H || ds[k] == tseq);
    if (!all) return kStatusPeerTimeout;
    nxt->ready_seq = tseq;
    return kRunning;
}

// ---------------------------------------------------------------------------------------------
// update_la_kernel -- see the header of this file.  256-bit accesses, 8 rows in flight per thread, tiles of
// (512 >> log2_tpr) * 8 list rows x one column chunk, handed out by the ticket word.
// ---------------------------------------------------------------------------------------------
// PERSIST = false: one pivot per launch (the launches of a batch are replayed as a CUDA graph).  PERSIST = true: up to `batch`
// pivots per cooperative launch; between two pivots the CTA that commits releases the new pivot count and everybody waits
// for it -- a grid barrier with the commit inside, ~2 us instead of a kernel boundary.  Data written during the launch is
// then read through L2 only (la_ld<true>, 256-bit ld.global.cg for the tiles).
template <typename real, int U, bool PERSIST>
__global__ void __launch_bounds__(kSelBlock, 1) update_la_kernel(const __grid_constant__ PivotParams<real> P, int batch)
{
    constexpr int VB = 32;
    constexpr int EPT = VB / (int)sizeof(real);
    __shared__ TreeSmem<real> sm;
    __shared__ real smax[32];
    __shared__ LaShared sh;
    __shared__ LaTileSmem<real> ts;

    DevState* st = P.st;
    LaState* la = P.la;
    // Programmatic dependent launch: this grid may have been scheduled while the previous pivot's grid was draining; nothing
    // may be read before that grid has completed.  Let the next pivot's grid be scheduled behind this one right away -- its
    // CTAs take the SMs as ours leave and wait here.
    if (!PERSIST) {
        pdl_wait();
        pdl_trigger();
    }
  for (int it = 0; it < (PERSIST ? batch : 1); ++it) {
    // One round trip for everything the kernel needs to know: the loop state and BOTH proposals are fetched together and
    // the right generation is picked afterwards (a dependent second and third fetch would cost a microsecond each).
    const int status = __ldcg(&st->status);
    const long long pivots = __ldcg(&st->pivots), limit = __ldcg(&st->limit);
    unsigned ready2[2];
    int p2[2], q2[2];
    long long nz2[2];
    double sc2[2], piv2[2];
#pragma unroll
    for (int g = 0; g < 2; ++g) {
        ready2[g] = __ldcg(&la->prop[g].ready_seq);
        p2[g] = __ldcg(&la->prop[g].p);
        q2[g] = __ldcg(&la->prop[g].q);
        nz2[g] = __ldcg(&la->prop[g].nz);
        sc2[g] = __ldcg(&la->prop[g].sc);
        piv2[g] = __ldcg(&la->prop[g].piv);
    }
    if (status != kRunning || pivots >= limit) return;   // (uniform over the grid: written before the last barrier / launch)
    const unsigned seq = (unsigned)(pivots + 1);
    const int par = (int)(seq & 1u);
    const Proposal* cur = &la->prop[par];
    if ((par ? ready2[1] : ready2[0]) != seq) {
        if (blockIdx.x == 0 && threadIdx.x == 0) st->status = kStatusInternal;
        return;
    }
    const int H = P.helpers;
    const bool helper = (int)blockIdx.x < H;
    const int p = par ? p2[1] : p2[0], q = par ? q2[1] : q2[0];
    const int lp = p - P.col0;   // outside [0, m_loc) when another rank owns the pivot column
    const real* rowp = la_rowp(P, par);
    const real* svec = P.s2 + (size_t)par * P.ld;
    const bool reverse = P.serpentine && (seq & 1u);
    const int rpp = kSelBlock >> P.log2_tpr;
    const int tile_rows = rpp * U;
    const int nlive = (int)(par ? nz2[1] : nz2[0]);
    const int ntiles = ((nlive + tile_rows - 1) / tile_rows) * P.nchunks;
    const real sc_cur = (real)(par ? sc2[1] : sc2[0]), piv_cur = (real)(par ? piv2[1] : piv2[0]);
    if (blockIdx.x == 0 && threadIdx.x == 0) la->stamps[0] = globaltimer();

    {
        // cost update + entering tournament on the first non-helper CTAs, so that the helpers start on the RHS row at once
        const int first = ((int)gridDim.x > H) ? H : 0;
        const int ncta = (int)gridDim.x - first;
        if ((int)blockIdx.x >= first && (int)blockIdx.x - first < P.Gc)
            la_cost_blocks<real, PERSIST>(P, la, seq + 1u, rowp, sc_cur, first, ncta, sm);
    }
    if (helper)
        la_chain<real, true, PERSIST>(P, la, seq, (int)blockIdx.x, H, rowp, svec, piv_cur, (long long)lp, p, q, reverse,
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
        int tile, skip_row = -1, skip_col = -1;
        bool want_rec = true;
        if (helper) {
            const LaClaim cl = la_decode(ts.word[0]);
            tile = cl.tile + base;
            skip_row = cl.skip_row;
            skip_col = cl.skip_col;
            want_rec = cl.record;
        } else {
            tile = (int)blockIdx.x - H;
        }
        if (tile < ntiles) {
            const int tmap0 = reverse ? (ntiles - 1 - tile) : tile;
            const int rb0 = tmap0 / P.nchunks;
            for (int e = threadIdx.x; e < tile_rows; e += kSelBlock) {
                const int k = rb0 * tile_rows + e;
                ts.row[0][e] = (k < nlive) ? la_ld<PERSIST>(rlist + k) : -1;
                ts.val[0][e] = (k < nlive) ? la_ld<PERSIST>(rval + k) : (real)0;
            }
        }
        __syncthreads();
        int rec_pending = -1;   // thread 0: tile whose completion record is still to be written
        unsigned rec_value = seq & kRecSeqMask;
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
                if (row[u] >= 0) v[u].p = ld_pack<PERSIST ? 3 : 0>(reinterpret_cast<const Pack<VB>*>(P.T + (long long)row[u] * P.ld + c));
            if (chunk != cur_chunk && c < P.ld) {
                cur_chunk = chunk;
#pragma unroll
                for (int e = 0; e < EPT; ++e) sreg[e] = la_ld<PERSIST>(svec + c + e);
            }
            if (threadIdx.x < 32) {
                // warp 0: the record of the previous tile goes out, the next tile's list entries come in
                if (threadIdx.x == 0 && rec_pending >= 0) st_release_u32(P.tile_rec + rec_pending, rec_value);
                wnext = __shfl_sync(0xffffffffu, wnext, 0);
                const int ntile = (int)(wnext & kTicketMask) + base;
                if (threadIdx.x == 0) ts.word[buf ^ 1] = wnext;
                if (ntile < ntiles) {
                    const int ntmap = reverse ? (ntiles - 1 - ntile) : ntile;
                    const int nrb = ntmap / P.nchunks;
                    for (int e = threadIdx.x; e < tile_rows; e += 32) {
                        const int k = nrb * tile_rows + e;
                        ts.row[buf ^ 1][e] = (k < nlive) ? la_ld<PERSIST>(rlist + k) : -1;
                        ts.val[buf ^ 1][e] = (k < nlive) ? la_ld<PERSIST>(rval + k) : (real)0;
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
                const real pv = (real)__ldcg(&cur->piv);   // (rare path: re-read rather than keep a register alive through the loop)
#pragma unroll 1
                for (int u = 0; u < U; ++u) {
                    const int r = ts.row[buf][ty + u * rpp];
                    if (r >= 0 && r != skip_row) P.T[(long long)r * P.ld + lp] = div_r(ts.val[buf][ty + u * rpp], pv);
                }
            }
            // the record says whether this tile held the published column old; once the helpers are done nobody reads records
            rec_pending = want_rec ? tmap : -1;
            rec_value = (seq & kRecSeqMask) | (skip_col >= 0 ? kRecHeld : 0u) | (skip_row >= 0 ? kRecRowHeld : 0u);
            __syncthreads();
            buf ^= 1;
            const LaClaim cl = la_decode(ts.word[buf]);
            tile = cl.tile + base;
            skip_row = cl.skip_row;
            skip_col = cl.skip_col;
            want_rec = cl.record;
        }
        if (threadIdx.x == 0 && rec_pending >= 0) st_release_u32(P.tile_rec + rec_pending, rec_value);
    }

    // ---- the last CTA to leave commits the pivot and re-arms the ticket word -------------------------------
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned done = atomicAdd(&la->tile_done, 1u);
        if (done == gridDim.x - 1) {
            __threadfence();
            Proposal* nxt = &la->prop[par ^ 1];
            P.base[p] = q;   // src/solver.cu:105
            if (pivots < P.trace_cap) P.trace[pivots] = make_int2(q, p);
            unsigned long long hsh = __ldcg(&st->hash);   // (L2: in the persistent variant another SM committed the previous pivot)
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
            st->rows_streamed = __ldcg(&st->rows_streamed) + (long long)nlive + 1;
            const int next_status = la_finalize(la, nxt, seq + 1u, H, true);
            st->status = next_status;
            la->word = 0ull;
            la->tile_done = 0u;
            la->stamps[6] = globaltimer();
            if (PERSIST) {
                st_release_s64(&st->pivots, pivots + 1);   // the barrier's release: everything above is visible with it
            } else {
                st->pivots = pivots + 1;
                __threadfence();
            }
        }
    }
    if (PERSIST) {
        // grid barrier: wait until this pivot is committed (bounded: a CTA that never arrives must not hang the GPU)
        if (threadIdx.x == 0) {
            const long long t0 = clock64();
            int ok = 1;
            while (ld_acquire_s64(&st->pivots) <= pivots) {
                if (clock64() - t0 > P.wait_cycles) {
                    ok = 0;
                    break;
                }
            }
            sh.ok = ok;
        }
        __syncthreads();
        const bool ok = sh.ok != 0;
        __syncthreads();
        if (!ok) {
            if (threadIdx.x == 0) {
                st->status = kStatusPeerTimeout;
                __threadfence();
            }
            return;
        }
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
    la_chain<real, false, false>(P, la, seq, (int)blockIdx.x, (int)gridDim.x, nullptr, nullptr, (real)0, -1, -1, -1, false, 0, sm, smax, sh);
}

// The prologue's verdicts (unbounded at the first pivot, a silent peer) have no running pivot to ride on.
template <typename real>
__global__ void la_prologue_commit_kernel(PivotParams<real> P)
{
    DevState* st = P.st;
    if (__ldcg(&st->status) != kRunning || __ldcg(&st->pivots) >= __ldcg(&st->limit)) return;
    const unsigned tseq = (unsigned)(__ldcg(&st->pivots) + 1);
    Proposal* nxt = &P.la->prop[tseq & 1u];
    if (__ldcg(&nxt->ready_seq) == tseq) return;
    const int sn = la_finalize(P.la, nxt, tseq, P.helpers, false);
    if (sn != kRunning) st->status = sn;
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
