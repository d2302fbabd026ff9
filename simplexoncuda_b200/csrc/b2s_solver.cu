// b2s_solver.cu -- host driver of the B200 two-phase simplex and the C ABI (include/b2s.h).
//
// The host never takes part in a pivot: it enqueues batches of {ratio, gather, update} launches
// (optionally replayed as one CUDA graph) and polls the device-resident DevState between
// batches.  The reference's host-driven loop is src/solver.cu:128-149 / src/twoPhaseMethod.cu:385-435.
#include "../../include/b2s.h"
#include "b2s_generator.cuh"
#include "b2s_kernels.cuh"
#include "b2s_p2p.cuh"
#include "b2s_persistent.cuh"
#include "b2s_lookahead.cuh"
#include "b2s_bulk.cuh"

#include <algorithm>
#include <chrono>
#include <climits>
#include <cstddef>
#include <cmath>
#include <map>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#ifdef B2S_WITH_NCCL
#include <nccl.h>
#endif

namespace b2s {

static thread_local std::string g_thread_error;

struct SolverBase {
    b2s_options opt{};
    std::string err;
    virtual ~SolverBase() {}
    virtual int load_host(int n, int m, const double* A, const double* b, const double* c) = 0;
    virtual int generate(int n, int m, const unsigned seeds[3], double lo, double hi) = 0;
    virtual int copy_problem(double* A, double* b, double* c) = 0;
    virtual int build_phase1() = 0;
    virtual int price_out() = 0;
    virtual int select_entering() = 0;
    virtual int iterate(long long max_pivots, int* status, long long* done) = 0;
    virtual int phase1_verdict(int* status) = 0;
    virtual int switch_phase2() = 0;
    virtual int extract(double* x, double* obj) = 0;
    virtual int solve(int* status, double* x, double* obj, int* basis, b2s_stats* stats) = 0;
    virtual int get_dims(int* n, int* m, long long* ra, long long* rs, long long* ld) const = 0;
    virtual int copy_tableau(double* out) = 0;
    virtual int copy_costs(double* out) = 0;
    virtual int copy_basis(int* out) = 0;
    virtual int copy_trace(int* qp, long long cap, long long* len, unsigned long long* hash) = 0;
    virtual int get_stats(b2s_stats* st) = 0;
    virtual int tournament(const double* vec, long long n, double* value, int* index) = 0;
    virtual int attach(double* table, size_t pitch, int rows, int cols, double* costs, int n_vars) = 0;
    virtual int set_basis(const int* base_host) = 0;
    virtual int min_element_device(const double* dvec, long long n, double* value, unsigned* index) = 0;
    virtual int ratio_min_device(const double* known, const double* column, long long n, double* value, unsigned* index) = 0;
    virtual int max_le_zero_device(const double* dvec, long long n, int* result) = 0;
    virtual int bench_update(int launches, int flush, float* ms, double* bytes) = 0;
    virtual int dist_init(int rank, int world, const char* id) = 0;
    virtual int dist_init_host(int rank, int world, b2s_allgather_fn fn, void* user) = 0;
    virtual int profile_pivots(int count, float* ms_ratio, float* ms_gather, float* ms_update, long long* done) = 0;
    virtual int profile_lookahead(int count, float* kernel_ms, double* stage_us, long long* done) = 0;
    virtual int loop_info(int* launches, int* lookahead, int* persistent) = 0;

    int fail(int code, const char* fmt, ...)
    {
        char buf[512];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof(buf), fmt, ap);
        va_end(ap);
        err = buf;
        g_thread_error = buf;
        return code;
    }
};

#define CK(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess)                                                                           \
            return fail(e_ == cudaErrorMemoryAllocation ? B2S_ERR_NOMEM : B2S_ERR_CUDA, "%s:%d %s: %s", \
                        __FILE__, __LINE__, #call, cudaGetErrorString(e_));                              \
    } while (0)

#ifdef B2S_WITH_NCCL
#define NK(call)                                                                                      \
    do {                                                                                              \
        ncclResult_t r_ = (call);                                                                     \
        if (r_ != ncclSuccess)                                                                        \
            return fail(B2S_ERR_NCCL, "%s:%d %s: %s", __FILE__, __LINE__, #call, ncclGetErrorString(r_)); \
    } while (0)
#endif

static double now_s()
{
    using namespace std::chrono;
    return duration<double>(steady_clock::now().time_since_epoch()).count();
}

enum Stage { kEmpty = 0, kLoaded, kBuilt, kPriced, kReady, kPhaseDone };

// Scratch device allocation of the kernel-level hooks: released on every return path.
template <typename X>
struct DevBuf {
    X* p = nullptr;
    cudaError_t alloc(size_t count) { return cudaMalloc(&p, sizeof(X) * (count ? count : 1)); }
    ~DevBuf() { cudaFree(p); }
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
};

template <typename real>
struct SolverImpl final : SolverBase {
    int dev = 0;
    int num_sms = 148;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int n = 0, m = 0, m_loc = 0, col0 = 0;
    long long ld = 0, Rs = 0, R1 = 0, Rc = 0;
    size_t cap_T = 0, cap_rows = 0, cap_cols = 0, cap_n = 0, cap_m = 0;  // allocated capacities
    int phase = 0;
    Stage stage = kEmpty;
    bool folded = true;
    bool attached = false;  // T / cost alias a caller-owned tabular_t (reference layout, no folding)

    real *T_own = nullptr, *cost_own = nullptr;  // allocations; T / cost may instead alias a caller's tabular_t
    real *T = nullptr, *cost = nullptr, *col = nullptr, *s = nullptr, *rowp = nullptr, *coef = nullptr, *c_dev = nullptr;
    real *rslot_v = nullptr, *rslot_max = nullptr, *cslot_v = nullptr;
    int *rslot_i = nullptr, *rslot_k = nullptr, *cslot_i = nullptr, *cslot_k = nullptr;
    int *base = nullptr, *neg = nullptr, *verdict = nullptr;
    double* x_dev = nullptr;
    double* stage_dev = nullptr;  // fp64 staging for fp32 loads / exports
    double* orig64 = nullptr;     // fp32 solves: the problem in fp64 ([b | A | c]) for the final polish
    size_t cap_orig = 0;
    bool have_orig = false;
    double polish_residual = -1.0; // last polish: max |dx| / max |x| of the final refinement step
    size_t cap_stage = 0;
    DevState* st = nullptr;
    DevState* st_host = nullptr;  // pinned, [0] = current snapshot, [1..2] = pipelined poll slots
    cudaEvent_t poll_ev[2] = {nullptr, nullptr};
    int2* trace = nullptr;
    long long trace_cap = 0;
    uint32_t* jump_tables = nullptr;
    void* flush_buf = nullptr;
    size_t flush_bytes = 0;

    // look-ahead pivot kernel (b2s_lookahead.cuh)
    LaState* la = nullptr;
    unsigned* tile_rec = nullptr;
    size_t cap_tiles = 0;
    real *col2 = nullptr, *s2 = nullptr, *rowp2 = nullptr, *rowval = nullptr;
    int *rowlist = nullptr, *rowpos = nullptr;
    int la_grid = 0;
    int la_u = 8;          // 256-bit loads in flight per streaming thread (set per problem in fill_params)
    int la_u_env = 0;      // B2S_LA_U: 0 = choose
    bool la_persist = false;  // B2S_LA_PERSIST=1
    bool la_pdl = false;  // programmatic dependent launch between consecutive pivots: measured, no gain (profiles/r02_lookahead.md)
    int la_helpers = 0;   // 0 = default: 8 on one GPU (chain hidden anyway), 16 when sharded (the chain is the critical path)
    long long wait_cycles = 4000000000ll;
    int fault_rank = -1;
    long long fault_pivot = 0;

    PivotParams<real> P{};
    int upd_grid = 0;
    size_t upd_smem = 0;  // dynamic shared memory of the update kernel (bulk-copy variant only)
    int loop_grid = 0;  // persistent loop kernel: co-resident CTAs
    std::map<int, cudaGraphExec_t> graphs;  // captured batches of pivots, keyed by batch length
    long long pivots_p1 = 0, pivots_p2 = 0;
    double cost0_phase1_start = 0.0;  // relative_infeasibility: cost[0] right after the phase-1 price-out
    double sec_load = 0, sec_p1 = 0, sec_p2 = 0;

    // sharding
    int rank = 0, world = 1;
#ifdef B2S_WITH_NCCL
    ncclComm_t comm = nullptr;
#endif
    // Host-layer collectives instead of NCCL (b2s_dist_init_host): the caller's all-gather moves the few bytes the slow paths
    // need (IPC handles, barriers, the price-out chain, the solution); the per-pivot exchanges are peer-memory kernels anyway.
    b2s_allgather_fn host_ag = nullptr;
    void* host_ag_user = nullptr;
    // peer-memory exchange (b2s_p2p.cuh)
    bool use_pdl = false;          // single-GPU launches path: programmatic dependent launch between the 3 kernels
    bool p2p = false;              // arenas mapped on every rank: the per-pivot exchanges bypass NCCL
    unsigned char* arena = nullptr;
    long long arena_rows = 0;
    unsigned char* peer_ptr[kMaxPeers] = {};

    ~SolverImpl() override { release(); }

    void release()
    {
        cudaSetDevice(dev);
        invalidate_graph();
        free_problem();
        cudaFree(st);
        cudaFree(la);
        la = nullptr;
        cudaFreeHost(st_host);
        cudaFree(trace);
        cudaFree(jump_tables);
        cudaFree(flush_buf);
        cudaFree(verdict);
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
        for (auto& e : poll_ev) {
            if (e) cudaEventDestroy(e);
            e = nullptr;
        }
        if (stream) cudaStreamDestroy(stream);
        close_arena();
#ifdef B2S_WITH_NCCL
        if (comm) ncclCommDestroy(comm);
        comm = nullptr;
#endif
        st = nullptr;
        st_host = nullptr;
        trace = nullptr;
        jump_tables = nullptr;
        flush_buf = nullptr;
        verdict = nullptr;
        ev0 = ev1 = nullptr;
        stream = nullptr;
    }

    void free_problem()
    {
        cudaFree(T_own);
        cudaFree(cost_own);
        T_own = cost_own = nullptr;
        cudaFree(col);
        cudaFree(s);
        cudaFree(rowp);
        cudaFree(coef);
        cudaFree(c_dev);
        cudaFree(rslot_v);
        cudaFree(rslot_max);
        cudaFree(cslot_v);
        cudaFree(rslot_i);
        cudaFree(rslot_k);
        cudaFree(cslot_i);
        cudaFree(cslot_k);
        cudaFree(base);
        cudaFree(neg);
        cudaFree(x_dev);
        cudaFree(stage_dev);
        cudaFree(orig64);
        orig64 = nullptr;
        cap_orig = 0;
        have_orig = false;
        cudaFree(tile_rec);
        cudaFree(col2);
        cudaFree(s2);
        cudaFree(rowp2);
        cudaFree(rowval);
        cudaFree(rowlist);
        cudaFree(rowpos);
        tile_rec = nullptr;
        col2 = s2 = rowp2 = rowval = nullptr;
        rowlist = rowpos = nullptr;
        cap_tiles = 0;
        T = cost = col = s = rowp = coef = c_dev = rslot_v = rslot_max = cslot_v = nullptr;
        rslot_i = rslot_k = cslot_i = cslot_k = base = neg = nullptr;
        x_dev = stage_dev = nullptr;
        cap_T = cap_rows = cap_cols = cap_n = cap_m = cap_stage = 0;
        stage = kEmpty;
    }

    int init()
    {
        int count = 0;
        if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0)
            return fail(B2S_ERR_NOGPU, "no CUDA device visible: this library has no CPU fallback");
        dev = opt.device;
        if (dev < 0 || dev >= count) return fail(B2S_ERR_ARG, "device %d out of range (%d visible)", dev, count);
        CK(cudaSetDevice(dev));
        CK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
        CK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        CK(cudaEventCreate(&ev0));
        CK(cudaEventCreate(&ev1));
        CK(cudaMalloc(&st, sizeof(DevState)));
        CK(cudaMemsetAsync(st, 0, sizeof(DevState), stream));  // `stream` is non-blocking: never mix in legacy-stream calls
        CK(cudaMalloc(&la, sizeof(LaState)));
        CK(cudaMemsetAsync(la, 0, sizeof(LaState), stream));
        if (const char* e = getenv("B2S_LA_U")) la_u_env = atoi(e) == 4 ? 4 : 8;
        if (const char* e = getenv("B2S_LA_PDL")) la_pdl = atoi(e) != 0;
        if (const char* e = getenv("B2S_LA_PERSIST")) la_persist = atoi(e) != 0;
        if (const char* e = getenv("B2S_LA_HELPERS")) la_helpers = std::max(1, std::min(kLaMaxHelpers, atoi(e)));
        if (const char* e = getenv("B2S_PEER_TIMEOUT_MS")) wait_cycles = std::max(1ll, atoll(e)) * 2000000ll;  // ~2 GHz
        if (const char* e = getenv("B2S_FAULT_RANK")) fault_rank = atoi(e);
        if (const char* e = getenv("B2S_FAULT_PIVOT")) fault_pivot = atoll(e);
        CK(cudaHostAlloc(&st_host, 3 * sizeof(DevState), cudaHostAllocDefault));
        CK(cudaEventCreateWithFlags(&poll_ev[0], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&poll_ev[1], cudaEventDisableTiming));
        CK(cudaMalloc(&verdict, 2 * sizeof(int)));
        trace_cap = opt.trace_capacity > 0 ? opt.trace_capacity : (1ll << 20);
        CK(cudaMalloc(&trace, sizeof(int2) * (size_t)trace_cap));
        folded = opt.fold_artificials != 0;
        if (const char* e = getenv("B2S_PDL")) use_pdl = atoi(e) != 0;
        return B2S_OK;
    }

    template <typename X>
    int dmalloc(X** p, size_t count)
    {
        CK(cudaMalloc(p, sizeof(X) * std::max<size_t>(count, 1)));
        return B2S_OK;
    }

    void close_arena()
    {
        for (int r = 0; r < kMaxPeers; ++r) {
            if (peer_ptr[r] && peer_ptr[r] != arena) cudaIpcCloseMemHandle(peer_ptr[r]);
            peer_ptr[r] = nullptr;
        }
        cudaFree(arena);
        arena = nullptr;
        arena_rows = 0;
        p2p = false;
    }

    bool have_collectives() const
    {
#ifdef B2S_WITH_NCCL
        if (comm) return true;
#endif
        return host_ag != nullptr;
    }
    // every rank contributes `bytes` bytes; recv holds world * bytes (host memory)
    int host_allgather(const void* send, void* recv, size_t bytes)
    {
        if (!host_ag) return fail(B2S_ERR_STATE, "no host all-gather registered");
        if (host_ag(host_ag_user, send, recv, bytes) != 0) return fail(B2S_ERR_NCCL, "host all-gather callback failed");
        return B2S_OK;
    }
    // all ranks have executed everything enqueued so far
    int rank_barrier()
    {
        if (world <= 1) return B2S_OK;
#ifdef B2S_WITH_NCCL
        if (comm) {
            NK(ncclAllReduce(verdict, verdict, 1, ncclInt, ncclSum, comm, stream));
            return B2S_OK;
        }
#endif
        CK(cudaStreamSynchronize(stream));
        std::vector<int> box((size_t)world);
        const int mine = rank;
        return host_allgather(&mine, box.data(), sizeof(int));
    }

    // Sharded solves: (re)create this rank's arena for `rows` pivot-constraint entries and map every
    // peer's arena (CUDA IPC handles travel through one ncclAllGather).  Collective over the ranks.
    int ensure_arena(long long rows)
    {
        if (world <= 1) return B2S_OK;
        const char* env = getenv("B2S_P2P");
        if ((env && atoi(env) == 0) || world > kMaxPeers) {
            if (host_ag) return fail(B2S_ERR_STATE, "a solver bootstrapped through the host layer needs the peer-memory exchanges (B2S_P2P=0 "
                                                    "and more than %d ranks are NCCL-only)", kMaxPeers);
            close_arena();
            return B2S_OK;
        }
        if (arena && rows <= arena_rows) return B2S_OK;
        close_arena();
        const long long cap = (rows + 1023) / 1024 * 1024;
        CK(cudaMalloc(&arena, arena_bytes<real>(cap)));
        CK(cudaMemsetAsync(arena, 0, arena_bytes<real>(cap), stream));
        cudaIpcMemHandle_t mine;
        CK(cudaIpcGetMemHandle(&mine, arena));
        std::vector<cudaIpcMemHandle_t> all((size_t)world);
        bool gathered = false;
#ifdef B2S_WITH_NCCL
        if (comm) {
            unsigned char* hbuf = nullptr;
            CK(cudaMalloc(&hbuf, sizeof(mine) * (size_t)world));
            CK(cudaMemcpyAsync(hbuf + sizeof(mine) * (size_t)rank, &mine, sizeof(mine), cudaMemcpyHostToDevice, stream));
            NK(ncclAllGather(hbuf + sizeof(mine) * (size_t)rank, hbuf, sizeof(mine), ncclChar, comm, stream));
            CK(cudaStreamSynchronize(stream));
            CK(cudaMemcpyAsync(all.data(), hbuf, sizeof(mine) * (size_t)world, cudaMemcpyDeviceToHost, stream));
            CK(cudaStreamSynchronize(stream));
            cudaFree(hbuf);
            gathered = true;
        }
#endif
        if (!gathered) {
            CK(cudaStreamSynchronize(stream));   // the arena is zeroed before anybody can map and write it
            int rc = host_allgather(&mine, all.data(), sizeof(mine));
            if (rc) return rc;
        }
        for (int r = 0; r < world; ++r) {
            if (r == rank) {
                peer_ptr[r] = arena;
            } else {
                void* ptr = nullptr;
                CK(cudaIpcOpenMemHandle(&ptr, all[(size_t)r], cudaIpcMemLazyEnablePeerAccess));
                peer_ptr[r] = (unsigned char*)ptr;
            }
        }
        arena_rows = cap;
        p2p = true;
        return B2S_OK;
    }

    // (Re)allocate for an n x m problem and reset the solver state.
    int prepare(int n_, int m_)
    {
        if (n_ < 1 || m_ < 1) return fail(B2S_ERR_ARG, "need n >= 1 and m >= 1 (got %d, %d)", n_, m_);
        CK(cudaSetDevice(dev));
        if (world > 1 && (m_ % (world * kSelBlock)) != 0)
            return fail(B2S_ERR_ARG, "sharded solve needs constraints %% (world*512) == 0 (m=%d, world=%d)", m_, world);
        if (world > 1 && (long long)m_ > (long long)kSelBlock * kMaxSlots)
            return fail(B2S_ERR_ARG, "sharded solve supports at most %d constraints", kSelBlock * kMaxSlots);
        folded = opt.fold_artificials != 0;  // attach() (caller-owned tabular_t) switches folding off for its own use
        n = n_;
        m = m_;
        m_loc = m / world;
        col0 = rank * m_loc;
        ld = ((long long)m_loc + 63) / 64 * 64;
        R1 = 1 + (long long)n + 2ll * m;
        Rs = folded ? 1 + (long long)n + m : R1;
        Rc = R1;
        phase = 0;
        const size_t needT = (size_t)Rs * (size_t)ld;
        if (needT > cap_T || (size_t)R1 > cap_rows || (size_t)ld > cap_cols || (size_t)n > cap_n || (size_t)m > cap_m) {
            free_problem();
            int rc;
            if ((rc = dmalloc(&T_own, needT))) return rc;
            if ((rc = dmalloc(&cost_own, (size_t)R1))) return rc;
            if ((rc = dmalloc(&rowp, (size_t)R1))) return rc;
            if ((rc = dmalloc(&col, (size_t)ld))) return rc;
            if ((rc = dmalloc(&s, (size_t)ld))) return rc;
            if ((rc = dmalloc(&coef, (size_t)ld))) return rc;
            if ((rc = dmalloc(&neg, (size_t)ld))) return rc;
            if ((rc = dmalloc(&c_dev, (size_t)n))) return rc;
            if ((rc = dmalloc(&x_dev, (size_t)n))) return rc;
            if ((rc = dmalloc(&base, (size_t)m))) return rc;
            if ((rc = dmalloc(&rslot_v, kMaxSlots))) return rc;
            if ((rc = dmalloc(&rslot_max, kMaxSlots))) return rc;
            if ((rc = dmalloc(&rslot_i, kMaxSlots))) return rc;
            if ((rc = dmalloc(&rslot_k, kMaxSlots))) return rc;
            if ((rc = dmalloc(&cslot_v, kMaxSlots))) return rc;
            if ((rc = dmalloc(&cslot_i, kMaxSlots))) return rc;
            if ((rc = dmalloc(&cslot_k, kMaxSlots))) return rc;
            if ((rc = dmalloc(&col2, 2 * (size_t)ld))) return rc;
            if ((rc = dmalloc(&s2, 2 * (size_t)ld))) return rc;
            if ((rc = dmalloc(&rowp2, 2 * (size_t)R1))) return rc;
            if ((rc = dmalloc(&rowval, 2 * (size_t)R1))) return rc;
            if ((rc = dmalloc(&rowlist, 2 * (size_t)R1))) return rc;
            if ((rc = dmalloc(&rowpos, 2 * (size_t)R1))) return rc;
            cap_T = needT;
            cap_rows = (size_t)R1;
            cap_cols = (size_t)ld;
            cap_n = (size_t)n;
            cap_m = (size_t)m;
        }
        T = T_own;
        cost = cost_own;
        attached = false;
        {
            int rc = ensure_arena(R1);
            if (rc) return rc;
        }
        CK(cudaMemsetAsync(st, 0, sizeof(DevState), stream));
        pivots_p1 = pivots_p2 = 0;
        sec_load = sec_p1 = sec_p2 = 0;
        invalidate_graph();
        fill_params();
        return B2S_OK;
    }

    void invalidate_graph()
    {
        for (auto& kv : graphs) cudaGraphExecDestroy(kv.second);
        graphs.clear();
    }

    static int ceil_log2(long long v)
    {
        int l = 0;
        while ((1ll << l) < v) ++l;
        return l;
    }

    // ---- update-kernel variants --------------------------------------------------------------
    struct Variant {
        int vb, u, hint, dyn;
    };
    static constexpr int kNumVariants = 15;  // 14 = bulk-copy (TMA engine) pipeline, b2s_bulk.cuh
    Variant variant() const
    {
        static const Variant table[kNumVariants] = {{16, 8, 0, 0}, {16, 8, 1, 0}, {32, 4, 0, 0}, {32, 4, 1, 0}, {32, 8, 0, 0},
                                                    {16, 4, 0, 0}, {16, 8, 2, 0}, {32, 8, 1, 0}, {32, 8, 0, 1}, {32, 4, 0, 1},
                                                    {16, 4, 0, 1}, {32, 8, 1, 1}, {32, 8, 2, 1}, {16, 8, 0, 1}, {32, 8, 0, 1}};
        return table[variant_index()];
    }
    // One index for the tiling (fill_params) and for the kernel (update_fn): a mismatch would leave columns uncovered.
    int variant_index() const
    {
        int v = opt.update_variant;
        if (v < 0 || v >= kNumVariants) v = 8;
        if (use_persistent()) v = 8;  // the loop kernel is built for the 256-bit / 8-row / ticketed geometry
        return v;
    }
    // persistent: 0 = three launches per pivot, 1 = loop kernel, 2 = auto.  Measured on B200 on complete solves
    // (profiles/r01_scaling_and_loop_modes.md): the loop kernel wins for small tableaux (47.6k vs 43.2k pivots/s
    // at 16 MB) and loses from ~30 MB upwards (37.6k vs 40.4k at 50 MB, 22.9k vs 27.6k at 100 MB, -2 % at 1 GB),
    // where launches replayed from a CUDA graph cost about as much as grid barriers and the stand-alone update
    // kernel streams a little faster than the loop kernel's phase D.
    bool use_persistent() const
    {
        if (opt.persistent == 0 || (world > 1 && !p2p)) return false;
        if (opt.persistent == 1) return true;
        return world == 1 && (double)Rs * (double)ld * sizeof(real) < 32e6;
    }
    // lookahead: 0 = three launches per pivot (ratio, gather, update), 1/2 = ONE launch per pivot whose helper CTAs prepare
    // the next pivot under the streaming update (b2s_lookahead.cuh) whenever the geometry allows it.
    bool use_lookahead() const
    {
        int mode = opt.lookahead;
        if (const char* e = getenv("B2S_LOOKAHEAD")) mode = atoi(e);
        if (mode == 0 || use_persistent() || variant_index() != 8 || !tile_rec) return false;
        if (world > 1 && !p2p) return false;
        // auto: the chain takes ~30 us even on an idle memory system (profiles/r02_lookahead.md); on one GPU it only pays once a
        // pivot streams longer than that, i.e. from ~280 MB of stored tableau (measured crossover, r02_loop_mode_sweep).  Sharded
        // solves always gain: the chain replaces two exposed NVLink exchanges and three extra launches.
        if (mode == 2 && world == 1 && (double)Rs * (double)ld * sizeof(real) < 280e6) return false;
        if ((long long)m > (long long)kSelBlock * kMaxSlots) return false;          // one ratio element per helper thread
        if (R1 >= (long long)kRowMask - 1 || ld >= (long long)kNoColumn - 1) return false;  // ticket-word fields
        return true;
    }
    typedef void (*LaFn)(const PivotParams<real>, int);
    LaFn la_fn() const { return la_u == 4 ? (LaFn)update_la_kernel<real, 4, false> : (LaFn)update_la_kernel<real, 8, false>; }
    LaFn la_persist_fn() const { return (LaFn)update_la_kernel<real, 8, true>; }
    // several pivots per cooperative launch of the look-ahead kernel (grid barrier instead of a kernel boundary)
    bool use_la_persist() const { return la_persist && la_u == 8 && use_lookahead(); }
    typedef void (*LoopFn)(PivotParams<real>, int);
    LoopFn loop_fn() const
    {
        return opt.skip_zero_rows ? (LoopFn)pivot_loop_kernel<real, 32, 8, true> : (LoopFn)pivot_loop_kernel<real, 32, 8, false>;
    }
    typedef void (*UpdateFn)(PivotParams<real>);
    template <int VB, int U, int HINT, bool DYN>
    UpdateFn pick_skip() const
    {
        return opt.skip_zero_rows ? (UpdateFn)update_kernel<real, VB, U, HINT, true, DYN>
                                  : (UpdateFn)update_kernel<real, VB, U, HINT, false, DYN>;
    }
    UpdateFn update_fn() const
    {
        const int v = variant_index();
        switch (v) {
            case 0: return pick_skip<16, 8, 0, false>();
            case 1: return pick_skip<16, 8, 1, false>();
            case 2: return pick_skip<32, 4, 0, false>();
            case 3: return pick_skip<32, 4, 1, false>();
            case 4: return pick_skip<32, 8, 0, false>();
            case 5: return pick_skip<16, 4, 0, false>();
            case 6: return pick_skip<16, 8, 2, false>();
            case 7: return pick_skip<32, 8, 1, false>();
            default:
            case 8: return pick_skip<32, 8, 0, true>();
            case 9: return pick_skip<32, 4, 0, true>();
            case 10: return pick_skip<16, 4, 0, true>();
            case 11: return pick_skip<32, 8, 1, true>();
            case 12: return pick_skip<32, 8, 2, true>();
            case 13: return pick_skip<16, 8, 0, true>();
            case 14: return opt.skip_zero_rows ? pick_skip<32, 8, 0, true>() : (UpdateFn)update_bulk_kernel<real>;
        }
    }

    void fill_params()
    {
        P.T = T;
        P.ld = ld;
        P.n = n;
        P.m = m;
        P.m_loc = m_loc;
        P.col0 = col0;
        P.Rs = Rs;
        P.Rc = Rc;
        P.fold_from = folded ? 1 + (long long)n + m : LLONG_MAX;
        P.cost = cost;
        P.base = base;
        P.col = col;
        P.s = s;
        P.rowp = rowp;
        P.rslot_v = rslot_v;
        P.rslot_i = rslot_i;
        P.rslot_k = rslot_k;
        P.rslot_max = rslot_max;
        P.cslot_v = cslot_v;
        P.cslot_i = cslot_i;
        P.cslot_k = cslot_k;
        P.st = st;
        P.trace = trace;
        P.trace_cap = trace_cap;
        P.rule = opt.pivot_rule;
        P.skip_zero = opt.skip_zero_rows;
        P.Gm = (int)std::min<long long>(((long long)m + kSelBlock - 1) / kSelBlock, kMaxSlots);
        P.Gm_loc0 = col0 / kSelBlock;
        P.Gm_loc = world > 1 ? m_loc / kSelBlock : P.Gm;
        P.Gc = (int)std::max<long long>(1, std::min<long long>((Rc - 1 + kSelBlock - 1) / kSelBlock, kMaxSlots));
        P.serpentine = 1;
        if (const char* e = getenv("B2S_SERPENTINE")) P.serpentine = atoi(e) != 0;
        P.rank = rank;
        P.world = world;
        P.arena_rows = arena_rows;
        for (int r = 0; r < kMaxPeers; ++r) P.peers[r] = peer_ptr[r];
        // update-kernel tiling
        const Variant v = variant();
        const int ept = v.vb / (int)sizeof(real);
        const long long thr_row = (ld + ept - 1) / ept;  // threads needed per row
        const int l2 = std::min(9, ceil_log2(thr_row));
        P.log2_tpr = l2;
        const long long chunk_cols = (long long)ept << l2;
        P.nchunks = (int)((ld + chunk_cols - 1) / chunk_cols);
        const int rpp = kSelBlock >> l2;
        int occ = 1;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, update_fn(), kSelBlock, 0);
        occ = std::max(1, occ);
        int grid = num_sms * occ;
        if (!v.dyn && P.nchunks <= grid) grid = grid / P.nchunks * P.nchunks;  // static striding keeps a CTA on one chunk
        int tg = 4;
        for (; tg > 1; tg >>= 1) {
            const long long rows_tile = (long long)rpp * v.u * tg;
            const long long nt = ((Rs + rows_tile - 1) / rows_tile) * P.nchunks;
            if (nt >= (v.dyn ? 16ll : 4ll) * grid) break;
        }
        if (v.dyn) tg = 1;  // finest tiles: measured best with the ticket scheduler (profiles/r01_kernel_sweep.md)
        if (const char* e = getenv("B2S_TILE_GROUPS")) tg = std::max(1, atoi(e));
        P.tile_groups = tg;
        const long long rows_tile = (long long)rpp * v.u * tg;
        P.ntiles = ((Rs + rows_tile - 1) / rows_tile) * P.nchunks;
        upd_grid = (int)std::max<long long>(std::min<long long>(grid, P.ntiles), std::min(P.Gc, grid));
        upd_grid = std::max(upd_grid, 1);
        upd_smem = 0;
        if (update_fn() == (UpdateFn)update_bulk_kernel<real>) {
            upd_smem = (size_t)kBulkStages * kBulkRows * kBulkCols * 8;
            cudaFuncSetAttribute(update_bulk_kernel<real>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)upd_smem);
            upd_grid = num_sms;
        }
        // look-ahead kernel: same tile geometry as variant 8; tiles are cut from the per-pivot row list, so size by the maximum
        {
            // Sharded slabs that stream in less time than the chain takes (< ~350 MB): half the bytes in flight per thread halves
            // the queueing delay every dependent access of the chain sees, and the lost streaming bandwidth is hidden
            // (N=4: 57.0 vs 61.4 us per pivot, N=8: 50.4 vs 52.2; one GPU / large slabs: 8 -- profiles/r02_lookahead.md).
            la_u = la_u_env ? la_u_env : ((world > 1 && (double)Rs * (double)ld * sizeof(real) < 350e6) ? 4 : 8);
            const long long rows_tile8 = (long long)rpp * la_u;
            const long long max_tiles = ((R1 + rows_tile8 - 1) / rows_tile8 + 1) * P.nchunks;
            P.la_u = la_u;
            if ((size_t)max_tiles > cap_tiles) {
                cudaFree(tile_rec);
                tile_rec = nullptr;
                cap_tiles = 0;
                if (cudaMalloc(&tile_rec, sizeof(unsigned) * (size_t)max_tiles) == cudaSuccess) {
                    cap_tiles = (size_t)max_tiles;
                    cudaMemsetAsync(tile_rec, 0, sizeof(unsigned) * cap_tiles, stream);
                }
            }
            const int want = la_helpers > 0 ? la_helpers : (world > 1 ? 16 : 8);   // measured: profiles/r02_lookahead.md   // measured: profiles/r02_lookahead.md
            la_grid = (int)std::max<long long>(1, std::min<long long>(num_sms, std::max<long long>(max_tiles + want, P.Gc)));
            P.helpers = std::min(want, la_grid);
            P.la = la;
            P.tile_rec = tile_rec;
            P.col2 = col2;
            P.s2 = s2;
            P.rowp2 = rowp2;
            P.rowlist = rowlist;
            P.rowval = rowval;
            P.rowpos = rowpos;
            P.rowp_stride = (long long)cap_rows;
            P.wait_cycles = wait_cycles;
            P.fault_rank = fault_rank;
            P.fault_pivot = fault_pivot;
        }
        if (opt.persistent) {
            int occ_l = 0;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_l, loop_fn(), kSelBlock, 0);
            loop_grid = num_sms * std::max(1, std::min(occ_l, 1));
        }
    }

    // ---- problem input -------------------------------------------------------------------------
    int zero_fill_for_load()
    {
        // slack (and artificial) rows entirely; padding columns of the data rows.
        const long long data_rows = 1 + (long long)n;
        CK(cudaMemsetAsync(T + data_rows * ld, 0, sizeof(real) * (size_t)(Rs - data_rows) * (size_t)ld, stream));
        if (ld > m_loc)
            CK(cudaMemset2DAsync(T + m_loc, sizeof(real) * ld, 0, sizeof(real) * (ld - m_loc), (size_t)data_rows, stream));
        return B2S_OK;
    }

    int ensure_stage(size_t count)
    {
        if (count > cap_stage) {
            cudaFree(stage_dev);
            stage_dev = nullptr;
            cap_stage = 0;
            CK(cudaMalloc(&stage_dev, sizeof(double) * count));
            cap_stage = count;
        }
        return B2S_OK;
    }

    int load_host(int n_, int m_, const double* A, const double* b, const double* c) override
    {
        if (!A || !b || !c) return fail(B2S_ERR_ARG, "null problem array");
        const double t0 = now_s();
        int rc = prepare(n_, m_);
        if (rc) return rc;
        if ((rc = zero_fill_for_load())) return rc;
        if (sizeof(real) == sizeof(double)) {
            // rows 1..n <- A (variable-major, row pitch m), local slab columns only; row 0 <- b
            CK(cudaMemcpy2DAsync(T + ld, sizeof(real) * ld, A + col0, sizeof(double) * m, sizeof(double) * m_loc,
                                 (size_t)n, cudaMemcpyHostToDevice, stream));
            CK(cudaMemcpyAsync(T, b + col0, sizeof(double) * m_loc, cudaMemcpyHostToDevice, stream));
            CK(cudaMemcpyAsync(c_dev, c, sizeof(double) * n, cudaMemcpyHostToDevice, stream));
        } else {
            if ((rc = ensure_stage((size_t)(n + 1) * (size_t)m_loc + (size_t)n))) return rc;
            CK(cudaMemcpy2DAsync(stage_dev + m_loc, sizeof(double) * m_loc, A + col0, sizeof(double) * m,
                                 sizeof(double) * m_loc, (size_t)n, cudaMemcpyHostToDevice, stream));
            CK(cudaMemcpyAsync(stage_dev, b + col0, sizeof(double) * m_loc, cudaMemcpyHostToDevice, stream));
            double* cst = stage_dev + (size_t)(n + 1) * m_loc;
            CK(cudaMemcpyAsync(cst, c, sizeof(double) * n, cudaMemcpyHostToDevice, stream));
            convert_rows<<<1024, 256, 0, stream>>>(T, ld, stage_dev, (long long)m_loc, (long long)(n + 1), m_loc);
            convert_rows<<<64, 256, 0, stream>>>(c_dev, (long long)n, cst, (long long)n, 1ll, n);
            if ((rc = keep_original_from_stage())) return rc;
        }
        CK(cudaStreamSynchronize(stream));
        stage = kLoaded;
        sec_load = now_s() - t0;
        return B2S_OK;
    }

    int ensure_jump_tables()
    {
        if (jump_tables) return B2S_OK;
        std::vector<uint32_t> host;
        xorwow_build_jump_tables(host);
        CK(cudaMalloc(&jump_tables, host.size() * sizeof(uint32_t)));
        CK(cudaMemcpyAsync(jump_tables, host.data(), host.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, stream));
        CK(cudaStreamSynchronize(stream));  // `host` goes out of scope
        return B2S_OK;
    }

    int generate(int n_, int m_, const unsigned seeds[3], double lo, double hi) override
    {
        const double t0 = now_s();
        int rc = prepare(n_, m_);
        if (rc) return rc;
        if ((rc = ensure_jump_tables())) return rc;
        if ((rc = zero_fill_for_load())) return rc;
        const double span = hi - lo;
        {  // b -> row 0 (local slab): element id = global constraint index
            const long long nthreads = ((long long)m_loc + kVecRun - 1) / kVecRun;
            generate_vector_kernel<real><<<(unsigned)((nthreads + 255) / 256), 256, 0, stream>>>(
                T, (long long)m_loc, (long long)col0, seeds[0], lo, span, jump_tables);
        }
        {  // c (replicated on every rank)
            const long long nthreads = ((long long)n + kVecRun - 1) / kVecRun;
            generate_vector_kernel<real><<<(unsigned)((nthreads + 255) / 256), 256, 0, stream>>>(
                c_dev, (long long)n, 0ll, seeds[1], lo, span, jump_tables);
        }
        {
            dim3 grid((unsigned)((m_loc + 255) / 256), (unsigned)((n + kMatRun - 1) / kMatRun));
            generate_matrix_kernel<real><<<grid, 256, 0, stream>>>(T, ld, 1ll, n, m_loc, col0, seeds[2], lo, span, jump_tables);
        }
        CK(cudaGetLastError());
        if (sizeof(real) != sizeof(double) && (rc = keep_original_from_tableau())) return rc;
        CK(cudaStreamSynchronize(stream));
        stage = kLoaded;
        sec_load = now_s() - t0;
        return B2S_OK;
    }

    // fp32 solves keep the problem in fp64 for the final polish (single GPU; the sharded path returns the plain fp32 answer)
    int ensure_orig()
    {
        have_orig = false;
        if (sizeof(real) == sizeof(double) || world > 1 || !opt.fp64_polish) return B2S_OK;
        const size_t need = (size_t)(n + 1) * (size_t)m + (size_t)n;
        if (need > cap_orig) {
            cudaFree(orig64);
            orig64 = nullptr;
            cap_orig = 0;
            if (cudaMalloc(&orig64, sizeof(double) * need) != cudaSuccess) {
                cudaGetLastError();
                return B2S_OK;   // no room for the copy: solve without polish
            }
            cap_orig = need;
        }
        have_orig = true;
        return B2S_OK;
    }
    int keep_original_from_stage()   // stage_dev = [b | A | c] in fp64, exactly what the caller passed
    {
        int rc = ensure_orig();
        if (rc || !have_orig) return rc;
        CK(cudaMemcpyAsync(orig64, stage_dev, sizeof(double) * ((size_t)(n + 1) * (size_t)m + (size_t)n), cudaMemcpyDeviceToDevice, stream));
        return B2S_OK;
    }
    int keep_original_from_tableau()  // generated in working precision: the fp64 problem is the widened one
    {
        int rc = ensure_orig();
        if (rc || !have_orig) return rc;
        widen_rows<<<1024, 256, 0, stream>>>(orig64, (long long)m, T, ld, (long long)(n + 1), m);
        widen_rows<<<64, 256, 0, stream>>>(orig64 + (size_t)(n + 1) * m, (long long)n, c_dev, (long long)n, 1ll, n);
        CK(cudaGetLastError());
        return B2S_OK;
    }

    // Iterative refinement of x_B against the fp64 problem with the fp32 tableau's slack block as approximate inverse.
    // Returns false (and leaves x_dev / the fp32 objective alone) if the refinement does not contract.
    int polish_fp64(double* obj, bool* used)
    {
        *used = false;
        DevBuf<double> xB, r, o;
        DevBuf<unsigned long long> norms;
        CK(xB.alloc((size_t)m));
        CK(r.alloc((size_t)m));
        CK(o.alloc(1));
        CK(norms.alloc(2));
        const unsigned blocks = (unsigned)((m + 255) / 256);
        polish_init_kernel<real><<<blocks, 256, 0, stream>>>(P, xB.p);
        double prev = -1.0, rel = 1.0;
        for (int it = 0; it < 8; ++it) {
            CK(cudaMemsetAsync(norms.p, 0, 2 * sizeof(unsigned long long), stream));
            polish_residual_kernel<real><<<blocks, 256, 0, stream>>>(P, orig64, xB.p, r.p);
            polish_correct_kernel<real><<<blocks, 256, 0, stream>>>(P, r.p, xB.p, norms.p);
            unsigned long long hn[2];
            CK(cudaMemcpyAsync(hn, norms.p, sizeof(hn), cudaMemcpyDeviceToHost, stream));
            CK(cudaStreamSynchronize(stream));
            double dx, xm;
            memcpy(&dx, &hn[0], 8);
            memcpy(&xm, &hn[1], 8);
            if (!(dx == dx) || !(xm == xm)) return B2S_OK;          // NaN: give up, keep the fp32 answer
            rel = dx / std::max(xm, 1e-300);
            if (it >= 2 && prev >= 0 && rel > prev && rel > 1e-6) return B2S_OK;   // not contracting
            prev = rel;
            if (rel < 1e-15) break;
        }
        polish_residual = rel;
        if (rel > 1e-9) return B2S_OK;
        CK(cudaMemsetAsync(x_dev, 0, sizeof(double) * n, stream));
        polish_finish_kernel<real><<<1, kSelBlock, 0, stream>>>(P, orig64, xB.p, x_dev, o.p);
        CK(cudaMemcpyAsync(obj, o.p, sizeof(double), cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
        *used = true;
        return B2S_OK;
    }

    int copy_problem(double* A, double* b, double* c) override
    {
        if (stage != kLoaded) return fail(B2S_ERR_STATE, "copy_problem is only valid between load/generate and build");
        if (world > 1) return fail(B2S_ERR_STATE, "copy_problem is not available on a sharded solver");
        CK(cudaSetDevice(dev));
        if (sizeof(real) == sizeof(double)) {
            if (A)
                CK(cudaMemcpy2DAsync(A, sizeof(double) * m, T + ld, sizeof(real) * ld, sizeof(double) * m, (size_t)n,
                                     cudaMemcpyDeviceToHost, stream));
            if (b) CK(cudaMemcpyAsync(b, T, sizeof(double) * m, cudaMemcpyDeviceToHost, stream));
            if (c) CK(cudaMemcpyAsync(c, c_dev, sizeof(double) * n, cudaMemcpyDeviceToHost, stream));
            CK(cudaStreamSynchronize(stream));
        } else {
            int rc = ensure_stage((size_t)(n + 1) * (size_t)m + (size_t)n);
            if (rc) return rc;
            widen_rows<<<1024, 256, 0, stream>>>(stage_dev, (long long)m, T, ld, (long long)(n + 1), m);
            widen_rows<<<64, 256, 0, stream>>>(stage_dev + (size_t)(n + 1) * m, (long long)n, c_dev, (long long)n, 1ll, n);
            CK(cudaStreamSynchronize(stream));
            if (b) CK(cudaMemcpyAsync(b, stage_dev, sizeof(double) * m, cudaMemcpyDeviceToHost, stream));
            if (A) CK(cudaMemcpyAsync(A, stage_dev + m, sizeof(double) * (size_t)n * m, cudaMemcpyDeviceToHost, stream));
            if (c) CK(cudaMemcpyAsync(c, stage_dev + (size_t)(n + 1) * m, sizeof(double) * n, cudaMemcpyDeviceToHost, stream));
            CK(cudaStreamSynchronize(stream));
        }
        return B2S_OK;
    }

    // ---- phases ------------------------------------------------------------------------------
    int build_phase1() override
    {
        if (stage != kLoaded) return fail(B2S_ERR_STATE, "build_phase1 needs a freshly loaded or generated problem");
        CK(cudaSetDevice(dev));
        Rc = R1;
        phase = 1;
        fill_params();
        invalidate_graph();
        CK(cudaMemsetAsync(st, 0, sizeof(DevState), stream));
        const long long work = std::max<long long>(std::max<long long>(m, Rc), m_loc);
        build_misc_kernel<real><<<(unsigned)((work + 255) / 256), 256, 0, stream>>>(P, neg, folded ? 0 : 1);
        negate_kernel<real><<<num_sms * 4, 256, 0, stream>>>(P, neg);
        state_reset_kernel<<<1, 1, 0, stream>>>(st, 1);
        CK(cudaGetLastError());
        {
            int rc = reset_lookahead();
            if (rc) return rc;
        }
        if (p2p) {
            // pivot sequence numbers restart at 1: clear this rank's flags, then make sure every rank
            // has done so before anyone can publish
            CK(cudaMemsetAsync(arena, 0, sizeof(ArenaHeader<real>), stream));
            int rc = rank_barrier();
            if (rc) return rc;
        }
        stage = kBuilt;
        return B2S_OK;
    }

    int price_out() override
    {
        if (stage != kBuilt) return fail(B2S_ERR_STATE, "price_out follows build_phase1 / switch_phase2");
        CK(cudaSetDevice(dev));
        coef_kernel<real><<<(unsigned)((m_loc + 255) / 256), 256, 0, stream>>>(P, coef);
        const long long threads = Rc * 32;
        const unsigned blocks = (unsigned)((threads + 255) / 256);
        if (world > 1) {
            // Chain the running sums through the ranks in ascending constraint order so the result is
            // bit-identical to the single-GPU order (slabs are multiples of 64 constraints).
            std::vector<real> hbuf;
            for (int r = 0; r < world; ++r) {
                if (r == rank) priceout_kernel<real><<<blocks, 256, 0, stream>>>(P, coef);
                bool sent = false;
#ifdef B2S_WITH_NCCL
                if (comm) {
                    NK(ncclBroadcast(cost, cost, (size_t)Rc * sizeof(real), ncclChar, r, comm, stream));
                    sent = true;
                }
#endif
                if (!sent) {   // host layer: everybody contributes its vector, rank r's is the one that counts
                    hbuf.resize((size_t)Rc * (size_t)(world + 1));
                    CK(cudaMemcpyAsync(hbuf.data(), cost, sizeof(real) * (size_t)Rc, cudaMemcpyDeviceToHost, stream));
                    CK(cudaStreamSynchronize(stream));
                    int rc = host_allgather(hbuf.data(), hbuf.data() + Rc, sizeof(real) * (size_t)Rc);
                    if (rc) return rc;
                    CK(cudaMemcpyAsync(cost, hbuf.data() + (size_t)Rc * (size_t)(1 + r), sizeof(real) * (size_t)Rc, cudaMemcpyHostToDevice, stream));
                    CK(cudaStreamSynchronize(stream));
                }
            }
        } else {
            priceout_kernel<real><<<blocks, 256, 0, stream>>>(P, coef);
        }
        CK(cudaGetLastError());
        if (phase == 1 && (opt.relative_infeasibility || sizeof(real) != sizeof(double))) {  // magnitude the phase-1 objective starts from: -(sum |b_i|)
            real c0 = 0;
            CK(cudaMemcpyAsync(&c0, cost, sizeof(real), cudaMemcpyDeviceToHost, stream));
            CK(cudaStreamSynchronize(stream));
            cost0_phase1_start = (double)c0;
        }
        stage = kPriced;
        return B2S_OK;
    }

    int select_entering() override
    {
        if (stage != kPriced && !(attached && stage == kBuilt))
            return fail(B2S_ERR_STATE, "select_entering follows price_out");
        CK(cudaSetDevice(dev));
        select_kernel<real><<<P.Gc, kSelBlock, 0, stream>>>(P);
        CK(cudaGetLastError());
        stage = kReady;
        return B2S_OK;
    }

    int enqueue_pivot()
    {
        if (use_lookahead()) {
            // one launch per pivot: streaming update + the next pivot's selection (and, sharded, its two exchanges) under it
            return launch_la();
        }
        if (world > 1 && p2p) {
            // exchanges done by the kernels themselves over NVLink peer memory (b2s_p2p.cuh)
            ratio_p2p_kernel<real><<<P.Gm_loc, kSelBlock, 0, stream>>>(P);
            gather_p2p_kernel<real><<<(unsigned)((Rs + 255) / 256), 256, 0, stream>>>(P);
            svec_p2p_kernel<real><<<(unsigned)((std::max(Rs, ld) + 255) / 256), 256, 0, stream>>>(P);
            update_fn()<<<upd_grid, kSelBlock, upd_smem, stream>>>(P);
            return B2S_OK;
        }
#ifdef B2S_WITH_NCCL
        if (world > 1 && comm) return enqueue_pivot_sharded();
#endif
        if (world > 1) return fail(B2S_ERR_STATE, "sharded solve without peer memory and without NCCL");
        const long long work = std::max(Rs, ld);
        if (use_pdl) {
            // programmatic dependent launch: each kernel may be scheduled while its predecessor drains and
            // blocks in griddepcontrol.wait until that predecessor's memory is visible
            int rc;
            if ((rc = launch_pdl((const void*)ratio_kernel<real, false>, (unsigned)P.Gm, kSelBlock))) return rc;
            if ((rc = launch_pdl((const void*)gather_kernel<real, false>, (unsigned)((work + 255) / 256), 256))) return rc;
            return launch_pdl((const void*)update_fn(), (unsigned)upd_grid, kSelBlock);
        }
        ratio_kernel<real, false><<<P.Gm, kSelBlock, 0, stream>>>(P);
        gather_kernel<real, false><<<(unsigned)((work + 255) / 256), 256, 0, stream>>>(P);
        update_fn()<<<upd_grid, kSelBlock, upd_smem, stream>>>(P);
        return B2S_OK;
    }

    // The look-ahead kernel; B2S_LA_PDL=1 adds programmatic stream serialization between consecutive pivots (measured: no gain).
    int launch_la()
    {
        if (!la_pdl) {
            la_fn()<<<la_grid, kSelBlock, 0, stream>>>(P, 1);
            return B2S_OK;
        }
        return launch_pdl((const void*)la_fn(), (unsigned)la_grid, kSelBlock);
    }

    int launch_pdl(const void* fn, unsigned grid, unsigned block)
    {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(block);
        cfg.dynamicSmemBytes = 0;
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        int one = 1;
        void* args[] = {&P, &one};   // (the three-launch kernels take P only: a trailing argument is ignored by the launch)
        CK(cudaLaunchKernelExC(&cfg, fn, args));
        return B2S_OK;
    }

#ifdef B2S_WITH_NCCL
    // One pivot on a constraint-sharded tableau.  Two small exchanges per pivot:
    //  (1) all-gather of the ratio-test stage-1 block winners (the epsilon tournament is not
    //      associative, so every rank replays the reference's stage 2 on the full slot list);
    //  (2) the owner of constraint p publishes the raw pivot-constraint vector through an integer
    //      sum all-reduce in which the other ranks contribute zeros (bit-exact broadcast whose root
    //      is only known on the device).
    int enqueue_pivot_sharded()
    {
        ratio_kernel<real, true><<<P.Gm_loc, kSelBlock, 0, stream>>>(P);
        const size_t cnt = (size_t)P.Gm_loc;
        const size_t off = (size_t)P.Gm_loc0;
        NK(ncclGroupStart());
        NK(ncclAllGather(rslot_v + off, rslot_v, cnt * sizeof(real), ncclChar, comm, stream));
        NK(ncclAllGather(rslot_max + off, rslot_max, cnt * sizeof(real), ncclChar, comm, stream));
        NK(ncclAllGather(rslot_i + off, rslot_i, cnt * sizeof(int), ncclChar, comm, stream));
        NK(ncclAllGather(rslot_k + off, rslot_k, cnt * sizeof(int), ncclChar, comm, stream));
        NK(ncclGroupEnd());
        ratio_finish_kernel<real><<<1, kSelBlock, 0, stream>>>(P);
        gather_kernel<real, true><<<(unsigned)((Rs + 255) / 256), 256, 0, stream>>>(P);
        NK(ncclAllReduce(rowp, rowp, (size_t)Rs, sizeof(real) == 8 ? ncclUint64 : ncclUint32, ncclSum, comm, stream));
        svec_kernel<real><<<(unsigned)((ld + 255) / 256), 256, 0, stream>>>(P);
        update_fn()<<<upd_grid, kSelBlock, upd_smem, stream>>>(P);
        return B2S_OK;
    }
#endif

    int pick_batch() const
    {
        if (opt.batch > 0) return opt.batch;
        const double bytes = 2.0 * (double)Rs * (double)ld * sizeof(real);
        const double t = std::max(12e-6, bytes / 5.0e12);
        return (int)std::min(256.0, std::max(4.0, 3e-3 / t));
    }

    int launch_batch(int batch)
    {
        if (use_la_persist()) {
            PivotParams<real> Pk = P;
            int bk = batch;
            void* args[] = {&Pk, &bk};
            cudaError_t ce = cudaLaunchCooperativeKernel((const void*)la_persist_fn(), dim3((unsigned)la_grid), dim3(kSelBlock), args, 0, stream);
            if (ce == cudaSuccess) return B2S_OK;
            if (ce != cudaErrorCooperativeLaunchTooLarge && ce != cudaErrorLaunchOutOfResources)
                return fail(B2S_ERR_CUDA, "cooperative launch of the look-ahead kernel: %s", cudaGetErrorString(ce));
            cudaGetLastError();
            la_persist = false;   // the SMs cannot host one CTA each right now: one launch per pivot from here on
        }
        if (use_persistent()) {
            // one cooperative launch runs the whole batch of pivots (b2s_persistent.cuh)
            CK(cudaMemsetAsync(&st->bar_count, 0, sizeof(unsigned), stream));
            PivotParams<real> Pk = P;
            int bk = batch;
            void* args[] = {&Pk, &bk};
            cudaError_t ce = cudaLaunchCooperativeKernel((const void*)loop_fn(), dim3((unsigned)loop_grid), dim3(kSelBlock),
                                                         args, 0, stream);
            if (ce == cudaSuccess) return reset_tickets();
            if (ce != cudaErrorCooperativeLaunchTooLarge && ce != cudaErrorLaunchOutOfResources)
                return fail(B2S_ERR_CUDA, "cooperative launch of the pivot loop kernel: %s", cudaGetErrorString(ce));
            // the SMs cannot host one CTA each right now (another context holds resources): use the other GPU
            // loop body from here on -- same kernels' arithmetic, three launches per pivot
            cudaGetLastError();
            opt.persistent = 0;
            fill_params();
            int rc = reset_tickets();
            if (rc) return rc;
        }
        const bool graphable = opt.use_graph && (world == 1 || p2p);
        if (!graphable) {
            for (int k = 0; k < batch; ++k) {
                int rc = enqueue_pivot();
                if (rc) return rc;
            }
            CK(cudaGetLastError());
            return B2S_OK;
        }
        auto it = graphs.find(batch);
        if (it == graphs.end()) {
            cudaGraph_t graph = nullptr;
            cudaGraphExec_t exec = nullptr;
            CK(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal));
            for (int k = 0; k < batch; ++k) enqueue_pivot();
            CK(cudaStreamEndCapture(stream, &graph));
            cudaError_t e = cudaGraphInstantiate(&exec, graph, 0);
            cudaGraphDestroy(graph);
            if (e != cudaSuccess) return fail(B2S_ERR_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(e));
            it = graphs.emplace(batch, exec).first;
        }
        CK(cudaGraphLaunch(it->second, stream));
        return B2S_OK;
    }

    // Look-ahead state: proposals, ticket word and tile records restart with the pivot counter.
    int reset_lookahead()
    {
        CK(cudaMemsetAsync(la, 0, sizeof(LaState), stream));
        if (tile_rec && cap_tiles) CK(cudaMemsetAsync(tile_rec, 0, sizeof(unsigned) * cap_tiles, stream));
        return B2S_OK;
    }
    // First pivot of a phase / of an iterate() call: the helpers' chain on the quiescent tableau (no-op when the previous
    // update left a complete proposal), then its verdict (unbounded at once, silent peer).
    int enqueue_prologue()
    {
        la_prologue_kernel<real><<<P.helpers, kSelBlock, 0, stream>>>(P);
        la_prologue_commit_kernel<real><<<1, 1, 0, stream>>>(P);
        CK(cudaGetLastError());
        return B2S_OK;
    }
    int enqueue_flush()
    {
        la_flush_kernel<real><<<(unsigned)std::min<long long>((Rs + 255) / 256, 4 * num_sms), 256, 0, stream>>>(P);
        CK(cudaGetLastError());
        return B2S_OK;
    }

    // The ticket scheduler of the per-launch update kernels re-arms itself; the persistent loop kernel leaves the
    // counter where its last pivot stopped.  Clear both words before any per-launch kernel may follow it.
    int reset_tickets()
    {
        static_assert(offsetof(DevState, tile_done) == offsetof(DevState, tile_ticket) + sizeof(unsigned), "adjacent");
        CK(cudaMemsetAsync(&st->tile_ticket, 0, 2 * sizeof(unsigned), stream));
        return B2S_OK;
    }

    int fetch_state()
    {
        CK(cudaMemcpyAsync(st_host, st, sizeof(DevState), cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
        return B2S_OK;
    }

    int iterate(long long max_pivots, int* status, long long* done) override
    {
        if (stage != kReady && stage != kPhaseDone) return fail(B2S_ERR_STATE, "iterate follows select_entering");
        CK(cudaSetDevice(dev));
        int rc = fetch_state();
        if (rc) return rc;
        const long long start = st_host->pivots;
        if (done) *done = 0;
        if (st_host->status != kRunning || max_pivots == 0) {
            if (status) *status = st_host->status;
            return B2S_OK;
        }
        const long long limit = max_pivots < 0 ? LLONG_MAX : start + max_pivots;
        CK(cudaMemcpyAsync(&st->limit, &limit, sizeof(long long), cudaMemcpyHostToDevice, stream));
        CK(cudaEventRecord(ev0, stream));
        const bool la_on = use_lookahead();
        if (la_on && (rc = enqueue_prologue())) return rc;
        // Batches are enqueued one ahead of the status poll, so the device never idles while the host
        // looks at the state; a batch enqueued after the phase ended (or the budget ran out) costs only
        // its early-exit launches because every kernel checks the device-resident status/limit first.
        const int full = pick_batch();
        const bool graphable = opt.use_graph && (world == 1 || p2p);
        long long enq = 0;  // pivots enqueued so far (upper bound on pivots made)
        auto enqueue = [&](int slot) -> int {
            long long left = limit - start - enq;
            int batch = full;
            if (left < batch) {
                // a short budget: do not replay a full batch of mostly idle iterations.  Graphs exist per
                // power-of-two length, so at most log2(full) extra captures ever happen per phase.
                batch = (int)std::max<long long>(left, 1);
                if (graphable) {
                    int p2 = 1;
                    while (p2 < batch) p2 <<= 1;
                    batch = std::min(p2, full);
                }
            }
            int rc2 = launch_batch(batch);
            if (rc2) return rc2;
            enq += batch;
            CK(cudaMemcpyAsync(st_host + 1 + slot, st, sizeof(DevState), cudaMemcpyDeviceToHost, stream));
            CK(cudaEventRecord(poll_ev[slot], stream));
            return B2S_OK;
        };
        int slot = 0;
        if ((rc = enqueue(slot))) return rc;
        while (true) {
            const bool more = (limit == LLONG_MAX) || (start + enq < limit);
            if (more && (rc = enqueue(slot ^ 1))) return rc;
            CK(cudaEventSynchronize(poll_ev[slot]));
            const DevState& snap = st_host[1 + slot];
            if (snap.status != kRunning || snap.pivots >= limit) break;
            if (!more) {  // budget fully enqueued but not yet consumed: cannot happen (limit enforced on device)
                break;
            }
            slot ^= 1;
        }
        if (la_on && (rc = enqueue_flush())) return rc;
        if ((rc = fetch_state())) return rc;
        CK(cudaEventRecord(ev1, stream));
        CK(cudaEventSynchronize(ev1));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, ev0, ev1));
        const long long made = st_host->pivots - start;
        if (phase == 2) {
            pivots_p2 += made;
            sec_p2 += ms * 1e-3;
        } else {
            pivots_p1 += made;
            sec_p1 += ms * 1e-3;
        }
        if (done) *done = made;
        if (status) *status = st_host->status;
        if (st_host->status != kRunning) stage = kPhaseDone;
        return check_device_status();
    }

    // Only RUNNING / FEASIBLE / UNBOUNDED are legitimate ends of a batch.  Anything else was stored by a bounded wait
    // that gave up (a peer rank or a CTA never published): the tableau may be half updated, so the solve must stop.
    int check_device_status()
    {
        const int s_ = st_host->status;
        if (s_ == kRunning || s_ == kFeasible || s_ == kUnbounded) return B2S_OK;
        stage = kEmpty;  // nothing can continue from this tableau
        if (s_ == kStatusPeerTimeout)
            return fail(B2S_ERR_PEER, "rank %d/%d: a peer rank (or a CTA of the loop kernel) did not publish within the bounded wait "
                                      "during pivot %lld; the tableau is no longer consistent", rank, world, st_host->pivots + 1);
        return fail(B2S_ERR_CUDA, "device loop stopped with undocumented status %d", s_);
    }

    // Opt-in (b2s_options.drive_out_artificials): pivot the artificials that are still basic after a feasible phase 1 out of
    // the basis, constraint by constraint, with the ordinary gather + update kernels (see driveout_*_kernel).
    int drive_out(long long* made_out)
    {
        if (made_out) *made_out = 0;
        if (world > 1) return fail(B2S_ERR_STATE, "drive_out_artificials is single-GPU only");
        CK(cudaSetDevice(dev));
        std::vector<int> hb((size_t)m);
        CK(cudaMemcpyAsync(hb.data(), base, sizeof(int) * (size_t)m, cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
        DevBuf<int> found;
        CK(found.alloc(1));
        long long made = 0;
        const long long work = std::max(Rs, ld);
        const unsigned fb = (unsigned)(((long long)n + m + 255) / 256);
        for (int i = 0; i < m; ++i) {
            if (hb[(size_t)i] < n + m) continue;
            const int big = INT_MAX;
            CK(cudaMemcpyAsync(found.p, &big, sizeof(int), cudaMemcpyHostToDevice, stream));
            driveout_find_kernel<real><<<fb, 256, 0, stream>>>(P, i, found.p);
            int q = INT_MAX;
            CK(cudaMemcpyAsync(&q, found.p, sizeof(int), cudaMemcpyDeviceToHost, stream));
            CK(cudaStreamSynchronize(stream));
            if (q == INT_MAX) continue;   // redundant constraint: its artificial stays basic at zero
            int rc = reset_tickets();
            if (rc) return rc;
            driveout_setup_kernel<real><<<(unsigned)((m_loc + 255) / 256), 256, 0, stream>>>(P, q, i);
            gather_kernel<real, false><<<(unsigned)((work + 255) / 256), 256, 0, stream>>>(P);
            update_fn()<<<upd_grid, kSelBlock, upd_smem, stream>>>(P);
            CK(cudaGetLastError());
            ++made;
        }
        CK(cudaStreamSynchronize(stream));
        pivots_p1 += made;
        if (made_out) *made_out = made;
        return B2S_OK;
    }

    int phase1_verdict(int* status) override
    {
        if (phase != 1 || (stage != kReady && stage != kPhaseDone))
            return fail(B2S_ERR_STATE, "phase1_verdict follows the phase-1 pivots");
        CK(cudaSetDevice(dev));
        CK(cudaMemsetAsync(verdict, 0, 2 * sizeof(int), stream));
        verdict_kernel<real><<<(unsigned)((m + 255) / 256), 256, 0, stream>>>(P, verdict);
        int h[2] = {0, 0};
        CK(cudaMemcpyAsync(h, verdict, sizeof(h), cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
        if (opt.relative_infeasibility) {
            // Opt-in, beyond the reference: the residual a feasible phase 1 leaves in cost[0] scales
            // with the magnitude it started from, so the -1e-9 tolerance is taken relative to that magnitude.
            real c0 = 0;
            CK(cudaMemcpyAsync(&c0, cost, sizeof(real), cudaMemcpyDeviceToHost, stream));
            CK(cudaStreamSynchronize(stream));
            const double scale = std::max(1.0, std::fabs(cost0_phase1_start));
            h[0] = ((double)c0 < -1e-9 * scale) ? 1 : 0;
        }
        if (sizeof(real) != sizeof(double)) {
            // fp32 (no reference counterpart): the absolute -1e-9 test is below fp32 resolution -- after ~m pivots the phase-1
            // objective carries a residual of order 1e-6 x its starting magnitude.  With no artificial variable left in the basis
            // the phase-1 optimum is exactly 0 whatever the accumulated residual says; otherwise the residual is judged against
            // fp32 resolution.
            real c0 = 0;
            CK(cudaMemcpyAsync(&c0, cost, sizeof(real), cudaMemcpyDeviceToHost, stream));
            CK(cudaStreamSynchronize(stream));
            const double scale = std::max(1.0, std::fabs(cost0_phase1_start));
            h[0] = (h[1] > 0 && (double)c0 < -1e-5 * scale) ? 1 : 0;
        }
        if (opt.drive_out_artificials && world == 1 && h[0] == 0 && h[1] > 0) {
            // beyond the reference: pivot the basic artificials out; what may remain sits in redundant constraints
            long long made = 0;
            int rc = drive_out(&made);
            if (rc) return rc;
            CK(cudaMemsetAsync(verdict, 0, 2 * sizeof(int), stream));
            driveout_verdict_kernel<real><<<(unsigned)m, 256, 0, stream>>>(P, verdict);
            CK(cudaMemcpyAsync(h, verdict, sizeof(h), cudaMemcpyDeviceToHost, stream));
            CK(cudaStreamSynchronize(stream));
        }
        *status = h[0] ? B2S_INFEASIBLE : (h[1] > 0 ? B2S_DEGENERATE : B2S_FEASIBLE);
        return B2S_OK;
    }

    int switch_phase2() override
    {
        if (phase != 1 || (stage != kReady && stage != kPhaseDone))
            return fail(B2S_ERR_STATE, "switch_phase2 follows phase 1");
        CK(cudaSetDevice(dev));
        phase = 2;
        Rc = 1 + (long long)n + m;  // rows -= cols (src/twoPhaseMethod.cu:288)
        Rs = Rc;                    // unfolded layout: the trailing artificial rows are simply ignored
        fill_params();
        invalidate_graph();
        phase2_costs_kernel<real><<<(unsigned)(((long long)n + m + 255) / 256), 256, 0, stream>>>(P, c_dev);
        state_reset_kernel<<<1, 1, 0, stream>>>(st, 0);  // back to RUNNING, keep counters / hash
        CK(cudaGetLastError());
        {
            int rc = reset_lookahead();
            if (rc) return rc;
        }
        if (p2p) {
            // a proposal prepared but not executed in phase 1 has left flags with the next pivot's number in the arenas
            int rc = rank_barrier();   // nobody is still inside a kernel that reads them
            if (rc) return rc;
            CK(cudaMemsetAsync(arena, 0, sizeof(ArenaHeader<real>), stream));
            if ((rc = rank_barrier())) return rc;
        }
        stage = kBuilt;
        return B2S_OK;
    }

    int extract(double* x, double* obj) override
    {
        if (stage == kEmpty || stage == kLoaded) return fail(B2S_ERR_STATE, "nothing to extract");
        CK(cudaSetDevice(dev));
        CK(cudaMemsetAsync(x_dev, 0, sizeof(double) * n, stream));
        solution_kernel<real><<<(unsigned)((m_loc + 255) / 256), 256, 0, stream>>>(P, x_dev);
        bool reduced = world <= 1;
#ifdef B2S_WITH_NCCL
        if (!reduced && comm) {
            NK(ncclAllReduce(x_dev, x_dev, (size_t)n, ncclDouble, ncclSum, comm, stream));
            reduced = true;
        }
#endif
        if (!reduced) {   // host layer: every entry is non-zero on at most one rank, so the sum is exact in any order
            std::vector<double> hx((size_t)n * (size_t)(world + 1));
            CK(cudaMemcpyAsync(hx.data(), x_dev, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, stream));
            CK(cudaStreamSynchronize(stream));
            int rc = host_allgather(hx.data(), hx.data() + n, sizeof(double) * (size_t)n);
            if (rc) return rc;
            for (int j = 0; j < n; ++j) {
                double acc = 0.0;
                for (int r = 0; r < world; ++r) acc += hx[(size_t)n * (size_t)(1 + r) + (size_t)j];
                hx[(size_t)j] = acc;
            }
            CK(cudaMemcpyAsync(x_dev, hx.data(), sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, stream));
            CK(cudaStreamSynchronize(stream));
        }
        real c0;
        CK(cudaMemcpyAsync(&c0, cost, sizeof(real), cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
        double objective = (double)c0;
        if (sizeof(real) != sizeof(double) && have_orig && phase == 2 && world == 1) {
            bool used = false;
            double polished = 0.0;
            int rc = polish_fp64(&polished, &used);   // rewrites x_dev on success
            if (rc) return rc;
            if (used) objective = polished;
        }
        if (x) CK(cudaMemcpyAsync(x, x_dev, sizeof(double) * n, cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
        if (obj) *obj = objective;
        return B2S_OK;
    }

    int solve(int* status, double* x, double* obj, int* basis, b2s_stats* stats) override
    {
        const double t0 = now_s();
        int rc, stt = B2S_RUNNING;
        long long done = 0;
        const long long cap = opt.max_pivots > 0 ? opt.max_pivots : -1;
        if ((rc = build_phase1())) return rc;
        if ((rc = price_out())) return rc;
        if ((rc = select_entering())) return rc;
        if ((rc = iterate(cap, &stt, &done))) return rc;
        int result;
        if (stt == B2S_RUNNING) {
            result = B2S_ITER_LIMIT;
        } else {
            if ((rc = phase1_verdict(&result))) return rc;  // the phase-1 solve status is ignored (src/twoPhaseMethod.cu:258)
            if (result == B2S_FEASIBLE) {
                if ((rc = switch_phase2())) return rc;
                if ((rc = price_out())) return rc;
                if ((rc = select_entering())) return rc;
                const long long left = cap < 0 ? -1 : std::max<long long>(0, cap - done);
                if ((rc = iterate(left, &stt, &done))) return rc;
                if (stt == B2S_RUNNING)
                    result = B2S_ITER_LIMIT;
                else if (stt != B2S_FEASIBLE)
                    result = stt;
                else if ((rc = extract(x, obj)))
                    return rc;
            }
        }
        if (basis && (rc = copy_basis(basis))) return rc;
        if (status) *status = result;
        if (stats) {
            get_stats(stats);
            stats->seconds_total = now_s() - t0 + sec_load;
        }
        return B2S_OK;
    }

    // ---- introspection -----------------------------------------------------------------------
    int get_dims(int* n_, int* m_, long long* ra, long long* rs, long long* ld_) const override
    {
        if (n_) *n_ = n;
        if (m_) *m_ = m;
        if (ra) *ra = Rc;
        if (rs) *rs = Rs;
        if (ld_) *ld_ = ld;
        return B2S_OK;
    }

    int copy_tableau(double* out) override
    {
        if (stage == kEmpty) return fail(B2S_ERR_STATE, "no problem loaded");
        if (world > 1) return fail(B2S_ERR_STATE, "copy_tableau is not available on a sharded solver");
        CK(cudaSetDevice(dev));
        const size_t count = (size_t)Rc * (size_t)m;
        int rc = ensure_stage(count);
        if (rc) return rc;
        export_kernel<real><<<num_sms * 8, 256, 0, stream>>>(P, stage_dev);
        CK(cudaMemcpyAsync(out, stage_dev, sizeof(double) * count, cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
        return B2S_OK;
    }

    int copy_costs(double* out) override
    {
        if (stage == kEmpty) return fail(B2S_ERR_STATE, "no problem loaded");
        CK(cudaSetDevice(dev));
        std::vector<real> tmp((size_t)Rc);
        CK(cudaMemcpyAsync(tmp.data(), cost, sizeof(real) * (size_t)Rc, cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
        for (long long i = 0; i < Rc; ++i) out[i] = (double)tmp[(size_t)i];
        return B2S_OK;
    }

    int copy_basis(int* out) override
    {
        if (stage == kEmpty || stage == kLoaded) return fail(B2S_ERR_STATE, "no basis yet");
        CK(cudaSetDevice(dev));
        CK(cudaMemcpyAsync(out, base, sizeof(int) * (size_t)m, cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
        return B2S_OK;
    }

    int copy_trace(int* qp, long long cap, long long* len, unsigned long long* hash) override
    {
        CK(cudaSetDevice(dev));
        int rc = fetch_state();
        if (rc) return rc;
        const long long have = st_host->pivots;
        if (len) *len = have;
        if (hash) *hash = st_host->hash;
        const long long cnt = std::min(std::min(have, cap), trace_cap);
        if (qp && cnt > 0) {
            CK(cudaMemcpyAsync(qp, trace, sizeof(int2) * (size_t)cnt, cudaMemcpyDeviceToHost, stream));
            CK(cudaStreamSynchronize(stream));
        }
        return B2S_OK;
    }

    int get_stats(b2s_stats* out) override
    {
        CK(cudaSetDevice(dev));
        int rc = fetch_state();
        if (rc) return rc;
        memset(out, 0, sizeof(*out));
        out->pivots_phase1 = pivots_p1;
        out->pivots_phase2 = pivots_p2;
        out->trace_hash = st_host->hash;
        out->seconds_load = sec_load;
        out->seconds_phase1 = sec_p1;
        out->seconds_phase2 = sec_p2;
        out->rows_streamed = st_host->rows_streamed;
        out->rows_total = pivots_p1 * (folded ? 1 + (long long)n + m : R1) + pivots_p2 * (1 + (long long)n + m);
        return B2S_OK;
    }

    // ---- caller-owned tableau (tabular_t of include/tabular.cuh:5-30) ---------------------------
    int attach(double* table, size_t pitch, int rows, int cols, double* costs, int n_vars) override
    {
        if (sizeof(real) != sizeof(double)) return fail(B2S_ERR_ARG, "attach needs an fp64 solver (reference TYPE)");
        if (world > 1) return fail(B2S_ERR_STATE, "attach is single-GPU only");
        if (!table || !costs || rows < 2 || cols < 1 || n_vars < 0) return fail(B2S_ERR_ARG, "bad tableau description");
        if (pitch % 32 != 0 || pitch < (size_t)cols * sizeof(double) || ((size_t)table % 32) != 0)
            return fail(B2S_ERR_ARG, "tableau pitch/base must be 32-byte aligned (cudaMallocPitch guarantees it)");
        CK(cudaSetDevice(dev));
        // scratch sized for this shape; no tableau storage of our own
        folded = false;
        world = 1;
        n = std::max(n_vars, 1);
        m = cols;
        m_loc = m;
        col0 = 0;
        const long long ld_ = (long long)(pitch / sizeof(double));
        if ((size_t)rows > cap_rows || (size_t)ld_ > cap_cols || (size_t)n > cap_n || (size_t)m > cap_m) {
            free_problem();
            int rc;
            if ((rc = dmalloc(&T_own, 1))) return rc;
            if ((rc = dmalloc(&cost_own, 1))) return rc;
            if ((rc = dmalloc(&rowp, (size_t)rows))) return rc;
            if ((rc = dmalloc(&col, (size_t)ld_))) return rc;
            if ((rc = dmalloc(&s, (size_t)ld_))) return rc;
            if ((rc = dmalloc(&coef, (size_t)ld_))) return rc;
            if ((rc = dmalloc(&neg, (size_t)ld_))) return rc;
            if ((rc = dmalloc(&c_dev, (size_t)n))) return rc;
            if ((rc = dmalloc(&x_dev, (size_t)n))) return rc;
            if ((rc = dmalloc(&base, (size_t)m))) return rc;
            if ((rc = dmalloc(&rslot_v, kMaxSlots))) return rc;
            if ((rc = dmalloc(&rslot_max, kMaxSlots))) return rc;
            if ((rc = dmalloc(&rslot_i, kMaxSlots))) return rc;
            if ((rc = dmalloc(&rslot_k, kMaxSlots))) return rc;
            if ((rc = dmalloc(&cslot_v, kMaxSlots))) return rc;
            if ((rc = dmalloc(&cslot_i, kMaxSlots))) return rc;
            if ((rc = dmalloc(&cslot_k, kMaxSlots))) return rc;
            if ((rc = dmalloc(&col2, 2 * (size_t)ld_))) return rc;
            if ((rc = dmalloc(&s2, 2 * (size_t)ld_))) return rc;
            if ((rc = dmalloc(&rowp2, 2 * (size_t)rows))) return rc;
            if ((rc = dmalloc(&rowval, 2 * (size_t)rows))) return rc;
            if ((rc = dmalloc(&rowlist, 2 * (size_t)rows))) return rc;
            if ((rc = dmalloc(&rowpos, 2 * (size_t)rows))) return rc;
            cap_T = 0;
            cap_rows = (size_t)rows;
            cap_cols = (size_t)ld_;
            cap_n = (size_t)n;
            cap_m = (size_t)m;
        }
        ld = ld_;
        R1 = Rs = Rc = rows;
        T = reinterpret_cast<real*>(table);
        cost = reinterpret_cast<real*>(costs);
        attached = true;
        phase = (rows == 1 + n_vars + cols) ? 2 : 1;
        pivots_p1 = pivots_p2 = 0;
        sec_load = sec_p1 = sec_p2 = 0;
        invalidate_graph();
        fill_params();
        CK(cudaMemsetAsync(st, 0, sizeof(DevState), stream));
        state_reset_kernel<<<1, 1, 0, stream>>>(st, 1);
        CK(cudaGetLastError());
        {
            int rc = reset_lookahead();
            if (rc) return rc;
        }
        stage = kBuilt;
        return B2S_OK;
    }

    int set_basis(const int* base_host) override
    {
        if (stage == kEmpty || stage == kLoaded || !base_host) return fail(B2S_ERR_STATE, "set_basis needs a built or attached tableau");
        CK(cudaSetDevice(dev));
        CK(cudaMemcpyAsync(base, base_host, sizeof(int) * (size_t)m, cudaMemcpyHostToDevice, stream));
        CK(cudaStreamSynchronize(stream));
        return B2S_OK;
    }

    // ---- device-vector primitives of include/reduction.cuh --------------------------------------
    int run_tournament_device(const real* dvals /* cnt+1 entries, [0] unused */, long long cnt, double* value, unsigned* index)
    {
        DevBuf<real> dv;
        DevBuf<int> di, dk;
        DevBuf<DevState> dst;
        CK(dv.alloc(kMaxSlots));
        CK(di.alloc(kMaxSlots));
        CK(dk.alloc(kMaxSlots));
        CK(dst.alloc(1));
        CK(cudaMemsetAsync(dst.p, 0, sizeof(DevState), stream));
        PivotParams<real> Q{};
        Q.cost = const_cast<real*>(dvals);
        Q.Rc = cnt + 1;
        Q.fold_from = LLONG_MAX;
        Q.cslot_v = dv.p;
        Q.cslot_i = di.p;
        Q.cslot_k = dk.p;
        Q.st = dst.p;
        Q.rule = opt.pivot_rule;
        Q.Gc = (int)std::max<long long>(1, std::min<long long>((cnt + kSelBlock - 1) / kSelBlock, kMaxSlots));
        select_kernel<real><<<Q.Gc, kSelBlock, 0, stream>>>(Q);
        DevState hs;
        CK(cudaMemcpyAsync(&hs, dst.p, sizeof(DevState), cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
        if (value) *value = hs.cq;
        if (index) *index = (unsigned)hs.q;
        return B2S_OK;
    }

    int min_element_device(const double* dvec, long long cnt, double* value, unsigned* index) override
    {
        if (sizeof(real) != sizeof(double)) return fail(B2S_ERR_ARG, "fp64 only");
        if (cnt < 1 || !dvec) return fail(B2S_ERR_ARG, "minElement needs a non-empty device vector");
        CK(cudaSetDevice(dev));
        DevBuf<real> tmp;  // the reference reduces a scratch copy too (src/reduction.cu:87-89)
        CK(tmp.alloc((size_t)cnt + 1));
        CK(cudaMemcpyAsync(tmp.p + 1, dvec, sizeof(real) * (size_t)cnt, cudaMemcpyDeviceToDevice, stream));
        return run_tournament_device(tmp.p, cnt, value, index);
    }

    int ratio_min_device(const double* known, const double* column, long long cnt, double* value, unsigned* index) override
    {
        if (sizeof(real) != sizeof(double)) return fail(B2S_ERR_ARG, "fp64 only");
        if (cnt < 1 || !known || !column) return fail(B2S_ERR_ARG, "ratio minElement needs two device vectors");
        CK(cudaSetDevice(dev));
        DevBuf<real> tmp;
        CK(tmp.alloc((size_t)cnt + 1));
        ratio_vector_kernel<real><<<(unsigned)((cnt + 255) / 256), 256, 0, stream>>>(
            reinterpret_cast<const real*>(known), reinterpret_cast<const real*>(column), cnt, tmp.p + 1);
        return run_tournament_device(tmp.p, cnt, value, index);
    }

    int max_le_zero_device(const double* dvec, long long cnt, int* result) override
    {
        if (sizeof(real) != sizeof(double)) return fail(B2S_ERR_ARG, "fp64 only");
        if (cnt < 1 || !dvec || !result) return fail(B2S_ERR_ARG, "isLessOrEqualThanZero needs a device vector");
        CK(cudaSetDevice(dev));
        DevBuf<real> out;
        CK(out.alloc(1));
        max_vector_kernel<real><<<1, kSelBlock, 0, stream>>>(reinterpret_cast<const real*>(dvec), cnt, out.p);
        real h = 0;
        CK(cudaMemcpyAsync(&h, out.p, sizeof(real), cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
        const double d = (double)h;  // compare(max) <= 0  <=>  max < 1e-9   (src/reduction.cu:200)
        *result = (fabs(d) < 1e-9 || d < 0.0) ? 1 : 0;
        return B2S_OK;
    }

    // ---- kernel-level hooks ------------------------------------------------------------------
    int tournament(const double* vec, long long cnt, double* value, int* index) override
    {
        if (cnt < 1 || !vec) return fail(B2S_ERR_ARG, "tournament needs a non-empty vector");
        CK(cudaSetDevice(dev));
        std::vector<real> h((size_t)cnt + 1);
        h[0] = 0;
        for (long long i = 0; i < cnt; ++i) h[(size_t)i + 1] = (real)vec[i];
        DevBuf<real> dcost;
        CK(dcost.alloc((size_t)cnt + 1));
        CK(cudaMemcpyAsync(dcost.p, h.data(), sizeof(real) * ((size_t)cnt + 1), cudaMemcpyHostToDevice, stream));
        unsigned idx = 0;
        int rc = run_tournament_device(dcost.p, cnt, value, &idx);
        if (index) *index = (int)idx;
        return rc;
    }

    int bench_update(int launches, int flush, float* ms, double* bytes) override
    {
        if (stage == kEmpty) return fail(B2S_ERR_STATE, "bench_update needs a loaded problem (dimensions)");
        CK(cudaSetDevice(dev));
        if (flush && !flush_buf) {
            flush_bytes = 512ull << 20;
            CK(cudaMalloc(&flush_buf, flush_bytes));
        }
        {
            int rc = reset_tickets();
            if (rc) return rc;
        }
        const long long work = std::max(Rs, ld);
        bench_fill_kernel<real><<<(unsigned)((work + 255) / 256), 256, 0, stream>>>(P);
        std::vector<cudaEvent_t> ev(2 * (size_t)launches);
        for (auto& e : ev) CK(cudaEventCreate(&e));
        for (int k = 0; k < launches; ++k) {
            if (flush) CK(cudaMemsetAsync(flush_buf, k & 0xff, flush_bytes, stream));
            bench_bump_kernel<<<1, 1, 0, stream>>>(st);
            CK(cudaEventRecord(ev[2 * k], stream));
            update_fn()<<<upd_grid, kSelBlock, upd_smem, stream>>>(P);
            CK(cudaEventRecord(ev[2 * k + 1], stream));
        }
        CK(cudaGetLastError());
        CK(cudaStreamSynchronize(stream));
        for (int k = 0; k < launches; ++k) CK(cudaEventElapsedTime(ms + k, ev[2 * k], ev[2 * k + 1]));
        for (auto& e : ev) cudaEventDestroy(e);
        if (bytes) *bytes = 2.0 * (double)Rs * (double)m_loc * sizeof(real);
        stage = kEmpty;  // tableau contents destroyed
        CK(cudaMemsetAsync(st, 0, sizeof(DevState), stream));
        CK(cudaStreamSynchronize(stream));
        return B2S_OK;
    }

    // Real pivots, launched one kernel at a time with CUDA events between the three launches.
    int profile_pivots(int count, float* ms_ratio, float* ms_gather, float* ms_update, long long* done) override
    {
        if (stage == kPhaseDone) {  // the phase ended exactly at the end of the previous chunk: nothing left to time
            if (done) *done = 0;
            return B2S_OK;
        }
        if (stage != kReady) return fail(B2S_ERR_STATE, "profile_pivots follows select_entering / iterate");
        if (use_lookahead()) {
            // one launch per pivot: the whole pivot is the "update" column
            int rc = profile_lookahead(count, ms_update, nullptr, done);
            const long long k = done ? *done : 0;
            for (long long i = 0; i < k; ++i) ms_ratio[i] = ms_gather[i] = 0.f;
            return rc;
        }
        if (world > 1) return fail(B2S_ERR_STATE, "profile_pivots is single-GPU only without the look-ahead kernel");
        CK(cudaSetDevice(dev));
        int rc = fetch_state();
        if (rc) return rc;
        if ((rc = reset_tickets())) return rc;
        const long long start = st_host->pivots;
        const long long limit = start + count;
        CK(cudaMemcpyAsync(&st->limit, &limit, sizeof(long long), cudaMemcpyHostToDevice, stream));
        std::vector<cudaEvent_t> ev(4 * (size_t)count);
        for (auto& e : ev) CK(cudaEventCreate(&e));
        const long long work = std::max(Rs, ld);
        for (int k = 0; k < count; ++k) {
            CK(cudaEventRecord(ev[4 * k + 0], stream));
            ratio_kernel<real, false><<<P.Gm, kSelBlock, 0, stream>>>(P);
            CK(cudaEventRecord(ev[4 * k + 1], stream));
            gather_kernel<real, false><<<(unsigned)((work + 255) / 256), 256, 0, stream>>>(P);
            CK(cudaEventRecord(ev[4 * k + 2], stream));
            update_fn()<<<upd_grid, kSelBlock, upd_smem, stream>>>(P);
            CK(cudaEventRecord(ev[4 * k + 3], stream));
        }
        CK(cudaGetLastError());
        if ((rc = fetch_state())) return rc;
        for (int k = 0; k < count; ++k) {
            CK(cudaEventElapsedTime(ms_ratio + k, ev[4 * k + 0], ev[4 * k + 1]));
            CK(cudaEventElapsedTime(ms_gather + k, ev[4 * k + 1], ev[4 * k + 2]));
            CK(cudaEventElapsedTime(ms_update + k, ev[4 * k + 2], ev[4 * k + 3]));
        }
        for (auto& e : ev) cudaEventDestroy(e);
        const long long made = st_host->pivots - start;
        if (phase == 2) pivots_p2 += made; else pivots_p1 += made;
        if (done) *done = made;
        if (st_host->status != kRunning) stage = kPhaseDone;
        return B2S_OK;
    }

    int dist_init_host(int rank_, int world_, b2s_allgather_fn fn, void* user) override
    {
        if (world_ < 1 || rank_ < 0 || rank_ >= world_ || world_ > kMaxPeers || (world_ > 1 && !fn))
            return fail(B2S_ERR_ARG, "bad rank/world %d/%d (at most %d ranks) or missing all-gather callback", rank_, world_, kMaxPeers);
        CK(cudaSetDevice(dev));
#ifdef B2S_WITH_NCCL
        if (comm) {
            ncclCommDestroy(comm);
            comm = nullptr;
        }
#endif
        close_arena();
        rank = rank_;
        world = world_;
        host_ag = fn;
        host_ag_user = user;
        free_problem();
        return B2S_OK;
    }

    int loop_info(int* launches, int* lookahead, int* persistent) override
    {
        const bool la_on = stage != kEmpty && use_lookahead();
        const bool pers = stage != kEmpty && use_persistent();
        if (launches) *launches = pers ? 0 : (la_on ? 1 : (world > 1 ? 4 : 3));
        if (lookahead) *lookahead = la_on ? 1 : 0;
        if (persistent) *persistent = pers ? 1 : 0;
        return B2S_OK;
    }

    // Look-ahead kernel, one launch at a time, with the chain's globaltimer stamps read back after each pivot.
    int profile_lookahead(int count, float* kernel_ms, double* stage_us, long long* done) override
    {
        if (stage == kPhaseDone) {
            if (done) *done = 0;
            return B2S_OK;
        }
        if (stage != kReady) return fail(B2S_ERR_STATE, "profile_lookahead follows select_entering / iterate");
        if (!use_lookahead()) return fail(B2S_ERR_STATE, "the look-ahead kernel is not in use for this problem");
        CK(cudaSetDevice(dev));
        int rc = fetch_state();
        if (rc) return rc;
        const long long start = st_host->pivots;
        const long long limit = start + count;
        CK(cudaMemcpyAsync(&st->limit, &limit, sizeof(long long), cudaMemcpyHostToDevice, stream));
        if ((rc = enqueue_prologue())) return rc;
        std::vector<cudaEvent_t> ev(2 * (size_t)count);
        for (auto& e : ev) CK(cudaEventCreate(&e));
        std::vector<LaState> snap((size_t)count);
        long long made = 0;
        for (int k = 0; k < count; ++k) {
            CK(cudaEventRecord(ev[2 * k], stream));
            if ((rc = launch_la())) return rc;
            CK(cudaEventRecord(ev[2 * k + 1], stream));
            if (stage_us) {   // the stamps are per pivot: read them before the next launch overwrites them
                CK(cudaMemcpyAsync(&snap[(size_t)k], la, sizeof(LaState), cudaMemcpyDeviceToHost, stream));
                CK(cudaStreamSynchronize(stream));
            }
        }
        CK(cudaGetLastError());
        if ((rc = enqueue_flush())) return rc;
        if ((rc = fetch_state())) return rc;
        made = st_host->pivots - start;
        for (long long k = 0; k < made; ++k) {
            CK(cudaEventElapsedTime(kernel_ms + k, ev[2 * k], ev[2 * k + 1]));
            if (stage_us) {
                const unsigned long long* t = snap[(size_t)k].stamps;
                const int order[6] = {1, 2, 3, 4, 5, 6};
                for (int j = 0; j < 6; ++j)
                    stage_us[6 * k + j] = t[order[j]] >= t[0] ? (double)(t[order[j]] - t[0]) * 1e-3 : -1.0;
            }
        }
        for (auto& e : ev) cudaEventDestroy(e);
        if (phase == 2) pivots_p2 += made; else pivots_p1 += made;
        if (done) *done = made;
        if (st_host->status != kRunning) stage = kPhaseDone;
        return check_device_status();
    }

    int dist_init(int rank_, int world_, const char* id) override
    {
#ifdef B2S_WITH_NCCL
        if (world_ < 1 || rank_ < 0 || rank_ >= world_) return fail(B2S_ERR_ARG, "bad rank/world %d/%d", rank_, world_);
        CK(cudaSetDevice(dev));
        if (comm) {
            ncclCommDestroy(comm);
            comm = nullptr;
        }
        close_arena();
        rank = rank_;
        world = world_;
        host_ag = nullptr;
        host_ag_user = nullptr;
        if (world > 1) {
            ncclUniqueId uid;
            static_assert(sizeof(uid) == B2S_NCCL_ID_BYTES, "unique id size");
            memcpy(&uid, id, sizeof(uid));
            NK(ncclCommInitRank(&comm, world, uid, rank));
        }
        free_problem();
        return B2S_OK;
#else
        (void)rank_;
        (void)world_;
        (void)id;
        return fail(B2S_ERR_NCCL, "library built without NCCL");
#endif
    }

};

}  // namespace b2s

// =============================================================================================
// C ABI
// =============================================================================================
using b2s::SolverBase;

struct b2s_solver {
    SolverBase* impl;
};

extern "C" {

void b2s_default_options(b2s_options* opt)
{
    memset(opt, 0, sizeof(*opt));
    opt->device = 0;
    opt->dtype = B2S_F64;
    opt->pivot_rule = B2S_RULE_REFERENCE;
    opt->fold_artificials = 1;
    opt->skip_zero_rows = 1;  /* value-exact (src/solver.cu:43: fma(s_i, 0, x) == x): rows with a_pr == 0 are not streamed */
    opt->use_graph = 1;
    opt->batch = 0;
    opt->max_pivots = 0;
    opt->trace_capacity = 0;
    opt->update_variant = 8;
    opt->persistent = 2;
    opt->lookahead = 2;
    opt->fp64_polish = 1;
    opt->drive_out_artificials = 0;
}

int b2s_device_count(void)
{
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess) return 0;
    return count;
}

int b2s_create(const b2s_options* opt, b2s_solver** out)
{
    if (!out) return B2S_ERR_ARG;
    *out = nullptr;
    b2s_options o;
    if (opt)
        o = *opt;
    else
        b2s_default_options(&o);
    if (o.pivot_rule < 0 || o.pivot_rule > 2) {
        b2s::g_thread_error = "unknown pivot_rule";
        return B2S_ERR_ARG;
    }
    SolverBase* impl = nullptr;
    int rc;
    if (o.dtype == B2S_F64) {
        auto* p = new b2s::SolverImpl<double>();
        p->opt = o;
        rc = p->init();
        impl = p;
    } else if (o.dtype == B2S_F32) {
        auto* p = new b2s::SolverImpl<float>();
        p->opt = o;
        rc = p->init();
        impl = p;
    } else {
        b2s::g_thread_error = "unknown dtype";
        return B2S_ERR_ARG;
    }
    if (rc != B2S_OK) {
        delete impl;
        return rc;
    }
    *out = new b2s_solver{impl};
    return B2S_OK;
}

void b2s_destroy(b2s_solver* s)
{
    if (!s) return;
    delete s->impl;
    delete s;
}

const char* b2s_last_error(const b2s_solver* s)
{
    if (s && s->impl) return s->impl->err.c_str();
    return b2s::g_thread_error.c_str();
}

#define B2S_FWD(expr)               \
    if (!s || !s->impl) return B2S_ERR_ARG; \
    return s->impl->expr

int b2s_load_problem_host(b2s_solver* s, int n, int m, const double* A, const double* b, const double* c) { B2S_FWD(load_host(n, m, A, b, c)); }
int b2s_generate_problem_device(b2s_solver* s, int n, int m, const unsigned seeds[3], double lo, double hi) { B2S_FWD(generate(n, m, seeds, lo, hi)); }
int b2s_copy_problem(b2s_solver* s, double* A, double* b, double* c) { B2S_FWD(copy_problem(A, b, c)); }
int b2s_solve_two_phase(b2s_solver* s, int* status, double* x, double* objective, int* basis, b2s_stats* stats) { B2S_FWD(solve(status, x, objective, basis, stats)); }
int b2s_build_phase1(b2s_solver* s) { B2S_FWD(build_phase1()); }
int b2s_price_out(b2s_solver* s) { B2S_FWD(price_out()); }
int b2s_select_entering(b2s_solver* s) { B2S_FWD(select_entering()); }
int b2s_iterate(b2s_solver* s, long long max_pivots, int* status, long long* pivots_done) { B2S_FWD(iterate(max_pivots, status, pivots_done)); }
int b2s_phase1_verdict(b2s_solver* s, int* status)
{
    if (!status) return B2S_ERR_ARG;
    B2S_FWD(phase1_verdict(status));
}
int b2s_switch_phase2(b2s_solver* s) { B2S_FWD(switch_phase2()); }
int b2s_extract_solution(b2s_solver* s, double* x, double* objective) { B2S_FWD(extract(x, objective)); }
int b2s_get_dims(const b2s_solver* s, int* n, int* m, long long* rows_active, long long* rows_stored, long long* ld) { B2S_FWD(get_dims(n, m, rows_active, rows_stored, ld)); }
int b2s_copy_tableau(b2s_solver* s, double* out) { B2S_FWD(copy_tableau(out)); }
int b2s_copy_costs(b2s_solver* s, double* out) { B2S_FWD(copy_costs(out)); }
int b2s_copy_basis(b2s_solver* s, int* out) { B2S_FWD(copy_basis(out)); }
int b2s_copy_trace(b2s_solver* s, int* qp_pairs, long long capacity, long long* length, unsigned long long* hash) { B2S_FWD(copy_trace(qp_pairs, capacity, length, hash)); }
int b2s_get_stats(b2s_solver* s, b2s_stats* stats)
{
    if (!stats) return B2S_ERR_ARG;
    B2S_FWD(get_stats(stats));
}
int b2s_tournament(b2s_solver* s, const double* vec, long long n, double* value, int* index) { B2S_FWD(tournament(vec, n, value, index)); }
int b2s_bench_update(b2s_solver* s, int launches, int flush_l2, float* ms_each, double* bytes_per_launch)
{
    if (launches < 1 || !ms_each) return B2S_ERR_ARG;
    B2S_FWD(bench_update(launches, flush_l2, ms_each, bytes_per_launch));
}
int b2s_dist_init(b2s_solver* s, int rank, int world, const char id[B2S_NCCL_ID_BYTES]) { B2S_FWD(dist_init(rank, world, id)); }
int b2s_attach_tableau_device(b2s_solver* s, double* table, size_t pitch_bytes, int rows, int cols, double* costs, int n_vars) { B2S_FWD(attach(table, pitch_bytes, rows, cols, costs, n_vars)); }
int b2s_set_basis(b2s_solver* s, const int* basis_host) { B2S_FWD(set_basis(basis_host)); }
int b2s_min_element_device(b2s_solver* s, const double* dvec, long long n, double* value, unsigned* index) { B2S_FWD(min_element_device(dvec, n, value, index)); }
int b2s_ratio_min_device(b2s_solver* s, const double* known_terms, const double* column, long long n, double* value, unsigned* index) { B2S_FWD(ratio_min_device(known_terms, column, n, value, index)); }
int b2s_max_le_zero_device(b2s_solver* s, const double* dvec, long long n, int* result) { B2S_FWD(max_le_zero_device(dvec, n, result)); }
int b2s_profile_pivots(b2s_solver* s, int count, float* ms_ratio, float* ms_gather, float* ms_update, long long* pivots_done)
{
    if (count < 1 || !ms_ratio || !ms_gather || !ms_update) return B2S_ERR_ARG;
    B2S_FWD(profile_pivots(count, ms_ratio, ms_gather, ms_update, pivots_done));
}

int b2s_get_loop_info(b2s_solver* s, int* launches_per_pivot, int* lookahead, int* persistent) { B2S_FWD(loop_info(launches_per_pivot, lookahead, persistent)); }
int b2s_profile_lookahead(b2s_solver* s, int count, float* kernel_ms, double* stage_us, long long* pivots_done)
{
    if (count < 1 || !kernel_ms) return B2S_ERR_ARG;
    B2S_FWD(profile_lookahead(count, kernel_ms, stage_us, pivots_done));
}

int b2s_dist_init_host(b2s_solver* s, int rank, int world, b2s_allgather_fn allgather, void* user) { B2S_FWD(dist_init_host(rank, world, allgather, user)); }

int b2s_dist_unique_id(char id[B2S_NCCL_ID_BYTES])
{
#ifdef B2S_WITH_NCCL
    ncclUniqueId uid;
    if (ncclGetUniqueId(&uid) != ncclSuccess) return B2S_ERR_NCCL;
    memcpy(id, &uid, sizeof(uid));
    return B2S_OK;
#else
    (void)id;
    return B2S_ERR_NCCL;
#endif
}

/* srand(seed); rand() x3 (src/problem.cu:63-67) for glibc's TYPE_3 generator or MSVC's LCG. */
void b2s_seed_triplet(unsigned seed, int rand_flavour, unsigned out[3])
{
    if (rand_flavour == B2S_RAND_MSVC) {
        unsigned state = seed;
        for (int k = 0; k < 3; ++k) {
            state = state * 214013u + 2531011u;
            out[k] = (state >> 16) & 0x7fffu;
        }
        return;
    }
    // glibc random_r, TYPE_3: 31-word additive feedback register seeded by a Lehmer sequence,
    // 310 outputs discarded, result = word >> 1.
    int reg[34];
    reg[0] = seed ? (int)seed : 1;
    for (int i = 1; i < 31; ++i) {
        const long long hi = reg[i - 1] / 127773, lo = reg[i - 1] % 127773;
        long long word = 16807 * lo - 2836 * hi;
        if (word < 0) word += 2147483647;
        reg[i] = (int)word;
    }
    unsigned ring[34];
    for (int i = 0; i < 31; ++i) ring[i] = (unsigned)reg[i];
    for (int i = 31; i < 34; ++i) ring[i] = ring[i - 31];
    // sliding window of the last 34 words
    std::vector<unsigned> w(ring, ring + 34);
    for (int i = 34; i < 344 + 3; ++i) w.push_back(w[i - 31] + w[i - 3]);
    for (int k = 0; k < 3; ++k) out[k] = w[344 + k] >> 1;
}

}  // extern "C"
