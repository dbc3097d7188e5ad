// qcb_internal.hpp -- process-wide context and state handle shared by the translation units
// of libqcb200.so.
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <map>
#include <string>
#include <vector>

#include "../../include/qcb200.h"
#include "schedule.hpp"

struct qcb_state {
    int n = 0;        // logical qubits
    int nloc = 0;     // local index bits held by this process (n unless sharded)
    int dtype = QCB_C128;
    void *d = nullptr;
    bool owned = false;
    bool sharded = false;
    size_t bytes = 0;
    std::vector<int> phys; // logical qubit -> physical index bit (>= nloc: rank bit nloc + k)
    void *alt = nullptr;   // sharded states: second buffer for out-of-place qubit swaps (nullptr: chunked in-place swaps)
    // peer-mapped views (CUDA IPC) of every rank's `d` and `alt`, indexed by rank; empty = swaps go through NCCL only
    std::vector<void *> peer_d, peer_alt;
};

namespace qcb {

struct Ctx {
    bool inited = false;
    int device = -1;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr; // non-blocking, for H2D/D2H chunks overlapped with compute
    cudaEvent_t chunk_ev[2][16] = {};
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // options
    long tile_bits = 11, low_bits = 3, max_gates_per_pass = 1 << 30, force_simple = 0, adjoint_tile_bits = 0, adjoint_mma = 1, adjoint_warps = 0, adjoint_prog_host = 0, measure_ctas = 0, contract_coop = 1, measure_persistent = 0, measure_tma = 1, pass_tma = 0, pass_tma_max_stages = 3, adjoint_f32 = 1, adjoint_ctas = 5, adjoint_defer = 1, f64_3m_min_gates = 4, pass_async = 0, small_cluster = 12, small_gbits = -1, ham_tile = 1;
    // stats
    double st_passes = 0, st_stages = 0, st_launches = 0, st_gates = 0, st_launches_total = 0, st_swaps = 0, st_fused_swaps = 0;
    long fuse_swaps = 1, peer_swap = 1, dist_second_buffer = 1;
    double st_peer_swaps = 0; // cumulative: in-place qubit swaps done by k_swap_peer (NVLink peer loads / stores, no staging)
    // per-launch device timing of the gate passes (option "time_passes"; read and reset by the stats "pass_ms" / "pass_count")
    long time_passes = 0;
    std::vector<cudaEvent_t> pass_ev; // pairs (start, stop)
    size_t pass_ev_used = 0;
    struct PassMeta { int stages, gates, first_window_bit, tma; unsigned long long masks; }; // masks: 4 bits per stage, stage 0 lowest
    std::vector<PassMeta> pass_meta; // one per timed launch (diagnostics: env QCB_PASS_LOG=<file> dumps them with their times)
    double st_tma_passes = 0, st_fw_passes = 0, st_bw_passes = 0, st_ham_passes = 0;
    double st_cluster_circuits = 0; // cumulative: gate lists run by the cluster-resident kernel (small_circuit.cuh)
    // gate table of the cluster-resident kernel: pinned host staging + device copy, and the event of the last upload
    void *h_gates = nullptr, *d_gates = nullptr;
    size_t gates_cap = 0;
    cudaEvent_t gates_ev = nullptr;
    double st_async_passes = 0; // cumulative: gate passes launched on the persistent cp.async kernel (tile_async.cuh)
    double st_gates_3m = 0; // cumulative: gates applied in the 3M arithmetic (full diamonds of ComplexF64 passes), see tile.cuh
    int *d_barrier = nullptr; // 1 int for stream-ordered inter-rank barriers (NCCL all-reduce)
    // scratch
    // backward sweep of an unsharded gradient: the per-pass accumulator rows of ALL passes stay in d_partial (one area per
    // pass) and are reduced by two launches at the end of the sweep instead of two or three per pass
    struct AdjDefer {
        bool on = false;
        int pass = 0, npasses = 0;
        size_t stride = 0;   // doubles per pass area
        int nrows = 0;       // accumulator rows per pass (mma: CTAs x warps; else CTAs)
        int kind = 0;        // 1: raw rows of the tensor-core pass (64 doubles per gate), 2: environment partials (32 doubles)
    } adj_defer;
    double *d_partial = nullptr; // reduction partials
    size_t partial_cap = 0;      // in doubles
    double *d_out = nullptr;     // small device result buffer
    size_t out_cap = 0;
    double *h_out = nullptr;     // pinned mirror
    void *scratch[4] = {nullptr, nullptr, nullptr, nullptr}; // 0, 1: gradient work states; 2: host-call state / swap staging; 3: spare shard
    size_t scratch_bytes[4] = {0, 0, 0, 0};
    std::map<std::string, std::vector<PassPlan>> plan_cache;
    // distributed
    int rank = 0, nranks = 1;
    void *nccl_comm = nullptr;
};

Ctx &ctx();
int set_error(int code, const char *fmt, ...);

#define QCB_CUDA(call)                                                                                   \
    do {                                                                                                 \
        cudaError_t e__ = (call);                                                                        \
        if (e__ != cudaSuccess)                                                                          \
            return qcb::set_error(e__ == cudaErrorMemoryAllocation ? QCB_ENOMEM : QCB_ECUDA, "%s: %s (%s:%d)", #call, \
                                  cudaGetErrorString(e__), __FILE__, __LINE__);                         \
    } while (0)

#define QCB_REQUIRE_INIT()                                                                               \
    do {                                                                                                 \
        if (!qcb::ctx().inited) return qcb::set_error(QCB_ECUDA, "qcb_init has not been called (no CUDA device bound; there is no CPU fallback)"); \
    } while (0)

#define QCB_TRY(expr)                  \
    do {                               \
        int rc__ = (expr);             \
        if (rc__ != QCB_OK) return rc__; \
    } while (0)

inline size_t amp_bytes(int dtype) { return dtype == QCB_C128 ? 16 : 8; }

struct HamEval;
bool small_gradient_eligible(const qcb_state *s, int ngates);
int run_small_gradient(const qcb_state *s, const std::vector<GateOp> &gates, const double *mats, const HamEval &H, double *d_env, double *d_energy2);
int ensure_scratch(int slot, size_t bytes);
int ensure_partial(size_t doubles);
int ensure_out(size_t doubles);
// brackets one gate-pass launch with CUDA events on the work stream when "time_passes" is on
struct PassTimerScope {
    cudaEvent_t stop = nullptr;
    PassTimerScope()
    {
        Ctx &c = ctx();
        if (!c.time_passes) return;
        if (c.pass_ev_used + 2 > c.pass_ev.size()) {
            cudaEvent_t a, b;
            if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) return;
            c.pass_ev.push_back(a);
            c.pass_ev.push_back(b);
        }
        cudaEventRecord(c.pass_ev[c.pass_ev_used], c.stream);
        stop = c.pass_ev[c.pass_ev_used + 1];
        c.pass_ev_used += 2;
    }
    ~PassTimerScope()
    {
        if (stop) cudaEventRecord(stop, ctx().stream);
    }
};

inline void count_launch(int k = 1)
{
    ctx().st_launches += k;
    ctx().st_launches_total += k;
}

// gate-list execution (passes.cu)
int run_gate_list(qcb_state *s, const std::vector<GateOp> &gates, const double *mats, bool adjoint, const void *first_src = nullptr);
int run_adjoint_fused(qcb_state *psi, qcb_state *lam, const std::vector<GateOp> &rev, const double *mats, double *d_env, bool *env_post);
int launch_adjoint_plan(qcb_state *psi, qcb_state *lam, const PassPlan &plan, const std::vector<GateOp> &rev, const double *mats, double *d_env);
bool adjoint_uses_mma(const qcb_state *psi);
SchedOptions adjoint_sched_options(const qcb_state *psi);
int launch_plan(qcb_state *s, const PassPlan &plan, const std::vector<GateOp> &gates, const double *mats, bool adjoint);
int launch_plan_peer(qcb_state *s, const PeerSrc &src, void *dst, const PassPlan &plan, const std::vector<GateOp> &gates, const double *mats, bool adjoint);
SchedOptions current_sched_options(const qcb_state *s);

// distributed (dist.cu)
int dist_state_canonicalize(qcb_state *s);
int dist_run_gate_list(qcb_state *s, const std::vector<GateOp> &gates, const double *mats, bool adjoint);
int dist_alloc_alt(qcb_state *s);
int dist_swap_top_now(qcb_state *s);                 // NCCL exchange of the rank bits with the top local bits + layout update
int dist_allreduce_sum(double *device_buf, int count); // in place, over all ranks, on the work stream
// backward sweep of the adjoint gradient on a pair of sharded states in the same layout (qubit swaps move both)
// *spare: a third buffer of the shards' size (or nullptr): qubit swaps then run out of place, the three buffers rotate
int dist_run_adjoint_pair(qcb_state *psi, qcb_state *lam, void **spare, const std::vector<GateOp> &rev, const double *mats, double *d_env);
int dist_swap_top_pair(qcb_state *psi, qcb_state *lam, void **spare); // swap_top on both states + layout update

} // namespace qcb
