// dist.cu -- multi-GPU sharding: one process per GPU, amplitudes sharded by the top log2(P) index bits,
// gates on those qubits handled by global<->local qubit swaps.  A swap is an all-to-all of contiguous blocks; it
// runs either as grouped NCCL send/recv over NVLink 5 / NVSwitch (swap_top_device) or -- when the ranks have mapped
// each other's shards through CUDA IPC -- fused into the next gate pass, whose tile loads then come straight from
// peer memory (GpuShardExec::run_pass, tile_pass_peer_kernel).  No reference counterpart; the
// schedule and the layout bookkeeping are in dist_plan.hpp and are replayed on the CPU by tests/emu with simulated ranks.
#include <nccl.h>

#include "dist_plan.hpp"
#include "qcb_internal.hpp"

namespace qcb {

#define QCB_NCCL(call)                                                                                  \
    do {                                                                                                \
        ncclResult_t r__ = (call);                                                                      \
        if (r__ != ncclSuccess) return set_error(QCB_ENCCL, "%s: %s", #call, ncclGetErrorString(r__)); \
    } while (0)

int dist_alloc_alt(qcb_state *s)
{
    if (!ctx().dist_second_buffer) return QCB_OK; // (option: exercise the paths of shards too large for a second buffer)
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) return QCB_OK;
    if (free_b < s->bytes + ((size_t)6 << 30)) return QCB_OK; // keep headroom for scratch and NCCL buffers
    if (cudaMalloc(&s->alt, s->bytes) != cudaSuccess) { s->alt = nullptr; cudaGetLastError(); }
    return QCB_OK;
}

int stream_barrier();

// In-place exchange over NVLink peer memory, for shards too large for a second buffer (34 qubits on 2 GPUs): block j of this
// rank <-> block `me` of rank j, element by element, no staging.  Each pair of ranks splits its exchange: the lower rank
// swaps the first half of the block pair, the higher rank the second half, so both directions of every link carry loads and
// stores at the same time and every element is touched by exactly one thread.  Ordering across ranks: a stream barrier
// before (every shard is final) and after (nobody still reads or writes a peer's shard).
struct PeerPtrs { uint4 *p[8]; };
static __global__ void __launch_bounds__(256) k_swap_peer(uint4 *__restrict__ mine, PeerPtrs peers, int P, int me, unsigned long long blk16)
{
    const unsigned long long half = blk16 >> 1, stride = (unsigned long long)gridDim.x * blockDim.x;
    for (int j = 0; j < P; j++) {
        if (j == me) continue;
        const unsigned long long off = me < j ? 0ull : half;
        uint4 *a = mine + (unsigned long long)j * blk16 + off;
        uint4 *b = peers.p[j] + (unsigned long long)me * blk16 + off;
        unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
        for (; i + 3 * stride < half; i += 4 * stride) { // four independent remote loads in flight per thread
            const uint4 b0 = b[i], b1 = b[i + stride], b2 = b[i + 2 * stride], b3 = b[i + 3 * stride];
            const uint4 a0 = a[i], a1 = a[i + stride], a2 = a[i + 2 * stride], a3 = a[i + 3 * stride];
            a[i] = b0; a[i + stride] = b1; a[i + 2 * stride] = b2; a[i + 3 * stride] = b3;
            b[i] = a0; b[i + stride] = a1; b[i + 2 * stride] = a2; b[i + 3 * stride] = a3;
        }
        for (; i < half; i += stride) {
            const uint4 bv = b[i], av = a[i];
            a[i] = bv;
            b[i] = av;
        }
    }
}

// exchange rank bit k with local bit nloc-g+k: block j of this rank <-> block `rank` of rank j
static int swap_top_device(qcb_state *s)
{
    Ctx &c = ctx();
    ncclComm_t comm = (ncclComm_t)c.nccl_comm;
    const int P = c.nranks, me = c.rank;
    int g = 0;
    while ((1 << g) < P) g++;
    const size_t blk = s->bytes >> g;
    char *d = (char *)s->d;
    if (s->alt) {
        char *o = (char *)s->alt;
        QCB_NCCL(ncclGroupStart());
        for (int j = 0; j < P; j++) {
            if (j == me) continue;
            QCB_NCCL(ncclSend(d + (size_t)j * blk, blk, ncclChar, j, comm, c.stream));
            QCB_NCCL(ncclRecv(o + (size_t)j * blk, blk, ncclChar, j, comm, c.stream));
        }
        QCB_NCCL(ncclGroupEnd());
        QCB_CUDA(cudaMemcpyAsync(o + (size_t)me * blk, d + (size_t)me * blk, blk, cudaMemcpyDeviceToDevice, c.stream));
        std::swap(s->d, s->alt);
        count_launch(2);
    } else if (c.peer_swap && (int)s->peer_d.size() == P && blk % 32 == 0) {
        PeerPtrs pp;
        for (int r = 0; r < 8; r++) pp.p[r] = r < P ? static_cast<uint4 *>(s->peer_d[r]) : nullptr;
        QCB_TRY(stream_barrier());
        k_swap_peer<<<(unsigned)c.sm_count * 8u, 256, 0, c.stream>>>(reinterpret_cast<uint4 *>(d), pp, P, me, (unsigned long long)(blk / 16));
        QCB_CUDA(cudaGetLastError());
        QCB_TRY(stream_barrier());
        count_launch(3);
        c.st_peer_swaps += 1;
    } else {
        // in place through a bounded staging buffer: (P-1) chunks
        const size_t CH = std::min(blk, (size_t)256 << 20);
        QCB_TRY(ensure_scratch(2, CH * (size_t)(P - 1)));
        char *stage = (char *)c.scratch[2];
        for (size_t off = 0; off < blk; off += CH) {
            const size_t len = std::min(CH, blk - off);
            QCB_NCCL(ncclGroupStart());
            int slot = 0;
            for (int j = 0; j < P; j++) {
                if (j == me) continue;
                QCB_NCCL(ncclSend(d + (size_t)j * blk + off, len, ncclChar, j, comm, c.stream));
                QCB_NCCL(ncclRecv(stage + (size_t)slot * CH, len, ncclChar, j, comm, c.stream));
                slot++;
            }
            QCB_NCCL(ncclGroupEnd());
            slot = 0;
            for (int j = 0; j < P; j++) {
                if (j == me) continue;
                QCB_CUDA(cudaMemcpyAsync(d + (size_t)j * blk + off, stage + (size_t)slot * CH, len, cudaMemcpyDeviceToDevice, c.stream));
                slot++;
            }
            count_launch(1 + (P - 1));
        }
    }
    c.st_swaps += 1;
    return QCB_OK;
}

// stream-ordered barrier across the ranks: a 1-int all-reduce on the work stream
int stream_barrier()
{
    Ctx &c = ctx();
    if (!c.d_barrier) {
        QCB_CUDA(cudaMalloc(&c.d_barrier, sizeof(int)));
        QCB_CUDA(cudaMemsetAsync(c.d_barrier, 0, sizeof(int), c.stream));
    }
    QCB_NCCL(ncclAllReduce(c.d_barrier, c.d_barrier, 1, ncclInt, ncclSum, (ncclComm_t)c.nccl_comm, c.stream));
    return QCB_OK;
}

struct GpuShardExec : ShardExecutor {
    qcb_state *s;
    bool pending = false; // a swap_top whose data movement rides on the loads of the next pass
    explicit GpuShardExec(qcb_state *st) : s(st) {}
    bool can_fuse() const { return ctx().fuse_swaps && s->alt && (int)s->peer_d.size() == ctx().nranks && (int)s->peer_alt.size() == ctx().nranks; }
    int run_pass(const PassPlan &pl, const std::vector<GateOp> &gates, const double *mats, bool adjoint) override
    {
        Ctx &c = ctx();
        int rc;
        if (pending) {
            // fused qubit swap: load this pass's tiles straight from the peers' shards (NVLink P2P), write the local
            // second buffer.  Barrier before: every rank has finished writing its shard; barrier after: nobody still reads
            // the buffer that becomes the next destination.
            PeerSrc src;
            int g = 0;
            while ((1 << g) < c.nranks) g++;
            src.shift = s->nloc - g;
            const size_t ab = amp_bytes(s->dtype);
            for (int r = 0; r < 8; r++)
                src.base[r] = r < c.nranks ? (const void *)((const char *)s->peer_d[r] + (((size_t)c.rank << src.shift) * ab)) : nullptr;
            QCB_TRY(stream_barrier());
            rc = launch_plan_peer(s, src, s->alt, pl, gates, mats, adjoint);
            if (rc != QCB_OK) return rc;
            QCB_TRY(stream_barrier());
            std::swap(s->d, s->alt);
            s->peer_d.swap(s->peer_alt);
            pending = false;
            c.st_fused_swaps += 1;
            count_launch(2);
        } else {
            rc = launch_plan(s, pl, gates, mats, adjoint);
            if (rc != QCB_OK) return rc;
        }
        c.st_passes += 1;
        c.st_stages += (double)pl.stages.size();
        return QCB_OK;
    }
    int swap_top() override
    {
        if (can_fuse()) { pending = true; ctx().st_swaps += 1; return QCB_OK; }
        const int rc = swap_top_device(s);
        if (rc == QCB_OK && s->alt && !s->peer_d.empty()) s->peer_d.swap(s->peer_alt);
        return rc;
    }
    int flush() override
    {
        if (!pending) return QCB_OK;
        pending = false;
        ctx().st_swaps -= 1; // swap_top_device counts it again
        const int rc = swap_top_device(s);
        // the NCCL path exchanges d <-> alt as well when alt exists: keep the peer views in step
        if (rc == QCB_OK && s->alt && !s->peer_d.empty()) s->peer_d.swap(s->peer_alt);
        return rc;
    }
};

static ShardLayout layout_of(const qcb_state *s)
{
    ShardLayout L;
    L.n = s->n; L.nloc = s->nloc; L.g = s->n - s->nloc;
    L.phys = s->phys;
    return L;
}

int dist_run_gate_list(qcb_state *s, const std::vector<GateOp> &gates, const double *mats, bool adjoint)
{
    if (!ctx().nccl_comm) return set_error(QCB_ENCCL, "qcb_dist_init has not been called");
    GpuShardExec ex(s);
    ShardLayout L = layout_of(s);
    int rc;
    try {
        rc = run_gates_sharded(ex, L, gates, mats, adjoint, current_sched_options(s), nullptr);
    } catch (const std::exception &e) {
        ex.flush();
        s->phys = L.phys;
        return set_error(QCB_EINVAL, "sharded scheduler: %s", e.what());
    }
    if (rc != QCB_OK) ex.flush(); // error return: a deferred qubit swap must still happen, the layout below says it did
    s->phys = L.phys;
    if (rc < 0) return set_error(QCB_EINVAL, "sharded scheduler could not make progress (code %d)", rc);
    return rc;
}

// ---- sharded adjoint gradients: the backward sweep runs on psi and lambda = H psi together ---------------------------
// Both states share one layout; a planned gate pass becomes one fused backward pass on the pair (passes.cu:
// launch_adjoint_plan, environments accumulated per rank), a bit-permuting pass and a qubit swap are applied to both.
static int swap_top_pair_device(qcb_state *psi, qcb_state *lam, void **spare)
{
    // with a spare buffer: psi <-> spare out of place (its old buffer becomes the spare), then lambda the same way
    for (qcb_state *s : {psi, lam}) {
        s->alt = spare ? *spare : nullptr;
        const int rc = swap_top_device(s); // exchanges s->d and s->alt when alt exists
        if (spare) *spare = s->alt;
        s->alt = nullptr;
        if (rc != QCB_OK) return rc;
    }
    return QCB_OK;
}

struct PairShardExec : ShardExecutor {
    qcb_state *psi, *lam;
    void **spare;
    double *d_env;
    PairShardExec(qcb_state *a, qcb_state *b, void **sp, double *env) : psi(a), lam(b), spare(sp), d_env(env) {}
    int run_pass(const PassPlan &pl, const std::vector<GateOp> &gates, const double *mats, bool) override
    {
        if (!pl.stages.empty()) return launch_adjoint_plan(psi, lam, pl, gates, mats, d_env);
        QCB_TRY(launch_plan(psi, pl, gates, mats, true)); // no gates: a pure bit permutation
        QCB_TRY(launch_plan(lam, pl, gates, mats, true));
        ctx().st_passes += 2;
        return QCB_OK;
    }
    int swap_top() override { return swap_top_pair_device(psi, lam, spare); }
};

int dist_run_adjoint_pair(qcb_state *psi, qcb_state *lam, void **spare, const std::vector<GateOp> &rev, const double *mats, double *d_env)
{
    if (!ctx().nccl_comm) return set_error(QCB_ENCCL, "qcb_dist_init has not been called");
    if (psi->phys != lam->phys) return set_error(QCB_EINVAL, "psi and lambda must share one layout");
    PairShardExec ex(psi, lam, spare, d_env);
    ShardLayout L = layout_of(psi);
    int rc;
    try {
        rc = run_gates_sharded(ex, L, rev, mats, true, adjoint_sched_options(psi), nullptr);
    } catch (const std::exception &e) {
        return set_error(QCB_EINVAL, "sharded scheduler: %s", e.what());
    }
    psi->phys = L.phys;
    lam->phys = L.phys;
    if (rc < 0) return set_error(QCB_EINVAL, "sharded scheduler could not make progress (code %d)", rc);
    return rc;
}

int dist_swap_top_pair(qcb_state *psi, qcb_state *lam, void **spare)
{
    if (!ctx().nccl_comm) return set_error(QCB_ENCCL, "qcb_dist_init has not been called");
    if (psi->phys != lam->phys) return set_error(QCB_EINVAL, "psi and lambda must share one layout");
    QCB_TRY(swap_top_pair_device(psi, lam, spare));
    ShardLayout L = layout_of(psi);
    swap_top_layout(L);
    psi->phys = L.phys;
    lam->phys = L.phys;
    return QCB_OK;
}

int dist_swap_top_now(qcb_state *s)
{
    if (!ctx().nccl_comm) return set_error(QCB_ENCCL, "qcb_dist_init has not been called");
    QCB_TRY(swap_top_device(s));
    if (s->alt && !s->peer_d.empty()) s->peer_d.swap(s->peer_alt);
    ShardLayout L = layout_of(s);
    swap_top_layout(L);
    s->phys = L.phys;
    return QCB_OK;
}

int dist_allreduce_sum(double *device_buf, int count)
{
    Ctx &c = ctx();
    if (!c.nccl_comm) return set_error(QCB_ENCCL, "qcb_dist_init has not been called");
    QCB_NCCL(ncclAllReduce(device_buf, device_buf, (size_t)count, ncclDouble, ncclSum, (ncclComm_t)c.nccl_comm, c.stream));
    return QCB_OK;
}

int dist_state_canonicalize(qcb_state *s)
{
    if (!ctx().nccl_comm) return set_error(QCB_ENCCL, "qcb_dist_init has not been called");
    GpuShardExec ex(s);
    ShardLayout L = layout_of(s);
    int rc;
    try {
        rc = canonicalize(ex, L, current_sched_options(s));
    } catch (const std::exception &e) {
        ex.flush();
        s->phys = L.phys;
        return set_error(QCB_EINVAL, "canonicalize: %s", e.what());
    }
    if (rc != QCB_OK) ex.flush();
    s->phys = L.phys;
    return rc;
}

} // namespace qcb

using namespace qcb;

extern "C" {

int qcb_dist_unique_id(void *out_128_bytes)
{
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    if (!out_128_bytes) return set_error(QCB_EINVAL, "null pointer");
    ncclUniqueId id;
    QCB_NCCL(ncclGetUniqueId(&id));
    memcpy(out_128_bytes, &id, sizeof id);
    return QCB_OK;
}

int qcb_dist_init(int rank, int nranks, const void *unique_id_128_bytes)
{
    QCB_REQUIRE_INIT();
    Ctx &c = ctx();
    if (nranks < 1 || (nranks & (nranks - 1)) || rank < 0 || rank >= nranks)
        return set_error(QCB_EINVAL, "nranks must be a power of two and 0 <= rank < nranks");
    if (!unique_id_128_bytes) return set_error(QCB_EINVAL, "null unique id");
    if (c.nccl_comm) qcb_dist_finalize();
    ncclUniqueId id;
    memcpy(&id, unique_id_128_bytes, sizeof id);
    ncclComm_t comm;
    QCB_NCCL(ncclCommInitRank(&comm, nranks, id, rank));
    c.nccl_comm = comm;
    c.rank = rank;
    c.nranks = nranks;
    return QCB_OK;
}

/* CUDA IPC handles of this rank's two state buffers (2 x 64 bytes; the second is all zero when the state has no second
 * buffer).  The host gathers the handles of all ranks (any channel) and passes them to qcb_state_ipc_import, after which
 * qubit swaps are fused into the following gate pass as peer-memory loads over NVLink. */
int qcb_state_ipc_export(const qcb_state *s, void *out_128_bytes)
{
    QCB_REQUIRE_INIT();
    if (!s || !s->d || !out_128_bytes) return set_error(QCB_EINVAL, "null pointer");
    if (!s->sharded || !s->owned) return set_error(QCB_EINVAL, "only library-owned sharded states can be exported");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    memset(out_128_bytes, 0, 128);
    cudaIpcMemHandle_t h;
    QCB_CUDA(cudaIpcGetMemHandle(&h, s->d));
    memcpy(out_128_bytes, &h, 64);
    if (s->alt) {
        QCB_CUDA(cudaIpcGetMemHandle(&h, s->alt));
        memcpy((char *)out_128_bytes + 64, &h, 64);
    }
    return QCB_OK;
}

int qcb_state_ipc_import(qcb_state *s, const void *all_handles /* nranks x 128 bytes, rank-major */)
{
    QCB_REQUIRE_INIT();
    Ctx &c = ctx();
    if (!s || !s->d || !all_handles) return set_error(QCB_EINVAL, "null pointer");
    if (!s->sharded || c.nranks > 8) return set_error(QCB_EINVAL, "peer mapping needs a sharded state and at most 8 ranks");
    if (!s->peer_d.empty()) qcb_state_ipc_close(s);
    const char zero[64] = {0};
    // some rank has no second buffer: map the first buffers only (in-place peer swaps, k_swap_peer); the swaps fused into
    // gate passes need both
    bool both = s->alt != nullptr, none = s->alt == nullptr;
    for (int r = 0; r < c.nranks; r++) {
        const bool has = memcmp((const char *)all_handles + 128 * (size_t)r + 64, zero, 64) != 0;
        both = both && has;
        none = none && !has;
    }
    if (!both && !none) return QCB_OK; // mixed: no mapping at all, every swap stays on NCCL
    s->peer_d.assign(c.nranks, nullptr);
    if (both) s->peer_alt.assign(c.nranks, nullptr);
    for (int r = 0; r < c.nranks; r++) {
        const char *h = (const char *)all_handles + 128 * (size_t)r;
        if (r == c.rank) { s->peer_d[r] = s->d; if (both) s->peer_alt[r] = s->alt; continue; }
        cudaIpcMemHandle_t hd, ha;
        memcpy(&hd, h, 64);
        memcpy(&ha, h + 64, 64);
        cudaError_t e1 = cudaIpcOpenMemHandle(&s->peer_d[r], hd, cudaIpcMemLazyEnablePeerAccess);
        cudaError_t e2 = (e1 == cudaSuccess && both) ? cudaIpcOpenMemHandle(&s->peer_alt[r], ha, cudaIpcMemLazyEnablePeerAccess) : e1;
        if (e1 != cudaSuccess || e2 != cudaSuccess) {
            cudaGetLastError();
            qcb_state_ipc_close(s); // unmaps what the lower ranks already opened
            return set_error(QCB_ECUDA, "cudaIpcOpenMemHandle(rank %d): %s", r, cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
        }
    }
    return QCB_OK;
}

/* Unmaps the peers' buffers again (call on every rank, then synchronise the ranks, before destroying the state). */
int qcb_state_ipc_close(qcb_state *s)
{
    if (!s) return QCB_OK;
    Ctx &c = ctx();
    if (c.inited) cudaStreamSynchronize(c.stream);
    for (size_t r = 0; r < s->peer_d.size(); r++) {
        if ((int)r == c.rank) continue;
        if (s->peer_d[r]) cudaIpcCloseMemHandle(s->peer_d[r]);
        if (r < s->peer_alt.size() && s->peer_alt[r]) cudaIpcCloseMemHandle(s->peer_alt[r]);
    }
    s->peer_d.clear();
    s->peer_alt.clear();
    cudaGetLastError();
    return QCB_OK;
}

int qcb_dist_finalize(void)
{
    Ctx &c = ctx();
    if (c.nccl_comm) {
        if (c.inited) cudaStreamSynchronize(c.stream);
        ncclCommDestroy((ncclComm_t)c.nccl_comm);
        c.nccl_comm = nullptr;
    }
    if (c.d_barrier) { cudaFree(c.d_barrier); c.d_barrier = nullptr; }
    c.rank = 0;
    c.nranks = 1;
    return QCB_OK;
}

} // extern "C"
