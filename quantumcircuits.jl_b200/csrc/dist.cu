// dist.cu -- multi-GPU sharding: one process per GPU, amplitudes sharded by the top log2(P) index bits,
// gates on those qubits handled by global<->local qubit swaps = an all-to-all of contiguous blocks over
// NCCL (NVLink 5 / NVSwitch).  No reference counterpart; the schedule and the layout
// bookkeeping are in dist_plan.hpp and are replayed on the CPU by tests/emu with simulated ranks.
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
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) return QCB_OK;
    if (free_b < s->bytes + ((size_t)6 << 30)) return QCB_OK; // keep headroom for scratch and NCCL buffers
    if (cudaMalloc(&s->alt, s->bytes) != cudaSuccess) { s->alt = nullptr; cudaGetLastError(); }
    return QCB_OK;
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

struct GpuShardExec : ShardExecutor {
    qcb_state *s;
    explicit GpuShardExec(qcb_state *st) : s(st) {}
    int run_pass(const PassPlan &pl, const std::vector<GateOp> &gates, const double *mats, bool adjoint) override
    {
        const int rc = launch_plan(s, pl, gates, mats, adjoint);
        if (rc == QCB_OK) { ctx().st_passes += 1; ctx().st_stages += (double)pl.stages.size(); }
        return rc;
    }
    int swap_top() override { return swap_top_device(s); }
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
        return set_error(QCB_EINVAL, "sharded scheduler: %s", e.what());
    }
    s->phys = L.phys;
    if (rc < 0) return set_error(QCB_EINVAL, "sharded scheduler could not make progress (code %d)", rc);
    return rc;
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
        return set_error(QCB_EINVAL, "canonicalize: %s", e.what());
    }
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

int qcb_dist_finalize(void)
{
    Ctx &c = ctx();
    if (c.nccl_comm) {
        if (c.inited) cudaStreamSynchronize(c.stream);
        ncclCommDestroy((ncclComm_t)c.nccl_comm);
        c.nccl_comm = nullptr;
    }
    c.rank = 0;
    c.nranks = 1;
    return QCB_OK;
}

} // extern "C"
