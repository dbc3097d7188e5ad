// api.cu -- the extern "C" surface of libqcb200.so (include/qcb200.h) and the host drivers
// behind it.  Each entry point replaces the reference interface named in the header.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "adjoint_mma.cuh"
#include "gates_hd.cuh"
#include "gates_host.hpp"
#include "kernels.cuh"
#include "flip_tile.cuh"
#include "measure_plan.hpp"
#include "qcb_internal.hpp"

namespace qcb {

static thread_local char g_err[1024] = "";

Ctx &ctx()
{
    static Ctx c;
    return c;
}

int set_error(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}

int ensure_scratch(int slot, size_t bytes)
{
    Ctx &c = ctx();
    if (c.scratch_bytes[slot] >= bytes) return QCB_OK;
    if (c.scratch[slot]) { cudaFree(c.scratch[slot]); c.scratch[slot] = nullptr; c.scratch_bytes[slot] = 0; }
    QCB_CUDA(cudaMalloc(&c.scratch[slot], bytes));
    c.scratch_bytes[slot] = bytes;
    return QCB_OK;
}
int ensure_partial(size_t doubles)
{
    Ctx &c = ctx();
    if (c.partial_cap >= doubles) return QCB_OK;
    if (c.d_partial) cudaFree(c.d_partial);
    c.d_partial = nullptr; c.partial_cap = 0;
    QCB_CUDA(cudaMalloc(&c.d_partial, doubles * sizeof(double)));
    c.partial_cap = doubles;
    return QCB_OK;
}
int ensure_out(size_t doubles)
{
    Ctx &c = ctx();
    if (c.out_cap >= doubles) return QCB_OK;
    if (c.d_out) cudaFree(c.d_out);
    if (c.h_out) cudaFreeHost(c.h_out);
    c.d_out = nullptr; c.h_out = nullptr; c.out_cap = 0;
    QCB_CUDA(cudaMalloc(&c.d_out, doubles * sizeof(double)));
    QCB_CUDA(cudaMallocHost(&c.h_out, doubles * sizeof(double)));
    c.out_cap = doubles;
    return QCB_OK;
}

static unsigned grid_for(unsigned long long items, int threads, int per_sm)
{
    const unsigned long long want = (items + threads - 1) / threads;
    const unsigned long long cap = (unsigned long long)ctx().sm_count * per_sm;
    return (unsigned)std::max<unsigned long long>(1, std::min(want, cap));
}

static int check_state(const qcb_state *s)
{
    if (!s || !s->d) return set_error(QCB_EINVAL, "null state handle");
    return QCB_OK;
}

// ---- measure ------------------------------------------------------------------------
template <typename R, int TB, int MINB, int KIND> static int launch_measure_pass(const qcb_state *s, const MeasParams<R> &P, double2 *partial)
{
    using V = typename Vec2<R>::type;
    auto kern = measure_tile_kernel<R, TB, MINB, KIND>;
    const size_t smem = sizeof(V) << TB;
    static bool attr_set = false;
    if (!attr_set) {
        QCB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    kern<<<(unsigned)(1ull << (s->nloc - TB)), 1 << (TB - kRegBits), smem, ctx().stream>>>((const V *)s->d, P, partial);
    QCB_CUDA(cudaGetLastError());
    count_launch();
    return QCB_OK;
}

// tiled path: ceil((n - TB) / (TB - low_bits)) + 1 reads of the state, flips done in shared memory.
// Sharded states: canonical layout first; the local passes see the rank bits through c_hi; the flips of the qubits that
// live in the rank bits need amplitudes of other ranks, so one qubit swap brings them to the top local bits for one more
// pass; the partial sums are all-reduced over the ranks (2 doubles).
template <typename R> static int measure_tiled(qcb_state *s, const HamEval &H, double *out2)
{
    Ctx &c = ctx();
    constexpr int TB = 12; // 64 KiB (ComplexF64) / 32 KiB (ComplexF32) tiles, 256 threads
    const int g = s->n - s->nloc;
    unsigned long long c_hi = 0;
    if (s->sharded) {
        QCB_TRY(dist_state_canonicalize(s));
        c_hi = (unsigned long long)c.rank << s->nloc;
    }
    const std::vector<MeasParams<R>> passes = plan_measure<R>(s->nloc, TB, (int)c.low_bits, H, c_hi);
    const size_t ntiles = (size_t)1 << (s->nloc - TB);
    const size_t npass = passes.size() + (g > 0 ? 1 : 0);
    QCB_TRY(ensure_partial(2 * ntiles * npass));
    QCB_TRY(ensure_out(64));
    // CTAs per SM the register budget is sized for ("measure_ctas": 2 = 104+ registers, 3 = 80 registers: the kernel is
    // bound by instruction issue and latency, not by HBM, so resident warps matter)
    const long ctas = c.measure_ctas ? c.measure_ctas : (sizeof(R) == 4 ? 3 : 2); // measured at 30 q: ComplexF32 9-15 % faster with 3, ComplexF64 not
    auto launch = [&](const MeasParams<R> &P, size_t slot) -> int {
        double2 *part = (double2 *)c.d_partial + slot * ntiles;
        if (ctas >= 3) return H.kind == 0 ? launch_measure_pass<R, TB, 3, 0>(s, P, part) : launch_measure_pass<R, TB, 3, 1>(s, P, part);
        return H.kind == 0 ? launch_measure_pass<R, TB, 2, 0>(s, P, part) : launch_measure_pass<R, TB, 2, 1>(s, P, part);
    };
    for (size_t p = 0; p < passes.size(); p++) QCB_TRY(launch(passes[p], p));
    if (g > 0) {
        QCB_TRY(dist_swap_top_now(s));
        QCB_TRY(launch(plan_measure_top<R>(s->nloc, g, TB, (int)c.low_bits, H, c_hi), passes.size()));
    }
    k_final_sum<<<2, 256, 0, c.stream>>>(c.d_partial, (int)(ntiles * npass), 2, c.d_out);
    QCB_CUDA(cudaGetLastError());
    count_launch();
    if (g > 0) QCB_TRY(dist_allreduce_sum(c.d_out, 2));
    QCB_CUDA(cudaMemcpyAsync(c.h_out, c.d_out, 2 * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    QCB_CUDA(cudaStreamSynchronize(c.stream));
    const double oc = H.kind == 0 ? -H.p0 * H.p2 : H.p1;
    out2[0] = c.h_out[0] + oc * c.h_out[1];
    out2[1] = 0.0; // the two orientations of every flip pair have exactly opposite imaginary parts
    c.st_passes = (double)npass;
    return QCB_OK;
}

// ---- persistent form of the tiled expectation value (measure_tile_persistent_kernel; option "measure_persistent") ------
template <typename R, int TB, int MINB, int KIND>
static int launch_measure_pass_persistent(const qcb_state *s, const MeasParams<R> &P, double2 *partial, unsigned *grid_out)
{
    using V = typename Vec2<R>::type;
    auto kern = measure_tile_persistent_kernel<R, TB, MINB, KIND>;
    const size_t smem = sizeof(V) << TB;
    static bool attr_set = false;
    static int ctas_per_sm = 1;
    if (!attr_set) {
        QCB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        QCB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, kern, 1 << (TB - kRegBits), smem));
        if (ctas_per_sm < 1) ctas_per_sm = 1;
        attr_set = true;
    }
    const unsigned long long ntiles = 1ull << (s->nloc - TB);
    const unsigned grid = (unsigned)std::min<unsigned long long>(ntiles, (unsigned long long)ctx().sm_count * ctas_per_sm);
    kern<<<grid, 1 << (TB - kRegBits), smem, ctx().stream>>>((const V *)s->d, P, ntiles, partial);
    QCB_CUDA(cudaGetLastError());
    count_launch();
    *grid_out = grid;
    return QCB_OK;
}

template <typename R> static int measure_tiled_persistent(qcb_state *s, const HamEval &H, double *out2)
{
    Ctx &c = ctx();
    constexpr int TB = 12;
    const int g = s->n - s->nloc;
    unsigned long long c_hi = 0;
    if (s->sharded) {
        QCB_TRY(dist_state_canonicalize(s));
        c_hi = (unsigned long long)c.rank << s->nloc;
    }
    const std::vector<MeasParams<R>> passes = plan_measure<R>(s->nloc, TB, (int)c.low_bits, H, c_hi);
    const size_t npass = passes.size() + (g > 0 ? 1 : 0);
    const size_t max_grid = (size_t)c.sm_count * 8; // upper bound of the persistent grid
    QCB_TRY(ensure_partial(2 * max_grid * npass));
    QCB_TRY(ensure_out(64));
    size_t filled = 0; // partial sums written so far (passes are packed back to back)
    auto launch = [&](const MeasParams<R> &P) -> int {
        unsigned grid = 0;
        double2 *part = (double2 *)c.d_partial + filled;
        const int rc = H.kind == 0 ? launch_measure_pass_persistent<R, TB, 2, 0>(s, P, part, &grid) : launch_measure_pass_persistent<R, TB, 2, 1>(s, P, part, &grid);
        filled += grid;
        return rc;
    };
    for (size_t p = 0; p < passes.size(); p++) QCB_TRY(launch(passes[p]));
    if (g > 0) {
        QCB_TRY(dist_swap_top_now(s));
        QCB_TRY(launch(plan_measure_top<R>(s->nloc, g, TB, (int)c.low_bits, H, c_hi)));
    }
    if (filled > max_grid * npass) return set_error(QCB_ECUDA, "internal: persistent grid larger than its bound");
    k_final_sum<<<2, 256, 0, c.stream>>>(c.d_partial, (int)filled, 2, c.d_out);
    QCB_CUDA(cudaGetLastError());
    count_launch();
    if (g > 0) QCB_TRY(dist_allreduce_sum(c.d_out, 2));
    QCB_CUDA(cudaMemcpyAsync(c.h_out, c.d_out, 2 * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    QCB_CUDA(cudaStreamSynchronize(c.stream));
    const double oc = H.kind == 0 ? -H.p0 * H.p2 : H.p1;
    out2[0] = c.h_out[0] + oc * c.h_out[1];
    out2[1] = 0.0;
    c.st_passes = (double)npass;
    return QCB_OK;
}

// ---- TMA-fed persistent form (flip_tile.cuh; default for >= 14 / 15 local qubits) ---------------------------------------
template <typename R, int KIND> static int launch_flip_measure(const FlipMaps &M, const FlipParams &P, double2 *partial, unsigned *grid_out)
{
    auto kern = flip_measure_kernel<R, KIND>;
    const size_t smem = flip_smem_bytes();
    static bool attr_set = false;
    if (!attr_set) {
        QCB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    const unsigned grid = (unsigned)std::min<unsigned long long>(P.n_items, (unsigned long long)ctx().sm_count);
    kern<<<grid, kFlipGroups * kFlipConsumers, smem, ctx().stream>>>(M, P, partial);
    QCB_CUDA(cudaGetLastError());
    count_launch();
    *grid_out = grid;
    return QCB_OK;
}

template <typename R> static int measure_flip(qcb_state *s, const HamEval &H, double *out2)
{
    Ctx &c = ctx();
    const int g = s->n - s->nloc;
    unsigned long long c_hi = 0;
    if (s->sharded) {
        QCB_TRY(dist_state_canonicalize(s));
        c_hi = (unsigned long long)c.rank << s->nloc;
    }
    FlipParams P;
    if (!flip_plan<R>(P, s->nloc, H, c_hi)) return set_error(QCB_EINVAL, "internal: state too small for the TMA expectation kernel");
    flip_fill_diag_table<R>(P);
    FlipMaps M;
    const char *why = "";
    if (flip_make_maps<R>(M, s->d, P, &why)) return set_error(QCB_ECUDA, "tensor map: %s", why);
    const size_t ntiles_top = g > 0 ? ((size_t)1 << (s->nloc - 12)) : 0;
    QCB_TRY(ensure_partial(2 * ((size_t)c.sm_count + ntiles_top)));
    QCB_TRY(ensure_out(64));
    unsigned grid = 0;
    QCB_TRY((H.kind == 0 ? launch_flip_measure<R, 0>(M, P, (double2 *)c.d_partial, &grid) : launch_flip_measure<R, 1>(M, P, (double2 *)c.d_partial, &grid)));
    size_t nparts = grid;
    if (g > 0) {
        // the flips of the qubits that live in the rank bits: one qubit swap brings them to the top local bits for one
        // more (non-persistent) tile pass
        QCB_TRY(dist_swap_top_now(s));
        const MeasParams<R> Pt = plan_measure_top<R>(s->nloc, g, 12, (int)c.low_bits, H, c_hi);
        double2 *part = (double2 *)c.d_partial + nparts;
        QCB_TRY((H.kind == 0 ? launch_measure_pass<R, 12, 2, 0>(s, Pt, part) : launch_measure_pass<R, 12, 2, 1>(s, Pt, part)));
        nparts += ntiles_top;
    }
    k_final_sum<<<2, 256, 0, c.stream>>>(c.d_partial, (int)nparts, 2, c.d_out);
    QCB_CUDA(cudaGetLastError());
    count_launch();
    if (g > 0) QCB_TRY(dist_allreduce_sum(c.d_out, 2));
    QCB_CUDA(cudaMemcpyAsync(c.h_out, c.d_out, 2 * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    QCB_CUDA(cudaStreamSynchronize(c.stream));
    const double oc = H.kind == 0 ? -H.p0 * H.p2 : H.p1;
    out2[0] = c.h_out[0] + oc * c.h_out[1];
    out2[1] = 0.0; // the two orientations of every flip pair have exactly opposite imaginary parts (Hermitian kernel Hamiltonians)
    c.st_passes = (double)P.nlevels + (g > 0 ? 1 : 0);
    return QCB_OK;
}

static int measure_impl(const qcb_state *s, int kind, double p0, double p1, double p2, double *out2)
{
    QCB_REQUIRE_INIT();
    QCB_TRY(check_state(s));
    Ctx &c = ctx();
    HamEval H{kind, s->n, p0, p1, p2};
    if (s->sharded || (!c.force_simple && s->nloc >= 13)) {
        // (a sharded state may be re-laid-out: the handle is logically const, the amplitudes it represents do not change)
        qcb_state *ms = const_cast<qcb_state *>(s);
        if (c.measure_tma && s->nloc >= (s->dtype == QCB_C128 ? 14 : 15))
            return s->dtype == QCB_C128 ? measure_flip<double>(ms, H, out2) : measure_flip<float>(ms, H, out2);
        if (c.measure_persistent) return s->dtype == QCB_C128 ? measure_tiled_persistent<double>(ms, H, out2) : measure_tiled_persistent<float>(ms, H, out2);
        return s->dtype == QCB_C128 ? measure_tiled<double>(ms, H, out2) : measure_tiled<float>(ms, H, out2);
    }
    const unsigned long long namps = 1ull << s->nloc;
    const unsigned blocks = grid_for(namps, kRedThreads, 8);
    QCB_TRY(ensure_partial((size_t)blocks * 2));
    QCB_TRY(ensure_out(64));
    if (s->dtype == QCB_C128)
        k_measure<double><<<blocks, kRedThreads, 0, c.stream>>>((const double2 *)s->d, namps, H, (double2 *)c.d_partial);
    else
        k_measure<float><<<blocks, kRedThreads, 0, c.stream>>>((const float2 *)s->d, namps, H, (double2 *)c.d_partial);
    QCB_CUDA(cudaGetLastError());
    k_final_sum<<<2, 256, 0, c.stream>>>(c.d_partial, (int)blocks, 2, c.d_out);
    QCB_CUDA(cudaGetLastError());
    count_launch(2);
    QCB_CUDA(cudaMemcpyAsync(c.h_out, c.d_out, 2 * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    QCB_CUDA(cudaStreamSynchronize(c.stream));
    out2[0] = c.h_out[0];
    out2[1] = c.h_out[1];
    return QCB_OK;
}

// ---- brickwork helpers ----------------------------------------------------------------
static int brickwork_gates(int nbits, int nlayers, const double *angles, int first, int last, bool reverse,
                           std::vector<GateOp> &gates, std::vector<double> &mats)
{
    if (nbits < 2) return set_error(QCB_EINVAL, "a brickwork circuit needs at least 2 qubits");
    const std::vector<int> pos = brickwork_positions(nbits, nlayers);
    const int G = (int)pos.size();
    if (first < 0 || last > G || first > last) return set_error(QCB_EINVAL, "gate range [%d, %d) outside [0, %d)", first, last, G);
    const int m = last - first;
    gates.resize(m);
    mats.resize((size_t)32 * m);
    for (int i = 0; i < m; i++) {
        const int g = reverse ? (last - 1 - i) : (first + i);
        to_doubles(general_unitary(angles + 15 * (size_t)g), mats.data() + 32 * (size_t)i);
        gates[i] = GateOp{pos[g], i};
    }
    return QCB_OK;
}

static int copy_state_raw(void *dst, const void *src, size_t bytes)
{
    QCB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, ctx().stream));
    return QCB_OK;
}

// An explicit Hamiltonian in CSR form, uploaded for the duration of one call (measure / gradients with H::AbstractMatrix)
struct DeviceCsr {
    char *base = nullptr;
    long long nrows = 0;
    long long *rowptr = nullptr, *colidx = nullptr;
    double2 *vals = nullptr;
    ~DeviceCsr() { if (base) cudaFree(base); }
};
static int upload_csr(DeviceCsr &D, int nqubits, long long nrows, const long long *rowptr, const long long *colidx, const double *vals_reim)
{
    if (!rowptr) return set_error(QCB_EINVAL, "null pointer");
    if (nrows != (1ll << nqubits)) return set_error(QCB_EINVAL, "H must be 2^n x 2^n with n = %d", nqubits);
    const long long nnz = rowptr[nrows];
    if (rowptr[0] != 0 || nnz < 0 || (nnz > 0 && (!colidx || !vals_reim))) return set_error(QCB_EINVAL, "bad CSR arrays");
    for (long long r = 0; r < nrows; r++)
        if (rowptr[r + 1] < rowptr[r]) return set_error(QCB_EINVAL, "rowptr must be non-decreasing");
    for (long long k = 0; k < nnz; k++)
        if (colidx[k] < 0 || colidx[k] >= nrows) return set_error(QCB_EINVAL, "column index out of range");
    Ctx &c = ctx();
    const size_t b_ptr = sizeof(long long) * (size_t)(nrows + 1), b_col = sizeof(long long) * (size_t)nnz, b_val = 16 * (size_t)nnz;
    QCB_CUDA(cudaMalloc(&D.base, b_ptr + b_col + b_val + 64));
    D.nrows = nrows;
    D.rowptr = (long long *)D.base;
    D.colidx = (long long *)(D.base + b_ptr);
    D.vals = (double2 *)(D.base + b_ptr + b_col + ((16 - (b_ptr + b_col) % 16) % 16));
    QCB_CUDA(cudaMemcpyAsync(D.rowptr, rowptr, b_ptr, cudaMemcpyHostToDevice, c.stream));
    if (nnz > 0) {
        QCB_CUDA(cudaMemcpyAsync(D.colidx, colidx, b_col, cudaMemcpyHostToDevice, c.stream));
        QCB_CUDA(cudaMemcpyAsync(D.vals, vals_reim, b_val, cudaMemcpyHostToDevice, c.stream));
    }
    return QCB_OK;
}

static int measure_csr_device(const qcb_state *s, const DeviceCsr &D, double *out_E2)
{
    Ctx &c = ctx();
    const unsigned blocks = grid_for((unsigned long long)D.nrows, kRedThreads, 8);
    QCB_TRY(ensure_partial((size_t)blocks * 2));
    QCB_TRY(ensure_out(64));
    if (s->dtype == QCB_C128) k_measure_csr<double><<<blocks, kRedThreads, 0, c.stream>>>((const double2 *)s->d, D.nrows, D.rowptr, D.colidx, D.vals, (double2 *)c.d_partial);
    else k_measure_csr<float><<<blocks, kRedThreads, 0, c.stream>>>((const float2 *)s->d, D.nrows, D.rowptr, D.colidx, D.vals, (double2 *)c.d_partial);
    QCB_CUDA(cudaGetLastError());
    k_final_sum<<<2, 256, 0, c.stream>>>(c.d_partial, (int)blocks, 2, c.d_out);
    QCB_CUDA(cudaGetLastError());
    count_launch(2);
    QCB_CUDA(cudaMemcpyAsync(c.h_out, c.d_out, 2 * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    QCB_CUDA(cudaStreamSynchronize(c.stream));
    out_E2[0] = c.h_out[0];
    out_E2[1] = c.h_out[1];
    return QCB_OK;
}

// Result area of a gradient on the device (c.d_out) and its pinned mirror (c.h_out), in doubles:
//   [ environments 32 G | energy parts 4 | gradients 15 G | angles 15 G ]
// The angles go up first (the stream is idle then, so the staging of a pageable source costs nothing), the backward sweep
// fills the environments, k_contract_env (gates_hd.cuh) turns them into the gradients next to them, and one download of
// [energy parts | gradients] with one synchronisation ends the call.
struct GradArea {
    size_t G;
    size_t env() const { return 0; }
    size_t energy() const { return 32 * G; }
    size_t grads() const { return 32 * G + 4; }
    size_t angles() const { return 32 * G + 4 + 15 * G; }
    size_t total() const { return 32 * G + 4 + 30 * G; }
};
static int grad_area_begin(const GradArea &ar, const double *angles)
{
    Ctx &c = ctx();
    QCB_TRY(ensure_out(ar.total()));
    QCB_CUDA(cudaMemcpyAsync(c.d_out + ar.angles(), angles, sizeof(double) * 15 * ar.G, cudaMemcpyHostToDevice, c.stream));
    QCB_CUDA(cudaMemsetAsync(c.d_out + ar.energy(), 0, 4 * sizeof(double), c.stream));
    return QCB_OK;
}
// grads[k, g] = 2 Re sum_{r,c} env_g[r,c] dU_g/dtheta_k [r,c] on the device, download, energy = sum of its two parts
static int grad_area_finish(const GradArea &ar, bool env_post, int want_energy, double *out_E, double *out_grads)
{
    Ctx &c = ctx();
    if (c.contract_coop)
        k_contract_env_coop<<<(unsigned)((ar.G + 7) / 8), 128, 0, c.stream>>>(c.d_out + ar.angles(), c.d_out + ar.env(), (int)ar.G, env_post ? 1 : 0, c.d_out + ar.grads());
    else
        k_contract_env<<<(unsigned)((4 * ar.G + 63) / 64), 64, 0, c.stream>>>(c.d_out + ar.angles(), c.d_out + ar.env(), (int)ar.G, env_post ? 1 : 0, c.d_out + ar.grads());
    QCB_CUDA(cudaGetLastError());
    count_launch();
    QCB_CUDA(cudaMemcpyAsync(c.h_out + ar.energy(), c.d_out + ar.energy(), sizeof(double) * (4 + 15 * ar.G), cudaMemcpyDeviceToHost, c.stream));
    QCB_CUDA(cudaStreamSynchronize(c.stream));
    if (want_energy) *out_E = c.h_out[ar.energy()] + c.h_out[ar.energy() + 2];
    std::memcpy(out_grads, c.h_out + ar.grads(), sizeof(double) * 15 * ar.G);
    return QCB_OK;
}

// Sharded states (collective call): forward circuit with qubit swaps, lambda = H psi in two sweeps around one swap (the
// flips of the qubits that live in the rank bits pair amplitudes of different ranks; in the layout the forward circuit
// ended in), backward sweep on the pair
// (dist.cu: PairShardExec), one all-reduce of the per-rank environments and energy parts.
static int gradients_sharded_impl(const qcb_state *psi0, int nbits, int nlayers, const double *angles, int ham_kind, const double *hp,
                                  int want_energy, double *out_E, double *out_grads)
{
    Ctx &c = ctx();
    if (psi0->nloc < kMinTB) return set_error(QCB_EUNSUPPORTED, "sharded gradients need at least 2^%d amplitudes per rank", kMinTB);
    if (nbits > 40) return set_error(QCB_EINVAL, "too many qubits");
    const std::vector<int> pos = brickwork_positions(nbits, nlayers);
    const int G = (int)pos.size();
    const int g = psi0->n - psi0->nloc;
    QCB_TRY(ensure_scratch(0, psi0->bytes));
    QCB_TRY(ensure_scratch(1, psi0->bytes));
    // a third buffer lets the qubit swaps of the pair run out of place (NCCL straight into it, then the buffers rotate);
    // without the memory for it they go through the bounded staging buffer in place
    {
        size_t free_b = 0, total_b = 0;
        if (c.scratch_bytes[3] < psi0->bytes && cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && free_b > psi0->bytes + ((size_t)6 << 30)) {
            if (ensure_scratch(3, psi0->bytes) != QCB_OK) cudaGetLastError();
        }
    }
    void *spare = c.scratch_bytes[3] >= psi0->bytes ? c.scratch[3] : nullptr;
    struct Buf { void *p; size_t bytes; } pool[3] = {{c.scratch[0], c.scratch_bytes[0]}, {c.scratch[1], c.scratch_bytes[1]}, {spare, spare ? c.scratch_bytes[3] : 0}};
    qcb_state A = *psi0, B = *psi0;
    A.d = c.scratch[0]; A.owned = false; A.alt = nullptr; A.peer_d.clear(); A.peer_alt.clear();
    B.d = c.scratch[1]; B.owned = false; B.alt = nullptr; B.peer_d.clear(); B.peer_alt.clear();
    // (whatever happens below, the scratch slots must own the three buffers again, in their rotated order)
    struct Restore {
        Ctx &c; qcb_state &A, &B; void *&spare; Buf (&pool)[3];
        ~Restore()
        {
            auto bytes_of = [&](void *p) { for (const Buf &b : pool) if (b.p == p) return b.bytes; return (size_t)0; };
            c.scratch[0] = A.d; c.scratch_bytes[0] = bytes_of(A.d);
            c.scratch[1] = B.d; c.scratch_bytes[1] = bytes_of(B.d);
            if (spare) { c.scratch[3] = spare; c.scratch_bytes[3] = bytes_of(spare); }
        }
    } restore{c, A, B, spare, pool};
    QCB_TRY(copy_state_raw(A.d, psi0->d, psi0->bytes));
    std::vector<GateOp> gates;
    std::vector<double> mats;
    QCB_TRY(brickwork_gates(nbits, nlayers, angles, 0, G, false, gates, mats));
    QCB_TRY(run_gate_list(&A, gates, mats.data(), false));
    if (G == 0) {
        if (want_energy) {
            double E2[2];
            QCB_TRY(measure_impl(&A, ham_kind, hp[0], hp[1], hp[2], E2));
            *out_E = E2[0];
        }
        return QCB_OK;
    }
    // (no canonicalisation: the two sweeps below work in any layout -- every logical qubit is local either now or after swap_top)
    B.phys = A.phys;
    const unsigned long long namps = 1ull << A.nloc;
    HamEval H{ham_kind, nbits, hp[0], hp[1], hp[2]};
    const unsigned blocks = grid_for(namps, kRedThreads, 8);
    QCB_TRY(ensure_partial((size_t)blocks * 2));
    const GradArea ar{(size_t)G};
    QCB_TRY(grad_area_begin(ar, angles));
    auto sweep = [&](unsigned long long mask, int first, int slot) -> int {
        ShardMap M;
        std::memset(&M, 0, sizeof M);
        M.n = nbits; M.nloc = A.nloc; M.hi = (unsigned long long)c.rank << A.nloc;
        for (int q = 0; q < nbits; q++) {
            M.phys[q] = (unsigned char)A.phys[q];
            if (((mask >> q) & 1ull) && A.phys[q] >= A.nloc) return set_error(QCB_EINVAL, "internal: flip of a qubit that is not local");
        }
        if (A.dtype == QCB_C128)
            k_apply_ham_sharded<double><<<blocks, kRedThreads, 0, c.stream>>>((double2 *)B.d, (const double2 *)A.d, namps, M, H, mask, first, (double2 *)c.d_partial);
        else
            k_apply_ham_sharded<float><<<blocks, kRedThreads, 0, c.stream>>>((float2 *)B.d, (const float2 *)A.d, namps, M, H, mask, first, (double2 *)c.d_partial);
        QCB_CUDA(cudaGetLastError());
        k_final_sum<<<2, 256, 0, c.stream>>>(c.d_partial, (int)blocks, 2, c.d_out + 32 * (size_t)G + 2 * slot);
        QCB_CUDA(cudaGetLastError());
        count_launch(2);
        return QCB_OK;
    };
    unsigned long long local_mask = 0;
    for (int q = 0; q < nbits; q++)
        if (A.phys[q] < A.nloc) local_mask |= 1ull << q;
    QCB_TRY(sweep(local_mask, 1, 0));
    if (g > 0) {
        QCB_TRY(dist_swap_top_pair(&A, &B, &spare));
        const unsigned long long all = nbits >= 64 ? ~0ull : ((1ull << nbits) - 1ull);
        QCB_TRY(sweep(all & ~local_mask, 0, 1));
    }
    std::vector<GateOp> rev(G);
    for (int i = 0; i < G; i++) rev[i] = GateOp{pos[G - 1 - i], G - 1 - i};
    QCB_TRY(dist_run_adjoint_pair(&A, &B, &spare, rev, mats.data(), c.d_out));
    QCB_TRY(dist_allreduce_sum(c.d_out, 32 * G + 4)); // environments and energy parts of all ranks
    return grad_area_finish(ar, adjoint_uses_mma(&A), want_energy, out_E, out_grads);
}

// adjoint-method gradient; numbers equal the reference's shift sweep.
// csr != nullptr: explicit Hamiltonian (ham_kind / hp unused)
static int gradients_impl(const qcb_state *psi0, int nbits, int nlayers, const double *angles, int ham_kind, const double *hp,
                          int want_energy, double *out_E, double *out_grads, const DeviceCsr *csr = nullptr)
{
    QCB_REQUIRE_INIT();
    QCB_TRY(check_state(psi0));
    if (psi0->n != nbits) return set_error(QCB_EINVAL, "circuit has %d qubits, state has %d", nbits, psi0->n);
    if (!csr && ham_kind != QCB_HAM_TFIM && ham_kind != QCB_HAM_EAST) return set_error(QCB_EINVAL, "unknown Hamiltonian kind %d", ham_kind);
    if (psi0->sharded && csr) return set_error(QCB_EUNSUPPORTED, "matrix Hamiltonians are not supported on sharded states");
    if (psi0->sharded) return gradients_sharded_impl(psi0, nbits, nlayers, angles, ham_kind, hp, want_energy, out_E, out_grads);
    Ctx &c = ctx();
    const std::vector<int> pos = brickwork_positions(nbits, nlayers);
    const int G = (int)pos.size();
    QCB_TRY(ensure_scratch(0, psi0->bytes));
    QCB_TRY(ensure_scratch(1, psi0->bytes));
    qcb_state A = *psi0, B = *psi0;
    A.d = c.scratch[0]; A.owned = false;
    B.d = c.scratch[1]; B.owned = false;
    std::vector<GateOp> gates;
    std::vector<double> mats;
    QCB_TRY(brickwork_gates(nbits, nlayers, angles, 0, G, false, gates, mats));
    if (!csr && small_gradient_eligible(psi0, G)) {
        // small states: forward circuit, lambda = H psi, backward sweep and the environments in ONE cluster launch
        // (small_circuit.cuh), then the contraction with dU/dtheta
        const GradArea ar{(size_t)G};
        QCB_TRY(grad_area_begin(ar, angles));
        c.st_launches = 0;
        HamEval H{ham_kind, nbits, hp[0], hp[1], hp[2]};
        QCB_TRY(run_small_gradient(psi0, gates, mats.data(), H, c.d_out + ar.env(), c.d_out + ar.energy()));
        QCB_TRY(grad_area_finish(ar, false, want_energy, out_E, out_grads));
        c.st_passes = 1; c.st_stages = 0; c.st_gates = 4.0 * G; c.st_fw_passes = 1; c.st_bw_passes = 0;
        return QCB_OK;
    }
    // psi .= psi0 (gradients.jl:90) and the forward pass (:92-101) in one: the first fused pass reads psi0 and writes psi
    QCB_TRY(run_gate_list(&A, gates, mats.data(), false, psi0->d));
    const double fw_passes = c.st_passes, fw_stages = c.st_stages, fw_launches = c.st_launches;
    if (G == 0) {
        if (want_energy) {
            double E2[2];
            if (csr) QCB_TRY(measure_csr_device(&A, *csr, E2));
            else QCB_TRY(measure_impl(&A, ham_kind, hp[0], hp[1], hp[2], E2)); // gradients.jl:104
            *out_E = E2[0];
        }
        return QCB_OK;
    }
    const GradArea ar{(size_t)G};
    QCB_TRY(grad_area_begin(ar, angles));
    // lambda = H psi_G, and E = <psi_G|lambda> from the same sweep (gradients.jl:104)
    const unsigned long long namps = 1ull << nbits;
    HamEval H{ham_kind, nbits, csr ? 0.0 : hp[0], csr ? 0.0 : hp[1], csr ? 0.0 : hp[2]};
    {
        // ham_tile: 0 = plain gather, 1 = low-bit tile in shared memory for ComplexF32 (default), 2 = for both types (measured slower)
        const bool f64 = psi0->dtype == QCB_C128;
        const bool tiled = !csr && nbits >= 13 && c.ham_tile >= (f64 ? 2 : 1);
        const unsigned blocks = tiled ? grid_for((namps >> (f64 ? 11 : 12)) * kRedThreads, kRedThreads, 6) : grid_for(namps, kRedThreads, 8);
        QCB_TRY(ensure_partial((size_t)blocks * 2));
        if (csr && psi0->dtype == QCB_C128)
            k_apply_csr<double><<<blocks, kRedThreads, 0, c.stream>>>((double2 *)B.d, (const double2 *)A.d, csr->nrows, csr->rowptr, csr->colidx, csr->vals, (double2 *)c.d_partial);
        else if (csr)
            k_apply_csr<float><<<blocks, kRedThreads, 0, c.stream>>>((float2 *)B.d, (const float2 *)A.d, csr->nrows, csr->rowptr, csr->colidx, csr->vals, (double2 *)c.d_partial);
        else if (tiled && f64)
            k_apply_ham_tiled<double, 11><<<blocks, kRedThreads, 0, c.stream>>>((double2 *)B.d, (const double2 *)A.d, namps >> 11, H, (double2 *)c.d_partial);
        else if (tiled)
            k_apply_ham_tiled<float, 12><<<blocks, kRedThreads, 0, c.stream>>>((float2 *)B.d, (const float2 *)A.d, namps >> 12, H, (double2 *)c.d_partial);
        else if (psi0->dtype == QCB_C128)
            k_apply_ham<double><<<blocks, kRedThreads, 0, c.stream>>>((double2 *)B.d, (const double2 *)A.d, namps, H, (double2 *)c.d_partial);
        else
            k_apply_ham<float><<<blocks, kRedThreads, 0, c.stream>>>((float2 *)B.d, (const float2 *)A.d, namps, H, (double2 *)c.d_partial);
        QCB_CUDA(cudaGetLastError());
        count_launch();
        if (want_energy) {
            k_final_sum<<<2, 256, 0, c.stream>>>(c.d_partial, (int)blocks, 2, c.d_out + ar.energy());
            QCB_CUDA(cudaGetLastError());
            count_launch();
        }
    }
    // backward sweep: un-apply gate g on psi and lambda, reduce the 4x4 environment
    double bw_passes = 0, bw_stages = 0;
    bool env_post = false;
    if (!c.force_simple && nbits >= kMinTB) {
        // fused: the reversed circuit goes through the pass planner, one launch per trapezoid of gates
        std::vector<GateOp> rev(G);
        for (int i = 0; i < G; i++) rev[i] = GateOp{pos[G - 1 - i], G - 1 - i};
        const double p0 = c.st_passes, s0 = c.st_stages;
        QCB_TRY(run_adjoint_fused(&A, &B, rev, mats.data(), c.d_out, &env_post));
        bw_passes = c.st_passes - p0; bw_stages = c.st_stages - s0;
    } else {
        const unsigned long long nquads = namps >> 2;
        const unsigned blocks = grid_for(nquads, kRedThreads, 4);
        QCB_TRY(ensure_partial((size_t)blocks * 32));
        for (int g = G - 1; g >= 0; g--) {
            const double *m = mats.data() + 32 * (size_t)g;
            if (psi0->dtype == QCB_C128) {
                GateMat<double> Ud;
                for (int r = 0; r < 4; r++)
                    for (int cc = 0; cc < 4; cc++) gate_set(Ud, r, cc, m[2 * (cc + 4 * r)], -m[2 * (cc + 4 * r) + 1]);
                k_adjoint_step<double><<<blocks, kRedThreads, 0, c.stream>>>((double2 *)A.d, (double2 *)B.d, nquads, pos[g], Ud, c.d_partial);
            } else {
                GateMat<float> Ud;
                for (int r = 0; r < 4; r++)
                    for (int cc = 0; cc < 4; cc++) gate_set(Ud, r, cc, m[2 * (cc + 4 * r)], -m[2 * (cc + 4 * r) + 1]);
                k_adjoint_step<float><<<blocks, kRedThreads, 0, c.stream>>>((float2 *)A.d, (float2 *)B.d, nquads, pos[g], Ud, c.d_partial);
            }
            QCB_CUDA(cudaGetLastError());
            k_final_sum<<<32, 128, 0, c.stream>>>(c.d_partial, (int)blocks, 32, c.d_out + 32 * (size_t)g);
            QCB_CUDA(cudaGetLastError());
            count_launch(2);
        }
        bw_passes = G;
    }
    QCB_TRY(grad_area_finish(ar, env_post, want_energy, out_E, out_grads));
    c.st_passes = fw_passes + bw_passes; c.st_stages = fw_stages + bw_stages; c.st_launches += fw_launches; c.st_gates = 4.0 * G;
    c.st_fw_passes = fw_passes; c.st_bw_passes = bw_passes;
    return QCB_OK;
}

} // namespace qcb

using namespace qcb;

extern "C" {

const char *qcb_last_error(void) { return g_err; }
const char *qcb_version(void) { return "qcb200 0.1.0 (sm_100a)"; }

int qcb_device_count(int *out)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { *out = 0; return set_error(QCB_ECUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e)); }
    *out = n;
    return QCB_OK;
}

int qcb_init(int device)
{
    Ctx &c = ctx();
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return set_error(QCB_ECUDA, "no CUDA device available (%s); this backend has no CPU fallback",
                         e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    if (device < 0 || device >= n) return set_error(QCB_EINVAL, "device %d out of range [0, %d)", device, n);
    if (c.inited && c.device == device) return QCB_OK;
    if (c.inited) qcb_finalize();
    QCB_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    QCB_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return set_error(QCB_ECUDA, "device %d is sm_%d%d; libqcb200 is built for sm_100a (B200) only", device, prop.major, prop.minor);
    c.sm_count = prop.multiProcessorCount;
    c.stream = nullptr; // the CUDA default stream, shared with the host framework unless qcb_set_stream is called
    QCB_CUDA(cudaEventCreate(&c.ev0));
    QCB_CUDA(cudaEventCreate(&c.ev1));
    QCB_CUDA(cudaStreamCreateWithFlags(&c.copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; i++)
        for (int k = 0; k < 16; k++) QCB_CUDA(cudaEventCreateWithFlags(&c.chunk_ev[i][k], cudaEventDisableTiming));
    c.device = device;
    c.inited = true;
    return QCB_OK;
}

int qcb_finalize(void)
{
    Ctx &c = ctx();
    if (!c.inited) return QCB_OK;
    cudaStreamSynchronize(c.stream);
    for (int i = 0; i < 4; i++) { if (c.scratch[i]) cudaFree(c.scratch[i]); c.scratch[i] = nullptr; c.scratch_bytes[i] = 0; }
    if (c.d_partial) cudaFree(c.d_partial);
    if (c.d_out) cudaFree(c.d_out);
    if (c.h_out) cudaFreeHost(c.h_out);
    c.d_partial = c.d_out = c.h_out = nullptr;
    c.partial_cap = c.out_cap = 0;
    if (c.h_gates) cudaFreeHost(c.h_gates);
    if (c.d_gates) cudaFree(c.d_gates);
    if (c.gates_ev) cudaEventDestroy(c.gates_ev);
    c.h_gates = c.d_gates = nullptr; c.gates_cap = 0; c.gates_ev = nullptr;
    cudaEventDestroy(c.ev0); cudaEventDestroy(c.ev1);
    for (int i = 0; i < 2; i++)
        for (int k = 0; k < 16; k++) if (c.chunk_ev[i][k]) { cudaEventDestroy(c.chunk_ev[i][k]); c.chunk_ev[i][k] = nullptr; }
    if (c.copy_stream) { cudaStreamDestroy(c.copy_stream); c.copy_stream = nullptr; }
    for (cudaEvent_t e : c.pass_ev) cudaEventDestroy(e);
    c.pass_ev.clear(); c.pass_ev_used = 0;
    c.stream = nullptr;
    c.plan_cache.clear();
    c.inited = false;
    return QCB_OK;
}

int qcb_trim(void)
{
    Ctx &c = ctx();
    if (!c.inited) return QCB_OK;
    QCB_CUDA(cudaStreamSynchronize(c.stream));
    QCB_CUDA(cudaStreamSynchronize(c.copy_stream));
    for (int i = 0; i < 4; i++) {
        if (c.scratch[i]) cudaFree(c.scratch[i]);
        c.scratch[i] = nullptr;
        c.scratch_bytes[i] = 0;
    }
    if (c.d_partial) cudaFree(c.d_partial);
    c.d_partial = nullptr;
    c.partial_cap = 0;
    return QCB_OK;
}

int qcb_sync(void)
{
    QCB_REQUIRE_INIT();
    QCB_CUDA(cudaStreamSynchronize(ctx().stream));
    return QCB_OK;
}

int qcb_set_stream(void *cuda_stream)
{
    QCB_REQUIRE_INIT();
    Ctx &c = ctx();
    QCB_CUDA(cudaStreamSynchronize(c.stream));
    c.stream = (cudaStream_t)cuda_stream; // NULL = the CUDA default stream
    return QCB_OK;
}

int qcb_timer_start(void)
{
    QCB_REQUIRE_INIT();
    QCB_CUDA(cudaEventRecord(ctx().ev0, ctx().stream));
    return QCB_OK;
}
int qcb_timer_stop(float *out_ms)
{
    QCB_REQUIRE_INIT();
    QCB_CUDA(cudaEventRecord(ctx().ev1, ctx().stream));
    QCB_CUDA(cudaEventSynchronize(ctx().ev1));
    QCB_CUDA(cudaEventElapsedTime(out_ms, ctx().ev0, ctx().ev1));
    return QCB_OK;
}

int qcb_set_option(const char *key, long value)
{
    Ctx &c = ctx();
    const std::string k(key ? key : "");
    if (k == "tile_bits") {
        if (value < kMinTB || value > kMaxTB) return set_error(QCB_EINVAL, "tile_bits must be in [%d, %d]", kMinTB, kMaxTB);
        c.tile_bits = value;
    } else if (k == "low_bits") {
        if (value < 0 || value > 6) return set_error(QCB_EINVAL, "low_bits must be in [0, 6]");
        c.low_bits = value;
    } else if (k == "max_gates_per_pass") {
        if (value < 1) return set_error(QCB_EINVAL, "max_gates_per_pass must be >= 1");
        c.max_gates_per_pass = value;
    } else if (k == "adjoint_tile_bits") {
        if (value != 0 && (value < kMinTB || value > 12)) return set_error(QCB_EINVAL, "adjoint_tile_bits must be 0 (auto) or in [%d, 12]", kMinTB);
        c.adjoint_tile_bits = value;
    } else if (k == "adjoint_defer") {
        c.adjoint_defer = value ? 1 : 0;
    } else if (k == "adjoint_ctas") {
        c.adjoint_ctas = value;
    } else if (k == "adjoint_f32") {
        c.adjoint_f32 = value ? 1 : 0;
    } else if (k == "ham_tile") {
        // lambda = H psi of the gradient: flips of the low index bits from shared memory (k_apply_ham_tiled) or the plain gather
        if (value < 0 || value > 2) return set_error(QCB_EINVAL, "ham_tile must be 0 (gather), 1 (low-bit tile for ComplexF32) or 2 (for both types)");
        c.ham_tile = value;
    } else if (k == "adjoint_mma") {
        c.adjoint_mma = value ? 1 : 0;
    } else if (k == "pass_tma") {
        if (value < 0 || value > 2) return set_error(QCB_EINVAL, "pass_tma must be 0 (off), 1 (every eligible pass) or 2 (HBM-bound passes only)");
        c.pass_tma = value;
    } else if (k == "pass_tma_max_stages") {
        c.pass_tma_max_stages = value;
    } else if (k == "small_cluster") {
        // gate lists on states of at most this many qubits run in one launch on a thread-block cluster (small_circuit.cuh);
        // 0 = never, 1 = up to the kernel's maximum (15 qubits ComplexF64 / 16 ComplexF32).  Default 12: above that the
        // distributed-shared-memory traffic of the gates on the CTA-rank bits makes it slower than the tile passes (measured)
        if (value < 0 || value > 16) return set_error(QCB_EINVAL, "small_cluster must be 0 (off), 1 (kernel maximum) or a qubit count <= 16");
        c.small_cluster = value;
    } else if (k == "small_gbits") {
        c.small_gbits = value; // cluster size of the small-state kernels: at most 2^value CTAs (-1: by state size)
    } else if (k == "pass_async") {
        c.pass_async = value ? 1 : 0; // gate passes on the persistent kernel with asynchronous tile loads (tile_async.cuh)
    } else if (k == "f64_3m_min_gates") {
        c.f64_3m_min_gates = value; // ComplexF64 passes with at least this many gates use the 3M arithmetic (0: all, large: none)
    } else if (k == "tma_l2_promotion") {
        if (value < 0 || value > 2) return set_error(QCB_EINVAL, "tma_l2_promotion must be 0 (none), 1 (128 bytes) or 2 (256 bytes)");
        tma_strided_promotion() = (int)value;
    } else if (k == "measure_tma") {
        c.measure_tma = value ? 1 : 0;
    } else if (k == "measure_persistent") {
        c.measure_persistent = value ? 1 : 0;
    } else if (k == "contract_coop") {
        c.contract_coop = value ? 1 : 0;
    } else if (k == "measure_ctas") {
        if (value != 0 && value != 2 && value != 3) return set_error(QCB_EINVAL, "measure_ctas must be 0 (auto), 2 or 3");
        c.measure_ctas = value;
    } else if (k == "adjoint_prog_host") {
        c.adjoint_prog_host = value ? 1 : 0;
    } else if (k == "adjoint_warps") {
        if (value != 0 && value != 4 && value != 8) return set_error(QCB_EINVAL, "adjoint_warps must be 0 (auto), 4 or 8");
        c.adjoint_warps = value;
    } else if (k == "time_passes") {
        c.time_passes = value ? 1 : 0;
        c.pass_ev_used = 0;
        c.pass_meta.clear();
        return QCB_OK; // (no effect on plans)
    } else if (k == "fuse_swaps") {
        c.fuse_swaps = value ? 1 : 0;
    } else if (k == "peer_swap") {
        c.peer_swap = value ? 1 : 0; // in-place qubit swaps over NVLink peer memory when a shard has no second buffer (else: staged NCCL)
    } else if (k == "dist_second_buffer") {
        c.dist_second_buffer = value ? 1 : 0; // 0: sharded states created from now on get no second buffer (tests of the in-place swap paths)
    } else if (k == "force_simple") {
        c.force_simple = value ? 1 : 0;
    } else
        return set_error(QCB_EINVAL, "unknown option '%s'", k.c_str());
    c.plan_cache.clear();
    return QCB_OK;
}

int qcb_get_stat(const char *key, double *out)
{
    Ctx &c = ctx();
    const std::string k(key ? key : "");
    if (k == "passes") *out = c.st_passes;
    else if (k == "stages") *out = c.st_stages;
    else if (k == "gates_3m") *out = c.st_gates_3m;
    else if (k == "async_passes") *out = c.st_async_passes;
    else if (k == "cluster_circuits") *out = c.st_cluster_circuits;
    else if (k == "peer_swaps") *out = c.st_peer_swaps;
    else if (k == "launches") *out = c.st_launches;
    else if (k == "gates") *out = c.st_gates;
    else if (k == "launches_total") *out = c.st_launches_total;
    else if (k == "swaps") *out = c.st_swaps;
    else if (k == "fused_swaps") *out = c.st_fused_swaps;
    else if (k == "sm_count") *out = c.sm_count;
    else if (k == "tma_passes") *out = c.st_tma_passes;
    else if (k == "fw_passes") *out = c.st_fw_passes;
    else if (k == "bw_passes") *out = c.st_bw_passes;
    else if (k == "ham_passes") *out = c.st_ham_passes;
    else if (k == "pass_count") *out = (double)(c.pass_ev_used / 2);
    else if (k == "pass_ms") {
        // sum of the event-timed durations of the gate-pass launches since the last read; resets the list
        double tot = 0;
        if (c.inited && c.pass_ev_used) {
            cudaStreamSynchronize(c.stream);
            const char *log = std::getenv("QCB_PASS_LOG");
            FILE *f = log && c.pass_meta.size() == c.pass_ev_used / 2 ? std::fopen(log, "a") : nullptr;
            for (size_t i = 0; i + 1 < c.pass_ev_used; i += 2) {
                float ms = 0;
                if (cudaEventElapsedTime(&ms, c.pass_ev[i], c.pass_ev[i + 1]) == cudaSuccess) tot += ms;
                if (f) {
                    const Ctx::PassMeta &m = c.pass_meta[i / 2];
                    std::fprintf(f, "%zu %.4f stages=%d gates=%d window_bit=%d tma=%d masks=%llx\n", i / 2, ms, m.stages, m.gates, m.first_window_bit, m.tma, m.masks);
                }
            }
            if (f) std::fclose(f);
        }
        c.pass_ev_used = 0;
        c.pass_meta.clear();
        *out = tot;
    }
    else return set_error(QCB_EINVAL, "unknown stat '%s'", k.c_str());
    return QCB_OK;
}

// ---- states -----------------------------------------------------------------------------
static int new_state(int nqubits, int dtype, void *wrap, bool sharded, qcb_state **out)
{
    QCB_REQUIRE_INIT();
    if (!out) return set_error(QCB_EINVAL, "null output pointer");
    if (dtype != QCB_C64 && dtype != QCB_C128) return set_error(QCB_EINVAL, "dtype must be QCB_C64 (0) or QCB_C128 (1)");
    Ctx &c = ctx();
    int gbits = 0;
    if (sharded) {
        if (c.nranks < 2 || !c.nccl_comm) return set_error(QCB_EINVAL, "qcb_dist_init must be called before creating sharded states");
        while ((1 << gbits) < c.nranks) gbits++;
    }
    if (nqubits < 1 || nqubits - gbits > 40 || nqubits - gbits < 1) return set_error(QCB_EINVAL, "nqubits = %d not supported", nqubits);
    if (sharded && nqubits - gbits < kMaxTB + gbits) return set_error(QCB_EINVAL, "sharded states need at least %d local qubits", kMaxTB + gbits);
    qcb_state *s = new qcb_state;
    s->n = nqubits;
    s->nloc = nqubits - gbits;
    s->dtype = dtype;
    s->sharded = sharded;
    s->bytes = amp_bytes(dtype) << s->nloc;
    s->phys.resize(nqubits);
    for (int q = 0; q < nqubits; q++) s->phys[q] = q;
    if (wrap) {
        s->d = wrap;
        s->owned = false;
    } else {
        cudaError_t e = cudaMalloc(&s->d, s->bytes);
        if (e != cudaSuccess) {
            delete s;
            return set_error(e == cudaErrorMemoryAllocation ? QCB_ENOMEM : QCB_ECUDA, "cudaMalloc(%zu bytes): %s", (size_t)(amp_bytes(dtype) << (nqubits - gbits)), cudaGetErrorString(e));
        }
        s->owned = true;
        if (sharded) dist_alloc_alt(s); // optional second buffer; failure just selects chunked in-place swaps
    }
    *out = s;
    return QCB_OK;
}

int qcb_state_create(int nqubits, int dtype, qcb_state **out) { return new_state(nqubits, dtype, nullptr, false, out); }
int qcb_state_wrap(int nqubits, int dtype, void *device_ptr, qcb_state **out)
{
    if (!device_ptr) return set_error(QCB_EINVAL, "null device pointer");
    if ((uintptr_t)device_ptr % 16 != 0) return set_error(QCB_EINVAL, "device pointer must be 16-byte aligned");
    return new_state(nqubits, dtype, device_ptr, false, out);
}
int qcb_state_create_sharded(int nqubits, int dtype, qcb_state **out) { return new_state(nqubits, dtype, nullptr, true, out); }

int qcb_state_destroy(qcb_state *s)
{
    if (!s) return QCB_OK;
    if (!s->peer_d.empty()) qcb_state_ipc_close(s);
    if (s->owned && s->d) {
        if (ctx().inited) cudaStreamSynchronize(ctx().stream);
        cudaFree(s->d);
        if (s->alt) cudaFree(s->alt);
    }
    delete s;
    return QCB_OK;
}
int qcb_state_nqubits(const qcb_state *s, int *out) { QCB_TRY(check_state(s)); *out = s->n; return QCB_OK; }
int qcb_state_local_qubits(const qcb_state *s, int *out) { QCB_TRY(check_state(s)); *out = s->nloc; return QCB_OK; }
int qcb_state_dtype(const qcb_state *s, int *out) { QCB_TRY(check_state(s)); *out = s->dtype; return QCB_OK; }
int qcb_state_device_ptr(const qcb_state *s, void **out) { QCB_TRY(check_state(s)); *out = s->d; return QCB_OK; }

int qcb_state_set_zero_state(qcb_state *s)
{
    QCB_REQUIRE_INIT();
    QCB_TRY(check_state(s));
    const unsigned long long namps = 1ull << s->nloc;
    const int set_one = (!s->sharded || ctx().rank == 0) ? 1 : 0;
    const unsigned blocks = grid_for(namps, 256, 16);
    if (s->dtype == QCB_C128) k_set_zero_state<double><<<blocks, 256, 0, ctx().stream>>>((double2 *)s->d, namps, set_one);
    else k_set_zero_state<float><<<blocks, 256, 0, ctx().stream>>>((float2 *)s->d, namps, set_one);
    QCB_CUDA(cudaGetLastError());
    count_launch();
    for (int q = 0; q < s->n; q++) s->phys[q] = q;
    return QCB_OK;
}

// psi' .= 0 (src/apply.jl:135) for callers that keep the reference's explicit zero fill
int qcb_state_fill_zero(qcb_state *s)
{
    QCB_REQUIRE_INIT();
    QCB_TRY(check_state(s));
    QCB_CUDA(cudaMemsetAsync(s->d, 0, s->bytes, ctx().stream));
    return QCB_OK;
}

int qcb_state_upload(qcb_state *s, const void *host)
{
    QCB_REQUIRE_INIT();
    QCB_TRY(check_state(s));
    if (!host) return set_error(QCB_EINVAL, "null host pointer");
    if (s->sharded) return set_error(QCB_EINVAL, "use qcb_state_upload_shard for sharded states");
    QCB_CUDA(cudaMemcpyAsync(s->d, host, s->bytes, cudaMemcpyHostToDevice, ctx().stream));
    QCB_CUDA(cudaStreamSynchronize(ctx().stream));
    return QCB_OK;
}
int qcb_state_download(const qcb_state *s, void *host)
{
    QCB_REQUIRE_INIT();
    QCB_TRY(check_state(s));
    if (!host) return set_error(QCB_EINVAL, "null host pointer");
    if (s->sharded) return set_error(QCB_EINVAL, "use qcb_state_download_shard for sharded states");
    QCB_CUDA(cudaMemcpyAsync(host, s->d, s->bytes, cudaMemcpyDeviceToHost, ctx().stream));
    QCB_CUDA(cudaStreamSynchronize(ctx().stream));
    return QCB_OK;
}
int qcb_state_upload_shard(qcb_state *s, const void *host)
{
    QCB_REQUIRE_INIT();
    QCB_TRY(check_state(s));
    if (!host) return set_error(QCB_EINVAL, "null host pointer");
    QCB_CUDA(cudaMemcpyAsync(s->d, host, s->bytes, cudaMemcpyHostToDevice, ctx().stream));
    QCB_CUDA(cudaStreamSynchronize(ctx().stream));
    for (int q = 0; q < s->n; q++) s->phys[q] = q;
    return QCB_OK;
}
int qcb_state_download_shard(qcb_state *s, void *host)
{
    QCB_REQUIRE_INIT();
    QCB_TRY(check_state(s));
    if (!host) return set_error(QCB_EINVAL, "null host pointer");
    if (s->sharded) QCB_TRY(dist_state_canonicalize(s));
    QCB_CUDA(cudaMemcpyAsync(host, s->d, s->bytes, cudaMemcpyDeviceToHost, ctx().stream));
    QCB_CUDA(cudaStreamSynchronize(ctx().stream));
    return QCB_OK;
}

int qcb_state_copy(qcb_state *dst, const qcb_state *src)
{
    QCB_REQUIRE_INIT();
    QCB_TRY(check_state(dst));
    QCB_TRY(check_state(src));
    if (dst->n != src->n || dst->dtype != src->dtype || dst->nloc != src->nloc)
        return set_error(QCB_EINVAL, "ψ′ must be the same size as ψ."); // src/apply.jl:108
    QCB_TRY(copy_state_raw(dst->d, src->d, src->bytes));
    dst->phys = src->phys;
    return QCB_OK;
}

int qcb_state_norm2(const qcb_state *s, double *out)
{
    QCB_REQUIRE_INIT();
    QCB_TRY(check_state(s));
    Ctx &c = ctx();
    const unsigned long long namps = 1ull << s->nloc;
    const unsigned blocks = grid_for(namps, kRedThreads, 8);
    QCB_TRY(ensure_partial(blocks));
    QCB_TRY(ensure_out(64));
    if (s->dtype == QCB_C128) k_norm2<double><<<blocks, kRedThreads, 0, c.stream>>>((const double2 *)s->d, namps, c.d_partial);
    else k_norm2<float><<<blocks, kRedThreads, 0, c.stream>>>((const float2 *)s->d, namps, c.d_partial);
    QCB_CUDA(cudaGetLastError());
    k_final_sum<<<1, 256, 0, c.stream>>>(c.d_partial, (int)blocks, 1, c.d_out);
    QCB_CUDA(cudaGetLastError());
    count_launch(2);
    QCB_CUDA(cudaMemcpyAsync(c.h_out, c.d_out, sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    QCB_CUDA(cudaStreamSynchronize(c.stream));
    *out = c.h_out[0]; // local shard only for sharded states
    return QCB_OK;
}

// ---- host gate algebra --------------------------------------------------------------------
int qcb_build_general_unitary_gate(const double *angles15, double *out_u32)
{
    if (!angles15 || !out_u32) return set_error(QCB_EINVAL, "null pointer");
    to_doubles(general_unitary(angles15), out_u32);
    return QCB_OK;
}
int qcb_general_unitary_gate_derivative(const double *angles15, int k, double *out_u32)
{
    if (!angles15 || !out_u32 || k < 0 || k > 14) return set_error(QCB_EINVAL, "bad arguments");
    to_doubles(general_unitary(angles15, k), out_u32);
    return QCB_OK;
}
int qcb_general_unitary_gate_gradient(const double *angles15, const double *env32, double *out15)
{
    if (!angles15 || !env32 || !out15) return set_error(QCB_EINVAL, "null pointer");
    general_unitary_gradient(angles15, env32, out15);
    return QCB_OK;
}
int qcb_brickwork_num_gates(int nbits, int nlayers) { return (int)brickwork_positions(nbits, nlayers).size(); }

// ---- gate application ---------------------------------------------------------------------
static int apply_gates_impl(qcb_state *s, const void *first_src, int ngates, const int *K, const int *arity, const double *mats_in)
{
    if (ngates < 0 || (ngates > 0 && (!K || !arity || !mats_in))) return set_error(QCB_EINVAL, "bad gate list");
    const int n = s->n;
    // a lone 1-qubit state has no pair to embed into: dedicated kernel
    if (n == 1) {
        if (first_src) QCB_CUDA(cudaMemcpyAsync(s->d, first_src, s->bytes, cudaMemcpyDeviceToDevice, ctx().stream));
        for (int g = 0; g < ngates; g++) {
            if (arity[g] != 1 || K[g] != 1) return set_error(QCB_EINVAL, "The gate cannot be applied as it sits outside the qubit space.");
            const double *u = mats_in + 32 * (size_t)g;
            if (s->dtype == QCB_C128) {
                Gate1<double> U; for (int i = 0; i < 8; i++) U.v[i] = u[i];
                k_apply_1q_simple<double><<<1, 32, 0, ctx().stream>>>((double2 *)s->d, 1ull, 0, U);
            } else {
                Gate1<float> U; for (int i = 0; i < 8; i++) U.v[i] = (float)u[i];
                k_apply_1q_simple<float><<<1, 32, 0, ctx().stream>>>((float2 *)s->d, 1ull, 0, U);
            }
            QCB_CUDA(cudaGetLastError());
            count_launch();
        }
        return QCB_OK;
    }
    std::vector<GateOp> gates(ngates);
    std::vector<double> mats((size_t)32 * ngates);
    for (int g = 0; g < ngates; g++) {
        const double *u = mats_in + 32 * (size_t)g;
        if (arity[g] == 2) {
            if (!(K[g] > 0 && K[g] < n)) // src/apply.jl:112
                return set_error(QCB_EINVAL, "The gate cannot be applied as it sits outside the qubit space.");
            std::memcpy(mats.data() + 32 * (size_t)g, u, 32 * sizeof(double));
            gates[g] = GateOp{K[g] - 1, g};
        } else if (arity[g] == 1) {
            if (!(K[g] >= 1 && K[g] <= n)) // src/apply.jl:27
                return set_error(QCB_EINVAL, "The gate cannot be applied as it sits outside the qubit space.");
            const bool on_low = K[g] < n; // pair (K, K+1) unless K is the last qubit
            embed_1q(u, on_low, mats.data() + 32 * (size_t)g);
            gates[g] = GateOp{on_low ? K[g] - 1 : K[g] - 2, g};
        } else
            return set_error(QCB_EINVAL, "gate %d: arity must be 1 or 2", g);
    }
    return run_gate_list(s, gates, mats.data(), false, first_src);
}
int qcb_apply_gates(qcb_state *s, int ngates, const int *K, const int *arity, const double *mats_in)
{
    QCB_REQUIRE_INIT();
    QCB_TRY(check_state(s));
    return apply_gates_impl(s, nullptr, ngates, K, arity, mats_in);
}
// apply!(psi', psi, gate) / apply!(cache, psi', psi, gate) with distinct states (src/apply.jl:20,39,117,121): dst = gates(src), src
// untouched; the first fused pass reads src and writes dst (no copy, no zero fill of psi' -- src/apply.jl:135)
int qcb_apply_gates_from(qcb_state *dst, const qcb_state *src, int ngates, const int *K, const int *arity, const double *mats_in)
{
    QCB_REQUIRE_INIT();
    QCB_TRY(check_state(dst));
    QCB_TRY(check_state(src));
    if (dst == src || dst->d == src->d) return apply_gates_impl(dst, nullptr, ngates, K, arity, mats_in);
    if (dst->n != src->n || dst->dtype != src->dtype || dst->sharded || src->sharded) return set_error(QCB_EINVAL, "psi' must be the same size as psi.");
    return apply_gates_impl(dst, src->d, ngates, K, arity, mats_in);
}

int qcb_apply_1q(qcb_state *s, const double *u8, int K)
{
    if (!u8) return set_error(QCB_EINVAL, "null gate");
    double m[32] = {0};
    std::memcpy(m, u8, 8 * sizeof(double));
    const int ar = 1;
    return qcb_apply_gates(s, 1, &K, &ar, m);
}
int qcb_apply_2q(qcb_state *s, const double *u32, int K)
{
    if (!u32) return set_error(QCB_EINVAL, "null gate");
    const int ar = 2;
    return qcb_apply_gates(s, 1, &K, &ar, u32);
}

int qcb_apply_brickwork_range(qcb_state *s, int nbits, int nlayers, const double *angles, int first_gate, int last_gate, int adjoint)
{
    QCB_REQUIRE_INIT();
    QCB_TRY(check_state(s));
    if (s->n != nbits) return set_error(QCB_EINVAL, "circuit has %d qubits, state has %d", nbits, s->n);
    if (!angles && last_gate > first_gate) return set_error(QCB_EINVAL, "null angles");
    std::vector<GateOp> gates;
    std::vector<double> mats;
    QCB_TRY(brickwork_gates(nbits, nlayers, angles, first_gate, last_gate, adjoint != 0, gates, mats));
    return run_gate_list(s, gates, mats.data(), adjoint != 0);
}
// apply!(cache, psi', psi, circuit) with distinct states, as the reference's out-of-place contract has it (src/circuits.jl:19-31):
// dst = circuit(src), src untouched.  The first fused pass reads src and writes dst, so no copy of the state precedes it.
int qcb_apply_brickwork_from(qcb_state *dst, const qcb_state *src, int nbits, int nlayers, const double *angles, int first_gate, int last_gate, int adjoint)
{
    QCB_REQUIRE_INIT();
    QCB_TRY(check_state(dst));
    QCB_TRY(check_state(src));
    if (dst == src || dst->d == src->d) return qcb_apply_brickwork_range(dst, nbits, nlayers, angles, first_gate, last_gate, adjoint);
    if (dst->n != src->n || dst->dtype != src->dtype || dst->sharded || src->sharded) return set_error(QCB_EINVAL, "psi' must be the same size as psi.");
    if (dst->n != nbits) return set_error(QCB_EINVAL, "circuit has %d qubits, state has %d", nbits, dst->n);
    if (!angles && last_gate > first_gate) return set_error(QCB_EINVAL, "null angles");
    std::vector<GateOp> gates;
    std::vector<double> mats;
    QCB_TRY(brickwork_gates(nbits, nlayers, angles, first_gate, last_gate, adjoint != 0, gates, mats));
    return run_gate_list(dst, gates, mats.data(), adjoint != 0, src->d);
}
int qcb_apply_brickwork(qcb_state *s, int nbits, int nlayers, const double *angles)
{
    return qcb_apply_brickwork_range(s, nbits, nlayers, angles, 0, qcb_brickwork_num_gates(nbits, nlayers), 0);
}

// Host-buffer form.  The transfers are pipelined with the part of the circuit that is local to a contiguous
// chunk of the state: the state is cut into 2^b chunks (the top b index bits); the gates that do not depend on
// the top b qubits ("head", a light-cone triangle at the start of the circuit) run on chunk k while chunk k+1 is
// still on the PCIe bus, and symmetrically the gates nothing else depends on ("tail") run chunk by chunk while
// finished chunks already travel back.  Only the middle part needs the whole state resident.
int qcb_apply_brickwork_host(int dtype, int nbits, int nlayers, const double *angles, const void *psi_in_host, void *psi_out_host)
{
    QCB_REQUIRE_INIT();
    if (!psi_out_host) return set_error(QCB_EINVAL, "null output buffer");
    if (dtype != QCB_C64 && dtype != QCB_C128) return set_error(QCB_EINVAL, "bad dtype");
    if (nbits < 1 || nbits > 40) return set_error(QCB_EINVAL, "bad qubit count");
    Ctx &c = ctx();
    const size_t bytes = amp_bytes(dtype) << nbits;
    QCB_TRY(ensure_scratch(2, bytes));
    qcb_state s;
    s.n = s.nloc = nbits; s.dtype = dtype; s.d = c.scratch[2]; s.bytes = bytes;
    s.phys.resize(nbits);
    for (int q = 0; q < nbits; q++) s.phys[q] = q;
    const int G = qcb_brickwork_num_gates(nbits, nlayers);
    std::vector<GateOp> gates;
    std::vector<double> mats;
    if (G > 0) QCB_TRY(brickwork_gates(nbits, nlayers, angles, 0, G, false, gates, mats));
    const int b = nbits >= 27 ? 3 : (nbits >= 25 ? 2 : 0); // chunk = 2^(n-b) amplitudes
    double tot_passes = 0, tot_stages = 0, tot_launches = 0;
    auto run = [&](qcb_state *st, const std::vector<GateOp> &g) -> int {
        if (g.empty()) return QCB_OK;
        const int rc = run_gate_list(st, g, mats.data(), false);
        tot_passes += c.st_passes; tot_stages += c.st_stages; tot_launches += c.st_launches;
        return rc;
    };
    if (b == 0 || c.force_simple) {
        if (psi_in_host) QCB_CUDA(cudaMemcpyAsync(s.d, psi_in_host, bytes, cudaMemcpyHostToDevice, c.stream));
        else QCB_TRY(qcb_state_set_zero_state(&s));
        QCB_TRY(run(&s, gates));
        QCB_CUDA(cudaMemcpyAsync(psi_out_host, s.d, bytes, cudaMemcpyDeviceToHost, c.stream));
        QCB_CUDA(cudaStreamSynchronize(c.stream));
    } else {
        // split: head (dependency-closed prefix below the top b qubits), tail (closed suffix), middle
        const int nc = nbits - b;
        std::vector<char> tag(G, 0); // 1 head, 2 tail
        {
            std::vector<char> dead(nbits, 0);
            for (int q = nc; q < nbits; q++) dead[q] = 1;
            if (psi_in_host)
                for (int g = 0; g < G; g++) {
                    const int q = gates[g].q;
                    if (dead[q] || dead[q + 1]) dead[q] = dead[q + 1] = 1;
                    else tag[g] = 1;
                }
            std::fill(dead.begin(), dead.end(), 0);
            for (int q = nc; q < nbits; q++) dead[q] = 1;
            for (int g = G - 1; g >= 0; g--) {
                const int q = gates[g].q;
                if (tag[g] == 1 || dead[q] || dead[q + 1]) dead[q] = dead[q + 1] = 1;
                else tag[g] = 2;
            }
        }
        std::vector<GateOp> head, mid, tail;
        for (int g = 0; g < G; g++) (tag[g] == 1 ? head : tag[g] == 2 ? tail : mid).push_back(gates[g]);
        const int nchunks = 1 << b;
        const size_t cbytes = bytes >> b;
        qcb_state view = s;
        view.n = view.nloc = nc; view.bytes = cbytes;
        view.phys.resize(nc);
        if (psi_in_host) {
            for (int k = 0; k < nchunks; k++) {
                char *dk = (char *)s.d + (size_t)k * cbytes;
                QCB_CUDA(cudaMemcpyAsync(dk, (const char *)psi_in_host + (size_t)k * cbytes, cbytes, cudaMemcpyHostToDevice, c.copy_stream));
                QCB_CUDA(cudaEventRecord(c.chunk_ev[0][k], c.copy_stream));
                QCB_CUDA(cudaStreamWaitEvent(c.stream, c.chunk_ev[0][k], 0));
                view.d = dk;
                QCB_TRY(run(&view, head));
            }
        } else {
            QCB_TRY(qcb_state_set_zero_state(&s));
        }
        QCB_TRY(run(&s, mid));
        for (int k = 0; k < nchunks; k++) {
            char *dk = (char *)s.d + (size_t)k * cbytes;
            view.d = dk;
            QCB_TRY(run(&view, tail));
            QCB_CUDA(cudaEventRecord(c.chunk_ev[1][k], c.stream));
            QCB_CUDA(cudaStreamWaitEvent(c.copy_stream, c.chunk_ev[1][k], 0));
            QCB_CUDA(cudaMemcpyAsync((char *)psi_out_host + (size_t)k * cbytes, dk, cbytes, cudaMemcpyDeviceToHost, c.copy_stream));
        }
        QCB_CUDA(cudaStreamSynchronize(c.copy_stream));
        QCB_CUDA(cudaStreamSynchronize(c.stream));
    }
    c.st_passes = tot_passes; c.st_stages = tot_stages; c.st_launches = tot_launches; c.st_gates = G;
    return QCB_OK;
}

// ---- expectation values ---------------------------------------------------------------------
int qcb_measure_tfim(const qcb_state *s, double J, double h, double g, double *out_E2)
{
    if (!out_E2) return set_error(QCB_EINVAL, "null output");
    return measure_impl(s, QCB_HAM_TFIM, J, h, g, out_E2);
}
int qcb_measure_east(const qcb_state *s, double c, double A, double B, double *out_E2)
{
    if (!out_E2) return set_error(QCB_EINVAL, "null output");
    return measure_impl(s, QCB_HAM_EAST, c, A, B, out_E2);
}

// measure(H::AbstractMatrix, psi)  src/circuits.jl:49-55 -- explicit Hamiltonian, CSR on the host (a dense matrix is the CSR
// with every entry).  The matrix is uploaded for the call (this is the reference's small-n cross-check path).
int qcb_measure_csr(const qcb_state *s, long long nrows, const long long *rowptr, const long long *colidx, const double *vals_reim, double *out_E2)
{
    QCB_REQUIRE_INIT();
    QCB_TRY(check_state(s));
    if (!out_E2) return set_error(QCB_EINVAL, "null pointer");
    if (s->sharded) return set_error(QCB_EUNSUPPORTED, "matrix Hamiltonians are not supported on sharded states");
    DeviceCsr D;
    QCB_TRY(upload_csr(D, s->n, nrows, rowptr, colidx, vals_reim));
    return measure_csr_device(s, D, out_E2);
}

// gradients(H::AbstractMatrix, psi0, circuit; calculate_energy): the reference reaches this through the generic
// gradients! + measure!(psi', H::AbstractMatrix, psi) (src/gradients.jl:88-152, src/circuits.jl:56-58;
// examples/test_gradients.jl:25).  Same adjoint sweep as for the kernel Hamiltonians with lambda = H psi as a CSR product;
// H must be Hermitian (the reference asserts a real expectation value).
int qcb_gradients_brickwork_csr(const qcb_state *psi0, int nbits, int nlayers, const double *angles, long long nrows, const long long *rowptr,
                                const long long *colidx, const double *vals_reim, int want_energy, double *out_E, double *out_grads)
{
    QCB_REQUIRE_INIT();
    QCB_TRY(check_state(psi0));
    if (!angles || !out_grads || (want_energy && !out_E)) return set_error(QCB_EINVAL, "null pointer");
    DeviceCsr D;
    QCB_TRY(upload_csr(D, psi0->n, nrows, rowptr, colidx, vals_reim));
    return gradients_impl(psi0, nbits, nlayers, angles, 0, nullptr, want_energy, out_E, out_grads, &D);
}

// ---- polar-optimiser building blocks -----------------------------------------------------------------------------
int qcb_state_inner(const qcb_state *a, const qcb_state *b, double *out2)
{
    QCB_REQUIRE_INIT();
    QCB_TRY(check_state(a));
    QCB_TRY(check_state(b));
    if (!out2) return set_error(QCB_EINVAL, "null output");
    if (a->n != b->n || a->dtype != b->dtype || a->sharded || b->sharded) return set_error(QCB_EINVAL, "states must have the same size and type");
    Ctx &c = ctx();
    const unsigned long long namps = 1ull << a->nloc;
    const unsigned blocks = grid_for(namps, kRedThreads, 8);
    QCB_TRY(ensure_partial((size_t)blocks * 2));
    QCB_TRY(ensure_out(64));
    if (a->dtype == QCB_C128) k_inner<double><<<blocks, kRedThreads, 0, c.stream>>>((const double2 *)a->d, (const double2 *)b->d, namps, (double2 *)c.d_partial);
    else k_inner<float><<<blocks, kRedThreads, 0, c.stream>>>((const float2 *)a->d, (const float2 *)b->d, namps, (double2 *)c.d_partial);
    QCB_CUDA(cudaGetLastError());
    k_final_sum<<<2, 256, 0, c.stream>>>(c.d_partial, (int)blocks, 2, c.d_out);
    QCB_CUDA(cudaGetLastError());
    count_launch(2);
    QCB_CUDA(cudaMemcpyAsync(c.h_out, c.d_out, 2 * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    QCB_CUDA(cudaStreamSynchronize(c.stream));
    out2[0] = c.h_out[0];
    out2[1] = c.h_out[1];
    return QCB_OK;
}

int qcb_environment_2q(const qcb_state *lam, const qcb_state *psi, int K, double *out_env32)
{
    QCB_REQUIRE_INIT();
    QCB_TRY(check_state(lam));
    QCB_TRY(check_state(psi));
    if (!out_env32) return set_error(QCB_EINVAL, "null output");
    if (lam->n != psi->n || lam->dtype != psi->dtype || lam->sharded || psi->sharded) return set_error(QCB_EINVAL, "states must have the same size and type");
    if (!(K > 0 && K < psi->n)) return set_error(QCB_EINVAL, "The gate cannot be applied as it sits outside the qubit space.");
    Ctx &c = ctx();
    const unsigned long long nquads = 1ull << (psi->nloc - 2);
    const unsigned blocks = grid_for(nquads, kRedThreads, 4);
    QCB_TRY(ensure_partial((size_t)blocks * 32));
    QCB_TRY(ensure_out(64));
    if (psi->dtype == QCB_C128) k_env_2q<double><<<blocks, kRedThreads, 0, c.stream>>>((const double2 *)psi->d, (const double2 *)lam->d, nquads, K - 1, c.d_partial);
    else k_env_2q<float><<<blocks, kRedThreads, 0, c.stream>>>((const float2 *)psi->d, (const float2 *)lam->d, nquads, K - 1, c.d_partial);
    QCB_CUDA(cudaGetLastError());
    k_final_sum<<<32, 128, 0, c.stream>>>(c.d_partial, (int)blocks, 32, c.d_out);
    QCB_CUDA(cudaGetLastError());
    count_launch(2);
    QCB_CUDA(cudaMemcpyAsync(c.h_out, c.d_out, 32 * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    QCB_CUDA(cudaStreamSynchronize(c.stream));
    std::memcpy(out_env32, c.h_out, 32 * sizeof(double));
    return QCB_OK;
}

// ---- gradients / optimisation -----------------------------------------------------------------
int qcb_gradients_brickwork(const qcb_state *psi0, int nbits, int nlayers, const double *angles, int ham_kind,
                            const double *ham_params3, int want_energy, double *out_E, double *out_grads)
{
    if (!angles || !ham_params3 || !out_grads || (want_energy && !out_E)) return set_error(QCB_EINVAL, "null pointer");
    return gradients_impl(psi0, nbits, nlayers, angles, ham_kind, ham_params3, want_energy, out_E, out_grads);
}

int qcb_optimise_brickwork(const qcb_state *psi0, int nbits, int nlayers, double *angles, int ham_kind, const double *hp,
                           int epochs, double lr, double *energies)
{
    QCB_REQUIRE_INIT();
    if (!angles || !hp || !energies || epochs < 0) return set_error(QCB_EINVAL, "bad arguments");
    const int G = qcb_brickwork_num_gates(nbits, nlayers);
    std::vector<double> grads((size_t)15 * std::max(G, 1));
    energies[0] = 0.0; // never written by the reference (src/gradients.jl:5,11)
    for (int e = 1; e <= epochs; e++) {
        double E = 0;
        QCB_TRY(gradients_impl(psi0, nbits, nlayers, angles, ham_kind, hp, 1, &E, grads.data()));
        energies[e] = E;                                                          // :11
        for (size_t i = 0; i < (size_t)15 * G; i++) angles[i] -= lr * grads[i];  // :13
    }
    // energies[end] = measure(H, psi0, circuit)   :15
    Ctx &c = ctx();
    QCB_TRY(check_state(psi0));
    QCB_TRY(ensure_scratch(0, psi0->bytes));
    qcb_state A = *psi0;
    // (a work copy must not inherit psi0's second buffer or its peer views: a fused qubit swap would read the peers' psi0)
    A.d = c.scratch[0]; A.owned = false; A.alt = nullptr; A.peer_d.clear(); A.peer_alt.clear();
    QCB_TRY(copy_state_raw(A.d, psi0->d, psi0->bytes));
    QCB_TRY(qcb_apply_brickwork(&A, nbits, nlayers, angles));
    double E2[2];
    QCB_TRY(measure_impl(&A, ham_kind, hp[0], hp[1], hp[2], E2));
    energies[epochs] = E2[0];
    return QCB_OK;
}

} // extern "C"
