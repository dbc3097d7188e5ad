// gates_hd.cuh -- the gate algebra of gates_host.hpp in a form that also runs on the device.
//
// The adjoint gradient ends with a small contraction per gate: grads[k, g] = 2 Re sum env_g .* dU_g/dtheta_k for the 15
// angles of build_general_unitary_gate (src/gates.jl:80-109).  On the host that is 4 us per gate -- most of a 12- or
// 20-qubit gradient (110 gates: 0.44 of 0.64 ms).  k_contract_env does it for all gates at once on the GPU, next to the
// environments it consumes, so that only 15 G doubles come back.  Plain structs instead of std::complex; the same
// functions compile for the host, where tests/test_abi.py checks them against gates_host.hpp.
#pragma once
#include "tile.cuh"

namespace qcb {
namespace hd {

struct Cx { double re, im; };
QCB_HD Cx cx(double r, double i = 0.0) { Cx z; z.re = r; z.im = i; return z; }
QCB_HD Cx operator*(Cx a, Cx b) { return cx(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re); }
QCB_HD Cx operator*(Cx a, double s) { return cx(a.re * s, a.im * s); }
QCB_HD Cx operator+(Cx a, Cx b) { return cx(a.re + b.re, a.im + b.im); }
QCB_HD Cx operator-(Cx a) { return cx(-a.re, -a.im); }
QCB_HD Cx times_i(Cx a) { return cx(-a.im, a.re); }

struct M2 { Cx m[4]; };  // column-major m[r + 2c]
struct M4 { Cx m[16]; }; // column-major m[r + 4c], index = (bit of qubit K) + 2 (bit of qubit K+1)

QCB_HD M2 mul(const M2 &a, const M2 &b)
{
    M2 o;
    for (int c = 0; c < 2; c++)
        for (int r = 0; r < 2; r++) o.m[r + 2 * c] = a.m[r] * b.m[2 * c] + a.m[r + 2] * b.m[1 + 2 * c];
    return o;
}
QCB_HD M4 mul(const M4 &a, const M4 &b)
{
    M4 o;
    for (int c = 0; c < 4; c++)
        for (int r = 0; r < 4; r++) {
            double sr = 0, si = 0;
            for (int k = 0; k < 4; k++) {
                const Cx x = a.m[r + 4 * k], y = b.m[k + 4 * c];
                sr += x.re * y.re - x.im * y.im;
                si += x.re * y.im + x.im * y.re;
            }
            o.m[r + 4 * c] = cx(sr, si);
        }
    return o;
}
// kron(hi, lo): `lo` acts on qubit K (the low index bit), `hi` on K+1
QCB_HD M4 kron(const M2 &hi, const M2 &lo)
{
    M4 o;
    for (int ch = 0; ch < 2; ch++)
        for (int cl = 0; cl < 2; cl++)
            for (int rh = 0; rh < 2; rh++)
                for (int rl = 0; rl < 2; rl++) o.m[(2 * rh + rl) + 4 * (2 * ch + cl)] = hi.m[rh + 2 * ch] * lo.m[rl + 2 * cl];
    return o;
}
QCB_HD M4 transpose(const M4 &a)
{
    M4 o;
    for (int c = 0; c < 4; c++)
        for (int r = 0; r < 4; r++) o.m[r + 4 * c] = a.m[c + 4 * r];
    return o;
}
// the CNOTs are permutation matrices: rcnot exchanges index 2 <-> 3, cnot 1 <-> 3 (rows from the left, columns from the right)
QCB_HD M4 swap_rows(M4 m, int i, int j)
{
    for (int c = 0; c < 4; c++) { const Cx t = m.m[i + 4 * c]; m.m[i + 4 * c] = m.m[j + 4 * c]; m.m[j + 4 * c] = t; }
    return m;
}
QCB_HD M4 swap_cols(M4 m, int i, int j)
{
    for (int r = 0; r < 4; r++) { const Cx t = m.m[r + 4 * i]; m.m[r + 4 * i] = m.m[r + 4 * j]; m.m[r + 4 * j] = t; }
    return m;
}

QCB_HD M2 identity2() { M2 o; o.m[0] = cx(1); o.m[1] = cx(0); o.m[2] = cx(0); o.m[3] = cx(1); return o; }
// RZ(a) = diag(e^{-ia/2}, e^{+ia/2})  src/gates.jl:36-43 ;  dRZ/da = RZ(a + pi) / 2
QCB_HD void rz_all(double a, M2 &r, M2 &d)
{
    double s, c;
    sincos(a / 2, &s, &c);
    r.m[0] = cx(c, -s); r.m[1] = cx(0); r.m[2] = cx(0); r.m[3] = cx(c, s);
    d.m[0] = cx(-0.5 * s, -0.5 * c); d.m[1] = cx(0); d.m[2] = cx(0); d.m[3] = cx(-0.5 * s, 0.5 * c);
}
// RY(a) = [[c, -s], [s, c]]  src/gates.jl:44-51 ;  dRY/da = RY(a + pi) / 2
QCB_HD void ry_all(double a, M2 &r, M2 &d)
{
    double s, c;
    sincos(a / 2, &s, &c);
    r.m[0] = cx(c); r.m[1] = cx(s); r.m[2] = cx(-s); r.m[3] = cx(c);
    d.m[0] = cx(-0.5 * s); d.m[1] = cx(0.5 * c); d.m[2] = cx(-0.5 * c); d.m[3] = cx(-0.5 * s);
}
// RotationGate(theta, phi, lambda) = [[c, -e^{i lam} s], [e^{i phi} s, e^{i(phi+lam)} c]]  src/gates.jl:61-73, and its
// three partial derivatives
QCB_HD void rotation_all(double theta, double phi, double lam, M2 &r, M2 (&d)[3])
{
    double s, c, sl, cl, sp, cp;
    sincos(theta / 2, &s, &c);
    sincos(lam, &sl, &cl);
    sincos(phi, &sp, &cp);
    const Cx p1 = cx(cl, sl), p2 = cx(cp, sp), p3 = p1 * p2;
    r.m[0] = cx(c); r.m[1] = p2 * s; r.m[2] = -(p1 * s); r.m[3] = p3 * c;
    d[0].m[0] = cx(-0.5 * s); d[0].m[1] = p2 * (0.5 * c); d[0].m[2] = p1 * (-0.5 * c); d[0].m[3] = p3 * (-0.5 * s);
    d[1].m[0] = cx(0); d[1].m[1] = times_i(p2) * s; d[1].m[2] = cx(0); d[1].m[3] = times_i(p3) * c;
    d[2].m[0] = cx(0); d[2].m[1] = cx(0); d[2].m[2] = -(times_i(p1) * s); d[2].m[3] = times_i(p3) * c;
}

// everything the two functions below share: the four layers l1..l4 and the factors their derivatives need
struct GeneralGate {
    M2 r[4], dr[4][3], rzm, z13, dz13, y14, dy14, y15, dy15, hi1, lo4;
    M4 l1, l2, l3, l4;
};
QCB_HD void general_gate_setup(const double *a, GeneralGate &G)
{
    for (int i = 0; i < 4; i++) rotation_all(a[3 * i], a[3 * i + 1], a[3 * i + 2], G.r[i], G.dr[i]);
    M2 unused;
    rz_all(-3.14159265358979323846 / 2, G.rzm, unused);
    rz_all(a[12], G.z13, G.dz13);
    ry_all(a[13], G.y14, G.dy14);
    ry_all(a[14], G.y15, G.dy15);
    G.hi1 = mul(G.rzm, G.r[3]);
    G.lo4 = mul(G.r[0], G.rzm);
    G.l1 = kron(G.hi1, G.r[2]);        // kron(RZ(-pi/2) r4, r3)       :97
    G.l2 = kron(G.y14, G.z13);         // kron(RY(phi), RZ(theta))     :98
    G.l3 = kron(G.y15, identity2());   // kron(RY(lambda), 1)          :99
    G.l4 = kron(G.r[1], G.lo4);        // kron(r2, r1 RZ(-pi/2))       :100
}
// U = (l4 rc) (l3 cn) ((l2 rc) l1)   src/gates.jl:105-107
QCB_HD M4 general_unitary_of(const GeneralGate &G)
{
    return mul(mul(swap_cols(G.l4, 2, 3), swap_cols(G.l3, 1, 3)), mul(swap_cols(G.l2, 2, 3), G.l1));
}
// out15[k] = 2 Re sum_{r,c} env[r,c] dU/dtheta_k [r,c]: dU = X_j dl_j Y_j for the layer l_j holding angle k, so
// sum env .* dU = sum dl_j .* W_j with W_j = X_j^T env Y_j^T (gates_host.hpp: general_unitary_gradient).  `layer` = 1..4
// computes the angles of that layer only (three 4x4 products each: the GPU gives every layer of every gate its own
// thread); layer = 0: all of them.
QCB_HD double dot_re(const M4 &d, const M4 &w)
{
    double s = 0;
    for (int i = 0; i < 16; i++) s += d.m[i].re * w.m[i].re - d.m[i].im * w.m[i].im;
    return 2 * s;
}
QCB_HD void general_gate_gradient(const GeneralGate &G, const M4 &E, double *out15, int layer = 0)
{
    if (layer == 0 || layer == 4) { // angles 1-6: r1 (qubit K), r2 (qubit K+1) in l4; X_4 = 1
        const M4 Y2 = swap_rows(G.l1, 2, 3), Y3 = swap_rows(mul(G.l2, Y2), 1, 3), Y4 = swap_rows(mul(G.l3, Y3), 2, 3);
        const M4 W4 = mul(E, transpose(Y4));
        for (int w = 0; w < 3; w++) {
            out15[w] = dot_re(kron(G.r[1], mul(G.dr[0][w], G.rzm)), W4);
            out15[3 + w] = dot_re(kron(G.dr[1][w], G.lo4), W4);
        }
    }
    if (layer == 0 || layer == 1) { // angles 7-12: r3 (qubit K), r4 (qubit K+1) in l1; Y_1 = 1
        const M4 X3 = swap_cols(G.l4, 2, 3), X2 = swap_cols(mul(X3, G.l3), 1, 3), X1 = swap_cols(mul(X2, G.l2), 2, 3);
        const M4 W1 = mul(transpose(X1), E);
        for (int w = 0; w < 3; w++) {
            out15[6 + w] = dot_re(kron(G.hi1, G.dr[2][w]), W1);
            out15[9 + w] = dot_re(kron(mul(G.rzm, G.dr[3][w]), G.r[2]), W1);
        }
    }
    if (layer == 0 || layer == 2) { // theta: RZ (qubit K), phi: RY (qubit K+1) in l2
        const M4 X2 = swap_cols(mul(swap_cols(G.l4, 2, 3), G.l3), 1, 3), Y2 = swap_rows(G.l1, 2, 3);
        const M4 W2 = mul(mul(transpose(X2), E), transpose(Y2));
        out15[12] = dot_re(kron(G.y14, G.dz13), W2);
        out15[13] = dot_re(kron(G.dy14, G.z13), W2);
    }
    if (layer == 0 || layer == 3) { // lambda: RY (qubit K+1) in l3
        const M4 X3 = swap_cols(G.l4, 2, 3), Y3 = swap_rows(mul(G.l2, swap_rows(G.l1, 2, 3)), 1, 3);
        const M4 W3 = mul(mul(transpose(X3), E), transpose(Y3));
        out15[14] = dot_re(kron(G.dy15, identity2()), W3);
    }
}

// One gate of the gradient: env32 = the 4x4 matrix the backward pass reduced for it, laid out [2 (r + 4c)] (Re, Im);
// post != 0: it is M = sum conj(lambda_{g-1}) psi_{g-1}^T (tensor-core pass) and env = conj(U) M comes first.
QCB_HD void contract_env_gate(const double *angles15, const double *env32, int post, double *out15, int layer = 0)
{
    GeneralGate G;
    general_gate_setup(angles15, G);
    M4 E;
    for (int i = 0; i < 16; i++) E.m[i] = cx(env32[2 * i], env32[2 * i + 1]);
    if (post) {
        M4 Uc = general_unitary_of(G);
        for (int i = 0; i < 16; i++) Uc.m[i].im = -Uc.m[i].im;
        E = mul(Uc, E);
    }
    general_gate_gradient(G, E, out15, layer);
}

} // namespace hd

#if defined(__CUDACC__)
// grads[15 g + k] for all gates; one thread per (gate, layer of the gate) -- the plain form (cross-check of the cooperative
// kernel below: option "contract_coop" = 0)
static __global__ void k_contract_env(const double *__restrict__ angles, const double *__restrict__ env, int ngates, int post, double *grads)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x, g = i >> 2;
    if (g < ngates) hd::contract_env_gate(angles + 15 * (size_t)g, env + 32 * (size_t)g, post, grads + 15 * (size_t)g, 1 + (i & 3));
}

// ---------------------------------------------------------------------------------------------------------------------
// The same contraction with 16 lanes per gate: every 4x4 matrix lives as ONE element per lane (lane e = r + 4c), products
// fetch their operands with shuffles inside the half warp, the 2x2 factors are recomputed by every lane.  No matrix ever
// sits in local memory: ~7 us instead of ~60 us per launch, which is 10 % of a 12-qubit gradient.  The lane algebra
// (which lane a shuffle reads for a product, a transpose, a CNOT permutation) was prototyped lane by lane against
// gates_host.hpp before it was written down here; tests/test_gpu_parity.py compares it with the plain form above.
namespace coop {
using hd::Cx;
using hd::cx;
using hd::M2;

__device__ __forceinline__ Cx m2_elem(const M2 &a, uint32_t i, uint32_t j)
{
    const Cx c0 = i ? a.m[1] : a.m[0], c1 = i ? a.m[3] : a.m[2];
    return j ? c1 : c0;
}
// element (r, c) of kron(hi, lo)
__device__ __forceinline__ Cx kron_elem(const M2 &hi, const M2 &lo, uint32_t r, uint32_t c) { return m2_elem(hi, r >> 1, c >> 1) * m2_elem(lo, r & 1u, c & 1u); }
__device__ __forceinline__ uint32_t s23(uint32_t x) { return x ^ ((x >> 1) & 1u); } // rcnot: 2 <-> 3
__device__ __forceinline__ uint32_t s13(uint32_t x) { return x ^ ((x & 1u) << 1); } // cnot:  1 <-> 3
__device__ __forceinline__ Cx shfl_cx(Cx v, uint32_t src) { return cx(__shfl_sync(0xffffffffu, v.re, (int)src, 16), __shfl_sync(0xffffffffu, v.im, (int)src, 16)); }
// lane (r, c): sum_k A'(r, k) B'(k, c) with A' = A^T if TA (A(k, r) sits on lane k + 4r), B' = B^T if TB
template <bool TA, bool TB> __device__ __forceinline__ Cx mm(Cx A, Cx B, uint32_t r, uint32_t c)
{
    double sr = 0, si = 0;
#pragma unroll
    for (uint32_t k = 0; k < 4; k++) {
        const Cx a = shfl_cx(A, TA ? (k + 4 * r) : (r + 4 * k)), b = shfl_cx(B, TB ? (c + 4 * k) : (k + 4 * c));
        sr += a.re * b.re - a.im * b.im;
        si += a.re * b.im + a.im * b.re;
    }
    return cx(sr, si);
}
// 2 Re sum over the 16 lanes of d .* w (valid on every lane)
__device__ __forceinline__ double dot_re(Cx d, Cx w)
{
    double v = d.re * w.re - d.im * w.im;
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o, 16);
    return 2 * v;
}
} // namespace coop

// blockDim.x = 128: 8 gates per block, one per half warp
static __global__ void __launch_bounds__(128) k_contract_env_coop(const double *__restrict__ angles, const double *__restrict__ env, int ngates, int post,
                                                                  double *grads)
{
    using namespace coop;
    const uint32_t e = threadIdx.x & 15u, r = e & 3u, c = e >> 2;
    const int gid = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 4);
    const int g = gid < ngates ? gid : ngates - 1; // whole warps stay active for the shuffles; surplus half warps do not write
    const double *a = angles + 15 * (size_t)g;
    M2 rot[4], unused3[3], rzm, z13, dz13, y14, dy14, y15, dy15, unused;
#pragma unroll
    for (int i = 0; i < 4; i++) hd::rotation_all(a[3 * i], a[3 * i + 1], a[3 * i + 2], rot[i], unused3);
    hd::rz_all(-3.14159265358979323846 / 2, rzm, unused);
    hd::rz_all(a[12], z13, dz13);
    hd::ry_all(a[13], y14, dy14);
    hd::ry_all(a[14], y15, dy15);
    const M2 hi1 = hd::mul(rzm, rot[3]), lo4 = hd::mul(rot[0], rzm), id = hd::identity2();
    // this lane's element of the four layers and of their CNOT-permuted forms
    const Cx l1 = kron_elem(hi1, rot[2], r, c), l2 = kron_elem(y14, z13, r, c), l3 = kron_elem(y15, id, r, c);
    const Cx Y2 = kron_elem(hi1, rot[2], s23(r), c);  // rcnot l1
    const Cx X3 = kron_elem(rot[1], lo4, r, s23(c));  // l4 rcnot
    const Cx D12 = kron_elem(y14, dz13, r, c), D13 = kron_elem(dy14, z13, r, c), D14 = kron_elem(dy15, id, r, c);
    Cx E = cx(env[32 * (size_t)g + 2 * e], env[32 * (size_t)g + 2 * e + 1]);
    if (post) { // env = conj(U) M,  U = (l4 rc) (l3 cn) ((l2 rc) l1)
        const Cx L3c = kron_elem(y15, id, r, s13(c)), L2c = kron_elem(y14, z13, r, s23(c));
        Cx U = mm<false, false>(mm<false, false>(X3, L3c, r, c), mm<false, false>(L2c, l1, r, c), r, c);
        U.im = -U.im;
        E = mm<false, false>(U, E, r, c);
    }
    // suffixes Y_j and prefixes X_j of the layers; a CNOT on a product is a shuffle
    const Cx Y3 = shfl_cx(mm<false, false>(l2, Y2, r, c), s13(r) + 4 * c);
    const Cx Y4 = shfl_cx(mm<false, false>(l3, Y3, r, c), s23(r) + 4 * c);
    const Cx X2 = shfl_cx(mm<false, false>(X3, l3, r, c), r + 4 * s13(c));
    const Cx X1 = shfl_cx(mm<false, false>(X2, l2, r, c), r + 4 * s23(c));
    const Cx W4 = mm<false, true>(E, Y4, r, c);                            // X_4 = 1
    const Cx W1 = mm<true, false>(X1, E, r, c);                            // Y_1 = 1
    const Cx W2 = mm<false, true>(mm<true, false>(X2, E, r, c), Y2, r, c);
    const Cx W3 = mm<false, true>(mm<true, false>(X3, E, r, c), Y3, r, c);
    double out[15];
    out[12] = dot_re(D12, W2);
    out[13] = dot_re(D13, W2);
    out[14] = dot_re(D14, W3);
#pragma unroll
    for (int i = 0; i < 4; i++) { // the rotations' derivatives (trigonometry recomputed: cheaper than keeping 12 more 2x2 matrices alive)
        M2 ri, d[3];
        hd::rotation_all(a[3 * i], a[3 * i + 1], a[3 * i + 2], ri, d);
#pragma unroll
        for (int w = 0; w < 3; w++) {
            Cx D;
            if (i == 0) D = kron_elem(rot[1], hd::mul(d[w], rzm), r, c);      // angles 1-3: r1 in l4 (qubit K)
            else if (i == 1) D = kron_elem(d[w], lo4, r, c);                  // angles 4-6: r2 in l4 (qubit K+1)
            else if (i == 2) D = kron_elem(hi1, d[w], r, c);                  // angles 7-9: r3 in l1 (qubit K)
            else D = kron_elem(hd::mul(rzm, d[w]), rot[2], r, c);             // angles 10-12: r4 in l1 (qubit K+1)
            out[3 * i + w] = dot_re(D, i < 2 ? W4 : W1);
        }
    }
    if (gid < ngates && e == 0) {
#pragma unroll
        for (int k = 0; k < 15; k++) grads[15 * (size_t)g + k] = out[k];
    }
}
#endif

} // namespace qcb
