// gates_host.hpp -- host-side gate algebra of the backend (tiny 2x2 / 4x4 complex work).
//
// Mirrors src/gates.jl of the reference: rotation gates (:36-73), the constant CNOT /
// ReverseCNOT tensors (:10-16) and the 15-angle general two-qubit gate (:80-109), plus the
// analytic derivative of that gate with respect to each angle, which the adjoint gradient
// sweep needs (the reference obtains the same numbers with +-pi/2
// shifts, src/gradients.jl:120-139).
//
// Matrix layout everywhere: column-major, M[r + 4c] (2x2: m[r + 2c]); for a 2-qubit gate on
// (K, K+1) the row/column index is (bit of K) + 2 (bit of K+1) -- the memory of the Julia
// 2x2x2x2 array u[i,j,a,b].
#pragma once
#include <cmath>
#include <complex>
#include <utility>

namespace qcb {

using cd = std::complex<double>;

struct M2 { cd m[4]; };
struct M4 { cd m[16]; };

// complex product without std::complex's operator* (which goes through the NaN-recovering __muldc3: 10x slower, and this
// algebra runs 15 G times per gradient on the host)
inline cd cmul(const cd &x, const cd &y) { return cd(x.real() * y.real() - x.imag() * y.imag(), x.real() * y.imag() + x.imag() * y.real()); }

inline M2 mul(const M2 &a, const M2 &b)
{
    M2 o;
    for (int c = 0; c < 2; c++)
        for (int r = 0; r < 2; r++) o.m[r + 2 * c] = cmul(a.m[r], b.m[2 * c]) + cmul(a.m[r + 2], b.m[1 + 2 * c]);
    return o;
}
inline M4 mul(const M4 &a, const M4 &b)
{
    M4 o;
    for (int c = 0; c < 4; c++)
        for (int r = 0; r < 4; r++) {
            double sr = 0, si = 0;
            for (int k = 0; k < 4; k++) {
                const cd &x = a.m[r + 4 * k], &y = b.m[k + 4 * c];
                sr += x.real() * y.real() - x.imag() * y.imag();
                si += x.real() * y.imag() + x.imag() * y.real();
            }
            o.m[r + 4 * c] = cd(sr, si);
        }
    return o;
}
// kron(hi, lo): `lo` acts on qubit K (the low index bit), `hi` on K+1.
inline M4 kron(const M2 &hi, const M2 &lo)
{
    M4 o;
    for (int ch = 0; ch < 2; ch++)
        for (int cl = 0; cl < 2; cl++)
            for (int rh = 0; rh < 2; rh++)
                for (int rl = 0; rl < 2; rl++) o.m[(2 * rh + rl) + 4 * (2 * ch + cl)] = cmul(hi.m[rh + 2 * ch], lo.m[rl + 2 * cl]);
    return o;
}

inline M2 identity2() { return M2{{1, 0, 0, 1}}; }
// RZ(a) = diag(e^{-ia/2}, e^{+ia/2})   src/gates.jl:36-43
inline M2 rz(double a)
{
    const double c = std::cos(a / 2), s = std::sin(a / 2);
    return M2{{cd(c, -s), 0, 0, cd(c, s)}};
}
// RY(a) = [[c, -s], [s, c]]            src/gates.jl:44-51
inline M2 ry(double a)
{
    const double c = std::cos(a / 2), s = std::sin(a / 2);
    return M2{{c, s, -s, c}};
}
// d/da of RZ / RY: both are exp(-i a P / 2), so the derivative is R(a + pi) / 2.
inline M2 scale(const M2 &a, double f) { return M2{{a.m[0] * f, a.m[1] * f, a.m[2] * f, a.m[3] * f}}; }
inline M2 drz(double a) { return scale(rz(a + M_PI), 0.5); }
inline M2 dry(double a) { return scale(ry(a + M_PI), 0.5); }

// RotationGate(theta, phi, lambda) = [[c, -e^{i lam} s], [e^{i phi} s, e^{i(phi+lam)} c]]   src/gates.jl:61-73
// which = -1: the gate; 0/1/2: derivative with respect to theta / phi / lambda.
inline M2 rotation(double theta, double phi, double lam, int which = -1)
{
    const double c = std::cos(theta / 2), s = std::sin(theta / 2);
    const cd p1(std::cos(lam), std::sin(lam)), p2(std::cos(phi), std::sin(phi));
    const cd p3 = p1 * p2;
    const cd I(0, 1);
    switch (which) {
    case 0: return M2{{-0.5 * s, 0.5 * p2 * c, -0.5 * p1 * c, -0.5 * s * p3}};
    case 1: return M2{{0, I * p2 * s, 0, I * p3 * c}};
    case 2: return M2{{0, 0, -I * p1 * s, I * p3 * c}};
    default: return M2{{c, p2 * s, -p1 * s, c * p3}};
    }
}

inline M4 cnot4()
{
    M4 o{};
    o.m[0 + 4 * 0] = 1; o.m[3 + 4 * 1] = 1; o.m[2 + 4 * 2] = 1; o.m[1 + 4 * 3] = 1; // control K, target K+1
    return o;
}
inline M4 rcnot4()
{
    M4 o{};
    o.m[0 + 4 * 0] = 1; o.m[1 + 4 * 1] = 1; o.m[3 + 4 * 2] = 1; o.m[2 + 4 * 3] = 1; // control K+1, target K
    return o;
}

// src/gates.jl:80-109.  dk = -1: the gate; dk = 0..14: its partial derivative w.r.t. angle dk
// (one factor replaced by its derivative; the association order of the product is the reference's).
inline M4 general_unitary(const double *a, int dk = -1)
{
    auto rot = [&](int i) { // i-th RotationGate uses angles 3i..3i+2
        const int w = (dk >= 3 * i && dk < 3 * i + 3) ? dk - 3 * i : -1;
        return rotation(a[3 * i], a[3 * i + 1], a[3 * i + 2], w);
    };
    const M2 r1 = rot(0), r2 = rot(1), r3 = rot(2), r4 = rot(3);
    const M2 rzm = rz(-M_PI / 2);
    const M2 z13 = dk == 12 ? drz(a[12]) : rz(a[12]); // theta
    const M2 y14 = dk == 13 ? dry(a[13]) : ry(a[13]); // phi
    const M2 y15 = dk == 14 ? dry(a[14]) : ry(a[14]); // lambda
    const M4 l1 = kron(mul(rzm, r4), r3);
    const M4 l2 = kron(y14, z13);
    const M4 l3 = kron(y15, identity2());
    const M4 l4 = kron(r2, mul(r1, rzm));
    const M4 rc = rcnot4(), cn = cnot4();
    return mul(mul(mul(l4, rc), mul(l3, cn)), mul(mul(l2, rc), l1));
}

inline M4 transpose(const M4 &a)
{
    M4 o;
    for (int c = 0; c < 4; c++)
        for (int r = 0; r < 4; r++) o.m[r + 4 * c] = a.m[c + 4 * r];
    return o;
}

// out15[k] = 2 Re sum_{r,c} env[r,c] dU/dtheta_k [r,c] for all 15 angles of the general gate at once.
// U = l4 rc l3 cn l2 rc l1 and every angle sits in one Kronecker factor of one l_j, so with dU = X_j dl_j Y_j
//     sum env .* dU = sum_{a,b} dl_j[a,b] W_j[a,b],   W_j = X_j^T env Y_j^T :
// four W matrices (shared prefix / suffix products) and one 16-term contraction per angle, instead of 15 full products of
// seven 4x4 matrices (the per-angle form, general_unitary(a, k), costs 1 us per angle on the host: 0.6 ms of a 0.9 ms
// 20-qubit gradient).  Same numbers up to the association order of the products.
inline void general_unitary_gradient(const double *a, const double *env32, double *out15)
{
    M2 r[4], dr[4][3];
    for (int i = 0; i < 4; i++) { // (one set of sines and cosines per rotation: the trigonometry is most of this function's time)
        const double c = std::cos(a[3 * i] / 2), sn = std::sin(a[3 * i] / 2);
        const cd p1(std::cos(a[3 * i + 2]), std::sin(a[3 * i + 2])), p2(std::cos(a[3 * i + 1]), std::sin(a[3 * i + 1]));
        const cd p3 = cmul(p1, p2), I(0, 1);
        r[i] = M2{{c, p2 * sn, -p1 * sn, c * p3}};
        dr[i][0] = M2{{-0.5 * sn, 0.5 * p2 * c, -0.5 * p1 * c, -0.5 * sn * p3}};
        dr[i][1] = M2{{0, cmul(I, p2) * sn, 0, cmul(I, p3) * c}};
        dr[i][2] = M2{{0, 0, -cmul(I, p1) * sn, cmul(I, p3) * c}};
    }
    const M2 rzm = rz(-M_PI / 2), id = identity2();
    const M2 z13 = rz(a[12]), y14 = ry(a[13]), y15 = ry(a[14]);
    // d/da exp(-i a P / 2) = R(a + pi) / 2, and cos((a + pi) / 2) = -sin(a / 2), sin((a + pi) / 2) = cos(a / 2): no new trigonometry
    const M2 dz13 = M2{{0.5 * cd(-z13.m[3].imag(), -z13.m[3].real()), 0, 0, 0.5 * cd(-z13.m[3].imag(), z13.m[3].real())}};
    auto dry_of = [](const M2 &y) { return M2{{-0.5 * y.m[1], 0.5 * y.m[0], -0.5 * y.m[0], -0.5 * y.m[1]}}; };
    const M2 hi1 = mul(rzm, r[3]), lo4 = mul(r[0], rzm);
    const M4 l1 = kron(hi1, r[2]), l2 = kron(y14, z13), l3 = kron(y15, id), l4 = kron(r[1], lo4);
    M4 E;
    for (int i = 0; i < 16; i++) E.m[i] = cd(env32[2 * i], env32[2 * i + 1]);
    // suffixes Y_j (everything to the right of l_j) and prefixes X_j (everything to its left).  The CNOTs are permutation
    // matrices: rcnot swaps index 2 <-> 3, cnot swaps 1 <-> 3 (rows when applied from the left, columns from the right)
    auto rows = [](M4 m, int i, int j) { for (int c = 0; c < 4; c++) std::swap(m.m[i + 4 * c], m.m[j + 4 * c]); return m; };
    auto cols = [](M4 m, int i, int j) { for (int r = 0; r < 4; r++) std::swap(m.m[r + 4 * i], m.m[r + 4 * j]); return m; };
    const M4 Y2 = rows(l1, 2, 3), Y3 = rows(mul(l2, Y2), 1, 3), Y4 = rows(mul(l3, Y3), 2, 3);
    const M4 X3 = cols(l4, 2, 3), X2 = cols(mul(X3, l3), 1, 3), X1 = cols(mul(X2, l2), 2, 3);
    const M4 W4 = mul(E, transpose(Y4));                        // X_4 = 1
    const M4 W3 = mul(mul(transpose(X3), E), transpose(Y3));
    const M4 W2 = mul(mul(transpose(X2), E), transpose(Y2));
    const M4 W1 = mul(transpose(X1), E);                        // Y_1 = 1
    auto dot = [](const M4 &d, const M4 &w) {
        double s = 0;
        for (int i = 0; i < 16; i++) s += d.m[i].real() * w.m[i].real() - d.m[i].imag() * w.m[i].imag();
        return 2 * s;
    };
    for (int w = 0; w < 3; w++) {
        out15[w] = dot(kron(r[1], mul(dr[0][w], rzm)), W4);      // angles 1-3: r1 in l4 (qubit K)
        out15[3 + w] = dot(kron(dr[1][w], lo4), W4);             // angles 4-6: r2 in l4 (qubit K+1)
        out15[6 + w] = dot(kron(hi1, dr[2][w]), W1);             // angles 7-9: r3 in l1 (qubit K)
        out15[9 + w] = dot(kron(mul(rzm, dr[3][w]), r[2]), W1);  // angles 10-12: r4 in l1 (qubit K+1)
    }
    out15[12] = dot(kron(y14, dz13), W2);                        // theta: RZ in l2 (qubit K)
    out15[13] = dot(kron(dry_of(y14), z13), W2);                 // phi: RY in l2 (qubit K+1)
    out15[14] = dot(kron(dry_of(y15), id), W3);                  // lambda: RY in l3 (qubit K+1)
}

inline void to_doubles(const M4 &u, double *out32)
{
    for (int i = 0; i < 16; i++) { out32[2 * i] = u.m[i].real(); out32[2 * i + 1] = u.m[i].imag(); }
}

// 1-qubit gate u (column-major u[out + 2 in], 8 doubles) embedded as a 2-qubit gate:
// on_low = true: acts on the low qubit of the pair (I (x) u), else on the high qubit (u (x) I).
inline void embed_1q(const double *u8, bool on_low, double *out32)
{
    M2 u;
    for (int i = 0; i < 4; i++) u.m[i] = cd(u8[2 * i], u8[2 * i + 1]);
    to_doubles(on_low ? kron(identity2(), u) : kron(u, identity2()), out32);
}

} // namespace qcb
