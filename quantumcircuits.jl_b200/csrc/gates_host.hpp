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

namespace qcb {

using cd = std::complex<double>;

struct M2 { cd m[4]; };
struct M4 { cd m[16]; };

inline M2 mul(const M2 &a, const M2 &b)
{
    M2 o;
    for (int c = 0; c < 2; c++)
        for (int r = 0; r < 2; r++) o.m[r + 2 * c] = a.m[r] * b.m[2 * c] + a.m[r + 2] * b.m[1 + 2 * c];
    return o;
}
inline M4 mul(const M4 &a, const M4 &b)
{
    M4 o;
    for (int c = 0; c < 4; c++)
        for (int r = 0; r < 4; r++) {
            cd s = 0;
            for (int k = 0; k < 4; k++) s += a.m[r + 4 * k] * b.m[k + 4 * c];
            o.m[r + 4 * c] = s;
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
                for (int rl = 0; rl < 2; rl++) o.m[(2 * rh + rl) + 4 * (2 * ch + cl)] = hi.m[rh + 2 * ch] * lo.m[rl + 2 * cl];
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
