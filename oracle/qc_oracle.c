/*
 * qc_oracle.c -- CPU restatement of the QuantumCircuits.jl dense-statevector path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under quantumcircuits.jl_b200/ may link, import
 * or call this file; it is the checker for tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py.
 *
 * Parity pin: the reference is pure Julia and Julia is not installed in this image,
 * so the reference itself cannot be executed here and it ships no stored golden
 * vectors (its tests draw unseeded randn).  The oracle is pinned instead by
 * re-running every known-answer / equivalence test the reference holds for this path
 * against it (tests/test_oracle.py): the GHZ test, apply == dense Kronecker operator
 * for random non-unitary 1- and 2-qubit gates, East kernel == dense mat(H, n)
 * (test/correctness_tests.jl:3-150), kernel-TFIM == dense TFIM
 * (examples/test_gradients.jl:10-28, examples/matrix_tfim.jl:7-32) and
 * gradients! == naive per-angle shift (examples/test_gradient_accuracy.jl:25-57).
 * In the task statement's terms this is "PARITY UNPINNED": there is neither a golden
 * vector of the reference nor an output of the reference itself behind this file.
 * What stands in for them, besides the reference's own equivalence tests, are answers
 * that come from outside any code base (tests/test_oracle.py, tests/test_abi.py):
 * the analytic product state of a brickwork circuit without entangling angles
 * (tests/product_circuit.py: gate builder, qubit order, layer geometry, application),
 * the free-fermion ground energy of the open transverse-field Ising chain (TFIM
 * matrix and kernel), the Makhlin invariants of the general two-qubit gate, and
 * central differences for the gradients.
 *
 * Every function cites the reference lines it follows (paths relative to
 * /root/reference).  Layout conventions:
 *   - state: 2^n complex amplitudes, interleaved (re, im); qubit q (1-based) is bit
 *     q-1 of the 0-based linear index (column-major 2x2x...x2 tensor).
 *   - 1-qubit gate: 4 complex doubles, column-major u[out + 2*in].
 *   - 2-qubit gate: 16 complex doubles, column-major u[i + 2j + 4a + 8b],
 *     (i, a) <-> qubit K, (j, b) <-> qubit K+1, (i, j) = output, (a, b) = input.
 *   - dtype 0 = ComplexF32 state ("c64"), 1 = ComplexF64 state ("c128").  Gates are
 *     always ComplexF64 (src/apply.jl:100-101); a ComplexF32 state is accumulated in
 *     double and rounded once per gate on store (src/apply.jl:72,78).
 */
#include <complex.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef double complex zc;
typedef float complex fc;

#define QCO_OK 0
#define QCO_EINVAL 1

static int g_threads = 0; /* 0 = OpenMP default */

void qco_set_threads(int t) { g_threads = t; }
int qco_get_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
static int nthreads(void)
{
#ifdef _OPENMP
    return g_threads > 0 ? g_threads : omp_get_max_threads();
#else
    return 1;
#endif
}

/* KernelAbstractions CPU workgroup size, src/apply.jl:86 */
#define KA_CPU_WG 16

/* ------------------------------------------------------------------------- */
/* gates: src/gates.jl                                                        */
/* ------------------------------------------------------------------------- */

/* 2x2 complex matrices are column-major m[r + 2c]. */
static void m2_mul(const zc *a, const zc *b, zc *o)
{
    zc t[4];
    for (int c = 0; c < 2; c++)
        for (int r = 0; r < 2; r++) t[r + 2 * c] = a[r] * b[2 * c] + a[r + 2] * b[1 + 2 * c];
    memcpy(o, t, sizeof t);
}
/* 4x4 complex, column-major m[r + 4c]; plain triple loop, k ascending. */
static void m4_mul(const zc *a, const zc *b, zc *o)
{
    zc t[16];
    for (int c = 0; c < 4; c++)
        for (int r = 0; r < 4; r++) {
            zc s = 0;
            for (int k = 0; k < 4; k++) s += a[r + 4 * k] * b[k + 4 * c];
            t[r + 4 * c] = s;
        }
    memcpy(o, t, sizeof t);
}
/* kron(A, B) for 2x2: B indexes the low bit (qubit K), A the high bit (K+1). */
static void kron2(const zc *a, const zc *b, zc *o)
{
    for (int ca = 0; ca < 2; ca++)
        for (int cb = 0; cb < 2; cb++)
            for (int ra = 0; ra < 2; ra++)
                for (int rb = 0; rb < 2; rb++)
                    o[(2 * ra + rb) + 4 * (2 * ca + cb)] = a[ra + 2 * ca] * b[rb + 2 * cb];
}

/* src/gates.jl:36-43 */
void qco_rz(double angle, zc *m)
{
    double c = cos(angle / 2), s = sin(angle / 2);
    m[0] = c - s * I; m[1] = 0; m[2] = 0; m[3] = c + s * I;
}
/* src/gates.jl:44-51 */
void qco_ry(double angle, zc *m)
{
    double c = cos(angle / 2), s = sin(angle / 2);
    m[0] = c; m[1] = s; m[2] = -s; m[3] = c;
}
/* src/gates.jl:52-59 */
void qco_rx(double angle, zc *m)
{
    double c = cos(angle / 2);
    zc s = -sin(angle / 2) * I;
    m[0] = c; m[1] = s; m[2] = s; m[3] = c;
}
/* src/gates.jl:61-73 */
void qco_rotation(double theta, double phi, double lambda, zc *m)
{
    double c = cos(theta / 2), s = sin(theta / 2);
    zc p1 = cos(lambda) + sin(lambda) * I; /* exp(im*lambda) */
    zc p2 = cos(phi) + sin(phi) * I;       /* exp(im*phi)    */
    zc p3 = p1 * p2;
    m[0] = c;      m[2] = -p1 * s;
    m[1] = p2 * s; m[3] = c * p3;
}

/* src/gates.jl:10-16 ; 4x4 form with qubit K as the low bit */
static const double CNOT4[16]  = {1,0,0,0, 0,0,0,1, 0,0,1,0, 0,1,0,0};
static const double RCNOT4[16] = {1,0,0,0, 0,1,0,0, 0,0,0,1, 0,0,1,0};

void qco_const_gate_2q(int which, zc *u) /* 0 = CNOT, 1 = ReverseCNOT */
{
    const double *s = which ? RCNOT4 : CNOT4;
    for (int i = 0; i < 16; i++) u[i] = s[i];
}

/* src/gates.jl:80-109 : 15 angles -> 4x4 (column-major == the 2x2x2x2 array). */
void qco_build_general_unitary_gate(const double *ang, zc *out)
{
    zc r1[4], r2[4], r3[4], r4[4], rzm[4], t[4], a[4], b[4], id[4] = {1, 0, 0, 1};
    zc l1[16], l2[16], l3[16], l4[16], cn[16], rc[16], x[16], y[16], z[16], w[16];
    qco_rotation(ang[0], ang[1], ang[2], r1);
    qco_rotation(ang[3], ang[4], ang[5], r2);
    qco_rotation(ang[6], ang[7], ang[8], r3);
    qco_rotation(ang[9], ang[10], ang[11], r4);
    double theta = ang[12], phi = ang[13], lambda = ang[14];
    qco_rz(-M_PI / 2, rzm);
    m2_mul(rzm, r4, t);            kron2(t, r3, l1);          /* :98  */
    qco_ry(phi, a); qco_rz(theta, b); kron2(a, b, l2);        /* :99  */
    qco_ry(lambda, a);             kron2(a, id, l3);          /* :100 */
    m2_mul(r1, rzm, t);            kron2(r2, t, l4);          /* :101 */
    qco_const_gate_2q(0, cn);
    qco_const_gate_2q(1, rc);
    m4_mul(l4, rc, x);  /* out = l4*rcnot            :106 */
    m4_mul(l3, cn, y);  /* out *= (l3*cnot)          :107 */
    m4_mul(x, y, z);
    m4_mul(l2, rc, x);  /* out *= (l2*rcnot)*l1      :108 */
    m4_mul(x, l1, w);
    m4_mul(z, w, out);
}

/* src/gradients.jl:27-32 : conjugate transpose of the 4x4 form. */
void qco_adjoint_2q(const zc *u, zc *o)
{
    zc t[16];
    for (int r = 0; r < 4; r++)
        for (int c = 0; c < 4; c++) t[c + 4 * r] = conj(u[r + 4 * c]);
    memcpy(o, t, sizeof t);
}

/* ------------------------------------------------------------------------- */
/* circuit geometry: src/circuits.jl:9-12                                      */
/* ------------------------------------------------------------------------- */
/* layer (1-based) starts at qubit 1 + (l-1)%2, step 2, up to nbits-1. */
static int layer_first(int l) { return 1 + (l - 1) % 2; }
static int layer_count(int l, int nbits)
{
    int f = layer_first(l), last = nbits - 1;
    return last >= f ? (last - f) / 2 + 1 : 0;
}
int qco_brickwork_num_gates(int nbits, int nlayers)
{
    int g = 0;
    for (int l = 1; l <= nlayers; l++) g += layer_count(l, nbits);
    return g;
}

/* ------------------------------------------------------------------------- */
/* gate application                                                            */
/* ------------------------------------------------------------------------- */

/* src/apply.jl:62-84 body for one 0-based output index, generic over state type. */
#define APPLY2_BODY(T, CT)                                                              \
    const int64_t lo = (int64_t)1 << (K - 1);                                           \
    const int64_t N = (int64_t)1 << n;                                                  \
    const int64_t ngroups = (N + KA_CPU_WG - 1) / KA_CPU_WG;                            \
    _Pragma("omp parallel for schedule(static) num_threads(nthreads())")                \
    for (int64_t g = 0; g < ngroups; g++) {                                             \
        int64_t end = (g + 1) * KA_CPU_WG < N ? (g + 1) * KA_CPU_WG : N;                \
        for (int64_t idx = g * KA_CPU_WG; idx < end; idx++) {                           \
            int64_t pre = idx % lo;                                                     \
            int i = (int)((idx >> (K - 1)) & 1), j = (int)((idx >> K) & 1);             \
            int64_t post = (idx >> (K + 1)) << (K + 1);                                 \
            zc acc = 0;                                                                 \
            for (int a = 0; a < 2; a++)                                                 \
                for (int b = 0; b < 2; b++) {                                           \
                    int64_t src = pre + ((int64_t)a << (K - 1)) + ((int64_t)b << K) + post; \
                    acc += (zc)in[src] * u[i + 2 * j + 4 * a + 8 * b];                  \
                }                                                                       \
            out[idx] = (CT)acc;                                                         \
        }                                                                               \
    }

static void zero_fill(int dtype, int n, void *p)
{
    /* src/apply.jl:135 : psi' .= 0 before every kernel launch (a full-state write). */
    size_t bytes = ((size_t)1 << n) * (dtype ? sizeof(zc) : sizeof(fc));
    size_t nt = (size_t)nthreads();
    size_t chunk = (bytes + nt - 1) / nt;
    #pragma omp parallel for schedule(static) num_threads(nthreads())
    for (size_t t = 0; t < nt; t++) {
        size_t b = t * chunk, e = b + chunk < bytes ? b + chunk : bytes;
        if (b < e) memset((char *)p + b, 0, e - b);
    }
}

/* src/apply.jl:117-142 (cache path, KA kernel).  K is 1-based, acts on K and K+1. */
int qco_apply_2q(int dtype, int n, void *out_, const void *in_, const zc *u, int K)
{
    if (!(K > 0 && K < n)) return QCO_EINVAL; /* src/apply.jl:112 */
    zero_fill(dtype, n, out_);
    if (dtype) {
        zc *out = out_; const zc *in = in_;
        APPLY2_BODY(double, zc)
    } else {
        fc *out = out_; const fc *in = in_;
        APPLY2_BODY(float, fc)
    }
    return QCO_OK;
}

/* src/apply.jl:39-60 (3-argument generic scatter path; input-major accumulation). */
int qco_apply_2q_scatter(int dtype, int n, void *out_, const void *in_, const zc *u, int K)
{
    if (!(K > 0 && K < n)) return QCO_EINVAL; /* src/apply.jl:45 */
    const int64_t N = (int64_t)1 << n;
    zc *acc = calloc((size_t)N, sizeof(zc));
    if (!acc) return QCO_EINVAL;
    for (int64_t idx = 0; idx < N; idx++) {
        zc psi = dtype ? ((const zc *)in_)[idx] : (zc)((const fc *)in_)[idx];
        int a = (int)((idx >> (K - 1)) & 1), b = (int)((idx >> K) & 1);
        int64_t base = idx & ~(((int64_t)3) << (K - 1));
        for (int j = 0; j < 2; j++)
            for (int i = 0; i < 2; i++) {
                int64_t dst = base | ((int64_t)i << (K - 1)) | ((int64_t)j << K);
                zc v = psi * u[i + 2 * j + 4 * a + 8 * b];
                if (dtype) acc[dst] += v;
                else acc[dst] = (zc)(fc)(acc[dst] + v); /* psi' is ComplexF32: rounds per += */
            }
    }
    for (int64_t idx = 0; idx < N; idx++) {
        if (dtype) ((zc *)out_)[idx] = acc[idx];
        else ((fc *)out_)[idx] = (fc)acc[idx];
    }
    free(acc);
    return QCO_OK;
}

/* src/apply.jl:20-38 (the reference's only 1-qubit path; generic scatter). */
int qco_apply_1q(int dtype, int n, void *out_, const void *in_, const zc *u, int K)
{
    if (!(K >= 1 && K <= n)) return QCO_EINVAL; /* src/apply.jl:27 */
    const int64_t N = (int64_t)1 << n;
    zc *acc = calloc((size_t)N, sizeof(zc));
    if (!acc) return QCO_EINVAL;
    for (int64_t idx = 0; idx < N; idx++) {
        zc psi = dtype ? ((const zc *)in_)[idx] : (zc)((const fc *)in_)[idx];
        int c = (int)((idx >> (K - 1)) & 1);
        int64_t base = idx & ~((int64_t)1 << (K - 1));
        for (int i = 0; i < 2; i++) {
            int64_t dst = base | ((int64_t)i << (K - 1));
            zc v = psi * u[i + 2 * c];
            if (dtype) acc[dst] += v;
            else acc[dst] = (zc)(fc)(acc[dst] + v);
        }
    }
    for (int64_t idx = 0; idx < N; idx++) {
        if (dtype) ((zc *)out_)[idx] = acc[idx];
        else ((fc *)out_)[idx] = (fc)acc[idx];
    }
    free(acc);
    return QCO_OK;
}

/* ------------------------------------------------------------------------- */
/* brickwork circuit driver: src/circuits.jl:19-31                             */
/* ------------------------------------------------------------------------- */
/* Applies gates [first_gate, ngates) given (layer, position-in-layer) of first_gate;
 * ping-pongs a/b, returns 0 if the result is in a, 1 if in b.  This one routine is
 * both apply!(cache, psi', psi, circuit) (start at layer 1) and
 * propagate_forwards! (src/gradients.jl:54-71). */
static int propagate(int dtype, int nbits, int nlayers, const double *angles, void *a,
                     void *b, int start_layer, int layer_gate_idx /*1-based*/, int gate_idx /*0-based*/)
{
    void *cur = a, *other = b;
    for (int l = start_layer; l <= nlayers; l++) {
        int cnt = layer_count(l, nbits), f = layer_first(l);
        int p0 = (l == start_layer) ? layer_gate_idx : 1;
        for (int p = p0; p <= cnt; p++) {
            zc u[16];
            qco_build_general_unitary_gate(angles + 15 * (size_t)gate_idx, u);
            qco_apply_2q(dtype, nbits, other, cur, u, f + 2 * (p - 1));
            void *t = cur; cur = other; other = t;
            gate_idx++;
        }
    }
    return cur == a ? 0 : 1;
}

/* psi holds the input and receives the output; scratch is a second state buffer.
 * gate range [g_begin, g_end) lets the bounded CPU baseline time a slice of a circuit. */
int qco_apply_brickwork_range(int dtype, int nbits, int nlayers, const double *angles, void *psi,
                              void *scratch, int g_begin, int g_end)
{
    if (nbits < 2) return QCO_EINVAL;
    int gate_idx = 0;
    void *cur = psi, *other = scratch;
    for (int l = 1; l <= nlayers; l++) {
        int cnt = layer_count(l, nbits), f = layer_first(l);
        for (int p = 1; p <= cnt; p++, gate_idx++) {
            if (gate_idx < g_begin || gate_idx >= g_end) continue;
            zc u[16];
            qco_build_general_unitary_gate(angles + 15 * (size_t)gate_idx, u);
            qco_apply_2q(dtype, nbits, other, cur, u, f + 2 * (p - 1));
            void *t = cur; cur = other; other = t;
        }
    }
    if (cur != psi) memcpy(psi, cur, ((size_t)1 << nbits) * (dtype ? sizeof(zc) : sizeof(fc)));
    return QCO_OK;
}
int qco_apply_brickwork(int dtype, int nbits, int nlayers, const double *angles, void *psi, void *scratch)
{
    return qco_apply_brickwork_range(dtype, nbits, nlayers, angles, psi, scratch, 0,
                                     qco_brickwork_num_gates(nbits, nlayers));
}

/* ------------------------------------------------------------------------- */
/* kernel Hamiltonians                                                         */
/* ------------------------------------------------------------------------- */

/* Julia's sum() over an array is pairwise with a 1024-element sequential base case
 * (Base.mapreduce_impl); restated here for both element types. */
static zc pairwise_z(const zc *v, int64_t lo, int64_t hi)
{
    if (hi - lo <= 1024) { zc s = 0; for (int64_t i = lo; i < hi; i++) s += v[i]; return s; }
    int64_t mid = lo + ((hi - lo) >> 1);
    return pairwise_z(v, lo, mid) + pairwise_z(v, mid, hi);
}
static fc pairwise_f(const fc *v, int64_t lo, int64_t hi)
{
    if (hi - lo <= 1024) { fc s = 0; for (int64_t i = lo; i < hi; i++) s += v[i]; return s; }
    int64_t mid = lo + ((hi - lo) >> 1);
    return pairwise_f(v, lo, mid) + pairwise_f(v, mid, hi);
}
/* hamiltonians.jl:41 E = sum(psi'), threaded over fixed 2^20 blocks for the baseline. */
static zc sum_state(int dtype, int n, const void *p)
{
    const int64_t N = (int64_t)1 << n, B = (int64_t)1 << 20;
    if (N <= B) return dtype ? pairwise_z(p, 0, N) : (zc)pairwise_f(p, 0, N);
    int64_t nb = N / B;
    zc *part = malloc((size_t)nb * sizeof(zc));
    #pragma omp parallel for schedule(static) num_threads(nthreads())
    for (int64_t b = 0; b < nb; b++)
        part[b] = dtype ? pairwise_z(p, b * B, (b + 1) * B) : (zc)pairwise_f(p, b * B, (b + 1) * B);
    zc s = pairwise_z(part, 0, nb);
    free(part);
    return s;
}

/* src/hamiltonians/tfim.jl:17-54 + src/hamiltonians/hamiltonians.jl:31-50.
 * scratch (psi') is clobbered exactly like the reference.  out_E = (re, im). */
int qco_measure_tfim(int dtype, int n, void *scratch, const void *psi_, double J, double h, double g,
                     double *out_E)
{
    const int64_t N = (int64_t)1 << n;
    #pragma omp parallel for schedule(static) num_threads(nthreads())
    for (int64_t grp = 0; grp < (N + KA_CPU_WG - 1) / KA_CPU_WG; grp++) {
        int64_t end = (grp + 1) * KA_CPU_WG < N ? (grp + 1) * KA_CPU_WG : N;
        for (int64_t C = grp * KA_CPU_WG; C < end; C++) {
            int zz = 0;
            for (int i = 0; i + 1 < n; i++) {                     /* :25-32 */
                int r = (int)((C >> i) & 3);
                zz += (r == 3 || r == 0);
                zz -= (r == 1 || r == 2);
            }
            int ones = __builtin_popcountll((unsigned long long)C);
            if (dtype) {
                const zc *psi = psi_;
                zc p = psi[C];
                double p2 = creal(p) * creal(p) + cimag(p) * cimag(p);
                double nn = zz * p2;                              /* :36 */
                zc x = 0;
                for (int i = 0; i < n; i++) x += conj(psi[C ^ ((int64_t)1 << i)]); /* :39-44 */
                zc tf = x * p;
                double z = (n - 2 * ones) * p2;                   /* :48-49 */
                ((zc *)scratch)[C] = (-J) * (g * tf + h * z + nn); /* :52 */
            } else {
                const fc *psi = psi_;
                fc p = psi[C];
                float p2 = crealf(p) * crealf(p) + cimagf(p) * cimagf(p);
                float nn = (float)zz * p2;
                fc x = 0;
                for (int i = 0; i < n; i++) x += conjf(psi[C ^ ((int64_t)1 << i)]);
                fc tf = x * p;
                float z = (float)(n - 2 * ones) * p2;
                /* Float64 Hamiltonian parameters promote the expression to ComplexF64;
                 * the store into the ComplexF32 scratch rounds it. */
                ((fc *)scratch)[C] = (fc)((-J) * (g * (zc)tf + h * (double)z + (double)nn));
            }
        }
    }
    zc E = sum_state(dtype, n, scratch);
    out_E[0] = creal(E); out_E[1] = cimag(E);
    return QCO_OK;
}

/* src/hamiltonians/east_model.jl:30-52 ; parameters as in the struct (c, A, B). */
int qco_measure_east(int dtype, int n, void *scratch, const void *psi_, double c, double A, double B,
                     double *out_E)
{
    const int64_t N = (int64_t)1 << n;
    #pragma omp parallel for schedule(static) num_threads(nthreads())
    for (int64_t grp = 0; grp < (N + KA_CPU_WG - 1) / KA_CPU_WG; grp++) {
        int64_t end = (grp + 1) * KA_CPU_WG < N ? (grp + 1) * KA_CPU_WG : N;
        for (int64_t C = grp * KA_CPU_WG; C < end; C++) {
            zc p = dtype ? ((const zc *)psi_)[C] : (zc)((const fc *)psi_)[C];
            int64_t f0 = C ^ 1;
            zc Ac; double Bc;
            if (dtype) {
                const zc *psi = psi_;
                Ac = conj(psi[f0]);                                /* :37 */
                Bc = 0.0 + (double)!(C & 1);                       /* :38 */
                for (int i = 1; i < n; i++) {                      /* :40-45 */
                    int njm1 = !((C >> (i - 1)) & 1), nj = !((C >> i) & 1);
                    Ac += (double)njm1 * conj(psi[C ^ ((int64_t)1 << i)]);
                    Bc += njm1 * nj + c * njm1;
                }
                Ac *= A; Bc *= B;                                  /* :47-48 */
                ((zc *)scratch)[C] = (Ac + (Bc + c) * conj(p)) * p; /* :50 */
            } else {
                const fc *psi = psi_;
                fc Af = conjf(psi[f0]);
                Bc = 0.0 + (double)!(C & 1);
                for (int i = 1; i < n; i++) {
                    int njm1 = !((C >> (i - 1)) & 1), nj = !((C >> i) & 1);
                    Af += (float)njm1 * conjf(psi[C ^ ((int64_t)1 << i)]);
                    Bc += njm1 * nj + c * njm1;
                }
                Ac = (zc)Af * A; Bc *= B;
                ((fc *)scratch)[C] = (fc)((Ac + (Bc + c) * conj(p)) * p);
            }
        }
    }
    zc E = sum_state(dtype, n, scratch);
    out_E[0] = creal(E); out_E[1] = cimag(E);
    return QCO_OK;
}

/* ham_kind 0 = TFIM params (J, h, g); 1 = East params (c, A, B). */
static double measure_kind(int dtype, int n, void *scratch, const void *psi, int ham_kind,
                           const double *hp)
{
    double E[2];
    if (ham_kind == 0) qco_measure_tfim(dtype, n, scratch, psi, hp[0], hp[1], hp[2], E);
    else qco_measure_east(dtype, n, scratch, psi, hp[0], hp[1], hp[2], E);
    return E[0]; /* hamiltonians.jl:49 real(E) */
}

/* ------------------------------------------------------------------------- */
/* gradients: src/gradients.jl:88-152 (the reference's O(G^2) shift sweep)     */
/* ------------------------------------------------------------------------- */
/* psi0: input state (not modified).  out_grads: 15 x G column-major.
 * The c64 energies are Float32 values like the reference's real(E).           */
int qco_gradients_brickwork(int dtype, int nbits, int nlayers, const double *angles_in, const void *psi0,
                            int ham_kind, const double *hp, int want_energy, double *out_E,
                            double *out_grads)
{
    if (nbits < 2) return QCO_EINVAL;
    const int G = qco_brickwork_num_gates(nbits, nlayers);
    const size_t bytes = ((size_t)1 << nbits) * (dtype ? sizeof(zc) : sizeof(fc));
    double *angles = malloc(sizeof(double) * 15 * (size_t)(G > 0 ? G : 1));
    void *psi = malloc(bytes), *psip = malloc(bytes), *psipp = malloc(bytes);
    if (!angles || !psi || !psip || !psipp) { free(angles); free(psi); free(psip); free(psipp); return QCO_EINVAL; }
    memcpy(angles, angles_in, sizeof(double) * 15 * (size_t)G);
    memcpy(psi, psi0, bytes);                                              /* :90 */
    if (propagate(dtype, nbits, nlayers, angles, psi, psip, 1, 1, 0)) {    /* :92-101 */
        void *t = psi; psi = psip; psip = t;
    }
    if (want_energy) *out_E = measure_kind(dtype, nbits, psip, psi, ham_kind, hp); /* :104 */

    int gate_idx = G - 1;
    for (int l = nlayers; l >= 1; l--) {
        int cnt = layer_count(l, nbits), f = layer_first(l);
        for (int p = cnt; p >= 1; p--, gate_idx--) {                       /* :107-108 */
            double *ang = angles + 15 * (size_t)gate_idx;
            zc u[16], ud[16];
            qco_build_general_unitary_gate(ang, u);
            qco_adjoint_2q(u, ud);                                         /* :110-111 */
            qco_apply_2q(dtype, nbits, psip, psi, ud, f + 2 * (p - 1));    /* :114 */
            { void *t = psi; psi = psip; psip = t; }
            memcpy(psipp, psi, bytes);                                     /* :118 */
            for (int k = 0; k < 15; k++) {                                 /* :120 */
                double orig = ang[k];
                ang[k] = orig + M_PI / 2;                                  /* :124 */
                if (propagate(dtype, nbits, nlayers, angles, psi, psip, l, p, gate_idx)) {
                    void *t = psi; psi = psip; psip = t;
                }
                double Ep = measure_kind(dtype, nbits, psip, psi, ham_kind, hp); /* :126 */
                memcpy(psi, psipp, bytes);                                 /* :129 */
                ang[k] = orig - M_PI / 2;                                  /* :130 */
                if (propagate(dtype, nbits, nlayers, angles, psi, psip, l, p, gate_idx)) {
                    void *t = psi; psi = psip; psip = t;
                }
                double Em = measure_kind(dtype, nbits, psip, psi, ham_kind, hp); /* :132 */
                if (dtype) out_grads[k + 15 * (size_t)gate_idx] = (Ep - Em) / 2;  /* :135 */
                else out_grads[k + 15 * (size_t)gate_idx] = (double)(((float)Ep - (float)Em) / 2);
                memcpy(psi, psipp, bytes);                                 /* :138 */
                ang[k] = orig;                                             /* :139 */
            }
        }
    }
    free(angles); free(psi); free(psip); free(psipp);
    return QCO_OK;
}

/* src/gradients.jl:1-17 with quirk Q1 resolved to the intended semantics
 * (calculate_energy=true; theta <- theta - lr * grad), see SURVEY.md 8a Q1.
 * energies has epochs+1 entries: [0] is never written by the reference (stays 0),
 * [i] (1 <= i < epochs) = energy at the START of epoch i, and [epochs] is overwritten
 * by the final measure(H, psi0, circuit) (:11,:15).                              */
int qco_optimise_brickwork(int dtype, int nbits, int nlayers, double *angles, const void *psi0,
                           int ham_kind, const double *hp, int epochs, double lr, double *energies)
{
    const int G = qco_brickwork_num_gates(nbits, nlayers);
    double *grads = malloc(sizeof(double) * 15 * (size_t)(G > 0 ? G : 1));
    if (!grads) return QCO_EINVAL;
    energies[0] = 0.0;
    for (int e = 1; e <= epochs; e++) {
        double E;
        int rc = qco_gradients_brickwork(dtype, nbits, nlayers, angles, psi0, ham_kind, hp, 1, &E, grads);
        if (rc) { free(grads); return rc; }
        energies[e] = E;
        for (size_t i = 0; i < 15 * (size_t)G; i++) angles[i] -= lr * grads[i];
    }
    /* :15  energies[end] = measure(H, psi0, circuit) */
    const size_t bytes = ((size_t)1 << nbits) * (dtype ? sizeof(zc) : sizeof(fc));
    void *a = malloc(bytes), *b = malloc(bytes);
    memcpy(a, psi0, bytes);
    qco_apply_brickwork(dtype, nbits, nlayers, angles, a, b);
    energies[epochs] = measure_kind(dtype, nbits, b, a, ham_kind, hp);
    free(a); free(b); free(grads);
    return QCO_OK;
}
