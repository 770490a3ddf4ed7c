/*
 * oracle/ref_quad.cpp -- TEST INFRASTRUCTURE ONLY.
 * Higher-precision truth for the AMIPS parity question: the reference's own text,
 * src/tetwild/LocalOperations.cpp:28-291 (the same line-range extract oracle/ref_build.sh writes to oracle/_ref/gen/),
 * compiled a second time with every `double` turned into IEEE binary128 (__float128, libquadmath). The numeric literals of
 * the text stay what the reference's compiler sees -- doubles (0.577350269189626, -0.333333333333333, ...) -- so the result is
 * the value of the reference's expression AS WRITTEN, to ~1e-30, for the same double inputs.
 * Second entry point: the mathematical conformal AMIPS energy (exact constants: E = tr(J^T J) / det(J)^(2/3) against the
 * regular tetrahedron) and its gradient / Hessian w.r.t. vertex 0, also in binary128, in edge-vector form.
 * Built into oracle/_ref/libtetwild_ref_quad.so (git-ignored, travels to the GPU box).
 */
#include <quadmath.h>
#include <cstdint>

typedef __float128 q128;

namespace tetwild_q {
inline q128 pow(q128 a, q128 b) { return powq(a, b); }
struct LocalOperations {
    static q128 comformalAMIPSEnergy_new(const q128* T);
    static void comformalAMIPSJacobian_new(const q128* T, q128* result_0);
    static void comformalAMIPSHessian_new(const q128* T, q128* result_0);
};
#define double q128
#include "amips_lines.inc"
#undef double
}  // namespace tetwild_q

extern "C" {

/* T: 12 SoA arrays of n doubles (the energy_ispc layout); outputs rounded to double (round-to-nearest of the binary128 value) */
void refq_amips_ejh_soa(const double* const* Ts, double* E, double* J3, double* H9, uint64_t n, int threads) {
#pragma omp parallel for schedule(static) num_threads(threads > 0 ? threads : 1)
    for (int64_t i = 0; i < (int64_t)n; ++i) {
        q128 T[12], j[3], h[9];
        for (int k = 0; k < 12; ++k) T[k] = (q128)Ts[k][i];
        if (E) E[i] = (double)tetwild_q::LocalOperations::comformalAMIPSEnergy_new(T);
        if (J3) { tetwild_q::LocalOperations::comformalAMIPSJacobian_new(T, j); for (int c = 0; c < 3; ++c) J3[3 * i + c] = (double)j[c]; }
        if (H9) { tetwild_q::LocalOperations::comformalAMIPSHessian_new(T, h); for (int c = 0; c < 9; ++c) H9[9 * i + c] = (double)h[c]; }
    }
}

/* the mathematical function (exact constants), binary128: with e_i = x_i - x_0, m = -(e1+e2+e3), n = -(x2-x1)x(x3-x1),
 * q = (4 sum|e_i|^2 - |m|^2)/2, d = -n.e1, f = (2 d^2)^(-1/3):  E = q f,  J = f (m - (2/3)(q/d) n),
 * H = f (3 I - (2/3)/d (m n^T + n m^T) + (10/9)(q/d^2) n n^T)   (same derivation as csrc/tw_math.cuh::amips_eval) */
void exactq_amips_ejh_soa(const double* const* Ts, double* E, double* J3, double* H9, uint64_t n, int threads) {
#pragma omp parallel for schedule(static) num_threads(threads > 0 ? threads : 1)
    for (int64_t i = 0; i < (int64_t)n; ++i) {
        q128 x[12];
        for (int k = 0; k < 12; ++k) x[k] = (q128)Ts[k][i];
        q128 e[3][3], m[3], nn[3], a[3], b[3];
        for (int v = 0; v < 3; ++v)
            for (int c = 0; c < 3; ++c) e[v][c] = x[3 * (v + 1) + c] - x[c];
        q128 s2 = 0, mm = 0;
        for (int c = 0; c < 3; ++c) {
            m[c] = -(e[0][c] + e[1][c] + e[2][c]);
            mm += m[c] * m[c];
            for (int v = 0; v < 3; ++v) s2 += e[v][c] * e[v][c];
            a[c] = e[1][c] - e[0][c];
            b[c] = e[2][c] - e[0][c];
        }
        const q128 q = (4 * s2 - mm) / 2;
        nn[0] = -(a[1] * b[2] - a[2] * b[1]); nn[1] = -(a[2] * b[0] - a[0] * b[2]); nn[2] = -(a[0] * b[1] - a[1] * b[0]);
        const q128 d = -(nn[0] * e[0][0] + nn[1] * e[0][1] + nn[2] * e[0][2]);
        const q128 f = powq(2 * d * d, -(q128)1 / (q128)3);
        if (E) E[i] = (double)(q * f);
        if (J3)
            for (int c = 0; c < 3; ++c) J3[3 * i + c] = (double)(f * (m[c] - ((q128)2 / (q128)3) * (q / d) * nn[c]));
        if (H9)
            for (int r = 0; r < 3; ++r)
                for (int c = 0; c < 3; ++c)
                    H9[9 * i + 3 * r + c] = (double)(f * ((q128)(r == c ? 3 : 0) - ((q128)2 / (q128)3) / d * (m[r] * nn[c] + nn[r] * m[c]) + ((q128)10 / (q128)9) * (q / (d * d)) * nn[r] * nn[c]));
    }
}

}  // extern "C"
