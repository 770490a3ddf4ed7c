/*
 * oracle/amips.c -- TEST INFRASTRUCTURE ONLY (CPU oracle; never on the product path).
 *
 * Restates the reference's conformal AMIPS energy of one tetrahedron and its first/second derivatives with
 * respect to vertex 0:
 *   energy    src/tetwild/LocalOperations.cpp:28-81    (comformalAMIPSEnergy_new)   == src/ispc/energy.ispc:23-63
 *   gradient  src/tetwild/LocalOperations.cpp:83-147   (comformalAMIPSJacobian_new)
 *   Hessian   src/tetwild/LocalOperations.cpp:149-291  (comformalAMIPSHessian_new)
 * and the gating / reduction logic wrapped around them:
 *   calTetQuality_AMIPS  LocalOperations.cpp:862-884   (orientation gate, MAX_ENERGY clamp)
 *   NewtonsUpdate        VertexSmoother.cpp:627-702    (one-ring E/J/H sum with the vertex rotated to slot 0)
 *   getNewEnergy         VertexSmoother.cpp:606-622    (one-ring energy sum in stored order + clamp)
 *
 * The reference functions are machine-generated straight-line code for
 *       E(T) = Q(T) * (det(T)^2)^(-0.333333333333333)
 * with Q the half sum of squared edge lengths written as a quadratic form in absolute coordinates and det the
 * determinant of the regular-tet-normalised edge matrix (constants 1/sqrt3, 2/sqrt3, 1/sqrt6, 3/sqrt6 as 15-digit
 * literals). The oracle evaluates that same expression (same literals, same absolute-coordinate formulation) and
 * obtains gradient and Hessian by forward-mode automatic differentiation (second-order jets in the 3 coordinates
 * of vertex 0) instead of restating ~250 generated lines; AD yields the exact derivatives of the expression, which is
 * what the generated code encodes. It is pinned against the reference's own text via oracle/_ref (tests/test_oracle_pin.py).
 */
#include <math.h>
#include <stddef.h>
#include "tw_oracle.h"

#pragma STDC FP_CONTRACT OFF

static const double K3 = 0.577350269189626;  /* 1/sqrt(3), literal of LocalOperations.cpp:47 */
static const double K23 = 1.15470053837925;  /* 2/sqrt(3) */
static const double K6 = 0.408248290463863;  /* 1/sqrt(6) */
static const double K36 = 1.22474487139159;  /* 3/sqrt(6) */
static const double KP = -0.333333333333333; /* exponent literal of LocalOperations.cpp:80 */

/* Q: LocalOperations.cpp:66-77, sign already folded in (the reference returns -(...)*pow) */
static double quad_form(const double *T) {
    double q = 0.0;
    for (int a = 0; a < 3; ++a) {
        double x0 = T[a], x1 = T[3 + a], x2 = T[6 + a], x3 = T[9 + a];
        q += x0 * (-1.5 * x0 + 0.5 * x1 + 0.5 * x2 + 0.5 * x3);
        q += x1 * (0.5 * x0 - 1.5 * x1 + 0.5 * x2 + 0.5 * x3);
        q += x2 * (0.5 * x0 + 0.5 * x1 - 1.5 * x2 + 0.5 * x3);
        q += x3 * (0.5 * x0 + 0.5 * x1 + 0.5 * x2 - 1.5 * x3);
    }
    return -q;
}

/* det: LocalOperations.cpp:47-62 and :78-80 */
static double ref_det(const double *T) {
    double A[3], B[3], C[3];
    for (int a = 0; a < 3; ++a) {
        double x0 = T[a], x1 = T[3 + a], x2 = T[6 + a], x3 = T[9 + a];
        A[a] = K3 * x0 - K23 * x1 + K3 * x3;
        B[a] = K6 * x0 + K6 * x1 - K36 * x2 + K6 * x3;
        C[a] = x0 - x3;
    }
    return C[2] * (B[1] * A[0] - A[1] * B[0]) - C[1] * (-B[0] * A[2] + B[2] * A[0]) + C[0] * (-B[1] * A[2] + A[1] * B[2]);
}

double ora_amips_energy(const double *T) {
    double d = ref_det(T);
    return quad_form(T) * pow(pow(d, 2), KP);
}

/* ---- second-order jets in (T[0], T[1], T[2]) ---- */
typedef struct {
    double v, g[3], h[6]; /* h: xx xy xz yy yz zz */
} jet;
static const int HI[3][3] = {{0, 1, 2}, {1, 3, 4}, {2, 4, 5}};

static jet jc(double c) {
    jet r = {c, {0, 0, 0}, {0, 0, 0, 0, 0, 0}};
    return r;
}
static jet jvar(double c, int k) {
    jet r = jc(c);
    r.g[k] = 1.0;
    return r;
}
static jet jadd(jet a, jet b) {
    jet r;
    r.v = a.v + b.v;
    for (int i = 0; i < 3; ++i) r.g[i] = a.g[i] + b.g[i];
    for (int i = 0; i < 6; ++i) r.h[i] = a.h[i] + b.h[i];
    return r;
}
static jet jscale(jet a, double s) {
    jet r;
    r.v = a.v * s;
    for (int i = 0; i < 3; ++i) r.g[i] = a.g[i] * s;
    for (int i = 0; i < 6; ++i) r.h[i] = a.h[i] * s;
    return r;
}
static jet jsub(jet a, jet b) { return jadd(a, jscale(b, -1.0)); }
static jet jmul(jet a, jet b) {
    jet r;
    r.v = a.v * b.v;
    for (int i = 0; i < 3; ++i) r.g[i] = a.g[i] * b.v + a.v * b.g[i];
    for (int i = 0; i < 3; ++i)
        for (int j = i; j < 3; ++j)
            r.h[HI[i][j]] = a.h[HI[i][j]] * b.v + a.g[i] * b.g[j] + a.g[j] * b.g[i] + a.v * b.h[HI[i][j]];
    return r;
}
static jet jpow(jet a, double p) {
    double f = pow(a.v, p);
    double f1 = p * pow(a.v, p - 1.0);
    double f2 = p * (p - 1.0) * pow(a.v, p - 2.0);
    jet r;
    r.v = f;
    for (int i = 0; i < 3; ++i) r.g[i] = f1 * a.g[i];
    for (int i = 0; i < 3; ++i)
        for (int j = i; j < 3; ++j) r.h[HI[i][j]] = f2 * a.g[i] * a.g[j] + f1 * a.h[HI[i][j]];
    return r;
}

static jet energy_jet(const double *T) {
    jet X[4][3];
    for (int i = 0; i < 4; ++i)
        for (int a = 0; a < 3; ++a) X[i][a] = (i == 0) ? jvar(T[a], a) : jc(T[3 * i + a]);
    jet q = jc(0.0);
    jet A[3], B[3], C[3];
    for (int a = 0; a < 3; ++a) {
        jet x0 = X[0][a], x1 = X[1][a], x2 = X[2][a], x3 = X[3][a];
        jet s;
        s = jadd(jadd(jscale(x0, -1.5), jscale(x1, 0.5)), jadd(jscale(x2, 0.5), jscale(x3, 0.5)));
        q = jadd(q, jmul(x0, s));
        s = jadd(jadd(jscale(x0, 0.5), jscale(x1, -1.5)), jadd(jscale(x2, 0.5), jscale(x3, 0.5)));
        q = jadd(q, jmul(x1, s));
        s = jadd(jadd(jscale(x0, 0.5), jscale(x1, 0.5)), jadd(jscale(x2, -1.5), jscale(x3, 0.5)));
        q = jadd(q, jmul(x2, s));
        s = jadd(jadd(jscale(x0, 0.5), jscale(x1, 0.5)), jadd(jscale(x2, 0.5), jscale(x3, -1.5)));
        q = jadd(q, jmul(x3, s));
        A[a] = jadd(jsub(jscale(x0, K3), jscale(x1, K23)), jscale(x3, K3));
        B[a] = jadd(jadd(jscale(x0, K6), jscale(x1, K6)), jsub(jscale(x3, K6), jscale(x2, K36)));
        C[a] = jsub(x0, x3);
    }
    jet m0 = jsub(jmul(B[1], A[0]), jmul(A[1], B[0]));
    jet m1 = jsub(jmul(B[2], A[0]), jmul(B[0], A[2]));
    jet m2 = jsub(jmul(A[1], B[2]), jmul(B[1], A[2]));
    jet det = jadd(jsub(jmul(C[2], m0), jmul(C[1], m1)), jmul(C[0], m2));
    jet e = jmul(jscale(q, -1.0), jpow(jmul(det, det), KP));
    return e;
}

void ora_amips_jacobian(const double *T, double *J3) {
    jet e = energy_jet(T);
    for (int i = 0; i < 3; ++i) J3[i] = e.g[i];
}

void ora_amips_hessian(const double *T, double *H9) {
    jet e = energy_jet(T);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) H9[3 * i + j] = e.h[HI[i][j]];
}

/* ---- batched forms ---- */

/* energy_ispc(V1_x..V4_z, E, count): src/ispc/energy.ispc:7-21 */
void ora_amips_energy_soa(const double *const Ts[12], double *E, uint64_t n, int threads) {
    (void)threads;
#pragma omp parallel for schedule(static) num_threads(threads > 0 ? threads : 1)
    for (int64_t i = 0; i < (int64_t)n; ++i) {
        double T[12];
        for (int k = 0; k < 12; ++k) T[k] = Ts[k][i];
        E[i] = ora_amips_energy(T);
    }
}

void ora_amips_ejh_soa(const double *const Ts[12], double *E, double *J3, double *H9, uint64_t n, int threads) {
    (void)threads;
#pragma omp parallel for schedule(static) num_threads(threads > 0 ? threads : 1)
    for (int64_t i = 0; i < (int64_t)n; ++i) {
        double T[12];
        for (int k = 0; k < 12; ++k) T[k] = Ts[k][i];
        jet e = energy_jet(T);
        if (E) E[i] = ora_amips_energy(T);
        if (J3)
            for (int a = 0; a < 3; ++a) J3[3 * i + a] = e.g[a];
        if (H9)
            for (int a = 0; a < 3; ++a)
                for (int b = 0; b < 3; ++b) H9[9 * i + 3 * a + b] = e.h[HI[a][b]];
    }
}

/* calTetQuality_AMIPS: LocalOperations.cpp:862-884 */
void ora_amips_quality(const double *V, const int32_t *tets, uint64_t nT, double *slim, int threads) {
    (void)threads;
#pragma omp parallel for schedule(static) num_threads(threads > 0 ? threads : 1)
    for (int64_t i = 0; i < (int64_t)nT; ++i) {
        const double *p[4];
        double T[12];
        for (int j = 0; j < 4; ++j) {
            p[j] = V + 3 * (size_t)tets[4 * i + j];
            for (int k = 0; k < 3; ++k) T[3 * j + k] = p[j][k];
        }
        double e;
        if (ora_cgal_orientation(p[0], p[1], p[2], p[3]) != 1) {
            e = ORA_MAX_ENERGY; /* :868-869 */
        } else {
            e = ora_amips_energy(T); /* :877 */
            if (isinf(e) || isnan(e)) e = ORA_MAX_ENERGY;
        }
        if (isinf(e) || isnan(e) || e <= 0) e = ORA_MAX_ENERGY; /* :882-883 */
        slim[i] = e;
    }
}

/* NewtonsUpdate: VertexSmoother.cpp:627-702 (non-ISPC build: energy summed in rotated order, :652-655) */
void ora_amips_ring_ejh(const double *V, const int32_t *tets, const int32_t *t_ids, const uint64_t *off,
                        const int32_t *center, uint64_t nG, double *E, double *J3, double *H9, uint8_t *ok,
                        int threads) {
    (void)threads;
#pragma omp parallel for schedule(dynamic, 256) num_threads(threads > 0 ? threads : 1)
    for (int64_t g = 0; g < (int64_t)nG; ++g) {
        double e = 0, J[3] = {0, 0, 0}, H[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        for (uint64_t k = off[g]; k < off[g + 1]; ++k) {
            const int32_t *tet = tets + 4 * (size_t)(t_ids ? t_ids[k] : (int64_t)k);
            int start = 0;
            for (int j = 0; j < 4; ++j)
                if (tet[j] == center[g]) { start = j; break; } /* :640-646 */
            double T[12];
            for (int j = 0; j < 4; ++j)
                for (int c = 0; c < 3; ++c) T[3 * j + c] = V[3 * (size_t)tet[(start + j) % 4] + c]; /* :647-651 */
            e += ora_amips_energy(T);
            jet je = energy_jet(T);
            for (int a = 0; a < 3; ++a) {
                J[a] += je.g[a];
                for (int b = 0; b < 3; ++b) H[3 * a + b] += je.h[HI[a][b]];
            }
        }
        int good = 1;
        if (isinf(e)) e = ORA_MAX_ENERGY; /* :680-683 */
        if (isnan(e)) good = 0;           /* :684-687 */
        if (e <= 0) good = 0;             /* :688-691 */
        for (int a = 0; a < 3; ++a)
            if (!isfinite(J[a])) good = 0; /* :692-695 */
        for (int a = 0; a < 9; ++a)
            if (!isfinite(H[a])) good = 0; /* :696-699 */
        E[g] = e;
        for (int a = 0; a < 3; ++a) J3[3 * g + a] = J[a];
        for (int a = 0; a < 9; ++a) H9[9 * g + a] = H[a];
        if (ok) ok[g] = (uint8_t)good;
    }
}

/* getNewEnergy: VertexSmoother.cpp:606-622 */
void ora_amips_ring_energy(const double *V, const int32_t *tets, const int32_t *t_ids, const uint64_t *off, uint64_t nG,
                           double *E, int threads) {
    (void)threads;
#pragma omp parallel for schedule(dynamic, 256) num_threads(threads > 0 ? threads : 1)
    for (int64_t g = 0; g < (int64_t)nG; ++g) {
        double s = 0;
        for (uint64_t k = off[g]; k < off[g + 1]; ++k) {
            const int32_t *tet = tets + 4 * (size_t)(t_ids ? t_ids[k] : (int64_t)k);
            double T[12];
            for (int j = 0; j < 4; ++j)
                for (int c = 0; c < 3; ++c) T[3 * j + c] = V[3 * (size_t)tet[j] + c];
            s += ora_amips_energy(T);
        }
        if (isinf(s) || isnan(s) || s <= 0 || s > ORA_MAX_ENERGY) s = ORA_MAX_ENERGY; /* :619-622 */
        E[g] = s;
    }
}

/* calTetQuality_AD: LocalOperations.cpp:783-860 (min / max dihedral angle of a tet).
 * Plane_3f(p,q,r) and Plane_3f::projection are CGAL constructions on Cartesian<double> (CGALTypes.h; CGAL is a system
 * package, not vendored -> restated from its published kernel_ftC3.h: plane_from_pointsC3, projection_planeC3,
 * squared_distanceC3). PARITY UNPINNED for those two constructions; everything else follows the reference text.
 * Built with -ffp-contract=off: plain IEEE double like the reference's x86-64 Release build. */
static int tet_dihedral(const double *x, double *amin, double *amax) {
    double nv[4][3], len[4];
    for (int i = 0; i < 4; ++i) {
        const double *P = x + 3 * ((i + 1) % 4), *Q = x + 3 * ((i + 2) % 4), *R = x + 3 * ((i + 3) % 4), *A = x + 3 * i;
        /* plane_from_pointsC3 */
        const double rpx = P[0] - R[0], rpy = P[1] - R[1], rpz = P[2] - R[2];
        const double rqx = Q[0] - R[0], rqy = Q[1] - R[1], rqz = Q[2] - R[2];
        const double pa = rpy * rqz - rqy * rpz, pb = rpz * rqx - rqz * rpx, pc = rpx * rqy - rqx * rpy;
        const double pd = -pa * R[0] - pb * R[1] - pc * R[2];
        if (pa == 0 && pb == 0 && pc == 0) return 0; /* pln.is_degenerate() :790 */
        /* projection_planeC3 */
        const double num = pa * A[0] + pb * A[1] + pc * A[2] + pd;
        const double den = pa * pa + pb * pb + pc * pc;
        const double lambda = num / den;
        const double t[3] = {A[0] - lambda * pa, A[1] - lambda * pb, A[2] - lambda * pc};
        if (t[0] == A[0] && t[1] == A[1] && t[2] == A[2]) return 0; /* :796 */
        for (int k = 0; k < 3; ++k) nv[i][k] = A[k] - t[k];          /* :801 */
        const double h = nv[i][0] * nv[i][0] + nv[i][1] * nv[i][1] + nv[i][2] * nv[i][2]; /* :802 */
        double m = fabs(nv[i][0]);
        if (fabs(nv[i][1]) > m) m = fabs(nv[i][1]);
        if (fabs(nv[i][2]) > m) m = fabs(nv[i][2]);
        if (m == 0 || h == 0) return 0; /* :820 */
        if (m < 1e-5) {                 /* :825-828 */
            for (int k = 0; k < 3; ++k) nv[i][k] = nv[i][k] / m;
            len[i] = sqrt(h / (m * m));
        } else {
            len[i] = sqrt(h);
        }
    }
    static const int ea[6] = {0, 1, 0, 2, 0, 3}, eb[6] = {1, 2, 2, 3, 3, 1}; /* opp_edges :834-838 */
    double ang[6];
    for (int k = 0; k < 6; ++k) {
        const double *a = nv[ea[k]], *b = nv[eb[k]];
        const double c = ((-a[0]) * b[0] + (-a[1]) * b[1] + (-a[2]) * b[2]) / (len[ea[k]] * len[eb[k]]); /* :843-844 */
        ang[k] = c > 1 ? 0.0 : (c < -1 ? M_PI : acos(c));
    }
    double lo = ang[0], hi = ang[0]; /* std::minmax_element :854: first smallest, last largest */
    for (int k = 1; k < 6; ++k) {
        if (ang[k] < lo) lo = ang[k];
        if (!(ang[k] < hi)) hi = ang[k];
    }
    *amin = lo;
    *amax = hi;
    return 1;
}

void ora_tet_dihedral(const double *V, const int32_t *tets, uint64_t nT, double *dmin, double *dmax, int threads) {
    (void)threads;
#pragma omp parallel for schedule(static) num_threads(threads > 0 ? threads : 1)
    for (int64_t i = 0; i < (int64_t)nT; ++i) {
        double T[12], lo = 0.0, hi = M_PI; /* degenerate answers :791-792 */
        if (tets[4 * i] >= 0) {
            for (int j = 0; j < 4; ++j)
                for (int c = 0; c < 3; ++c) T[3 * j + c] = V[3 * (size_t)tets[4 * i + j] + c];
            double a, b;
            if (tet_dihedral(T, &a, &b)) { lo = a; hi = b; }
        }
        dmin[i] = lo;
        dmax[i] = hi;
    }
}
