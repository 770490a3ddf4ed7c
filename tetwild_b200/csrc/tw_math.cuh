// tw_math.cuh -- numeric core shared by every kernel of libtetwild_gpu.
//
// "Strict" arithmetic: the envelope decisions of the reference (src/tetwild/LocalOperations.cpp:1083-1093,
// src/tetwild/geogram/mesh_AABB.cpp:482-548) compare per-triangle squared distances, computed in plain IEEE double
// WITHOUT fused multiply-add (a default x86-64 build of the reference has no FMA), against eps_2. To take the
// same decisions the device code evaluates those expressions with the explicitly rounded intrinsics
// (__dmul_rn/__dadd_rn/...: never contracted by nvcc) in the same operation order. The functions are
// __host__ __device__ so the CPU-only unit tests (tests/host_harness.cu) can execute the very same source.
#pragma once
#include <cstdint>
#include <cmath>
#include <cfloat>

#if defined(__CUDACC__)
#define TW_HD __host__ __device__ __forceinline__
#define TW_HD_NOINLINE inline __host__ __device__ __noinline__
#else
#define TW_HD inline
#define TW_HD_NOINLINE inline
#endif

namespace tw {

#if defined(__CUDA_ARCH__)
TW_HD double dmul(double a, double b) { return __dmul_rn(a, b); }
TW_HD double dadd(double a, double b) { return __dadd_rn(a, b); }
TW_HD double dsub(double a, double b) { return __dsub_rn(a, b); }
TW_HD double ddiv(double a, double b) { return __ddiv_rn(a, b); }
TW_HD double dsqrt(double a) { return __dsqrt_rn(a); }
TW_HD double dfma(double a, double b, double c) { return __fma_rn(a, b, c); }
#else
// host build: compiled with -ffp-contract=off (see tests/host_harness build line)
TW_HD double dmul(double a, double b) { volatile double r = a * b; return r; }
TW_HD double dadd(double a, double b) { volatile double r = a + b; return r; }
TW_HD double dsub(double a, double b) { volatile double r = a - b; return r; }
TW_HD double ddiv(double a, double b) { return a / b; }
TW_HD double dsqrt(double a) { return std::sqrt(a); }
TW_HD double dfma(double a, double b, double c) { return std::fma(a, b, c); }
#endif

struct V3 {
    double x, y, z;
};
TW_HD V3 mk(double x, double y, double z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
// geogram vec3 operator semantics (see oracle/envelope.c header): component-wise, left to right
TW_HD V3 vsub(V3 a, V3 b) { return mk(dsub(a.x, b.x), dsub(a.y, b.y), dsub(a.z, b.z)); }
TW_HD V3 vadd(V3 a, V3 b) { return mk(dadd(a.x, b.x), dadd(a.y, b.y), dadd(a.z, b.z)); }
TW_HD V3 vscale(V3 a, double s) { return mk(dmul(a.x, s), dmul(a.y, s), dmul(a.z, s)); }
TW_HD double vdot(V3 a, V3 b) { return dadd(dadd(dmul(a.x, b.x), dmul(a.y, b.y)), dmul(a.z, b.z)); }
TW_HD double vlen2(V3 a) { return vdot(a, a); }
TW_HD V3 vcross(V3 a, V3 b) {
    return mk(dsub(dmul(a.y, b.z), dmul(a.z, b.y)), dsub(dmul(a.z, b.x), dmul(a.x, b.z)),
              dsub(dmul(a.x, b.y), dmul(a.y, b.x)));
}
TW_HD double vdist2(V3 a, V3 b) { return vlen2(vsub(b, a)); }
TW_HD double vdist(V3 a, V3 b) { return dsqrt(vdist2(a, b)); }
TW_HD V3 vnormalize(V3 a) {
    double s = dsqrt(vlen2(a));
    if (s > 1e-30) s = ddiv(1.0, s);
    return mk(dmul(s, a.x), dmul(s, a.y), dmul(s, a.z));
}

// ------------------------------------------------------------------------------------------------------------
// Conservative single-precision box test (used by every traversal of surface.cuh / envelope.cu).
// The query point is bracketed by two floats (p_lo <= p <= p_hi), every operation is rounded DOWN, so the result is a
// rigorous lower bound of the true squared distance from p to the (outward-rounded) float box -- and the box contains
// every facet below it. A subtree is therefore never skipped when it holds a facet with d2 <= eps2 (thr = eps2 rounded
// UP to float); the bound merely admits a few more boxes than the exact double test would (relative slack ~1e-7).
// FP32 runs at twice the FP64 rate on B200 and leaves the FP64 pipe to the leaf arithmetic.
// Host build: the same operations under fesetround (tests/test_host_core.py checks the bound against exact rationals).
// ------------------------------------------------------------------------------------------------------------
struct PointF {
    float lx, ly, lz, hx, hy, hz;
};
#if defined(__CUDA_ARCH__)
TW_HD float f_down(double x) { return __double2float_rd(x); }
TW_HD float f_up(double x) { return __double2float_ru(x); }
TW_HD float fsub_down(float a, float b) { return __fsub_rd(a, b); }
TW_HD float fmul_down(float a, float b) { return __fmul_rd(a, b); }
TW_HD float ffma_down(float a, float b, float c) { return __fmaf_rd(a, b, c); }
#else
}  // namespace tw
#include <cfenv>
namespace tw {
struct RoundGuard {
    int old;
    explicit RoundGuard(int mode) : old(std::fegetround()) { std::fesetround(mode); }
    ~RoundGuard() { std::fesetround(old); }
};
// (operands go through volatiles: the compiler does not know that the conversions / operations depend on the rounding mode)
inline float f_down(double x) { RoundGuard g(FE_DOWNWARD); volatile double vx = x; volatile float r = (float)vx; return r; }
inline float f_up(double x) { RoundGuard g(FE_UPWARD); volatile double vx = x; volatile float r = (float)vx; return r; }
inline float fsub_down(float a, float b) { RoundGuard g(FE_DOWNWARD); volatile float va = a, vb = b; volatile float r = va - vb; return r; }
inline float fmul_down(float a, float b) { RoundGuard g(FE_DOWNWARD); volatile float va = a, vb = b; volatile float r = va * vb; return r; }
inline float ffma_down(float a, float b, float c) { RoundGuard g(FE_DOWNWARD); volatile float va = a, vb = b, vc = c; volatile float r = std::fmaf(va, vb, vc); return r; }
#endif
TW_HD PointF bracket(V3 p) {
    PointF q;
    q.lx = f_down(p.x); q.ly = f_down(p.y); q.lz = f_down(p.z);
    q.hx = f_up(p.x); q.hy = f_up(p.y); q.hz = f_up(p.z);
    return q;
}
TW_HD float box_d2_lb(const PointF& q, float lx, float ly, float lz, float hx, float hy, float hz) {
    const float dx = fmaxf(fmaxf(fsub_down(lx, q.hx), fsub_down(q.lx, hx)), 0.0f);
    const float dy = fmaxf(fmaxf(fsub_down(ly, q.hy), fsub_down(q.ly, hy)), 0.0f);
    const float dz = fmaxf(fmaxf(fsub_down(lz, q.hz), fsub_down(q.lz, hz)), 0.0f);
    return ffma_down(dz, dz, ffma_down(dy, dy, fmul_down(dx, dx)));
}

// ------------------------------------------------------------------------------------------------------------
// Oriented bound of one facet (32 B): centre c and radius R of a ball that holds the facet, a unit vector n (the facet's normal,
// rounded to float; zero for degenerate facets) and the half-thickness w of the facet along n. For ANY unit vector u,
// |p - x|^2 = (u.(p-x))^2 + |(p-x) - u (u.(p-x))|^2, so for every x of the facet
//   |p - x|^2 >= max(0, |n.(p-c)| - w)^2 + max(0, sqrt(|p-c|^2 - (n.(p-c))^2) - R)^2.
// w and R are computed (rounded up) from the STORED float n and c against the exact vertices, so the bound is rigorous for
// the stored values. Where the box bound of a leaf only knows the facet's axis-aligned extent, this bound knows its plane:
// the exact nearest search runs the full point-triangle routine on ~10x fewer facets (envelope.cu, nearest_packet_kernel).
// Host build: the same source (tests/test_host_core.py checks the bound against the exact routine and exact rationals).
// ------------------------------------------------------------------------------------------------------------
struct __attribute__((aligned(16))) TriBound {
    float cx, cy, cz, R;
    float nx, ny, nz, w;
};
static_assert(sizeof(TriBound) == 32, "TriBound is two 128-bit words");

TW_HD void make_bound(const double* tv /*9 doubles: V0 V1 V2*/, bool degenerate, TriBound& B) {
    const double e0[3] = {tv[3] - tv[0], tv[4] - tv[1], tv[5] - tv[2]}, e1[3] = {tv[6] - tv[0], tv[7] - tv[1], tv[8] - tv[2]};
    const double e2[3] = {tv[6] - tv[3], tv[7] - tv[4], tv[8] - tv[5]};
    const double l01 = e0[0] * e0[0] + e0[1] * e0[1] + e0[2] * e0[2], l02 = e1[0] * e1[0] + e1[1] * e1[1] + e1[2] * e1[2];
    const double l12 = e2[0] * e2[0] + e2[1] * e2[1] + e2[2] * e2[2];
    const double cr[3] = {e0[1] * e1[2] - e0[2] * e1[1], e0[2] * e1[0] - e0[0] * e1[2], e0[0] * e1[1] - e0[1] * e1[0]};
    const double cr2 = cr[0] * cr[0] + cr[1] * cr[1] + cr[2] * cr[2];
    // centre of the smallest enclosing circle: midpoint of the longest edge for an obtuse facet, else the circumcentre
    double c[3];
    if (l12 >= l01 + l02) { for (int k = 0; k < 3; ++k) c[k] = 0.5 * (tv[3 + k] + tv[6 + k]); }
    else if (l02 >= l01 + l12) { for (int k = 0; k < 3; ++k) c[k] = 0.5 * (tv[k] + tv[6 + k]); }
    else if (l01 >= l02 + l12) { for (int k = 0; k < 3; ++k) c[k] = 0.5 * (tv[k] + tv[3 + k]); }
    else {  // V0 + (|e1|^2 (e0 x e1) x e0 + |e0|^2 e1 x (e0 x e1)) / (2 |e0 x e1|^2)
        const double a[3] = {cr[1] * e0[2] - cr[2] * e0[1], cr[2] * e0[0] - cr[0] * e0[2], cr[0] * e0[1] - cr[1] * e0[0]};
        const double b[3] = {e1[1] * cr[2] - e1[2] * cr[1], e1[2] * cr[0] - e1[0] * cr[2], e1[0] * cr[1] - e1[1] * cr[0]};
        for (int k = 0; k < 3; ++k) c[k] = tv[k] + (l02 * a[k] + l01 * b[k]) / (2.0 * cr2);
    }
    B.cx = (float)c[0]; B.cy = (float)c[1]; B.cz = (float)c[2];
    if (!(fabs((double)B.cx) <= 3e38) || !(fabs((double)B.cy) <= 3e38) || !(fabs((double)B.cz) <= 3e38)) { B.cx = (float)tv[0]; B.cy = (float)tv[1]; B.cz = (float)tv[2]; }
    const double inv = (cr2 > 0.0 && !degenerate) ? 1.0 / sqrt(cr2) : 0.0;
    B.nx = (float)(cr[0] * inv); B.ny = (float)(cr[1] * inv); B.nz = (float)(cr[2] * inv);
    if (!(fabs((double)B.nx) <= 2.0) || !(fabs((double)B.ny) <= 2.0) || !(fabs((double)B.nz) <= 2.0)) { B.nx = B.ny = B.nz = 0.f; }
    // R and w from the ROUNDED centre / normal against the exact vertices, rounded up
    double R2 = 0.0, wmax = 0.0;
    for (int k = 0; k < 3; ++k) {
        const double dx = tv[3 * k] - (double)B.cx, dy = tv[3 * k + 1] - (double)B.cy, dz = tv[3 * k + 2] - (double)B.cz;
        R2 = fmax(R2, dx * dx + dy * dy + dz * dz);
        wmax = fmax(wmax, fabs(dx * (double)B.nx + dy * (double)B.ny + dz * (double)B.nz));
    }
    B.R = f_up(sqrt(R2) * (1.0 + 1e-6));
    B.w = f_up(wmax * (1.0 + 1e-6) + 1e-300);
}

// rigorous lower bound of the squared distance from p to the facet behind `b`; evaluated in double: the float fields
// convert exactly, |n| is within 1.2e-7 of 1 (covered by the 5e-7 deflations), and the rounding of the double operations
// (a few 1e-16 |p-c|^2) by the last term
TW_HD double bound_lb2(const TriBound& b, V3 p) {
    const double dx = p.x - (double)b.cx, dy = p.y - (double)b.cy, dz = p.z - (double)b.cz;
    const double pi = dx * (double)b.nx + dy * (double)b.ny + dz * (double)b.nz;
    const double r2 = dx * dx + dy * dy + dz * dz;
    const double a = fmax(fabs(pi) - (double)b.w, 0.0);
    // lateral distance: a LOWER bound of sqrt(q) is all that is needed, so the root is taken in single precision (ncu r02: the
    // double root was 6-7 % of the instructions of the point and nearest kernels). float(q) <= q (1 + 6e-8), the correctly
    // rounded sqrtf adds 6e-8, the product another 6e-8; 0.9999996f = 1 - 4.2e-7 leaves sf < sqrt(q). q beyond the float range is
    // clamped (still a lower bound), q below it rounds to 0 (l = 0).
    const double q = fmax(r2 - pi * pi * (1.0 + 5e-7), 0.0);
    const float sf = sqrtf(fminf((float)q, 3.0e38f)) * 0.9999996f;
    const double l = fmax((double)sf - (double)b.R, 0.0);
    return a * a * (1.0 - 5e-7) + l * l - 1e-14 * r2;
}

// ------------------------------------------------------------------------------------------------------------
// Point-triangle squared distance: the 7-region minimisation of Q(s,t) = |V0 + s e0 + t e1 - p|^2 (D. Eberly).
// This is the leaf routine of the reference tree (mesh_AABB.cpp:153-171 -> geogram point_triangle_squared_distance).
// The query-independent terms are precomputed once per facet into a 128-byte record (TriRec) with the same
// operations the reference performs per call, so results stay bit-identical.
// ------------------------------------------------------------------------------------------------------------
struct __attribute__((aligned(16))) TriRec {
    double v0[3];   // V0
    double e0[3];   // V1 - V0
    double e1[3];   // V2 - V0
    double a00, a01, a11;
    double det;     // |a00*a11 - a01*a01|
    double inv;     // 1/det
    double den;     // a00 - 2 a01 + a11
    uint32_t facet; // caller's facet id
    uint32_t flags; // bit0: degenerate (det < 1e-30)
};
static_assert(sizeof(TriRec) == 128, "TriRec must be one 128-byte line");

TW_HD void make_trirec(const double* V0, const double* V1, const double* V2, uint32_t facet, TriRec& r) {
    V3 a = mk(V0[0], V0[1], V0[2]), b = mk(V1[0], V1[1], V1[2]), c = mk(V2[0], V2[1], V2[2]);
    V3 e0 = vsub(b, a), e1 = vsub(c, a);
    r.v0[0] = a.x; r.v0[1] = a.y; r.v0[2] = a.z;
    r.e0[0] = e0.x; r.e0[1] = e0.y; r.e0[2] = e0.z;
    r.e1[0] = e1.x; r.e1[1] = e1.y; r.e1[2] = e1.z;
    r.a00 = vlen2(e0);
    r.a01 = vdot(e0, e1);
    r.a11 = vlen2(e1);
    r.det = fabs(dsub(dmul(r.a00, r.a11), dmul(r.a01, r.a01)));
    r.inv = ddiv(1.0, r.det);
    r.den = dadd(dsub(r.a00, dmul(2.0, r.a01)), r.a11);
    r.facet = facet;
    r.flags = (r.det < 1e-30) ? 1u : 0u;
}

TW_HD double seg_sqdist(V3 p, V3 v0, V3 v1, V3& nearest) {
    double l2 = vdist2(v0, v1);
    double t = vdot(vsub(p, v0), vsub(v1, v0));
    if (t <= 0.0 || l2 == 0.0) { nearest = v0; return vdist2(p, v0); }
    if (t > l2) { nearest = v1; return vdist2(p, v1); }
    double l1 = ddiv(t, l2);
    double l0 = dsub(1.0, l1);
    nearest = mk(dadd(dmul(l0, v0.x), dmul(l1, v1.x)), dadd(dmul(l0, v0.y), dmul(l1, v1.y)),
                 dadd(dmul(l0, v0.z), dmul(l1, v1.z)));
    return vdist2(p, nearest);
}

// degenerate facets need the caller's exact vertices (V0 + e0 is not V1 in floating point)
TW_HD double tri_sqdist_degenerate(V3 p, const double* tv /*9 doubles*/, V3& nearest) {
    V3 a = mk(tv[0], tv[1], tv[2]), b = mk(tv[3], tv[4], tv[5]), c = mk(tv[6], tv[7], tv[8]);
    V3 cur;
    double best = seg_sqdist(p, a, b, nearest);
    double d = seg_sqdist(p, a, c, cur);
    if (d < best) { best = d; nearest = cur; }
    d = seg_sqdist(p, b, c, cur);
    if (d < best) { best = d; nearest = cur; }
    return best;
}

// quadratic Q(s,t) + c, evaluated exactly like  s*(a00*s + a01*t + 2*b0) + t*(a01*s + a11*t + 2*b1) + c
TW_HD double quad_st(double s, double t, double a00, double a01, double a11, double b0, double b1, double c) {
    double l = dmul(s, dadd(dadd(dmul(a00, s), dmul(a01, t)), dmul(2.0, b0)));
    double r = dmul(t, dadd(dadd(dmul(a01, s), dmul(a11, t)), dmul(2.0, b1)));
    return dadd(dadd(l, r), c);
}
// edge minimisers shared by several regions
TW_HD double edge_s(double a00, double b0, double c, double& s) {  // t = 0, s in [0,1]
    if (b0 >= 0.0) { s = 0.0; return c; }
    if (-b0 >= a00) { s = 1.0; return dadd(dadd(a00, dmul(2.0, b0)), c); }
    s = ddiv(-b0, a00);
    return dadd(dmul(b0, s), c);
}
TW_HD double edge_t(double a11, double b1, double c, double& t) {  // s = 0, t in [0,1]
    if (b1 >= 0.0) { t = 0.0; return c; }
    if (-b1 >= a11) { t = 1.0; return dadd(dadd(a11, dmul(2.0, b1)), c); }
    t = ddiv(-b1, a11);
    return dadd(dmul(b1, t), c);
}

// Non-degenerate facet. Returns d^2; s,t = barycentric parameters of the nearest point (V0 + s e0 + t e1).
TW_HD double tri_sqdist_rec(V3 p, const TriRec& r, double& s_out, double& t_out) {
    V3 diff = mk(dsub(r.v0[0], p.x), dsub(r.v0[1], p.y), dsub(r.v0[2], p.z));
    V3 e0 = mk(r.e0[0], r.e0[1], r.e0[2]), e1 = mk(r.e1[0], r.e1[1], r.e1[2]);
    const double a00 = r.a00, a01 = r.a01, a11 = r.a11, det = r.det;
    double b0 = vdot(diff, e0), b1 = vdot(diff, e1), c = vlen2(diff);
    double s = dsub(dmul(a01, b1), dmul(a11, b0));
    double t = dsub(dmul(a01, b0), dmul(a00, b1));
    double d2;
    if (dadd(s, t) <= det) {
        if (s < 0.0) {
            if (t < 0.0) {  // region 4
                if (b0 < 0.0) {
                    t = 0.0;
                    if (-b0 >= a00) { s = 1.0; d2 = dadd(dadd(a00, dmul(2.0, b0)), c); }
                    else { s = ddiv(-b0, a00); d2 = dadd(dmul(b0, s), c); }
                } else {
                    s = 0.0;
                    d2 = edge_t(a11, b1, c, t);
                }
            } else {  // region 3
                s = 0.0;
                d2 = edge_t(a11, b1, c, t);
            }
        } else if (t < 0.0) {  // region 5
            t = 0.0;
            d2 = edge_s(a00, b0, c, s);
        } else {  // region 0
            s = dmul(s, r.inv);
            t = dmul(t, r.inv);
            d2 = quad_st(s, t, a00, a01, a11, b0, b1, c);
        }
    } else {
        if (s < 0.0) {  // region 2
            double tmp0 = dadd(a01, b0), tmp1 = dadd(a11, b1);
            if (tmp1 > tmp0) {
                double numer = dsub(tmp1, tmp0);
                if (numer >= r.den) { s = 1.0; t = 0.0; d2 = dadd(dadd(a00, dmul(2.0, b0)), c); }
                else { s = ddiv(numer, r.den); t = dsub(1.0, s); d2 = quad_st(s, t, a00, a01, a11, b0, b1, c); }
            } else {
                s = 0.0;
                if (tmp1 <= 0.0) { t = 1.0; d2 = dadd(dadd(a11, dmul(2.0, b1)), c); }
                else if (b1 >= 0.0) { t = 0.0; d2 = c; }
                else { t = ddiv(-b1, a11); d2 = dadd(dmul(b1, t), c); }
            }
        } else if (t < 0.0) {  // region 6
            double tmp0 = dadd(a01, b1), tmp1 = dadd(a00, b0);
            if (tmp1 > tmp0) {
                double numer = dsub(tmp1, tmp0);
                if (numer >= r.den) { t = 1.0; s = 0.0; d2 = dadd(dadd(a11, dmul(2.0, b1)), c); }
                else { t = ddiv(numer, r.den); s = dsub(1.0, t); d2 = quad_st(s, t, a00, a01, a11, b0, b1, c); }
            } else {
                t = 0.0;
                if (tmp1 <= 0.0) { s = 1.0; d2 = dadd(dadd(a00, dmul(2.0, b0)), c); }
                else if (b0 >= 0.0) { s = 0.0; d2 = c; }
                else { s = ddiv(-b0, a00); d2 = dadd(dmul(b0, s), c); }
            }
        } else {  // region 1
            double numer = dsub(dsub(dadd(a11, b1), a01), b0);
            if (numer <= 0.0) { s = 0.0; t = 1.0; d2 = dadd(dadd(a11, dmul(2.0, b1)), c); }
            else if (numer >= r.den) { s = 1.0; t = 0.0; d2 = dadd(dadd(a00, dmul(2.0, b0)), c); }
            else { s = ddiv(numer, r.den); t = dsub(1.0, s); d2 = quad_st(s, t, a00, a01, a11, b0, b1, c); }
        }
    }
    if (d2 < 0.0) d2 = 0.0;
    s_out = s;
    t_out = t;
    return d2;
}

TW_HD V3 tri_nearest_point(const TriRec& r, double s, double t) {  // V0 + s*e0 + t*e1
    return mk(dadd(dadd(r.v0[0], dmul(s, r.e0[0])), dmul(t, r.e1[0])), dadd(dadd(r.v0[1], dmul(s, r.e0[1])), dmul(t, r.e1[1])),
              dadd(dadd(r.v0[2], dmul(s, r.e0[2])), dmul(t, r.e1[2])));
}

// ------------------------------------------------------------------------------------------------------------
// Conformal AMIPS energy of a tetrahedron and its gradient / Hessian with respect to vertex 0, in closed form.
//
// Reference: comformalAMIPS{Energy,Jacobian,Hessian}_new (src/tetwild/LocalOperations.cpp:28-291) evaluate
//   E = Q * (det^2)^(-0.333333333333333) with Q = half the sum of the six squared edge lengths and det = -sqrt2 * d,
//   d = det[x1-x0, x2-x0, x3-x0], through ~750 machine-generated flops in absolute coordinates.
// Here: with e_i = x_i - x0, m = -(e1+e2+e3) (= grad Q), n = -(x2-x1)x(x3-x1) (= grad d, independent of x0):
//   q = (4 sum|e_i|^2 - |m|^2)/2,   d = -n.e1,   f = (2 d^2)^(-1/3)
//   E = q f
//   J = f (m - (2/3)(q/d) n)
//   H = f (3 I - (2/3)/d (m n^T + n m^T) + (10/9)(q/d^2) n n^T)
// (d is linear and q quadratic in x0, so these are exact). ~130 flops + one rcbrt + one division; working in edge
// vectors is also better conditioned than the reference's absolute-coordinate form (DESIGN.md "AMIPS tolerance").
// Plain (contractable) arithmetic: the bar for this part is 1e-9 relative, not bit equality.
// ------------------------------------------------------------------------------------------------------------
TW_HD double tw_rcbrt(double x) {
#if defined(__CUDA_ARCH__)
    return rcbrt(x);
#else
    return 1.0 / std::cbrt(x);
#endif
}

struct Amips {
    double E;
    double J[3];
    double H[6];  // xx xy xz yy yz zz (the closed form is symmetric)
};

template <bool WANT_JH>
TW_HD void amips_eval(const double* x /*12: v0 v1 v2 v3*/, Amips& o) {
    const double e1x = x[3] - x[0], e1y = x[4] - x[1], e1z = x[5] - x[2];
    const double e2x = x[6] - x[0], e2y = x[7] - x[1], e2z = x[8] - x[2];
    const double e3x = x[9] - x[0], e3y = x[10] - x[1], e3z = x[11] - x[2];
    const double mx = -(e1x + e2x + e3x), my = -(e1y + e2y + e3y), mz = -(e1z + e2z + e3z);
    const double s2 = e1x * e1x + e1y * e1y + e1z * e1z + e2x * e2x + e2y * e2y + e2z * e2z + e3x * e3x + e3y * e3y + e3z * e3z;
    const double q = 0.5 * (4.0 * s2 - (mx * mx + my * my + mz * mz));
    const double ax = e2x - e1x, ay = e2y - e1y, az = e2z - e1z;
    const double bx = e3x - e1x, by = e3y - e1y, bz = e3z - e1z;
    const double nx = -(ay * bz - az * by), ny = -(az * bx - ax * bz), nz = -(ax * by - ay * bx);
    const double d = -(nx * e1x + ny * e1y + nz * e1z);
    const double f = tw_rcbrt(2.0 * d * d);
    o.E = q * f;
    if (WANT_JH) {
        const double invd = 1.0 / d;
        const double k1 = (2.0 / 3.0) * invd;
        const double qd = q * k1;  // (2/3) q/d
        o.J[0] = f * (mx - qd * nx);
        o.J[1] = f * (my - qd * ny);
        o.J[2] = f * (mz - qd * nz);
        const double k2 = (5.0 / 3.0) * qd * invd;  // (10/9) q/d^2
        const double ux = k2 * nx - k1 * mx, uy = k2 * ny - k1 * my, uz = k2 * nz - k1 * mz;  // H = f(3I + n u^T - k1 m n^T)
        o.H[0] = f * (3.0 + nx * ux - k1 * mx * nx);
        o.H[1] = f * (nx * uy - k1 * mx * ny);
        o.H[2] = f * (nx * uz - k1 * mx * nz);
        o.H[3] = f * (3.0 + ny * uy - k1 * my * ny);
        o.H[4] = f * (ny * uz - k1 * my * nz);
        o.H[5] = f * (3.0 + nz * uz - k1 * mz * nz);
    }
}

// ------------------------------------------------------------------------------------------------------------
// Exact sign predicates on doubles (what the reference gets from CGAL's Epick kernel, CGALTypes.h:40-41):
// static forward-error filter, then exact evaluation with floating-point expansions (Shewchuk 1997).
// ------------------------------------------------------------------------------------------------------------
namespace exact {
constexpr double EPS_HALF = 1.1102230246251565e-16;  // 2^-53

TW_HD void two_sum(double a, double b, double& x, double& y) {
    x = dadd(a, b);
    double bv = dsub(x, a);
    double av = dsub(x, bv);
    y = dadd(dsub(a, av), dsub(b, bv));
}
TW_HD void fast_two_sum(double a, double b, double& x, double& y) {
    x = dadd(a, b);
    y = dsub(b, dsub(x, a));
}
TW_HD void two_prod(double a, double b, double& x, double& y) {
    x = dmul(a, b);
    y = dfma(a, b, -x);
}

// h = e + f (expansions by increasing magnitude, zeros eliminated); returns |h|
TW_HD_NOINLINE int exp_sum(const double* e, int ne, const double* f, int nf, double* h) {
    int i = 0, j = 0, n = 0;
    double Q = 0.0, q;
    int taken = 0;
    double g0 = 0.0;
    while (i < ne || j < nf) {
        double g;
        if (j >= nf || (i < ne && fabs(e[i]) < fabs(f[j]))) g = e[i++]; else g = f[j++];
        if (taken == 0) { g0 = g; taken = 1; continue; }
        if (taken == 1) { fast_two_sum(g, g0, Q, q); taken = 2; }
        else two_sum(Q, g, Q, q);
        if (q != 0.0) h[n++] = q;
    }
    if (taken == 0) return 0;
    if (taken == 1) { if (g0 != 0.0) { h[0] = g0; return 1; } return 0; }
    if (Q != 0.0) h[n++] = Q;
    return n;
}

// h = e * b
TW_HD_NOINLINE int exp_scale(const double* e, int ne, double b, double* h) {
    if (ne == 0 || b == 0.0) return 0;
    int n = 0;
    double Q, q, P, p, s;
    two_prod(e[0], b, Q, q);
    if (q != 0.0) h[n++] = q;
    for (int k = 1; k < ne; ++k) {
        two_prod(e[k], b, P, p);
        two_sum(Q, p, s, q);
        if (q != 0.0) h[n++] = q;
        fast_two_sum(P, s, Q, q);
        if (q != 0.0) h[n++] = q;
    }
    if (Q != 0.0) h[n++] = Q;
    return n;
}

// px*qy - qx*py  as an expansion (<= 4 terms)
TW_HD int cross2(double px, double py, double qx, double qy, double* h) {
    double a[2], b[2], x1, x0;
    int na = 0, nb = 0;
    two_prod(px, qy, x1, x0);
    if (x0 != 0.0) a[na++] = x0;
    if (x1 != 0.0) a[na++] = x1;
    two_prod(qx, py, x1, x0);
    if (x0 != 0.0) b[nb++] = -x0;
    if (x1 != 0.0) b[nb++] = -x1;
    return exp_sum(a, na, b, nb, h);
}

TW_HD int exp_sign(const double* e, int n) { return n == 0 ? 0 : ((e[n - 1] > 0.0) - (e[n - 1] < 0.0)); }

// exact sign of det[a-d; b-d; c-d] = det | a 1; b 1; c 1; d 1 |, cofactor expansion along z
TW_HD_NOINLINE int orient3d_exact(const double* a, const double* b, const double* c, const double* d) {
    double ab[4], bc[4], cd[4], da[4], ac[4], bd[4];
    int nab = cross2(a[0], a[1], b[0], b[1], ab), nbc = cross2(b[0], b[1], c[0], c[1], bc);
    int ncd = cross2(c[0], c[1], d[0], d[1], cd), nda = cross2(d[0], d[1], a[0], a[1], da);
    int nac = cross2(a[0], a[1], c[0], c[1], ac), nbd = cross2(b[0], b[1], d[0], d[1], bd);
    double t8[8], m[12], part[24], acc[96], tmp[96];
    int nt, nm, np, nacc = 0;
    // + az * (bc + cd - bd)
    for (int k = 0; k < nbd; ++k) bd[k] = -bd[k];
    nt = exp_sum(bc, nbc, cd, ncd, t8); nm = exp_sum(t8, nt, bd, nbd, m);
    nacc = exp_scale(m, nm, a[2], acc);
    // - bz * (ac + cd + da)
    nt = exp_sum(cd, ncd, da, nda, t8); nm = exp_sum(t8, nt, ac, nac, m);
    np = exp_scale(m, nm, -b[2], part);
    nacc = exp_sum(acc, nacc, part, np, tmp); for (int k = 0; k < nacc; ++k) acc[k] = tmp[k];
    // + cz * (ab + bd + da)     (bd currently negated -> restore)
    for (int k = 0; k < nbd; ++k) bd[k] = -bd[k];
    nt = exp_sum(da, nda, ab, nab, t8); nm = exp_sum(t8, nt, bd, nbd, m);
    np = exp_scale(m, nm, c[2], part);
    nacc = exp_sum(acc, nacc, part, np, tmp); for (int k = 0; k < nacc; ++k) acc[k] = tmp[k];
    // - dz * (ab + bc - ac)
    for (int k = 0; k < nac; ++k) ac[k] = -ac[k];
    nt = exp_sum(ab, nab, bc, nbc, t8); nm = exp_sum(t8, nt, ac, nac, m);
    np = exp_scale(m, nm, -d[2], part);
    nacc = exp_sum(acc, nacc, part, np, tmp);
    return exp_sign(tmp, nacc);
}

// sign of det[a-d; b-d; c-d]
TW_HD int orient3d(const double* a, const double* b, const double* c, const double* d) {
    double adx = dsub(a[0], d[0]), bdx = dsub(b[0], d[0]), cdx = dsub(c[0], d[0]);
    double ady = dsub(a[1], d[1]), bdy = dsub(b[1], d[1]), cdy = dsub(c[1], d[1]);
    double adz = dsub(a[2], d[2]), bdz = dsub(b[2], d[2]), cdz = dsub(c[2], d[2]);
    double bdxcdy = dmul(bdx, cdy), cdxbdy = dmul(cdx, bdy);
    double cdxady = dmul(cdx, ady), adxcdy = dmul(adx, cdy);
    double adxbdy = dmul(adx, bdy), bdxady = dmul(bdx, ady);
    double det = dadd(dadd(dmul(adz, dsub(bdxcdy, cdxbdy)), dmul(bdz, dsub(cdxady, adxcdy))), dmul(cdz, dsub(adxbdy, bdxady)));
    double perm = dadd(dadd(dmul(dadd(fabs(bdxcdy), fabs(cdxbdy)), fabs(adz)), dmul(dadd(fabs(cdxady), fabs(adxcdy)), fabs(bdz))),
                       dmul(dadd(fabs(adxbdy), fabs(bdxady)), fabs(cdz)));
    double bound = dmul((7.0 + 56.0 * EPS_HALF) * EPS_HALF, perm);
    if (det > bound) return 1;
    if (-det > bound) return -1;
    return orient3d_exact(a, b, c, d);
}

// CGAL::orientation(p,q,r,s) = sign det[q-p; r-p; s-p]   (LocalOperations.cpp:755-758, :864-867)
TW_HD int cgal_orientation(const double* p, const double* q, const double* r, const double* s) { return orient3d(q, r, s, p); }

TW_HD_NOINLINE int orient2d_exact(double px, double py, double qx, double qy, double rx, double ry) {
    double pq[4], qr[4], rp[4], t[8], det[12];
    int npq = cross2(px, py, qx, qy, pq), nqr = cross2(qx, qy, rx, ry, qr), nrp = cross2(rx, ry, px, py, rp);
    int nt = exp_sum(pq, npq, qr, nqr, t);
    int nd = exp_sum(t, nt, rp, nrp, det);
    return exp_sign(det, nd);
}
TW_HD int orient2d(double px, double py, double qx, double qy, double rx, double ry) {
    double l = dmul(dsub(px, rx), dsub(qy, ry));
    double r = dmul(dsub(py, ry), dsub(qx, rx));
    double det = dsub(l, r);
    double sum;
    if (l > 0.0) { if (r <= 0.0) return (det > 0.0) - (det < 0.0); sum = dadd(l, r); }
    else if (l < 0.0) { if (r >= 0.0) return (det > 0.0) - (det < 0.0); sum = dsub(-l, r); }
    else return (det > 0.0) - (det < 0.0);
    double bound = dmul((3.0 + 16.0 * EPS_HALF) * EPS_HALF, sum);
    if (det >= bound || -det >= bound) return (det > 0.0) - (det < 0.0);
    return orient2d_exact(px, py, qx, qy, rx, ry);
}
// Triangle_3::is_degenerate() == collinear (LocalOperations.cpp:1048)
TW_HD bool triangle_is_degenerate(const double* p, const double* q, const double* r) {
    if (orient2d(p[0], p[1], q[0], q[1], r[0], r[1]) != 0) return false;
    if (orient2d(p[0], p[2], q[0], q[2], r[0], r[2]) != 0) return false;
    if (orient2d(p[1], p[2], q[1], q[2], r[1], r[2]) != 0) return false;
    return true;
}
}  // namespace exact

}  // namespace tw
