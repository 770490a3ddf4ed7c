// sampling.cuh -- sampleTriangle (src/tetwild/Common.cpp:143-255) decomposed into independent "runs" so that a warp
// can generate the samples of one face in parallel while producing exactly the points (bit for bit, including the
// int truncations that decide the sample count) that the sequential reference pushes into `ps`.
//
// Reference order of `ps`:
//   A  base edge       v0 + (n01*sd)*n          n = 0 .. floor(N)          (:167-169)   then v1 (:170)
//      [if M < 1: v2, return (:174-177)]
//   R  rows m=1..M     v + (i*n01)*sd           i = 0 .. N1(m)             (:187-206), stops at the first row whose
//                                                                            end points are closer than sd (:197-198)
//      v2 (:207)
//   B  edge v1->v2     v1 + (n12*sd)*n          n = 1 .. floor(N_B)        (:210-218)
//   C  edge v2->v0     v2 + (n20*sd)*n          n = 1 .. floor(N_C)        (:220-228)
// All arithmetic is "strict" (tw_math.cuh): same operations, same order, no FMA.
#pragma once
#include "tw_math.cuh"

namespace tw {

struct SamplePlan {
    int kind;        // 0: three vertices only (:154-158); 1: base edge + v1 + v2 (M < 1); 2: full
    V3 v0, v1, v2;   // rotated so that v0->v1 is the longest edge
    V3 n01, n12, n20, n02;
    double sd, sqrt3_2;
    double tan_v0, sin_v0, sin_v1;
    int nA;          // base-edge samples n = 0 .. nA-1
    int M;           // candidate rows
    int nB, nC;      // edge samples n = 1 .. nB / nC
};

TW_HD int count_le(double N) {  // number of integers n >= 0 with n <= N  (loop `for (int n = 0; n <= N; n++)`)
    return N < 0.0 ? 0 : (int)floor(N) + 1;
}

TW_HD void make_plan(const double* tri9, double sd, SamplePlan& P) {
    V3 vs[3] = {mk(tri9[0], tri9[1], tri9[2]), mk(tri9[3], tri9[4], tri9[5]), mk(tri9[6], tri9[7], tri9[8])};
    P.sd = sd;
    P.sqrt3_2 = ddiv(dsqrt(3.0), 2.0);
    P.n01 = P.n12 = P.n20 = P.n02 = mk(0.0, 0.0, 0.0);
    P.tan_v0 = P.sin_v0 = P.sin_v1 = 0.0;
    double ls[3];
    for (int i = 0; i < 3; ++i) ls[i] = vlen2(vsub(vs[i], vs[(i + 1) % 3]));
    int max_i = 0;  // std::minmax_element returns the LAST largest element
    for (int i = 1; i < 3; ++i)
        if (!(ls[i] < ls[max_i])) max_i = i;
    double N = ddiv(dsqrt(ls[max_i]), sd);
    P.nA = P.M = P.nB = P.nC = 0;
    if (N <= 1.0) { P.kind = 0; P.v0 = vs[0]; P.v1 = vs[1]; P.v2 = vs[2]; return; }  // caller's order (:155-156)
    P.v0 = vs[max_i]; P.v1 = vs[(max_i + 1) % 3]; P.v2 = vs[(max_i + 2) % 3];
    if (N == (double)(int)N) N = dsub(N, 1.0);
    V3 e01 = vsub(P.v1, P.v0);
    P.n01 = vnormalize(e01);
    P.nA = count_le(N);
    V3 e02 = vsub(P.v2, P.v0);
    double dt = vdot(e02, e01);
    V3 foot = mk(dadd(ddiv(dmul(dt, e01.x), ls[max_i]), P.v0.x), dadd(ddiv(dmul(dt, e01.y), ls[max_i]), P.v0.y),
                 dadd(ddiv(dmul(dt, e01.z), ls[max_i]), P.v0.z));
    double h = vdist(foot, P.v2);
    P.M = (int)ddiv(h, dmul(P.sqrt3_2, sd));
    if (P.M < 1) { P.kind = 1; P.M = 0; return; }
    P.kind = 2;
    P.n02 = vnormalize(e02);
    V3 e12 = vsub(P.v2, P.v1);
    P.n12 = vnormalize(e12);
    V3 e10 = vsub(P.v0, P.v1);
    double c0 = dsqrt(vlen2(vcross(e02, e01)));
    double c1 = dsqrt(vlen2(vcross(e12, e10)));
    P.sin_v0 = ddiv(c0, dmul(vdist(P.v0, P.v2), vdist(P.v0, P.v1)));
    P.tan_v0 = ddiv(c0, vdot(e02, e01));
    P.sin_v1 = ddiv(c1, dmul(vdist(P.v1, P.v2), vdist(P.v0, P.v1)));
    double NB = ddiv(dsqrt(ls[(max_i + 1) % 3]), sd);
    if (NB > 1.0) {
        if (NB == (double)(int)NB) NB = dsub(NB, 1.0);
        P.nB = count_le(NB) - 1;  // n = 1 .. floor(NB)
        if (P.nB < 0) P.nB = 0;
    }
    double NC = ddiv(dsqrt(ls[(max_i + 2) % 3]), sd);
    if (NC > 1.0) {
        if (NC == (double)(int)NC) NC = dsub(NC, 1.0);
        P.nC = count_le(NC) - 1;
        if (P.nC < 0) P.nC = 0;
        P.n20 = vnormalize(vsub(P.v0, P.v2));
    }
}

struct RowPlan {
    bool stop;  // reference `break` at this row
    V3 v;       // first sample of the row
    int N1;     // samples i = 0 .. N1
};

TW_HD void make_row(const SamplePlan& P, int m, RowPlan& R) {
    const double k = ddiv(P.sqrt3_2, P.tan_v0);
    int n = (int)dadd(dmul(k, (double)m), 0.5);
    int n1 = (int)dmul(k, (double)m);
    if (m % 2 == 0 && n == n1) n += 1;
    const double msd = dmul(dmul((double)m, P.sqrt3_2), P.sd);
    const double s0 = ddiv(msd, P.sin_v0), s1 = ddiv(msd, P.sin_v1);
    V3 v0m = mk(dadd(P.v0.x, dmul(s0, P.n02.x)), dadd(P.v0.y, dmul(s0, P.n02.y)), dadd(P.v0.z, dmul(s0, P.n02.z)));
    V3 v1m = mk(dadd(P.v1.x, dmul(s1, P.n12.x)), dadd(P.v1.y, dmul(s1, P.n12.y)), dadd(P.v1.z, dmul(s1, P.n12.z)));
    R.stop = vdist(v0m, v1m) <= P.sd;
    if (R.stop) { R.N1 = -1; R.v = v0m; return; }
    const double delta = dmul(dsub(dadd((double)n, ddiv((double)(m % 2), 2.0)), ddiv(dmul((double)m, P.sqrt3_2), P.tan_v0)), P.sd);
    R.v = mk(dadd(v0m.x, dmul(delta, P.n01.x)), dadd(v0m.y, dmul(delta, P.n01.y)), dadd(v0m.z, dmul(delta, P.n01.z)));
    R.N1 = (int)ddiv(vdist(R.v, v1m), P.sd);
}

// o + (dir*sd)*n     (base edge and the two other edges)
TW_HD V3 edge_sample(V3 o, V3 dir, double sd, int n) {
    return mk(dadd(o.x, dmul(dmul(dir.x, sd), (double)n)), dadd(o.y, dmul(dmul(dir.y, sd), (double)n)), dadd(o.z, dmul(dmul(dir.z, sd), (double)n)));
}
// v + (i*dir)*sd     (row interior)
TW_HD V3 row_sample(V3 v, V3 dir, double sd, int i) {
    return mk(dadd(v.x, dmul(dmul((double)i, dir.x), sd)), dadd(v.y, dmul(dmul((double)i, dir.y), sd)), dadd(v.z, dmul(dmul((double)i, dir.z), sd)));
}

// Sequential enumeration in the reference's order (used by twg_sample_triangle and the host harness).
template <class SINK>
TW_HD void enumerate_samples(const SamplePlan& P, SINK& sink) {
    if (P.kind == 0) { sink(P.v0); sink(P.v1); sink(P.v2); return; }
    for (int n = 0; n < P.nA; ++n) sink(edge_sample(P.v0, P.n01, P.sd, n));
    sink(P.v1);
    if (P.kind == 1) { sink(P.v2); return; }
    for (int m = 1; m <= P.M; ++m) {
        RowPlan R;
        make_row(P, m, R);
        if (R.stop) break;
        for (int i = 0; i <= R.N1; ++i) sink(row_sample(R.v, P.n01, P.sd, i));
    }
    sink(P.v2);
    for (int n = 1; n <= P.nB; ++n) sink(edge_sample(P.v1, P.n12, P.sd, n));
    for (int n = 1; n <= P.nC; ++n) sink(edge_sample(P.v2, P.n20, P.sd, n));
}

}  // namespace tw
