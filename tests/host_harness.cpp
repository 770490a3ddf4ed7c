// tests/host_harness.cpp -- TEST INFRASTRUCTURE ONLY.
// Compiles the __host__ __device__ numeric core of the product (tetwild_b200/csrc/tw_math.cuh, sampling.cuh) as plain
// C++ so the CPU-only test tier can check the exact source the kernels run (closed-form AMIPS, record-based
// point-triangle distance, run-decomposed sampleTriangle, exact predicates) against the oracle without a GPU.
// It is not part of libtetwild_gpu.so and is never used by the product path.
#include <cstdint>
#include <cstring>
#include "../tetwild_b200/csrc/tw_math.cuh"
#include "../tetwild_b200/csrc/sampling.cuh"
#include "../tetwild_b200/csrc/winding_math.cuh"

extern "C" {

// running sum of atan2(y_k, x_k) through the product's complex-product accumulator (winding_math.cuh::Angle), renormalised
// every `tile` factors like the kernel does per staged tile
double hh_angle_sum(const double* x, const double* y, const uint8_t* skip, uint64_t n, int tile, int* k_out) {
    tww::Angle a;
    a.init();
    for (uint64_t i = 0; i < n; ++i) {
        a.mul(x[i], y[i], skip && skip[i]);
        if ((i + 1) % (uint64_t)tile == 0) a.renorm();
    }
    a.renorm();
    if (k_out) *k_out = a.k;
    return a.total();
}
double hh_norm3(double x, double y, double z) { return tww::norm3(x, y, z); }
// the conservative FP32 lower bound of the squared distance from a double point to a float box (tw_math.cuh)
float hh_box_d2_lb(const double* p, const float* box6) {
    const tw::PointF q = tw::bracket(tw::mk(p[0], p[1], p[2]));
    return tw::box_d2_lb(q, box6[0], box6[1], box6[2], box6[3], box6[4], box6[5]);
}
// one Van Oosterom-Strackee factor (x + i y) of triangle (a, b, c) seen from p, with the kernel's arithmetic
void hh_solid_angle_factor(const double* p, const double* A, const double* B, const double* C, double* xy) {
    const double ax = A[0] - p[0], ay = A[1] - p[1], az = A[2] - p[2];
    const double bx = B[0] - p[0], by = B[1] - p[1], bz = B[2] - p[2];
    const double cx = C[0] - p[0], cy = C[1] - p[1], cz = C[2] - p[2];
    const double la = tww::norm3(ax, ay, az), lb = tww::norm3(bx, by, bz), lc = tww::norm3(cx, cy, cz);
    xy[1] = ax * (by * cz - bz * cy) + bx * (cy * az - cz * ay) + cx * (ay * bz - az * by);
    xy[0] = la * lb * lc + (bx * cx + by * cy + bz * cz) * la + (cx * ax + cy * ay + cz * az) * lb + (ax * bx + ay * by + az * bz) * lc;
}

void hh_amips_ejh(const double* T12, double* E, double* J3, double* H9) {
    tw::Amips r;
    tw::amips_eval<true>(T12, r);
    *E = r.E;
    for (int k = 0; k < 3; ++k) J3[k] = r.J[k];
    H9[0] = r.H[0]; H9[1] = r.H[1]; H9[2] = r.H[2];
    H9[3] = r.H[1]; H9[4] = r.H[3]; H9[5] = r.H[4];
    H9[6] = r.H[2]; H9[7] = r.H[4]; H9[8] = r.H[5];
}

void hh_amips_ejh_batch(const double* T /* 12 x n, row-major */, uint64_t n, double* E, double* J3, double* H9) {
    for (uint64_t i = 0; i < n; ++i) {
        double x[12];
        for (int k = 0; k < 12; ++k) x[k] = T[(size_t)k * n + i];
        hh_amips_ejh(x, E + i, J3 + 3 * i, H9 + 9 * i);
    }
}

double hh_amips_energy(const double* T12) {
    tw::Amips r;
    tw::amips_eval<false>(T12, r);
    return r.E;
}

double hh_tri_sqdist(const double* p, const double* v0, const double* v1, const double* v2, double* nearest) {
    tw::TriRec r;
    tw::make_trirec(v0, v1, v2, 0, r);
    tw::V3 P = tw::mk(p[0], p[1], p[2]);
    if (r.flags & 1u) {
        double tv[9];
        memcpy(tv, v0, 24); memcpy(tv + 3, v1, 24); memcpy(tv + 6, v2, 24);
        tw::V3 q;
        double d = tw::tri_sqdist_degenerate(P, tv, q);
        nearest[0] = q.x; nearest[1] = q.y; nearest[2] = q.z;
        return d;
    }
    double s, t;
    double d = tw::tri_sqdist_rec(P, r, s, t);
    tw::V3 q = tw::tri_nearest_point(r, s, t);
    nearest[0] = q.x; nearest[1] = q.y; nearest[2] = q.z;
    return d;
}

// oriented facet bound (tw_math.cuh::TriBound): the lower bound the nearest-facet kernels prune leaves with
double hh_tri_bound_lb2(const double* p, const double* tri9) {
    tw::TriRec r;
    tw::make_trirec(tri9, tri9 + 3, tri9 + 6, 0, r);
    tw::TriBound B;
    tw::make_bound(tri9, (r.flags & 1u) != 0, B);
    return tw::bound_lb2(B, tw::mk(p[0], p[1], p[2]));
}

struct Sink {
    double* out; uint64_t cap, n;
    void operator()(tw::V3 p) { if (n < cap) { out[3 * n] = p.x; out[3 * n + 1] = p.y; out[3 * n + 2] = p.z; } ++n; }
};
uint64_t hh_sample_triangle(const double* tri9, double sd, double* out, uint64_t cap) {
    tw::SamplePlan P;
    tw::make_plan(tri9, sd, P);
    Sink s{out, cap, 0};
    tw::enumerate_samples(P, s);
    return s.n;
}

int hh_orient3d(const double* a, const double* b, const double* c, const double* d) { return tw::exact::orient3d(a, b, c, d); }
int hh_orient3d_exact(const double* a, const double* b, const double* c, const double* d) { return tw::exact::orient3d_exact(a, b, c, d); }
int hh_cgal_orientation(const double* p, const double* q, const double* r, const double* s) { return tw::exact::cgal_orientation(p, q, r, s); }
int hh_triangle_is_degenerate(const double* p, const double* q, const double* r) { return tw::exact::triangle_is_degenerate(p, q, r) ? 1 : 0; }

}  // extern "C"
