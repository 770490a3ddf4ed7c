/*
 * oracle/ref_wrap.cpp -- TEST INFRASTRUCTURE ONLY.
 * extern "C" wrapper around pieces of the UNMODIFIED reference, compiled by oracle/ref_build.sh from where the
 * sources lie under /root/reference into oracle/_ref/libtetwild_ref.so (git-ignored; never copied into the repo):
 *   (1) src/tetwild/LocalOperations.cpp:28-291  AMIPS energy / Jacobian / Hessian (needs only <cmath>)
 *   (2) src/tetwild/Common.cpp:143-256          sampleTriangle, over the vec3 of oracle/shim
 *   (3) src/tetwild/geogram/mesh_AABB.{h,cpp}   the whole tree, over the geogram API shim of oracle/shim
 * The generated .inc files are line-range extracts written to oracle/_ref/gen/ at build time.
 * Used to pin the oracle restatement (tests/test_oracle_pin.py, tests/golden/make_golden.py) and, where it is the
 * reference's own code, as bench.py's cpu_baseline with kind "reference".
 */
#include <cmath>
#include <array>
#include <vector>
#include <cstdint>
#include <cstring>
#include <limits>
#include <geogram/basic/geometry.h>
#include <geogram/mesh/mesh.h>
#include <tetwild/geogram/mesh_AABB.h>
#include <tetwild/DistanceQuery.h>  /* the reference's own get_point_facet_nearest_point (DistanceQuery.h:20-39) */

extern "C" int ora_triangle_is_degenerate(const double* p, const double* q, const double* r);  /* exact: CGAL's Triangle_3::is_degenerate on doubles */

using std::pow;

namespace tetwild {
struct LocalOperations {
    static double comformalAMIPSEnergy_new(const double* T);
    static void comformalAMIPSJacobian_new(const double* T, double* result_0);
    static void comformalAMIPSHessian_new(const double* T, double* result_0);
};
#include "amips_lines.inc"
#include "sample_lines.inc"
}  // namespace tetwild

extern "C" {

double ref_amips_energy(const double* T) { return tetwild::LocalOperations::comformalAMIPSEnergy_new(T); }
void ref_amips_jacobian(const double* T, double* J) { tetwild::LocalOperations::comformalAMIPSJacobian_new(T, J); }
void ref_amips_hessian(const double* T, double* H) { tetwild::LocalOperations::comformalAMIPSHessian_new(T, H); }

void ref_amips_ejh_soa(const double* const* Ts, double* E, double* J3, double* H9, uint64_t n, int threads) {
#pragma omp parallel for schedule(static) num_threads(threads > 0 ? threads : 1)
    for (int64_t i = 0; i < (int64_t)n; ++i) {
        double T[12];
        for (int k = 0; k < 12; ++k) T[k] = Ts[k][i];
        if (E) E[i] = ref_amips_energy(T);
        if (J3) ref_amips_jacobian(T, J3 + 3 * i);
        if (H9) ref_amips_hessian(T, H9 + 9 * i);
    }
}

// VertexSmoother::NewtonsUpdate (VertexSmoother.cpp:627-702) over many one-rings, restated around the reference's own
// E / J / H text: gather the member tet, rotate the centre vertex to slot 0 (:640-651), add the three results (:652-672),
// apply the acceptance rules (:680-699). OpenMP over rings (the reference itself runs one ring at a time).
void ref_amips_ring_ejh(const double* V, const int32_t* tets4, const int32_t* t_ids, const uint64_t* off, const int32_t* center, uint64_t nG,
                        double* E, double* J3, double* H9, uint8_t* ok, int threads) {
#pragma omp parallel for schedule(dynamic, 64) num_threads(threads > 0 ? threads : 1)
    for (int64_t g = 0; g < (int64_t)nG; ++g) {
        double e = 0, J[3] = {0, 0, 0}, H[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        for (uint64_t k = off[g]; k < off[g + 1]; ++k) {
            const int32_t* t = tets4 + 4 * (size_t)(t_ids ? t_ids[k] : (int64_t)k);
            int start = 0;
            for (int j = 0; j < 4; ++j)
                if (t[j] == center[g]) { start = j; break; }
            double T[12], j3[3], h9[9];
            for (int j = 0; j < 4; ++j)
                for (int c = 0; c < 3; ++c) T[3 * j + c] = V[3 * (size_t)t[(start + j) % 4] + c];
            e += ref_amips_energy(T);
            ref_amips_jacobian(T, j3);
            ref_amips_hessian(T, h9);
            for (int c = 0; c < 3; ++c) J[c] += j3[c];
            for (int c = 0; c < 9; ++c) H[c] += h9[c];
        }
        bool good = true;
        if (std::isinf(e)) e = 1e50;
        if (std::isnan(e) || e <= 0) good = false;
        for (int c = 0; c < 3; ++c) if (!std::isfinite(J[c])) good = false;
        for (int c = 0; c < 9; ++c) if (!std::isfinite(H[c])) good = false;
        E[g] = e;
        for (int c = 0; c < 3; ++c) J3[3 * g + c] = J[c];
        for (int c = 0; c < 9; ++c) H9[9 * g + c] = H[c];
        if (ok) ok[g] = good;
    }
}

uint64_t ref_sample_triangle(const double* tri9, double sampling_dist, double* out_xyz, uint64_t cap) {
    std::array<GEO::vec3, 3> vs;
    for (int i = 0; i < 3; ++i) vs[i] = GEO::vec3(tri9[3 * i], tri9[3 * i + 1], tri9[3 * i + 2]);
    std::vector<GEO::vec3> ps;
    tetwild::sampleTriangle(vs, ps, sampling_dist);
    for (size_t i = 0; i < ps.size() && i < cap; ++i) { out_xyz[3 * i] = ps[i].x; out_xyz[3 * i + 1] = ps[i].y; out_xyz[3 * i + 2] = ps[i].z; }
    return ps.size();
}

struct ref_tree {
    GEO::Mesh mesh;
    GEO::MeshFacetsAABBWithEps* aabb;
};

/* F must already be in the desired (spatially sorted) order: the tree is built with reorder=false */
ref_tree* ref_tree_create(const double* V, uint32_t nV, const uint32_t* F, uint32_t nF) {
    ref_tree* t = new ref_tree;
    t->mesh.vertices.xyz.assign(V, V + 3 * (size_t)nV);
    t->mesh.facet_corners.v.assign(F, F + 3 * (size_t)nF);
    t->mesh.facets.n = nF;
    t->aabb = new GEO::MeshFacetsAABBWithEps(t->mesh, false);
    return t;
}
void ref_tree_destroy(ref_tree* t) { if (t) { delete t->aabb; delete t; } }

void ref_tree_nearest(const ref_tree* t, const double* P, uint64_t n, uint32_t* facet, double* nearest, double* d2, int threads) {
#pragma omp parallel for schedule(dynamic, 1024) num_threads(threads > 0 ? threads : 1)
    for (int64_t i = 0; i < (int64_t)n; ++i) {
        GEO::vec3 p(P[3 * i], P[3 * i + 1], P[3 * i + 2]), np;
        double d;
        GEO::index_t f = t->aabb->nearest_facet(p, np, d);
        if (facet) facet[i] = f;
        if (nearest) { nearest[3 * i] = np.x; nearest[3 * i + 1] = np.y; nearest[3 * i + 2] = np.z; }
        if (d2) d2[i] = d;
    }
}

/* per point: facet_in_envelope_with_hint(p, eps2, NO_FACET, ...) then sq_dist > eps2, as at LocalOperations.cpp:1083-1088 */
void ref_tree_envelope_points_out(const ref_tree* t, const double* P, uint64_t n, double eps2, uint8_t* out,
                                  uint32_t* facet, double* d2, int threads) {
#pragma omp parallel for schedule(dynamic, 1024) num_threads(threads > 0 ? threads : 1)
    for (int64_t i = 0; i < (int64_t)n; ++i) {
        GEO::vec3 p(P[3 * i], P[3 * i + 1], P[3 * i + 2]), np;
        double d = std::numeric_limits<double>::max();
        GEO::index_t f = GEO::NO_FACET;
        t->aabb->facet_in_envelope_with_hint(p, eps2, f, np, d);
        out[i] = d > eps2;
        if (facet) facet[i] = f;
        if (d2) d2[i] = d;
    }
}

/* LocalOperations::isFaceOutEnvelop_sampling (LocalOperations.cpp:1046-1109) restated around the reference's OWN
 * sampleTriangle (Common.cpp:143-255), get_point_facet_nearest_point (DistanceQuery.h) and tree (mesh_AABB.cpp): degenerate
 * face -> IN (:1048, exact predicate from oracle/predicates.c), samples visited from the middle, wrapping (:1078), previous
 * facet as hint (:1080-1086), first OUT sample decides (:1088-1093). degenerate_shortcut = 0: the per-face body of
 * Preprocess::isOutEnvelop (Preprocess.cpp:652-739), which samples every face and visits the samples in order. */
void ref_tree_faces_out(const ref_tree* t, const double* tris9, uint64_t n, double sampling_dist, double eps2, int degenerate_shortcut,
                        uint8_t* out, uint64_t* num_samples, int threads) {
#pragma omp parallel num_threads(threads > 0 ? threads : 1)
    {
        std::vector<GEO::vec3> ps;
#pragma omp for schedule(dynamic, 16)
        for (int64_t f = 0; f < (int64_t)n; ++f) {
            const double* T = tris9 + 9 * f;
            out[f] = 0;
            if (num_samples) num_samples[f] = 0;
            if (degenerate_shortcut && ora_triangle_is_degenerate(T, T + 3, T + 6)) continue;
            std::array<GEO::vec3, 3> vs;
            for (int i = 0; i < 3; ++i) vs[i] = GEO::vec3(T[3 * i], T[3 * i + 1], T[3 * i + 2]);
            ps.clear();
            tetwild::sampleTriangle(vs, ps, sampling_dist);
            if (num_samples) num_samples[f] = ps.size();
            GEO::vec3 nearest_point;
            double sq_dist = std::numeric_limits<double>::max();
            GEO::index_t prev_facet = GEO::NO_FACET;
            const size_t ps_size = ps.size();
            size_t cnt = 0;
            for (size_t i = degenerate_shortcut ? ps_size / 2 : 0; cnt < ps_size; i = (i + 1) % ps_size, ++cnt) {
                const GEO::vec3& current_point = ps[i];
                if (prev_facet != GEO::NO_FACET) tetwild::get_point_facet_nearest_point(t->mesh, current_point, prev_facet, nearest_point, sq_dist);
                if (sq_dist > eps2) t->aabb->facet_in_envelope_with_hint(current_point, eps2, prev_facet, nearest_point, sq_dist);
                if (sq_dist > eps2) { out[f] = 1; break; }
            }
        }
    }
}

}  // extern "C"
