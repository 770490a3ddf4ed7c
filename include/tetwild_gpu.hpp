/*
 * tetwild_gpu.hpp -- C++11 host-side adapters over the C ABI (tetwild_gpu.h), header-only.
 *
 * They keep the reference's call signatures so that TetWild's sequential host code (MeshRefinement scheduler, the
 * four local operations, InoutFiltering) can be pointed at the GPU path by swapping a type, not by rewriting call
 * sites. Nothing here needs CGAL, Eigen, geogram or libigl: point / triangle / matrix arguments are templates that
 * only use what the reference's own call sites use (operator[] on points, operator()(i,j) / rows() / cols() /
 * resize() on matrices, vertices.nb() / facets.vertex(f,lv) on GEO::Mesh), so the reference's types bind as they are.
 *
 *   reference (Yixin-Hu/TetWild @49de8cd)                                        adapter
 *   ---------------------------------------------------------------------------  ----------------------------------
 *   GEO::MeshFacetsAABBWithEps      src/tetwild/geogram/mesh_AABB.h:64-226       twg::MeshFacetsAABBWithEps
 *   LocalOperations::isFaceOutEnvelop / isPointOutEnvelop /                      twg::LocalOperations
 *     isPointOutBoundaryEnvelop / calTetQualities / comformalAMIPS*_new
 *                                   src/tetwild/LocalOperations.h:69,97-100,109-111
 *   VertexSmoother::NewtonsUpdate / getNewEnergy                                 twg::VertexSmoother
 *                                   src/tetwild/VertexSmoother.h:30-31
 *   ispc::energy_ispc               src/ispc/energy.ispc:7-21                    twg::energy_ispc
 *   igl::winding_number(V,F,O,W)    src/tetwild/InoutFiltering.cpp:45            twg::winding_number
 *   InoutFiltering::filter          src/tetwild/InoutFiltering.cpp:23-82         twg::InoutFiltering::filter
 *
 * Every adapter has the reference's one-query signature AND a batched overload (std::vector in / out): the scheduler
 * stays sequential, but a call that goes to the GPU costs a launch + two PCIe hops (~20 us), so callers that already
 * hold many independent queries (VertexSmoother.cpp:216-241, MeshRefinement.cpp:51, EdgeSplitter.cpp:130-147,
 * EdgeCollapser.cpp:727-775, the whole winding filter) should use the batched form. INTEGRATION.md shows the patches.
 *
 * Errors: the C ABI returns codes; the adapters throw twg::Error (the reference throws TetWildError, Exception.h:19).
 * There is no CPU fallback anywhere: without libtetwild_gpu.so and an sm_100-class device construction throws.
 */
#ifndef TETWILD_GPU_HPP
#define TETWILD_GPU_HPP

#include <array>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <cmath>
#include <vector>

#include "tetwild_gpu.h"

namespace twg {

typedef uint32_t index_t;                      /* GEO::index_t */
static const index_t NO_FACET = TWG_NO_FACET;  /* GEO::NO_FACET */

struct Error : public std::runtime_error {
    int code;
    Error(int c, const std::string& what) : std::runtime_error(what), code(c) {}
};

/* One device. Not copyable; shared by reference like the reference shares `State&`. */
class Context {
public:
    explicit Context(int device_id = 0) : h_(nullptr) {
        int rc = twg_create(&h_, device_id);
        if (rc != 0) throw Error(rc, "twg_create failed (no sm_100-class GPU or library built for another arch; there is no CPU fallback)");
    }
    /* one context over several devices: handles are replicated, host-buffer batches are split by index range */
    explicit Context(const std::vector<int>& device_ids) : h_(nullptr) {
        int rc = twg_create_multi(&h_, device_ids.data(), (int)device_ids.size());
        if (rc != 0) throw Error(rc, "twg_create_multi failed (a device is missing or not sm_100-class; there is no CPU fallback)");
    }
    ~Context() { twg_destroy(h_); }
    twg_ctx* handle() const { return h_; }
    int num_devices() const { return twg_num_devices(h_); }
    void set_option(const char* name, double value) const { check(twg_set_option(h_, name, value)); }
    void check(int rc) const {
        if (rc != 0) throw Error(rc, std::string("libtetwild_gpu: ") + twg_last_error(h_));
    }
    uint64_t launches() const { return twg_launch_count(h_); }

private:
    Context(const Context&);
    Context& operator=(const Context&);
    twg_ctx* h_;
};

/* minimal GEO::vec3 stand-in for callers that have no geogram; any type with operator[] works in the templates */
struct vec3 {
    double x, y, z;
    vec3() : x(0), y(0), z(0) {}
    vec3(double a, double b, double c) : x(a), y(b), z(c) {}
    double& operator[](int i) { return (&x)[i]; }
    const double& operator[](int i) const { return (&x)[i]; }
};

/* ------------------------------------------------------------------------------------------------------------------
 * GEO::MeshFacetsAABBWithEps (mesh_AABB.h:64-226). Differences that a maintainer must know:
 *   - the reference reorders the caller's mesh in place (mesh_AABB.cpp:368-370) and returns facet ids of the REORDERED
 *     mesh; this class never touches caller data and returns ids in the caller's numbering;
 *   - facet_in_envelope*: the reference stops at the first facet within sq_epsilon, so WHICH facet / point it reports is
 *     traversal dependent; here the exact nearest facet is reported (it satisfies the same contract: sq_dist <=
 *     sq_epsilon iff some facet is within the envelope). The batched decision path uses the early-exit kernel.
 * ---------------------------------------------------------------------------------------------------------------- */
class MeshFacetsAABBWithEps {
public:
    /* raw arrays: V = nV*3 doubles xyz-interleaved, F = nF*3 vertex indices */
    MeshFacetsAABBWithEps(Context& ctx, const double* V, uint32_t nV, const uint32_t* F, uint32_t nF) : ctx_(ctx), s_(nullptr) {
        ctx_.check(twg_surface_create(ctx_.handle(), V, nV, F, nF, &s_));
    }
    /* GEO::Mesh-like: M.vertices.nb(), M.vertices.point_ptr(v), M.facets.nb(), M.facets.vertex(f, lv)
     * (the accessors mesh_AABB.cpp:63-79 itself uses). `reorder` is accepted for signature compatibility. */
    template <class MESH>
    MeshFacetsAABBWithEps(Context& ctx, const MESH& M, bool reorder = true) : ctx_(ctx), s_(nullptr) {
        (void)reorder;
        const uint32_t nV = (uint32_t)M.vertices.nb(), nF = (uint32_t)M.facets.nb();
        std::vector<double> V(3 * (size_t)nV);
        std::vector<uint32_t> F(3 * (size_t)nF);
        for (uint32_t v = 0; v < nV; ++v)
            for (int c = 0; c < 3; ++c) V[3 * (size_t)v + c] = M.vertices.point_ptr(v)[c];
        for (uint32_t f = 0; f < nF; ++f)
            for (int lv = 0; lv < 3; ++lv) F[3 * (size_t)f + lv] = (uint32_t)M.facets.vertex(f, lv);
        ctx_.check(twg_surface_create(ctx_.handle(), V.data(), nV, F.data(), nF, &s_));
    }
    ~MeshFacetsAABBWithEps() { twg_surface_destroy(s_); }

    /* mesh_AABB.h:130-141 */
    template <class VEC3>
    index_t nearest_facet(const VEC3& p, VEC3& nearest_point, double& sq_dist) const {
        const double P[3] = {p[0], p[1], p[2]};
        double q[3];
        index_t f = NO_FACET;
        ctx_.check(twg_nearest(s_, P, 1, &f, q, &sq_dist));
        nearest_point[0] = q[0]; nearest_point[1] = q[1]; nearest_point[2] = q[2];
        return f;
    }
    /* mesh_AABB.h:162-176: the hint only accelerates the reference's search; the result is the same nearest facet */
    template <class VEC3>
    void nearest_facet_with_hint(const VEC3& p, index_t& nearest_facet_io, VEC3& nearest_point, double& sq_dist) const {
        nearest_facet_io = nearest_facet(p, nearest_point, sq_dist);
    }
    /* mesh_AABB.h:182-193 */
    template <class VEC3>
    index_t facet_in_envelope(const VEC3& p, double sq_epsilon, VEC3& nearest_point, double& sq_dist) const {
        (void)sq_epsilon;
        return nearest_facet(p, nearest_point, sq_dist);
    }
    /* mesh_AABB.h:199-213 */
    template <class VEC3>
    void facet_in_envelope_with_hint(const VEC3& p, double sq_epsilon, index_t& nearest_facet_io, VEC3& nearest_point, double& sq_dist) const {
        (void)sq_epsilon;
        nearest_facet_io = nearest_facet(p, nearest_point, sq_dist);
    }
    /* mesh_AABB.h:221-226 */
    template <class VEC3>
    double squared_distance(const VEC3& p) const {
        const double P[3] = {p[0], p[1], p[2]};
        double d2 = 0.0;
        ctx_.check(twg_nearest(s_, P, 1, nullptr, nullptr, &d2));
        return d2;
    }

    /* ---- batched forms (P = n*3 doubles) ---- */
    void nearest_facets(const double* P, uint64_t n, index_t* facets, double* nearest_points, double* sq_dists) const {
        ctx_.check(twg_nearest(s_, P, n, facets, nearest_points, sq_dists));
    }
    void squared_distances(const double* P, uint64_t n, double* sq_dists) const { ctx_.check(twg_nearest(s_, P, n, nullptr, nullptr, sq_dists)); }
    /* out[i] = 1 iff no facet lies within sqrt(sq_epsilon) of P_i (early-exit traversal, mesh_AABB.cpp:482-548) */
    void points_out_of_envelope(const double* P, uint64_t n, double sq_epsilon, uint8_t* out) const {
        ctx_.check(twg_envelope_points_out(s_, P, n, sq_epsilon, out));
    }
    /* out[i] = isFaceOutEnvelop_sampling(tris[i]), tris = n*9 doubles */
    void faces_out_of_envelope(const double* tris, uint64_t n, double sampling_dist, double sq_epsilon, uint8_t* out) const {
        ctx_.check(twg_envelope_faces_out(s_, tris, n, sampling_dist, sq_epsilon, out));
    }
    /* Preprocess::isOutEnvelop(new_f_ids, geo_sf_mesh, geo_face_tree) (Preprocess.cpp:643-747): true iff ANY face of the set
     * has a sample farther than eps; unlike isFaceOutEnvelop degenerate faces are sampled too. tris = n*9 doubles (the
     * V_in / F_in rows of new_f_ids); sq_epsilon is the 0.8-scaled eps_2 Preprocess runs with (:201-205). */
    bool isOutEnvelop(const double* tris, uint64_t n, double sampling_dist, double sq_epsilon) const {
        if (n == 0) return false;
        std::vector<uint8_t> out(n);
        ctx_.check(twg_envelope_faces_out_ex(s_, tris, n, sampling_dist, sq_epsilon, TWG_FACES_NO_DEGENERATE_SHORTCUT, out.data()));
        for (uint64_t i = 0; i < n; ++i)
            if (out[i]) return true;
        return false;
    }
    uint32_t nb_facets() const { return twg_surface_num_facets(s_); }
    twg_surface* handle() const { return s_; }
    Context& context() const { return ctx_; }

private:
    MeshFacetsAABBWithEps(const MeshFacetsAABBWithEps&);
    MeshFacetsAABBWithEps& operator=(const MeshFacetsAABBWithEps&);
    Context& ctx_;
    twg_surface* s_;
};

/* ispc::energy_ispc(V1_x, ..., V4_z, E, count) (src/ispc/energy.ispc:7-21) with the context in front */
inline void energy_ispc(Context& ctx, const double* V1_x, const double* V1_y, const double* V1_z, const double* V2_x, const double* V2_y,
                        const double* V2_z, const double* V3_x, const double* V3_y, const double* V3_z, const double* V4_x,
                        const double* V4_y, const double* V4_z, double* E, int count) {
    const double* T[12] = {V1_x, V1_y, V1_z, V2_x, V2_y, V2_z, V3_x, V3_y, V3_z, V4_x, V4_y, V4_z};
    ctx.check(twg_amips_energy_soa(ctx.handle(), T, E, (uint64_t)(count < 0 ? 0 : count)));
}

/* DelaunayTetrahedralization::getVoxelPoints (DelaunayTetrahedralization.cpp:61-106): the voxel-stuffing grid over [p_min, p_max]
 * -- N[i] = (int)(D_i / voxel_resolution) + 1 cells per axis (:76-80), planes p_min + d (j+1) plus the two box faces (:82-88), the eight
 * box corners left out (:95-97) -- filtered by distance to the input surface: a grid point is kept iff squared_distance(p) >=
 * voxel_resolution^2 / 4 (:92,:99-100). The reference asks the tree once per grid point; here ONE batched nearest query.
 * voxel_resolution is bbox_diag / 20 when the relative target edge length is below 5 (%), else the absolute edge length (:67-71).
 * The reference forms the plane coordinates in exact rationals and converts them to double; here one fused multiply-add
 * (a single rounding of the same exact value). Appends to voxel_points in the reference's order (i, j, k nested). */
inline double voxel_resolution(double relative_edge_length_percent, double absolute_edge_length, double bbox_diag) {
    return relative_edge_length_percent < 5.0 ? bbox_diag / 20.0 : absolute_edge_length;
}
inline void getVoxelPoints(const double p_min[3], const double p_max[3], const MeshFacetsAABBWithEps& geo_face_tree, double voxel_resolution,
                           std::vector<std::array<double, 3> >& voxel_points) {
    std::vector<double> ds[3];
    for (int i = 0; i < 3; ++i) {
        const double D = p_max[i] - p_min[i];
        const int N = (int)(D / voxel_resolution) + 1;
        const double d = D / N;
        ds[i].push_back(p_min[i]);
        for (int j = 0; j < N - 1; ++j) ds[i].push_back(std::fma(d, (double)(j + 1), p_min[i]));
        ds[i].push_back(p_max[i]);
    }
    std::vector<double> P;
    P.reserve(3 * ds[0].size() * ds[1].size() * ds[2].size());
    for (size_t i = 0; i < ds[0].size(); ++i)
        for (size_t j = 0; j < ds[1].size(); ++j)
            for (size_t k = 0; k < ds[2].size(); ++k) {
                if ((i == 0 || i == ds[0].size() - 1) && (j == 0 || j == ds[1].size() - 1) && (k == 0 || k == ds[2].size() - 1)) continue;
                P.push_back(ds[0][i]); P.push_back(ds[1][j]); P.push_back(ds[2][k]);
            }
    const uint64_t n = P.size() / 3;
    if (n == 0) return;
    std::vector<double> d2(n);
    geo_face_tree.squared_distances(P.data(), n, d2.data());
    const double min_dis = voxel_resolution * voxel_resolution / 4;
    for (uint64_t q = 0; q < n; ++q) {
        if (d2[q] < min_dis) continue;
        std::array<double, 3> pt = {{P[3 * q], P[3 * q + 1], P[3 * q + 2]}};
        voxel_points.push_back(pt);
    }
}

/* State::State (State.cpp:24-41): the kernel parameters derived from the user's eps_rel, --stage and the bbox diagonal, with
 * the reference's own expressions (eps_2 is compared bit for bit, so the threshold must be the same double):
 *   sampling_dist = eps_input / stage;  eps = eps_input - sampling_dist / sqrt(3) * (stage + 1 - sub_stage);  eps_2 = eps * eps
 * and the per-sub-stage growth of MeshRefinement.cpp:317-320,340-344:  eps += eps_delta;  eps_2 = eps * eps. */
struct EnvelopeParams {
    double eps, eps_2, sampling_dist, eps_delta;
    int stage, sub_stage;
    static EnvelopeParams from_args(double bbox_diag, double eps_rel, int stage = 1, int sub_stage = 1) {
        EnvelopeParams e;
        const double eps_input = bbox_diag * eps_rel;
        e.stage = stage;
        e.sub_stage = sub_stage;
        e.eps_delta = eps_input / stage / std::sqrt(3);                                   /* State.cpp:25 */
        e.sampling_dist = eps_input / stage;                                              /* State.cpp:37 */
        e.eps = eps_input - e.sampling_dist / std::sqrt(3) * (stage + 1 - sub_stage);     /* State.cpp:38 */
        e.eps_2 = e.eps * e.eps;                                                          /* State.cpp:39 */
        return e;
    }
    /* MeshRefinement.cpp:317-320 / :340-344 */
    void next_sub_stage() {
        eps += eps_delta;
        eps_2 = eps * eps;
        ++sub_stage;
    }
};

/* ------------------------------------------------------------------------------------------------------------------
 * LocalOperations (LocalOperations.h:32-127): the hot-path members only. Holds references exactly like the reference
 * (tet vertices are read through an accessor so that TetVertex::posf binds without a copy of the struct).
 *   POS: callable  const double* pos(int v_id)  -> the 3 doubles of tet_vertices[v_id].posf
 * ---------------------------------------------------------------------------------------------------------------- */
struct TetQuality {  /* TetmeshElements.h:66-111, the field the AMIPS path writes */
    double slim_energy;
    TetQuality() : slim_energy(0) {}
};

class LocalOperations {
public:
    LocalOperations(Context& ctx, const MeshFacetsAABBWithEps& geo_sf_tree, const MeshFacetsAABBWithEps& geo_b_tree, double eps_2,
                    double sampling_dist)
        : eps_2(eps_2), sampling_dist(sampling_dist), ctx_(ctx), geo_sf_tree_(geo_sf_tree), geo_b_tree_(geo_b_tree) {}

    double eps_2, sampling_dist;  /* State::eps_2, State::sampling_dist (State.h), mutable like the reference (eps grows per sub-stage) */

    /* LocalOperations.cpp:967-976 -> :1046-1109; TRI: tri[k][c] like Triangle_3f */
    template <class TRI>
    bool isFaceOutEnvelop(const TRI& tri) const {
        double t9[9];
        for (int k = 0; k < 3; ++k)
            for (int c = 0; c < 3; ++c) t9[3 * k + c] = tri[k][c];
        uint8_t out = 0;
        geo_sf_tree_.faces_out_of_envelope(t9, 1, sampling_dist, eps_2, &out);
        return out != 0;
    }
    /* batched: every candidate face of one collapse / smoothing step at once (EdgeCollapser.cpp:727-775, VertexSmoother.cpp:425) */
    template <class TRI>
    void isFaceOutEnvelop(const std::vector<TRI>& tris, std::vector<uint8_t>& is_out) const {
        std::vector<double> t9(9 * tris.size());
        for (size_t i = 0; i < tris.size(); ++i)
            for (int k = 0; k < 3; ++k)
                for (int c = 0; c < 3; ++c) t9[9 * i + 3 * k + c] = tris[i][k][c];
        is_out.assign(tris.size(), 0);
        geo_sf_tree_.faces_out_of_envelope(t9.data(), tris.size(), sampling_dist, eps_2, is_out.data());
    }
    /* LocalOperations.cpp:1034-1044 */
    template <class POINT>
    bool isPointOutEnvelop(const POINT& p) const { return geo_sf_tree_.squared_distance(p) > eps_2; }
    /* LocalOperations.cpp:1111-1121 */
    template <class POINT>
    bool isPointOutBoundaryEnvelop(const POINT& p) const { return geo_b_tree_.squared_distance(p) > eps_2; }

    /* LocalOperations.cpp:695-773 (+ :862-884). V = nV*3 doubles (posf of every tet vertex, refreshed by the caller when
     * vertices move), new_tets as in the reference. all_measure is ignored there too. */
    void calTetQualities(const double* V, uint32_t nV, const std::vector<std::array<int, 4> >& new_tets, std::vector<TetQuality>& tet_qs,
                         bool all_measure = false) const {
        (void)all_measure;
        tet_qs.resize(new_tets.size());
        if (new_tets.empty()) return;
        std::vector<double> e(new_tets.size());
        /* std::array<int,4> is 16 contiguous bytes: the vector IS the int32 tets4 array of the C ABI */
        const int32_t* t4 = reinterpret_cast<const int32_t*>(new_tets.data());
        ctx_.check(twg_amips_quality(ctx_.handle(), V, nV, t4, new_tets.size(), e.data()));
        for (size_t i = 0; i < e.size(); ++i) tet_qs[i].slim_energy = e[i];
    }

    /* LocalOperations.h:109-111: single-tet static forms, T = 12 doubles (4 vertices x xyz) */
    static double comformalAMIPSEnergy_new(Context& ctx, const double* T) {
        double E = 0;
        const double* p[12];
        for (int k = 0; k < 12; ++k) p[k] = T + k;
        ctx.check(twg_amips_ejh_soa(ctx.handle(), p, &E, nullptr, nullptr, 1));
        return E;
    }
    static void comformalAMIPSJacobian_new(Context& ctx, const double* T, double* result_0) {
        const double* p[12];
        for (int k = 0; k < 12; ++k) p[k] = T + k;
        ctx.check(twg_amips_ejh_soa(ctx.handle(), p, nullptr, result_0, nullptr, 1));
    }
    static void comformalAMIPSHessian_new(Context& ctx, const double* T, double* result_0) {
        const double* p[12];
        for (int k = 0; k < 12; ++k) p[k] = T + k;
        ctx.check(twg_amips_ejh_soa(ctx.handle(), p, nullptr, nullptr, result_0, 1));
    }

    Context& context() const { return ctx_; }

private:
    Context& ctx_;
    const MeshFacetsAABBWithEps& geo_sf_tree_;
    const MeshFacetsAABBWithEps& geo_b_tree_;
};

/* ------------------------------------------------------------------------------------------------------------------
 * VertexSmoother::NewtonsUpdate / getNewEnergy (VertexSmoother.cpp:627-702, :544-625).
 * Batched form = one call for the one-rings of MANY vertices (an independent set of the smoothing pass).
 * ---------------------------------------------------------------------------------------------------------------- */
class VertexSmoother {
public:
    /* V: nV*3 doubles; tets: the mesh's tet array (std::vector<std::array<int,4>>, 16-byte aligned storage) */
    VertexSmoother(Context& ctx, const double* V, uint32_t nV, const std::vector<std::array<int, 4> >& tets) : ctx_(ctx), V_(V), nV_(nV), tets_(tets) {}

    /* one vertex: J = 3 doubles, H = 9 doubles row-major (Eigen::Vector3d / Matrix3d .data() of a RowMajor or symmetric use) */
    bool NewtonsUpdate(const std::vector<int>& t_ids, int v_id, double& energy, double* J, double* H) const {
        const uint64_t off[2] = {0, (uint64_t)t_ids.size()};
        uint8_t ok = 0;
        ctx_.check(twg_amips_ring_ejh(ctx_.handle(), V_, nV_, reinterpret_cast<const int32_t*>(tets_.data()), tets_.size(), t_ids.data(), off, &v_id, 1,
                                      &energy, J, H, &ok));
        return ok != 0;
    }
    double getNewEnergy(const std::vector<int>& t_ids) const {
        const uint64_t off[2] = {0, (uint64_t)t_ids.size()};
        double e = 0;
        ctx_.check(twg_amips_ring_energy(ctx_.handle(), V_, nV_, reinterpret_cast<const int32_t*>(tets_.data()), tets_.size(), t_ids.data(), off, 1, &e));
        return e;
    }
    /* batched over groups: ring g = t_ids[group_off[g] .. group_off[g+1]) around vertex v_ids[g] */
    void NewtonsUpdate(const std::vector<int>& t_ids, const std::vector<uint64_t>& group_off, const std::vector<int>& v_ids, std::vector<double>& energy,
                       std::vector<double>& J3, std::vector<double>& H9, std::vector<uint8_t>& ok) const {
        const size_t g = v_ids.size();
        energy.resize(g); J3.resize(3 * g); H9.resize(9 * g); ok.resize(g);
        ctx_.check(twg_amips_ring_ejh(ctx_.handle(), V_, nV_, reinterpret_cast<const int32_t*>(tets_.data()), tets_.size(), t_ids.data(), group_off.data(),
                                      v_ids.data(), g, energy.data(), J3.data(), H9.data(), ok.data()));
    }
    void getNewEnergy(const std::vector<int>& t_ids, const std::vector<uint64_t>& group_off, std::vector<double>& energy) const {
        const size_t g = group_off.empty() ? 0 : group_off.size() - 1;
        energy.resize(g);
        ctx_.check(twg_amips_ring_energy(ctx_.handle(), V_, nV_, reinterpret_cast<const int32_t*>(tets_.data()), tets_.size(), t_ids.data(), group_off.data(),
                                         g, energy.data()));
    }

private:
    Context& ctx_;
    const double* V_;
    uint32_t nV_;
    const std::vector<std::array<int, 4> >& tets_;
};

/* ------------------------------------------------------------------------------------------------------------------
 * TetMesh: the scheduler's mesh kept resident on the device (twg_mesh, include/tetwild_gpu.h "resident tet mesh").
 * Mirrors `tet_vertices[].posf`, `tets`, `t_is_removed` and `tet_vertices[].conn_tets` (LocalOperations.h:35-45): create
 * once after the front end (MeshRefinement.cpp:208), call sync_vertices / sync_tets for what an accepted operation
 * changed, and the AMIPS batches of a pass then ship only ids in and results out.
 *   POS: callable  const double* pos(int v_id)  -> the 3 doubles of tet_vertices[v_id].posf
 * ---------------------------------------------------------------------------------------------------------------- */
class TetMesh {
public:
    template <class POS>
    TetMesh(Context& ctx, size_t n_vertices, POS pos, const std::vector<std::array<int, 4> >& tets, const std::vector<bool>& t_is_removed)
        : ctx_(ctx), m_(nullptr) {
        std::vector<double> V(3 * n_vertices);
        for (size_t v = 0; v < n_vertices; ++v) {
            const double* p = pos((int)v);
            V[3 * v] = p[0]; V[3 * v + 1] = p[1]; V[3 * v + 2] = p[2];
        }
        std::vector<int32_t> T(4 * tets.size());
        for (size_t t = 0; t < tets.size(); ++t) pack(tets[t], t < t_is_removed.size() && t_is_removed[t], &T[4 * t]);
        ctx_.check(twg_mesh_create(ctx_.handle(), V.data(), (uint32_t)n_vertices, T.data(), tets.size(), &m_));
    }
    ~TetMesh() { twg_mesh_destroy(m_); }
    twg_mesh* handle() const { return m_; }
    size_t num_vertices() const { return twg_mesh_num_vertices(m_); }
    size_t num_tets() const { return (size_t)twg_mesh_num_tets(m_); }

    /* after tet_vertices / tets grew (EdgeSplitter pushes new slots, EdgeSplitter.cpp:130-147) */
    void resize(size_t n_vertices, size_t n_tets) { ctx_.check(twg_mesh_resize(m_, (uint32_t)n_vertices, n_tets)); }
    /* posf of the listed vertices changed (accepted smoothing step, VertexSmoother.cpp:139-150; split / collapse) */
    template <class POS>
    void sync_vertices(const std::vector<int>& v_ids, POS pos) {
        std::vector<double> xyz(3 * v_ids.size());
        for (size_t i = 0; i < v_ids.size(); ++i) {
            const double* p = pos(v_ids[i]);
            xyz[3 * i] = p[0]; xyz[3 * i + 1] = p[1]; xyz[3 * i + 2] = p[2];
        }
        ctx_.check(twg_mesh_set_vertices(m_, v_ids.data(), xyz.data(), v_ids.size()));
    }
    /* the listed tet slots were rewritten or removed */
    void sync_tets(const std::vector<int>& t_ids, const std::vector<std::array<int, 4> >& tets, const std::vector<bool>& t_is_removed) {
        std::vector<int32_t> T(4 * t_ids.size());
        for (size_t i = 0; i < t_ids.size(); ++i) {
            const size_t t = (size_t)t_ids[i];
            pack(tets[t], t < t_is_removed.size() && t_is_removed[t], &T[4 * i]);
        }
        ctx_.check(twg_mesh_set_tets(m_, t_ids.data(), T.data(), t_ids.size()));
    }

    /* calTetQuality_AMIPS of the listed tets (LocalOperations.cpp:862-884); the whole-mesh loops of
     * VertexSmoother.cpp:216-241 and MeshRefinement.cpp:51 pass every live tet id */
    void calTetQualities(const std::vector<int>& t_ids, std::vector<TetQuality>& tet_qs) const {
        std::vector<double> e(t_ids.size());
        tet_qs.resize(t_ids.size());
        if (t_ids.empty()) return;
        ctx_.check(twg_mesh_quality(m_, t_ids.data(), t_ids.size(), e.data()));
        for (size_t i = 0; i < e.size(); ++i) tet_qs[i].slim_energy = e[i];
    }
    /* calTetQuality_AD (LocalOperations.cpp:783-860) of the listed tets */
    void calTetQuality_AD(const std::vector<int>& t_ids, std::vector<double>& min_d_angle, std::vector<double>& max_d_angle) const {
        min_d_angle.resize(t_ids.size()); max_d_angle.resize(t_ids.size());
        if (t_ids.empty()) return;
        ctx_.check(twg_mesh_dihedral(m_, t_ids.data(), t_ids.size(), min_d_angle.data(), max_d_angle.data()));
    }
    /* VertexSmoother::NewtonsUpdate (VertexSmoother.cpp:627-702) for the one-rings conn_tets[v] of many vertices */
    void NewtonsUpdate(const std::vector<int>& v_ids, std::vector<double>& energy, std::vector<double>& J3, std::vector<double>& H9,
                       std::vector<uint8_t>& ok) const {
        const size_t g = v_ids.size();
        energy.resize(g); J3.resize(3 * g); H9.resize(9 * g); ok.resize(g);
        if (g) ctx_.check(twg_mesh_vertex_ring_ejh(m_, v_ids.data(), g, energy.data(), J3.data(), H9.data(), ok.data()));
    }
    /* one vertex with the reference's own member list (t_ids = conn_tets[v_id]) */
    bool NewtonsUpdate(const std::vector<int>& t_ids, int v_id, double& energy, double* J, double* H) const {
        const uint64_t off[2] = {0, (uint64_t)t_ids.size()};
        uint8_t ok = 0;
        ctx_.check(twg_mesh_ring_ejh(m_, t_ids.data(), off, &v_id, 1, &energy, J, H, &ok));
        return ok != 0;
    }
    /* VertexSmoother::getNewEnergy (VertexSmoother.cpp:544-625) */
    double getNewEnergy(const std::vector<int>& t_ids) const {
        const uint64_t off[2] = {0, (uint64_t)t_ids.size()};
        double e = 0;
        ctx_.check(twg_mesh_ring_energy(m_, t_ids.data(), off, 1, &e));
        return e;
    }
    /* conn_tets as the device rebuilt it: ring of vertex v = tets[off[v] .. off[v+1]) in ascending tet id */
    void conn_tets(std::vector<uint64_t>& off, std::vector<int>& tets) const {
        off.resize(num_vertices() + 1);
        ctx_.check(twg_mesh_get_rings(m_, off.data(), nullptr));
        tets.resize((size_t)off.back());
        if (!tets.empty()) ctx_.check(twg_mesh_get_rings(m_, off.data(), tets.data()));
    }

private:
    TetMesh(const TetMesh&);
    TetMesh& operator=(const TetMesh&);
    static void pack(const std::array<int, 4>& t, bool removed, int32_t* out) {
        out[0] = removed ? -1 : t[0]; out[1] = t[1]; out[2] = t[2]; out[3] = t[3];
    }
    Context& ctx_;
    twg_mesh* m_;
};

/* ------------------------------------------------------------------------------------------------------------------
 * igl::winding_number(V, F, O, W) (called at InoutFiltering.cpp:45,66; MeshRefinement.cpp:614,1056).
 * MATD / MATI / VECD: Eigen-like (rows(), cols(), operator()(i,j), resize(n)); any storage order.
 * ---------------------------------------------------------------------------------------------------------------- */
namespace detail {
template <class MATD>
inline void pack_rows3(const MATD& M, std::vector<double>& out) {
    const size_t n = (size_t)M.rows();
    out.resize(3 * n);
    for (size_t i = 0; i < n; ++i)
        for (int c = 0; c < 3; ++c) out[3 * i + c] = M(i, c);
}
template <class MATI>
inline void pack_faces(const MATI& F, std::vector<uint32_t>& out) {
    const size_t n = (size_t)F.rows();
    out.resize(3 * n);
    for (size_t i = 0; i < n; ++i)
        for (int c = 0; c < 3; ++c) out[3 * i + c] = (uint32_t)F(i, c);
}
}  // namespace detail

template <class MATD, class MATI, class MATO, class VECD>
inline void winding_number(Context& ctx, const MATD& V, const MATI& F, const MATO& O, VECD& W) {
    std::vector<double> v, o;
    std::vector<uint32_t> f;
    detail::pack_rows3(V, v);
    detail::pack_faces(F, f);
    detail::pack_rows3(O, o);
    std::vector<double> w(o.size() / 3);
    ctx.check(twg_winding_number(ctx.handle(), v.data(), (uint32_t)(v.size() / 3), f.data(), (uint32_t)(f.size() / 3), o.data(), w.size(), w.data(), nullptr));
    W.resize(w.size());
    for (size_t i = 0; i < w.size(); ++i) W(i) = w[i];
}

/* InoutFiltering::filter (InoutFiltering.cpp:23-82) on packed arrays: centroids of the live tets are formed here the way
 * :26-39 does (CGAL::centroid of 4 points = arithmetic mean), then keep = W > 0.5 with the flip-and-retry of :56-75. */
struct InoutFiltering {
    /* V: tet vertex positions (posf), tets + t_is_removed as in the reference; SV/SF: the tracked surface of getSurface (:84-124) */
    static void filter(Context& ctx, const double* V, const std::vector<std::array<int, 4> >& tets, std::vector<bool>& t_is_removed, const double* SV,
                       uint32_t nSV, const uint32_t* SF, uint32_t nSF, bool* retried = nullptr) {
        std::vector<double> C;
        C.reserve(3 * tets.size());
        for (size_t i = 0; i < tets.size(); ++i) {
            if (t_is_removed[i]) continue;
            for (int c = 0; c < 3; ++c) {
                /* CGAL::centroid(4 points, Dimension_tag<0>): ((p0 + p1) + p2 + p3) / 4 in Cartesian<double> */
                const double s = V[3 * (size_t)tets[i][0] + c] + V[3 * (size_t)tets[i][1] + c] + V[3 * (size_t)tets[i][2] + c] + V[3 * (size_t)tets[i][3] + c];
                C.push_back(s / 4.0);
            }
        }
        const uint64_t nC = C.size() / 3;
        std::vector<uint8_t> keep(nC);
        int r = 0;
        if (nC) ctx.check(twg_inout_filter(ctx.handle(), SV, nSV, SF, nSF, C.data(), nC, keep.data(), &r));
        if (retried) *retried = r != 0;
        size_t cnt = 0;
        for (size_t i = 0; i < tets.size(); ++i) {
            if (t_is_removed[i]) continue;
            t_is_removed[i] = !keep[cnt++];
        }
    }
    /* MeshRefinement::markInOut (MeshRefinement.cpp:592-624) and the twin inside outputMidResult (:1036-1068): the same
     * centroids and W > 0.5 rule WITHOUT the flip-and-retry of filter(); writes a copy, t_is_removed itself is untouched */
    static void markInOut(Context& ctx, const double* V, const std::vector<std::array<int, 4> >& tets, const std::vector<bool>& t_is_removed,
                          std::vector<bool>& tmp_t_is_removed, const double* SV, uint32_t nSV, const uint32_t* SF, uint32_t nSF) {
        tmp_t_is_removed = t_is_removed;
        std::vector<double> C;
        C.reserve(3 * tets.size());
        for (size_t i = 0; i < tets.size(); ++i) {
            if (tmp_t_is_removed[i]) continue;
            for (int c = 0; c < 3; ++c) {
                const double s = V[3 * (size_t)tets[i][0] + c] + V[3 * (size_t)tets[i][1] + c] + V[3 * (size_t)tets[i][2] + c] + V[3 * (size_t)tets[i][3] + c];
                C.push_back(s / 4.0);
            }
        }
        const uint64_t nC = C.size() / 3;
        std::vector<uint8_t> keep(nC);
        if (nC) ctx.check(twg_winding_number(ctx.handle(), SV, nSV, SF, nSF, C.data(), nC, nullptr, keep.data()));
        size_t cnt = 0;
        for (size_t i = 0; i < tets.size(); ++i) {
            if (tmp_t_is_removed[i]) continue;
            tmp_t_is_removed[i] = !keep[cnt++];
        }
    }
};

}  // namespace twg

#endif /* TETWILD_GPU_HPP */
