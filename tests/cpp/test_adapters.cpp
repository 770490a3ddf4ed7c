// tests/cpp/test_adapters.cpp -- GPU tier (built and run by tests/test_gpu_cpp_adapters.py).
// Drives the C++ host adapters of include/tetwild_gpu.hpp the way TetWild's own call sites drive the reference
// (MeshRefinement.cpp:209-226, EdgeCollapser.cpp:311-329,:727-775, VertexSmoother.cpp:627-702, InoutFiltering.cpp:23-82)
// with stand-ins for the reference's CGAL / Eigen / geogram types, and checks every answer against the CPU oracle
// (oracle/tw_oracle.h -- test infrastructure; the adapters themselves never see it).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "tetwild_gpu.hpp"
extern "C" {
#include "tw_oracle.h"
}

// ---- stand-ins with the accessors the reference's types offer ----
struct Point_3f {  // CGAL::Point_3<Epick>: operator[]
    double c[3];
    double operator[](int i) const { return c[i]; }
};
struct Triangle_3f {  // CGAL::Triangle_3<Epick>: operator[] -> vertex
    Point_3f v[3];
    const Point_3f& operator[](int i) const { return v[i]; }
};
struct MatrixXd {  // Eigen::MatrixXd (column-major like Eigen's default)
    std::vector<double> d;
    long r, cdim;
    MatrixXd(long rows = 0, long cols = 0) : d((size_t)(rows * cols)), r(rows), cdim(cols) {}
    long rows() const { return r; }
    long cols() const { return cdim; }
    double& operator()(long i, long j) { return d[(size_t)(j * r + i)]; }
    double operator()(long i, long j) const { return d[(size_t)(j * r + i)]; }
};
struct MatrixXi {
    std::vector<int> d;
    long r, cdim;
    MatrixXi(long rows = 0, long cols = 0) : d((size_t)(rows * cols)), r(rows), cdim(cols) {}
    long rows() const { return r; }
    int& operator()(long i, long j) { return d[(size_t)(j * r + i)]; }
    int operator()(long i, long j) const { return d[(size_t)(j * r + i)]; }
};
struct VectorXd {
    std::vector<double> d;
    void resize(size_t n) { d.resize(n); }
    double& operator()(size_t i) { return d[i]; }
    size_t size() const { return d.size(); }
};
struct GeoMesh {  // GEO::Mesh: vertices.nb()/point_ptr(), facets.nb()/vertex()
    struct Verts {
        std::vector<double> xyz;
        unsigned nb() const { return (unsigned)(xyz.size() / 3); }
        const double* point_ptr(unsigned v) const { return &xyz[3 * (size_t)v]; }
    } vertices;
    struct Facets {
        std::vector<unsigned> idx;
        unsigned nb() const { return (unsigned)(idx.size() / 3); }
        unsigned vertex(unsigned f, unsigned lv) const { return idx[3 * (size_t)f + lv]; }
    } facets;
};

static int g_fail = 0;
#define EXPECT(cond, ...)                                         \
    do {                                                          \
        if (!(cond)) {                                            \
            ++g_fail;                                             \
            std::fprintf(stderr, "FAIL %s:%d: ", __FILE__, __LINE__); \
            std::fprintf(stderr, __VA_ARGS__);                    \
            std::fprintf(stderr, "\n");                           \
        }                                                         \
    } while (0)

static void uv_sphere(int nu, int nv, double radius, GeoMesh& M) {
    const double PI = 3.14159265358979323846;
    M.vertices.xyz.clear();
    M.facets.idx.clear();
    auto push = [&](double x, double y, double z) { M.vertices.xyz.push_back(x); M.vertices.xyz.push_back(y); M.vertices.xyz.push_back(z); };
    push(0, 0, radius);
    for (int i = 1; i < nv; ++i)
        for (int j = 0; j < nu; ++j) {
            const double th = PI * i / nv, ph = 2 * PI * j / nu;
            push(radius * std::sin(th) * std::cos(ph), radius * std::sin(th) * std::sin(ph), radius * std::cos(th));
        }
    push(0, 0, -radius);
    const unsigned south = (unsigned)(M.vertices.nb() - 1);
    auto ring = [&](int i, int j) { return (unsigned)(1 + (i - 1) * nu + (j % nu)); };
    auto tri = [&](unsigned a, unsigned b, unsigned c) { M.facets.idx.push_back(a); M.facets.idx.push_back(b); M.facets.idx.push_back(c); };
    for (int j = 0; j < nu; ++j) tri(0, ring(1, j), ring(1, j + 1));
    for (int i = 1; i < nv - 1; ++i)
        for (int j = 0; j < nu; ++j) {
            tri(ring(i, j), ring(i + 1, j), ring(i + 1, j + 1));
            tri(ring(i, j), ring(i + 1, j + 1), ring(i, j + 1));
        }
    for (int j = 0; j < nu; ++j) tri(south, ring(nv - 1, j + 1), ring(nv - 1, j));
}

int main() {
    std::mt19937_64 rng(20240501);
    std::uniform_real_distribution<double> U(-1.0, 1.0);
    std::normal_distribution<double> N(0.0, 1.0);

    twg::Context ctx(0);
    const uint64_t l0 = ctx.launches();

    // ---------------- surface + the two trees of MeshRefinement.cpp:209-213 ----------------
    GeoMesh geo_sf_mesh, geo_b_mesh;
    uv_sphere(96, 64, 0.5, geo_sf_mesh);
    // boundary mesh: edges stored as degenerate triangles (Preprocess.cpp:192-197)
    geo_b_mesh.vertices.xyz = {0, 0, 0, 0.3, 0, 0, 0.3, 0.3, 0, 0, 0.3, 0.1};
    geo_b_mesh.facets.idx = {0, 1, 1, 1, 2, 2, 2, 3, 3, 3, 0, 0};
    twg::MeshFacetsAABBWithEps geo_sf_tree(ctx, geo_sf_mesh), geo_b_tree(ctx, geo_b_mesh);
    const twg::EnvelopeParams ep = twg::EnvelopeParams::from_args(std::sqrt(3.0), 1e-3);
    twg::LocalOperations lo(ctx, geo_sf_tree, geo_b_tree, ep.eps_2, ep.sampling_dist);
    ora_surface* osf = ora_surface_create(geo_sf_mesh.vertices.xyz.data(), geo_sf_mesh.vertices.nb(), geo_sf_mesh.facets.idx.data(), geo_sf_mesh.facets.nb(), 1);
    ora_surface* ob = ora_surface_create(geo_b_mesh.vertices.xyz.data(), geo_b_mesh.vertices.nb(), geo_b_mesh.facets.idx.data(), geo_b_mesh.facets.nb(), 1);

    // ---------------- voxel stuffing: DelaunayTetrahedralization::getVoxelPoints (DelaunayTetrahedralization.cpp:61-106) ----------------
    {
        const double p_min[3] = {-0.55, -0.6, -0.52}, p_max[3] = {0.55, 0.58, 0.61};
        const double diag = std::sqrt(1.1 * 1.1 + 1.18 * 1.18 + 1.13 * 1.13);
        const double vr = twg::voxel_resolution(/*relative edge length, %*/ 2.0, /*absolute*/ 0.02 * diag, diag);
        EXPECT(vr == diag / 20.0, "voxel resolution rule");
        std::vector<std::array<double, 3> > vox;
        twg::getVoxelPoints(p_min, p_max, geo_sf_tree, vr, vox);
        // the reference's loop, one tree query per grid point, against the oracle
        std::vector<double> ds[3];
        for (int i = 0; i < 3; ++i) {
            const double D = p_max[i] - p_min[i];
            const int Ni = (int)(D / vr) + 1;
            const double d = D / Ni;
            ds[i].push_back(p_min[i]);
            for (int j = 0; j < Ni - 1; ++j) ds[i].push_back(std::fma(d, (double)(j + 1), p_min[i]));
            ds[i].push_back(p_max[i]);
        }
        size_t k_out = 0, grid = 0, mism = 0;
        for (size_t i = 0; i < ds[0].size(); ++i)
            for (size_t j = 0; j < ds[1].size(); ++j)
                for (size_t k = 0; k < ds[2].size(); ++k) {
                    if ((i == 0 || i == ds[0].size() - 1) && (j == 0 || j == ds[1].size() - 1) && (k == 0 || k == ds[2].size() - 1)) continue;
                    ++grid;
                    const double q[3] = {ds[0][i], ds[1][j], ds[2][k]};
                    double od = 0;
                    ora_point_sqdist(osf, q, 1, &od, 1);
                    if (od < vr * vr / 4) continue;
                    if (k_out >= vox.size() || vox[k_out][0] != q[0] || vox[k_out][1] != q[1] || vox[k_out][2] != q[2]) ++mism;
                    ++k_out;
                }
        EXPECT(mism == 0 && k_out == vox.size(), "getVoxelPoints: %zu kept vs %zu by the reference loop, %zu differ", vox.size(), k_out, mism);
        EXPECT(k_out > grid / 2 && k_out < grid, "degenerate voxel test: %zu of %zu grid points kept", k_out, grid);
    }

    // isPointOutEnvelop / isPointOutBoundaryEnvelop, one point at a time like EdgeCollapser.cpp:311-329
    int n_out = 0;
    for (int i = 0; i < 300; ++i) {
        double d[3] = {N(rng), N(rng), N(rng)};
        const double l = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
        const double r = 0.5 + (i % 3 == 0 ? 0.0 : ep.eps * 1.5 * U(rng));
        Point_3f p = {{d[0] / l * r, d[1] / l * r, d[2] / l * r}};
        uint8_t o1 = 0, o2 = 0;
        double od = 0;
        ora_point_sqdist(osf, p.c, 1, &od, 1);
        o1 = od > ep.eps_2;
        EXPECT(lo.isPointOutEnvelop(p) == (o1 != 0), "isPointOutEnvelop differs at point %d", i);
        EXPECT(geo_sf_tree.squared_distance(p) == od, "squared_distance differs at point %d", i);
        ora_point_sqdist(ob, p.c, 1, &od, 1);
        o2 = od > ep.eps_2;
        EXPECT(lo.isPointOutBoundaryEnvelop(p) == (o2 != 0), "isPointOutBoundaryEnvelop differs at point %d", i);
        n_out += o1;
        twg::vec3 q(p[0], p[1], p[2]), nearest;
        double sq = 0;
        twg::index_t f = geo_sf_tree.nearest_facet(q, nearest, sq);
        EXPECT(f < geo_sf_tree.nb_facets(), "nearest_facet id out of range");
        const double dd = (q.x - nearest.x) * (q.x - nearest.x) + (q.y - nearest.y) * (q.y - nearest.y) + (q.z - nearest.z) * (q.z - nearest.z);
        EXPECT(std::fabs(dd - sq) <= 1e-6 * sq + 1e-18, "nearest point inconsistent with sq_dist");
    }
    EXPECT(n_out > 20 && n_out < 280, "degenerate test: %d of 300 points out", n_out);

    // isFaceOutEnvelop: single + batched, candidate faces near the surface (EdgeCollapser.cpp:727-775)
    std::vector<Triangle_3f> tris;
    for (int i = 0; i < 400; ++i) {
        Triangle_3f t;
        double c[3] = {N(rng), N(rng), N(rng)};
        const double l = std::sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);
        for (int k = 0; k < 3; ++k) {
            double v[3] = {c[0] / l + 0.03 * U(rng), c[1] / l + 0.03 * U(rng), c[2] / l + 0.03 * U(rng)};
            const double lv = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
            const double r = 0.5 - 0.0002 + (i % 2 ? ep.eps * 2.5 * U(rng) : 0.0);
            for (int a = 0; a < 3; ++a) t.v[k].c[a] = v[a] / lv * r;
        }
        if (i % 50 == 0) t.v[2] = t.v[1];  // degenerate -> IN (LocalOperations.cpp:1048)
        tris.push_back(t);
    }
    std::vector<uint8_t> batch;
    lo.isFaceOutEnvelop(tris, batch);
    int f_out = 0;
    for (size_t i = 0; i < tris.size(); ++i) {
        double t9[9];
        for (int k = 0; k < 3; ++k)
            for (int a = 0; a < 3; ++a) t9[3 * k + a] = tris[i][k][a];
        uint8_t o = 0;
        uint64_t ns = 0;
        ora_envelope_faces_out(osf, t9, 1, ep.sampling_dist, ep.eps_2, &o, &ns, 1);
        EXPECT(batch[i] == o, "batched isFaceOutEnvelop differs at face %zu", i);
        if (i < 40) EXPECT(lo.isFaceOutEnvelop(tris[i]) == (o != 0), "isFaceOutEnvelop differs at face %zu", i);
        f_out += o;
    }
    EXPECT(f_out > 10 && f_out < 390, "degenerate test: %d of 400 faces out", f_out);
    // Preprocess::isOutEnvelop (Preprocess.cpp:643-747): sets of faces, no degenerate shortcut, eps scaled by 0.8
    {
        const double eps2p = ep.eps_2 * 0.8 * 0.8;
        int n_true = 0;
        const int kSet = 2;
        int n_sets = 0;
        for (size_t b = 0; b + kSet <= tris.size(); b += kSet, ++n_sets) {
            double t9[kSet * 9];
            for (int f = 0; f < kSet; ++f)
                for (int k = 0; k < 3; ++k)
                    for (int a = 0; a < 3; ++a) t9[9 * f + 3 * k + a] = tris[b + f][k][a];
            uint8_t o[kSet];
            ora_envelope_faces_out_ex(osf, t9, kSet, ep.sampling_dist, eps2p, 0, o, nullptr, 1);
            bool any = false;
            for (int f = 0; f < kSet; ++f) any = any || o[f];
            EXPECT(geo_sf_tree.isOutEnvelop(t9, kSet, ep.sampling_dist, eps2p) == any, "Preprocess isOutEnvelop differs at set %zu", b / kSet);
            n_true += any;
        }
        EXPECT(n_true > 0 && n_true < n_sets, "degenerate test: %d of %d face sets out", n_true, n_sets);
    }

    // ---------------- AMIPS: calTetQualities, NewtonsUpdate, getNewEnergy, energy_ispc ----------------
    const int nV = 3000;
    std::vector<double> V(3 * nV);
    for (auto& x : V) x = U(rng);
    std::vector<std::array<int, 4> > tets;
    std::uniform_int_distribution<int> pick(0, nV - 1);
    for (int i = 0; i < 20000; ++i) {
        std::array<int, 4> t = {{pick(rng), pick(rng), pick(rng), pick(rng)}};
        tets.push_back(t);  // random orientation, some degenerate (repeated vertices): exercises the MAX_ENERGY gate
    }
    std::vector<twg::TetQuality> tet_qs;
    lo.calTetQualities(V.data(), nV, tets, tet_qs);
    std::vector<double> oe(tets.size());
    ora_amips_quality(V.data(), reinterpret_cast<const int32_t*>(tets.data()), tets.size(), oe.data(), 4);
    int n_max = 0;
    for (size_t i = 0; i < tets.size(); ++i) {
        const bool gmax = tet_qs[i].slim_energy == TWG_MAX_ENERGY, omax = oe[i] == TWG_MAX_ENERGY;
        EXPECT(gmax == omax, "MAX_ENERGY gate differs at tet %zu", i);
        if (!omax) EXPECT(std::fabs(tet_qs[i].slim_energy - oe[i]) <= 1e-9 * std::fabs(oe[i]), "slim_energy differs at tet %zu: %.17g vs %.17g", i, tet_qs[i].slim_energy, oe[i]);
        n_max += omax;
    }
    EXPECT(n_max > 5000 && n_max < 15000, "degenerate test: %d gated tets", n_max);

    // one-rings: a vertex and the tets that contain it (made positively oriented so the sums are meaningful)
    twg::VertexSmoother sm(ctx, V.data(), nV, tets);
    std::vector<int> t_ids, v_ids;
    std::vector<uint64_t> group_off(1, 0);
    for (int v = 0; v < 200; ++v) {
        int cnt = 0;
        for (size_t i = 0; i < tets.size() && cnt < 30; ++i)
            if (oe[i] != TWG_MAX_ENERGY && (tets[i][0] == v || tets[i][1] == v || tets[i][2] == v || tets[i][3] == v)) { t_ids.push_back((int)i); ++cnt; }
        if (cnt == 0) continue;
        v_ids.push_back(v);
        group_off.push_back(t_ids.size());
    }
    EXPECT(v_ids.size() > 100, "too few rings");
    std::vector<double> E, J3, H9, En;
    std::vector<uint8_t> ok;
    sm.NewtonsUpdate(t_ids, group_off, v_ids, E, J3, H9, ok);
    sm.getNewEnergy(t_ids, group_off, En);
    const size_t G = v_ids.size();
    std::vector<double> oE(G), oJ(3 * G), oH(9 * G), oEn(G);
    std::vector<uint8_t> ook(G);
    ora_amips_ring_ejh(V.data(), reinterpret_cast<const int32_t*>(tets.data()), t_ids.data(), group_off.data(), v_ids.data(), G, oE.data(), oJ.data(), oH.data(), ook.data(), 4);
    ora_amips_ring_energy(V.data(), reinterpret_cast<const int32_t*>(tets.data()), t_ids.data(), group_off.data(), G, oEn.data(), 4);
    for (size_t g = 0; g < G; ++g) {
        EXPECT(ok[g] == ook[g], "NewtonsUpdate ok flag differs at ring %zu", g);
        EXPECT(std::fabs(E[g] - oE[g]) <= 1e-9 * std::fabs(oE[g]), "ring E differs at %zu", g);
        EXPECT(std::fabs(En[g] - oEn[g]) <= 1e-9 * std::fabs(oEn[g]), "getNewEnergy differs at %zu", g);
        double jmax = 0, hmax = 0;
        for (int k = 0; k < 3; ++k) jmax = std::fmax(jmax, std::fabs(oJ[3 * g + k]));
        for (int k = 0; k < 9; ++k) hmax = std::fmax(hmax, std::fabs(oH[9 * g + k]));
        for (int k = 0; k < 3; ++k) EXPECT(std::fabs(J3[3 * g + k] - oJ[3 * g + k]) <= 1e-9 * jmax, "ring J differs at %zu", g);
        for (int k = 0; k < 9; ++k) EXPECT(std::fabs(H9[9 * g + k] - oH[9 * g + k]) <= 1e-9 * hmax, "ring H differs at %zu", g);
        if (g < 10) {  // the reference's one-vertex signature
            std::vector<int> ring(t_ids.begin() + group_off[g], t_ids.begin() + group_off[g + 1]);
            double e1, j1[3], h1[9];
            const bool good = sm.NewtonsUpdate(ring, v_ids[g], e1, j1, h1);
            EXPECT(good == (ook[g] != 0) && e1 == E[g] && j1[0] == J3[3 * g] && h1[8] == H9[9 * g + 8], "single-ring NewtonsUpdate differs from the batched call at %zu", g);
            EXPECT(sm.getNewEnergy(ring) == En[g], "single-ring getNewEnergy differs at %zu", g);
        }
    }
    // ---------------- resident tet mesh: the same answers with only ids crossing the bus ----------------
    {
        std::vector<bool> removed(tets.size(), false);
        for (size_t i = 0; i < tets.size(); i += 13) removed[i] = true;
        auto posf = [&](int v) { return &V[3 * (size_t)v]; };
        twg::TetMesh mesh(ctx, nV, posf, tets, removed);
        EXPECT(mesh.num_vertices() == (size_t)nV && mesh.num_tets() == tets.size(), "TetMesh sizes");
        std::vector<int> live;
        for (size_t i = 0; i < tets.size(); ++i)
            if (!removed[i]) live.push_back((int)i);
        std::vector<twg::TetQuality> q2;
        mesh.calTetQualities(live, q2);
        for (size_t k = 0; k < live.size(); ++k) EXPECT(q2[k].slim_energy == tet_qs[live[k]].slim_energy, "resident calTetQualities differs at tet %d", live[k]);
        std::vector<double> amin, amax, omin(tets.size()), omax(tets.size());
        mesh.calTetQuality_AD(live, amin, amax);
        ora_tet_dihedral(V.data(), reinterpret_cast<const int32_t*>(tets.data()), tets.size(), omin.data(), omax.data(), 4);
        for (size_t k = 0; k < live.size(); ++k)
            EXPECT(std::fabs(amin[k] - omin[live[k]]) < 1e-13 && std::fabs(amax[k] - omax[live[k]]) < 1e-13, "calTetQuality_AD differs at tet %d", live[k]);
        // conn_tets rebuilt on the device == the scheduler's own bookkeeping
        std::vector<uint64_t> coff;
        std::vector<int> ctets;
        mesh.conn_tets(coff, ctets);
        std::vector<std::vector<int> > conn(nV);
        for (size_t i = 0; i < tets.size(); ++i)
            if (!removed[i])
                for (int k = 0; k < 4; ++k) conn[tets[i][k]].push_back((int)i);
        size_t bad = 0;
        for (int v = 0; v < nV; ++v) {
            if (coff[v + 1] - coff[v] != conn[v].size()) { ++bad; continue; }
            for (size_t k = 0; k < conn[v].size(); ++k) bad += ctets[coff[v] + k] != conn[v][k];
        }
        EXPECT(bad == 0, "conn_tets differs at %zu places", bad);
        // NewtonsUpdate of a vertex through the resident mesh == through the ship-everything adapter (same kernel)
        for (int v = 0; v < 20; ++v) {
            if (conn[v].empty()) continue;
            bool dup = false;  // a random tet may name v twice; the reference never has such tets
            for (int t : conn[v]) { int c = 0; for (int k = 0; k < 4; ++k) c += tets[t][k] == v; dup = dup || c > 1; }
            if (dup) continue;
            double e1, j1[3], h1[9], e2, j2[3], h2[9];
            const bool g1 = sm.NewtonsUpdate(conn[v], v, e1, j1, h1), g2 = mesh.NewtonsUpdate(conn[v], v, e2, j2, h2);
            EXPECT(g1 == g2 && (e1 == e2 || (e1 != e1 && e2 != e2)), "resident NewtonsUpdate differs at vertex %d", v);
            if (g1) EXPECT(j1[1] == j2[1] && h1[5] == h2[5], "resident NewtonsUpdate J/H differ at vertex %d", v);
            const double n1 = sm.getNewEnergy(conn[v]), n2 = mesh.getNewEnergy(conn[v]);
            EXPECT(n1 == n2, "resident getNewEnergy differs at vertex %d", v);
        }
        // an accepted smoothing step moves a vertex: sync and re-evaluate
        V[3 * 5] += 0.01; V[3 * 5 + 2] -= 0.02;
        mesh.sync_vertices(std::vector<int>(1, 5), posf);
        std::vector<twg::TetQuality> qa, qb;
        std::vector<int> ring5;
        for (int t : conn[5]) ring5.push_back(t);
        mesh.calTetQualities(ring5, qa);
        std::vector<std::array<int, 4> > sub;
        for (int t : ring5) sub.push_back(tets[t]);
        lo.calTetQualities(V.data(), nV, sub, qb);
        for (size_t k = 0; k < ring5.size(); ++k) EXPECT(qa[k].slim_energy == qb[k].slim_energy, "after sync_vertices: tet %d differs", ring5[k]);
        V[3 * 5] -= 0.01; V[3 * 5 + 2] += 0.02;
    }
    // energy_ispc argument list (LocalOperations.cpp:750)
    {
        const int n = 5000;
        std::vector<double> T[12];
        for (int k = 0; k < 12; ++k) T[k].resize(n);
        for (int i = 0; i < n; ++i)
            for (int k = 0; k < 12; ++k) T[k][i] = V[3 * (size_t)tets[i][k / 3] + k % 3];
        std::vector<double> Eg(n), Eo(n);
        twg::energy_ispc(ctx, T[0].data(), T[1].data(), T[2].data(), T[3].data(), T[4].data(), T[5].data(), T[6].data(), T[7].data(), T[8].data(),
                         T[9].data(), T[10].data(), T[11].data(), Eg.data(), n);
        const double* Tp[12];
        for (int k = 0; k < 12; ++k) Tp[k] = T[k].data();
        ora_amips_energy_soa(Tp, Eo.data(), n, 4);
        for (int i = 0; i < n; ++i)
            if (std::isfinite(Eo[i]) && oe[i] != TWG_MAX_ENERGY) EXPECT(std::fabs(Eg[i] - Eo[i]) <= 1e-9 * std::fabs(Eo[i]), "energy_ispc differs at %d", i);
        double t12[12], j[3], h[9], oj[3], oh[9];
        for (int k = 0; k < 12; ++k) t12[k] = T[k][7];
        const double e = twg::LocalOperations::comformalAMIPSEnergy_new(ctx, t12);
        twg::LocalOperations::comformalAMIPSJacobian_new(ctx, t12, j);
        twg::LocalOperations::comformalAMIPSHessian_new(ctx, t12, h);
        ora_amips_jacobian(t12, oj);
        ora_amips_hessian(t12, oh);
        EXPECT(std::fabs(e - ora_amips_energy(t12)) <= 1e-9 * std::fabs(e), "comformalAMIPSEnergy_new differs");
        for (int k = 0; k < 3; ++k) EXPECT(std::fabs(j[k] - oj[k]) <= 1e-9 * (std::fabs(oj[0]) + std::fabs(oj[1]) + std::fabs(oj[2])), "Jacobian differs");
        for (int k = 0; k < 9; ++k) EXPECT(std::fabs(h[k] - oh[k]) <= 1e-9 * (std::fabs(oh[0]) + std::fabs(oh[4]) + std::fabs(oh[8])), "Hessian differs");
    }

    // ---------------- winding: igl::winding_number signature + InoutFiltering::filter ----------------
    {
        const long nSV = geo_sf_mesh.vertices.nb(), nSF = geo_sf_mesh.facets.nb();
        MatrixXd SV(nSV, 3), O(4000, 3);
        MatrixXi SF(nSF, 3);
        for (long i = 0; i < nSV; ++i)
            for (int c = 0; c < 3; ++c) SV(i, c) = geo_sf_mesh.vertices.xyz[3 * (size_t)i + c];
        for (long i = 0; i < nSF; ++i)
            for (int c = 0; c < 3; ++c) SF(i, c) = (int)geo_sf_mesh.facets.idx[3 * (size_t)i + c];
        std::vector<double> Oc(3 * 4000);
        for (long i = 0; i < 4000; ++i)
            for (int c = 0; c < 3; ++c) { O(i, c) = 0.6 * U(rng); Oc[3 * (size_t)i + c] = O(i, c); }
        VectorXd W;
        twg::winding_number(ctx, SV, SF, O, W);
        std::vector<double> Wo(4000);
        ora_winding_direct(geo_sf_mesh.vertices.xyz.data(), (uint32_t)nSV, geo_sf_mesh.facets.idx.data(), (uint32_t)nSF, Oc.data(), 4000, Wo.data(), 4);
        int inside = 0;
        for (size_t i = 0; i < 4000; ++i) {
            EXPECT(std::fabs(W(i) - Wo[i]) < 1e-10, "W differs at %zu: %.17g vs %.17g", i, W(i), Wo[i]);
            EXPECT((W(i) > 0.5) == (Wo[i] > 0.5), "inside/outside decision differs at %zu", i);
            inside += Wo[i] > 0.5;
        }
        EXPECT(inside > 500 && inside < 3500, "degenerate test: %d inside", inside);

        // filter(): tets around random centres; the reversed surface must trigger the retry (InoutFiltering.cpp:56-75)
        std::vector<double> TV;
        std::vector<std::array<int, 4> > tt;
        for (int i = 0; i < 3000; ++i) {
            const double c[3] = {0.6 * U(rng), 0.6 * U(rng), 0.6 * U(rng)};
            std::array<int, 4> t;
            for (int k = 0; k < 4; ++k) {
                t[k] = (int)(TV.size() / 3);
                for (int a = 0; a < 3; ++a) TV.push_back(c[a] + 0.01 * U(rng));
            }
            tt.push_back(t);
        }
        std::vector<double> cen(3 * tt.size());
        for (size_t i = 0; i < tt.size(); ++i)
            for (int a = 0; a < 3; ++a) cen[3 * i + a] = (TV[3 * (size_t)tt[i][0] + a] + TV[3 * (size_t)tt[i][1] + a] + TV[3 * (size_t)tt[i][2] + a] + TV[3 * (size_t)tt[i][3] + a]) / 4.0;
        for (int pass = 0; pass < 2; ++pass) {
            std::vector<unsigned> Fq = geo_sf_mesh.facets.idx;
            if (pass == 1)
                for (size_t f = 0; f < Fq.size() / 3; ++f) std::swap(Fq[3 * f + 1], Fq[3 * f + 2]);
            std::vector<bool> removed(tt.size(), false);
            removed[5] = removed[17] = true;  // already-removed tets stay removed and are not queried (:28-29)
            bool retried = false;
            twg::InoutFiltering::filter(ctx, TV.data(), tt, removed, geo_sf_mesh.vertices.xyz.data(), (uint32_t)nSV, Fq.data(), (uint32_t)nSF, &retried);
            std::vector<uint8_t> okeep(tt.size());
            const int oretry = ora_inout_filter(geo_sf_mesh.vertices.xyz.data(), (uint32_t)nSV, Fq.data(), (uint32_t)nSF, cen.data(), tt.size(), okeep.data(), nullptr, 1, 4);
            EXPECT(retried == (pass == 1) && (oretry != 0) == retried, "flip-and-retry: pass %d retried=%d oracle=%d", pass, (int)retried, oretry);
            for (size_t i = 0; i < tt.size(); ++i) {
                const bool expect_removed = (i == 5 || i == 17) ? true : !okeep[i];
                EXPECT(removed[i] == expect_removed, "filter decision differs at tet %zu (pass %d)", i, pass);
            }
            // MeshRefinement::markInOut (:592-624): same rule, no flip-and-retry, result in a copy
            std::vector<bool> before(tt.size(), false), marked;
            before[5] = before[17] = true;
            twg::InoutFiltering::markInOut(ctx, TV.data(), tt, before, marked, geo_sf_mesh.vertices.xyz.data(), (uint32_t)nSV, Fq.data(), (uint32_t)nSF);
            std::vector<double> Wd(tt.size());
            ora_winding_direct(geo_sf_mesh.vertices.xyz.data(), (uint32_t)nSV, Fq.data(), (uint32_t)nSF, cen.data(), tt.size(), Wd.data(), 4);
            EXPECT(before[5] && before[17] && !before[0], "markInOut must not touch t_is_removed");
            for (size_t i = 0; i < tt.size(); ++i) {
                const bool expect_removed = (i == 5 || i == 17) ? true : !(Wd[i] > 0.5);
                EXPECT(marked[i] == expect_removed, "markInOut decision differs at tet %zu (pass %d)", i, pass);
            }
        }
    }

    ora_surface_destroy(osf);
    ora_surface_destroy(ob);
    std::printf("%s: %d failures, %llu kernel launches\n", g_fail ? "FAILED" : "ok", g_fail, (unsigned long long)(ctx.launches() - l0));
    return g_fail ? 1 : 0;
}
