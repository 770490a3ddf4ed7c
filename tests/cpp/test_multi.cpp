// tests/cpp/test_multi.cpp -- GPU tier (built and run by tests/test_gpu_multi.py).
// The caller of SURVEY.md 8(e) is a single C++ process (TetWild): InoutFiltering::filter hands ALL tet centroids to one
// winding call (InoutFiltering.cpp:40-52), MeshRefinement builds ONE envelope structure (MeshRefinement.cpp:209-226). This
// driver opens twg::Context over several devices (twg_create_multi) and checks that every batch call through the C++
// adapters gives bit-identical results to the one-device context, and prints the wall time of both.
//   usage: test_multi <n_contexts> <real_devices: 1 = device ids 0..n-1, 0 = n worker contexts on device 0>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include "tetwild_gpu.hpp"

static int failures = 0;
#define EXPECT(cond, what)                                      \
    do {                                                        \
        if (!(cond)) { ++failures; printf("FAIL: %s\n", what); } \
    } while (0)

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

static void uv_sphere(int nu, int nv, std::vector<double>& V, std::vector<uint32_t>& F) {
    const double PI = 3.14159265358979323846;
    V.insert(V.end(), {0, 0, 0.5});
    for (int i = 1; i < nv; ++i)
        for (int j = 0; j < nu; ++j) {
            const double th = PI * i / nv, ph = 2 * PI * j / nu, r = 0.5 * (1 + 0.01 * std::sin(17 * th) * std::cos(13 * ph));
            V.insert(V.end(), {r * std::sin(th) * std::cos(ph), r * std::sin(th) * std::sin(ph), r * std::cos(th)});
        }
    V.insert(V.end(), {0, 0, -0.5});
    const uint32_t south = (uint32_t)(V.size() / 3 - 1);
    auto ring = [&](int i, int j) { return (uint32_t)(1 + (i - 1) * nu + (j % nu)); };
    for (int j = 0; j < nu; ++j) F.insert(F.end(), {0u, ring(1, j), ring(1, j + 1)});
    for (int i = 1; i < nv - 1; ++i)
        for (int j = 0; j < nu; ++j) {
            F.insert(F.end(), {ring(i, j), ring(i + 1, j), ring(i + 1, j + 1)});
            F.insert(F.end(), {ring(i, j), ring(i + 1, j + 1), ring(i, j + 1)});
        }
    for (int j = 0; j < nu; ++j) F.insert(F.end(), {south, ring(nv - 1, j + 1), ring(nv - 1, j)});
}

int main(int argc, char** argv) {
    const int n_ctx = argc > 1 ? atoi(argv[1]) : 2;
    const bool real = argc > 2 ? atoi(argv[2]) != 0 : false;
    std::vector<int> ids;
    for (int k = 0; k < n_ctx; ++k) ids.push_back(real ? k : 0);
    try {
        twg::Context one(0), many(ids);
        EXPECT(many.num_devices() == n_ctx && one.num_devices() == 1, "device counts");
        std::vector<double> V;
        std::vector<uint32_t> F;
        uv_sphere(300, 300, V, F);
        const uint32_t nV = (uint32_t)(V.size() / 3), nF = (uint32_t)(F.size() / 3);
        std::mt19937_64 rng(5);
        std::uniform_real_distribution<double> U(-0.6, 0.6), S(0.49, 0.51);
        std::normal_distribution<double> N(0.0, 1.0);
        // ---- envelope: points in a thin shell around the sphere
        const uint64_t n = 2000003;
        std::vector<double> P(3 * n);
        for (uint64_t i = 0; i < n; ++i) {
            double x = N(rng), y = N(rng), z = N(rng), l = std::sqrt(x * x + y * y + z * z), r = S(rng);
            P[3 * i] = r * x / l; P[3 * i + 1] = r * y / l; P[3 * i + 2] = r * z / l;
        }
        const twg::EnvelopeParams ep = twg::EnvelopeParams::from_args(1.0, 4e-3);
        twg::MeshFacetsAABBWithEps t1(one, V.data(), nV, F.data(), nF), tm(many, V.data(), nV, F.data(), nF);
        std::vector<uint8_t> o1(n), om(n);
        t1.points_out_of_envelope(P.data(), n, ep.eps_2, o1.data());
        tm.points_out_of_envelope(P.data(), n, ep.eps_2, om.data());   // warm-up of both paths (scratch allocation)
        double a = now();
        t1.points_out_of_envelope(P.data(), n, ep.eps_2, o1.data());
        double b = now();
        tm.points_out_of_envelope(P.data(), n, ep.eps_2, om.data());
        double c = now();
        EXPECT(o1 == om, "envelope decisions: multi == single");
        uint64_t outs = 0;
        for (uint64_t i = 0; i < n; ++i) outs += o1[i];
        EXPECT(outs > n / 20 && outs < n - n / 20, "envelope decisions are a mix");
        printf("envelope  %llu points: 1 device %.2f ms, %d contexts %.2f ms\n", (unsigned long long)n, (b - a) * 1e3, n_ctx, (c - b) * 1e3);
        std::vector<double> d1(n), dm(n);
        t1.squared_distances(P.data(), n, d1.data());
        tm.squared_distances(P.data(), n, dm.data());
        EXPECT(memcmp(d1.data(), dm.data(), n * 8) == 0, "squared distances: multi == single");
        // ---- winding: InoutFiltering-shaped call
        const uint64_t nq = 3000001;
        std::vector<double> C(3 * nq);
        for (auto& x : C) x = U(rng);
        std::vector<uint8_t> k1(nq), km(nq);
        int r1 = -1, rm = -1;
        one.check(twg_inout_filter(one.handle(), V.data(), nV, F.data(), nF, C.data(), nq, k1.data(), &r1));
        a = now();
        one.check(twg_inout_filter(one.handle(), V.data(), nV, F.data(), nF, C.data(), nq, k1.data(), &r1));
        b = now();
        many.check(twg_inout_filter(many.handle(), V.data(), nV, F.data(), nF, C.data(), nq, km.data(), &rm));
        c = now();
        EXPECT(k1 == km && r1 == 0 && rm == 0, "inout filter: multi == single");
        uint64_t kept = 0;
        for (uint64_t i = 0; i < nq; ++i) kept += k1[i];
        EXPECT(std::fabs((double)kept / nq - 0.30) < 0.05, "about 30 % of the box is inside the sphere");
        printf("inout     %llu centroids (one-shot, build included): 1 device %.1f ms, %d contexts %.1f ms\n", (unsigned long long)nq, (b - a) * 1e3, n_ctx, (c - b) * 1e3);
        // ---- flat AMIPS batch (energy_ispc argument list)
        const uint64_t nt = 1500001;
        std::vector<std::vector<double> > T(12, std::vector<double>(nt));
        const double base[12] = {0, 0, 0, 1, 0, 0, 0.5, std::sqrt(3.0) / 2, 0, 0.5, std::sqrt(3.0) / 6, std::sqrt(6.0) / 3};
        for (uint64_t i = 0; i < nt; ++i)
            for (int k = 0; k < 12; ++k) T[k][i] = base[k] + 0.1 * N(rng);
        const double* Tp[12];
        for (int k = 0; k < 12; ++k) Tp[k] = T[k].data();
        std::vector<double> E1(nt), Em(nt), H1(9 * nt), Hm(9 * nt), J1(3 * nt), Jm(3 * nt);
        one.check(twg_amips_ejh_soa(one.handle(), Tp, E1.data(), J1.data(), H1.data(), nt));
        many.check(twg_amips_ejh_soa(many.handle(), Tp, Em.data(), Jm.data(), Hm.data(), nt));
        EXPECT(memcmp(E1.data(), Em.data(), nt * 8) == 0 && memcmp(J1.data(), Jm.data(), nt * 24) == 0 && memcmp(H1.data(), Hm.data(), nt * 72) == 0,
               "AMIPS E/J/H: multi == single");
        // ---- errors surface through the multi context with the device named
        std::vector<uint32_t> Fbad(F.begin(), F.begin() + 30);
        Fbad[7] = nV + 5;
        twg_surface* sb = nullptr;
        const int rc = twg_surface_create(many.handle(), V.data(), nV, Fbad.data(), 10, &sb);
        EXPECT(rc != 0 && sb == nullptr && strstr(twg_last_error(many.handle()), "out of range") != nullptr, "bad facet index refused through the multi context");
    } catch (const twg::Error& e) {
        printf("twg::Error %d: %s\n", e.code, e.what());
        return 2;
    }
    printf("ok: %d failures\n", failures);
    return failures ? 1 : 0;
}
