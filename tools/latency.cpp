// tools/latency.cpp -- host-side latency of SMALL calls through the C ABI (what an un-batched call site of TetWild would see).
// Built and driven by scripts/latency.py, which writes the test geometry to a binary file. Not part of the library.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "tetwild_gpu.h"

template <class T>
static std::vector<T> rd(FILE* f) {
    unsigned long long n = 0;
    if (fread(&n, 8, 1, f) != 1) { fprintf(stderr, "short file\n"); exit(2); }
    std::vector<T> v(n);
    if (n && fread(v.data(), sizeof(T), n, f) != n) { fprintf(stderr, "short file\n"); exit(2); }
    return v;
}
template <class F>
static double time_us(F fn, int reps) {
    for (int i = 0; i < 5; ++i) fn();
    const auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < reps; ++i) fn();
    return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() / reps;
}
#define CK(x) do { int rc_ = (x); if (rc_) { fprintf(stderr, "%s failed: %d %s\n", #x, rc_, twg_last_error(ctx)); return 1; } } while (0)

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    FILE* f = fopen(argv[1], "rb");
    if (!f) return 2;
    std::vector<double> V = rd<double>(f);        // surface vertices
    std::vector<unsigned> F = rd<unsigned>(f);    // surface facets
    std::vector<double> T = rd<double>(f);        // candidate faces (9 doubles each)
    std::vector<double> P = rd<double>(f);        // query points
    std::vector<double> MV = rd<double>(f);       // tet mesh vertices
    std::vector<int> MT = rd<int>(f);             // tets
    std::vector<double> par = rd<double>(f);      // sampling_dist, eps2
    fclose(f);
    twg_ctx* ctx = nullptr;
    if (twg_create(&ctx, 0)) { fprintf(stderr, "twg_create failed (no CPU fallback)\n"); return 1; }
    twg_surface* s = nullptr;
    CK(twg_surface_create(ctx, V.data(), (unsigned)(V.size() / 3), F.data(), (unsigned)(F.size() / 3), &s));
    twg_mesh* m = nullptr;
    CK(twg_mesh_create(ctx, MV.data(), (unsigned)(MV.size() / 3), MT.data(), MT.size() / 4, &m));
    CK(twg_mesh_build_rings(m));
    const double sd = par[0], eps2 = par[1];
    std::vector<unsigned char> out(1 << 16);
    printf("{\"faces_out_us\": {");
    const int nf[] = {1, 4, 16, 64, 256, 1024, 4096};
    for (int k = 0; k < 7; ++k) {
        const int n = nf[k];
        const double us = time_us([&] { twg_envelope_faces_out(s, T.data(), n, sd, eps2, out.data()); }, n <= 256 ? 300 : 60);
        printf("%s\"%d\": %.1f", k ? ", " : "", n, us);
    }
    printf("}, \"points_out_us\": {");
    const int np_[] = {1, 32, 256, 2048, 16384};
    for (int k = 0; k < 5; ++k) {
        const int n = np_[k];
        const double us = time_us([&] { twg_envelope_points_out(s, P.data(), n, eps2, out.data()); }, 300);
        printf("%s\"%d\": %.1f", k ? ", " : "", n, us);
    }
    printf("}, \"mesh_vertex_ring_ejh_us\": {");
    std::vector<int> ids(4096);
    for (int i = 0; i < 4096; ++i) ids[i] = (int)((i * 7919u) % (MV.size() / 3));
    std::vector<double> E(4096), J(3 * 4096), H(9 * 4096);
    const int nr[] = {1, 16, 128, 1024, 4096};
    for (int k = 0; k < 5; ++k) {
        const int n = nr[k];
        const double us = time_us([&] { twg_mesh_vertex_ring_ejh(m, ids.data(), n, E.data(), J.data(), H.data(), out.data()); }, 300);
        printf("%s\"%d\": %.1f", k ? ", " : "", n, us);
    }
    printf("}, \"mesh_quality_us\": {");
    std::vector<int> tids(4096);
    for (int i = 0; i < 4096; ++i) tids[i] = (int)((i * 104729u) % (MT.size() / 4));
    for (int k = 0; k < 5; ++k) {
        const int n = nr[k];
        const double us = time_us([&] { twg_mesh_quality(m, tids.data(), n, E.data()); }, 300);
        printf("%s\"%d\": %.1f", k ? ", " : "", n, us);
    }
    printf("}, \"mesh_vertex_trial_energy_us\": {");
    std::vector<double> X(3 * 32);
    for (int i = 0; i < 32; ++i)
        for (int a = 0; a < 3; ++a) X[3 * i + a] = MV[3 * (size_t)ids[i] + a] + 1e-4 * (a + 1);
    const int nt[] = {1, 4, 16, 32};
    for (int k = 0; k < 4; ++k) {
        const int n = nt[k];
        const double us = time_us([&] { twg_mesh_vertex_trial_energy(m, ids.data(), X.data(), n, E.data()); }, 300);
        printf("%s\"%d\": %.1f", k ? ", " : "", n, us);
    }
    printf("}, \"nearest_us\": {");
    std::vector<unsigned> fac(64);
    std::vector<double> npt(3 * 64), nd2(64);
    const int nn[] = {1, 8, 64};
    for (int k = 0; k < 3; ++k) {
        const int n = nn[k];
        const double us = time_us([&] { twg_nearest(s, P.data(), n, fac.data(), npt.data(), nd2.data()); }, 300);
        printf("%s\"%d\": %.1f", k ? ", " : "", n, us);
    }
    double floor_us = 0.0;
    CK(twg_debug_roundtrip(ctx, 2000, &floor_us));
    printf("}, \"empty_kernel_roundtrip_us\": %.2f}\n", floor_us);
    twg_mesh_destroy(m);
    twg_surface_destroy(s);
    twg_destroy(ctx);
    return 0;
}
