// tools/wbuild_prof.cu -- host-only profile of the winding hierarchy build (no GPU needed): nvcc -I.. wbuild_prof.cu
#include "../tetwild_b200/csrc/winding.cu"
#include <cmath>
int main(int argc, char** argv) {
    const int nu = argc > 1 ? atoi(argv[1]) : 708, nv = nu;
    std::vector<double> V;
    std::vector<uint32_t> F;
    const double PI = 3.14159265358979323846;
    V.insert(V.end(), {0, 0, 0.5});
    for (int i = 1; i < nv; ++i)
        for (int j = 0; j < nu; ++j) {
            const double th = PI * i / nv, ph = 2 * PI * j / nu, r = 0.5 * (1 + 0.01 * sin(17 * th) * cos(13 * ph));
            V.insert(V.end(), {r * sin(th) * cos(ph), r * sin(th) * sin(ph), r * cos(th)});
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
    setenv("TWG_TRACE", "1", 1);
    HostTree T;
    const auto t0 = std::chrono::steady_clock::now();
    build_host_tree(V.data(), (uint32_t)(V.size() / 3), F.data(), (uint32_t)(F.size() / 3), 64, T);
    auto fnv = [](const void* p, size_t n) { const unsigned char* b = (const unsigned char*)p; unsigned long long h = 1469598103934665603ull; for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; } return h; };
    printf("checksums nodes %016llx caps %016llx tris %016llx\n", fnv(T.nodes.data(), T.nodes.size() * sizeof(WNode)), fnv(T.caps.data(), T.caps.size() * 8), fnv(T.tris.data(), T.tris.size() * 8));
    printf("root cap points %u, children cap points %u %u\n", T.nodes[1].cap_cnt, T.nodes.size() > 3 ? T.nodes[2].cap_cnt : 0u, T.nodes.size() > 3 ? T.nodes[3].cap_cnt : 0u);
    printf("facets %zu nodes %zu cap points %zu total %.1f ms\n", F.size() / 3, T.nodes.size(), T.caps.size() / 4,
           std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
    return 0;
}
