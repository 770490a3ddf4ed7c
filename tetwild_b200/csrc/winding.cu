// winding.cu -- generalized winding number of query points against a triangle surface, and the W > 0.5 filter.
//
// Replaces igl::winding_number(V,F,O,W) as called by InoutFiltering::filter (src/tetwild/InoutFiltering.cpp:45-75)
// and MeshRefinement::markInOut / outputMidResult (src/tetwild/MeshRefinement.cpp:592-624, :1036-1068).
//
// Algorithm: the exact hierarchical evaluation of Jacobson et al. 2013 (the one libigl runs), laid out for the GPU.
//   * facets are Morton-sorted and cut into leaf blocks of kLeaf triangles; an implicit binary heap of nodes covers
//     contiguous block ranges. Each node stores its bounding box and its CAP: the exterior (unmatched) directed
//     edges of its sub-mesh, fanned to one apex vertex. For a query outside the node's box the solid angle of the
//     sub-mesh equals that of the cap exactly (they share their boundary and the closed difference lies inside the
//     convex box), so far sub-meshes cost O(sqrt(#facets)) instead of O(#facets).
//   * queries are Morton-sorted on the device; a warp owns 32 consecutive (hence spatially coherent) queries and
//     traverses the heap with ONE shared stack: a node is opened iff some lane lies inside its box, otherwise all
//     lanes add its cap. Caps and leaf triangles are streamed, tile by tile, into a per-warp shared-memory ring with
//     1-D bulk async copies (TMA, cp.async.bulk + mbarrier, double-buffered) and consumed by all 32 lanes with
//     conflict-free broadcast reads.
//   * per (query, triangle): Van Oosterom-Strackee solid angle, atan2(det, |a||b||c| + (a.b)|c| + (b.c)|a| + (c.a)|b|).
// FP64-pipe-bound by design (no tensor-core formulation exists for atan2/sqrt chains).
#include <algorithm>
#include <cub/device/device_radix_sort.cuh>
#include "common.cuh"

namespace {

constexpr uint32_t kLeaf = 64;        // triangles per leaf block
constexpr int kWThreads = 256;        // 8 warps per CTA
constexpr int kWarps = kWThreads / 32;
constexpr int kTileSeg = 32;          // cap segments per staged tile (32 * 48 B = 1536 B)
constexpr int kTileTri = 32;          // triangles per staged tile    (32 * 72 B = 2304 B)
constexpr int kStageBytes = 2304;
constexpr int kStackDepth = 64;

struct __align__(16) WNode {
    float lo[3], hi[3];      // bounding box, rounded outward
    uint32_t cap_off, cap_cnt;
    double apex[3];
    uint32_t tri_off, tri_cnt;  // facets of the whole subtree (contiguous in sorted order)
};
static_assert(sizeof(WNode) == 64, "WNode is 64 bytes");

struct WView {
    const WNode* nodes;    // heap, index 1 .. 2*nBlkP-1
    const double* caps;    // 6 doubles per segment (A, B)
    const double* tris;    // 9 doubles per facet, sorted
    uint32_t nBlkP;        // leaf blocks, power of two
    uint32_t nF;
};

__device__ __forceinline__ double solid_angle_2pi(double ax, double ay, double az, double la, double bx, double by, double bz, double lb,
                                                  double cx, double cy, double cz, double lc) {
    const double det = ax * (by * cz - bz * cy) + bx * (cy * az - cz * ay) + cx * (ay * bz - az * by);
    const double ab = ax * bx + ay * by + az * bz;
    const double bc = bx * cx + by * cy + bz * cz;
    const double ca = cx * ax + cy * ay + cz * az;
    return atan2(det, la * lb * lc + bc * la + ca * lb + ab * lc);
}

template <bool USE_TMA>
struct Stream {  // per-warp tile streamer
    unsigned char* buf;  // 2 stages of kStageBytes
    uint64_t* bars;      // 2 mbarriers
    uint32_t phase[2];
    int lane;
};

// sum over fan triangles (apex, A_k, B_k), k in [0,cnt)
template <bool USE_TMA>
__device__ __forceinline__ double eval_cap(const WView& W, const WNode& nd, double px, double py, double pz, Stream<USE_TMA>& st) {
    const double ox = nd.apex[0] - px, oy = nd.apex[1] - py, oz = nd.apex[2] - pz;
    const double lo = sqrt(ox * ox + oy * oy + oz * oz);
    double acc = 0.0;
    const double* src = W.caps + (size_t)nd.cap_off * 6;
    const uint32_t cnt = nd.cap_cnt;
    if (!USE_TMA) {
        for (uint32_t k = 0; k < cnt; ++k) {
            const double2* q = reinterpret_cast<const double2*>(src + (size_t)k * 6);
            const double2 u = __ldg(q), v = __ldg(q + 1), w = __ldg(q + 2);
            const double ax = u.x - px, ay = u.y - py, az = v.x - pz;
            const double bx = v.y - px, by = w.x - py, bz = w.y - pz;
            acc += solid_angle_2pi(ox, oy, oz, lo, ax, ay, az, sqrt(ax * ax + ay * ay + az * az), bx, by, bz, sqrt(bx * bx + by * by + bz * bz));
        }
        return acc;
    }
    const uint32_t ntile = (cnt + kTileSeg - 1) / kTileSeg;
    if (st.lane == 0 && ntile) {
        const uint32_t m = cnt < (uint32_t)kTileSeg ? cnt : (uint32_t)kTileSeg;
        mbar_expect_tx(&st.bars[0], m * 48u);
        tma_bulk_g2s(st.buf, src, m * 48u, &st.bars[0]);
    }
    for (uint32_t t = 0; t < ntile; ++t) {
        const int s = t & 1;
        if (st.lane == 0 && t + 1 < ntile) {
            const uint32_t rem = cnt - (t + 1) * kTileSeg;
            const uint32_t m = rem < (uint32_t)kTileSeg ? rem : (uint32_t)kTileSeg;
            mbar_expect_tx(&st.bars[s ^ 1], m * 48u);
            tma_bulk_g2s(st.buf + (s ^ 1) * kStageBytes, src + (size_t)(t + 1) * kTileSeg * 6, m * 48u, &st.bars[s ^ 1]);
        }
        mbar_wait(&st.bars[s], st.phase[s]);
        st.phase[s] ^= 1u;
        const uint32_t rem = cnt - t * kTileSeg;
        const uint32_t m = rem < (uint32_t)kTileSeg ? rem : (uint32_t)kTileSeg;
        const double* tile = reinterpret_cast<const double*>(st.buf + s * kStageBytes);
        for (uint32_t k = 0; k < m; ++k) {
            const double2* q = reinterpret_cast<const double2*>(tile + k * 6);
            const double2 u = q[0], v = q[1], w = q[2];
            const double ax = u.x - px, ay = u.y - py, az = v.x - pz;
            const double bx = v.y - px, by = w.x - py, bz = w.y - pz;
            acc += solid_angle_2pi(ox, oy, oz, lo, ax, ay, az, sqrt(ax * ax + ay * ay + az * az), bx, by, bz, sqrt(bx * bx + by * by + bz * bz));
        }
        __syncwarp();  // every lane is done with stage s before it is refilled (two tiles later)
    }
    return acc;
}

// sum over facets [off, off+cnt) of the sorted triangle array
template <bool USE_TMA>
__device__ __forceinline__ double eval_tris(const WView& W, uint32_t off, uint32_t cnt, double px, double py, double pz, Stream<USE_TMA>& st) {
    double acc = 0.0;
    const double* src = W.tris + (size_t)off * 9;
    if (!USE_TMA) {
        for (uint32_t k = 0; k < cnt; ++k) {
            const double* q = src + (size_t)k * 9;
            const double ax = __ldg(q) - px, ay = __ldg(q + 1) - py, az = __ldg(q + 2) - pz;
            const double bx = __ldg(q + 3) - px, by = __ldg(q + 4) - py, bz = __ldg(q + 5) - pz;
            const double cx = __ldg(q + 6) - px, cy = __ldg(q + 7) - py, cz = __ldg(q + 8) - pz;
            acc += solid_angle_2pi(ax, ay, az, sqrt(ax * ax + ay * ay + az * az), bx, by, bz, sqrt(bx * bx + by * by + bz * bz), cx, cy, cz,
                                   sqrt(cx * cx + cy * cy + cz * cz));
        }
        return acc;
    }
    const uint32_t ntile = (cnt + kTileTri - 1) / kTileTri;
    auto bytes_of = [](uint32_t m) { return ((m + 1u) & ~1u) * 72u; };  // even count -> multiple of 16 bytes (array is padded)
    if (st.lane == 0 && ntile) {
        const uint32_t m = cnt < (uint32_t)kTileTri ? cnt : (uint32_t)kTileTri;
        mbar_expect_tx(&st.bars[0], bytes_of(m));
        tma_bulk_g2s(st.buf, src, bytes_of(m), &st.bars[0]);
    }
    for (uint32_t t = 0; t < ntile; ++t) {
        const int s = t & 1;
        if (st.lane == 0 && t + 1 < ntile) {
            const uint32_t rem = cnt - (t + 1) * kTileTri;
            const uint32_t m = rem < (uint32_t)kTileTri ? rem : (uint32_t)kTileTri;
            mbar_expect_tx(&st.bars[s ^ 1], bytes_of(m));
            tma_bulk_g2s(st.buf + (s ^ 1) * kStageBytes, src + (size_t)(t + 1) * kTileTri * 9, bytes_of(m), &st.bars[s ^ 1]);
        }
        mbar_wait(&st.bars[s], st.phase[s]);
        st.phase[s] ^= 1u;
        const uint32_t rem = cnt - t * kTileTri;
        const uint32_t m = rem < (uint32_t)kTileTri ? rem : (uint32_t)kTileTri;
        const double* tile = reinterpret_cast<const double*>(st.buf + s * kStageBytes);
        for (uint32_t k = 0; k < m; ++k) {
            const double* q = tile + k * 9;
            const double ax = q[0] - px, ay = q[1] - py, az = q[2] - pz;
            const double bx = q[3] - px, by = q[4] - py, bz = q[5] - pz;
            const double cx = q[6] - px, cy = q[7] - py, cz = q[8] - pz;
            acc += solid_angle_2pi(ax, ay, az, sqrt(ax * ax + ay * ay + az * az), bx, by, bz, sqrt(bx * bx + by * by + bz * bz), cx, cy, cz,
                                   sqrt(cx * cx + cy * cy + cz * cz));
        }
        __syncwarp();
    }
    return acc;
}

template <bool USE_TMA>
__global__ void __launch_bounds__(kWThreads) winding_kernel(WView W, const double* __restrict__ Q, const uint32_t* __restrict__ perm, uint64_t n,
                                                           double* __restrict__ Wout, uint8_t* __restrict__ keep) {
    __shared__ __align__(128) unsigned char sbuf[USE_TMA ? kWarps * 2 * kStageBytes : 16];
    __shared__ __align__(8) uint64_t sbar[kWarps * 2];
    __shared__ uint32_t sstack[kWarps][kStackDepth];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    Stream<USE_TMA> st;
    st.buf = sbuf + (USE_TMA ? wib * 2 * kStageBytes : 0);
    st.bars = sbar + wib * 2;
    st.phase[0] = st.phase[1] = 0;
    st.lane = lane;
    if (USE_TMA) {
        if (lane == 0) {
            mbar_init(&st.bars[0], 1);
            mbar_init(&st.bars[1], 1);
            mbar_fence_init();
        }
        __syncwarp();
    }
    uint32_t* stack = sstack[wib];
    const uint64_t ngroups = (n + 31) / 32;
    const uint64_t warp = (uint64_t)blockIdx.x * kWarps + wib;
    const uint64_t nwarps = (uint64_t)gridDim.x * kWarps;
    const double inv2pi = 0.15915494309189535;
    for (uint64_t g = warp; g < ngroups; g += nwarps) {
        const uint64_t i = g * 32 + lane;
        const bool valid = i < n;
        const uint64_t src = perm ? (uint64_t)perm[valid ? i : n - 1] : (valid ? i : n - 1);
        const double px = __ldg(Q + 3 * src), py = __ldg(Q + 3 * src + 1), pz = __ldg(Q + 3 * src + 2);
        double acc = 0.0;
        int sp = 0;
        if (lane == 0) stack[0] = 1u;
        sp = 1;
        __syncwarp();
        while (sp > 0) {
            const uint32_t node = stack[sp - 1];
            --sp;
            __syncwarp();
            WNode nd;
            {
                const uint4* q = reinterpret_cast<const uint4*>(W.nodes + node);
                uint4* d = reinterpret_cast<uint4*>(&nd);
                d[0] = __ldg(q); d[1] = __ldg(q + 1); d[2] = __ldg(q + 2); d[3] = __ldg(q + 3);
            }
            if (nd.tri_cnt == 0) continue;
            const bool inside = valid && px >= (double)nd.lo[0] && px <= (double)nd.hi[0] && py >= (double)nd.lo[1] && py <= (double)nd.hi[1] &&
                                pz >= (double)nd.lo[2] && pz <= (double)nd.hi[2];
            const bool any = __any_sync(0xffffffffu, inside);
            if (!any) {
                if (nd.cap_cnt < nd.tri_cnt) acc += eval_cap<USE_TMA>(W, nd, px, py, pz, st);   // (cap smaller than the sub-mesh)
                else acc += eval_tris<USE_TMA>(W, nd.tri_off, nd.tri_cnt, px, py, pz, st);
            } else if (node >= W.nBlkP) {
                acc += eval_tris<USE_TMA>(W, nd.tri_off, nd.tri_cnt, px, py, pz, st);
            } else {
                if (lane == 0) { stack[sp] = 2u * node + 1u; stack[sp + 1] = 2u * node; }
                sp += 2;
                __syncwarp();
            }
        }
        if (valid) {
            const double w = acc * inv2pi;
            if (Wout) Wout[src] = w;
            if (keep) keep[src] = w > 0.5 ? 1 : 0;
        }
        __syncwarp();
    }
}

// ---- query Morton sort ----
__device__ __forceinline__ unsigned long long enc(double d) {
    unsigned long long b = (unsigned long long)__double_as_longlong(d);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double dec(unsigned long long u) {
    unsigned long long b = (u >> 63) ? (u & 0x7fffffffffffffffull) : ~u;
    return __longlong_as_double((long long)b);
}
__global__ void qinit_kernel(unsigned long long* bounds) {
    if (threadIdx.x < 3) bounds[threadIdx.x] = ~0ull;
    else if (threadIdx.x < 6) bounds[threadIdx.x] = 0ull;
}
__global__ void __launch_bounds__(256) qbounds_kernel(const double* __restrict__ Q, uint64_t n, unsigned long long* bounds) {
    double lo[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, hi[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
#pragma unroll
        for (int c = 0; c < 3; ++c) { const double x = Q[3 * i + c]; lo[c] = fmin(lo[c], x); hi[c] = fmax(hi[c], x); }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[c] = fmin(lo[c], __shfl_xor_sync(0xffffffffu, lo[c], o));
            hi[c] = fmax(hi[c], __shfl_xor_sync(0xffffffffu, hi[c], o));
        }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int c = 0; c < 3; ++c) { atomicMin(bounds + c, enc(lo[c])); atomicMax(bounds + 3 + c, enc(hi[c])); }
    }
}
__device__ __forceinline__ uint32_t spread10(uint32_t v) {
    v &= 0x3ffu;
    v = (v | (v << 16)) & 0x030000ffu;
    v = (v | (v << 8)) & 0x0300f00fu;
    v = (v | (v << 4)) & 0x030c30c3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}
__global__ void __launch_bounds__(256) qkeys_kernel(const double* __restrict__ Q, uint64_t n, const unsigned long long* __restrict__ bounds,
                                                    uint32_t* keys, uint32_t* vals) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t code = 0;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const double lo = dec(bounds[c]), hi = dec(bounds[3 + c]);
        const double ext = hi - lo;
        double u = ext > 0.0 ? (Q[3 * i + c] - lo) / ext : 0.0;
        u = fmin(fmax(u, 0.0), 1.0);
        code |= spread10((uint32_t)(u * 1023.0)) << c;
    }
    keys[i] = code;
    vals[i] = (uint32_t)i;
}

// ---- host-side hierarchy construction ----
struct HalfEdge {
    uint64_t key;
    uint32_t blk;
    int32_t dir;
};
struct CapRec {
    uint32_t node, a, b;
};

inline uint64_t spread3h(uint64_t v) {
    v &= 0x1fffffull;
    v = (v | v << 32) & 0x1f00000000ffffull;
    v = (v | v << 16) & 0x1f0000ff0000ffull;
    v = (v | v << 8) & 0x100f00f00f00f00full;
    v = (v | v << 4) & 0x10c30c30c30c30c3ull;
    v = (v | v << 2) & 0x1249249249249249ull;
    return v;
}

struct HostTree {
    std::vector<WNode> nodes;
    std::vector<double> caps;
    std::vector<double> tris;
    uint32_t nBlkP = 1;
};

void build_host_tree(const double* V, uint32_t nV, const uint32_t* F, uint32_t nF, HostTree& T) {
    // 1. merge exactly coincident vertices (libigl: remove_duplicate_vertices(V,F,0.0,...))
    std::vector<uint32_t> idx(nV), canon(nV);
    for (uint32_t i = 0; i < nV; ++i) idx[i] = i;
    auto vless = [&](uint32_t a, uint32_t b) {
        const double *p = V + 3 * (size_t)a, *q = V + 3 * (size_t)b;
        if (p[0] != q[0]) return p[0] < q[0];
        if (p[1] != q[1]) return p[1] < q[1];
        if (p[2] != q[2]) return p[2] < q[2];
        return a < b;
    };
    std::sort(idx.begin(), idx.end(), vless);
    for (uint32_t i = 0; i < nV; ++i) {
        if (i > 0) {
            const double *p = V + 3 * (size_t)idx[i], *q = V + 3 * (size_t)idx[i - 1];
            if (p[0] == q[0] && p[1] == q[1] && p[2] == q[2]) { canon[idx[i]] = canon[idx[i - 1]]; continue; }
        }
        canon[idx[i]] = idx[i];
    }
    // 2. Morton order of facet centroids
    double lo[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, hi[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
    for (size_t k = 0; k < 3 * (size_t)nF; ++k)
        for (int c = 0; c < 3; ++c) {
            const double x = V[3 * (size_t)F[k] + c];
            lo[c] = std::min(lo[c], x);
            hi[c] = std::max(hi[c], x);
        }
    std::vector<std::pair<uint64_t, uint32_t>> order(nF);
    for (uint32_t f = 0; f < nF; ++f) {
        uint64_t code = 0;
        for (int c = 0; c < 3; ++c) {
            const double ctr = (V[3 * (size_t)F[3 * (size_t)f] + c] + V[3 * (size_t)F[3 * (size_t)f + 1] + c] + V[3 * (size_t)F[3 * (size_t)f + 2] + c]) / 3.0;
            const double ext = hi[c] - lo[c];
            double u = ext > 0 ? (ctr - lo[c]) / ext : 0.0;
            u = std::min(std::max(u, 0.0), 1.0);
            code |= spread3h((uint64_t)(u * 2097151.0)) << c;
        }
        order[f] = {code, f};
    }
    std::sort(order.begin(), order.end());
    std::vector<uint32_t> SF(3 * (size_t)nF);
    for (uint32_t j = 0; j < nF; ++j)
        for (int k = 0; k < 3; ++k) SF[3 * (size_t)j + k] = canon[F[3 * (size_t)order[j].second + k]];
    // 3. heap over leaf blocks
    const uint32_t nBlk = (nF + kLeaf - 1) / kLeaf;
    uint32_t nBlkP = 1;
    while (nBlkP < nBlk) nBlkP <<= 1;
    int depth = 0;
    while ((1u << depth) < nBlkP) ++depth;
    T.nBlkP = nBlkP;
    const uint32_t nNodes = 2 * nBlkP;
    std::vector<double> blo(3 * (size_t)nNodes, DBL_MAX), bhi(3 * (size_t)nNodes, -DBL_MAX);
    std::vector<uint32_t> nfac(nNodes, 0), foff(nNodes, 0);
    T.tris.assign(9 * ((size_t)nF + 2), 0.0);  // +2: bulk copies round the triangle count up to even
    for (uint32_t j = 0; j < nF; ++j) {
        const uint32_t node = nBlkP + j / kLeaf;
        for (int k = 0; k < 3; ++k) {
            const double* p = V + 3 * (size_t)SF[3 * (size_t)j + k];
            for (int c = 0; c < 3; ++c) {
                T.tris[9 * (size_t)j + 3 * k + c] = p[c];
                blo[3 * (size_t)node + c] = std::min(blo[3 * (size_t)node + c], p[c]);
                bhi[3 * (size_t)node + c] = std::max(bhi[3 * (size_t)node + c], p[c]);
            }
        }
        nfac[node]++;
    }
    for (uint32_t b = 0; b < nBlkP; ++b) foff[nBlkP + b] = std::min((uint64_t)b * kLeaf, (uint64_t)nF);
    for (uint32_t i = nBlkP - 1; i >= 1; --i) {
        for (int c = 0; c < 3; ++c) {
            blo[3 * (size_t)i + c] = std::min(blo[3 * (size_t)(2 * i) + c], blo[3 * (size_t)(2 * i + 1) + c]);
            bhi[3 * (size_t)i + c] = std::max(bhi[3 * (size_t)(2 * i) + c], bhi[3 * (size_t)(2 * i + 1) + c]);
        }
        nfac[i] = nfac[2 * i] + nfac[2 * i + 1];
        foff[i] = foff[2 * i];
    }
    // 4. exterior edges of every node
    std::vector<HalfEdge> he;
    he.reserve(3 * (size_t)nF);
    for (uint32_t j = 0; j < nF; ++j)
        for (int k = 0; k < 3; ++k) {
            const uint32_t a = SF[3 * (size_t)j + k], b = SF[3 * (size_t)j + (k + 1) % 3];
            if (a == b) continue;
            HalfEdge h;
            h.blk = j / kLeaf;
            if (a < b) { h.key = ((uint64_t)a << 32) | b; h.dir = 1; }
            else { h.key = ((uint64_t)b << 32) | a; h.dir = -1; }
            he.push_back(h);
        }
    std::sort(he.begin(), he.end(), [](const HalfEdge& x, const HalfEdge& y) { return x.key < y.key; });
    std::vector<CapRec> recs;
    recs.reserve(2 * (size_t)nF);
    std::vector<std::pair<uint32_t, int32_t>> per;  // (node, net) scratch
    for (size_t s = 0; s < he.size();) {
        size_t e = s;
        while (e < he.size() && he[e].key == he[s].key) ++e;
        const uint32_t u = (uint32_t)(he[s].key >> 32), v = (uint32_t)(he[s].key & 0xffffffffu);
        for (int sh = 0; sh <= depth; ++sh) {
            per.clear();
            for (size_t k = s; k < e; ++k) {
                const uint32_t node = (nBlkP + he[k].blk) >> sh;
                bool found = false;
                for (auto& pr : per)
                    if (pr.first == node) { pr.second += he[k].dir; found = true; break; }
                if (!found) per.push_back({node, he[k].dir});
            }
            bool any = false;
            for (auto& pr : per) {
                if (pr.second == 0) continue;
                any = true;
                const uint32_t a = pr.second > 0 ? u : v, b = pr.second > 0 ? v : u;
                for (int32_t r = 0; r < std::abs(pr.second); ++r) recs.push_back({pr.first, a, b});
            }
            if (!any && per.size() == 1) break;  // all members in one node and cancelling: same for every ancestor
        }
        s = e;
    }
    // 5. bucket by node, choose apex, drop segments touching it
    std::vector<uint32_t> cnt(nNodes + 1, 0);
    for (auto& r : recs) cnt[r.node + 1]++;
    for (uint32_t i = 0; i < nNodes; ++i) cnt[i + 1] += cnt[i];
    std::vector<CapRec> sorted(recs.size());
    {
        std::vector<uint32_t> cur(cnt.begin(), cnt.end() - 1);
        for (auto& r : recs) sorted[cur[r.node]++] = r;
    }
    T.nodes.assign(nNodes, WNode{});
    T.caps.clear();
    T.caps.reserve(6 * recs.size() + 8);
    for (uint32_t i = 1; i < nNodes; ++i) {
        WNode& nd = T.nodes[i];
        for (int c = 0; c < 3; ++c) {
            nd.lo[c] = nfac[i] ? nextafterf((float)blo[3 * (size_t)i + c], -INFINITY) : INFINITY;
            nd.hi[c] = nfac[i] ? nextafterf((float)bhi[3 * (size_t)i + c], INFINITY) : -INFINITY;
        }
        nd.tri_off = foff[i];
        nd.tri_cnt = nfac[i];
        nd.cap_off = (uint32_t)(T.caps.size() / 6);
        nd.cap_cnt = 0;
        nd.apex[0] = nd.apex[1] = nd.apex[2] = 0.0;
        if (cnt[i + 1] > cnt[i]) {
            const uint32_t apex = sorted[cnt[i]].a;
            for (int c = 0; c < 3; ++c) nd.apex[c] = V[3 * (size_t)apex + c];
            for (uint32_t k = cnt[i]; k < cnt[i + 1]; ++k) {
                const CapRec& r = sorted[k];
                if (r.a == apex || r.b == apex) continue;
                for (int c = 0; c < 3; ++c) T.caps.push_back(V[3 * (size_t)r.a + c]);
                for (int c = 0; c < 3; ++c) T.caps.push_back(V[3 * (size_t)r.b + c]);
                nd.cap_cnt++;
            }
        }
    }
    T.caps.resize(T.caps.size() + 8, 0.0);
}

cudaStream_t pick(twg_ctx* c, void* stream) { return stream ? (cudaStream_t)stream : c->streams[0]; }

}  // namespace

struct twg_winding {
    twg_ctx* ctx = nullptr;
    WNode* nodes = nullptr;
    double* caps = nullptr;
    double* tris = nullptr;
    uint32_t nBlkP = 1, nF = 0;
    uint64_t n_nodes = 0, n_caps = 0;
    bool use_tma = true;
    bool sort_queries = true;
    WView view() const { return WView{nodes, caps, tris, nBlkP, nF}; }
};

extern "C" {

void twg_winding_destroy(twg_winding* w) {
    if (!w) return;
    if (w->ctx) cudaSetDevice(w->ctx->device);
    cudaFree(w->nodes);
    cudaFree(w->caps);
    cudaFree(w->tris);
    delete w;
}

int twg_winding_create(twg_ctx* c, const double* V, uint32_t nV, const uint32_t* F, uint32_t nF, twg_winding** out) {
    TWG_CHECK(c, c && out && (nF == 0 || (V && F)), TWG_ERR_INVALID_ARG, "null argument");
    TWG_CUDA(c, cudaSetDevice(c->device));
    twg_winding* w = new twg_winding;
    w->ctx = c;
    w->nF = nF;
    const char* e1 = getenv("TWG_WINDING_TMA");
    if (e1 && e1[0] == '0') w->use_tma = false;
    const char* e2 = getenv("TWG_WINDING_SORT");
    if (e2 && e2[0] == '0') w->sort_queries = false;
    if (nF == 0) { *out = w; return 0; }
    HostTree T;
    build_host_tree(V, nV, F, nF, T);
    w->nBlkP = T.nBlkP;
    w->n_nodes = T.nodes.size();
    w->n_caps = T.caps.size() / 6;
    cudaStream_t st = c->streams[0];
    cudaError_t e = cudaMalloc(&w->nodes, T.nodes.size() * sizeof(WNode));
    if (e == cudaSuccess) e = cudaMalloc(&w->caps, T.caps.size() * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&w->tris, T.tris.size() * sizeof(double));
    if (e == cudaSuccess) e = cudaMemcpyAsync(w->nodes, T.nodes.data(), T.nodes.size() * sizeof(WNode), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(w->caps, T.caps.data(), T.caps.size() * sizeof(double), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(w->tris, T.tris.data(), T.tris.size() * sizeof(double), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { twg_winding_destroy(w); return twg_fail(c, (int)e, cudaGetErrorString(e), __FILE__, __LINE__); }
    *out = w;
    return 0;
}

int twg_winding_stats(const twg_winding* w, uint64_t* n_nodes, uint64_t* n_caps, uint64_t* n_tris) {
    if (!w) return TWG_ERR_INVALID_ARG;
    if (n_nodes) *n_nodes = w->n_nodes;
    if (n_caps) *n_caps = w->n_caps;
    if (n_tris) *n_tris = w->nF;
    return 0;
}

int twg_winding_eval_dev(twg_winding* w, const double* dC, uint64_t nC, double* dW, uint8_t* dKeep, void* stream) {
    twg_ctx* c = w ? w->ctx : nullptr;
    TWG_CHECK(c, w && dC && (dW || dKeep), TWG_ERR_INVALID_ARG, "null argument");
    TWG_CHECK(c, nC < 0xffffffffull, TWG_ERR_INVALID_ARG, "at most 2^32-2 queries per call");
    if (nC == 0) return 0;
    TWG_CUDA(c, cudaSetDevice(c->device));
    cudaStream_t st = pick(c, stream);
    if (w->nF == 0) {
        if (dW) TWG_CUDA(c, cudaMemsetAsync(dW, 0, nC * sizeof(double), st));
        if (dKeep) TWG_CUDA(c, cudaMemsetAsync(dKeep, 0, nC, st));
        return 0;
    }
    uint32_t* perm = nullptr;
    if (w->sort_queries && nC > 32) {
        // scratch slot 2: bounds | keys | keys2 | vals | vals2 | cub temp
        auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
        size_t tmp_bytes = 0;
        TWG_CUDA(c, cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, (uint32_t*)nullptr, (uint32_t*)nullptr, (uint32_t*)nullptr, (uint32_t*)nullptr,
                                                    (int)nC, 0, 30, st));
        const size_t kb = up(nC * 4);
        TWG_TRY(twg_ensure_scratch(c, 2, 256 + 4 * kb + up(tmp_bytes)));
        char* base = (char*)c->dscratch[2];
        unsigned long long* bounds = (unsigned long long*)base;
        uint32_t *keys = (uint32_t*)(base + 256), *keys2 = (uint32_t*)(base + 256 + kb), *vals = (uint32_t*)(base + 256 + 2 * kb),
                 *vals2 = (uint32_t*)(base + 256 + 3 * kb);
        void* tmp = base + 256 + 4 * kb;
        TWG_LAUNCH(c, qinit_kernel, 1, 32, 0, st, bounds);
        unsigned g = (unsigned)std::min<uint64_t>((nC + 255) / 256, (uint64_t)c->sm_count * 8);
        TWG_LAUNCH(c, qbounds_kernel, g, 256, 0, st, dC, nC, bounds);
        TWG_LAUNCH(c, qkeys_kernel, (unsigned)((nC + 255) / 256), 256, 0, st, dC, nC, bounds, keys, vals);
        TWG_CUDA(c, cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys, keys2, vals, vals2, (int)nC, 0, 30, st));
        c->launches += 4;
        perm = vals2;
    }
    const uint64_t ngroups = (nC + 31) / 32;
    unsigned grid = (unsigned)std::min<uint64_t>((ngroups + kWarps - 1) / kWarps, (uint64_t)c->sm_count * 32);
    if (w->use_tma) TWG_LAUNCH(c, (winding_kernel<true>), grid, kWThreads, 0, st, w->view(), dC, perm, nC, dW, dKeep);
    else TWG_LAUNCH(c, (winding_kernel<false>), grid, kWThreads, 0, st, w->view(), dC, perm, nC, dW, dKeep);
    return 0;
}

int twg_winding_eval(twg_winding* w, const double* C, uint64_t nC, double* W, uint8_t* keep) {
    twg_ctx* c = w ? w->ctx : nullptr;
    TWG_CHECK(c, w && C && (W || keep), TWG_ERR_INVALID_ARG, "null argument");
    if (nC == 0) return 0;
    TWG_CUDA(c, cudaSetDevice(c->device));
    const uint64_t chunk = 1ull << 23;  // 8 Mi queries: 192 MiB in per slot
    auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const uint64_t cmax = nC < chunk ? nC : chunk;
    const size_t qb = up(cmax * 24), wb = up(cmax * 8), kb = up(cmax);
    // slots 0 and 1 alternate (slot 2 is the sort scratch, shared: chunks are serialised on the sort by stream order)
    for (int k = 0; k < 2; ++k) TWG_TRY(twg_ensure_scratch(c, k, qb + wb + kb));
    int slot = 0;
    for (uint64_t b = 0; b < nC; b += chunk, slot ^= 1) {
        const uint64_t m = (nC - b < chunk) ? (nC - b) : chunk;
        cudaStream_t st = c->streams[slot];
        char* base = (char*)c->dscratch[slot];
        double* dQ = (double*)base;
        double* dW = (double*)(base + qb);
        uint8_t* dK = (uint8_t*)(base + qb + wb);
        TWG_CUDA(c, cudaMemcpyAsync(dQ, C + 3 * b, m * 24, cudaMemcpyHostToDevice, st));
        // the sort scratch (slot 2) is shared by both streams: order the kernels of consecutive chunks
        if (b > 0) TWG_CUDA(c, cudaStreamWaitEvent(st, c->ev[slot ^ 1], 0));
        TWG_TRY(twg_winding_eval_dev(w, dQ, m, W ? dW : nullptr, keep ? dK : nullptr, st));
        TWG_CUDA(c, cudaEventRecord(c->ev[slot], st));
        if (W) TWG_CUDA(c, cudaMemcpyAsync(W + b, dW, m * 8, cudaMemcpyDeviceToHost, st));
        if (keep) TWG_CUDA(c, cudaMemcpyAsync(keep + b, dK, m, cudaMemcpyDeviceToHost, st));
    }
    for (int k = 0; k < 2; ++k) TWG_CUDA(c, cudaStreamSynchronize(c->streams[k]));
    return 0;
}

int twg_winding_number(twg_ctx* c, const double* V, uint32_t nV, const uint32_t* F, uint32_t nF, const double* C, uint64_t nC, double* W,
                       uint8_t* keep) {
    twg_winding* w = nullptr;
    TWG_TRY(twg_winding_create(c, V, nV, F, nF, &w));
    int rc = twg_winding_eval(w, C, nC, W, keep);
    twg_winding_destroy(w);
    return rc;
}

int twg_inout_filter(twg_ctx* c, const double* V, uint32_t nV, const uint32_t* F, uint32_t nF, const double* C, uint64_t nC, uint8_t* keep,
                     int* retried) {
    TWG_CHECK(c, c && keep, TWG_ERR_INVALID_ARG, "null argument");
    if (retried) *retried = 0;
    TWG_TRY(twg_winding_number(c, V, nV, F, nF, C, nC, nullptr, keep));
    uint64_t kept = 0;
    for (uint64_t i = 0; i < nC; ++i) kept += keep[i];
    if (kept != 0) return 0;
    // InoutFiltering.cpp:56-75: the surface may be totally reversed -> swap columns 1,2 and try again
    if (retried) *retried = 1;
    std::vector<uint32_t> F2(3 * (size_t)nF);
    for (uint32_t f = 0; f < nF; ++f) { F2[3 * (size_t)f] = F[3 * (size_t)f]; F2[3 * (size_t)f + 1] = F[3 * (size_t)f + 2]; F2[3 * (size_t)f + 2] = F[3 * (size_t)f + 1]; }
    return twg_winding_number(c, V, nV, F2.data(), nF, C, nC, nullptr, keep);
}

}  // extern "C"
