// winding.cu -- generalized winding number of query points against a triangle surface, and the W > 0.5 filter.
//
// Replaces igl::winding_number(V,F,O,W) as called by InoutFiltering::filter (src/tetwild/InoutFiltering.cpp:45-75)
// and MeshRefinement::markInOut / outputMidResult (src/tetwild/MeshRefinement.cpp:592-624, :1036-1068).
//
// Algorithm: the exact hierarchical evaluation of Jacobson et al. 2013 (the one libigl runs), laid out for the GPU.
//   * facets are Morton-sorted and cut into leaf blocks; an implicit binary heap of nodes covers contiguous block
//     ranges. Each node stores its bounding box and its CAP: the exterior (unmatched) directed edges of its sub-mesh,
//     fanned to one apex vertex. For a query outside the node's box the solid angle of the sub-mesh equals that of
//     the cap exactly (they share their boundary and the closed difference lies inside the convex box), so far
//     sub-meshes cost O(sqrt(#facets)) instead of O(#facets).
//   * cap edges are traced into POLYLINES on the host (a closed loop for a manifold patch): consecutive fan triangles
//     (apex, P_k, P_k+1) share P_k+1, so each extra boundary edge costs one point (24 B + a chain-start flag), one
//     norm and one square root instead of two of each.
//   * the per-triangle angle atan2(y, x) of the Van Oosterom-Strackee formula, y = det[a b c],
//     x = |a||b||c| + (a.b)|c| + (b.c)|a| + (c.a)|b|, is NOT evaluated per triangle: sum_f atan2(y_f, x_f) =
//     arg(prod_f (x_f + i y_f)) + 2 pi k. Each lane keeps the running complex product z (factors rescaled by a power
//     of two, z renormalised once per tile) and the integer k, which changes exactly when a factor rotates z across
//     the negative real axis (sign test on Im z before/after). One atan2 per QUERY instead of one per (query,
//     triangle) pair; 4 FP64 multiply-adds per pair instead of ~80 FP64 instructions.
//   * queries are Morton-sorted on the device (qsort.cu); a warp owns 32 consecutive (hence spatially coherent)
//     queries and traverses the heap with ONE shared stack: a node is opened iff some lane lies inside its box,
//     otherwise all lanes add its cap. Cap polylines and leaf triangles are streamed, tile by tile, into a per-warp
//     shared-memory ring with 1-D bulk async copies (TMA, cp.async.bulk + mbarrier, double-buffered) and consumed by
//     all 32 lanes with conflict-free broadcast reads.
// FP64-pipe-bound by design (no tensor-core formulation exists for sqrt / cross-product chains).
#include <algorithm>
#include <atomic>
#include <chrono>
#include <thread>
#include "common.cuh"
#include "winding_math.cuh"
#include "winding.cuh"

namespace {

constexpr int kWThreads = 256;        // 8 warps per CTA
constexpr int kWarps = kWThreads / 32;
constexpr int kTilePts = 32;          // cap points per staged tile   (32 * 32 B = 1024 B)
constexpr int kTileTri = 32;          // triangles per staged tile    (32 * 72 B = 2304 B)
constexpr int kStageBytes = 2304;
constexpr int kStackDepth = 64;

using tww::norm3;
using tww::Angle;

struct Stream {  // per-warp tile streamer
    unsigned char* buf;  // 2 stages of kStageBytes
    uint64_t* bars;      // 2 mbarriers
    uint32_t phase[2];
    int lane;
};

// fan triangles (apex, P_k, P_k+1) over the cap polylines of node nd
__device__ __forceinline__ void eval_cap(const WView& W, const WNode& nd, double px, double py, double pz, Stream& st, Angle& acc) {
    const double ox = nd.apex[0] - px, oy = nd.apex[1] - py, oz = nd.apex[2] - pz;
    const double lo = norm3(ox, oy, oz);
    const double* src = W.caps + (size_t)nd.cap_off * 4;
    const uint32_t cnt = nd.cap_cnt;
    const uint32_t ntile = (cnt + kTilePts - 1) / kTilePts;
    if (st.lane == 0 && ntile) {
        const uint32_t m = cnt < (uint32_t)kTilePts ? cnt : (uint32_t)kTilePts;
        mbar_expect_tx(&st.bars[0], m * 32u);
        tma_bulk_g2s(st.buf, src, m * 32u, &st.bars[0]);
    }
    double ax = 0.0, ay = 0.0, az = 0.0, la = 0.0, oa = 0.0;
    for (uint32_t t = 0; t < ntile; ++t) {
        const int s = t & 1;
        if (st.lane == 0 && t + 1 < ntile) {
            const uint32_t rem = cnt - (t + 1) * kTilePts;
            const uint32_t m = rem < (uint32_t)kTilePts ? rem : (uint32_t)kTilePts;
            mbar_expect_tx(&st.bars[s ^ 1], m * 32u);
            tma_bulk_g2s(st.buf + (s ^ 1) * kStageBytes, src + (size_t)(t + 1) * kTilePts * 4, m * 32u, &st.bars[s ^ 1]);
        }
        mbar_wait(&st.bars[s], st.phase[s]);
        st.phase[s] ^= 1u;
        const uint32_t rem = cnt - t * kTilePts;
        const uint32_t m = rem < (uint32_t)kTilePts ? rem : (uint32_t)kTilePts;
        const double2* tile = reinterpret_cast<const double2*>(st.buf + s * kStageBytes);
#pragma unroll 4
        for (uint32_t k = 0; k < m; ++k) {
            const double2 u = tile[2 * k], v = tile[2 * k + 1];  // (x, y), (z, flag): the same address for every lane
            const double bx = u.x - px, by = u.y - py, bz = v.x - pz;
            const double lb = norm3(bx, by, bz);
            const double ob = ox * bx + oy * by + oz * bz;
            const double ab = ax * bx + ay * by + az * bz;
            const double y = ox * (ay * bz - az * by) + oy * (az * bx - ax * bz) + oz * (ax * by - ay * bx);
            const double x = lo * (la * lb + ab) + ob * la + oa * lb;
            acc.mul(x, y, tww::hi32(v.y) != 0);  // flag 1.0 (integer test): the first point of a chain closes no triangle
            ax = bx; ay = by; az = bz; la = lb; oa = ob;
        }
        acc.renorm();
        __syncwarp();  // every lane is done with stage s before it is refilled (two tiles later)
    }
}

// facets [off, off+cnt) of the sorted triangle array
__device__ __forceinline__ void eval_tris(const WView& W, uint32_t off, uint32_t cnt, double px, double py, double pz, Stream& st, Angle& acc) {
    const double* src = W.tris + (size_t)off * 9;
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
#pragma unroll 2
        for (uint32_t k = 0; k < m; ++k) {
            const double* q = tile + k * 9;
            const double ax = q[0] - px, ay = q[1] - py, az = q[2] - pz;
            const double bx = q[3] - px, by = q[4] - py, bz = q[5] - pz;
            const double cx = q[6] - px, cy = q[7] - py, cz = q[8] - pz;
            const double la = norm3(ax, ay, az), lb = norm3(bx, by, bz), lc = norm3(cx, cy, cz);
            const double y = ax * (by * cz - bz * cy) + bx * (cy * az - cz * ay) + cx * (ay * bz - az * by);
            const double x = la * lb * lc + (bx * cx + by * cy + bz * cz) * la + (cx * ax + cy * ay + cz * az) * lb + (ax * bx + ay * by + az * bz) * lc;
            acc.mul(x, y, false);
        }
        acc.renorm();
        __syncwarp();
    }
}

// MINB = resident CTAs per SM the register allocation is capped for: 4 -> 64 registers (32 warps per SM, a few bytes of
// spill outside the tile loops), 3 -> 76 registers (24 warps). The kernel is bound by FP64 issue latency, not bandwidth.
template <int MINB>
__global__ void __launch_bounds__(kWThreads, MINB) winding_kernel(WView W, const double* __restrict__ Q, const uint32_t* __restrict__ perm, uint64_t n,
                                                           double* __restrict__ Wout, uint8_t* __restrict__ keep, unsigned long long* dbg) {
    __shared__ __align__(128) unsigned char sbuf[kWarps * 2 * kStageBytes];
    __shared__ __align__(8) uint64_t sbar[kWarps * 2];
    __shared__ uint32_t sstack[kWarps][kStackDepth];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    Stream st;
    st.buf = sbuf + wib * 2 * kStageBytes;
    st.bars = sbar + wib * 2;
    st.phase[0] = st.phase[1] = 0;
    st.lane = lane;
    if (lane == 0) {
        mbar_init(&st.bars[0], 1);
        mbar_init(&st.bars[1], 1);
        mbar_fence_init();
    }
    __syncwarp();
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
        Angle acc;
        acc.init();
        int sp = 0;
        unsigned long long pairs = 0;  // (query, cap point or facet) evaluations of this group (diagnostic: twg_debug_counter 2)
        const unsigned nvalid = __popc(__ballot_sync(0xffffffffu, valid));
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
                // a cap point costs about 2/3 of a leaf triangle (one norm instead of three)
                if (2u * nd.cap_cnt < 3u * nd.tri_cnt) { eval_cap(W, nd, px, py, pz, st, acc); pairs += (unsigned long long)nd.cap_cnt * nvalid; }
                else { eval_tris(W, nd.tri_off, nd.tri_cnt, px, py, pz, st, acc); pairs += (unsigned long long)nd.tri_cnt * nvalid; }
            } else if (node >= W.nBlkP) {
                eval_tris(W, nd.tri_off, nd.tri_cnt, px, py, pz, st, acc);
                pairs += (unsigned long long)nd.tri_cnt * nvalid;
            } else {
                if (lane == 0) { stack[sp] = 2u * node + 1u; stack[sp + 1] = 2u * node; }
                sp += 2;
                __syncwarp();
            }
        }
        if (valid) {
            const double w = acc.total() * inv2pi;
            if (Wout) Wout[src] = w;
            if (keep) keep[src] = w > 0.5 ? 1 : 0;
        }
        if (lane == 0) atomicAdd(dbg + TWG_DBG_WINDING_PAIRS, pairs);
        __syncwarp();
    }
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


// ---- small host-side parallel helpers for the hierarchy build (std::thread, no OpenMP dependency)
unsigned host_threads() {
    unsigned n = std::thread::hardware_concurrency();
    if (n == 0) n = 1;
    return n > 16 ? 16 : n;
}
// fn(begin, end, part) over [0, n) cut into `parts` contiguous ranges, one thread each
template <class FN>
void parallel_ranges(size_t n, unsigned parts, FN fn, size_t min_items = 4096) {
    if (parts <= 1 || n < min_items) { fn((size_t)0, n, 0u); return; }
    std::vector<std::thread> pool;
    for (unsigned p = 0; p < parts; ++p) pool.emplace_back([&, p] { fn(n * p / parts, n * (p + 1) / parts, p); });
    for (auto& t : pool) t.join();
}
// sort = per-thread std::sort of equal chunks + rounds of pairwise in-place merges (deterministic for a strict weak order;
// elements that compare equal may land in any relative order, like std::sort)
template <class IT, class CMP>
void parallel_sort(IT first, IT last, CMP cmp) {
    const size_t n = (size_t)(last - first);
    unsigned parts = host_threads();
    while (parts > 1 && n / parts < 65536) parts >>= 1;
    unsigned p2 = 1;
    while (p2 * 2 <= parts) p2 *= 2;  // power of two chunks
    if (p2 <= 1) { std::sort(first, last, cmp); return; }
    auto bound = [&](unsigned k) { return first + (ptrdiff_t)(n * k / p2); };
    parallel_ranges(p2, p2, [&](size_t b, size_t e, unsigned) { for (size_t k = b; k < e; ++k) std::sort(bound((unsigned)k), bound((unsigned)k + 1), cmp); }, 0);
    for (unsigned width = 1; width < p2; width *= 2) {
        const unsigned pairs = p2 / (2 * width);
        parallel_ranges(pairs, pairs, [&](size_t b, size_t e, unsigned) {
            for (size_t k = b; k < e; ++k) {
                const unsigned lo = (unsigned)k * 2 * width;
                std::inplace_merge(bound(lo), bound(lo + width), bound(lo + 2 * width), cmp);
            }
        }, 0);
    }
}

struct PhaseTimer {  // TWG_TRACE=1: wall time of the build phases on stderr
    bool on;
    std::chrono::steady_clock::time_point t0;
    PhaseTimer() : on(getenv("TWG_TRACE") != nullptr), t0(std::chrono::steady_clock::now()) {}
    void lap(const char* what) {
        if (!on) return;
        const auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "[twg] winding build: %-28s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    }
};

void build_host_tree(const double* V, uint32_t nV, const uint32_t* F, uint32_t nF, uint32_t kLeaf, HostTree& T) {
    PhaseTimer tm;
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
    parallel_sort(idx.begin(), idx.end(), vless);
    for (uint32_t i = 0; i < nV; ++i) {
        if (i > 0) {
            const double *p = V + 3 * (size_t)idx[i], *q = V + 3 * (size_t)idx[i - 1];
            if (p[0] == q[0] && p[1] == q[1] && p[2] == q[2]) { canon[idx[i]] = canon[idx[i - 1]]; continue; }
        }
        canon[idx[i]] = idx[i];
    }
    tm.lap("1 vertex merge");
    // 2. kd order of the facets: the heap node that owns blocks [i*2^h, (i+1)*2^h) must be a COMPACT patch, because the
    //    price of a far sub-mesh is the length of its boundary. Recursive split of the centroid set at the block-aligned
    //    midpoint along the longest axis of its bounding box (libigl's WindingNumberAABB splits the same way); a plain
    //    Morton order gives z-curve ranges with boundaries ~2.5x longer (measured: 7.6 sqrt(n) vs 3 sqrt(n) edges).
    //    The leaf size is shrunk so that the blocks fill the power-of-two heap evenly.
    uint32_t nBlkP = 1;
    while ((uint64_t)nBlkP * kLeaf < nF) nBlkP <<= 1;
    kLeaf = ((nF + nBlkP - 1) / nBlkP + 1u) & ~1u;  // even: a block then starts on a 16-byte boundary (72 B per facet) for the bulk copies
    if (kLeaf < 2) kLeaf = 2;
    std::vector<float> ctr(3 * (size_t)nF);
    for (uint32_t f = 0; f < nF; ++f)
        for (int c = 0; c < 3; ++c)
            ctr[3 * (size_t)f + c] = (float)((V[3 * (size_t)F[3 * (size_t)f] + c] + V[3 * (size_t)F[3 * (size_t)f + 1] + c] + V[3 * (size_t)F[3 * (size_t)f + 2] + c]) / 3.0);
    std::vector<uint32_t> order(nF);
    for (uint32_t f = 0; f < nF; ++f) order[f] = f;
    {
        struct Task { uint32_t b, e, blocks; };
        // breadth-first to a fixed depth, then the subtrees are independent: split them over host threads
        std::vector<Task> todo(1, Task{0, nF, nBlkP});
        auto split = [&](const Task& t, Task& l, Task& r) -> bool {
            if (t.blocks <= 1 || t.e - t.b <= kLeaf) return false;
            const uint64_t left_cap = (uint64_t)(t.blocks / 2) * kLeaf;
            const uint32_t mid = (uint32_t)std::min<uint64_t>(t.e, t.b + left_cap);
            if (mid < t.e) {
                float lo3[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi3[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
                for (uint32_t k = t.b; k < t.e; ++k)
                    for (int c = 0; c < 3; ++c) {
                        const float x = ctr[3 * (size_t)order[k] + c];
                        lo3[c] = std::min(lo3[c], x);
                        hi3[c] = std::max(hi3[c], x);
                    }
                int ax = 0;
                if (hi3[1] - lo3[1] > hi3[ax] - lo3[ax]) ax = 1;
                if (hi3[2] - lo3[2] > hi3[ax] - lo3[ax]) ax = 2;
                std::nth_element(order.begin() + t.b, order.begin() + mid, order.begin() + t.e, [&](uint32_t x, uint32_t y) {
                    const float a = ctr[3 * (size_t)x + ax], b2 = ctr[3 * (size_t)y + ax];
                    return a < b2 || (a == b2 && x < y);
                });
            }
            l = Task{t.b, mid, t.blocks / 2};
            r = Task{mid, t.e, t.blocks / 2};
            return true;
        };
        auto run = [&](Task t0) {
            std::vector<Task> st(1, t0);
            while (!st.empty()) {
                Task t = st.back(), l, r;
                st.pop_back();
                if (split(t, l, r)) { st.push_back(l); st.push_back(r); }
            }
        };
        unsigned nthreads = std::thread::hardware_concurrency();
        if (nthreads == 0) nthreads = 1;
        if (nthreads > 16) nthreads = 16;
        if (nF < 100000 || nthreads == 1) {
            run(todo[0]);
        } else {
            for (int lvl = 0; lvl < 5; ++lvl) {  // 32 independent subtrees
                std::vector<Task> next;
                for (auto& t : todo) { Task l, r; if (split(t, l, r)) { next.push_back(l); next.push_back(r); } }
                todo.swap(next);
            }
            std::atomic<size_t> cursor(0);
            std::vector<std::thread> pool;
            for (unsigned w = 0; w < nthreads; ++w)
                pool.emplace_back([&]() { for (size_t k; (k = cursor.fetch_add(1)) < todo.size();) run(todo[k]); });
            for (auto& th : pool) th.join();
        }
    }
    // canonical order inside every leaf block: by facet index. The kd splits fix which facets share a block (std::nth_element
    // leaves their arrangement unspecified); sorting each block makes the facet array -- hence every sum -- a function of the
    // input alone, and lets the device build (winding_build.cu) reproduce this hierarchy bit for bit.
    parallel_ranges(nBlkP, host_threads(), [&](size_t b0, size_t b1, unsigned) {
        for (size_t b = b0; b < b1; ++b) {
            const uint64_t j0 = std::min<uint64_t>((uint64_t)b * kLeaf, nF), j1 = std::min<uint64_t>((uint64_t)(b + 1) * kLeaf, nF);
            std::sort(order.begin() + j0, order.begin() + j1);
        }
    });
    std::vector<uint32_t> SF(3 * (size_t)nF);
    for (uint32_t j = 0; j < nF; ++j)
        for (int k = 0; k < 3; ++k) SF[3 * (size_t)j + k] = canon[F[3 * (size_t)order[j] + k]];
    tm.lap("2 kd order");
    // 3. heap over leaf blocks
    int depth = 0;
    while ((1u << depth) < nBlkP) ++depth;
    T.nBlkP = nBlkP;
    const uint32_t nNodes = 2 * nBlkP;
    std::vector<double> blo(3 * (size_t)nNodes, DBL_MAX), bhi(3 * (size_t)nNodes, -DBL_MAX);
    std::vector<uint32_t> nfac(nNodes, 0), foff(nNodes, 0);
    T.tris.assign(9 * ((size_t)nF + 2), 0.0);  // +2: bulk copies round the triangle count up to even
    parallel_ranges(nBlkP, host_threads(), [&](size_t b0, size_t b1, unsigned) {  // a leaf block is owned by one thread
        const uint64_t j0 = std::min<uint64_t>((uint64_t)b0 * kLeaf, nF), j1 = std::min<uint64_t>((uint64_t)b1 * kLeaf, nF);
        for (uint64_t j = j0; j < j1; ++j) {
            const uint32_t node = nBlkP + (uint32_t)(j / kLeaf);
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
    });
    for (uint32_t b = 0; b < nBlkP; ++b) foff[nBlkP + b] = std::min((uint64_t)b * kLeaf, (uint64_t)nF);
    for (uint32_t i = nBlkP - 1; i >= 1; --i) {
        for (int c = 0; c < 3; ++c) {
            blo[3 * (size_t)i + c] = std::min(blo[3 * (size_t)(2 * i) + c], blo[3 * (size_t)(2 * i + 1) + c]);
            bhi[3 * (size_t)i + c] = std::max(bhi[3 * (size_t)(2 * i) + c], bhi[3 * (size_t)(2 * i + 1) + c]);
        }
        nfac[i] = nfac[2 * i] + nfac[2 * i + 1];
        foff[i] = foff[2 * i];
    }
    tm.lap("3 heap boxes");
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
    tm.lap("4a half-edge list");
    parallel_sort(he.begin(), he.end(), [](const HalfEdge& x, const HalfEdge& y) { return x.key < y.key; });
    tm.lap("4b half-edge sort");
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
    tm.lap("4c exterior edges");
    // 5. bucket by node, trace each node's exterior edges into polylines, choose the apex, cut the polylines at it
    std::vector<uint32_t> cnt(nNodes + 1, 0);
    for (auto& r : recs) cnt[r.node + 1]++;
    for (uint32_t i = 0; i < nNodes; ++i) cnt[i + 1] += cnt[i];
    std::vector<CapRec> sorted(recs.size());
    {
        std::vector<uint32_t> cur(cnt.begin(), cnt.end() - 1);
        for (auto& r : recs) sorted[cur[r.node]++] = r;
    }
    T.nodes.assign(nNodes, WNode{});
    // nodes are independent: every thread traces a contiguous range of nodes into its own cap array (offsets local to the
    // part), the parts are then concatenated in node order, which reproduces the serial layout exactly
    const unsigned parts5 = (nNodes >= 4096) ? host_threads() : 1u;
    std::vector<std::vector<double>> part_caps(parts5);
    // ranges of nodes with equal numbers of edge records (the few nodes near the root carry the longest caps)
    std::vector<uint32_t> part_first(parts5 + 1, nNodes);
    part_first[0] = 0;
    for (unsigned p = 1; p < parts5; ++p) {
        const uint64_t target = (uint64_t)recs.size() * p / parts5;
        part_first[p] = (uint32_t)(std::upper_bound(cnt.begin(), cnt.begin() + nNodes, (uint32_t)target) - cnt.begin());
        if (part_first[p] < part_first[p - 1]) part_first[p] = part_first[p - 1];
        if (part_first[p] > nNodes) part_first[p] = nNodes;
    }
    {
    std::vector<std::thread> pool5;
    auto trace_part = [&](unsigned part) {
    const size_t n0 = part_first[part], n1 = part_first[part + 1];
    std::vector<double>& caps = part_caps[part];
    if (n1 > n0) caps.reserve(4 * ((size_t)(cnt[n1] - cnt[n0]) + (cnt[n1] - cnt[n0]) / 4) + 16);
    std::vector<uint32_t> chain;        // vertex ids of all polylines of one node, back to back
    std::vector<uint32_t> chain_start;  // offsets into `chain`
    std::vector<uint8_t> used;
    for (uint32_t i = (uint32_t)std::max<size_t>(n0, 1); i < (uint32_t)n1; ++i) {
        WNode& nd = T.nodes[i];
        for (int c = 0; c < 3; ++c) {
            nd.lo[c] = nfac[i] ? nextafterf((float)blo[3 * (size_t)i + c], -INFINITY) : INFINITY;
            nd.hi[c] = nfac[i] ? nextafterf((float)bhi[3 * (size_t)i + c], INFINITY) : -INFINITY;
        }
        nd.tri_off = foff[i];
        nd.tri_cnt = nfac[i];
        nd.cap_off = (uint32_t)(caps.size() / 4);  // local to the part; rebased below
        nd.cap_cnt = 0;
        nd.apex[0] = nd.apex[1] = nd.apex[2] = 0.0;
        const uint32_t e0 = cnt[i], e1 = cnt[i + 1];
        if (e1 == e0) continue;
        // edges of this node sorted by start vertex; greedy walks a -> b -> ... consume them (a directed multigraph:
        // closed loops for a manifold patch, arbitrary trails otherwise -- every edge is emitted exactly once)
        std::sort(sorted.begin() + e0, sorted.begin() + e1, [](const CapRec& x, const CapRec& y) { return x.a < y.a || (x.a == y.a && x.b < y.b); });
        used.assign(e1 - e0, 0);
        chain.clear();
        chain_start.clear();
        auto first_unused_from = [&](uint32_t v) -> int64_t {
            uint32_t lo_ = e0, hi_ = e1;
            while (lo_ < hi_) { const uint32_t mid = (lo_ + hi_) / 2; if (sorted[mid].a < v) lo_ = mid + 1; else hi_ = mid; }
            for (uint32_t k = lo_; k < e1 && sorted[k].a == v; ++k)
                if (!used[k - e0]) return (int64_t)k;
            return -1;
        };
        for (uint32_t k = e0; k < e1; ++k) {
            if (used[k - e0]) continue;
            chain_start.push_back((uint32_t)chain.size());
            chain.push_back(sorted[k].a);
            int64_t cur = k;
            while (cur >= 0) {
                used[(uint32_t)cur - e0] = 1;
                chain.push_back(sorted[(uint32_t)cur].b);
                cur = first_unused_from(sorted[(uint32_t)cur].b);
            }
        }
        chain_start.push_back((uint32_t)chain.size());
        const uint32_t apex = chain[0];
        for (int c = 0; c < 3; ++c) nd.apex[c] = V[3 * (size_t)apex + c];
        // a fan triangle with the apex as one of its corners is degenerate (zero solid angle): cut the polylines there
        for (size_t ci = 0; ci + 1 < chain_start.size(); ++ci) {
            uint32_t run = 0;
            for (uint32_t k = chain_start[ci]; k <= chain_start[ci + 1]; ++k) {
                const bool end = (k == chain_start[ci + 1]) || chain[k] == apex;
                if (!end) { ++run; continue; }
                if (run >= 2) {
                    for (uint32_t q = k - run; q < k; ++q) {
                        for (int c = 0; c < 3; ++c) caps.push_back(V[3 * (size_t)chain[q] + c]);
                        caps.push_back(q == k - run ? 1.0 : 0.0);
                        nd.cap_cnt++;
                    }
                }
                run = 0;
            }
        }
    }
    };
    if (parts5 == 1) trace_part(0);
    else {
        for (unsigned p = 0; p < parts5; ++p) pool5.emplace_back(trace_part, p);
        for (auto& t : pool5) t.join();
    }
    }
    {
        size_t total = 0;
        for (auto& pc : part_caps) total += pc.size();
        T.caps.clear();
        T.caps.reserve(total + 8);
        for (unsigned part = 0; part < parts5; ++part) {
            const uint32_t base = (uint32_t)(T.caps.size() / 4);
            const uint32_t n0 = part_first[part], n1 = part_first[part + 1];
            for (uint32_t i = std::max(n0, 1u); i < n1; ++i) T.nodes[i].cap_off += base;
            T.caps.insert(T.caps.end(), part_caps[part].begin(), part_caps[part].end());
        }
    }
    T.caps.resize(T.caps.size() + 8, 0.0);
    tm.lap("5 caps traced");
}

cudaStream_t pick(twg_ctx* c, void* stream) { return stream ? (cudaStream_t)stream : c->streams[0]; }

}  // namespace

extern "C" {

void twg_winding_destroy(twg_winding* w) {
    if (!w) return;
    if (!w->replicas.empty()) {
        for (twg_winding* r : w->replicas) twg_winding_destroy(r);
        delete w;
        return;
    }
    if (w->ctx) cudaSetDevice(w->ctx->device);
    cudaFree(w->nodes);
    cudaFree(w->caps);
    cudaFree(w->tris);
    delete w;
}

// uploads a host-built hierarchy to one device
static int winding_upload(twg_ctx* c, const HostTree& T, uint32_t nF, twg_winding** out) {
    TWG_CUDA(c, cudaSetDevice(c->device));
    twg_winding* w = new twg_winding;
    w->ctx = c;
    w->nF = nF;
    w->leaf = (uint32_t)c->opt.winding_leaf;
    w->sort_queries = c->opt.winding_sort != 0;
    if (nF == 0) { *out = w; return 0; }
    for (int k = 0; k < 3; ++k) {  // root box
        const double lo = T.nodes[1].lo[k], hi = T.nodes[1].hi[k], m = 0.1 * (hi - lo);
        w->sort_box[k] = lo - m;
        w->sort_box[3 + k] = hi + m;
    }
    w->nBlkP = T.nBlkP;
    w->n_nodes = T.nodes.size();
    w->n_caps = T.caps.size() / 4;
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

// surface to the device, hierarchy built there (winding_build.cu)
static int winding_build_on_device(twg_ctx* c, const double* V, uint32_t nV, const uint32_t* F, uint32_t nF, twg_winding** out) {
    TWG_CUDA(c, cudaSetDevice(c->device));
    cudaStream_t st = c->streams[0];
    double* dV = nullptr;
    uint32_t* dF = nullptr;
    TWG_CUDA(c, cudaMalloc(&dV, sizeof(double) * 3 * (size_t)nV));
    cudaError_t e = cudaMalloc(&dF, sizeof(uint32_t) * 3 * (size_t)nF);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dV, V, sizeof(double) * 3 * (size_t)nV, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dF, F, sizeof(uint32_t) * 3 * (size_t)nF, cudaMemcpyHostToDevice, st);
    DeviceTree T;
    int rc = e == cudaSuccess ? twg_winding_build_device(c, dV, nV, dF, nF, (uint32_t)c->opt.winding_leaf, &T) : twg_fail(c, (int)e, cudaGetErrorString(e), __FILE__, __LINE__);
    cudaStreamSynchronize(st);
    cudaFree(dV);
    cudaFree(dF);
    if (rc != 0) { cudaFree(T.nodes); cudaFree(T.caps); cudaFree(T.tris); return rc; }
    twg_winding* w = new twg_winding;
    w->ctx = c;
    w->nF = nF;
    w->leaf = (uint32_t)c->opt.winding_leaf;
    w->sort_queries = c->opt.winding_sort != 0;
    w->nodes = T.nodes; w->caps = T.caps; w->tris = T.tris;
    w->nBlkP = T.nBlkP; w->n_nodes = T.n_nodes; w->n_caps = T.n_caps;
    for (int k = 0; k < 3; ++k) {
        const double lo = T.root_lo[k], hi = T.root_hi[k], m = 0.1 * (hi - lo);
        w->sort_box[k] = lo - m;
        w->sort_box[3 + k] = hi + m;
    }
    *out = w;
    return 0;
}

int twg_winding_create(twg_ctx* c, const double* V, uint32_t nV, const uint32_t* F, uint32_t nF, twg_winding** out) {
    TWG_CHECK(c, c && out && (nF == 0 || (V && F)), TWG_ERR_INVALID_ARG, "null argument");
    TWG_CHECK(c, nF < 0x7fffffffu, TWG_ERR_INVALID_ARG, "at most 2^31-2 facets");
    for (size_t k = 0; k < 3 * (size_t)nF; ++k) TWG_CHECK(c, F[k] < nV, TWG_ERR_INVALID_ARG, "facet references a vertex out of range");
    if (nF && c->opt.winding_device_build) {  // the whole construction on the device (winding_build.cu); every device builds its own replica
        if (twg_is_multi(c)) {
            twg_winding* w = new twg_winding;
            w->ctx = c;
            w->nF = nF;
            w->replicas.assign(c->children.size(), nullptr);
            const int rc = twg_multi_run(c, [&](int k, twg_ctx* child) { return winding_build_on_device(child, V, nV, F, nF, &w->replicas[k]); });
            if (rc != 0) { twg_winding_destroy(w); return rc; }
            w->nBlkP = w->replicas[0]->nBlkP;
            w->n_nodes = w->replicas[0]->n_nodes;
            w->n_caps = w->replicas[0]->n_caps;
            *out = w;
            return 0;
        }
        return winding_build_on_device(c, V, nV, F, nF, out);
    }
    HostTree T;
    if (nF) build_host_tree(V, nV, F, nF, (uint32_t)c->opt.winding_leaf, T);
    if (twg_is_multi(c)) {  // the hierarchy is built once and uploaded to every device by that device's thread
        twg_winding* w = new twg_winding;
        w->ctx = c;
        w->nF = nF;
        w->replicas.assign(c->children.size(), nullptr);
        const int rc = twg_multi_run(c, [&](int k, twg_ctx* child) { return winding_upload(child, T, nF, &w->replicas[k]); });
        if (rc != 0) { twg_winding_destroy(w); return rc; }
        w->nBlkP = w->replicas[0]->nBlkP;
        w->n_nodes = w->replicas[0]->n_nodes;
        w->n_caps = w->replicas[0]->n_caps;
        *out = w;
        return 0;
    }
    return winding_upload(c, T, nF, out);
}

twg_winding* twg_winding_replica(twg_winding* w, int k) {
    if (!w) return nullptr;
    if (w->replicas.empty()) return k == 0 ? w : nullptr;
    return (k >= 0 && k < (int)w->replicas.size()) ? w->replicas[k] : nullptr;
}

// test hook: the three arrays of the hierarchy as they live on the device (replica 0 of a multi-device handle).
// nodes_out: n_nodes * 64 bytes, caps_out: n_cap_points * 4 doubles, tris_out: (n_triangles + 2) * 9 doubles
int twg_debug_winding_download(twg_winding* w, void* nodes_out, double* caps_out, double* tris_out) {
    twg_ctx* c = w ? w->ctx : nullptr;
    TWG_CHECK(c, w != nullptr, TWG_ERR_INVALID_ARG, "null argument");
    if (!w->replicas.empty()) return twg_forward0(c, twg_debug_winding_download(w->replicas[0], nodes_out, caps_out, tris_out));
    if (w->nF == 0) return 0;
    TWG_CUDA(c, cudaSetDevice(c->device));
    if (nodes_out) TWG_CUDA(c, cudaMemcpy(nodes_out, w->nodes, w->n_nodes * sizeof(WNode), cudaMemcpyDeviceToHost));
    if (caps_out) TWG_CUDA(c, cudaMemcpy(caps_out, w->caps, w->n_caps * 4 * sizeof(double), cudaMemcpyDeviceToHost));
    if (tris_out) TWG_CUDA(c, cudaMemcpy(tris_out, w->tris, ((size_t)w->nF + 2) * 9 * sizeof(double), cudaMemcpyDeviceToHost));
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
    TWG_CHECK(c, !twg_is_multi(c), TWG_ERR_INVALID_ARG, "_dev entry points take a one-device handle (twg_winding_replica)");
    TWG_CHECK(c, nC <= 0x7fffffffull, TWG_ERR_INVALID_ARG, "at most 2^31-1 queries per device call (the host entry point chunks larger batches)");
    if (nC == 0) return 0;
    TWG_CUDA(c, cudaSetDevice(c->device));
    cudaStream_t st = pick(c, stream);
    if (w->nF == 0) {
        if (dW) TWG_CUDA(c, cudaMemsetAsync(dW, 0, nC * sizeof(double), st));
        if (dKeep) TWG_CUDA(c, cudaMemsetAsync(dKeep, 0, nC, st));
        return 0;
    }
    const uint32_t* perm = nullptr;
    twg_lane* lane = nullptr;
    TWG_TRY(twg_get_lane(c, st, &lane));
    if (w->sort_queries && nC > 32) TWG_TRY(twg_sort_points(c, lane, st, dC, nC, &perm, w->sort_box));
    const uint64_t ngroups = (nC + 31) / 32;
    unsigned grid = (unsigned)std::min<uint64_t>((ngroups + kWarps - 1) / kWarps, (uint64_t)c->sm_count * 32);
    if (c->opt.winding_minb >= 4) TWG_LAUNCH(c, winding_kernel<4>, grid, kWThreads, 0, st, w->view(), dC, perm, nC, dW, dKeep, c->dcounters);
    else TWG_LAUNCH(c, winding_kernel<3>, grid, kWThreads, 0, st, w->view(), dC, perm, nC, dW, dKeep, c->dcounters);
    return twg_lane_mark(c, lane);
}

int twg_winding_eval(twg_winding* w, const double* C, uint64_t nC, double* W, uint8_t* keep) {
    twg_ctx* c = w ? w->ctx : nullptr;
    TWG_CHECK(c, w && C && (W || keep), TWG_ERR_INVALID_ARG, "null argument");
    if (nC == 0) return 0;
    if (!w->replicas.empty()) {  // multi-device handle: contiguous index ranges, one per device (multi.cu)
        if (nC < TWG_MULTI_MIN_POINTS) return twg_forward0(c, twg_winding_eval(w->replicas[0], C, nC, W, keep));
        const uint64_t G = w->replicas.size();
        return twg_multi_run(c, [&](int k, twg_ctx*) {
            const uint64_t b = nC * (uint64_t)k / G, e = nC * (uint64_t)(k + 1) / G;
            return twg_winding_eval(w->replicas[k], C + 3 * b, e - b, W ? W + b : nullptr, keep ? keep + b : nullptr);
        });
    }
    TWG_CUDA(c, cudaSetDevice(c->device));
    const uint64_t chunk = 1ull << 23;  // 8 Mi queries: 192 MiB in per slot
    auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const uint64_t cmax = nC < chunk ? nC : chunk;
    const size_t qb = up(cmax * 24), wb = up(cmax * 8), kb = up(cmax);
    for (int k = 0; k < TWG_NUM_STREAMS; ++k) TWG_TRY(twg_ensure_scratch(c, k, qb + wb + kb));
    int slot = 0;
    for (uint64_t b = 0; b < nC; b += chunk, slot = (slot + 1) % TWG_NUM_STREAMS) {
        const uint64_t m = (nC - b < chunk) ? (nC - b) : chunk;
        cudaStream_t st = c->streams[slot];
        char* base = (char*)c->dscratch[slot];
        double* dQ = (double*)base;
        double* dW = (double*)(base + qb);
        uint8_t* dK = (uint8_t*)(base + qb + wb);
        TWG_CUDA(c, cudaMemcpyAsync(dQ, C + 3 * b, m * 24, cudaMemcpyHostToDevice, st));
        TWG_TRY(twg_winding_eval_dev(w, dQ, m, W ? dW : nullptr, keep ? dK : nullptr, st));
        if (W) TWG_CUDA(c, cudaMemcpyAsync(W + b, dW, m * 8, cudaMemcpyDeviceToHost, st));
        if (keep) TWG_CUDA(c, cudaMemcpyAsync(keep + b, dK, m, cudaMemcpyDeviceToHost, st));
    }
    for (int k = 0; k < TWG_NUM_STREAMS; ++k) TWG_CUDA(c, cudaStreamSynchronize(c->streams[k]));
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
