// envelope.cu -- batched envelope / nearest-facet queries against a twg_surface (C ABI: include/tetwild_gpu.h).
//
// Replaces LocalOperations::isPointOutEnvelop / isFaceOutEnvelop_sampling (src/tetwild/LocalOperations.cpp:1034-1109)
// and the tree queries behind them (src/tetwild/geogram/mesh_AABB.cpp:381-548).
//
// Kernels (details above each one)
//   env_points_kernel   persistent warps over groups of Morton-sorted queries: warp-cooperative group frontier, per-lane
//                       8-wide conservative FP32 steps, parked leaf tests run in rounds, early exit at the first facet
//                       within eps. The top kTopNodes pair records are staged once per CTA into shared memory by ONE 1-D
//                       bulk async copy (TMA, cp.async.bulk + mbarrier).
//   nearest_kernel      exact nearest facet / point / d2: best-first binary descent per lane, previous facet as first
//                       bound, FP32 box bounds, dynamic work claiming.
//   env_faces_kernel    one warp per candidate face: sampleTriangle's runs (sampling.cuh) flattened into equal chunks per
//                       lane, one candidate-facet collection per face, SCAN / FETCH / LEAF stage rounds, previous facet
//                       as hint like the reference (:1080-1086); the first OUT sample stops the warp.
#include "surface.cuh"
#include "sampling.cuh"

namespace {

// Pair records [0, kTopNodes) (the top of the heap) are staged in shared memory. Measured on B200 (10 M points, 200 k
// facets): 8 records 2.45, 64 records 2.49, 512 records (24 KiB) 2.37 G points/s -- the L1 already holds the hot top of
// the tree, and shared memory beyond a few KiB only costs occupancy. Default: 64 records = heap levels 0..5 = 3 KiB.
// (option env_top, 4 .. 512)
constexpr int kEnvThreads = 128;

__device__ __forceinline__ uint32_t stage_top(const SurfaceView& S, NodePair* top, uint64_t* bar) {
    const uint32_t topN = S.topN;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(bar, topN * (uint32_t)sizeof(NodePair));
        tma_bulk_g2s(top, S.pairs, topN * (uint32_t)sizeof(NodePair), bar);
    }
    mbar_wait(bar, 0);
    return topN;
}

// Point kernel: persistent warps, WARP-COOPERATIVE descent of the upper tree + per-lane finish, dynamic lane refill and
// batched leaf tests.
//   * perm (optional) is the Morton order of the batch (qsort.cu). A warp claims GROUPS of `group` consecutive sorted
//     queries from a global counter. For each group the 32 lanes together walk the upper levels ONCE: they reduce the
//     group's bounding box, dilate it by eps (outward-rounded floats) and refine a FRONTIER of heap nodes three levels
//     at a time -- lane j tests candidate box j against the dilated group box, survivors are compacted with a ballot into
//     a per-warp shared-memory list (node id + its float box) -- until the frontier's children are facets or it would
//     exceed kFrontMax nodes. Box-vs-group-box overlap is a superset of "some query of the group is within eps of the
//     box", so no subtree that any query needs is lost; a group that straddles a jump of the Morton curve merely keeps a
//     shallower frontier (worst case: the root level = the old per-lane traversal).
//   * a lane then starts its query from the frontier (one conservative FP32 box test per frontier node, warp-uniform loop
//     over shared memory), nearest node first, and continues with private 8-wide steps (surface.cuh); typically one or
//     two steps remain instead of ceil(depth/3).
//   * a lane that finishes its query immediately takes the next one of the group; the frontier is only read when a query
//     starts, so the warp moves on to the next group while slower lanes are still finishing;
//   * a lane that reaches facets parks them as `pending` instead of running the ~170-instruction point-triangle routine
//     on its own; the warp runs the routine when kLeafQuorum lanes are parked (or nothing else can advance), one facet
//     per parked lane per round. Early exit is per query, as in the reference (mesh_AABB.cpp:489).
constexpr int kFrontMaxDefault = 32;
constexpr int kLeafQuorumDefault = 16;
template <int FM>
struct FrontT {
    uint32_t node[FM];
    float box[FM][6];
};

// MINB: resident CTAs per SM the register allocation is capped for (8 -> 64 registers, 32 warps per SM); FM: frontier cap
// The per-lane stack holds kEnvStack subtree roots. A query needs at most FM (frontier) + 7 per remaining wide step; when
// eps is large against the facet size that can exceed the stack. A push that does not fit is NOT dropped silently: the lane
// raises `ovf`, and if such a query would end as OUT it is decided again by the exact binary descent twd::in_envelope
// (32-entry stack, enough for any 2^32-leaf heap), so the answer never depends on the stack size.
constexpr int kEnvStack = 64;
__device__ __noinline__ bool env_overflow_fallback(const SurfaceView& S, tw::V3 p, double eps2, const NodePair* top, uint32_t topN, unsigned long long* dbg) {
    atomicAdd(dbg + TWG_DBG_ENV_STACK_OVERFLOW, 1ull);
    uint32_t pos;
    return twd::in_envelope(S, p, eps2, pos, top, topN);
}

template <int MINB, int FM>
__global__ void __launch_bounds__(kEnvThreads, MINB) env_points_kernel(SurfaceView S, const double* __restrict__ P, const uint32_t* __restrict__ perm,
                                                                uint64_t n, double eps2, uint8_t* __restrict__ out, unsigned long long* counter, int group, int policy, int kLeafQuorum,
                                                                unsigned long long* dbg, int use_bound) {
    static_assert(FM <= kEnvStack, "the frontier of a group must fit the per-lane stack");
    extern __shared__ __align__(128) unsigned char smraw[];
    NodePair* top = reinterpret_cast<NodePair*>(smraw);
    __shared__ __align__(8) uint64_t bar;
    typedef FrontT<FM> Front;
    constexpr int kFrontMax = FM;
    __shared__ Front fronts[kEnvThreads / 32][2];
    const uint32_t topN = stage_top(S, top, &bar);
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const unsigned lt = (1u << lane) - 1u;
    const float thr = __double2float_ru(eps2);
    const float epsf = __fsqrt_ru(thr);
    const uint32_t leaf0 = S.nLeafP;
    const int L = 31 - __clz(leaf0);
    const uint32_t first = 1u << (L % 3);  // start level: a multiple of three levels above the facets
    const float inf = __int_as_float(0x7f800000);
    uint64_t cb = 0, ce = 0;  // unassigned part of the warp's current group (warp-uniform)
    bool more = true, active = false;
    const Front* fr = &fronts[wib][0];  // frontier of the current group (warp-uniform)
    int fcount = 0;
    uint32_t pend_mask = 0, pend_c0 = 0;
    tw::V3 p = tw::mk(0, 0, 0);
    twd::PointF q = twd::bracket(p);
    uint64_t src = 0;
    uint32_t stack[kEnvStack];
    int sp = 0;
    bool ovf = false;
    for (;;) {
        const unsigned need = __ballot_sync(full, !active);
        const unsigned act = ~need;
        const unsigned pend = __ballot_sync(full, active && pend_mask != 0);
        const int n_pend = __popc(pend), n_step = __popc(act & ~pend);
        const bool can_refill = need != 0 && (cb < ce || more);
        if (act == 0 && !can_refill) break;
        // Round scheduling: every round the warp runs ONE of three phases -- start queries on idle lanes (scan of the group's
        // frontier), leaf tests of parked lanes, or a traversal step of the rest -- and the lanes of the other two wait.
        // policy 1 picks the phase with the most lanes ready (a start scan is ~500 instructions: running it for one or two
        // lanes at a time is the costliest divergence of this kernel); policy 0 starts queries as soon as a lane is idle.
        const int n_need = can_refill ? __popc(need) : 0;
        const bool do_refill = n_need > 0 && (policy == 0 || act == 0 || (n_need >= n_pend && n_need >= n_step));
        if (do_refill) {
            if (cb >= ce && more) {
                unsigned long long c = 0;
                if (lane == 0) c = atomicAdd(counter, 1ull);
                c = __shfl_sync(full, c, 0);
                cb = c * (uint64_t)group;
                ce = (cb + group < n) ? cb + group : n;
                if (cb >= n) { more = false; cb = ce = 0; }
                if (cb < ce) {
                    // ---- group bounding box, dilated by eps (outward-rounded floats)
                    float lo[3] = {inf, inf, inf}, hi[3] = {-inf, -inf, -inf};
                    for (uint64_t j = cb + lane; j < ce; j += 32) {
#pragma unroll
                        for (int k = 0; k < 3; ++k) {
                            const double x = __ldg(P + 3 * j + k);
                            lo[k] = fminf(lo[k], __double2float_rd(x));
                            hi[k] = fmaxf(hi[k], __double2float_ru(x));
                        }
                    }
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) {
                            lo[k] = fminf(lo[k], __shfl_xor_sync(full, lo[k], o));
                            hi[k] = fmaxf(hi[k], __shfl_xor_sync(full, hi[k], o));
                        }
                        lo[k] = __fsub_rd(lo[k], epsf);
                        hi[k] = __fadd_ru(hi[k], epsf);
                    }
                    // ---- cooperative refinement of the frontier; the buffer not referenced by `fr` is free: lanes still
                    // working on the previous group never look at a frontier again
                    Front* A = const_cast<Front*>(fr == &fronts[wib][0] ? &fronts[wib][1] : &fronts[wib][0]);
                    Front* B = const_cast<Front*>(fr);
                    __syncwarp();
                    if ((uint32_t)lane < first) {
                        A->node[lane] = first + lane;
                        A->box[lane][0] = A->box[lane][1] = A->box[lane][2] = -inf;  // root-level boxes are not stored: always admitted
                        A->box[lane][3] = A->box[lane][4] = A->box[lane][5] = inf;
                    }
                    __syncwarp();
                    int count = (int)first;
                    for (int d = L % 3; d + 3 <= L - 3 && count > 0; d += 3) {
                        const int total = 8 * count;
                        int ncount = 0;
                        bool overflow = false;
                        for (int base = 0; base < total; base += 32) {
                            const int idx = base + lane;
                            bool ok = false;
                            uint32_t child = 0;
                            float b0 = 0, b1 = 0, b2 = 0, b3 = 0, b4 = 0, b5 = 0;
                            if (idx < total) {
                                child = 8u * A->node[idx >> 3] + (uint32_t)(idx & 7);
                                const float* rec = reinterpret_cast<const float*>(S.pairs + (child >> 1)) + 6 * (child & 1u);
                                const float2 u = __ldg(reinterpret_cast<const float2*>(rec));
                                const float2 v = __ldg(reinterpret_cast<const float2*>(rec) + 1);
                                const float2 w = __ldg(reinterpret_cast<const float2*>(rec) + 2);
                                b0 = u.x; b1 = u.y; b2 = v.x; b3 = v.y; b4 = w.x; b5 = w.y;
                                ok = b0 <= hi[0] && b3 >= lo[0] && b1 <= hi[1] && b4 >= lo[1] && b2 <= hi[2] && b5 >= lo[2];
                            }
                            const unsigned m = __ballot_sync(full, ok);
                            if (ncount + __popc(m) > kFrontMax) { overflow = true; break; }
                            if (ok) {
                                const int pos = ncount + __popc(m & lt);
                                B->node[pos] = child;
                                B->box[pos][0] = b0; B->box[pos][1] = b1; B->box[pos][2] = b2;
                                B->box[pos][3] = b3; B->box[pos][4] = b4; B->box[pos][5] = b5;
                            }
                            ncount += __popc(m);
                        }
                        __syncwarp();
                        if (overflow) break;
                        Front* t = A; A = B; B = t;
                        count = ncount;
                    }
                    fr = A;
                    fcount = count;
                }
            }
            if (cb < ce) {
                const uint64_t mine = cb + __popc(need & lt);
                if (!active && mine < ce) {
                    src = perm ? (uint64_t)__ldg(perm + mine) : mine;  // P is already in sorted order; only the result is scattered
                    p = tw::mk(__ldg(P + 3 * mine), __ldg(P + 3 * mine + 1), __ldg(P + 3 * mine + 2));
                    q = twd::bracket(p);
                    sp = 0;
                    ovf = false;
                    int best = -1;
                    float bd = 0.f;
                    for (int i = 0; i < fcount; ++i) {
                        const float d = twd::box_d2_lb(q, fr->box[i][0], fr->box[i][1], fr->box[i][2], fr->box[i][3], fr->box[i][4], fr->box[i][5]);
                        if (d <= thr) {
                            if (best < 0 || d < bd) { best = sp; bd = d; }
                            stack[sp++] = fr->node[i];
                        }
                    }
                    if (best >= 0 && best != sp - 1) { const uint32_t t = stack[best]; stack[best] = stack[sp - 1]; stack[sp - 1] = t; }  // nearest first
                    pend_mask = 0;
                    active = true;
                }
                const uint64_t adv = cb + __popc(need);
                cb = adv < ce ? adv : ce;
            }
            continue;
        }
        if (policy == 0 ? (n_pend >= kLeafQuorum || pend == act) : (n_pend >= kLeafQuorum || n_pend >= n_step)) {
            if (active && pend_mask != 0) {
                // the lane's next parked facet whose ORIENTED bound (surface.cuh::TriBound, ~20 FP64 instructions) is within eps: a
                // leaf box admits every facet whose axis-aligned extent comes within eps of the query, the plane of most of them
                // does not -- those never reach the ~170-instruction exact routine
                uint32_t pos = 0;
                bool have = false;
                while (pend_mask != 0) {
                    const int c = __ffs(pend_mask) - 1;
                    pend_mask &= pend_mask - 1;
                    pos = pend_c0 + (uint32_t)c - leaf0;
                    if (pos >= S.nF) continue;
                    if (!use_bound) { have = true; break; }
                    const TriBound tb = twd::load_bound(S.tb + pos);
                    if (twd::bound_lb2(tb, p) <= eps2 * twd::kSlack) { have = true; break; }
                }
                if (have) {
                    double s, t; tw::V3 nd; bool deg;
                    if (twd::facet_d2(S, pos, p, s, t, nd, deg) <= eps2) { out[src] = 0; active = false; pend_mask = 0; }
                }
            }
        } else if (active && pend_mask == 0) {
            if (sp == 0) {
                out[src] = (ovf && env_overflow_fallback(S, p, eps2, top, topN, dbg)) ? 0 : 1;
                active = false;
            } else {
                const uint32_t node = stack[--sp];
                if (node >= leaf0) {  // trees with fewer than three levels
                    pend_mask = 1u; pend_c0 = node;
                } else {
                    float d[8];
                    const uint32_t mask = twd::wide_step(S, q, thr, node, top, topN, d);
                    const uint32_t c0 = 8u * node;
                    if (c0 >= leaf0) {
                        pend_mask = mask; pend_c0 = c0;
                    } else {
                        int best = -1;
                        float bd = 0.f;
#pragma unroll
                        for (int c = 0; c < 8; ++c)
                            if (((mask >> c) & 1u) && (best < 0 || d[c] < bd)) { best = c; bd = d[c]; }
#pragma unroll
                        for (int c = 0; c < 8; ++c)
                            if (((mask >> c) & 1u) && c != best) {
                                if (sp < kEnvStack - 1) stack[sp++] = c0 + (uint32_t)c;  // one slot stays free for the nearest child
                                else ovf = true;
                            }
                        if (best >= 0) stack[sp++] = c0 + (uint32_t)best;  // nearest subtree is popped first
                    }
                }
            }
        }
    }
}

constexpr int kNearRun = 8;  // consecutive sorted queries per lane and claim
__global__ void __launch_bounds__(kEnvThreads, 8) nearest_kernel(SurfaceView S, const double* __restrict__ P, const uint32_t* __restrict__ perm, uint64_t n,
                                                             uint32_t* __restrict__ facet, double* __restrict__ nearest, double* __restrict__ d2out,
                                                             unsigned long long* counter) {
    extern __shared__ __align__(128) unsigned char smraw[];
    NodePair* top = reinterpret_cast<NodePair*>(smraw);
    __shared__ __align__(8) uint64_t bar;
    const uint32_t topN = stage_top(S, top, &bar);
    // every lane walks a RUN of consecutive Morton-sorted queries and starts each search from the facet nearest to
    // its previous query (nearest_facet_with_hint, mesh_AABB.h:162-176 / get_nearest_facet_hint :381-416 play the same
    // role): the initial bound is already within a facet or two of the answer, so far queries no longer open every
    // box on the way down. The result is the minimum over all facets either way.
    // Warps claim 32 x kNearRun queries at a time from a global counter: far queries cost 10-100x a near one and come
    // in spatial clusters of the sorted order, so a static split left most of the machine idle behind a few long runs
    // (ncu: 13 % warps active).
    const int lane = threadIdx.x & 31;
    for (;;) {
    unsigned long long claim = 0;
    if (lane == 0) claim = atomicAdd(counter, 1ull);
    claim = __shfl_sync(0xffffffffu, claim, 0);
    const uint64_t cb = claim * (uint64_t)(32 * kNearRun);
    if (cb >= n) break;
    const uint64_t b0 = cb + (uint64_t)lane * kNearRun, e0 = (b0 + kNearRun < n) ? b0 + kNearRun : n;
    uint32_t hint = TWG_NO_FACET;
    for (uint64_t j = b0; j < e0; ++j) {
        const uint64_t i = perm ? (uint64_t)__ldg(perm + j) : j;
        const tw::V3 p = tw::mk(__ldg(P + 3 * i), __ldg(P + 3 * i + 1), __ldg(P + 3 * i + 2));
        twd::Nearest b;
        b.d2 = DBL_MAX; b.s = b.t = 0.0; b.pos = 0; b.deg = false; b.pt_deg = p;
        if (hint != TWG_NO_FACET) {
            b.pos = hint;
            b.d2 = twd::facet_d2(S, hint, p, b.s, b.t, b.pt_deg, b.deg);
        }
        twd::nearest_facet(S, p, b, top, topN);
        hint = b.pos;
        if (d2out) d2out[i] = b.d2;
        if (facet || nearest) {
            const tw::TriRec r = twd::load_tri(S.tris + b.pos);
            if (facet) facet[i] = r.facet;
            if (nearest) {
                const tw::V3 q = b.deg ? b.pt_deg : tw::tri_nearest_point(r, b.s, b.t);
                nearest[3 * i] = q.x; nearest[3 * i + 1] = q.y; nearest[3 * i + 2] = q.z;
            }
        }
    }
    }
}

// The tiny call of the sequential scheduler (ONE point: EdgeCollapser.cpp:312,322; VertexSmoother.cpp:358; n <= 64): the points
// travel in the kernel's parameters, one lane per point walks the tree on its own (in_envelope / nearest_facet, the routines the
// face kernel and the per-lane nearest kernel use: same exact predicates, same answers), results go straight into the mapped
// slab and the kernel raises the completion word itself -- ONE driver call per host call (the batched path needs a counter reset,
// the kernel and a signal launch).
struct TwgTinyPoints {
    double xyz[192];
};
__global__ void __launch_bounds__(64) points_tiny_kernel(SurfaceView S, const __grid_constant__ TwgTinyPoints in, uint32_t n, int what, double eps2,
                                                         uint8_t* __restrict__ out, uint32_t* __restrict__ facet, double* __restrict__ nearest,
                                                         double* __restrict__ d2out, twg_done done) {
    const uint32_t i = threadIdx.x;
    if (i < n) {
        const tw::V3 p = tw::mk(in.xyz[3 * i], in.xyz[3 * i + 1], in.xyz[3 * i + 2]);
        if (what == 0) {
            uint32_t pos;
            out[i] = twd::in_envelope(S, p, eps2, pos, nullptr, 0u) ? 0 : 1;
        } else {
            twd::Nearest b;
            b.d2 = DBL_MAX; b.s = b.t = 0.0; b.pos = 0; b.deg = false; b.pt_deg = p;
            twd::nearest_facet(S, p, b, nullptr, 0u);
            if (d2out) d2out[i] = b.d2;
            if (facet || nearest) {
                const tw::TriRec r = twd::load_tri(S.tris + b.pos);
                if (facet) facet[i] = r.facet;
                if (nearest) {
                    const tw::V3 q = b.deg ? b.pt_deg : tw::tri_nearest_point(r, b.s, b.t);
                    nearest[3 * i] = q.x; nearest[3 * i + 1] = q.y; nearest[3 * i + 2] = q.z;
                }
            }
        }
    }
    twg_signal_done(done);
}

// Exact nearest facet for large sorted batches (nearest_facet, mesh_AABB.cpp:418-480): two warp-cooperative forms, no per-thread
// stack anywhere (round 1: one descent per lane with two 32-entry stacks in local memory -- 9.7 active lanes per instruction,
// 232 M local-memory accesses per 10 M points).
//
// (1) PACKET traversal -- queries near the surface. A warp owns 32 consecutive Morton-sorted queries and walks the tree ONCE for
//     all of them with one shared stack (shared memory): a node is entered iff SOME lane still needs it (conservative FP32 box
//     bound <= that lane's current best), every lane tests the same pair record (one broadcast load, no divergence), the
//     nearer child by majority goes first. Queries that are close in space need nearly the same nodes (|d_i - d_j| <=
//     |p_i - p_j|), so the union costs little more than one traversal.
// (2) ONE QUERY PER WARP -- far queries. A far query's candidates are the ~10^2 leaves whose boxes intrude the sphere of its
//     nearest distance (a cap of radius sqrt(2 d h)); 32 far queries of one packet lie ~0.02 apart (their density is low),
//     their caps barely overlap, and a packet would run the exact point-triangle routine once per (leaf, interested lane).
//     So when a packet has not finished within `budget` node visits, its queries are finished one at a time by the whole
//     warp: the 32 lanes test 32 DIFFERENT boxes -- the eight children of four nodes popped from a shared stack -- against
//     the SAME query, survivors are compacted back with a ballot, and at the leaf level every lane owns one facet:
//     oriented bound, then the exact routine for the few that remain, then a warp minimum. The packet phase is not
//     wasted: each query starts from the (already near-exact) bound it reached there.
// At the leaves the oriented bound (surface.cuh::TriBound, ~20 FP64 instructions) decides whether the exact ~170-instruction
// routine runs at all. Each lane starts from the facet that answered the same lane of the previous packet of its chunk
// (nearest_facet_with_hint, mesh_AABB.h:162-176). The result is the exact minimum over all facets (same d2 bits as brute
// force; ties between equidistant facets go to the one met first).
constexpr int kPacketChunk = 4;   // consecutive packets per claim: three of four start from a neighbour's facet
constexpr int kPacketStack = 40;  // one sibling per level of a heap with at most 2^32 leaves
constexpr int kWarpStack = 1536;  // one-query-per-warp form: nodes pending on the shared stack (32 lanes x up to 8 survivors per sweep step)
constexpr int kWarpLeaves = 512;  // ... and facets waiting for their batch

struct PacketBest {
    double d2, s, t;
    uint32_t pos;
};

__device__ __forceinline__ void packet_leaf(const SurfaceView& S, uint32_t pos, bool want, tw::V3 p, PacketBest& b) {
    // `want`: this lane's box bound admits the leaf. Oriented bound first, then the exact routine for the lanes that remain.
    if (pos >= S.nF) return;  // padding leaf (warp-uniform)
    bool need = false;
    if (want) {
        const TriBound tb = twd::load_bound(S.tb + pos);
        need = twd::bound_lb2(tb, p) <= b.d2 * twd::kSlack;
    }
    if (__ballot_sync(0xffffffffu, need) == 0u) return;
    if (need) {
        double s, t; tw::V3 nd; bool deg;
        const double d2 = twd::facet_d2(S, pos, p, s, t, nd, deg);
        if (d2 < b.d2) { b.d2 = d2; b.s = s; b.t = t; b.pos = pos; }
    }
}

__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// form (2): the whole warp finishes ONE query. p and `b` are warp-uniform on entry (b: the bound reached so far, a real facet);
// on exit b is the exact nearest facet, warp-uniform. Needs a heap of at least three levels (nLeafP >= 8).
//   phase 0  greedy descent: lanes 0..7 bound the eight descendants of the current node, the nearest one is entered, down to
//            eight facets that are tested exactly -- the bound is now within a facet size of the answer, which is what keeps the
//            sweep small (the candidate cap grows with the square root of the slack);
//   phase 1  sweep: every lane pops ONE node and bounds its eight descendants (one 192-byte run, twd::wide_step), the survivors
//            of all lanes are compacted onto the shared stack with one warp scan; survivors on the facet level go to a leaf
//            list instead;
//   phase 2  whenever 32 facets are listed (or nothing else is left): one facet per lane, oriented bound, exact routine for
//            the lanes that remain, warp minimum.
struct WarpScratch {
    uint32_t stack[kWarpStack];
    uint32_t leaves[kWarpLeaves];
};

__device__ __forceinline__ void warp_leaf_batch(const SurfaceView& S, tw::V3 p, uint32_t pos, bool on, PacketBest& mine, double& best, float& thr) {
    const unsigned full = 0xffffffffu;
    bool need = false;
    if (on && pos < S.nF) {
        const TriBound tb = twd::load_bound(S.tb + pos);
        need = twd::bound_lb2(tb, p) <= best * twd::kSlack;
    }
    if (__ballot_sync(full, need) == 0u) return;
    bool improved = false;
    if (need) {
        double s_, t_; tw::V3 nd; bool deg;
        const double d2 = twd::facet_d2(S, pos, p, s_, t_, nd, deg);
        if (d2 < mine.d2) { mine.d2 = d2; mine.s = s_; mine.t = t_; mine.pos = pos; improved = d2 < best; }
    }
    if (__ballot_sync(full, improved)) {
        best = warp_min(mine.d2);
        thr = __double2float_ru(best * twd::kSlack);
    }
}

__device__ __noinline__ void warp_query_nearest(const SurfaceView& S, tw::V3 p, PacketBest& b, WarpScratch* ws, const NodePair* top, uint32_t topN, int lane) {
    const unsigned full = 0xffffffffu;
    const twd::PointF q = twd::bracket(p);
    const uint32_t leaf0 = S.nLeafP;
    const int L = 31 - __clz(leaf0);
    const uint32_t first = 1u << (L % 3);
    uint32_t* stack = ws->stack;
    uint32_t* leaves = ws->leaves;
    // lane-local candidate (the incumbent lives in lane 0), warp-uniform pruning bound
    PacketBest mine;
    mine.d2 = (lane == 0) ? b.d2 : DBL_MAX; mine.s = b.s; mine.t = b.t; mine.pos = b.pos;
    double best = b.d2;
    float thr = (best < 1e37) ? __double2float_ru(best * twd::kSlack) : __int_as_float(0x7f7fffff);
    // ---- phase 0: greedy descent
    {
        uint32_t node = 0;  // 0: the (virtual) parent of the `first` root-level nodes
        for (;;) {
            const uint32_t c0 = node ? 8u * node : first;
            const uint32_t nchild = node ? 8u : first;
            if (c0 >= leaf0) {  // eight facets: exact tests, no bound needed
                warp_leaf_batch(S, p, c0 + (uint32_t)lane - leaf0, (uint32_t)lane < nchild, mine, best, thr);
                break;
            }
            unsigned long long keyv = ~0ull;
            if ((uint32_t)lane < nchild) {
                const uint32_t child = c0 + (uint32_t)lane;
                const float* rec = reinterpret_cast<const float*>(S.pairs + (child >> 1)) + 6 * (child & 1u);
                const float2 u = __ldg(reinterpret_cast<const float2*>(rec));
                const float2 v = __ldg(reinterpret_cast<const float2*>(rec) + 1);
                const float2 w = __ldg(reinterpret_cast<const float2*>(rec) + 2);
                const float d = twd::box_d2_lb(q, u.x, u.y, v.x, v.y, w.x, w.y);  // >= 0, +inf for padding: bit pattern is order preserving
                keyv = ((unsigned long long)__float_as_uint(d) << 32) | child;
            }
#pragma unroll
            for (int o = 4; o > 0; o >>= 1) {
                const unsigned long long other = __shfl_xor_sync(full, keyv, o);
                keyv = other < keyv ? other : keyv;
            }
            keyv = __shfl_sync(full, keyv, 0);
            node = (uint32_t)(keyv & 0xffffffffu);
        }
    }
    // ---- phase 1 + 2: sweep from the root level with the tight bound
    __syncwarp();
    if ((uint32_t)lane < first) stack[lane] = first + (uint32_t)lane;
    int sp = (int)first, lc = 0;
    __syncwarp();
    while (sp > 0 || lc > 0) {
        if (sp > 0) {
            int take = sp < 32 ? sp : 32;
            const int room = (kWarpStack - sp) / 7;  // every popped node frees one slot and may push eight
            if (take > room) take = room > 0 ? room : 1;
            const bool on = lane < take;
            const uint32_t node = on ? stack[sp - take + lane] : 0u;
            sp -= take;
            __syncwarp();
            uint32_t mask = 0;
            const uint32_t c0 = 8u * node;
            if (on) {
                float d[8];
                mask = twd::wide_step(S, q, thr, node, top, topN, d);
            }
            const bool leaf_level = c0 >= leaf0;
            // one warp scan for both destinations: low half = children pushed back, high half = facets listed
            const uint32_t mine_cnt = on ? (leaf_level ? ((uint32_t)__popc(mask) << 16) : (uint32_t)__popc(mask)) : 0u;
            uint32_t incl = mine_cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(full, incl, o);
                if (lane >= o) incl += t;
            }
            const uint32_t tot = __shfl_sync(full, incl, 31);
            const uint32_t excl = incl - mine_cnt;
            if (sp + (int)(tot & 0xffffu) > kWarpStack || lc + (int)(tot >> 16) > kWarpLeaves) {
                // (cannot happen with the room / flush rules below for heaps of up to 2^31 leaves; kept as a hard stop)
                best = -1.0;
                break;
            }
            if (on && mask) {
                uint32_t* dst = leaf_level ? leaves + lc + (excl >> 16) : stack + sp + (excl & 0xffffu);
                const uint32_t bias = leaf_level ? leaf0 : 0u;
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    if ((mask >> c) & 1u) *dst++ = c0 + (uint32_t)c - bias;
            }
            sp += (int)(tot & 0xffffu);
            lc += (int)(tot >> 16);
            __syncwarp();
        }
        // facets: a full batch at a time while the sweep goes on, everything once the stack is empty or the list is nearly full
        while (lc >= 32 || (lc > 0 && (sp == 0 || lc > kWarpLeaves - 256))) {
            const int take = lc < 32 ? lc : 32;
            const bool on = lane < take;
            const uint32_t pos = on ? leaves[lc - take + lane] : 0u;
            lc -= take;
            __syncwarp();
            warp_leaf_batch(S, p, pos, on, mine, best, thr);
        }
    }
    if (best < 0.0) {  // hard stop above: finish with the exact per-lane search (lane 0), never with a partial answer
        twd::Nearest nb;
        nb.d2 = mine.d2; nb.s = mine.s; nb.t = mine.t; nb.pos = mine.pos; nb.deg = false; nb.pt_deg = p;
        if (lane == 0) {
            nb.d2 = b.d2; nb.s = b.s; nb.t = b.t; nb.pos = b.pos;
            twd::nearest_facet(S, p, nb, top, topN);
            mine.d2 = nb.d2; mine.s = nb.s; mine.t = nb.t; mine.pos = nb.pos;
        } else {
            mine.d2 = DBL_MAX;
        }
        best = __shfl_sync(full, mine.d2, 0);
    }
    // the winner: smallest d2, lowest lane among equals (deterministic)
    const unsigned win = __ballot_sync(full, mine.d2 == best);
    const int src = __ffs(win) - 1;
    b.d2 = best;
    b.s = __shfl_sync(full, mine.s, src);
    b.t = __shfl_sync(full, mine.t, src);
    b.pos = __shfl_sync(full, mine.pos, src);
}

__global__ void __launch_bounds__(kEnvThreads, 6) nearest_packet_kernel(SurfaceView S, const double* __restrict__ Ps /*sorted*/, const uint32_t* __restrict__ perm,
                                                                    uint64_t n, uint32_t* __restrict__ facet, double* __restrict__ nearest,
                                                                    double* __restrict__ d2out, unsigned long long* counter, int budget) {
    extern __shared__ __align__(128) unsigned char smraw[];
    NodePair* top = reinterpret_cast<NodePair*>(smraw);
    __shared__ __align__(8) uint64_t bar;
    __shared__ WarpScratch scratch[kEnvThreads / 32];
    const uint32_t topN = stage_top(S, top, &bar);
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    uint32_t* stack = scratch[wib].stack;
    const uint32_t leaf0 = S.nLeafP;
    const uint64_t npackets = (n + 31) / 32;
    for (;;) {
        unsigned long long claim = 0;
        if (lane == 0) claim = atomicAdd(counter, 1ull);
        claim = __shfl_sync(full, claim, 0);
        const uint64_t g0 = claim * (uint64_t)kPacketChunk;
        if (g0 >= npackets) break;
        uint32_t hint = TWG_NO_FACET;
        for (uint64_t g = g0; g < g0 + kPacketChunk && g < npackets; ++g) {
            const uint64_t j = g * 32 + lane;
            const bool valid = j < n;
            const uint64_t jj = valid ? j : n - 1;
            const tw::V3 p = tw::mk(__ldg(Ps + 3 * jj), __ldg(Ps + 3 * jj + 1), __ldg(Ps + 3 * jj + 2));
            const twd::PointF q = twd::bracket(p);
            PacketBest b;
            b.d2 = DBL_MAX; b.s = b.t = 0.0; b.pos = 0;
            if (hint != TWG_NO_FACET) {
                tw::V3 nd; bool deg;
                b.pos = hint;
                b.d2 = twd::facet_d2(S, hint, p, b.s, b.t, nd, deg);
            }
            float thr = (b.d2 < 1e37) ? __double2float_ru(b.d2 * twd::kSlack) : __int_as_float(0x7f7fffff);  // box bounds above this cannot hold a nearer facet
            int sp = 0, visits = 0;
            uint32_t node = 1;
            bool finished = true;
            for (;;) {
                if (++visits > budget) { finished = false; break; }
                const NodePair np = (node < topN) ? top[node] : twd::load_pair(S.pairs + node);
                const float dl = twd::box_d2_lb(q, np.a.x, np.a.y, np.a.z, np.a.w, np.b.x, np.b.y);
                const float dr = twd::box_d2_lb(q, np.b.z, np.b.w, np.c.x, np.c.y, np.c.z, np.c.w);
                const bool wl = valid && dl <= thr, wr = valid && dr <= thr;  // empty (padding) boxes give +inf
                const unsigned bl = __ballot_sync(full, wl), br = __ballot_sync(full, wr);
                const uint32_t cl = 2u * node;
                const bool lfirst = __popc(__ballot_sync(full, dl <= dr)) >= 16;
                if (cl >= leaf0) {
                    const uint32_t pl = cl - leaf0;
                    const double before = b.d2;
#pragma unroll 1
                    for (int k = 0; k < 2; ++k) {
                        const bool left = (k == 0) == lfirst;
                        if (left ? bl : br) packet_leaf(S, left ? pl : pl + 1u, (left ? wl : wr) && (double)(left ? dl : dr) <= b.d2 * twd::kSlack, p, b);
                    }
                    if (b.d2 != before) thr = __double2float_ru(b.d2 * twd::kSlack);
                } else if (bl | br) {
                    if (bl && br) {
                        if (lane == 0) stack[sp] = lfirst ? cl + 1u : cl;
                        ++sp;
                        node = lfirst ? cl : cl + 1u;
                    } else {
                        node = bl ? cl : cl + 1u;
                    }
                    continue;
                }
                if (sp == 0) break;
                __syncwarp();
                node = stack[--sp];
                __syncwarp();
            }
            if (!finished) {
                // form (2): the packet's queries one at a time, each by the whole warp, from the bound it has reached
                if (b.d2 == DBL_MAX) {  // (a heap deeper than the budget) start every query from facet 0
                    tw::V3 nd; bool deg;
                    b.pos = 0;
                    b.d2 = twd::facet_d2(S, 0u, p, b.s, b.t, nd, deg);
                }
                const unsigned vmask = __ballot_sync(full, valid);
                for (int k = 0; k < 32; ++k) {
                    if (!((vmask >> k) & 1u)) continue;
                    const tw::V3 pk = tw::mk(__shfl_sync(full, p.x, k), __shfl_sync(full, p.y, k), __shfl_sync(full, p.z, k));
                    PacketBest bk;
                    bk.d2 = __shfl_sync(full, b.d2, k); bk.s = __shfl_sync(full, b.s, k); bk.t = __shfl_sync(full, b.t, k); bk.pos = __shfl_sync(full, b.pos, k);
                    warp_query_nearest(S, pk, bk, &scratch[wib], top, topN, lane);
                    if (lane == k) b = bk;
                }
            }
            // ---- results (scattered to the caller's order)
            if (valid) {
                const uint64_t i = perm ? (uint64_t)__ldg(perm + j) : j;
                if (d2out) d2out[i] = b.d2;
                if (facet || nearest) {
                    const tw::TriRec r = twd::load_tri(S.tris + b.pos);
                    if (facet) facet[i] = r.facet;
                    if (nearest) {
                        tw::V3 pt;
                        if (r.flags & 1u) {  // degenerate facet: the three-segment routine returns the point itself
                            double tv[9];
#pragma unroll
                            for (int k = 0; k < 9; ++k) tv[k] = __ldg(S.triV + (size_t)b.pos * 9 + k);
                            tw::tri_sqdist_degenerate(p, tv, pt);
                        } else {
                            pt = tw::tri_nearest_point(r, b.s, b.t);
                        }
                        nearest[3 * i] = pt.x; nearest[3 * i + 1] = pt.y; nearest[3 * i + 2] = pt.z;
                    }
                }
            }
            hint = b.pos;
        }
    }
}

// Exact nearest facet, form (3): ROUND-SCHEDULED lanes -- the machinery of env_points_kernel applied to the full search. One
// query per lane, but no lane ever runs the long routines alone:
//   * persistent warps claim groups of consecutive Morton-sorted queries; a lane that finishes takes the next query of the
//     group at once, so a far query (10-100x the work of a near one) delays nobody;
//   * every round the warp runs ONE phase for all lanes that are ready for it: REFILL (start queries on idle lanes), LEAF (one
//     exact point-triangle test per parked lane) or STEP (pop a subtree, bound its eight descendants with one 192-byte run) --
//     the phase with the most lanes ready;
//   * a lane that reaches facets parks them; in the LEAF phase it first drops parked facets whose ORIENTED bound
//     (tw_math.cuh::TriBound) cannot beat its current best, and only runs the exact routine on the rest (~1 in 10 for far
//     queries);
//   * the per-lane stack holds (node, lower bound) pairs and is re-checked when popped, nearest descendant on top; the first
//     thing a query tests is the facet that answered the lane's previous query (nearest_facet_with_hint, mesh_AABB.h:162-176),
//     so the search starts with a bound that is already within a facet or two of the answer.
// Unlike the packet form, a lane only ever looks at ITS OWN candidates: far queries of one warp lie ~0.02 apart (their density
// is low) and share few of them. The stack cannot overflow for heaps of up to 2^24 leaves (7 per 8-wide level + the root level);
// deeper heaps that do are finished by the exact binary descent twd::nearest_facet. Result: the exact minimum over all facets.
constexpr int kNrStack = 64;

template <int MINB>
__global__ void __launch_bounds__(kEnvThreads, MINB) nearest_rounds_kernel(SurfaceView S, const double* __restrict__ Ps /*sorted*/, const uint32_t* __restrict__ perm,
                                                                       uint64_t n, uint32_t* __restrict__ facet, double* __restrict__ nearest,
                                                                       double* __restrict__ d2out, unsigned long long* counter, int group, int quorum,
                                                                       unsigned long long* dbg /*NULL unless option trace: work counters 3..6*/) {
    extern __shared__ __align__(128) unsigned char smraw[];
    NodePair* top = reinterpret_cast<NodePair*>(smraw);
    __shared__ __align__(8) uint64_t bar;
    const uint32_t topN = stage_top(S, top, &bar);
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    const uint32_t leaf0 = S.nLeafP;
    const int L = 31 - __clz(leaf0);
    const uint32_t first = 1u << (L % 3);
    const float fmax_ = __int_as_float(0x7f7fffff);
    uint32_t n_steps = 0, n_bounds = 0, n_exact = 0, n_rounds = 0;
    uint64_t cb = 0, ce = 0;
    bool more = true, active = false, hint_pending = false;
    uint32_t pend_mask = 0, pend_c0 = 0, hint = TWG_NO_FACET;
    tw::V3 p = tw::mk(0, 0, 0);
    twd::PointF q = twd::bracket(p);
    uint64_t src = 0;
    double bd2 = DBL_MAX, bs = 0.0, bt = 0.0;
    uint32_t bpos = 0;
    float thr = fmax_;
    uint32_t stk[kNrStack];
    float stl[kNrStack];
    int sp = 0;
    for (;;) {
        const unsigned need = __ballot_sync(full, !active);
        const unsigned act = ~need;
        const unsigned pend = __ballot_sync(full, active && (pend_mask != 0 || hint_pending));
        const int n_pend = __popc(pend), n_step = __popc(act & ~pend);
        const bool can_refill = need != 0 && (cb < ce || more);
        if (act == 0 && !can_refill) break;
        const int n_need = can_refill ? __popc(need) : 0;
        ++n_rounds;
        if (n_need > 0 && (act == 0 || (n_need >= n_pend && n_need >= n_step))) {
            // ---- REFILL
            if (cb >= ce && more) {
                unsigned long long c = 0;
                if (lane == 0) c = atomicAdd(counter, 1ull);
                c = __shfl_sync(full, c, 0);
                cb = c * (uint64_t)group;
                ce = (cb + group < n) ? cb + group : n;
                if (cb >= n) { more = false; cb = ce = 0; }
            }
            if (cb < ce) {
                const uint64_t mine = cb + __popc(need & lt);
                if (!active && mine < ce) {
                    src = mine;
                    p = tw::mk(__ldg(Ps + 3 * mine), __ldg(Ps + 3 * mine + 1), __ldg(Ps + 3 * mine + 2));
                    q = twd::bracket(p);
                    bd2 = DBL_MAX; bs = bt = 0.0; bpos = 0; thr = fmax_;
                    sp = 0;
                    for (uint32_t k = 0; k < first; ++k) { stk[sp] = first + k; stl[sp] = 0.f; ++sp; }
                    pend_mask = 0;
                    hint_pending = hint != TWG_NO_FACET;
                    active = true;
                }
                const uint64_t adv = cb + __popc(need);
                cb = adv < ce ? adv : ce;
            }
            continue;
        }
        if (n_pend >= quorum || n_pend >= n_step) {
            // ---- LEAF: one exact test per parked lane; facets whose oriented bound cannot win are dropped first
            if (active && (hint_pending || pend_mask != 0)) {
                uint32_t pos = 0;
                bool have = false;
                if (hint_pending) {
                    pos = hint; hint_pending = false; have = true;
                } else {
                    while (pend_mask != 0) {
                        const int c = __ffs(pend_mask) - 1;
                        pend_mask &= pend_mask - 1;
                        pos = pend_c0 + (uint32_t)c - leaf0;
                        if (pos >= S.nF) continue;
                        const TriBound tb = twd::load_bound(S.tb + pos);
                        ++n_bounds;
                        if (twd::bound_lb2(tb, p) <= bd2 * twd::kSlack) { have = true; break; }
                    }
                }
                if (have) {
                    double s_, t_; tw::V3 nd; bool deg;
                    const double d2 = twd::facet_d2(S, pos, p, s_, t_, nd, deg);
                    ++n_exact;
                    if (d2 < bd2) {
                        bd2 = d2; bs = s_; bt = t_; bpos = pos;
                        thr = (bd2 < 1e37) ? __double2float_ru(bd2 * twd::kSlack) : fmax_;
                    }
                }
            }
            continue;
        }
        // ---- STEP
        if (active && pend_mask == 0 && !hint_pending) {
            uint32_t node = 0;
            bool got = false;
            while (sp > 0) {
                --sp;
                if (stl[sp] <= thr) { node = stk[sp]; got = true; break; }
            }
            if (!got) {
                // ---- the query is finished: results to the caller's position
                const uint64_t i = perm ? (uint64_t)__ldg(perm + src) : src;
                if (d2out) d2out[i] = bd2;
                if (facet || nearest) {
                    const tw::TriRec r = twd::load_tri(S.tris + bpos);
                    if (facet) facet[i] = r.facet;
                    if (nearest) {
                        tw::V3 pt;
                        if (r.flags & 1u) {
                            double tv[9];
#pragma unroll
                            for (int k = 0; k < 9; ++k) tv[k] = __ldg(S.triV + (size_t)bpos * 9 + k);
                            tw::tri_sqdist_degenerate(p, tv, pt);
                        } else {
                            pt = tw::tri_nearest_point(r, bs, bt);
                        }
                        nearest[3 * i] = pt.x; nearest[3 * i + 1] = pt.y; nearest[3 * i + 2] = pt.z;
                    }
                }
                hint = bpos;
                active = false;
            } else {
                float d[8];
                const uint32_t mask = twd::wide_step(S, q, thr, node, top, topN, d);
                ++n_steps;
                const uint32_t c0 = 8u * node;
                if (c0 >= leaf0) {
                    pend_mask = mask; pend_c0 = c0;
                } else if (mask != 0) {
                    if (sp + 8 > kNrStack) {
                        // (heaps deeper than 2^24 leaves only) finish this query with the exact binary descent
                        twd::Nearest nb;
                        nb.d2 = bd2; nb.s = bs; nb.t = bt; nb.pos = bpos; nb.deg = false; nb.pt_deg = p;
                        twd::nearest_facet(S, p, nb, top, topN);
                        bd2 = nb.d2; bs = nb.s; bt = nb.t; bpos = nb.pos;
                        sp = 0;
                    } else {
                        int best = -1;
                        float bdm = 0.f;
#pragma unroll
                        for (int c = 0; c < 8; ++c)
                            if (((mask >> c) & 1u) && (best < 0 || d[c] < bdm)) { best = c; bdm = d[c]; }
#pragma unroll
                        for (int c = 7; c >= 0; --c)
                            if (((mask >> c) & 1u) && c != best) { stk[sp] = c0 + (uint32_t)c; stl[sp] = d[c]; ++sp; }
                        stk[sp] = c0 + (uint32_t)best; stl[sp] = bdm; ++sp;  // nearest descendant is popped first
                    }
                }
            }
        }
    }
    if (dbg) {  // diagnostics (option trace): 3 = 8-wide steps, 4 = oriented bounds, 5 = exact tests, 6 = warp rounds
        atomicAdd(dbg + 3, (unsigned long long)n_steps);
        atomicAdd(dbg + 4, (unsigned long long)n_bounds);
        atomicAdd(dbg + 5, (unsigned long long)n_exact);
        if (lane == 0) atomicAdd(dbg + 6, (unsigned long long)n_rounds);
    }
}

// one sample of isFaceOutEnvelop_sampling (:1079-1093): hint facet first, then the tree. Returns true if OUT.
__device__ __forceinline__ bool sample_out(const SurfaceView& S, tw::V3 p, double eps2, uint32_t& prev, const NodePair* top, uint32_t topN) {
    if (prev != TWG_NO_FACET) {
        double s, t; tw::V3 nd; bool deg;
        if (twd::facet_d2(S, prev, p, s, t, nd, deg) <= eps2) return false;
    }
    uint32_t pos;
    if (twd::in_envelope(S, p, eps2, pos, top, topN)) { prev = pos; return false; }
    return true;
}

// Face kernel: one warp per candidate face, the samples of sampleTriangle FLATTENED and dealt out in equal contiguous
// chunks. The sample set is a list of runs (sampling.cuh): base edge, the two single vertices, the two other edges, and one
// run per lattice row. The lanes build the run table together (lane j plans row base+j: the row's first sample and its
// length, with the reference's own int truncations), a warp scan turns the lengths into offsets, and lane l then walks
// samples [l*c, (l+1)*c) of the concatenation, c = ceil(S / 32): every lane gets the same number of samples whatever the
// shape of the triangle (rows of a triangle shrink towards the apex; one lane per row left two thirds of the warp idle),
// and consecutive samples of a lane are neighbours on the face, so the previous facet is the right hint
// (LocalOperations.cpp:1080-1086). The first OUT sample raises a warp-shared flag and the warp stops; the answer does not
// depend on the order in which samples are visited (the reference starts from the middle sample only to exit earlier).
// Candidate facets of one face. Every sample of a face lies in the face's bounding box, so a facet within eps of ANY sample
// has its (outward-rounded) leaf box inside that box dilated by eps. The warp therefore walks the tree ONCE per face --
// the same 8-wide cooperative refinement as the point kernel's group frontier, continued down to the leaves: lane j tests
// candidate box j against the dilated face box, survivors are compacted with a ballot -- and a sample only tests the
// facets of that list (FP32 lower bound of the box distance first, the exact point-triangle routine for the boxes within
// eps). One traversal per face instead of one per sample, and the per-sample work is a uniform loop over shared memory.
// A face too large for a leaf list (more than kCandMax leaves) falls back to per-sample descents from the root with the
// previous facet as hint (measured: scanning a 96-node frontier per sample costs more than the 18-level descent it saves).
constexpr int kCandMax = 96;
struct Cands {
    uint32_t node[kCandMax];
    float box[kCandMax][6];
};
// returns the number of nodes left in *res: leaves (heap ids >= nLeafP) when *leaves, else the subtree roots of the deepest
// level that fitted
__device__ __forceinline__ int collect_leaves(const SurfaceView& S, const float lo[3], const float hi[3], Cands* A, Cands* B, int lane, const Cands** res,
                                              bool* leaves) {
    const unsigned full = 0xffffffffu;
    const unsigned lt = (1u << lane) - 1u;
    const float inf = __int_as_float(0x7f800000);
    const int L = 31 - __clz(S.nLeafP);
    const uint32_t first = 1u << (L % 3);
    __syncwarp();
    if ((uint32_t)lane < first) {
        A->node[lane] = first + lane;  // root-level boxes are not stored: always admitted
        A->box[lane][0] = A->box[lane][1] = A->box[lane][2] = -inf;
        A->box[lane][3] = A->box[lane][4] = A->box[lane][5] = inf;
    }
    __syncwarp();
    int count = (int)first;
    int d = L % 3;
    for (; d + 3 <= L && count > 0; d += 3) {
        const int total = 8 * count;
        int ncount = 0;
        bool overflow = false;
        for (int base = 0; base < total; base += 32) {
            const int idx = base + lane;
            bool ok = false;
            uint32_t child = 0;
            float b0 = 0, b1 = 0, b2 = 0, b3 = 0, b4 = 0, b5 = 0;
            if (idx < total) {
                child = 8u * A->node[idx >> 3] + (uint32_t)(idx & 7);
                const float* rec = reinterpret_cast<const float*>(S.pairs + (child >> 1)) + 6 * (child & 1u);
                const float2 u = __ldg(reinterpret_cast<const float2*>(rec));
                const float2 v = __ldg(reinterpret_cast<const float2*>(rec) + 1);
                const float2 w = __ldg(reinterpret_cast<const float2*>(rec) + 2);
                b0 = u.x; b1 = u.y; b2 = v.x; b3 = v.y; b4 = w.x; b5 = w.y;
                ok = b0 <= hi[0] && b3 >= lo[0] && b1 <= hi[1] && b4 >= lo[1] && b2 <= hi[2] && b5 >= lo[2];
            }
            const unsigned m = __ballot_sync(full, ok);
            if (ncount + __popc(m) > kCandMax) { overflow = true; break; }
            if (ok) {
                const int pos = ncount + __popc(m & lt);
                B->node[pos] = child;
                B->box[pos][0] = b0; B->box[pos][1] = b1; B->box[pos][2] = b2;
                B->box[pos][3] = b3; B->box[pos][4] = b4; B->box[pos][5] = b5;
            }
            ncount += __popc(m);
        }
        __syncwarp();
        if (overflow) break;  // A still holds the last level that fitted
        Cands* t = A; A = B; B = t;
        count = ncount;
    }
    *res = A;
    *leaves = (d == L) || count == 0;
    return count;
}

// one sample against the candidate list: hint facet first (LocalOperations.cpp:1080-1086), then the listed facets
__device__ __forceinline__ bool sample_out_cands(const SurfaceView& S, tw::V3 p, double eps2, float thr, uint32_t& prev, const Cands* C, int ncand) {
    double s, t; tw::V3 nd; bool deg;
    if (prev != TWG_NO_FACET && twd::facet_d2(S, prev, p, s, t, nd, deg) <= eps2) return false;
    const twd::PointF q = twd::bracket(p);
    for (int i = 0; i < ncand; ++i) {
        const float d = twd::box_d2_lb(q, C->box[i][0], C->box[i][1], C->box[i][2], C->box[i][3], C->box[i][4], C->box[i][5]);
        if (d <= thr) {
            const uint32_t pos = C->node[i] - S.nLeafP;
            if (pos != prev && pos < S.nF && twd::facet_d2(S, pos, p, s, t, nd, deg) <= eps2) { prev = pos; return false; }
        }
    }
    return true;
}

constexpr int kRunCap = 128;  // runs per pass: 4 fixed + up to 124 rows; taller triangles take more passes
struct __align__(16) FaceRun {
    double ox, oy, oz;  // first sample of a row (rows only)
    int cnt;            // samples in this run
    int off;            // samples before this run (exclusive scan)
};

// LDG: `tris` is global memory (read-only path); else it may be shared memory (the tiny-call kernel below): plain loads
template <bool LDG>
__device__ __forceinline__ void env_faces_body(const SurfaceView& S, const double* tris, uint64_t n, double sd, double eps2, uint32_t flags,
                                               uint8_t* __restrict__ out) {
    extern __shared__ __align__(128) unsigned char smraw[];
    NodePair* top = reinterpret_cast<NodePair*>(smraw);
    __shared__ __align__(8) uint64_t bar;
    __shared__ volatile int flag[kEnvThreads / 32];
    __shared__ FaceRun runs_all[kEnvThreads / 32][kRunCap];
    __shared__ Cands cands_all[kEnvThreads / 32][2];
    const uint32_t topN = stage_top(S, top, &bar);
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    FaceRun* runs = runs_all[wib];
    const float thr = __double2float_ru(eps2);
    const float epsf = __fsqrt_ru(thr);
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t f = warp; f < n; f += nwarps) {
        double t9[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) t9[k] = LDG ? __ldg(tris + f * 9 + k) : tris[f * 9 + k];
        // :1048 -- Preprocess::isOutEnvelop (Preprocess.cpp:643-747) has no such shortcut
        if (!(flags & TWG_FACES_NO_DEGENERATE_SHORTCUT) && tw::exact::triangle_is_degenerate(t9, t9 + 3, t9 + 6)) {
            if (lane == 0) out[f] = 0;
            continue;
        }
        // ---- candidate facets: leaves meeting the face's bounding box dilated by eps. The samples are convex combinations of
        // the vertices up to rounding (a few 1e-16 relative); the dilation is widened by 1e-4 eps + 4 float ulps of the
        // largest coordinate, orders of magnitude more than that.
        float flo[3], fhi[3];
        float amax = 0.f;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const double mn = fmin(t9[k], fmin(t9[3 + k], t9[6 + k])), mx = fmax(t9[k], fmax(t9[3 + k], t9[6 + k]));
            flo[k] = __double2float_rd(mn); fhi[k] = __double2float_ru(mx);
            amax = fmaxf(amax, fmaxf(fabsf(flo[k]), fabsf(fhi[k])));
        }
        const float pad = __fadd_ru(__fmul_ru(epsf, 1.0001f), __fmul_ru(amax, 4.8e-7f));
#pragma unroll
        for (int k = 0; k < 3; ++k) { flo[k] = __fsub_rd(flo[k], pad); fhi[k] = __fadd_ru(fhi[k], pad); }
        const Cands* cand = nullptr;
        bool leaves = false;
        const int ncand = collect_leaves(S, flo, fhi, &cands_all[wib][0], &cands_all[wib][1], lane, &cand, &leaves);
        if (ncand == 0) {  // nothing of the surface within eps of the face's box: every sample is out
            if (lane == 0) out[f] = 1;
            continue;
        }
        tw::SamplePlan P;
        tw::make_plan(t9, sd, P);
        // number of rows before the reference's `break` (:197-198): first row whose stop test fires
        int rows = 0;
        if (P.kind == 2) {
            rows = P.M;
            for (int base = 0; base < P.M; base += 32) {
                const int m = base + lane + 1;
                bool stop = false;
                if (m <= P.M) { tw::RowPlan R; tw::make_row(P, m, R); stop = R.stop; }
                const unsigned b = __ballot_sync(full, stop);
                if (b) { rows = base + __ffs(b) - 1; break; }
            }
        }
        if (lane == 0) flag[wib] = 0;
        __syncwarp();
        uint32_t prev = TWG_NO_FACET;
        bool found_out = false;
        // passes over the run list: pass 0 holds the four fixed runs + the first rows, later passes the remaining rows
        for (int row0 = 0; row0 == 0 || row0 < rows; ) {
            const int fixed = (row0 == 0) ? 4 : 0;
            const int nrow = min(rows - row0, kRunCap - fixed);
            const int nrun = fixed + nrow;
            // ---- run table: lengths
            if (row0 == 0 && lane < 4) {
                // run 0: base edge n = 0..nA-1; run 1: the single vertices (kind 0: v0 v1 v2, else v1 and v2);
                // run 2: edge v1->v2, n = 1..nB; run 3: edge v2->v0, n = 1..nC
                const int c = (lane == 0) ? P.nA : (lane == 1) ? ((P.kind == 0) ? 3 : 2) : (lane == 2) ? P.nB : P.nC;
                runs[lane].cnt = c;
            }
            for (int base = 0; base < nrow; base += 32) {
                const int j = base + lane;
                if (j < nrow) {
                    tw::RowPlan R;
                    tw::make_row(P, row0 + j + 1, R);
                    FaceRun& r = runs[fixed + j];
                    r.ox = R.v.x; r.oy = R.v.y; r.oz = R.v.z;
                    r.cnt = R.N1 + 1 > 0 ? R.N1 + 1 : 0;
                }
            }
            __syncwarp();
            // ---- exclusive scan of the lengths
            int total = 0;
            for (int base = 0; base < nrun; base += 32) {
                const int j = base + lane;
                const int c = (j < nrun) ? runs[j].cnt : 0;
                int incl = c;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int t = __shfl_up_sync(full, incl, o);
                    if (lane >= o) incl += t;
                }
                if (j < nrun) runs[j].off = total + incl - c;
                total += __shfl_sync(full, incl, 31);
            }
            __syncwarp();
            // ---- every lane walks its chunk of the concatenated runs
            const int chunk = (total + 31) / 32;
            int sidx = lane * chunk;
            const int send = min(total, sidx + chunk);
            int r = 0, i = 0;
            if (sidx < send) {
                int lo = 0, hi = nrun - 1;  // last run with off <= sidx: the run that covers sample sidx
                while (lo < hi) {
                    const int mid = (lo + hi + 1) >> 1;
                    if (runs[mid].off <= sidx) lo = mid; else hi = mid - 1;
                }
                r = lo;
                while (runs[r].cnt == 0 || runs[r].off + runs[r].cnt <= sidx) ++r;  // skip empty runs that share the offset
                i = sidx - runs[r].off;
            }
            auto next_sample = [&]() -> tw::V3 {  // sample sidx of the concatenated runs; advances the cursor
                while (i >= runs[r].cnt) { ++r; i = 0; }
                tw::V3 p;
                if (r >= fixed) p = tw::row_sample(tw::mk(runs[r].ox, runs[r].oy, runs[r].oz), P.n01, sd, i);
                else if (r == 0) p = tw::edge_sample(P.v0, P.n01, sd, i);
                else if (r == 1) p = (P.kind == 0) ? (i == 0 ? P.v0 : (i == 1 ? P.v1 : P.v2)) : (i == 0 ? P.v1 : P.v2);
                else if (r == 2) p = tw::edge_sample(P.v1, P.n12, sd, i + 1);
                else p = tw::edge_sample(P.v2, P.n20, sd, i + 1);
                ++i; ++sidx;
                return p;
            };
            if (!leaves) {
                // face too large for a candidate list: every lane descends from the root per sample (hint first)
                while (sidx < send && !found_out) {
                    if (flag[wib]) break;
                    const tw::V3 p = next_sample();
                    if (sample_out(S, p, eps2, prev, top, topN)) { found_out = true; flag[wib] = 1; }
                }
            } else {
                // ---- warp-synchronous rounds over three stages:
                //   FETCH  generate the lane's next sample; it goes to LEAF on the hint facet (the facet that held the lane's
                //          previous sample, LocalOperations.cpp:1080-1082) or, without a hint, to SCAN
                //   LEAF   exact point-triangle test of one facet; within eps -> the sample is IN (next FETCH), else SCAN
                //   SCAN   FP32 box bounds of the face's candidate leaves from where the lane stopped; first box within eps ->
                //          LEAF on that facet; list exhausted -> the sample, hence the face, is OUT
                // A face inside the envelope keeps all lanes in FETCH -> LEAF(hint) lockstep; per-lane loops left 8 of 32
                // lanes active per instruction (ncu) because hint hits, scans and leaf tests interleaved at random.
                enum { FETCH = 0, LEAF = 1, SCAN = 2, DONE = 3 };
                int stage = (sidx < send && !found_out) ? FETCH : DONE;
                tw::V3 p = tw::mk(0, 0, 0);
                twd::PointF q = twd::bracket(p);
                uint32_t leaf_pos = 0;
                int scan_i = 0;
                for (;;) {
                    // the flag is read BEFORE the vote: a lane that leaves the vote early may already be raising it in its SCAN stage
                    // below (racecheck, round 2), and lanes that disagreed about it would leave the loop at different rounds
                    const int stop = flag[wib];
                    if (__ballot_sync(full, stage != DONE) == 0u || __any_sync(full, stop != 0)) break;
                    // fixed order SCAN -> FETCH -> LEAF: scanning lanes find their next candidate and fetching lanes their next
                    // sample first (cheap stages, subsets of the warp), so that the expensive exact test then runs ONCE for
                    // every lane that has a facet to test -- hint facets and scan candidates in the same round
                    if (stage == SCAN) {
                        bool hit = false;
                        for (; scan_i < ncand; ++scan_i) {
                            const float d = twd::box_d2_lb(q, cand->box[scan_i][0], cand->box[scan_i][1], cand->box[scan_i][2], cand->box[scan_i][3],
                                                           cand->box[scan_i][4], cand->box[scan_i][5]);
                            const uint32_t pos = cand->node[scan_i] - S.nLeafP;
                            if (d <= thr && pos != prev && pos < S.nF) { leaf_pos = pos; hit = true; ++scan_i; break; }
                        }
                        if (hit) stage = LEAF;
                        else { found_out = true; flag[wib] = 1; stage = DONE; }
                    }
                    __syncwarp();
                    if (flag[wib]) break;
                    if (stage == FETCH) {
                        p = next_sample();
                        q = twd::bracket(p);
                        scan_i = 0;
                        if (prev != TWG_NO_FACET) { leaf_pos = prev; stage = LEAF; }
                        else stage = SCAN;
                    }
                    __syncwarp();
                    if (stage == LEAF) {
                        double s_, t_; tw::V3 nd; bool deg;
                        if (twd::facet_d2(S, leaf_pos, p, s_, t_, nd, deg) <= eps2) { prev = leaf_pos; stage = (sidx < send) ? FETCH : DONE; }
                        else stage = SCAN;
                    }
                    __syncwarp();
                }
            }
            __syncwarp();
            if (flag[wib]) break;
            row0 += nrow;
            if (nrow == 0) break;
            __syncwarp();
        }
        __syncwarp();
        const unsigned any = __ballot_sync(full, found_out);
        if (lane == 0) out[f] = any ? 1 : 0;
        __syncwarp();
    }
}

__global__ void __launch_bounds__(kEnvThreads, 4) env_faces_kernel(SurfaceView S, const double* __restrict__ tris, uint64_t n, double sd, double eps2,
                                                               uint32_t flags, uint8_t* __restrict__ out) {
    env_faces_body<true>(S, tris, n, sd, eps2, flags, out);
}

// The faces of ONE candidate operation (EdgeCollapser.cpp:770, VertexSmoother.cpp:425; n <= 16): the triangles travel in the
// kernel's parameters, the decisions go straight into the mapped slab, the last CTA raises the completion word itself.
struct TwgTinyFaces {
    double tri[16 * 9];
};
__global__ void __launch_bounds__(kEnvThreads, 4) env_faces_tiny_kernel(SurfaceView S, const __grid_constant__ TwgTinyFaces in, uint32_t n, double sd, double eps2,
                                                                    uint32_t flags, uint8_t* __restrict__ out, twg_done done) {
    __shared__ double st[16 * 9];
    for (int i = threadIdx.x; i < 16 * 9; i += blockDim.x) st[i] = in.tri[i];
    __syncthreads();
    env_faces_body<false>(S, st, n, sd, eps2, flags, out);
    twg_signal_done(done);
}

struct CollectSink {
    double* out;
    uint64_t cap, n;
    __device__ void operator()(tw::V3 p) {
        if (n < cap) { out[3 * n] = p.x; out[3 * n + 1] = p.y; out[3 * n + 2] = p.z; }
        ++n;
    }
};

__global__ void sample_triangle_kernel(const double* tri9, double sd, double* out, uint64_t cap, uint64_t* count) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double t9[9];
    for (int k = 0; k < 9; ++k) t9[k] = tri9[k];
    tw::SamplePlan P;
    tw::make_plan(t9, sd, P);
    CollectSink sink{out, cap, 0};
    tw::enumerate_samples(P, sink);
    *count = sink.n;
}

cudaStream_t pick(twg_ctx* c, void* stream) { return stream ? (cudaStream_t)stream : c->streams[0]; }

unsigned grid_persistent(twg_ctx* c, uint64_t items, int per_block, int ctas_per_sm) {
    uint64_t b = (items + per_block - 1) / per_block;
    const uint64_t m = (uint64_t)c->sm_count * ctas_per_sm;
    if (b > m) b = m;
    if (b == 0) b = 1;
    return (unsigned)b;
}

uint32_t top_n(const twg_surface* s) { const uint32_t k = (uint32_t)s->ctx->opt.env_top; return s->nLeafP < k ? s->nLeafP : k; }
size_t top_smem(const twg_surface* s) { return (size_t)top_n(s) * sizeof(NodePair); }
SurfaceView view_of(const twg_surface* s) { SurfaceView v = s->view(); v.topN = top_n(s); return v; }

}  // namespace

extern "C" {

int twg_envelope_points_out_dev(twg_surface* s, const double* dP, uint64_t n, double eps2, uint8_t* dOut, void* stream) {
    twg_ctx* c = s ? s->ctx : nullptr;
    TWG_CHECK(c, s && dP && dOut, TWG_ERR_INVALID_ARG, "null argument");
    TWG_CHECK(c, !twg_is_multi(c), TWG_ERR_INVALID_ARG, "_dev entry points take a one-device handle (twg_surface_replica)");
    TWG_CHECK(c, eps2 >= 0.0, TWG_ERR_INVALID_ARG, "eps2 must be >= 0");
    if (n == 0) return 0;
    TWG_CUDA(c, cudaSetDevice(c->device));
    cudaStream_t st = pick(c, stream);
    twg_lane* lane = nullptr;
    TWG_TRY(twg_get_lane(c, st, &lane));
    const uint32_t* perm = nullptr;
    const double* Pq = dP;  // queries in traversal order
    if (n >= TWG_SORT_MIN && c->opt.envelope_sort) TWG_TRY(twg_sort_points(c, lane, st, dP, n, &perm, s->sort_box, &Pq));
    TWG_CUDA(c, cudaMemsetAsync(lane->counters, 0, sizeof(unsigned long long), st));
    const int group = c->opt.env_group, policy = c->opt.env_policy, front = c->opt.env_front, quorum = c->opt.env_quorum;
    const unsigned grid = grid_persistent(c, (n + group - 1) / group, kEnvThreads / 32, 8);
    if (front <= 16)
        TWG_LAUNCH(c, (env_points_kernel<8, 16>), grid, kEnvThreads, top_smem(s), st, view_of(s), Pq, perm, n, eps2, dOut, lane->counters, group, policy, quorum, c->dcounters, c->opt.env_bound);
    else if (front >= 64)
        TWG_LAUNCH(c, (env_points_kernel<8, 64>), grid, kEnvThreads, top_smem(s), st, view_of(s), Pq, perm, n, eps2, dOut, lane->counters, group, policy, quorum, c->dcounters, c->opt.env_bound);
    else
        TWG_LAUNCH(c, (env_points_kernel<8, 32>), grid, kEnvThreads, top_smem(s), st, view_of(s), Pq, perm, n, eps2, dOut, lane->counters, group, policy, quorum, c->dcounters, c->opt.env_bound);
    return twg_lane_mark(c, lane);
}

int twg_nearest_dev(twg_surface* s, const double* dP, uint64_t n, uint32_t* dFacet, double* dNearest, double* dD2, void* stream) {
    twg_ctx* c = s ? s->ctx : nullptr;
    TWG_CHECK(c, s && dP, TWG_ERR_INVALID_ARG, "null argument");
    TWG_CHECK(c, !twg_is_multi(c), TWG_ERR_INVALID_ARG, "_dev entry points take a one-device handle (twg_surface_replica)");
    if (n == 0) return 0;
    TWG_CUDA(c, cudaSetDevice(c->device));
    cudaStream_t st = pick(c, stream);
    twg_lane* lane = nullptr;
    TWG_TRY(twg_get_lane(c, st, &lane));
    const uint32_t* perm = nullptr;
    const double* Pq = nullptr;
    const bool sorted = n >= TWG_SORT_MIN && c->opt.envelope_sort;
    if (sorted) TWG_TRY(twg_sort_points(c, lane, st, dP, n, &perm, s->sort_box, &Pq, nullptr, c->opt.nearest_curve));
    TWG_CUDA(c, cudaMemsetAsync(lane->counters, 0, sizeof(unsigned long long), st));
    if (sorted && c->opt.nearest_mode == 2 && s->nLeafP >= 8) {  // round-scheduled lanes (form 3)
        const int group = c->opt.nearest_group;
        const unsigned grid = grid_persistent(c, (n + group - 1) / group, kEnvThreads / 32, 6);
        TWG_LAUNCH(c, (nearest_rounds_kernel<6>), grid, kEnvThreads, top_smem(s), st, view_of(s), Pq, perm, n, dFacet, dNearest, dD2, lane->counters, group,
                   c->opt.env_quorum, c->opt.trace ? c->dcounters : (unsigned long long*)nullptr);
        return twg_lane_mark(c, lane);
    }
    if (sorted && c->opt.nearest_mode == 1 && s->nLeafP >= 8) {  // packets of 32 neighbouring queries share one traversal (default)
        const uint64_t claims = ((n + 31) / 32 + kPacketChunk - 1) / kPacketChunk;
        TWG_LAUNCH(c, nearest_packet_kernel, grid_persistent(c, claims, kEnvThreads / 32, 6), kEnvThreads, top_smem(s), st, view_of(s), Pq, perm, n, dFacet,
                   dNearest, dD2, lane->counters, c->opt.nearest_budget);
        return twg_lane_mark(c, lane);
    }
    TWG_LAUNCH(c, nearest_kernel, grid_persistent(c, (n + 32 * kNearRun - 1) / (32 * kNearRun), kEnvThreads / 32, 8), kEnvThreads, top_smem(s), st, view_of(s), dP,
               perm, n, dFacet, dNearest, dD2, lane->counters);
    return twg_lane_mark(c, lane);
}

int twg_envelope_faces_out_dev(twg_surface* s, const double* dTris, uint64_t n, double sd, double eps2, uint8_t* dOut, void* stream) {
    return twg_envelope_faces_out_ex_dev(s, dTris, n, sd, eps2, 0u, dOut, stream);
}

int twg_envelope_faces_out_ex_dev(twg_surface* s, const double* dTris, uint64_t n, double sd, double eps2, uint32_t flags, uint8_t* dOut, void* stream) {
    twg_ctx* c = s ? s->ctx : nullptr;
    TWG_CHECK(c, s && dTris && dOut, TWG_ERR_INVALID_ARG, "null argument");
    TWG_CHECK(c, eps2 >= 0.0 && sd > 0.0 && isfinite(sd), TWG_ERR_INVALID_ARG, "need eps2 >= 0 and finite sampling_dist > 0");
    TWG_CHECK(c, !twg_is_multi(c), TWG_ERR_INVALID_ARG, "_dev entry points take a one-device handle (twg_surface_replica)");
    if (n == 0) return 0;
    TWG_CUDA(c, cudaSetDevice(c->device));
    TWG_LAUNCH(c, env_faces_kernel, grid_persistent(c, n, kEnvThreads / 32, 8), kEnvThreads, top_smem(s), pick(c, stream), view_of(s), dTris, n, sd, eps2, flags, dOut);
    return 0;
}

// ---- host-buffer entry points ----
static int points_host(twg_surface* s, int what, const double* P, uint64_t n, double eps2, uint8_t* out, uint32_t* facet, double* nearest, double* d2) {
    twg_ctx* c = s->ctx;
    if (n == 0) return 0;
    if (!s->replicas.empty()) {  // multi-device handle: contiguous index ranges, one per device (multi.cu)
        if (n < TWG_MULTI_MIN_POINTS) return twg_forward0(c, points_host(s->replicas[0], what, P, n, eps2, out, facet, nearest, d2));
        const uint64_t G = s->replicas.size();
        return twg_multi_run(c, [&](int k, twg_ctx*) {
            const uint64_t b = n * (uint64_t)k / G, e = n * (uint64_t)(k + 1) / G;
            return points_host(s->replicas[k], what, P + 3 * b, e - b, eps2, out ? out + b : nullptr, facet ? facet + b : nullptr,
                               nearest ? nearest + 3 * b : nullptr, d2 ? d2 + b : nullptr);
        });
    }
    TWG_CUDA(c, cudaSetDevice(c->device));
    // 1 Mi points (24 MiB in) per slot by default: small enough that the first copy and the last kernel, which nothing
    // overlaps, are a small part of a 10 M batch; large enough for the sort + traversal to run at full rate
    const uint64_t chunk_points = (uint64_t)c->opt.chunk_points;
    auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
    // the exact nearest search is compute-bound (6 ns per point against 1.2 ns of PCIe) and loses efficiency on small
    // launches (far queries cluster; measured 17 ns per point at 2 M, 6.4 ns at 10 M): it gets 8 Mi-point chunks
    const uint64_t chunk = (what == 0) ? chunk_points : (chunk_points > (8ull << 20) ? chunk_points : (8ull << 20));
    const uint64_t cmax = n < chunk ? n : chunk;
    const size_t pb = up(cmax * 24), ob = up(cmax), fb = up(cmax * 4), nb = up(cmax * 24), db = up(cmax * 8);
    for (int k = 0; k < TWG_NUM_STREAMS; ++k) TWG_TRY(twg_ensure_scratch(c, k, pb + ob + fb + nb + db));
    if (n <= 64 && c->opt.fast_calls) {
        // ONE point of the sequential scheduler (EdgeCollapser.cpp:312,322; VertexSmoother.cpp:358): the query and the answer live in
        // the mapped slab, the kernel reads and writes them over PCIe, the host spins on the completion word
        char* slab;
        const size_t qb = up(n * 24), rb = up(n), fb2 = up(n * 4), nb2 = up(n * 24), db2 = up(n * 8);
        TWG_TRY(twg_fast_slab(c, qb + rb + fb2 + nb2 + db2, &slab));
        cudaStream_t st = c->streams[0];
        TwgTinyPoints in;
        memcpy(in.xyz, P, n * 24);
        twg_done done;
        TWG_TRY(twg_fast_arm(c, &done));
        TWG_LAUNCH(c, points_tiny_kernel, 1, 64, 0, st, view_of(s), in, (uint32_t)n, what, eps2, (uint8_t*)(slab + qb), facet ? (uint32_t*)(slab + qb + rb) : nullptr,
                   nearest ? (double*)(slab + qb + rb + fb2) : nullptr, d2 ? (double*)(slab + qb + rb + fb2 + nb2) : nullptr, done);
        TWG_TRY(twg_fast_spin(c, st, done.seq));
        if (what == 0) memcpy(out, slab + qb, n);
        if (facet) memcpy(facet, slab + qb + rb, n * 4);
        if (nearest) memcpy(nearest, slab + qb + rb + fb2, n * 24);
        if (d2) memcpy(d2, slab + qb + rb + fb2 + nb2, n * 8);
        return 0;
    }
    if (what == 0 && n <= 16384) {  // a few points (EdgeCollapser.cpp:322 asks for one): pinned slabs, one copy each way
        TWG_TRY(twg_ensure_pinned(c, up(n * 24), up(n)));
        cudaStream_t st = c->streams[0];
        char* base = (char*)c->dscratch[0];
        memcpy(c->pin_in[0], P, n * 24);
        TWG_CUDA(c, cudaMemcpyAsync(base, c->pin_in[0], n * 24, cudaMemcpyHostToDevice, st));
        TWG_TRY(twg_envelope_points_out_dev(s, (const double*)base, n, eps2, (uint8_t*)(base + pb), st));
        TWG_CUDA(c, cudaMemcpyAsync(c->pin_out[0], base + pb, n, cudaMemcpyDeviceToHost, st));
        TWG_CUDA(c, cudaStreamSynchronize(st));
        memcpy(out, c->pin_out[0], n);
        return 0;
    }
    int slot = 0;
    for (uint64_t b = 0; b < n; b += chunk, slot = (slot + 1) % TWG_NUM_STREAMS) {
        const uint64_t m = (n - b < chunk) ? (n - b) : chunk;
        cudaStream_t st = c->streams[slot];
        char* base = (char*)c->dscratch[slot];
        double* dP = (double*)base;
        uint8_t* dO = (uint8_t*)(base + pb);
        uint32_t* dF = (uint32_t*)(base + pb + ob);
        double* dN = (double*)(base + pb + ob + fb);
        double* dD = (double*)(base + pb + ob + fb + nb);
        TWG_CUDA(c, cudaMemcpyAsync(dP, P + 3 * b, m * 24, cudaMemcpyHostToDevice, st));
        if (what == 0) {
            TWG_TRY(twg_envelope_points_out_dev(s, dP, m, eps2, dO, st));
            TWG_CUDA(c, cudaMemcpyAsync(out + b, dO, m, cudaMemcpyDeviceToHost, st));
        } else {
            TWG_TRY(twg_nearest_dev(s, dP, m, facet ? dF : nullptr, nearest ? dN : nullptr, d2 ? dD : nullptr, st));
            if (facet) TWG_CUDA(c, cudaMemcpyAsync(facet + b, dF, m * 4, cudaMemcpyDeviceToHost, st));
            if (nearest) TWG_CUDA(c, cudaMemcpyAsync(nearest + 3 * b, dN, m * 24, cudaMemcpyDeviceToHost, st));
            if (d2) TWG_CUDA(c, cudaMemcpyAsync(d2 + b, dD, m * 8, cudaMemcpyDeviceToHost, st));
        }
    }
    for (int k = 0; k < TWG_NUM_STREAMS; ++k) TWG_CUDA(c, cudaStreamSynchronize(c->streams[k]));
    return 0;
}

int twg_envelope_points_out(twg_surface* s, const double* P, uint64_t n, double eps2, uint8_t* out) {
    twg_ctx* c = s ? s->ctx : nullptr;
    TWG_CHECK(c, s && P && out, TWG_ERR_INVALID_ARG, "null argument");
    TWG_CHECK(c, eps2 >= 0.0, TWG_ERR_INVALID_ARG, "eps2 must be >= 0");
    return points_host(s, 0, P, n, eps2, out, nullptr, nullptr, nullptr);
}

int twg_nearest(twg_surface* s, const double* P, uint64_t n, uint32_t* facet, double* nearest_xyz, double* d2) {
    twg_ctx* c = s ? s->ctx : nullptr;
    TWG_CHECK(c, s && P, TWG_ERR_INVALID_ARG, "null argument");
    return points_host(s, 1, P, n, 0.0, nullptr, facet, nearest_xyz, d2);
}

int twg_envelope_faces_out(twg_surface* s, const double* tris, uint64_t n, double sd, double eps2, uint8_t* out) {
    return twg_envelope_faces_out_ex(s, tris, n, sd, eps2, 0u, out);
}

int twg_envelope_faces_out_ex(twg_surface* s, const double* tris, uint64_t n, double sd, double eps2, uint32_t flags, uint8_t* out) {
    twg_ctx* c = s ? s->ctx : nullptr;
    TWG_CHECK(c, s && tris && out, TWG_ERR_INVALID_ARG, "null argument");
    if (n == 0) return 0;
    if (!s->replicas.empty()) {
        if (n < TWG_MULTI_MIN_FACES) return twg_forward0(c, twg_envelope_faces_out_ex(s->replicas[0], tris, n, sd, eps2, flags, out));
        const uint64_t G = s->replicas.size();
        return twg_multi_run(c, [&](int k, twg_ctx*) {
            const uint64_t b = n * (uint64_t)k / G, e = n * (uint64_t)(k + 1) / G;
            return twg_envelope_faces_out_ex(s->replicas[k], tris + 9 * b, e - b, sd, eps2, flags, out + b);
        });
    }
    TWG_CUDA(c, cudaSetDevice(c->device));
    auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
    if (n <= 16 && c->opt.fast_calls) {  // the faces of ONE candidate (EdgeCollapser.cpp:770, VertexSmoother.cpp:425): zero-copy slab
        char* slab;
        TWG_TRY(twg_fast_slab(c, up(n * 72) + up(n), &slab));
        cudaStream_t st = c->streams[0];
        TWG_CHECK(c, eps2 >= 0.0 && sd > 0.0 && isfinite(sd), TWG_ERR_INVALID_ARG, "need eps2 >= 0 and finite sampling_dist > 0");
        TwgTinyFaces in;
        memcpy(in.tri, tris, n * 72);
        twg_done done;
        TWG_TRY(twg_fast_arm(c, &done));
        // one warp per face, kEnvThreads / 32 warps per CTA
        TWG_LAUNCH(c, env_faces_tiny_kernel, (unsigned)((n + kEnvThreads / 32 - 1) / (kEnvThreads / 32)), kEnvThreads, top_smem(s), st, view_of(s), in, (uint32_t)n, sd, eps2,
                   flags, (uint8_t*)(slab + up(n * 72)), done);
        TWG_TRY(twg_fast_spin(c, st, done.seq));
        memcpy(out, slab + up(n * 72), n);
        return 0;
    }
    if (n <= 4096) {  // the faces of one local operation: pinned slabs, one copy each way (pageable copies cost ~10 us each)
        TWG_TRY(twg_ensure_scratch(c, 0, up(n * 72) + up(n)));
        TWG_TRY(twg_ensure_pinned(c, up(n * 72), up(n)));
        char* base = (char*)c->dscratch[0];
        cudaStream_t st = c->streams[0];
        memcpy(c->pin_in[0], tris, n * 72);
        TWG_CUDA(c, cudaMemcpyAsync(base, c->pin_in[0], n * 72, cudaMemcpyHostToDevice, st));
        TWG_TRY(twg_envelope_faces_out_ex_dev(s, (const double*)base, n, sd, eps2, flags, (uint8_t*)(base + up(n * 72)), st));
        TWG_CUDA(c, cudaMemcpyAsync(c->pin_out[0], base + up(n * 72), n, cudaMemcpyDeviceToHost, st));
        TWG_CUDA(c, cudaStreamSynchronize(st));
        memcpy(out, c->pin_out[0], n);
        return 0;
    }
    // large batches: 1 Mi-face chunks over the context's streams (bounds the scratch memory; smaller chunks cost more in kernel tails than the overlapped copy saves: 60.6 vs 67.5 M faces/s at 64 Ki)
    const uint64_t chunk = 1ull << 20;
    const uint64_t cmaxf = n < chunk ? n : chunk;
    for (int k = 0; k < TWG_NUM_STREAMS; ++k) TWG_TRY(twg_ensure_scratch(c, k, up(cmaxf * 72) + up(cmaxf)));
    int slot = 0;
    for (uint64_t b = 0; b < n; b += chunk, slot = (slot + 1) % TWG_NUM_STREAMS) {
        const uint64_t m = (n - b < chunk) ? (n - b) : chunk;
        cudaStream_t st = c->streams[slot];
        char* base = (char*)c->dscratch[slot];
        TWG_CUDA(c, cudaMemcpyAsync(base, tris + 9 * b, m * 72, cudaMemcpyHostToDevice, st));
        TWG_TRY(twg_envelope_faces_out_ex_dev(s, (const double*)base, m, sd, eps2, flags, (uint8_t*)(base + up(cmaxf * 72)), st));
        TWG_CUDA(c, cudaMemcpyAsync(out + b, base + up(cmaxf * 72), m, cudaMemcpyDeviceToHost, st));
    }
    for (int k = 0; k < TWG_NUM_STREAMS; ++k) TWG_CUDA(c, cudaStreamSynchronize(c->streams[k]));
    return 0;
}

int twg_sample_triangle(twg_ctx* c, const double* tri9, double sd, double* out_xyz, uint64_t cap, uint64_t* count) {
    TWG_CHECK(c, c && tri9 && count, TWG_ERR_INVALID_ARG, "null argument");
    if (twg_is_multi(c)) return twg_forward0(c, twg_sample_triangle(c->children[0], tri9, sd, out_xyz, cap, count));
    TWG_CHECK(c, sd > 0.0 && isfinite(sd), TWG_ERR_INVALID_ARG, "need finite sampling_dist > 0");
    TWG_CUDA(c, cudaSetDevice(c->device));
    auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
    if (!out_xyz) cap = 0;
    TWG_TRY(twg_ensure_scratch(c, 0, 256 + 256 + up(cap * 24)));
    char* base = (char*)c->dscratch[0];
    cudaStream_t st = c->streams[0];
    TWG_CUDA(c, cudaMemcpyAsync(base, tri9, 72, cudaMemcpyHostToDevice, st));
    TWG_LAUNCH(c, sample_triangle_kernel, 1, 32, 0, st, (const double*)base, sd, (double*)(base + 512), cap, (uint64_t*)(base + 256));
    TWG_CUDA(c, cudaMemcpyAsync(count, base + 256, 8, cudaMemcpyDeviceToHost, st));
    TWG_CUDA(c, cudaStreamSynchronize(st));
    const uint64_t m = *count < cap ? *count : cap;
    if (m) {
        TWG_CUDA(c, cudaMemcpyAsync(out_xyz, base + 512, m * 24, cudaMemcpyDeviceToHost, st));
        TWG_CUDA(c, cudaStreamSynchronize(st));
    }
    return 0;
}

}  // extern "C"
