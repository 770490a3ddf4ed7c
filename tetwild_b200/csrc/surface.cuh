// surface.cuh -- device-resident envelope structure (replaces GEO::MeshFacetsAABBWithEps, mesh_AABB.{h,cpp}).
//
// Layout in HBM (all built on the device, see surface.cu):
//   tris   [nF]      one 128-byte TriRec per facet, in Morton order of the facet centroids (tw_math.cuh)
//   triV   [nF*9]    the caller's exact vertex coordinates per sorted facet (degenerate facets, hint output)
//   pairs  [nLeafP]  implicit binary heap over the sorted facets, padded to a power of two: record i (1 <= i <
//                    nLeafP) holds the bounding boxes of BOTH children 2i and 2i+1 as 12 floats = three 128-bit
//                    loads per visit. Boxes are rounded OUTWARD to float (conservative: a node is never pruned
//                    that the exact double box would have kept); nodes >= nLeafP are the facets themselves
//                    (leaf j <-> sorted facet j), padding leaves have empty boxes (+inf/-inf).
// The reference's tree is the same implicit balanced tree over Morton-sorted facets (mesh_AABB.cpp:88-141) with
// double boxes, one per node; results do not depend on the tree shape.
#pragma once
#include "common.cuh"

struct __align__(16) NodePair {
    float4 a, b, c;  // left: lo=(a.x,a.y,a.z) hi=(a.w,b.x,b.y); right: lo=(b.z,b.w,c.x) hi=(c.y,c.z,c.w)
};
static_assert(sizeof(NodePair) == 48, "NodePair is three 128-bit words");

using tw::TriBound;  // oriented bound of one facet: tw_math.cuh

struct SurfaceView {
    const NodePair* pairs;
    const tw::TriRec* tris;
    const double* triV;
    const TriBound* tb;
    uint32_t nF;
    uint32_t nLeafP;  // power of two >= max(nF, 2)
    uint32_t topN;    // pair records [0, topN) are staged in shared memory by the query kernels
};

struct twg_surface {
    twg_ctx* ctx = nullptr;
    uint32_t nF = 0, nLeafP = 0;
    NodePair* pairs = nullptr;
    tw::TriRec* tris = nullptr;
    double* triV = nullptr;
    TriBound* tb = nullptr;
    std::vector<twg_surface*> replicas;  // handle made on a multi-device context: one replica per device (multi.cu); else empty
    double bbox[6] = {0, 0, 0, 0, 0, 0};     // lo xyz, hi xyz of the surface
    double sort_box[6] = {0, 0, 0, 0, 0, 0}; // bbox grown by 5 %: Morton quantisation box of query batches (qsort.cu)
    SurfaceView view() const { return SurfaceView{pairs, tris, triV, tb, nF, nLeafP, 0}; }
};

#if defined(__CUDACC__)
namespace twd {

// squared distance from p to an (outward-rounded) float box, in double; 0 inside. A lower bound of the distance
// to anything inside the box; callers compare with a relative slack (kSlack) so that rounding in this bound can
// never prune a facet whose own computed d2 passes the exact comparison.
__device__ __forceinline__ double box_d2(double px, double py, double pz, float lx, float ly, float lz, float hx, float hy, float hz) {
    double dx = fmax(fmax((double)lx - px, px - (double)hx), 0.0);
    double dy = fmax(fmax((double)ly - py, py - (double)hy), 0.0);
    double dz = fmax(fmax((double)lz - pz, pz - (double)hz), 0.0);
    return dx * dx + dy * dy + dz * dz;
}
constexpr double kSlack = 1.0 + 1e-9;

__device__ __forceinline__ NodePair load_pair(const NodePair* p) {
    NodePair r;
    const float4* q = reinterpret_cast<const float4*>(p);
    r.a = __ldg(q); r.b = __ldg(q + 1); r.c = __ldg(q + 2);
    return r;
}

__device__ __forceinline__ tw::TriRec load_tri(const tw::TriRec* p) {
    tw::TriRec r;
    const double2* q = reinterpret_cast<const double2*>(p);
    double2* d = reinterpret_cast<double2*>(&r);
#pragma unroll
    for (int k = 0; k < 8; ++k) d[k] = __ldg(q + k);
    return r;
}

// d2 from p to sorted facet `pos` (any facet, degenerate included); s,t valid only when !degenerate
__device__ __forceinline__ double facet_d2(const SurfaceView& S, uint32_t pos, tw::V3 p, double& s, double& t, tw::V3& near_deg, bool& deg) {
    tw::TriRec r = load_tri(S.tris + pos);
    deg = (r.flags & 1u) != 0;
    if (!deg) return tw::tri_sqdist_rec(p, r, s, t);
    double tv[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) tv[k] = __ldg(S.triV + (size_t)pos * 9 + k);
    s = t = 0.0;
    return tw::tri_sqdist_degenerate(p, tv, near_deg);
}

__device__ __forceinline__ TriBound load_bound(const TriBound* p) {
    TriBound r;
    const float4* q = reinterpret_cast<const float4*>(p);
    const float4 a = __ldg(q), b = __ldg(q + 1);
    r.cx = a.x; r.cy = a.y; r.cz = a.z; r.R = a.w;
    r.nx = b.x; r.ny = b.y; r.nz = b.z; r.w = b.w;
    return r;
}
using tw::bound_lb2;

// Conservative single-precision box test: tw_math.cuh (host/device, so that the CPU tier can check its rigor)
using tw::PointF;
using tw::bracket;
using tw::box_d2_lb;

// Is some facet within sqrt(eps2) of p?  (facet_in_envelope_recursive, mesh_AABB.cpp:482-548: stop at the first
// facet with d2 <= eps2, never enter a box farther than eps.)  Binary descent, one query per lane; used by the
// face kernel, whose lanes each walk their own run of samples. top/topN: pair records [0, topN) staged in shared memory.
__device__ __forceinline__ bool in_envelope(const SurfaceView& S, tw::V3 p, double eps2, uint32_t& hit_pos, const NodePair* top, uint32_t topN,
                                            uint32_t start = 1u) {  // start: root of the subtree to search (an internal node)
    const float thr = __double2float_ru(eps2);
    const PointF q = bracket(p);
    uint32_t stack[32];
    int sp = 0;
    uint32_t node = start;
    const uint32_t leaf0 = S.nLeafP;
    for (;;) {
        NodePair np = (node < topN) ? top[node] : load_pair(S.pairs + node);
        const float dl = box_d2_lb(q, np.a.x, np.a.y, np.a.z, np.a.w, np.b.x, np.b.y);
        const float dr = box_d2_lb(q, np.b.z, np.b.w, np.c.x, np.c.y, np.c.z, np.c.w);
        const bool hl = dl <= thr, hr = dr <= thr;
        const uint32_t cl = 2u * node;
        if (cl >= leaf0) {
            const bool lfirst = !(hr && dr < dl);
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const bool left = (k == 0) ? lfirst : !lfirst;
                const uint32_t pos = (left ? cl : cl + 1u) - leaf0;
                if ((left ? hl : hr) && pos < S.nF) {
                    double s, t; tw::V3 nd; bool deg;
                    const double d2 = facet_d2(S, pos, p, s, t, nd, deg);
                    if (d2 <= eps2) { hit_pos = pos; return true; }
                }
            }
        } else if (hl || hr) {
            if (hl && hr) {
                const bool lnear = dl <= dr;
                stack[sp++] = lnear ? cl + 1u : cl;
                node = lnear ? cl : cl + 1u;
            } else {
                node = hl ? cl : cl + 1u;
            }
            continue;
        }
        if (sp == 0) return false;
        node = stack[--sp];
    }
}

// One step of the 8-wide traversal used by the point kernel: the boxes of the eight descendants 8i .. 8i+7 of heap
// node i (three levels down) are the four CONSECUTIVE pair records 4i .. 4i+3 = one 192-byte run, fetched with twelve
// independent 128-bit loads. A query near the surface makes ceil(depth / 3) dependent memory round trips instead of
// depth (6 instead of 18 for 200k facets, the first three from the shared-memory copy of the top of the tree).
// Returns the bit mask of admitted descendants and their lower-bound distances.
__device__ __forceinline__ uint32_t wide_step(const SurfaceView& S, const PointF& q, float thr, uint32_t node, const NodePair* top, uint32_t topN, float d[8]) {
    const uint32_t pr = 4u * node;
#pragma unroll
    for (int h = 0; h < 4; ++h) {
        const NodePair np = (pr + 3u < topN) ? top[pr + h] : load_pair(S.pairs + pr + h);
        d[2 * h] = box_d2_lb(q, np.a.x, np.a.y, np.a.z, np.a.w, np.b.x, np.b.y);
        d[2 * h + 1] = box_d2_lb(q, np.b.z, np.b.w, np.c.x, np.c.y, np.c.z, np.c.w);
    }
    uint32_t mask = 0;
#pragma unroll
    for (int c = 0; c < 8; ++c) mask |= (d[c] <= thr) ? (1u << c) : 0u;
    return mask;
}

struct Nearest {
    double d2, s, t;
    uint32_t pos;
    bool deg;
    tw::V3 pt_deg;
};

// Exact nearest facet (nearest_facet_recursive, mesh_AABB.cpp:418-480). Boxes are pruned with the conservative FP32 lower
// bound of the box distance (never larger than the true distance, so no subtree that could hold a nearer facet is skipped;
// the reference prunes with the double box distance, visits a subset of these boxes and finds the same minimum).
__device__ __forceinline__ void nearest_facet(const SurfaceView& S, tw::V3 p, Nearest& best, const NodePair* top, uint32_t topN) {
    uint32_t stack[32];
    float dstack[32];
    const PointF q = bracket(p);
    int sp = 0;
    uint32_t node = 1;
    const uint32_t leaf0 = S.nLeafP;
    for (;;) {
        NodePair np = (node < topN) ? top[node] : load_pair(S.pairs + node);
        const float dl = box_d2_lb(q, np.a.x, np.a.y, np.a.z, np.a.w, np.b.x, np.b.y);
        const float dr = box_d2_lb(q, np.b.z, np.b.w, np.c.x, np.c.y, np.c.z, np.c.w);
        const uint32_t cl = 2u * node;
        if (cl >= leaf0) {
            const bool lfirst = dl <= dr;
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const bool left = (k == 0) ? lfirst : !lfirst;
                const float db = left ? dl : dr;
                const uint32_t pos = (left ? cl : cl + 1u) - leaf0;
                if ((double)db <= best.d2 * kSlack && pos < S.nF) {  // padding leaves (pos >= nF) have empty boxes
                    double s, t; tw::V3 nd; bool deg;
                    const double d2 = facet_d2(S, pos, p, s, t, nd, deg);
                    if (d2 < best.d2) { best.d2 = d2; best.s = s; best.t = t; best.pos = pos; best.deg = deg; best.pt_deg = nd; }
                }
            }
        } else {
            const bool hl = ((double)dl <= best.d2 * kSlack) && (dl < 1e30f), hr = ((double)dr <= best.d2 * kSlack) && (dr < 1e30f);
            if (hl && hr) {
                const bool lnear = dl <= dr;
                stack[sp] = lnear ? cl + 1u : cl;
                dstack[sp++] = lnear ? dr : dl;
                node = lnear ? cl : cl + 1u;
                continue;
            } else if (hl || hr) {
                node = hl ? cl : cl + 1u;
                continue;
            }
        }
        for (;;) {
            if (sp == 0) return;
            --sp;
            if ((double)dstack[sp] <= best.d2 * kSlack) { node = stack[sp]; break; }
        }
    }
}

}  // namespace twd
#endif
