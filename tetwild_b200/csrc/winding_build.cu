// winding_build.cu -- the winding-number hierarchy built ON THE DEVICE, bit-identical to build_host_tree (winding.cu).
//
// What igl::winding_number builds per call (WindingNumberAABB over the surface handed over at InoutFiltering.cpp:40-45) took
// 0.27 s on 16 host cores for 1.0 M facets -- the dominant cost of a one-shot call at the sizes TetWild meets (10^5 facets,
// 10^6 centroids). Every phase of the host build is restated as data-parallel passes with the SAME decisions:
//   1  vertex merge        three stable radix sorts of the vertex ids by z, y, x (order-preserving keys); canon = first id of a run
//   2  kd order            level by level: per segment the centroid box -> longest axis; ONE stable sort of all facets on
//                          (segment, centroid coordinate) in facet-id order = the host's nth_element with its (coordinate, id)
//                          order; left child = the first (blocks / 2) * leaf facets. Blocks are then put in facet-id order
//                          (the canonical order the host build uses too).
//   3  facets, leaf boxes, heap boxes, subtree ranges
//   4  exterior edges      half-edges sorted by undirected key; one thread per group replays the host's per-level net count
//                          (count pass, exclusive scan, write pass: deterministic); records sorted by (node, a, b)
//   5  cap polylines       one thread per node replays the host's greedy walk over its sorted edges; the apex is the first
//                          edge's start vertex, so runs are cut and emitted on the fly into an over-allocated region
//                          (<= 2 points per edge), then compacted with a scan.
// Sorts on this (build-time) path use cub::DeviceRadixSort; the per-query path has its own sort (qsort.cu).
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include "winding.cuh"

namespace {

struct Arena {  // device scratch of one build: stream-ordered allocations from the device's pool (no synchronising cudaMalloc / cudaFree)
    cudaStream_t st;
    std::vector<void*> ptrs;
    void* tmp = nullptr;     // temporary storage of the library sorts / scans, grown on demand (stream-ordered reuse)
    size_t tmp_bytes = 0;
    explicit Arena(cudaStream_t s) : st(s) {}
    ~Arena() { for (void* p : ptrs) cudaFreeAsync(p, st); }
    cudaError_t temp(size_t bytes, void** out) {
        if (bytes > tmp_bytes) {
            const size_t want = bytes + bytes / 4 + 256;
            void* p = nullptr;
            const cudaError_t e = cudaMallocAsync(&p, want, st);
            if (e != cudaSuccess) return e;
            ptrs.push_back(p);   // the old one stays alive until the arena dies: earlier work on the stream may still use it
            tmp = p;
            tmp_bytes = want;
        }
        *out = tmp;
        return cudaSuccess;
    }
    template <class T>
    cudaError_t get(T** p, size_t n) {
        *p = nullptr;
        const cudaError_t e = cudaMallocAsync((void**)p, (n ? n : 1) * sizeof(T), st);
        if (e == cudaSuccess) ptrs.push_back(*p);
        return e;
    }
};

__device__ __forceinline__ unsigned long long enc64(double d) {  // order-preserving; -0.0 and +0.0 give one key (they compare equal)
    d = d + 0.0;
    const unsigned long long b = (unsigned long long)__double_as_longlong(d);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ uint32_t enc32(float f) {
    f = f + 0.0f;
    const uint32_t b = __float_as_uint(f);
    return (b >> 31) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float dec32(uint32_t u) { return __uint_as_float((u >> 31) ? (u & 0x7fffffffu) : ~u); }

__global__ void iota_kernel(uint32_t* a, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = i;
}
__global__ void vkey_kernel(const double* __restrict__ V, const uint32_t* __restrict__ idx, uint32_t nV, int comp, unsigned long long* keys) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nV) keys[i] = enc64(V[3 * (size_t)idx[i] + comp]);
}
__global__ void vrun_kernel(const double* __restrict__ V, const uint32_t* __restrict__ idx, uint32_t nV, uint32_t* start) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nV) return;
    bool head = i == 0;
    if (!head) {
        const double *p = V + 3 * (size_t)idx[i], *q = V + 3 * (size_t)idx[i - 1];
        head = !(p[0] == q[0] && p[1] == q[1] && p[2] == q[2]);
    }
    start[i] = head ? i : 0u;
}
__global__ void canon_kernel(const uint32_t* __restrict__ idx, const uint32_t* __restrict__ start, uint32_t nV, uint32_t* canon) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nV) canon[idx[i]] = idx[start[i]];
}

// ---- kd order
__global__ void ctr_kernel(const double* __restrict__ V, const uint32_t* __restrict__ F, uint32_t nF, float* ctr) {
    const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nF) return;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const double a = V[3 * (size_t)F[3 * (size_t)f] + c], b = V[3 * (size_t)F[3 * (size_t)f + 1] + c], d = V[3 * (size_t)F[3 * (size_t)f + 2] + c];
        ctr[3 * (size_t)f + c] = __double2float_rn(__ddiv_rn(__dadd_rn(__dadd_rn(a, b), d), 3.0));
    }
}
__global__ void bb_init_kernel(uint32_t* bb, size_t n) {  // min slots start at the largest key, max slots at the smallest
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) bb[i] = (i % 6) < 3 ? 0xffffffffu : 0u;
}
struct MaxU32 {
    __device__ __forceinline__ uint32_t operator()(uint32_t a, uint32_t b) const { return a > b ? a : b; }
};
// bb[s][0..2] = min (encoded), bb[s][3..5] = max over the facets of segment s. min / max do not depend on the order they are
// taken in, so the reduction is free to be hierarchical: lanes of a warp that share a segment (__match_any_sync) reduce with
// redux.sync and their leader alone goes on; while the level has at most kSegSmem segments the leaders meet in shared memory and
// each CTA sends one atomic per segment and component to global memory. (One atomic per facet, as in the first version, spent
// 25 of the build's 33 ms of kernel time serialised on the few boxes of the top levels: profiles/r02_launches_bench_scale0.1.txt.)
constexpr uint32_t kSegSmem = 64;
__global__ void __launch_bounds__(256) seg_bbox_kernel(const float* __restrict__ ctr, const uint32_t* __restrict__ seg, uint32_t nF, uint32_t nseg,
                                                       uint32_t* bb) {
    __shared__ uint32_t sbb[kSegSmem * 6];
    const bool staged = nseg <= kSegSmem;
    if (staged) {
        for (uint32_t i = threadIdx.x; i < nseg * 6; i += blockDim.x) sbb[i] = (i % 6) < 3 ? 0xffffffffu : 0u;
        __syncthreads();
    }
    const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
    const bool in = f < nF;
    const uint32_t s = in ? seg[f] : 0xffffffffu;  // lanes past the end form a group of their own and send nothing
    const unsigned grp = __match_any_sync(0xffffffffu, s);
    const bool leader = (threadIdx.x & 31) == (unsigned)(__ffs(grp) - 1);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const uint32_t k = in ? enc32(ctr[3 * (size_t)f + c]) : 0u;
        const uint32_t lo = __reduce_min_sync(grp, k), hi = __reduce_max_sync(grp, k);
        if (leader && in) {
            if (staged) {
                atomicMin(sbb + 6 * s + c, lo);
                atomicMax(sbb + 6 * s + 3 + c, hi);
            } else {
                atomicMin(bb + 6 * (size_t)s + c, lo);
                atomicMax(bb + 6 * (size_t)s + 3 + c, hi);
            }
        }
    }
    if (staged) {
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < nseg * 6; i += blockDim.x) {
            const uint32_t v = sbb[i];
            if ((i % 6) < 3) { if (v != 0xffffffffu) atomicMin(bb + i, v); }
            else if (v != 0u) atomicMax(bb + i, v);
        }
    }
}
struct SegPlan {
    uint32_t mid;  // absolute position: facets before it go to the left child
    int ax;        // axis to order by, -1: keep (not split, or everything fits the left child)
};
// one thread per segment of this level (build_host_tree::split)
__global__ void seg_plan_kernel(const uint32_t* __restrict__ sb, const uint32_t* __restrict__ se, uint32_t nseg, uint32_t blocks, uint32_t kLeaf,
                                const uint32_t* __restrict__ bb, SegPlan* plan, uint32_t* sb_next, uint32_t* se_next) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nseg) return;
    const uint32_t b = sb[s], e = se[s];
    SegPlan p;
    p.mid = e;
    p.ax = -1;
    if (blocks > 1 && e - b > kLeaf) {
        const unsigned long long left_cap = (unsigned long long)(blocks / 2) * kLeaf;
        const unsigned long long m = (unsigned long long)b + left_cap;
        p.mid = m < e ? (uint32_t)m : e;
        if (p.mid < e) {
            float ext[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) ext[c] = __fsub_rn(dec32(bb[6 * (size_t)s + 3 + c]), dec32(bb[6 * (size_t)s + c]));
            int ax = 0;
            if (ext[1] > ext[ax]) ax = 1;
            if (ext[2] > ext[ax]) ax = 2;
            p.ax = ax;
        }
    }
    plan[s] = p;
    sb_next[2 * s] = b; se_next[2 * s] = p.mid;
    sb_next[2 * s + 1] = p.mid; se_next[2 * s + 1] = e;
}
__global__ void seg_keys_kernel(const float* __restrict__ ctr, const uint32_t* __restrict__ seg, const SegPlan* __restrict__ plan, uint32_t nF,
                                unsigned long long* keys, uint32_t* vals) {
    const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nF) return;
    const uint32_t s = seg[f];
    const int ax = plan[s].ax;
    keys[f] = ((unsigned long long)s << 32) | (ax >= 0 ? (unsigned long long)enc32(ctr[3 * (size_t)f + ax]) : 0ull);
    vals[f] = f;
}
__global__ void seg_assign_kernel(const uint32_t* __restrict__ sorted_f, const SegPlan* __restrict__ plan, uint32_t nF, uint32_t* seg) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nF) return;
    const uint32_t f = sorted_f[p];
    const uint32_t s = seg[f];
    seg[f] = 2u * s + (p >= plan[s].mid ? 1u : 0u);
}
// ---- split axis by surface area (twg_kd_order_device, `sah` form): the facets of a node, sorted along axis a and cut at the heap's
// position, give two child boxes; the axis with the smallest  area(L) nL + area(R) nR  wins (the counts are the same for all axes)
__global__ void fbox_kernel(const double* __restrict__ V, const uint32_t* __restrict__ F, uint32_t nF, float* fbox) {
    const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nF) return;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const double a = V[3 * (size_t)F[3 * (size_t)f] + c], b = V[3 * (size_t)F[3 * (size_t)f + 1] + c], d = V[3 * (size_t)F[3 * (size_t)f + 2] + c];
        fbox[6 * (size_t)f + c] = __double2float_rd(fmin(a, fmin(b, d)));
        fbox[6 * (size_t)f + 3 + c] = __double2float_ru(fmax(a, fmax(b, d)));
    }
}
// cb[2 s + child][0..2] = min, [3..5] = max (encoded) over the facet boxes of that child; same hierarchical reduction as seg_bbox_kernel
__global__ void __launch_bounds__(256) child_box_kernel(const float* __restrict__ fbox, const uint32_t* __restrict__ sorted_f, const uint32_t* __restrict__ seg,
                                                        const SegPlan* __restrict__ plan, uint32_t nF, uint32_t ngrp, uint32_t* cb) {
    __shared__ uint32_t sbb[kSegSmem * 6];
    const bool staged = ngrp <= kSegSmem;
    if (staged) {
        for (uint32_t i = threadIdx.x; i < ngrp * 6; i += blockDim.x) sbb[i] = (i % 6) < 3 ? 0xffffffffu : 0u;
        __syncthreads();
    }
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    const bool in = p < nF;
    uint32_t f = 0, g = 0xffffffffu;
    if (in) {
        f = sorted_f[p];
        const uint32_t s = seg[f];
        g = 2u * s + (p >= plan[s].mid ? 1u : 0u);
    }
    const unsigned grp = __match_any_sync(0xffffffffu, g);
    const bool leader = (threadIdx.x & 31) == (unsigned)(__ffs(grp) - 1);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const uint32_t klo = in ? enc32(fbox[6 * (size_t)f + c]) : 0xffffffffu, khi = in ? enc32(fbox[6 * (size_t)f + 3 + c]) : 0u;
        const uint32_t lo = __reduce_min_sync(grp, klo), hi = __reduce_max_sync(grp, khi);
        if (leader && in) {
            uint32_t* dst = staged ? sbb : cb;
            atomicMin(dst + 6 * (size_t)g + c, lo);
            atomicMax(dst + 6 * (size_t)g + 3 + c, hi);
        }
    }
    if (staged) {
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < ngrp * 6; i += blockDim.x) {
            const uint32_t v = sbb[i];
            if ((i % 6) < 3) { if (v != 0xffffffffu) atomicMin(cb + i, v); }
            else if (v != 0u) atomicMax(cb + i, v);
        }
    }
}
__global__ void axis_cost_kernel(const uint32_t* __restrict__ cb, const SegPlan* __restrict__ plan, const uint32_t* __restrict__ sb, const uint32_t* __restrict__ se,
                                 uint32_t nseg, float* cost) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nseg) return;
    float total = 0.f;
    for (int child = 0; child < 2; ++child) {
        const uint32_t n = child == 0 ? plan[s].mid - sb[s] : se[s] - plan[s].mid;
        if (n == 0) continue;
        const uint32_t* b = cb + 6 * (size_t)(2 * s + child);
        const float dx = dec32(b[3]) - dec32(b[0]), dy = dec32(b[4]) - dec32(b[1]), dz = dec32(b[5]) - dec32(b[2]);
        total += (dx * dy + dy * dz + dz * dx) * (float)n;
    }
    cost[s] = total;
}
// ax[s] = the cheapest of the three axes (ties: the lower axis), for segments that are split at all
__global__ void axis_pick_kernel(const float* __restrict__ cost3, uint32_t nseg, SegPlan* plan) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nseg || plan[s].ax < 0) return;
    int best = 0;
    if (cost3[nseg + s] < cost3[best * (size_t)nseg + s]) best = 1;
    if (cost3[2 * (size_t)nseg + s] < cost3[best * (size_t)nseg + s]) best = 2;
    plan[s].ax = best;
}
__global__ void axis_keys_kernel(const float* __restrict__ ctr, const uint32_t* __restrict__ seg, const SegPlan* __restrict__ plan, uint32_t nF, int ax,
                                 unsigned long long* keys, uint32_t* vals) {
    const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nF) return;
    const uint32_t s = seg[f];
    keys[f] = ((unsigned long long)s << 32) | (plan[s].ax >= 0 ? (unsigned long long)enc32(ctr[3 * (size_t)f + ax]) : 0ull);
    vals[f] = f;
}
// the order of the level: position p takes the facet the winning axis of its segment put there (a segment is the same range of
// positions under all three sorts)
__global__ void axis_merge_kernel(const uint32_t* __restrict__ o0, const uint32_t* __restrict__ o1, const uint32_t* __restrict__ o2, const uint32_t* __restrict__ seg,
                                  const SegPlan* __restrict__ plan, uint32_t nF, uint32_t* order) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nF) return;
    const int ax = plan[seg[o0[p]]].ax;
    order[p] = ax == 1 ? o1[p] : (ax == 2 ? o2[p] : o0[p]);
}
__global__ void block_keys_kernel(const uint32_t* __restrict__ sorted_f, uint32_t nF, uint32_t kLeaf, unsigned long long* keys, uint32_t* vals) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nF) return;
    keys[p] = ((unsigned long long)(p / kLeaf) << 32) | sorted_f[p];
    vals[p] = sorted_f[p];
}

// ---- facets, leaf boxes, heap
__global__ void sf_tris_kernel(const double* __restrict__ V, const uint32_t* __restrict__ F, const uint32_t* __restrict__ order, const uint32_t* __restrict__ canon,
                               uint32_t nF, uint32_t* SF, double* tris) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nF) return;
    const uint32_t f = order[j];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const uint32_t v = canon[F[3 * (size_t)f + k]];
        SF[3 * (size_t)j + k] = v;
#pragma unroll
        for (int c = 0; c < 3; ++c) tris[9 * (size_t)j + 3 * k + c] = V[3 * (size_t)v + c];
    }
}
struct NodeBox {
    double lo[3], hi[3];
    uint32_t nfac, foff;
};
__global__ void leaf_box_kernel(const double* __restrict__ tris, uint32_t nF, uint32_t nBlkP, uint32_t kLeaf, NodeBox* nb) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nBlkP) return;
    const unsigned long long j0u = (unsigned long long)b * kLeaf, j1u = j0u + kLeaf;
    const uint32_t j0 = j0u < nF ? (uint32_t)j0u : nF, j1 = j1u < nF ? (uint32_t)j1u : nF;
    NodeBox x;
#pragma unroll
    for (int c = 0; c < 3; ++c) { x.lo[c] = DBL_MAX; x.hi[c] = -DBL_MAX; }
    for (uint32_t j = j0; j < j1; ++j)
        for (int k = 0; k < 3; ++k)
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const double p = tris[9 * (size_t)j + 3 * k + c];
                x.lo[c] = fmin(x.lo[c], p);
                x.hi[c] = fmax(x.hi[c], p);
            }
    x.nfac = j1 - j0;
    x.foff = j0;
    nb[nBlkP + b] = x;
}
__global__ void heap_up_kernel(NodeBox* nb, uint32_t first) {
    const uint32_t i = first + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 2u * first) return;
    const NodeBox l = nb[2 * i], r = nb[2 * i + 1];
    NodeBox x;
#pragma unroll
    for (int c = 0; c < 3; ++c) { x.lo[c] = fmin(l.lo[c], r.lo[c]); x.hi[c] = fmax(l.hi[c], r.hi[c]); }
    x.nfac = l.nfac + r.nfac;
    x.foff = l.foff;
    nb[i] = x;
}

// ---- exterior edges
constexpr unsigned long long kNoEdge = ~0ull;
__global__ void he_kernel(const uint32_t* __restrict__ SF, uint32_t nF, uint32_t kLeaf, unsigned long long* keys, uint32_t* pay) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 3 * (size_t)nF) return;
    const uint32_t j = (uint32_t)(i / 3), k = (uint32_t)(i % 3);
    const uint32_t a = SF[3 * (size_t)j + k], b = SF[3 * (size_t)j + (k + 1) % 3];
    if (a == b) { keys[i] = kNoEdge; pay[i] = 0; return; }
    const uint32_t blk = j / kLeaf;
    if (a < b) { keys[i] = ((unsigned long long)a << 32) | b; pay[i] = blk; }
    else { keys[i] = ((unsigned long long)b << 32) | a; pay[i] = blk | 0x80000000u; }
}
// one thread per group of equal keys (its first element): the per-level net counts of build_host_tree step 4.
// WRITE = false: gcount[i] = records the group emits; WRITE = true: records at goff[i].
template <bool WRITE>
__global__ void he_group_kernel(const unsigned long long* __restrict__ keys, const uint32_t* __restrict__ pay, size_t n, uint32_t nBlkP, int depth,
                                uint32_t* gcount, const uint32_t* __restrict__ goff, uint32_t* rnode, uint32_t* ra, uint32_t* rb) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long key = keys[i];
    const bool head = key != kNoEdge && (i == 0 || keys[i - 1] != key);
    if (!head) {
        if (!WRITE) gcount[i] = 0;
        return;
    }
    size_t e = i + 1;
    while (e < n && keys[e] == key) ++e;
    const uint32_t u = (uint32_t)(key >> 32), v = (uint32_t)(key & 0xffffffffu);
    uint32_t emitted = 0;
    const uint32_t base = WRITE ? goff[i] : 0u;
    for (int sh = 0; sh <= depth; ++sh) {
        bool any = false;
        int distinct = 0;
        for (size_t k = i; k < e; ++k) {
            const uint32_t node = (nBlkP + (pay[k] & 0x7fffffffu)) >> sh;
            bool first = true;
            for (size_t j = i; j < k; ++j)
                if (((nBlkP + (pay[j] & 0x7fffffffu)) >> sh) == node) { first = false; break; }
            if (!first) continue;
            ++distinct;
            int net = 0;
            for (size_t j = k; j < e; ++j)
                if (((nBlkP + (pay[j] & 0x7fffffffu)) >> sh) == node) net += (pay[j] & 0x80000000u) ? -1 : 1;
            if (net == 0) continue;
            any = true;
            const uint32_t a = net > 0 ? u : v, b = net > 0 ? v : u;
            const int reps = net > 0 ? net : -net;
            for (int r = 0; r < reps; ++r) {
                if (WRITE) { rnode[base + emitted] = node; ra[base + emitted] = a; rb[base + emitted] = b; }
                ++emitted;
            }
        }
        if (!any && distinct == 1) break;  // all members in one node and cancelling: the same for every ancestor
    }
    if (!WRITE) gcount[i] = emitted;
}
__global__ void gather_u32_kernel(const uint32_t* __restrict__ src, const uint32_t* __restrict__ idx, size_t n, uint32_t* dst) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[idx[i]];
}
__global__ void rec_key2_kernel(const uint32_t* __restrict__ rnode, const uint32_t* __restrict__ ra, const uint32_t* __restrict__ idx, size_t n, unsigned long long* keys) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) keys[i] = ((unsigned long long)rnode[idx[i]] << 32) | ra[idx[i]];
}
// roff[node] = first sorted record with rnode >= node, for node in [0, nNodes]
__global__ void rec_offsets_kernel(const uint32_t* __restrict__ rnode_sorted, size_t n, uint32_t nNodes, uint32_t* roff) {
    const uint32_t node = blockIdx.x * blockDim.x + threadIdx.x;
    if (node > nNodes) return;
    size_t lo = 0, hi = n;
    while (lo < hi) {
        const size_t mid = (lo + hi) >> 1;
        if (rnode_sorted[mid] < node) lo = mid + 1; else hi = mid;
    }
    roff[node] = (uint32_t)lo;
}

// ---- cap polylines: build_host_tree step 5 for one node per thread. Edges [e0, e1) of the node are sorted by (a, b).
__global__ void trace_kernel(const double* __restrict__ V, const uint32_t* __restrict__ ra, const uint32_t* __restrict__ rb, const uint32_t* __restrict__ roff,
                             uint32_t nNodes, uint8_t* used, double* scratch /*4 doubles per point, 2 points per edge*/, uint32_t* cap_cnt, double* apex_out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nNodes) return;
    uint32_t cnt = 0;
    double ap[3] = {0.0, 0.0, 0.0};
    const uint32_t e0 = i >= 1 ? roff[i] : 0u, e1 = i >= 1 ? roff[i + 1] : 0u;
    if (e1 > e0) {
        const uint32_t apex = ra[e0];
#pragma unroll
        for (int c = 0; c < 3; ++c) ap[c] = V[3 * (size_t)apex + c];
        double* out = scratch + 8 * (size_t)e0;  // region of 2 * (e1 - e0) points
        uint32_t wp = 0, run_start = 0, run_len = 0;
        auto flush = [&]() {
            if (run_len < 2) wp = run_start;  // a run of fewer than two points closes no fan triangle
            run_start = wp;
            run_len = 0;
        };
        auto visit = [&](uint32_t v) {
            if (v == apex) { flush(); return; }  // a fan triangle with the apex as a corner is degenerate: cut the polyline there
            double* q = out + 4 * (size_t)wp;
            q[0] = V[3 * (size_t)v]; q[1] = V[3 * (size_t)v + 1]; q[2] = V[3 * (size_t)v + 2];
            q[3] = run_len == 0 ? 1.0 : 0.0;
            ++wp; ++run_len;
        };
        for (uint32_t k = e0; k < e1; ++k) {
            if (used[k]) continue;
            visit(ra[k]);
            long long cur = k;
            while (cur >= 0) {
                used[cur] = 1;
                const uint32_t v = rb[cur];
                visit(v);
                // first unused edge that starts at v
                uint32_t lo = e0, hi = e1;
                while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (ra[mid] < v) lo = mid + 1; else hi = mid; }
                cur = -1;
                for (uint32_t t = lo; t < e1 && ra[t] == v; ++t)
                    if (!used[t]) { cur = t; break; }
            }
            flush();
        }
        cnt = wp;
    }
    cap_cnt[i] = cnt;
    apex_out[3 * (size_t)i] = ap[0]; apex_out[3 * (size_t)i + 1] = ap[1]; apex_out[3 * (size_t)i + 2] = ap[2];
}
// one warp per node: its points from the over-allocated region to their final place
__global__ void cap_compact_kernel(const double* __restrict__ scratch, const uint32_t* __restrict__ roff, const uint32_t* __restrict__ cap_cnt,
                                   const uint32_t* __restrict__ cap_off, uint32_t nNodes, double* caps) {
    const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= nNodes || w == 0) return;
    const uint32_t n4 = 4 * cap_cnt[w];
    const double* src = scratch + 8 * (size_t)roff[w];
    double* dst = caps + 4 * (size_t)cap_off[w];
    for (uint32_t k = lane; k < n4; k += 32) dst[k] = src[k];
}
__global__ void nodes_kernel(const NodeBox* __restrict__ nb, const uint32_t* __restrict__ cap_cnt, const uint32_t* __restrict__ cap_off,
                             const double* __restrict__ apex, uint32_t nNodes, WNode* nodes) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nNodes) return;
    WNode nd;
    memset(&nd, 0, sizeof(nd));
    if (i >= 1) {
        const NodeBox x = nb[i];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float inf = __int_as_float(0x7f800000);
            nd.lo[c] = x.nfac ? nextafterf(__double2float_rn(x.lo[c]), -inf) : inf;
            nd.hi[c] = x.nfac ? nextafterf(__double2float_rn(x.hi[c]), inf) : -inf;
        }
        nd.tri_off = x.foff;
        nd.tri_cnt = x.nfac;
        nd.cap_off = cap_off[i];
        nd.cap_cnt = cap_cnt[i];
        nd.apex[0] = apex[3 * (size_t)i]; nd.apex[1] = apex[3 * (size_t)i + 1]; nd.apex[2] = apex[3 * (size_t)i + 2];
    }
    nodes[i] = nd;
}

inline unsigned gridof(size_t n, int block = 256) { return (unsigned)((n + block - 1) / block > 0 ? (n + block - 1) / block : 1); }

}  // namespace

#define WB_CUDA(call)                                                                                \
    do {                                                                                             \
        cudaError_t e__ = (call);                                                                    \
        if (e__ != cudaSuccess) return twg_fail(c, (int)e__, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

template <class K, class Vv>
static int sort_pairs(twg_ctx* c, Arena& A, cudaStream_t st, K*& keys, K*& keys_alt, Vv*& vals, Vv*& vals_alt, size_t n, int begin_bit, int end_bit) {
    TWG_CHECK(c, n <= 0x7fffffffull, TWG_ERR_INVALID_ARG, "too many items for one device sort");
    size_t tmp_bytes = 0;
    WB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys, keys_alt, vals, vals_alt, (int)n, begin_bit, end_bit, st));
    void* tmp = nullptr;
    WB_CUDA(A.temp(tmp_bytes ? tmp_bytes : 16, &tmp));
    WB_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys, keys_alt, vals, vals_alt, (int)n, begin_bit, end_bit, st));
    c->launches += 2 + (end_bit - begin_bit + 7) / 8;
    std::swap(keys, keys_alt);
    std::swap(vals, vals_alt);
    return 0;
}

// Facet order of a count-balanced kd hierarchy whose splits sit exactly where an IMPLICIT HEAP over nLeafP one-facet leaves puts
// them (the surface structure of surface.cu): the facets of a heap node of `blocks` leaves go, sorted along the longest axis of
// their centroid box, blocks / 2 to the left child and the rest to the right one, level by level until a node has `stop_leaves`
// leaves (one 8-wide step of the traversal: the order inside does not change any box above the facets). Same level kernels as the
// winding hierarchy below (leaf size 1); d_order[p] = facet at leaf p. Deterministic: ties keep their previous order (stable sort).
int twg_kd_order_device(twg_ctx* c, cudaStream_t st, const double* dV, const uint32_t* dF, uint32_t nF, uint32_t nLeafP, uint32_t stop_leaves, int sah,
                        uint32_t* d_order) {
    Arena A(st);
    float* ctr;
    uint32_t *seg, *fv, *fv2, *sb[2], *se[2], *bb;
    unsigned long long *fk, *fk2;
    SegPlan* plan;
    if (stop_leaves < 1) stop_leaves = 1;
    uint32_t max_seg = nLeafP / stop_leaves;
    if (max_seg < 1) max_seg = 1;
    WB_CUDA(A.get(&ctr, 3 * (size_t)nF)); WB_CUDA(A.get(&seg, nF)); WB_CUDA(A.get(&fv, nF)); WB_CUDA(A.get(&fv2, nF));
    WB_CUDA(A.get(&fk, nF)); WB_CUDA(A.get(&fk2, nF));
    for (int k = 0; k < 2; ++k) { WB_CUDA(A.get(&sb[k], (size_t)max_seg * 2)); WB_CUDA(A.get(&se[k], (size_t)max_seg * 2)); }
    WB_CUDA(A.get(&bb, 6 * (size_t)max_seg)); WB_CUDA(A.get(&plan, max_seg));
    float *fbox = nullptr, *cost3 = nullptr;
    uint32_t *cb = nullptr, *oax[3] = {nullptr, nullptr, nullptr};
    if (sah) {
        WB_CUDA(A.get(&fbox, 6 * (size_t)nF)); WB_CUDA(A.get(&cost3, 3 * (size_t)max_seg)); WB_CUDA(A.get(&cb, 12 * (size_t)max_seg));
        for (int ax = 0; ax < 3; ++ax) WB_CUDA(A.get(&oax[ax], nF));
        TWG_LAUNCH(c, fbox_kernel, gridof(nF), 256, 0, st, dV, dF, nF, fbox);
    }
    TWG_LAUNCH(c, ctr_kernel, gridof(nF), 256, 0, st, dV, dF, nF, ctr);
    WB_CUDA(cudaMemsetAsync(seg, 0, sizeof(uint32_t) * (size_t)nF, st));
    {
        const uint32_t be[2] = {0u, nF};
        WB_CUDA(cudaMemcpyAsync(sb[0], &be[0], 4, cudaMemcpyHostToDevice, st));
        WB_CUDA(cudaMemcpyAsync(se[0], &be[1], 4, cudaMemcpyHostToDevice, st));
        WB_CUDA(cudaStreamSynchronize(st));  // `be` lives on this stack frame
    }
    TWG_LAUNCH(c, iota_kernel, gridof(nF), 256, 0, st, fv, nF);
    for (int lvl = 0; (nLeafP >> lvl) > stop_leaves; ++lvl) {
        const uint32_t nseg = 1u << lvl, blocks = nLeafP >> lvl;
        const int cur = lvl & 1, nxt = cur ^ 1;
        TWG_LAUNCH(c, bb_init_kernel, gridof(6 * (size_t)nseg), 256, 0, st, bb, 6 * (size_t)nseg);
        TWG_LAUNCH(c, seg_bbox_kernel, gridof(nF), 256, 0, st, (const float*)ctr, (const uint32_t*)seg, nF, nseg, bb);
        TWG_LAUNCH(c, seg_plan_kernel, gridof(nseg), 256, 0, st, (const uint32_t*)sb[cur], (const uint32_t*)se[cur], nseg, blocks, 1u, (const uint32_t*)bb, plan, sb[nxt],
                   se[nxt]);
        if (sah) {
            // three sorts, two child boxes each, cheapest axis per segment; ties inside a sort keep the facet id order (keys are
            // rebuilt from the ids, values start as iota: the result does not depend on the previous level's order)
            for (int ax = 0; ax < 3; ++ax) {
                TWG_LAUNCH(c, axis_keys_kernel, gridof(nF), 256, 0, st, (const float*)ctr, (const uint32_t*)seg, (const SegPlan*)plan, nF, ax, fk, fv);
                TWG_TRY(sort_pairs(c, A, st, fk, fk2, fv, fv2, nF, 0, 32 + (lvl > 0 ? lvl : 1)));
                WB_CUDA(cudaMemcpyAsync(oax[ax], fv, sizeof(uint32_t) * (size_t)nF, cudaMemcpyDeviceToDevice, st));
                TWG_LAUNCH(c, bb_init_kernel, gridof(12 * (size_t)nseg), 256, 0, st, cb, 12 * (size_t)nseg);
                TWG_LAUNCH(c, child_box_kernel, gridof(nF), 256, 0, st, (const float*)fbox, (const uint32_t*)oax[ax], (const uint32_t*)seg, (const SegPlan*)plan, nF, 2 * nseg,
                           cb);
                TWG_LAUNCH(c, axis_cost_kernel, gridof(nseg), 256, 0, st, (const uint32_t*)cb, (const SegPlan*)plan, (const uint32_t*)sb[cur], (const uint32_t*)se[cur], nseg,
                           cost3 + (size_t)ax * nseg);
            }
            TWG_LAUNCH(c, axis_pick_kernel, gridof(nseg), 256, 0, st, (const float*)cost3, nseg, plan);
            TWG_LAUNCH(c, axis_merge_kernel, gridof(nF), 256, 0, st, (const uint32_t*)oax[0], (const uint32_t*)oax[1], (const uint32_t*)oax[2], (const uint32_t*)seg,
                       (const SegPlan*)plan, nF, fv);
        } else {
            TWG_LAUNCH(c, seg_keys_kernel, gridof(nF), 256, 0, st, (const float*)ctr, (const uint32_t*)seg, (const SegPlan*)plan, nF, fk, fv);
            TWG_TRY(sort_pairs(c, A, st, fk, fk2, fv, fv2, nF, 0, 32 + (lvl > 0 ? lvl : 1)));
        }
        TWG_LAUNCH(c, seg_assign_kernel, gridof(nF), 256, 0, st, (const uint32_t*)fv, (const SegPlan*)plan, nF, seg);
    }
    WB_CUDA(cudaMemcpyAsync(d_order, fv, sizeof(uint32_t) * (size_t)nF, cudaMemcpyDeviceToDevice, st));
    WB_CUDA(cudaStreamSynchronize(st));  // the arena frees behind the copy (stream-ordered); callers may use d_order on any stream
    return 0;
}

int twg_winding_build_device(twg_ctx* c, const double* dV, uint32_t nV, const uint32_t* dF, uint32_t nF, uint32_t kLeaf, DeviceTree* out) {
    cudaStream_t st = c->streams[0];
    Arena A(st);
    // ---- geometry of the heap (the same arithmetic as the host build)
    uint32_t nBlkP = 1;
    while ((uint64_t)nBlkP * kLeaf < nF) nBlkP <<= 1;
    kLeaf = ((nF + nBlkP - 1) / nBlkP + 1u) & ~1u;
    if (kLeaf < 2) kLeaf = 2;
    int depth = 0;
    while ((1u << depth) < nBlkP) ++depth;
    const uint32_t nNodes = 2 * nBlkP;

    // ---- 1. vertex merge
    uint32_t *idx, *idx2, *canon, *vstart;
    unsigned long long *vk, *vk2;
    WB_CUDA(A.get(&idx, nV)); WB_CUDA(A.get(&idx2, nV)); WB_CUDA(A.get(&canon, nV)); WB_CUDA(A.get(&vstart, nV));
    WB_CUDA(A.get(&vk, nV)); WB_CUDA(A.get(&vk2, nV));
    TWG_LAUNCH(c, iota_kernel, gridof(nV), 256, 0, st, idx, nV);
    for (int comp = 2; comp >= 0; --comp) {
        TWG_LAUNCH(c, vkey_kernel, gridof(nV), 256, 0, st, dV, (const uint32_t*)idx, nV, comp, vk);
        TWG_TRY(sort_pairs(c, A, st, vk, vk2, idx, idx2, nV, 0, 64));
    }
    TWG_LAUNCH(c, vrun_kernel, gridof(nV), 256, 0, st, dV, (const uint32_t*)idx, nV, vstart);
    {
        size_t tb = 0;
        WB_CUDA(cub::DeviceScan::InclusiveScan(nullptr, tb, vstart, vstart, MaxU32(), (int)nV, st));
        void* tmp;
        WB_CUDA(A.temp(tb ? tb : 16, &tmp));
        WB_CUDA(cub::DeviceScan::InclusiveScan(tmp, tb, vstart, vstart, MaxU32(), (int)nV, st));
        c->launches += 2;
    }
    TWG_LAUNCH(c, canon_kernel, gridof(nV), 256, 0, st, (const uint32_t*)idx, (const uint32_t*)vstart, nV, canon);

    // ---- 2. kd order
    float* ctr;
    uint32_t *seg, *fv, *fv2, *sb[2], *se[2], *bb;
    unsigned long long *fk, *fk2;
    SegPlan* plan;
    WB_CUDA(A.get(&ctr, 3 * (size_t)nF)); WB_CUDA(A.get(&seg, nF)); WB_CUDA(A.get(&fv, nF)); WB_CUDA(A.get(&fv2, nF));
    WB_CUDA(A.get(&fk, nF)); WB_CUDA(A.get(&fk2, nF));
    for (int k = 0; k < 2; ++k) { WB_CUDA(A.get(&sb[k], (size_t)nBlkP * 2)); WB_CUDA(A.get(&se[k], (size_t)nBlkP * 2)); }
    WB_CUDA(A.get(&bb, 6 * (size_t)nBlkP)); WB_CUDA(A.get(&plan, nBlkP));
    TWG_LAUNCH(c, ctr_kernel, gridof(nF), 256, 0, st, dV, dF, nF, ctr);
    WB_CUDA(cudaMemsetAsync(seg, 0, sizeof(uint32_t) * (size_t)nF, st));
    {
        const uint32_t b0 = 0, e0 = nF;
        WB_CUDA(cudaMemcpyAsync(sb[0], &b0, 4, cudaMemcpyHostToDevice, st));
        WB_CUDA(cudaMemcpyAsync(se[0], &e0, 4, cudaMemcpyHostToDevice, st));
        WB_CUDA(cudaStreamSynchronize(st));  // b0 / e0 live on this stack frame
    }
    TWG_LAUNCH(c, iota_kernel, gridof(nF), 256, 0, st, fv, nF);  // level-0 "sorted order" (also the final order of a one-block heap)
    for (int lvl = 0; lvl < depth; ++lvl) {
        const uint32_t nseg = 1u << lvl, blocks = nBlkP >> lvl;
        const int cur = lvl & 1, nxt = cur ^ 1;
        TWG_LAUNCH(c, bb_init_kernel, gridof(6 * (size_t)nseg), 256, 0, st, bb, 6 * (size_t)nseg);
        TWG_LAUNCH(c, seg_bbox_kernel, gridof(nF), 256, 0, st, (const float*)ctr, (const uint32_t*)seg, nF, nseg, bb);
        TWG_LAUNCH(c, seg_plan_kernel, gridof(nseg), 256, 0, st, (const uint32_t*)sb[cur], (const uint32_t*)se[cur], nseg, blocks, kLeaf, (const uint32_t*)bb, plan,
                   sb[nxt], se[nxt]);
        TWG_LAUNCH(c, seg_keys_kernel, gridof(nF), 256, 0, st, (const float*)ctr, (const uint32_t*)seg, (const SegPlan*)plan, nF, fk, fv);
        TWG_TRY(sort_pairs(c, A, st, fk, fk2, fv, fv2, nF, 0, 32 + (lvl > 0 ? lvl : 1)));
        TWG_LAUNCH(c, seg_assign_kernel, gridof(nF), 256, 0, st, (const uint32_t*)fv, (const SegPlan*)plan, nF, seg);
    }
    // canonical order inside the blocks: by facet id
    uint32_t* order;
    WB_CUDA(A.get(&order, nF));
    TWG_LAUNCH(c, block_keys_kernel, gridof(nF), 256, 0, st, (const uint32_t*)fv, nF, kLeaf, fk, fv2);
    {
        uint32_t* v1 = fv2;
        uint32_t* v2 = order;
        TWG_TRY(sort_pairs(c, A, st, fk, fk2, v1, v2, nF, 0, 64));
        order = v1;  // sort_pairs swapped: v1 now names the sorted values
    }

    // ---- 3. facets, boxes, heap
    uint32_t* SF;
    double* tris;
    NodeBox* nb;
    WB_CUDA(A.get(&SF, 3 * (size_t)nF));
    WB_CUDA(cudaMalloc(&tris, sizeof(double) * 9 * ((size_t)nF + 2)));
    out->tris = tris;
    WB_CUDA(cudaMemsetAsync(tris + 9 * (size_t)nF, 0, sizeof(double) * 18, st));
    WB_CUDA(A.get(&nb, nNodes));
    TWG_LAUNCH(c, sf_tris_kernel, gridof(nF), 256, 0, st, dV, dF, (const uint32_t*)order, (const uint32_t*)canon, nF, SF, tris);
    TWG_LAUNCH(c, leaf_box_kernel, gridof(nBlkP, 128), 128, 0, st, (const double*)tris, nF, nBlkP, kLeaf, nb);
    for (uint32_t first = nBlkP / 2; first >= 1; first >>= 1) {
        TWG_LAUNCH(c, heap_up_kernel, gridof(first), 256, 0, st, nb, first);
        if (first == 1) break;
    }

    // ---- 4. exterior edges
    const size_t nHE = 3 * (size_t)nF;
    unsigned long long *hk, *hk2;
    uint32_t *hp, *hp2, *gcount, *goff;
    WB_CUDA(A.get(&hk, nHE)); WB_CUDA(A.get(&hk2, nHE)); WB_CUDA(A.get(&hp, nHE)); WB_CUDA(A.get(&hp2, nHE));
    WB_CUDA(A.get(&gcount, nHE + 1)); WB_CUDA(A.get(&goff, nHE + 1));
    TWG_LAUNCH(c, he_kernel, gridof(nHE), 256, 0, st, (const uint32_t*)SF, nF, kLeaf, hk, hp);
    TWG_TRY(sort_pairs(c, A, st, hk, hk2, hp, hp2, nHE, 0, 64));
    TWG_LAUNCH(c, (he_group_kernel<false>), gridof(nHE), 256, 0, st, (const unsigned long long*)hk, (const uint32_t*)hp, nHE, nBlkP, depth, gcount,
               (const uint32_t*)nullptr, (uint32_t*)nullptr, (uint32_t*)nullptr, (uint32_t*)nullptr);
    WB_CUDA(cudaMemsetAsync(gcount + nHE, 0, 4, st));
    {
        size_t tb = 0;
        WB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, gcount, goff, (int)(nHE + 1), st));
        void* tmp;
        WB_CUDA(A.temp(tb ? tb : 16, &tmp));
        WB_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tb, gcount, goff, (int)(nHE + 1), st));
        c->launches += 2;
    }
    uint32_t nRec = 0;
    WB_CUDA(cudaMemcpyAsync(&nRec, goff + nHE, 4, cudaMemcpyDeviceToHost, st));
    WB_CUDA(cudaStreamSynchronize(st));
    uint32_t *rnode, *ra, *rb, *ri, *ri2, *k32, *k32b, *rnode_s, *ra_s, *rb_s, *roff;
    unsigned long long *k64, *k64b;
    WB_CUDA(A.get(&rnode, nRec)); WB_CUDA(A.get(&ra, nRec)); WB_CUDA(A.get(&rb, nRec));
    WB_CUDA(A.get(&ri, nRec)); WB_CUDA(A.get(&ri2, nRec)); WB_CUDA(A.get(&k32, nRec)); WB_CUDA(A.get(&k32b, nRec));
    WB_CUDA(A.get(&k64, nRec)); WB_CUDA(A.get(&k64b, nRec));
    WB_CUDA(A.get(&rnode_s, nRec)); WB_CUDA(A.get(&ra_s, nRec)); WB_CUDA(A.get(&rb_s, nRec));
    WB_CUDA(A.get(&roff, (size_t)nNodes + 2));
    if (nRec) {
        TWG_LAUNCH(c, (he_group_kernel<true>), gridof(nHE), 256, 0, st, (const unsigned long long*)hk, (const uint32_t*)hp, nHE, nBlkP, depth, (uint32_t*)nullptr,
                   (const uint32_t*)goff, rnode, ra, rb);
        // sort the records by (node, a, b): stable by b, then stable by (node, a)
        TWG_LAUNCH(c, iota_kernel, gridof(nRec), 256, 0, st, ri, nRec);
        WB_CUDA(cudaMemcpyAsync(k32, rb, sizeof(uint32_t) * (size_t)nRec, cudaMemcpyDeviceToDevice, st));
        TWG_TRY(sort_pairs(c, A, st, k32, k32b, ri, ri2, nRec, 0, 32));
        TWG_LAUNCH(c, rec_key2_kernel, gridof(nRec), 256, 0, st, (const uint32_t*)rnode, (const uint32_t*)ra, (const uint32_t*)ri, (size_t)nRec, k64);
        TWG_TRY(sort_pairs(c, A, st, k64, k64b, ri, ri2, nRec, 0, 64));
        TWG_LAUNCH(c, gather_u32_kernel, gridof(nRec), 256, 0, st, (const uint32_t*)rnode, (const uint32_t*)ri, (size_t)nRec, rnode_s);
        TWG_LAUNCH(c, gather_u32_kernel, gridof(nRec), 256, 0, st, (const uint32_t*)ra, (const uint32_t*)ri, (size_t)nRec, ra_s);
        TWG_LAUNCH(c, gather_u32_kernel, gridof(nRec), 256, 0, st, (const uint32_t*)rb, (const uint32_t*)ri, (size_t)nRec, rb_s);
    }
    TWG_LAUNCH(c, rec_offsets_kernel, gridof((size_t)nNodes + 1), 256, 0, st, (const uint32_t*)rnode_s, (size_t)nRec, nNodes, roff);

    // ---- 5. cap polylines
    uint8_t* used;
    double *scratch, *apex;
    uint32_t *cap_cnt, *cap_off;
    WB_CUDA(A.get(&used, (size_t)nRec)); WB_CUDA(A.get(&scratch, 8 * (size_t)nRec)); WB_CUDA(A.get(&apex, 3 * (size_t)nNodes));
    WB_CUDA(A.get(&cap_cnt, (size_t)nNodes + 1)); WB_CUDA(A.get(&cap_off, (size_t)nNodes + 1));
    WB_CUDA(cudaMemsetAsync(used, 0, nRec ? nRec : 1, st));
    TWG_LAUNCH(c, trace_kernel, gridof(nNodes, 64), 64, 0, st, dV, (const uint32_t*)ra_s, (const uint32_t*)rb_s, (const uint32_t*)roff, nNodes, used, scratch, cap_cnt,
               apex);
    WB_CUDA(cudaMemsetAsync(cap_cnt + nNodes, 0, 4, st));
    {
        size_t tb = 0;
        WB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, cap_cnt, cap_off, (int)(nNodes + 1), st));
        void* tmp;
        WB_CUDA(A.temp(tb ? tb : 16, &tmp));
        WB_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tb, cap_cnt, cap_off, (int)(nNodes + 1), st));
        c->launches += 2;
    }
    uint32_t nCap = 0;
    WB_CUDA(cudaMemcpyAsync(&nCap, cap_off + nNodes, 4, cudaMemcpyDeviceToHost, st));
    WB_CUDA(cudaStreamSynchronize(st));
    WB_CUDA(cudaMalloc(&out->caps, sizeof(double) * (4 * (size_t)nCap + 8)));
    WB_CUDA(cudaMemsetAsync(out->caps + 4 * (size_t)nCap, 0, sizeof(double) * 8, st));
    TWG_LAUNCH(c, cap_compact_kernel, gridof((size_t)nNodes * 32), 256, 0, st, (const double*)scratch, (const uint32_t*)roff, (const uint32_t*)cap_cnt,
               (const uint32_t*)cap_off, nNodes, out->caps);
    WB_CUDA(cudaMalloc(&out->nodes, sizeof(WNode) * (size_t)nNodes));
    TWG_LAUNCH(c, nodes_kernel, gridof(nNodes), 256, 0, st, (const NodeBox*)nb, (const uint32_t*)cap_cnt, (const uint32_t*)cap_off, (const double*)apex, nNodes,
               out->nodes);
    WNode root;
    WB_CUDA(cudaMemcpyAsync(&root, out->nodes + 1, sizeof(WNode), cudaMemcpyDeviceToHost, st));
    WB_CUDA(cudaStreamSynchronize(st));
    for (int k = 0; k < 3; ++k) { out->root_lo[k] = root.lo[k]; out->root_hi[k] = root.hi[k]; }
    out->nBlkP = nBlkP;
    out->n_nodes = nNodes;
    out->n_caps = (uint64_t)nCap + 2;  // the host build counts its 8 padding doubles as two points
    return 0;
}
