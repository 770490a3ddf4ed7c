// surface.cu -- on-device construction of the envelope structure (replaces the constructor of
// GEO::MeshFacetsAABBWithEps, src/tetwild/geogram/mesh_AABB.cpp:356-379: Morton reorder :368-370, bbox fill :63-141).
//
// Pipeline (all on the device): facet bbox reduction -> facet order (option surface_order: space-filling curve of the centroids
// + radix sort, or -- the default -- the kd order of winding_build.cu: median splits along the longest axis, cut where the
// implicit heap cuts its leaf range) -> per-facet TriRec + leaf boxes -> bottom-up union of the implicit heap, one launch per
// level. Every heap node is a contiguous range of the order, so the order is what decides how much sibling boxes overlap:
// envelope 2.25 (Z curve) -> 2.04 (Hilbert) -> 1.78 ms (kd) per 10 M points with the same kernels and the same answers.
#include <cub/device/device_radix_sort.cuh>
#include "surface.cuh"

namespace {

__device__ __forceinline__ unsigned long long enc(double d) {  // order-preserving double -> u64
    unsigned long long b = (unsigned long long)__double_as_longlong(d);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__host__ __device__ inline double dec(unsigned long long u) {
    unsigned long long b = (u >> 63) ? (u & 0x7fffffffffffffffull) : ~u;
    double d;
    memcpy(&d, &b, 8);
    return d;
}

__global__ void init_bounds_kernel(unsigned long long* bounds) {
    if (threadIdx.x < 3) bounds[threadIdx.x] = ~0ull;      // min slots
    else if (threadIdx.x < 6) bounds[threadIdx.x] = 0ull;  // max slots
    else if (threadIdx.x < 8) bounds[threadIdx.x] = 0ull;  // [6]: facets that name a vertex >= nV
}

// also validates the facet indices (a facet that names a vertex >= nV is counted, never dereferenced; the build then fails)
__global__ void __launch_bounds__(256) bounds_kernel(const double* __restrict__ V, uint32_t nV, const uint32_t* __restrict__ F, uint32_t nF,
                                                     unsigned long long* bounds) {
    double lo[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, hi[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
    for (uint32_t f = blockIdx.x * blockDim.x + threadIdx.x; f < nF; f += gridDim.x * blockDim.x) {
        if (F[3 * (size_t)f] >= nV || F[3 * (size_t)f + 1] >= nV || F[3 * (size_t)f + 2] >= nV) {
            atomicAdd(bounds + 6, 1ull);
            continue;
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const double* p = V + 3 * (size_t)F[3 * (size_t)f + k];
#pragma unroll
            for (int c = 0; c < 3; ++c) { lo[c] = fmin(lo[c], p[c]); hi[c] = fmax(hi[c], p[c]); }
        }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[c] = fmin(lo[c], __shfl_xor_sync(0xffffffffu, lo[c], o));
            hi[c] = fmax(hi[c], __shfl_xor_sync(0xffffffffu, hi[c], o));
        }
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int c = 0; c < 3; ++c) { atomicMin(bounds + c, enc(lo[c])); atomicMax(bounds + 3 + c, enc(hi[c])); }
    }
}

__device__ __forceinline__ unsigned long long spread3(unsigned long long v) {
    v &= 0x1fffffull;
    v = (v | v << 32) & 0x1f00000000ffffull;
    v = (v | v << 16) & 0x1f0000ff0000ffull;
    v = (v | v << 8) & 0x100f00f00f00f00full;
    v = (v | v << 4) & 0x10c30c30c30c30c3ull;
    v = (v | v << 2) & 0x1249249249249249ull;
    return v;
}

// Hilbert index of a point on a 2^21 grid (J. Skilling, "Programming the Hilbert curve", 2004: Gray-code untangling of the
// transposed axes, then bit interleaving). Facets are ordered along the Hilbert curve rather than the Z curve: every range of
// consecutive facets -- and every node of the implicit heap is one -- is then a CONNECTED piece of the curve, without the long
// jumps of the Z order, so sibling boxes overlap less and a query enters fewer subtrees (option surface_order: 1 Hilbert, 0 Morton;
// 2, the default, replaces the curve by the kd order, see the top of this file).
__device__ __forceinline__ unsigned long long hilbert63(uint32_t x, uint32_t y, uint32_t z) {
    uint32_t X[3] = {x, y, z};
    const uint32_t M = 1u << 20;
    // inverse undo excess work
    for (uint32_t Q = M; Q > 1; Q >>= 1) {
        const uint32_t P = Q - 1;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            if (X[i] & Q) X[0] ^= P;
            else { const uint32_t t = (X[0] ^ X[i]) & P; X[0] ^= t; X[i] ^= t; }
        }
    }
    // Gray encode
    X[1] ^= X[0];
    X[2] ^= X[1];
    uint32_t t = 0;
    for (uint32_t Q = M; Q > 1; Q >>= 1)
        if (X[2] & Q) t ^= Q - 1;
#pragma unroll
    for (int i = 0; i < 3; ++i) X[i] ^= t;
    // interleave: bit b of X[0] is the most significant of the three bits of level b
    return (spread3(X[0]) << 2) | (spread3(X[1]) << 1) | spread3(X[2]);
}

__global__ void __launch_bounds__(256) morton_kernel(const double* __restrict__ V, const uint32_t* __restrict__ F, uint32_t nF,
                                                     const unsigned long long* __restrict__ bounds, unsigned long long* keys, uint32_t* vals, int hilbert) {
    const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nF) return;
    unsigned long long code = 0;
    uint32_t qd[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const double lo = dec(bounds[c]), hi = dec(bounds[3 + c]);
        double ctr = (V[3 * (size_t)F[3 * (size_t)f] + c] + V[3 * (size_t)F[3 * (size_t)f + 1] + c] + V[3 * (size_t)F[3 * (size_t)f + 2] + c]) * (1.0 / 3.0);
        const double ext = hi - lo;
        double u = ext > 0.0 ? (ctr - lo) / ext : 0.0;
        u = fmin(fmax(u, 0.0), 1.0);
        qd[c] = (uint32_t)(u * 2097151.0);
        code |= spread3((unsigned long long)qd[c]) << c;
    }
    keys[f] = hilbert ? hilbert63(qd[0], qd[1], qd[2]) : code;
    vals[f] = f;
}

__device__ __forceinline__ void store_half(NodePair* pairs, uint32_t node /*child index*/, float lx, float ly, float lz, float hx, float hy, float hz) {
    float* dst = reinterpret_cast<float*>(pairs + (node >> 1)) + (node & 1u) * 6;
    dst[0] = lx; dst[1] = ly; dst[2] = lz; dst[3] = hx; dst[4] = hy; dst[5] = hz;
}

__global__ void __launch_bounds__(128) leaf_kernel(const double* __restrict__ V, const uint32_t* __restrict__ F, const uint32_t* __restrict__ order,
                                                   uint32_t nF, uint32_t nLeafP, tw::TriRec* tris, double* triV, NodePair* pairs, TriBound* tb) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nLeafP) return;
    if (j >= nF) {
        const float inf = __int_as_float(0x7f800000);
        store_half(pairs, nLeafP + j, inf, inf, inf, -inf, -inf, -inf);
        return;
    }
    const uint32_t f = order[j];
    double tv[9];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const double* p = V + 3 * (size_t)F[3 * (size_t)f + k];
        tv[3 * k] = p[0]; tv[3 * k + 1] = p[1]; tv[3 * k + 2] = p[2];
    }
    tw::TriRec r;
    tw::make_trirec(tv, tv + 3, tv + 6, f, r);
    tris[j] = r;
#pragma unroll
    for (int k = 0; k < 9; ++k) triV[(size_t)j * 9 + k] = tv[k];
    float lo[3], hi[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        lo[c] = __double2float_rd(fmin(fmin(tv[c], tv[3 + c]), tv[6 + c]));
        hi[c] = __double2float_ru(fmax(fmax(tv[c], tv[3 + c]), tv[6 + c]));
    }
    store_half(pairs, nLeafP + j, lo[0], lo[1], lo[2], hi[0], hi[1], hi[2]);
    TriBound B;
    tw::make_bound(tv, (r.flags & 1u) != 0, B);
    tb[j] = B;
}

// nodes [first, 2*first): union of their two child boxes -> their slot in the parent record
__global__ void __launch_bounds__(256) level_kernel(NodePair* pairs, uint32_t first) {
    const uint32_t i = first + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 2u * first) return;
    const float* s = reinterpret_cast<const float*>(pairs + i);
    store_half(pairs, i, fminf(s[0], s[6]), fminf(s[1], s[7]), fminf(s[2], s[8]), fmaxf(s[3], s[9]), fmaxf(s[4], s[10]), fmaxf(s[5], s[11]));
}

int build(twg_ctx* c, const double* dV, uint32_t nV, const uint32_t* dF, uint32_t nF, twg_surface** out) {
    cudaStream_t st = c->streams[0];
    twg_surface* s = new twg_surface;
    s->ctx = c;
    s->nF = nF;
    uint32_t lp = 2;
    while (lp < nF) lp <<= 1;
    s->nLeafP = lp;
    unsigned long long *bounds = nullptr, *keys = nullptr, *keys2 = nullptr;
    uint32_t *vals = nullptr, *vals2 = nullptr;
    void* tmp = nullptr;
    size_t tmp_bytes = 0;
    int rc = 0;
    auto fail = [&](int code) {
        cudaFree(bounds); cudaFree(keys); cudaFree(keys2); cudaFree(vals); cudaFree(vals2); cudaFree(tmp);
        twg_surface_destroy(s);
        return code;
    };
#define B_CUDA(call)                                                                                              \
    do {                                                                                                          \
        cudaError_t e__ = (call);                                                                                 \
        if (e__ != cudaSuccess) return fail(twg_fail(c, (int)e__, cudaGetErrorString(e__), __FILE__, __LINE__)); \
    } while (0)
    B_CUDA(cudaMalloc(&s->pairs, sizeof(NodePair) * (size_t)lp));
    B_CUDA(cudaMalloc(&s->tris, sizeof(tw::TriRec) * (size_t)nF));
    B_CUDA(cudaMalloc(&s->triV, sizeof(double) * 9 * (size_t)nF));
    B_CUDA(cudaMalloc(&s->tb, sizeof(TriBound) * (size_t)nF));
    B_CUDA(cudaMalloc(&bounds, 8 * sizeof(unsigned long long)));
    B_CUDA(cudaMalloc(&keys, sizeof(unsigned long long) * (size_t)nF));
    B_CUDA(cudaMalloc(&keys2, sizeof(unsigned long long) * (size_t)nF));
    B_CUDA(cudaMalloc(&vals, sizeof(uint32_t) * (size_t)nF));
    B_CUDA(cudaMalloc(&vals2, sizeof(uint32_t) * (size_t)nF));
    B_CUDA(cudaMemsetAsync(s->pairs, 0, sizeof(NodePair) * (size_t)lp, st));
    init_bounds_kernel<<<1, 32, 0, st>>>(bounds);
    c->launches++;
    {
        unsigned g = (nF + 255) / 256;
        if (g > (unsigned)c->sm_count * 8) g = c->sm_count * 8;
        bounds_kernel<<<g, 256, 0, st>>>(dV, nV, dF, nF, bounds);
        c->launches++;
        // the later kernels dereference F: stop here if an index is out of range
        unsigned long long bad = 0;
        B_CUDA(cudaMemcpyAsync(&bad, bounds + 6, sizeof(bad), cudaMemcpyDeviceToHost, st));
        B_CUDA(cudaStreamSynchronize(st));
        if (bad != 0) return fail(twg_fail(c, TWG_ERR_INVALID_ARG, "facet references a vertex out of range", __FILE__, __LINE__));
        morton_kernel<<<(nF + 255) / 256, 256, 0, st>>>(dV, dF, nF, bounds, keys, vals, c->opt.surface_order);
        c->launches++;
    }
    static_assert(sizeof(int) == 4, "cub takes int item counts; nF < 2^31 is checked by the callers");
    B_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys, keys2, vals, vals2, (int)nF, 0, 63, st));
    B_CUDA(cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 16));
    B_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys, keys2, vals, vals2, (int)nF, 0, 63, st));
    c->launches += 4;  // cub: histogram + onesweep passes (approximate, library kernels)
    if (c->opt.surface_order >= 2 && lp > 8) {
        // kd order (winding_build.cu): count-balanced median splits along the longest axis, aligned with the heap's node ranges
        const int krc = twg_kd_order_device(c, st, dV, dF, nF, lp, 8u, c->opt.surface_order >= 3 ? 1 : 0, vals2);
        if (krc != 0) return fail(krc);
    }
    leaf_kernel<<<(lp + 127) / 128, 128, 0, st>>>(dV, dF, vals2, nF, lp, s->tris, s->triV, s->pairs, s->tb);
    c->launches++;
    for (uint32_t first = lp / 2; first >= 2; first >>= 1) {
        level_kernel<<<(first + 255) / 256, 256, 0, st>>>(s->pairs, first);
        c->launches++;
    }
    B_CUDA(cudaGetLastError());
    {
        unsigned long long hb[6];
        B_CUDA(cudaMemcpyAsync(hb, bounds, sizeof(hb), cudaMemcpyDeviceToHost, st));
        B_CUDA(cudaStreamSynchronize(st));
        for (int k = 0; k < 6; ++k) s->bbox[k] = dec(hb[k]);
        for (int k = 0; k < 3; ++k) {
            const double m = 0.05 * (s->bbox[3 + k] - s->bbox[k]);
            s->sort_box[k] = s->bbox[k] - m;
            s->sort_box[3 + k] = s->bbox[3 + k] + m;
        }
    }
#undef B_CUDA
    cudaFree(bounds); cudaFree(keys); cudaFree(keys2); cudaFree(vals); cudaFree(vals2); cudaFree(tmp);
    *out = s;
    return rc;
}

}  // namespace

extern "C" {

void twg_surface_destroy(twg_surface* s) {
    if (!s) return;
    if (!s->replicas.empty()) {
        for (twg_surface* r : s->replicas) twg_surface_destroy(r);
        delete s;
        return;
    }
    if (s->ctx) cudaSetDevice(s->ctx->device);
    cudaFree(s->pairs);
    cudaFree(s->tris);
    cudaFree(s->triV);
    cudaFree(s->tb);
    delete s;
}

uint32_t twg_surface_num_facets(const twg_surface* s) { return s ? s->nF : 0; }
twg_surface* twg_surface_replica(twg_surface* s, int k) {
    if (!s) return nullptr;
    if (s->replicas.empty()) return k == 0 ? s : nullptr;
    return (k >= 0 && k < (int)s->replicas.size()) ? s->replicas[k] : nullptr;
}

int twg_surface_create_dev(twg_ctx* c, const double* dV, uint32_t nV, const uint32_t* dF, uint32_t nF, twg_surface** out) {
    TWG_CHECK(c, c && dV && dF && out, TWG_ERR_INVALID_ARG, "null argument");
    TWG_CHECK(c, nF > 0 && nF < 0x7fffffffu, TWG_ERR_INVALID_ARG, "surface must have 1 .. 2^31-2 facets");
    TWG_CHECK(c, !twg_is_multi(c), TWG_ERR_INVALID_ARG, "_dev entry points take a one-device context (twg_device_context)");
    TWG_CUDA(c, cudaSetDevice(c->device));
    return build(c, dV, nV, dF, nF, out);
}

int twg_surface_create(twg_ctx* c, const double* V, uint32_t nV, const uint32_t* F, uint32_t nF, twg_surface** out) {
    TWG_CHECK(c, c && V && F && out, TWG_ERR_INVALID_ARG, "null argument");
    TWG_CHECK(c, nF > 0 && nF < 0x7fffffffu, TWG_ERR_INVALID_ARG, "surface must have 1 .. 2^31-2 facets");
    if (twg_is_multi(c)) {  // one replica per device, each built by its own device from the caller's arrays
        twg_surface* s = new twg_surface;
        s->ctx = c;
        s->nF = nF;
        s->replicas.assign(c->children.size(), nullptr);
        const int rc = twg_multi_run(c, [&](int k, twg_ctx* child) { return twg_surface_create(child, V, nV, F, nF, &s->replicas[k]); });
        if (rc != 0) { twg_surface_destroy(s); return rc; }
        s->nLeafP = s->replicas[0]->nLeafP;
        for (int k = 0; k < 6; ++k) { s->bbox[k] = s->replicas[0]->bbox[k]; s->sort_box[k] = s->replicas[0]->sort_box[k]; }
        *out = s;
        return 0;
    }
    TWG_CUDA(c, cudaSetDevice(c->device));
    double* dV = nullptr;
    uint32_t* dF = nullptr;
    TWG_CUDA(c, cudaMalloc(&dV, sizeof(double) * 3 * (size_t)nV));
    cudaError_t e = cudaMalloc(&dF, sizeof(uint32_t) * 3 * (size_t)nF);
    if (e != cudaSuccess) { cudaFree(dV); return twg_fail(c, (int)e, cudaGetErrorString(e), __FILE__, __LINE__); }
    int rc = 0;
    e = cudaMemcpyAsync(dV, V, sizeof(double) * 3 * (size_t)nV, cudaMemcpyHostToDevice, c->streams[0]);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dF, F, sizeof(uint32_t) * 3 * (size_t)nF, cudaMemcpyHostToDevice, c->streams[0]);
    if (e != cudaSuccess) rc = twg_fail(c, (int)e, cudaGetErrorString(e), __FILE__, __LINE__);
    if (rc == 0) rc = build(c, dV, nV, dF, nF, out);
    cudaFree(dV);
    cudaFree(dF);
    return rc;
}

}  // extern "C"
