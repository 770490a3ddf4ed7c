// qsort.cu -- Morton ordering of a query batch on the device (shared by the envelope and winding kernels).
//
// Both traversals are one-query-per-lane; a warp only runs at full SIMT width when its 32 queries take the same path
// through the hierarchy and touch the same cache lines, so every large batch is first ordered along a 30-bit Morton
// curve over its own bounding box: bbox reduction -> key kernel -> cub::DeviceRadixSort (keys + original indices).
// The kernels then read P[perm[i]] and scatter their 1-byte / 8-byte results to the caller's positions.
// The reference gets the same effect for free: its samples come out of sampleTriangle in spatial order and it passes
// the previous facet as a hint (LocalOperations.cpp:1078-1086).
#include <cub/device/device_radix_sort.cuh>
#include "common.cuh"

namespace {

__device__ __forceinline__ unsigned long long enc(double d) {  // order-preserving double -> u64
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
    // flat coalesced sweep over the 3n doubles; component = index % 3
    const uint64_t tot = 3 * n;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < tot; i += (uint64_t)gridDim.x * blockDim.x) {
        const double x = __ldg(Q + i);
        const int c = (int)(i % 3);
        if (isfinite(x)) { lo[c] = fmin(lo[c], x); hi[c] = fmax(hi[c], x); }
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
struct Box6 {
    double v[6];  // lo xyz, hi xyz
};
__global__ void __launch_bounds__(256) qkeys_kernel(const double* __restrict__ Q, uint64_t n, const unsigned long long* __restrict__ bounds, Box6 known,
                                                    uint32_t* keys, uint32_t* vals) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t code = 0;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const double lo = bounds ? dec(bounds[c]) : known.v[c], hi = bounds ? dec(bounds[3 + c]) : known.v[3 + c];
        const double ext = hi - lo;
        const double x = __ldg(Q + 3 * i + c);
        double u = (ext > 0.0 && isfinite(x)) ? (x - lo) / ext : 0.0;
        u = fmin(fmax(u, 0.0), 1.0);
        code |= spread10((uint32_t)(u * 1023.0)) << c;
    }
    keys[i] = code;
    vals[i] = (uint32_t)i;
}

// Pd[i] = P[perm[i]]: the traversal kernels then read their queries as a coalesced stream (one memory round trip per
// refill instead of the dependent perm -> point pair), only the 1-byte / 8-byte results are scattered back.
__global__ void __launch_bounds__(256) qgather_kernel(const double* __restrict__ P, const uint32_t* __restrict__ perm, uint64_t n, double* __restrict__ Pd) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t s = (uint64_t)__ldg(perm + i);
    const double x = __ldg(P + 3 * s), y = __ldg(P + 3 * s + 1), z = __ldg(P + 3 * s + 2);
    Pd[3 * i] = x; Pd[3 * i + 1] = y; Pd[3 * i + 2] = z;
}

}  // namespace

// perm_out[i] = index (in the caller's order) of the i-th point along the Morton curve. `lane` is the caller stream's lane
// (twg_get_lane): its sort scratch is only ever touched by work queued on that one stream.
// known_box (optional, host: lo xyz, hi xyz): quantise over this box (points outside are clamped to its faces)
// instead of reducing the batch's own bounding box first -- the surface's box is what matters to both traversals.
int twg_sort_points(twg_ctx* c, twg_lane* lane, cudaStream_t st, const double* dP, uint64_t n, const uint32_t** perm_out, const double* known_box,
                    const double** sorted_out) {
    TWG_CHECK(c, n <= 0x7fffffffull, TWG_ERR_INVALID_ARG, "at most 2^31-1 queries per device call (the host entry points chunk larger batches)");
    auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
    // The keys are 30-bit Morton codes; only the top `bits` are sorted (stable): 24 bits = a 256^3 grid over the surface's
    // box = three 8-bit radix passes instead of four. Order inside a cell does not matter to the traversals (a warp's group
    // of 64 queries spans a cell or two either way).
    const int bits = c->opt.sort_bits;
    const int begin_bit = 30 - bits;
    size_t tmp_bytes = 0;
    TWG_CUDA(c, cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, (uint32_t*)nullptr, (uint32_t*)nullptr, (uint32_t*)nullptr, (uint32_t*)nullptr,
                                                (int)n, begin_bit, 30, st));
    const size_t kb = up(n * 4);
    const size_t pb = sorted_out ? up(n * 24) : 0;
    const size_t need = 256 + 4 * kb + up(tmp_bytes) + pb;
    if (lane->dsort_bytes < need) {
        if (lane->dsort) {
            TWG_CUDA(c, cudaStreamSynchronize(st));
            TWG_CUDA(c, cudaFree(lane->dsort));
            lane->dsort = nullptr;
            lane->dsort_bytes = 0;
        }
        const size_t want = need + need / 8;
        TWG_CUDA(c, cudaMalloc(&lane->dsort, want));
        lane->dsort_bytes = want;
    }
    char* base = (char*)lane->dsort;
    unsigned long long* bounds = (unsigned long long*)base;
    uint32_t *keys = (uint32_t*)(base + 256), *keys2 = (uint32_t*)(base + 256 + kb), *vals = (uint32_t*)(base + 256 + 2 * kb),
             *vals2 = (uint32_t*)(base + 256 + 3 * kb);
    void* tmp = base + 256 + 4 * kb;
    Box6 known;
    for (int k = 0; k < 6; ++k) known.v[k] = known_box ? known_box[k] : 0.0;
    if (!known_box) {
        TWG_LAUNCH(c, qinit_kernel, 1, 32, 0, st, bounds);
        uint64_t g = (3 * n + 255) / 256;
        if (g > (uint64_t)c->sm_count * 16) g = (uint64_t)c->sm_count * 16;
        TWG_LAUNCH(c, qbounds_kernel, (unsigned)g, 256, 0, st, dP, n, bounds);
    }
    TWG_LAUNCH(c, qkeys_kernel, (unsigned)((n + 255) / 256), 256, 0, st, dP, n, known_box ? (const unsigned long long*)nullptr : bounds, known, keys, vals);
    TWG_CUDA(c, cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys, keys2, vals, vals2, (int)n, begin_bit, 30, st));
    c->launches += 2 + (bits + 7) / 8;  // cub: histogram + exclusive-sum + one onesweep pass per 8 key bits (library kernels)
    *perm_out = vals2;
    if (sorted_out) {
        double* Pd = (double*)(base + 256 + 4 * kb + up(tmp_bytes));
        TWG_LAUNCH(c, qgather_kernel, (unsigned)((n + 255) / 256), 256, 0, st, dP, vals2, n, Pd);
        *sorted_out = Pd;
    }
    return 0;
}
