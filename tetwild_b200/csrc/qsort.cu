// qsort.cu -- Morton ordering of a query batch on the device (shared by the envelope, nearest-facet and winding kernels).
//
// The traversals run at full SIMT width only when the 32 queries of a warp (the 64 of a group) walk one neighbourhood of the
// hierarchy, so every large batch is ordered along a Morton curve over the surface's bounding box first. The reference gets
// the same effect for free: its samples come out of sampleTriangle in spatial order and it passes the previous facet as a
// hint (LocalOperations.cpp:1078-1086).
//
// Own radix sort (no library on the query path): a STABLE least-significant-digit sort of (key, original index) pairs on the
// top `sort_bits` of a 30-bit Morton key, 8 bits per pass, written for this one job:
//   keys_hist   one read of the points: Morton key of every point + the per-tile histogram of the first digit
//   tile_hist   per-tile histogram of the next digit (passes 2..)
//   row_scan    exclusive scan of every digit's counts over the tiles (one CTA per digit) + the digit totals
//   scatter     per tile of 4096 pairs: warp-synchronous stable ranking (__match_any_sync groups lanes with equal digits; the
//               group's first lane advances the warp's counter of that digit), the tile is staged in shared memory in digit
//               order and written out in runs, so the global stores of one digit are contiguous. The pairs of the first pass
//               carry an implicit index (no iota array); the LAST pass gathers the points themselves -- the kernels then
//               read their queries as one coalesced stream and only scatter 1 B / 8 B results back through `perm`.
// Deterministic: equal keys keep the caller's order, so a batch is traversed in the same order on every run.
#include "common.cuh"

namespace {

constexpr int kSortThreads = 256;
constexpr int kSortItems = 16;
constexpr int kSortTile = kSortThreads * kSortItems;  // 4096 pairs per CTA
constexpr int kSortWarps = kSortThreads / 32;

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
// 30-bit Hilbert index of a point on a 1024^3 grid (Skilling's transpose form, as in surface.cu): consecutive indices are
// adjacent cells, so a run of sorted queries is a connected blob without the jumps of the Z order (option sort_curve)
__device__ __forceinline__ uint32_t hilbert30(uint32_t x, uint32_t y, uint32_t z) {
    uint32_t X[3] = {x, y, z};
    const uint32_t M = 1u << 9;
    for (uint32_t Q = M; Q > 1; Q >>= 1) {
        const uint32_t P = Q - 1;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            if (X[i] & Q) X[0] ^= P;
            else { const uint32_t t = (X[0] ^ X[i]) & P; X[0] ^= t; X[i] ^= t; }
        }
    }
    X[1] ^= X[0];
    X[2] ^= X[1];
    uint32_t t = 0;
    for (uint32_t Q = M; Q > 1; Q >>= 1)
        if (X[2] & Q) t ^= Q - 1;
    return (spread10(X[0] ^ t) << 2) | (spread10(X[1] ^ t) << 1) | spread10(X[2] ^ t);
}
struct Box6 {
    double v[6];  // lo xyz, hi xyz
};
struct Digit {
    int shift;
    uint32_t mask;
    __device__ __forceinline__ uint32_t of(uint32_t key) const { return (key >> shift) & mask; }
};

// element e of the tile owned by (warp w, round r, lane l): e = w * 512 + r * 32 + l -- increasing with (w, r, l), which is the
// order the stable ranking below counts in
__device__ __forceinline__ uint32_t tile_elem(int w, int r, int l) { return (uint32_t)(w * (kSortItems * 32) + r * 32 + l); }

// Morton keys of the points + histogram of the first digit of every tile: hist[d * ntiles + tile]
__global__ void __launch_bounds__(kSortThreads) qs_keys_hist_kernel(const double* __restrict__ Q, uint64_t n, const unsigned long long* __restrict__ bounds, Box6 known,
                                                                uint32_t* __restrict__ keys, uint32_t* __restrict__ hist, uint32_t ntiles, Digit dg, int hilbert) {
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    double lo[3], inv[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const double l = bounds ? dec(bounds[c]) : known.v[c], hgh = bounds ? dec(bounds[3 + c]) : known.v[3 + c];
        const double ext = hgh - l;
        lo[c] = l;
        inv[c] = ext > 0.0 ? 1.0 / ext : 0.0;
    }
    const uint64_t base = (uint64_t)blockIdx.x * kSortTile;
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
#pragma unroll 4
    for (int r = 0; r < kSortItems; ++r) {
        const uint64_t i = base + tile_elem(w, r, l);
        if (i < n) {
            uint32_t code = 0, qd[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const double x = __ldg(Q + 3 * i + c);
                double u = isfinite(x) ? (x - lo[c]) * inv[c] : 0.0;
                u = fmin(fmax(u, 0.0), 1.0);
                qd[c] = (uint32_t)(u * 1023.0);
                code |= spread10(qd[c]) << c;
            }
            if (hilbert) code = hilbert30(qd[0], qd[1], qd[2]);
            keys[i] = code;
            atomicAdd(&h[dg.of(code)], 1u);
        }
    }
    __syncthreads();
    hist[(size_t)threadIdx.x * ntiles + blockIdx.x] = h[threadIdx.x];
}

__global__ void __launch_bounds__(kSortThreads) qs_tile_hist_kernel(const uint32_t* __restrict__ keys, uint64_t n, uint32_t* __restrict__ hist, uint32_t ntiles, Digit dg) {
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const uint64_t base = (uint64_t)blockIdx.x * kSortTile;
#pragma unroll 4
    for (int r = 0; r < kSortItems; ++r) {
        const uint64_t i = base + (uint64_t)r * kSortThreads + threadIdx.x;
        if (i < n) atomicAdd(&h[dg.of(__ldg(keys + i))], 1u);
    }
    __syncthreads();
    hist[(size_t)threadIdx.x * ntiles + blockIdx.x] = h[threadIdx.x];
}

// CTA d: exclusive scan of row d of hist over the tiles (in place), total[d] = the row's sum
__global__ void __launch_bounds__(256) qs_row_scan_kernel(uint32_t* __restrict__ hist, uint32_t ntiles, uint32_t* __restrict__ total) {
    __shared__ uint32_t wsum[8];
    __shared__ uint32_t carry_s;
    uint32_t* row = hist + (size_t)blockIdx.x * ntiles;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (uint32_t b = 0; b < ntiles; b += 256) {
        const uint32_t i = b + threadIdx.x;
        const uint32_t v = i < ntiles ? row[i] : 0u;
        uint32_t incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) wsum[w] = incl;
        __syncthreads();
        uint32_t woff = 0;
        for (int k = 0; k < w; ++k) woff += wsum[k];
        const uint32_t carry = carry_s;
        if (i < ntiles) row[i] = carry + woff + incl - v;
        __syncthreads();
        if (threadIdx.x == 255) carry_s = carry + woff + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) total[blockIdx.x] = carry_s;
}

// One pass over one tile: stable ranks by digit, staging in digit order, contiguous runs out.
//   FIRST: the pairs' indices are implicit (position in the batch); LAST with Pin: write the POINTS in sorted order + perm
template <bool FIRST>
__global__ void __launch_bounds__(kSortThreads, 4) qs_scatter_kernel(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in, uint64_t n,
                                                              const uint32_t* __restrict__ hist /*scanned rows*/, const uint32_t* __restrict__ total, uint32_t ntiles,
                                                              Digit dg, uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out,
                                                              const double* __restrict__ Pin, double* __restrict__ Pout) {
    __shared__ uint32_t cnt[kSortWarps][256];   // per-warp digit counters -> per-warp bases
    __shared__ uint32_t dstart[256];            // first staged slot of every digit in this tile
    __shared__ uint32_t gbase[256];             // global position of the tile's first element of every digit
    __shared__ uint32_t skey[kSortTile], sval[kSortTile];
    __shared__ uint32_t wtot[kSortWarps], wtot2[kSortWarps];
    const unsigned full = 0xffffffffu;
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    const unsigned lt = (1u << l) - 1u;
    const uint64_t base = (uint64_t)blockIdx.x * kSortTile;
    const uint32_t tile_n = (uint32_t)((n - base < (uint64_t)kSortTile) ? (n - base) : (uint64_t)kSortTile);
#pragma unroll
    for (int k = 0; k < kSortWarps; ++k) cnt[k][threadIdx.x] = 0;
    __syncthreads();
    uint32_t key[kSortItems];
    uint16_t rank[kSortItems];  // < 512: position among the warp's elements with the same digit
#pragma unroll
    for (int r = 0; r < kSortItems; ++r) {
        const uint32_t e = tile_elem(w, r, l);
        key[r] = e < tile_n ? __ldg(keys_in + base + e) : 0xffffffffu;
    }
    // ---- stable rank inside the warp's 512 elements: round after round, lanes with equal digits form a group; the first lane of the
    // group reads and advances the warp's counter of that digit, the others add their position in the group
#pragma unroll
    for (int r = 0; r < kSortItems; ++r) {
        const bool valid = tile_elem(w, r, l) < tile_n;
        const uint32_t d = dg.of(key[r]);
        const unsigned m = __match_any_sync(full, valid ? d : (256u + (uint32_t)l));
        const int leader = __ffs(m) - 1;
        uint32_t b = 0;
        if (valid && l == leader) {
            b = cnt[w][d];
            cnt[w][d] = b + (uint32_t)__popc(m);
        }
        b = __shfl_sync(full, b, leader);
        rank[r] = (uint16_t)(b + (uint32_t)__popc(m & lt));
        __syncwarp();
    }
    __syncthreads();
    // ---- thread d: bases of digit d over the warps, the tile's count, and (block scans) the digit's first staged slot and the
    // number of elements with a smaller digit in the whole batch
    {
        const int d = threadIdx.x;
        uint32_t run = 0;
#pragma unroll
        for (int k = 0; k < kSortWarps; ++k) {
            const uint32_t c = cnt[k][d];
            cnt[k][d] = run;
            run += c;
        }
        const uint32_t tv = __ldg(total + d);
        uint32_t incl = run, ti = tv;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(full, incl, o);
            const uint32_t t2 = __shfl_up_sync(full, ti, o);
            if (l >= o) { incl += t; ti += t2; }
        }
        if (l == 31) { wtot[w] = incl; wtot2[w] = ti; }
        __syncthreads();
        uint32_t woff = 0, wo2 = 0;
        for (int k = 0; k < w; ++k) { woff += wtot[k]; wo2 += wtot2[k]; }
        dstart[d] = woff + incl - run;
        // global: digits below d (all tiles) + digit d in the tiles before this one
        gbase[d] = (wo2 + ti - tv) + __ldg(hist + (size_t)d * ntiles + blockIdx.x);
    }
    __syncthreads();
    // ---- stage the tile in digit order (the indices are only read now: they are not live across the ranking rounds)
#pragma unroll
    for (int r = 0; r < kSortItems; ++r) {
        const uint32_t e = tile_elem(w, r, l);
        if (e < tile_n) {
            const uint32_t d = dg.of(key[r]);
            const uint32_t slot = dstart[d] + cnt[w][d] + (uint32_t)rank[r];
            skey[slot] = key[r];
            sval[slot] = FIRST ? (uint32_t)(base + e) : __ldg(vals_in + base + e);
        }
    }
    __syncthreads();
    // ---- write runs: staged slot i of digit d goes to gbase[d] + (i - dstart[d])
    for (uint32_t i = threadIdx.x; i < tile_n; i += kSortThreads) {
        const uint32_t k = skey[i], v = sval[i];
        const uint32_t d = dg.of(k);
        const uint32_t dst = gbase[d] + (i - dstart[d]);
        if (Pout) {
            const double x = __ldg(Pin + 3 * (size_t)v), y = __ldg(Pin + 3 * (size_t)v + 1), z = __ldg(Pin + 3 * (size_t)v + 2);
            Pout[3 * (size_t)dst] = x; Pout[3 * (size_t)dst + 1] = y; Pout[3 * (size_t)dst + 2] = z;
        } else if (keys_out) {
            keys_out[dst] = k;
        }
        vals_out[dst] = v;
    }
}

}  // namespace

// perm_out[i] = index (in the caller's order) of the i-th point along the Morton curve. `lane` is the caller stream's lane
// (twg_get_lane): its sort scratch is only ever touched by work queued on that one stream.
// known_box (optional, host: lo xyz, hi xyz): quantise over this box (points outside are clamped to its faces)
// instead of reducing the batch's own bounding box first -- the surface's box is what matters to both traversals.
// sorted_out (optional): the points themselves in sorted order (gathered by the last pass).
int twg_sort_points(twg_ctx* c, twg_lane* lane, cudaStream_t st, const double* dP, uint64_t n, const uint32_t** perm_out, const double* known_box,
                    const double** sorted_out, const uint32_t** keys_out_dbg, int curve) {
    if (curve < 0) curve = c->opt.sort_curve;
    TWG_CHECK(c, n <= 0x7fffffffull, TWG_ERR_INVALID_ARG, "at most 2^31-1 queries per device call (the host entry points chunk larger batches)");
    auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
    // The keys are 30-bit Morton codes; only the top `bits` are sorted (stable): 24 bits = a 256^3 grid over the surface's
    // box = three 8-bit passes. Order inside a cell does not matter to the traversals (a warp's group of 64 queries spans a
    // cell or two either way).
    const int bits = c->opt.sort_bits;
    const int passes = (bits + 7) / 8;
    const uint32_t ntiles = (uint32_t)((n + kSortTile - 1) / kSortTile);
    const size_t kb = up(n * 4);
    const size_t hb = up((size_t)256 * ntiles * 4);
    const size_t pb = sorted_out ? up(n * 24) : 0;
    const size_t need = 1280 + 4 * kb + hb + pb;
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
    uint32_t* total = (uint32_t*)(base + 256);
    uint32_t* kbuf[2] = {(uint32_t*)(base + 1280), (uint32_t*)(base + 1280 + kb)};
    uint32_t* vbuf[2] = {(uint32_t*)(base + 1280 + 2 * kb), (uint32_t*)(base + 1280 + 3 * kb)};
    uint32_t* hist = (uint32_t*)(base + 1280 + 4 * kb);
    double* Pd = sorted_out ? (double*)(base + 1280 + 4 * kb + hb) : nullptr;
    Box6 known;
    for (int k = 0; k < 6; ++k) known.v[k] = known_box ? known_box[k] : 0.0;
    if (!known_box) {
        TWG_LAUNCH(c, qinit_kernel, 1, 32, 0, st, bounds);
        uint64_t g = (3 * n + 255) / 256;
        if (g > (uint64_t)c->sm_count * 16) g = (uint64_t)c->sm_count * 16;
        TWG_LAUNCH(c, qbounds_kernel, (unsigned)g, 256, 0, st, dP, n, bounds);
    }
    int cur = 0;
    for (int p = 0; p < passes; ++p) {
        Digit dg;
        dg.shift = (30 - bits) + 8 * p;
        const int width = (bits - 8 * p) < 8 ? (bits - 8 * p) : 8;
        dg.mask = (1u << width) - 1u;
        if (p == 0)
            TWG_LAUNCH(c, qs_keys_hist_kernel, ntiles, kSortThreads, 0, st, dP, n, known_box ? (const unsigned long long*)nullptr : bounds, known, kbuf[0], hist, ntiles, dg, curve);
        else
            TWG_LAUNCH(c, qs_tile_hist_kernel, ntiles, kSortThreads, 0, st, (const uint32_t*)kbuf[cur], n, hist, ntiles, dg);
        TWG_LAUNCH(c, qs_row_scan_kernel, 256, 256, 0, st, hist, ntiles, total);
        const bool last = p + 1 == passes;
        const double* Pin = (last && Pd) ? dP : nullptr;
        double* Pout = (last && Pd) ? Pd : nullptr;
        if (p == 0)
            TWG_LAUNCH(c, (qs_scatter_kernel<true>), ntiles, kSortThreads, 0, st, (const uint32_t*)kbuf[cur], (const uint32_t*)nullptr, n, (const uint32_t*)hist,
                       (const uint32_t*)total, ntiles, dg, kbuf[cur ^ 1], vbuf[cur ^ 1], Pin, Pout);
        else
            TWG_LAUNCH(c, (qs_scatter_kernel<false>), ntiles, kSortThreads, 0, st, (const uint32_t*)kbuf[cur], (const uint32_t*)vbuf[cur], n, (const uint32_t*)hist,
                       (const uint32_t*)total, ntiles, dg, kbuf[cur ^ 1], vbuf[cur ^ 1], Pin, Pout);
        cur ^= 1;
    }
    *perm_out = vbuf[cur];
    if (sorted_out) *sorted_out = Pd;
    if (keys_out_dbg) *keys_out_dbg = (Pd ? nullptr : kbuf[cur]);  // the last pass does not write keys when it gathers the points
    return 0;
}

// test hook (tests/test_gpu_robustness.py): sorts n host points over `box`, returns the permutation and the sorted keys
extern "C" int twg_debug_sort_points(twg_ctx* c, const double* P, uint64_t n, const double* box6, uint32_t* perm, uint32_t* keys, double* sorted_xyz) {
    TWG_CHECK(c, c && P && perm && n > 0, TWG_ERR_INVALID_ARG, "null argument");
    if (twg_is_multi(c)) return twg_forward0(c, twg_debug_sort_points(c->children[0], P, n, box6, perm, keys, sorted_xyz));
    TWG_CUDA(c, cudaSetDevice(c->device));
    cudaStream_t st = c->streams[0];
    twg_lane* lane = nullptr;
    TWG_TRY(twg_get_lane(c, st, &lane));
    TWG_TRY(twg_ensure_scratch(c, 0, n * 24));
    TWG_CUDA(c, cudaMemcpyAsync(c->dscratch[0], P, n * 24, cudaMemcpyHostToDevice, st));
    const uint32_t *dperm = nullptr, *dkeys = nullptr;
    const double* dsorted = nullptr;
    if (keys) {  // keys-only form (what the winding kernel uses)
        TWG_TRY(twg_sort_points(c, lane, st, (const double*)c->dscratch[0], n, &dperm, box6, nullptr, &dkeys));
        TWG_CUDA(c, cudaMemcpyAsync(keys, dkeys, n * 4, cudaMemcpyDeviceToHost, st));
        TWG_CUDA(c, cudaMemcpyAsync(perm, dperm, n * 4, cudaMemcpyDeviceToHost, st));
        TWG_CUDA(c, cudaStreamSynchronize(st));
    }
    if (sorted_xyz) {  // gathering form (envelope / nearest kernels)
        TWG_TRY(twg_sort_points(c, lane, st, (const double*)c->dscratch[0], n, &dperm, box6, &dsorted, nullptr));
        TWG_CUDA(c, cudaMemcpyAsync(sorted_xyz, dsorted, n * 24, cudaMemcpyDeviceToHost, st));
        TWG_CUDA(c, cudaMemcpyAsync(perm, dperm, n * 4, cudaMemcpyDeviceToHost, st));
        TWG_CUDA(c, cudaStreamSynchronize(st));
    }
    return twg_lane_mark(c, lane);
}
