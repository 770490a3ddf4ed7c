// common.cuh -- context, error handling and small device helpers of libtetwild_gpu (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <string>
#include <vector>
#include "../../include/tetwild_gpu.h"
#include "tw_math.cuh"

#define TWG_NUM_STREAMS 3

// Tuning knobs of one context. Defaults come from the environment ONCE, at twg_create (TWG_ENV_GROUP, TWG_SORT_BITS, ...:
// the variable is "TWG_" + the upper-cased option name); twg_set_option changes them per context afterwards.
struct twg_options {
    int env_group = 128;       // queries per cooperative group of the envelope point kernel
    int env_policy = 1;        // round scheduling (1) or eager refill (0)
    int env_front = 32;        // frontier cap of a group: 16 / 32 / 64
    int env_quorum = 16;       // parked lanes that trigger a leaf round
    int env_top = 64;          // pair records staged in shared memory
    int env_bound = 1;         // oriented facet bound before the exact leaf routine
    int envelope_sort = 1;     // Morton-order large batches before traversal
    int surface_order = 2;     // facet order of the envelope structure: 0 Z (Morton) curve, 1 Hilbert curve, 2 kd (splits aligned with the heap, longest centroid axis), 3 kd with the split axis chosen by child surface area
    int sort_bits = 24;        // key bits that are sorted
    int sort_curve = 0;        // query order of the envelope / winding batches: 0 Z (Morton) curve, 1 Hilbert curve
    int nearest_curve = 1;     // query order of the nearest-facet batches (Hilbert: 24.9 -> 22.7 ms per 10 M points; the envelope test prefers Z: 1.75 vs 1.79 ms)
    long long chunk_points = 1ll << 20;  // points per staging chunk of the host entry points
    int ring_waves = 3;        // resident CTAs per SM of the one-ring kernels
    int wide_gather = 1;       // resident-mesh quality pass: 24-byte vertices as one 16-byte + one 8-byte load instead of three 8-byte loads
    int ring_minb = 3;         // resident CTAs per SM the one-ring kernel's registers are capped for (3: 80 registers, 4: 64)
    int winding_minb = 3;
    int winding_sort = 1;
    int winding_leaf = 32;
    int winding_device_build = 1;
    int amips_tma = 1;
    int nearest_mode = 1;      // 1: packets of 32 queries (+ one query per warp past the budget), 2: round-scheduled lanes, 0: per-lane descents (round 1)
    int nearest_group = 64;    // queries per claimed group of the round-scheduled nearest kernel
    int nearest_budget = 1 << 30;  // node visits a packet of 32 queries may spend before its queries are finished one per warp (measured: never pays)
    int fast_calls = 1;        // tiny host calls go through the zero-copy slab (twg_fast_*)
    int trace = 0;
};

// One "lane" per CUDA stream that launches sorted / persistent kernels: its own Morton-sort scratch and its own work
// counters, so that asynchronous calls on DIFFERENT streams never share mutable device state. Lanes 0..TWG_NUM_STREAMS-1
// belong to the context's staging streams; a lane is created for every distinct caller stream passed to a _dev entry point
// (at most TWG_MAX_LANES; beyond that the least recently used one is recycled after its last launch has completed).
#define TWG_MAX_LANES 16
struct twg_lane {
    cudaStream_t stream = nullptr;
    void* dsort = nullptr;
    size_t dsort_bytes = 0;
    unsigned long long* counters = nullptr;  // device, 8 slots
    cudaEvent_t done = nullptr;              // recorded after the last launch that used this lane
    uint64_t tick = 0;
};

struct twg_worker;  // host thread that drives one device of a multi-device context (multi.cu)

struct twg_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t streams[TWG_NUM_STREAMS] = {nullptr, nullptr, nullptr};
    cudaEvent_t ev[TWG_NUM_STREAMS] = {nullptr, nullptr, nullptr};
    // pinned staging (host-buffer entry points): one in + one out slab per stream
    void* pin_in[TWG_NUM_STREAMS] = {nullptr, nullptr, nullptr};
    void* pin_out[TWG_NUM_STREAMS] = {nullptr, nullptr, nullptr};
    size_t pin_in_bytes = 0, pin_out_bytes = 0;
    // device scratch, grown on demand
    void* dscratch[TWG_NUM_STREAMS] = {nullptr, nullptr, nullptr};
    size_t dscratch_bytes[TWG_NUM_STREAMS] = {0, 0, 0};
    // zero-copy slab for tiny calls (twg_fast_*): mapped pinned host memory the kernels read and write directly, plus the
    // completion word the host spins on -- no cudaMemcpy, no cudaStreamSynchronize on the latency path
    void* fast_slab = nullptr;
    size_t fast_bytes = 0;
    uint32_t fast_seq = 0;
    uint32_t* fast_counter = nullptr;  // device: CTAs of the tiny kernel in flight that are through (twg_signal_done)
    std::vector<twg_lane> lanes;
    uint64_t lane_tick = 0;
    unsigned long long* dcounters = nullptr;  // device, TWG_NUM_DEBUG_COUNTERS slots (twg_debug_counter)
    twg_options opt;
    uint64_t launches = 0;
    // multi-device context (twg_create_multi): children[k] is an ordinary one-device context driven by workers[k]
    std::vector<twg_ctx*> children;
    std::vector<twg_worker*> workers;
    twg_ctx* parent = nullptr;
    mutable char err[512] = {0};
};
#define TWG_NUM_DEBUG_COUNTERS 8
#define TWG_DBG_ENV_STACK_OVERFLOW 0
#define TWG_DBG_BAD_INDEX 1
#define TWG_DBG_WINDING_PAIRS 2

inline int twg_fail(const twg_ctx* c, int code, const char* what, const char* file, int line) {
    if (c) snprintf(c->err, sizeof(c->err), "%s (%s:%d) code=%d", what, file, line, code);
    return code ? code : TWG_ERR_INTERNAL;
}

#define TWG_CUDA(ctx, call)                                                                   \
    do {                                                                                      \
        cudaError_t e__ = (call);                                                             \
        if (e__ != cudaSuccess) return twg_fail((ctx), (int)e__, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

#define TWG_CHECK(ctx, cond, code, msg)                                        \
    do {                                                                       \
        if (!(cond)) return twg_fail((ctx), (code), (msg), __FILE__, __LINE__); \
    } while (0)

#define TWG_TRY(expr)              \
    do {                           \
        int rc__ = (expr);         \
        if (rc__ != 0) return rc__; \
    } while (0)

// every kernel launch of the library goes through this (gpu_launches accounting for bench.py)
#define TWG_LAUNCH(ctx, kernel, grid, block, smem, stream, ...)                      \
    do {                                                                             \
        kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);                  \
        (ctx)->launches++;                                                           \
        cudaError_t e__ = cudaGetLastError();                                        \
        if (e__ != cudaSuccess) return twg_fail((ctx), (int)e__, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

int twg_ensure_scratch(twg_ctx* c, int slot, size_t bytes);
// Tiny calls (the un-batched call sites of the sequential scheduler: one face, one point, one one-ring): the arguments are
// written into a mapped pinned slab that the kernel reads over PCIe, the results are written by the kernel straight into the
// same slab, and completion is a word in that slab written by a one-warp kernel queued behind the work; the host spins on it.
// twg_fast_slab returns the slab (host pointer == device pointer under unified addressing), at least `bytes` long.
// facet order of a heap-aligned kd hierarchy (winding_build.cu); d_order: nF entries of device memory
int twg_kd_order_device(twg_ctx* c, cudaStream_t st, const double* dV, const uint32_t* dF, uint32_t nF, uint32_t nLeafP, uint32_t stop_leaves, int sah, uint32_t* d_order);
int twg_fast_slab(twg_ctx* c, size_t bytes, char** slab);
int twg_fast_wait(twg_ctx* c, cudaStream_t st);
// The same without the second launch: a kernel that takes a twg_done raises the completion word itself when its last CTA is
// through (twg_signal_done as the kernel's last statement, reached by every thread). twg_fast_arm hands out the word, the
// CTA counter and the sequence number for ONE launch on `st`; twg_fast_spin waits for that number.
struct twg_done {
    volatile uint32_t* flag;  // NULL: nothing to signal
    uint32_t* counter;        // device memory, zero between launches
    uint32_t seq;
};
int twg_fast_arm(twg_ctx* c, twg_done* done);
int twg_fast_spin(twg_ctx* c, cudaStream_t st, uint32_t seq);
#ifdef __CUDACC__
__device__ __forceinline__ void twg_signal_done(const twg_done& d) {
    if (!d.flag) return;
    __threadfence_system();  // this thread's results are visible to the host before the word is
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t total = gridDim.x * gridDim.y * gridDim.z;
        if (total == 1u || atomicAdd(d.counter, 1u) == total - 1u) {
            if (total != 1u) *d.counter = 0u;
            __threadfence_system();
            *d.flag = d.seq;
        }
    }
}
#endif
int twg_ensure_pinned(twg_ctx* c, size_t in_bytes, size_t out_bytes);
#define TWG_SORT_MIN 4096  /* batches below this are traversed in the caller's order */
// the lane of stream `st` (created on first use); *out stays valid until the next twg_get_lane call on this context
int twg_get_lane(twg_ctx* c, cudaStream_t st, twg_lane** out);
// marks the lane busy until everything queued on its stream so far has completed (call after the last launch of an entry point)
int twg_lane_mark(twg_ctx* c, twg_lane* lane);
int twg_sort_points(twg_ctx* c, twg_lane* lane, cudaStream_t st, const double* dP, uint64_t n, const uint32_t** perm_out, const double* known_box = nullptr,
                    const double** sorted_out = nullptr, const uint32_t** keys_out_dbg = nullptr, int curve = -1 /* -1: option sort_curve */);
inline bool twg_is_multi(const twg_ctx* c) { return c && !c->children.empty(); }
// multi.cu: fn(k, child_k) on every device's own host thread, concurrently; first non-zero return code wins
int twg_multi_run(twg_ctx* c, const std::function<int(int, twg_ctx*)>& fn);
int twg_forward0(twg_ctx* c, int rc);  // rc of a call forwarded to device 0 (copies its error message)
// batches below these sizes stay on device 0 of a multi-device context
#define TWG_MULTI_MIN_POINTS 65536
#define TWG_MULTI_MIN_FACES 4096
#define TWG_MULTI_MIN_TETS 262144

// ---- device helpers ----
#if defined(__CUDACC__)
__device__ __forceinline__ double ldg_d(const double* p) { return __ldg(p); }
// 128-bit streaming load / store (read-once inputs, write-once outputs: keep them out of L1)
__device__ __forceinline__ double2 ld_stream2(const double* p) {
    double2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream2(double* p, double2 v) {
    asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
}

// ---- per-thread asynchronous global -> shared copies (LDGSTS): 8 bytes each, both addresses 8-byte aligned ----
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ---- mbarrier + 1-D bulk async copy (TMA) ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared bulk copy; bytes % 16 == 0, both addresses 16-byte aligned; completes on `bar`
__device__ __forceinline__ void tma_bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// shared -> global bulk copy (TMA store), bulk-group completion. bytes % 16 == 0, both addresses 16-byte aligned.
__device__ __forceinline__ void tma_bulk_s2g(void* gmem_dst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all committed bulk stores have finished READING their shared-memory source (it may be overwritten)
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// generic-proxy writes to shared memory become visible to the async proxy (TMA) after this fence + a barrier
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
#endif
