// ctx.cu -- context lifetime, options, per-stream lanes, scratch / pinned staging, error reporting (C ABI: include/tetwild_gpu.h).
#include <cctype>
#include <time.h>
#include "common.cuh"

void twg_multi_teardown(twg_ctx* c);  // multi.cu

namespace {

struct OptDesc {
    const char* name;
    int twg_options::*i;
    long long twg_options::*ll;
    long long lo, hi;
};
const OptDesc kOpts[] = {
    {"env_group", &twg_options::env_group, nullptr, 32, 4096},
    {"env_policy", &twg_options::env_policy, nullptr, 0, 1},
    {"env_front", &twg_options::env_front, nullptr, 1, 64},
    {"env_quorum", &twg_options::env_quorum, nullptr, 1, 32},
    {"env_top", &twg_options::env_top, nullptr, 4, 512},
    {"env_bound", &twg_options::env_bound, nullptr, 0, 1},
    {"envelope_sort", &twg_options::envelope_sort, nullptr, 0, 1},
    {"surface_order", &twg_options::surface_order, nullptr, 0, 3},
    {"sort_bits", &twg_options::sort_bits, nullptr, 8, 30},
    {"sort_curve", &twg_options::sort_curve, nullptr, 0, 1},
    {"nearest_curve", &twg_options::nearest_curve, nullptr, 0, 1},
    {"chunk_points", nullptr, &twg_options::chunk_points, 1024, 1ll << 31},
    {"ring_waves", &twg_options::ring_waves, nullptr, 1, 32},
    {"ring_minb", &twg_options::ring_minb, nullptr, 3, 4},
    {"wide_gather", &twg_options::wide_gather, nullptr, 0, 1},
    {"winding_minb", &twg_options::winding_minb, nullptr, 1, 8},
    {"winding_sort", &twg_options::winding_sort, nullptr, 0, 1},
    {"winding_leaf", &twg_options::winding_leaf, nullptr, 2, 4096},
    {"winding_device_build", &twg_options::winding_device_build, nullptr, 0, 1},
    {"amips_tma", &twg_options::amips_tma, nullptr, 0, 1},
    {"nearest_mode", &twg_options::nearest_mode, nullptr, 0, 2},
    {"nearest_group", &twg_options::nearest_group, nullptr, 32, 4096},
    {"nearest_budget", &twg_options::nearest_budget, nullptr, 1, 1 << 30},
    {"fast_calls", &twg_options::fast_calls, nullptr, 0, 1},
    {"trace", &twg_options::trace, nullptr, 0, 1},
};

const OptDesc* find_opt(const char* name) {
    for (const OptDesc& d : kOpts)
        if (strcmp(d.name, name) == 0) return &d;
    return nullptr;
}

void set_opt(twg_options& o, const OptDesc& d, long long v) {
    if (v < d.lo) v = d.lo;
    if (v > d.hi) v = d.hi;
    if (d.i) o.*(d.i) = (int)v;
    else o.*(d.ll) = v;
}

void options_from_env(twg_options& o) {
    for (const OptDesc& d : kOpts) {
        char var[64] = "TWG_";
        size_t k = 4;
        for (const char* p = d.name; *p && k + 1 < sizeof(var); ++p) var[k++] = (char)toupper((unsigned char)*p);
        var[k] = 0;
        const char* e = getenv(var);
        if (e && *e) set_opt(o, d, strtoll(e, nullptr, 10));
    }
}

void free_lane(twg_lane& l) {
    if (l.dsort) cudaFree(l.dsort);
    if (l.counters) cudaFree(l.counters);
    if (l.done) cudaEventDestroy(l.done);
    l = twg_lane();
}

int init_lane(twg_ctx* c, twg_lane& l, cudaStream_t st) {
    l.stream = st;
    TWG_CUDA(c, cudaMalloc(&l.counters, 8 * sizeof(unsigned long long)));
    TWG_CUDA(c, cudaMemset(l.counters, 0, 8 * sizeof(unsigned long long)));
    TWG_CUDA(c, cudaEventCreateWithFlags(&l.done, cudaEventDisableTiming));
    return 0;
}

}  // namespace

extern "C" {

const char* twg_version(void) { return "tetwild_b200 0.2 (sm_100a)"; }

int twg_create(twg_ctx** out, int device_id) {
    if (!out) return TWG_ERR_INVALID_ARG;
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) return TWG_ERR_NO_DEVICE;  // no CPU fallback, by design
    if (device_id < 0 || device_id >= ndev) return TWG_ERR_INVALID_ARG;
    twg_ctx* c = new twg_ctx;
    c->device = device_id;
    options_from_env(c->opt);
    if (cudaSetDevice(device_id) != cudaSuccess) { delete c; return TWG_ERR_NO_DEVICE; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device_id) != cudaSuccess) { delete c; return TWG_ERR_NO_DEVICE; }
    if (prop.major < 10) {  // the fatbin holds sm_100a SASS only
        delete c;
        return TWG_ERR_NO_DEVICE;
    }
    c->sm_count = prop.multiProcessorCount;
    {   // the stream-ordered pool keeps what the device-side builds free (winding_build.cu), so a rebuild does not go back to the driver
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device_id) == cudaSuccess) {
            unsigned long long keep = 1ull << 32;  // up to 4 GiB cached
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
    }
    c->lanes.reserve(TWG_MAX_LANES);
    for (int i = 0; i < TWG_NUM_STREAMS; ++i) {
        if (cudaStreamCreateWithFlags(&c->streams[i], cudaStreamNonBlocking) != cudaSuccess) { twg_destroy(c); return TWG_ERR_INTERNAL; }
        if (cudaEventCreateWithFlags(&c->ev[i], cudaEventDisableTiming) != cudaSuccess) { twg_destroy(c); return TWG_ERR_INTERNAL; }
        c->lanes.emplace_back();
        if (init_lane(c, c->lanes.back(), c->streams[i]) != 0) { twg_destroy(c); return TWG_ERR_INTERNAL; }
    }
    if (cudaMalloc(&c->dcounters, TWG_NUM_DEBUG_COUNTERS * sizeof(unsigned long long)) != cudaSuccess ||
        cudaMemset(c->dcounters, 0, TWG_NUM_DEBUG_COUNTERS * sizeof(unsigned long long)) != cudaSuccess) {
        twg_destroy(c);
        return TWG_ERR_INTERNAL;
    }
    *out = c;
    return 0;
}

void twg_destroy(twg_ctx* c) {
    if (!c) return;
    if (!c->children.empty()) {
        twg_multi_teardown(c);
        delete c;
        return;
    }
    cudaSetDevice(c->device);
    for (int i = 0; i < TWG_NUM_STREAMS; ++i)
        if (c->streams[i]) cudaStreamSynchronize(c->streams[i]);
    for (twg_lane& l : c->lanes) {
        if (l.done) cudaEventSynchronize(l.done);
        free_lane(l);
    }
    for (int i = 0; i < TWG_NUM_STREAMS; ++i) {
        if (c->streams[i]) cudaStreamDestroy(c->streams[i]);
        if (c->ev[i]) cudaEventDestroy(c->ev[i]);
        if (c->pin_in[i]) cudaFreeHost(c->pin_in[i]);
        if (c->pin_out[i]) cudaFreeHost(c->pin_out[i]);
        if (c->dscratch[i]) cudaFree(c->dscratch[i]);
    }
    if (c->dcounters) cudaFree(c->dcounters);
    if (c->fast_slab) cudaFreeHost(c->fast_slab);
    if (c->fast_counter) cudaFree(c->fast_counter);
    delete c;
}

const char* twg_last_error(const twg_ctx* c) { return c ? c->err : "null context"; }
int twg_device(const twg_ctx* c) { return c ? c->device : -1; }
int twg_num_devices(const twg_ctx* c) { return c ? (c->children.empty() ? 1 : (int)c->children.size()) : 0; }
twg_ctx* twg_device_context(twg_ctx* c, int k) {
    if (!c) return nullptr;
    if (c->children.empty()) return k == 0 ? c : nullptr;
    return (k >= 0 && k < (int)c->children.size()) ? c->children[k] : nullptr;
}

uint64_t twg_launch_count(const twg_ctx* c) {
    if (!c) return 0;
    uint64_t n = c->launches;
    for (const twg_ctx* k : c->children) n += k->launches;
    return n;
}

int twg_synchronize(twg_ctx* c) {
    if (!c) return TWG_ERR_INVALID_ARG;
    if (!c->children.empty()) {
        for (twg_ctx* k : c->children) TWG_TRY(twg_synchronize(k));
        return 0;
    }
    TWG_CUDA(c, cudaSetDevice(c->device));
    for (int i = 0; i < TWG_NUM_STREAMS; ++i) TWG_CUDA(c, cudaStreamSynchronize(c->streams[i]));
    return 0;
}

int twg_set_option(twg_ctx* c, const char* name, double value) {
    TWG_CHECK(c, c && name, TWG_ERR_INVALID_ARG, "null argument");
    const OptDesc* d = find_opt(name);
    TWG_CHECK(c, d != nullptr, TWG_ERR_INVALID_ARG, "unknown option");
    set_opt(c->opt, *d, (long long)value);
    for (twg_ctx* k : c->children) set_opt(k->opt, *d, (long long)value);
    return 0;
}

int twg_get_option(const twg_ctx* c, const char* name, double* value) {
    TWG_CHECK(c, c && name && value, TWG_ERR_INVALID_ARG, "null argument");
    const OptDesc* d = find_opt(name);
    TWG_CHECK(c, d != nullptr, TWG_ERR_INVALID_ARG, "unknown option");
    *value = d->i ? (double)(c->opt.*(d->i)) : (double)(c->opt.*(d->ll));
    return 0;
}

int twg_debug_counter(twg_ctx* c, int which, uint64_t* value) {
    TWG_CHECK(c, c && value && which >= 0 && which < TWG_NUM_DEBUG_COUNTERS, TWG_ERR_INVALID_ARG, "bad argument");
    *value = 0;
    if (!c->children.empty()) {
        for (twg_ctx* k : c->children) {
            uint64_t v = 0;
            TWG_TRY(twg_debug_counter(k, which, &v));
            *value += v;
        }
        return 0;
    }
    TWG_CUDA(c, cudaSetDevice(c->device));
    TWG_TRY(twg_synchronize(c));
    for (twg_lane& l : c->lanes)
        if (l.done) TWG_CUDA(c, cudaEventSynchronize(l.done));
    unsigned long long v = 0;
    TWG_CUDA(c, cudaMemcpy(&v, c->dcounters + which, sizeof(v), cudaMemcpyDeviceToHost));
    *value = v;
    return 0;
}

}  // extern "C"

int twg_get_lane(twg_ctx* c, cudaStream_t st, twg_lane** out) {
    ++c->lane_tick;
    for (twg_lane& l : c->lanes)
        if (l.stream == st) { l.tick = c->lane_tick; *out = &l; return 0; }
    if (c->lanes.size() < TWG_MAX_LANES) {
        c->lanes.emplace_back();
        TWG_TRY(init_lane(c, c->lanes.back(), st));
        c->lanes.back().tick = c->lane_tick;
        *out = &c->lanes.back();
        return 0;
    }
    // recycle the least recently used caller lane once everything it launched has completed
    twg_lane* lru = nullptr;
    for (size_t k = TWG_NUM_STREAMS; k < c->lanes.size(); ++k)
        if (!lru || c->lanes[k].tick < lru->tick) lru = &c->lanes[k];
    TWG_CUDA(c, cudaEventSynchronize(lru->done));
    lru->stream = st;
    lru->tick = c->lane_tick;
    *out = lru;
    return 0;
}

int twg_lane_mark(twg_ctx* c, twg_lane* lane) {
    // only caller-stream lanes are ever recycled; the context's own streams are synchronised by the host entry points
    if (lane >= c->lanes.data() && lane < c->lanes.data() + TWG_NUM_STREAMS) return 0;
    TWG_CUDA(c, cudaEventRecord(lane->done, lane->stream));
    return 0;
}

namespace {
__global__ void fast_signal_kernel(volatile uint32_t* flag, uint32_t seq) {
    if (threadIdx.x == 0) {
        __threadfence_system();
        *flag = seq;
    }
}
}  // namespace

int twg_fast_slab(twg_ctx* c, size_t bytes, char** slab) {
    const size_t need = bytes + 256;  // the first 256 bytes hold the completion word
    if (c->fast_bytes < need) {
        if (c->fast_slab) {
            TWG_CUDA(c, cudaStreamSynchronize(c->streams[0]));
            TWG_CUDA(c, cudaFreeHost(c->fast_slab));
            c->fast_slab = nullptr;
            c->fast_bytes = 0;
        }
        const size_t want = need < (1u << 16) ? (1u << 16) : need * 2;
        TWG_CUDA(c, cudaHostAlloc(&c->fast_slab, want, cudaHostAllocMapped | cudaHostAllocPortable));
        memset(c->fast_slab, 0, 256);
        c->fast_bytes = want;
    }
    *slab = (char*)c->fast_slab + 256;
    return 0;
}

int twg_fast_arm(twg_ctx* c, twg_done* done) {
    if (!c->fast_counter) {
        TWG_CUDA(c, cudaMalloc(&c->fast_counter, 256));
        TWG_CUDA(c, cudaMemset(c->fast_counter, 0, 256));
    }
    if (!c->fast_slab) {
        char* slab;
        TWG_TRY(twg_fast_slab(c, 0, &slab));
    }
    done->flag = (volatile uint32_t*)c->fast_slab;
    done->counter = c->fast_counter;
    done->seq = ++c->fast_seq ? c->fast_seq : ++c->fast_seq;  // never 0
    return 0;
}

namespace {
__global__ void fast_null_kernel(twg_done done) { twg_signal_done(done); }
}  // namespace

// the floor under every tiny call on this machine: launch an EMPTY kernel that raises the completion word, spin until the host
// sees it (host clock around `reps` such round trips)
int twg_debug_roundtrip(twg_ctx* c, int reps, double* us_per_call) {
    TWG_CHECK(c, c && us_per_call && reps > 0, TWG_ERR_INVALID_ARG, "bad argument");
    if (twg_is_multi(c)) return twg_forward0(c, twg_debug_roundtrip(c->children[0], reps, us_per_call));
    TWG_CUDA(c, cudaSetDevice(c->device));
    cudaStream_t st = c->streams[0];
    timespec t0, t1;
    for (int i = -5; i < reps; ++i) {
        if (i == 0) clock_gettime(CLOCK_MONOTONIC, &t0);
        twg_done done;
        TWG_TRY(twg_fast_arm(c, &done));
        TWG_LAUNCH(c, fast_null_kernel, 1, 32, 0, st, done);
        TWG_TRY(twg_fast_spin(c, st, done.seq));
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    *us_per_call = ((t1.tv_sec - t0.tv_sec) * 1e6 + (t1.tv_nsec - t0.tv_nsec) * 1e-3) / reps;
    return 0;
}

int twg_fast_wait(twg_ctx* c, cudaStream_t st) {
    volatile uint32_t* flag = (volatile uint32_t*)c->fast_slab;
    const uint32_t seq = ++c->fast_seq ? c->fast_seq : ++c->fast_seq;  // never 0
    fast_signal_kernel<<<1, 32, 0, st>>>(flag, seq);
    c->launches++;
    TWG_CUDA(c, cudaGetLastError());
    return twg_fast_spin(c, st, seq);
}

int twg_fast_spin(twg_ctx* c, cudaStream_t st, uint32_t seq) {
    volatile uint32_t* flag = (volatile uint32_t*)c->fast_slab;
    // spin: the word arrives over PCIe right after the results (posted writes of one device stay in order)
    for (unsigned long long spins = 0; *flag != seq; ++spins) {
        if ((spins & 0xfffff) == 0xfffff) {  // ~ every few ms: has the stream failed?
            const cudaError_t e = cudaStreamQuery(st);
            if (e != cudaSuccess && e != cudaErrorNotReady) return twg_fail(c, (int)e, cudaGetErrorString(e), __FILE__, __LINE__);
            if (e == cudaSuccess && *flag != seq) {  // finished, word not seen yet: one proper synchronize settles it
                TWG_CUDA(c, cudaStreamSynchronize(st));
                break;
            }
        }
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
    }
    return 0;
}

int twg_ensure_scratch(twg_ctx* c, int slot, size_t bytes) {
    if (c->dscratch_bytes[slot] >= bytes) return 0;
    if (c->dscratch[slot]) {
        TWG_CUDA(c, cudaStreamSynchronize(c->streams[slot]));
        TWG_CUDA(c, cudaFree(c->dscratch[slot]));
        c->dscratch[slot] = nullptr;
        c->dscratch_bytes[slot] = 0;
    }
    size_t want = bytes + bytes / 8 + 256;
    TWG_CUDA(c, cudaMalloc(&c->dscratch[slot], want));
    c->dscratch_bytes[slot] = want;
    return 0;
}

int twg_ensure_pinned(twg_ctx* c, size_t in_bytes, size_t out_bytes) {
    if (c->pin_in_bytes < in_bytes) {
        for (int i = 0; i < TWG_NUM_STREAMS; ++i) {
            if (c->pin_in[i]) { TWG_CUDA(c, cudaStreamSynchronize(c->streams[i])); TWG_CUDA(c, cudaFreeHost(c->pin_in[i])); c->pin_in[i] = nullptr; }
            TWG_CUDA(c, cudaHostAlloc(&c->pin_in[i], in_bytes, cudaHostAllocDefault));
        }
        c->pin_in_bytes = in_bytes;
    }
    if (c->pin_out_bytes < out_bytes) {
        for (int i = 0; i < TWG_NUM_STREAMS; ++i) {
            if (c->pin_out[i]) { TWG_CUDA(c, cudaStreamSynchronize(c->streams[i])); TWG_CUDA(c, cudaFreeHost(c->pin_out[i])); c->pin_out[i] = nullptr; }
            TWG_CUDA(c, cudaHostAlloc(&c->pin_out[i], out_bytes, cudaHostAllocDefault));
        }
        c->pin_out_bytes = out_bytes;
    }
    return 0;
}
