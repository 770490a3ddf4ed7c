// ctx.cu -- context lifetime, scratch / pinned staging, error reporting (C ABI: include/tetwild_gpu.h).
#include "common.cuh"

extern "C" {

const char* twg_version(void) { return "tetwild_b200 0.1 (sm_100a)"; }

int twg_create(twg_ctx** out, int device_id) {
    if (!out) return TWG_ERR_INVALID_ARG;
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) return TWG_ERR_NO_DEVICE;  // no CPU fallback, by design
    if (device_id < 0 || device_id >= ndev) return TWG_ERR_INVALID_ARG;
    twg_ctx* c = new twg_ctx;
    c->device = device_id;
    if (cudaSetDevice(device_id) != cudaSuccess) { delete c; return TWG_ERR_NO_DEVICE; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device_id) != cudaSuccess) { delete c; return TWG_ERR_NO_DEVICE; }
    if (prop.major < 10) {  // the fatbin holds sm_100a SASS only
        delete c;
        return TWG_ERR_NO_DEVICE;
    }
    c->sm_count = prop.multiProcessorCount;
    for (int i = 0; i < TWG_NUM_STREAMS; ++i) {
        if (cudaStreamCreateWithFlags(&c->streams[i], cudaStreamNonBlocking) != cudaSuccess) { twg_destroy(c); return TWG_ERR_INTERNAL; }
        if (cudaEventCreateWithFlags(&c->ev[i], cudaEventDisableTiming) != cudaSuccess) { twg_destroy(c); return TWG_ERR_INTERNAL; }
    }
    *out = c;
    return 0;
}

void twg_destroy(twg_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    for (int i = 0; i < TWG_NUM_STREAMS; ++i) {
        if (c->streams[i]) { cudaStreamSynchronize(c->streams[i]); cudaStreamDestroy(c->streams[i]); }
        if (c->ev[i]) cudaEventDestroy(c->ev[i]);
        if (c->pin_in[i]) cudaFreeHost(c->pin_in[i]);
        if (c->pin_out[i]) cudaFreeHost(c->pin_out[i]);
        if (c->dscratch[i]) cudaFree(c->dscratch[i]);
    }
    for (int i = 0; i <= TWG_NUM_STREAMS; ++i)
        if (c->dsort[i]) cudaFree(c->dsort[i]);
    delete c;
}

const char* twg_last_error(const twg_ctx* c) { return c ? c->err : "null context"; }
int twg_device(const twg_ctx* c) { return c ? c->device : -1; }
uint64_t twg_launch_count(const twg_ctx* c) { return c ? c->launches : 0; }

int twg_synchronize(twg_ctx* c) {
    if (!c) return TWG_ERR_INVALID_ARG;
    TWG_CUDA(c, cudaSetDevice(c->device));
    for (int i = 0; i < TWG_NUM_STREAMS; ++i) TWG_CUDA(c, cudaStreamSynchronize(c->streams[i]));
    return 0;
}

}  // extern "C"

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
