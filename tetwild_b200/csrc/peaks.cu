// peaks.cu -- roofline denominators measured on the device the context owns ("measure a DFMA
// microbenchmark and record it next to the HBM number"). Not on the hot path; bench.py calls these once per run.
//
//   twg_measure_fp64_tflops   dependent-chain DFMA kernel, 8 independent accumulators per thread, 148 x 8 CTAs x 256
//                             threads: the FP64 vector pipe at full issue rate (2 flops per DFMA)
//   twg_measure_copy_gbs      128-bit grid-stride device copy (read + write bytes), the same definition as the
//                             driver's MEASURED_PEAKS.json hbm_gbs
#include "common.cuh"

namespace {

constexpr int kChains = 8;
constexpr int kInner = 512;

__global__ void __launch_bounds__(256) dfma_kernel(double* out, double a, double b, int outer) {
    double x[kChains];
#pragma unroll
    for (int k = 0; k < kChains; ++k) x[k] = (double)(threadIdx.x + k) * 1e-3;
    for (int o = 0; o < outer; ++o) {
#pragma unroll 16
        for (int i = 0; i < kInner; ++i) {
#pragma unroll
            for (int k = 0; k < kChains; ++k) x[k] = __fma_rn(x[k], a, b);
        }
    }
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < kChains; ++k) s += x[k];
    if (s == 123.456) out[0] = s;  // never true for the chosen a, b: keeps the chain alive without a store
}

// The same chain with THREE DISTINCT register operands per DFMA (a[k], b[k] live in registers, loaded from memory so that
// they cannot be folded): what real FP64 code (cross products, dot products) looks like to the register file. If this
// variant is slower than dfma_kernel, the gap is operand bandwidth, and it is the ceiling a kernel like winding_kernel
// (FP64 pipe ~67 % active with math_pipe_throttle stalls) actually runs against.
__global__ void __launch_bounds__(256) dfma3_kernel(double* out, const double* __restrict__ coef, int outer) {
    double x[kChains], a[kChains], b[kChains];
#pragma unroll
    for (int k = 0; k < kChains; ++k) {
        x[k] = (double)(threadIdx.x + k) * 1e-3;
        a[k] = coef[k] - 1e-12 * threadIdx.x;           // per-thread values: vector registers, not uniform ones
        b[k] = coef[kChains + k] + 1e-12 * threadIdx.x;
    }
    for (int o = 0; o < outer; ++o) {
#pragma unroll 16
        for (int i = 0; i < kInner; ++i) {
#pragma unroll
            for (int k = 0; k < kChains; ++k) x[k] = __fma_rn(x[k], a[k], b[k]);
        }
    }
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < kChains; ++k) s += x[k];
    if (s == 123.456) out[0] = s;
}

__global__ void __launch_bounds__(256) copy_kernel(const double2* __restrict__ src, double2* __restrict__ dst, uint64_t n2) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (uint64_t)gridDim.x * blockDim.x) {
        double2 v = ld_stream2(reinterpret_cast<const double*>(src + i));
        st_stream2(reinterpret_cast<double*>(dst + i), v);
    }
}

}  // namespace

extern "C" {

int twg_measure_fp64_tflops(twg_ctx* c, double* tflops) {
    TWG_CHECK(c, c && tflops, TWG_ERR_INVALID_ARG, "null argument");
    if (twg_is_multi(c)) return twg_forward0(c, twg_measure_fp64_tflops(c->children[0], tflops));
    TWG_CUDA(c, cudaSetDevice(c->device));
    TWG_TRY(twg_ensure_scratch(c, 0, 256));
    cudaStream_t st = c->streams[0];
    cudaEvent_t e0, e1;
    TWG_CUDA(c, cudaEventCreate(&e0));
    TWG_CUDA(c, cudaEventCreate(&e1));
    const unsigned grid = (unsigned)c->sm_count * 8;
    const int outer = 64;
    double best = 0.0;
    for (int rep = 0; rep < 6; ++rep) {
        TWG_CUDA(c, cudaEventRecord(e0, st));
        TWG_LAUNCH(c, dfma_kernel, grid, 256, 0, st, (double*)c->dscratch[0], 0.999999, 1e-7, outer);
        TWG_CUDA(c, cudaEventRecord(e1, st));
        TWG_CUDA(c, cudaEventSynchronize(e1));
        float ms = 0.f;
        TWG_CUDA(c, cudaEventElapsedTime(&ms, e0, e1));
        const double flops = 2.0 * kChains * kInner * (double)outer * 256.0 * grid;
        const double tf = flops / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;  // first launch is the warm-up
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *tflops = best;
    return 0;
}

int twg_measure_fp64_tflops_distinct(twg_ctx* c, double* tflops) {
    TWG_CHECK(c, c && tflops, TWG_ERR_INVALID_ARG, "null argument");
    if (twg_is_multi(c)) return twg_forward0(c, twg_measure_fp64_tflops_distinct(c->children[0], tflops));
    TWG_CUDA(c, cudaSetDevice(c->device));
    TWG_TRY(twg_ensure_scratch(c, 0, 4096));
    cudaStream_t st = c->streams[0];
    double h[2 * kChains];
    for (int k = 0; k < kChains; ++k) { h[k] = 0.999999 - 1e-7 * k; h[kChains + k] = 1e-7 * (k + 1); }
    double* coef = (double*)c->dscratch[0] + 32;
    TWG_CUDA(c, cudaMemcpyAsync(coef, h, sizeof(h), cudaMemcpyHostToDevice, st));
    cudaEvent_t e0, e1;
    TWG_CUDA(c, cudaEventCreate(&e0));
    TWG_CUDA(c, cudaEventCreate(&e1));
    const unsigned grid = (unsigned)c->sm_count * 8;
    const int outer = 64;
    double best = 0.0;
    for (int rep = 0; rep < 6; ++rep) {
        TWG_CUDA(c, cudaEventRecord(e0, st));
        TWG_LAUNCH(c, dfma3_kernel, grid, 256, 0, st, (double*)c->dscratch[0], coef, outer);
        TWG_CUDA(c, cudaEventRecord(e1, st));
        TWG_CUDA(c, cudaEventSynchronize(e1));
        float ms = 0.f;
        TWG_CUDA(c, cudaEventElapsedTime(&ms, e0, e1));
        const double tf = 2.0 * kChains * kInner * (double)outer * 256.0 * grid / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *tflops = best;
    return 0;
}

int twg_measure_copy_gbs(twg_ctx* c, uint64_t bytes, double* gbs) {
    TWG_CHECK(c, c && gbs, TWG_ERR_INVALID_ARG, "null argument");
    TWG_CHECK(c, bytes >= (1ull << 20), TWG_ERR_INVALID_ARG, "need at least 1 MiB");
    if (twg_is_multi(c)) return twg_forward0(c, twg_measure_copy_gbs(c->children[0], bytes, gbs));
    TWG_CUDA(c, cudaSetDevice(c->device));
    bytes &= ~(uint64_t)255;
    void *a = nullptr, *b = nullptr;
    TWG_CUDA(c, cudaMalloc(&a, bytes));
    if (cudaMalloc(&b, bytes) != cudaSuccess) { cudaFree(a); return twg_fail(c, TWG_ERR_INTERNAL, "cudaMalloc", __FILE__, __LINE__); }
    cudaStream_t st = c->streams[0];
    cudaMemsetAsync(a, 1, bytes, st);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    double best = 0.0;
    int rc = 0;
    for (int rep = 0; rep < 8 && rc == 0; ++rep) {
        cudaEventRecord(e0, st);
        copy_kernel<<<(unsigned)c->sm_count * 16, 256, 0, st>>>((const double2*)a, (double2*)b, bytes / 16);
        c->launches++;
        cudaEventRecord(e1, st);
        if (cudaEventSynchronize(e1) != cudaSuccess) { rc = twg_fail(c, TWG_ERR_INTERNAL, "copy kernel failed", __FILE__, __LINE__); break; }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double g = 2.0 * (double)bytes / (ms * 1e-3) / 1e9;
        if (rep > 0 && g > best) best = g;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(a);
    cudaFree(b);
    *gbs = best;
    return rc;
}

}  // extern "C"
