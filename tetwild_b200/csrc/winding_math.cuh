// winding_math.cuh -- the per-factor arithmetic of the winding-number kernel (csrc/winding.cu), kept in a header of
// __host__ __device__ functions so that the CPU-only tests (tests/host_harness.cpp) execute the very same source:
// |v| with one third-order correction, and the branch-free running sum of atan2 angles as a complex product.
// The device path uses the intrinsics; the host path restates them with memcpy / std::sqrt (same results for the integer
// bookkeeping; the seed of the reciprocal square root differs, the corrected value agrees to ~1e-16).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include "tw_math.cuh"

namespace tww {

#if defined(__CUDA_ARCH__)
TW_HD int hi32(double x) { return __double2hiint(x); }
TW_HD double from_hi(int hi) { return __hiloint2double(hi, 0); }
TW_HD int imax(int a, int b) { return max(a, b); }
TW_HD double rsqrt_seed(double l2) {
    double y0;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(l2));
    return y0;
}
#else
TW_HD int hi32(double x) { uint64_t u; std::memcpy(&u, &x, 8); return (int)(uint32_t)(u >> 32); }
TW_HD double from_hi(int hi) { const uint64_t u = (uint64_t)(uint32_t)hi << 32; double x; std::memcpy(&x, &u, 8); return x; }
TW_HD int imax(int a, int b) { return a > b ? a : b; }
TW_HD double rsqrt_seed(double l2) {  // ~22 correct bits, like MUFU.RSQ64H: 1/sqrt with the low 31 mantissa bits dropped
    const double y = 1.0 / std::sqrt(l2);
    uint64_t u;
    std::memcpy(&u, &y, 8);
    u &= 0xffffffff80000000ull;
    double r;
    std::memcpy(&r, &u, 8);
    return r;
}
#endif

// |v| for the solid-angle terms. Hardware seed y0 ~ 1/sqrt(l2) (MUFU.RSQ64H, relative error < 2^-21) followed by ONE
// third-order correction: with s = l2*y0 and e = 1 - s*y0 (exact to rounding through the fma), sqrt(l2) =
// s*(1 + e/2 + 3e^2/8 + O(e^3)); the neglected term is < 2^-63 relative. 5 FP64 instructions after the seed, branch free
// (CUDA's sqrt() adds a slow-path call per use; two Newton steps on the reciprocal root cost 8).
// The 1e-300 folded into the sum of squares only matters for a query ON a vertex (length 0): the factor then degenerates
// to a positive real.
TW_HD double norm3(double x, double y, double z) {
    const double l2 = fma(x, x, fma(y, y, fma(z, z, 1e-300)));
    const double y0 = rsqrt_seed(l2);
    const double s = l2 * y0;
    const double e = fma(-s, y0, 1.0);
    const double p = fma(0.375, e, 0.5);
    return fma(s * e, p, s);
}

// Running sum of atan2(y_f, x_f) as arg(z) + 2 pi k, z = prod (x_f + i y_f) (see the header comment).
// BRANCH FREE: the loop bodies that call mul() must stay straight-line code so that the compiler can interleave the
// unrolled iterations (the complex product is a serial chain of dependent FP64 operations; everything else of the next
// point overlaps it). Half planes are told apart by the SIGN BIT of Im z, exactly like atan2 treats signed zeros:
// U = {sign clear, arg in [+0, pi]}, L = {sign set, arg in [-pi, -0]}. A counter-clockwise factor (sign of y clear,
// angle in [0, pi]) that takes z from U to L went through the negative real axis (k += 1); a clockwise factor from L to U
// went through it the other way (k -= 1); U<->L moves in the other pairings cross the POSITIVE axis and change nothing.
// A sign of Im z that rounding gets "wrong" can only happen next to the negative axis (next to the positive axis both
// products of zr*y + zi*x have the same sign), where arg + 2 pi k is continuous, so the bookkeeping stays exact.
struct Angle {
    double zr, zi;
    int k;
    TW_HD void init() { zr = 1.0; zi = 0.0; k = 0; }
    // z *= (x + i y) * 2^-e, e = the larger binary exponent of x, y (integer pipe); skip = chain start / zero factor
    TW_HD void mul(double x, double y, bool skip) {
        const int e = imax(hi32(x) & 0x7ff00000, hi32(y) & 0x7ff00000);
        skip = skip || e == 0;       // x = y = 0 (atan2(0,0) = 0 in the reference) or no triangle here: multiply by 1
        const double sc = from_hi(0x7fe00000 - e);
        x *= sc; y *= sc;            // |x + i y| in [1, 2 sqrt 2)
        x = skip ? 1.0 : x;
        y = skip ? 0.0 : y;
        const double nr = fma(zr, x, -(zi * y));
        const double ni = fma(zr, y, zi * x);
        const int hz = hi32(zi), hn = hi32(ni), hy = hi32(y);
        const int m = (hz ^ hn) & ~(hy ^ hz);  // sign bit: half plane changed AND the factor turns away from the old half plane
        k += (m >> 31) & (2 * (hz >> 31) + 1);
        zr = nr; zi = ni;
    }
    TW_HD void renorm() {  // after at most 32 factors: |z| < 2^49 -> back to [1, 2 sqrt 2)
        const int e = imax(hi32(zr) & 0x7ff00000, hi32(zi) & 0x7ff00000);
        const double sc = from_hi(0x7fe00000 - e);
        zr *= sc; zi *= sc;
    }
    TW_HD double total() const { return atan2(zi, zr) + 6.283185307179586476925 * (double)k; }
};

}  // namespace tww
