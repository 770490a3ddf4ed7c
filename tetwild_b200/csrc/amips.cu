// amips.cu -- batched conformal-AMIPS kernels (FP64) and their C ABI.
//
// Replaces, for batches: energy_ispc (src/ispc/energy.ispc:7-65), comformalAMIPS{Energy,Jacobian,Hessian}_new
// (src/tetwild/LocalOperations.cpp:28-291), calTetQualities (:695-773, :862-884), VertexSmoother::NewtonsUpdate
// (src/tetwild/VertexSmoother.cpp:627-702) and getNewEnergy (:544-625). Math: tw_math.cuh::amips_eval.
//
// Kernels
//   amips_soa_tma_kernel flat SoA E+J+H over full 256-tet tiles: two coalesced loads per coordinate array, J3/H9 rows (AoS per
//                        tet, as the reference lays them out) written conflict-free into shared memory and stored by TMA
//                        bulk shared->global copies. HBM-bound: 96 B in + 104 B out per tet (0.93 of the copy peak).
//   amips_soa_kernel     the generic form (energy only, ragged tails, unaligned buffers): 128-bit streaming loads, J3/H9
//                        transposed through shared memory and copied out with 128-bit stores.
//   amips_quality_kernel indexed gather (int4 tet load + 4 vertex gathers), exact orientation gate.
//   amips_ring_kernel    one warp per one-ring, lanes over member tets, software-pipelined over the warp's rings
//                        (index chain of later rings in flight), 10 sums reduced through a shared-memory transpose.
//                        Bound by instruction issue + gather latency (0.40 of the HBM peak; five other shapes measured, see below).
//   amips_ring_tiny_kernel the same body for ONE un-batched call (<= 32 rings): vertex ids / trial positions in the kernel
//                        parameters, completion word raised by the kernel (common.cuh::twg_signal_done).
#include "common.cuh"

namespace {

constexpr int AM_THREADS = 128;

struct SoaArgs {
    const double* T[12];
    double* E;
    double* J3;
    double* H9;
    uint64_t n;
};

template <int VEC, bool JH>
__global__ void __launch_bounds__(AM_THREADS) amips_soa_kernel(SoaArgs a) {
    constexpr int TILE = AM_THREADS * VEC;
    extern __shared__ __align__(16) double sm[];  // JH: [TILE*3] J then [TILE*9] H
    double* sJ = sm;
    double* sH = sm + TILE * 3;
    const uint64_t n = a.n;
    for (uint64_t base = (uint64_t)blockIdx.x * TILE; base < n; base += (uint64_t)gridDim.x * TILE) {
        const uint64_t i0 = base + (uint64_t)threadIdx.x * VEC;
        double x[VEC][12];
        if (i0 + VEC <= n) {
            if (VEC == 2) {
#pragma unroll
                for (int k = 0; k < 12; ++k) {
                    double2 v = ld_stream2(a.T[k] + i0);
                    x[0][k] = v.x;
                    x[VEC - 1][k] = v.y;
                }
            } else {
#pragma unroll
                for (int k = 0; k < 12; ++k) x[0][k] = __ldg(a.T[k] + i0);
            }
        } else {
#pragma unroll
            for (int v = 0; v < VEC; ++v)
#pragma unroll
                for (int k = 0; k < 12; ++k) x[v][k] = (i0 + v < n) ? __ldg(a.T[k] + i0 + v) : (double)((k * 7 + k / 3) % 5);
        }
        tw::Amips r[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) tw::amips_eval<JH>(x[v], r[v]);
        if (a.E) {
            if (VEC == 2 && i0 + VEC <= n) st_stream2(a.E + i0, make_double2(r[0].E, r[VEC - 1].E));
            else {
#pragma unroll
                for (int v = 0; v < VEC; ++v)
                    if (i0 + v < n) a.E[i0 + v] = r[v].E;
            }
        }
        if (JH) {
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                const int tl = threadIdx.x * VEC + v;
                sJ[tl * 3 + 0] = r[v].J[0]; sJ[tl * 3 + 1] = r[v].J[1]; sJ[tl * 3 + 2] = r[v].J[2];
                double* h = sH + tl * 9;
                h[0] = r[v].H[0]; h[1] = r[v].H[1]; h[2] = r[v].H[2];
                h[3] = r[v].H[1]; h[4] = r[v].H[3]; h[5] = r[v].H[4];
                h[6] = r[v].H[2]; h[7] = r[v].H[4]; h[8] = r[v].H[5];
            }
            __syncthreads();
            const uint64_t cnt = (n - base < (uint64_t)TILE) ? (n - base) : (uint64_t)TILE;
            if (a.J3) {
                double* dst = a.J3 + base * 3;
                const int tot = (int)cnt * 3;
                if (VEC == 2) {  // launcher guarantees 16-byte aligned outputs; TILE*3 is even
                    for (int k = threadIdx.x * 2; k + 1 < tot; k += AM_THREADS * 2) st_stream2(dst + k, make_double2(sJ[k], sJ[k + 1]));
                    if ((tot & 1) && threadIdx.x == 0) dst[tot - 1] = sJ[tot - 1];
                } else {
                    for (int k = threadIdx.x; k < tot; k += AM_THREADS) dst[k] = sJ[k];
                }
            }
            if (a.H9) {
                double* dst = a.H9 + base * 9;
                const int tot = (int)cnt * 9;
                if (VEC == 2) {
                    for (int k = threadIdx.x * 2; k + 1 < tot; k += AM_THREADS * 2) st_stream2(dst + k, make_double2(sH[k], sH[k + 1]));
                    if ((tot & 1) && threadIdx.x == 0) dst[tot - 1] = sH[tot - 1];
                } else {
                    for (int k = threadIdx.x; k < tot; k += AM_THREADS) dst[k] = sH[k];
                }
            }
            __syncthreads();
        }
    }
}

// E+J+H for full, 16-byte aligned tiles: the output path is TMA. Thread t evaluates tets t and t+128 of a 256-tet tile
// (two coalesced 64-bit loads per coordinate array), writes its J rows (3 doubles) and H rows (9 doubles) into shared
// memory in the FINAL AoS layout -- consecutive lanes are 3 resp. 9 doubles apart, which is conflict-free for 64-bit
// stores -- and one thread hands the tile's 6 KiB of J and 18 KiB of H to two bulk shared->global copies
// (cp.async.bulk, UBLKCP in SASS). The copy-out loop of amips_soa_kernel (24 LDS.128 + STG.128 per thread and tile) and
// one of its two barriers disappear, and the store drains while the next tile is loaded and evaluated: the buffer is
// only waited for (wait_group.read) right before it is overwritten. The ragged last tile goes through amips_soa_kernel.
__global__ void __launch_bounds__(AM_THREADS) amips_soa_tma_kernel(SoaArgs a, uint64_t n_tiles) {
    constexpr int TILE = AM_THREADS * 2;
    extern __shared__ __align__(128) double sm[];  // [TILE*3] J then [TILE*9] H
    double* sJ = sm;
    double* sH = sm + TILE * 3;
    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint64_t base = tile * TILE;
        double x[2][12];
#pragma unroll
        for (int k = 0; k < 12; ++k) {
            x[0][k] = __ldg(a.T[k] + base + threadIdx.x);
            x[1][k] = __ldg(a.T[k] + base + AM_THREADS + threadIdx.x);
        }
        tw::Amips r[2];
#pragma unroll
        for (int v = 0; v < 2; ++v) tw::amips_eval<true>(x[v], r[v]);
        if (a.E) {
            a.E[base + threadIdx.x] = r[0].E;
            a.E[base + AM_THREADS + threadIdx.x] = r[1].E;
        }
        if (threadIdx.x == 0) tma_store_wait_read();  // the previous tile's bulk stores have read the buffer
        __syncthreads();
#pragma unroll
        for (int v = 0; v < 2; ++v) {
            const int row = v * AM_THREADS + threadIdx.x;
            double* j = sJ + row * 3;
            j[0] = r[v].J[0]; j[1] = r[v].J[1]; j[2] = r[v].J[2];
            double* h = sH + row * 9;
            h[0] = r[v].H[0]; h[1] = r[v].H[1]; h[2] = r[v].H[2];
            h[3] = r[v].H[1]; h[4] = r[v].H[3]; h[5] = r[v].H[4];
            h[6] = r[v].H[2]; h[7] = r[v].H[4]; h[8] = r[v].H[5];
        }
        fence_proxy_async_smem();
        __syncthreads();
        if (threadIdx.x == 0) {
            if (a.J3) tma_bulk_s2g(a.J3 + base * 3, sJ, TILE * 3 * (uint32_t)sizeof(double));
            if (a.H9) tma_bulk_s2g(a.H9 + base * 9, sH, TILE * 9 * (uint32_t)sizeof(double));
            tma_store_commit();
        }
    }
    if (threadIdx.x == 0) tma_store_wait_all();  // global writes complete before the CTA exits
}

__device__ __forceinline__ void gather_vertex(const double* __restrict__ V, int32_t v, double* dst) {
    const double* p = V + 3 * (size_t)v;
    dst[0] = __ldg(p); dst[1] = __ldg(p + 1); dst[2] = __ldg(p + 2);
}

// calTetQuality_AMIPS (LocalOperations.cpp:862-884)
// A tet that names a vertex outside [0, nV) is never dereferenced: it gives MAX_ENERGY and raises the context's bad-index
// counter (the host entry point turns that into TWG_ERR_INVALID_ARG).
__global__ void __launch_bounds__(256, 3) amips_quality_kernel(const double* __restrict__ V, uint32_t nV, const int4* __restrict__ tets, uint64_t nT,
                                                            double* __restrict__ slim, unsigned long long* dbg) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nT; i += (uint64_t)gridDim.x * blockDim.x) {
        const int4 t = __ldg(tets + i);
        if ((uint32_t)t.x >= nV || (uint32_t)t.y >= nV || (uint32_t)t.z >= nV || (uint32_t)t.w >= nV) {
            slim[i] = TWG_MAX_ENERGY;
            atomicAdd(dbg + TWG_DBG_BAD_INDEX, 1ull);
            continue;
        }
        double x[12];
        gather_vertex(V, t.x, x); gather_vertex(V, t.y, x + 3); gather_vertex(V, t.z, x + 6); gather_vertex(V, t.w, x + 9);
        double e;
        if (tw::exact::cgal_orientation(x, x + 3, x + 6, x + 9) != 1) {
            e = TWG_MAX_ENERGY;
        } else {
            tw::Amips r;
            tw::amips_eval<false>(x, r);
            e = r.E;
        }
        if (isinf(e) || isnan(e) || e <= 0.0) e = TWG_MAX_ENERGY;
        slim[i] = e;
    }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// NewtonsUpdate (VertexSmoother.cpp:627-702) / getNewEnergy (:544-625): one warp per one-ring, lanes over member tets.
//
// The address chain of a ring is four dependent loads deep (vids[g] -> off[row] -> t_ids[k] -> tets[ti] -> V[v]) and a ring
// gives a warp only ~24 tets of arithmetic, so the kernel is a SOFTWARE PIPELINE over the warp's rings: while ring i is
// gathered and evaluated, the tet records of ring i+1, the member ids of ring i+2, the CSR bounds of ring i+3 and the row of
// ring i+4 are in flight (each stage consumes what the previous iteration loaded). Only the vertex gather of the current
// ring is waited for.
//   * the centre vertex is rotated to slot 0 (:640-651), so it is the same vertex for every member: loaded once per ring;
//   * the closed-form Hessian is symmetric: 10 sums per ring (E, J[3], H[6]), reduced through a per-warp shared-memory
//     transpose (10 stores + 16 loads per lane) instead of 13 x 5 64-bit shuffle steps; lane (c, h) adds the members
//     16h .. 16h+15 of component c in lane order, the two halves meet in one shuffle -- deterministic, and closer to the
//     reference's sequential ring order than a butterfly.
//
// Round 2 measured three other shapes of this kernel against it (profiles/r02_ring_variants.txt) and kept this one:
//   * ncu (profiles/r02_amips_ring_kernel.txt): 21 warp instructions per tet at 58 % issue utilisation, FP64 pipe 30 %, L1 59 %,
//     DRAM 34 % -- the kernel is bound by instruction issue and latency together, not by HBM; the closed-form evaluation is
//     only ~200 of the ~520 instructions a ring costs, the rest is gather, address chain and reduction;
//   * vertices of ring i+1 staged in shared memory with cp.async, or prefetched into L2: 11.7 / 17.2 G tets/s (more
//     instructions than the latency they hide);
//   * rings of a warp walked as ONE dense member stream, 32 members per step, in-order segmented reduction in shared memory:
//     lanes full (31.4 in the evaluation) but 340 instructions per step of reduction: 25 instructions per tet, 24.8 G tets/s
//     (profiles/r02_amips_ring_flat_kernel_not_kept.txt);
//   * one ring per LANE, members added in registers (no reduction at all, 12.9 instructions per tet): 29.1 G tets/s on 16 M
//     tets, but every lane is then its own memory stream -- 113 664 of them -- and beyond ~25 M tets (2 GB of mesh) the kernel
//     drops to 6.4 G tets/s with every unit idle (DRAM 9 %, L1 16 %, issue 34 %, long-scoreboard 10.8: profiles/
//     r02_amips_ring_lane_kernel_not_kept.txt), whether a lane's rings are interleaved or consecutive.
struct RingHead {   // stage result: CSR bounds and centre of one ring (member offsets fit 32 bits: at most 4 * 2^29 members)
    uint32_t b, cnt;
    int32_t c;
};

// `vids` and `trial` are read with plain loads: the tiny-call kernel below passes them in shared memory
template <bool ENERGY_ONLY>
__device__ __forceinline__ void ring_warp_body(const double* __restrict__ V, const int4* __restrict__ tets, const int32_t* __restrict__ t_ids,
                                               const uint64_t* __restrict__ off, const int32_t* __restrict__ center, const int32_t* vids, uint64_t nG,
                                               double* __restrict__ E, double* __restrict__ J3, double* __restrict__ H9, uint8_t* __restrict__ ok,
                                               uint32_t nV, uint64_t nT, unsigned long long* dbg, const double* trial) {
    constexpr int NRED = ENERGY_ONLY ? 1 : 10;
    __shared__ double red[8][NRED][33];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const uint64_t gl = nG - 1;  // prefetches past the end re-read the last ring (valid addresses, never consumed)

    // ---- pipeline stages: straight-line, every load unconditional or predicated (no branches)
    auto stage_row = [&](uint64_t g) -> uint32_t {  // vids != NULL: ring g is the one-ring of vertex vids[g] in a vertex -> tets CSR
        const uint64_t gg = g < gl ? g : gl;
        return vids ? (uint32_t)vids[gg] : (uint32_t)gg;
    };
    auto stage_head = [&](uint64_t g, uint32_t row) -> RingHead {
        const uint64_t gg = g < gl ? g : gl;
        const uint64_t b = __ldg(off + row), e = __ldg(off + row + 1);
        RingHead h;
        h.b = (uint32_t)b; h.cnt = (uint32_t)(e - b);
        h.c = vids ? (int32_t)row : (ENERGY_ONLY ? 0 : __ldg(center + gg));
        return h;
    };
    auto stage_tid = [&](const RingHead& h, uint32_t k) -> uint32_t {  // member k of the ring -> tet id
        const bool in = k < h.cnt;
        return t_ids ? (in ? (uint32_t)__ldg(t_ids + h.b + k) : 0u) : h.b + (in ? k : 0u);
    };
    auto stage_tet = [&](const RingHead& h, uint32_t k, uint32_t ti) -> int4 {  // a member id outside [0, nT) reads nothing
        int4 t = make_int4(-1, -1, -1, -1);
        if (k < h.cnt && (uint64_t)ti < nT) t = __ldg(tets + ti);
        return t;
    };
    // one member tet: gather, evaluate, add into the lane's partial sums
    // trial position of the centre vertex (twg_mesh_vertex_trial_energy: the line search of VertexSmoother.cpp:505-541 moves the
    // vertex, asks getNewEnergy, and moves it back -- here the mesh stays untouched and the ring sees the vertex at `trial`)
    uint64_t g_cur = 0;
    auto override_centre = [&](double* x, int32_t a0, int32_t a1, int32_t a2, int32_t a3, int32_t c) {
        const double tx = trial[3 * g_cur], ty = trial[3 * g_cur + 1], tz = trial[3 * g_cur + 2];
        const int32_t a[4] = {a0, a1, a2, a3};
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (a[j] == c) { x[3 * j] = tx; x[3 * j + 1] = ty; x[3 * j + 2] = tz; }
    };
    auto member = [&](int4 t, int32_t c, double* acc) {
        int32_t a0 = t.x, a1 = t.y, a2 = t.z, a3 = t.w;
        // removed slots (negative first index) and out-of-range ids contribute nothing and are never dereferenced
        if ((uint32_t)a0 >= nV || (uint32_t)a1 >= nV || (uint32_t)a2 >= nV || (uint32_t)a3 >= nV) {
            atomicAdd(dbg + TWG_DBG_BAD_INDEX, 1ull);
            return;
        }
        if (!ENERGY_ONLY) {
            // :640-651, rotate so that the first slot holding the centre vertex comes first (slot 0 when absent, like the
            // reference); getNewEnergy keeps the stored order. Two conditional register rotations: no branches, no indexing.
            const int start = (a0 == c) ? 0 : (a1 == c) ? 1 : (a2 == c) ? 2 : (a3 == c) ? 3 : 0;
            const bool r1 = (start & 1) != 0, r2 = (start & 2) != 0;
            const int32_t b0 = r1 ? a1 : a0, b1 = r1 ? a2 : a1, b2 = r1 ? a3 : a2, b3 = r1 ? a0 : a3;
            a0 = r2 ? b2 : b0; a1 = r2 ? b3 : b1; a2 = r2 ? b0 : b2; a3 = r2 ? b1 : b3;
        }
        double x[12];
        gather_vertex(V, a0, x);  // the centre: the same address in every lane, one L1 transaction
        gather_vertex(V, a1, x + 3);
        gather_vertex(V, a2, x + 6);
        gather_vertex(V, a3, x + 9);
        if (trial) override_centre(x, a0, a1, a2, a3, c);
        tw::Amips r;
        tw::amips_eval<!ENERGY_ONLY>(x, r);
        acc[0] += r.E;
        if (!ENERGY_ONLY) {
            acc[1] += r.J[0]; acc[2] += r.J[1]; acc[3] += r.J[2];
#pragma unroll
            for (int k = 0; k < 6; ++k) acc[4 + k] += r.H[k];
        }
    };

    // ---- prologue: fill the pipeline for rings i = 0 .. 3 of this warp
    uint64_t g = warp;
    RingHead h_cur = stage_head(g, stage_row(g));
    int4 tet_cur = stage_tet(h_cur, lane, stage_tid(h_cur, lane));
    RingHead h_n1 = stage_head(g + nwarps, stage_row(g + nwarps));
    uint32_t ti_n1 = stage_tid(h_n1, lane);
    RingHead h_n2 = stage_head(g + 2 * nwarps, stage_row(g + 2 * nwarps));
    uint32_t row_n3 = stage_row(g + 3 * nwarps);

    for (; g < nG; g += nwarps) {
        g_cur = g;
        // ---- issue the loads of the later rings first
        const uint32_t row_n4 = stage_row(g + 4 * nwarps);
        const RingHead h_n3 = stage_head(g + 3 * nwarps, row_n3);
        const uint32_t ti_n2 = stage_tid(h_n2, lane);
        const int4 tet_n1 = stage_tet(h_n1, lane, ti_n1);
        // ---- ring g
        double acc[NRED];
#pragma unroll
        for (int k = 0; k < NRED; ++k) acc[k] = 0.0;
        if ((uint32_t)lane < h_cur.cnt) member(tet_cur, h_cur.c, acc);
        for (uint32_t k = 32 + lane; k < h_cur.cnt; k += 32)  // rings of more than 32 tets: the rest is fetched on demand
            member(stage_tet(h_cur, k, stage_tid(h_cur, k)), h_cur.c, acc);
        if (ENERGY_ONLY) {
            double en = warp_sum(acc[0]);
            if (lane == 0) {  // getNewEnergy :619-622
                if (isinf(en) || isnan(en) || en <= 0.0 || en > TWG_MAX_ENERGY) en = TWG_MAX_ENERGY;
                E[g] = en;
            }
        } else {
#pragma unroll
            for (int k = 0; k < NRED; ++k) red[wib][k][lane] = acc[k];
            __syncwarp();
            const int cmp = lane & 15, half = lane >> 4;
            double sum = 0.0;
            if (cmp < NRED) {
#pragma unroll
                for (int j = 0; j < 16; ++j) sum += red[wib][cmp][half * 16 + j];
            }
            sum += __shfl_xor_sync(0xffffffffu, sum, 16);
            __syncwarp();
            // :680-699 -- E: +inf -> MAX_ENERGY, NaN or <= 0 rejects; any non-finite J / H rejects
            bool bad = false;
            if (cmp == 0) {
                if (isinf(sum)) sum = TWG_MAX_ENERGY;
                bad = isnan(sum) || sum <= 0.0;
            } else if (cmp < NRED) {
                bad = !isfinite(sum);
            }
            const bool good = !__any_sync(0xffffffffu, bad);
            if (half == 0) {
                if (cmp == 0) { E[g] = sum; if (ok) ok[g] = good ? 1 : 0; }
                else if (cmp < 4) J3[g * 3 + (cmp - 1)] = sum;
                else if (cmp < NRED) {
                    // H6 = xx xy xz yy yz zz -> row-major 3x3: first slot of every component, mirror of the off-diagonal ones
                    const int s0 = (cmp == 4) ? 0 : (cmp == 5) ? 1 : (cmp == 6) ? 2 : (cmp == 7) ? 4 : (cmp == 8) ? 5 : 8;
                    const int s1 = (cmp == 5) ? 3 : (cmp == 6) ? 6 : (cmp == 8) ? 7 : s0;
                    double* Hg = H9 + g * 9;
                    Hg[s0] = sum;
                    Hg[s1] = sum;
                }
            }
        }
        // ---- advance the pipeline
        h_cur = h_n1; tet_cur = tet_n1;
        h_n1 = h_n2; ti_n1 = ti_n2;
        h_n2 = h_n3;
        row_n3 = row_n4;
    }
}

template <bool ENERGY_ONLY, int MINB>
__global__ void __launch_bounds__(256, MINB) amips_ring_kernel(const double* __restrict__ V, const int4* __restrict__ tets,
                                                            const int32_t* __restrict__ t_ids, const uint64_t* __restrict__ off,
                                                            const int32_t* __restrict__ center, const int32_t* __restrict__ vids, uint64_t nG,
                                                            double* __restrict__ E, double* __restrict__ J3, double* __restrict__ H9,
                                                            uint8_t* __restrict__ ok, uint32_t nV, uint64_t nT, unsigned long long* dbg,
                                                            const double* __restrict__ trial) {
    ring_warp_body<ENERGY_ONLY>(V, tets, t_ids, off, center, vids, nG, E, J3, H9, ok, nV, nT, dbg, trial);
}

// The tiny call of the sequential scheduler (ONE Newton step, the step sizes of ONE line search: twg_mesh_vertex_ring_ejh /
// twg_mesh_vertex_trial_energy with n <= 32): the vertex ids and trial positions travel in the kernel's PARAMETERS (no read
// over PCIe, no copy), the results go straight to the mapped slab and the last CTA raises the completion word itself (no
// second launch): one driver call per host call.
struct TwgTinyRings {
    int32_t vid[32];
    double xyz[96];
};
template <bool ENERGY_ONLY>
__global__ void __launch_bounds__(256, 3) amips_ring_tiny_kernel(const double* __restrict__ V, const int4* __restrict__ tets, const int32_t* __restrict__ adj,
                                                                 const uint64_t* __restrict__ off, const __grid_constant__ TwgTinyRings in, uint32_t n,
                                                                 double* __restrict__ E, double* __restrict__ J3, double* __restrict__ H9,
                                                                 uint8_t* __restrict__ ok, uint32_t nV, uint64_t nT, unsigned long long* dbg, twg_done done) {
    __shared__ int32_t sv[32];
    __shared__ double sx[96];
    if (threadIdx.x < 32) sv[threadIdx.x] = in.vid[threadIdx.x];
    if (ENERGY_ONLY && threadIdx.x < 96) sx[threadIdx.x] = in.xyz[threadIdx.x];
    __syncthreads();
    ring_warp_body<ENERGY_ONLY>(V, tets, adj, off, nullptr, sv, n, E, J3, H9, ok, nV, nT, dbg, ENERGY_ONLY ? sx : nullptr);
    twg_signal_done(done);
}

inline bool aligned16(const void* p) { return (((uintptr_t)p) & 15u) == 0; }

int launch_soa(twg_ctx* c, const double* const dT[12], double* dE, double* dJ3, double* dH9, uint64_t n, cudaStream_t st) {
    if (n == 0) return 0;
    SoaArgs a;
    bool vec = true;  // 128-bit / TMA paths need 16-byte aligned arrays (checked below); odd n is handled by the tail code
    for (int k = 0; k < 12; ++k) { a.T[k] = dT[k]; vec = vec && aligned16(dT[k]); }
    a.E = dE; a.J3 = dJ3; a.H9 = dH9; a.n = n;
    vec = vec && (!dE || aligned16(dE)) && (!dJ3 || aligned16(dJ3)) && (!dH9 || aligned16(dH9));
    const bool jh = (dJ3 != nullptr) || (dH9 != nullptr);
    const bool use_tma = c->opt.amips_tma != 0;
    if (vec && jh && use_tma && n >= 2 * AM_THREADS) {
        // full tiles through the TMA-store kernel, the ragged rest (< 256 tets) through the generic one
        const uint64_t n_tiles = n / (2 * AM_THREADS);
        uint64_t blocks = n_tiles;
        const uint64_t maxb = (uint64_t)c->sm_count * 8;
        if (blocks > maxb) blocks = maxb;
        TWG_LAUNCH(c, amips_soa_tma_kernel, (unsigned)blocks, AM_THREADS, (size_t)2 * AM_THREADS * 12 * sizeof(double), st, a, n_tiles);
        const uint64_t done = n_tiles * 2 * AM_THREADS;
        if (done == n) return 0;
        SoaArgs t = a;
        for (int k = 0; k < 12; ++k) t.T[k] = a.T[k] + done;
        if (a.E) t.E = a.E + done;
        if (a.J3) t.J3 = a.J3 + done * 3;
        if (a.H9) t.H9 = a.H9 + done * 9;
        t.n = n - done;
        TWG_LAUNCH(c, (amips_soa_kernel<2, true>), 1, AM_THREADS, (size_t)2 * AM_THREADS * 12 * sizeof(double), st, t);
        return 0;
    }
    const int tile = AM_THREADS * (vec ? 2 : 1);
    uint64_t blocks = (n + tile - 1) / tile;
    const uint64_t maxb = (uint64_t)c->sm_count * 16;
    if (blocks > maxb) blocks = maxb;
    const size_t smem = jh ? (size_t)tile * 12 * sizeof(double) : 0;
    if (vec) {
        if (jh) TWG_LAUNCH(c, (amips_soa_kernel<2, true>), (unsigned)blocks, AM_THREADS, smem, st, a);
        else TWG_LAUNCH(c, (amips_soa_kernel<2, false>), (unsigned)blocks, AM_THREADS, smem, st, a);
    } else {
        if (jh) TWG_LAUNCH(c, (amips_soa_kernel<1, true>), (unsigned)blocks, AM_THREADS, smem, st, a);
        else TWG_LAUNCH(c, (amips_soa_kernel<1, false>), (unsigned)blocks, AM_THREADS, smem, st, a);
    }
    return 0;
}

cudaStream_t pick(twg_ctx* c, void* stream) { return stream ? (cudaStream_t)stream : c->streams[0]; }

unsigned grid_for(twg_ctx* c, uint64_t items, int per_block, int waves) {
    uint64_t b = (items + per_block - 1) / per_block;
    uint64_t m = (uint64_t)c->sm_count * waves;
    if (b > m) b = m;
    if (b == 0) b = 1;
    return (unsigned)b;
}

// persistent grid of the ring kernels: 3 resident CTAs per SM (80 registers x 256 threads), every warp pipelines over its rings
int ring_waves(const twg_ctx* c) { return c->opt.ring_waves; }

template <bool ENERGY_ONLY>
int launch_ring(twg_ctx* c, cudaStream_t st, const double* dV, uint32_t nV, const int4* dTets, uint64_t nT, const int32_t* dTids, const uint64_t* dOff,
                const int32_t* dCenter, const int32_t* dVids, uint64_t nG, double* dE, double* dJ3, double* dH9, uint8_t* dOk,
                const double* dTrial = nullptr) {
    if (c->opt.ring_minb >= 4) {  // 64 registers, 32 warps per SM
        TWG_LAUNCH(c, (amips_ring_kernel<ENERGY_ONLY, 4>), grid_for(c, nG, 8, ring_waves(c) > 4 ? ring_waves(c) : 4), 256, 0, st, dV, dTets, dTids, dOff, dCenter, dVids,
                   nG, dE, dJ3, dH9, dOk, nV, nT, c->dcounters, dTrial);
        return 0;
    }
    TWG_LAUNCH(c, (amips_ring_kernel<ENERGY_ONLY, 3>), grid_for(c, nG, 8, ring_waves(c)), 256, 0, st, dV, dTets, dTids, dOff, dCenter, dVids, nG, dE, dJ3, dH9, dOk,
               nV, nT, c->dcounters, dTrial);
    return 0;
}

}  // namespace

extern "C" {

int twg_amips_energy_soa_dev(twg_ctx* c, const double* const dT[12], double* dE, uint64_t n, void* stream) {
    TWG_CHECK(c, c && dT && dE, TWG_ERR_INVALID_ARG, "null argument");
    TWG_CHECK(c, !twg_is_multi(c), TWG_ERR_INVALID_ARG, "_dev entry points take a one-device context (twg_device_context)");
    TWG_CUDA(c, cudaSetDevice(c->device));
    return launch_soa(c, dT, dE, nullptr, nullptr, n, pick(c, stream));
}

int twg_amips_ejh_soa_dev(twg_ctx* c, const double* const dT[12], double* dE, double* dJ3, double* dH9, uint64_t n, void* stream) {
    TWG_CHECK(c, c && dT, TWG_ERR_INVALID_ARG, "null argument");
    TWG_CHECK(c, !twg_is_multi(c), TWG_ERR_INVALID_ARG, "_dev entry points take a one-device context (twg_device_context)");
    TWG_CUDA(c, cudaSetDevice(c->device));
    return launch_soa(c, dT, dE, dJ3, dH9, n, pick(c, stream));
}

int twg_amips_quality_dev(twg_ctx* c, const double* dV, uint32_t nV, const int32_t* dTets, uint64_t nT, double* dSlim, void* stream) {
    TWG_CHECK(c, c && dV && dTets && dSlim, TWG_ERR_INVALID_ARG, "null argument");
    TWG_CHECK(c, aligned16(dTets), TWG_ERR_ALIGNMENT, "tets4 must be 16-byte aligned");
    if (nT == 0) return 0;
    TWG_CHECK(c, !twg_is_multi(c), TWG_ERR_INVALID_ARG, "_dev entry points take a one-device context (twg_device_context)");
    TWG_CUDA(c, cudaSetDevice(c->device));
    TWG_LAUNCH(c, amips_quality_kernel, grid_for(c, nT, 256, 8), 256, 0, pick(c, stream), dV, nV, (const int4*)dTets, nT, dSlim, c->dcounters);
    return 0;
}

int twg_amips_ring_ejh_dev(twg_ctx* c, const double* dV, uint32_t nV, const int32_t* dTets, uint64_t nT, const int32_t* dTids,
                           const uint64_t* dOff, const int32_t* dCenter, uint64_t nG, double* dE, double* dJ3, double* dH9,
                           uint8_t* dOk, void* stream) {
    TWG_CHECK(c, c && dV && dTets && dOff && dCenter && dE && dJ3 && dH9, TWG_ERR_INVALID_ARG, "null argument");
    TWG_CHECK(c, aligned16(dTets), TWG_ERR_ALIGNMENT, "tets4 must be 16-byte aligned");
    if (nG == 0) return 0;
    TWG_CHECK(c, !twg_is_multi(c), TWG_ERR_INVALID_ARG, "_dev entry points take a one-device context (twg_device_context)");
    TWG_CUDA(c, cudaSetDevice(c->device));
    return launch_ring<false>(c, pick(c, stream), dV, nV, (const int4*)dTets, nT, dTids, dOff, dCenter, (const int32_t*)nullptr, nG, dE, dJ3, dH9, dOk);
}

// one-rings named by their centre vertex: members of ring g are adj_tets[adj_off[v] .. adj_off[v+1]) with v = dVids[g]
int twg_amips_vertex_ring_ejh_dev(twg_ctx* c, const double* dV, uint32_t nV, const int32_t* dTets, uint64_t nT, const int32_t* dAdjTets,
                                  const uint64_t* dAdjOff, const int32_t* dVids, uint64_t nG, double* dE, double* dJ3, double* dH9, uint8_t* dOk,
                                  void* stream) {
    TWG_CHECK(c, c && dV && dTets && dAdjTets && dAdjOff && dVids && dE && dJ3 && dH9, TWG_ERR_INVALID_ARG, "null argument");
    TWG_CHECK(c, aligned16(dTets), TWG_ERR_ALIGNMENT, "tets4 must be 16-byte aligned");
    if (nG == 0) return 0;
    TWG_CHECK(c, !twg_is_multi(c), TWG_ERR_INVALID_ARG, "_dev entry points take a one-device context (twg_device_context)");
    TWG_CUDA(c, cudaSetDevice(c->device));
    return launch_ring<false>(c, pick(c, stream), dV, nV, (const int4*)dTets, nT, dAdjTets, dAdjOff, (const int32_t*)nullptr, dVids, nG, dE, dJ3, dH9, dOk);
}

// n <= 32 one-rings of the resident mesh named by HOST vertex ids; trial_xyz != NULL: getNewEnergy with the vertex at the trial
// position (energies only), else NewtonsUpdate's E / J / H / ok. Output pointers are device-visible (the mapped slab).
int twg_amips_ring_tiny(twg_ctx* c, const double* dV, uint32_t nV, const int32_t* dTets, uint64_t nT, const int32_t* dAdjTets, const uint64_t* dAdjOff,
                        const int32_t* v_ids, const double* trial_xyz, uint32_t n, double* E, double* J3, double* H9, uint8_t* ok, cudaStream_t st,
                        const twg_done* done) {
    TWG_CHECK(c, c && dV && dTets && dAdjTets && dAdjOff && v_ids && E && done && n >= 1 && n <= 32, TWG_ERR_INVALID_ARG, "bad tiny ring call");
    TWG_CHECK(c, aligned16(dTets), TWG_ERR_ALIGNMENT, "tets4 must be 16-byte aligned");
    TwgTinyRings in;
    memcpy(in.vid, v_ids, (size_t)n * 4);
    for (uint32_t k = n; k < 32; ++k) in.vid[k] = v_ids[0];
    const unsigned blocks = (n + 7) / 8;
    if (trial_xyz) {
        memcpy(in.xyz, trial_xyz, (size_t)n * 24);
        TWG_LAUNCH(c, (amips_ring_tiny_kernel<true>), blocks, 256, 0, st, dV, (const int4*)dTets, dAdjTets, dAdjOff, in, n, E, (double*)nullptr, (double*)nullptr,
                   (uint8_t*)nullptr, nV, nT, c->dcounters, *done);
    } else {
        TWG_CHECK(c, J3 && H9, TWG_ERR_INVALID_ARG, "null argument");
        TWG_LAUNCH(c, (amips_ring_tiny_kernel<false>), blocks, 256, 0, st, dV, (const int4*)dTets, dAdjTets, dAdjOff, in, n, E, J3, H9, ok, nV, nT, c->dcounters, *done);
    }
    return 0;
}

// getNewEnergy of the one-ring of vertex dVids[g] with that vertex at dTrial[3g..3g+2] (the mesh itself is not modified)
int twg_amips_vertex_trial_energy_dev(twg_ctx* c, const double* dV, uint32_t nV, const int32_t* dTets, uint64_t nT, const int32_t* dAdjTets,
                                      const uint64_t* dAdjOff, const int32_t* dVids, const double* dTrial, uint64_t nG, double* dE, void* stream) {
    TWG_CHECK(c, c && dV && dTets && dAdjTets && dAdjOff && dVids && dTrial && dE, TWG_ERR_INVALID_ARG, "null argument");
    TWG_CHECK(c, aligned16(dTets), TWG_ERR_ALIGNMENT, "tets4 must be 16-byte aligned");
    TWG_CHECK(c, !twg_is_multi(c), TWG_ERR_INVALID_ARG, "_dev entry points take a one-device context (twg_device_context)");
    if (nG == 0) return 0;
    TWG_CUDA(c, cudaSetDevice(c->device));
    return launch_ring<true>(c, pick(c, stream), dV, nV, (const int4*)dTets, nT, dAdjTets, dAdjOff, (const int32_t*)nullptr, dVids, nG, dE, (double*)nullptr,
                             (double*)nullptr, (uint8_t*)nullptr, dTrial);
}

int twg_amips_ring_energy_dev(twg_ctx* c, const double* dV, uint32_t nV, const int32_t* dTets, uint64_t nT, const int32_t* dTids,
                              const uint64_t* dOff, uint64_t nG, double* dE, void* stream) {
    TWG_CHECK(c, c && dV && dTets && dOff && dE, TWG_ERR_INVALID_ARG, "null argument");
    TWG_CHECK(c, aligned16(dTets), TWG_ERR_ALIGNMENT, "tets4 must be 16-byte aligned");
    if (nG == 0) return 0;
    TWG_CHECK(c, !twg_is_multi(c), TWG_ERR_INVALID_ARG, "_dev entry points take a one-device context (twg_device_context)");
    TWG_CUDA(c, cudaSetDevice(c->device));
    return launch_ring<true>(c, pick(c, stream), dV, nV, (const int4*)dTets, nT, dTids, dOff, (const int32_t*)nullptr, (const int32_t*)nullptr, nG, dE,
                             (double*)nullptr, (double*)nullptr, (uint8_t*)nullptr);
}

// ---- host-buffer entry points: chunked H2D -> kernel -> D2H pipeline over TWG_NUM_STREAMS streams ----
int twg_amips_ejh_soa(twg_ctx* c, const double* const T[12], double* E, double* J3, double* H9, uint64_t n) {
    TWG_CHECK(c, c && T, TWG_ERR_INVALID_ARG, "null argument");
    if (n == 0) return 0;
    if (twg_is_multi(c)) {  // flat batch: contiguous index ranges, one per device
        if (n < TWG_MULTI_MIN_TETS) return twg_forward0(c, twg_amips_ejh_soa(c->children[0], T, E, J3, H9, n));
        const uint64_t G = c->children.size();
        return twg_multi_run(c, [&](int k, twg_ctx* child) {
            const uint64_t b = (n * (uint64_t)k / G) & ~1ull, e = (k + 1 == (int)G) ? n : ((n * (uint64_t)(k + 1) / G) & ~1ull);  // even starts keep 16-byte alignment
            const double* Tk[12];
            for (int a = 0; a < 12; ++a) Tk[a] = T[a] + b;
            return twg_amips_ejh_soa(child, Tk, E ? E + b : nullptr, J3 ? J3 + 3 * b : nullptr, H9 ? H9 + 9 * b : nullptr, e - b);
        });
    }
    TWG_CUDA(c, cudaSetDevice(c->device));
    const uint64_t chunk = 1ull << 21;  // 2 Mi tets: 192 MiB in, up to 208 MiB out per slot
    const int nout = (E ? 1 : 0) + (J3 ? 3 : 0) + (H9 ? 9 : 0);
    const size_t per = (size_t)(12 + nout) * sizeof(double);
    const uint64_t cmax = n < chunk ? ((n + 1) & ~1ull) : chunk;
    for (int s = 0; s < TWG_NUM_STREAMS; ++s) TWG_TRY(twg_ensure_scratch(c, s, per * cmax));
    int slot = 0;
    for (uint64_t b = 0; b < n; b += chunk, slot = (slot + 1) % TWG_NUM_STREAMS) {
        const uint64_t m = (n - b < chunk) ? (n - b) : chunk;
        cudaStream_t st = c->streams[slot];
        double* base = (double*)c->dscratch[slot];
        const double* dT[12];
        for (int k = 0; k < 12; ++k) {
            dT[k] = base + (size_t)k * cmax;
            TWG_CUDA(c, cudaMemcpyAsync((void*)dT[k], T[k] + b, m * sizeof(double), cudaMemcpyHostToDevice, st));
        }
        double* dE = E ? base + (size_t)12 * cmax : nullptr;
        double* dJ = J3 ? base + (size_t)(12 + (E ? 1 : 0)) * cmax : nullptr;
        double* dH = H9 ? base + (size_t)(12 + (E ? 1 : 0) + (J3 ? 3 : 0)) * cmax : nullptr;
        TWG_TRY(launch_soa(c, dT, dE, dJ, dH, m, st));
        if (E) TWG_CUDA(c, cudaMemcpyAsync(E + b, dE, m * sizeof(double), cudaMemcpyDeviceToHost, st));
        if (J3) TWG_CUDA(c, cudaMemcpyAsync(J3 + b * 3, dJ, m * 3 * sizeof(double), cudaMemcpyDeviceToHost, st));
        if (H9) TWG_CUDA(c, cudaMemcpyAsync(H9 + b * 9, dH, m * 9 * sizeof(double), cudaMemcpyDeviceToHost, st));
    }
    for (int s = 0; s < TWG_NUM_STREAMS; ++s) TWG_CUDA(c, cudaStreamSynchronize(c->streams[s]));
    return 0;
}

int twg_amips_energy_soa(twg_ctx* c, const double* const T[12], double* E, uint64_t n) {
    TWG_CHECK(c, c && T && E, TWG_ERR_INVALID_ARG, "null argument");
    return twg_amips_ejh_soa(c, T, E, nullptr, nullptr, n);
}

static int begin_checked(twg_ctx* c, cudaStream_t st);
static int finish_checked(twg_ctx* c, cudaStream_t st);

int twg_amips_quality(twg_ctx* c, const double* V, uint32_t nV, const int32_t* tets, uint64_t nT, double* slim) {
    TWG_CHECK(c, c && V && tets && slim, TWG_ERR_INVALID_ARG, "null argument");
    if (nT == 0) return 0;
    if (twg_is_multi(c)) {  // vertices replicated, tets split by index range
        if (nT < TWG_MULTI_MIN_TETS) return twg_forward0(c, twg_amips_quality(c->children[0], V, nV, tets, nT, slim));
        const uint64_t G = c->children.size();
        return twg_multi_run(c, [&](int k, twg_ctx* child) {
            const uint64_t b = nT * (uint64_t)k / G, e = nT * (uint64_t)(k + 1) / G;
            return twg_amips_quality(child, V, nV, tets + 4 * b, e - b, slim + b);
        });
    }
    TWG_CUDA(c, cudaSetDevice(c->device));
    const size_t vb = ((size_t)nV * 3 * sizeof(double) + 255) & ~(size_t)255;
    const size_t tb = ((size_t)nT * 16 + 255) & ~(size_t)255;
    TWG_TRY(twg_ensure_scratch(c, 0, vb + tb + nT * sizeof(double)));
    char* base = (char*)c->dscratch[0];
    cudaStream_t st = c->streams[0];
    TWG_TRY(begin_checked(c, st));
    TWG_CUDA(c, cudaMemcpyAsync(base, V, (size_t)nV * 3 * sizeof(double), cudaMemcpyHostToDevice, st));
    TWG_CUDA(c, cudaMemcpyAsync(base + vb, tets, (size_t)nT * 16, cudaMemcpyHostToDevice, st));
    TWG_TRY(twg_amips_quality_dev(c, (const double*)base, nV, (const int32_t*)(base + vb), nT, (double*)(base + vb + tb), st));
    TWG_CUDA(c, cudaMemcpyAsync(slim, base + vb + tb, nT * sizeof(double), cudaMemcpyDeviceToHost, st));
    return finish_checked(c, st);
}

// The kernels never dereference an out-of-range index; they count them. The host entry points clear the counter before
// their launch, read it with the results and report TWG_ERR_INVALID_ARG (results of the offending tets / rings are MAX_ENERGY / partial sums).
static int begin_checked(twg_ctx* c, cudaStream_t st) {
    TWG_CUDA(c, cudaMemsetAsync(c->dcounters + TWG_DBG_BAD_INDEX, 0, sizeof(unsigned long long), st));
    return 0;
}
static int finish_checked(twg_ctx* c, cudaStream_t st) {
    unsigned long long bad = 0;
    TWG_CUDA(c, cudaMemcpyAsync(&bad, c->dcounters + TWG_DBG_BAD_INDEX, sizeof(bad), cudaMemcpyDeviceToHost, st));
    TWG_CUDA(c, cudaStreamSynchronize(st));
    TWG_CHECK(c, bad == 0, TWG_ERR_INVALID_ARG, "a tet or ring references a vertex / tet out of range");
    return 0;
}

static int ring_host(twg_ctx* c, bool energy_only, const double* V, uint32_t nV, const int32_t* tets, uint64_t nT, const int32_t* t_ids,
                     const uint64_t* off, const int32_t* center, uint64_t nG, double* E, double* J3, double* H9, uint8_t* ok) {
    if (nG == 0) return 0;
    if (twg_is_multi(c)) return twg_forward0(c, ring_host(c->children[0], energy_only, V, nV, tets, nT, t_ids, off, center, nG, E, J3, H9, ok));
    TWG_CUDA(c, cudaSetDevice(c->device));
    const uint64_t nM = off[nG];
    TWG_CHECK(c, nM < (1ull << 32) && off[0] <= nM, TWG_ERR_INVALID_ARG, "group_off must be non-decreasing and below 2^32");
    for (uint64_t g = 0; g < nG; ++g) TWG_CHECK(c, off[g] <= off[g + 1], TWG_ERR_INVALID_ARG, "group_off must be non-decreasing");
    if (!t_ids) TWG_CHECK(c, nM <= nT, TWG_ERR_INVALID_ARG, "group_off runs past the tet array");
    auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t vb = up((size_t)nV * 24), tb = up((size_t)nT * 16), ib = t_ids ? up((size_t)nM * 4) : 0, ob = up((size_t)(nG + 1) * 8),
                 cb = up((size_t)nG * 4), eb = up((size_t)nG * 8), jb = up((size_t)nG * 24), hb = up((size_t)nG * 72), kb = up((size_t)nG);
    TWG_TRY(twg_ensure_scratch(c, 0, vb + tb + ib + ob + cb + eb + jb + hb + kb));
    char* p = (char*)c->dscratch[0];
    cudaStream_t st = c->streams[0];
    char* dV = p; p += vb;
    char* dT = p; p += tb;
    char* dI = t_ids ? p : nullptr; p += ib;
    char* dO = p; p += ob;
    char* dC = p; p += cb;
    char* dE = p; p += eb;
    char* dJ = p; p += jb;
    char* dH = p; p += hb;
    char* dK = p;
    TWG_TRY(begin_checked(c, st));
    TWG_CUDA(c, cudaMemcpyAsync(dV, V, (size_t)nV * 24, cudaMemcpyHostToDevice, st));
    TWG_CUDA(c, cudaMemcpyAsync(dT, tets, (size_t)nT * 16, cudaMemcpyHostToDevice, st));
    if (t_ids) TWG_CUDA(c, cudaMemcpyAsync(dI, t_ids, (size_t)nM * 4, cudaMemcpyHostToDevice, st));
    TWG_CUDA(c, cudaMemcpyAsync(dO, off, (size_t)(nG + 1) * 8, cudaMemcpyHostToDevice, st));
    if (!energy_only) TWG_CUDA(c, cudaMemcpyAsync(dC, center, (size_t)nG * 4, cudaMemcpyHostToDevice, st));
    if (energy_only) {
        TWG_TRY(twg_amips_ring_energy_dev(c, (const double*)dV, nV, (const int32_t*)dT, nT, (const int32_t*)dI, (const uint64_t*)dO, nG,
                                          (double*)dE, st));
    } else {
        TWG_TRY(twg_amips_ring_ejh_dev(c, (const double*)dV, nV, (const int32_t*)dT, nT, (const int32_t*)dI, (const uint64_t*)dO,
                                       (const int32_t*)dC, nG, (double*)dE, (double*)dJ, (double*)dH, (uint8_t*)dK, st));
    }
    TWG_CUDA(c, cudaMemcpyAsync(E, dE, (size_t)nG * 8, cudaMemcpyDeviceToHost, st));
    if (!energy_only) {
        TWG_CUDA(c, cudaMemcpyAsync(J3, dJ, (size_t)nG * 24, cudaMemcpyDeviceToHost, st));
        TWG_CUDA(c, cudaMemcpyAsync(H9, dH, (size_t)nG * 72, cudaMemcpyDeviceToHost, st));
        if (ok) TWG_CUDA(c, cudaMemcpyAsync(ok, dK, (size_t)nG, cudaMemcpyDeviceToHost, st));
    }
    return finish_checked(c, st);
}

int twg_amips_ring_ejh(twg_ctx* c, const double* V, uint32_t nV, const int32_t* tets, uint64_t nT, const int32_t* t_ids,
                       const uint64_t* off, const int32_t* center, uint64_t nG, double* E, double* J3, double* H9, uint8_t* ok) {
    TWG_CHECK(c, c && V && tets && off && center && E && J3 && H9, TWG_ERR_INVALID_ARG, "null argument");
    return ring_host(c, false, V, nV, tets, nT, t_ids, off, center, nG, E, J3, H9, ok);
}

int twg_amips_ring_energy(twg_ctx* c, const double* V, uint32_t nV, const int32_t* tets, uint64_t nT, const int32_t* t_ids,
                          const uint64_t* off, uint64_t nG, double* E) {
    TWG_CHECK(c, c && V && tets && off && E, TWG_ERR_INVALID_ARG, "null argument");
    return ring_host(c, true, V, nV, tets, nT, t_ids, off, nullptr, nG, E, nullptr, nullptr, nullptr);
}

}  // extern "C"
