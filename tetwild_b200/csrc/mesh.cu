// mesh.cu -- device-resident mirror of the scheduler's tet mesh (C ABI: include/tetwild_gpu.h, "resident tet mesh").
//
// The reference keeps `tet_vertices[].posf`, `tets` and `tet_vertices[].conn_tets` in host vectors that every local
// operation reads in place (src/tetwild/LocalOperations.h:35-45). Shipping them with each AMIPS call costs ~90 B per tet
// over PCIe, 20x what the kernel needs to run; twg_mesh keeps them in HBM instead:
//   V     [capV*3] f64   posf, updated by scatter after accepted operations (twg_mesh_set_vertices)
//   T     [capT]   int4  tets; a removed tet (t_is_removed) is marked by a negative first index (twg_mesh_set_tets)
//   adj   CSR            conn_tets: vertex -> incident live tets in ascending tet id, built on the device
//                        (key = vertex, value = tet, cub radix sort, offsets by binary search)
// so that a whole-mesh quality pass moves 8 B per tet and a batch of one-ring Newton evaluations moves 4 B in and 105 B
// out per RING. Kernels are the ones of amips.cu (same arithmetic, same gates) plus the dihedral-angle pass
// calTetQuality_AD (LocalOperations.cpp:783-860), which LocalOperations::outputInfo (:348-354) runs over all live tets
// after every operation.
#include <cub/device/device_radix_sort.cuh>
#include "common.cuh"

extern "C" int twg_amips_vertex_ring_ejh_dev(twg_ctx* c, const double* dV, uint32_t nV, const int32_t* dTets, uint64_t nT, const int32_t* dAdjTets,
                                             const uint64_t* dAdjOff, const int32_t* dVids, uint64_t nG, double* dE, double* dJ3,
                                             double* dH9, uint8_t* dOk, void* stream);

// tiny-call variant (amips.cu): n <= 32 rings named by HOST vertex ids (and host trial positions -> energies only), results
// into the mapped slab, completion through `done`
extern "C" int twg_amips_ring_tiny(twg_ctx* c, const double* dV, uint32_t nV, const int32_t* dTets, uint64_t nT, const int32_t* dAdjTets, const uint64_t* dAdjOff,
                                   const int32_t* v_ids, const double* trial_xyz, uint32_t n, double* E, double* J3, double* H9, uint8_t* ok, cudaStream_t st,
                                   const twg_done* done);

extern "C" int twg_amips_vertex_trial_energy_dev(twg_ctx* c, const double* dV, uint32_t nV, const int32_t* dTets, uint64_t nT, const int32_t* dAdjTets,
                                                 const uint64_t* dAdjOff, const int32_t* dVids, const double* dTrial, uint64_t nG, double* dE, void* stream);

struct twg_mesh {
    twg_ctx* ctx = nullptr;
    uint32_t nV = 0, capV = 0;
    uint64_t nT = 0, capT = 0;
    double* V = nullptr;
    int4* T = nullptr;
    int32_t* adj_tets = nullptr;   // 4*nT entries (live ones first)
    uint64_t* adj_off = nullptr;   // nV+1
    uint64_t adj_cap = 0, adj_off_cap = 0;
    bool adj_valid = false;
};

namespace {

constexpr double kPi = 3.14159265358979323846;  // M_PI

__global__ void __launch_bounds__(256) scatter_vertices_kernel(double* __restrict__ V, const int32_t* __restrict__ ids, const double* __restrict__ xyz,
                                                               uint64_t n, uint32_t nV) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t v = ids[i];
    if (v < 0 || (uint32_t)v >= nV) return;  // validated on the host; never write out of bounds
    V[3 * (size_t)v] = xyz[3 * i]; V[3 * (size_t)v + 1] = xyz[3 * i + 1]; V[3 * (size_t)v + 2] = xyz[3 * i + 2];
}
__global__ void __launch_bounds__(256) scatter_tets_kernel(int4* __restrict__ T, const int32_t* __restrict__ ids, const int4* __restrict__ tets, uint64_t n,
                                                           uint64_t nT) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t t = ids[i];
    if (t < 0 || (uint64_t)t >= nT) return;
    T[t] = tets[i];
}

// conn_tets: (vertex, tet) pairs of the live tets; removed tets sort to the end under key 0xffffffff
__global__ void __launch_bounds__(256) adj_pairs_kernel(const int4* __restrict__ T, uint64_t nT, uint32_t* __restrict__ keys, int32_t* __restrict__ vals) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nT) return;
    const int4 t = T[i];
    const bool live = t.x >= 0;
    keys[4 * i] = live ? (uint32_t)t.x : 0xffffffffu; keys[4 * i + 1] = live ? (uint32_t)t.y : 0xffffffffu;
    keys[4 * i + 2] = live ? (uint32_t)t.z : 0xffffffffu; keys[4 * i + 3] = live ? (uint32_t)t.w : 0xffffffffu;
    vals[4 * i] = vals[4 * i + 1] = vals[4 * i + 2] = vals[4 * i + 3] = (int32_t)i;
}
__global__ void __launch_bounds__(256) adj_offsets_kernel(const uint32_t* __restrict__ sorted_keys, uint64_t m, uint32_t nV, uint64_t* __restrict__ off) {
    const uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v > nV) return;
    uint64_t lo = 0, hi = m;  // first position with key >= v
    while (lo < hi) {
        const uint64_t mid = (lo + hi) >> 1;
        if ((uint64_t)sorted_keys[mid] < v) lo = mid + 1; else hi = mid;
    }
    off[v] = lo;
}

__device__ __forceinline__ void load_tet(const double* __restrict__ V, int4 t, double* x) {
    const int32_t v[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const double* p = V + 3 * (size_t)v[j];
        x[3 * j] = __ldg(p); x[3 * j + 1] = __ldg(p + 1); x[3 * j + 2] = __ldg(p + 2);
    }
}

// calTetQuality_AMIPS over the resident mesh (LocalOperations.cpp:862-884); t_ids == NULL: tets 0..n-1.
// A removed tet (negative first index) gives MAX_ENERGY: the reference never evaluates those (t_is_removed).
// a 24-byte vertex with one 16-byte and one 8-byte load (whichever half is 16-byte aligned), never reading outside the vertex:
// two L1 wavefronts per lane instead of three (ncu r02: the quality pass keeps L1 77 % busy at 24 % issue utilisation)
__device__ __forceinline__ void load_tet_wide(const double* __restrict__ V, int4 t, double* x) {
    const int32_t v[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const double* p = V + 3 * (size_t)v[j];
        const bool odd = (reinterpret_cast<uintptr_t>(p) & 8u) != 0;
        const double2 a = __ldg(reinterpret_cast<const double2*>(p + (odd ? 1 : 0)));
        const double b = __ldg(p + (odd ? 0 : 2));
        x[3 * j] = odd ? b : a.x; x[3 * j + 1] = odd ? a.x : a.y; x[3 * j + 2] = odd ? a.y : b;
    }
}
template <bool WIDE = false>
__device__ __forceinline__ double mesh_quality_one(const double* __restrict__ V, const int4* __restrict__ T, uint64_t ti) {
    const int4 t = __ldg(T + ti);
    double e = TWG_MAX_ENERGY;
    if (t.x >= 0) {
        double x[12];
        if (WIDE) load_tet_wide(V, t, x); else load_tet(V, t, x);
        if (tw::exact::cgal_orientation(x, x + 3, x + 6, x + 9) == 1) {
            tw::Amips r;
            tw::amips_eval<false>(x, r);
            e = r.E;
        }
        if (isinf(e) || isnan(e) || e <= 0.0) e = TWG_MAX_ENERGY;
    }
    return e;
}
template <bool WIDE>
__global__ void __launch_bounds__(256, 3) mesh_quality_kernel(const double* __restrict__ V, const int4* __restrict__ T, const int32_t* __restrict__ t_ids,
                                                           uint64_t n, double* __restrict__ slim) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        slim[i] = mesh_quality_one<WIDE>(V, T, t_ids ? (uint64_t)__ldg(t_ids + i) : i);
}
// the tets of ONE local operation (n <= 256): ids in the kernel parameters, results straight into the mapped slab, completion
// word raised by the kernel itself -- one driver call per host call
struct TwgTinyTets {
    int32_t id[256];
};
__global__ void __launch_bounds__(256) mesh_quality_tiny_kernel(const double* __restrict__ V, const int4* __restrict__ T, const __grid_constant__ TwgTinyTets in,
                                                                uint32_t n, double* __restrict__ slim, twg_done done) {
    if (threadIdx.x < n) slim[threadIdx.x] = mesh_quality_one(V, T, (uint64_t)in.id[threadIdx.x]);
    twg_signal_done(done);
}

// calTetQuality_AD (LocalOperations.cpp:783-860). The plane through the three other vertices and the projection onto
// it are CGAL constructions on Cartesian<double> (Plane_3(p,q,r) -> plane_from_pointsC3, Plane_3::projection ->
// projection_planeC3; CGAL is not vendored, restated from its published kernel_ftC3.h): plain IEEE double, evaluated
// here without contraction so that the arguments of acos are the doubles an x86-64 build without FMA produces.
__device__ __forceinline__ bool tet_dihedral(const double* x, double& amin, double& amax) {
    using tw::V3; using tw::mk; using tw::vdot; using tw::vlen2;
    using tw::dadd; using tw::dsub; using tw::dmul; using tw::ddiv; using tw::dsqrt;  // hide CUDA's ::dadd(a, b, mode) family
    V3 nv[4];
    double len[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const double* P = x + 3 * ((i + 1) & 3);
        const double* Q = x + 3 * ((i + 2) & 3);
        const double* R = x + 3 * ((i + 3) & 3);
        const double* A = x + 3 * i;
        const double rpx = dsub(P[0], R[0]), rpy = dsub(P[1], R[1]), rpz = dsub(P[2], R[2]);
        const double rqx = dsub(Q[0], R[0]), rqy = dsub(Q[1], R[1]), rqz = dsub(Q[2], R[2]);
        const double pa = dsub(dmul(rpy, rqz), dmul(rqy, rpz));
        const double pb = dsub(dmul(rpz, rqx), dmul(rqz, rpx));
        const double pc = dsub(dmul(rpx, rqy), dmul(rqx, rpy));
        const double pd = dsub(dsub(dmul(-pa, R[0]), dmul(pb, R[1])), dmul(pc, R[2]));
        if (pa == 0.0 && pb == 0.0 && pc == 0.0) return false;  // pln.is_degenerate() :790
        const double num = dadd(dadd(dadd(dmul(pa, A[0]), dmul(pb, A[1])), dmul(pc, A[2])), pd);
        const double den = dadd(dadd(dmul(pa, pa), dmul(pb, pb)), dmul(pc, pc));
        const double lambda = ddiv(num, den);
        const double tx = dsub(A[0], dmul(lambda, pa)), ty = dsub(A[1], dmul(lambda, pb)), tz = dsub(A[2], dmul(lambda, pc));
        if (tx == A[0] && ty == A[1] && tz == A[2]) return false;  // :796
        nv[i] = mk(dsub(A[0], tx), dsub(A[1], ty), dsub(A[2], tz));
        const double h = vlen2(nv[i]);  // CGAL::squared_distance(posf, tmp_p)
        const double m = fmax(fmax(fabs(nv[i].x), fabs(nv[i].y)), fabs(nv[i].z));
        if (m == 0.0 || h == 0.0) return false;  // :820
        if (m < 1e-5) {                          // :826-828
            nv[i] = mk(ddiv(nv[i].x, m), ddiv(nv[i].y, m), ddiv(nv[i].z, m));
            len[i] = dsqrt(ddiv(h, dmul(m, m)));
        } else {
            len[i] = dsqrt(h);
        }
    }
    const int ea[6] = {0, 1, 0, 2, 0, 3}, eb[6] = {1, 2, 2, 3, 3, 1};  // opp_edges :834-838
    amin = DBL_MAX; amax = -DBL_MAX;
    bool first = true;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        const V3 a = mk(-nv[ea[k]].x, -nv[ea[k]].y, -nv[ea[k]].z);
        const double c = ddiv(vdot(a, nv[eb[k]]), dmul(len[ea[k]], len[eb[k]]));
        const double ang = (c > 1.0) ? 0.0 : ((c < -1.0) ? kPi : acos(c));
        // std::minmax_element: first smallest, last largest under operator<
        if (first) { amin = amax = ang; first = false; }
        else {
            if (ang < amin) amin = ang;
            if (!(ang < amax)) amax = ang;
        }
    }
    return true;
}

__global__ void __launch_bounds__(256) mesh_dihedral_kernel(const double* __restrict__ V, const int4* __restrict__ T, const int32_t* __restrict__ t_ids,
                                                            uint64_t n, double* __restrict__ dmin, double* __restrict__ dmax) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const int4 t = __ldg(T + (t_ids ? (uint64_t)__ldg(t_ids + i) : i));
        double a = 0.0, b = kPi;  // the degenerate answers of :791-792, :797-798, :821-822
        if (t.x >= 0) {
            double x[12];
            load_tet(V, t, x);
            double lo, hi;
            if (tet_dihedral(x, lo, hi)) { a = lo; b = hi; }
        }
        dmin[i] = a;
        dmax[i] = b;
    }
}

unsigned grid_for(twg_ctx* c, uint64_t items, int per_block, int waves) {
    uint64_t b = (items + per_block - 1) / per_block;
    const uint64_t m = (uint64_t)c->sm_count * waves;
    if (b > m) b = m;
    if (b == 0) b = 1;
    return (unsigned)b;
}
size_t up256(size_t x) { return (x + 255) & ~(size_t)255; }
constexpr uint64_t kSmallCall = 8192;  // calls up to this many units take the packed, pinned, one-copy-each-way path
constexpr uint64_t kFastRings = 32;    // ... and these few go through the zero-copy slab (twg_fast_*): what ONE un-batched call site asks for
constexpr uint64_t kFastTets = 256;

int grow(twg_mesh* m, uint32_t nV, uint64_t nT) {
    twg_ctx* c = m->ctx;
    cudaStream_t st = c->streams[0];
    if (nV > m->capV) {
        const uint32_t cap = nV + nV / 4 + 64;
        double* p = nullptr;
        TWG_CUDA(c, cudaMalloc(&p, (size_t)cap * 24));
        if (m->V) {
            TWG_CUDA(c, cudaMemcpyAsync(p, m->V, (size_t)m->nV * 24, cudaMemcpyDeviceToDevice, st));
            TWG_CUDA(c, cudaStreamSynchronize(st));
            TWG_CUDA(c, cudaFree(m->V));
        }
        m->V = p; m->capV = cap;
    }
    if (nT > m->capT) {
        const uint64_t cap = nT + nT / 4 + 64;
        int4* p = nullptr;
        TWG_CUDA(c, cudaMalloc(&p, (size_t)cap * 16));
        if (m->T) {
            TWG_CUDA(c, cudaMemcpyAsync(p, m->T, (size_t)m->nT * 16, cudaMemcpyDeviceToDevice, st));
            TWG_CUDA(c, cudaStreamSynchronize(st));
            TWG_CUDA(c, cudaFree(m->T));
        }
        m->T = p; m->capT = cap;
    }
    if (nV > m->nV) TWG_CUDA(c, cudaMemsetAsync(m->V + 3 * (size_t)m->nV, 0, (size_t)(nV - m->nV) * 24, st));
    if (nT > m->nT) TWG_CUDA(c, cudaMemsetAsync(m->T + m->nT, 0xff, (size_t)(nT - m->nT) * 16, st));  // new slots start as removed tets
    m->nV = nV; m->nT = nT;
    m->adj_valid = false;
    return 0;
}

// stage a small host array into the context's scratch slot 0 (after `offset` bytes); returns the device address
template <typename T>
int stage(twg_ctx* c, size_t& offset, const T* host, size_t count, T** dev) {
    char* base = (char*)c->dscratch[0];
    *dev = (T*)(base + offset);
    if (host && count) TWG_CUDA(c, cudaMemcpyAsync(*dev, host, count * sizeof(T), cudaMemcpyHostToDevice, c->streams[0]));
    offset += up256(count * sizeof(T));
    return 0;
}

}  // namespace

extern "C" {

int twg_mesh_create(twg_ctx* c, const double* V, uint32_t nV, const int32_t* tets4, uint64_t nT, twg_mesh** out) {
    TWG_CHECK(c, c && out && (V || nV == 0) && (tets4 || nT == 0), TWG_ERR_INVALID_ARG, "null argument");
    // the resident mesh lives on ONE device (the scheduler that mutates it is sequential): device 0 of a multi-device context
    if (twg_is_multi(c)) return twg_forward0(c, twg_mesh_create(c->children[0], V, nV, tets4, nT, out));
    TWG_CHECK(c, nT < (1ull << 29), TWG_ERR_INVALID_ARG, "at most 2^29 tets");
    for (uint64_t t = 0; t < nT; ++t)
        if (tets4[4 * t] >= 0)  // a negative first index marks a removed slot
            for (int k = 0; k < 4; ++k)
                TWG_CHECK(c, tets4[4 * t + k] >= 0 && (uint32_t)tets4[4 * t + k] < nV, TWG_ERR_INVALID_ARG, "tet references a vertex out of range");
    *out = nullptr;
    TWG_CUDA(c, cudaSetDevice(c->device));
    twg_mesh* m = new twg_mesh;
    m->ctx = c;
    int rc = grow(m, nV, nT);
    if (rc == 0 && nV) rc = (int)cudaMemcpyAsync(m->V, V, (size_t)nV * 24, cudaMemcpyHostToDevice, c->streams[0]);
    if (rc == 0 && nT) rc = (int)cudaMemcpyAsync(m->T, tets4, (size_t)nT * 16, cudaMemcpyHostToDevice, c->streams[0]);
    if (rc == 0) rc = (int)cudaStreamSynchronize(c->streams[0]);
    if (rc != 0) {
        twg_fail(c, rc, "twg_mesh_create failed", __FILE__, __LINE__);
        twg_mesh_destroy(m);
        return rc;
    }
    *out = m;
    return 0;
}

void twg_mesh_destroy(twg_mesh* m) {
    if (!m) return;
    cudaSetDevice(m->ctx->device);
    cudaStreamSynchronize(m->ctx->streams[0]);
    if (m->V) cudaFree(m->V);
    if (m->T) cudaFree(m->T);
    if (m->adj_tets) cudaFree(m->adj_tets);
    if (m->adj_off) cudaFree(m->adj_off);
    delete m;
}

uint32_t twg_mesh_num_vertices(const twg_mesh* m) { return m ? m->nV : 0; }
uint64_t twg_mesh_num_tets(const twg_mesh* m) { return m ? m->nT : 0; }
const double* twg_mesh_vertices_dev(const twg_mesh* m) { return m ? m->V : nullptr; }
const int32_t* twg_mesh_tets_dev(const twg_mesh* m) { return m ? (const int32_t*)m->T : nullptr; }

int twg_mesh_resize(twg_mesh* m, uint32_t nV, uint64_t nT) {
    twg_ctx* c = m ? m->ctx : nullptr;
    TWG_CHECK(c, m, TWG_ERR_INVALID_ARG, "null argument");
    TWG_CHECK(c, nV >= m->nV && nT >= m->nT, TWG_ERR_INVALID_ARG, "the mesh only grows (removed tets keep their slot, like t_is_removed)");
    TWG_CHECK(c, nT < (1ull << 29), TWG_ERR_INVALID_ARG, "at most 2^29 tets");
    TWG_CUDA(c, cudaSetDevice(c->device));
    return grow(m, nV, nT);
}

int twg_mesh_set_vertices(twg_mesh* m, const int32_t* ids, const double* xyz, uint64_t n) {
    twg_ctx* c = m ? m->ctx : nullptr;
    TWG_CHECK(c, m && (n == 0 || (ids && xyz)), TWG_ERR_INVALID_ARG, "null argument");
    if (n == 0) return 0;
    for (uint64_t i = 0; i < n; ++i) TWG_CHECK(c, ids[i] >= 0 && (uint32_t)ids[i] < m->nV, TWG_ERR_INVALID_ARG, "vertex id out of range");
    TWG_CUDA(c, cudaSetDevice(c->device));
    TWG_TRY(twg_ensure_scratch(c, 0, up256(n * 4) + up256(n * 24)));
    size_t o = 0;
    int32_t* dI; double* dX;
    TWG_TRY(stage(c, o, ids, n, &dI));
    TWG_TRY(stage(c, o, xyz, 3 * n, &dX));
    TWG_LAUNCH(c, scatter_vertices_kernel, (unsigned)((n + 255) / 256), 256, 0, c->streams[0], m->V, dI, dX, n, m->nV);
    TWG_CUDA(c, cudaStreamSynchronize(c->streams[0]));  // the caller may reuse ids/xyz
    return 0;
}

int twg_mesh_set_tets(twg_mesh* m, const int32_t* ids, const int32_t* tets4, uint64_t n) {
    twg_ctx* c = m ? m->ctx : nullptr;
    TWG_CHECK(c, m && (n == 0 || (ids && tets4)), TWG_ERR_INVALID_ARG, "null argument");
    if (n == 0) return 0;
    for (uint64_t i = 0; i < n; ++i) {
        TWG_CHECK(c, ids[i] >= 0 && (uint64_t)ids[i] < m->nT, TWG_ERR_INVALID_ARG, "tet id out of range");
        if (tets4[4 * i] >= 0)
            for (int k = 0; k < 4; ++k) TWG_CHECK(c, tets4[4 * i + k] >= 0 && (uint32_t)tets4[4 * i + k] < m->nV, TWG_ERR_INVALID_ARG, "tet references a vertex out of range");
    }
    TWG_CUDA(c, cudaSetDevice(c->device));
    TWG_TRY(twg_ensure_scratch(c, 0, up256(n * 4) + up256(n * 16)));
    size_t o = 0;
    int32_t *dI, *dT;
    TWG_TRY(stage(c, o, ids, n, &dI));
    TWG_TRY(stage(c, o, tets4, 4 * n, &dT));
    TWG_LAUNCH(c, scatter_tets_kernel, (unsigned)((n + 255) / 256), 256, 0, c->streams[0], m->T, dI, (const int4*)dT, n, m->nT);
    TWG_CUDA(c, cudaStreamSynchronize(c->streams[0]));
    m->adj_valid = false;
    return 0;
}

int twg_mesh_get_vertices(twg_mesh* m, double* xyz_out) {
    twg_ctx* c = m ? m->ctx : nullptr;
    TWG_CHECK(c, m && xyz_out, TWG_ERR_INVALID_ARG, "null argument");
    TWG_CUDA(c, cudaSetDevice(c->device));
    TWG_CUDA(c, cudaMemcpyAsync(xyz_out, m->V, (size_t)m->nV * 24, cudaMemcpyDeviceToHost, c->streams[0]));
    TWG_CUDA(c, cudaStreamSynchronize(c->streams[0]));
    return 0;
}

int twg_mesh_build_rings(twg_mesh* m) {
    twg_ctx* c = m ? m->ctx : nullptr;
    TWG_CHECK(c, m, TWG_ERR_INVALID_ARG, "null argument");
    if (m->adj_valid) return 0;
    TWG_CUDA(c, cudaSetDevice(c->device));
    cudaStream_t st = c->streams[0];
    const uint64_t np = 4 * m->nT;
    if (m->adj_cap < np) {
        if (m->adj_tets) TWG_CUDA(c, cudaFree(m->adj_tets));
        m->adj_tets = nullptr;
        m->adj_cap = np + np / 4 + 64;
        TWG_CUDA(c, cudaMalloc(&m->adj_tets, m->adj_cap * 4));
    }
    if (m->adj_off_cap < (uint64_t)m->nV + 1) {
        if (m->adj_off) TWG_CUDA(c, cudaFree(m->adj_off));
        m->adj_off = nullptr;
        m->adj_off_cap = (uint64_t)m->nV + 1 + m->nV / 4 + 64;
        TWG_CUDA(c, cudaMalloc(&m->adj_off, m->adj_off_cap * 8));
    }
    if (np == 0) {
        TWG_CUDA(c, cudaMemsetAsync(m->adj_off, 0, ((size_t)m->nV + 1) * 8, st));
        m->adj_valid = true;
        return 0;
    }
    size_t tmp_bytes = 0;
    TWG_CUDA(c, cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, (uint32_t*)nullptr, (uint32_t*)nullptr, (int32_t*)nullptr, (int32_t*)nullptr, (int)np, 0, 32, st));
    const size_t kb = up256(np * 4);
    TWG_TRY(twg_ensure_scratch(c, 0, 3 * kb + up256(tmp_bytes)));
    char* base = (char*)c->dscratch[0];
    uint32_t *keys = (uint32_t*)base, *keys2 = (uint32_t*)(base + kb);
    int32_t* vals = (int32_t*)(base + 2 * kb);
    void* tmp = base + 3 * kb;
    TWG_LAUNCH(c, adj_pairs_kernel, (unsigned)((m->nT + 255) / 256), 256, 0, st, m->T, m->nT, keys, vals);
    TWG_CUDA(c, cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys, keys2, vals, m->adj_tets, (int)np, 0, 32, st));  // stable: tets ascend within a vertex
    c->launches += 6;
    TWG_LAUNCH(c, adj_offsets_kernel, (unsigned)(((uint64_t)m->nV + 1 + 255) / 256), 256, 0, st, keys2, np, m->nV, m->adj_off);
    TWG_CUDA(c, cudaStreamSynchronize(st));
    m->adj_valid = true;
    return 0;
}

int twg_mesh_get_rings(twg_mesh* m, uint64_t* off_out /* nV+1 */, int32_t* tets_out /* off_out[nV] entries, may be NULL */) {
    twg_ctx* c = m ? m->ctx : nullptr;
    TWG_CHECK(c, m && off_out, TWG_ERR_INVALID_ARG, "null argument");
    TWG_TRY(twg_mesh_build_rings(m));
    cudaStream_t st = c->streams[0];
    TWG_CUDA(c, cudaMemcpyAsync(off_out, m->adj_off, ((size_t)m->nV + 1) * 8, cudaMemcpyDeviceToHost, st));
    TWG_CUDA(c, cudaStreamSynchronize(st));
    if (tets_out && off_out[m->nV]) {
        TWG_CUDA(c, cudaMemcpyAsync(tets_out, m->adj_tets, (size_t)off_out[m->nV] * 4, cudaMemcpyDeviceToHost, st));
        TWG_CUDA(c, cudaStreamSynchronize(st));
    }
    return 0;
}

// a5 over the resident mesh, rings named by their centre vertex (conn_tets): 4 B in, 105 B out per ring
int twg_mesh_vertex_ring_ejh(twg_mesh* m, const int32_t* v_ids, uint64_t n, double* E, double* J3, double* H9, uint8_t* ok) {
    twg_ctx* c = m ? m->ctx : nullptr;
    TWG_CHECK(c, m && (n == 0 || (v_ids && E && J3 && H9)), TWG_ERR_INVALID_ARG, "null argument");
    if (n == 0) return 0;
    for (uint64_t i = 0; i < n; ++i) TWG_CHECK(c, v_ids[i] >= 0 && (uint32_t)v_ids[i] < m->nV, TWG_ERR_INVALID_ARG, "vertex id out of range");
    TWG_TRY(twg_mesh_build_rings(m));
    TWG_CUDA(c, cudaSetDevice(c->device));
    cudaStream_t st = c->streams[0];
    if (n <= kFastRings && c->opt.fast_calls) {
        // ONE Newton step of the sequential scheduler: ids and results live in the mapped slab, the kernel reads and writes them over
        // PCIe, the host spins on the completion word -- one kernel launch is the only driver call that waits on anything
        char* slab;
        const size_t ib = up256(n * 4), eb = up256(n * 8), jb = up256(n * 24), hb = up256(n * 72), kb = up256(n);
        TWG_TRY(twg_fast_slab(c, ib + eb + jb + hb + kb, &slab));
        twg_done done;
        TWG_TRY(twg_fast_arm(c, &done));
        TWG_TRY(twg_amips_ring_tiny(c, m->V, m->nV, (const int32_t*)m->T, m->nT, m->adj_tets, m->adj_off, v_ids, nullptr, (uint32_t)n, (double*)(slab + ib),
                                    (double*)(slab + ib + eb), (double*)(slab + ib + eb + jb), (uint8_t*)(slab + ib + eb + jb + hb), st, &done));
        TWG_TRY(twg_fast_spin(c, st, done.seq));
        memcpy(E, slab + ib, n * 8);
        memcpy(J3, slab + ib + eb, n * 24);
        memcpy(H9, slab + ib + eb + jb, n * 72);
        if (ok) memcpy(ok, slab + ib + eb + jb + hb, n);
        return 0;
    }
    if (n <= kSmallCall) {
        // a handful of rings (what one un-batched Newton step of the scheduler asks for): latency is everything. Ids go
        // through a pinned slab in ONE copy, the four result arrays come back packed in ONE copy (each extra cudaMemcpy of a
        // pageable buffer costs ~10 us: 53 -> ~25 us per call).
        const size_t ib = up256(n * 4), eb = up256(n * 8), jb = up256(n * 24), hb = up256(n * 72), kb = up256(n);
        TWG_TRY(twg_ensure_pinned(c, ib, eb + jb + hb + kb));
        TWG_TRY(twg_ensure_scratch(c, 0, ib + eb + jb + hb + kb));
        char* d = (char*)c->dscratch[0];
        char* ho = (char*)c->pin_out[0];
        memcpy(c->pin_in[0], v_ids, n * 4);
        TWG_CUDA(c, cudaMemcpyAsync(d, c->pin_in[0], n * 4, cudaMemcpyHostToDevice, st));
        TWG_TRY(twg_amips_vertex_ring_ejh_dev(c, m->V, m->nV, (const int32_t*)m->T, m->nT, m->adj_tets, m->adj_off, (const int32_t*)d, n, (double*)(d + ib),
                                              (double*)(d + ib + eb), (double*)(d + ib + eb + jb), (uint8_t*)(d + ib + eb + jb + hb), st));
        TWG_CUDA(c, cudaMemcpyAsync(ho, d + ib, eb + jb + hb + (ok ? n : 0), cudaMemcpyDeviceToHost, st));
        TWG_CUDA(c, cudaStreamSynchronize(st));
        memcpy(E, ho, n * 8);
        memcpy(J3, ho + eb, n * 24);
        memcpy(H9, ho + eb + jb, n * 72);
        if (ok) memcpy(ok, ho + eb + jb + hb, n);
        return 0;
    }
    // chunks on two streams so that the result copies of one chunk overlap the kernel of the next
    const uint64_t chunk = 1ull << 20;
    const uint64_t cm = n < chunk ? n : chunk;
    const size_t per = up256(cm * 4) + up256(cm * 8) + up256(cm * 24) + up256(cm * 72) + up256(cm);
    for (int k = 0; k < 2; ++k) TWG_TRY(twg_ensure_scratch(c, k, per));
    int slot = 0;
    for (uint64_t b = 0; b < n; b += chunk, slot ^= 1) {
        const uint64_t k = (n - b < chunk) ? n - b : chunk;
        st = c->streams[slot];
        char* p = (char*)c->dscratch[slot];
        int32_t* dI = (int32_t*)p; p += up256(cm * 4);
        double* dE = (double*)p; p += up256(cm * 8);
        double* dJ = (double*)p; p += up256(cm * 24);
        double* dH = (double*)p; p += up256(cm * 72);
        uint8_t* dK = (uint8_t*)p;
        TWG_CUDA(c, cudaMemcpyAsync(dI, v_ids + b, k * 4, cudaMemcpyHostToDevice, st));
        TWG_TRY(twg_amips_vertex_ring_ejh_dev(c, m->V, m->nV, (const int32_t*)m->T, m->nT, m->adj_tets, m->adj_off, dI, k, dE, dJ, dH, dK, st));
        TWG_CUDA(c, cudaMemcpyAsync(E + b, dE, k * 8, cudaMemcpyDeviceToHost, st));
        TWG_CUDA(c, cudaMemcpyAsync(J3 + 3 * b, dJ, k * 24, cudaMemcpyDeviceToHost, st));
        TWG_CUDA(c, cudaMemcpyAsync(H9 + 9 * b, dH, k * 72, cudaMemcpyDeviceToHost, st));
        if (ok) TWG_CUDA(c, cudaMemcpyAsync(ok + b, dK, k, cudaMemcpyDeviceToHost, st));
    }
    for (int k = 0; k < 2; ++k) TWG_CUDA(c, cudaStreamSynchronize(c->streams[k]));
    return 0;
}

// a5 / a6 with explicit member lists (t_ids, CSR group_off) against the resident V / T: only the index lists travel
int twg_mesh_ring_ejh(twg_mesh* m, const int32_t* t_ids, const uint64_t* group_off, const int32_t* center, uint64_t nG, double* E, double* J3,
                      double* H9, uint8_t* ok) {
    twg_ctx* c = m ? m->ctx : nullptr;
    TWG_CHECK(c, m && (nG == 0 || (t_ids && group_off && center && E && J3 && H9)), TWG_ERR_INVALID_ARG, "null argument");
    if (nG == 0) return 0;
    const uint64_t nM = group_off[nG];
    for (uint64_t i = 0; i < nM; ++i) TWG_CHECK(c, t_ids[i] >= 0 && (uint64_t)t_ids[i] < m->nT, TWG_ERR_INVALID_ARG, "tet id out of range");
    TWG_CUDA(c, cudaSetDevice(c->device));
    cudaStream_t st = c->streams[0];
    TWG_TRY(twg_ensure_scratch(c, 0, up256(nM * 4) + up256((nG + 1) * 8) + up256(nG * 4) + up256(nG * 8) + up256(nG * 24) + up256(nG * 72) + up256(nG)));
    size_t o = 0;
    int32_t *dI, *dC; uint64_t* dO; double *dE, *dJ, *dH; uint8_t* dK;
    TWG_TRY(stage(c, o, t_ids, nM, &dI));
    TWG_TRY(stage(c, o, group_off, nG + 1, &dO));
    TWG_TRY(stage(c, o, center, nG, &dC));
    TWG_TRY(stage(c, o, (const double*)nullptr, nG, &dE));
    TWG_TRY(stage(c, o, (const double*)nullptr, 3 * nG, &dJ));
    TWG_TRY(stage(c, o, (const double*)nullptr, 9 * nG, &dH));
    TWG_TRY(stage(c, o, (const uint8_t*)nullptr, nG, &dK));
    TWG_TRY(twg_amips_ring_ejh_dev(c, m->V, m->nV, (const int32_t*)m->T, m->nT, dI, dO, dC, nG, dE, dJ, dH, dK, st));
    if (nG <= kSmallCall) {  // E | J | H | ok are consecutive in the scratch slab: one packed copy through pinned memory
        const size_t eb = up256(nG * 8), jb = up256(nG * 24), hb = up256(nG * 72);
        TWG_TRY(twg_ensure_pinned(c, 256, eb + jb + hb + up256(nG)));
        char* ho = (char*)c->pin_out[0];
        TWG_CUDA(c, cudaMemcpyAsync(ho, dE, eb + jb + hb + (ok ? nG : 0), cudaMemcpyDeviceToHost, st));
        TWG_CUDA(c, cudaStreamSynchronize(st));
        memcpy(E, ho, nG * 8);
        memcpy(J3, ho + eb, nG * 24);
        memcpy(H9, ho + eb + jb, nG * 72);
        if (ok) memcpy(ok, ho + eb + jb + hb, nG);
        return 0;
    }
    TWG_CUDA(c, cudaMemcpyAsync(E, dE, nG * 8, cudaMemcpyDeviceToHost, st));
    TWG_CUDA(c, cudaMemcpyAsync(J3, dJ, nG * 24, cudaMemcpyDeviceToHost, st));
    TWG_CUDA(c, cudaMemcpyAsync(H9, dH, nG * 72, cudaMemcpyDeviceToHost, st));
    if (ok) TWG_CUDA(c, cudaMemcpyAsync(ok, dK, nG, cudaMemcpyDeviceToHost, st));
    TWG_CUDA(c, cudaStreamSynchronize(st));
    return 0;
}

int twg_mesh_ring_energy(twg_mesh* m, const int32_t* t_ids, const uint64_t* group_off, uint64_t nG, double* E) {
    twg_ctx* c = m ? m->ctx : nullptr;
    TWG_CHECK(c, m && (nG == 0 || (t_ids && group_off && E)), TWG_ERR_INVALID_ARG, "null argument");
    if (nG == 0) return 0;
    const uint64_t nM = group_off[nG];
    for (uint64_t i = 0; i < nM; ++i) TWG_CHECK(c, t_ids[i] >= 0 && (uint64_t)t_ids[i] < m->nT, TWG_ERR_INVALID_ARG, "tet id out of range");
    TWG_CUDA(c, cudaSetDevice(c->device));
    cudaStream_t st = c->streams[0];
    TWG_TRY(twg_ensure_scratch(c, 0, up256(nM * 4) + up256((nG + 1) * 8) + up256(nG * 8)));
    size_t o = 0;
    int32_t* dI; uint64_t* dO; double* dE;
    TWG_TRY(stage(c, o, t_ids, nM, &dI));
    TWG_TRY(stage(c, o, group_off, nG + 1, &dO));
    TWG_TRY(stage(c, o, (const double*)nullptr, nG, &dE));
    TWG_TRY(twg_amips_ring_energy_dev(c, m->V, m->nV, (const int32_t*)m->T, m->nT, dI, dO, nG, dE, st));
    TWG_CUDA(c, cudaMemcpyAsync(E, dE, nG * 8, cudaMemcpyDeviceToHost, st));
    TWG_CUDA(c, cudaStreamSynchronize(st));
    return 0;
}

// The line search of the smoother (VertexSmoother.cpp:505-541): "move v to p, getNewEnergy(conn_tets[v]), move it back" for n
// (vertex, trial position) pairs at once. The resident mesh is NOT modified: every ring is evaluated with its centre vertex at
// the trial position, so the pairs are independent -- all step sizes of one Newton step, or the steps of many candidates.
int twg_mesh_vertex_trial_energy(twg_mesh* m, const int32_t* v_ids, const double* xyz, uint64_t n, double* E) {
    twg_ctx* c = m ? m->ctx : nullptr;
    TWG_CHECK(c, m && (n == 0 || (v_ids && xyz && E)), TWG_ERR_INVALID_ARG, "null argument");
    if (n == 0) return 0;
    for (uint64_t i = 0; i < n; ++i) TWG_CHECK(c, v_ids[i] >= 0 && (uint32_t)v_ids[i] < m->nV, TWG_ERR_INVALID_ARG, "vertex id out of range");
    TWG_TRY(twg_mesh_build_rings(m));
    TWG_CUDA(c, cudaSetDevice(c->device));
    cudaStream_t st = c->streams[0];
    const size_t ib = up256(n * 4), xb = up256(n * 24), eb = up256(n * 8);
    if (n <= kFastRings && c->opt.fast_calls) {  // the step sizes of one line search: zero-copy slab
        char* slab;
        TWG_TRY(twg_fast_slab(c, ib + xb + eb, &slab));
        twg_done done;
        TWG_TRY(twg_fast_arm(c, &done));
        TWG_TRY(twg_amips_ring_tiny(c, m->V, m->nV, (const int32_t*)m->T, m->nT, m->adj_tets, m->adj_off, v_ids, xyz, (uint32_t)n, (double*)(slab + ib + xb), nullptr,
                                    nullptr, nullptr, st, &done));
        TWG_TRY(twg_fast_spin(c, st, done.seq));
        memcpy(E, slab + ib + xb, n * 8);
        return 0;
    }
    TWG_TRY(twg_ensure_scratch(c, 0, ib + xb + eb));
    char* d = (char*)c->dscratch[0];
    if (n <= kSmallCall) {  // one packed copy each way through the pinned slabs
        TWG_TRY(twg_ensure_pinned(c, ib + xb, eb));
        memcpy(c->pin_in[0], v_ids, n * 4);
        memcpy((char*)c->pin_in[0] + ib, xyz, n * 24);
        TWG_CUDA(c, cudaMemcpyAsync(d, c->pin_in[0], ib + n * 24, cudaMemcpyHostToDevice, st));
    } else {
        TWG_CUDA(c, cudaMemcpyAsync(d, v_ids, n * 4, cudaMemcpyHostToDevice, st));
        TWG_CUDA(c, cudaMemcpyAsync(d + ib, xyz, n * 24, cudaMemcpyHostToDevice, st));
    }
    TWG_TRY(twg_amips_vertex_trial_energy_dev(c, m->V, m->nV, (const int32_t*)m->T, m->nT, m->adj_tets, m->adj_off, (const int32_t*)d, (const double*)(d + ib), n,
                                              (double*)(d + ib + xb), st));
    if (n <= kSmallCall) {
        TWG_CUDA(c, cudaMemcpyAsync(c->pin_out[0], d + ib + xb, n * 8, cudaMemcpyDeviceToHost, st));
        TWG_CUDA(c, cudaStreamSynchronize(st));
        memcpy(E, c->pin_out[0], n * 8);
    } else {
        TWG_CUDA(c, cudaMemcpyAsync(E, d + ib + xb, n * 8, cudaMemcpyDeviceToHost, st));
        TWG_CUDA(c, cudaStreamSynchronize(st));
    }
    return 0;
}

int twg_mesh_vertex_trial_energy_dev(twg_mesh* m, const int32_t* dVids, const double* dXyz, uint64_t n, double* dE, void* stream) {
    twg_ctx* c = m ? m->ctx : nullptr;
    TWG_CHECK(c, m && dVids && dXyz && dE, TWG_ERR_INVALID_ARG, "null argument");
    if (n == 0) return 0;
    TWG_TRY(twg_mesh_build_rings(m));
    return twg_amips_vertex_trial_energy_dev(c, m->V, m->nV, (const int32_t*)m->T, m->nT, m->adj_tets, m->adj_off, dVids, dXyz, n, dE, stream);
}

static int per_tet_host(twg_mesh* m, int what, const int32_t* t_ids, uint64_t n, double* out0, double* out1) {
    twg_ctx* c = m->ctx;
    if (!t_ids) n = m->nT;
    if (n == 0) return 0;
    if (t_ids)
        for (uint64_t i = 0; i < n; ++i) TWG_CHECK(c, t_ids[i] >= 0 && (uint64_t)t_ids[i] < m->nT, TWG_ERR_INVALID_ARG, "tet id out of range");
    TWG_CUDA(c, cudaSetDevice(c->device));
    cudaStream_t st = c->streams[0];
    if (t_ids && n <= kFastTets && c->opt.fast_calls) {  // the tets of one local operation: zero-copy slab, no cudaMemcpy, no synchronize
        char* slab;
        const size_t ib = up256(n * 4), eb = up256(n * 8);
        TWG_TRY(twg_fast_slab(c, ib + 2 * eb, &slab));
        if (what == 0) {
            TwgTinyTets in;
            memcpy(in.id, t_ids, n * 4);
            twg_done done;
            TWG_TRY(twg_fast_arm(c, &done));
            TWG_LAUNCH(c, mesh_quality_tiny_kernel, 1, 256, 0, st, m->V, m->T, in, (uint32_t)n, (double*)(slab + ib), done);
            TWG_TRY(twg_fast_spin(c, st, done.seq));
        } else {
            memcpy(slab, t_ids, n * 4);
            TWG_LAUNCH(c, mesh_dihedral_kernel, grid_for(c, n, 256, 8), 256, 0, st, m->V, m->T, (const int32_t*)slab, n, (double*)(slab + ib), (double*)(slab + ib + eb));
            TWG_TRY(twg_fast_wait(c, st));
        }
        memcpy(out0, slab + ib, n * 8);
        if (what == 1) memcpy(out1, slab + ib + eb, n * 8);
        return 0;
    }
    TWG_TRY(twg_ensure_scratch(c, 0, up256(t_ids ? n * 4 : 0) + 2 * up256(n * 8)));
    size_t o = 0;
    int32_t* dI = nullptr; double *d0, *d1;
    if (t_ids) TWG_TRY(stage(c, o, t_ids, n, &dI));
    TWG_TRY(stage(c, o, (const double*)nullptr, n, &d0));
    TWG_TRY(stage(c, o, (const double*)nullptr, n, &d1));
    if (what == 0 && c->opt.wide_gather) TWG_LAUNCH(c, mesh_quality_kernel<true>, grid_for(c, n, 256, 8), 256, 0, st, m->V, m->T, dI, n, d0);
    else if (what == 0) TWG_LAUNCH(c, mesh_quality_kernel<false>, grid_for(c, n, 256, 8), 256, 0, st, m->V, m->T, dI, n, d0);
    else TWG_LAUNCH(c, mesh_dihedral_kernel, grid_for(c, n, 256, 8), 256, 0, st, m->V, m->T, dI, n, d0, d1);
    TWG_CUDA(c, cudaMemcpyAsync(out0, d0, n * 8, cudaMemcpyDeviceToHost, st));
    if (what == 1) TWG_CUDA(c, cudaMemcpyAsync(out1, d1, n * 8, cudaMemcpyDeviceToHost, st));
    TWG_CUDA(c, cudaStreamSynchronize(st));
    return 0;
}

int twg_mesh_quality(twg_mesh* m, const int32_t* t_ids, uint64_t n, double* slim_energy) {
    twg_ctx* c = m ? m->ctx : nullptr;
    TWG_CHECK(c, m && slim_energy, TWG_ERR_INVALID_ARG, "null argument");
    return per_tet_host(m, 0, t_ids, n, slim_energy, nullptr);
}

int twg_mesh_dihedral(twg_mesh* m, const int32_t* t_ids, uint64_t n, double* min_d_angle, double* max_d_angle) {
    twg_ctx* c = m ? m->ctx : nullptr;
    TWG_CHECK(c, m && min_d_angle && max_d_angle, TWG_ERR_INVALID_ARG, "null argument");
    return per_tet_host(m, 1, t_ids, n, min_d_angle, max_d_angle);
}

/* device-pointer variants of the two whole-mesh passes (results stay on the device, asynchronous on `stream`) */
int twg_mesh_quality_dev(twg_mesh* m, const int32_t* dTids, uint64_t n, double* dSlim, void* stream) {
    twg_ctx* c = m ? m->ctx : nullptr;
    TWG_CHECK(c, m && dSlim, TWG_ERR_INVALID_ARG, "null argument");
    if (!dTids) n = m->nT;
    if (n == 0) return 0;
    TWG_CUDA(c, cudaSetDevice(c->device));
    if (c->opt.wide_gather) TWG_LAUNCH(c, mesh_quality_kernel<true>, grid_for(c, n, 256, 8), 256, 0, stream ? (cudaStream_t)stream : c->streams[0], m->V, m->T, dTids, n, dSlim);
    else TWG_LAUNCH(c, mesh_quality_kernel<false>, grid_for(c, n, 256, 8), 256, 0, stream ? (cudaStream_t)stream : c->streams[0], m->V, m->T, dTids, n, dSlim);
    return 0;
}
int twg_mesh_dihedral_dev(twg_mesh* m, const int32_t* dTids, uint64_t n, double* dMin, double* dMax, void* stream) {
    twg_ctx* c = m ? m->ctx : nullptr;
    TWG_CHECK(c, m && dMin && dMax, TWG_ERR_INVALID_ARG, "null argument");
    if (!dTids) n = m->nT;
    if (n == 0) return 0;
    TWG_CUDA(c, cudaSetDevice(c->device));
    TWG_LAUNCH(c, mesh_dihedral_kernel, grid_for(c, n, 256, 8), 256, 0, stream ? (cudaStream_t)stream : c->streams[0], m->V, m->T, dTids, n, dMin, dMax);
    return 0;
}
int twg_mesh_vertex_ring_ejh_dev(twg_mesh* m, const int32_t* dVids, uint64_t n, double* dE, double* dJ3, double* dH9, uint8_t* dOk, void* stream) {
    twg_ctx* c = m ? m->ctx : nullptr;
    TWG_CHECK(c, m && dVids && dE && dJ3 && dH9, TWG_ERR_INVALID_ARG, "null argument");
    TWG_TRY(twg_mesh_build_rings(m));
    return twg_amips_vertex_ring_ejh_dev(c, m->V, m->nV, (const int32_t*)m->T, m->nT, m->adj_tets, m->adj_off, dVids, n, dE, dJ3, dH9, dOk, stream);
}

}  // extern "C"
