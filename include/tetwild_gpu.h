/*
 * tetwild_gpu.h -- C ABI of libtetwild_gpu.so: TetWild's data-parallel hot path on NVIDIA B200 (sm_100a).
 *
 * The reference (Yixin-Hu/TetWild @49de8cd) has no plugin/FFI layer for this path except one true C ABI, the ISPC
 * operator energy_ispc (src/ispc/energy.ispc:7-21). This header is the boundary a maintainer binds instead of:
 *
 *   S1  ispc::energy_ispc(V1_x..V4_z, E, count)                   src/ispc/energy.ispc:7-21, LocalOperations.h:21-23
 *   S2  GEO::MeshFacetsAABBWithEps (ctor, nearest_facet,          src/tetwild/geogram/mesh_AABB.h:77,130,162,182,199,221
 *       facet_in_envelope[_with_hint], squared_distance)
 *   S3  LocalOperations::isFaceOutEnvelop / isPointOutEnvelop /   src/tetwild/LocalOperations.h:69,97-100,109-111
 *       calTetQualities / comformalAMIPS{Energy,Jacobian,Hessian}_new,
 *       VertexSmoother::NewtonsUpdate / getNewEnergy              src/tetwild/VertexSmoother.cpp:627-702, :544-625
 *   S4  igl::winding_number(V,F,O,W) + the W > 0.5 rule           src/tetwild/InoutFiltering.cpp:45-75,
 *                                                                 src/tetwild/MeshRefinement.cpp:614,1056
 *
 * Conventions
 *   - plain pointers and sizes only; every function returns 0 on success, a cudaError_t value or a TWG_ERR_* code
 *     otherwise, never throws; twg_last_error() gives the message. There is NO CPU fallback: if no sm_100-class
 *     device can be opened twg_create fails.
 *   - functions without suffix take HOST buffers (caller-owned; pageable or pinned) and include the host<->device
 *     copies; functions ending in _dev take DEVICE pointers on the context's device plus a cudaStream_t (as void*,
 *     NULL = the context's stream) and are asynchronous on that stream.
 *   - vertices are xyz-interleaved doubles, facet / tet indices are 32-bit, facet ids returned are in the CALLER's
 *     numbering (the library sorts facets internally, like mesh_reorder at mesh_AABB.cpp:368-370, but never
 *     mutates caller data).
 *   - calls on one context are serialised by the caller (the reference is single-threaded). _dev calls on DIFFERENT
 *     caller streams may overlap on the device: every stream owns its sort scratch and work counters (a "lane"; at
 *     most 13 caller streams are tracked, further ones recycle the least recently used lane after its work finished).
 *     A _dev call takes at most 2^31-1 queries; the host entry points chunk any size.
 *   - multi-GPU: twg_create_multi opens one context over several devices of this process. Surface and
 *     winding handles made on it are replicated on every device; every host-buffer batch call splits [0, n) into
 *     contiguous index ranges, one per device, each staged and evaluated by that device's own host thread, results
 *     written straight into the caller's buffer. _dev entry points need the one-device handles
 *     (twg_device_context / twg_surface_replica / twg_winding_replica). The resident tet mesh (twg_mesh) lives on device 0.
 */
#ifndef TETWILD_GPU_H
#define TETWILD_GPU_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TWG_ERR_INTERNAL 10001
#define TWG_ERR_INVALID_ARG 10002
#define TWG_ERR_NO_DEVICE 10003
#define TWG_ERR_ALIGNMENT 10004

#define TWG_MAX_ENERGY 1e50        /* State::MAX_ENERGY, src/tetwild/State.h:29 */
#define TWG_NO_FACET 0xffffffffu   /* GEO::NO_FACET */

typedef struct twg_ctx twg_ctx;
typedef struct twg_surface twg_surface;   /* S2: envelope / nearest-facet structure over the input surface */
typedef struct twg_winding twg_winding;   /* S4: winding-number hierarchy over a (tracked) surface */

/* ---- context ------------------------------------------------------------------------------------------------ */
int twg_create(twg_ctx** ctx, int device_id);
/* one context over n_devices devices (n_devices == 1 is twg_create; an id may repeat, which puts several worker contexts on
 * one device -- useful for testing the split on a one-GPU box). SURVEY.md 8(b): twg_create(ctx, device_ids, n) */
int twg_create_multi(twg_ctx** ctx, const int* device_ids, int n_devices);
int twg_num_devices(const twg_ctx* ctx);
twg_ctx* twg_device_context(twg_ctx* ctx, int k);  /* the one-device context of device k (k = 0: ctx itself when single) */
void twg_destroy(twg_ctx* ctx);
const char* twg_last_error(const twg_ctx* ctx);
int twg_device(const twg_ctx* ctx);
int twg_synchronize(twg_ctx* ctx);
uint64_t twg_launch_count(const twg_ctx* ctx);  /* kernels launched by this context so far */
const char* twg_version(void);
/* Tuning knobs, per context (defaults: environment variable TWG_<NAME> read once at twg_create). Names: env_group, env_policy,
 * env_front, env_quorum, env_top, env_bound, envelope_sort, surface_order, sort_bits, sort_curve, nearest_curve, chunk_points, ring_waves, ring_minb, wide_gather, winding_minb, winding_sort,
 * winding_leaf, winding_device_build, amips_tma, nearest_mode, nearest_group, nearest_budget, fast_calls, trace. Values are clamped to their valid range. */
int twg_set_option(twg_ctx* ctx, const char* name, double value);
int twg_get_option(const twg_ctx* ctx, const char* name, double* value);
/* diagnostics (cumulative per context) */
#define TWG_COUNTER_ENV_STACK_OVERFLOW 0
#define TWG_COUNTER_WINDING_PAIRS 2 /* (query, cap point or facet) evaluations of the winding kernel so far */
int twg_debug_counter(twg_ctx* ctx, int which, uint64_t* value);

/* ---- S2: surface build (replaces MeshFacetsAABBWithEps::MeshFacetsAABBWithEps, mesh_AABB.cpp:356-379) ------------ */
int twg_surface_create(twg_ctx* ctx, const double* V, uint32_t nV, const uint32_t* F, uint32_t nF, twg_surface** out);
int twg_surface_create_dev(twg_ctx* ctx, const double* dV, uint32_t nV, const uint32_t* dF, uint32_t nF, twg_surface** out);
void twg_surface_destroy(twg_surface* s);
uint32_t twg_surface_num_facets(const twg_surface* s);
twg_surface* twg_surface_replica(twg_surface* s, int k);  /* replica on device k of a multi-device context (k = 0: s itself when single) */

/* a10/a12: out[i] = 1 iff min_f d2(P_i, f) > eps2
 * (isPointOutEnvelop LocalOperations.cpp:1034-1044; per-sample test of :1083-1093 via facet_in_envelope_with_hint) */
int twg_envelope_points_out(twg_surface* s, const double* P, uint64_t n, double eps2, uint8_t* out);
int twg_envelope_points_out_dev(twg_surface* s, const double* dP, uint64_t n, double eps2, uint8_t* dOut, void* stream);

/* a9: isFaceOutEnvelop(tri) (LocalOperations.cpp:967-976, :1046-1109): tris = n*9 doubles; triangles are sampled on
 * the device with sampleTriangle semantics (Common.cpp:143-255); out[i] = 1 iff some sample is farther than eps from
 * the surface; collinear triangles give 0 (:1048). */
int twg_envelope_faces_out(twg_surface* s, const double* tris, uint64_t n, double sampling_dist, double eps2, uint8_t* out);
int twg_envelope_faces_out_dev(twg_surface* s, const double* dTris, uint64_t n, double sampling_dist, double eps2,
                               uint8_t* dOut, void* stream);
/* flags for the _ex variants */
#define TWG_FACES_NO_DEGENERATE_SHORTCUT 1u /* sample degenerate faces too, like Preprocess::isOutEnvelop (Preprocess.cpp:643-747) */
/* the per-face body of Preprocess::isOutEnvelop when flags has TWG_FACES_NO_DEGENERATE_SHORTCUT (the caller ORs the faces of
 * one candidate set; its eps_2 is the 0.8-scaled one of Preprocess.cpp:201-205); flags = 0 is twg_envelope_faces_out */
int twg_envelope_faces_out_ex(twg_surface* s, const double* tris, uint64_t n, double sampling_dist, double eps2, uint32_t flags,
                              uint8_t* out);
int twg_envelope_faces_out_ex_dev(twg_surface* s, const double* dTris, uint64_t n, double sampling_dist, double eps2, uint32_t flags,
                                  uint8_t* dOut, void* stream);

/* a13: nearest_facet / squared_distance (mesh_AABB.h:130-141,221-226). Any output pointer may be NULL.
 * facet ids are unique up to exact distance ties. */
int twg_nearest(twg_surface* s, const double* P, uint64_t n, uint32_t* facet, double* nearest_xyz, double* d2);
int twg_nearest_dev(twg_surface* s, const double* dP, uint64_t n, uint32_t* dFacet, double* dNearest, double* dD2, void* stream);

/* a8 (debug / parity): the samples sampleTriangle would generate for one triangle, in the reference's order.
 * Writes at most cap points, *count = number the reference generates. */
int twg_sample_triangle(twg_ctx* ctx, const double* tri9, double sampling_dist, double* out_xyz, uint64_t cap, uint64_t* count);

/* test hook: the library's own Morton radix sort of a query batch (csrc/qsort.cu). box6 = lo xyz, hi xyz of the quantisation box
 * (NULL: the batch's own bounding box). perm[i] = caller index of the i-th point in sorted order; keys (nullable) = the sorted
 * 30-bit Morton keys (pairs-only form); sorted_xyz (nullable) = the points in sorted order (gathering form). */
int twg_debug_sort_points(twg_ctx* ctx, const double* P, uint64_t n, const double* box6, uint32_t* perm, uint32_t* keys, double* sorted_xyz);

/* measurement hook: the floor under every tiny call on this machine -- launch an empty kernel that raises the completion word in
 * mapped host memory, spin until the host sees it; *us_per_call = host wall time per round trip over `reps` of them. */
int twg_debug_roundtrip(twg_ctx* ctx, int reps, double* us_per_call);

/* ---- roofline denominators measured on the context's device (not on the hot path; bench.py calls them once) -------- */
/* FP64 vector pipe: DFMA microbenchmark, 2 flops per DFMA, TFLOP/s.  Copy: 128-bit grid-stride copy, read+write GB/s. */
int twg_measure_fp64_tflops(twg_ctx* ctx, double* tflops);
int twg_measure_fp64_tflops_distinct(twg_ctx* ctx, double* tflops);  /* three distinct register operands per DFMA */
int twg_measure_copy_gbs(twg_ctx* ctx, uint64_t bytes, double* gbs);

/* ---- S1/S3: AMIPS ------------------------------------------------------------------------------------------------ */
/* S1 mirror: exactly energy_ispc's argument list (12 SoA coordinate arrays, E, count) plus the context. */
int twg_amips_energy_soa(twg_ctx* ctx, const double* const T[12], double* E, uint64_t n);
int twg_amips_energy_soa_dev(twg_ctx* ctx, const double* const dT[12], double* dE, uint64_t n, void* stream);
/* a1+a2+a3 per tet: E[n], J3[n*3], H9[n*9] (row-major 3x3); any of E/J3/H9 may be NULL */
int twg_amips_ejh_soa(twg_ctx* ctx, const double* const T[12], double* E, double* J3, double* H9, uint64_t n);
int twg_amips_ejh_soa_dev(twg_ctx* ctx, const double* const dT[12], double* dE, double* dJ3, double* dH9, uint64_t n, void* stream);
/* a4: calTetQualities / calTetQuality_AMIPS (LocalOperations.cpp:695-773, :862-884): indexed gather, exact
 * orientation gate, slim_energy = MAX_ENERGY when not POSITIVE / inf / NaN / <= 0 */
int twg_amips_quality(twg_ctx* ctx, const double* Vxyz, uint32_t nV, const int32_t* tets4, uint64_t nT, double* slim_energy);
int twg_amips_quality_dev(twg_ctx* ctx, const double* dVxyz, uint32_t nV, const int32_t* dTets4, uint64_t nT, double* dSlim, void* stream);
/* a5: NewtonsUpdate (VertexSmoother.cpp:627-702) for nGroups one-rings at once. Group g owns members
 * k in [group_off[g], group_off[g+1]); member k is tet tets4[t_ids ? t_ids[k] : k]; the tet is rotated so that
 * center[g] sits in slot 0 (:640-651). Outputs E[g], J3[g*3], H9[g*9], ok[g] (0 where the reference returns false:
 * E NaN / <= 0, J or H not finite; E = MAX_ENERGY where it is +inf, :680-699). ok may be NULL. */
int twg_amips_ring_ejh(twg_ctx* ctx, const double* Vxyz, uint32_t nV, const int32_t* tets4, uint64_t nT, const int32_t* t_ids,
                       const uint64_t* group_off, const int32_t* center, uint64_t nGroups, double* E, double* J3, double* H9,
                       uint8_t* ok);
int twg_amips_ring_ejh_dev(twg_ctx* ctx, const double* dVxyz, uint32_t nV, const int32_t* dTets4, uint64_t nT,
                           const int32_t* dTids, const uint64_t* dGroupOff, const int32_t* dCenter, uint64_t nGroups,
                           double* dE, double* dJ3, double* dH9, uint8_t* dOk, void* stream);
/* a6: getNewEnergy (VertexSmoother.cpp:544-625): sum of energies over each ring in stored vertex order, clamped to
 * MAX_ENERGY when inf / NaN / <= 0 / > MAX_ENERGY (:619-622) */
int twg_amips_ring_energy(twg_ctx* ctx, const double* Vxyz, uint32_t nV, const int32_t* tets4, uint64_t nT, const int32_t* t_ids,
                          const uint64_t* group_off, uint64_t nGroups, double* E);
int twg_amips_ring_energy_dev(twg_ctx* ctx, const double* dVxyz, uint32_t nV, const int32_t* dTets4, uint64_t nT,
                              const int32_t* dTids, const uint64_t* dGroupOff, uint64_t nGroups, double* dE, void* stream);

/* ---- resident tet mesh (S3 callers: whole-mesh passes and one-ring batches without re-shipping the mesh) ---------- */
/* Device mirror of the scheduler's `tet_vertices[].posf`, `tets`, `t_is_removed` and `tet_vertices[].conn_tets`
 * (src/tetwild/LocalOperations.h:35-45). Created once after the front end (MeshRefinement.cpp:208), then kept in step
 * with accepted operations through the two scatter calls; AMIPS batches then move only ids in and results out.
 * A removed tet (t_is_removed[t]) is a tet whose first index is negative. Slots only grow, like the host vectors. */
typedef struct twg_mesh twg_mesh;
int twg_mesh_create(twg_ctx* ctx, const double* Vxyz, uint32_t nV, const int32_t* tets4, uint64_t nT, twg_mesh** out);
void twg_mesh_destroy(twg_mesh* m);
uint32_t twg_mesh_num_vertices(const twg_mesh* m);
uint64_t twg_mesh_num_tets(const twg_mesh* m);
const double* twg_mesh_vertices_dev(const twg_mesh* m);  /* device pointers, valid until the next resize */
const int32_t* twg_mesh_tets_dev(const twg_mesh* m);
int twg_mesh_resize(twg_mesh* m, uint32_t nV, uint64_t nT);  /* new vertex slots are 0, new tet slots are removed */
int twg_mesh_set_vertices(twg_mesh* m, const int32_t* v_ids, const double* xyz, uint64_t n);    /* posf[v_ids[i]] = xyz[i] */
int twg_mesh_set_tets(twg_mesh* m, const int32_t* t_ids, const int32_t* tets4, uint64_t n);     /* tets[t_ids[i]] = tets4[i] */
int twg_mesh_get_vertices(twg_mesh* m, double* xyz_out /* nV*3 */);
/* conn_tets rebuilt on the device (vertex -> incident live tets, ascending tet id); get_rings copies it back:
 * off_out[nV+1], tets_out[off_out[nV]] (tets_out may be NULL to query sizes) */
int twg_mesh_build_rings(twg_mesh* m);
int twg_mesh_get_rings(twg_mesh* m, uint64_t* off_out, int32_t* tets_out);
/* a4 over the resident mesh: calTetQuality_AMIPS (LocalOperations.cpp:862-884) of tets t_ids[0..n) (t_ids NULL: all
 * nT slots; removed tets give MAX_ENERGY). The whole-mesh callers: VertexSmoother.cpp:216-241, MeshRefinement.cpp:51 */
int twg_mesh_quality(twg_mesh* m, const int32_t* t_ids, uint64_t n, double* slim_energy);
int twg_mesh_quality_dev(twg_mesh* m, const int32_t* dTids, uint64_t n, double* dSlim, void* stream);
/* calTetQuality_AD (LocalOperations.cpp:783-860): min / max dihedral angle per tet (0 and pi for degenerate tets),
 * what LocalOperations::outputInfo (:348-354) and getFilteredAngles collect over all live tets */
int twg_mesh_dihedral(twg_mesh* m, const int32_t* t_ids, uint64_t n, double* min_d_angle, double* max_d_angle);
int twg_mesh_dihedral_dev(twg_mesh* m, const int32_t* dTids, uint64_t n, double* dMin, double* dMax, void* stream);
/* a5 (NewtonsUpdate, VertexSmoother.cpp:627-702) for the one-rings of vertices v_ids[0..n): members are conn_tets[v] */
int twg_mesh_vertex_ring_ejh(twg_mesh* m, const int32_t* v_ids, uint64_t n, double* E, double* J3, double* H9, uint8_t* ok);
int twg_mesh_vertex_ring_ejh_dev(twg_mesh* m, const int32_t* dVids, uint64_t n, double* dE, double* dJ3, double* dH9, uint8_t* dOk,
                                 void* stream);
/* a5 / a6 with explicit member lists (CSR over t_ids) against the resident mesh */
int twg_mesh_ring_ejh(twg_mesh* m, const int32_t* t_ids, const uint64_t* group_off, const int32_t* center, uint64_t nGroups,
                      double* E, double* J3, double* H9, uint8_t* ok);
int twg_mesh_ring_energy(twg_mesh* m, const int32_t* t_ids, const uint64_t* group_off, uint64_t nGroups, double* E);
/* the smoother's line search (VertexSmoother.cpp:505-541: move v, getNewEnergy(conn_tets[v]), move it back) for n (vertex, trial
 * position) pairs at once: E[i] = getNewEnergy of the one-ring of v_ids[i] with that vertex at xyz[i]. The resident mesh is not
 * modified, so the pairs are independent: all step sizes of one Newton step, or the steps of many candidates, in one call. */
int twg_mesh_vertex_trial_energy(twg_mesh* m, const int32_t* v_ids, const double* xyz, uint64_t n, double* E);
int twg_mesh_vertex_trial_energy_dev(twg_mesh* m, const int32_t* dVids, const double* dXyz, uint64_t n, double* dE, void* stream);

/* ---- S4: generalized winding number ------------------------------------------------------------------------------ */
/* F may contain repeated faces (InoutFiltering.cpp:99-100). */
int twg_winding_create(twg_ctx* ctx, const double* V, uint32_t nV, const uint32_t* F, uint32_t nF, twg_winding** out);
void twg_winding_destroy(twg_winding* w);
twg_winding* twg_winding_replica(twg_winding* w, int k);
/* W[i] and/or keep[i] = (W[i] > 0.5) for query i (either may be NULL) */
int twg_winding_eval(twg_winding* w, const double* C, uint64_t nC, double* W, uint8_t* keep);
int twg_winding_eval_dev(twg_winding* w, const double* dC, uint64_t nC, double* dW, uint8_t* dKeep, void* stream);
/* igl::winding_number(V,F,O,W) one-shot mirror (build + eval + free) */
int twg_winding_number(twg_ctx* ctx, const double* V, uint32_t nV, const uint32_t* F, uint32_t nF, const double* C,
                       uint64_t nC, double* W, uint8_t* keep);
/* InoutFiltering::filter decision (InoutFiltering.cpp:45-75): keep = W > 0.5; if nothing is kept, faces are flipped
 * (columns 1,2 swapped) and the test repeated; *retried tells which happened */
int twg_inout_filter(twg_ctx* ctx, const double* V, uint32_t nV, const uint32_t* F, uint32_t nF, const double* C,
                     uint64_t nC, uint8_t* keep, int* retried);
/* test hook: the three arrays of the hierarchy as they live on the device. nodes_out: n_nodes * 64 bytes, caps_out: n_cap_points * 4
 * doubles, tris_out: (n_triangles + 2) * 9 doubles (sizes from twg_winding_stats; any pointer may be NULL) */
int twg_debug_winding_download(twg_winding* w, void* nodes_out, double* caps_out, double* tris_out);
/* statistics of the hierarchy: number of nodes, total cap segments, leaf triangles */
int twg_winding_stats(const twg_winding* w, uint64_t* n_nodes, uint64_t* n_cap_segments, uint64_t* n_triangles);

#ifdef __cplusplus
}
#endif
#endif /* TETWILD_GPU_H */
