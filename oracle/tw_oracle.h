/*
 * oracle/tw_oracle.h -- C interface of the CPU oracle.
 *
 * TEST INFRASTRUCTURE ONLY. The oracle restates, on the CPU, the reference algorithms of TetWild's data-parallel
 * hot path. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load it -- as the checker or the timed CPU baseline, never as a product code path.
 *
 * Pinning status (see DESIGN.md "Oracle"):
 *   AMIPS (a1-a7)      pinned: checked against the reference's own text compiled into oracle/_ref (ref_build.sh)
 *                      and against the committed golden vectors tests/golden/amips_golden.json.
 *   sampleTriangle(a8) pinned against the reference's own text compiled over a vec3 shim (oracle/_ref).
 *   AABB tree (a11-13) pinned against the reference's mesh_AABB.cpp compiled unmodified over a geogram API shim.
 *   point-triangle distance (a14, geogram b613750), winding number (a16, libigl 45cfc79), CGAL predicates:
 *                      third-party code NOT under /root/reference -> restated from the published algorithms;
 *                      "parity unpinned" for these three leaf routines.
 */
#ifndef TW_ORACLE_H
#define TW_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORA_MAX_EXPANSION 96
#define ORA_MAX_ENERGY 1e50      /* reference: src/tetwild/State.h:29 */
#define ORA_NO_FACET 0xffffffffu /* GEO::NO_FACET */

/* ---- exact predicates (predicates.c) ---- */
int ora_orient3d_exact(const double *a, const double *b, const double *c, const double *d);
int ora_orient3d(const double *a, const double *b, const double *c, const double *d);
int ora_cgal_orientation(const double *p, const double *q, const double *r, const double *s);
int ora_triangle_is_degenerate(const double *p, const double *q, const double *r);

/* ---- AMIPS (amips.c) ---- */
double ora_amips_energy(const double *T12);
void ora_amips_jacobian(const double *T12, double *J3);
void ora_amips_hessian(const double *T12, double *H9);
void ora_amips_energy_soa(const double *const T[12], double *E, uint64_t n, int threads);
void ora_amips_ejh_soa(const double *const T[12], double *E, double *J3, double *H9, uint64_t n, int threads);
void ora_amips_quality(const double *Vxyz, const int32_t *tets4, uint64_t nT, double *slim_energy, int threads);
void ora_amips_ring_ejh(const double *Vxyz, const int32_t *tets4, const int32_t *t_ids, const uint64_t *group_off,
                        const int32_t *center, uint64_t nGroups, double *E, double *J3, double *H9, uint8_t *ok,
                        int threads);
void ora_amips_ring_energy(const double *Vxyz, const int32_t *tets4, const int32_t *t_ids, const uint64_t *group_off,
                           uint64_t nGroups, double *E, int threads);

/* calTetQuality_AD (LocalOperations.cpp:783-860); a tet with a negative first index is a removed slot -> (0, pi) */
void ora_tet_dihedral(const double *Vxyz, const int32_t *tets4, uint64_t nT, double *dmin, double *dmax, int threads);

/* ---- envelope (envelope.c) ---- */
typedef struct ora_surface ora_surface;
/* order: 0 = keep caller's facet order, 1 = sort facets along a Morton curve (stand-in for geogram mesh_reorder) */
ora_surface *ora_surface_create(const double *V, uint32_t nV, const uint32_t *F, uint32_t nF, int order);
void ora_surface_destroy(ora_surface *s);
uint32_t ora_surface_num_facets(const ora_surface *s);
void ora_surface_get_order(const ora_surface *s, uint32_t *orig); /* tree position -> caller's facet id */
/* facet ids below are in the CALLER's original numbering */
double ora_point_triangle_sqdist(const double *p, const double *v0, const double *v1, const double *v2,
                                 double *nearest3);
void ora_nearest(const ora_surface *s, const double *P, uint64_t n, uint32_t *facet, double *nearest_xyz, double *d2,
                 int threads);
void ora_envelope_points_out(const ora_surface *s, const double *P, uint64_t n, double eps2, uint8_t *out, int threads);
void ora_envelope_points_out_brute(const ora_surface *s, const double *P, uint64_t n, double eps2, uint8_t *out,
                                   int threads);
void ora_point_sqdist(const ora_surface *s, const double *P, uint64_t n, double *d2, int threads);
void ora_point_sqdist_brute(const ora_surface *s, const double *P, uint64_t n, double *d2, uint32_t *facet, int threads);
/* sampleTriangle: writes up to cap samples, returns the number the reference would generate */
uint64_t ora_sample_triangle(const double *tri9, double sampling_dist, double *samples_xyz, uint64_t cap);
void ora_envelope_faces_out(const ora_surface *s, const double *tris9, uint64_t n, double sampling_dist, double eps2,
                            uint8_t *out, uint64_t *num_samples, int threads);

void ora_envelope_faces_out_ex(const ora_surface *s, const double *tris9, uint64_t n, double sampling_dist, double eps2,
                               int degenerate_shortcut, uint8_t *out, uint64_t *num_samples, int threads);

/* ---- winding number (winding.c) ---- */
double ora_solid_angle_w(const double *a, const double *b, const double *c, const double *p);
void ora_winding_direct(const double *V, uint32_t nV, const uint32_t *F, uint32_t nF, const double *C, uint64_t nC,
                        double *W, int threads);
typedef struct ora_wtree ora_wtree;
ora_wtree *ora_wtree_create(const double *V, uint32_t nV, const uint32_t *F, uint32_t nF);
void ora_wtree_destroy(ora_wtree *t);
void ora_wtree_eval(const ora_wtree *t, const double *C, uint64_t nC, double *W, int threads);
uint64_t ora_wtree_stats(const ora_wtree *t, uint64_t *n_nodes, uint64_t *cap_faces_total);
/* InoutFiltering::filter restated: keep[i] = W>0.5, with the flip-and-retry rule; returns 1 if the retry was taken */
int ora_inout_filter(const double *V, uint32_t nV, const uint32_t *F, uint32_t nF, const double *C, uint64_t nC,
                     uint8_t *keep, double *W, int hierarchical, int threads);

int ora_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
