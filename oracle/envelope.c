/*
 * oracle/envelope.c -- TEST INFRASTRUCTURE ONLY (CPU oracle; never on the product path).
 *
 * Restates the reference's sampled-point envelope test:
 *   tree build         src/tetwild/geogram/mesh_AABB.cpp:63-141, :356-379   (implicit balanced bbox tree)
 *   box distances      src/tetwild/geogram/mesh_AABB.cpp:180-238
 *   hint               src/tetwild/geogram/mesh_AABB.cpp:381-416
 *   nearest facet      src/tetwild/geogram/mesh_AABB.cpp:418-480, mesh_AABB.h:130-176,221-226
 *   eps early exit     src/tetwild/geogram/mesh_AABB.cpp:482-548, mesh_AABB.h:182-213
 *   sampleTriangle     src/tetwild/Common.cpp:143-255
 *   face / point test  src/tetwild/LocalOperations.cpp:1034-1109, src/tetwild/DistanceQuery.h:20-39
 *
 * Third-party arithmetic NOT under /root/reference (geogram fork b613750, cmake/TetWildDownloadExternal.cmake:26-29):
 *   GEO::Geom::point_triangle_squared_distance / point_segment_squared_distance, vec3 operators, normalize(),
 *   mesh_reorder(MESH_ORDER_MORTON). These are restated from the published algorithm (D. Eberly, "Distance Between
 *   Point and Triangle in 3D", the 7-region minimisation of the squared-distance quadratic) -- PARITY UNPINNED for the
 *   leaf arithmetic. Facet order only changes traversal cost, not results (any spatial sort is acceptable).
 *
 * All arithmetic is plain IEEE double without FMA contraction, matching a default x86-64 build of the reference
 * (no -march / -ffast-math flags in CMakeLists.txt:54-56, cmake/Warnings.cmake).
 */
#include <math.h>
#include <float.h>
#include <stdlib.h>
#include <string.h>
#include "tw_oracle.h"

#pragma STDC FP_CONTRACT OFF

typedef struct { double lo[3], hi[3]; } box_t;

struct ora_surface {
    uint32_t nV, nF;
    double *V;      /* nV*3 */
    uint32_t *F;    /* nF*3, in tree order */
    uint32_t *orig; /* tree position -> caller's facet id */
    box_t *boxes;   /* index 1..max_node */
    uint32_t n_boxes;
};

/* ---- vec3 helpers with geogram's operation order (RECOLLECTED; see header) ---- */
static inline double v_dot(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static inline double v_len2(const double *a) { return a[0] * a[0] + a[1] * a[1] + a[2] * a[2]; }
static inline void v_sub(const double *a, const double *b, double *r) { r[0] = a[0] - b[0]; r[1] = a[1] - b[1]; r[2] = a[2] - b[2]; }
static inline void v_cross(const double *a, const double *b, double *r) {
    r[0] = a[1] * b[2] - a[2] * b[1];
    r[1] = a[2] * b[0] - a[0] * b[2];
    r[2] = a[0] * b[1] - a[1] * b[0];
}
static inline double v_dist2(const double *a, const double *b) { double d[3]; v_sub(b, a, d); return v_len2(d); }
static inline double v_dist(const double *a, const double *b) { return sqrt(v_dist2(a, b)); }
static inline void v_normalize(const double *a, double *r) {
    double s = sqrt(v_len2(a));
    if (s > 1e-30) s = 1.0 / s;
    r[0] = s * a[0]; r[1] = s * a[1]; r[2] = s * a[2];
}

/* ---- point-segment and point-triangle squared distance (Eberly) ---- */
static double point_segment_sqdist(const double *p, const double *v0, const double *v1, double *nearest) {
    double l2 = v_dist2(v0, v1);
    double d0[3], d1[3];
    v_sub(p, v0, d0);
    v_sub(v1, v0, d1);
    double t = v_dot(d0, d1);
    if (t <= 0.0 || l2 == 0.0) {
        memcpy(nearest, v0, 24);
        return v_dist2(p, v0);
    } else if (t > l2) {
        memcpy(nearest, v1, 24);
        return v_dist2(p, v1);
    }
    double l1 = t / l2;
    double l0 = 1.0 - l1;
    for (int k = 0; k < 3; ++k) nearest[k] = l0 * v0[k] + l1 * v1[k];
    return v_dist2(p, nearest);
}

double ora_point_triangle_sqdist(const double *p, const double *V0, const double *V1, const double *V2, double *nearest) {
    double diff[3], e0[3], e1[3];
    v_sub(V0, p, diff);
    v_sub(V1, V0, e0);
    v_sub(V2, V0, e1);
    double a00 = v_len2(e0), a01 = v_dot(e0, e1), a11 = v_len2(e1);
    double b0 = v_dot(diff, e0), b1 = v_dot(diff, e1), c = v_len2(diff);
    double det = fabs(a00 * a11 - a01 * a01);
    double s = a01 * b1 - a11 * b0;
    double t = a01 * b0 - a00 * b1;
    double d2;

    if (det < 1e-30) { /* degenerate triangle: nearest of the three edges */
        double cur[3];
        double best = point_segment_sqdist(p, V0, V1, nearest);
        double d = point_segment_sqdist(p, V0, V2, cur);
        if (d < best) { best = d; memcpy(nearest, cur, 24); }
        d = point_segment_sqdist(p, V1, V2, cur);
        if (d < best) { best = d; memcpy(nearest, cur, 24); }
        return best;
    }

    if (s + t <= det) {
        if (s < 0.0) {
            if (t < 0.0) { /* region 4 */
                if (b0 < 0.0) {
                    t = 0.0;
                    if (-b0 >= a00) { s = 1.0; d2 = a00 + 2.0 * b0 + c; }
                    else { s = -b0 / a00; d2 = b0 * s + c; }
                } else {
                    s = 0.0;
                    if (b1 >= 0.0) { t = 0.0; d2 = c; }
                    else if (-b1 >= a11) { t = 1.0; d2 = a11 + 2.0 * b1 + c; }
                    else { t = -b1 / a11; d2 = b1 * t + c; }
                }
            } else { /* region 3 */
                s = 0.0;
                if (b1 >= 0.0) { t = 0.0; d2 = c; }
                else if (-b1 >= a11) { t = 1.0; d2 = a11 + 2.0 * b1 + c; }
                else { t = -b1 / a11; d2 = b1 * t + c; }
            }
        } else if (t < 0.0) { /* region 5 */
            t = 0.0;
            if (b0 >= 0.0) { s = 0.0; d2 = c; }
            else if (-b0 >= a00) { s = 1.0; d2 = a00 + 2.0 * b0 + c; }
            else { s = -b0 / a00; d2 = b0 * s + c; }
        } else { /* region 0: interior */
            double inv = 1.0 / det;
            s *= inv;
            t *= inv;
            d2 = s * (a00 * s + a01 * t + 2.0 * b0) + t * (a01 * s + a11 * t + 2.0 * b1) + c;
        }
    } else {
        double tmp0, tmp1, numer, denom;
        if (s < 0.0) { /* region 2 */
            tmp0 = a01 + b0;
            tmp1 = a11 + b1;
            if (tmp1 > tmp0) {
                numer = tmp1 - tmp0;
                denom = a00 - 2.0 * a01 + a11;
                if (numer >= denom) { s = 1.0; t = 0.0; d2 = a00 + 2.0 * b0 + c; }
                else {
                    s = numer / denom; t = 1.0 - s;
                    d2 = s * (a00 * s + a01 * t + 2.0 * b0) + t * (a01 * s + a11 * t + 2.0 * b1) + c;
                }
            } else {
                s = 0.0;
                if (tmp1 <= 0.0) { t = 1.0; d2 = a11 + 2.0 * b1 + c; }
                else if (b1 >= 0.0) { t = 0.0; d2 = c; }
                else { t = -b1 / a11; d2 = b1 * t + c; }
            }
        } else if (t < 0.0) { /* region 6 */
            tmp0 = a01 + b1;
            tmp1 = a00 + b0;
            if (tmp1 > tmp0) {
                numer = tmp1 - tmp0;
                denom = a00 - 2.0 * a01 + a11;
                if (numer >= denom) { t = 1.0; s = 0.0; d2 = a11 + 2.0 * b1 + c; }
                else {
                    t = numer / denom; s = 1.0 - t;
                    d2 = s * (a00 * s + a01 * t + 2.0 * b0) + t * (a01 * s + a11 * t + 2.0 * b1) + c;
                }
            } else {
                t = 0.0;
                if (tmp1 <= 0.0) { s = 1.0; d2 = a00 + 2.0 * b0 + c; }
                else if (b0 >= 0.0) { s = 0.0; d2 = c; }
                else { s = -b0 / a00; d2 = b0 * s + c; }
            }
        } else { /* region 1 */
            numer = a11 + b1 - a01 - b0;
            if (numer <= 0.0) { s = 0.0; t = 1.0; d2 = a11 + 2.0 * b1 + c; }
            else {
                denom = a00 - 2.0 * a01 + a11;
                if (numer >= denom) { s = 1.0; t = 0.0; d2 = a00 + 2.0 * b0 + c; }
                else {
                    s = numer / denom; t = 1.0 - s;
                    d2 = s * (a00 * s + a01 * t + 2.0 * b0) + t * (a01 * s + a11 * t + 2.0 * b1) + c;
                }
            }
        }
    }
    if (d2 < 0.0) d2 = 0.0; /* round-off guard */
    for (int k = 0; k < 3; ++k) nearest[k] = V0[k] + s * e0[k] + t * e1[k];
    return d2;
}

/* ---- tree ---- */
static uint32_t max_node_index(uint32_t n, uint32_t b, uint32_t e) { /* mesh_AABB.cpp:88-100 */
    if (b + 1 == e) return n;
    uint32_t m = b + (e - b) / 2;
    uint32_t l = max_node_index(2 * n, b, m), r = max_node_index(2 * n + 1, m, e);
    return l > r ? l : r;
}

static void facet_box(const ora_surface *s, uint32_t f, box_t *B) { /* mesh_AABB.cpp:63-79 */
    for (int c = 0; c < 3; ++c) { B->lo[c] = DBL_MAX; B->hi[c] = -DBL_MAX; }
    for (int k = 0; k < 3; ++k) {
        const double *p = s->V + 3 * (size_t)s->F[3 * (size_t)f + k];
        for (int c = 0; c < 3; ++c) {
            if (p[c] < B->lo[c]) B->lo[c] = p[c];
            if (p[c] > B->hi[c]) B->hi[c] = p[c];
        }
    }
}

static void init_boxes(ora_surface *s, uint32_t n, uint32_t b, uint32_t e) { /* mesh_AABB.cpp:119-141 */
    if (b + 1 == e) { facet_box(s, b, &s->boxes[n]); return; }
    uint32_t m = b + (e - b) / 2;
    init_boxes(s, 2 * n, b, m);
    init_boxes(s, 2 * n + 1, m, e);
    for (int c = 0; c < 3; ++c) {
        s->boxes[n].lo[c] = fmin(s->boxes[2 * n].lo[c], s->boxes[2 * n + 1].lo[c]);
        s->boxes[n].hi[c] = fmax(s->boxes[2 * n].hi[c], s->boxes[2 * n + 1].hi[c]);
    }
}

typedef struct { uint64_t code; uint32_t id; } mkey;
static int mkey_cmp(const void *a, const void *b) {
    const mkey *x = (const mkey *)a, *y = (const mkey *)b;
    if (x->code != y->code) return x->code < y->code ? -1 : 1;
    return x->id < y->id ? -1 : (x->id > y->id);
}
static uint64_t spread3(uint64_t v) { /* 21 bits -> every third bit */
    v &= 0x1fffffULL;
    v = (v | v << 32) & 0x1f00000000ffffULL;
    v = (v | v << 16) & 0x1f0000ff0000ffULL;
    v = (v | v << 8) & 0x100f00f00f00f00fULL;
    v = (v | v << 4) & 0x10c30c30c30c30c3ULL;
    v = (v | v << 2) & 0x1249249249249249ULL;
    return v;
}

ora_surface *ora_surface_create(const double *V, uint32_t nV, const uint32_t *F, uint32_t nF, int order) {
    if (nF == 0) return NULL;
    ora_surface *s = (ora_surface *)calloc(1, sizeof(*s));
    s->nV = nV; s->nF = nF;
    s->V = (double *)malloc(sizeof(double) * 3 * (size_t)nV);
    memcpy(s->V, V, sizeof(double) * 3 * (size_t)nV);
    s->F = (uint32_t *)malloc(sizeof(uint32_t) * 3 * (size_t)nF);
    s->orig = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)nF);
    if (order) {
        double lo[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, hi[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
        for (uint32_t f = 0; f < nF; ++f)
            for (int k = 0; k < 3; ++k)
                for (int c = 0; c < 3; ++c) {
                    double x = V[3 * (size_t)F[3 * (size_t)f + k] + c];
                    if (x < lo[c]) lo[c] = x;
                    if (x > hi[c]) hi[c] = x;
                }
        mkey *keys = (mkey *)malloc(sizeof(mkey) * (size_t)nF);
        for (uint32_t f = 0; f < nF; ++f) {
            uint64_t code = 0;
            for (int c = 0; c < 3; ++c) {
                double ctr = 0;
                for (int k = 0; k < 3; ++k) ctr += V[3 * (size_t)F[3 * (size_t)f + k] + c];
                ctr /= 3.0;
                double ext = hi[c] - lo[c];
                double u = ext > 0 ? (ctr - lo[c]) / ext : 0.0;
                if (u < 0) u = 0;
                if (u > 1) u = 1;
                uint64_t q = (uint64_t)(u * 2097151.0);
                code |= spread3(q) << c;
            }
            keys[f].code = code; keys[f].id = f;
        }
        qsort(keys, nF, sizeof(mkey), mkey_cmp);
        for (uint32_t i = 0; i < nF; ++i) s->orig[i] = keys[i].id;
        free(keys);
    } else {
        for (uint32_t i = 0; i < nF; ++i) s->orig[i] = i;
    }
    for (uint32_t i = 0; i < nF; ++i) memcpy(s->F + 3 * (size_t)i, F + 3 * (size_t)s->orig[i], 12);
    s->n_boxes = max_node_index(1, 0, nF) + 1; /* mesh_AABB.cpp:371-375 */
    s->boxes = (box_t *)malloc(sizeof(box_t) * (size_t)s->n_boxes);
    init_boxes(s, 1, 0, nF);
    return s;
}

void ora_surface_destroy(ora_surface *s) {
    if (!s) return;
    free(s->V); free(s->F); free(s->orig); free(s->boxes); free(s);
}
uint32_t ora_surface_num_facets(const ora_surface *s) { return s->nF; }
void ora_surface_get_order(const ora_surface *s, uint32_t *orig) { memcpy(orig, s->orig, sizeof(uint32_t) * (size_t)s->nF); }

static inline double sqr(double x) { return x * x; }

static double inner_box_sqdist(const double *p, const box_t *B) { /* mesh_AABB.cpp:180-192 */
    double r = sqr(p[0] - B->lo[0]);
    r = fmin(r, sqr(p[0] - B->hi[0]));
    for (int c = 1; c < 3; ++c) {
        r = fmin(r, sqr(p[c] - B->lo[c]));
        r = fmin(r, sqr(p[c] - B->hi[c]));
    }
    return r;
}
static double box_signed_sqdist(const double *p, const box_t *B) { /* mesh_AABB.cpp:201-220 */
    int inside = 1;
    double r = 0.0;
    for (int c = 0; c < 3; ++c) {
        if (p[c] < B->lo[c]) { inside = 0; r += sqr(p[c] - B->lo[c]); }
        else if (p[c] > B->hi[c]) { inside = 0; r += sqr(p[c] - B->hi[c]); }
    }
    if (inside) r = -inner_box_sqdist(p, B);
    return r;
}
static double box_center_sqdist(const double *p, const box_t *B) { /* mesh_AABB.cpp:229-238 */
    double r = 0.0;
    for (int c = 0; c < 3; ++c) {
        double d = p[c] - 0.5 * (B->lo[c] + B->hi[c]);
        r += sqr(d);
    }
    return r;
}

static void facet_nearest(const ora_surface *s, const double *p, uint32_t f, double *np, double *d2) {
    const uint32_t *t = s->F + 3 * (size_t)f; /* mesh_AABB.cpp:153-171 */
    *d2 = ora_point_triangle_sqdist(p, s->V + 3 * (size_t)t[0], s->V + 3 * (size_t)t[1], s->V + 3 * (size_t)t[2], np);
}

static void nearest_hint(const ora_surface *s, const double *p, uint32_t *nf, double *np, double *d2) {
    uint32_t b = 0, e = s->nF, n = 1; /* mesh_AABB.cpp:381-416 */
    while (e != b + 1) {
        uint32_t m = b + (e - b) / 2;
        if (box_center_sqdist(p, &s->boxes[2 * n]) < box_center_sqdist(p, &s->boxes[2 * n + 1])) { e = m; n = 2 * n; }
        else { b = m; n = 2 * n + 1; }
    }
    *nf = b;
    const double *v = s->V + 3 * (size_t)s->F[3 * (size_t)b];
    memcpy(np, v, 24);
    *d2 = v_dist2(p, np);
}

static void nearest_rec(const ora_surface *s, const double *p, uint32_t *nf, double *np, double *d2, uint32_t n,
                        uint32_t b, uint32_t e) { /* mesh_AABB.cpp:418-480 */
    if (b + 1 == e) {
        double cp[3], cd;
        facet_nearest(s, p, b, cp, &cd);
        if (cd < *d2) { *nf = b; memcpy(np, cp, 24); *d2 = cd; }
        return;
    }
    uint32_t m = b + (e - b) / 2, l = 2 * n, r = 2 * n + 1;
    double dl = box_signed_sqdist(p, &s->boxes[l]);
    double dr = box_signed_sqdist(p, &s->boxes[r]);
    if (dl < dr) {
        if (dl < *d2) nearest_rec(s, p, nf, np, d2, l, b, m);
        if (dr < *d2) nearest_rec(s, p, nf, np, d2, r, m, e);
    } else {
        if (dr < *d2) nearest_rec(s, p, nf, np, d2, r, m, e);
        if (dl < *d2) nearest_rec(s, p, nf, np, d2, l, b, m);
    }
}

static void envelope_rec(const ora_surface *s, const double *p, double eps2, uint32_t *nf, double *np, double *d2,
                         uint32_t n, uint32_t b, uint32_t e) { /* mesh_AABB.cpp:482-548 */
    if (*d2 <= eps2) return;
    if (b + 1 == e) {
        double cp[3], cd;
        facet_nearest(s, p, b, cp, &cd);
        if (cd < *d2) { *nf = b; memcpy(np, cp, 24); *d2 = cd; }
        return;
    }
    uint32_t m = b + (e - b) / 2, l = 2 * n, r = 2 * n + 1;
    double dl = box_signed_sqdist(p, &s->boxes[l]);
    double dr = box_signed_sqdist(p, &s->boxes[r]);
    if (dl < dr) {
        if (dl < *d2 && dl <= eps2) envelope_rec(s, p, eps2, nf, np, d2, l, b, m);
        if (dr < *d2 && dr <= eps2) envelope_rec(s, p, eps2, nf, np, d2, r, m, e);
    } else {
        if (dr < *d2 && dr <= eps2) envelope_rec(s, p, eps2, nf, np, d2, r, m, e);
        if (dl < *d2 && dl <= eps2) envelope_rec(s, p, eps2, nf, np, d2, l, b, m);
    }
}

/* MeshFacetsAABBWithEps::nearest_facet, mesh_AABB.h:130-141 */
static uint32_t tree_nearest(const ora_surface *s, const double *p, double *np, double *d2) {
    uint32_t nf;
    nearest_hint(s, p, &nf, np, d2);
    nearest_rec(s, p, &nf, np, d2, 1, 0, s->nF);
    return nf;
}
/* facet_in_envelope_with_hint, mesh_AABB.h:199-213 */
static void tree_envelope_with_hint(const ora_surface *s, const double *p, double eps2, uint32_t *nf, double *np, double *d2) {
    if (*nf == ORA_NO_FACET) nearest_hint(s, p, nf, np, d2);
    envelope_rec(s, p, eps2, nf, np, d2, 1, 0, s->nF);
}

void ora_nearest(const ora_surface *s, const double *P, uint64_t n, uint32_t *facet, double *nearest, double *d2, int threads) {
    (void)threads;
#pragma omp parallel for schedule(dynamic, 1024) num_threads(threads > 0 ? threads : 1)
    for (int64_t i = 0; i < (int64_t)n; ++i) {
        double np[3], d;
        uint32_t f = tree_nearest(s, P + 3 * i, np, &d);
        if (facet) facet[i] = s->orig[f];
        if (nearest) memcpy(nearest + 3 * i, np, 24);
        if (d2) d2[i] = d;
    }
}

/* isPointOutEnvelop: LocalOperations.cpp:1034-1044 applies squared_distance() > eps_2 (full nearest search);
 * the per-sample query of isFaceOutEnvelop_sampling (:1083-1086) uses facet_in_envelope_with_hint with no previous
 * facet. Both give the same decision; this entry point uses the early-exit form (the C2 workload of BASELINE.json). */
void ora_envelope_points_out(const ora_surface *s, const double *P, uint64_t n, double eps2, uint8_t *out, int threads) {
    (void)threads;
#pragma omp parallel for schedule(dynamic, 1024) num_threads(threads > 0 ? threads : 1)
    for (int64_t i = 0; i < (int64_t)n; ++i) {
        double np[3], d = DBL_MAX;
        uint32_t f = ORA_NO_FACET;
        tree_envelope_with_hint(s, P + 3 * i, eps2, &f, np, &d);
        out[i] = d > eps2;
    }
}

void ora_point_sqdist(const ora_surface *s, const double *P, uint64_t n, double *d2, int threads) {
    ora_nearest(s, P, n, NULL, NULL, d2, threads); /* squared_distance(): mesh_AABB.h:221-226 */
}

void ora_point_sqdist_brute(const ora_surface *s, const double *P, uint64_t n, double *d2, uint32_t *facet, int threads) {
    (void)threads;
#pragma omp parallel for schedule(dynamic, 64) num_threads(threads > 0 ? threads : 1)
    for (int64_t i = 0; i < (int64_t)n; ++i) {
        double best = DBL_MAX, np[3];
        uint32_t bf = 0;
        for (uint32_t f = 0; f < s->nF; ++f) {
            double d;
            facet_nearest(s, P + 3 * i, f, np, &d);
            if (d < best) { best = d; bf = f; }
        }
        d2[i] = best;
        if (facet) facet[i] = s->orig[bf];
    }
}

void ora_envelope_points_out_brute(const ora_surface *s, const double *P, uint64_t n, double eps2, uint8_t *out, int threads) {
    (void)threads;
#pragma omp parallel for schedule(dynamic, 64) num_threads(threads > 0 ? threads : 1)
    for (int64_t i = 0; i < (int64_t)n; ++i) {
        double best = DBL_MAX, np[3];
        for (uint32_t f = 0; f < s->nF; ++f) {
            double d;
            facet_nearest(s, P + 3 * i, f, np, &d);
            if (d < best) best = d;
        }
        out[i] = best > eps2;
    }
}

/* ---- sampleTriangle: Common.cpp:143-255 ---- */
typedef struct { double *buf; uint64_t cap, n; } sink_t;
static inline void push(sink_t *k, const double *p) {
    if (k->buf && k->n < k->cap) memcpy(k->buf + 3 * k->n, p, 24);
    k->n++;
}
static inline void push3(sink_t *k, double x, double y, double z) { double p[3] = {x, y, z}; push(k, p); }

static void sample_triangle(const double vs[3][3], sink_t *ps, double sd) {
    double sqrt3_2 = sqrt(3) / 2; /* :144 (std::sqrt(int) -> double) */
    double ls[3];
    for (int i = 0; i < 3; ++i) { double d[3]; v_sub(vs[i], vs[(i + 1) % 3], d); ls[i] = v_len2(d); } /* :146-149 */
    int max_i = 0; /* std::minmax_element: first smallest / LAST largest (:150-152) */
    for (int i = 1; i < 3; ++i) if (!(ls[i] < ls[max_i])) max_i = i;
    double N = sqrt(ls[max_i]) / sd; /* :153 */
    if (N <= 1) { for (int i = 0; i < 3; ++i) push(ps, vs[i]); return; } /* :154-158 */
    if (N == (int)N) N -= 1; /* :159-160 */
    const double *v0 = vs[max_i], *v1 = vs[(max_i + 1) % 3], *v2 = vs[(max_i + 2) % 3];
    double e01[3], n01[3];
    v_sub(v1, v0, e01);
    v_normalize(e01, n01); /* :166 */
    for (int n = 0; n <= N; ++n) /* :167-169: v0 + (n_v0v1 * sd) * n */
        push3(ps, v0[0] + n01[0] * sd * n, v0[1] + n01[1] * sd * n, v0[2] + n01[2] * sd * n);
    push(ps, v1); /* :170 */

    double e02[3];
    v_sub(v2, v0, e02);
    double dt = v_dot(e02, e01);
    double foot[3];
    for (int k = 0; k < 3; ++k) foot[k] = dt * e01[k] / ls[max_i] + v0[k]; /* :172 */
    double h = v_dist(foot, v2);
    int M = (int)(h / (sqrt3_2 * sd)); /* :173 */
    if (M < 1) { push(ps, v2); return; } /* :174-177 */

    double n02[3], e12[3], n12[3], e10[3];
    v_normalize(e02, n02); /* :179 */
    v_sub(v2, v1, e12);
    v_normalize(e12, n12); /* :180 */
    v_sub(v0, v1, e10);
    double c0[3], c1[3];
    v_cross(e02, e01, c0);
    v_cross(e12, e10, c1);
    double sin_v0 = sqrt(v_len2(c0)) / (v_dist(v0, v2) * v_dist(v0, v1)); /* :182 */
    double tan_v0 = sqrt(v_len2(c0)) / v_dot(e02, e01);                   /* :183 */
    double tan_v1 = sqrt(v_len2(c1)) / v_dot(e12, e10);                   /* :184 (unused, as in the reference) */
    double sin_v1 = sqrt(v_len2(c1)) / (v_dist(v1, v2) * v_dist(v0, v1)); /* :185 */
    (void)tan_v1;

    for (int m = 1; m <= M; ++m) { /* :187-206 */
        int n = (int)(sqrt3_2 / tan_v0 * m + 0.5);
        int n1 = (int)(sqrt3_2 / tan_v0 * m);
        if (m % 2 == 0 && n == n1) n += 1;
        double s0 = m * sqrt3_2 * sd / sin_v0, s1 = m * sqrt3_2 * sd / sin_v1;
        double v0m[3], v1m[3];
        for (int k = 0; k < 3; ++k) { v0m[k] = v0[k] + s0 * n02[k]; v1m[k] = v1[k] + s1 * n12[k]; }
        if (v_dist(v0m, v1m) <= sd) break;
        double delta_d = ((n + (m % 2) / 2.0) - m * sqrt3_2 / tan_v0) * sd;
        double v[3];
        for (int k = 0; k < 3; ++k) v[k] = v0m[k] + delta_d * n01[k];
        int N1 = (int)(v_dist(v, v1m) / sd);
        for (int i = 0; i <= N1; ++i) /* v + (i * n_v0v1) * sd */
            push3(ps, v[0] + i * n01[0] * sd, v[1] + i * n01[1] * sd, v[2] + i * n01[2] * sd);
    }
    push(ps, v2); /* :207 */

    N = sqrt(ls[(max_i + 1) % 3]) / sd; /* :210-218 */
    if (N > 1) {
        if (N == (int)N) N -= 1;
        for (int n = 1; n <= N; ++n) push3(ps, v1[0] + n12[0] * sd * n, v1[1] + n12[1] * sd * n, v1[2] + n12[2] * sd * n);
    }
    N = sqrt(ls[(max_i + 2) % 3]) / sd; /* :220-228 */
    if (N > 1) {
        if (N == (int)N) N -= 1;
        double e20[3], n20[3];
        v_sub(v0, v2, e20);
        v_normalize(e20, n20);
        for (int n = 1; n <= N; ++n) push3(ps, v2[0] + n20[0] * sd * n, v2[1] + n20[1] * sd * n, v2[2] + n20[2] * sd * n);
    }
}

uint64_t ora_sample_triangle(const double *tri9, double sd, double *samples, uint64_t cap) {
    double vs[3][3];
    memcpy(vs, tri9, 72);
    sink_t k = {samples, cap, 0};
    sample_triangle(vs, &k, sd);
    return k.n;
}

/* isFaceOutEnvelop_sampling: LocalOperations.cpp:1046-1109 */
static int face_out(const ora_surface *s, const double *tri9, double sd, double eps2, double **scratch, uint64_t *cap,
                    uint64_t *nsamp, int degenerate_shortcut) {
    /* :1048; Preprocess::isOutEnvelop (Preprocess.cpp:643-747) samples every face, degenerate or not */
    if (degenerate_shortcut && ora_triangle_is_degenerate(tri9, tri9 + 3, tri9 + 6)) { if (nsamp) *nsamp = 0; return 0; }
    uint64_t n = ora_sample_triangle(tri9, sd, *scratch, *cap);
    if (n > *cap) {
        *cap = n + n / 2;
        *scratch = (double *)realloc(*scratch, sizeof(double) * 3 * (*cap));
        n = ora_sample_triangle(tri9, sd, *scratch, *cap);
    }
    if (nsamp) *nsamp = n;
    const double *ps = *scratch;
    double np[3], d2 = DBL_MAX; /* :1071-1073 */
    uint32_t prev = ORA_NO_FACET;
    uint64_t cnt = 0;
    for (uint64_t i = n / 2;; i = (i + 1) % n) { /* :1078 from the middle, wrapping */
        const double *p = ps + 3 * i;
        if (prev != ORA_NO_FACET) facet_nearest(s, p, prev, np, &d2); /* :1080-1082 */
        if (d2 > eps2) tree_envelope_with_hint(s, p, eps2, &prev, np, &d2); /* :1083-1086 */
        if (d2 > eps2) return 1; /* :1088-1093 */
        cnt++;
        if (cnt >= n) break; /* :1095-1097 */
    }
    return 0;
}

void ora_envelope_faces_out(const ora_surface *s, const double *tris9, uint64_t n, double sd, double eps2, uint8_t *out,
                            uint64_t *num_samples, int threads) {
    ora_envelope_faces_out_ex(s, tris9, n, sd, eps2, 1, out, num_samples, threads);
}

/* degenerate_shortcut = 1: LocalOperations::isFaceOutEnvelop_sampling (:1046-1109); 0: the per-face body of
 * Preprocess::isOutEnvelop (Preprocess.cpp:652-739), which has no such shortcut (the caller ORs the faces of a set) */
void ora_envelope_faces_out_ex(const ora_surface *s, const double *tris9, uint64_t n, double sd, double eps2,
                               int degenerate_shortcut, uint8_t *out, uint64_t *num_samples, int threads) {
    (void)threads;
#pragma omp parallel num_threads(threads > 0 ? threads : 1)
    {
        uint64_t cap = 4096;
        double *scratch = (double *)malloc(sizeof(double) * 3 * cap);
#pragma omp for schedule(dynamic, 16)
        for (int64_t i = 0; i < (int64_t)n; ++i) {
            uint64_t ns = 0;
            out[i] = (uint8_t)face_out(s, tris9 + 9 * i, sd, eps2, &scratch, &cap, &ns, degenerate_shortcut);
            if (num_samples) num_samples[i] = ns;
        }
        free(scratch);
    }
}
