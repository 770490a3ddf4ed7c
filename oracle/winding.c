/*
 * oracle/winding.c -- TEST INFRASTRUCTURE ONLY (CPU oracle; never on the product path).
 *
 * Restates the generalized-winding-number inside/outside filter of the reference:
 *   InoutFiltering::filter      src/tetwild/InoutFiltering.cpp:23-82   (W > 0.5 keeps a tet; flip-and-retry :56-75)
 *   twins                       src/tetwild/MeshRefinement.cpp:592-624, :1036-1068
 * whose arithmetic lives in igl::winding_number (libigl 45cfc79, cmake/TetWildDownloadExternal.cmake:18-21),
 * a third-party dependency NOT under /root/reference -> restated from the published algorithm
 * (A. Jacobson, L. Kavan, O. Sorkine-Hornung, "Robust Inside-Outside Segmentation using Generalized Winding
 * Numbers", SIGGRAPH 2013): per-triangle solid angle by the Van Oosterom-Strackee formula, summed directly
 * (ora_winding_direct), or evaluated through the paper's exact bounding-volume hierarchy in which a sub-mesh whose
 * bounding box does not contain the query is replaced by a fan over its exterior (unmatched) edges
 * (ora_wtree_*; section 4.1 of the paper; median split on the longest axis, leaf at <= 100 facets).
 * PARITY UNPINNED for this leaf arithmetic (no reference-owned vector exists). Decisions only depend
 * on W > 0.5; both evaluations are equal in exact arithmetic and differ by rounding only.
 */
#include <math.h>
#include <float.h>
#include <stdlib.h>
#include <string.h>
#include "tw_oracle.h"

#pragma STDC FP_CONTRACT OFF

#define WT_MIN_F 100

/* solid angle of triangle (a,b,c) seen from p, divided by 4*pi (contribution to the winding number) */
double ora_solid_angle_w(const double *A, const double *B, const double *C, const double *P) {
    double v[3][3];
    for (int d = 0; d < 3; ++d) { v[0][d] = A[d] - P[d]; v[1][d] = B[d] - P[d]; v[2][d] = C[d] - P[d]; }
    double vl[3];
    for (int i = 0; i < 3; ++i) vl[i] = sqrt(v[i][0] * v[i][0] + v[i][1] * v[i][1] + v[i][2] * v[i][2]);
    double detf = v[0][0] * v[1][1] * v[2][2] + v[1][0] * v[2][1] * v[0][2] + v[2][0] * v[0][1] * v[1][2] -
                  v[2][0] * v[1][1] * v[0][2] - v[1][0] * v[0][1] * v[2][2] - v[0][0] * v[2][1] * v[1][2];
    double dp0 = v[1][0] * v[2][0] + v[1][1] * v[2][1] + v[1][2] * v[2][2];
    double dp1 = v[2][0] * v[0][0] + v[2][1] * v[0][1] + v[2][2] * v[0][2];
    double dp2 = v[0][0] * v[1][0] + v[0][1] * v[1][1] + v[0][2] * v[1][2];
    return atan2(detf, vl[0] * vl[1] * vl[2] + dp0 * vl[0] + dp1 * vl[1] + dp2 * vl[2]) / (2.0 * M_PI);
}

static double sum_faces(const double *V, const uint32_t *F, uint64_t nF, const double *p) {
    double w = 0.0;
    for (uint64_t f = 0; f < nF; ++f)
        w += ora_solid_angle_w(V + 3 * (size_t)F[3 * f], V + 3 * (size_t)F[3 * f + 1], V + 3 * (size_t)F[3 * f + 2], p);
    return w;
}

void ora_winding_direct(const double *V, uint32_t nV, const uint32_t *F, uint32_t nF, const double *C, uint64_t nC,
                        double *W, int threads) {
    (void)nV; (void)threads;
#pragma omp parallel for schedule(dynamic, 64) num_threads(threads > 0 ? threads : 1)
    for (int64_t i = 0; i < (int64_t)nC; ++i) W[i] = sum_faces(V, F, nF, C + 3 * i);
}

/* ---- hierarchy ---- */
typedef struct {
    double lo[3], hi[3];
    uint32_t fb, fe;       /* facet range in the permuted facet array */
    uint64_t cb, ce;       /* cap range (triangles) in the cap array */
    int32_t child[2];      /* -1 = leaf */
} wnode;

struct ora_wtree {
    uint32_t nV, nF;
    double *V;
    uint32_t *F;           /* permuted, vertex ids canonicalised (exact duplicates merged) */
    wnode *nodes; uint64_t n_nodes, cap_nodes;
    uint32_t *cap; uint64_t n_cap, cap_cap; /* cap triangles, 3 ids each */
};

typedef struct { uint64_t key; int32_t sgn; } dedge;
static int dedge_cmp(const void *a, const void *b) {
    uint64_t x = ((const dedge *)a)->key, y = ((const dedge *)b)->key;
    return x < y ? -1 : (x > y);
}

/* exterior (unmatched) directed edges of facets [fb,fe) fanned to one apex */
static void build_cap(ora_wtree *t, wnode *nd) {
    uint32_t nf = nd->fe - nd->fb;
    dedge *E = (dedge *)malloc(sizeof(dedge) * 3 * (size_t)nf);
    for (uint32_t f = 0; f < nf; ++f) {
        const uint32_t *tri = t->F + 3 * (size_t)(nd->fb + f);
        for (int k = 0; k < 3; ++k) {
            uint32_t i = tri[k], j = tri[(k + 1) % 3];
            dedge e;
            if (i < j) { e.key = ((uint64_t)i << 32) | j; e.sgn = 1; }
            else { e.key = ((uint64_t)j << 32) | i; e.sgn = -1; }
            E[3 * (size_t)f + k] = e;
        }
    }
    qsort(E, 3 * (size_t)nf, sizeof(dedge), dedge_cmp);
    nd->cb = t->n_cap;
    int have_apex = 0;
    uint32_t apex = 0;
    for (size_t a = 0; a < 3 * (size_t)nf;) {
        size_t b = a;
        int32_t net = 0;
        while (b < 3 * (size_t)nf && E[b].key == E[a].key) { net += E[b].sgn; ++b; }
        uint32_t i = (uint32_t)(E[a].key >> 32), j = (uint32_t)(E[a].key & 0xffffffffu);
        if (i != j && net != 0) {
            if (net < 0) { uint32_t tmp = i; i = j; j = tmp; net = -net; }
            if (!have_apex) { apex = i; have_apex = 1; }
            if (i != apex && j != apex) {
                for (int32_t r = 0; r < net; ++r) {
                    if (t->n_cap + 1 > t->cap_cap) {
                        t->cap_cap = t->cap_cap ? t->cap_cap * 2 : 1024;
                        t->cap = (uint32_t *)realloc(t->cap, sizeof(uint32_t) * 3 * t->cap_cap);
                    }
                    uint32_t *c = t->cap + 3 * t->n_cap++;
                    c[0] = apex; c[1] = i; c[2] = j;
                }
            }
        }
        a = b;
    }
    nd->ce = t->n_cap;
    free(E);
}

static int dbl_cmp(const void *a, const void *b) {
    double x = *(const double *)a, y = *(const double *)b;
    return x < y ? -1 : (x > y);
}

static int32_t grow(ora_wtree *t, uint32_t fb, uint32_t fe) {
    if (t->n_nodes + 1 > t->cap_nodes) {
        t->cap_nodes = t->cap_nodes ? t->cap_nodes * 2 : 64;
        t->nodes = (wnode *)realloc(t->nodes, sizeof(wnode) * t->cap_nodes);
    }
    int32_t id = (int32_t)t->n_nodes++;
    wnode nd;
    nd.fb = fb; nd.fe = fe; nd.child[0] = nd.child[1] = -1;
    for (int c = 0; c < 3; ++c) { nd.lo[c] = DBL_MAX; nd.hi[c] = -DBL_MAX; }
    for (uint32_t f = fb; f < fe; ++f)
        for (int k = 0; k < 3; ++k) {
            const double *p = t->V + 3 * (size_t)t->F[3 * (size_t)f + k];
            for (int c = 0; c < 3; ++c) { if (p[c] < nd.lo[c]) nd.lo[c] = p[c]; if (p[c] > nd.hi[c]) nd.hi[c] = p[c]; }
        }
    build_cap(t, &nd);
    uint32_t nf = fe - fb;
    uint64_t ncap = nd.ce - nd.cb;
    if (!(nf <= WT_MIN_F || (int64_t)ncap - 2 >= (int64_t)nf)) {
        int ax = 0;
        double len = -DBL_MAX;
        for (int c = 0; c < 3; ++c) if (nd.hi[c] - nd.lo[c] > len) { len = nd.hi[c] - nd.lo[c]; ax = c; }
        double *bc = (double *)malloc(sizeof(double) * nf), *srt = (double *)malloc(sizeof(double) * nf);
        for (uint32_t f = 0; f < nf; ++f) {
            const uint32_t *tri = t->F + 3 * (size_t)(fb + f);
            bc[f] = (t->V[3 * (size_t)tri[0] + ax] + t->V[3 * (size_t)tri[1] + ax] + t->V[3 * (size_t)tri[2] + ax]) / 3.0;
        }
        memcpy(srt, bc, sizeof(double) * nf);
        qsort(srt, nf, sizeof(double), dbl_cmp);
        double split = (nf % 2) ? srt[nf / 2] : 0.5 * (srt[nf / 2 - 1] + srt[nf / 2]);
        uint32_t *tmp = (uint32_t *)malloc(sizeof(uint32_t) * 3 * (size_t)nf);
        uint32_t nl = 0;
        for (uint32_t f = 0; f < nf; ++f) if (bc[f] <= split) nl++;
        if (nl != 0 && nl != nf) {
            uint32_t a = 0, b = nl;
            for (uint32_t f = 0; f < nf; ++f) {
                uint32_t dst = (bc[f] <= split) ? a++ : b++;
                memcpy(tmp + 3 * (size_t)dst, t->F + 3 * (size_t)(fb + f), 12);
            }
            memcpy(t->F + 3 * (size_t)fb, tmp, sizeof(uint32_t) * 3 * (size_t)nf);
            free(tmp); free(bc); free(srt);
            t->nodes[id] = nd;
            int32_t l = grow(t, fb, fb + nl);
            int32_t r = grow(t, fb + nl, fe);
            nd.child[0] = l; nd.child[1] = r;
        } else { free(tmp); free(bc); free(srt); }
    }
    t->nodes[id] = nd;
    return id;
}

typedef struct { double x, y, z; uint32_t id; } vkey;
static int vkey_cmp(const void *a, const void *b) {
    const vkey *p = (const vkey *)a, *q = (const vkey *)b;
    if (p->x != q->x) return p->x < q->x ? -1 : 1;
    if (p->y != q->y) return p->y < q->y ? -1 : 1;
    if (p->z != q->z) return p->z < q->z ? -1 : 1;
    return p->id < q->id ? -1 : (p->id > q->id);
}

ora_wtree *ora_wtree_create(const double *V, uint32_t nV, const uint32_t *F, uint32_t nF) {
    ora_wtree *t = (ora_wtree *)calloc(1, sizeof(*t));
    t->nV = nV; t->nF = nF;
    t->V = (double *)malloc(sizeof(double) * 3 * (size_t)(nV ? nV : 1));
    memcpy(t->V, V, sizeof(double) * 3 * (size_t)nV);
    /* merge exactly coincident vertices (remove_duplicate_vertices with epsilon 0) */
    uint32_t *canon = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(nV ? nV : 1));
    vkey *vk = (vkey *)malloc(sizeof(vkey) * (size_t)(nV ? nV : 1));
    for (uint32_t i = 0; i < nV; ++i) { vk[i].x = V[3 * (size_t)i]; vk[i].y = V[3 * (size_t)i + 1]; vk[i].z = V[3 * (size_t)i + 2]; vk[i].id = i; }
    qsort(vk, nV, sizeof(vkey), vkey_cmp);
    for (uint32_t i = 0; i < nV; ++i) {
        if (i > 0 && vk[i].x == vk[i - 1].x && vk[i].y == vk[i - 1].y && vk[i].z == vk[i - 1].z) canon[vk[i].id] = canon[vk[i - 1].id];
        else canon[vk[i].id] = vk[i].id;
    }
    free(vk);
    t->F = (uint32_t *)malloc(sizeof(uint32_t) * 3 * (size_t)(nF ? nF : 1));
    for (size_t k = 0; k < 3 * (size_t)nF; ++k) t->F[k] = canon[F[k]];
    free(canon);
    if (nF) grow(t, 0, nF);
    return t;
}

void ora_wtree_destroy(ora_wtree *t) {
    if (!t) return;
    free(t->V); free(t->F); free(t->nodes); free(t->cap); free(t);
}

uint64_t ora_wtree_stats(const ora_wtree *t, uint64_t *n_nodes, uint64_t *cap_total) {
    if (n_nodes) *n_nodes = t->n_nodes;
    if (cap_total) *cap_total = t->n_cap;
    return t->n_nodes;
}

static double wt_eval(const ora_wtree *t, int32_t id, const double *p, uint64_t *work) {
    const wnode *nd = &t->nodes[id];
    int inside = 1;
    for (int c = 0; c < 3; ++c) if (p[c] < nd->lo[c] || p[c] > nd->hi[c]) { inside = 0; break; }
    if (inside) {
        if (nd->child[0] >= 0) return wt_eval(t, nd->child[0], p, work) + wt_eval(t, nd->child[1], p, work);
        if (work) *work += nd->fe - nd->fb;
        return sum_faces(t->V, t->F + 3 * (size_t)nd->fb, nd->fe - nd->fb, p);
    }
    uint64_t ncap = nd->ce - nd->cb;
    uint32_t nf = nd->fe - nd->fb;
    if ((int64_t)ncap - 2 < (int64_t)nf) {
        if (work) *work += ncap;
        return sum_faces(t->V, t->cap + 3 * nd->cb, ncap, p);
    }
    if (work) *work += nf;
    return sum_faces(t->V, t->F + 3 * (size_t)nd->fb, nf, p);
}

void ora_wtree_eval(const ora_wtree *t, const double *C, uint64_t nC, double *W, int threads) {
    (void)threads;
    if (t->nF == 0) { for (uint64_t i = 0; i < nC; ++i) W[i] = 0.0; return; }
#pragma omp parallel for schedule(dynamic, 256) num_threads(threads > 0 ? threads : 1)
    for (int64_t i = 0; i < (int64_t)nC; ++i) W[i] = wt_eval(t, 0, C + 3 * i, NULL);
}

/* InoutFiltering::filter, InoutFiltering.cpp:45-75: keep = W > 0.5; if nothing survives swap F columns 1,2 and retry */
int ora_inout_filter(const double *V, uint32_t nV, const uint32_t *F, uint32_t nF, const double *C, uint64_t nC,
                     uint8_t *keep, double *W, int hierarchical, int threads) {
    double *w = W ? W : (double *)malloc(sizeof(double) * (size_t)(nC ? nC : 1));
    int retried = 0;
    uint32_t *F2 = NULL;
    const uint32_t *Fc = F;
    for (int pass = 0; pass < 2; ++pass) {
        if (hierarchical) {
            ora_wtree *t = ora_wtree_create(V, nV, Fc, nF);
            ora_wtree_eval(t, C, nC, w, threads);
            ora_wtree_destroy(t);
        } else {
            ora_winding_direct(V, nV, Fc, nF, C, nC, w, threads);
        }
        uint64_t kept = 0;
        for (uint64_t i = 0; i < nC; ++i) { keep[i] = w[i] > 0.5; kept += keep[i]; }
        if (kept != 0 || pass == 1) break;
        retried = 1;
        F2 = (uint32_t *)malloc(sizeof(uint32_t) * 3 * (size_t)(nF ? nF : 1));
        for (uint32_t f = 0; f < nF; ++f) { F2[3 * f] = F[3 * f]; F2[3 * f + 1] = F[3 * f + 2]; F2[3 * f + 2] = F[3 * f + 1]; }
        Fc = F2;
    }
    free(F2);
    if (!W) free(w);
    return retried;
}

#ifdef _OPENMP
#include <omp.h>
int ora_max_threads(void) { return omp_get_max_threads(); }
#else
int ora_max_threads(void) { return 1; }
#endif
