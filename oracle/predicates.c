/*
 * oracle/predicates.c -- TEST INFRASTRUCTURE ONLY (CPU oracle, never shipped, never on the product path).
 *
 * Exact geometric sign predicates on IEEE doubles, restating what the reference obtains from CGAL's
 * Exact_predicates_inexact_constructions_kernel (reference: src/tetwild/CGALTypes.h:40-41):
 *
 *   - CGAL::orientation(p,q,r,s)        used at src/tetwild/LocalOperations.cpp:755-758 and :864-867
 *   - Triangle_3f::is_degenerate()      used at src/tetwild/LocalOperations.cpp:1048   (== collinear(p,q,r))
 *
 * CGAL is a system dependency of the reference (README.md:93-103, Dockerfile:6), not vendored under
 * /root/reference. Exact predicates are mathematically defined (the sign of a polynomial in the input doubles),
 * so any exact evaluation agrees with CGAL's. Here: a forward-error static filter followed by an exact evaluation
 * with floating-point expansions (J.R. Shewchuk, "Adaptive Precision Floating-Point Arithmetic and Fast Robust
 * Geometric Predicates", 1997 -- published algorithm, restated, not copied).
 *
 * Sign convention: CGAL POSITIVE  <=>  det[q-p; r-p; s-p] > 0.
 */
#include <math.h>
#include <float.h>
#include "tw_oracle.h"

#pragma STDC FP_CONTRACT OFF

#define EPS_HALF (DBL_EPSILON * 0.5) /* 2^-53 */

/* ---- error-free transformations ------------------------------------------------------------------------------ */
static inline void two_sum(double a, double b, double *x, double *y) {
    double s = a + b;
    double bv = s - a;
    double av = s - bv;
    double br = b - bv;
    double ar = a - av;
    *x = s;
    *y = ar + br;
}
static inline void fast_two_sum(double a, double b, double *x, double *y) { /* requires |a| >= |b| */
    double s = a + b;
    double bv = s - a;
    *x = s;
    *y = b - bv;
}
static inline void two_prod(double a, double b, double *x, double *y) {
    double p = a * b;
    *x = p;
    *y = fma(a, b, -p);
}

/* h = e + f, all expansions sorted by increasing magnitude, zero components dropped. Returns length of h. */
static int exp_sum(const double *e, int ne, const double *f, int nf, double *h) {
    double g[2 * ORA_MAX_EXPANSION];
    int i = 0, j = 0, m = 0;
    while (i < ne && j < nf) {
        if (fabs(e[i]) < fabs(f[j])) g[m++] = e[i++]; else g[m++] = f[j++];
    }
    while (i < ne) g[m++] = e[i++];
    while (j < nf) g[m++] = f[j++];
    if (m == 0) return 0;
    if (m == 1) { h[0] = g[0]; return (g[0] != 0.0) ? 1 : 0; }
    int n = 0;
    double Q, q;
    fast_two_sum(g[1], g[0], &Q, &q);
    if (q != 0.0) h[n++] = q;
    for (int k = 2; k < m; ++k) {
        two_sum(Q, g[k], &Q, &q);
        if (q != 0.0) h[n++] = q;
    }
    if (Q != 0.0 || n == 0) h[n++] = Q;
    if (n == 1 && h[0] == 0.0) return 0;
    return n;
}

/* h = e * b */
static int exp_scale(const double *e, int ne, double b, double *h) {
    if (ne == 0 || b == 0.0) return 0;
    int n = 0;
    double Q, q, P, p, s;
    two_prod(e[0], b, &Q, &q);
    if (q != 0.0) h[n++] = q;
    for (int k = 1; k < ne; ++k) {
        two_prod(e[k], b, &P, &p);
        two_sum(Q, p, &s, &q);
        if (q != 0.0) h[n++] = q;
        fast_two_sum(P, s, &Q, &q);
        if (q != 0.0) h[n++] = q;
    }
    if (Q != 0.0 || n == 0) h[n++] = Q;
    if (n == 1 && h[0] == 0.0) return 0;
    return n;
}

/* pq = px*qy - qx*py as an expansion of at most 4 terms */
static int cross2(double px, double py, double qx, double qy, double *h) {
    double a[2], b[2];
    two_prod(px, qy, &a[1], &a[0]);
    double t1, t0;
    two_prod(qx, py, &t1, &t0);
    b[0] = -t0;
    b[1] = -t1;
    /* drop zeros so inputs to exp_sum are valid expansions */
    double aa[2], bb[2];
    int na = 0, nb = 0;
    if (a[0] != 0.0) aa[na++] = a[0];
    if (a[1] != 0.0) aa[na++] = a[1];
    if (b[0] != 0.0) bb[nb++] = b[0];
    if (b[1] != 0.0) bb[nb++] = b[1];
    return exp_sum(aa, na, bb, nb, h);
}

static int exp_sign(const double *e, int n) {
    if (n == 0) return 0;
    return (e[n - 1] > 0.0) - (e[n - 1] < 0.0);
}

static void exp_neg(double *e, int n) { for (int i = 0; i < n; ++i) e[i] = -e[i]; }

/* exact sign of det | a 1; b 1; c 1; d 1 | == det[a-d; b-d; c-d] */
int ora_orient3d_exact(const double *a, const double *b, const double *c, const double *d) {
    double ab[4], bc[4], cd[4], da[4], ac[4], bd[4];
    int nab = cross2(a[0], a[1], b[0], b[1], ab);
    int nbc = cross2(b[0], b[1], c[0], c[1], bc);
    int ncd = cross2(c[0], c[1], d[0], d[1], cd);
    int nda = cross2(d[0], d[1], a[0], a[1], da);
    int nac = cross2(a[0], a[1], c[0], c[1], ac);
    int nbd = cross2(b[0], b[1], d[0], d[1], bd);
    double t8[8], bcd[12], cda[12], dab[12], abc[12];
    int nt;
    /* minors over rows (x,y,1): see derivation in DESIGN.md (cofactor expansion along z) */
    nt = exp_sum(cd, ncd, da, nda, t8);
    int ncda = exp_sum(t8, nt, ac, nac, cda); /* ac + cd + da */
    nt = exp_sum(da, nda, ab, nab, t8);
    int ndab = exp_sum(t8, nt, bd, nbd, dab); /* ab + bd + da */
    exp_neg(bd, nbd);
    exp_neg(ac, nac);
    nt = exp_sum(ab, nab, bc, nbc, t8);
    int nabc = exp_sum(t8, nt, ac, nac, abc); /* ab + bc - ac */
    nt = exp_sum(bc, nbc, cd, ncd, t8);
    int nbcd = exp_sum(t8, nt, bd, nbd, bcd); /* bc + cd - bd */
    double adet[24], bdet[24], cdet[24], ddet[24];
    int na = exp_scale(bcd, nbcd, a[2], adet);
    int nb = exp_scale(cda, ncda, -b[2], bdet);
    int nc = exp_scale(dab, ndab, c[2], cdet);
    int nd = exp_scale(abc, nabc, -d[2], ddet);
    double abdet[48], cddet[48], det[96];
    int nabd = exp_sum(adet, na, bdet, nb, abdet);
    int ncdd = exp_sum(cdet, nc, ddet, nd, cddet);
    int ndet = exp_sum(abdet, nabd, cddet, ncdd, det);
    return exp_sign(det, ndet);
}

/* sign of det[a-d; b-d; c-d]; static filter then exact */
int ora_orient3d(const double *a, const double *b, const double *c, const double *d) {
    double adx = a[0] - d[0], bdx = b[0] - d[0], cdx = c[0] - d[0];
    double ady = a[1] - d[1], bdy = b[1] - d[1], cdy = c[1] - d[1];
    double adz = a[2] - d[2], bdz = b[2] - d[2], cdz = c[2] - d[2];
    double bdxcdy = bdx * cdy, cdxbdy = cdx * bdy;
    double cdxady = cdx * ady, adxcdy = adx * cdy;
    double adxbdy = adx * bdy, bdxady = bdx * ady;
    double det = adz * (bdxcdy - cdxbdy) + bdz * (cdxady - adxcdy) + cdz * (adxbdy - bdxady);
    double permanent = (fabs(bdxcdy) + fabs(cdxbdy)) * fabs(adz) + (fabs(cdxady) + fabs(adxcdy)) * fabs(bdz) +
                       (fabs(adxbdy) + fabs(bdxady)) * fabs(cdz);
    double errbound = (7.0 + 56.0 * EPS_HALF) * EPS_HALF * permanent;
    if (det > errbound) return 1;
    if (-det > errbound) return -1;
    return ora_orient3d_exact(a, b, c, d);
}

/* CGAL::orientation(p,q,r,s): sign det[q-p; r-p; s-p]  (reference call sites LocalOperations.cpp:755,864) */
int ora_cgal_orientation(const double *p, const double *q, const double *r, const double *s) {
    return ora_orient3d(q, r, s, p);
}

/* exact sign of (px-rx)(qy-ry) - (py-ry)(qx-rx) */
static int orient2d_exact(double px, double py, double qx, double qy, double rx, double ry) {
    double pq[4], qr[4], rp[4], t[8], det[12];
    int npq = cross2(px, py, qx, qy, pq);
    int nqr = cross2(qx, qy, rx, ry, qr);
    int nrp = cross2(rx, ry, px, py, rp);
    int nt = exp_sum(pq, npq, qr, nqr, t);
    int nd = exp_sum(t, nt, rp, nrp, det);
    return exp_sign(det, nd);
}

static int orient2d(double px, double py, double qx, double qy, double rx, double ry) {
    double l = (px - rx) * (qy - ry);
    double r = (py - ry) * (qx - rx);
    double det = l - r;
    double detsum;
    if (l > 0.0) {
        if (r <= 0.0) return (det > 0.0) - (det < 0.0);
        detsum = l + r;
    } else if (l < 0.0) {
        if (r >= 0.0) return (det > 0.0) - (det < 0.0);
        detsum = -l - r;
    } else {
        return (det > 0.0) - (det < 0.0);
    }
    double errbound = (3.0 + 16.0 * EPS_HALF) * EPS_HALF * detsum;
    if (det >= errbound || -det >= errbound) return (det > 0.0) - (det < 0.0);
    return orient2d_exact(px, py, qx, qy, rx, ry);
}

/* CGAL Triangle_3::is_degenerate() == collinear(p,q,r): all three axis-projected 2D orientations vanish
 * (reference call site LocalOperations.cpp:1048). */
int ora_triangle_is_degenerate(const double *p, const double *q, const double *r) {
    if (orient2d(p[0], p[1], q[0], q[1], r[0], r[1]) != 0) return 0;
    if (orient2d(p[0], p[2], q[0], q[2], r[0], r[2]) != 0) return 0;
    if (orient2d(p[1], p[2], q[1], q[2], r[1], r[2]) != 0) return 0;
    return 1;
}
