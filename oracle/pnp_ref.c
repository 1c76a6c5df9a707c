/*
 * pnp_ref.c — plain-C restatement of the deterministic PnP + RANSAC solver (TEST ORACLE ONLY).
 *
 * Reference call this stage replaces (ros/gisnav/gisnav/core/_shared.py:95-119):
 *     z = elevation[floor(y), floor(x)]                       (_shared.py:100-102, raw uint8 DEM)
 *     cv2.solvePnPRansac(obj f32[K,3], img f32[K,2], k, zeros(4,1),
 *                        useExtrinsicGuess=False, iterationsCount=10)   (_shared.py:109-116)
 *     cv2.Rodrigues(rvec)                                     (_shared.py:117)
 * OpenCV's arithmetic is not under /root/reference (opencv-python-headless, un-pinned,
 * ros/gisnav/setup.py:116); oracle/cv2_ref.py runs that very call (4.13.0 here) for the pose-level
 * comparison.  This file is the *bit-exact* oracle for the index work (which hypotheses are drawn,
 * per-hypothesis inlier counts, the winning hypothesis, the inlier mask): OpenCV draws minimal
 * sets from an internal fixed-seed RNG that a GPU scorer cannot reproduce, so the solver is
 * defined here (SURVEY.md §7 "hard parts") and the CUDA kernels in gisnav_b200/csrc/pnp.cu follow
 * the same operation order:
 *
 *   - counter-based RNG (lowbias32 hash of seed/hypothesis/draw) -> 4 distinct point indices;
 *   - P3P on the first three (Grunert's distance-ratio quartic, Ferrari with the resolvent root
 *     found by bisection + Newton: only + - * / sqrt, all IEEE-754 correctly rounded, so CPU and GPU
 *     agree to the bit), disambiguated by the reprojection error of the fourth point;
 *   - fp32 scoring: squared reprojection error <= thr^2 and positive depth (OpenCV:
 *     RANSACPointSetRegistrator::findInliers uses err <= thr*thr on float errors);
 *   - winner = most inliers, lowest hypothesis index on ties;
 *   - fp64 Levenberg-Marquardt refit on the winner's inliers (OpenCV refits with
 *     SOLVEPNP_ITERATIVE); this part is compared to tolerance, not to the bit.
 *
 * Compile with -ffp-contract=off (the CUDA side uses --fmad=false): every multiply-add below is
 * either two roundings, or an explicit fma()/fmaf().
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#define GNB_MAX_DRAW_TRIES 64

static inline uint32_t hash32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}

static inline uint32_t rnd32(uint32_t seed, uint32_t hyp, uint32_t ctr) {
    return hash32(seed ^ hash32(hyp * 0x9E3779B9U + ctr * 0x85EBCA6BU + 0x165667B1U));
}

/* 4 distinct indices in [0,n); returns 0 when n < 4 or the tries run out */
static int draw4(uint32_t seed, uint32_t hyp, uint32_t n, uint32_t idx[4]) {
    if (n < 4) return 0;
    uint32_t ctr = 0;
    for (int j = 0; j < 4; ++j) {
        int ok = 0;
        while (!ok && ctr < GNB_MAX_DRAW_TRIES) {
            uint32_t r = rnd32(seed, hyp, ctr++);
            uint32_t c = (uint32_t)(((uint64_t)r * (uint64_t)n) >> 32);
            ok = 1;
            for (int q = 0; q < j; ++q) if (idx[q] == c) ok = 0;
            if (ok) idx[j] = c;
        }
        if (!ok) return 0;
    }
    return 1;
}

/* ---- small fp64 helpers ------------------------------------------------------------------- */
static inline void cross3(const double a[3], const double b[3], double c[3]) {
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
}
static inline double dot3(const double a[3], const double b[3]) {
    return a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
}

/* inverse of the 3x3 matrix with COLUMNS c0,c1,c2; returns 0 if singular */
static int inv3_cols(const double c0[3], const double c1[3], const double c2[3], double inv[9]) {
    double x12[3], x20[3], x01[3];
    cross3(c1, c2, x12); cross3(c2, c0, x20); cross3(c0, c1, x01);
    double det = dot3(c0, x12);
    if (!(fabs(det) > 1e-300)) return 0;
    double id = 1.0 / det;
    for (int j = 0; j < 3; ++j) {
        inv[0 * 3 + j] = x12[j] * id;
        inv[1 * 3 + j] = x20[j] * id;
        inv[2 * 3 + j] = x01[j] * id;
    }
    return 1;
}

/* real roots of y^2 + b y + c */
static int quad_roots(double b, double c, double r[2]) {
    double disc = b * b - 4.0 * c;
    if (disc < 0.0) return 0;
    double s = sqrt(disc);
    double q = (b >= 0.0) ? -0.5 * (b + s) : -0.5 * (b - s);
    r[0] = q;
    r[1] = (q != 0.0) ? c / q : 0.0;
    return 2;
}

/* real roots of a4 v^4 + a3 v^3 + a2 v^2 + a1 v + a0 (Ferrari; algebraic operations only) */
static int quartic_roots(double a4, double a3, double a2, double a1, double a0, double roots[4]) {
    if (!(fabs(a4) > 1e-300)) return 0;
    double ia = 1.0 / a4;
    double a = a3 * ia, b = a2 * ia, c = a1 * ia, d = a0 * ia;
    double aa = a * a;
    double p = b - 0.375 * aa;
    double q = c - 0.5 * a * b + 0.125 * aa * a;
    double r = d - 0.25 * a * c + 0.0625 * aa * b - (3.0 / 256.0) * aa * aa;
    double sh = 0.25 * a;
    int n = 0;
    double scale = fabs(p) + sqrt(fabs(r)) + 1e-300;
    if (fabs(q) <= 1e-14 * scale * sqrt(scale)) {
        /* biquadratic: y^4 + p y^2 + r */
        double z[2];
        int nz = quad_roots(p, r, z);
        for (int i = 0; i < nz; ++i) {
            if (z[i] >= 0.0) {
                double s = sqrt(z[i]);
                roots[n++] = s - sh;
                roots[n++] = -s - sh;
            }
        }
        return n;
    }
    /* resolvent g(m) = 8 m^3 + 8 p m^2 + (2 p^2 - 8 r) m - q^2, g(0) < 0: positive root by bisection */
    double c2 = 8.0 * p, c1 = 2.0 * p * p - 8.0 * r, c0 = -q * q;
    double hi = 1.0 + fabs(c2) * 0.125;
    double t1 = sqrt(fabs(c1) * 0.125), t0 = fabs(c0) * 0.125;
    if (t1 > hi) hi = t1;
    if (t0 > hi) hi = t0;
    hi = 2.0 * hi + 1.0;
    double lo = 0.0;
    for (int it = 0; it < 8 && (((8.0 * hi + c2) * hi + c1) * hi + c0) <= 0.0; ++it) hi *= 4.0;
    for (int it = 0; it < 80; ++it) {
        double mid = 0.5 * (lo + hi);
        double g = ((8.0 * mid + c2) * mid + c1) * mid + c0;
        if (g <= 0.0) lo = mid; else hi = mid;
    }
    double m = 0.5 * (lo + hi);
    for (int it = 0; it < 3; ++it) { /* Newton polish */
        double g = ((8.0 * m + c2) * m + c1) * m + c0;
        double dg = (24.0 * m + 2.0 * c2) * m + c1;
        if (fabs(dg) > 1e-300) {
            double mn = m - g / dg;
            if (mn > 0.0) m = mn;
        }
    }
    if (!(m > 0.0)) return 0;
    double s = sqrt(2.0 * m);
    double hq = q / (2.0 * s);
    double base = 0.5 * p + m;
    double z[2];
    int nz = quad_roots(-s, base + hq, z);
    for (int i = 0; i < nz; ++i) roots[n++] = z[i] - sh;
    nz = quad_roots(s, base - hq, z);
    for (int i = 0; i < nz; ++i) roots[n++] = z[i] - sh;
    return n;
}

typedef struct { double r[9]; double t[3]; } pose_t;

/*
 * P3P.  f[i] = unit bearing of image point i in the camera frame, x[i] = world point.
 * Unknown depths s_i: with s2 = u s1, s3 = v s1 the three law-of-cosines equations reduce to
 *   u = N(v)/D(v),  N = (K-1) v^2 - 2 K cb v + (K+1),  D = 2 (cg - v ca),  K = (a^2-c^2)/b^2
 *   b^2 (D^2 + N^2 - 2 cg N D) - c^2 (1 + v^2 - 2 cb v) D^2 = 0           (quartic in v)
 * where a=|x2-x3|, b=|x1-x3|, c=|x1-x2|, ca=f2.f3, cb=f1.f3, cg=f1.f2.
 */
static int p3p(const double f[3][3], const double x[3][3], pose_t sol[4]) {
    double d12[3], d13[3], d23[3];
    for (int k = 0; k < 3; ++k) {
        d12[k] = x[1][k] - x[0][k];
        d13[k] = x[2][k] - x[0][k];
        d23[k] = x[2][k] - x[1][k];
    }
    double a2 = dot3(d23, d23), b2 = dot3(d13, d13), c2 = dot3(d12, d12);
    if (!(a2 > 1e-12 && b2 > 1e-12 && c2 > 1e-12)) return 0;
    double nx[3];
    cross3(d12, d13, nx);
    if (!(dot3(nx, nx) > 1e-12 * b2 * c2)) return 0; /* collinear */
    double ca = dot3(f[1], f[2]), cb = dot3(f[0], f[2]), cg = dot3(f[0], f[1]);
    double K = (a2 - c2) / b2;
    double n2 = K - 1.0, n1 = -2.0 * K * cb, n0 = K + 1.0;
    double e1 = -2.0 * ca, e0 = 2.0 * cg;
    /* D^2 */
    double dd2 = e1 * e1, dd1 = 2.0 * e1 * e0, dd0 = e0 * e0;
    /* N^2 */
    double nn4 = n2 * n2, nn3 = 2.0 * n2 * n1, nn2 = 2.0 * n2 * n0 + n1 * n1, nn1 = 2.0 * n1 * n0, nn0 = n0 * n0;
    /* N*D */
    double nd3 = n2 * e1, nd2 = n2 * e0 + n1 * e1, nd1 = n1 * e0 + n0 * e1, nd0 = n0 * e0;
    /* (1 + v^2 - 2 cb v) D^2 */
    double q1 = -2.0 * cb;
    double w4 = dd2, w3 = dd1 + q1 * dd2, w2 = dd0 + q1 * dd1 + dd2, w1 = q1 * dd0 + dd1, w0 = dd0;
    double tc = 2.0 * cg;
    double A4 = b2 * nn4 - c2 * w4;
    double A3 = b2 * (nn3 - tc * nd3) - c2 * w3;
    double A2 = b2 * (dd2 + nn2 - tc * nd2) - c2 * w2;
    double A1 = b2 * (dd1 + nn1 - tc * nd1) - c2 * w1;
    double A0 = b2 * (dd0 + nn0 - tc * nd0) - c2 * w0;
    double vs[4];
    int nv = quartic_roots(A4, A3, A2, A1, A0, vs);
    double xinv[9];
    if (!inv3_cols(d12, d13, nx, xinv)) return 0;
    int ns = 0;
    for (int i = 0; i < nv; ++i) {
        double v = vs[i];
        if (!(v > 0.0)) continue;
        double den = e0 + e1 * v;
        if (!(fabs(den) > 1e-12)) continue;
        double u = ((n2 * v + n1) * v + n0) / den;
        if (!(u > 0.0)) continue;
        double s1sq = b2 / ((v + q1) * v + 1.0);
        if (!(s1sq > 0.0)) continue;
        double s[3];
        s[0] = sqrt(s1sq); s[1] = u * s[0]; s[2] = v * s[0];
        /* two Gauss-Newton steps on the three distance equations */
        for (int it = 0; it < 2; ++it) {
            double r0 = s[1] * s[1] + s[2] * s[2] - 2.0 * s[1] * s[2] * ca - a2;
            double r1 = s[0] * s[0] + s[2] * s[2] - 2.0 * s[0] * s[2] * cb - b2;
            double r2 = s[0] * s[0] + s[1] * s[1] - 2.0 * s[0] * s[1] * cg - c2;
            double j0[3] = {0.0, 2.0 * (s[1] - s[2] * ca), 2.0 * (s[2] - s[1] * ca)};
            double j1[3] = {2.0 * (s[0] - s[2] * cb), 0.0, 2.0 * (s[2] - s[0] * cb)};
            double j2[3] = {2.0 * (s[0] - s[1] * cg), 2.0 * (s[1] - s[0] * cg), 0.0};
            /* solve J ds = r with J rows j0,j1,j2: columns of J^T are the rows */
            double jc0[3] = {j0[0], j1[0], j2[0]}, jc1[3] = {j0[1], j1[1], j2[1]}, jc2[3] = {j0[2], j1[2], j2[2]};
            double ji[9];
            if (!inv3_cols(jc0, jc1, jc2, ji)) break;
            double ds0 = ji[0] * r0 + ji[1] * r1 + ji[2] * r2;
            double ds1 = ji[3] * r0 + ji[4] * r1 + ji[5] * r2;
            double ds2 = ji[6] * r0 + ji[7] * r1 + ji[8] * r2;
            s[0] -= ds0; s[1] -= ds1; s[2] -= ds2;
        }
        if (!(s[0] > 0.0 && s[1] > 0.0 && s[2] > 0.0)) continue;
        double p1[3], y12[3], y13[3], ny[3];
        for (int k = 0; k < 3; ++k) {
            p1[k] = s[0] * f[0][k];
            y12[k] = s[1] * f[1][k] - p1[k];
            y13[k] = s[2] * f[2][k] - p1[k];
        }
        cross3(y12, y13, ny);
        /* R = [y12 y13 ny] * inv([d12 d13 nx]) */
        pose_t *o = &sol[ns];
        for (int rr = 0; rr < 3; ++rr)
            for (int cc = 0; cc < 3; ++cc)
                o->r[rr * 3 + cc] = y12[rr] * xinv[0 * 3 + cc] + y13[rr] * xinv[1 * 3 + cc] + ny[rr] * xinv[2 * 3 + cc];
        for (int rr = 0; rr < 3; ++rr)
            o->t[rr] = p1[rr] - (o->r[rr * 3 + 0] * x[0][0] + o->r[rr * 3 + 1] * x[0][1] + o->r[rr * 3 + 2] * x[0][2]);
        ++ns;
    }
    return ns;
}

/* camera intrinsics, general upper-triangular K (row-major 3x3) */
typedef struct { double fx, fy, cx, cy, skew; } intr_t;

static void bearing(const intr_t *k, double u, double v, double f[3]) {
    double yn = (v - k->cy) / k->fy;
    double xn = (u - k->cx - k->skew * yn) / k->fx;
    double inv = 1.0 / sqrt(xn * xn + yn * yn + 1.0);
    f[0] = xn * inv; f[1] = yn * inv; f[2] = inv;
}

static double reproj_err2_d(const intr_t *k, const pose_t *p, const double x[3], double u, double v) {
    double xc = p->r[0] * x[0] + p->r[1] * x[1] + p->r[2] * x[2] + p->t[0];
    double yc = p->r[3] * x[0] + p->r[4] * x[1] + p->r[5] * x[2] + p->t[1];
    double zc = p->r[6] * x[0] + p->r[7] * x[1] + p->r[8] * x[2] + p->t[2];
    if (!(zc > 1e-9)) return 1e300;
    double xn = xc / zc, yn = yc / zc;
    double du = k->fx * xn + k->skew * yn + k->cx - u;
    double dv = k->fy * yn + k->cy - v;
    return du * du + dv * dv;
}

/* one hypothesis: returns 1 and fills rt12 (fp32 R row-major 9 + t 3) or returns 0 */
static int make_hypothesis(const float *obj, const float *img, uint32_t n, const intr_t *k,
                           uint32_t seed, uint32_t hyp, float rt12[12]) {
    uint32_t idx[4];
    if (!draw4(seed, hyp, n, idx)) return 0;
    double f[3][3], x[4][3];
    for (int j = 0; j < 4; ++j)
        for (int c = 0; c < 3; ++c) x[j][c] = (double)obj[3 * idx[j] + c];
    for (int j = 0; j < 3; ++j) bearing(k, (double)img[2 * idx[j]], (double)img[2 * idx[j] + 1], f[j]);
    pose_t sol[4];
    int ns = p3p(f, (const double(*)[3])x, sol);
    int best = -1;
    double beste = 1e300;
    for (int s = 0; s < ns; ++s) {
        double e = reproj_err2_d(k, &sol[s], x[3], (double)img[2 * idx[3]], (double)img[2 * idx[3] + 1]);
        if (e < beste) { beste = e; best = s; }
    }
    if (best < 0) return 0;
    for (int i = 0; i < 9; ++i) rt12[i] = (float)sol[best].r[i];
    for (int i = 0; i < 3; ++i) rt12[9 + i] = (float)sol[best].t[i];
    for (int i = 0; i < 12; ++i) if (!(fabsf(rt12[i]) < 3.0e38f)) return 0; /* inf/nan guard */
    return 1;
}

/* fp32 inlier test, operation order shared with the CUDA scorer */
static inline int is_inlier_f(const float rt[12], const float kf[5], float X, float Y, float Z,
                              float u, float v, float thr2) {
    float xc = fmaf(rt[0], X, fmaf(rt[1], Y, fmaf(rt[2], Z, rt[9])));
    float yc = fmaf(rt[3], X, fmaf(rt[4], Y, fmaf(rt[5], Z, rt[10])));
    float zc = fmaf(rt[6], X, fmaf(rt[7], Y, fmaf(rt[8], Z, rt[11])));
    if (!(zc > 1e-6f)) return 0;
    float iz = 1.0f / zc;
    float xn = xc * iz, yn = yc * iz;
    float du = fmaf(kf[0], xn, fmaf(kf[4], yn, kf[2])) - u;
    float dv = fmaf(kf[1], yn, kf[3]) - v;
    float e = fmaf(du, du, dv * dv);
    return e <= thr2;
}

/* ---- LM refit (fp64) ---------------------------------------------------------------------- */
static void rodrigues_exp(const double w[3], double r[9]) {
    double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
    double th = sqrt(th2);
    double a, b;
    if (th < 1e-8) { a = 1.0 - th2 / 6.0; b = 0.5 - th2 / 24.0; }
    else { a = sin(th) / th; b = (1.0 - cos(th)) / th2; }
    double wx = w[0], wy = w[1], wz = w[2];
    r[0] = 1.0 - b * (wy * wy + wz * wz); r[1] = -a * wz + b * wx * wy;      r[2] = a * wy + b * wx * wz;
    r[3] = a * wz + b * wx * wy;          r[4] = 1.0 - b * (wx * wx + wz * wz); r[5] = -a * wx + b * wy * wz;
    r[6] = -a * wy + b * wx * wz;         r[7] = a * wx + b * wy * wz;       r[8] = 1.0 - b * (wx * wx + wy * wy);
}

static void mat3_mul(const double a[9], const double b[9], double c[9]) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            c[i * 3 + j] = a[i * 3] * b[j] + a[i * 3 + 1] * b[3 + j] + a[i * 3 + 2] * b[6 + j];
}

/* re-orthonormalise rows by Gram-Schmidt on columns (keeps R a rotation over many updates) */
static void orthonormalize(double r[9]) {
    double c0[3] = {r[0], r[3], r[6]}, c1[3] = {r[1], r[4], r[7]}, c2[3];
    double n0 = 1.0 / sqrt(dot3(c0, c0));
    for (int k = 0; k < 3; ++k) c0[k] *= n0;
    double d = dot3(c0, c1);
    for (int k = 0; k < 3; ++k) c1[k] -= d * c0[k];
    double n1 = 1.0 / sqrt(dot3(c1, c1));
    for (int k = 0; k < 3; ++k) c1[k] *= n1;
    cross3(c0, c1, c2);
    for (int k = 0; k < 3; ++k) { r[k * 3] = c0[k]; r[k * 3 + 1] = c1[k]; r[k * 3 + 2] = c2[k]; }
}

/* accumulate normal equations over masked points; returns cost (sum of squared residuals) */
static double lm_accumulate(const float *obj, const float *img, const uint8_t *mask, uint32_t n,
                            const intr_t *k, const pose_t *p, double H[36], double g[6]) {
    double cost = 0.0;
    if (H) { memset(H, 0, 36 * sizeof(double)); memset(g, 0, 6 * sizeof(double)); }
    for (uint32_t i = 0; i < n; ++i) {
        if (!mask[i]) continue;
        double X = obj[3 * i], Y = obj[3 * i + 1], Z = obj[3 * i + 2];
        double xc = p->r[0] * X + p->r[1] * Y + p->r[2] * Z + p->t[0];
        double yc = p->r[3] * X + p->r[4] * Y + p->r[5] * Z + p->t[1];
        double zc = p->r[6] * X + p->r[7] * Y + p->r[8] * Z + p->t[2];
        if (!(zc > 1e-9)) { cost += 1e12; continue; }
        double iz = 1.0 / zc;
        double xn = xc * iz, yn = yc * iz;
        double ru = k->fx * xn + k->skew * yn + k->cx - (double)img[2 * i];
        double rv = k->fy * yn + k->cy - (double)img[2 * i + 1];
        cost += ru * ru + rv * rv;
        if (!H) continue;
        /* d(u,v)/d(Xc) */
        double a00 = k->fx * iz, a01 = k->skew * iz, a02 = -(k->fx * xn + k->skew * yn) * iz;
        double a11 = k->fy * iz, a12 = -k->fy * yn * iz;
        /* dXc = -[Xc]x w + tau  =>  columns for w: -[Xc]x */
        double ju[6], jv[6];
        /* -[Xc]x = [[0, zc, -yc], [-zc, 0, xc], [yc, -xc, 0]] */
        ju[0] = a01 * (-zc) + a02 * yc;
        ju[1] = a00 * zc + a02 * (-xc);
        ju[2] = a00 * (-yc) + a01 * xc;
        ju[3] = a00; ju[4] = a01; ju[5] = a02;
        jv[0] = a11 * (-zc) + a12 * yc;
        jv[1] = a12 * (-xc);
        jv[2] = a11 * xc;
        jv[3] = 0.0; jv[4] = a11; jv[5] = a12;
        for (int r = 0; r < 6; ++r) {
            g[r] += ju[r] * ru + jv[r] * rv;
            for (int c = r; c < 6; ++c) H[r * 6 + c] += ju[r] * ju[c] + jv[r] * jv[c];
        }
    }
    if (H) for (int r = 0; r < 6; ++r) for (int c = 0; c < r; ++c) H[r * 6 + c] = H[c * 6 + r];
    return cost;
}

/* solve A x = b for SPD 6x6 by Cholesky; returns 0 if not positive definite */
static int chol_solve6(const double A[36], const double b[6], double x[6]) {
    double L[36];
    memset(L, 0, sizeof(L));
    for (int i = 0; i < 6; ++i) {
        for (int j = 0; j <= i; ++j) {
            double s = A[i * 6 + j];
            for (int k = 0; k < j; ++k) s -= L[i * 6 + k] * L[j * 6 + k];
            if (i == j) {
                if (!(s > 0.0)) return 0;
                L[i * 6 + i] = sqrt(s);
            } else L[i * 6 + j] = s / L[j * 6 + j];
        }
    }
    double y[6];
    for (int i = 0; i < 6; ++i) {
        double s = b[i];
        for (int k = 0; k < i; ++k) s -= L[i * 6 + k] * y[k];
        y[i] = s / L[i * 6 + i];
    }
    for (int i = 5; i >= 0; --i) {
        double s = y[i];
        for (int k = i + 1; k < 6; ++k) s -= L[k * 6 + i] * x[k];
        x[i] = s / L[i * 6 + i];
    }
    return 1;
}

#define GNB_LM_MAX_ITERS 30

static void lm_refit(const float *obj, const float *img, const uint8_t *mask, uint32_t n,
                     const intr_t *k, pose_t *p) {
    double H[36], g[6], lambda = 1e-3;
    double cost = lm_accumulate(obj, img, mask, n, k, p, H, g);
    for (int it = 0; it < GNB_LM_MAX_ITERS; ++it) {
        double A[36], nb[6], dx[6];
        memcpy(A, H, sizeof(A));
        for (int i = 0; i < 6; ++i) { A[i * 6 + i] += lambda * (H[i * 6 + i] + 1e-12); nb[i] = -g[i]; }
        if (!chol_solve6(A, nb, dx)) { lambda *= 10.0; if (lambda > 1e12) break; continue; }
        pose_t q;
        double dr[9];
        rodrigues_exp(dx, dr);
        mat3_mul(dr, p->r, q.r);
        orthonormalize(q.r);
        /* t' = dR t + tau */
        for (int i = 0; i < 3; ++i)
            q.t[i] = dr[i * 3] * p->t[0] + dr[i * 3 + 1] * p->t[1] + dr[i * 3 + 2] * p->t[2] + dx[3 + i];
        double c2 = lm_accumulate(obj, img, mask, n, k, &q, 0, 0);
        if (c2 < cost) {
            double rel = (cost - c2) / (cost + 1e-300);
            double step = sqrt(dx[0] * dx[0] + dx[1] * dx[1] + dx[2] * dx[2] + dx[3] * dx[3] + dx[4] * dx[4] + dx[5] * dx[5]);
            *p = q;
            cost = lm_accumulate(obj, img, mask, n, k, p, H, g);
            lambda *= 0.1; if (lambda < 1e-12) lambda = 1e-12;
            if (rel < 1e-12 || step < 1e-10) break;
        } else {
            double step = sqrt(dx[0] * dx[0] + dx[1] * dx[1] + dx[2] * dx[2] + dx[3] * dx[3] + dx[4] * dx[4] + dx[5] * dx[5]);
            if (step < 1e-9) break; /* no measurable improvement left */
            lambda *= 10.0;
            if (lambda > 1e12) break;
        }
    }
}

/*
 * Full solve.  Outputs (any may be NULL): counts[iters] (-1 = invalid hypothesis), hyp_rt
 * [iters*12] fp32, mask[n], R[9], t[3], stats[2] = {best hypothesis, inlier count}.
 * Returns 0 ok, 1 = no valid model / fewer than 4 inliers.
 */
int gnbref_solve_pnp_ransac(const float *obj, const float *img, uint32_t n, const double kmat[9],
                            uint32_t iters, float thr_px, uint32_t seed, int refine,
                            int32_t *counts, float *hyp_rt, uint8_t *mask, double *R, double *t,
                            int32_t *stats) {
    intr_t k = {kmat[0], kmat[4], kmat[2], kmat[5], kmat[1]};
    float kf[5] = {(float)k.fx, (float)k.fy, (float)k.cx, (float)k.cy, (float)k.skew};
    float thr2 = thr_px * thr_px;
    int32_t best = -1, bestc = -1;
    float best_rt[12];
    for (uint32_t h = 0; h < iters; ++h) {
        float rt[12];
        int32_t c = -1;
        if (make_hypothesis(obj, img, n, &k, seed, h, rt)) {
            c = 0;
            for (uint32_t i = 0; i < n; ++i)
                c += is_inlier_f(rt, kf, obj[3 * i], obj[3 * i + 1], obj[3 * i + 2], img[2 * i], img[2 * i + 1], thr2);
        } else {
            for (int i = 0; i < 12; ++i) rt[i] = 0.0f;
        }
        if (counts) counts[h] = c;
        if (hyp_rt) memcpy(hyp_rt + 12 * h, rt, sizeof(rt));
        if (c > bestc) { bestc = c; best = (int32_t)h; memcpy(best_rt, rt, sizeof(rt)); }
    }
    if (stats) { stats[0] = best; stats[1] = bestc; }
    if (best < 0 || bestc < 4) {
        if (mask) memset(mask, 0, n);
        return 1;
    }
    uint8_t local_mask[1];
    (void)local_mask;
    pose_t p;
    for (int i = 0; i < 9; ++i) p.r[i] = (double)best_rt[i];
    for (int i = 0; i < 3; ++i) p.t[i] = (double)best_rt[9 + i];
    if (mask) {
        for (uint32_t i = 0; i < n; ++i)
            mask[i] = (uint8_t)is_inlier_f(best_rt, kf, obj[3 * i], obj[3 * i + 1], obj[3 * i + 2], img[2 * i], img[2 * i + 1], thr2);
        if (refine) { orthonormalize(p.r); lm_refit(obj, img, mask, n, &k, &p); }
    }
    if (R) memcpy(R, p.r, sizeof(p.r));
    if (t) memcpy(t, p.t, sizeof(p.t));
    return 0;
}

/* 3-D points from reference keypoints + DEM: _shared.py:95-102. Returns 1 if an index is out of
 * range (numpy would raise IndexError / wrap negatives; the oracle reports instead). */
int gnbref_points3d(const float *mkp_ref, uint32_t n, const uint8_t *dem, int32_t h, int32_t w, float *obj) {
    int bad = 0;
    for (uint32_t i = 0; i < n; ++i) {
        float x = mkp_ref[2 * i], y = mkp_ref[2 * i + 1];
        int32_t xi = (int32_t)floorf(x), yi = (int32_t)floorf(y);
        float z = 0.0f;
        if (dem) {
            if (xi < 0 || yi < 0 || xi >= w || yi >= h) { bad = 1; }
            else z = (float)dem[(size_t)yi * (size_t)w + (size_t)xi];
        }
        obj[3 * i] = x; obj[3 * i + 1] = y; obj[3 * i + 2] = z;
    }
    return bad;
}
