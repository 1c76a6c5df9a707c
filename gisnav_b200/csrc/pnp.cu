// pnp.cu — K5 (batched P3P hypothesis generation + warp-shuffle inlier scoring) and K6 (fp64 LM
// refit + pixel->WGS84->ECEF tail).  Replaces compute_pose
// (ros/gisnav/gisnav/core/_shared.py:89-125: DEM lookup :95-102, cv2.solvePnPRansac :109-116,
// cv2.Rodrigues :117) and the pose tail (pose_node.py:333-381 with
// _transformations.py:301-327,330-346,369-393).
//
// COMPILED WITH --fmad=false: the hypothesis generator and the scorer follow oracle/pnp_ref.c
// operation for operation (only + - * / sqrt and explicit fma), so hypotheses, inlier counts, the
// winning hypothesis and the inlier mask are bit-identical to the CPU oracle.  The LM refit and
// the tail use tree reductions and device sin/cos and are compared to tolerance.
#include "common.cuh"

#include <math.h>

#define PNP_MAX_DRAW_TRIES 64
#define PNP_LM_MAX_ITERS 30

struct Pose { double r[9]; double t[3]; };
struct Intr { double fx, fy, cx, cy, skew; };

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}
__device__ __forceinline__ uint32_t rnd32(uint32_t seed, uint32_t hyp, uint32_t ctr) {
    return hash32(seed ^ hash32(hyp * 0x9E3779B9U + ctr * 0x85EBCA6BU + 0x165667B1U));
}
__device__ int draw4(uint32_t seed, uint32_t hyp, uint32_t n, uint32_t idx[4]) {
    if (n < 4) return 0;
    uint32_t ctr = 0;
    for (int j = 0; j < 4; ++j) {
        int ok = 0;
        while (!ok && ctr < PNP_MAX_DRAW_TRIES) {
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

__device__ __forceinline__ void cross3(const double a[3], const double b[3], double c[3]) {
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
}
__device__ __forceinline__ double dot3(const double a[3], const double b[3]) {
    return a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
}
__device__ int inv3_cols(const double c0[3], const double c1[3], const double c2[3], double inv[9]) {
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
__device__ int quad_roots(double b, double c, double r[2]) {
    double disc = b * b - 4.0 * c;
    if (disc < 0.0) return 0;
    double s = sqrt(disc);
    double q = (b >= 0.0) ? -0.5 * (b + s) : -0.5 * (b - s);
    r[0] = q;
    r[1] = (q != 0.0) ? c / q : 0.0;
    return 2;
}
__device__ int quartic_roots(double a4, double a3, double a2, double a1, double a0, double roots[4]) {
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
    for (int it = 0; it < 3; ++it) {
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

// P3P, see oracle/pnp_ref.c for the derivation
__device__ int p3p(const double f[3][3], const double x[4][3], Pose sol[4]) {
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
    if (!(dot3(nx, nx) > 1e-12 * b2 * c2)) return 0;
    double ca = dot3(f[1], f[2]), cb = dot3(f[0], f[2]), cg = dot3(f[0], f[1]);
    double K = (a2 - c2) / b2;
    double n2 = K - 1.0, n1 = -2.0 * K * cb, n0 = K + 1.0;
    double e1 = -2.0 * ca, e0 = 2.0 * cg;
    double dd2 = e1 * e1, dd1 = 2.0 * e1 * e0, dd0 = e0 * e0;
    double nn4 = n2 * n2, nn3 = 2.0 * n2 * n1, nn2 = 2.0 * n2 * n0 + n1 * n1, nn1 = 2.0 * n1 * n0, nn0 = n0 * n0;
    double nd3 = n2 * e1, nd2 = n2 * e0 + n1 * e1, nd1 = n1 * e0 + n0 * e1, nd0 = n0 * e0;
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
        for (int it = 0; it < 2; ++it) {
            double r0 = s[1] * s[1] + s[2] * s[2] - 2.0 * s[1] * s[2] * ca - a2;
            double r1 = s[0] * s[0] + s[2] * s[2] - 2.0 * s[0] * s[2] * cb - b2;
            double r2 = s[0] * s[0] + s[1] * s[1] - 2.0 * s[0] * s[1] * cg - c2;
            double j0[3] = {0.0, 2.0 * (s[1] - s[2] * ca), 2.0 * (s[2] - s[1] * ca)};
            double j1[3] = {2.0 * (s[0] - s[2] * cb), 0.0, 2.0 * (s[2] - s[0] * cb)};
            double j2[3] = {2.0 * (s[0] - s[1] * cg), 2.0 * (s[1] - s[0] * cg), 0.0};
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
        Pose* o = &sol[ns];
        for (int rr = 0; rr < 3; ++rr)
            for (int cc = 0; cc < 3; ++cc)
                o->r[rr * 3 + cc] = y12[rr] * xinv[0 * 3 + cc] + y13[rr] * xinv[1 * 3 + cc] + ny[rr] * xinv[2 * 3 + cc];
        for (int rr = 0; rr < 3; ++rr)
            o->t[rr] = p1[rr] - (o->r[rr * 3 + 0] * x[0][0] + o->r[rr * 3 + 1] * x[0][1] + o->r[rr * 3 + 2] * x[0][2]);
        ++ns;
    }
    return ns;
}

__device__ void bearing(const Intr& k, double u, double v, double f[3]) {
    double yn = (v - k.cy) / k.fy;
    double xn = (u - k.cx - k.skew * yn) / k.fx;
    double inv = 1.0 / sqrt(xn * xn + yn * yn + 1.0);
    f[0] = xn * inv; f[1] = yn * inv; f[2] = inv;
}

__device__ double reproj_err2_d(const Intr& k, const Pose& p, const double x[3], double u, double v) {
    double xc = p.r[0] * x[0] + p.r[1] * x[1] + p.r[2] * x[2] + p.t[0];
    double yc = p.r[3] * x[0] + p.r[4] * x[1] + p.r[5] * x[2] + p.t[1];
    double zc = p.r[6] * x[0] + p.r[7] * x[1] + p.r[8] * x[2] + p.t[2];
    if (!(zc > 1e-9)) return 1e300;
    double xn = xc / zc, yn = yc / zc;
    double du = k.fx * xn + k.skew * yn + k.cx - u;
    double dv = k.fy * yn + k.cy - v;
    return du * du + dv * dv;
}

__device__ __forceinline__ int is_inlier_f(const float rt[12], const float kf[5], float X, float Y, float Z, float u,
                                           float v, float thr2) {
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

// ------------------------------------------------------------------------------------------------
// 3-D points: (x, y, dem[floor(y), floor(x)])  (_shared.py:95-102)
__global__ void __launch_bounds__(256) points3d_kernel(const float* __restrict__ mkp_ref, const int* __restrict__ match_count,
                                                       int k_cap, const uint8_t* __restrict__ dem, int dem_h, int dem_w,
                                                       int has_dem, float* __restrict__ obj, int* __restrict__ range_flag) {
    const int pair = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = match_count[pair];
    if (i >= n) return;
    const float x = mkp_ref[((size_t)pair * k_cap + i) * 2 + 0], y = mkp_ref[((size_t)pair * k_cap + i) * 2 + 1];
    float z = 0.f;
    if (has_dem) {
        const int xi = (int)floorf(x), yi = (int)floorf(y);
        if (xi < 0 || yi < 0 || xi >= dem_w || yi >= dem_h) range_flag[pair] = 1;
        else z = (float)dem[(size_t)pair * dem_h * dem_w + (size_t)yi * dem_w + xi];
    }
    float* o = obj + ((size_t)pair * k_cap + i) * 3;
    o[0] = x; o[1] = y; o[2] = z;
}

// one thread per hypothesis (fp64 P3P)
__global__ void __launch_bounds__(128) hypothesis_kernel(const float* __restrict__ obj, const float* __restrict__ img,
                                                         const int* __restrict__ match_count, int k_cap,
                                                         const double* __restrict__ kmat, int iters, uint32_t seed,
                                                         int min_matches, float* __restrict__ hyp,
                                                         int* __restrict__ hyp_count) {
    const int pair = blockIdx.y;
    const int h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= iters) return;
    const int n = match_count[pair];
    float rt[12];
    for (int i = 0; i < 12; ++i) rt[i] = 0.f;
    int valid = 0;
    if (n >= min_matches && n >= 4) {
        const double* km = kmat + (size_t)pair * 9;
        Intr k = {km[0], km[4], km[2], km[5], km[1]};
        const float* o = obj + (size_t)pair * k_cap * 3;
        const float* im = img + (size_t)pair * k_cap * 2;
        uint32_t idx[4];
        if (draw4(seed, (uint32_t)h, (uint32_t)n, idx)) {
            double f[3][3], x[4][3];
            for (int j = 0; j < 4; ++j)
                for (int c = 0; c < 3; ++c) x[j][c] = (double)o[3 * idx[j] + c];
            for (int j = 0; j < 3; ++j) bearing(k, (double)im[2 * idx[j]], (double)im[2 * idx[j] + 1], f[j]);
            Pose sol[4];
            int ns = p3p(f, x, sol);
            int best = -1;
            double beste = 1e300;
            for (int s = 0; s < ns; ++s) {
                double e = reproj_err2_d(k, sol[s], x[3], (double)im[2 * idx[3]], (double)im[2 * idx[3] + 1]);
                if (e < beste) { beste = e; best = s; }
            }
            if (best >= 0) {
                valid = 1;
                for (int i = 0; i < 9; ++i) rt[i] = (float)sol[best].r[i];
                for (int i = 0; i < 3; ++i) rt[9 + i] = (float)sol[best].t[i];
                for (int i = 0; i < 12; ++i) if (!(fabsf(rt[i]) < 3.0e38f)) valid = 0;
                if (!valid) for (int i = 0; i < 12; ++i) rt[i] = 0.f;
            }
        }
    }
    float* ho = hyp + ((size_t)pair * iters + h) * 12;
    for (int i = 0; i < 12; ++i) ho[i] = rt[i];
    hyp_count[(size_t)pair * iters + h] = valid ? 0 : -1;
}

// scoring: points staged in shared memory once per CTA; one warp per hypothesis, lanes stride over
// points, inlier count by warp-shuffle reduction.
__global__ void __launch_bounds__(256) score_kernel(const float* __restrict__ obj, const float* __restrict__ img,
                                                    const int* __restrict__ match_count, int k_cap,
                                                    const double* __restrict__ kmat, int iters, float thr_px,
                                                    const float* __restrict__ hyp, int* __restrict__ hyp_count) {
    extern __shared__ float pts[];  // [5][n]: X,Y,Z,u,v
    const int pair = blockIdx.y;
    const int n = match_count[pair];
    if (n < 4) return;
    float* X = pts; float* Y = X + n; float* Z = Y + n; float* U = Z + n; float* V = U + n;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        X[i] = obj[((size_t)pair * k_cap + i) * 3 + 0];
        Y[i] = obj[((size_t)pair * k_cap + i) * 3 + 1];
        Z[i] = obj[((size_t)pair * k_cap + i) * 3 + 2];
        U[i] = img[((size_t)pair * k_cap + i) * 2 + 0];
        V[i] = img[((size_t)pair * k_cap + i) * 2 + 1];
    }
    __syncthreads();
    const double* km = kmat + (size_t)pair * 9;
    const float kf[5] = {(float)km[0], (float)km[4], (float)km[2], (float)km[5], (float)km[1]};
    const float thr2 = thr_px * thr_px;
    const int warps = blockDim.x >> 5, wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int h = blockIdx.x * warps + wid; h < iters; h += gridDim.x * warps) {
        if (hyp_count[(size_t)pair * iters + h] < 0) continue;
        float rt[12];
        const float* hp = hyp + ((size_t)pair * iters + h) * 12;
#pragma unroll
        for (int i = 0; i < 12; ++i) rt[i] = hp[i];
        int c = 0;
        for (int i = lane; i < n; i += 32) c += is_inlier_f(rt, kf, X[i], Y[i], Z[i], U[i], V[i], thr2);
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) c += __shfl_xor_sync(0xffffffffu, c, s);
        if (lane == 0) hyp_count[(size_t)pair * iters + h] = c;
    }
}

// ------------------------------------------------------------------------------------------------
// K6 helpers
__device__ void rodrigues_exp(const double w[3], double r[9]) {
    double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
    double th = sqrt(th2);
    double a, b;
    if (th < 1e-8) { a = 1.0 - th2 / 6.0; b = 0.5 - th2 / 24.0; }
    else { a = sin(th) / th; b = (1.0 - cos(th)) / th2; }
    double wx = w[0], wy = w[1], wz = w[2];
    r[0] = 1.0 - b * (wy * wy + wz * wz); r[1] = -a * wz + b * wx * wy;         r[2] = a * wy + b * wx * wz;
    r[3] = a * wz + b * wx * wy;          r[4] = 1.0 - b * (wx * wx + wz * wz); r[5] = -a * wx + b * wy * wz;
    r[6] = -a * wy + b * wx * wz;         r[7] = a * wx + b * wy * wz;          r[8] = 1.0 - b * (wx * wx + wy * wy);
}
__device__ void mat3_mul(const double a[9], const double b[9], double c[9]) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) c[i * 3 + j] = a[i * 3] * b[j] + a[i * 3 + 1] * b[3 + j] + a[i * 3 + 2] * b[6 + j];
}
__device__ void orthonormalize(double r[9]) {
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
__device__ int chol_solve6(const double A[36], const double b[6], double x[6]) {
    double L[36];
    for (int i = 0; i < 36; ++i) L[i] = 0.0;
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

// pose tail — mirrors oracle/tail_ref.py::pose_tail. Returns 0 ok / GNB_SOFT_OUT_OF_BOUNDS.
__device__ int pose_tail_dev(const double r[9], const double t[3], const double A[12], int ref_h, int ref_w,
                             double ecef[3], double quat[4], double lla[3]) {
    // C = -R^T t
    double c[3];
    for (int i = 0; i < 3; ++i) c[i] = -(r[0 * 3 + i] * t[0] + r[1 * 3 + i] * t[1] + r[2 * 3 + i] * t[2]);
    // int() truncation toward zero; NaN/huge values fail the range test
    if (!(fabs(c[0]) < 2.0e9 && fabs(c[1]) < 2.0e9)) return GNB_SOFT_OUT_OF_BOUNDS;
    const long long xi = (long long)c[0], yi = (long long)c[1];
    if (!(0 <= xi && xi <= ref_h && 0 <= yi && yi <= ref_w)) return GNB_SOFT_OUT_OF_BOUNDS;  // sic: pose_node.py:340
    for (int i = 0; i < 3; ++i) lla[i] = A[i * 4 + 0] * c[0] + A[i * 4 + 1] * c[1] + A[i * 4 + 2] * c[2] + A[i * 4 + 3];
    const double WA = 6378137.0, WF = 1.0 / 298.257223563, WE2 = WF * (2.0 - WF);
    const double D2R = 0.017453292519943295;
    const double lam = lla[0] * D2R, phi = lla[1] * D2R;
    const double sl = sin(lam), cl = cos(lam), sp = sin(phi), cp = cos(phi);
    const double nn = WA / sqrt(1.0 - WE2 * sp * sp);
    ecef[0] = (nn + lla[2]) * cp * cl;
    ecef[1] = (nn + lla[2]) * cp * sl;
    ecef[2] = (nn * (1.0 - WE2) + lla[2]) * sp;
    // R_n = A[:3,:3] / column norms; rot_enu = R_n R^T
    double rn[9];
    for (int j = 0; j < 3; ++j) {
        double nrm = sqrt(A[0 * 4 + j] * A[0 * 4 + j] + A[1 * 4 + j] * A[1 * 4 + j] + A[2 * 4 + j] * A[2 * 4 + j]);
        for (int i = 0; i < 3; ++i) rn[i * 3 + j] = A[i * 4 + j] / nrm;
    }
    double rt[9], renu[9], e2e[9], m[9];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) rt[i * 3 + j] = r[j * 3 + i];
    mat3_mul(rn, rt, renu);
    e2e[0] = -sl; e2e[1] = -sp * cl; e2e[2] = cp * cl;
    e2e[3] = cl;  e2e[4] = -sp * sl; e2e[5] = cp * sl;
    e2e[6] = 0.0; e2e[7] = cp;       e2e[8] = sp;
    mat3_mul(e2e, renu, m);
    // quaternion_from_matrix: Gram-Schmidt on columns (== transforms3d.affines.decompose),
    // flip first column if det < 0, then rotation -> quaternion with w >= 0
    double c0[3] = {m[0], m[3], m[6]}, c1[3] = {m[1], m[4], m[7]}, c2[3] = {m[2], m[5], m[8]};
    double n0 = sqrt(dot3(c0, c0));
    for (int k = 0; k < 3; ++k) c0[k] /= n0;
    double d01 = dot3(c0, c1);
    for (int k = 0; k < 3; ++k) c1[k] -= d01 * c0[k];
    double n1 = sqrt(dot3(c1, c1));
    for (int k = 0; k < 3; ++k) c1[k] /= n1;
    double d02 = dot3(c0, c2), d12 = dot3(c1, c2);
    for (int k = 0; k < 3; ++k) c2[k] -= d02 * c0[k] + d12 * c1[k];
    double n2 = sqrt(dot3(c2, c2));
    for (int k = 0; k < 3; ++k) c2[k] /= n2;
    double cr[3];
    cross3(c1, c2, cr);
    if (dot3(c0, cr) < 0.0) for (int k = 0; k < 3; ++k) c0[k] = -c0[k];
    const double q00 = c0[0], q10 = c0[1], q20 = c0[2], q01 = c1[0], q11 = c1[1], q21 = c1[2], q02 = c2[0], q12 = c2[1], q22 = c2[2];
    const double tr = q00 + q11 + q22;
    double qw, qx, qy, qz;
    if (tr > 0.0) {
        double s = sqrt(tr + 1.0) * 2.0;
        qw = 0.25 * s; qx = (q21 - q12) / s; qy = (q02 - q20) / s; qz = (q10 - q01) / s;
    } else if (q00 > q11 && q00 > q22) {
        double s = sqrt(1.0 + q00 - q11 - q22) * 2.0;
        qw = (q21 - q12) / s; qx = 0.25 * s; qy = (q01 + q10) / s; qz = (q02 + q20) / s;
    } else if (q11 > q22) {
        double s = sqrt(1.0 + q11 - q00 - q22) * 2.0;
        qw = (q02 - q20) / s; qx = (q01 + q10) / s; qy = 0.25 * s; qz = (q12 + q21) / s;
    } else {
        double s = sqrt(1.0 + q22 - q00 - q11) * 2.0;
        qw = (q10 - q01) / s; qx = (q02 + q20) / s; qy = (q12 + q21) / s; qz = 0.25 * s;
    }
    if (qw < 0.0) { qw = -qw; qx = -qx; qy = -qy; qz = -qz; }
    quat[0] = qx; quat[1] = qy; quat[2] = qz; quat[3] = qw;
    return 0;
}

#define FIN_THREADS 256
#define FIN_NACC 28  // 21 upper-triangular H + 6 g + cost

// block-wide deterministic sum of FIN_NACC doubles per thread -> red[0..FIN_NACC) (valid for all)
__device__ void block_reduce_acc(double acc[FIN_NACC], double* red /* [8][FIN_NACC] + [FIN_NACC] */) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < FIN_NACC; ++i) {
        double v = acc[i];
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
        if (lane == 0) red[wid * FIN_NACC + i] = v;
    }
    __syncthreads();
    if (threadIdx.x < FIN_NACC) {
        double v = 0.0;
        for (int w = 0; w < FIN_THREADS / 32; ++w) v += red[w * FIN_NACC + threadIdx.x];
        red[(FIN_THREADS / 32) * FIN_NACC + threadIdx.x] = v;
    }
    __syncthreads();
}

__device__ void lm_accumulate_thread(const float* o, const float* im, const uint8_t* mask, int n, const Intr& k,
                                     const Pose& p, double acc[FIN_NACC]) {
    for (int i = 0; i < FIN_NACC; ++i) acc[i] = 0.0;
    for (int i = threadIdx.x; i < n; i += FIN_THREADS) {
        if (!mask[i]) continue;
        double X = o[3 * i], Y = o[3 * i + 1], Z = o[3 * i + 2];
        double xc = p.r[0] * X + p.r[1] * Y + p.r[2] * Z + p.t[0];
        double yc = p.r[3] * X + p.r[4] * Y + p.r[5] * Z + p.t[1];
        double zc = p.r[6] * X + p.r[7] * Y + p.r[8] * Z + p.t[2];
        if (!(zc > 1e-9)) { acc[27] += 1e12; continue; }
        double iz = 1.0 / zc;
        double xn = xc * iz, yn = yc * iz;
        double ru = k.fx * xn + k.skew * yn + k.cx - (double)im[2 * i];
        double rv = k.fy * yn + k.cy - (double)im[2 * i + 1];
        acc[27] += ru * ru + rv * rv;
        double a00 = k.fx * iz, a01 = k.skew * iz, a02 = -(k.fx * xn + k.skew * yn) * iz;
        double a11 = k.fy * iz, a12 = -k.fy * yn * iz;
        double ju[6], jv[6];
        ju[0] = a01 * (-zc) + a02 * yc;
        ju[1] = a00 * zc + a02 * (-xc);
        ju[2] = a00 * (-yc) + a01 * xc;
        ju[3] = a00; ju[4] = a01; ju[5] = a02;
        jv[0] = a11 * (-zc) + a12 * yc;
        jv[1] = a12 * (-xc);
        jv[2] = a11 * xc;
        jv[3] = 0.0; jv[4] = a11; jv[5] = a12;
        int q = 0;
        for (int r = 0; r < 6; ++r) {
            acc[21 + r] += ju[r] * ru + jv[r] * rv;
            for (int c = r; c < 6; ++c) acc[q++] += ju[r] * ju[c] + jv[r] * jv[c];
        }
    }
}

// one CTA per pair: winner selection, inlier mask, LM refit, tail
__global__ void __launch_bounds__(FIN_THREADS) finalize_kernel(
    const float* __restrict__ obj, const float* __restrict__ img, const int* __restrict__ match_count, int k_cap,
    const double* __restrict__ kmat, const double* __restrict__ affine, int iters, float thr_px, int min_matches,
    int refine, const float* __restrict__ hyp, const int* __restrict__ hyp_count, const int* __restrict__ range_flag,
    const int* __restrict__ kp_count, int slot_a0, int stride_a, int slot_b0, int ref_h, int ref_w, int do_tail,
    uint8_t* __restrict__ inlier_mask, PairOut* __restrict__ out) {
    __shared__ int s_best_c[FIN_THREADS / 32], s_best_h[FIN_THREADS / 32];
    __shared__ int s_winner_h, s_winner_c, s_ninl;
    __shared__ double red[(FIN_THREADS / 32 + 1) * FIN_NACC];
    __shared__ Pose s_p, s_q;
    __shared__ int s_state;  // 0 = keep iterating, 1 = stop
    const int pair = blockIdx.x, tid = threadIdx.x;
    const int n = match_count[pair];
    PairOut* po = out + pair;
    if (tid == 0) {
        po->n_kp_qry = kp_count ? kp_count[slot_a0 + pair * stride_a] : 0;
        po->n_kp_ref = kp_count ? kp_count[slot_b0 + pair] : 0;
        po->n_matches = n;
        po->n_inliers = 0;
        po->best_hypothesis = -1;
        po->status = GNB_OK;
        for (int i = 0; i < 9; ++i) po->r[i] = 0.0;
        for (int i = 0; i < 3; ++i) { po->t[i] = 0.0; po->ecef[i] = 0.0; po->lla[i] = 0.0; }
        for (int i = 0; i < 4; ++i) po->quat[i] = 0.0;
    }
    uint8_t* mask = inlier_mask + (size_t)pair * k_cap;
    if (n < min_matches || n < 4) {
        if (tid == 0) po->status = GNB_SOFT_TOO_FEW_MATCHES;
        for (int i = tid; i < max(n, 0); i += FIN_THREADS) mask[i] = 0;
        return;
    }
    if (range_flag[pair]) {
        if (tid == 0) po->status = GNB_E_RANGE;
        return;
    }
    // winner: max count, lowest index on ties
    int bc = -1, bh = 0x7fffffff;
    for (int h = tid; h < iters; h += FIN_THREADS) {
        const int c = hyp_count[(size_t)pair * iters + h];
        if (c > bc) { bc = c; bh = h; }
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        const int oc = __shfl_xor_sync(0xffffffffu, bc, s), oh = __shfl_xor_sync(0xffffffffu, bh, s);
        if (oc > bc || (oc == bc && oh < bh)) { bc = oc; bh = oh; }
    }
    if ((tid & 31) == 0) { s_best_c[tid >> 5] = bc; s_best_h[tid >> 5] = bh; }
    __syncthreads();
    if (tid == 0) {
        int c = s_best_c[0], h = s_best_h[0];
        for (int w = 1; w < FIN_THREADS / 32; ++w)
            if (s_best_c[w] > c || (s_best_c[w] == c && s_best_h[w] < h)) { c = s_best_c[w]; h = s_best_h[w]; }
        s_winner_c = c; s_winner_h = h; s_ninl = 0;
    }
    __syncthreads();
    if (s_winner_c < 4) {
        if (tid == 0) { po->status = GNB_SOFT_PNP_FAILED; po->best_hypothesis = s_winner_c < 0 ? -1 : s_winner_h; }
        for (int i = tid; i < n; i += FIN_THREADS) mask[i] = 0;
        return;
    }
    const double* km = kmat + (size_t)pair * 9;
    const Intr k = {km[0], km[4], km[2], km[5], km[1]};
    const float kf[5] = {(float)k.fx, (float)k.fy, (float)k.cx, (float)k.cy, (float)k.skew};
    const float thr2 = thr_px * thr_px;
    const float* o = obj + (size_t)pair * k_cap * 3;
    const float* im = img + (size_t)pair * k_cap * 2;
    float rt[12];
    const float* hp = hyp + ((size_t)pair * iters + s_winner_h) * 12;
    for (int i = 0; i < 12; ++i) rt[i] = hp[i];
    int cnt = 0;
    for (int i = tid; i < n; i += FIN_THREADS) {
        const int in = is_inlier_f(rt, kf, o[3 * i], o[3 * i + 1], o[3 * i + 2], im[2 * i], im[2 * i + 1], thr2);
        mask[i] = (uint8_t)in;
        cnt += in;
    }
    atomicAdd(&s_ninl, cnt);
    if (tid == 0) {
        for (int i = 0; i < 9; ++i) s_p.r[i] = (double)rt[i];
        for (int i = 0; i < 3; ++i) s_p.t[i] = (double)rt[9 + i];
        if (refine) orthonormalize(s_p.r);
        s_state = refine ? 0 : 1;
    }
    __syncthreads();
    if (refine) {
        double acc[FIN_NACC];
        double H[36], g[6], cost = 0.0, lambda = 1e-3;  // thread 0 only
        lm_accumulate_thread(o, im, mask, n, k, s_p, acc);
        block_reduce_acc(acc, red);
        const double* tot = red + (FIN_THREADS / 32) * FIN_NACC;
        if (tid == 0) {
            int q = 0;
            for (int r = 0; r < 6; ++r) for (int c = r; c < 6; ++c) { H[r * 6 + c] = tot[q]; H[c * 6 + r] = tot[q]; ++q; }
            for (int r = 0; r < 6; ++r) g[r] = tot[21 + r];
            cost = tot[27];
        }
        double dxs[6] = {0, 0, 0, 0, 0, 0};
        for (int it = 0; it < PNP_LM_MAX_ITERS; ++it) {
            // thread 0 proposes a candidate (possibly raising lambda until the system is PD)
            if (tid == 0) {
                int have = 0;
                while (!have) {
                    double A[36], nb[6];
                    for (int i = 0; i < 36; ++i) A[i] = H[i];
                    for (int i = 0; i < 6; ++i) { A[i * 6 + i] += lambda * (H[i * 6 + i] + 1e-12); nb[i] = -g[i]; }
                    if (chol_solve6(A, nb, dxs)) have = 1;
                    else { lambda *= 10.0; if (lambda > 1e12) break; }
                }
                if (!have) s_state = 1;
                else {
                    double dr[9];
                    rodrigues_exp(dxs, dr);
                    mat3_mul(dr, s_p.r, s_q.r);
                    orthonormalize(s_q.r);
                    for (int i = 0; i < 3; ++i)
                        s_q.t[i] = dr[i * 3] * s_p.t[0] + dr[i * 3 + 1] * s_p.t[1] + dr[i * 3 + 2] * s_p.t[2] + dxs[3 + i];
                }
            }
            __syncthreads();
            const int stop_a = s_state;
            __syncthreads();
            if (stop_a) break;
            lm_accumulate_thread(o, im, mask, n, k, s_q, acc);
            block_reduce_acc(acc, red);
            if (tid == 0) {
                const double c2 = tot[27];
                if (c2 < cost) {
                    const double rel = (cost - c2) / (cost + 1e-300);
                    const double step = sqrt(dxs[0] * dxs[0] + dxs[1] * dxs[1] + dxs[2] * dxs[2] + dxs[3] * dxs[3] +
                                             dxs[4] * dxs[4] + dxs[5] * dxs[5]);
                    s_p = s_q;
                    int q = 0;
                    for (int r = 0; r < 6; ++r) for (int c = r; c < 6; ++c) { H[r * 6 + c] = tot[q]; H[c * 6 + r] = tot[q]; ++q; }
                    for (int r = 0; r < 6; ++r) g[r] = tot[21 + r];
                    cost = c2;
                    lambda *= 0.1; if (lambda < 1e-12) lambda = 1e-12;
                    if (rel < 1e-12 || step < 1e-10) s_state = 1;
                } else {
                    const double step = sqrt(dxs[0] * dxs[0] + dxs[1] * dxs[1] + dxs[2] * dxs[2] + dxs[3] * dxs[3] +
                                             dxs[4] * dxs[4] + dxs[5] * dxs[5]);
                    if (step < 1e-9) s_state = 1;  // no measurable improvement left
                    lambda *= 10.0;
                    if (lambda > 1e12) s_state = 1;
                }
            }
            __syncthreads();
            const int stop_b = s_state;
            __syncthreads();
            if (stop_b) break;
        }
    }
    __syncthreads();
    if (tid == 0) {
        po->n_inliers = s_ninl;
        po->best_hypothesis = s_winner_h;
        for (int i = 0; i < 9; ++i) po->r[i] = s_p.r[i];
        for (int i = 0; i < 3; ++i) po->t[i] = s_p.t[i];
        if (do_tail) {
            const int rc = pose_tail_dev(s_p.r, s_p.t, affine + (size_t)pair * 12, ref_h, ref_w, po->ecef, po->quat, po->lla);
            if (rc) po->status = rc;
        }
    }
}

int gnb_pnp_pairs(gnb_ctx* ctx, int pairs, int dem_h, int dem_w, int has_dem, int ref_h, int ref_w, int do_tail,
                  int min_matches, int use_kp_counts, int stride_a) {
    const int k = ctx->cfg.max_keypoints, iters = ctx->cfg.ransac_iters;
    GNB_CUDA(ctx, cudaMemsetAsync(ctx->range_flag, 0, sizeof(int) * pairs, ctx->stream));
    {
        dim3 grid(ceil_div(k, 256), pairs);
        GNB_KERNEL(ctx, "points3d_kernel", points3d_kernel<<<grid, 256, 0, ctx->stream>>>(ctx->mkp_ref, ctx->match_count, k, ctx->dem, dem_h, dem_w, has_dem,
                                                       ctx->obj, ctx->range_flag));
    }
    {
        dim3 grid(ceil_div(iters, 128), pairs);
        GNB_KERNEL(ctx, "hypothesis_kernel", hypothesis_kernel<<<grid, 128, 0, ctx->stream>>>(ctx->obj, ctx->mkp_qry, ctx->match_count, k, ctx->kmat, iters,
                                                         ctx->cfg.ransac_seed, min_matches, ctx->hyp, ctx->hyp_count));
    }
    {
        const size_t smem = (size_t)5 * k * sizeof(float);
        GNB_CUDA(ctx, gnb_func_smem(ctx, score_kernel, (int)(5 * GNB_MAX_KP * sizeof(float))));
        dim3 grid(ceil_div(iters, 8 * 8), pairs);  // 8 warps per CTA, 8 hypotheses per warp
        GNB_KERNEL(ctx, "score_kernel", score_kernel<<<grid, 256, smem, ctx->stream>>>(ctx->obj, ctx->mkp_qry, ctx->match_count, k, ctx->kmat, iters,
                                                       ctx->cfg.reproj_px, ctx->hyp, ctx->hyp_count));
    }
    GNB_KERNEL(ctx, "finalize_kernel", finalize_kernel<<<pairs, FIN_THREADS, 0, ctx->stream>>>(
        ctx->obj, ctx->mkp_qry, ctx->match_count, k, ctx->kmat, ctx->affine, iters, ctx->cfg.reproj_px,
        min_matches, ctx->cfg.refine, ctx->hyp, ctx->hyp_count, ctx->range_flag, use_kp_counts ? ctx->kp_count : nullptr, 0, stride_a,
        ctx->cfg.max_batch, ref_h, ref_w, do_tail, ctx->inlier_mask, ctx->out_dev));
    return GNB_OK;
}

// standalone tail (gnb_geodetic_tail): one thread, same device function as the fused path
__global__ void tail_kernel(const double* __restrict__ in /* r9 t3 A12 */, int ref_h, int ref_w, double* __restrict__ outv,
                            int* __restrict__ status) {
    double ecef[3], quat[4], lla[3];
    for (int i = 0; i < 3; ++i) { ecef[i] = 0; lla[i] = 0; }
    for (int i = 0; i < 4; ++i) quat[i] = 0;
    *status = pose_tail_dev(in, in + 9, in + 12, ref_h, ref_w, ecef, quat, lla);
    for (int i = 0; i < 3; ++i) { outv[i] = ecef[i]; outv[7 + i] = lla[i]; }
    for (int i = 0; i < 4; ++i) outv[3 + i] = quat[i];
}

int gnb_tail_device(gnb_ctx* ctx, const double* r9, const double* t3, const double* affine12, int ref_h, int ref_w,
                    double* ecef3, double* quat4, double* lla3) {
    double h_in[24];
    memcpy(h_in, r9, 72); memcpy(h_in + 9, t3, 24); memcpy(h_in + 12, affine12, 96);
    double* d = reinterpret_cast<double*>(ctx->stage_a);
    GNB_CUDA(ctx, cudaMemcpyAsync(d, h_in, sizeof(h_in), cudaMemcpyHostToDevice, ctx->stream));
    GNB_KERNEL(ctx, "tail_kernel", tail_kernel<<<1, 1, 0, ctx->stream>>>(d, ref_h, ref_w, d + 24, reinterpret_cast<int*>(d + 40)));
    double h_out[10];
    int st = 0;
    GNB_CUDA(ctx, cudaMemcpyAsync(h_out, d + 24, sizeof(h_out), cudaMemcpyDeviceToHost, ctx->stream));
    GNB_CUDA(ctx, cudaMemcpyAsync(&st, d + 40, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    GNB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    memcpy(ecef3, h_out, 24); memcpy(quat4, h_out + 3, 32); memcpy(lla3, h_out + 7, 24);
    return st;
}
