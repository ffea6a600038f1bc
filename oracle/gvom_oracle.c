/*
 * gvom_oracle.c -- CPU restatement of G-VOM's per-scan voxel-mapping path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity oracle for the CUDA path
 * in gvom_b200/csrc.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it; the product never does.
 *
 * It restates, function by function, what the reference's Numba kernels in
 * /root/reference/scripts/gvom.py compute (cited as gvom.py:LINE below).  The
 * floating-point contraction pattern (which products are fused into FMAs)
 * follows the PTX that Numba 0.65 / NVVM (CUDA 12.9) emits for each kernel
 * (re-derived in DESIGN.md), because voxel indices and
 * ray trip counts are compared bit-exactly.  Compile with -ffp-contract=off so
 * that gcc adds no contractions of its own; every fused product is an explicit
 * fma()/fmaf() call.
 *
 * Parity pinning: checked against dumps of the executed reference
 * (the .npz files under tests/golden, made by tests/golden/make_golden.py on a B200 through
 * Numba-CUDA and in the CPU container through NUMBA_ENABLE_CUDASIM).
 *
 * Compact cell ids: the reference hands them out with an atomic counter in
 * scheduling order (gvom.py:1238,1031,1059); here they are handed out in
 * linear-voxel order within each launch.  Nothing observable depends on it.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    double xy_res, z_res;
    int64_t xy_size, z_size;
    double min_distance;
    double pos_thr, neg_thr, slope_thr;
    double robot_height, robot_radius, ground_to_lidar;
    int64_t xy_eigen_dist, z_eigen_dist;
} gvo_params;

/* point accessors: the reference keeps the caller's dtype on the device
 * (cuda.to_device, gvom.py:115), so float32 clouds are squared in float32
 * and widened only for the divisions. */
typedef struct { const void* p; int is_f32; int64_t stride; } gvo_pts;

static inline double P(const gvo_pts* c, int64_t i, int k) {
    return c->is_f32 ? (double)((const float*)c->p)[i * c->stride + k]
                     : ((const double*)c->p)[i * c->stride + k];
}

/* gvom.py:1145-1149: d2 = x*x + y*y + z*z, contracted as fma(z,z,fma(x,x,y*y)),
 * in the cloud's own precision; rejected when d2 < min_distance^2 (world frame!). */
static inline int too_close(const gvo_pts* c, int64_t i, double min_distance) {
    double d2;
    if (c->is_f32) {
        const float* q = (const float*)c->p + i * c->stride;
        float t = q[1] * q[1];
        t = fmaf(q[0], q[0], t);
        t = fmaf(q[2], q[2], t);
        d2 = (double)t;
    } else {
        const double* q = (const double*)c->p + i * c->stride;
        double t = q[1] * q[1];
        t = fma(q[0], q[0], t);
        t = fma(q[2], q[2], t);
        d2 = t;
    }
    return d2 < min_distance * min_distance;
}

/* ------------------------------------------------------------------------
 * gvom.py:1121-1138  __transform_pointcloud (in place)
 * row r: T[r,3] + fma(p2,T[r,2], fma(p0,T[r,0], p1*T[r,1])), float64 math,
 * stored back in the cloud's dtype.
 * ---------------------------------------------------------------------- */
void gvo_transform(void* pts, int is_f32, int64_t stride, int64_t n, const double* T) {
    for (int64_t i = 0; i < n; ++i) {
        double p0, p1, p2, o[3];
        if (is_f32) { float* q = (float*)pts + i * stride; p0 = q[0]; p1 = q[1]; p2 = q[2]; }
        else { double* q = (double*)pts + i * stride; p0 = q[0]; p1 = q[1]; p2 = q[2]; }
        for (int r = 0; r < 3; ++r) {
            double t = p1 * T[4 * r + 1];
            t = fma(p0, T[4 * r + 0], t);
            t = fma(p2, T[4 * r + 2], t);
            o[r] = T[4 * r + 3] + t;
        }
        if (is_f32) { float* q = (float*)pts + i * stride; q[0] = (float)o[0]; q[1] = (float)o[1]; q[2] = (float)o[2]; }
        else { double* q = (double*)pts + i * stride; q[0] = o[0]; q[1] = o[1]; q[2] = o[2]; }
    }
}

/* ------------------------------------------------------------------------
 * gvom.py:1140-1231  __point_2_map: hit counts + ego->point DDA pass counts.
 * Returns the number of DDA steps that incremented a voxel (work count S).
 * ---------------------------------------------------------------------- */
int64_t gvo_point_2_map(const gvo_params* g, const void* pts, int is_f32, int64_t stride,
                        int64_t n, const double* ego, const double* origin,
                        int32_t* hit, int32_t* total) {
    gvo_pts c = {pts, is_f32, stride};
    const double sxy = (double)g->xy_size, sz = (double)g->z_size;
    int64_t steps = 0;
#pragma omp parallel for schedule(dynamic, 1024) reduction(+ : steps)
    for (int64_t i = 0; i < n; ++i) {
        if (too_close(&c, i, g->min_distance)) continue;             /* :1148 */
        double ex = P(&c, i, 0) / g->xy_res, ey = P(&c, i, 1) / g->xy_res, ez = P(&c, i, 2) / g->z_res;
        double xi = floor(ex - origin[0]), yi = floor(ey - origin[1]), zi = floor(ez - origin[2]);
        int oob = (xi < 0 || xi >= sxy) || (yi < 0 || yi >= sxy) || (zi < 0 || zi >= sz);
        if (!oob) {                                                   /* :1165-1171 */
            int64_t v = (int64_t)xi + (int64_t)yi * g->xy_size + (int64_t)zi * g->xy_size * g->xy_size;
#pragma omp atomic
            hit[v] += 1;
#pragma omp atomic
            total[v] += 1;
        }
        /* ray trace (:1174-1231): float32 state, float64 length */
        float pt[3], s[3];
        pt[0] = (float)(ego[0] / g->xy_res);
        pt[1] = (float)(ego[1] / g->xy_res);
        pt[2] = (float)(ego[2] / g->z_res);
        s[0] = (float)ex - pt[0];
        s[1] = (float)ey - pt[1];
        s[2] = (float)ez - pt[2];
        float l2 = s[0] * s[0];
        l2 = fmaf(s[1], s[1], l2);
        l2 = fmaf(s[2], s[2], l2);
        float L = sqrtf(l2);
        s[0] = s[0] / L; s[1] = s[1] / L; s[2] = s[2] / L;
        float a0 = fabsf(s[0]), a1 = fabsf(s[1]), a2 = fabsf(s[2]);
        float m = fmaxf(a0, fmaxf(a1, a2));
        int k = 0;                                                    /* :1199-1204: later axis wins ties */
        if (m == a1) k = 1;
        if (m == a2) k = 2;
        double lim = (double)L - 1.0;
        if (!(lim > 0.0)) continue;                                   /* while(length < ray_length-1), length=0 */
        float ak = fabsf(s[k]);
        float inc_k = s[k] / ak;
        int k1 = (k + 1) % 3, k2 = (k + 2) % 3;
        float inc1 = s[k1] / ak, inc2 = s[k2] / ak;
        double dlen = fabs(1.0 / (double)s[k]);
        double length = 0.0;
        do {
            pt[k] += inc_k; pt[k1] += inc1; pt[k2] += inc2;
            double x = floor((double)pt[0] - origin[0]);
            if (x < 0 || x >= sxy) break;
            double y = floor((double)pt[1] - origin[1]);
            if (y < 0 || y >= sxy) break;
            double z = floor((double)pt[2] - origin[2]);
            if (z < 0 || z >= sz) break;
            int64_t v = (int64_t)x + (int64_t)y * g->xy_size + (int64_t)z * g->xy_size * g->xy_size;
#pragma omp atomic
            total[v] += 1;
            steps++;
            length += dlen;
        } while (length < lim);
    }
    return steps;
}

/* ------------------------------------------------------------------------
 * gvom.py:1233-1247  __assign_indices + __move_data x2.
 * index_map: >=0 compact id, -1 unknown, < -1 free with passes = -code-1.
 * Returns the cell count; hit_c/total_c must hold at least that many.
 * ---------------------------------------------------------------------- */
int64_t gvo_assign_indices(int64_t V, const int32_t* hit, const int32_t* total,
                           int32_t* index_map, int32_t* hit_c, int32_t* total_c) {
    int64_t c = 0;
    for (int64_t v = 0; v < V; ++v) {
        if (hit[v] > 0) {
            index_map[v] = (int32_t)c;
            if (hit_c) { hit_c[c] = hit[v]; total_c[c] = total[v]; }
            ++c;
        } else {
            index_map[v] = -total[v] - 1;
        }
    }
    return c;
}

/* shared neighbourhood walk of gvom.py:1262-1298 and :1323-1371 */
#define NEIGHBOUR_LOOP(BODY)                                                              \
    for (int64_t i = 0; i < n; ++i) {                                                     \
        if (too_close(&c, i, g->min_distance)) continue;                                  \
        double fx = P(&c, i, 0) / g->xy_res - origin[0];                                  \
        double fy = P(&c, i, 1) / g->xy_res - origin[1];                                  \
        double fz = P(&c, i, 2) / g->z_res - origin[2];                                   \
        double bx = floor(fx), by = floor(fy), bz = floor(fz);                            \
        int32_t x0 = (int32_t)(bx - (double)g->xy_eigen_dist), x1 = (int32_t)(bx + 1.0 + (double)g->xy_eigen_dist); \
        int32_t y0 = (int32_t)(by - (double)g->xy_eigen_dist), y1 = (int32_t)(by + 1.0 + (double)g->xy_eigen_dist); \
        int32_t z0 = (int32_t)(bz - (double)g->z_eigen_dist), z1 = (int32_t)(bz + 1.0 + (double)g->z_eigen_dist);   \
        for (int32_t x = x0; x < x1; ++x) {                                               \
            if (x < 0 || x >= g->xy_size) continue;                                       \
            for (int32_t y = y0; y < y1; ++y) {                                           \
                if (y < 0 || y >= g->xy_size) continue;                                   \
                for (int32_t z = z0; z < z1; ++z) {                                       \
                    if (z < 0 || z >= g->z_size) continue;                                \
                    double lp0 = fx - (double)x, lp1 = fy - (double)y, lp2 = fz - (double)z; \
                    int32_t idx = index_map[(int64_t)x + (int64_t)y * g->xy_size + (int64_t)z * g->xy_size * g->xy_size]; \
                    if (idx < 0) continue;                                                \
                    double* M = metrics + (int64_t)idx * 10;                              \
                    BODY                                                                  \
                }                                                                         \
            }                                                                             \
        }                                                                                 \
    }

/* ------------------------------------------------------------------------
 * gvom.py:1249-1387  __calculate_mean, __normalize_mean, __calculate_covariance,
 * __normalize_covariance.  metrics is float64 [C,10]: mean xyz, cov xx xy xz yy yz
 * zz, count; zero-initialised here (gvom.py:1079-1080).  Points whose own voxel is
 * outside the grid still contribute to in-grid neighbours (quirk, :1262-1279).
 * Returns K = number of (point, occupied neighbour) pairs.
 * ---------------------------------------------------------------------- */
int64_t gvo_moments(const gvo_params* g, const void* pts, int is_f32, int64_t stride, int64_t n,
                    const double* origin, const int32_t* index_map, int64_t C, double* metrics) {
    gvo_pts c = {pts, is_f32, stride};
    int64_t K = 0;
    memset(metrics, 0, sizeof(double) * 10 * (size_t)C);
    NEIGHBOUR_LOOP(M[0] += lp0; M[1] += lp1; M[2] += lp2; M[9] += 1.0; ++K;)
    for (int64_t j = 0; j < C; ++j)                                   /* :1300-1308 */
        for (int a = 0; a < 3; ++a) metrics[j * 10 + a] = metrics[j * 10 + a] / metrics[j * 10 + 9];
    NEIGHBOUR_LOOP(
        M[3] += (lp0 - M[0]) * (lp0 - M[0]);
        M[4] += (lp0 - M[0]) * (lp1 - M[1]);
        M[5] += (lp0 - M[0]) * (lp2 - M[2]);
        M[6] += (lp1 - M[1]) * (lp1 - M[1]);
        M[7] += (lp1 - M[1]) * (lp2 - M[2]);
        M[8] += (lp2 - M[2]) * (lp2 - M[2]);)
    for (int64_t j = 0; j < C; ++j)                                   /* :1375-1387 */
        for (int a = 3; a < 9; ++a) {
            if (metrics[j * 10 + 9] <= 0) metrics[j * 10 + a] = 0;
            else metrics[j * 10 + a] = metrics[j * 10 + a] / metrics[j * 10 + 9];
        }
    return K;
}

/* ------------------------------------------------------------------------
 * gvom.py:1389-1421  __calculate_min_height (float32 min, init 1 at :1085)
 * ---------------------------------------------------------------------- */
void gvo_min_height(const gvo_params* g, const void* pts, int is_f32, int64_t stride, int64_t n,
                    const double* origin, const int32_t* index_map, int64_t C, float* min_height) {
    gvo_pts c = {pts, is_f32, stride};
    const double sxy = (double)g->xy_size, sz = (double)g->z_size;
    for (int64_t j = 0; j < C; ++j) min_height[j] = 1.0f;
    for (int64_t i = 0; i < n; ++i) {
        if (too_close(&c, i, g->min_distance)) continue;
        double xi = floor(P(&c, i, 0) / g->xy_res - origin[0]);
        if (xi < 0 || xi >= sxy) continue;
        double yi = floor(P(&c, i, 1) / g->xy_res - origin[1]);
        if (yi < 0 || yi >= sxy) continue;
        double fz = P(&c, i, 2) / g->z_res - origin[2];
        double zi = floor(fz);
        if (zi < 0 || zi >= sz) continue;
        float lz = (float)(fz - zi);
        int32_t idx = index_map[(int64_t)xi + (int64_t)yi * g->xy_size + (int64_t)zi * g->xy_size * g->xy_size];
        if (idx >= 0 && lz < min_height[idx]) min_height[idx] = lz;
    }
}

/* ------------------------------------------------------------------------
 * gvom.py:1009-1035 __combine_indices (is_last = 0) and :1037-1063
 * __combine_old_indices (is_last = 1: an occupied voxel of the previous
 * combined map survives only where the current code is in [-11,-1]).
 * counter is the running compact-cell count (host int64, gvom.py:231).
 * ---------------------------------------------------------------------- */
void gvo_combine_indices(const gvo_params* g, int64_t* counter, int32_t* combined,
                         const double* combined_origin, const int32_t* old_map,
                         const double* old_origin, int is_last) {
    const int64_t S = g->xy_size, Z = g->z_size;
    const double dx = combined_origin[0] - old_origin[0];
    const double dy = combined_origin[1] - old_origin[1];
    const double dz = combined_origin[2] - old_origin[2];
    for (int64_t z = 0; z < Z; ++z)
        for (int64_t y = 0; y < S; ++y)
            for (int64_t x = 0; x < S; ++x) {
                double ox = (double)x + dx, oy = (double)y + dy, oz = (double)z + dz;
                if (ox >= (double)S || oy >= (double)S || oz >= (double)Z || ox < 0 || oy < 0 || oz < 0) continue;
                int64_t v = x + y * S + z * S * S;
                int64_t vo = (int64_t)(ox + oy * (double)S + oz * (double)S * (double)S);
                int32_t o = old_map[vo], cur = combined[v];
                if (o >= 0 && cur <= -1 && (!is_last || cur >= -11)) {
                    combined[v] = (int32_t)(*counter);
                    *counter += 1;
                } else if (o < -1 && cur <= -1) {
                    combined[v] = cur + o + 1;
                }
            }
}

/* ------------------------------------------------------------------------
 * gvom.py:888-980  __combine_metrics.  combined_* are float32/int32 compact
 * arrays of the combined map; old metrics are float64 for ring-buffer slots and
 * float32 for the previous combined map (old_is_f32).  Math in float64, stores
 * round to float32.
 * ---------------------------------------------------------------------- */
void gvo_combine_metrics(const gvo_params* g, float* cm, int32_t* chit, int32_t* ctot, float* cminh,
                         const int32_t* combined, const double* combined_origin,
                         const void* om, int old_is_f32, const int32_t* ohit, const int32_t* otot,
                         const float* ominh, const int32_t* old_map, const double* old_origin) {
    const int64_t S = g->xy_size, Z = g->z_size;
    const double dx = combined_origin[0] - old_origin[0];
    const double dy = combined_origin[1] - old_origin[1];
    const double dz = combined_origin[2] - old_origin[2];
#pragma omp parallel for schedule(static)
    for (int64_t z = 0; z < Z; ++z)
        for (int64_t y = 0; y < S; ++y)
            for (int64_t x = 0; x < S; ++x) {
                double ox = (double)x + dx, oy = (double)y + dy, oz = (double)z + dz;
                if (ox >= (double)S || oy >= (double)S || oz >= (double)Z || ox < 0 || oy < 0 || oz < 0) continue;
                int32_t ic = combined[x + y * S + z * S * S];
                int32_t io = old_map[(int64_t)(ox + oy * (double)S + oz * (double)S * (double)S)];
                if (ic < 0 || io < 0) continue;
                float* c = cm + (int64_t)ic * 10;
                double o[10];
                for (int a = 0; a < 10; ++a)
                    o[a] = old_is_f32 ? (double)((const float*)om)[(int64_t)io * 10 + a]
                                      : ((const double*)om)[(int64_t)io * 10 + a];
                double n1 = c[9], n2 = o[9], nt = n1 + n2;
                double c0 = c[0], c1 = c[1], c2 = c[2];
                double mx = (c0 * n1 + o[0] * n2) / nt;
                double my = (c1 * n1 + o[1] * n2) / nt;
                double mz = (c2 * n1 + o[2] * n2) / nt;
                const double cd[3] = {c0 - mx, c1 - my, c2 - mz};
                const double od[3] = {o[0] - mx, o[1] - my, o[2] - mz};
                static const int A[6] = {0, 0, 0, 1, 1, 2}, B[6] = {0, 1, 2, 1, 2, 2};
                for (int e = 0; e < 6; ++e) {
                    double v = (n1 * (double)c[3 + e] + n2 * o[3 + e] + n1 * cd[A[e]] * cd[B[e]] +
                                n2 * od[A[e]] * od[B[e]]) / nt;
                    c[3 + e] = (float)v;
                }
                c[0] = (float)mx; c[1] = (float)my; c[2] = (float)mz;
                c[9] = (float)nt;
                chit[ic] += ohit[io];
                ctot[ic] += otot[io];
                float mh = ominh[io];
                if (mh < cminh[ic]) cminh[ic] = mh;
            }
}

/* ------------------------------------------------------------------------
 * gvom.py:1423-1487  __calculate_eigenvalues (closed-form symmetric 3x3)
 * ---------------------------------------------------------------------- */
void gvo_eigenvalues(int64_t C, const float* cm, float* eig) {
    for (int64_t i = 0; i < C; ++i) {
        const float* m = cm + i * 10;
        float xx = m[3], xy = m[4], xz = m[5], yy = m[6], yz = m[7], zz = m[8];
        float p1 = xz * xz;
        p1 = fmaf(xy, xy, p1);
        p1 = fmaf(yz, yz, p1);
        double q = (double)((xx + yy) + zz) / 3.0;
        float* e = eig + i * 3;
        if (p1 == 0.0f) {
            e[0] = fmaxf(xx, fmaxf(yy, zz));
            e[2] = fminf(xx, fminf(yy, zz));
            e[1] = (float)((q * 3.0 - (double)e[0]) - (double)e[2]);
        } else {
            double ax = (double)xx - q, ay = (double)yy - q, az = (double)zz - q;
            double p2 = ay * ay;
            p2 = fma(ax, ax, p2);
            p2 = fma(az, az, p2);
            p2 = fma((double)p1, 2.0, p2);
            double p = sqrt(p2 / 6.0);
            double B0 = ax / p, B1 = (double)xy / p, B2 = (double)xz / p, B3 = ay / p, B4 = (double)yz / p, B5 = az / p;
            double r = B0 * (B3 * B5 - B4 * B4) - B1 * (B1 * B5 - B4 * B2);
            r = fma(B2, B1 * B4 - B3 * B2, r);
            r = r * 0.5;
            double phi;
            if (r <= -1.0) phi = M_PI / 3.0;
            else if (r >= 1.0) phi = 0.0;
            else phi = acos(r) / 3.0;
            e[0] = (float)(q + 2.0 * p * cos(phi));
            e[2] = (float)(q + 2.0 * p * cos(phi + (2.0 * M_PI / 3.0)));
            e[1] = (float)((3.0 * q - (double)e[0]) - (double)e[2]);
        }
    }
}

#define HM(a, x, y) (a)[(int64_t)(x) * S + (int64_t)(y)] /* 2-D maps are indexed [x,y], C order */

/* ------------------------------------------------------------------------
 * gvom.py:560-590  __make_height_map and __make_inferred_height_map
 * (both pre-filled with -1000, :327-331).
 * ---------------------------------------------------------------------- */
void gvo_height_maps(const gvo_params* g, const double* origin, const int32_t* combined,
                     const float* cminh, const double* ego, double* height, double* inferred) {
    const int64_t S = g->xy_size, Z = g->z_size;
    for (int64_t x = 0; x < S; ++x)
        for (int64_t y = 0; y < S; ++y) {
            double h = -1000.0, inf = -1000.0;
            /* PTX has mul+sub here; ptxas (sm_100) contracts them: xp = fma(o+x, res, -ego) */
            double xp = fma(origin[0] + (double)x, g->xy_res, -ego[0]);
            double yp = fma(origin[1] + (double)y, g->xy_res, -ego[1]);
            if (fma(xp, xp, yp * yp) <= g->robot_radius * g->robot_radius) h = ego[2] - g->ground_to_lidar;
            for (int64_t z = 0; z < Z; ++z) {
                int32_t idx = combined[x + y * S + z * S * S];
                if (idx >= 0) { h = (((double)z + (double)cminh[idx]) + origin[2]) * g->z_res; break; }
            }
            for (int64_t z = 0; z < Z; ++z) {
                int32_t idx = combined[x + y * S + z * S * S];
                if (idx < -1) { inf = (origin[2] + (double)z) * g->z_res; break; }
            }
            HM(height, x, y) = h;
            HM(inferred, x, y) = inf;
        }
}

/* ------------------------------------------------------------------------
 * gvom.py:717-805  __calculate_slope: 3x3 LSQ plane.  roughness pre-filled -1,
 * slopes 0 (:342-349).  Residuals use the NORMALISED a0/m, a1/m (quirk).
 * ---------------------------------------------------------------------- */
void gvo_slope(const gvo_params* g, const double* height, double* xs, double* ys, double* rough) {
    const int64_t S = g->xy_size;
    for (int64_t x0 = 0; x0 < S; ++x0)
        for (int64_t y0 = 0; y0 < S; ++y0) {
            HM(xs, x0, y0) = 0.0; HM(ys, x0, y0) = 0.0; HM(rough, x0, y0) = -1.0;
            double px[9], py[9], pz[9];
            int n = 0;
            double sx = 0, sy = 0, sz = 0;
            int64_t xa = x0 - 1 < 0 ? 0 : x0 - 1, xb = x0 + 2 > S ? S : x0 + 2;
            int64_t ya = y0 - 1 < 0 ? 0 : y0 - 1, yb = y0 + 2 > S ? S : y0 + 2;
            for (int64_t x = xa; x < xb; ++x)
                for (int64_t y = ya; y < yb; ++y) {
                    double h = HM(height, x, y);
                    if (h > -1000.0) {
                        px[n] = (double)x * g->xy_res; py[n] = (double)y * g->xy_res; pz[n] = h;
                        sx += px[n]; sy += py[n]; sz += pz[n];
                        ++n;
                    }
                }
            if (n < 3) continue;
            double mx = sx / (double)n, my = sy / (double)n, mz = sz / (double)n;
            double xx = 0, xy = 0, xz = 0, yy = 0, yz = 0;
            for (int i = 0; i < n; ++i) {
                double dx = px[i] - mx, dy = py[i] - my, dz = pz[i] - mz;
                xx = fma(dx, dx, xx); xy = fma(dx, dy, xy); xz = fma(dx, dz, xz);
                yy = fma(dy, dy, yy); yz = fma(dy, dz, yz);
            }
            /* SASS (ptxas sm_100 on Numba's PTX): the second product of each difference
             * is rounded, the first is fused: det = fma(xx,yy,-(xy*xy)) etc. */
            double det = fma(xx, yy, -(xy * xy));
            if (det == 0.0) continue;
            double a0 = fma(xz, yy, -(xy * yz)) / det;
            double a1 = fma(xx, yz, -(xy * xz)) / det;
            double m = sqrt(fma(a0, a0, a1 * a1) + 1.0);
            a0 = a0 / m; a1 = a1 / m;
            double err = 0.0;
            for (int i = 0; i < n; ++i) {
                double e = (pz[i] - mz) - fma(a0, px[i] - mx, a1 * (py[i] - my));
                err = fma(e, e, err);
            }
            err = err / (double)n;
            if (err > 0) err = log(err);
            HM(rough, x0, y0) = err;
            HM(xs, x0, y0) = atan2(a0, 1.0 / m);
            HM(ys, x0, y0) = atan2(a1, 1.0 / m);
        }
}

/* ------------------------------------------------------------------------
 * gvom.py:592-713  __guess_height (output pre-filled 0, :355-356).  Quirks kept:
 * the loop condition tests x_n_done twice (:619); y_nh is folded in under the
 * x_nh guard (:704-706); asymmetric lateral ranges (:631,646,661,676).
 * ---------------------------------------------------------------------- */
void gvo_guess_height(const gvo_params* g, const double* height, const double* inferred, double* guessed) {
    const int64_t S = g->xy_size;
    for (int64_t x0 = 0; x0 < S; ++x0)
        for (int64_t y0 = 0; y0 < S; ++y0) {
            HM(guessed, x0, y0) = 0.0;
            if (HM(height, x0, y0) > -1000.0) continue;
            if (HM(inferred, x0, y0) == -1000.0) continue;
            int xpd = 0, xnd = 0, ypd = 0, ynd = 0;
            int64_t x_p = x0, x_n = x0, y_p = y0, y_n = y0;
            double x_ph = -1000, x_nh = -1000, y_ph = -1000, y_nh = -1000;
            int64_t i = 0;
            while (i < 15 && !(xnd && xnd && ypd && ynd)) {
                x_p += 1; x_n -= 1; y_p += 1; y_n -= 1; i += 1;
                if (!xpd) {
                    if (x_p < S) {
                        for (int64_t d = -i; d < i; ++d) {
                            if (y0 + d >= S || y0 + d < 0) continue;
                            if (HM(height, x_p, y0 + d) > -1000.0) { x_ph = HM(height, x_p, y0 + d); xpd = 1; break; }
                        }
                    } else xpd = 1;
                }
                if (!xnd) {
                    if (x_n >= 0) {
                        for (int64_t d = -i + 1; d < i + 1; ++d) {
                            if (y0 + d >= S || y0 + d < 0) continue;
                            if (HM(height, x_n, y0 + d) > -1000.0) { x_nh = HM(height, x_n, y0 + d); xnd = 1; break; }
                        }
                    } else xnd = 1;
                }
                if (!ypd) {
                    if (y_p < S) {
                        for (int64_t d = -i + 1; d < i + 1; ++d) {
                            if (x0 + d >= S || x0 + d < 0) continue;
                            if (HM(height, x0 + d, y_p) > -1000.0) { y_ph = HM(height, x0 + d, y_p); ypd = 1; break; }
                        }
                    } else ypd = 1;
                }
                if (!ynd) {
                    if (y_n >= 0) {
                        for (int64_t d = -i; d < i; ++d) {
                            if (x0 + d >= S || x0 + d < 0) continue;
                            if (HM(height, x0 + d, y_n) > -1000.0) { y_nh = HM(height, x0 + d, y_n); ynd = 1; break; }
                        }
                    } else ynd = 1;
                }
            }
            double mn = 1000.0, mx = HM(inferred, x0, y0);
            if (x_ph > -1000) { mn = fmin(x_ph, mn); mx = fmax(x_ph, mx); }
            if (x_nh > -1000) { mn = fmin(x_nh, mn); mx = fmax(x_nh, mx); }
            if (y_ph > -1000) { mn = fmin(y_ph, mn); mx = fmax(y_ph, mx); }
            if (x_nh > -1000) { mn = fmin(y_nh, mn); mx = fmax(y_nh, mx); }   /* sic */
            double dh = mx - mn;
            if (dh > 0) HM(guessed, x0, y0) = dh;
        }
}

/* ------------------------------------------------------------------------
 * gvom.py:515-555 positive, :505-512 negative, :444-452 visibility (int32 maps)
 * ---------------------------------------------------------------------- */
void gvo_obstacle_maps(const gvo_params* g, const double* origin, const int32_t* combined,
                       const int32_t* chit, const int32_t* ctot, const double* height,
                       const double* xs, const double* ys, const double* guessed,
                       int32_t* pos, int32_t* neg, int32_t* vis) {
    const int64_t S = g->xy_size, Z = g->z_size;
    for (int64_t x = 0; x < S; ++x)
        for (int64_t y = 0; y < S; ++y) {
            HM(neg, x, y) = HM(guessed, x, y) > g->neg_thr ? 100 : 0;
            HM(vis, x, y) = HM(height, x, y) > -1000.0 ? 1 : 0;
            HM(pos, x, y) = 0;
            double sx = HM(xs, x, y), sy = HM(ys, x, y);
            double sl = sqrt(fma(sx, sx, sy * sy));
            if (!(sl < g->slope_thr)) { HM(pos, x, y) = 100; continue; }
            double h = HM(height, x, y);
            double lo = floor((h + g->pos_thr) / g->z_res - origin[2]);
            double hi = floor((h + g->robot_height) / g->z_res - origin[2]);
            int64_t zlo = (int64_t)lo + 1, zhi = (int64_t)hi;
            if (!(zlo >= 0 && zlo < Z)) continue;
            if (!(zhi >= 0 && zhi < Z)) continue;
            double density = 0.0, n = 0.0;
            for (int64_t z = zlo; z <= zhi; ++z) {
                int32_t idx = combined[x + y * S + z * S * S];
                if (idx >= 0 && chit[idx] > 10) { n += (double)ctot[idx]; density += (double)chit[idx]; }
            }
            if (n > 0.0) density = density / n;
            HM(pos, x, y) = (int32_t)(density * 100.0);
        }
}

/* ------------------------------------------------------------------------
 * gvom.py:481-503, :455-468, :470-479  debug exports (float32 rows)
 * ---------------------------------------------------------------------- */
void gvo_debug_voxel_map(const gvo_params* g, const double* origin, const int32_t* combined,
                         const int32_t* chit, const int32_t* ctot, const float* eig, float* out) {
    const int64_t S = g->xy_size, Z = g->z_size;
    for (int64_t z = 0; z < Z; ++z)
        for (int64_t y = 0; y < S; ++y)
            for (int64_t x = 0; x < S; ++x) {
                int32_t i = combined[x + y * S + z * S * S];
                if (i < 0) continue;
                float* o = out + (int64_t)i * 8;
                o[0] = (float)(((double)x + origin[0]) * g->xy_res);
                o[1] = (float)(((double)y + origin[1]) * g->xy_res);
                o[2] = (float)(((double)z + origin[2]) * g->z_res);
                o[3] = (float)((double)chit[i] / (double)ctot[i]);
                o[4] = (float)chit[i];
                o[5] = eig[i * 3 + 0] - eig[i * 3 + 1];
                o[6] = eig[i * 3 + 1] - eig[i * 3 + 2];
                o[7] = eig[i * 3 + 2];
            }
}

void gvo_debug_height_map(const gvo_params* g, const double* origin, const double* height,
                          const double* rough, const double* xs, const double* ys, float* out) {
    const int64_t S = g->xy_size;
    for (int64_t x = 0; x < S; ++x)
        for (int64_t y = 0; y < S; ++y) {
            float* o = out + (x + y * S) * 7;
            double sx = HM(xs, x, y), sy = HM(ys, x, y);
            o[0] = (float)(((double)x + origin[0]) * g->xy_res);
            o[1] = (float)(((double)y + origin[1]) * g->xy_res);
            o[2] = (float)(HM(height, x, y) - g->z_res);
            o[3] = (float)HM(rough, x, y);
            o[4] = (float)sx;
            o[5] = (float)sy;
            o[6] = (float)sqrt(fma(sx, sx, sy * sy));
        }
}

void gvo_debug_inferred_height_map(const gvo_params* g, const double* origin, const double* guessed, float* out) {
    const int64_t S = g->xy_size;
    for (int64_t x = 0; x < S; ++x)
        for (int64_t y = 0; y < S; ++y) {
            float* o = out + (x + y * S) * 3;
            o[0] = (float)(((double)x + origin[0]) * g->xy_res);
            o[1] = (float)(((double)y + origin[1]) * g->xy_res);
            o[2] = (float)(HM(guessed, x, y) - g->z_res);
        }
}

int gvo_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void gvo_set_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}
