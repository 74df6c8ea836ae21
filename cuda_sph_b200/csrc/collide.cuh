// Wall response in fp64 (inputs are the fp32 state, promoted): box walls (reference base_kernels.py:75-98) and the
// pipe-segment solver (base_kernels.py:56-72,101-127; util_kernels.py:38-204) incl. the xoroshiro128+ outlet recycle
// (numba/cuda/random.py, pinned numba==0.54.1).  fp64 because the pipe solver divides by sin(alpha)+0.001 with
// sin = sqrt(1 - cos^2): in fp32 the cancellation would eat the 1e-4 position tolerance for grazing impacts.
#pragma once
#include "sph_common.cuh"

namespace sph {

// ---- box: per dimension, strict comparisons, clamp to 1e-3 inside, v *= -1 then *= DAMP --------------------------
__device__ __forceinline__ void collide_box(double *x, double *v, const StepConsts &c) {
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        bool bounced = false;
        if (x[d] < 0.0) { x[d] = 1e-3; bounced = true; }
        if (x[d] > c.space[d]) { x[d] = c.space[d] - 1e-3; bounced = true; }
        if (bounced) { v[d] *= -1.0; v[d] *= c.damp; }
    }
}

// ---- pipe table helpers: rows of [x_start, y_c, z_c, r_start, length], last row = pipe end ------------------------
struct PipeView {
    const double *t;
    int rows;
    __device__ __forceinline__ double at(int s, int col) const { return t[5 * s + col]; }
};

__device__ inline int pipe_find_segment(const PipeView &p, double x) {
    for (int j = 0; j + 1 < p.rows; ++j)
        if (p.at(j, 0) <= x && x < p.at(j + 1, 0)) return j;
    return -1;
}

__device__ inline double pipe_x_begin(const PipeView &p, int s) {
    double xb = p.at(0, 0);
    for (int i = 0; i < s; ++i) xb += p.at(i, 4);
    return xb;
}

__device__ inline double pipe_radius(const PipeView &p, int s, double x) {
    const double r0 = p.at(s, 3), r1 = p.at(s + 1, 3);
    if (r0 == r1) return r0;
    if (r0 < r1) {
        const double delta = x - pipe_x_begin(p, s);
        const double trunc_len = p.at(s, 4) * r0 / (r1 - r0);
        return r0 * (1.0 + delta / trunc_len);
    }
    const double delta = pipe_x_begin(p, s) + p.at(s, 4) - x;
    const double trunc_len = p.at(s, 4) * r1 / (r0 - r1);
    return r1 * (1.0 + delta / trunc_len);
}

__device__ inline bool pipe_is_out(const PipeView &p, int s, const double *x) {
    const double yn = x[1] - p.at(s, 1), zn = x[2] - p.at(s, 2);
    return sqrt(yn * yn + zn * zn) > pipe_radius(p, s, x[0]);
}

__device__ __forceinline__ double len3(const double *a) { return sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); }

// Single bounce about the wall generatrix lying in the particle's azimuthal plane (util_kernels.py:152-204).
__device__ inline void pipe_bounce(const PipeView &p, int s, double *x, double *v) {
    const double r0 = p.at(s, 3), r1 = p.at(s + 1, 3);
    const double yn = x[1] - p.at(s, 1), zn = x[2] - p.at(s, 2);
    const double hh = sqrt(yn * yn + zn * zn);
    double a[3], b[3], e[3];  // generatrix end points and direction
    a[0] = p.at(s, 0);
    b[0] = p.at(s + 1, 0);
    a[1] = yn / hh * r0;  b[1] = yn / hh * r1;
    a[2] = zn / hh * r0;  b[2] = zn / hh * r1;
    e[0] = b[0] - a[0];  e[1] = b[1] - a[1];  e[2] = b[2] - a[2];
    a[1] += p.at(s, 1);
    a[2] += p.at(s, 2);

    // back-track along v to the wall: t = -(d / (sin(alpha) + 0.001)) / |v|   (util_kernels.py:113-121)
    const double vl = len3(v), el = len3(e);
    const double cos_a = (v[0] * e[0] + v[1] * e[1] + v[2] * e[2]) / (vl * el);
    const double w[3] = {a[0] - x[0], a[1] - x[1], a[2] - x[2]};
    const double cr[3] = {e[1] * w[2] - e[2] * w[1], e[2] * w[0] - e[0] * w[2], e[0] * w[1] - e[1] * w[0]};
    const double dist = len3(cr) / el;
    const double sin_a = sqrt(1.0 - cos_a * cos_a);
    const double tback = -(dist / (sin_a + 0.001)) / vl;
    const double cp[3] = {x[0] + v[0] * tback, x[1] + v[1] * tback, x[2] + v[2] * tback};

    if (r0 == r1) {  // cylinder: flip the radial components
        v[1] = -v[1];
        v[2] = -v[2];
    } else {         // cone: v' = 2 proj_e(v) - v
        const double c2 = (e[0] * v[0] + e[1] * v[1] + e[2] * v[2]) / (el * vl);
#pragma unroll
        for (int d = 0; d < 3; ++d) v[d] = 2.0 * e[d] / el * c2 * vl - v[d];
    }
    const double dd[3] = {x[0] - cp[0], x[1] - cp[1], x[2] - cp[2]};
    const double way = len3(dd), nv = len3(v);
#pragma unroll
    for (int d = 0; d < 3; ++d) x[d] = cp[d] + v[d] * way / nv;
}

// ---- xoroshiro128+ (2016 constants 55/14/36, as in numba) -----------------------------------------------------------
__host__ __device__ __forceinline__ uint64_t xoro_rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
__host__ __device__ __forceinline__ uint64_t xoro_next(uint64_t &s0, uint64_t &s1) {
    const uint64_t r = s0 + s1;
    s1 ^= s0;
    s0 = xoro_rotl(s0, 55) ^ s1 ^ (s1 << 14);
    s1 = xoro_rotl(s1, 36);
    return r;
}
__host__ __device__ __forceinline__ double xoro_unit(uint64_t &s0, uint64_t &s1) {
    return (double)(xoro_next(s0, s1) >> 11) * (1.0 / 9007199254740992.0);
}

// collision_kernel (base_kernels.py:56-72).  rng = this particle's two state words (indexed by particle id).
__device__ inline void collide_pipe(const PipeView &p, double *x, double *v, uint64_t *rng) {
    const int s = pipe_find_segment(p, x[0]);
    if (s >= 0) {
        if (pipe_is_out(p, s, x)) pipe_bounce(p, s, x, v);
        return;
    }
    if (x[0] < 0.0) {  // left of the inlet: mirror, then bounce against segment 0 if outside
        x[0] = -x[0];
        v[0] = -v[0];
        if (pipe_is_out(p, 0, x)) pipe_bounce(p, 0, x, v);
    } else {           // past the outlet (or NaN): recycle to x = 0 with a uniform point of the inlet disc
        x[0] = 0.0;
        const int s0 = pipe_find_segment(p, 0.0);
        const double R = pipe_radius(p, s0 < 0 ? 0 : s0, 0.0);
        uint64_t a = rng[0], b = rng[1];
        const double r = R * sqrt(xoro_unit(a, b));
        const double theta = xoro_unit(a, b) * 2.0 * 3.141592653589793;
        rng[0] = a;
        rng[1] = b;
        x[1] = p.at(0, 1) + r * cos(theta);
        x[2] = p.at(0, 2) + r * sin(theta);
    }
}

}  // namespace sph
