/*
 * TEST INFRASTRUCTURE -- NOT PRODUCT CODE.
 *
 * fp64 CPU restatement of the reference's Voxel SPH step (iwoplaza/cuda-sph), used only as the parity
 * checker by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.  The
 * product path (cuda_sph_b200/csrc) never links, imports or calls anything in this file.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks this file against golden vectors produced by
 * executing the reference's own kernels under numba's CUDA simulator (tests/golden/generate_golden.py, via
 * oracle/ref_shim.py), against SURVEY.md Appendix A, and against the known-answer values of the reference's
 * own sim/tests/test_collisions.py:90-228.
 *
 * Each function cites the reference lines it restates.  Arithmetic is written in the reference's operand
 * order.  Where the reference's Python source uses `**` the simulator evaluates libm pow(); with
 * orc_set_exact_pow(1) this file does the same so that results are bit-identical to the simulator run; with
 * orc_set_exact_pow(0) squares/cubes are plain multiplications (what numba's NVVM backend would emit on a
 * real GPU; differs from pow() by <= 1 ulp in ~0.1 % of squares) -- that is the mode timed as the CPU
 * baseline.
 *
 * Deviations from reference undefined behaviour (documented in DESIGN.md, mirrored by the CUDA engine):
 *   D1  a particle whose position is non-finite, whose cell coordinate does not fit int32, or whose linear
 *       cell key falls outside [0, n_cells) gets key = n_cells ("dead cell"): it sorts to the tail, is nobody's
 *       neighbour candidate and has zero neighbours itself (reference: out-of-bounds index, voxel_kernels.py:12,17).
 *   D2  missing neighbour cells are skipped (reference leaves neigh_voxels uninitialised, voxel_kernels.py:44,64;
 *       the `i == -1` test at :65 shows the intent).
 *   D3  the end of a cell's map range is the begin of the next non-empty cell or (N - n_dead)
 *       (reference: len(map), voxel_kernels.py:72-76; identical when there are no dead particles).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_MAX_NEIGHBOURS 32 /* config.py:30 */

typedef struct {
    int32_t n;            /* particle count */
    int32_t mode;         /* 0 = BOX, 1 = PIPE  (config.py:13, abstract_sph_strategy.py:102) */
    double h;             /* INF_R   config.py:20 */
    double mass;          /* MASS    config.py:18 */
    double rho0;          /* RHO_0   config.py:19 */
    double k;             /* K       config.py:22 */
    double visc;          /* VISC    config.py:21 */
    double damp;          /* DAMP    config.py:23 */
    double dt;            /* 1/fps   abstract_sph_strategy.py:20 */
    double ext[3];        /* external_force */
    double space[3];      /* space_size */
    double voxel[3];      /* voxel_size */
    int32_t pipe_rows;    /* S+1 rows of [x, y, z, r, len]  (common/data_classes.py:34-44) */
    int32_t pad_;
    const double *pipe;
} OrcParams;

static int g_exact_pow = 1;
void orc_set_exact_pow(int on) { g_exact_pow = on; }
int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void orc_set_num_threads(int t) {
#ifdef _OPENMP
    if (t > 0) omp_set_num_threads(t);
#else
    (void)t;
#endif
}

static inline double p2(double x) { return g_exact_pow ? pow(x, 2.0) : x * x; }
static inline double p3(double x) { return g_exact_pow ? pow(x, 3.0) : x * x * x; }
static inline double phalf(double x) { return g_exact_pow ? pow(x, 0.5) : sqrt(x); }

/* ---- SPH constants: config.py:24-29 ---- */
static double w_const(double h) { return 315.0 / (64.0 * M_PI * pow(h, 9.0)); }
static double grad_w_const(double h) { return -45.0 / (M_PI * pow(h, 6.0)); }
static double lap_w_const(double h) { return 45.0 / (M_PI * pow(h, 6.0)); }

void orc_constants(double h, double *out3) {
    out3[0] = w_const(h);
    out3[1] = grad_w_const(h);
    out3[2] = lap_w_const(h);
}

/* ceil dims for keys (voxel_sph_strategy.py:70-73), trunc dims for neighbour bounds (:110-116) */
void orc_dims(const OrcParams *P, int32_t *ceil3, int32_t *trunc3) {
    for (int d = 0; d < 3; ++d) {
        double q = P->space[d] / P->voxel[d];
        ceil3[d] = (int32_t)ceil(q);
        trunc3[d] = (int32_t)q;
    }
}

static int64_t n_cells_of(const OrcParams *P) {
    int32_t c[3], t[3];
    orc_dims(P, c, t);
    return (int64_t)c[0] * c[1] * c[2];
}
int64_t orc_n_cells(const OrcParams *P) { return n_cells_of(P); }

/* voxel_kernels.py:9-12  v_d = int32(pos_d / voxel_size_d), C truncation.  Returns 0 if not representable. */
static int cell_coords(const OrcParams *P, const double *pos, int32_t *v) {
    for (int d = 0; d < 3; ++d) {
        double q = pos[d] / P->voxel[d];
        if (!(fabs(q) < 2147483648.0)) return 0; /* NaN, inf, overflow -> D1 */
        v[d] = (int32_t)q;
    }
    return 1;
}

/* voxel_kernels.py:15-17, 88-105 (+ D1) */
void orc_cell_keys(const OrcParams *P, const double *pos, int32_t *keys) {
    int32_t cd[3], td[3];
    orc_dims(P, cd, td);
    const int64_t ncell = (int64_t)cd[0] * cd[1] * cd[2];
#pragma omp parallel for schedule(static)
    for (int32_t i = 0; i < P->n; ++i) {
        int32_t v[3];
        int64_t key = ncell;
        if (cell_coords(P, pos + 3 * (size_t)i, v)) {
            int64_t k = (int64_t)v[0] + (int64_t)v[1] * cd[0] + (int64_t)v[2] * cd[0] * cd[1];
            if (k >= 0 && k < ncell) key = k;
        }
        keys[i] = (int32_t)key;
    }
}

/* voxel_sph_strategy.py:81-107: map sorted by (voxel_id, particle_id); voxel_begin[c] = first map index of
 * cell c or -1.  A stable counting sort by key is exactly numpy's structured sort order (== lexsort((id,key))).
 * voxel_begin has n_cells entries; returns the number of dead particles (key == n_cells). */
int32_t orc_sort_cells(const OrcParams *P, const int32_t *keys, int32_t *map_ids, int32_t *map_keys,
                       int32_t *voxel_begin) {
    const int64_t ncell = n_cells_of(P);
    const int32_t n = P->n;
    int32_t *start = (int32_t *)calloc((size_t)ncell + 2, sizeof(int32_t));
    for (int32_t i = 0; i < n; ++i) start[keys[i] + 1]++;
    for (int64_t c = 0; c <= ncell; ++c) start[c + 1] += start[c];
    const int32_t n_dead = n - start[ncell];
    if (voxel_begin)
        for (int64_t c = 0; c < ncell; ++c) voxel_begin[c] = (start[c + 1] > start[c]) ? start[c] : -1;
    for (int32_t i = 0; i < n; ++i) {
        int32_t at = start[keys[i]]++;
        map_ids[at] = i;
        if (map_keys) map_keys[at] = keys[i];
    }
    free(start);
    return n_dead;
}

/* Exclusive cell ranges [begin, end) used by the neighbour walk: begin = voxel_begin, end per D3.
 * ends[c] is only meaningful for non-empty cells. */
static void cell_ends(int64_t ncell, const int32_t *voxel_begin, int32_t n_live, int32_t *ends) {
    int32_t next = n_live;
    for (int64_t c = ncell - 1; c >= 0; --c) {
        ends[c] = next;
        if (voxel_begin[c] != -1) next = voxel_begin[c];
    }
}

/* voxel_kernels.py:20-26 */
static inline int are_neighbours(const double *a, const double *b, double h) {
    return sqrt(p2(a[0] - b[0]) + p2(a[1] - b[1]) + p2(a[2] - b[2])) <= h;
}

/* voxel_kernels.py:29-85 (+ D1, D2, D3).  Writes up to 32 particle ids, returns the count. */
static int get_neighbours(const OrcParams *P, const int32_t *td, int32_t i, const double *pos,
                          const int32_t *voxel_begin, const int32_t *ends, const int32_t *map_ids,
                          int64_t ncell, int32_t *out) {
    int32_t v[3];
    const double *pi = pos + 3 * (size_t)i;
    if (!cell_coords(P, pi, v)) return 0;
    {   /* D1: a dead particle has no neighbours */
        int32_t cd[3], t2[3];
        orc_dims(P, cd, t2);
        int64_t k = (int64_t)v[0] + (int64_t)v[1] * cd[0] + (int64_t)v[2] * cd[0] * cd[1];
        if (k < 0 || k >= ncell) return 0;
    }
    int cnt = 0;
    for (int dx = -1; dx <= 1; ++dx)
        for (int dy = -1; dy <= 1; ++dy)
            for (int dz = -1; dz <= 1; ++dz) {
                int64_t x = (int64_t)v[0] + dx, y = (int64_t)v[1] + dy, z = (int64_t)v[2] + dz;
                if (x < 0 || x >= td[0] || y < 0 || y >= td[1] || z < 0 || z >= td[2]) continue;
                /* compute_1d_idx(neigh_voxel, space_dim) with the TRUNC dims: voxel_kernels.py:60 */
                int64_t c = x + y * td[0] + z * (int64_t)td[0] * td[1];
                if (c >= ncell || voxel_begin[c] == -1) continue;
                for (int32_t m = voxel_begin[c]; m < ends[c]; ++m) {
                    int32_t j = map_ids[m];
                    if (are_neighbours(pi, pos + 3 * (size_t)j, P->h)) {
                        out[cnt++] = j;
                        if (cnt >= ORC_MAX_NEIGHBOURS) return cnt;
                    }
                }
            }
    return cnt;
}

/* base_kernels.py:6-9 + util_kernels.py:21-28 */
static inline double norm_squared(const double *a, const double *b) {
    double res = 0.0;
    for (int d = 0; d < 3; ++d) res += p2(a[d] - b[d]);
    return res;
}

/* ---------------- pipe geometry: util_kernels.py:38-204 ---------------- */
#define PR(s, c) pipe[5 * (s) + (c)]

/* util_kernels.py:72-77 */
int orc_find_segment(const double *pipe, int rows, double x) {
    for (int j = 0; j < rows - 1; ++j)
        if (PR(j, 0) <= x && x < PR(j + 1, 0)) return j;
    return -1;
}
/* util_kernels.py:38-43 */
double orc_x_at_segment_beginning(const double *pipe, int s) {
    double xb = PR(0, 0);
    for (int i = 0; i < s; ++i) xb += PR(i, 4);
    return xb;
}
/* util_kernels.py:97-102 */
double orc_vector_length(const double *v) {
    double len = 0;
    for (int i = 0; i < 3; ++i) len += p2(v[i]);
    return sqrt(len);
}
/* util_kernels.py:105-110 */
double orc_distance_between_points(const double *a, const double *b) {
    double len = 0;
    for (int i = 0; i < 3; ++i) len += p2(a[i] - b[i]);
    return sqrt(len);
}
/* util_kernels.py:56-69 */
static double radius_in_position(const double *pos, const double *pipe, int s) {
    const double r0 = PR(s, 3), r1 = PR(s + 1, 3);
    if (r0 == r1) return r0;
    if (r0 < r1) {
        double delta = pos[0] - orc_x_at_segment_beginning(pipe, s);
        double truncated = PR(s, 4) * r0 / (r1 - r0);
        return r0 * (1.0 + delta / truncated);
    }
    double delta = orc_x_at_segment_beginning(pipe, s) + PR(s, 4) - pos[0];
    double truncated = PR(s, 4) * r1 / (r0 - r1);
    return r1 * (1.0 + delta / truncated);
}
/* util_kernels.py:46-53 */
int orc_is_out_of_pipe(const double *pos, const double *pipe, int s) {
    double yn = pos[1] - PR(s, 1), zn = pos[2] - PR(s, 2);
    double hh = phalf(p2(yn) + p2(zn));
    return hh > radius_in_position(pos, pipe, s);
}
/* util_kernels.py:80-85 */
static double cos_between(const double *a, const double *b, double la, double lb) {
    double scalar = 0;
    for (int d = 0; d < 3; ++d) scalar = scalar + a[d] * b[d];
    return scalar / (la * lb);
}
/* util_kernels.py:124-142 */
static double distance_to_pipe(const double *point, const double *edge, const double *lpoint) {
    double pv[3], cr[3];
    for (int d = 0; d < 3; ++d) pv[d] = lpoint[d] - point[d];
    cr[0] = edge[1] * pv[2] - edge[2] * pv[1];
    cr[1] = edge[2] * pv[0] - edge[0] * pv[2];
    cr[2] = edge[0] * pv[1] - edge[1] * pv[0];
    return orc_vector_length(cr) / orc_vector_length(edge);
}
/* util_kernels.py:113-121 */
static double calc_dt(const double *pos, const double *speed, const double *edge, const double *lpoint) {
    double sl = orc_vector_length(speed), el = orc_vector_length(edge);
    double cos_a = cos_between(speed, edge, sl, el);
    double d = distance_to_pipe(pos, edge, lpoint);
    double sin_a = sqrt(1 - p2(cos_a));
    double d_to_collision = d / (sin_a + 0.001);
    return -d_to_collision / sl;
}
/* util_kernels.py:182-204 (calc_edge_vector :152-171, calc_collision_point :174-179, calc_summary_vector :88-94) */
void orc_solve_collision(double *pos, double *speed, const double *pipe, int s) {
    const double r0 = PR(s, 3), r1 = PR(s + 1, 3);
    double yn = pos[1] - PR(s, 1), zn = pos[2] - PR(s, 2);
    double hh = phalf(p2(yn) + p2(zn));
    double first[3], second[3], edge[3], cp[3];
    first[0] = PR(s, 0);
    second[0] = PR(s + 1, 0);
    for (int d = 1; d < 3; ++d) {
        first[d] = (pos[d] - PR(s, d)) / hh * r0;
        second[d] = (pos[d] - PR(s, d)) / hh * r1;
    }
    for (int d = 0; d < 3; ++d) edge[d] = second[d] - first[d];
    for (int d = 1; d < 3; ++d) first[d] = first[d] + PR(s, d);

    double dt = calc_dt(pos, speed, edge, first);
    for (int d = 0; d < 3; ++d) cp[d] = pos[d] + speed[d] * dt;

    if (r0 == r1) {
        for (int d = 1; d < 3; ++d) speed[d] = -speed[d];
    } else {
        double ll = orc_vector_length(edge), pl = orc_vector_length(speed);
        double c = cos_between(edge, speed, ll, pl);
        double sum[3];
        for (int d = 0; d < 3; ++d) sum[d] = 2.0 * edge[d] / ll * c * pl;
        for (int d = 0; d < 3; ++d) speed[d] = sum[d] - speed[d];
    }
    double way = orc_distance_between_points(pos, cp);
    double sv = orc_vector_length(speed);
    for (int d = 0; d < 3; ++d) pos[d] = cp[d] + speed[d] * way / sv;
}

/* ---------------- xoroshiro128+ as in numba/cuda/random.py (pinned numba==0.54.1, sim/requirements.txt:2) -------- */
static inline uint64_t rotl64(uint64_t x, unsigned k) { return (x << k) | (x >> (64 - k)); }
static inline uint64_t xoro_next(uint64_t *s) {
    uint64_t s0 = s[0], s1 = s[1], result = s0 + s1;
    s1 ^= s0;
    s[0] = rotl64(s0, 55) ^ s1 ^ (s1 << 14);
    s[1] = rotl64(s1, 36);
    return result;
}
static void xoro_jump(uint64_t *s) {
    static const uint64_t jump[2] = {0xbeac0467eba5facbULL, 0xd86b048b86aa9922ULL};
    uint64_t s0 = 0, s1 = 0;
    for (int i = 0; i < 2; ++i)
        for (int b = 0; b < 64; ++b) {
            if (jump[i] & (1ULL << b)) { s0 ^= s[0]; s1 ^= s[1]; }
            xoro_next(s);
        }
    s[0] = s0;
    s[1] = s1;
}
/* create_xoroshiro128p_states(n, seed): state 0 = SplitMix64(seed) in both words, state i = jump(state i-1).
 * abstract_sph_strategy.py:27 uses n = grid*block, seed = 16435234. */
void orc_rng_init(uint64_t *states, int64_t n, uint64_t seed) {
    if (n < 1) return;
    uint64_t z = seed + 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    z = z ^ (z >> 31);
    states[0] = z;
    states[1] = z;
    for (int64_t i = 1; i < n; ++i) {
        states[2 * i] = states[2 * i - 2];
        states[2 * i + 1] = states[2 * i - 1];
        xoro_jump(states + 2 * i);
    }
}
static inline double xoro_uniform(uint64_t *s) {
    return (double)(xoro_next(s) >> 11) * (1.0 / 9007199254740992.0);
}
double orc_rng_uniform(uint64_t *state2) { return xoro_uniform(state2); }

/* base_kernels.py:101-127 */
static void put_particle_at_pipe_begin(double *pos, double *vel, const double *pipe, int rows, uint64_t *rng) {
    if (pos[0] < 0) {
        pos[0] = -pos[0];
        vel[0] = -vel[0];
        if (orc_is_out_of_pipe(pos, pipe, 0)) orc_solve_collision(pos, vel, pipe, 0);
    } else {
        pos[0] = 0.0;
        int seg = orc_find_segment(pipe, rows, pos[0]);
        double R = radius_in_position(pos, pipe, seg);
        double r = R * sqrt(xoro_uniform(rng));
        double theta = xoro_uniform(rng) * 2.0 * M_PI;
        pos[1] = PR(0, 1) + r * cos(theta);
        pos[2] = PR(0, 2) + r * sin(theta);
    }
}

/* base_kernels.py:56-72.  In place on pos/vel (N x 3); rng = N x 2 uint64. */
void orc_collide_pipe(const OrcParams *P, double *pos, double *vel, uint64_t *rng) {
    const double *pipe = P->pipe;
#pragma omp parallel for schedule(static)
    for (int32_t i = 0; i < P->n; ++i) {
        double *p = pos + 3 * (size_t)i, *v = vel + 3 * (size_t)i;
        int s = orc_find_segment(pipe, P->pipe_rows, p[0]);
        if (s == -1)
            put_particle_at_pipe_begin(p, v, pipe, P->pipe_rows, rng + 2 * (size_t)i);
        else if (orc_is_out_of_pipe(p, pipe, s))
            orc_solve_collision(p, v, pipe, s);
    }
}

/* base_kernels.py:75-98 */
void orc_collide_box(const OrcParams *P, double *pos, double *vel) {
#pragma omp parallel for schedule(static)
    for (int32_t i = 0; i < P->n; ++i)
        for (int d = 0; d < 3; ++d) {
            double *x = pos + 3 * (size_t)i + d, *v = vel + 3 * (size_t)i + d;
            int bounced = 0;
            if (*x < 0) { *x = 1e-3; bounced = 1; }
            if (*x > P->space[d]) { *x = P->space[d] - 1e-3; bounced = 1; }
            if (bounced) { *v *= -1; *v *= P->damp; }
        }
}

/* base_kernels.py:30-53 */
void orc_integrate(const OrcParams *P, double *pos, double *vel, const double *rho, const double *pressure,
                   const double *viscosity, double *force) {
#pragma omp parallel for schedule(static)
    for (int32_t i = 0; i < P->n; ++i)
        for (int d = 0; d < 3; ++d) {
            size_t a = 3 * (size_t)i + d;
            force[a] = P->ext[d] + -pressure[a] + viscosity[a];
            vel[a] += force[a] / rho[i] * P->dt;
            pos[a] += vel[a] * P->dt;
        }
}

/* Neighbour lists for every particle (voxel_kernels.py:29-85).  neigh = N x 32 int32, cnt = N int32. */
void orc_neighbours(const OrcParams *P, const double *pos, const int32_t *map_ids, const int32_t *voxel_begin,
                    int32_t n_dead, int32_t *neigh, int32_t *cnt) {
    int32_t cd[3], td[3];
    orc_dims(P, cd, td);
    const int64_t ncell = (int64_t)cd[0] * cd[1] * cd[2];
    int32_t *ends = (int32_t *)malloc((size_t)(ncell > 0 ? ncell : 1) * sizeof(int32_t));
    cell_ends(ncell, voxel_begin, P->n - n_dead, ends);
#pragma omp parallel for schedule(dynamic, 256)
    for (int32_t i = 0; i < P->n; ++i)
        cnt[i] = get_neighbours(P, td, i, pos, voxel_begin, ends, map_ids, ncell,
                                neigh + (size_t)ORC_MAX_NEIGHBOURS * i);
    free(ends);
}

/* voxel_kernels.py:108-132 */
void orc_density(const OrcParams *P, const double *pos, const int32_t *neigh, const int32_t *cnt, double *rho) {
    const double wc = w_const(P->h), h2 = p2(P->h); /* INF_R_2 = INF_R ** 2, config.py:24 */
#pragma omp parallel for schedule(static)
    for (int32_t i = 0; i < P->n; ++i) {
        double acc = 0;
        const int32_t *nb = neigh + (size_t)ORC_MAX_NEIGHBOURS * i;
        for (int q = 0; q < cnt[i]; ++q) {
            int32_t j = nb[q];
            if (j == i) continue;
            acc += wc * p3(h2 - norm_squared(pos + 3 * (size_t)i, pos + 3 * (size_t)j));
        }
        rho[i] = acc * P->mass;
    }
}

/* voxel_kernels.py:135-171 + base_kernels.py:12-21 */
void orc_pressure(const OrcParams *P, const double *pos, const double *rho, const int32_t *neigh,
                  const int32_t *cnt, double *out) {
    const double gc = grad_w_const(P->h);
#pragma omp parallel for schedule(static)
    for (int32_t i = 0; i < P->n; ++i) {
        double acc[3] = {0.0, 0.0, 0.0};
        const double *pi = pos + 3 * (size_t)i;
        const int32_t *nb = neigh + (size_t)ORC_MAX_NEIGHBOURS * i;
        for (int q = 0; q < cnt[i]; ++q) {
            int32_t j = nb[q];
            if (j == i) continue;
            const double *pj = pos + 3 * (size_t)j;
            double p_i = P->k * (rho[i] - P->rho0);
            double p_j = P->k * (rho[j] - P->rho0);
            double factor = p_i / p2(rho[i]) + p_j / p2(rho[j]);
            double dist = sqrt(norm_squared(pi, pj));
            double gf = gc * p2(P->h - dist);
            for (int d = 0; d < 3; ++d) acc[d] += factor * (gf * (pi[d] - pj[d]) / dist);
        }
        for (int d = 0; d < 3; ++d) out[3 * (size_t)i + d] = acc[d];
    }
}

/* voxel_kernels.py:174-211 + base_kernels.py:24-27.  min(1, x) follows Python: x if x < 1 else 1 (so NaN -> 1). */
void orc_viscosity(const OrcParams *P, const double *pos, const double *vel, const double *rho,
                   const int32_t *neigh, const int32_t *cnt, double *out) {
    const double lc = lap_w_const(P->h);
#pragma omp parallel for schedule(static)
    for (int32_t i = 0; i < P->n; ++i) {
        double acc[3] = {0.0, 0.0, 0.0};
        const double *pi = pos + 3 * (size_t)i, *vi = vel + 3 * (size_t)i;
        const int32_t *nb = neigh + (size_t)ORC_MAX_NEIGHBOURS * i;
        for (int q = 0; q < cnt[i]; ++q) {
            int32_t j = nb[q];
            if (j == i) continue;
            double lap = lc * (P->h - sqrt(norm_squared(pi, pos + 3 * (size_t)j)));
            for (int d = 0; d < 3; ++d) {
                double term = (vel[3 * (size_t)j + d] - vi[d]) / rho[j] * lap;
                acc[d] += term * P->mass * P->visc / rho[i];
            }
        }
        for (int d = 0; d < 3; ++d) out[3 * (size_t)i + d] = (acc[d] < 1) ? acc[d] : 1.0;
    }
}

/*
 * One full VoxelSPHStrategy.compute_next_state (abstract_sph_strategy.py:31-46, voxel_sph_strategy.py:19-116).
 * pos/vel are updated in place; every other output pointer may be NULL.  Returns the number of dead particles.
 */
int32_t orc_step(const OrcParams *P, double *pos, double *vel, uint64_t *rng, double *rho_out, double *force_out,
                 double *pressure_out, double *viscosity_out, int32_t *keys_out, int32_t *map_ids_out,
                 int32_t *voxel_begin_out, int32_t *neigh_cnt_out, int32_t *neigh_out) {
    const int32_t n = P->n;
    const int64_t ncell = n_cells_of(P);
    int32_t *keys = keys_out ? keys_out : (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
    int32_t *ids = map_ids_out ? map_ids_out : (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
    int32_t *vb = voxel_begin_out ? voxel_begin_out : (int32_t *)malloc(sizeof(int32_t) * (size_t)(ncell + 1));
    int32_t *cnt = neigh_cnt_out ? neigh_cnt_out : (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
    int32_t *neigh = neigh_out ? neigh_out : (int32_t *)malloc(sizeof(int32_t) * (size_t)n * ORC_MAX_NEIGHBOURS);
    double *rho = rho_out ? rho_out : (double *)malloc(sizeof(double) * (size_t)n);
    double *force = force_out ? force_out : (double *)malloc(sizeof(double) * 3 * (size_t)n);
    double *pr = pressure_out ? pressure_out : (double *)malloc(sizeof(double) * 3 * (size_t)n);
    double *vi = viscosity_out ? viscosity_out : (double *)malloc(sizeof(double) * 3 * (size_t)n);

    orc_cell_keys(P, pos, keys);
    int32_t n_dead = orc_sort_cells(P, keys, ids, NULL, vb);
    orc_neighbours(P, pos, ids, vb, n_dead, neigh, cnt);
    orc_density(P, pos, neigh, cnt, rho);
    orc_pressure(P, pos, rho, neigh, cnt, pr);
    orc_viscosity(P, pos, vel, rho, neigh, cnt, vi);
    orc_integrate(P, pos, vel, rho, pr, vi, force);
    if (P->mode == 1)
        orc_collide_pipe(P, pos, vel, rng);
    else
        orc_collide_box(P, pos, vel);

    if (!keys_out) free(keys);
    if (!map_ids_out) free(ids);
    if (!voxel_begin_out) free(vb);
    if (!neigh_cnt_out) free(cnt);
    if (!neigh_out) free(neigh);
    if (!rho_out) free(rho);
    if (!force_out) free(force);
    if (!pressure_out) free(pr);
    if (!viscosity_out) free(vi);
    return n_dead;
}
