// Shared pieces of the neighbour sweeps (pair terms, the fp64 integrate + collide epilogue `finish_particle`, the exact
// one-thread `thread_walk` fallback) and the FIRST-GENERATION sweeps: density_kernel / force_kernel with per-warp
// shared-memory tiles.  The engine runs the row-staged sweeps of sweep_rows.cuh by default; these are selected with
// SPH_SWEEP=warp and kept for A/B measurements (tests/test_gpu_parity.py::test_kernel_variants_agree).
//
// Reference semantics (voxel_kernels.py:29-85): a particle's neighbour list is the first 32 candidates (self included)
// that pass sqrt(r^2) <= INF_R when the <= 27 cells around it are walked with dx outermost and dz innermost and the
// particles inside a cell in ascending particle id.  The sorted arrays are cell-contiguous with ascending id inside a
// cell, so the candidate sequence of a cell is the concatenation of <= 27 contiguous ranges ("virtual list"), and it
// is THE SAME for every particle of that cell.
//
// Work item = a chunk of <= 32 consecutive particles of one cell, owned by one warp (the warp that holds the chunk's
// first particle among its 32 sorted slots).  Per item:
//   1. lanes 0..26 fetch the 27 cell ranges, a warp scan turns the counts into offsets of the virtual list (T entries);
//   2. the virtual list is staged, TILE entries at a time, into the warp's private shared-memory slice with coalesced
//      loads (the force sweep also stages velocity and the per-candidate factors p_j/rho_j^2, LAP_W_CONST/rho_j, so
//      they are computed once per candidate, not once per pair);
//   3. density sweep: for each particle of the chunk, the 32 LANES TEST 32 CANDIDATES PER ROUND straight from shared
//      memory; a ballot + popc prefix implements "first 32 hits" exactly, so no lane ever idles behind a slower
//      neighbour (scan lengths inside one cell differ by 5x).  Accepted candidates are written as 16-bit virtual
//      indices to a per-particle list (64 B / particle in HBM) and their poly6 terms are warp-reduced;
//   4. force sweep: rebuilds the same tile, then every LANE OWNS ONE PARTICLE and runs down its own list (the pair
//      math happens once per accepted pair, in list order == the reference's summation order, no reductions), and
//      finishes with the fp64 integrate + collide epilogue and the scatter to the id-ordered master arrays.
// Particles whose own cell coordinates differ from their sort cell (aliased keys, reference quirk Q5) and chunks whose
// virtual list does not fit 16-bit indices take a plain one-thread-per-particle walk instead.
#pragma once
#include "collide.cuh"
#include "sph_common.cuh"

namespace sph {

constexpr int SW_THREADS = 128;
constexpr int SW_WARPS = SW_THREADS / 32;
constexpr int TILE = 256;            // candidates staged per piece (per warp)
constexpr int LIST_STRIDE = 34;      // uint16 entries per list row in shared memory (32 + pad: conflict-free rows)
constexpr unsigned FULL = 0xffffffffu;
constexpr uint8_t CNT_WALK = 0x80;   // neighbour-count flag: this particle is handled by thread_walk

struct TilePlan;   // sweep_rows.cuh

struct SweepArgs {
    const TilePlan *plans;   // row-staged sweeps: one plan per 128-particle tile (rows_plan_kernel)
    const int *n_items;      // work items of the sweeps: the 32-particle passes of tiles whose rows do not fit
    const int *items;        //   tile * 8 + 1 + pass
    const int *n_dense;      // ... and the tiles of a few dense cells (sweep_dense.cuh): tile * 8
    const int *dense_items;
    const float4 *spos;
    const float4 *svel;
    const uint32_t *skeys;
    const uint32_t *sids;
    const int2 *cell_range;
    float *srho;          // density sweep: out; force sweep: in
    uint16_t *nlist;      // [n][32] virtual-list indices of the accepted neighbours (density out, force in)
    uint8_t *ncnt;        // [n] min(32, #in range) | CNT_WALK
    float4 *pos_m, *vel_m, *sforce, *spress, *svisc;  // force sweep outputs
    const double *pipe;
    uint64_t *rng;
    const int32_t *gid;   // x-slab mode: global particle id per local index (rng states are indexed by it); else null
    int n;                // local particles (owned + ghosts)
    int n_own;            // local indices < n_own are owned: only those are integrated and written back
};

// Exact fp64 predicate sqrt(dx^2+dy^2+dz^2) <= INF_R on the promoted fp32 coordinates, no FMA contraction (the
// simulator has none).  sqrt is monotone and correctly rounded, hence equivalent to r2 <= r2_max.
__device__ __noinline__ bool in_range_exact(float ax, float ay, float az, float bx, float by, float bz,
                                            double r2_max) {
    const double dx = (double)ax - (double)bx, dy = (double)ay - (double)by, dz = (double)az - (double)bz;
    const double r2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
    return r2 <= r2_max;
}

// fp64 r^2 of two fp32 points, same operation order as in_range_exact
__device__ __noinline__ double r2_exact(float ax, float ay, float az, float bx, float by, float bz) {
    const double dx = (double)ax - (double)bx, dy = (double)ay - (double)by, dz = (double)az - (double)bz;
    return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}

// poly6 term (h^2 - r^2)^3.  Close to the cut-off h^2 - r^2 cancels: there it is formed from the fp64 r^2 (the reference
// is fp64 throughout), which keeps the density of a particle whose only neighbours sit at the cut-off within tolerance.
// The density of a particle is ONE of two sums over its list, chosen by its neighbour count alone (so that every code
// path -- staged rows, 32-particle passes, one-thread walk, any slab decomposition -- produces the same bits):
//   count > 8 : sum of poly6_fast (fp32; a term near the cut-off is negligible among >= 8 others)
//   count <= 8: sum of poly6_term (terms near the cut-off from the fp64 r^2)
// Both are written with explicit roundings: no FMA contraction may differ between call sites.
__device__ __forceinline__ float poly6_fast(const StepConsts &c, float r2) {
    const float d = __fsub_rn(c.h2, r2);
    return __fmul_rn(__fmul_rn(d, d), d);
}
__device__ __forceinline__ float poly6_term(const StepConsts &c, float r2, float ax, float ay, float az, float bx,
                                            float by, float bz) {
    if (r2 > c.h2_near) {
        const double d = c.h2_d - r2_exact(ax, ay, az, bx, by, bz);
        return (float)__dmul_rn(__dmul_rn(d, d), d);
    }
    return poly6_fast(c, r2);
}
constexpr int kSparseCount = 8;   // neighbour counts up to this use the poly6_term sum

__device__ __forceinline__ float pressure_coeff(const StepConsts &c, float rho) {
    return c.k * (rho - c.rho0) / (rho * rho);   // p / rho^2 with p = K (rho - RHO_0)
}

// Pressure + viscosity pair terms (voxel_kernels.py:155-168,195-208; base_kernels.py:12-27).
struct ForceAcc {
    float px = 0.f, py = 0.f, pz = 0.f, ux = 0.f, uy = 0.f, uz = 0.f;
    bool any = false;
    __device__ __forceinline__ void pair(const StepConsts &c, float dx, float dy, float dz, float r2, float a_i,
                                         float a_j, float b_j, const float4 &vi, const float4 &vj) {
        any = true;
        const float rinv = rsqrtf(r2);
        const float r = (r2 > 0.f) ? r2 * rinv : 0.f;
        const float hr = c.h - r;
        const float gw = (a_i + a_j) * (c.grad_c * hr * hr) * rinv;   // (p_i/rho_i^2 + p_j/rho_j^2) GRAD_W (h-r)^2 / r
        px = fmaf(gw, dx, px);
        py = fmaf(gw, dy, py);
        pz = fmaf(gw, dz, pz);
        const float lw = hr * b_j;                                     // LAP_W_CONST (h-r) / rho_j
        ux = fmaf(vj.x - vi.x, lw, ux);
        uy = fmaf(vj.y - vi.y, lw, uy);
        uz = fmaf(vj.z - vi.z, lw, uz);
    }
    // branch-free variant for the list-driven sweep: pj = (x, y, z, a_j), vj = (vx, vy, vz, b_j); `on` masks the pair
    __device__ __forceinline__ void pair(const StepConsts &c, const float4 &pi, const float4 &vi, float a_i,
                                         const float4 &pj, const float4 &vj, bool on) {
        const float dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
        const float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
        const float rinv = rsqrtf(r2);
        const float r = (r2 > 0.f) ? r2 * rinv : 0.f;
        const float hr = c.h - r;
        float gw = (a_i + pj.w) * (c.grad_c * hr * hr) * rinv;
        float lw = hr * vj.w;
        gw = on ? gw : 0.f;
        lw = on ? lw : 0.f;
        any = any || on;
        px = fmaf(gw, dx, px);
        py = fmaf(gw, dy, py);
        pz = fmaf(gw, dz, pz);
        ux = fmaf(vj.x - vi.x, lw, ux);
        uy = fmaf(vj.y - vi.y, lw, uy);
        uz = fmaf(vj.z - vi.z, lw, uz);
    }
};

// integrating_kernel + collision kernel in fp64 (base_kernels.py:30-98), scatter to the id-ordered master arrays.
template <bool RECORD_TERMS>
__device__ __forceinline__ void finish_particle(const SweepArgs &a, const StepConsts &c, int t, const float4 pi,
                                                const float4 vi, float rho_i, ForceAcc f, uint32_t id,
                                                bool maybe_empty = true) {
    if ((int)id >= a.n_own) return;              // ghost particle of an x-slab: its owner integrates it
    // empty slot of an x-slab (hole left by an emigrant / unused capacity): x = NaN, so only a DEAD particle can be one --
    // callers that know their particle is alive skip this random 4-byte gather
    if (maybe_empty && a.gid && a.gid[id] < 0) return;
    if (f.any) {  // with no neighbour besides self the reference's sums stay exactly 0 (and rho_i is 0)
        const float s = c.mass_visc / rho_i;
        f.ux *= s;
        f.uy *= s;
        f.uz *= s;
    }
    // min(1, x) with Python semantics: x if x < 1 else 1   (voxel_kernels.py:211; NaN -> 1)
    f.ux = (f.ux < 1.f) ? f.ux : 1.f;
    f.uy = (f.uy < 1.f) ? f.uy : 1.f;
    f.uz = (f.uz < 1.f) ? f.uz : 1.f;
    double F[3] = {c.ext[0] + -(double)f.px + (double)f.ux, c.ext[1] + -(double)f.py + (double)f.uy,
                   c.ext[2] + -(double)f.pz + (double)f.uz};
    double v[3] = {vi.x, vi.y, vi.z};
    double x[3] = {pi.x, pi.y, pi.z};
    const double s_dt = c.dt / (double)rho_i;   // F / rho * dt with one division (rho = 0 keeps the inf / NaN semantics)
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        v[d] += F[d] * s_dt;
        x[d] += v[d] * c.dt;
    }
    if (c.mode == 1) {
        PipeView pv{a.pipe, c.pipe_rows};
        collide_pipe(pv, x, v, a.rng + 2 * (size_t)(a.gid ? (uint32_t)a.gid[id] : id));
    } else {
        collide_box(x, v, c);
    }
    a.pos_m[MI(id)] = make_float4((float)x[0], (float)x[1], (float)x[2], rho_i);
    a.vel_m[MI(id)] = make_float4((float)v[0], (float)v[1], (float)v[2], 0.f);
    a.sforce[t] = make_float4((float)F[0], (float)F[1], (float)F[2], 0.f);
    if (RECORD_TERMS) {
        a.spress[t] = make_float4(f.px, f.py, f.pz, 0.f);
        a.svisc[t] = make_float4(f.ux, f.uy, f.uz, 0.f);
    }
}

// ---- fallback: one thread, one particle, candidates straight from global memory -------------------------------------
template <bool FORCE>
__device__ __noinline__ int thread_walk(const SweepArgs &a, const GridDesc &g, const StepConsts &c, int t,
                                        const float4 pi, const float4 vi, float a_i, float &dens, ForceAcc &f) {
    int vx, vy, vz;
    if (!cell_of(g, pi.x, pi.y, pi.z, vx, vy, vz)) return 0;
    int cnt = 0;
    float dens_fast = 0.f, dens_exact = 0.f;   // see poly6_fast / poly6_term: the count decides which one is the density
    for (int dx = -1; dx <= 1; ++dx) {
        const int x = vx + dx;
        if (x < 0 || x >= g.tx) continue;
        for (int dy = -1; dy <= 1; ++dy) {
            const int y = vy + dy;
            if (y < 0 || y >= g.ty) continue;
            for (int dz = -1; dz <= 1; ++dz) {
                const int z = vz + dz;
                if (z < 0 || z >= g.tz) continue;
                const long long cl = (long long)x - g.xoff + (long long)y * g.wn + (long long)z * g.wn * g.hn;
                if (cl < 0 || cl >= g.ncells) continue;
                const int2 r = __ldg(&a.cell_range[cl]);
                for (int j = r.x; j < r.y; ++j) {
                    const float4 pj = __ldg(&a.spos[j]);
                    const float ddx = pi.x - pj.x, ddy = pi.y - pj.y, ddz = pi.z - pj.z;
                    const float r2 = fmaf(ddz, ddz, fmaf(ddy, ddy, ddx * ddx));
                    bool in = r2 <= c.h2_lo;
                    if (!in && r2 < c.h2_hi) in = in_range_exact(pi.x, pi.y, pi.z, pj.x, pj.y, pj.z, c.r2_max);
                    if (in) {
                        if (j != t) {
                            if (FORCE) {
                                const float rho_j = __ldg(&a.srho[j]);
                                const float4 vj = __ldg(&a.svel[j]);
                                f.pair(c, ddx, ddy, ddz, r2, a_i, pressure_coeff(c, rho_j), c.lap_c / rho_j, vi, vj);
                            } else {
                                dens_fast = __fadd_rn(dens_fast, poly6_fast(c, r2));
                                if (cnt < kSparseCount)
                                    dens_exact = __fadd_rn(dens_exact, poly6_term(c, r2, pi.x, pi.y, pi.z, pj.x, pj.y, pj.z));
                            }
                        }
                        if (++cnt >= kMaxNeighbours) {
                            dens = dens_fast;
                            return cnt;
                        }
                    }
                }
            }
        }
    }
    dens = (cnt <= kSparseCount) ? dens_exact : dens_fast;
    return cnt;
}

// Segment table of a work item: the 27 neighbour cells as ranges of the sorted arrays + offsets into the virtual list.
struct SegTable {
    int seg_start[32];   // first sorted index of each of the 27 neighbour cells
    int seg_off[33];     // exclusive offsets into the virtual list; [27..32] = T
};

// Per-warp shared-memory slice of the density sweep.  The tile keeps NEGATED coordinates, two candidates per entry
// (q and q + 32 of a 64-wide round), so a round is three LDS.64 + packed f32x2 math (sm_100 FADD2 / FMUL2 / FFMA2).
struct DensitySmem {
    float2 tx[TILE / 2], ty[TILE / 2], tz[TILE / 2];
    float4 ppos[32];                    // positions of the chunk's particles (broadcast reads)
    uint16_t list[32 * LIST_STRIDE];    // per particle: accepted virtual indices, in scan order
    SegTable seg;
};

// Per-warp shared-memory slice of the force sweep.
struct ForceSmem {
    float4 tile_pos[TILE];              // (x, y, z, a_j = p_j / rho_j^2)
    float4 tile_vel[TILE];              // (vx, vy, vz, b_j = LAP_W_CONST / rho_j)
    uint16_t list[32 * LIST_STRIDE];
    SegTable seg;
};

// Work-item header shared by both sweeps: decode the chunk, build the segment table.  Returns T (virtual list length).
__device__ __forceinline__ void decode_cell(const GridDesc &g, uint32_t ckey, int &cx, int &cy, int &cz) {
    cz = (int)(ckey / (uint32_t)(g.wk * g.hk));
    const int rem = (int)(ckey - (uint32_t)cz * (uint32_t)(g.wk * g.hk));
    cy = rem / g.wk;
    cx = rem - cy * g.wk + g.xoff;
}

__device__ __forceinline__ int build_segments(const SweepArgs &a, const GridDesc &g, SegTable &sm, int lane, int cx,
                                              int cy, int cz) {
    int2 r = make_int2(0, 0);
    if (lane < 27) {   // lane -> (dx, dy, dz) with dx outermost, dz innermost (voxel_kernels.py:46-48)
        const int x = cx + lane / 9 - 1, y = cy + (lane / 3) % 3 - 1, z = cz + lane % 3 - 1;
        if (x >= 0 && x < g.tx && y >= 0 && y < g.ty && z >= 0 && z < g.tz) {
            const long long cl = (long long)x - g.xoff + (long long)y * g.wn + (long long)z * g.wn * g.hn;
            if (cl >= 0 && cl < g.ncells) r = __ldg(&a.cell_range[cl]);
        }
    }
    const int count = r.y - r.x;
    int inc = count;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(FULL, inc, o);
        if (lane >= o) inc += u;
    }
    sm.seg_start[lane] = r.x;
    sm.seg_off[lane] = inc - count;
    const int T = __shfl_sync(FULL, inc, 31);
    if (lane == 0) sm.seg_off[32] = T;
    __syncwarp();
    return T;
}

// sorted index of virtual-list entry v (binary search over the 32 offsets)
__device__ __forceinline__ int virtual_to_sorted(const SegTable &sm, int v) {
    int s = 0;
#pragma unroll
    for (int step = 16; step > 0; step >>= 1)
        if (sm.seg_off[s + step] <= v) s += step;
    return sm.seg_start[s] + (v - sm.seg_off[s]);
}

// ---------------------------------------------------------------------------------------------------------------------
// density_kernel (voxel_kernels.py:108-132): rho_i = MASS * sum_{j != i} W_CONST (h^2 - r^2)^3 + neighbour lists
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SW_THREADS)
density_kernel(const SweepArgs a, const GridDesc g, const StepConsts c) {
    __shared__ DensitySmem smem[SW_WARPS];
    DensitySmem &sm = smem[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    const int t = blockIdx.x * SW_THREADS + threadIdx.x;
    const bool valid = t < a.n;
    const uint32_t key = valid ? a.skeys[t] : (uint32_t)g.ncells;
    const bool live = key != (uint32_t)g.ncells;
    const int2 cr = live ? __ldg(&a.cell_range[key]) : make_int2(0, 0);
    if (valid && !live) {   // dead particle: no neighbours
        a.srho[t] = 0.f;
        a.ncnt[t] = 0;
    }
    unsigned leaders = __ballot_sync(FULL, live && (((t - cr.x) & 31) == 0));
    const uint32_t lt = (1u << lane) - 1u;
    const float SENTINEL = -1e30f;   // negated coordinate of a padding candidate: r2 = +inf, never accepted

    while (leaders) {
        const int src = __ffs(leaders) - 1;
        leaders &= leaders - 1;
        const int cs = __shfl_sync(FULL, t, src);            // first particle of this chunk
        const int ce = __shfl_sync(FULL, cr.y, src);         // end of the cell
        const uint32_t ckey = __shfl_sync(FULL, key, src);
        const int m = min(32, ce - cs);
        const int tt = cs + lane;
        const bool mine = lane < m;
        int cx, cy, cz;
        decode_cell(g, ckey, cx, cy, cz);
        if (cx < g.own_lo - 1 || cx > g.own_hi) continue;   // x-slab: outer ghost column, nobody needs its density
        const int T = build_segments(a, g, sm.seg, lane, cx, cy, cz);

        float4 pi = make_float4(0.f, 0.f, 0.f, 0.f);
        bool walk = false;
        if (mine) {
            pi = a.spos[tt];
            int vx, vy, vz;
            walk = !cell_of(g, pi.x, pi.y, pi.z, vx, vy, vz) || vx != cx || vy != cy || vz != cz;
        }
        if (T > 65535) walk = mine;
        sm.ppos[lane] = pi;
        // virtual index of this lane's own particle: own cell is segment 13 (dx = dy = dz = 0)
        const int self_v = sm.seg.seg_off[13] + (tt - sm.seg.seg_start[13]);
        int my_cnt = 0;          // neighbours accepted so far == entries in my list
        float my_rho = 0.f;
        unsigned pending = __ballot_sync(FULL, mine && !walk);   // particles still collecting neighbours

        for (int pbase = 0; pbase < T && pending; pbase += TILE) {
            const int plen = min(TILE, T - pbase);
            __syncwarp();
            // ---- stage the piece (coalesced gathers, negated coordinates, sentinel padding) + clear the masks ----
#pragma unroll
            for (int k = 0; k < TILE / 32; ++k) {
                const int q = k * 32 + lane;
                float nx = SENTINEL, ny = SENTINEL, nz = SENTINEL;
                if (q < plen) {
                    const float4 pj = __ldg(&a.spos[virtual_to_sorted(sm.seg, pbase + q)]);
                    nx = -pj.x;
                    ny = -pj.y;
                    nz = -pj.z;
                }
                float *fx = reinterpret_cast<float *>(sm.tx), *fy = reinterpret_cast<float *>(sm.ty),
                      *fz = reinterpret_cast<float *>(sm.tz);
                const int slot = ((k >> 1) * 32 + lane) * 2 + (k & 1);
                fx[slot] = nx;
                fy[slot] = ny;
                fz[slot] = nz;
            }
            __syncwarp();

            // ---- rounds: 32 lanes x 2 candidates against one particle at a time ----
            unsigned todo = pending;
            while (todo) {
                const int i = __ffs(todo) - 1;
                todo &= todo - 1;
                const float4 p = sm.ppos[i];                     // broadcast
                const float2 px2 = make_float2(p.x, p.x), py2 = make_float2(p.y, p.y), pz2 = make_float2(p.z, p.z);
                const int self_i = __shfl_sync(FULL, self_v, i) - pbase - lane;   // == r*64 (+32) for the own slot
                int cnt = __shfl_sync(FULL, my_cnt, i);
                uint16_t *wr = &sm.list[i * LIST_STRIDE + cnt];   // next free list slot of particle i
                float part = 0.f;
#pragma unroll
                for (int r = 0; r < TILE / 64; ++r) {
                    if (r * 64 >= plen) break;
                    const float2 dx = __fadd2_rn(px2, sm.tx[r * 32 + lane]);
                    const float2 dy = __fadd2_rn(py2, sm.ty[r * 32 + lane]);
                    const float2 dz = __fadd2_rn(pz2, sm.tz[r * 32 + lane]);
                    const float2 r2 = __ffma2_rn(dz, dz, __ffma2_rn(dy, dy, __fmul2_rn(dx, dx)));
                    unsigned m0 = __ballot_sync(FULL, r2.x <= c.h2_lo), m1 = __ballot_sync(FULL, r2.y <= c.h2_lo);
                    // rare (warp-uniform branch): some candidate sits inside the rounding band -> fp64 decides
                    if (__any_sync(FULL, (r2.x > c.h2_lo && r2.x < c.h2_hi) || (r2.y > c.h2_lo && r2.y < c.h2_hi))) {
                        bool in0 = r2.x <= c.h2_lo, in1 = r2.y <= c.h2_lo;
                        if (!in0 && r2.x < c.h2_hi)
                            in0 = in_range_exact(p.x, p.y, p.z, p.x - dx.x, p.y - dy.x, p.z - dz.x, c.r2_max);
                        if (!in1 && r2.y < c.h2_hi)
                            in1 = in_range_exact(p.x, p.y, p.z, p.x - dx.y, p.y - dy.y, p.z - dz.y, c.r2_max);
                        m0 = __ballot_sync(FULL, in0);
                        m1 = __ballot_sync(FULL, in1);
                    }
                    // list positions: "first 32 hits" == position < 32
                    const int n0 = __popc(m0);
                    const int k0 = __popc(m0 & lt), k1 = n0 + __popc(m1 & lt);
                    const int room = kMaxNeighbours - cnt;
                    const bool t0 = ((m0 >> lane) & 1u) && k0 < room, t1 = ((m1 >> lane) & 1u) && k1 < room;
                    if (t0) wr[k0] = (uint16_t)(pbase + r * 64 + lane);
                    if (t1) wr[k1] = (uint16_t)(pbase + r * 64 + 32 + lane);
                    // poly6 terms of the accepted candidates (self excluded): (h^2 - r^2)^3
                    const float d0 = c.h2 - r2.x, d1 = c.h2 - r2.y;
                    const float w0 = (t0 && self_i != r * 64) ? d0 * d0 * d0 : 0.f;
                    const float w1 = (t1 && self_i != r * 64 + 32) ? d1 * d1 * d1 : 0.f;
                    part += w0 + w1;
                    const int n = n0 + __popc(m1);
                    cnt += n;
                    wr += n;
                    if (cnt >= kMaxNeighbours) {
                        cnt = kMaxNeighbours;
                        pending &= ~(1u << i);
                        break;
                    }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(FULL, part, o);
                if (lane == i) {
                    my_cnt = cnt;
                    my_rho += part;
                }
            }
        }
        __syncwarp();
        // results: density, count, neighbour list (rows are contiguous in HBM: 32 x 64 B per chunk)
        if (mine) {
            uint8_t cflag = (uint8_t)my_cnt;
            if (walk) {
                ForceAcc dummy;
                float dens = 0.f;
                const int wc = thread_walk<false>(a, g, c, tt, pi, pi, 0.f, dens, dummy);
                my_rho = dens;
                cflag = (uint8_t)wc | CNT_WALK;
            }
            a.srho[tt] = my_rho * c.w_mass;
            a.ncnt[tt] = cflag;
        }
        uint32_t *gl = reinterpret_cast<uint32_t *>(a.nlist + (size_t)cs * 32);
#pragma unroll 4
        for (int it = 0; it < 16; ++it) {
            const int row = it * 2 + (lane >> 4), col = (lane & 15) * 2;
            if (row < m) {
                const uint32_t lo = sm.list[row * LIST_STRIDE + col], hi = sm.list[row * LIST_STRIDE + col + 1];
                gl[row * 16 + (lane & 15)] = lo | (hi << 16);
            }
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// pressure_kernel + viscosity_kernel + integrating_kernel + collision kernel in one sweep
// (voxel_kernels.py:135-211, base_kernels.py:30-98), driven by the neighbour lists of the density sweep.
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async4(void *smem_dst, const void *gmem_src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

template <bool RECORD>
__global__ void __launch_bounds__(SW_THREADS, 4)
force_kernel(const SweepArgs a, const GridDesc g, const StepConsts c) {
    __shared__ ForceSmem smem[SW_WARPS];
    ForceSmem &sm = smem[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    const int t = blockIdx.x * SW_THREADS + threadIdx.x;
    const bool valid = t < a.n;
    const uint32_t key = valid ? a.skeys[t] : (uint32_t)g.ncells;
    const bool live = key != (uint32_t)g.ncells;
    const int2 cr = live ? __ldg(&a.cell_range[key]) : make_int2(0, 0);
    if (valid && !live)   // dead particle: F = external force, rho = 0 (reference NaN semantics carry on)
        finish_particle<RECORD>(a, c, t, a.spos[t], a.svel[t], a.srho[t], ForceAcc(), a.sids[t]);
    unsigned leaders = __ballot_sync(FULL, live && (((t - cr.x) & 31) == 0));

    while (leaders) {
        const int src = __ffs(leaders) - 1;
        leaders &= leaders - 1;
        const int cs = __shfl_sync(FULL, t, src);
        const int ce = __shfl_sync(FULL, cr.y, src);
        const uint32_t ckey = __shfl_sync(FULL, key, src);
        const int m = min(32, ce - cs);
        const int tt = cs + lane;
        const bool mine = lane < m;
        int cx, cy, cz;
        decode_cell(g, ckey, cx, cy, cz);
        if (cx < g.own_lo || cx >= g.own_hi) continue;      // x-slab: ghost cell, its owner computes the forces
        const int T = build_segments(a, g, sm.seg, lane, cx, cy, cz);

        float4 pi = make_float4(0.f, 0.f, 0.f, 0.f), vi = pi;
        float rho_i = 0.f, a_i = 0.f;
        int my_cnt = 0;
        bool walk = false;
        if (mine) {
            pi = a.spos[tt];
            vi = a.svel[tt];
            rho_i = a.srho[tt];
            a_i = pressure_coeff(c, rho_i);
            const uint8_t cf = a.ncnt[tt];
            walk = (cf & CNT_WALK) != 0;
            my_cnt = walk ? 0 : cf;
        }
        // neighbour lists of the chunk: HBM -> shared (coalesced 2 KiB)
        const uint32_t *gl = reinterpret_cast<const uint32_t *>(a.nlist + (size_t)cs * 32);
#pragma unroll 4
        for (int it = 0; it < 16; ++it) {
            const int row = it * 2 + (lane >> 4), col = (lane & 15) * 2;
            if (row < m) {
                const uint32_t w = __ldg(&gl[row * 16 + (lane & 15)]);
                sm.list[row * LIST_STRIDE + col] = (uint16_t)(w & 0xffffu);
                sm.list[row * LIST_STRIDE + col + 1] = (uint16_t)(w >> 16);
            }
        }
        const int self_v = sm.seg.seg_off[13] + (tt - sm.seg.seg_start[13]);
        ForceAcc f;
        int cur = 0;
        unsigned pending = __ballot_sync(FULL, cur < my_cnt);

        for (int pbase = 0; pbase < T && pending; pbase += TILE) {
            const int plen = min(TILE, T - pbase), pend = pbase + plen;
            __syncwarp();
            // ---- stage the piece: positions and velocities by cp.async, densities through registers ----
            float rho_j[TILE / 32];
#pragma unroll
            for (int k = 0; k < TILE / 32; ++k) {
                const int q = k * 32 + lane;
                rho_j[k] = 1.f;
                if (q < plen) {
                    const int j = virtual_to_sorted(sm.seg, pbase + q);
                    cp_async16(&sm.tile_pos[q], &a.spos[j]);
                    cp_async16(&sm.tile_vel[q], &a.svel[j]);
                    rho_j[k] = __ldg(&a.srho[j]);
                }
            }
            cp_async_wait_all();
#pragma unroll
            for (int k = 0; k < TILE / 32; ++k) {
                const int q = k * 32 + lane;
                if (q < plen) {   // per-candidate factors, once per candidate instead of once per pair
                    sm.tile_pos[q].w = pressure_coeff(c, rho_j[k]);
                    sm.tile_vel[q].w = c.lap_c / rho_j[k];
                }
            }
            __syncwarp();
            // ---- lane = particle: run down the own list while its entries fall into this piece (2 per trip) ----
            const uint16_t *row = &sm.list[lane * LIST_STRIDE];
            while (cur < my_cnt) {
                const int v0 = row[cur];
                if (v0 >= pend) break;
                const bool two = (cur + 1 < my_cnt) && (row[cur + 1] < pend);
                const int v1 = two ? row[cur + 1] : v0;
                const float4 p0 = sm.tile_pos[v0 - pbase], p1 = sm.tile_pos[v1 - pbase];
                const float4 w0 = sm.tile_vel[v0 - pbase], w1 = sm.tile_vel[v1 - pbase];
                f.pair(c, pi, vi, a_i, p0, w0, v0 != self_v);
                f.pair(c, pi, vi, a_i, p1, w1, two && v1 != self_v);
                cur += two ? 2 : 1;
            }
            pending = __ballot_sync(FULL, cur < my_cnt);
        }
        __syncwarp();
        if (mine) {
            if (walk) {
                float dens = 0.f;
                f = ForceAcc();
                thread_walk<true>(a, g, c, tt, pi, vi, a_i, dens, f);
            }
            finish_particle<RECORD>(a, c, tt, pi, vi, rho_i, f, a.sids[tt]);
        }
        __syncwarp();
    }
}

}  // namespace sph
