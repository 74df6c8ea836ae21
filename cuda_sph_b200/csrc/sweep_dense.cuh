// Dense tiles (sm_100a): neighbour lists + poly6 density and the list-driven force sweep for the 128-particle tiles whose
// candidate rows do not fit the shared-memory staging of the main sweeps (rows_plan_kernel makes them work items).  The
// reference's own pipe workload lives here: config.py's pipe holds 110 .. 840 particles per cell, and from its second
// step on the recycle rule piles 1e5 particles into each inlet cell (base_kernels.py:56-72).
//
// Reference semantics (voxel_kernels.py:29-85): first 32 candidates with sqrt(r^2) <= INF_R when the <= 27 cells are
// walked dx outermost / dz innermost, every cell in ascending particle id.  At these densities a particle tests
// 600 .. 6000 candidates before its 32nd hit (the whole dx = -1 block of a particle with a large fractional x holds
// almost none), so the sweep is a brute-force scan organised for instruction issue:
//   * one CELL at a time: all particles of a cell share one candidate sequence (the cell's 27 ranges of the sorted arrays),
//     which is streamed through shared memory in chunks of DN_CHUNK candidates (SoA x | y | z, coalesced 16-B loads);
//   * lane = particle, lockstep over the warp, 32 candidates per round with the packed f32x2 superset test of
//     density_flat_kernel (sign bits shifted into a hit mask; non-zero masks kept per lane, a lane stops at FL_KEEP hits);
//     the chunk loop ends as soon as every lane of the cell is done, so nobody stages what nobody reads;
//   * lists: the set bits are mapped back to sorted indices through the cell's range table, r^2 is recomputed the
//     canonical way, the fp64 predicate decides inside the rounding band, the density is summed in list order.  Lists
//     leave as 32-bit sorted indices (dlist) for force_gather_kernel.
// Same bits as every other path (canonical density sums, sweep.cuh).
#pragma once
#include "sweep_flat.cuh"

namespace sph {

constexpr int DN_THREADS = RB_THREADS;
constexpr int DN_CHUNK = 2048;                 // candidates staged per chunk
constexpr int DENSE_MAX_CELLS = 3;             // tiles of at most this many cells (> 42 particles per cell) come here
constexpr int DN_MROWS = FL_KEEP + 1;          // non-zero masks per lane (+ one spare row)
static_assert(DN_CHUNK % (4 * DN_THREADS) == 0, "staging runs four candidates per thread and trip");

struct DenseSmem {
    float x[DN_CHUNK], y[DN_CHUNK], z[DN_CHUNK];
    uint32_t masks[DN_MROWS * DN_THREADS];     // per thread: its non-zero hit masks in scan order (first candidate = bit 31)
    uint32_t mround[DN_MROWS * DN_THREADS];    // round (32 candidates) of the mask word, counted over the whole window
    int seg_g[27];                             // the current cell's walk: first sorted index of every segment
    int seg_off[28];                           // exclusive prefix of the segment lengths; [27] = window length
    uint32_t ckey[DN_THREADS];                 // the tile's non-empty cells
    int cfirst[DN_THREADS + 1];                // first lane of every cell; [ncell] = live lanes
    int wcount[DN_THREADS / 32], wlive[DN_THREADS / 32];
};
static_assert(sizeof(DenseSmem) <= 75 * 1024, "three CTAs per SM");

// walk position -> sorted index; s is a running segment pointer (positions ascend)
__device__ __forceinline__ int dense_index(const DenseSmem &sm, int w, int &s) {
    while (w >= sm.seg_off[s + 1]) ++s;
    return sm.seg_g[s] + (w - sm.seg_off[s]);
}

__device__ __forceinline__ void dense_tile(const SweepArgs &a, const GridDesc &g, const StepConsts &c,
                                           uint32_t *__restrict__ dlist, const int tile, DenseSmem &sm) {
    const int j = threadIdx.x, lane = j & 31, warp = j >> 5;
    const int p0 = tile * DN_THREADS;
    const int t = p0 + j;
    const int nb = min(DN_THREADS, a.n - p0);
    const uint32_t key = (j < nb) ? a.skeys[t] : (uint32_t)g.ncells;
    const uint32_t prev_key = (j < nb && j > 0) ? a.skeys[t - 1] : 0xffffffffu;
    const float4 pi = (j < nb) ? a.spos[t] : make_float4(0.f, 0.f, 0.f, 0.f);
    const bool live = key != (uint32_t)g.ncells;
    if (j < nb && !live) {   // dead particle: no neighbours
        a.srho[t] = 0.f;
        a.ncnt[t] = 0;
    }
    // ---- the tile's non-empty cells (dead particles sort to the tail: the live lanes are a prefix) --------------------
    const bool first = live && (j == 0 || prev_key != key);
    const unsigned bal = __ballot_sync(FULL, first), lbal = __ballot_sync(FULL, live);
    if (lane == 0) {
        sm.wcount[warp] = __popc(bal);
        sm.wlive[warp] = __popc(lbal);
    }
    __syncthreads();
    int coff = 0, ncell = 0, nlive = 0;
#pragma unroll
    for (int w = 0; w < DN_THREADS / 32; ++w) {
        const int n = sm.wcount[w];
        if (w < warp) coff += n;
        ncell += n;
        nlive += sm.wlive[w];
    }
    if (first) {
        const int ci = coff + __popc(bal & lanemask_le_()) - 1;
        sm.ckey[ci] = key;
        sm.cfirst[ci] = j;
    }
    if (j == 0) sm.cfirst[ncell] = nlive;
    int cx = 0, cy = 0, cz = 0;
    bool want = false, walk = false;
    if (live) {
        decode_cell(g, key, cx, cy, cz);
        want = !(cx < g.own_lo - 1 || cx > g.own_hi);   // x-slab: nobody needs the density of the outer ghost column
        walk = want && !own_cell_matches(g, pi, cx, cy, cz);   // aliased key (quirk Q5)
    }
    const bool scan = want && !walk;
    __syncthreads();

    const float2 npx = make_float2(-pi.x, -pi.x), npy = make_float2(-pi.y, -pi.y), npz = make_float2(-pi.z, -pi.z);
    const float2 lim2 = make_float2(-c.h2_hi, -c.h2_hi);
    for (int ci = 0; ci < ncell; ++ci) {
        const int lo = sm.cfirst[ci], hi = sm.cfirst[ci + 1];
        const bool mine = j >= lo && j < hi;
        // ---- the cell's walk: 27 ranges of the sorted arrays ----------------------------------------------------------
        if (warp == 0) {
            int len = 0, gs = 0;
            if (lane < 27) {
                int qx, qy, qz;
                decode_cell(g, sm.ckey[ci], qx, qy, qz);
                const int2 r = neighbour_range(a, g, lane, qx, qy, qz);
                const int x = qx + lane / 9 - 1;   // columns outside the local table of an x-slab do not exist here
                if (x >= g.xoff && x < g.xoff + g.wk) len = max(r.y - r.x, 0);
                gs = r.x;
            }
            int inc = len;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int u = __shfl_up_sync(FULL, inc, o);
                if (lane >= o) inc += u;
            }
            if (lane < 27) {
                sm.seg_g[lane] = gs;
                sm.seg_off[lane] = inc - len;
            }
            if (lane == 27) sm.seg_off[27] = inc;   // lanes >= 27 carry the total
        }
        // does any lane of this cell scan at all?  (CTA-uniform; also the barrier that publishes the table)
        const int any_scan = __syncthreads_or(mine && scan);
        const int T = sm.seg_off[27];
        bool act = mine && scan && T > 0;
        int cs = 0, nz = 0;
        if (any_scan) {
            for (int chunk0 = 0; chunk0 < T; chunk0 += DN_CHUNK) {
                const int n = min(DN_CHUNK, T - chunk0);
                // ---- stage: consecutive threads take consecutive candidates, four independent loads per trip ----------
                {
                    int s = 0;
                    for (int i0 = j; i0 < n; i0 += 4 * DN_THREADS) {
                        float4 v[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int i = i0 + u * DN_THREADS;
                            if (i < n) v[u] = __ldg(&a.spos[dense_index(sm, chunk0 + i, s)]);
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int i = i0 + u * DN_THREADS;
                            if (i < n) {
                                sm.x[i] = v[u].x;
                                sm.y[i] = v[u].y;
                                sm.z[i] = v[u].z;
                            }
                        }
                    }
                }
                __syncthreads();
                // ---- scan the chunk: 32 candidates per round, lockstep over the warp ---------------------------------
                {
                    const float4 *xq = reinterpret_cast<const float4 *>(sm.x);
                    const float4 *yq = reinterpret_cast<const float4 *>(sm.y);
                    const float4 *zq = reinterpret_cast<const float4 *>(sm.z);
                    const int round0 = chunk0 >> 5;
#pragma unroll 1
                    for (int k0 = 0; k0 < n; k0 += 32) {
                        if (!__any_sync(FULL, act)) break;
                        uint32_t m = 0;
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            const float4 X = xq[u], Y = yq[u], Z = zq[u];
                            const float2 dx0 = __fadd2_rn(make_float2(X.x, X.y), npx), dx1 = __fadd2_rn(make_float2(X.z, X.w), npx);
                            const float2 dy0 = __fadd2_rn(make_float2(Y.x, Y.y), npy), dy1 = __fadd2_rn(make_float2(Y.z, Y.w), npy);
                            const float2 dz0 = __fadd2_rn(make_float2(Z.x, Z.y), npz), dz1 = __fadd2_rn(make_float2(Z.z, Z.w), npz);
                            float2 s0 = __ffma2_rn(dx0, dx0, lim2), s1 = __ffma2_rn(dx1, dx1, lim2);
                            s0 = __ffma2_rn(dy0, dy0, s0);
                            s1 = __ffma2_rn(dy1, dy1, s1);
                            s0 = __ffma2_rn(dz0, dz0, s0);
                            s1 = __ffma2_rn(dz1, dz1, s1);
                            m = __funnelshift_l(__float_as_uint(s0.x), m, 1);
                            m = __funnelshift_l(__float_as_uint(s0.y), m, 1);
                            m = __funnelshift_l(__float_as_uint(s1.x), m, 1);
                            m = __funnelshift_l(__float_as_uint(s1.y), m, 1);
                        }
                        const int nv = n - k0;   // candidates of this round that belong to the chunk
                        if (nv < 32) m &= 0xffffffffu << (32 - nv);
                        if (act && m) {
                            sm.masks[nz * DN_THREADS + j] = m;
                            sm.mround[nz * DN_THREADS + j] = (uint32_t)(round0 + (k0 >> 5));
                            cs += __popc(m);
                            ++nz;
                        }
                        act = act && cs < FL_KEEP;
                        xq += 8;
                        yq += 8;
                        zq += 8;
                    }
                }
                // next chunk only while somebody is still looking (also the barrier before restaging)
                if (!__syncthreads_or(act && chunk0 + DN_CHUNK < T)) break;
            }
        }
        // ---- lists + density: run down the set bits (first 32 accepted), canonical r^2, poly6 in list order ------------
        if (mine && scan) {
            // act still set: the lane saw its whole window (the superset is complete)
            const bool complete = act || cs < FL_KEEP;
            uint32_t *gl = dlist + (size_t)t * kMaxNeighbours;
            int k = 0, s = 0;
            float dens_fast = 0.f, dens_exact = 0.f;
            for (int w = 0; w < nz && k < kMaxNeighbours; ++w) {
                uint32_t m = sm.masks[w * DN_THREADS + j];
                const int base = 32 * (int)sm.mround[w * DN_THREADS + j];
                while (m && k < kMaxNeighbours) {
                    const int b = __clz(m);
                    m &= 0x7fffffffu >> b;
                    const int gi = dense_index(sm, base + b, s);
                    const float4 pj = __ldg(&a.spos[gi]);
                    const float dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
                    const float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                    bool in = r2 <= c.h2_lo;
                    if (!in && r2 < c.h2_hi) in = in_range_exact(pi.x, pi.y, pi.z, pj.x, pj.y, pj.z, c.r2_max);
                    if (!in) continue;
                    gl[k] = (uint32_t)gi;
                    if (gi != t) {
                        dens_fast = __fadd_rn(dens_fast, poly6_fast(c, r2));
                        if (k < kSparseCount)
                            dens_exact = __fadd_rn(dens_exact, poly6_term(c, r2, pi.x, pi.y, pi.z, pj.x, pj.y, pj.z));
                    }
                    ++k;
                }
            }
            if (k < kMaxNeighbours && !complete) {
                walk = true;   // the spares went to band rejections before the 32nd neighbour: exact walk (below)
            } else {
                publish_density(a, c, t, (k <= kSparseCount) ? dens_exact : dens_fast, (uint8_t)k);
            }
        }
        __syncthreads();   // tables and masks are rewritten by the next cell
    }
    if (live && walk) {
        ForceAcc dummy;
        float dw = 0.f;
        const int wc = thread_walk<false>(a, g, c, t, pi, pi, 0.f, dw, dummy);
        publish_density(a, c, t, dw, (uint8_t)wc | CNT_WALK);
    }
}

__global__ void __launch_bounds__(DN_THREADS, 3)
density_dense_kernel(const SweepArgs a, const GridDesc g, const StepConsts c, uint32_t *__restrict__ dlist) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    DenseSmem &sm = *reinterpret_cast<DenseSmem *>(smem_raw);
    const int n = *a.n_dense;
    for (int i = blockIdx.x; i < n; i += gridDim.x) {
        dense_tile(a, g, c, dlist, a.dense_items[i] >> 3, sm);
        __syncthreads();
    }
}

// pressure + viscosity + integrate + collide for the dense tiles: lane = particle runs down its list of sorted indices and
// gathers the neighbours from the sorted arrays (the particles of a dense cell share their neighbourhood, so the gathers
// hit L1 / L2).  Same pair arithmetic and order as force_rows_kernel.
template <bool RECORD>
__global__ void __launch_bounds__(RB_THREADS)
force_gather_kernel(const SweepArgs a, const GridDesc g, const StepConsts c, const uint32_t *__restrict__ dlist) {
    const int n = *a.n_dense;
    const int j = threadIdx.x;
    for (int i = blockIdx.x; i < n; i += gridDim.x) {
        const int tile = a.dense_items[i] >> 3;
        const int t = tile * RB_THREADS + j;
        if (t >= a.n) continue;
        const uint32_t key = a.skeys[t];
        const uint32_t my_id = a.sids[t];
        const float4 pi = a.spos[t], vi = a.svel[t];
        const float rho = a.srho[t];
        ForceAcc f;
        if (key != (uint32_t)g.ncells) {
            int cx, cy, cz;
            decode_cell(g, key, cx, cy, cz);
            if (cx < g.own_lo || cx >= g.own_hi) continue;   // x-slab: ghost cell, its owner computes the forces
            const uint8_t cf = a.ncnt[t];
            if (cf & CNT_WALK) {
                float dens = 0.f;
                thread_walk<true>(a, g, c, t, pi, vi, pressure_coeff(c, rho), dens, f);
            } else {
                const uint4 *lg = reinterpret_cast<const uint4 *>(dlist + (size_t)t * kMaxNeighbours);
                const int cnt = cf;
                for (int q = 0; q * 4 < cnt; ++q) {
                    const uint4 e = __ldg(&lg[q]);
                    const uint32_t w[4] = {e.x, e.y, e.z, e.w};
                    float4 pj[4], vj[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int jx = (q * 4 + u < cnt) ? (int)w[u] : t;
                        pj[u] = __ldg(&a.spos[jx]);
                        vj[u] = __ldg(&a.svel[jx]);
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (q * 4 + u < cnt) f.pair(c, pi, vi, pi.w, pj[u], vj[u], (int)w[u] != t);
                }
            }
        }
        // dead particle: F = external force, rho = 0 (reference NaN semantics carry on)
        finish_particle<RECORD>(a, c, t, pi, vi, rho, f, my_id);
    }
}

}  // namespace sph
