// density_flat_kernel (sm_100a): neighbour lists + poly6 density, "column-block" organisation.
//
// Reference semantics (voxel_kernels.py:29-85,108-132): the neighbour list of a particle is the first 32 candidates (self
// included) with sqrt(r^2) <= INF_R when its <= 27 cells are walked dx outermost / dz innermost, every cell in ascending
// particle id; the density is the poly6 sum over that list (self skipped).
//
// What bounds this sweep is the shared-memory data pipe (78 % busy at 8 particles per cell: a broadcast LDS.128 still
// delivers 16 bytes to each of 32 lanes) and instruction issue (47 %), not HBM (7 %): a capped particle tests ~215
// candidates to find its 32.  So the organisation is about instructions and shared-memory bytes per candidate test:
//   * column blocks.  The walk of a cell (cx, cy, cz) is B(cx-1) ++ B(cx) ++ B(cx+1) with B(x) = the 9 cells
//     (x, cy+dy, cz+dz), dy outer / dz inner.  For the x-consecutive cells of a tile the blocks are staged ONCE, in
//     column order, so the candidate sequence of EVERY cell is a contiguous window of the staged arrays (as many staged
//     candidates as the row staging of sweep_rows.cuh, but one flat loop per particle instead of 27 segment loops).
//   * staging: one 16-byte descriptor per segment (first sorted index, length, staged position, first row slot), eight
//     lanes per segment copy the candidates straight from the sorted positions into three SoA arrays (x | y | z, 12 B
//     per candidate): short segments (<= 16 candidates: everything at <= 8 particles per cell) by 4-byte cp.async, so that
//     all copies of the tile are in flight before the first wait; the tail of longer segments by 16-byte loads through
//     registers (a 4-byte cp.async costs the L1 as many sectors as a 16-byte load).  The copies are gathers of short
//     runs into a transposed layout, which TMA bulk copies cannot produce (a first version staged AoS rows by
//     cp.async.bulk and transposed in place: the transposed groups cost 4-way bank conflicts in every later access).
//   * scan: lane = particle, all lanes of a warp step through their windows in lockstep (broadcast LDS.128, four
//     candidates per vector).  Per PAIR of candidates three FADD2 + three FFMA2 (packed f32x2) evaluate
//     s = r^2 - h^2 (1 + 1e-5) and one funnel shift per candidate moves the sign bit of s into a 32-candidate hit mask:
//     no compare, no store, no branch per candidate (the row-staged sweep spends 12.75 instructions per candidate, this
//     loop 5.1).  Non-zero masks are kept per lane in shared memory.
//   * lists: a second pass runs down the set bits of the masks (first 32), recomputes r^2 the canonical way, applies
//     the reference's fp64 predicate inside the rounding band (rare, out of line) and accumulates the density in
//     list order.  Lists leave as row slots of the force sweep's staging (force_rows_kernel, sweep_rows.cuh), packed in
//     registers and written as 4 x 16 B per particle.
// Tiles this kernel cannot take run through the row-staged sweep of sweep_rows.cuh: tiles whose rows exceed the force
// sweep's staging are split by rows_plan_kernel into 32-particle work items (density_rows_items_kernel on a second stream,
// side by side with this kernel); Q2 grids and tiles whose blocks exceed this kernel's staging are appended to a second
// list that density_rows_items_kernel runs afterwards.  Single particles it cannot take (aliased keys Q5, windows longer
// than FL_MAXR rounds) use the one-thread walk.  All paths produce the same bits.
#pragma once
#include "sweep_rows.cuh"

namespace sph {

constexpr int FL_THREADS = RB_THREADS;   // one tile of the force sweep (its TilePlan defines the list slots)
#ifndef SPH_FL_CAP
#define SPH_FL_CAP 2560
#define SPH_FL_MAXR 24
#define SPH_FL_CTAS 4
#endif
constexpr int FL_CAP = SPH_FL_CAP;       // staged candidates per tile
constexpr int FL_SLACK = 32;             // the scan reads up to one round past a window
constexpr int FL_MAXC = RB_MAXC;         // non-empty cells per tile
constexpr int FL_MAXB = 80;              // column blocks per tile
constexpr int FL_MAXR = SPH_FL_MAXR;              // 32-candidate rounds per particle (longer windows: one-thread walk)
constexpr int FL_KEEP = 34;              // superset hits after which a lane stops scanning (32 + spares for rejections)
constexpr int FL_CTAS = SPH_FL_CTAS;
static_assert(FL_THREADS == 128, "density_flat_kernel is written for 128-particle tiles");

struct FlatSmem {
    // staged candidates, SoA; a window starts at a multiple of four (blocks are padded with far-away candidates)
    float x[FL_CAP + FL_SLACK], y[FL_CAP + FL_SLACK], z[FL_CAP + FL_SLACK];
    uint16_t slot[FL_CAP];               // staged position -> row slot of the force sweep
    union {
        struct {                              // staging table, dead once the candidates are staged: segment (block, dy, dz)
            int4 seg[FL_MAXB * 9];            // = (first sorted index, length, staged position, first row slot)
        } st;
        // per thread: its NON-ZERO hit masks in scan order (first candidate of a round in bit 31); one spare row so the
        // second pass may read one word ahead
        uint32_t masks[(FL_MAXR + 1) * FL_THREADS];
    } u;
    uint8_t mround[(FL_MAXR + 1) * FL_THREADS];   // round of the mask word
    int blk_start[FL_MAXB + 1];          // staged position of block b (multiple of 4); [nblk] = end
    int blk_x[FL_MAXB], blk_cy[FL_MAXB], blk_cz[FL_MAXB];
    uint32_t ckey[FL_MAXC];
    int cell_xyz[FL_MAXC][3];
    int cell_win[FL_MAXC], cell_len[FL_MAXC];   // window of a cell: staged position and length
    int row_lo[9], row_base[10];
    int wsum[FL_THREADS / 32], wsum2[FL_THREADS / 32], wcount[FL_THREADS / 32];
};
static_assert(sizeof(FlatSmem) <= (228 / FL_CTAS - 1) * 1024, "FL_CTAS CTAs per SM");
// finished lanes keep stepping with their warp: everything they may read lies inside the arrays
static_assert((3 * (FL_CAP + FL_SLACK) + FL_MAXR * 32) * 4 <= (int)sizeof(FlatSmem), "scan overrun must stay inside FlatSmem");

struct FlatArgs {
    int *refused;      // whole tiles handed over to density_rows_items_kernel: tile * 8
    int *n_refused;    // zeroed by reorder_kernel
};

// Exact (slow, rare) evaluation of one particle's masks: fp64 predicate inside the rounding band, first 32 accepted,
// list entries straight to global memory.  Returns the neighbour count, or -1 if the superset ran out before the
// window did (the caller walks).
__device__ __noinline__ int flat_exact_particle(const FlatSmem &sm, const SweepArgs &a, const StepConsts &c, int j, int t,
                                                int win, int nz, bool exhausted, float4 pi, int selfslot,
                                                float &dens_out) {
    uint16_t *gl = a.nlist + (size_t)t * 32;
    int k = 0;
    float dens_fast = 0.f, dens_exact = 0.f;
    for (int w = 0; w < nz && k < kMaxNeighbours; ++w) {
        uint32_t m = sm.u.masks[w * FL_THREADS + j];
        const int base = win + 32 * (int)sm.mround[w * FL_THREADS + j];
        while (m && k < kMaxNeighbours) {
            const int b = __clz(m);
            m &= 0x7fffffffu >> b;
            const int f = base + b;
            const float xs = sm.x[f], ys = sm.y[f], zs = sm.z[f];
            const float dx = pi.x - xs, dy = pi.y - ys, dz = pi.z - zs;
            const float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
            bool in = r2 <= c.h2_lo;
            if (!in && r2 < c.h2_hi) in = in_range_exact(pi.x, pi.y, pi.z, xs, ys, zs, c.r2_max);
            if (!in) continue;
            const int slot = sm.slot[f];
            gl[k] = (uint16_t)slot;
            if (slot != selfslot) {
                dens_fast = __fadd_rn(dens_fast, poly6_fast(c, r2));
                if (k < kSparseCount) dens_exact = __fadd_rn(dens_exact, poly6_term(c, r2, pi.x, pi.y, pi.z, xs, ys, zs));
            }
            ++k;
        }
    }
    if (k < kMaxNeighbours && !exhausted) return -1;
    dens_out = (k <= kSparseCount) ? dens_exact : dens_fast;
    return k;
}

__global__ void __launch_bounds__(FL_THREADS, FL_CTAS)
density_flat_kernel(const SweepArgs a, const GridDesc g, const StepConsts c, const FlatArgs fa) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    FlatSmem &sm = *reinterpret_cast<FlatSmem *>(smem_raw);
    const int j = threadIdx.x, lane = j & 31, warp = j >> 5;
    const int tile = blockIdx.x;
    if (!g.aligned) {   // Q2 grid: every particle walks (row-staged sweep, whole tile)
        if (j == 0) fa.refused[atomicAdd(fa.n_refused, 1)] = tile * 8;
        return;
    }
    const int p0 = tile * FL_THREADS;
    const int t = p0 + j;
    const int nb = min(FL_THREADS, a.n - p0);
    const TilePlan &tp = a.plans[tile];

    // ---- prologue loads, all issued up front -------------------------------------------------------------------------
    const uint32_t key = (j < nb) ? a.skeys[t] : (uint32_t)g.ncells;
    const uint32_t prev_key = (j < nb && j > 0) ? a.skeys[t - 1] : 0xffffffffu;
    const float4 pi = (j < nb) ? a.spos[t] : make_float4(0.f, 0.f, 0.f, 0.f);
    const int tp_fits = tp.fits;
    int rlen = 0, rlo = 0;
    if (j < 9) {
        rlen = tp.row_len[j];
        rlo = tp.row_lo[j];
    }
    {   // streaming inputs of the tile that runs PF_AHEAD CTAs later -> L2
        const int pp0 = (tile + PF_AHEAD) * FL_THREADS;
        if (pp0 + FL_THREADS <= a.n && j >= 32 && j < 32 + 21) {
            if (j == 32) prefetch_l2(&a.plans[tile + PF_AHEAD]);
            else if (j < 32 + 5) prefetch_l2(reinterpret_cast<const char *>(a.skeys + pp0) + (j - 33) * 128);
            else prefetch_l2(reinterpret_cast<const char *>(a.spos + pp0) + (j - 37) * 128);
        }
    }
    const bool live = key != (uint32_t)g.ncells;
    if (j < nb && !live) {   // dead particle: no neighbours
        a.srho[t] = 0.f;
        a.ncnt[t] = 0;
    }
    // ---- the tile's non-empty cells ----------------------------------------------------------------------------------
    const bool first = live && (j == 0 || prev_key != key);
    const unsigned bal = __ballot_sync(FULL, first);
    if (lane == 0) sm.wcount[warp] = __popc(bal);
    if (warp == 0) {   // row slots of the force sweep's staging (same prefix as rows_issue_planned)
        int inc = rlen;
#pragma unroll
        for (int o = 1; o < 16; o <<= 1) {
            const int u = __shfl_up_sync(FULL, inc, o);
            if (lane >= o) inc += u;
        }
        if (lane < 9) {
            sm.row_base[lane] = inc - rlen;
            sm.row_lo[lane] = rlo;
        }
    }
    __syncthreads();
    int coff = 0, ncell = 0;
#pragma unroll
    for (int w = 0; w < FL_THREADS / 32; ++w) {
        const int n = sm.wcount[w];
        if (w < warp) coff += n;
        ncell += n;
    }
    if (ncell == 0) return;   // nothing alive in this tile
    const int ci = coff + __popc(bal & lanemask_le_()) - 1;
    if (!tp_fits) return;   // rows do not fit the force sweep's staging either: its passes are work items (rows_plan_kernel)
    bool refuse = ncell > FL_MAXC;   // CTA-uniform (cannot happen while FL_MAXC == RB_MAXC: the plan fits)
    int cx = 0, cy = 0, cz = 0;
    if (live) decode_cell(g, key, cx, cy, cz);
    if (!refuse && first) {
        sm.ckey[ci] = key;
        sm.cell_xyz[ci][0] = cx;
        sm.cell_xyz[ci][1] = cy;
        sm.cell_xyz[ci][2] = cz;
    }
    __syncthreads();

    // ---- column blocks: cell c needs columns cx-1 .. cx+1 of its (cy, cz) row; consecutive cells share them ----------
    int nbk = 0, reuse = 0, nf = 0, bcy = 0, bcz = 0, span = 0;
    if (!refuse && j < ncell) {
        const int bx = sm.cell_xyz[j][0];
        bcy = sm.cell_xyz[j][1];
        bcz = sm.cell_xyz[j][2];
        const int xlo = max(0, g.xoff), xhi = min(g.tx, g.xoff + g.wk) - 1;   // columns of the domain and the local table
        const int fcol = max(bx - 1, xlo), lcol = min(bx + 1, xhi);
        int prev_last = INT_MIN;
        if (j > 0 && sm.cell_xyz[j - 1][1] == bcy && sm.cell_xyz[j - 1][2] == bcz)
            prev_last = min(sm.cell_xyz[j - 1][0] + 1, xhi);
        nf = max(fcol, prev_last + 1);
        nbk = max(lcol - nf + 1, 0);
        reuse = (prev_last >= fcol) ? prev_last - fcol + 1 : 0;
        span = max(lcol - fcol + 1, 0);
    }
    int binc = nbk;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(FULL, binc, o);
        if (lane >= o) binc += u;
    }
    if (lane == 31) sm.wsum[warp] = binc;
    __syncthreads();
    int bbase = 0, nblk = 0;
#pragma unroll
    for (int w = 0; w < FL_THREADS / 32; ++w) {
        const int n = sm.wsum[w];
        if (w < warp) bbase += n;
        nblk += n;
    }
    refuse = refuse || nblk > FL_MAXB;
    int my_fb = 0;
    if (!refuse && j < ncell) {
        const int pfx = bbase + binc - nbk;
        for (int i = 0; i < nbk; ++i) {
            sm.blk_x[pfx + i] = nf + i;
            sm.blk_cy[pfx + i] = bcy;
            sm.blk_cz[pfx + i] = bcz;
        }
        my_fb = pfx - reuse;
    }
    __syncthreads();

    // ---- segments: the 9 cell ranges of every block; the range loads of all trips are issued together ----------------
    const int nseg = refuse ? 0 : nblk * 9;
    {
        constexpr int SEG_TRIPS = (FL_MAXB * 9 + FL_THREADS - 1) / FL_THREADS;
        int2 rr[SEG_TRIPS];
#pragma unroll
        for (int u = 0; u < SEG_TRIPS; ++u) {
            const int it = j + u * FL_THREADS;
            rr[u] = make_int2(0, 0);
            if (it < nseg) {
                const int b = it / 9, s = it - b * 9;
                const int x = sm.blk_x[b], y = sm.blk_cy[b] + s / 3 - 1, z = sm.blk_cz[b] + s % 3 - 1;
                if (y >= 0 && y < g.ty && z >= 0 && z < g.tz) {
                    const long long cl = (long long)x - g.xoff + (long long)y * g.wn + (long long)z * g.wn * g.hn;
                    if (cl >= 0 && cl < g.ncells) rr[u] = __ldg(&a.cell_range[cl]);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < SEG_TRIPS; ++u) {
            const int it = j + u * FL_THREADS;
            if (it < nseg) {
                sm.u.st.seg[it].x = rr[u].x;
                sm.u.st.seg[it].y = min(max(rr[u].y - rr[u].x, 0), FL_CAP + 1);
            }
        }
    }
    __syncthreads();

    // ---- staged position of every block (padded to groups of four) and of its segments -------------------------------
    int tot = 0;
    if (j < nseg / 9) {
#pragma unroll
        for (int s = 0; s < 9; ++s) tot += sm.u.st.seg[j * 9 + s].y;
    }
    const int padded = (tot + 3) & ~3;
    int sinc = padded;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(FULL, sinc, o);
        if (lane >= o) sinc += u;
    }
    if (lane == 31) sm.wsum2[warp] = sinc;
    __syncthreads();
    int sbase = 0, staged = 0;
#pragma unroll
    for (int w = 0; w < FL_THREADS / 32; ++w) {
        if (w < warp) sbase += sm.wsum2[w];
        staged += sm.wsum2[w];
    }
    refuse = refuse || staged > FL_CAP;
    if (refuse) {   // more blocks / staged candidates than this kernel holds: the row-staged sweep takes the whole tile
        if (j == 0) fa.refused[atomicAdd(fa.n_refused, 1)] = tile * 8;
        return;
    }
    if (j < nblk) {
        const int start = sbase + sinc - padded;
        sm.blk_start[j] = start;
        if (j == nblk - 1) sm.blk_start[nblk] = start + padded;
        int f = start;
#pragma unroll
        for (int s = 0; s < 9; ++s) {   // complete the segment descriptors: staged position, first row slot
            int4 &d = sm.u.st.seg[j * 9 + s];
            d.z = f;
            d.w = sm.row_base[s] + (d.x - sm.row_lo[s]);
            f += d.y;
        }
        for (; f < start + padded; ++f) {   // pad candidates: far away, never a hit
            sm.x[f] = 1e18f;
            sm.y[f] = 1e18f;
            sm.z[f] = 1e18f;
            sm.slot[f] = 0xffffu;
        }
    }
    __syncthreads();

    // ---- stage: eight lanes per segment, every candidate by three 4-byte cp.async (x | y | z land transposed without a
    //      register round trip, so ALL copies of the tile are in flight before the first one is waited for; the first
    //      version loaded float4s into registers and stored them, which exposed one global round trip per trip of the loop)
    {
        const int sub = lane & 7;
        auto stage_segment = [&](const int4 d) {   // (first sorted index, length, staged position, first row slot)
            // short segments (<= 16 candidates: everything at <= 8 particles per cell) entirely by cp.async, of long ones
            // the first eight
            const int na = d.y <= 16 ? d.y : 8;
            for (int q = sub; q < na; q += 8) {
                const float *src = reinterpret_cast<const float *>(a.spos + d.x + q);
                cp_async4(&sm.x[d.z + q], src);
                cp_async4(&sm.y[d.z + q], src + 1);
                cp_async4(&sm.z[d.z + q], src + 2);
                sm.slot[d.z + q] = (uint16_t)(d.w + q);
            }
            // the tail of long segments: 16-byte loads, three in flight per lane (a 4-byte cp.async costs the L1 as many
            // sectors as a 16-byte load, so a dense cell is cheaper through registers)
            for (int q0 = sub + na; q0 < d.y; q0 += 24) {
                float4 w[3];
#pragma unroll
                for (int k = 0; k < 3; ++k)
                    if (q0 + 8 * k < d.y) w[k] = __ldg(&a.spos[d.x + q0 + 8 * k]);
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const int q = q0 + 8 * k;
                    if (q < d.y) {
                        sm.x[d.z + q] = w[k].x;
                        sm.y[d.z + q] = w[k].y;
                        sm.z[d.z + q] = w[k].z;
                        sm.slot[d.z + q] = (uint16_t)(d.w + q);
                    }
                }
            }
        };
        for (int it = j >> 3; it < nseg; it += 2 * (FL_THREADS / 8)) {   // two descriptors per trip: one LDS.128 each
            const int it1 = it + FL_THREADS / 8;
            const int4 d0 = sm.u.st.seg[it];
            const int4 d1 = (it1 < nseg) ? sm.u.st.seg[it1] : make_int4(0, 0, 0, 0);
            stage_segment(d0);
            stage_segment(d1);
        }
        cp_async_wait_all();
    }
    // windows of the cells (block tables are final since the previous barrier)
    if (j < ncell) {
        const int fcolb = my_fb;
        sm.cell_win[j] = sm.blk_start[fcolb];
        sm.cell_len[j] = sm.blk_start[fcolb + span] - sm.blk_start[fcolb];
    }
    __syncthreads();   // staging tables are dead from here on: u.masks may be written

    // ---- this lane's particle ----------------------------------------------------------------------------------------
    bool want = false, walk = false;
    if (live) {
        want = !(cx < g.own_lo - 1 || cx > g.own_hi);   // x-slab: nobody needs the density of the outer ghost column
        walk = want && !own_cell_matches(g, pi, cx, cy, cz);
    }
    const bool scan = live && want && !walk;
    int win = 0, len = 0;
    if (scan) {
        win = sm.cell_win[ci];
        len = sm.cell_len[ci];
    }

    // ---- scan: 32 candidates per round, lockstep over the warp --------------------------------------------------------
    // s = (px - x)^2 + (py - y)^2 + (pz - z)^2 - h2_hi, sign bit = inside the superset r^2 < h^2 (1 + 1e-5)
    int cs = 0, nz = 0, nrs = 0;
    bool act = len > 0;
    {
        const float2 npx = make_float2(-pi.x, -pi.x), npy = make_float2(-pi.y, -pi.y), npz = make_float2(-pi.z, -pi.z);
        const float2 lim2 = make_float2(-c.h2_hi, -c.h2_hi);
        const float4 *xq = reinterpret_cast<const float4 *>(sm.x + win);
        const float4 *yq = reinterpret_cast<const float4 *>(sm.y + win);
        const float4 *zq = reinterpret_cast<const float4 *>(sm.z + win);
        uint32_t *mrow = sm.u.masks + j;
        uint8_t *rrow = sm.mround + j;
        int k0 = 0;
#pragma unroll 1
        for (int round = 0; round < FL_MAXR; ++round) {
            if (!__any_sync(FULL, act)) break;
            uint32_t m = 0;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const float4 X = xq[u], Y = yq[u], Z = zq[u];
                // x - px (the sign does not matter: the difference is squared)
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
            const int nv = len - k0;   // candidates of this round that belong to the window
            const uint32_t vm = nv >= 32 ? 0xffffffffu : (nv > 0 ? 0xffffffffu << (32 - nv) : 0u);
            m &= vm;
            if (act) {
                nrs = round + 1;
                if (m) {
                    mrow[nz * FL_THREADS] = m;
                    rrow[nz * FL_THREADS] = (uint8_t)round;
                    cs += __popc(m);
                    ++nz;
                }
            }
            k0 += 32;
            xq += 8;
            yq += 8;
            zq += 8;
            act = act && cs < FL_KEEP && k0 < len;
        }
    }
    // a lane that is still active ran out of rounds: its window is longer than FL_MAXR * 32 candidates
    const bool exhausted = nrs * 32 >= len;   // this lane's masks cover its whole window (the superset is complete)
    bool need_walk = walk || (scan && act);

    // ---- lists + density: run down the set bits (first 32), canonical r^2, poly6 in list order ------------------------
    const int selfslot = sm.row_base[4] + (t - sm.row_lo[4]);   // row 4 = (dy, dz) = (0, 0): the tile's own cells
    uint32_t pk[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) pk[i] = 0u;
    float dens = 0.f;
    int count = 0;
    bool slow = false;
    if (scan && !act) {
        const int kn = min(cs, kMaxNeighbours);
        uint32_t m = sm.u.masks[j];                       // kn > 0 implies a stored word
        int base = win + 32 * (int)sm.mround[j];
        const uint32_t *mnext = sm.u.masks + j + FL_THREADS;
        const uint8_t *rnext = sm.mround + j + FL_THREADS;
        unsigned band = 0u, near = 0u;
#pragma unroll
        for (int i = 0; i < kMaxNeighbours; ++i) {
            if (i < kn) {
                const int b = __clz(m);
                const int f = base + b;
                m &= 0x7fffffffu >> b;
                if (m == 0u) {   // stored words are non-zero: one reload, never a loop (reads at most the spare row)
                    m = *mnext;
                    base = win + 32 * (int)*rnext;
                    mnext += FL_THREADS;
                    rnext += FL_THREADS;
                }
                const float dx = pi.x - sm.x[f], dy = pi.y - sm.y[f], dz = pi.z - sm.z[f];
                const float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                const uint32_t slot = sm.slot[f];
                const bool other = slot != (uint32_t)selfslot;
                band |= (r2 > c.h2_lo) ? 1u : 0u;
                if (i < kSparseCount) near |= (other && r2 > c.h2_near) ? 1u : 0u;
                dens = __fadd_rn(dens, other ? poly6_fast(c, r2) : 0.f);
                pk[i >> 1] |= slot << ((i & 1) * 16);
            }
        }
        count = kn;
        // inside the rounding band the fp64 predicate decides; a sparse particle with a neighbour near the cut-off sums
        // poly6_term (fp64 r^2): both out of line
        slow = band != 0u || (kn <= kSparseCount && near != 0u);
    }
    if (slow) {
        float d2 = 0.f;
        const int k = flat_exact_particle(sm, a, c, j, t, win, nz, exhausted, pi, selfslot, d2);
        if (k < 0) {
            need_walk = true;
        } else {
            count = k;
            dens = d2;
        }
    }
    if (need_walk) {   // aliased key (quirk Q5), over-long window or exhausted spares: exact one-thread walk
        ForceAcc dummy;
        float dw = 0.f;
        const int wc = thread_walk<false>(a, g, c, t, pi, pi, 0.f, dw, dummy);
        publish_density(a, c, t, dw, (uint8_t)wc | CNT_WALK);
    } else if (scan) {
        publish_density(a, c, t, dens, (uint8_t)count);
        if (!slow) {   // the exact path wrote its entries itself
            uint4 *gl = reinterpret_cast<uint4 *>(a.nlist + (size_t)t * 32);
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (q * 8 < count) gl[q] = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
        }
    }
}

}  // namespace sph
