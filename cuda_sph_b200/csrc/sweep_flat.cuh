// density_flat_kernel (sm_100a): neighbour lists + poly6 density, "column-block" organisation.
//
// Reference semantics (voxel_kernels.py:29-85,108-132): the neighbour list of a particle is the first 32 candidates (self
// included) with sqrt(r^2) <= INF_R when its <= 27 cells are walked dx outermost / dz innermost, every cell in ascending
// particle id; the density is the poly6 sum over that list (self skipped).
//
// What bounds this sweep is instruction issue, not memory (a capped particle tests ~215 candidates to find its 32), so
// the organisation below is about instructions per candidate test and idle lanes:
//   * column blocks.  The walk of a cell (cx, cy, cz) is B(cx-1) ++ B(cx) ++ B(cx+1) with B(x) = the 9 cells
//     (x, cy+dy, cz+dz), dy outer / dz inner.  For the x-consecutive cells of a tile the blocks are staged ONCE, in
//     column order, so the candidate sequence of EVERY cell is a contiguous window of the staged array (as many staged
//     candidates as the row staging of sweep_rows.cuh, but no per-segment loops: one flat loop per particle).
//   * staging: segments of >= 16 candidates by 1-D TMA bulk copies (cp.async.bulk -> UBLKCP, mbarrier completion), short
//     ones by plain 16-B loads; then an in-place 4x4 transpose per group of four candidates into packed-SoA form
//     (-x | y - y0 | z - z0 | (y-y0)^2 + (z-z0)^2, four candidates per 16-B vector).  y0 / z0 are chosen so that the
//     shifted coordinates are EXACT (Sterbenz), hence differences of shifted coordinates equal differences of the
//     originals bit for bit and the canonical density sum of sweep.cuh is unchanged.
//   * scan: lane = particle, all lanes of a warp step through their windows in lockstep (broadcast LDS.128).  Per pair
//     of candidates: one FADD2 (dx), one FADD2 + three FFMA2 evaluate
//         s = dx^2 + [(y'-py')^2 + (z'-pz')^2 expanded] - h^2 (1 + band)
//     and one funnel shift per candidate moves the sign bit of s into a 32-candidate hit mask (no compare, no store, no
//     branch per candidate).  s is a SUPERSET test: `band` covers the rounding of the expanded form (StepConsts::h2_sup).
//   * lists: a second pass runs down the set bits of the masks (first 32), recomputes r^2 the canonical way, applies
//     the reference's fp64 predicate inside the narrow rounding band (rare, out of line) and accumulates the density in
//     list order.  Lists leave as row slots of the force sweep's staging (force_rows_kernel, sweep_rows.cuh).
// Tiles this kernel cannot take (Q2 grids, tiles whose rows or blocks exceed shared memory) are appended to a list and
// processed by density_rows_fallback_kernel (the row-staged sweep of sweep_rows.cuh); single particles it cannot take
// (aliased keys Q5, windows longer than FL_MAXR rounds) use the one-thread walk.  All paths produce the same bits.
#pragma once
#include "sweep_rows.cuh"

namespace sph {

constexpr int FL_THREADS = RB_THREADS;   // one tile of the force sweep (its TilePlan defines the list slots)
constexpr int FL_CAP = 2048;             // staged candidates per tile
constexpr int FL_SLACK = 32;             // the scan reads up to one round past a window
constexpr int FL_MAXC = RB_MAXC;         // non-empty cells per tile
constexpr int FL_MAXB = 80;              // column blocks per tile
constexpr int FL_MAXR = 24;              // 32-candidate rounds per particle (longer windows: one-thread walk)
constexpr int FL_KEEP = 36;              // superset hits after which a lane stops scanning (32 + spares for rejections)
constexpr int FL_TMA_MIN = 16;           // candidates per segment from which the copy is a TMA bulk copy
constexpr int FL_CTAS = 4;
static_assert(FL_THREADS == 128, "density_flat_kernel is written for 128-particle tiles");

struct FlatSmem {
    // staged candidates; after the transpose group q (candidates 4q..4q+3) = { -x[4], y'[4], z'[4], w'[4] }
    float4 cand[FL_CAP + FL_SLACK];
    uint16_t slot[FL_CAP];               // staged position -> row slot of the force sweep
    union {
        struct {
            int seg_g[FL_MAXB * 9];           // first sorted index of segment (block, dy, dz)
            uint16_t seg_n[FL_MAXB * 9];      // its length
        } st;
        uint32_t masks[FL_MAXR * FL_THREADS]; // [round][thread] hit masks, first candidate of a round in bit 31
    } u;
    int blk_start[FL_MAXB + 1];          // staged position of block b (multiple of 4); [nblk] = end
    int blk_x[FL_MAXB], blk_cy[FL_MAXB], blk_cz[FL_MAXB];
    float blk_y0[FL_MAXB], blk_z0[FL_MAXB];
    uint8_t grp_blk[FL_CAP / 4];         // group of four -> block
    uint32_t ckey[FL_MAXC];
    int cell_fb[FL_MAXC], cell_nbw[FL_MAXC];   // window of a cell: first block, number of blocks
    int row_lo[9], row_base[10];
    int wsum[FL_THREADS / 32], wsum2[FL_THREADS / 32], wsum3[FL_THREADS / 32], wcount[FL_THREADS / 32];
    unsigned long long mbar;
};

// finished lanes keep stepping with their warp: everything they may touch lies inside the structure
static_assert(sizeof(FlatSmem) >= (size_t)(FL_CAP + FL_MAXR * 32 + 32) * 16, "scan overrun must stay inside FlatSmem");

struct FlatArgs {
    int *refused;      // tiles left to density_rows_fallback_kernel
    int *n_refused;    // zeroed by rows_plan_kernel
};

// Origin of the shifted y / z coordinates of the candidates around cell index c: for c >= 2 every coordinate of the
// cells c-1 .. c+1 lies within [y0 / 2, 2 y0] of y0 = fl(c * voxel), so y - y0 is exact; cells 0 and 1 keep y0 = 0.
__device__ __forceinline__ float flat_origin(int c, double voxel) { return c >= 2 ? (float)((double)c * voxel) : 0.f; }

// Exact (slow, rare) evaluation of one particle's masks: fp64 predicate inside the rounding band, first 32 accepted,
// list entries straight to global memory.  Returns the neighbour count, or -1 if the superset ran out before the
// sequence did (the caller walks).
__device__ __noinline__ int flat_exact_particle(const FlatSmem &sm, const SweepArgs &a, const StepConsts &c, int j, int t,
                                                int win, int nr, bool exhausted, float4 pi, float y0, float z0,
                                                int selfslot, float &dens_out) {
    const float *cf = reinterpret_cast<const float *>(sm.cand);
    uint16_t *gl = a.nlist + (size_t)t * 32;
    const float pys = pi.y - y0, pzs = pi.z - z0;
    int k = 0;
    float dens_fast = 0.f, dens_exact = 0.f;
    for (int r = 0; r < nr && k < kMaxNeighbours; ++r) {
        uint32_t m = sm.u.masks[r * FL_THREADS + j];
        while (m && k < kMaxNeighbours) {
            const int b = __clz(m);
            m &= 0x7fffffffu >> b;
            const int f = win + r * 32 + b;
            const int o = (f >> 2) * 16 + (f & 3);
            const float xs = -cf[o], ys = cf[o + 4], zs = cf[o + 8];
            const float dx = pi.x - xs, dy = pys - ys, dz = pzs - zs;
            const float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
            bool in = r2 <= c.h2_lo;
            // ys + y0 / zs + z0 are the original coordinates (the shift was exact)
            if (!in && r2 < c.h2_hi) in = in_range_exact(pi.x, pi.y, pi.z, xs, ys + y0, zs + z0, c.r2_max);
            if (!in) continue;
            const int slot = sm.slot[f];
            gl[k] = (uint16_t)slot;
            if (slot != selfslot) {
                dens_fast = __fadd_rn(dens_fast, poly6_fast(c, r2));
                if (k < kSparseCount)
                    dens_exact = __fadd_rn(dens_exact, poly6_term(c, r2, pi.x, pi.y, pi.z, xs, ys + y0, zs + z0));
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
    const bool live = key != (uint32_t)g.ncells;
    if (j == 0) {
        mbar_init(&sm.mbar, 1);
        fence_mbar_init();
    }
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
    bool refuse = !g.aligned || !tp_fits || ncell > FL_MAXC;   // CTA-uniform
    if (!refuse && first) sm.ckey[ci] = key;
    __syncthreads();

    // ---- column blocks: cell c needs columns cx-1 .. cx+1 of its (cy, cz) row; consecutive cells share them ----------
    int nbk = 0, reuse = 0, nf = 0, bcy = 0, bcz = 0, span = 0;
    if (!refuse && j < ncell) {
        int cx;
        decode_cell(g, sm.ckey[j], cx, bcy, bcz);
        const int xlo = max(0, g.xoff), xhi = min(g.tx, g.xoff + g.wk) - 1;   // columns of the domain and the local table
        const int fcol = max(cx - 1, xlo), lcol = min(cx + 1, xhi);
        int prev_last = INT_MIN;
        if (j > 0) {
            int qx, qy, qz;
            decode_cell(g, sm.ckey[j - 1], qx, qy, qz);
            if (qy == bcy && qz == bcz) prev_last = min(qx + 1, xhi);
        }
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
    if (!refuse && j < ncell) {
        const int pfx = bbase + binc - nbk;
        for (int i = 0; i < nbk; ++i) {
            sm.blk_x[pfx + i] = nf + i;
            sm.blk_cy[pfx + i] = bcy;
            sm.blk_cz[pfx + i] = bcz;
        }
        sm.cell_fb[j] = pfx - reuse;
        sm.cell_nbw[j] = span;
    }
    __syncthreads();

    // ---- segments: the 9 cell ranges of every block ------------------------------------------------------------------
    const int nseg = refuse ? 0 : nblk * 9;
    for (int it = j; it < nseg; it += FL_THREADS) {
        const int b = it / 9, s = it - b * 9;
        const int x = sm.blk_x[b], y = sm.blk_cy[b] + s / 3 - 1, z = sm.blk_cz[b] + s % 3 - 1;
        int2 r = make_int2(0, 0);
        if (y >= 0 && y < g.ty && z >= 0 && z < g.tz) {
            const long long cl = (long long)x - g.xoff + (long long)y * g.wn + (long long)z * g.wn * g.hn;
            if (cl >= 0 && cl < g.ncells) r = __ldg(&a.cell_range[cl]);
        }
        sm.u.st.seg_g[it] = r.x;
        sm.u.st.seg_n[it] = (uint16_t)min(max(r.y - r.x, 0), FL_CAP + 1);
    }
    __syncthreads();

    // ---- staged position of every block (padded to groups of four) ---------------------------------------------------
    int tot = 0, bulk = 0;
    if (j < nseg / 9) {
#pragma unroll
        for (int s = 0; s < 9; ++s) {
            const int n = sm.u.st.seg_n[j * 9 + s];
            tot += n;
            if (n >= FL_TMA_MIN) bulk += n;
        }
    }
    const int padded = (tot + 3) & ~3;
    int sinc = padded, binc2 = bulk;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(FULL, sinc, o);
        const int v = __shfl_up_sync(FULL, binc2, o);
        if (lane >= o) {
            sinc += u;
            binc2 += v;
        }
    }
    if (lane == 31) {
        sm.wsum2[warp] = sinc;
        sm.wsum3[warp] = binc2;
    }
    __syncthreads();
    int sbase = 0, staged = 0, nbulk = 0;
#pragma unroll
    for (int w = 0; w < FL_THREADS / 32; ++w) {
        if (w < warp) sbase += sm.wsum2[w];
        staged += sm.wsum2[w];
        nbulk += sm.wsum3[w];
    }
    refuse = refuse || staged > FL_CAP;
    if (refuse) {   // CTA-uniform; nothing was issued
        if (j == 0) fa.refused[atomicAdd(fa.n_refused, 1)] = tile;
        return;
    }
    if (j == 0 && nbulk > 0) mbar_expect_tx(&sm.mbar, (uint32_t)nbulk * 16u);
    if (j < nblk) {
        const int start = sbase + sinc - padded;
        sm.blk_start[j] = start;
        if (j == nblk - 1) sm.blk_start[nblk] = start + padded;
        sm.blk_y0[j] = flat_origin(sm.blk_cy[j], g.voxel[1]);
        sm.blk_z0[j] = flat_origin(sm.blk_cz[j], g.voxel[2]);
        for (int f = start + tot; f < start + padded; ++f) {   // pad candidates: far away, never a hit
            sm.cand[f] = make_float4(1e18f, 1e18f, 1e18f, 0.f);
            sm.slot[f] = 0xffffu;
        }
    }
    __syncthreads();   // the barrier is armed before any copy is issued

    // ---- issue the copies; slot and group tables ---------------------------------------------------------------------
    for (int it = j; it < nseg; it += FL_THREADS) {
        const int n = sm.u.st.seg_n[it];
        if (n == 0) continue;
        const int b = it / 9, s = it - b * 9;
        int f0 = sm.blk_start[b];
        for (int q = 0; q < s; ++q) f0 += sm.u.st.seg_n[b * 9 + q];
        const int gs = sm.u.st.seg_g[it];
        if (n >= FL_TMA_MIN) {
            bulk_g2s(&sm.cand[f0], &a.spos[gs], (uint32_t)n * 16u, &sm.mbar);
        } else {
            for (int q = 0; q < n; q += 4) {
                float4 v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (q + u < n) v[u] = __ldg(&a.spos[gs + q + u]);
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (q + u < n) sm.cand[f0 + q + u] = v[u];
            }
        }
        const int slot0 = sm.row_base[s] + (gs - sm.row_lo[s]);
        for (int q = 0; q < n; ++q) sm.slot[f0 + q] = (uint16_t)(slot0 + q);
        for (int q = f0 >> 2; q <= (f0 + n - 1) >> 2; ++q) sm.grp_blk[q] = (uint8_t)b;
    }
    __syncthreads();
    if (nbulk > 0) {
        while (!mbar_try_wait(&sm.mbar, 0u)) {
        }
    }

    // ---- in-place 4x4 transpose into packed-SoA form ------------------------------------------------------------------
    for (int q = j; q < (staged >> 2); q += FL_THREADS) {
        const int b = sm.grp_blk[q];
        const float y0 = sm.blk_y0[b], z0 = sm.blk_z0[b];
        const float4 c0 = sm.cand[4 * q], c1 = sm.cand[4 * q + 1], c2 = sm.cand[4 * q + 2], c3 = sm.cand[4 * q + 3];
        const float4 ys = make_float4(c0.y - y0, c1.y - y0, c2.y - y0, c3.y - y0);
        const float4 zs = make_float4(c0.z - z0, c1.z - z0, c2.z - z0, c3.z - z0);
        sm.cand[4 * q] = make_float4(-c0.x, -c1.x, -c2.x, -c3.x);
        sm.cand[4 * q + 1] = ys;
        sm.cand[4 * q + 2] = zs;
        sm.cand[4 * q + 3] = make_float4(fmaf(ys.x, ys.x, zs.x * zs.x), fmaf(ys.y, ys.y, zs.y * zs.y),
                                         fmaf(ys.z, ys.z, zs.z * zs.z), fmaf(ys.w, ys.w, zs.w * zs.w));
    }
    __syncthreads();   // staging tables are dead from here on: u.masks may be written

    // ---- this lane's particle ----------------------------------------------------------------------------------------
    int cx = 0, cy = 0, cz = 0;
    bool want = false, walk = false;
    if (live) {
        decode_cell(g, key, cx, cy, cz);
        want = !(cx < g.own_lo - 1 || cx > g.own_hi);   // x-slab: nobody needs the density of the outer ghost column
        walk = want && !own_cell_matches(g, pi, cx, cy, cz);
    }
    const bool scan = live && want && !walk;
    int win = 0, len = 0;
    if (scan) {
        const int fb = sm.cell_fb[ci];
        win = sm.blk_start[fb];
        len = sm.blk_start[fb + sm.cell_nbw[ci]] - win;
    }
    const float y0 = flat_origin(cy, g.voxel[1]), z0 = flat_origin(cz, g.voxel[2]);
    const float pys = pi.y - y0, pzs = pi.z - z0;
    // s = (px - x)^2 + w' + K + qy y' + qz z'  =  r^2 - h2_sup   (expanded in y and z)
    const float2 px2 = make_float2(pi.x, pi.x);
    const float qy = -2.f * pys, qz = -2.f * pzs;
    const float2 qy2 = make_float2(qy, qy), qz2 = make_float2(qz, qz);
    const float kk = fmaf(pzs, pzs, fmaf(pys, pys, -c.h2_sup));
    const float2 k2 = make_float2(kk, kk);

    // ---- scan: 32 candidates per round, lockstep over the warp --------------------------------------------------------
    int cs = 0, nr = 0, k0 = 0;
    bool act = len > 0;
    {
        const float4 *cq = sm.cand + win;
        uint32_t *mrow = sm.u.masks + j;
#pragma unroll 1
        for (int round = 0; round < FL_MAXR; ++round) {
            if (!__any_sync(FULL, act)) break;
            uint32_t m = 0;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const float4 X = cq[4 * u], Y = cq[4 * u + 1], Z = cq[4 * u + 2], W = cq[4 * u + 3];
                const float2 d0 = __fadd2_rn(px2, make_float2(X.x, X.y));
                const float2 d1 = __fadd2_rn(px2, make_float2(X.z, X.w));
                float2 s0 = __fadd2_rn(make_float2(W.x, W.y), k2);
                float2 s1 = __fadd2_rn(make_float2(W.z, W.w), k2);
                s0 = __ffma2_rn(d0, d0, s0);
                s1 = __ffma2_rn(d1, d1, s1);
                s0 = __ffma2_rn(qy2, make_float2(Y.x, Y.y), s0);
                s1 = __ffma2_rn(qy2, make_float2(Y.z, Y.w), s1);
                s0 = __ffma2_rn(qz2, make_float2(Z.x, Z.y), s0);
                s1 = __ffma2_rn(qz2, make_float2(Z.z, Z.w), s1);
                m = __funnelshift_l(__float_as_uint(s0.x), m, 1);
                m = __funnelshift_l(__float_as_uint(s0.y), m, 1);
                m = __funnelshift_l(__float_as_uint(s1.x), m, 1);
                m = __funnelshift_l(__float_as_uint(s1.y), m, 1);
            }
            const int nv = len - k0;   // candidates of this round that belong to the window
            const uint32_t vm = nv >= 32 ? 0xffffffffu : (nv > 0 ? 0xffffffffu << (32 - nv) : 0u);
            m &= vm;
            if (act) {
                mrow[round * FL_THREADS] = m;
                cs += __popc(m);
                nr = round + 1;
            }
            k0 += 32;
            cq += 32;
            act = act && cs < FL_KEEP && k0 < len;
        }
    }
    // a lane that is still active ran out of rounds: its window is longer than FL_MAXR * 32 candidates
    const bool exhausted = nr * 32 >= len;   // this lane's masks cover its whole window (the superset is complete)
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
        const float *cf = reinterpret_cast<const float *>(sm.cand);
        const int kn = min(cs, kMaxNeighbours);
        int r = 0;
        uint32_t m = nr > 0 ? sm.u.masks[j] : 0u;
        bool near = false;
#pragma unroll
        for (int i = 0; i < kMaxNeighbours; ++i) {
            if (i < kn) {
                while (m == 0u) {
                    ++r;
                    m = sm.u.masks[r * FL_THREADS + j];
                }
                const int b = __clz(m);
                m &= 0x7fffffffu >> b;
                const int f = win + r * 32 + b;
                const int o = (f >> 2) * 16 + (f & 3);
                const float dx = pi.x + cf[o], dy = pys - cf[o + 4], dz = pzs - cf[o + 8];
                const float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                slow = slow || r2 > c.h2_lo;
                const uint32_t slot = sm.slot[f];
                const bool other = slot != (uint32_t)selfslot;
                if (i < kSparseCount) near = near || (other && r2 > c.h2_near);
                dens = __fadd_rn(dens, other ? poly6_fast(c, r2) : 0.f);
                pk[i >> 1] |= slot << ((i & 1) * 16);
            }
        }
        count = kn;
        // a sparse particle with a neighbour near the cut-off sums poly6_term (fp64 r^2): out of line
        slow = slow || (kn <= kSparseCount && near);
        // fewer than 32 superset hits although the scan stopped early cannot happen (it stops on FL_KEEP >= 32 hits)
    }
    if (slow) {
        float d2 = 0.f;
        const int k = flat_exact_particle(sm, a, c, j, t, win, nr, exhausted, pi, y0, z0, selfslot, d2);
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
