// Neighbour sweeps, "row-staged" organisation (sm_100a): density_rows_kernel and force_rows_kernel.
//
// Reference semantics (voxel_kernels.py:29-85): a particle's neighbour list is the first 32 candidates (self included)
// that pass sqrt(r^2) <= INF_R when the <= 27 cells around it are walked dx outermost / dz innermost and every cell in
// ascending particle id; all sums run over that list in order.  The sorted arrays are cell-contiguous with ascending id
// inside a cell and the cell key is x-fastest, so for a group of consecutive sorted particles the candidates of ALL
// their cells lie in 9 contiguous ranges of the sorted arrays ("rows": one per (dy, dz), covering the x-neighbours).
//
// One CTA = 128 consecutive sorted particles:
//   setup   find the CTA's non-empty cells, fetch their 27 cell ranges, take per-row min/max -> 9 row ranges; warp 0
//           issues 9 (density) / 18 (force) 1-D TMA bulk copies (cp.async.bulk -> UBLKCP) that land the rows in shared
//           memory and signal an mbarrier; the per-cell segment tables are turned into row slots meanwhile.
//   density lane = particle.  Each lane walks the 27 segments of its cell in reference order straight over the staged
//           rows (LDS.128 per candidate, software-prefetched) and keeps the first 32 candidates of the fp32 superset
//           r2 < h^2 (1 + 1e-5); a second, short pass over the kept candidates applies the fp64 predicate inside the
//           rounding band (so neighbour lists are bit-identical to the fp64 reference) and accumulates the poly6
//           density in list order.  Consecutive sorted particles share cells, so the lanes of a warp read the same
//           candidate most of the time (shared-memory broadcast) and all 32 lanes are busy at any particles-per-cell.
//           Outputs: rho, neighbour count, the list of row SLOTS, and the per-particle pair factors p/rho^2 and
//           LAP_W_CONST/rho (into the .w lanes of the sorted position / velocity, picked up by the force rows).
//   force   same setup (positions + velocities), then lane = particle runs down its slot list: 2 LDS.128 + ~40 FP
//           instructions per pair, no reductions, list order == the reference's summation order; fp64 integrate +
//           collide epilogue and scatter to the id-ordered master arrays (finish_particle, sweep.cuh).
// Both kernels derive the row layout from (sorted keys, cell table) with the same code, so a slot means the same thing
// in both.  A CTA whose rows do not fit shared memory falls back to 32-particle passes (one per warp); a pass that still does
// not fit, particles whose own cell differs from their sort cell (aliased keys, reference quirk Q5), grids whose
// trunc and ceil dims differ (Q2) and particles with a rejected band candidate take the plain one-thread walk of
// sweep.cuh over global memory (exact, slow, rare).
#pragma once
#include "sweep.cuh"

namespace sph {

#ifndef SPH_RB_THREADS
// measured at 2^20 dam-break (25/cell) / 2^22 box (8/cell), density + force ms:  64: 0.41+0.20 / 1.58+0.76;
// 128: 0.42+0.20 / 1.40+0.67;  256: 0.49+0.21 / 1.48+0.72
#define SPH_RB_THREADS 128
#endif
constexpr int RB_THREADS = SPH_RB_THREADS;   // threads == particles per CTA (a tile)
constexpr int RB_WARPS = RB_THREADS / 32;
#ifndef SPH_RB_CAP
// 2240: the largest staging that keeps 3 force CTAs and 4 row-density CTAs per SM (static_asserts below); at 2048, 3.4 % of
// the dam-break tiles (25 per cell) did not fit and ran as 32-particle passes: step 0.600 -> 0.563 ms at 2^20 particles
#define SPH_RB_CAP (RB_THREADS == 128 ? 2240 : RB_THREADS == 64 ? 1536 : 3328)
#endif
constexpr int RB_CAP = SPH_RB_CAP;   // row slots per CTA pass (candidates staged in shared memory)
constexpr int RB_MAXC = RB_THREADS / 2;                   // non-empty cells per CTA pass
constexpr int RB_DENSITY_CTAS = RB_THREADS == 128 ? 4 : RB_THREADS == 64 ? 6 : 2;   // CTAs per SM the kernels are sized for
#ifndef SPH_RB_FORCE_CTAS
#define SPH_RB_FORCE_CTAS (RB_THREADS == 128 ? 3 : RB_THREADS == 64 ? 4 : 2)
#endif
constexpr int RB_FORCE_CTAS = SPH_RB_FORCE_CTAS;
#ifndef SPH_PF_AHEAD
#define SPH_PF_AHEAD (148 * 4)
#endif
constexpr int PF_AHEAD = SPH_PF_AHEAD;   // L2 prefetch distance of the sweeps, in tiles (about one wave of resident CTAs)
constexpr int RB_KEEP = 33;       // superset candidates kept per particle (32 + spares for rejected band candidates)
constexpr int RB_LSTRIDE = 38;    // uint16 per list row (19 words: conflict-free rows; >= RB_KEEP + 1)

// ---- mbarrier + 1-D TMA bulk copy -----------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(void *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(void *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, void *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// L2 prefetch of a line another CTA will read soon (streaming inputs are cold in L2: the plans, lists and counts were
// written before the 2 GB of lists pushed them out).  Out-of-range addresses are the caller's business.
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void mbar_inval(void *bar) {
    asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(void *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}

__device__ __forceinline__ uint32_t lanemask_le_() {
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_le;" : "=r"(m));
    return m;
}

// ---- per-CTA plan (shared memory) ---------------------------------------------------------------------------------------
struct RowPlan {
    unsigned long long mbar;
    int row_lo[9], row_hi[9];   // sorted-index range of each row (min / max over the CTA's cells)
    int row_base[10];           // first slot of each row; [9] = slots used
    int wcount[RB_WARPS];
    int ncell;                  // non-empty cells of the pass
    int fits;
    uint32_t ckey[RB_MAXC];     // key of local cell ci
};

struct DensityRowsSmem {
    float4 rows[RB_CAP + 8];                 // [RB_CAP..] = far-away sentinel candidates (prefetch may touch them)
    // per local cell, its 27 segments in walk order: setup writes the sorted start (int), then slot | count << 16
    uint32_t seg[RB_MAXC * 27];
    uint16_t seg_cnt[RB_MAXC * 27];
    uint16_t list[RB_THREADS * RB_LSTRIDE];
    RowPlan plan;
};

struct ForceRowsSmem {
    float4 rpos[RB_CAP + 2];   // (x, y, z, p/rho^2)
    float4 rvel[RB_CAP + 2];   // (vx, vy, vz, LAP_W_CONST/rho)
    RowPlan plan;
};

// Neighbour cell `s` (0..26, dx outermost, dz innermost: voxel_kernels.py:46-48) of cell (cx, cy, cz): range of the
// sorted arrays, (0, 0) if the cell is outside the domain / the local table.
// 227 KB of shared memory per SM, 1 KB of it reserved per resident CTA
static_assert((sizeof(ForceRowsSmem) + 1024) * RB_FORCE_CTAS <= 227 * 1024, "RB_FORCE_CTAS force CTAs per SM");
static_assert((sizeof(DensityRowsSmem) + 1024) * RB_DENSITY_CTAS <= 227 * 1024, "RB_DENSITY_CTAS density CTAs per SM");
__device__ __forceinline__ int2 neighbour_range(const SweepArgs &a, const GridDesc &g, int s, int cx, int cy, int cz) {
    const int x = cx + s / 9 - 1, y = cy + (s / 3) % 3 - 1, z = cz + s % 3 - 1;
    if (x >= 0 && x < g.tx && y >= 0 && y < g.ty && z >= 0 && z < g.tz) {
        const long long cl = (long long)x - g.xoff + (long long)y * g.wn + (long long)z * g.wn * g.hn;
        if (cl >= 0 && cl < g.ncells) return __ldg(&a.cell_range[cl]);
    }
    return make_int2(0, 0);
}

// Setup of one pass over the CTA-local particles [j0, j1).  Returns false (CTA-uniform) if the pass does not fit; then
// nothing was issued.  On success the TMA copies are in flight on plan.mbar (wait with the caller's parity).
//   key, live : this thread's particle (j = threadIdx.x, sorted index t = p0 + j)
//   ci        : out, local cell index of this thread's particle (valid if live and in range)
template <bool DENSITY>
__device__ __forceinline__ bool rows_setup(const SweepArgs &a, const GridDesc &g, RowPlan &plan, float4 *rows_a,
                                           float4 *rows_b, DensityRowsSmem *ds, int j0, int j1, int t, uint32_t key,
                                           bool live, int &ci) {
    const int j = threadIdx.x, lane = j & 31, warp = j >> 5;
    const bool mine = live && j >= j0 && j < j1;
    bool first = false;
    if (mine) first = (j == j0) || (a.skeys[t - 1] != key);
    const unsigned bal = __ballot_sync(FULL, first);
    if (lane == 0) plan.wcount[warp] = __popc(bal);
    if (j < 9) {
        plan.row_lo[j] = INT_MAX;
        plan.row_hi[j] = 0;
    }
    __syncthreads();
    int off = 0, total = 0;
#pragma unroll
    for (int w = 0; w < RB_WARPS; ++w) {
        const int c = plan.wcount[w];
        if (w < warp) off += c;
        total += c;
    }
    ci = off + __popc(bal & lanemask_le_()) - 1;
    if (total > RB_MAXC) {
        __syncthreads();   // wcount is rewritten by the next pass
        return false;
    }
    if (first) plan.ckey[ci] = key;
    __syncthreads();

    // ---- per cell: the 27 neighbour ranges, row min / max, virtual-list offsets ----
    for (int c = warp; c < total; c += RB_WARPS) {
        int cx, cy, cz;
        decode_cell(g, plan.ckey[c], cx, cy, cz);
        int2 r = make_int2(0, 0);
        if (lane < 27) r = neighbour_range(a, g, lane, cx, cy, cz);
        const int cnt = r.y - r.x;
        if (cnt > 0) {
            atomicMin(&plan.row_lo[lane % 9], r.x);
            atomicMax(&plan.row_hi[lane % 9], r.y);
        }
        if (DENSITY && lane < 27) {
            ds->seg[c * 27 + lane] = (uint32_t)r.x;
            ds->seg_cnt[c * 27 + lane] = (uint16_t)min(cnt, 65535);
        }
    }
    __syncthreads();

    // ---- row bases, capacity check, TMA issue (warp 0) ----
    if (warp == 0) {
        const int len = (lane < 9) ? max(plan.row_hi[lane] - plan.row_lo[lane], 0) : 0;
        int inc = len;
#pragma unroll
        for (int o = 1; o < 16; o <<= 1) {
            const int u = __shfl_up_sync(FULL, inc, o);
            if (lane >= o) inc += u;
        }
        const int slots = __shfl_sync(FULL, inc, 8);
        if (lane < 9) plan.row_base[lane] = inc - len;
        const bool fits = slots <= RB_CAP;
        if (lane == 0) {
            plan.row_base[9] = slots;
            plan.ncell = total;
            plan.fits = fits ? 1 : 0;
        }
        // lane 0 arms the barrier with the byte count of all rows before any copy is issued
        if (fits && lane == 0) {
            fence_proxy_async();   // earlier generic-proxy reads of the rows (previous pass) vs. the async-proxy writes
            mbar_expect_tx(&plan.mbar, (uint32_t)slots * (DENSITY ? 16u : 32u));
        }
        __syncwarp();
        if (fits && lane < 9 && len > 0) {
            const int base = inc - len, lo = plan.row_lo[lane];
            bulk_g2s(&rows_a[base], &a.spos[lo], (uint32_t)len * 16u, &plan.mbar);
            if (!DENSITY) bulk_g2s(&rows_b[base], &a.svel[lo], (uint32_t)len * 16u, &plan.mbar);
        }
    }
    __syncthreads();
    if (!plan.fits) return false;

    // ---- per cell: slot of the own segment; density: sorted starts -> row slots ----
    for (int c = warp; c < total; c += RB_WARPS) {
        if (DENSITY) {
            if (lane < 27) {
                const int start = (int)ds->seg[c * 27 + lane];
                const int cnt = ds->seg_cnt[c * 27 + lane];
                const int slot0 = plan.row_base[lane % 9] + (start - plan.row_lo[lane % 9]);
                ds->seg[c * 27 + lane] = cnt ? ((uint32_t)slot0 | ((uint32_t)cnt << 16)) : 0u;
            }
        }
    }
    __syncthreads();
    return true;
}

// ---- row plans, computed once per step for both sweeps ---------------------------------------------------------------------
// One warp per tile (= the 128 sorted particles of one sweep CTA): the 9 row ranges (min / max over the 27 neighbour
// ranges of the tile's non-empty cells).  The sweeps then start with one 80-byte load and the TMA copies; without the
// plan every CTA of both kernels repeats this chain of dependent loads before its first copy can be issued.
struct TilePlan {
    int row_lo[9];
    int row_len[9];
    int slots;   // sum of row_len
    int fits;    // slots <= RB_CAP and at most RB_MAXC cells: the whole tile is one pass
};

constexpr int RP_WARPS = 8;
constexpr int OWN_TILE = -2;         // only_pass value: the CTA's own tile (skipped if its rows do not fit: work items)
constexpr int ITEM_CTAS = 148 * 2;   // grid of the work-item kernels

__global__ void __launch_bounds__(RP_WARPS * 32)
rows_plan_kernel(const SweepArgs a, const GridDesc g, TilePlan *__restrict__ plans, int ntiles,
                 int *__restrict__ n_items, int *__restrict__ items, int *__restrict__ n_dense,
                 int *__restrict__ dense_items, int dense_max_cells, int dense_min_slots) {
    const int lane = threadIdx.x & 31;
    const int tile = blockIdx.x * RP_WARPS + (threadIdx.x >> 5);
    if (tile >= ntiles) return;
    const int p0 = tile * RB_THREADS;
    // The tile's cells come in RUNS of ascending x inside one grid row (cy, cz) -- one run, two where the tile wraps into
    // the next grid row.  Row (dy, dz) of a run with cells xa .. xb is the hull of the non-empty cells
    // (xa - 1 .. xb + 1, cy + dy, cz + dz): 32 cell ranges per trip, one lane each, first / last non-empty by ballot.
    // (The first version fetched the 27 neighbour ranges of every cell: 432 loads and index computations for 16 cells where
    // this one needs ~170, and 2030 warp instructions per tile.)
    int lo[9], hi[9];   // warp-uniform
#pragma unroll
    for (int r = 0; r < 9; ++r) {
        lo[r] = INT_MAX;
        hi[r] = 0;
    }
    auto flush = [&](int xa, int xb, int cy, int cz) {
#pragma unroll
        for (int r = 0; r < 9; ++r) {
            const int y = cy + r / 3 - 1, z = cz + r % 3 - 1;
            if (y < 0 || y >= g.ty || z < 0 || z >= g.tz) continue;
            const long long rowbase = (long long)y * g.wn + (long long)z * g.wn * g.hn - g.xoff;
            for (int x0 = xa - 1; x0 <= xb + 1; x0 += 32) {
                const int x = x0 + lane;
                int2 rg = make_int2(0, 0);
                if (x <= xb + 1 && x >= 0 && x < g.tx) {
                    const long long cl = rowbase + x;
                    if (cl >= 0 && cl < g.ncells) rg = __ldg(&a.cell_range[cl]);
                }
                const unsigned ne = __ballot_sync(FULL, rg.y > rg.x);
                if (ne) {
                    lo[r] = min(lo[r], __shfl_sync(FULL, rg.x, __ffs(ne) - 1));
                    hi[r] = max(hi[r], __shfl_sync(FULL, rg.y, 31 - __clz(ne)));
                }
            }
        }
    };
    int ncell = 0;
    bool run = false;
    int run_xa = 0, run_xb = 0, run_y = 0, run_z = 0;   // warp-uniform
    for (int chunk = 0; chunk < RB_THREADS / 32; ++chunk) {
        const int t = p0 + chunk * 32 + lane;
        const uint32_t key = (t < a.n) ? a.skeys[t] : (uint32_t)g.ncells;
        const bool live = key != (uint32_t)g.ncells;
        const bool first = live && ((chunk == 0 && lane == 0) || a.skeys[t - 1] != key);
        unsigned todo = __ballot_sync(FULL, first);
        ncell += __popc(todo);
        int cx = 0, cy = 0, cz = 0;
        if (first) decode_cell(g, key, cx, cy, cz);
        while (todo) {
            const int src = __ffs(todo) - 1;
            todo &= todo - 1;
            const int x = __shfl_sync(FULL, cx, src), y = __shfl_sync(FULL, cy, src), z = __shfl_sync(FULL, cz, src);
            if (run && y == run_y && z == run_z && x > run_xb) {
                run_xb = x;
            } else {
                if (run) flush(run_xa, run_xb, run_y, run_z);
                run = true;
                run_xa = run_xb = x;
                run_y = y;
                run_z = z;
            }
        }
    }
    if (run) flush(run_xa, run_xb, run_y, run_z);
    int lo_l = INT_MAX, hi_l = 0;   // lanes 0..8: row `lane`
#pragma unroll
    for (int r = 0; r < 9; ++r)
        if (lane == r) {
            lo_l = lo[r];
            hi_l = hi[r];
        }
    const int len = (lane < 9) ? max(hi_l - lo_l, 0) : 0;
    int sum = len;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(FULL, sum, o);
    TilePlan &tp = plans[tile];
    if (lane < 9) {
        tp.row_lo[lane] = (len > 0) ? lo_l : 0;
        tp.row_len[lane] = len;
    }
    if (lane == 0) {
        tp.slots = sum;
        const bool fits = sum <= RB_CAP && ncell <= RB_MAXC;
        tp.fits = fits ? 1 : 0;
        if (!fits && g.aligned) {
            // rows that do not fit the sweeps' shared-memory staging.  A tile of a few dense cells becomes a work item of the
            // dense kernels (density_dense_kernel / force_gather_kernel, one cell at a time: slots < 0 marks it); a
            // tile of many cells -- or any tile with the row-staged density (SPH_DENSITY=rows) -- is taken as 32-particle
            // passes (tile * 8 + 1 + pass) by the *_rows_items kernels.  Either way small grids on a second stream run
            // them next to the main sweeps.
            // (rows far beyond the staging mean dense neighbour cells: 32-particle passes would not fit either)
            if (dense_max_cells > 0 && (ncell <= dense_max_cells || sum > dense_min_slots)) {
                tp.slots = -max(sum, 1);
                dense_items[atomicAdd(n_dense, 1)] = tile * 8;   // counters are zeroed by reorder_kernel
            } else {
                const int np = (min(RB_THREADS, a.n - p0) + 31) / 32;
                const int q = atomicAdd(n_items, np);
                for (int u = 0; u < np; ++u) items[q + u] = tile * 8 + 1 + u;
            }
        }
    }
}

// First thing a sweep CTA does (warp 0 only, before any CTA-wide barrier): initialise the mbarrier and, if the tile has a
// fitting plan, put its TMA row copies in flight -- every other load of the CTA's prologue then overlaps with them.
template <bool DENSITY>
__device__ __forceinline__ void rows_issue_planned(const SweepArgs &a, const GridDesc &g, RowPlan &plan,
                                                   const TilePlan &tp, float4 *rows_a, float4 *rows_b) {
    const int lane = threadIdx.x & 31;
    if (lane == 0) {
        mbar_init(&plan.mbar, 1);
        fence_mbar_init();
    }
    __syncwarp();
    if (!g.aligned || !tp.fits) return;
    const int len = (lane < 9) ? tp.row_len[lane] : 0;
    const int lo = (lane < 9) ? tp.row_lo[lane] : 0;
    int inc = len;
#pragma unroll
    for (int o = 1; o < 16; o <<= 1) {
        const int u = __shfl_up_sync(FULL, inc, o);
        if (lane >= o) inc += u;
    }
    const int slots = __shfl_sync(FULL, inc, 8);
    if (lane < 9) {
        plan.row_base[lane] = inc - len;
        plan.row_lo[lane] = lo;
    }
    if (lane == 0) {
        plan.row_base[9] = slots;
        mbar_expect_tx(&plan.mbar, (uint32_t)slots * (DENSITY ? 16u : 32u));
    }
    __syncwarp();
    if (lane < 9 && len > 0) {
        bulk_g2s(&rows_a[inc - len], &a.spos[lo], (uint32_t)len * 16u, &plan.mbar);
        if (!DENSITY) bulk_g2s(&rows_b[inc - len], &a.svel[lo], (uint32_t)len * 16u, &plan.mbar);
    }
}

// Rest of the whole-tile setup (its TMA copies were issued by rows_issue_planned): the density sweep builds its per-cell
// segment tables while the rows are in flight.  ci: local cell index (density only).
template <bool DENSITY>
__device__ __forceinline__ void rows_setup_planned(const SweepArgs &a, const GridDesc &g, RowPlan &plan,
                                                   DensityRowsSmem *ds, int nb, int t, uint32_t key, uint32_t prev_key,
                                                   bool live, int &ci) {
    const int j = threadIdx.x, lane = j & 31, warp = j >> 5;
    if (DENSITY) {
        // local cells of the tile and their 27 segments (as row slots)
        const bool mine = live && j < nb;
        bool first = false;
        if (mine) first = (j == 0) || (prev_key != key);
        const unsigned bal = __ballot_sync(FULL, first);
        if (lane == 0) plan.wcount[warp] = __popc(bal);
        __syncthreads();   // also publishes row_lo / row_base
        int off = 0, total = 0;
#pragma unroll
        for (int w = 0; w < RB_WARPS; ++w) {
            const int c = plan.wcount[w];
            if (w < warp) off += c;
            total += c;
        }
        ci = off + __popc(bal & lanemask_le_()) - 1;
        if (first) plan.ckey[ci] = key;
        __syncthreads();
        // four cells per trip: their 27 range loads are independent, so issue them together (a tile of sparse
        // particles has ~50 cells and this loop would otherwise pay one global round trip per cell)
        const int rbase = plan.row_base[lane % 9] - plan.row_lo[lane % 9];
        for (int c0 = warp; c0 < total; c0 += 4 * RB_WARPS) {
            int2 r[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int c = c0 + u * RB_WARPS;
                r[u] = make_int2(0, 0);
                if (c < total && lane < 27) {
                    int cx, cy, cz;
                    decode_cell(g, plan.ckey[c], cx, cy, cz);
                    r[u] = neighbour_range(a, g, lane, cx, cy, cz);
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int c = c0 + u * RB_WARPS;
                if (c < total && lane < 27) {
                    const int cnt = min(r[u].y - r[u].x, 65535);
                    ds->seg[c * 27 + lane] = cnt > 0 ? ((uint32_t)(rbase + r[u].x) | ((uint32_t)cnt << 16)) : 0u;
                }
            }
        }
    }
    __syncthreads();
}

// lk + 2 if r < lim else lk: one FSETP + one predicated add (the list offset advances only on a hit)
__device__ __forceinline__ int advance_if_less(int lk, float r, float lim) {
    asm("{\n\t.reg .pred p;\n\tsetp.lt.f32 p, %1, %2;\n\t@p add.s32 %0, %0, 2;\n\t}" : "+r"(lk) : "f"(r), "f"(lim));
    return lk;
}

// Does the particle's own cell (from its position, fp64 division + C truncation: voxel_kernels.py:40-41) equal the
// cell it was sorted into?  Cheap fp32 interior test first; the exact test only near a cell face.
__device__ __forceinline__ bool own_cell_matches(const GridDesc &g, const float4 &p, int cx, int cy, int cz) {
    const float fx = p.x * g.inv_voxel[0] - (float)cx, fy = p.y * g.inv_voxel[1] - (float)cy,
                fz = p.z * g.inv_voxel[2] - (float)cz;
    const float lo = 1e-3f, hi = 1.f - 1e-3f;
    if (fx > lo && fx < hi && fy > lo && fy < hi && fz > lo && fz < hi) return true;
    int vx, vy, vz;
    return cell_of(g, p.x, p.y, p.z, vx, vy, vz) && vx == cx && vy == cy && vz == cz;
}

__device__ __noinline__ float sparse_density(const StepConsts &c, const float4 *rows, const uint16_t *lrow, int k,
                                             int self, float4 p) {
    float dens = 0.f;
    for (int i = 0; i < k; ++i) {
        const int slot = lrow[i];
        if (slot == self) continue;
        const float4 cj = rows[slot];
        const float dx = p.x - cj.x, dy = p.y - cj.y, dz = p.z - cj.z;
        dens = __fadd_rn(dens, poly6_term(c, fmaf(dz, dz, fmaf(dy, dy, dx * dx)), p.x, p.y, p.z, cj.x, cj.y, cj.z));
    }
    return dens;
}

// ---------------------------------------------------------------------------------------------------------------------
// density_kernel (voxel_kernels.py:108-132) + neighbour lists
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void publish_density(const SweepArgs &a, const StepConsts &c, int t, float dens,
                                                uint8_t cflag) {
    const float rho = dens * c.w_mass;
    a.srho[t] = rho;
    a.ncnt[t] = cflag;
    // per-particle pair factors ride in the .w lanes of the sorted arrays (read by the force rows)
    const_cast<float *>(reinterpret_cast<const float *>(a.spos + t))[3] = pressure_coeff(c, rho);
    const_cast<float *>(reinterpret_cast<const float *>(a.svel + t))[3] = c.lap_c / rho;
}

// One tile (128 consecutive sorted particles) of the row-staged density sweep; CTA-uniform control flow.
// only_pass >= 0: just that 32-particle pass (the fallback kernel spreads the passes of a tile whose rows do not fit
// over four CTAs instead of running them one after the other with three warps idle).
__device__ __forceinline__ void density_rows_tile(const SweepArgs &a, const GridDesc &g, const StepConsts &c,
                                                  const int tile, DensityRowsSmem &sm, const int only_pass = -1) {
    RowPlan &plan = sm.plan;
    const int j = threadIdx.x, lane = j & 31, warp = j >> 5;
    const int p0 = tile * RB_THREADS;
    const int t = p0 + j;
    const int nb = min(RB_THREADS, a.n - p0);
    if (warp == 0) rows_issue_planned<true>(a, g, plan, a.plans[tile], sm.rows, nullptr);
    // the prologue's loads, all issued up front (their use depends on the key, their addresses do not)
    const uint32_t key = (j < nb) ? a.skeys[t] : (uint32_t)g.ncells;
    const uint32_t prev_key = (j < nb && j > 0) ? a.skeys[t - 1] : 0xffffffffu;
    const float4 pi_raw = (j < nb) ? a.spos[t] : make_float4(0.f, 0.f, 0.f, 0.f);
    const bool live = key != (uint32_t)g.ncells;
    if (j < 8) sm.rows[RB_CAP + j] = make_float4(1e18f, 1e18f, 1e18f, 0.f);
    if (j < nb && !live) {   // dead particle: no neighbours
        a.srho[t] = 0.f;
        a.ncnt[t] = 0;
    }
    __syncthreads();   // mbarrier initialised, row_lo / row_base of a planned tile published
    {
        const TilePlan &tp = a.plans[tile];
        const bool plan_fits = g.aligned && tp.fits != 0 && tp.slots > 0;   // implies live particles
        if (only_pass == OWN_TILE && g.aligned && tp.fits == 0) return;     // its passes are work items
        if (!plan_fits && __syncthreads_count(live) == 0) return;
    }

    // this thread's particle
    float4 pi = make_float4(0.f, 0.f, 0.f, 0.f);
    int cx = 0, cy = 0, cz = 0;
    bool want = false, walk = false;
    if (live) {
        pi = pi_raw;
        decode_cell(g, key, cx, cy, cz);
        want = !(cx < g.own_lo - 1 || cx > g.own_hi);   // x-slab: nobody needs the density of the outer ghost column
        walk = want && (!g.aligned || !own_cell_matches(g, pi, cx, cy, cz));
    }

    uint32_t parity = 0;
    const bool tp_fits = g.aligned && a.plans[tile].fits != 0;
    // -1: whole CTA; 0..RB_WARPS-1: one warp's particles; RB_WARPS: no staging at all (everyone walks)
    int pass = g.aligned ? -1 : RB_WARPS;
    int j0 = 0, j1 = nb;
    if (only_pass >= 0 && g.aligned) {
        pass = only_pass;
        j0 = pass * 32;
        j1 = min(nb, j0 + 32);
        if (j0 >= nb) return;
    }
    while (pass < RB_WARPS) {
        int ci = 0;
        bool ok;
        if (pass < 0) {   // whole tile: planned by rows_plan_kernel, or known not to fit
            ok = tp_fits;
            if (ok) rows_setup_planned<true>(a, g, plan, &sm, nb, t, key, prev_key, live, ci);
        } else {
            ok = rows_setup<true>(a, g, plan, sm.rows, nullptr, &sm, j0, j1, t, key, live, ci);
        }
        const bool in_pass = live && want && j >= j0 && j < j1;
        if (ok) {
            while (!mbar_try_wait(&plan.mbar, parity)) {
            }
            parity ^= 1u;
            const bool active = in_pass && !walk;
            if (active) {
                // lane = particle, in sorted order: the lanes of a warp share one or two cells, hence read the same
                // candidate most of the time (broadcast).  (Regrouping the CTA's particles by fractional x, which
                // governs the scan length, was measured slower: lanes of different cells diverge at every segment.)
                const int jj = j;
                const int tj = t;
                const int selfj = plan.row_base[4] + (t - plan.row_lo[4]);   // row 4 = (dy, dz) = (0, 0): the tile's own cells
                const float4 pj = pi;
                uint16_t *const lrow = sm.list + jj * RB_LSTRIDE;
                // ---- scan: walk the cell's 27 segments in reference order over the staged rows, keep the first
                //      RB_KEEP candidates of the fp32 superset r2 < h^2 (1 + 1e-5); two candidates per trip ----
                // next free entry of the list row, as a BYTE offset into sm.list: every candidate is stored, the offset
                // only advances on a hit (one STS + one predicated add per candidate)
                char *const lbytes = reinterpret_cast<char *>(sm.list);
                int lk = jj * RB_LSTRIDE * 2;
                const int lend = lk + RB_KEEP * 2;
                const uint32_t *seg = sm.seg + ci * 27;
                for (int s = 0; s < 27 && lk < lend; ++s) {
                    const uint32_t e = seg[s];
                    int slot = (int)(e & 0xffffu);
                    int n = (int)(e >> 16);
                    // four candidates per trip, double-buffered: (a0, a1) were loaded one trip ahead, (b0, b1) are
                    // loaded at the top of the trip they are used in -- LDS latency hides behind the other pair's math
                    float4 a0 = sm.rows[slot], a1 = sm.rows[slot + 1];
#define SPH_TEST_CANDIDATE(cand, sl)                                                        \
    {                                                                                       \
        const float dx_ = pj.x - (cand).x, dy_ = pj.y - (cand).y, dz_ = pj.z - (cand).z;    \
        const float r_ = fmaf(dz_, dz_, fmaf(dy_, dy_, dx_ * dx_));                         \
        *reinterpret_cast<uint16_t *>(lbytes + lk) = (uint16_t)(sl);                        \
        lk = advance_if_less(lk, r_, c.h2_hi);                                              \
    }
                    while (n >= 4) {
                        const float4 b0 = sm.rows[slot + 2], b1 = sm.rows[slot + 3];
                        SPH_TEST_CANDIDATE(a0, slot)
                        SPH_TEST_CANDIDATE(a1, slot + 1)
                        a0 = sm.rows[slot + 4];   // next trip (at worst sentinels)
                        a1 = sm.rows[slot + 5];
                        SPH_TEST_CANDIDATE(b0, slot + 2)
                        SPH_TEST_CANDIDATE(b1, slot + 3)
                        slot += 4;
                        n -= 4;
                        if (lk >= lend) break;
                    }
                    if (n >= 2 && n < 4 && lk < lend) {
                        const float4 b0 = sm.rows[slot + 2];
                        SPH_TEST_CANDIDATE(a0, slot)
                        SPH_TEST_CANDIDATE(a1, slot + 1)
                        a0 = b0;
                        slot += 2;
                        n -= 2;
                    }
                    if (n == 1 && lk < lend) SPH_TEST_CANDIDATE(a0, slot)
#undef SPH_TEST_CANDIDATE
                }
                const int kept = (lk >> 1) - jj * RB_LSTRIDE;   // <= RB_KEEP + 1
                // ---- second pass over the kept candidates: poly6 density in list order + the fp64 predicate ----
                int k = min(kept, kMaxNeighbours);
                float dens = 0.f;
                bool band = false;
                // common case: no kept candidate inside the rounding band -> the list is final as it stands; four
                // candidates per trip (loads first), poly6 sum over the first 32 in list order
#pragma unroll 1
                for (int i = 0; i < k; i += 4) {
                    float4 cj[4];
                    int sl[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) sl[u] = lrow[min(i + u, k - 1)];
#pragma unroll
                    for (int u = 0; u < 4; ++u) cj[u] = sm.rows[sl[u]];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const float dx = pj.x - cj[u].x, dy = pj.y - cj[u].y, dz = pj.z - cj[u].z;
                        const float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                        band = band || r2 > c.h2_lo;
                        const float w = (i + u < k && sl[u] != selfj) ? poly6_fast(c, r2) : 0.f;
                        dens = __fadd_rn(dens, w);
                    }
                }
                for (int i = k; i < kept; ++i) {   // spares beyond the 32nd: only matter if something gets rejected
                    const float4 cj = sm.rows[lrow[i]];
                    const float dx = pj.x - cj.x, dy = pj.y - cj.y, dz = pj.z - cj.z;
                    band = band || fmaf(dz, dz, fmaf(dy, dy, dx * dx)) > c.h2_lo;
                }
                if (band) {
                    // rare: fp64 predicate for the candidates inside the band, first 32 accepted stay (compacted in place)
                    k = 0;
                    dens = 0.f;
                    for (int i = 0; i < kept && k < kMaxNeighbours; ++i) {
                        const int slot = lrow[i];
                        const float4 cj = sm.rows[slot];
                        const float dx = pj.x - cj.x, dy = pj.y - cj.y, dz = pj.z - cj.z;
                        const float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                        if (r2 > c.h2_lo && !in_range_exact(pj.x, pj.y, pj.z, cj.x, cj.y, cj.z, c.r2_max)) continue;
                        lrow[k++] = (uint16_t)slot;
                        if (slot != selfj) dens = __fadd_rn(dens, poly6_fast(c, r2));
                    }
                }
                // a sparse particle's density can hinge on one neighbour at the cut-off, where h^2 - r^2 cancels in fp32:
                // redo the few terms with the fp64 r^2 (never taken by capped lists, where such a term is negligible)
                if (k <= kSparseCount) dens = sparse_density(c, sm.rows, lrow, k, selfj, pj);
                uint8_t cflag = (uint8_t)k;
                if (k < kMaxNeighbours && kept >= RB_KEEP) {
                    // the scan stopped on its budget and too many band candidates were rejected: redo exactly
                    ForceAcc dummy;
                    dens = 0.f;
                    const int wc = thread_walk<false>(a, g, c, tj, pj, pj, 0.f, dens, dummy);
                    cflag = (uint8_t)wc | CNT_WALK;
                }
                publish_density(a, c, tj, dens, cflag);
            }
            if (in_pass && walk) {   // aliased key (quirk Q5): plain walk by the particle's own thread
                ForceAcc dummy;
                float dens = 0.f;
                const int wc = thread_walk<false>(a, g, c, t, pi, pi, 0.f, dens, dummy);
                publish_density(a, c, t, dens, (uint8_t)wc | CNT_WALK);
            }
            __syncwarp();
            // lists: shared -> HBM, 64 B per particle, 16 B per lane and trip (rows of this warp's own particles)
            {
                const int wbase = warp * 32;
                uint4 *gl = reinterpret_cast<uint4 *>(a.nlist + (size_t)(p0 + wbase) * 32);
#pragma unroll
                for (int it = 0; it < 4; ++it) {
                    const int row = it * 8 + (lane >> 2), part = lane & 3;   // 4 lanes x 16 B per row
                    const int jj = wbase + row;
                    if (jj >= j0 && jj < j1 && jj < nb) {
                        const uint32_t *src = reinterpret_cast<const uint32_t *>(sm.list + jj * RB_LSTRIDE) + part * 4;
                        gl[row * 4 + part] = make_uint4(src[0], src[1], src[2], src[3]);
                    }
                }
            }
        } else if (pass >= 0) {
            // a 32-particle pass that still does not fit: everyone walks
            if (in_pass) {
                ForceAcc dummy;
                float dens = 0.f;
                const int wc = thread_walk<false>(a, g, c, t, pi, pi, 0.f, dens, dummy);
                publish_density(a, c, t, dens, (uint8_t)wc | CNT_WALK);
            }
        }
        if ((ok && pass < 0) || only_pass >= 0) break;
        ++pass;
        if (pass * 32 >= nb) break;
        j0 = pass * 32;
        j1 = min(nb, j0 + 32);
        __syncthreads();
    }
    if (pass == RB_WARPS && !g.aligned) {   // Q2 grid: no staging, every particle walks
        if (live && want) {
            ForceAcc dummy;
            float dens = 0.f;
            const int wc = thread_walk<false>(a, g, c, t, pi, pi, 0.f, dens, dummy);
            publish_density(a, c, t, dens, (uint8_t)wc | CNT_WALK);
        }
    }
}

// Work items: the 32-particle passes of tiles whose rows do not fit (rows_plan_kernel) and whole tiles density_flat_kernel
// hands over.  A small grid on a second, higher-priority stream runs them next to the main sweep, so the long items
// start first and run side by side instead of one after the other inside one CTA of the main grid.
__global__ void __launch_bounds__(RB_THREADS, RB_DENSITY_CTAS)
density_rows_items_kernel(const SweepArgs a, const GridDesc g, const StepConsts c, const int *__restrict__ n_items,
                          const int *__restrict__ items) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    DensityRowsSmem &sm = *reinterpret_cast<DensityRowsSmem *>(smem_raw);
    const int n = *n_items;
    for (int i = blockIdx.x; i < n; i += gridDim.x) {
        const int item = items[i];   // tile * 8 + (0: whole tile | 1 + pass)
        density_rows_tile(a, g, c, item >> 3, sm, (item & 7) - 1);
        __syncthreads();
        if (threadIdx.x == 0) mbar_inval(&sm.plan.mbar);   // the next item initialises it again
        __syncthreads();
    }
}

// Whole grid through the row-staged sweep (SPH_DENSITY=rows: A/B against density_flat_kernel of sweep_flat.cuh).
__global__ void __launch_bounds__(RB_THREADS, RB_DENSITY_CTAS)
density_rows_kernel(const SweepArgs a, const GridDesc g, const StepConsts c) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    density_rows_tile(a, g, c, blockIdx.x, *reinterpret_cast<DensityRowsSmem *>(smem_raw), OWN_TILE);
}

// ---------------------------------------------------------------------------------------------------------------------
// pressure_kernel + viscosity_kernel + integrating_kernel + collision kernel in one sweep
// (voxel_kernels.py:135-211, base_kernels.py:30-98), driven by the slot lists of density_rows_kernel.
// ---------------------------------------------------------------------------------------------------------------------
template <bool RECORD>
__device__ __forceinline__ void force_rows_tile(const SweepArgs &a, const GridDesc &g, const StepConsts &c,
                                                const int tile, ForceRowsSmem &sm, const int only_pass = -1) {
    RowPlan &plan = sm.plan;
    const int j = threadIdx.x;
    const int p0 = tile * RB_THREADS;
    const int t = p0 + j;
    const int nb = min(RB_THREADS, a.n - p0);
    const uint32_t key = (j < nb) ? a.skeys[t] : (uint32_t)g.ncells;
    const bool live = key != (uint32_t)g.ncells;
    if ((j >> 5) == 0) rows_issue_planned<false>(a, g, plan, a.plans[tile], sm.rpos, sm.rvel);
    // issued before the key arrives (their use depends on it, their address does not)
    const uint8_t cf_raw = (j < nb) ? a.ncnt[t] : (uint8_t)0;
    const float rho_raw = (j < nb) ? a.srho[t] : 0.f;
    const uint32_t my_id = (j < nb) ? a.sids[t] : 0u;   // master slot of the epilogue's scatter
    // the lists too: all four 16-byte pieces, whatever the count turns out to be (one round trip instead of two)
    uint4 e[4];
    {
        const uint4 *lg = reinterpret_cast<const uint4 *>(a.nlist + (size_t)t * 32);
#pragma unroll
        for (int q = 0; q < 4; ++q) e[q] = (j < nb) ? __ldg(&lg[q]) : make_uint4(0u, 0u, 0u, 0u);
    }
    if (only_pass == OWN_TILE) {   // streaming inputs of the tile that will run PF_AHEAD CTAs later -> L2
        const int pt = tile + PF_AHEAD;
        const int pp0 = pt * RB_THREADS;
        if (pp0 + RB_THREADS <= a.n) {
            if ((j >> 5) == 1) {
                if (j == 32) prefetch_l2(&a.plans[pt]);
                else if (j < 32 + 5) prefetch_l2(reinterpret_cast<const char *>(a.skeys + pp0) + (j - 33) * 128);
                else if (j < 32 + 9) prefetch_l2(reinterpret_cast<const char *>(a.srho + pp0) + (j - 37) * 128);
                else if (j < 32 + 13) prefetch_l2(reinterpret_cast<const char *>(a.sids + pp0) + (j - 41) * 128);
                else if (j == 32 + 13) prefetch_l2(a.ncnt + pp0);
            } else if ((j >> 5) >= 2) {   // 64 lines of lists
                prefetch_l2(reinterpret_cast<const char *>(a.nlist + (size_t)pp0 * 32) + (j - 64) * 128);
            }
        }
    }
    __syncthreads();   // mbarrier initialised, row_lo / row_base of a planned tile published
    // a fitting plan implies live particles; otherwise count them (a tile of dead particles stages nothing)
    const bool plan_fits = g.aligned && a.plans[tile].fits != 0 && a.plans[tile].slots > 0;
    if (only_pass == OWN_TILE && g.aligned && a.plans[tile].fits == 0) return;   // its passes are work items
    const bool any_live = plan_fits || __syncthreads_count(live) != 0;

    int cx = 0, cy = 0, cz = 0, my_cnt = 0;
    bool want = false, walk = false;
    float rho_i = 0.f;
    if (live) {
        decode_cell(g, key, cx, cy, cz);
        want = !(cx < g.own_lo || cx >= g.own_hi);   // x-slab: ghost cell, its owner computes the forces
        if (want) {
            walk = (cf_raw & CNT_WALK) != 0;
            my_cnt = walk ? 0 : cf_raw;
            rho_i = rho_raw;
        }
    }
    // dead particle: F = external force, rho = 0 (reference NaN semantics carry on); finished in the first trip (of the
    // pass that covers it)
    bool dead_todo = j < nb && !live && (only_pass < 0 || (j >> 5) == only_pass);

    uint32_t parity = 0;
    const bool tp_fits = g.aligned && any_live && a.plans[tile].fits != 0;
    // -1: whole tile; 0..RB_WARPS-1: one warp's particles; RB_WARPS: no staging (Q2 grid: everyone walks; or nothing alive)
    int pass = (g.aligned && any_live) ? -1 : RB_WARPS;
    int j0 = 0, j1 = nb;
    if (only_pass >= 0 && pass < 0) {
        pass = only_pass;
        j0 = pass * 32;
        j1 = min(nb, j0 + 32);
        if (j0 >= nb) return;
    }
    for (;;) {
        // every path that completes a particle funnels into the ONE finish_particle call at the bottom of the trip
        ForceAcc f;
        float4 pi = make_float4(0.f, 0.f, 0.f, 0.f), vi = pi;
        float rho_f = rho_i;
        bool fin = false, need_walk = false;
        const bool was_dead = dead_todo;
        if (dead_todo) {
            pi = a.spos[t];
            vi = a.svel[t];
            rho_f = rho_raw;
            fin = true;
            dead_todo = false;
        }
        if (pass < RB_WARPS) {
            int ci = 0;
            bool ok;
            if (pass < 0) {
                ok = tp_fits;
                // planned tile: nothing left to set up (rows in flight since the first instruction of warp 0)
            } else {
                ok = rows_setup<false>(a, g, plan, sm.rpos, sm.rvel, nullptr, j0, j1, t, key, live, ci);
            }
            const bool in_pass = live && want && j >= j0 && j < j1;
            if (ok) {
                while (!mbar_try_wait(&plan.mbar, parity)) {
                }
                parity ^= 1u;
                if (in_pass && !walk) {
                    const int self = plan.row_base[4] + (t - plan.row_lo[4]);
                    pi = sm.rpos[self];
                    vi = sm.rvel[self];
                    const float a_i = pi.w;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        if (q * 8 < my_cnt) {
                            const uint32_t w[4] = {e[q].x, e[q].y, e[q].z, e[q].w};
#pragma unroll
                            for (int u = 0; u < 8; ++u) {
                                const int raw = (int)((w[u >> 1] >> ((u & 1) * 16)) & 0xffffu);
                                const int slot = (q * 8 + u < my_cnt) ? raw : self;
                                f.pair(c, pi, vi, a_i, sm.rpos[slot], sm.rvel[slot], slot != self);
                            }
                        }
                    }
                    fin = true;
                } else if (in_pass) {
                    need_walk = true;
                }
            } else if (pass >= 0 && in_pass) {
                need_walk = true;   // a 32-particle pass that still does not fit
            }
        } else if (!g.aligned && live && want) {
            need_walk = true;       // Q2 grid
        }
        if (need_walk) {   // exact one-thread walk over global memory (rare)
            ForceAcc fw;
            float dens = 0.f;
            pi = a.spos[t];
            vi = a.svel[t];
            thread_walk<true>(a, g, c, t, pi, vi, pressure_coeff(c, rho_i), dens, fw);
            f = fw;
            fin = true;
        }
        if (fin) finish_particle<RECORD>(a, c, t, pi, vi, rho_f, f, my_id, was_dead);

        if (pass >= RB_WARPS || (pass < 0 && tp_fits) || only_pass >= 0) break;
        ++pass;
        if (pass * 32 >= nb) break;
        j0 = pass * 32;
        j1 = min(nb, j0 + 32);
        __syncthreads();
    }
}

// Parity tap (voxel_kernels.py:78-85, `neighbours`): the lists of the most recent density sweep as PARTICLE IDS, id
// order, -1 padded.  Particles of planned tiles are decoded from the slot lists the sweeps really used; walk-path
// particles and tiles taken as 32-particle passes (pass-local slots) are re-walked.
__global__ void __launch_bounds__(128)
neighbour_lists_kernel(const SweepArgs a, const GridDesc g, const StepConsts c, int32_t *__restrict__ out,
                       const uint32_t *__restrict__ dlist) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.n) return;
    const uint32_t key = a.skeys[t];
    int32_t *row = out + (size_t)a.sids[t] * kMaxNeighbours;
    for (int k = 0; k < kMaxNeighbours; ++k) row[k] = -1;
    if (key == (uint32_t)g.ncells) return;
    const TilePlan &tp = a.plans[t / RB_THREADS];
    const uint8_t cf = a.ncnt[t];
    if (g.aligned && !(cf & CNT_WALK) && !tp.fits && tp.slots < 0 && dlist) {   // dense tile: sorted indices (density_dense_kernel)
        for (int k = 0; k < (int)cf; ++k) row[k] = (int32_t)a.sids[dlist[(size_t)t * kMaxNeighbours + k]];
        return;
    }
    if (g.aligned && !(cf & CNT_WALK) && tp.fits) {
        int base[10];
        base[0] = 0;
        for (int r = 0; r < 9; ++r) base[r + 1] = base[r] + tp.row_len[r];
        const uint16_t *lst = a.nlist + (size_t)t * kMaxNeighbours;
        for (int k = 0; k < (int)cf; ++k) {
            const int slot = lst[k];
            int r = 0;
            while (r < 8 && slot >= base[r + 1]) ++r;
            row[k] = (int32_t)a.sids[tp.row_lo[r] + (slot - base[r])];
        }
        return;
    }
    const float4 pi = a.spos[t];
    int vx, vy, vz, k = 0;
    if (!cell_of(g, pi.x, pi.y, pi.z, vx, vy, vz)) return;
    for (int dx = -1; dx <= 1; ++dx)
        for (int dy = -1; dy <= 1; ++dy)
            for (int dz = -1; dz <= 1; ++dz) {
                const int x = vx + dx, y = vy + dy, z = vz + dz;
                if (x < 0 || x >= g.tx || y < 0 || y >= g.ty || z < 0 || z >= g.tz) continue;
                const long long cl = (long long)x - g.xoff + (long long)y * g.wn + (long long)z * g.wn * g.hn;
                if (cl < 0 || cl >= g.ncells) continue;
                const int2 r = __ldg(&a.cell_range[cl]);
                for (int j = r.x; j < r.y; ++j) {
                    const float4 pj = __ldg(&a.spos[j]);
                    if (!in_range_exact(pi.x, pi.y, pi.z, pj.x, pj.y, pj.z, c.r2_max)) continue;
                    row[k] = (int32_t)a.sids[j];
                    if (++k >= kMaxNeighbours) return;
                }
            }
}

template <bool RECORD>
__global__ void __launch_bounds__(RB_THREADS, RB_FORCE_CTAS)
force_rows_kernel(const SweepArgs a, const GridDesc g, const StepConsts c) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    force_rows_tile<RECORD>(a, g, c, blockIdx.x, *reinterpret_cast<ForceRowsSmem *>(smem_raw), OWN_TILE);
}

// The passes of tiles whose rows do not fit (work items of rows_plan_kernel), next to force_rows_kernel.
template <bool RECORD>
__global__ void __launch_bounds__(RB_THREADS, RB_FORCE_CTAS)
force_rows_items_kernel(const SweepArgs a, const GridDesc g, const StepConsts c) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    ForceRowsSmem &sm = *reinterpret_cast<ForceRowsSmem *>(smem_raw);
    const int n = *a.n_items;
    for (int i = blockIdx.x; i < n; i += gridDim.x) {
        const int item = a.items[i];
        force_rows_tile<RECORD>(a, g, c, item >> 3, sm, (item & 7) - 1);
        __syncthreads();
        if (threadIdx.x == 0) mbar_inval(&sm.plan.mbar);
        __syncthreads();
    }
}

}  // namespace sph
