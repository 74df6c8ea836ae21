// Stable LSD radix sort of (cell key, particle index) pairs, "onesweep" organisation -- the device replacement of the
// reference's host-side numpy structured sort (voxel_sph_strategy.py:81-88).  Stability + values starting as iota give
// exactly numpy's (voxel_id, particle_id) order.
//
//   os_hist     one pass over the keys: the global histogram of EVERY digit position at once
//   os_pass x P one launch per 8-bit digit: a CTA takes the next tile (atomic ticket, so a tile's predecessors are
//               always resident or finished), ranks its 2048 pairs stably (warp multisplit by ballots, warps in input
//               order), publishes its per-digit counts and obtains its global offsets by decoupled look-back over the
//               preceding tiles' status words (flag and value share one 32-bit word: no fences, nothing to tear),
//               then scatters.
// Keys are read P + 1 times and written P times (the three-kernel version in radix_sort.cuh reads them 2 P times).
#pragma once
#include "sph_common.cuh"

namespace sph {

constexpr int OS_THREADS = 256;
constexpr int OS_ITEMS = 8;
constexpr int OS_TILE = OS_THREADS * OS_ITEMS;   // 2048 pairs per tile
constexpr int OS_WARPS = OS_THREADS / 32;
constexpr int OS_RADIX = 256;
constexpr int OS_MAX_PASSES = 4;
constexpr uint32_t OS_FLAG_AGG = 1u << 30;       // status word: tile aggregate available
constexpr uint32_t OS_FLAG_PREFIX = 2u << 30;    // status word: inclusive prefix available
constexpr uint32_t OS_VALUE_MASK = (1u << 30) - 1u;

struct OsPasses {
    int n_passes;
    int shift[OS_MAX_PASSES];
    uint32_t mask[OS_MAX_PASSES];
};

// Control block layout (uint32 words), zeroed by one memset per sort:
//   [0, P*256)                         global digit histograms
//   [P*256, P*256 + P)                 tile tickets
//   [P*256 + 8, ... + P*ntiles*256)    look-back status words
__host__ __device__ inline size_t os_ctrl_words(int passes, int ntiles) {
    return (size_t)passes * OS_RADIX + 8 + (size_t)passes * ntiles * OS_RADIX;
}

__global__ void __launch_bounds__(OS_THREADS)
os_hist(const uint32_t *__restrict__ keys, int n, OsPasses ps, uint32_t *__restrict__ ctrl) {
    __shared__ uint32_t hist[OS_MAX_PASSES][OS_RADIX];
    const int tid = threadIdx.x;
    for (int p = 0; p < ps.n_passes; ++p) hist[p][tid] = 0;
    __syncthreads();
    for (int i = blockIdx.x * OS_THREADS + tid; i < n; i += gridDim.x * OS_THREADS) {
        const uint32_t k = keys[i];
        for (int p = 0; p < ps.n_passes; ++p) atomicAdd(&hist[p][(k >> ps.shift[p]) & ps.mask[p]], 1u);
    }
    __syncthreads();
    for (int p = 0; p < ps.n_passes; ++p) {
        const uint32_t v = hist[p][tid];
        if (v) atomicAdd(&ctrl[p * OS_RADIX + tid], v);
    }
}

// hash_kernel + os_hist in one pass over the positions (assign_voxels_to_particles_kernel, voxel_kernels.py:88-105, and
// the digit histograms of every pass of the sort that follows): the keys are written once and not read back here.
__global__ void __launch_bounds__(OS_THREADS)
hash_hist_kernel(const float4 *__restrict__ pos_m, uint32_t *__restrict__ keys, int n, GridDesc g, OsPasses ps,
                 uint32_t *__restrict__ ctrl) {
    __shared__ uint32_t hist[OS_MAX_PASSES][OS_RADIX];
    const int tid = threadIdx.x;
    for (int p = 0; p < ps.n_passes; ++p) hist[p][tid] = 0;
    __syncthreads();
    // four independent position loads in flight per thread (the fp64 division of key_of sits behind each of them)
    const int stride = gridDim.x * OS_THREADS;
    for (int i0 = blockIdx.x * OS_THREADS + tid; i0 < n; i0 += 4 * stride) {
        float4 q[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + u * stride;
            if (i < n) q[u] = pos_m[MI(i)];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + u * stride;
            if (i < n) {
                const uint32_t k = key_of(g, q[u].x, q[u].y, q[u].z);
                keys[i] = k;
                for (int p = 0; p < ps.n_passes; ++p) atomicAdd(&hist[p][(k >> ps.shift[p]) & ps.mask[p]], 1u);
            }
        }
    }
    __syncthreads();
    for (int p = 0; p < ps.n_passes; ++p) {
        const uint32_t v = hist[p][tid];
        if (v) atomicAdd(&ctrl[p * OS_RADIX + tid], v);
    }
}

// peers of this lane's digit (8 bits) among the valid lanes, by ballots
__device__ __forceinline__ uint32_t os_match(uint32_t d, bool valid) {
    uint32_t peers = __ballot_sync(0xffffffffu, valid);
#pragma unroll
    for (int b = 0; b < 8; ++b) {
        const uint32_t bal = __ballot_sync(0xffffffffu, (d >> b) & 1u);
        peers &= ((d >> b) & 1u) ? bal : ~bal;
    }
    return peers;
}

// vin == nullptr means "values are iota" (first pass).
__global__ void __launch_bounds__(OS_THREADS)
os_pass(const uint32_t *__restrict__ kin, const uint32_t *__restrict__ vin, uint32_t *__restrict__ kout,
        uint32_t *__restrict__ vout, int n, int pass, int n_passes, int shift, uint32_t mask, int ntiles,
        uint32_t *__restrict__ ctrl) {
    __shared__ uint32_t wcnt[OS_WARPS][OS_RADIX];
    __shared__ uint32_t warp_sum[OS_WARPS];
    __shared__ int tile_s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) tile_s = (int)atomicAdd(&ctrl[n_passes * OS_RADIX + pass], 1u);
#pragma unroll
    for (int w = 0; w < OS_WARPS; ++w) wcnt[w][tid] = 0;
    __syncthreads();
    const int tile = tile_s;
    volatile uint32_t *status = ctrl + (size_t)n_passes * OS_RADIX + 8 + ((size_t)pass * ntiles) * OS_RADIX;

    // stable ranks inside the warp's 256 pairs (the warp visits them 32 at a time, in input order)
    const int wbase = tile * OS_TILE + warp * (32 * OS_ITEMS);
    uint32_t key[OS_ITEMS], rank[OS_ITEMS];
    const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
    for (int k = 0; k < OS_ITEMS; ++k) {
        const int i = wbase + k * 32 + lane;
        key[k] = (i < n) ? kin[i] : 0u;
    }
#pragma unroll
    for (int k = 0; k < OS_ITEMS; ++k) {
        const int i = wbase + k * 32 + lane;
        const bool valid = i < n;
        const uint32_t d = (key[k] >> shift) & mask;
        const uint32_t peers = os_match(d, valid);
        uint32_t pre = 0;
        if (valid) pre = wcnt[warp][d];
        __syncwarp();
        if (valid && (peers & lt) == 0) wcnt[warp][d] = pre + __popc(peers);
        __syncwarp();
        rank[k] = pre + __popc(peers & lt);
    }
    __syncthreads();

    // digit `tid`: tile count, publish, look back
    uint32_t tile_cnt = 0;
#pragma unroll
    for (int w = 0; w < OS_WARPS; ++w) tile_cnt += wcnt[w][tid];
    uint32_t excl = 0;
    if (tile == 0) {
        status[tid] = OS_FLAG_PREFIX | tile_cnt;
    } else {
        status[(size_t)tile * OS_RADIX + tid] = OS_FLAG_AGG | tile_cnt;
        int prev = tile - 1;
        while (true) {
            const uint32_t s = status[(size_t)prev * OS_RADIX + tid];
            const uint32_t flag = s & ~OS_VALUE_MASK;
            if (flag == 0) continue;   // predecessor has a ticket, hence is resident or done: it will publish
            excl += s & OS_VALUE_MASK;
            if (flag == OS_FLAG_PREFIX) break;
            --prev;
        }
        status[(size_t)tile * OS_RADIX + tid] = OS_FLAG_PREFIX | (excl + tile_cnt);
    }

    // exclusive scan of the 256 global digit totals -> global base of digit `tid`
    const uint32_t tot = ctrl[pass * OS_RADIX + tid];
    uint32_t inc = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += u;
    }
    if (lane == 31) warp_sum[warp] = inc;
    __syncthreads();
    uint32_t running = inc - tot + excl;
#pragma unroll
    for (int w = 0; w < OS_WARPS; ++w)
        if (w < warp) running += warp_sum[w];
    // per-warp start offsets (warp order == input order)
#pragma unroll
    for (int w = 0; w < OS_WARPS; ++w) {
        const uint32_t c = wcnt[w][tid];
        wcnt[w][tid] = running;
        running += c;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < OS_ITEMS; ++k) {
        const int i = wbase + k * 32 + lane;
        if (i < n) {
            const uint32_t d = (key[k] >> shift) & mask;
            const uint32_t dst = wcnt[warp][d] + rank[k];
            kout[dst] = key[k];
            vout[dst] = vin ? vin[i] : (uint32_t)i;
        }
    }
}

// ---- chain-free variant: per-tile digit counts + one scan per digit instead of the look-back ----------------------------
// The look-back chain costs about L * sqrt(2 * tiles) of serial L2 round trips per pass when all tiles are resident at
// once (26 us per pass at 512 tiles); counting first and scanning per digit removes the chain for two extra launches.
//   ts_hist    per-tile digit counts of this pass's input order       -> tile_cnt[digit][tile]
//   ts_scan    one CTA per digit: exclusive scan over the tiles + the global digit base (from os_hist's totals)
//   ts_scatter stable ranking as in os_pass, offsets read from tile_cnt
__global__ void __launch_bounds__(OS_THREADS)
ts_hist(const uint32_t *__restrict__ keys, int n, int shift, uint32_t mask, uint32_t *__restrict__ tile_cnt,
        int ntiles) {
    __shared__ uint32_t hist[OS_RADIX];
    const int tid = threadIdx.x;
    hist[tid] = 0;
    __syncthreads();
    const int base = blockIdx.x * OS_TILE;
#pragma unroll
    for (int k = 0; k < OS_ITEMS; ++k) {
        const int i = base + k * OS_THREADS + tid;
        if (i < n) atomicAdd(&hist[(keys[i] >> shift) & mask], 1u);
    }
    __syncthreads();
    tile_cnt[(size_t)tid * ntiles + blockIdx.x] = hist[tid];
}

// grid = OS_RADIX CTAs; CTA d: tile_cnt[d][0..ntiles) -> exclusive prefix + global base of digit d
__global__ void __launch_bounds__(OS_THREADS)
ts_scan(uint32_t *__restrict__ tile_cnt, int ntiles, const uint32_t *__restrict__ digit_total) {
    __shared__ uint32_t warp_sum[OS_WARPS];
    __shared__ uint32_t carry_s;
    const int d = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // global base of digit d = sum of the totals of the smaller digits
    uint32_t part = (tid < d) ? digit_total[tid] : 0u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if (lane == 0) warp_sum[warp] = part;
    __syncthreads();
    if (tid == 0) {
        uint32_t b = 0;
        for (int w = 0; w < OS_WARPS; ++w) b += warp_sum[w];
        carry_s = b;
    }
    __syncthreads();
    uint32_t *row = tile_cnt + (size_t)d * ntiles;
    for (int base = 0; base < ntiles; base += OS_THREADS) {
        const int i = base + tid;
        const uint32_t v = (i < ntiles) ? row[i] : 0;
        uint32_t inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += u;
        }
        __syncthreads();   // warp_sum / carry_s of the previous round have been read
        if (lane == 31) warp_sum[warp] = inc;
        __syncthreads();
        uint32_t woff = 0;
#pragma unroll
        for (int w = 0; w < OS_WARPS; ++w)
            if (w < warp) woff += warp_sum[w];
        const uint32_t carry = carry_s;
        if (i < ntiles) row[i] = carry + woff + inc - v;
        __syncthreads();
        if (tid == OS_THREADS - 1) carry_s = carry + woff + inc;
    }
}

__global__ void __launch_bounds__(OS_THREADS)
ts_scatter(const uint32_t *__restrict__ kin, const uint32_t *__restrict__ vin, uint32_t *__restrict__ kout,
           uint32_t *__restrict__ vout, int n, int shift, uint32_t mask, const uint32_t *__restrict__ tile_off,
           int ntiles) {
    __shared__ uint32_t wcnt[OS_WARPS][OS_RADIX];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = blockIdx.x;
#pragma unroll
    for (int w = 0; w < OS_WARPS; ++w) wcnt[w][tid] = 0;
    const uint32_t tbase = tile_off[(size_t)tid * ntiles + tile];   // global start of (digit tid, this tile)
    __syncthreads();
    const int wbase = tile * OS_TILE + warp * (32 * OS_ITEMS);
    uint32_t key[OS_ITEMS], rank[OS_ITEMS];
    const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
    for (int k = 0; k < OS_ITEMS; ++k) {
        const int i = wbase + k * 32 + lane;
        key[k] = (i < n) ? kin[i] : 0u;
    }
#pragma unroll
    for (int k = 0; k < OS_ITEMS; ++k) {
        const int i = wbase + k * 32 + lane;
        const bool valid = i < n;
        const uint32_t d = (key[k] >> shift) & mask;
        const uint32_t peers = os_match(d, valid);
        uint32_t pre = 0;
        if (valid) pre = wcnt[warp][d];
        __syncwarp();
        if (valid && (peers & lt) == 0) wcnt[warp][d] = pre + __popc(peers);
        __syncwarp();
        rank[k] = pre + __popc(peers & lt);
    }
    __syncthreads();
    uint32_t running = tbase;
#pragma unroll
    for (int w = 0; w < OS_WARPS; ++w) {
        const uint32_t c = wcnt[w][tid];
        wcnt[w][tid] = running;
        running += c;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < OS_ITEMS; ++k) {
        const int i = wbase + k * 32 + lane;
        if (i < n) {
            const uint32_t d = (key[k] >> shift) & mask;
            const uint32_t dst = wcnt[warp][d] + rank[k];
            kout[dst] = key[k];
            vout[dst] = vin ? vin[i] : (uint32_t)i;
        }
    }
}


// ---- second generation pass (the default): tile sorted in shared memory, coalesced scatter --------------------------------
// The passes above store every pair straight from its register to its final place: 4-byte stores to (up to) 32 different
// sectors per warp instruction, which is what bounds them (ncu: 0.48 ms per pass at 2^25 pairs, 1.1 TB/s).  Here a tile
// is first sorted by its digit INSIDE shared memory (stable: position = digit start in the tile + pairs of the earlier
// warps + rank inside the warp), then written out in that order, so that consecutive threads write consecutive
// addresses of a digit's run (16 pairs = 64 B on average at 4096-pair tiles and 256 digits).  Ranks come from
// eight ballots per key (two keys in flight per thread).  LOOKBACK selects how a tile learns its global offsets:
// decoupled look-back over the status words (one launch per pass) or the per-tile counts of ts2_hist + ts_scan.
constexpr int OS2_ITEMS = 16;
constexpr int OS2_TILE = OS_THREADS * OS2_ITEMS;   // 4096 pairs per tile

constexpr int OS2_RUNS = 2 * OS_WARPS;

struct Os2Smem {
    // digit counts per HALF warp-tile (a warp's 512 pairs are ranked as two runs of 256 with their own counters, so two
    // independent match / count / update chains are in flight per thread), then the run's first local position of a digit
    uint32_t wcnt[OS2_RUNS][OS_RADIX];
    uint32_t delta[OS_RADIX];            // global position of a pair = delta[digit] + its position in the sorted tile
    uint32_t skey[OS2_TILE], sval[OS2_TILE];
    uint32_t warp_sum[OS_WARPS], warp_sum2[OS_WARPS];
    int tile;
};

__global__ void __launch_bounds__(OS_THREADS)
ts2_hist(const uint32_t *__restrict__ keys, int n, int shift, uint32_t mask, uint32_t *__restrict__ tile_cnt,
         int ntiles) {
    __shared__ uint32_t hist[OS_RADIX];
    const int tid = threadIdx.x;
    hist[tid] = 0;
    __syncthreads();
    const int base = blockIdx.x * OS2_TILE;
#pragma unroll
    for (int k = 0; k < OS2_ITEMS; ++k) {
        const int i = base + k * OS_THREADS + tid;
        if (i < n) atomicAdd(&hist[(keys[i] >> shift) & mask], 1u);
    }
    __syncthreads();
    tile_cnt[(size_t)tid * ntiles + blockIdx.x] = hist[tid];
}

#ifndef SPH_OS2_CTAS
#define SPH_OS2_CTAS 3
#endif
template <bool LOOKBACK>
__global__ void __launch_bounds__(OS_THREADS, SPH_OS2_CTAS)
os2_pass(const uint32_t *__restrict__ kin, const uint32_t *__restrict__ vin, uint32_t *__restrict__ kout,
         uint32_t *__restrict__ vout, int n, int pass, int n_passes, int shift, uint32_t mask, int ntiles,
         uint32_t *__restrict__ ctrl, const uint32_t *__restrict__ tile_off) {
    extern __shared__ __align__(16) unsigned char os2_raw[];
    Os2Smem &sm = *reinterpret_cast<Os2Smem *>(os2_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (LOOKBACK && tid == 0) sm.tile = (int)atomicAdd(&ctrl[n_passes * OS_RADIX + pass], 1u);
#pragma unroll
    for (int w = 0; w < OS2_RUNS; ++w) sm.wcnt[w][tid] = 0;
    __syncthreads();
    const int tile = LOOKBACK ? sm.tile : (int)blockIdx.x;
    const int nvalid = min(OS2_TILE, n - tile * OS2_TILE);

    // the warp's 512 pairs, 32 at a time in input order: stable rank among the warp's pairs of the same digit
    const int wbase = tile * OS2_TILE + warp * (32 * OS2_ITEMS);
    uint32_t key[OS2_ITEMS], val[OS2_ITEMS], rank2[OS2_ITEMS / 2];   // two 16-bit ranks per register
    const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
    for (int k = 0; k < OS2_ITEMS; ++k) {
        const int i = wbase + k * 32 + lane;
        key[k] = (i < n) ? kin[i] : 0u;
    }
#pragma unroll
    for (int k = 0; k < OS2_ITEMS; ++k) {
        const int i = wbase + k * 32 + lane;
        val[k] = (i < n) ? (vin ? vin[i] : (uint32_t)i) : 0u;
    }
#pragma unroll
    for (int k = 0; k < OS2_ITEMS / 2; ++k) {   // items k (first run) and k + 8 (second run) together
        const int ia = wbase + k * 32 + lane, ib = ia + (OS2_ITEMS / 2) * 32;
        const bool va = ia < n, vb = ib < n;
        const uint32_t da = (key[k] >> shift) & mask, db = (key[k + OS2_ITEMS / 2] >> shift) & mask;
#ifndef SPH_OS2_MATCH
        // eight ballots per key; match.any measured 22 % slower per pass on sm_100a (profiles/r2/ab_sort.txt)
        const uint32_t pa = os_match(da, va), pb = os_match(db, vb);
#else
        const uint32_t pa = __match_any_sync(0xffffffffu, va ? da : (0x10000u + lane));
        const uint32_t pb = __match_any_sync(0xffffffffu, vb ? db : (0x10000u + lane));
#endif
        uint32_t prea = 0, preb = 0;
        if (va) prea = sm.wcnt[2 * warp][da];
        if (vb) preb = sm.wcnt[2 * warp + 1][db];
        __syncwarp();
        if (va && (pa & lt) == 0) sm.wcnt[2 * warp][da] = prea + __popc(pa);
        if (vb && (pb & lt) == 0) sm.wcnt[2 * warp + 1][db] = preb + __popc(pb);
        __syncwarp();
        const uint32_t ra = prea + __popc(pa & lt), rb = preb + __popc(pb & lt);
        rank2[k] = ra | (rb << 16);   // low half: item k, high half: item k + 8
    }
    __syncthreads();

    // digit `tid`: count in this tile, start inside the sorted tile, global offset
    uint32_t tile_cnt = 0;
#pragma unroll
    for (int w = 0; w < OS2_RUNS; ++w) tile_cnt += sm.wcnt[w][tid];
    uint32_t gbase;   // global position of the tile's first pair of digit `tid`
    if (LOOKBACK) {
        volatile uint32_t *status = ctrl + (size_t)n_passes * OS_RADIX + 8 + ((size_t)pass * ntiles) * OS_RADIX;
        uint32_t excl = 0;
        if (tile == 0) {
            status[tid] = OS_FLAG_PREFIX | tile_cnt;
        } else {
            status[(size_t)tile * OS_RADIX + tid] = OS_FLAG_AGG | tile_cnt;
            // (looking back four predecessors per trip was measured: +7 % per pass -- the chains are short)
            int prev = tile - 1;
            while (true) {
                const uint32_t s = status[(size_t)prev * OS_RADIX + tid];
                const uint32_t flag = s & ~OS_VALUE_MASK;
                if (flag == 0) continue;   // predecessor has a ticket, hence is resident or done: it will publish
                excl += s & OS_VALUE_MASK;
                if (flag == OS_FLAG_PREFIX) break;
                --prev;
            }
            status[(size_t)tile * OS_RADIX + tid] = OS_FLAG_PREFIX | (excl + tile_cnt);
        }
        const uint32_t tot = ctrl[pass * OS_RADIX + tid];   // exclusive scan of the 256 global digit totals
        uint32_t inc = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += u;
        }
        if (lane == 31) sm.warp_sum[warp] = inc;
        gbase = inc - tot + excl;
    } else {
        gbase = tile_off[(size_t)tid * ntiles + tile];
    }
    uint32_t linc = tile_cnt;   // exclusive scan of the tile's digit counts
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t u = __shfl_up_sync(0xffffffffu, linc, o);
        if (lane >= o) linc += u;
    }
    if (lane == 31) sm.warp_sum2[warp] = linc;
    __syncthreads();
    uint32_t lstart = linc - tile_cnt;
#pragma unroll
    for (int w = 0; w < OS_WARPS; ++w) {
        if (w < warp) {
            lstart += sm.warp_sum2[w];
            if (LOOKBACK) gbase += sm.warp_sum[w];
        }
    }
    sm.delta[tid] = gbase - lstart;
    uint32_t running = lstart;   // per-run first positions (run order == input order)
#pragma unroll
    for (int w = 0; w < OS2_RUNS; ++w) {
        const uint32_t c = sm.wcnt[w][tid];
        sm.wcnt[w][tid] = running;
        running += c;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < OS2_ITEMS; ++k) {
        const int i = wbase + k * 32 + lane;
        if (i < n) {
            const uint32_t d = (key[k] >> shift) & mask;
            const int h = k / (OS2_ITEMS / 2), kk = k % (OS2_ITEMS / 2);
            const uint32_t p = sm.wcnt[2 * warp + h][d] + ((rank2[kk] >> (h * 16)) & 0xffffu);
            sm.skey[p] = key[k];
            sm.sval[p] = val[k];
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < OS2_ITEMS; ++k) {
        const int p = k * OS_THREADS + tid;
        if (p < nvalid) {
            const uint32_t kk = sm.skey[p];
            const uint32_t dst = sm.delta[(kk >> shift) & mask] + (uint32_t)p;
            kout[dst] = kk;
            vout[dst] = sm.sval[p];
        }
    }
}

}  // namespace sph
