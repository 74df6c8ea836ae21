// FIRST-GENERATION sort (SPH_SORT=classic); the engine default is radix_onesweep.cuh.
// Stable LSD radix sort of (cell key, particle index) pairs -- the device replacement of the reference's host-side
// numpy structured sort (voxel_sph_strategy.py:81-88).  Stability + values starting as iota give exactly numpy's
// (voxel_id, particle_id) order.
//
// Three launches per digit pass, no inter-block spinning (so nothing can hang):
//   rs_hist    per-tile digit histogram            -> block_hist[digit][tile]
//   rs_scan    one CTA per digit: exclusive scan over tiles, digit totals
//   rs_scatter stable in-tile ranking with warp match + scatter
// Tiles are 4096 pairs (256 threads x 16); a warp owns 512 consecutive pairs and visits them 32 at a time, so
// (warp, round, lane) order == input order, which is what makes the ranking stable.
#pragma once
#include "sph_common.cuh"

namespace sph {

constexpr int RS_THREADS = 256;
constexpr int RS_ITEMS = 16;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_RADIX = 256;

__device__ __forceinline__ uint32_t lanemask_lt() {
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

__global__ void __launch_bounds__(RS_THREADS)
rs_hist(const uint32_t *__restrict__ keys, int n, int shift, uint32_t mask, uint32_t *__restrict__ block_hist,
        int ntiles) {
    __shared__ uint32_t hist[RS_RADIX];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    hist[tid] = 0;
    __syncthreads();
    const int wbase = blockIdx.x * RS_TILE + warp * (32 * RS_ITEMS);
#pragma unroll
    for (int k = 0; k < RS_ITEMS; ++k) {
        const int i = wbase + k * 32 + lane;
        const bool valid = i < n;
        const uint32_t d = valid ? ((keys[i] >> shift) & mask) : RS_RADIX;
        const uint32_t peers = __match_any_sync(0xffffffffu, d);
        if (valid && (peers & lanemask_lt()) == 0) atomicAdd(&hist[d], (uint32_t)__popc(peers));
    }
    __syncthreads();
    block_hist[(size_t)tid * ntiles + blockIdx.x] = hist[tid];
}

// grid = RS_RADIX CTAs; CTA d turns block_hist[d][0..ntiles) into its exclusive prefix and writes the digit total.
__global__ void __launch_bounds__(RS_THREADS)
rs_scan(uint32_t *__restrict__ block_hist, int ntiles, uint32_t *__restrict__ digit_total) {
    __shared__ uint32_t warp_sum[RS_WARPS];
    __shared__ uint32_t carry_s;
    uint32_t *row = block_hist + (size_t)blockIdx.x * ntiles;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < ntiles; base += RS_THREADS) {
        const int i = base + tid;
        const uint32_t v = (i < ntiles) ? row[i] : 0;
        uint32_t inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += u;
        }
        if (lane == 31) warp_sum[warp] = inc;
        __syncthreads();
        uint32_t woff = 0;
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w)
            if (w < warp) woff += warp_sum[w];
        const uint32_t carry = carry_s;
        if (i < ntiles) row[i] = carry + woff + inc - v;
        __syncthreads();
        if (tid == RS_THREADS - 1) carry_s = carry + woff + inc;
        __syncthreads();
    }
    if (tid == 0) digit_total[blockIdx.x] = carry_s;
}

// vin == nullptr means "values are iota" (first pass).
__global__ void __launch_bounds__(RS_THREADS)
rs_scatter(const uint32_t *__restrict__ kin, const uint32_t *__restrict__ vin, uint32_t *__restrict__ kout,
           uint32_t *__restrict__ vout, int n, int shift, uint32_t mask, const uint32_t *__restrict__ block_hist,
           int ntiles, const uint32_t *__restrict__ digit_total) {
    __shared__ uint32_t wcnt[RS_WARPS][RS_RADIX];
    __shared__ uint32_t warp_sum[RS_WARPS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int w = 0; w < RS_WARPS; ++w) wcnt[w][tid] = 0;

    // exclusive scan of the 256 digit totals -> global base of digit `tid`
    const uint32_t tot = digit_total[tid];
    uint32_t inc = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += u;
    }
    if (lane == 31) warp_sum[warp] = inc;
    __syncthreads();
    uint32_t dbase = inc - tot;
#pragma unroll
    for (int w = 0; w < RS_WARPS; ++w)
        if (w < warp) dbase += warp_sum[w];
    dbase += block_hist[(size_t)tid * ntiles + blockIdx.x];

    // stable ranks inside the warp's 512 pairs
    const int wbase = blockIdx.x * RS_TILE + warp * (32 * RS_ITEMS);
    uint32_t key[RS_ITEMS];
    uint32_t rank[RS_ITEMS];
    const uint32_t lt = lanemask_lt();
#pragma unroll
    for (int k = 0; k < RS_ITEMS; ++k) {
        const int i = wbase + k * 32 + lane;
        const bool valid = i < n;
        key[k] = valid ? kin[i] : 0u;
        const uint32_t d = valid ? ((key[k] >> shift) & mask) : RS_RADIX;
        const uint32_t peers = __match_any_sync(0xffffffffu, d);
        uint32_t pre = 0;
        if (valid) pre = wcnt[warp][d];
        __syncwarp();
        if (valid && (peers & lt) == 0) wcnt[warp][d] = pre + __popc(peers);
        __syncwarp();
        rank[k] = pre + __popc(peers & lt);
    }
    __syncthreads();
    // digit `tid`: turn per-warp counts into global start offsets (warp order == input order)
    uint32_t running = dbase;
#pragma unroll
    for (int w = 0; w < RS_WARPS; ++w) {
        const uint32_t c = wcnt[w][tid];
        wcnt[w][tid] = running;
        running += c;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < RS_ITEMS; ++k) {
        const int i = wbase + k * 32 + lane;
        if (i < n) {
            const uint32_t d = (key[k] >> shift) & mask;
            const uint32_t dst = wcnt[warp][d] + rank[k];
            kout[dst] = key[k];
            vout[dst] = vin ? vin[i] : (uint32_t)i;
        }
    }
}

}  // namespace sph
