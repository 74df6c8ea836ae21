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

}  // namespace sph
