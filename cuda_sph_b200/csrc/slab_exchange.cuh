// x-slab exchange kernels (multi-GPU; no reference counterpart -- the reference is single-GPU).
//
// Slot layout of the master arrays on one rank (capacity = own_cap + ghost_cap):
//   [0, own_cap)         owned region: particles this rank integrates; may contain holes (gid = -1, x = NaN) where a
//                        particle emigrated, and unused slots above the high-water mark
//   [own_cap, capacity)  ghost region: copies of the neighbours' boundary columns, rebuilt every step
// Empty slots hash to the dead cell, sort to the tail and are skipped by every kernel, so a step always runs over the
// whole capacity: no particle count ever has to travel to the host and nothing in the step loop synchronises with it.
//
// Per step:  slab_route_kernel packs, for every owned particle, (a) a MIGRANT record for its new owner if its column
// left the slab and (b) GHOST records for the ranks whose two-column halo contains its column, into one send block
// per destination rank (a migrant region and a ghost region, counts in the block header);
// one fixed-size all_to_all (NCCL over NVLink; block sizes are static, adjacent ranks get large blocks, the others
// small ones for the rare long-distance migrant / the pipe outlet -> inlet recycle);  slab_unpack_kernel appends
// the migrants above the high-water mark and the ghosts into the ghost region.
#pragma once
#include "sph_common.cuh"

namespace sph {

constexpr int SLAB_MAX_WORLD = 16;
constexpr int SLAB_FLAG_BYTES = 256;   // head of the receive allocation: one arrival flag per source rank
constexpr int SLAB_HALO = 2;   // ghost columns per side (two: the density of first-column ghosts is recomputed locally)

struct SlabRoute {
    int32_t world, rank;
    int32_t bounds[SLAB_MAX_WORLD + 1];   // rank r owns columns [bounds[r], bounds[r + 1])
    int32_t n_cols;
    double voxel_x;
    int32_t own_cap, capacity;
    int32_t rec_bytes;                    // 32: (pos4, vel4 with gid in .w); 48: + xoroshiro state (PIPE)
    int32_t cap_m[SLAB_MAX_WORLD];        // migrant records per destination block
    int32_t cap_g[SLAB_MAX_WORLD];        // ghost records per destination block (stored after the migrants)
    int64_t peer_off[SLAB_MAX_WORLD + 1]; // byte offset of each block in the send / receive buffer;
                                          // block = 16-byte header (n_migrants, n_ghosts) + records
};

// counters (device): [0] high-water mark of the owned region, [1] ghosts of this step, [2] overflow flags,
// [3] scratch (compaction), [4] live owned particles (filled by slab_count_kernel)
constexpr int SLAB_HWM = 0, SLAB_NGHOST = 1, SLAB_OVERFLOW = 2, SLAB_SCRATCH = 3, SLAB_NLIVE = 4;

// int32(x / voxel) with fp64 division and C truncation, as the hash kernel does; -1 if not representable
__device__ __forceinline__ int slab_column(float x, double voxel_x) {
    const double q = (double)x / voxel_x;
    if (!(fabs(q) < 2147483648.0)) return -1;
    return (int)q;
}

__device__ __forceinline__ int slab_owner(const SlabRoute &r, int col) {
    col = min(max(col, 0), r.n_cols - 1);
    int o = 0;
    for (int k = 1; k < r.world; ++k) o += (col >= r.bounds[k]) ? 1 : 0;
    return o;
}

// One record into the block of destination `d` (all lanes of the warp call this; `emit` says which lanes have one).
// kind 0 = migrant (first region of the block), 1 = ghost (second region).  Warp-aggregated: one atomic per (warp, destination, kind).
__device__ __forceinline__ bool slab_emit(const SlabRoute &r, unsigned char *sendbuf, int32_t *counters, bool emit,
                                          int d, int kind, const float4 &p, const float4 &v, int gid,
                                          const uint64_t *rng) {
    const unsigned lane = threadIdx.x & 31;
    bool stored = false;   // false for a lane whose record did not fit its block (overflow: the caller keeps the particle)
    unsigned pending = __ballot_sync(0xffffffffu, emit);
    while (pending) {
        const int leader = __ffs(pending) - 1;
        const int dl = __shfl_sync(0xffffffffu, d, leader);
        const unsigned same = __ballot_sync(0xffffffffu, emit && d == dl) & pending;
        pending &= ~same;
        unsigned char *block = sendbuf + r.peer_off[dl];
        int32_t *hdr = reinterpret_cast<int32_t *>(block);
        int base = 0;
        if ((int)lane == leader) base = atomicAdd(&hdr[kind], __popc(same));
        base = __shfl_sync(0xffffffffu, base, leader);
        if ((same >> lane) & 1u) {
            const int mine = base + __popc(same & ((1u << lane) - 1u));
            const int cap = kind == 0 ? r.cap_m[dl] : r.cap_g[dl];
            if (mine >= cap) {
                atomicOr(&counters[SLAB_OVERFLOW], 1);
            } else {
                const int slot = kind == 0 ? mine : r.cap_m[dl] + mine;
                unsigned char *rec = block + 16 + (size_t)slot * r.rec_bytes;
                *reinterpret_cast<float4 *>(rec) = p;
                *reinterpret_cast<float4 *>(rec + 16) = make_float4(v.x, v.y, v.z, __int_as_float(gid));
                if (r.rec_bytes == 48) {
                    uint64_t s0 = 0, s1 = 0;
                    if (kind == 0 && rng) {
                        s0 = rng[2 * (size_t)gid];
                        s1 = rng[2 * (size_t)gid + 1];
                    }
                    *reinterpret_cast<ulonglong2 *>(rec + 32) = make_ulonglong2(s0, s1);
                }
                stored = true;
            }
        }
    }
    return stored;
}

__global__ void __launch_bounds__(256)
slab_route_kernel(SlabRoute r, float4 *__restrict__ pos_m, const float4 *__restrict__ vel_m, int32_t *__restrict__ gid,
                  const uint64_t *__restrict__ rng, unsigned char *__restrict__ sendbuf,
                  int32_t *__restrict__ counters) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool in = i < r.own_cap;
    const int g = in ? gid[i] : -1;
    const bool have = g >= 0;
    float4 p = make_float4(0.f, 0.f, 0.f, 0.f), v = p;
    int col = -1, o = r.rank;
    if (have) {
        p = pos_m[MI(i)];
        v = vel_m[MI(i)];
        col = slab_column(p.x, r.voxel_x);
        if (col >= 0) o = slab_owner(r, col);   // a particle without a column (non-finite x) stays where it is, dead
    }
    const bool leaves = have && o != r.rank;
    const int lo = r.bounds[o], hi = r.bounds[o + 1];
    const bool ghost_l = have && o > 0 && col >= lo && col < lo + SLAB_HALO;
    const bool ghost_r = have && o < r.world - 1 && col >= hi - SLAB_HALO && col < hi;
    const bool sent = slab_emit(r, sendbuf, counters, leaves, o, 0, p, v, g, rng);
    slab_emit(r, sendbuf, counters, ghost_l, o - 1, 1, p, v, g, nullptr);
    slab_emit(r, sendbuf, counters, ghost_r, o + 1, 1, p, v, g, nullptr);
    if (leaves && sent) {   // the slot becomes a hole -- only if the record left: on overflow the particle stays (flagged)
        gid[i] = -1;
        pos_m[MI(i)] = make_float4(__int_as_float(0x7fc00000), 0.f, 0.f, 0.f);
    }
}

// ---- exchange over peer memory ------------------------------------------------------------------------------------------
// With peer pointers (CUDA IPC handles of the receive buffers, sph_slab_open_peers) the step has no collective:
//   slab_route_cta_kernel  routes the freshly integrated owned particles: the records of a 1024-slot CTA are counted in
//                          shared memory, ONE global atomic per CTA and region reserves their run in the send block,
//                          the records are stored as contiguous runs (slab_route_kernel spends one global atomic per
//                          warp and region on four words of the same 128-byte line: 0.18 ms at 9 M slots, this 0.05);
//   slab_push_kernel       copies the USED part of every block into the block "from me" of the destination's receive
//                          buffer over NVLink (consecutive threads store consecutive 16-byte pieces: full lines);
//   slab_signal_wait_kernel publishes the record counts, raises a flag at every peer and waits for theirs.
// Receive buffers are double-buffered by exchange parity, so a fast rank's next push never writes what a slow rank has
// not unpacked yet.  Measured and rejected: emitting the records from the force sweep's epilogue (into the peers' memory
// or into the local send blocks): the reservation's atomic round trip sits on the critical path of 60 % of the sweep's
// CTAs, +0.23 ms on a 1.57 ms sweep, four times what the separate routing pass costs.
struct SlabEmit {
    SlabRoute r;
    unsigned char *src[SLAB_MAX_WORLD];   // my send block for rank d (local)
    unsigned char *dst[SLAB_MAX_WORLD];   // my block in rank d's receive buffer of this parity (d == rank: my own)
    int32_t *cnt;                         // [world][2] records emitted to rank d (migrants, ghosts); zeroed per step
    int32_t *counters;                    // slab counters (overflow flags)
    float inv_vx;                         // fp32 1 / voxel_x and the columns [in_lo, in_hi) of the slab that are in nobody's
    int32_t in_lo, in_hi;                 //   halo: a particle safely inside them needs no record (fast exit)
};

constexpr int ROUTE_CTA = 1024;
constexpr int PUSH_CTAS = 64;        // CTAs per destination of slab_push_kernel

__global__ void __launch_bounds__(ROUTE_CTA)
slab_route_cta_kernel(const __grid_constant__ SlabEmit em, float4 *__restrict__ pos_m,
                      const float4 *__restrict__ vel_m, int32_t *__restrict__ gid, const uint64_t *__restrict__ rng) {
    __shared__ int s_cnt[2 * SLAB_MAX_WORLD], s_base[2 * SLAB_MAX_WORLD];
    const SlabRoute &r = em.r;
    const unsigned lane = threadIdx.x & 31;
    if (threadIdx.x < 2 * SLAB_MAX_WORLD) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const int i = blockIdx.x * ROUTE_CTA + threadIdx.x;
    const int g = (i < r.own_cap) ? gid[i] : -1;
    const bool have = g >= 0;
    float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
    if (have) p = pos_m[MI(i)];
    // most particles sit in the interior of the slab: decide that without the fp64 division (the fp32 estimate of
    // x / voxel is off by < 1e-3 for |q| < 4096; whoever is near a column boundary takes the exact path)
    const float q = p.x * em.inv_vx, fl = floorf(q), fr = q - fl;
    const bool interior = fr > 1e-3f && fr < 1.f - 1e-3f && fabsf(q) < 4096.f && (int)fl >= em.in_lo && (int)fl < em.in_hi;
    int o = r.rank, col = -1;
    bool rec[3] = {false, false, false};   // migrant to o, ghost to o - 1, ghost to o + 1
    if (have && !interior) {
        col = slab_column(p.x, r.voxel_x);
        if (col >= 0) o = slab_owner(r, col);   // a particle without a column (non-finite x) stays where it is, dead
        const int lo = r.bounds[o], hi = r.bounds[o + 1];
        rec[0] = o != r.rank;
        rec[1] = o > 0 && col >= lo && col < lo + SLAB_HALO;
        rec[2] = o < r.world - 1 && col >= hi - SLAB_HALO && col < hi;
    }
    const int dest[3] = {o, o - 1, o + 1};
    int idx[3] = {-1, -1, -1};
    // CTA-local record indices.  The two common regions -- ghosts of particles that stay, for the left / right neighbour
    // -- are aggregated by ballot (one shared-memory atomic per warp); the rare rest (migrants and their ghosts) goes
    // lane by lane
    const unsigned lt = (1u << lane) - 1u;
#pragma unroll
    for (int u = 1; u < 3; ++u) {
        const bool common = rec[u] && o == r.rank;
        const unsigned m = __ballot_sync(0xffffffffu, common);
        if (m) {
            int base = 0;
            if (lane == 0) base = atomicAdd(&s_cnt[(r.rank + (u == 1 ? -1 : 1)) * 2 + 1], __popc(m));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (common) idx[u] = base + __popc(m & lt);
        }
    }
#pragma unroll
    for (int u = 0; u < 3; ++u)
        if (rec[u] && idx[u] < 0) idx[u] = atomicAdd(&s_cnt[dest[u] * 2 + (u ? 1 : 0)], 1);
    __syncthreads();
    if (threadIdx.x < 2 * r.world && s_cnt[threadIdx.x] > 0)
        s_base[threadIdx.x] = atomicAdd(&em.cnt[threadIdx.x], s_cnt[threadIdx.x]);
    __syncthreads();
    if (rec[0] || rec[1] || rec[2]) {
        const float4 v = vel_m[MI(i)];
        bool kept = false;   // migrant record did not fit: the particle stays with this rank (and the overflow flag is up)
#pragma unroll
        for (int u = 0; u < 3; ++u) {
            if (!rec[u]) continue;
            const int d = dest[u], kind = u ? 1 : 0;
            const int mine = s_base[d * 2 + kind] + idx[u];
            const int cap = kind == 0 ? r.cap_m[d] : r.cap_g[d];
            if (mine >= cap) {
                atomicOr(&em.counters[SLAB_OVERFLOW], 1);
                if (u == 0) kept = true;
                continue;
            }
            const int slot = kind == 0 ? mine : r.cap_m[d] + mine;
            unsigned char *out = em.src[d] + 16 + (size_t)slot * r.rec_bytes;
            *reinterpret_cast<float4 *>(out) = p;
            *reinterpret_cast<float4 *>(out + 16) = make_float4(v.x, v.y, v.z, __int_as_float(g));
            if (r.rec_bytes == 48) {
                uint64_t s0 = 0, s1 = 0;
                if (kind == 0 && rng) {
                    s0 = rng[2 * (size_t)g];
                    s1 = rng[2 * (size_t)g + 1];
                }
                *reinterpret_cast<ulonglong2 *>(out + 32) = make_ulonglong2(s0, s1);
            }
        }
        if (rec[0] && !kept) {   // the slot becomes a hole
            gid[i] = -1;
            pos_m[MI(i)] = make_float4(__int_as_float(0x7fc00000), 0.f, 0.f, 0.f);
        }
    }
}

// Used records of every send block -> the receivers' memory.  grid.y = destination rank; consecutive threads store
// consecutive 16-byte pieces (full 128-byte lines on the NVLink side).
__global__ void __launch_bounds__(256)
slab_push_kernel(const __grid_constant__ SlabEmit em) {
    const int d = blockIdx.y;
    const int rec16 = em.r.rec_bytes / 16;
    const long long n_m = (long long)min(em.cnt[d * 2], em.r.cap_m[d]) * rec16;
    const long long n_g = (long long)min(em.cnt[d * 2 + 1], em.r.cap_g[d]) * rec16;
    const uint4 *src = reinterpret_cast<const uint4 *>(em.src[d]) + 1;   // behind the 16-byte header
    uint4 *dst = reinterpret_cast<uint4 *>(em.dst[d]) + 1;
    const long long g0 = (long long)em.r.cap_m[d] * rec16;               // ghost region of the block
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_m + n_g;
         i += (long long)gridDim.x * blockDim.x) {
        const long long k = i < n_m ? i : g0 + (i - n_m);
        dst[k] = src[k];
    }
}

// Exchange epilogue of a step: thread d publishes the counts of my block at rank d, raises my flag there (release at
// system scope: the records were stored by the force sweep earlier in this stream) and waits for rank d's flag here.
// flags[s] of a rank = last epoch rank s has completed.  The wait is bounded: a peer that never arrives (crashed
// process) sets overflow bit 8 instead of hanging the GPU.
__global__ void slab_signal_wait_kernel(const SlabEmit *__restrict__ em, int32_t *const *__restrict__ peer_flags,
                                        volatile int32_t *my_flags, int epoch, long long timeout_cycles) {
    const int d = threadIdx.x;
    if (d >= em->r.world) return;
    int32_t *hdr = reinterpret_cast<int32_t *>(em->dst[d]);
    hdr[0] = em->cnt[d * 2];
    hdr[1] = em->cnt[d * 2 + 1];
    __threadfence_system();
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(peer_flags[d] + em->r.rank), "r"(epoch) : "memory");
    const long long t0 = clock64();
    for (;;) {
        int v;
        asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(my_flags + d) : "memory");
        if (v - epoch >= 0) break;
        if (clock64() - t0 > timeout_cycles) {
            atomicOr(&em->counters[SLAB_OVERFLOW], 8);
            break;
        }
    }
}

// ghost region -> empty
__global__ void __launch_bounds__(256)
slab_clear_ghosts_kernel(SlabRoute r, float4 *__restrict__ pos_m, int32_t *__restrict__ gid,
                         int32_t *__restrict__ counters) {
    const int i = r.own_cap + blockIdx.x * blockDim.x + threadIdx.x;
    if (i == r.own_cap) counters[SLAB_NGHOST] = 0;
    if (i >= r.capacity) return;
    if (gid[i] >= 0) {
        gid[i] = -1;
        pos_m[MI(i)] = make_float4(__int_as_float(0x7fc00000), 0.f, 0.f, 0.f);
    }
}

// grid.y = source rank.  Migrants are appended above the high-water mark, ghosts into the ghost region.
__global__ void __launch_bounds__(256)
slab_unpack_kernel(SlabRoute r, const unsigned char *__restrict__ recvbuf, float4 *__restrict__ pos_m,
                   float4 *__restrict__ vel_m, int32_t *__restrict__ gid, uint64_t *__restrict__ rng,
                   int32_t *__restrict__ counters) {
    const int src = blockIdx.y;
    const unsigned char *block = recvbuf + r.peer_off[src];
    const int32_t *hdr = reinterpret_cast<const int32_t *>(block);
    const int n_front = min(hdr[0], r.cap_m[src]), n_back = min(hdr[1], r.cap_g[src]);
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;   // over [0, n_front + n_back)
    if (blockIdx.x * blockDim.x >= n_front + n_back) return;
    const unsigned lane = threadIdx.x & 31;
    const bool valid = idx < n_front + n_back;
    const bool migrant = valid && idx < n_front;
    const bool ghost = valid && !migrant;
    const int slot_in = migrant ? idx : r.cap_m[src] + (idx - n_front);
    // aggregated appends
    const unsigned mm = __ballot_sync(0xffffffffu, migrant), gm = __ballot_sync(0xffffffffu, ghost);
    int mbase = 0, gbase = 0;
    if (lane == 0) {
        if (mm) mbase = atomicAdd(&counters[SLAB_HWM], __popc(mm));
        if (gm) gbase = atomicAdd(&counters[SLAB_NGHOST], __popc(gm));
    }
    mbase = __shfl_sync(0xffffffffu, mbase, 0);
    gbase = __shfl_sync(0xffffffffu, gbase, 0);
    if (!valid) return;
    const unsigned lt = (1u << lane) - 1u;
    int dst;
    if (migrant) {
        dst = mbase + __popc(mm & lt);
        if (dst >= r.own_cap) {
            atomicOr(&counters[SLAB_OVERFLOW], 2);
            return;
        }
    } else {
        dst = r.own_cap + gbase + __popc(gm & lt);
        if (dst >= r.capacity) {
            atomicOr(&counters[SLAB_OVERFLOW], 4);
            return;
        }
    }
    const unsigned char *rec = block + 16 + (size_t)slot_in * r.rec_bytes;
    const float4 p = *reinterpret_cast<const float4 *>(rec);
    const float4 v = *reinterpret_cast<const float4 *>(rec + 16);
    const int g = __float_as_int(v.w);
    pos_m[MI(dst)] = p;
    vel_m[MI(dst)] = make_float4(v.x, v.y, v.z, 0.f);
    gid[dst] = g;
    if (migrant && r.rec_bytes == 48 && rng) {
        const ulonglong2 s = *reinterpret_cast<const ulonglong2 *>(rec + 32);
        rng[2 * (size_t)g] = s.x;
        rng[2 * (size_t)g + 1] = s.y;
    }
}

// live owned particles -> counters[SLAB_NLIVE] (counters[SLAB_NLIVE] must be zero)
__global__ void __launch_bounds__(256)
slab_count_kernel(SlabRoute r, const int32_t *__restrict__ gid, int32_t *__restrict__ counters) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool have = i < r.own_cap && gid[i] >= 0;
    const unsigned m = __ballot_sync(0xffffffffu, have);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(&counters[SLAB_NLIVE], __popc(m));
}

// Compaction of the owned region (closes the holes): live particles -> tmp arrays at consecutive slots ...
__global__ void __launch_bounds__(256)
slab_compact_gather_kernel(SlabRoute r, const float4 *__restrict__ pos_m, const float4 *__restrict__ vel_m,
                           const int32_t *__restrict__ gid, float4 *__restrict__ tpos, float4 *__restrict__ tvel,
                           int32_t *__restrict__ tgid, int32_t *__restrict__ counters) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool have = i < r.own_cap && gid[i] >= 0;
    const unsigned m = __ballot_sync(0xffffffffu, have);
    const unsigned lane = threadIdx.x & 31;
    int base = 0;
    if (lane == 0 && m) base = atomicAdd(&counters[SLAB_SCRATCH], __popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (!have) return;
    const int dst = base + __popc(m & ((1u << lane) - 1u));
    tpos[dst] = pos_m[MI(i)];
    tvel[dst] = vel_m[MI(i)];
    tgid[dst] = gid[i];
}

// ... and back; slots above the new high-water mark become empty.  counters[SLAB_SCRATCH] holds the live count.
__global__ void __launch_bounds__(256)
slab_compact_scatter_kernel(SlabRoute r, float4 *__restrict__ pos_m, float4 *__restrict__ vel_m,
                            int32_t *__restrict__ gid, const float4 *__restrict__ tpos,
                            const float4 *__restrict__ tvel, const int32_t *__restrict__ tgid,
                            int32_t *__restrict__ counters) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= r.own_cap) return;
    const int n = counters[SLAB_SCRATCH];
    if (i < n) {
        pos_m[MI(i)] = tpos[i];
        vel_m[MI(i)] = tvel[i];
        gid[i] = tgid[i];
    } else {
        gid[i] = -1;
        pos_m[MI(i)] = make_float4(__int_as_float(0x7fc00000), 0.f, 0.f, 0.f);
    }
}

__global__ void slab_compact_finish_kernel(int32_t *__restrict__ counters) {
    counters[SLAB_HWM] = counters[SLAB_SCRATCH];
    counters[SLAB_SCRATCH] = 0;
}

}  // namespace sph
