// SPH step kernels (sm_100a) other than the neighbour sweeps (sweep.cuh): cell hashing, cell table + SoA reorder,
// the fp64 <-> fp32 state converters at the host boundary and the parity-tap helpers.
//
// Data layout in HBM:
//   master  one 32-byte record per particle, particle-id order: float4 (x, y, z, density) | float4 (vx, vy, vz, 0).
//           pos_m points at record 0, vel_m = pos_m + 1, and particle i is pos_m[MI(i)] / vel_m[MI(i)] (MI(i) = 2 i):
//           the random gather of the reorder and the random scatter of the force epilogue then touch ONE 32-byte sector
//           per particle (two arrays cost two sectors = two 64-byte DRAM atoms: ncu, reorder_kernel at 77 % DRAM)
//   sorted  spos[N]  float4 (x, y, z, -),       svel[N]  float4, srho[N] float        -- cell-contiguous order
//   cells   cell_range[ncells + 1] int2 (begin, end) into the sorted arrays; entry ncells is the dead cell
#pragma once
#include "collide.cuh"
#include "sph_common.cuh"

namespace sph {

// ---- assign_voxels_to_particles_kernel (voxel_kernels.py:88-105) ------------------------------------------------
__global__ void __launch_bounds__(256)
hash_kernel(const float4 *__restrict__ pos_m, uint32_t *__restrict__ keys, int n, GridDesc g) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = pos_m[MI(i)];
    keys[i] = key_of(g, p.x, p.y, p.z);
}

// ---- __populate_voxel_begins (voxel_sph_strategy.py:98-107) + gather into cell-contiguous SoA ---------------------
// cell_range must be zeroed before the launch: an empty cell keeps (0, 0).
template <bool WITH_VEL>
__global__ void __launch_bounds__(256)
reorder_kernel(const uint32_t *__restrict__ skeys, const uint32_t *__restrict__ sids,
               const float4 *__restrict__ pos_m, const float4 *__restrict__ vel_m, float4 *__restrict__ spos,
               float4 *__restrict__ svel, int2 *__restrict__ cell_range, int n, int *__restrict__ n_items) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    if (t == 0) n_items[0] = n_items[1] = n_items[2] = 0;   // work-item counters of the sweeps (rows_plan_kernel / density_flat_kernel)
    const uint32_t key = skeys[t];
    const uint32_t id = sids[t];
    if (t == 0) {
        cell_range[key].x = 0;
    } else {
        const uint32_t prev = skeys[t - 1];
        if (prev != key) {
            cell_range[key].x = t;
            cell_range[prev].y = t;
        }
    }
    if (t == n - 1) cell_range[key].y = n;
    spos[t] = pos_m[MI(id)];
    if (WITH_VEL) svel[t] = vel_m[MI(id)];
}

// Velocity half of the reorder, for callers whose velocities arrive late (sph_compute_next_state uploads them while the
// density sweep runs): x, y, z only -- by then .w holds the pair factor the density sweep left there.
__global__ void __launch_bounds__(256)
gather_vel_kernel(const uint32_t *__restrict__ sids, const float4 *__restrict__ vel_m, float4 *__restrict__ svel, int n) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const float4 v = vel_m[MI(sids[t])];
    float *o = reinterpret_cast<float *>(svel + t);
    o[0] = v.x;
    o[1] = v.y;
    o[2] = v.z;
}

// ---- x-slab mode only: make the order inside every cell ascending in GLOBAL particle id ------------------------------
// On one GPU the local index is the particle id, so the stable sort already yields the reference's (voxel_id,
// particle_id) order.  On a slab the local arrays hold owned particles in arrival order followed by ghosts, so the
// in-cell order is repaired here: rank = number of particles of the same cell with a smaller global id.
__global__ void __launch_bounds__(256)
cell_range_kernel(const uint32_t *__restrict__ skeys, int2 *__restrict__ cell_range, int n) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const uint32_t key = skeys[t];
    if (t == 0) {
        cell_range[key].x = 0;
    } else {
        const uint32_t prev = skeys[t - 1];
        if (prev != key) {
            cell_range[key].x = t;
            cell_range[prev].y = t;
        }
    }
    if (t == n - 1) cell_range[key].y = n;
}

__global__ void __launch_bounds__(256)
fix_order_kernel(const uint32_t *__restrict__ skeys, const uint32_t *__restrict__ sids_in,
                 uint32_t *__restrict__ sids_out, const int32_t *__restrict__ gid,
                 const int2 *__restrict__ cell_range, int n, uint32_t dead_key) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const bool in = t < n;
    const uint32_t key = in ? skeys[t] : dead_key;
    const uint32_t me = in ? sids_in[t] : 0u;
    const bool live = in && key != dead_key;
    const int32_t g = live ? gid[me] : 0;
    // A cell that lies inside this warp's 32 sorted positions is ranked by shuffles (one gather of the global id per
    // particle instead of one per pair); a cell that continues into a neighbouring warp takes the table walk below.
    const uint32_t up = __shfl_up_sync(0xffffffffu, key, 1);
    const unsigned heads = __ballot_sync(0xffffffffu, lane == 0 || up != key);
    const int start = 31 - __clz(heads & (0xffffffffu >> (31 - lane)));            // first lane of my cell
    const unsigned after = heads & ~(0xffffffffu >> (31 - lane));                     // heads behind me
    const int end = after ? __ffs(after) - 1 : 32;                                     // one past its last lane
    bool inside = live;
    if (live && start == 0 && t - lane > 0 && skeys[t - lane - 1] == key) inside = false;          // began in the previous warp
    if (live && end == 32 && t - lane + 32 < n && skeys[t - lane + 32] == key) inside = false;     // continues in the next
    int len = inside ? end - start : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) len = max(len, __shfl_xor_sync(0xffffffffu, len, o));
    int rank = 0;
    for (int k = 0; k < len; ++k) {
        const int src = start + k;
        const int32_t v = __shfl_sync(0xffffffffu, g, src & 31);
        if (inside && src < end) rank += (v < g) ? 1 : 0;
    }
    if (!in) return;
    if (!live) {   // dead cell (incl. every empty slot): nobody's candidate, keep the order
        sids_out[t] = me;
        return;
    }
    if (inside) {
        sids_out[t - (lane - start) + rank] = me;
        return;
    }
    const int2 r = cell_range[key];
    for (int u = r.x; u < r.y; ++u) rank += (gid[sids_in[u]] < g) ? 1 : 0;
    sids_out[r.x + rank] = me;
}

// ---- host boundary converters -------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
pack_state_kernel(const T *__restrict__ pos3, const T *__restrict__ vel3, float4 *__restrict__ pos_m,
                  float4 *__restrict__ vel_m, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    pos_m[MI(i)] = make_float4((float)pos3[3 * (size_t)i], (float)pos3[3 * (size_t)i + 1], (float)pos3[3 * (size_t)i + 2],
                           0.f);
    vel_m[MI(i)] = make_float4((float)vel3[3 * (size_t)i], (float)vel3[3 * (size_t)i + 1], (float)vel3[3 * (size_t)i + 2],
                           0.f);
}

// one (N,3) host-layout array -> float4 (w = 0)
template <typename T>
__global__ void __launch_bounds__(256)
pack_vec_kernel(const T *__restrict__ src3, float4 *__restrict__ dst, int n) {   // dst: pos_m or vel_m (record stride)
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    dst[MI(i)] = make_float4((float)src3[3 * (size_t)i], (float)src3[3 * (size_t)i + 1], (float)src3[3 * (size_t)i + 2], 0.f);
}

// sorted fp32 scalar (density) -> id-ordered fp64
__global__ void __launch_bounds__(256)
unsort_scalar_kernel(const float *__restrict__ sorted, const uint32_t *__restrict__ sids, double *__restrict__ out,
                     int n) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) out[sids[t]] = (double)sorted[t];
}

template <typename T>
__global__ void __launch_bounds__(256)
unpack_state_kernel(const float4 *__restrict__ pos_m, const float4 *__restrict__ vel_m, T *__restrict__ pos3,
                    T *__restrict__ vel3, T *__restrict__ rho, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = pos_m[MI(i)], v = vel_m[MI(i)];
    if (pos3) {
        pos3[3 * (size_t)i] = (T)p.x;
        pos3[3 * (size_t)i + 1] = (T)p.y;
        pos3[3 * (size_t)i + 2] = (T)p.z;
    }
    if (vel3) {
        vel3[3 * (size_t)i] = (T)v.x;
        vel3[3 * (size_t)i + 1] = (T)v.y;
        vel3[3 * (size_t)i + 2] = (T)v.z;
    }
    if (rho) rho[i] = (T)p.w;
}

// sorted float4 -> id-ordered (N,3) fp64 (result_force and the recorded terms)
__global__ void __launch_bounds__(256)
unsort_vec3_kernel(const float4 *__restrict__ sorted, const uint32_t *__restrict__ sids, double *__restrict__ out3,
                   int n) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const float4 f = sorted[t];
    const size_t id = sids[t];
    out3[3 * id] = f.x;
    out3[3 * id + 1] = f.y;
    out3[3 * id + 2] = f.z;
}

// sorted neighbour counts (low 7 bits of ncnt) -> id-ordered int32
__global__ void __launch_bounds__(256)
unsort_count_kernel(const uint8_t *__restrict__ sorted, const uint32_t *__restrict__ sids, int32_t *__restrict__ out,
                    int n) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) out[sids[t]] = sorted[t] & 0x7f;
}

// voxel_begin as the reference stores it: first map index of the cell, -1 if empty (voxel_sph_strategy.py:92-107)
__global__ void __launch_bounds__(256)
voxel_begin_kernel(const int2 *__restrict__ cell_range, int32_t *__restrict__ begin, int ncells) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncells) return;
    const int2 r = cell_range[c];
    begin[c] = (r.y > r.x) ? r.x : -1;
}

// analize.py-style reductions: [0] non-finite count, [1] max density bits, [2] max speed bits
__global__ void __launch_bounds__(256)
stats_kernel(const float4 *__restrict__ pos_m, const float4 *__restrict__ vel_m, int n, uint32_t *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t bad = 0;
    float rho = 0.f, sp = 0.f;
    if (i < n) {
        const float4 p = pos_m[MI(i)], v = vel_m[MI(i)];
        const bool fin = isfinite(p.x) && isfinite(p.y) && isfinite(p.z) && isfinite(v.x) && isfinite(v.y) &&
                         isfinite(v.z);
        bad = fin ? 0u : 1u;
        if (isfinite(p.w)) rho = fmaxf(p.w, 0.f);
        if (fin) sp = sqrtf(v.x * v.x + v.y * v.y + v.z * v.z);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        bad += __shfl_xor_sync(0xffffffffu, bad, o);
        rho = fmaxf(rho, __shfl_xor_sync(0xffffffffu, rho, o));
        sp = fmaxf(sp, __shfl_xor_sync(0xffffffffu, sp, o));
    }
    if ((threadIdx.x & 31) == 0) {
        if (bad) atomicAdd(&out[0], bad);
        atomicMax(&out[1], __float_as_uint(rho));
        atomicMax(&out[2], __float_as_uint(sp));
    }
}

// ---- frame export: fp64, id order, every stride-th particle (section 8(f)1) -----------------------------------------
__global__ void __launch_bounds__(256)
export_pack_kernel(const float4 *__restrict__ pos_m, const float4 *__restrict__ vel_m, double *__restrict__ out, int n_out,
                   int stride) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_out) return;
    const size_t i = (size_t)k * stride;
    const float4 p = pos_m[MI(i)], v = vel_m[MI(i)];
    double *pos3 = out, *vel3 = out + 3 * (size_t)n_out, *rho = out + 6 * (size_t)n_out;
    pos3[3 * (size_t)k] = p.x;
    pos3[3 * (size_t)k + 1] = p.y;
    pos3[3 * (size_t)k + 2] = p.z;
    vel3[3 * (size_t)k] = v.x;
    vel3[3 * (size_t)k + 1] = v.y;
    vel3[3 * (size_t)k + 2] = v.z;
    rho[k] = p.w;
}

// ---- seeded start states (section 8(f)2): counter-based draws, mirrored bit for bit by cuda_sph_b200/config.py -------
// draw `stream` of particle `id`: SplitMix64 finaliser of seed + golden * (16 id + stream + 1), top 24 bits -> [0, 1)
__host__ __device__ inline float gen_u24(uint64_t seed, uint64_t id, uint32_t stream) {
    uint64_t z = seed + 0x9E3779B97F4A7C15ULL * (id * 16ULL + stream + 1ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    z ^= z >> 31;
    return (float)(uint32_t)(z >> 40) * 5.9604644775390625e-08f;   // exact: 24 bits * 2^-24
}

struct GenArgs {
    int32_t kind;
    uint64_t seed;
    float ext[3];        // box kinds: extent of the filled region per dimension (fp32)
    float top[3];        // largest fp32 below the space size (positions stay inside the domain)
    const double *pipe;  // pipe kind: rows x 5 table
    int32_t pipe_rows;
};

__global__ void __launch_bounds__(256)
generate_kernel(float4 *__restrict__ pos_m, float4 *__restrict__ vel_m, int n, GenArgs ga) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float x, y, z, vx = 0.f, vy = 0.f, vz = 0.f;
    if (ga.kind != 2) {
        x = fminf(__fmul_rn(gen_u24(ga.seed, i, 0), ga.ext[0]), ga.top[0]);
        y = fminf(__fmul_rn(gen_u24(ga.seed, i, 1), ga.ext[1]), ga.top[1]);
        z = fminf(__fmul_rn(gen_u24(ga.seed, i, 2), ga.ext[2]), ga.top[2]);
        vx = __fadd_rn(__fadd_rn(gen_u24(ga.seed, i, 3), -0.5f), 1.5f);
        vy = __fadd_rn(__fadd_rn(gen_u24(ga.seed, i, 4), -0.5f), -5.0f);
        vz = __fadd_rn(__fadd_rn(gen_u24(ga.seed, i, 5), -0.5f), -5.0f);
    } else {
        // config.py:105-115: x uniform along the pipe; (y, z) uniform over 98 % of the local disc.  fp64 with explicit
        // roundings (no transcendental functions: the disc is sampled by rejection from the draw sequence)
        const int rows = ga.pipe_rows;
        const double len = ga.pipe[5 * (rows - 1)] - ga.pipe[0];
        const double xd = __dmul_rn((double)gen_u24(ga.seed, i, 0), len);
        int s = 0;
        while (s + 2 < rows && xd >= ga.pipe[5 * (s + 1)] - ga.pipe[0]) ++s;
        const double r0 = ga.pipe[5 * s + 3], r1 = ga.pipe[5 * (s + 1) + 3];
        const double frac = __ddiv_rn(__dsub_rn(xd, ga.pipe[5 * s] - ga.pipe[0]), ga.pipe[5 * s + 4]);
        const double rmax = __dmul_rn(__dadd_rn(r0, __dmul_rn(__dsub_rn(r1, r0), frac)), 0.98);
        double a = 0.0, b = 0.0;
        for (int tr = 0; tr < 7; ++tr) {
            const double ta = __dsub_rn(__dmul_rn((double)gen_u24(ga.seed, i, 1 + 2 * tr), 2.0), 1.0);
            const double tb = __dsub_rn(__dmul_rn((double)gen_u24(ga.seed, i, 2 + 2 * tr), 2.0), 1.0);
            if (__dadd_rn(__dmul_rn(ta, ta), __dmul_rn(tb, tb)) <= 1.0) {
                a = ta;
                b = tb;
                break;
            }
        }
        x = (float)__dadd_rn(xd, ga.pipe[0]);
        y = (float)__dadd_rn(ga.pipe[1], __dmul_rn(a, rmax));
        z = (float)__dadd_rn(ga.pipe[2], __dmul_rn(b, rmax));
    }
    pos_m[MI(i)] = make_float4(x, y, z, 0.f);
    vel_m[MI(i)] = make_float4(vx, vy, vz, 0.f);
}

// ---- per-frame reductions (section 8(f)4; analize.py:9-14) ----------------------------------------------------------
// out: [0] non-finite particles, [1] max position bits (ordered-int encoding), [2] min position, [3] max velocity
// component, [4] max speed, [5] max density, [6..38] neighbour-count histogram
__device__ __forceinline__ uint32_t ord_f32(float f) {   // monotone float -> uint mapping
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
// out[0..39] as above, out[40..41] = the dead cell's range at this moment (begin, end)
__global__ void frame_stats_init_kernel(uint32_t *__restrict__ out, const int2 *__restrict__ dead_range, int have_range) {
    const int k = threadIdx.x;
    if (k < 40) out[k] = (k == 2) ? 0xffffffffu : 0u;
    if (k == 40) out[40] = have_range ? (uint32_t)dead_range->x : 0u;
    if (k == 41) out[41] = have_range ? (uint32_t)dead_range->y : 0u;
}

__global__ void __launch_bounds__(256)
frame_stats_kernel(const float4 *__restrict__ pos_m, const float4 *__restrict__ vel_m, const uint8_t *__restrict__ ncnt,
                   int n, int have_counts, uint32_t *__restrict__ out) {
    __shared__ uint32_t hist[33];
    if (threadIdx.x < 33) hist[threadIdx.x] = 0;
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t bad = 0, pmax = 0, pmin = 0xffffffffu, vmax = 0, smax = 0, dmax = 0;
    if (i < n) {
        const float4 p = pos_m[MI(i)], v = vel_m[MI(i)];
        const bool fin = isfinite(p.x) && isfinite(p.y) && isfinite(p.z) && isfinite(v.x) && isfinite(v.y) &&
                         isfinite(v.z);
        bad = fin ? 0u : 1u;
        if (fin) {
            pmax = ord_f32(fmaxf(p.x, fmaxf(p.y, p.z)));
            pmin = ord_f32(fminf(p.x, fminf(p.y, p.z)));
            vmax = ord_f32(fmaxf(v.x, fmaxf(v.y, v.z)));
            smax = ord_f32(sqrtf(v.x * v.x + v.y * v.y + v.z * v.z));
        }
        if (isfinite(p.w)) dmax = ord_f32(p.w);
        if (have_counts) atomicAdd(&hist[min((int)(ncnt[i] & 0x7f), 32)], 1u);   // ncnt is in sorted order: a histogram does not care
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        bad += __shfl_xor_sync(0xffffffffu, bad, o);
        pmax = max(pmax, __shfl_xor_sync(0xffffffffu, pmax, o));
        pmin = min(pmin, __shfl_xor_sync(0xffffffffu, pmin, o));
        vmax = max(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
        smax = max(smax, __shfl_xor_sync(0xffffffffu, smax, o));
        dmax = max(dmax, __shfl_xor_sync(0xffffffffu, dmax, o));
    }
    if ((threadIdx.x & 31) == 0) {
        if (bad) atomicAdd(&out[0], bad);
        atomicMax(&out[1], pmax);
        atomicMin(&out[2], pmin);
        atomicMax(&out[3], vmax);
        atomicMax(&out[4], smax);
        atomicMax(&out[5], dmax);
    }
    __syncthreads();
    if (threadIdx.x < 33 && hist[threadIdx.x]) atomicAdd(&out[6 + threadIdx.x], hist[threadIdx.x]);
}

}  // namespace sph
