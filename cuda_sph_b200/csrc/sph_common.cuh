// Shared device-side types of the B200 SPH step engine.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sph {

constexpr int kMaxNeighbours = 32;  // MAX_NEIGHBOURS, reference config.py:30

// Index of particle i in the master arrays: one 32-byte record (position | velocity) per particle, pos_m and
// vel_m = pos_m + 1 both step by two float4 (layout note in sph_kernels.cuh).
__host__ __device__ __forceinline__ size_t MI(size_t i) { return 2 * i; }

// Uniform grid.  Keys are linearised with the CEIL dims (voxel_sph_strategy.py:70-73) while the neighbour walk bounds
// and linearises with the TRUNC dims (voxel_sph_strategy.py:110-116, voxel_kernels.py:56,60) -- reference quirk Q2,
// reproduced literally.  xoff shifts the x cell coordinate for slab-local tables (0 on a single GPU).
struct GridDesc {
    double voxel[3];
    int32_t xoff;        // first x cell column of the local table
    int32_t wk, hk;      // key strides: key = (vx - xoff) + vy * wk + vz * wk * hk
    int32_t wn, hn;      // neighbour-cell strides (trunc dims; == wk, hk when space is a multiple of voxel)
    int32_t tx, ty, tz;  // trunc dims: a neighbour cell must satisfy 0 <= c_d < t_d
    int32_t ncells;      // table size; key == ncells is the dead cell
    // x-slab decomposition (multi-GPU).  Single GPU: strict_x = 0, own_lo = INT_MIN/2, own_hi = INT_MAX/2.
    int32_t strict_x;    // 1: a particle whose x column falls outside the local table [xoff, xoff + wk) is dead
    int32_t own_lo, own_hi;  // owned x columns [own_lo, own_hi); density is also computed one column beyond
    int32_t aligned;     // 1: trunc dims == ceil dims (no quirk Q2), the row-staged sweeps apply; 0: every particle walks
    float inv_voxel[3];  // fp32 1 / voxel, only for the cheap "strictly inside its cell" pre-test
};

struct StepConsts {
    // fp32 sweep constants
    float h;
    float h2_lo, h2_hi;  // band around h^2 inside which the fp64 predicate decides
    float h2;
    float h2_sup;        // h^2 (1 + band) of density_flat_kernel's superset test (band covers its expanded-form rounding)
    float h2_near;       // above this r^2 the poly6 term (h^2 - r^2)^3 is evaluated from an fp64 r^2 (cancellation)
    float w_mass;        // W_CONST * MASS                       (config.py:27, voxel_kernels.py:132)
    float grad_c;        // GRAD_W_CONST                         (config.py:28)
    float lap_c;         // LAP_W_CONST                          (config.py:29)
    float k, rho0;       // K, RHO_0
    float mass_visc;     // MASS * VISC                          (voxel_kernels.py:208)
    // fp64 epilogue constants
    double r2_max;       // largest double r2 with sqrt(r2) <= INF_R (voxel_kernels.py:20-26)
    double h2_d;         // INF_R^2 in fp64
    double dt;
    double ext[3];
    double space[3];
    double damp;
    int32_t mode;        // 0 box, 1 pipe
    int32_t pipe_rows;
};

__device__ __forceinline__ bool cell_of(const GridDesc &g, float x, float y, float z, int &vx, int &vy, int &vz) {
    // v_d = int32(pos_d / voxel_size_d): fp64 division, C truncation (voxel_kernels.py:9-12)
    const double qx = (double)x / g.voxel[0], qy = (double)y / g.voxel[1], qz = (double)z / g.voxel[2];
    if (!(fabs(qx) < 2147483648.0) || !(fabs(qy) < 2147483648.0) || !(fabs(qz) < 2147483648.0)) return false;
    vx = (int)qx;
    vy = (int)qy;
    vz = (int)qz;
    return true;
}

__device__ __forceinline__ uint32_t key_of(const GridDesc &g, float x, float y, float z) {
    int vx, vy, vz;
    if (!cell_of(g, x, y, z, vx, vy, vz)) return (uint32_t)g.ncells;
    // x-slab mode: no x aliasing across slabs -- a column outside the domain or the local table is dead
    if (g.strict_x && (vx < 0 || vx >= g.tx || vx < g.xoff || vx >= g.xoff + g.wk)) return (uint32_t)g.ncells;
    const long long k = (long long)vx - g.xoff + (long long)vy * g.wk + (long long)vz * g.wk * g.hk;
    return (k >= 0 && k < g.ncells) ? (uint32_t)k : (uint32_t)g.ncells;
}

}  // namespace sph
