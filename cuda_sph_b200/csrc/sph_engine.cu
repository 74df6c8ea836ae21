// libsph_b200.so -- engine object + C ABI (include/sph_b200.h).  One handle = one GPU = one stream.
//
// Step pipeline (device resident, replayed as a CUDA graph):
//   hash_kernel -> [rs_hist, rs_scan, rs_scatter] x passes -> memset(cell_range) -> reorder_kernel
//   -> density_kernel -> force_kernel (pressure + viscosity + integrate + collide, scatter to master)
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/sph_b200.h"
#include "radix_onesweep.cuh"
#include "radix_sort.cuh"
#include "slab_exchange.cuh"
#include "sph_kernels.cuh"
#include "sweep.cuh"
#include "sweep_dense.cuh"
#include "sweep_flat.cuh"
#include "sweep_rows.cuh"

using namespace sph;

static thread_local std::string g_err;
static int fail(const std::string &m) {
    g_err = m;
    return 1;
}
#define CK(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess)                                                                           \
            return fail(std::string(#call) + ": " + cudaGetErrorString(e_) + " (" __FILE__ ":" +         \
                        std::to_string(__LINE__) + ")");                                                 \
    } while (0)

struct SphEngine {
    SphParams p{};
    int device = 0;
    int n = 0;
    GridDesc grid{};
    StepConsts consts{};
    int32_t ceil_dims[3]{}, trunc_dims[3]{};
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaStream_t copy_stream = nullptr;   // second stream of sph_compute_next_state (created on first use)
    cudaEvent_t ev_copy[3]{};

    // master (id order) and sorted (cell order) state
    float4 *pos_m = nullptr, *vel_m = nullptr, *spos = nullptr, *svel = nullptr, *sforce = nullptr;
    float4 *spress = nullptr, *svisc = nullptr;
    float *srho = nullptr;
    uint16_t *nlist = nullptr;   // [n][32] neighbour lists (virtual-list indices), density -> force
    uint32_t *dlist = nullptr;   // [n][32] neighbour lists of the dense tiles (sorted indices; sweep_dense.cuh)
    uint8_t *ncnt = nullptr;     // [n] neighbour counts
    // keys / sort buffers
    uint32_t *keys = nullptr, *ka = nullptr, *va = nullptr, *kb = nullptr, *vb = nullptr;
    uint32_t *skeys = nullptr, *sids = nullptr;  // aliases of the final sort buffers
    uint32_t *block_hist = nullptr, *digit_total = nullptr;
    TilePlan *tile_plans = nullptr;   // one row plan per 128-particle tile of the sweeps (rows_plan_kernel)
    // work items of the sweeps (tile * 8 + (0: whole tile | 1 + pass)): [0] = count of list A (passes of tiles whose rows
    // do not fit; rows_plan_kernel), [1] = count of list B (whole tiles density_flat_kernel hands over), then the lists
    int *refused = nullptr;
    int ntiles_rb = 0;
    cudaStream_t aux_stream = nullptr;   // the work-item kernels run here, next to the main sweeps
    cudaEvent_t ev_fork[2]{}, ev_join[2]{};
    bool flat_density = true;         // SPH_DENSITY=rows: every tile through density_rows_kernel
    int dense_min_slots = RB_CAP * 3 / 2;   // non-fitting tiles with longer rows go to the dense kernels whatever their
                                          // cell count (measured: 2304 costs dam1m +40 %, 4096 costs pipe4m +60 %)
    uint32_t *os_ctrl = nullptr;      // onesweep control block (histograms, tickets, look-back status)
    bool onesweep = true;         // SPH_SORT=classic selects the three-kernel passes of radix_sort.cuh
    bool sort_gen2 = true;        // tiles sorted in shared memory + coalesced scatter, hash fused with the histograms (SPH_SORT=count|lookback: first generation)
    bool sort_lookback = true;   // SPH_SORT=lookback: decoupled look-back passes (one kernel per digit) instead of count + scan
    int ntiles = 0, passes = 0, pass_bits[8]{}, key_bits = 0;
    int2 *cell_range = nullptr;
    // pipe
    double *pipe_d = nullptr;
    int pipe_rows = 0;
    uint64_t *rng = nullptr;
    int64_t rng_count = 0;
    // x-slab mode (multi-GPU)
    bool slab = false;
    bool rows_sweeps = true;      // row-staged sweeps (sweep_rows.cuh); false: per-warp tiles (sweep.cuh), SPH_SWEEP=warp
    int32_t *gid = nullptr;       // global particle id per local index
    int32_t slab_lo = 0, slab_hi = 0;
    int64_t cell_capacity = 0;    // entries allocated in cell_range
    SlabRoute route{};            // native exchange (sph_slab_exchange_init)
    bool route_ready = false;
    unsigned char *sendbuf = nullptr, *recvbuf = nullptr;   // recvbuf: parity 0, parity 1 back to back (inside recv_alloc)
    unsigned char *recv_alloc = nullptr;   // [SLAB_FLAG_BYTES flags][receive buffer parity 0][parity 1]: the IPC-exported allocation
    size_t xchg_bytes = 0;                 // one receive buffer
    int parity = 0;                        // receive buffer the NEXT exchange fills
    // fused routing over peer memory (sph_slab_open_peers)
    bool p2p_ready = false;
    void *peer_base[SLAB_MAX_WORLD]{};     // recv_alloc of every rank, mapped here (CUDA IPC); [rank] = my own
    SlabEmit *emit_d = nullptr;            // [2] device copies, one per parity
    SlabEmit emit_h[2]{};                  // the same on the host (kernel parameters of the route / push kernels)
    int32_t **peer_flags_d = nullptr;      // [world] flag arrays of the peers
    int32_t *emit_cnt = nullptr;           // [world][2]
    int epoch = 0;
    int32_t *slab_counters = nullptr;   // SLAB_HWM, SLAB_NGHOST, SLAB_OVERFLOW, SLAB_SCRATCH, SLAB_NLIVE (+ pad)
    int32_t *tmp_gid = nullptr;         // compaction scratch
    // host-boundary staging (fp64 / fp32 (N,3) + rho), grown lazily
    void *stage = nullptr;
    size_t stage_bytes = 0;
    uint32_t *stats_d = nullptr;
    // frame export (sph_export_begin / sph_export_wait): device staging + pinned host buffer per slot
    double *exp_dev[3]{}, *exp_host[3]{};
    int64_t exp_cap[3]{}, exp_n[3]{};
    cudaEvent_t exp_ev[3]{}, exp_packed = nullptr;
    uint32_t *exp_stats_dev[3]{}, *exp_stats_host[3]{};   // frame statistics of the exported frame (40 words + dead range)
    int64_t exp_steps[3]{};
    uint32_t *fstats_d = nullptr;
    // snapshot (sph_save_state)
    float4 *snap_pos = nullptr, *snap_vel = nullptr;
    uint64_t *snap_rng = nullptr;
    int64_t snap_rng_count = 0;
    int64_t snap_steps = -1;
    // graph
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t graph_exec = nullptr;
    bool graph_valid = false;
    // events for sph_step_timed
    cudaEvent_t ev[8]{};
    bool has_state = false;
    int64_t steps_done = 0;
    int64_t launches = 0;
    int launches_per_step = 0;
};

const char *sph_last_error(void) { return g_err.c_str(); }
int sph_version(void) { return 1; }

static int ensure_stage(SphEngine *e, size_t bytes) {
    if (e->stage_bytes >= bytes) return 0;
    if (e->stage) cudaFree(e->stage);
    e->stage = nullptr;
    e->stage_bytes = 0;
    CK(cudaMalloc(&e->stage, bytes));
    e->stage_bytes = bytes;
    return 0;
}

static void invalidate_graph(SphEngine *e) {
    if (e->graph_exec) cudaGraphExecDestroy(e->graph_exec);
    if (e->graph) cudaGraphDestroy(e->graph);
    e->graph_exec = nullptr;
    e->graph = nullptr;
    e->graph_valid = false;
}

// numba create_xoroshiro128p_states: state 0 = SplitMix64(seed) in both words, state i = state i-1 jumped 2^64.
static void init_rng_host(std::vector<uint64_t> &st, int64_t n, uint64_t seed) {
    st.resize(2 * (size_t)n);
    if (n < 1) return;
    uint64_t z = seed + 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    z ^= z >> 31;
    uint64_t s0 = z, s1 = z;
    st[0] = s0;
    st[1] = s1;
    static const uint64_t J[2] = {0xbeac0467eba5facbULL, 0xd86b048b86aa9922ULL};
    for (int64_t i = 1; i < n; ++i) {
        uint64_t a = 0, b = 0;
        for (int w = 0; w < 2; ++w)
            for (int bit = 0; bit < 64; ++bit) {
                if (J[w] & (1ULL << bit)) { a ^= s0; b ^= s1; }
                xoro_next(s0, s1);
            }
        s0 = a;
        s1 = b;
        st[2 * i] = s0;
        st[2 * i + 1] = s1;
    }
}

static int validate(const SphParams *p) {
    if (!p) return fail("params is NULL");
    if (p->particle_count <= 0) return fail("particle_count must be > 0");
    if (p->max_neighbours != kMaxNeighbours) return fail("only max_neighbours == 32 is supported");
    if (p->mode != SPH_MODE_BOX && p->mode != SPH_MODE_PIPE) return fail("mode must be SPH_MODE_BOX or SPH_MODE_PIPE");
    if (!(p->h > 0) || !(p->dt > 0)) return fail("h and dt must be > 0");
    for (int d = 0; d < 3; ++d)
        if (!(p->voxel_size[d] > 0) || !(p->space_size[d] > 0)) return fail("space_size / voxel_size must be > 0");
    return 0;
}

int sph_create(const SphParams *params, int device, sph_handle_t *out) {
    if (!out) return fail("out is NULL");
    *out = nullptr;
    if (validate(params)) return 1;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail("no CUDA device: libsph_b200 has no CPU fallback");
    if (device < 0 || device >= ndev) return fail("bad device index");
    CK(cudaSetDevice(device));

    SphEngine *e = new SphEngine();
    e->p = *params;
    e->device = device;
    e->n = params->particle_count;
    const int n = e->n;

    // grid: ceil dims for keys, trunc dims for the neighbour walk (reference quirk Q2)
    long long ncells = 1;
    for (int d = 0; d < 3; ++d) {
        const double q = params->space_size[d] / params->voxel_size[d];
        e->ceil_dims[d] = (int32_t)std::ceil(q);
        e->trunc_dims[d] = (int32_t)q;
        e->grid.voxel[d] = params->voxel_size[d];
        ncells *= e->ceil_dims[d];
    }
    if (ncells <= 0 || ncells >= (1LL << 31) - 2) {
        delete e;
        return fail("cell table too large for int32 keys");
    }
    e->grid.xoff = 0;
    e->grid.wk = e->ceil_dims[0];
    e->grid.hk = e->ceil_dims[1];
    e->grid.wn = e->trunc_dims[0];
    e->grid.hn = e->trunc_dims[1];
    e->grid.tx = e->trunc_dims[0];
    e->grid.ty = e->trunc_dims[1];
    e->grid.tz = e->trunc_dims[2];
    e->grid.ncells = (int32_t)ncells;
    e->grid.strict_x = 0;
    e->grid.own_lo = INT32_MIN / 2;
    e->grid.own_hi = INT32_MAX / 2;
    e->grid.aligned = 1;
    for (int d = 0; d < 3; ++d) {
        if (e->ceil_dims[d] != e->trunc_dims[d]) e->grid.aligned = 0;
        e->grid.inv_voxel[d] = (float)(1.0 / params->voxel_size[d]);
    }
    if (const char *sw = getenv("SPH_SWEEP")) e->rows_sweeps = strcmp(sw, "warp") != 0;
    if (const char *sd = getenv("SPH_DENSITY")) e->flat_density = strcmp(sd, "rows") != 0;
    if (const char *sd = getenv("SPH_DENSE_MIN_SLOTS")) e->dense_min_slots = atoi(sd);
    e->slab = (params->flags & SPH_FLAG_SLAB) != 0;
    long long table_cells = ncells;
    if (e->slab) {
        for (int d = 0; d < 3; ++d)
            if (e->ceil_dims[d] != e->trunc_dims[d]) {
                delete e;
                return fail("x-slab mode needs space_size to be a multiple of voxel_size");
            }
        table_cells = (long long)(e->ceil_dims[0] + 4) * e->ceil_dims[1] * e->ceil_dims[2];
        if (table_cells >= (1LL << 31) - 2) {
            delete e;
            return fail("cell table too large for int32 keys");
        }
    }
    e->cell_capacity = table_cells + 1;

    // constants (config.py:24-29)
    const double h = params->h;
    StepConsts &c = e->consts;
    c.h = (float)h;
    c.h2 = (float)(h * h);
    c.h2_near = (float)(h * h * 0.98);
    c.h2_d = h * h;
    c.h2_lo = (float)(h * h * (1.0 - 1e-5));
    c.h2_hi = (float)(h * h * (1.0 + 1e-5));
    {   // superset band of density_flat_kernel: 16 x 2^-24 x (sum of the magnitudes its expanded r^2 adds up), >= 2e-5
        const double *v = params->voxel_size;
        const double mag = 25.0 * (v[1] * v[1] + v[2] * v[2]) + 4.0 * v[0] * v[0] + h * h;
        const double band = std::max(2e-5, 16.0 * std::ldexp(1.0, -24) * mag / (h * h));
        c.h2_sup = (float)(h * h * (1.0 + band));
    }
    c.w_mass = (float)(315.0 / (64.0 * M_PI * std::pow(h, 9.0)) * params->mass);
    c.grad_c = (float)(-45.0 / (M_PI * std::pow(h, 6.0)));
    c.lap_c = (float)(45.0 / (M_PI * std::pow(h, 6.0)));
    c.k = (float)params->k;
    c.rho0 = (float)params->rho0;
    c.mass_visc = (float)(params->mass * params->visc);
    // largest double r2 with sqrt(r2) <= h  (sqrt is correctly rounded and monotone)
    double r2m = h * h;
    while (std::sqrt(std::nextafter(r2m, INFINITY)) <= h) r2m = std::nextafter(r2m, INFINITY);
    while (std::sqrt(r2m) > h) r2m = std::nextafter(r2m, -INFINITY);
    c.r2_max = r2m;
    c.dt = params->dt;
    c.damp = params->damp;
    for (int d = 0; d < 3; ++d) {
        c.ext[d] = params->external_force[d];
        c.space[d] = params->space_size[d];
    }
    c.mode = params->mode;
    c.pipe_rows = 0;

    // radix passes over the bits of [0, ncells] (ncells itself = dead cell)
    int bits = 1;
    while ((1LL << bits) <= table_cells) ++bits;
    e->key_bits = bits;
    e->passes = (bits + 7) / 8;
    for (int i = 0; i < e->passes; ++i) e->pass_bits[i] = bits / e->passes + (i < bits % e->passes ? 1 : 0);
    e->ntiles = (n + RS_TILE - 1) / RS_TILE;

#define ALLOC(ptr, count)                                                                   \
    do {                                                                                    \
        cudaError_t e_ = cudaMalloc((void **)&(ptr), sizeof(*(ptr)) * (size_t)(count));     \
        if (e_ != cudaSuccess) {                                                            \
            fail(std::string("cudaMalloc " #ptr ": ") + cudaGetErrorString(e_));            \
            sph_destroy(e);                                                                 \
            return 1;                                                                       \
        }                                                                                   \
    } while (0)
    ALLOC(e->pos_m, 2 * (size_t)n);   // 32-byte records: position | velocity (sph_kernels.cuh)
    e->vel_m = e->pos_m + 1;
    ALLOC(e->spos, n);
    ALLOC(e->svel, n);
    ALLOC(e->sforce, n);
    ALLOC(e->srho, n);
    ALLOC(e->keys, n);
    ALLOC(e->ka, n);
    ALLOC(e->va, n);
    ALLOC(e->kb, n);
    ALLOC(e->vb, n);
    ALLOC(e->block_hist, (size_t)RS_RADIX * e->ntiles);
    ALLOC(e->digit_total, RS_RADIX);
    if (const char *so = getenv("SPH_SORT")) {
        e->onesweep = strcmp(so, "classic") != 0;
        // lookback2 (default) | count2: second generation, tile offsets by decoupled look-back | per-tile counts + scan;
        // lookback | count: first generation (register-to-global scatter); classic: three-kernel passes of radix_sort.cuh
        e->sort_lookback = strcmp(so, "lookback") == 0 || strcmp(so, "lookback2") == 0;
        e->sort_gen2 = strcmp(so, "lookback2") == 0 || strcmp(so, "count2") == 0;
    }
    if (e->passes > OS_MAX_PASSES) e->onesweep = false;
    ALLOC(e->os_ctrl, os_ctrl_words(e->passes, (n + OS_TILE - 1) / OS_TILE));
    ALLOC(e->tile_plans, (n + RB_THREADS - 1) / RB_THREADS);
    e->ntiles_rb = (n + RB_THREADS - 1) / RB_THREADS;
    ALLOC(e->refused, 6 * (size_t)e->ntiles_rb + 4);
    ALLOC(e->cell_range, e->cell_capacity);
    if (e->slab) ALLOC(e->gid, n);
    ALLOC(e->stats_d, 4);
    ALLOC(e->nlist, (size_t)n * 32);
    ALLOC(e->dlist, (size_t)n * 32);
    ALLOC(e->ncnt, n);
    if (params->flags & SPH_FLAG_RECORD_TERMS) {
        ALLOC(e->spress, n);
        ALLOC(e->svisc, n);
    }
    if (params->mode == SPH_MODE_PIPE && !e->slab) {   // slab mode: sph_slab_configure sizes it by the global count
        e->rng_count = n;
        ALLOC(e->rng, 2 * (size_t)n);
        std::vector<uint64_t> st;
        init_rng_host(st, n, params->rng_seed);
        if (cudaMemcpy(e->rng, st.data(), st.size() * sizeof(uint64_t), cudaMemcpyHostToDevice) != cudaSuccess) {
            sph_destroy(e);
            return fail("rng upload failed");
        }
    }
#undef ALLOC
    // final sort output buffer: pass 0 writes A, pass 1 writes B, ...
    const bool final_in_a = (e->passes % 2) == 1;
    e->skeys = final_in_a ? e->ka : e->kb;
    e->sids = final_in_a ? e->va : e->vb;

    cudaMemset(e->pos_m, 0, 2 * sizeof(float4) * (size_t)n);
    cudaMemset(e->sforce, 0, sizeof(float4) * (size_t)n);
    cudaMemset(e->srho, 0, sizeof(float) * (size_t)n);
    cudaMemset(e->keys, 0, sizeof(uint32_t) * (size_t)n);
    cudaMemset(e->cell_range, 0, sizeof(int2) * (size_t)e->cell_capacity);
    if (cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking) != cudaSuccess) {
        sph_destroy(e);
        return fail("cudaStreamCreate failed");
    }
    e->own_stream = true;
    {
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        if (cudaStreamCreateWithPriority(&e->aux_stream, cudaStreamNonBlocking, hi) != cudaSuccess) {
            sph_destroy(e);
            return fail("cudaStreamCreate failed");
        }
        for (int i = 0; i < 2; ++i) {
            cudaEventCreateWithFlags(&e->ev_fork[i], cudaEventDisableTiming);
            cudaEventCreateWithFlags(&e->ev_join[i], cudaEventDisableTiming);
        }
    }
    for (auto &ev : e->ev) cudaEventCreate(&ev);
    cudaFuncSetAttribute(os2_pass<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Os2Smem));
    cudaFuncSetAttribute(os2_pass<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Os2Smem));
    cudaFuncSetAttribute(density_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(DensityRowsSmem));
    cudaFuncSetAttribute(density_flat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FlatSmem));
    cudaFuncSetAttribute(density_dense_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(DenseSmem));
    cudaFuncSetAttribute(density_rows_items_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)sizeof(DensityRowsSmem));
    cudaFuncSetAttribute(force_rows_items_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)sizeof(ForceRowsSmem));
    cudaFuncSetAttribute(force_rows_items_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)sizeof(ForceRowsSmem));
    cudaFuncSetAttribute(force_rows_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)sizeof(ForceRowsSmem));
    cudaFuncSetAttribute(force_rows_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)sizeof(ForceRowsSmem));
    // launches of one step (kernels + the two memsets), as enqueue_step issues them on a single GPU
    e->launches_per_step =
        ((e->onesweep && e->sort_gen2) ? 0 : 1)                                                      // hash (fused into hash_hist otherwise)
        + (e->onesweep ? 2 + (e->sort_lookback ? 1 : 3) * e->passes : 3 * e->passes)                  // memset + histogram + passes
        + 2                                                                                          // memset(cell_range), reorder
        + (e->rows_sweeps ? 1 /* rows_plan */ + (e->flat_density ? 4 + 3 : 2 + 2) : 2);              // density + force kernels
    if (cudaDeviceSynchronize() != cudaSuccess) {
        sph_destroy(e);
        return fail("device error during create");
    }
    *out = e;
    return 0;
}

int sph_destroy(sph_handle_t e) {
    if (!e) return 0;
    cudaSetDevice(e->device);
    cudaDeviceSynchronize();
    invalidate_graph(e);
    void *ptrs[] = {e->pos_m, e->spos, e->svel, e->sforce, e->spress, e->svisc, e->srho, e->nlist, e->dlist, e->ncnt,
                    e->keys, e->ka, e->va, e->kb, e->vb, e->block_hist, e->digit_total, e->os_ctrl, e->tile_plans, e->refused, e->cell_range, e->pipe_d,
                    e->rng, e->gid, e->stage, e->stats_d, e->snap_pos, e->snap_vel, e->snap_rng, e->sendbuf, e->recv_alloc, e->emit_d, e->peer_flags_d, e->emit_cnt,
                    e->slab_counters, e->tmp_gid};
    if (e->p2p_ready)
        for (int k = 0; k < e->route.world; ++k)
            if (k != e->route.rank && e->peer_base[k]) cudaIpcCloseMemHandle(e->peer_base[k]);
    for (void *q : ptrs)
        if (q) cudaFree(q);
    for (auto &ev : e->ev)
        if (ev) cudaEventDestroy(ev);
    for (int k = 0; k < 3; ++k) {
        if (e->exp_dev[k]) cudaFree(e->exp_dev[k]);
        if (e->exp_host[k]) cudaFreeHost(e->exp_host[k]);
        if (e->exp_ev[k]) cudaEventDestroy(e->exp_ev[k]);
        if (e->exp_stats_dev[k]) cudaFree(e->exp_stats_dev[k]);
        if (e->exp_stats_host[k]) cudaFreeHost(e->exp_stats_host[k]);
    }
    if (e->exp_packed) cudaEventDestroy(e->exp_packed);
    if (e->fstats_d) cudaFree(e->fstats_d);
    if (e->own_stream && e->stream) cudaStreamDestroy(e->stream);
    if (e->copy_stream) cudaStreamDestroy(e->copy_stream);
    if (e->aux_stream) cudaStreamDestroy(e->aux_stream);
    for (int i = 0; i < 2; ++i) {
        if (e->ev_fork[i]) cudaEventDestroy(e->ev_fork[i]);
        if (e->ev_join[i]) cudaEventDestroy(e->ev_join[i]);
    }
    for (auto &ev : e->ev_copy)
        if (ev) cudaEventDestroy(ev);
    delete e;
    return 0;
}

int sph_set_stream(sph_handle_t e, void *cuda_stream) {
    if (!e) return fail("null handle");
    CK(cudaSetDevice(e->device));
    CK(cudaStreamSynchronize(e->stream));
    if (e->own_stream && e->stream) cudaStreamDestroy(e->stream);
    e->stream = (cudaStream_t)cuda_stream;
    e->own_stream = false;
    invalidate_graph(e);
    return 0;
}

int sph_set_pipe(sph_handle_t e, const double *table, int32_t rows) {
    if (!e) return fail("null handle");
    if (rows < 2 || !table) return fail("pipe table needs >= 2 rows of 5 doubles");
    CK(cudaSetDevice(e->device));
    CK(cudaStreamSynchronize(e->stream));
    if (e->pipe_d) cudaFree(e->pipe_d);
    e->pipe_d = nullptr;
    CK(cudaMalloc((void **)&e->pipe_d, sizeof(double) * 5 * (size_t)rows));
    CK(cudaMemcpy(e->pipe_d, table, sizeof(double) * 5 * (size_t)rows, cudaMemcpyHostToDevice));
    e->pipe_rows = rows;
    e->consts.pipe_rows = rows;
    invalidate_graph(e);
    return 0;
}

template <typename T>
static int upload_impl(SphEngine *e, const T *pos, const T *vel) {
    if (!e) return fail("null handle");
    if (!pos || !vel) return fail("position / velocity is NULL");
    CK(cudaSetDevice(e->device));
    const size_t n = e->n, bytes = 3 * n * sizeof(T);
    if (ensure_stage(e, 7 * n * sizeof(double))) return 1;
    T *dpos = (T *)e->stage, *dvel = dpos + 3 * n;
    CK(cudaMemcpyAsync(dpos, pos, bytes, cudaMemcpyHostToDevice, e->stream));
    CK(cudaMemcpyAsync(dvel, vel, bytes, cudaMemcpyHostToDevice, e->stream));
    pack_state_kernel<T><<<(e->n + 255) / 256, 256, 0, e->stream>>>(dpos, dvel, e->pos_m, e->vel_m, e->n);
    CK(cudaGetLastError());
    e->launches += 1;
    e->has_state = true;
    return 0;
}
int sph_upload(sph_handle_t e, const double *pos, const double *vel) { return upload_impl<double>(e, pos, vel); }
int sph_upload_f32(sph_handle_t e, const float *pos, const float *vel) { return upload_impl<float>(e, pos, vel); }

template <typename T>
static int download_impl(SphEngine *e, T *pos, T *vel, T *rho) {
    if (!e) return fail("null handle");
    CK(cudaSetDevice(e->device));
    const size_t n = e->n;
    if (ensure_stage(e, 7 * n * sizeof(double))) return 1;
    T *dpos = (T *)e->stage, *dvel = dpos + 3 * n, *drho = dvel + 3 * n;
    unpack_state_kernel<T><<<(e->n + 255) / 256, 256, 0, e->stream>>>(e->pos_m, e->vel_m, pos ? dpos : nullptr,
                                                                       vel ? dvel : nullptr, rho ? drho : nullptr,
                                                                       e->n);
    CK(cudaGetLastError());
    e->launches += 1;
    if (pos) CK(cudaMemcpyAsync(pos, dpos, 3 * n * sizeof(T), cudaMemcpyDeviceToHost, e->stream));
    if (vel) CK(cudaMemcpyAsync(vel, dvel, 3 * n * sizeof(T), cudaMemcpyDeviceToHost, e->stream));
    if (rho) CK(cudaMemcpyAsync(rho, drho, n * sizeof(T), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    return 0;
}
int sph_download(sph_handle_t e, double *p, double *v, double *r) { return download_impl<double>(e, p, v, r); }
int sph_download_f32(sph_handle_t e, float *p, float *v, float *r) { return download_impl<float>(e, p, v, r); }

static SweepArgs sweep_args(SphEngine *e, const uint32_t *sids, int n, int n_own) {
    SweepArgs sa{};
    sa.spos = e->spos;
    sa.svel = e->svel;
    sa.skeys = e->skeys;
    sa.sids = sids;
    sa.cell_range = e->cell_range;
    sa.srho = e->srho;
    sa.nlist = e->nlist;
    sa.ncnt = e->ncnt;
    sa.pos_m = e->pos_m;
    sa.vel_m = e->vel_m;
    sa.sforce = e->sforce;
    sa.spress = e->spress;
    sa.svisc = e->svisc;
    sa.pipe = e->pipe_d;
    sa.rng = e->rng;
    sa.gid = e->slab ? e->gid : nullptr;
    sa.n = n;
    sa.n_own = n_own;
    sa.plans = e->tile_plans;
    sa.n_items = e->refused;       // [0] passes (list A), [1] tiles refused by density_flat_kernel (B), [2] dense tiles (C)
    sa.items = e->refused + 4;
    sa.n_dense = e->refused + 2;
    sa.dense_items = e->refused + 4 + 5 * (size_t)e->ntiles_rb;
    return sa;
}

// Enqueue one step on e->stream.  If evs != nullptr, records stage boundaries into e->ev[0..5].
// Enqueue one step on e->stream; `stages` selects parts of it (1: hash + sort, 2: cell table + reorder + row plans +
// density sweep, 4: force sweep) so that sph_compute_next_state can interleave them with its host copies.
// split_vel: the reorder leaves the velocities out and stage 4 starts with their gather (the caller uploads them late).
static int enqueue_step(SphEngine *e, bool timed, int n, int n_own, int stages = 7, bool split_vel = false) {
    cudaStream_t s = e->stream;
    const int g256 = (n + 255) / 256;
    const int ntiles = (n + RS_TILE - 1) / RS_TILE;
    if (timed) cudaEventRecord(e->ev[0], s);
    const bool gen2 = e->onesweep && e->sort_gen2;
    if (!(stages & 1)) {
    } else if (gen2) {
        // second-generation sort: hash + digit histograms of every pass in one kernel, then one (look-back) or three
        // (count, scan, scatter) launches per 8-bit digit with 4096-pair tiles sorted in shared memory
        const int otiles = (n + OS2_TILE - 1) / OS2_TILE;
        OsPasses ps{};
        ps.n_passes = e->passes;
        int shift = 0;
        for (int p = 0; p < e->passes; ++p) {
            ps.shift[p] = shift;
            ps.mask[p] = (1u << e->pass_bits[p]) - 1u;
            shift += e->pass_bits[p];
        }
        cudaMemsetAsync(e->os_ctrl, 0,
                        sizeof(uint32_t) * (e->sort_lookback ? os_ctrl_words(e->passes, otiles)
                                                             : (size_t)e->passes * OS_RADIX + 8), s);
        hash_hist_kernel<<<std::min((n + OS_THREADS - 1) / OS_THREADS, 148 * 8), OS_THREADS, 0, s>>>(
            e->pos_m, e->keys, n, e->grid, ps, e->os_ctrl);
        if (timed) cudaEventRecord(e->ev[1], s);
        const uint32_t *kin = e->keys, *vin = nullptr;
        uint32_t *tile_cnt = e->os_ctrl + (size_t)e->passes * OS_RADIX + 8;
        for (int p = 0; p < e->passes; ++p) {
            uint32_t *kout = (p % 2 == 0) ? e->ka : e->kb;
            uint32_t *vout = (p % 2 == 0) ? e->va : e->vb;
            if (e->sort_lookback) {
                os2_pass<true><<<otiles, OS_THREADS, sizeof(Os2Smem), s>>>(kin, vin, kout, vout, n, p, e->passes,
                                                                            ps.shift[p], ps.mask[p], otiles, e->os_ctrl,
                                                                            nullptr);
            } else {
                ts2_hist<<<otiles, OS_THREADS, 0, s>>>(kin, n, ps.shift[p], ps.mask[p], tile_cnt, otiles);
                ts_scan<<<OS_RADIX, OS_THREADS, 0, s>>>(tile_cnt, otiles, e->os_ctrl + p * OS_RADIX);
                os2_pass<false><<<otiles, OS_THREADS, sizeof(Os2Smem), s>>>(kin, vin, kout, vout, n, p, e->passes,
                                                                             ps.shift[p], ps.mask[p], otiles, e->os_ctrl,
                                                                             tile_cnt);
            }
            kin = kout;
            vin = vout;
        }
    }
    if ((stages & 1) && !gen2) hash_kernel<<<g256, 256, 0, s>>>(e->pos_m, e->keys, n, e->grid);
    if (timed && !((stages & 1) && gen2)) cudaEventRecord(e->ev[1], s);
    // LSD radix sort of (key, id): pass 0 reads keys with implicit iota values
    if (!(stages & 1) || gen2) {
    } else if (e->onesweep) {
        const int otiles = (n + OS_TILE - 1) / OS_TILE;
        OsPasses ps{};
        ps.n_passes = e->passes;
        int shift = 0;
        for (int p = 0; p < e->passes; ++p) {
            ps.shift[p] = shift;
            ps.mask[p] = (1u << e->pass_bits[p]) - 1u;
            shift += e->pass_bits[p];
        }
        cudaMemsetAsync(e->os_ctrl, 0,
                        sizeof(uint32_t) * (e->sort_lookback ? os_ctrl_words(e->passes, otiles)
                                                             : (size_t)e->passes * OS_RADIX + 8), s);
        os_hist<<<std::min(otiles, 148 * 8), OS_THREADS, 0, s>>>(e->keys, n, ps, e->os_ctrl);
        const uint32_t *kin = e->keys, *vin = nullptr;
        // per-tile counts live behind the control words of the look-back variant (same buffer, never used together)
        uint32_t *tile_cnt = e->os_ctrl + (size_t)e->passes * OS_RADIX + 8;
        for (int p = 0; p < e->passes; ++p) {
            uint32_t *kout = (p % 2 == 0) ? e->ka : e->kb;
            uint32_t *vout = (p % 2 == 0) ? e->va : e->vb;
            if (e->sort_lookback) {
                os_pass<<<otiles, OS_THREADS, 0, s>>>(kin, vin, kout, vout, n, p, e->passes, ps.shift[p], ps.mask[p],
                                                      otiles, e->os_ctrl);
            } else {
                ts_hist<<<otiles, OS_THREADS, 0, s>>>(kin, n, ps.shift[p], ps.mask[p], tile_cnt, otiles);
                ts_scan<<<OS_RADIX, OS_THREADS, 0, s>>>(tile_cnt, otiles, e->os_ctrl + p * OS_RADIX);
                ts_scatter<<<otiles, OS_THREADS, 0, s>>>(kin, vin, kout, vout, n, ps.shift[p], ps.mask[p], tile_cnt,
                                                         otiles);
            }
            kin = kout;
            vin = vout;
        }
    } else {
        const uint32_t *kin = e->keys, *vin = nullptr;
        int shift = 0;
        for (int p = 0; p < e->passes; ++p) {
            uint32_t *kout = (p % 2 == 0) ? e->ka : e->kb;
            uint32_t *vout = (p % 2 == 0) ? e->va : e->vb;
            const uint32_t mask = (1u << e->pass_bits[p]) - 1u;
            rs_hist<<<ntiles, RS_THREADS, 0, s>>>(kin, n, shift, mask, e->block_hist, ntiles);
            rs_scan<<<RS_RADIX, RS_THREADS, 0, s>>>(e->block_hist, ntiles, e->digit_total);
            rs_scatter<<<ntiles, RS_THREADS, 0, s>>>(kin, vin, kout, vout, n, shift, mask, e->block_hist, ntiles,
                                                     e->digit_total);
            kin = kout;
            vin = vout;
            shift += e->pass_bits[p];
        }
    }
    if (timed) cudaEventRecord(e->ev[2], s);
    const uint32_t *sids = e->sids;
    if (e->slab) sids = (e->sids == e->va) ? e->vb : e->va;   // in-cell order repaired below
    if (stages & 2) {
        cudaMemsetAsync(e->cell_range, 0, sizeof(int2) * ((size_t)e->grid.ncells + 1), s);
        if (e->slab) {   // in-cell order by GLOBAL id (the local index order is arrival order on a slab)
            cell_range_kernel<<<g256, 256, 0, s>>>(e->skeys, e->cell_range, n);
            fix_order_kernel<<<g256, 256, 0, s>>>(e->skeys, e->sids, const_cast<uint32_t *>(sids), e->gid, e->cell_range, n,
                                                  (uint32_t)e->grid.ncells);
        }
        if (split_vel)
            reorder_kernel<false><<<g256, 256, 0, s>>>(e->skeys, sids, e->pos_m, e->vel_m, e->spos, e->svel, e->cell_range, n,
                                                       e->refused);
        else
            reorder_kernel<true><<<g256, 256, 0, s>>>(e->skeys, sids, e->pos_m, e->vel_m, e->spos, e->svel, e->cell_range, n,
                                                      e->refused);
    }
    if ((stages & 4) && split_vel) gather_vel_kernel<<<g256, 256, 0, s>>>(sids, e->vel_m, e->svel, n);
    if (timed) cudaEventRecord(e->ev[3], s);
    const SweepArgs sa = sweep_args(e, sids, n, n_own);
    if (e->rows_sweeps) {
        const int grb = (n + RB_THREADS - 1) / RB_THREADS;
        cudaStream_t x = e->aux_stream;
        int *list_b = e->refused + 4 + 4 * (size_t)e->ntiles_rb;
        if (stages & 2) {
            rows_plan_kernel<<<(grb + RP_WARPS - 1) / RP_WARPS, RP_WARPS * 32, 0, s>>>(
                sa, e->grid, e->tile_plans, grb, e->refused, e->refused + 4, e->refused + 2,
                e->refused + 4 + 5 * (size_t)e->ntiles_rb, e->flat_density ? DENSE_MAX_CELLS : 0, e->dense_min_slots);
            // fork: the tiles whose rows do not fit run next to the main density sweep
            cudaEventRecord(e->ev_fork[0], s);
            cudaStreamWaitEvent(x, e->ev_fork[0], 0);
            if (e->flat_density)
                density_dense_kernel<<<148 * 3, DN_THREADS, sizeof(DenseSmem), x>>>(sa, e->grid, e->consts, e->dlist);
            density_rows_items_kernel<<<ITEM_CTAS, RB_THREADS, sizeof(DensityRowsSmem), x>>>(sa, e->grid, e->consts,
                                                                                             sa.n_items, sa.items);
            cudaEventRecord(e->ev_join[0], x);
            if (e->flat_density) {
                const FlatArgs fa{list_b, e->refused + 1};
                density_flat_kernel<<<grb, FL_THREADS, sizeof(FlatSmem), s>>>(sa, e->grid, e->consts, fa);
                density_rows_items_kernel<<<ITEM_CTAS, RB_THREADS, sizeof(DensityRowsSmem), s>>>(
                    sa, e->grid, e->consts, e->refused + 1, list_b);
            } else {
                density_rows_kernel<<<grb, RB_THREADS, sizeof(DensityRowsSmem), s>>>(sa, e->grid, e->consts);
            }
            cudaStreamWaitEvent(s, e->ev_join[0], 0);
        }
        if (timed) cudaEventRecord(e->ev[4], s);
        if (stages & 4) {
            cudaEventRecord(e->ev_fork[1], s);
            cudaStreamWaitEvent(x, e->ev_fork[1], 0);
            if (e->spress) {
                if (e->flat_density)
                    force_gather_kernel<true><<<148 * 8, RB_THREADS, 0, x>>>(sa, e->grid, e->consts, e->dlist);
                force_rows_items_kernel<true><<<ITEM_CTAS, RB_THREADS, sizeof(ForceRowsSmem), x>>>(sa, e->grid, e->consts);
                cudaEventRecord(e->ev_join[1], x);
                force_rows_kernel<true><<<grb, RB_THREADS, sizeof(ForceRowsSmem), s>>>(sa, e->grid, e->consts);
            } else {
                if (e->flat_density)
                    force_gather_kernel<false><<<148 * 8, RB_THREADS, 0, x>>>(sa, e->grid, e->consts, e->dlist);
                force_rows_items_kernel<false><<<ITEM_CTAS, RB_THREADS, sizeof(ForceRowsSmem), x>>>(sa, e->grid, e->consts);
                cudaEventRecord(e->ev_join[1], x);
                force_rows_kernel<false><<<grb, RB_THREADS, sizeof(ForceRowsSmem), s>>>(sa, e->grid, e->consts);
            }
            cudaStreamWaitEvent(s, e->ev_join[1], 0);
        }
    } else {
        const int gsw = (n + SW_THREADS - 1) / SW_THREADS;
        if (stages & 2) density_kernel<<<gsw, SW_THREADS, 0, s>>>(sa, e->grid, e->consts);
        if (timed) cudaEventRecord(e->ev[4], s);
        if (!(stages & 4)) {
        } else if (e->spress)
            force_kernel<true><<<gsw, SW_THREADS, 0, s>>>(sa, e->grid, e->consts);
        else
            force_kernel<false><<<gsw, SW_THREADS, 0, s>>>(sa, e->grid, e->consts);
    }
    if (timed) cudaEventRecord(e->ev[5], s);
    CK(cudaGetLastError());
    return 0;
}

static int check_ready(SphEngine *e) {
    if (!e) return fail("null handle");
    if (e->slab) return fail("this handle is in x-slab mode: use sph_slab_step");
    if (!e->has_state) return fail("no particle state: call sph_upload first");
    if (e->p.mode == SPH_MODE_PIPE && !e->pipe_d) return fail("PIPE mode needs sph_set_pipe before stepping");
    return 0;
}

int sph_step(sph_handle_t e, int32_t n_steps) {
    if (check_ready(e)) return 1;
    if (n_steps <= 0) return 0;
    CK(cudaSetDevice(e->device));
    if (e->p.flags & SPH_FLAG_NO_GRAPH) {
        for (int i = 0; i < n_steps; ++i)
            if (enqueue_step(e, false, e->n, e->n)) return 1;
    } else {
        if (!e->graph_valid) {
            CK(cudaStreamBeginCapture(e->stream, cudaStreamCaptureModeThreadLocal));
            const int rc = enqueue_step(e, false, e->n, e->n);
            cudaGraph_t g = nullptr;
            cudaError_t ce = cudaStreamEndCapture(e->stream, &g);
            if (rc || ce != cudaSuccess) {
                if (g) cudaGraphDestroy(g);
                return rc ? 1 : fail(std::string("graph capture: ") + cudaGetErrorString(ce));
            }
            e->graph = g;
            CK(cudaGraphInstantiate(&e->graph_exec, e->graph, 0));
            e->graph_valid = true;
        }
        for (int i = 0; i < n_steps; ++i) CK(cudaGraphLaunch(e->graph_exec, e->stream));
    }
    e->steps_done += n_steps;
    e->launches += (int64_t)n_steps * e->launches_per_step;
    return 0;
}

int sph_step_timed(sph_handle_t e, int32_t n_steps, SphTimings *t) {
    if (check_ready(e)) return 1;
    if (!t) return fail("timings is NULL");
    CK(cudaSetDevice(e->device));
    memset(t, 0, sizeof(*t));
    for (int i = 0; i < n_steps; ++i) {
        if (enqueue_step(e, true, e->n, e->n)) return 1;
        CK(cudaEventSynchronize(e->ev[5]));
        float ms[5];
        for (int k = 0; k < 5; ++k) CK(cudaEventElapsedTime(&ms[k], e->ev[k], e->ev[k + 1]));
        t->hash_ms += ms[0];
        t->sort_ms += ms[1];
        t->reorder_ms += ms[2];
        t->density_ms += ms[3];
        t->force_ms += ms[4];
        float tot;
        CK(cudaEventElapsedTime(&tot, e->ev[0], e->ev[5]));
        t->total_ms += tot;
    }
    t->steps = n_steps;
    t->launches_per_step = e->launches_per_step;
    t->sort_passes = e->passes;
    e->steps_done += n_steps;
    e->launches += (int64_t)n_steps * e->launches_per_step;
    return 0;
}

static int ensure_copy_stream(SphEngine *e) {
    if (e->copy_stream) return 0;
    CK(cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking));
    for (auto &ev : e->ev_copy) CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    return 0;
}

// The reference-facing call.  Same result as sph_upload + sph_step(1) + sph_download, but the host copies are
// interleaved with the step on two streams: everything up to and including the density sweep needs positions only and
// runs while the velocities are still on their way in, and the density goes out while the force sweep runs (the copies are PCIe-bound: 104 B per particle against a 0.7 us step).
int sph_compute_next_state(sph_handle_t e, const double *pos_in, const double *vel_in, double *pos_out,
                           double *vel_out, double *density_out) {
    if (!e) return fail("null handle");
    if (!pos_in || !vel_in) return fail("position / velocity is NULL");
    if (e->slab) return fail("this handle is in x-slab mode: use sph_slab_step");
    if (e->p.mode == SPH_MODE_PIPE && !e->pipe_d) return fail("PIPE mode needs sph_set_pipe before stepping");
    CK(cudaSetDevice(e->device));
    const size_t n = e->n, vec = 3 * n * sizeof(double);
    if (ensure_stage(e, 7 * n * sizeof(double))) return 1;
    if (ensure_copy_stream(e)) return 1;
    cudaStream_t s = e->stream, c2 = e->copy_stream;
    double *dpos = (double *)e->stage, *dvel = dpos + 3 * n, *drho = dvel + 3 * n;
    const int g256 = (e->n + 255) / 256;
    // the staging buffer may still be in use by earlier work on the main stream
    CK(cudaEventRecord(e->ev_copy[2], s));
    CK(cudaStreamWaitEvent(c2, e->ev_copy[2], 0));
    CK(cudaMemcpyAsync(dpos, pos_in, vec, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(dvel, vel_in, vec, cudaMemcpyHostToDevice, c2));
    CK(cudaEventRecord(e->ev_copy[0], c2));
    pack_vec_kernel<double><<<g256, 256, 0, s>>>(dpos, e->pos_m, e->n);
    // hash, sort, cell table, position reorder, row plans and the density sweep need positions only: they run while the
    // velocities are still on their way in
    if (enqueue_step(e, false, e->n, e->n, 1 | 2, true)) return 1;
    CK(cudaEventRecord(e->ev_copy[1], s));
    if (density_out) {                                              // rho leaves while the forces are computed
        CK(cudaStreamWaitEvent(c2, e->ev_copy[1], 0));
        unsort_scalar_kernel<<<g256, 256, 0, c2>>>(e->srho, e->sids, drho, e->n);
        CK(cudaMemcpyAsync(density_out, drho, n * sizeof(double), cudaMemcpyDeviceToHost, c2));
    }
    CK(cudaStreamWaitEvent(s, e->ev_copy[0], 0));
    pack_vec_kernel<double><<<g256, 256, 0, s>>>(dvel, e->vel_m, e->n);
    if (enqueue_step(e, false, e->n, e->n, 4, true)) return 1;      // velocity gather, force + integrate + collide
    unpack_state_kernel<double><<<g256, 256, 0, s>>>(e->pos_m, e->vel_m, pos_out ? dpos : nullptr,
                                                     vel_out ? dvel : nullptr, (double *)nullptr, e->n);
    CK(cudaGetLastError());
    if (pos_out) CK(cudaMemcpyAsync(pos_out, dpos, vec, cudaMemcpyDeviceToHost, s));
    if (vel_out) CK(cudaMemcpyAsync(vel_out, dvel, vec, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    CK(cudaStreamSynchronize(c2));
    e->has_state = true;
    e->steps_done += 1;
    e->launches += e->launches_per_step + 5;   // + 2 packs, velocity gather, density unsort, unpack
    return 0;
}

int sph_export_begin(sph_handle_t e, int32_t slot, int32_t stride) {
    if (!e) return fail("null handle");
    if (slot < 0 || slot > 2) return fail("export slot must be 0, 1 or 2");
    if (stride < 1) return fail("stride must be >= 1");
    if (e->slab) return fail("frame export works on single-GPU handles");
    if (!e->has_state) return fail("no particle state");
    CK(cudaSetDevice(e->device));
    if (ensure_copy_stream(e)) return 1;
    const int64_t n_out = ((int64_t)e->n + stride - 1) / stride;
    if (!e->exp_ev[slot]) CK(cudaEventCreateWithFlags(&e->exp_ev[slot], cudaEventDisableTiming));
    if (!e->exp_packed) CK(cudaEventCreateWithFlags(&e->exp_packed, cudaEventDisableTiming));
    if (e->exp_cap[slot] < n_out) {
        CK(cudaEventSynchronize(e->exp_ev[slot]));
        if (e->exp_dev[slot]) cudaFree(e->exp_dev[slot]);
        if (e->exp_host[slot]) cudaFreeHost(e->exp_host[slot]);
        e->exp_dev[slot] = e->exp_host[slot] = nullptr;
        e->exp_cap[slot] = 0;
        CK(cudaMalloc((void **)&e->exp_dev[slot], 7 * sizeof(double) * (size_t)n_out));
        CK(cudaHostAlloc((void **)&e->exp_host[slot], 7 * sizeof(double) * (size_t)n_out, cudaHostAllocDefault));
        e->exp_cap[slot] = n_out;
    }
    // the slot's previous copy must have left the device staging buffer before it is overwritten
    CK(cudaStreamWaitEvent(e->stream, e->exp_ev[slot], 0));
    export_pack_kernel<<<(int)((n_out + 255) / 256), 256, 0, e->stream>>>(e->pos_m, e->vel_m, e->exp_dev[slot],
                                                                        (int)n_out, stride);
    CK(cudaGetLastError());
    // the frame's statistics ride along (section 8(f)4): reduced on the main stream right after the frame's steps
    if (!e->exp_stats_dev[slot]) {
        CK(cudaMalloc((void **)&e->exp_stats_dev[slot], 42 * sizeof(uint32_t)));
        CK(cudaHostAlloc((void **)&e->exp_stats_host[slot], 42 * sizeof(uint32_t), cudaHostAllocDefault));
    }
    frame_stats_init_kernel<<<1, 64, 0, e->stream>>>(e->exp_stats_dev[slot], e->cell_range + e->grid.ncells,
                                                     e->steps_done > 0 ? 1 : 0);
    frame_stats_kernel<<<(e->n + 255) / 256, 256, 0, e->stream>>>(e->pos_m, e->vel_m, e->ncnt, e->n,
                                                                  e->steps_done > 0 ? 1 : 0, e->exp_stats_dev[slot]);
    CK(cudaGetLastError());
    e->exp_steps[slot] = e->steps_done;
    CK(cudaEventRecord(e->exp_packed, e->stream));
    CK(cudaStreamWaitEvent(e->copy_stream, e->exp_packed, 0));
    CK(cudaMemcpyAsync(e->exp_host[slot], e->exp_dev[slot], 7 * sizeof(double) * (size_t)n_out, cudaMemcpyDeviceToHost,
                       e->copy_stream));
    CK(cudaMemcpyAsync(e->exp_stats_host[slot], e->exp_stats_dev[slot], 42 * sizeof(uint32_t), cudaMemcpyDeviceToHost,
                       e->copy_stream));
    CK(cudaEventRecord(e->exp_ev[slot], e->copy_stream));
    e->exp_n[slot] = n_out;
    e->launches += 3;
    return 0;
}

int sph_export_wait(sph_handle_t e, int32_t slot, double **pos, double **vel, double **rho, int64_t *n_out) {
    if (!e) return fail("null handle");
    if (slot < 0 || slot > 2 || !e->exp_ev[slot] || e->exp_n[slot] == 0) return fail("no export in flight on this slot");
    CK(cudaSetDevice(e->device));
    CK(cudaEventSynchronize(e->exp_ev[slot]));
    const size_t m = (size_t)e->exp_n[slot];
    if (pos) *pos = e->exp_host[slot];
    if (vel) *vel = e->exp_host[slot] + 3 * m;
    if (rho) *rho = e->exp_host[slot] + 6 * m;
    if (n_out) *n_out = (int64_t)m;
    return 0;
}

int sph_generate_state(sph_handle_t e, int32_t kind, uint64_t seed) {
    if (!e) return fail("null handle");
    if (e->slab) return fail("start-state generators work on single-GPU handles");
    if (kind < 0 || kind > 2) return fail("unknown generator kind");
    if (kind == SPH_GEN_PIPE && !e->pipe_d) return fail("the pipe generator needs sph_set_pipe");
    CK(cudaSetDevice(e->device));
    GenArgs ga{};
    ga.kind = kind;
    ga.seed = seed;
    for (int d = 0; d < 3; ++d) {
        const double full = e->p.space_size[d];
        ga.ext[d] = (float)((kind == SPH_GEN_BOX_WALL && d == 0) ? full * 0.1 : full);
        ga.top[d] = std::nextafter((float)full, 0.f);
    }
    ga.pipe = e->pipe_d;
    ga.pipe_rows = e->pipe_rows;
    generate_kernel<<<(e->n + 255) / 256, 256, 0, e->stream>>>(e->pos_m, e->vel_m, e->n, ga);
    CK(cudaGetLastError());
    e->launches += 1;
    e->has_state = true;
    return 0;
}

static void decode_frame_stats(const SphEngine *e, const uint32_t *h, int64_t steps, SphFrameStats *st) {
    auto unord = [](uint32_t u) {
        u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
        float f;
        memcpy(&f, &u, 4);
        return f;
    };
    memset(st, 0, sizeof(*st));
    st->steps_done = steps;
    st->n_particles = e->n;
    st->n_dead = (int32_t)h[41] - (int32_t)h[40];
    st->n_nonfinite = (int32_t)h[0];
    const bool any = h[1] != 0;   // at least one finite particle
    st->max_position = any ? unord(h[1]) : NAN;
    st->min_position = any ? unord(h[2]) : NAN;
    st->max_velocity = any ? unord(h[3]) : NAN;
    st->max_speed = any ? unord(h[4]) : NAN;
    st->max_density = h[5] ? unord(h[5]) : NAN;
    for (int k = 0; k < 33; ++k) st->neighbour_hist[k] = (int32_t)h[6 + k];
}

int sph_export_stats(sph_handle_t e, int32_t slot, SphFrameStats *st) {
    if (!e || !st) return fail("null argument");
    if (slot < 0 || slot > 2 || !e->exp_ev[slot] || e->exp_n[slot] == 0) return fail("no export in flight on this slot");
    CK(cudaSetDevice(e->device));
    CK(cudaEventSynchronize(e->exp_ev[slot]));
    decode_frame_stats(e, e->exp_stats_host[slot], e->exp_steps[slot], st);
    return 0;
}

int sph_get_frame_stats(sph_handle_t e, SphFrameStats *st) {
    if (!e || !st) return fail("null argument");
    if (e->slab) return fail("frame statistics work on single-GPU handles");
    CK(cudaSetDevice(e->device));
    if (!e->fstats_d) CK(cudaMalloc((void **)&e->fstats_d, 42 * sizeof(uint32_t)));
    frame_stats_init_kernel<<<1, 64, 0, e->stream>>>(e->fstats_d, e->cell_range + e->grid.ncells, e->steps_done > 0 ? 1 : 0);
    frame_stats_kernel<<<(e->n + 255) / 256, 256, 0, e->stream>>>(e->pos_m, e->vel_m, e->ncnt, e->n,
                                                                  e->steps_done > 0 ? 1 : 0, e->fstats_d);
    CK(cudaGetLastError());
    uint32_t h[42];
    CK(cudaMemcpyAsync(h, e->fstats_d, sizeof(h), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    decode_frame_stats(e, h, e->steps_done, st);
    e->launches += 2;
    return 0;
}

int sph_save_state(sph_handle_t e) {
    if (!e) return fail("null handle");
    if (!e->has_state) return fail("no particle state to save");
    CK(cudaSetDevice(e->device));
    const size_t n = e->n;
    if (!e->snap_pos) CK(cudaMalloc((void **)&e->snap_pos, 2 * sizeof(float4) * n));   // whole records (position | velocity)
    const size_t rng_bytes = 2 * sizeof(uint64_t) * (size_t)e->rng_count;   // slab mode: one state per GLOBAL particle
    if (e->rng && e->snap_rng && e->snap_rng_count != e->rng_count) {
        cudaFree(e->snap_rng);
        e->snap_rng = nullptr;
    }
    if (e->rng && !e->snap_rng) {
        CK(cudaMalloc((void **)&e->snap_rng, rng_bytes));
        e->snap_rng_count = e->rng_count;
    }
    CK(cudaMemcpyAsync(e->snap_pos, e->pos_m, 2 * sizeof(float4) * n, cudaMemcpyDeviceToDevice, e->stream));
    if (e->rng) CK(cudaMemcpyAsync(e->snap_rng, e->rng, rng_bytes, cudaMemcpyDeviceToDevice, e->stream));
    e->snap_steps = e->steps_done;
    return 0;
}

int sph_restore_state(sph_handle_t e) {
    if (!e) return fail("null handle");
    if (e->snap_steps < 0) return fail("sph_save_state has not been called");
    CK(cudaSetDevice(e->device));
    const size_t n = e->n;
    CK(cudaMemcpyAsync(e->pos_m, e->snap_pos, 2 * sizeof(float4) * n, cudaMemcpyDeviceToDevice, e->stream));
    if (e->rng && e->snap_rng && e->snap_rng_count == e->rng_count)
        CK(cudaMemcpyAsync(e->rng, e->snap_rng, 2 * sizeof(uint64_t) * (size_t)e->rng_count, cudaMemcpyDeviceToDevice,
                           e->stream));
    e->steps_done = e->snap_steps;
    return 0;
}

int sph_slab_configure(sph_handle_t e, int32_t x_lo, int32_t x_hi, int64_t n_global) {
    if (!e) return fail("null handle");
    if (!e->slab) return fail("create the engine with SPH_FLAG_SLAB");
    if (x_lo < 0 || x_hi > e->ceil_dims[0] || x_lo >= x_hi) return fail("bad slab column range");
    CK(cudaSetDevice(e->device));
    CK(cudaStreamSynchronize(e->stream));
    e->slab_lo = x_lo;
    e->slab_hi = x_hi;
    GridDesc &g = e->grid;
    g.xoff = x_lo - 2;                       // two ghost columns on each side
    g.wk = (x_hi - x_lo) + 4;
    g.wn = g.wk;
    g.ncells = g.wk * e->ceil_dims[1] * e->ceil_dims[2];
    g.strict_x = 1;
    g.own_lo = x_lo;
    g.own_hi = x_hi;
    if (e->p.mode == SPH_MODE_PIPE && e->rng_count != n_global) {
        if (e->rng) cudaFree(e->rng);
        e->rng = nullptr;
        CK(cudaMalloc((void **)&e->rng, 2 * sizeof(uint64_t) * (size_t)n_global));
        std::vector<uint64_t> st;
        init_rng_host(st, n_global, e->p.rng_seed);
        CK(cudaMemcpy(e->rng, st.data(), st.size() * sizeof(uint64_t), cudaMemcpyHostToDevice));
        e->rng_count = n_global;
    }
    e->has_state = true;
    return 0;
}

int sph_slab_step(sph_handle_t e, int32_t n_own, int32_t n_local) {
    if (!e) return fail("null handle");
    if (!e->slab) return fail("create the engine with SPH_FLAG_SLAB");
    if (e->slab_hi <= e->slab_lo) return fail("call sph_slab_configure first");
    if (n_own < 0 || n_local < n_own || n_local > e->n) return fail("bad particle counts (capacity exceeded?)");
    if (e->p.mode == SPH_MODE_PIPE && !e->pipe_d) return fail("PIPE mode needs sph_set_pipe before stepping");
    if (n_local == 0) return 0;
    CK(cudaSetDevice(e->device));
    if (enqueue_step(e, false, n_local, n_own)) return 1;
    e->steps_done += 1;
    e->launches += e->launches_per_step + 2;
    return 0;
}

// ---- native x-slab exchange (slab_exchange.cuh) ------------------------------------------------------------------
int sph_slab_exchange_init(sph_handle_t e, int32_t world, int32_t rank, const int32_t *bounds, int32_t own_cap,
                           const int32_t *cap_migrants, const int32_t *cap_ghosts, void **sendbuf, void **recvbuf,
                           int64_t *block_bytes) {
    if (!e || !bounds || !cap_migrants || !cap_ghosts) return fail("null argument");
    if (!e->slab) return fail("create the engine with SPH_FLAG_SLAB");
    if (e->slab_hi <= e->slab_lo) return fail("call sph_slab_configure first");
    if (world < 1 || world > SLAB_MAX_WORLD || rank < 0 || rank >= world) return fail("bad world / rank");
    if (own_cap <= 0 || own_cap >= e->n) return fail("own_cap must leave room for the ghost region");
    if (bounds[rank] != e->slab_lo || bounds[rank + 1] != e->slab_hi) return fail("bounds disagree with sph_slab_configure");
    CK(cudaSetDevice(e->device));
    CK(cudaStreamSynchronize(e->stream));
    SlabRoute &r = e->route;
    r = SlabRoute{};
    r.world = world;
    r.rank = rank;
    for (int k = 0; k <= world; ++k) r.bounds[k] = bounds[k];
    r.n_cols = e->ceil_dims[0];
    r.voxel_x = e->p.voxel_size[0];
    r.own_cap = own_cap;
    r.capacity = e->n;
    r.rec_bytes = (e->p.mode == SPH_MODE_PIPE) ? 48 : 32;
    int64_t off = 0;
    for (int k = 0; k < world; ++k) {
        if (cap_migrants[k] < 0 || cap_ghosts[k] < 0) return fail("negative block capacity");
        r.cap_m[k] = cap_migrants[k];
        r.cap_g[k] = cap_ghosts[k];
        r.peer_off[k] = off;
        const int64_t bytes = 16 + (int64_t)(cap_migrants[k] + cap_ghosts[k]) * r.rec_bytes;
        if (block_bytes) block_bytes[k] = bytes;
        off += bytes;
    }
    r.peer_off[world] = off;
    if (e->p2p_ready) return fail("sph_slab_exchange_init after sph_slab_open_peers: create a new handle");
    for (void *q : {(void *)e->sendbuf, (void *)e->recv_alloc, (void *)e->slab_counters, (void *)e->tmp_gid})
        if (q) cudaFree(q);
    e->sendbuf = e->recvbuf = e->recv_alloc = nullptr;
    off = (off + 255) & ~(int64_t)255;
    e->xchg_bytes = (size_t)off;
    e->parity = 0;
    e->slab_counters = e->tmp_gid = nullptr;
    CK(cudaMalloc((void **)&e->sendbuf, (size_t)off));
    CK(cudaMalloc((void **)&e->recv_alloc, SLAB_FLAG_BYTES + 2 * (size_t)off));
    CK(cudaMemset(e->recv_alloc, 0, SLAB_FLAG_BYTES + 2 * (size_t)off));
    e->recvbuf = e->recv_alloc + SLAB_FLAG_BYTES;
    CK(cudaMalloc((void **)&e->slab_counters, 8 * sizeof(int32_t)));
    CK(cudaMalloc((void **)&e->tmp_gid, sizeof(int32_t) * (size_t)own_cap));
    CK(cudaMemset(e->sendbuf, 0, (size_t)off));
    CK(cudaMemset(e->slab_counters, 0, 8 * sizeof(int32_t)));
    // every slot starts empty
    CK(cudaMemset(e->gid, 0xff, sizeof(int32_t) * (size_t)e->n));
    CK(cudaMemset2D(e->pos_m, 2 * sizeof(float4), 0xff, sizeof(float4), (size_t)e->n));   // positions only (record stride)
    if (sendbuf) *sendbuf = e->sendbuf;
    if (recvbuf) *recvbuf = e->recvbuf;
    e->route_ready = true;
    return 0;
}

static int check_route(SphEngine *e) {
    if (!e) return fail("null handle");
    if (!e->route_ready) return fail("call sph_slab_exchange_init first");
    return 0;
}

int sph_slab_route(sph_handle_t e) {
    if (check_route(e)) return 1;
    CK(cudaSetDevice(e->device));
    const SlabRoute &r = e->route;
    for (int k = 0; k < r.world; ++k) CK(cudaMemsetAsync(e->sendbuf + r.peer_off[k], 0, 16, e->stream));
    slab_route_kernel<<<(r.own_cap + 255) / 256, 256, 0, e->stream>>>(r, e->pos_m, e->vel_m, e->gid, e->rng, e->sendbuf,
                                                                      e->slab_counters);
    CK(cudaGetLastError());
    e->launches += 1 + r.world;
    return 0;
}

int sph_slab_unpack(sph_handle_t e) {
    if (check_route(e)) return 1;
    CK(cudaSetDevice(e->device));
    const SlabRoute &r = e->route;
    slab_clear_ghosts_kernel<<<(r.capacity - r.own_cap + 255) / 256, 256, 0, e->stream>>>(r, e->pos_m, e->gid,
                                                                                         e->slab_counters);
    int maxrec = 1;
    for (int k = 0; k < r.world; ++k) maxrec = std::max(maxrec, r.cap_m[k] + r.cap_g[k]);
    dim3 grid((maxrec + 255) / 256, r.world);
    slab_unpack_kernel<<<grid, 256, 0, e->stream>>>(r, e->recvbuf + (size_t)e->parity * e->xchg_bytes, e->pos_m, e->vel_m,
                                                    e->gid, e->rng, e->slab_counters);
    CK(cudaGetLastError());
    e->parity ^= 1;
    e->launches += 2;
    return 0;
}

int sph_slab_parity(sph_handle_t e, int32_t *parity) {
    if (check_route(e)) return 1;
    if (!parity) return fail("null argument");
    *parity = e->parity;
    return 0;
}

int sph_slab_ipc_handle(sph_handle_t e, void *handle64) {
    if (check_route(e)) return 1;
    if (!handle64) return fail("null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    CK(cudaSetDevice(e->device));
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, e->recv_alloc));
    memcpy(handle64, &h, 64);
    return 0;
}

int sph_slab_open_peers(sph_handle_t e, const void *handles, const int64_t *remote_off) {
    if (check_route(e)) return 1;
    if (!handles || !remote_off) return fail("null argument");
    if (e->p2p_ready) return fail("peers are already open");
    CK(cudaSetDevice(e->device));
    CK(cudaStreamSynchronize(e->stream));
    const SlabRoute &r = e->route;
    for (int k = 0; k < r.world; ++k) {
        if (k == r.rank) {
            e->peer_base[k] = e->recv_alloc;
            continue;
        }
        cudaIpcMemHandle_t h;
        memcpy(&h, (const unsigned char *)handles + 64 * (size_t)k, 64);
        CK(cudaIpcOpenMemHandle(&e->peer_base[k], h, cudaIpcMemLazyEnablePeerAccess));
    }
    CK(cudaMalloc((void **)&e->emit_d, 2 * sizeof(SlabEmit)));
    CK(cudaMalloc((void **)&e->peer_flags_d, sizeof(int32_t *) * SLAB_MAX_WORLD));
    CK(cudaMalloc((void **)&e->emit_cnt, sizeof(int32_t) * 2 * SLAB_MAX_WORLD));
    SlabEmit em[2];
    int32_t *flags[SLAB_MAX_WORLD] = {};
    for (int q = 0; q < 2; ++q) {
        em[q] = SlabEmit{};
        em[q].r = r;
        em[q].cnt = e->emit_cnt;
        em[q].counters = e->slab_counters;
        em[q].inv_vx = (float)(1.0 / r.voxel_x);
        em[q].in_lo = r.bounds[r.rank] + (r.rank > 0 ? SLAB_HALO : 0);
        em[q].in_hi = r.bounds[r.rank + 1] - (r.rank < r.world - 1 ? SLAB_HALO : 0);
        for (int k = 0; k < r.world; ++k) em[q].src[k] = e->sendbuf + r.peer_off[k];
        for (int k = 0; k < r.world; ++k)   // every rank lays out its two receive buffers by ITS block sizes
            em[q].dst[k] = (unsigned char *)e->peer_base[k] + SLAB_FLAG_BYTES + remote_off[2 * k + q];
    }
    for (int k = 0; k < r.world; ++k) flags[k] = (int32_t *)e->peer_base[k];
    CK(cudaMemcpy(e->emit_d, em, sizeof(em), cudaMemcpyHostToDevice));
    e->emit_h[0] = em[0];
    e->emit_h[1] = em[1];
    CK(cudaMemcpy(e->peer_flags_d, flags, sizeof(flags), cudaMemcpyHostToDevice));
    e->p2p_ready = true;
    return 0;
}

// The exchange over peer memory: route (CTA-aggregated) -> push the used records into the receivers' buffers -> counts +
// flags to the peers, wait for theirs.  sph_slab_unpack follows.
static int exchange_p2p(SphEngine *e, float *ms3) {
    if (check_route(e)) return 1;
    if (!e->p2p_ready) return fail("call sph_slab_open_peers first");
    CK(cudaSetDevice(e->device));
    e->epoch += 1;
    const SlabRoute &r = e->route;
    cudaStream_t s = e->stream;
    if (ms3) cudaEventRecord(e->ev[0], s);
    CK(cudaMemsetAsync(e->emit_cnt, 0, sizeof(int32_t) * 2 * SLAB_MAX_WORLD, s));
    slab_route_cta_kernel<<<(r.own_cap + ROUTE_CTA - 1) / ROUTE_CTA, ROUTE_CTA, 0, s>>>(e->emit_h[e->parity], e->pos_m,
                                                                                      e->vel_m, e->gid, e->rng);
    if (ms3) cudaEventRecord(e->ev[1], s);
    slab_push_kernel<<<dim3(PUSH_CTAS, r.world), 256, 0, s>>>(e->emit_h[e->parity]);
    if (ms3) cudaEventRecord(e->ev[2], s);
    slab_signal_wait_kernel<<<1, 32, 0, s>>>(e->emit_d + e->parity, e->peer_flags_d, (volatile int32_t *)e->recv_alloc,
                                             e->epoch, 4000000000LL);
    if (ms3) cudaEventRecord(e->ev[3], s);
    CK(cudaGetLastError());
    e->launches += 4;
    if (ms3) {
        CK(cudaEventSynchronize(e->ev[3]));
        for (int k = 0; k < 3; ++k) CK(cudaEventElapsedTime(&ms3[k], e->ev[k], e->ev[k + 1]));
    }
    return 0;
}
int sph_slab_exchange_p2p(sph_handle_t e) { return exchange_p2p(e, nullptr); }
/* same with CUDA events around the three kernels: ms3 = {route, push, flag barrier}; synchronises */
int sph_slab_exchange_p2p_timed(sph_handle_t e, float *ms3) {
    if (!ms3) return fail("null argument");
    return exchange_p2p(e, ms3);
}

int sph_slab_step_all(sph_handle_t e) {
    if (check_route(e)) return 1;
    if (e->p.mode == SPH_MODE_PIPE && !e->pipe_d) return fail("PIPE mode needs sph_set_pipe before stepping");
    CK(cudaSetDevice(e->device));
    if (enqueue_step(e, false, e->n, e->route.own_cap)) return 1;
    e->steps_done += 1;
    e->launches += e->launches_per_step + 2;
    return 0;
}

int sph_slab_step_all_timed(sph_handle_t e, SphTimings *t) {
    if (check_route(e)) return 1;
    if (!t) return fail("timings is NULL");
    if (e->p.mode == SPH_MODE_PIPE && !e->pipe_d) return fail("PIPE mode needs sph_set_pipe before stepping");
    CK(cudaSetDevice(e->device));
    memset(t, 0, sizeof(*t));
    if (enqueue_step(e, true, e->n, e->route.own_cap)) return 1;
    CK(cudaEventSynchronize(e->ev[5]));
    float ms[5];
    for (int k = 0; k < 5; ++k) CK(cudaEventElapsedTime(&ms[k], e->ev[k], e->ev[k + 1]));
    t->hash_ms = ms[0];
    t->sort_ms = ms[1];
    t->reorder_ms = ms[2];
    t->density_ms = ms[3];
    t->force_ms = ms[4];
    CK(cudaEventElapsedTime(&t->total_ms, e->ev[0], e->ev[5]));
    t->steps = 1;
    t->launches_per_step = e->launches_per_step + 2;
    t->sort_passes = e->passes;
    e->steps_done += 1;
    e->launches += e->launches_per_step + 2;
    return 0;
}

int sph_slab_compact(sph_handle_t e) {
    if (check_route(e)) return 1;
    CK(cudaSetDevice(e->device));
    const SlabRoute &r = e->route;
    const int g = (r.own_cap + 255) / 256;
    // spos / svel are free between steps: use them as the compaction scratch
    slab_compact_gather_kernel<<<g, 256, 0, e->stream>>>(r, e->pos_m, e->vel_m, e->gid, e->spos, e->svel, e->tmp_gid,
                                                         e->slab_counters);
    slab_compact_scatter_kernel<<<g, 256, 0, e->stream>>>(r, e->pos_m, e->vel_m, e->gid, e->spos, e->svel, e->tmp_gid,
                                                          e->slab_counters);
    slab_compact_finish_kernel<<<1, 1, 0, e->stream>>>(e->slab_counters);
    CK(cudaGetLastError());
    e->launches += 3;
    return 0;
}

int sph_slab_counters(sph_handle_t e, int32_t *out5) {
    if (check_route(e)) return 1;
    if (!out5) return fail("null argument");
    CK(cudaSetDevice(e->device));
    const SlabRoute &r = e->route;
    CK(cudaMemsetAsync(e->slab_counters + SLAB_NLIVE, 0, sizeof(int32_t), e->stream));
    slab_count_kernel<<<(r.own_cap + 255) / 256, 256, 0, e->stream>>>(r, e->gid, e->slab_counters);
    CK(cudaGetLastError());
    int32_t h[8];
    CK(cudaMemcpyAsync(h, e->slab_counters, sizeof(h), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    out5[0] = h[SLAB_HWM];
    out5[1] = h[SLAB_NGHOST];
    out5[2] = h[SLAB_OVERFLOW];
    out5[3] = h[SLAB_NLIVE];
    out5[4] = r.own_cap;
    return 0;
}

int sph_sync(sph_handle_t e) {
    if (!e) return fail("null handle");
    CK(cudaSetDevice(e->device));
    CK(cudaStreamSynchronize(e->stream));
    return 0;
}

// ---- parity taps ----------------------------------------------------------------------------------------------
static int d2h(SphEngine *e, void *dst, const void *src, size_t bytes) {
    CK(cudaSetDevice(e->device));
    CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    return 0;
}
int sph_get_keys(sph_handle_t e, int32_t *keys) {
    if (!e || !keys) return fail("null argument");
    return d2h(e, keys, e->keys, sizeof(int32_t) * (size_t)e->n);
}
int sph_get_sorted_ids(sph_handle_t e, int32_t *ids) {
    if (!e || !ids) return fail("null argument");
    return d2h(e, ids, e->sids, sizeof(int32_t) * (size_t)e->n);
}
int sph_get_sorted_keys(sph_handle_t e, int32_t *keys) {
    if (!e || !keys) return fail("null argument");
    return d2h(e, keys, e->skeys, sizeof(int32_t) * (size_t)e->n);
}
int sph_get_voxel_begin(sph_handle_t e, int32_t *begin, int64_t n_cells) {
    if (!e || !begin) return fail("null argument");
    if (n_cells != e->grid.ncells) return fail("n_cells mismatch (use sph_n_cells)");
    CK(cudaSetDevice(e->device));
    if (ensure_stage(e, sizeof(int32_t) * (size_t)n_cells)) return 1;
    voxel_begin_kernel<<<(int)((n_cells + 255) / 256), 256, 0, e->stream>>>(e->cell_range, (int32_t *)e->stage,
                                                                            (int)n_cells);
    CK(cudaGetLastError());
    return d2h(e, begin, e->stage, sizeof(int32_t) * (size_t)n_cells);
}
int sph_get_neighbour_counts(sph_handle_t e, int32_t *counts) {
    if (!e || !counts) return fail("null argument");
    if (e->steps_done == 0) return fail("no step has run yet");
    CK(cudaSetDevice(e->device));
    if (ensure_stage(e, sizeof(int32_t) * (size_t)e->n)) return 1;
    unsort_count_kernel<<<(e->n + 255) / 256, 256, 0, e->stream>>>(e->ncnt, e->sids, (int32_t *)e->stage, e->n);
    CK(cudaGetLastError());
    return d2h(e, counts, e->stage, sizeof(int32_t) * (size_t)e->n);
}
int sph_get_neighbour_lists(sph_handle_t e, int32_t *lists) {
    if (!e || !lists) return fail("null argument");
    if (e->steps_done == 0) return fail("no step has run yet");
    if (e->slab) return fail("the neighbour-list tap works on single-GPU handles");
    if (!e->rows_sweeps) return fail("the neighbour-list tap needs the row-staged sweeps");
    CK(cudaSetDevice(e->device));
    const size_t bytes = sizeof(int32_t) * (size_t)e->n * kMaxNeighbours;
    if (ensure_stage(e, std::max(bytes, 7 * sizeof(double) * (size_t)e->n))) return 1;
    const SweepArgs sa = sweep_args(e, e->sids, e->n, e->n);
    neighbour_lists_kernel<<<(e->n + 127) / 128, 128, 0, e->stream>>>(sa, e->grid, e->consts, (int32_t *)e->stage,
                                                                      e->flat_density ? e->dlist : nullptr);
    CK(cudaGetLastError());
    return d2h(e, lists, e->stage, bytes);
}
static int get_vec3(SphEngine *e, const float4 *sorted, double *out) {
    CK(cudaSetDevice(e->device));
    if (ensure_stage(e, 7 * sizeof(double) * (size_t)e->n)) return 1;
    unsort_vec3_kernel<<<(e->n + 255) / 256, 256, 0, e->stream>>>(sorted, e->sids, (double *)e->stage, e->n);
    CK(cudaGetLastError());
    return d2h(e, out, e->stage, 3 * sizeof(double) * (size_t)e->n);
}
int sph_get_forces(sph_handle_t e, double *force) {
    if (!e || !force) return fail("null argument");
    if (e->steps_done == 0) return fail("no step has run yet");
    return get_vec3(e, e->sforce, force);
}
int sph_get_terms(sph_handle_t e, double *pressure, double *viscosity) {
    if (!e) return fail("null handle");
    if (!e->spress) return fail("create the engine with SPH_FLAG_RECORD_TERMS");
    if (e->steps_done == 0) return fail("no step has run yet");
    if (pressure && get_vec3(e, e->spress, pressure)) return 1;
    if (viscosity && get_vec3(e, e->svisc, viscosity)) return 1;
    return 0;
}
int sph_get_rng_states(sph_handle_t e, uint64_t *states) {
    if (!e || !states) return fail("null argument");
    if (!e->rng) return fail("rng states exist in PIPE mode only");
    return d2h(e, states, e->rng, 2 * sizeof(uint64_t) * (size_t)e->rng_count);
}
int sph_set_rng_states(sph_handle_t e, const uint64_t *states) {
    if (!e || !states) return fail("null argument");
    if (!e->rng) return fail("rng states exist in PIPE mode only");
    CK(cudaSetDevice(e->device));
    CK(cudaMemcpyAsync(e->rng, states, 2 * sizeof(uint64_t) * (size_t)e->rng_count, cudaMemcpyHostToDevice,
                       e->stream));
    CK(cudaStreamSynchronize(e->stream));
    return 0;
}
int sph_get_stats(sph_handle_t e, SphStats *st) {
    if (!e || !st) return fail("null argument");
    CK(cudaSetDevice(e->device));
    CK(cudaMemsetAsync(e->stats_d, 0, 4 * sizeof(uint32_t), e->stream));
    stats_kernel<<<(e->n + 255) / 256, 256, 0, e->stream>>>(e->pos_m, e->vel_m, e->n, e->stats_d);
    CK(cudaGetLastError());
    uint32_t h[4];
    int2 dead{0, 0};
    CK(cudaMemcpyAsync(h, e->stats_d, sizeof(h), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaMemcpyAsync(&dead, e->cell_range + e->grid.ncells, sizeof(int2), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    st->n_particles = e->n;
    st->n_dead = dead.y - dead.x;
    st->n_nonfinite = (int32_t)h[0];
    st->n_cells = e->grid.ncells;
    memcpy(&st->max_density, &h[1], 4);
    memcpy(&st->max_speed, &h[2], 4);
    st->steps_done = e->steps_done;
    return 0;
}
int64_t sph_n_cells(sph_handle_t e) { return e ? e->grid.ncells : -1; }
int sph_path_counters(sph_handle_t e, int32_t *out4) {
    if (!e || !out4) return fail("null argument");
    CK(cudaSetDevice(e->device));
    int h[3] = {0, 0, 0};
    CK(cudaMemcpyAsync(h, e->refused, sizeof(h), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    out4[0] = h[0];
    out4[1] = h[1];
    out4[2] = h[2];
    out4[3] = e->ntiles_rb;
    return 0;
}
int sph_cell_dims(sph_handle_t e, int32_t *ceil3, int32_t *trunc3) {
    if (!e) return fail("null handle");
    for (int d = 0; d < 3; ++d) {
        if (ceil3) ceil3[d] = e->ceil_dims[d];
        if (trunc3) trunc3[d] = e->trunc_dims[d];
    }
    return 0;
}
int sph_device_ptr(sph_handle_t e, int32_t which, void **ptr, int64_t *n_elements) {
    if (!e || !ptr) return fail("null argument");
    void *q = nullptr;
    switch (which) {
        case 0: q = e->pos_m; break;
        case 1: q = e->vel_m; break;
        case 2: q = e->sids; break;
        case 3: q = e->spos; break;
        case 4: q = e->gid; break;
        case 5: q = e->rng; break;
        case 6: q = e->slab_counters; break;
        default: return fail("unknown buffer id");
    }
    *ptr = q;
    if (n_elements) *n_elements = e->n;
    return 0;
}
int64_t sph_launch_count(sph_handle_t e) { return e ? e->launches : -1; }
