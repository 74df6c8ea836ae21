/*
 * sph_b200.h -- C ABI of the B200-native SPH step engine (libsph_b200.so).
 *
 * This is the drop-in boundary for ONE hot path of iwoplaza/cuda-sph: the body of
 *     VoxelSPHStrategy.compute_next_state(old_state) -> new_state
 * (reference: sim/src/sph/strategies/abstract_sph_strategy.py:31-46 and voxel_sph_strategy.py:19-116, kernels in
 * sim/src/sph/kernels/{voxel,base,util}_kernels.py).  Plain pointers and sizes only; no torch / numpy types.
 * The reference binds it from Python with ctypes; the host-side mirror of the reference's
 * strategy interface lives in cuda_sph_b200/strategy.py.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; sph_last_error() gives the message
 *     (the reference raises Python exceptions: thread_layout.py:30-31, numba CudaAPIError).
 *   - host arrays are C-contiguous; positions / velocities are (N,3) fp64 exactly like SimulationState
 *     (common/data_classes.py:81-85).  The engine computes in fp32 (+ fp64 where parity needs it) and never keeps a
 *     host pointer after the call returns (reference: copies to device, abstract_sph_strategy.py:66-67).
 *   - one handle == one GPU == one CUDA stream; a handle is not thread-safe (the reference is single-threaded and
 *     synchronises after every launch: abstract_sph_strategy.py:98,115).
 *   - step calls are asynchronous on the handle's stream; download / get_* calls synchronise.
 */
#ifndef SPH_B200_H
#define SPH_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPH_MODE_BOX 0   /* config.SIM_MODE == 'BOX'  -> collision_kernel_box  (base_kernels.py:75-98)  */
#define SPH_MODE_PIPE 1  /* config.SIM_MODE == 'PIPE' -> collision_kernel      (base_kernels.py:56-72)  */

#define SPH_FLAG_RECORD_NEIGHBOUR_COUNTS 1u /* accepted for compatibility: neighbour counts are always kept (1 B/particle) */
#define SPH_FLAG_RECORD_TERMS 2u            /* force sweep also stores the pressure and viscosity terms (parity tap)  */
#define SPH_FLAG_NO_GRAPH 4u                /* launch kernels eagerly instead of replaying a CUDA graph             */
#define SPH_FLAG_SLAB 8u                    /* x-slab mode (one handle per GPU of a multi-GPU run); particle_count is the
                                               CAPACITY of the local arrays (owned + ghost particles)              */

/* Replaces: SimulationParameters (common/data_classes.py:70-78) + the module constants of config.py:18-36 that the
 * reference freezes into its kernels at import time (base_kernels.py:2, voxel_kernels.py:5). */
typedef struct SphParams {
    int32_t particle_count;    /* params.particle_count                                   */
    int32_t mode;              /* SPH_MODE_BOX | SPH_MODE_PIPE   (config.py:13)           */
    double h;                  /* INF_R    config.py:20                                   */
    double mass;               /* MASS     config.py:18                                   */
    double rho0;               /* RHO_0    config.py:19                                   */
    double k;                  /* K        config.py:22                                   */
    double visc;               /* VISC     config.py:21                                   */
    double damp;               /* DAMP     config.py:23                                   */
    double dt;                 /* 1 / params.fps   abstract_sph_strategy.py:20            */
    double external_force[3];  /* params.external_force                                   */
    double space_size[3];      /* params.space_size                                       */
    double voxel_size[3];      /* params.voxel_size                                       */
    int32_t max_neighbours;    /* MAX_NEIGHBOURS config.py:30; only 32 is supported       */
    uint32_t flags;            /* SPH_FLAG_*                                              */
    uint64_t rng_seed;         /* 16435234 in abstract_sph_strategy.py:27                 */
} SphParams;

typedef struct SphEngine *sph_handle_t;

/* Per-stage device time of the most recent sph_step_timed() call, milliseconds, summed over its steps. */
typedef struct SphTimings {
    float hash_ms;     /* assign_voxels_to_particles_kernel        voxel_kernels.py:88-105        */
    float sort_ms;     /* host structured sort                     voxel_sph_strategy.py:81-88    */
    float reorder_ms;  /* __populate_voxel_begins + SoA gather     voxel_sph_strategy.py:98-107   */
    float density_ms;  /* density_kernel                           voxel_kernels.py:108-132       */
    float force_ms;    /* pressure + viscosity + integrate + collide  voxel_kernels.py:135-211, base_kernels.py:30-98 */
    float total_ms;
    int32_t steps;
    int32_t launches_per_step; /* kernels + memsets this library launches per step */
    int32_t sort_passes;       /* 8-bit digit passes of the radix sort (ceil(log2(n_cells + 1)) bits)  */
} SphTimings;

typedef struct SphStats {
    int32_t n_particles;
    int32_t n_dead;       /* particles in the "dead cell" (non-finite / out-of-table position) */
    int32_t n_nonfinite;  /* particles with a non-finite position or velocity component                      */
    int32_t n_cells;
    float max_density;    /* analize.py:14 */
    float max_speed;
    int64_t steps_done;
} SphStats;

const char *sph_last_error(void);
int sph_version(void);

/* Replaces AbstractSPHStrategy.__init__ (abstract_sph_strategy.py:18-29): allocates device state, creates the
 * xoroshiro128+ states (PIPE mode; numba create_xoroshiro128p_states(grid*block, seed)). */
int sph_create(const SphParams *params, int device, sph_handle_t *out);
int sph_destroy(sph_handle_t h);

/* Replaces cuda.to_device(params.pipe.to_numpy()) (abstract_sph_strategy.py:74): rows x 5 table
 * [x_start, y_c, z_c, r_start, length], last row = pipe end (common/data_classes.py:34-44). */
int sph_set_pipe(sph_handle_t h, const double *table, int32_t rows);

/* Use an existing CUDA stream (e.g. torch.cuda.current_stream().cuda_stream) instead of the engine's own. */
int sph_set_stream(sph_handle_t h, void *cuda_stream);

/* Replaces _send_arrays_to_gpu (abstract_sph_strategy.py:64-74): position, velocity (N,3) fp64 host arrays. */
int sph_upload(sph_handle_t h, const double *position, const double *velocity);
/* Same, from fp32 (N,3) host arrays (start states generated in fp32). */
int sph_upload_f32(sph_handle_t h, const float *position, const float *velocity);

/* n device-resident steps, no host round trip: hash -> sort -> cell table + SoA reorder -> density sweep ->
 * fused pressure/viscosity/integrate/collide sweep.  Replaces the body of compute_next_state (:31-46) n times. */
int sph_step(sph_handle_t h, int32_t n_steps);
/* Same, launched eagerly with CUDA events around every stage; fills *t. */
int sph_step_timed(sph_handle_t h, int32_t n_steps, SphTimings *t);

/* Replaces __finalize_computation (abstract_sph_strategy.py:76-84): (N,3),(N,3),(N,) fp64, particle-id order.
 * Any pointer may be NULL. */
int sph_download(sph_handle_t h, double *position, double *velocity, double *density);
int sph_download_f32(sph_handle_t h, float *position, float *velocity, float *density);

/* The reference-facing call: upload + one step + download == compute_next_state(old_state). */
int sph_compute_next_state(sph_handle_t h, const double *pos_in, const double *vel_in, double *pos_out,
                           double *vel_out, double *density_out);

int sph_sync(sph_handle_t h);

/* Device-side snapshot / restore of the particle state (position, velocity, density, rng states, step counter).
 * No reference counterpart (the reference restarts from config.start_state, sim/src/main.py:14); used for
 * checkpoint/resume and by bench.py to keep a workload inside its first steps. */
int sph_save_state(sph_handle_t h);
int sph_restore_state(sph_handle_t h);

/* ---- x-slab domain decomposition (no reference counterpart: the reference is single-GPU) --------------------------
 * The handle owns the cell columns [x_lo, x_hi) of the global grid and keeps a two-column ghost halo on each side.
 * The host layer (cuda_sph_b200/slab.py, torch.distributed) writes owned particles to local indices [0, n_own) and
 * ghost particles to [n_own, n_local) of the master arrays (sph_device_ptr 0/1) and their global ids to buffer 4,
 * then calls sph_slab_step: hash -> sort -> in-cell order by global id -> cell table + reorder -> density for owned
 * and first-ghost columns -> forces / integrate / collide for owned particles only.  n_global sizes the xoroshiro
 * state table (PIPE mode; states are indexed by global particle id). */
int sph_slab_configure(sph_handle_t h, int32_t x_lo, int32_t x_hi, int64_t n_global);
int sph_slab_step(sph_handle_t h, int32_t n_own, int32_t n_local);

/* Native exchange path (csrc/slab_exchange.cuh): nothing in the step loop synchronises with the host.
 * Slots [0, own_cap) of the master arrays are the owned region (holes and unused slots are EMPTY: global id -1),
 * [own_cap, particle_count) the ghost region.  bounds[world + 1] are the slab column boundaries of all ranks;
 * cap_migrants / cap_ghosts[world] the record capacities of the block sent to (and received from) each rank --
 * symmetric by construction (adjacent ranks large, the others small).  Returns the device send / receive buffers and the
 * byte size of every block: the host layer moves them with ONE fixed-size all_to_all per step
 * (cuda_sph_b200/slab.py NativeSlabRunner; torch.distributed, NCCL over NVLink).
 *   sph_slab_route     pack migrant + ghost records of the owned particles into the send blocks, open the holes
 *   sph_slab_unpack    empty the ghost region, append received migrants / ghosts
 *   sph_slab_step_all  one local step over the whole capacity (empty slots are dead)
 *   sph_slab_compact   close the holes of the owned region (call every few dozen steps)
 *   sph_slab_counters  out5 = {high-water mark, ghosts, overflow flags, live owned particles, own_cap}; synchronises */
int sph_slab_exchange_init(sph_handle_t h, int32_t world, int32_t rank, const int32_t *bounds, int32_t own_cap,
                           const int32_t *cap_migrants, const int32_t *cap_ghosts, void **sendbuf, void **recvbuf,
                           int64_t *block_bytes);
int sph_slab_route(sph_handle_t h);
int sph_slab_unpack(sph_handle_t h);
int sph_slab_step_all(sph_handle_t h);
int sph_slab_step_all_timed(sph_handle_t h, SphTimings *t); /* same, with per-stage CUDA events; synchronises */
int sph_slab_compact(sph_handle_t h);
int sph_slab_counters(sph_handle_t h, int32_t *out5);

/* Exchange over peer memory (csrc/slab_exchange.cuh "exchange over peer memory"): the receive buffer is double-buffered
 * by exchange parity -- *recvbuf of sph_slab_exchange_init points at buffer 0, buffer 1 follows at
 * (sum(block_bytes) rounded up to 256) bytes; sph_slab_unpack consumes buffer `parity` and flips it, an all_to_all of the
 * host layer must therefore land in buffer sph_slab_parity().  With the receive allocations of all ranks mapped
 * (sph_slab_ipc_handle -> all-gather of the 64-byte CUDA IPC handles -> sph_slab_open_peers; remote_off[2 k + q] = byte
 * offset of the block "from this rank" inside rank k's receive buffers, q = 0 / 1 the parity: ranks differ in size) a
 * step needs no collective:
 *   sph_slab_step_all -> sph_slab_exchange_p2p (CTA-aggregated routing into the send blocks, used records pushed into
 *   the receivers' buffers over NVLink, counts + arrival flags to all peers, bounded wait for theirs) -> sph_slab_unpack;
 * sph_slab_route + all_to_all + sph_slab_unpack remain the portable path (and what the host layer uses without IPC). */
int sph_slab_parity(sph_handle_t h, int32_t *parity);
int sph_slab_ipc_handle(sph_handle_t h, void *handle64);
int sph_slab_open_peers(sph_handle_t h, const void *handles /* world x 64 bytes */, const int64_t *remote_off /* 2 x world */);
int sph_slab_exchange_p2p(sph_handle_t h);
int sph_slab_exchange_p2p_timed(sph_handle_t h, float *ms3); /* + CUDA events: {route, push, flag barrier} ms; synchronises */

/* ---- parity taps: state of the most recent step --------------------------------------------------------------- */
int sph_get_keys(sph_handle_t h, int32_t *keys);                 /* self.voxels            voxel_sph_strategy.py:82 */
int sph_get_sorted_ids(sph_handle_t h, int32_t *ids);            /* voxel_particle_map['particle_id']        :85-88 */
int sph_get_sorted_keys(sph_handle_t h, int32_t *keys);          /* voxel_particle_map['voxel_id']           :85-88 */
int sph_get_voxel_begin(sph_handle_t h, int32_t *begin, int64_t n_cells); /* self.voxel_begin (-1 = empty)  :92-107 */
int sph_get_neighbour_counts(sph_handle_t h, int32_t *counts);   /* get_neighbours return value   voxel_kernels.py:85 */
int sph_get_neighbour_lists(sph_handle_t h, int32_t *lists);     /* N x 32 particle ids, -1 padded: `neighbours` of
                                                                    get_neighbours          voxel_kernels.py:29-85 */
int sph_get_forces(sph_handle_t h, double *force);               /* self.result_force  abstract_sph_strategy.py:83 */
int sph_get_terms(sph_handle_t h, double *pressure, double *viscosity); /* d_new_pressure_term / d_new_viscosity_term */
int sph_get_rng_states(sph_handle_t h, uint64_t *states);        /* N x 2 uint64 (PIPE mode)                        */
int sph_set_rng_states(sph_handle_t h, const uint64_t *states);
int sph_get_stats(sph_handle_t h, SphStats *stats);
int64_t sph_n_cells(sph_handle_t h);
/* Diagnostics (no reference counterpart): work-item counts of the most recent step, out[4] = { 32-particle passes of
 * tiles whose rows do not fit the staging, tiles handed over by density_flat_kernel, dense tiles (density_dense_kernel),
 * number of 128-particle tiles }.  Says which share of a workload runs on which sweep path. */
int sph_path_counters(sph_handle_t h, int32_t *out4);
int sph_cell_dims(sph_handle_t h, int32_t *ceil3, int32_t *trunc3);

/* ---- frame export pipeline (section 8(f)1).  Replaces the per-frame __finalize_computation + np.save sequence of
 * state_generator.py:26-37 / saver.py:28-36 for device-resident runs.  sph_export_begin snapshots the state (fp64,
 * particle-id order, every `stride`-th id: stride > 1 down-samples, e.g. for the viewer's 100 000-point cap,
 * gl_point_field.py:11) into one of three engine-owned PINNED host buffers on a second stream and returns at once, so
 * the copy runs under the next steps; sph_export_wait blocks until that buffer is complete and returns pointers into
 * it (valid until the slot is used again): position (n_out,3), velocity (n_out,3), density (n_out). */
int sph_export_begin(sph_handle_t h, int32_t slot, int32_t stride);
int sph_export_wait(sph_handle_t h, int32_t slot, double **position, double **velocity, double **density,
                    int64_t *n_out);

/* ---- seeded start states generated ON THE DEVICE (section 8(f)2).  Replaces the per-particle Python loops with
 * unseeded `random` of config.py:79-120.  kind: 0 = dam-break column (first 10 % of x, velocity [1.5,-5,-5] +- 0.5,
 * config.py:84-95), 1 = uniform box (same velocities), 2 = inside the pipe (uniform in x, uniform over 98 % of the
 * local disc, zero velocity, config.py:105-115; needs sph_set_pipe).  Counter-based (SplitMix64 of seed, particle id
 * and draw index), so the state is a pure function of (kind, seed, id) and cuda_sph_b200/config.py mirrors it on the
 * host bit for bit. */
#define SPH_GEN_BOX_WALL 0
#define SPH_GEN_UNIFORM 1
#define SPH_GEN_PIPE 2
int sph_generate_state(sph_handle_t h, int32_t kind, uint64_t seed);

/* ---- per-frame reductions on the device (section 8(f)4; analize.py:9-14 prints max position / velocity / density per
 * epoch, np.max over all components).  Maxima are over finite values; non-finite particles are counted. */
typedef struct SphFrameStats {
    int64_t steps_done;
    int32_t n_particles;
    int32_t n_dead;              /* particles in the dead cell at the most recent step       */
    int32_t n_nonfinite;         /* particles with a non-finite position or velocity component              */
    int32_t reserved;
    float max_position;          /* np.max(position)  analize.py:12 */
    float min_position;
    float max_velocity;          /* np.max(velocity)  analize.py:13 */
    float max_speed;
    float max_density;           /* np.max(density)   analize.py:14 */
    int32_t neighbour_hist[33];  /* particles by neighbour count (0..32, self included) of the most recent step */
} SphFrameStats;
int sph_get_frame_stats(sph_handle_t h, SphFrameStats *stats);               /* current state; synchronises */
int sph_export_stats(sph_handle_t h, int32_t slot, SphFrameStats *stats);    /* of the frame exported into `slot` */

/* Device pointers for zero-copy wrapping (torch / __cuda_array_interface__).  which: 0 = master records, 32 bytes per
 * particle in id order: float4 (x,y,z,density) | float4 (vx,vy,vz,0), i.e. float[N][8]; 1 = the velocity half of
 * record 0 (= pointer 0 + 16 bytes, same 32-byte stride); 2 = sorted ids int32[N], 3 = sorted position float4[N],
 * 4 = global ids int32[capacity] (x-slab mode), 5 = xoroshiro states uint64[2 * count] (PIPE mode),
 * 6 = slab counters int32[8] (native exchange: high-water mark, ghosts, overflow, scratch, live). */
int sph_device_ptr(sph_handle_t h, int32_t which, void **ptr, int64_t *n_elements);

/* Total kernels/memsets launched by this handle so far (bench.py's "gpu_launches"). */
int64_t sph_launch_count(sph_handle_t h);

#ifdef __cplusplus
}
#endif
#endif /* SPH_B200_H */
