/* include/sph_b200.h -- C ABI of the B200-native SPH step (libsph_b200.so).
 *
 * This is the drop-in boundary for ONE path of Allkams/Fluid-Simulation-3D: the
 * per-timestep update Physics::Fluid::FluidSimulation::Update(dt)
 * (engine/physics/physicsWorld.cc:39-111) and the neighbour structure it builds
 * (UpdateSpatialLookup, :466-498).  The reference has no FFI layer; its seam is
 * the C++ class in engine/physics/physicsWorld.h:32-80.  Every entry point below
 * names the reference member it replaces.  The C++ class of the same shape that
 * sits on top of this ABI is fluid-simulation-3d_b200/host/FluidSimulation.h;
 * INTEGRATION.md shows the binding a maintainer of the reference would add.
 *
 * Conventions: plain C, opaque handle, int status (0 = SPH_OK); the message of
 * the last failure is sph_last_error(ctx).  Host buffers are caller-owned,
 * tightly packed fp32 in the reference's AoS layouts (vec3 = 3 floats, vec4 = 4,
 * vec2 = 2) and always in ORIGINAL PARTICLE INDEX order, whatever order the
 * device keeps internally.  One host thread per context.  All device work of a
 * context is ordered on one CUDA stream; calls that fill host memory return
 * after the data is there.  There is no CPU fallback: without a CUDA device
 * sph_create fails with SPH_ERR_CUDA.
 */
#ifndef SPH_B200_H
#define SPH_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPH_B200_ABI_VERSION 1

enum SphStatus {
    SPH_OK = 0,
    SPH_ERR_INVALID = 1,   /* bad argument / bad state */
    SPH_ERR_CUDA = 2,      /* CUDA runtime failure (message holds cudaGetErrorString) */
    SPH_ERR_NCCL = 3,      /* NCCL failure */
    SPH_ERR_CAPACITY = 4,  /* particle count exceeds the context's capacity */
    SPH_ERR_UNSUPPORTED = 5
};

/* Solver parameters = the private data members of the reference class with their
 * in-class defaults (physicsWorld.h:96-106,145), set through its setters
 * (physicsWorld.cc:214-302). */
typedef struct SphParams {
    float   interaction_radius;        /* interactionRadius  0.35 : cell size and kernel support */
    float   sqr_radius;                /* sqrRadius          0.35f*0.35f : the d^2 cull. A separate
                                          constant in the reference (const member, never derived
                                          from the radius) -- kept separate here on purpose. */
    float   target_density;            /* TargetDensity      99.7 */
    float   pressure_multiplier;       /* pressureMultiplier 300  */
    float   near_pressure_multiplier;  /* nearPressureMultiplier 20 */
    float   viscosity_strength;        /* viscosityStrength  0.5  */
    float   gravity_scale;             /* gravityScale       10   */
    int32_t gravity;                   /* gravity            false */
    float   bound[3];                  /* BoundScale         (20,20,20): full box extents, centred on 0 */
} SphParams;

/* Solver features from the reference's TODO list (README.md:38,43) that it has not implemented: SURVEY 8(f) rank 4.
 * Both are OFF by default (identity rotation, zero strength) and then S6 is the reference's, bit for bit.
 *  - rotatable bound: the unused members boundRotation / boundTransform (physicsWorld.h:146-147).  The box BoundScale is
 *    rotated about the origin by the unit quaternion (x, y, z, w); S6's clamp and -0.95 reflection (physicsWorld.cc:88-106)
 *    run on the position / velocity expressed in the box's own axes.  Gravity stays along world -y.  The GRID table
 *    covers the rotated box's bounding box.
 *  - stickiness ("Apply stickyness to the particles to mimic water better"): the wall adhesion impulse of the double-density
 *    relaxation scheme the reference follows (README.md:64): a particle closer than stick_distance to a wall of the box
 *    gets  v -= dt * stick_strength * d * (1 - d / stick_distance) * n  with d its distance to that wall and n the wall's
 *    inward normal, before S6 moves it. */
typedef struct SphExtras {
    float bound_rotation[4];           /* unit quaternion (x, y, z, w); (0,0,0,1) = the reference's axis-aligned box */
    float stick_strength;              /* 0 = off */
    float stick_distance;              /* > 0 when stick_strength != 0 */
} SphExtras;

/* Which neighbour table the step builds and walks.
 *  SPH_TABLE_GRID            B200 layout (default): cell key = linear index of the reference's
 *                            integer cell floor(pred/r) inside the bounding box, x fastest, so the
 *                            3 x-adjacent cells of each of the 9 (y,z) rows are ONE contiguous run
 *                            of the sorted arrays and neighbouring particles sit in neighbouring
 *                            memory.  Same neighbour sets as the reference.
 *  SPH_TABLE_REFERENCE_HASH  the reference's own structure on the device: key = HashCell % N
 *                            (physicsWorld.cc:505-516), table length N, bucket walk with the
 *                            float-rounded hash filter (:343-346).  Bit-exact keys / sorted order /
 *                            start table; scatters neighbouring cells across memory. */
enum SphTableMode { SPH_TABLE_GRID = 0, SPH_TABLE_REFERENCE_HASH = 1 };

/* Arrays sph_download can return (original particle index order). */
enum SphField {
    SPH_FIELD_POSITIONS = 0,      /* float[N][3]  FluidSimulation::positions      (physicsWorld.h:79) */
    SPH_FIELD_OUT_POSITIONS = 1,  /* float[N][4]  FluidSimulation::OutPositions, w = 0.34 (:80, .cc:107) */
    SPH_FIELD_VELOCITIES = 2,     /* float[N][3]  velocity            (getVelocity, .cc:155-159) */
    SPH_FIELD_DENSITIES = 3,      /* float[N][2]  densities (rho, near rho) (getDensity/getNearDensity, .cc:161-170) */
    SPH_FIELD_PREDICTED = 4,      /* float[N][3]  predictedPositions of the last step (.cc:47) */
    SPH_FIELD_VEL_AFTER_PRESSURE = 5,   /* float[N][3] velocity after CalculatePressureForce (.cc:421) */
    SPH_FIELD_VEL_AFTER_VISCOSITY = 6,  /* float[N][3] velocity after CalculateViscosityForce (.cc:463), before collision */
    SPH_FIELD_HASH = 7,           /* uint32[N]    HashCell(PositionToCellCoord(pred))  (.cc:477-478) */
    SPH_FIELD_KEY = 8,            /* uint32[N]    hash % N                              (.cc:479) */
    SPH_FIELD_NEIGHBOUR_COUNT = 9,/* uint32[N]    candidates surviving every filter of CalculateDensity (.cc:339-357), incl. self */
    SPH_FIELD_SPEED_NORMALIZED = 10, /* float[N]  getSpeedNormalzied (.cc:178-182) */
    SPH_FIELD_COLORS = 11         /* float[N][4]  speed gradient of FluidSimCPU::updateColors (fluidSimCPU.cc:100-125) */
};

/* Tables sph_download_table can return (sorted sequence / table order, not particle order). */
enum SphTable {
    SPH_TABLE_SORTED_INDEX = 0,   /* uint32[N]  particle index of each sorted row  (spatialLookup[].x, .cc:480-484) */
    SPH_TABLE_SORTED_KEY = 1,     /* uint32[N]  key of each sorted row             (spatialLookup[].z) */
    SPH_TABLE_START_INDICES = 2,  /* uint32[len] REFERENCE_HASH: startIndices, len N, 0x7FFFFFFF = empty (.cc:481,486-496)
                                                 GRID: prefix table, len cells+1: rows of cell c are [t[c], t[c+1]) */
    SPH_TABLE_SORTED_HASH = 3     /* uint32[N]  exact u32 hash of each sorted row  (spatialLookup[].y before float rounding) */
};

typedef struct SphContext SphContext;

/* -- lifetime ------------------------------------------------------------- */
/* getInstance() (.cc:32-37): one context owns every device buffer for up to `capacity` particles. */
int  sph_create(SphContext** out, int device, uint32_t capacity);
int  sph_destroy(SphContext* ctx);
const char* sph_last_error(const SphContext* ctx);   /* ctx may be NULL: error of the last failed sph_create */
int  sph_abi_version(void);

/* -- parameters (setters/getters .cc:214-302) ------------------------------ */
void sph_default_params(SphParams* p);
int  sph_set_params(SphContext* ctx, const SphParams* p);
/* the optional features above; rejected (state unchanged) for a zero quaternion, a negative strength or a distance <= 0 with a
 * non-zero strength.  The quaternion is normalised. */
int  sph_set_extras(SphContext* ctx, const SphExtras* e);
int  sph_get_extras(const SphContext* ctx, SphExtras* e);
int  sph_get_params(const SphContext* ctx, SphParams* p);
int  sph_set_table_mode(SphContext* ctx, int mode);
int  sph_get_table_mode(const SphContext* ctx);
/* 1 (default): record the six stage timers with CUDA events each step (getElapsedTime*, .cc:184-212) */
int  sph_set_stage_timing(SphContext* ctx, int enabled);
/* 1: also keep neighbour counts during the density pass (debug tap, off by default) */
int  sph_set_neighbour_count_tap(SphContext* ctx, int enabled);
/* Entries per particle of the neighbour list the density pass records for the pressure and viscosity
 * passes (0 = no list, every pass walks the table).  A particle with more neighbours than this is still
 * exact: the later passes walk the table for it.  Without this call the capacity starts at 64 and grows
 * by itself when the density pass meets denser particles. */
int  sph_set_neighbour_list_capacity(SphContext* ctx, uint32_t entries);

/* -- state ---------------------------------------------------------------- */
/* InitializeData(n) (.cc:112-147): cube lattice spawn (GridArrangement :518-557), velocities zero,
 * then the initial lookup + densities.  */
int  sph_spawn_grid(SphContext* ctx, uint32_t n);
/* Device-side scene spawn (no host array, no upload): an nx*ny*nz lattice block, particle id = (iy*nx + ix)*nz + iz
 * in the reference's fill order (GridArrangement :526-530: y outer from the TOP layer down, then x, then z):
 *   pos = fp32(origin + (ix, ny-1-iy, iz) * gap)  [fp64 lattice, rounded once]
 *         + (u(seed, id, axis) - 0.5) * jitter_amp,     vel = (u(seed ^ 0x5EED, id, axis) - 0.5) * velocity_scale
 * with u = top 24 bits of splitmix64(seed ^ (3*id + axis)) / 2^24: counter-based, so the host generator
 * (fluid-simulation-3d_b200/scenes.py) produces the same bits.  Velocities are zero when velocity_scale is 0.
 * Asynchronous like sph_step.  Dam-break / column emitters are this call with the origin chosen against the bounds. */
typedef struct SphBlockSpawn {
    uint32_t nx, ny, nz;
    uint32_t reserved;        /* 0 */
    double   gap;             /* lattice spacing (reference spawn gap: 0.215, physicsWorld.cc:140) */
    double   origin[3];       /* position of the lattice site with the smallest x, y, z */
    float    jitter_amp;      /* full width of the uniform jitter per axis (0 = regular lattice) */
    float    velocity_scale;  /* full width of the uniform initial velocity per axis (0 = at rest) */
    uint64_t seed;
} SphBlockSpawn;
int  sph_spawn_block(SphContext* ctx, const SphBlockSpawn* block);
/* Replace the particle state: positions[N][3], velocity[N][3] (NULL = zeros). */
int  sph_upload_state(SphContext* ctx, uint32_t n, const float* pos3, const float* vel3);
uint32_t sph_num_particles(const SphContext* ctx);

/* -- the hot path ---------------------------------------------------------- */
/* Update(deltatime) (.cc:39-111): S1 predict, S2 lookup, S3 density, S4 pressure, S5 viscosity
 * (snapshot semantics), S6 integrate + box collision.  Asynchronous: returns when
 * the work is enqueued; sph_download / sph_synchronize wait for it. */
int  sph_step(SphContext* ctx, float dt);
/* nsteps consecutive Update(dt) calls (sub-stepping, offline pre-computation).  From the second step on the launch
 * sequence of a step is replayed as ONE CUDA graph launch per step (the step is launch-bound at the reference's own
 * scene sizes); results are bit-identical to nsteps calls of sph_step.  The last step runs plainly, so the stage
 * timers describe it.  Environment SPH_GRAPH=0 disables the replay. */
int  sph_step_n(SphContext* ctx, float dt, uint32_t nsteps);
/* steps executed by graph replay since creation (diagnostic) */
uint64_t sph_graph_replays(const SphContext* ctx);
/* sph_step / sph_step_n replay the step as a CUDA graph from the second consecutive step of an unchanged configuration on
 * (one cudaGraphLaunch instead of a dozen kernel launches).  The recording carries no stage timers: while the timers are on
 * (sph_set_stage_timing, the default) every 16th step runs through plain launches and refreshes them, and sph_get_timings
 * reports the last such step.  On by default; enabled = 0 makes every step a sequence of plain launches. */
int  sph_set_graph_replay(SphContext* ctx, int enabled);
/* GRID table, counting sort: cells that held more than 16384 rows (a blow-up clamping much of the scene into one rim cell) keep
 * the arrival order of their rows instead of the canonical ascending-index order: neighbour sets are unaffected, float sums lose
 * run-to-run reproducibility in the last bits.  Number of such cells over all steps so far (0 in any sane scene); synchronises. */
uint64_t sph_noncanonical_cells(SphContext* ctx);
/* Rows of the density pass's per-thread survivor stack in the last step: 24, or 72 once the neighbour lists the pass itself
 * measures are long throughout (mean above 56 rows per warp; back to 24 below 40).  A tuning tap: results do not depend on it. */
int  sph_density_stack_rows(const SphContext* ctx);
int  sph_synchronize(SphContext* ctx);
/* rebuild lookup + densities for the current positions without advancing (InitializeData's tail, .cc:144-145) */
int  sph_refresh_densities(SphContext* ctx);

/* -- read-back ------------------------------------------------------------- */
int  sph_download(SphContext* ctx, int field, void* host, size_t host_bytes);
int  sph_download_table(SphContext* ctx, int table, void* host, size_t host_bytes, size_t* out_len);
/* getPosition/getVelocity/getDensity/getNearDensity/getSpeed/getSpeedNormalzied (.cc:149-182):
 * out10 = pos xyz, vel xyz, rho, near rho, speed, speed normalised; all zero when index >= N. */
int  sph_get_particle(SphContext* ctx, uint32_t index, float* out10);
/* ms: gravity(predict), spatial, density, pressure, viscosity, position+collision (.cc:184-212) */
int  sph_get_timings(SphContext* ctx, double* out6);
/* kernels launched by this context since creation (bench.py's gpu_launches) */
uint64_t sph_launch_count(const SphContext* ctx);
/* the context's cudaStream_t (as void*), so a caller can order its own events / copies against the step */
void* sph_stream(const SphContext* ctx);
/* grid-table geometry: dims[3], origin cell[3] (for tests) */
int  sph_get_grid(const SphContext* ctx, int32_t* dims3, int32_t* origin3);
/* GRID table: the cell is subdivided this many times in x (dims3[0] counts the fine cells) */
int  sph_grid_x_subdivision(const SphContext* ctx);

/* -- state snapshots (the reference has none: Reset re-spawns, physicsWorld.cc:112) ---- */
/* Little-endian file: "SPHB2002", u32 n, u32 sizeof(SphParams), SphParams, then n x pos3, n x vel3 (fp32, particle index
 * order); "SPHB2001" files (no size word) are still read.  n is checked against the file size before anything is allocated.
 * Loading replaces parameters and state (n must fit the capacity). */
int  sph_save_state(SphContext* ctx, const char* path);
int  sph_load_state(SphContext* ctx, const char* path);

/* -- host memory helpers ---------------------------------------------------- */
/* Page-lock / unlock a caller-owned host range in place (cudaHostRegister) so uploads and downloads
 * through it run at full PCIe rate; optional, purely a performance hint. */
int  sph_host_register(void* ptr, size_t bytes);
int  sph_host_unregister(void* ptr);

/* -- pipelined transfers ------------------------------------------------------ */
/* The reference's Update() leaves OutPositions in host memory every frame (physicsWorld.cc:107, consumed at
 * fluidSimCPU.cc:58); a caller that also feeds new state every frame pays PCIe twice.  These four calls move both
 * copies off the solver's stream (own staging buffers, own copy streams), so the upload of frame k+1 and the
 * download of frame k-1 overlap the step of frame k:
 *
 *     sph_upload_state_begin(ctx, n, pos, vel);             // H2D starts, returns at once
 *     for (k = 0; k < frames; k++) {
 *         sph_upload_state_commit(ctx);                     // state := upload k, ordered after all enqueued work
 *         if (k + 1 < frames) sph_upload_state_begin(...);  // frame k+1's input travels during frame k
 *         sph_step(ctx, dt);
 *         if (k) sph_download_wait(ctx);                    // frame k-1's result is in host memory now
 *         sph_download_begin(ctx, SPH_FIELD_OUT_POSITIONS, out[k & 1], bytes);
 *     }
 *     sph_download_wait(ctx);
 *
 * Host buffers must stay valid (and should be page-locked: sph_host_register) until the matching commit has been
 * followed by a sph_synchronize / sph_download_wait, resp. until sph_download_wait returns.  One upload and one
 * download may be in flight per context.  Results are bit-identical to the blocking calls.  Single-GPU contexts. */
int  sph_upload_state_begin(SphContext* ctx, uint32_t n, const float* pos3, const float* vel3);
int  sph_upload_state_commit(SphContext* ctx);
int  sph_download_begin(SphContext* ctx, int field, void* host, size_t host_bytes);
int  sph_download_wait(SphContext* ctx);

/* -- slab-decomposed multi-GPU (one context per rank / GPU) ---------------- */
/* Size of the opaque rendezvous blob (an ncclUniqueId). Rank 0 fills it, the host runtime
 * broadcasts it (torch.distributed / MPI / file), every rank passes it to sph_comm_init. */
size_t sph_comm_id_bytes(void);
int  sph_comm_get_id(void* id_out, size_t id_bytes);
int  sph_comm_init(SphContext* ctx, int rank, int nranks, const void* id, size_t id_bytes);
/* Slab planes along z: rank k owns z in [planes[k], planes[k+1]); planes has nranks+1 entries. */
int  sph_comm_set_planes(SphContext* ctx, const float* planes);
/* Upload this rank's OWNED particles with their global ids. */
int  sph_upload_owned(SphContext* ctx, uint32_t n, const uint32_t* global_id, const float* pos3, const float* vel3);
/* Download this rank's owned particles (device order): ids + requested field. */
int  sph_download_owned(SphContext* ctx, int field, uint32_t* global_id, void* host, size_t host_bytes, uint32_t* out_n);
/* The same download WITHOUT the gather on the host: the export kernel writes every owned row straight to
 * host_base[global id] (element = the field's size), so after all ranks have called it the caller's array is complete
 * in particle index order -- no staging copy, no ids, no host-side scatter.  host_base (host_elems elements, at least
 * the largest id + 1) must be page-locked and device-mapped: sph_host_register it first.  Returns when this rank's rows
 * are in host memory.  Ranks write disjoint elements, so they may call it concurrently from their own threads. */
int  sph_download_owned_scatter(SphContext* ctx, int field, void* host_base, size_t host_elems, uint32_t* out_n);
/* Pipelined forms of the two calls above (same rules as sph_upload_state_begin / sph_download_begin; committed with
 * sph_upload_state_commit, completed with sph_download_wait).  out_n is this rank's owned count at the time of the call. */
int  sph_upload_owned_begin(SphContext* ctx, uint32_t n, const uint32_t* global_id, const float* pos3, const float* vel3);
int  sph_download_owned_begin(SphContext* ctx, int field, uint32_t* global_id, void* host, size_t host_bytes, uint32_t* out_n);
/* counters of the last step: owned, ghosts received (lo, hi), migrated out (lo, hi) */
int  sph_comm_stats(const SphContext* ctx, uint32_t* out5);

/* -- slab re-balancing (planes at particle-count quantiles, re-balanced every K steps) ---- */
/* The slabs are cut at cell-layer granularity; this returns the nranks+1 global layer indices in force
 * (rank k owns layers [layers[k], layers[k+1]); layer 0 is the first layer of the grid table, sph_get_grid). */
int  sph_comm_get_layers(const SphContext* ctx, int32_t* layers_out);
/* COLLECTIVE (every rank calls it between the same two steps).  Builds the global histogram of particles per
 * z layer from the tables of the last step (one tiny kernel + one ncclAllReduce of dims[2] counters), cuts it at the
 * particle-count quantiles with sph_slab_balance_layers below -- every plane moves by at most max_shift layers
 * (capped at 3, and at what the exchange buffers hold) -- and puts the new planes in force.  The next sph_step
 * migrates the rows of the layers that changed owner through the ordinary migration path; results stay identical
 * to the single-GPU step.  max_shift = 0 only measures.  layers_out (nranks+1), hist_out (hist_entries >= dims[2]
 * counters, the global histogram) and changed_out may be NULL.  Needs a completed sph_step. */
int  sph_comm_rebalance(SphContext* ctx, uint32_t max_shift, int32_t* layers_out, uint32_t* hist_out,
                        size_t hist_entries, int* changed_out);
/* Pure host function (no device, no context): cut hist[0, gz) into nranks contiguous runs of layers with particle
 * counts as equal as layer granularity allows, every run at least three layers thick (the slab protocol's minimum).
 * With layers_old != NULL every plane stays within max_shift layers of its old place, never crosses its old
 * neighbours (migration is single-hop), moves at most row_budget particles (0: unlimited) and moves only if that
 * brings it nearer to its quantile by at least a quarter of the rows moved (hysteresis: no flipping between the two
 * boundaries of a layer the quantile falls into); with layers_old == NULL the cut is unconstrained.  Deterministic: every rank derives the same planes from the same histogram. */
/* Pure host function: the verdict both ends of a slab link reach about it from the two 8-word count messages that crossed it
 * (migrants, ghosts, kept migrants, status, exchange-buffer rows, free rows, free ghost rows, 0): 1 = the link carries its
 * payloads and halos this step, 0 = both ends skip it and return a capacity error after the step, -1 = NULL argument.
 * Symmetric in its arguments. */
int  sph_slab_link_ok(const uint32_t* mine8, const uint32_t* theirs8);
int  sph_slab_balance_layers(const uint32_t* hist, int32_t gz, int32_t nranks, const int32_t* layers_old,
                             uint32_t max_shift, uint64_t row_budget, int32_t* layers_new);

#ifdef __cplusplus
}
#endif
#endif /* SPH_B200_H */
