/* oracle/sph_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C restatement of the reference's per-timestep SPH update
 * (Physics::Fluid::FluidSimulation::Update, engine/physics/physicsWorld.cc:39-111,
 * 304-557 and engine/physics/kernels.h:25-82).  Pinned bit-for-bit against the
 * UNMODIFIED reference TU (oracle/_ref/libsph_ref.so) by tests/test_oracle.py and
 * against the committed fixtures under tests/golden/.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use it.
 */
#ifndef SPH_ORACLE_H
#define SPH_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct OracleParams {
    float interaction_radius;        /* physicsWorld.h:97  (0.35)  */
    float sqr_radius;                /* physicsWorld.h:96  const 0.35f*0.35f, NOT derived from the radius (Q2) */
    float target_density;            /* :98  (99.7) */
    float pressure_multiplier;       /* :99  (300)  */
    float near_pressure_multiplier;  /* :100 (20)   */
    float viscosity_strength;        /* :101 (0.5)  */
    float gravity_scale;             /* :106 (10)   */
    int   gravity;                   /* :105 (false) */
    float bound[3];                  /* :145 (20,20,20) */
} OracleParams;

typedef struct Oracle Oracle;

Oracle* oracle_create(int n);
void    oracle_destroy(Oracle* o);
void    oracle_default_params(OracleParams* p);
void    oracle_set_params(Oracle* o, const OracleParams* p);
void    oracle_set_threads(int nthreads);           /* OpenMP team size for the par-for stages */
void    oracle_set_wide_lookup(Oracle* o, int wide);/* 1: keep (index,hash,key) as u32 instead of the reference's floats (defined above 2^24) */

void    oracle_spawn_grid(Oracle* o);               /* InitializeData :112-147 incl. initial lookup + densities */
void    oracle_set_state(Oracle* o, const float* pos3, const float* vel3);

void    oracle_set_predicted(Oracle* o, const float* pred3);   /* slab model: inject ghost rows */
void    oracle_set_densities(Oracle* o, const float* dens2);

/* stages */
void    oracle_stage_predict(Oracle* o, float dt);                    /* :42-48  */
void    oracle_stage_spatial(Oracle* o, const uint32_t* forced_order);/* :466-498; forced_order (particle ids in sorted sequence) replaces the sort's tie order */
void    oracle_stage_density(Oracle* o);                              /* :304-311, 325-365 */
void    oracle_stage_pressure(Oracle* o, float dt);                   /* :367-422 */
void    oracle_stage_viscosity(Oracle* o, float dt, int jacobi);      /* :424-464 */
void    oracle_stage_integrate(Oracle* o, float dt);                  /* :81-108 */
void    oracle_step(Oracle* o, float dt, int jacobi);                 /* S1..S6 */

/* read-back (original particle index order) */
int     oracle_num_particles(const Oracle* o);
void    oracle_get_positions(const Oracle* o, float* out3);
void    oracle_get_out_positions(const Oracle* o, float* out4);
void    oracle_get_velocities(const Oracle* o, float* out3);
void    oracle_get_predicted(const Oracle* o, float* out3);
void    oracle_get_densities(const Oracle* o, float* out2);
void    oracle_get_vel_after_pressure(const Oracle* o, float* out3);
void    oracle_get_vel_after_viscosity(const Oracle* o, float* out3);
void    oracle_get_hash_key(const Oracle* o, uint32_t* hash, uint32_t* key, int32_t* cell3);
void    oracle_get_sorted(const Oracle* o, uint32_t* idx, uint32_t* hash_as_stored, uint32_t* key);
void    oracle_get_start_indices(const Oracle* o, uint32_t* out);
void    oracle_get_neighbour_counts(const Oracle* o, uint32_t* out);
/* per-particle tolerance scales: sum of |term| of the pressure / viscosity sums */
void    oracle_get_force_scales(const Oracle* o, float dt, float* pressure_scale, float* viscosity_scale);
void    oracle_get_timings(const Oracle* o, double* out6);

void    oracle_kernels(float dist, float radius, float* out5);         /* kernels.h:25-82 */

#ifdef __cplusplus
}
#endif
#endif
