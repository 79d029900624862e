/* oracle/sph_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C restatement of the reference CPU SPH step.  Every function cites the
 * reference lines it follows (paths relative to the reference root,
 * engine/physics/physicsWorld.cc unless stated).  Expression order mirrors the
 * reference's glm expressions operation by operation so that, compiled for
 * baseline x86-64 without FMA contraction, results are BIT-IDENTICAL to the
 * unmodified reference (oracle/_ref) given the same sorted tie order; that is
 * asserted by tests/test_oracle.py.  Parity: PINNED (against oracle/_ref built
 * from the unmodified reference sources, and the fixtures in tests/golden/).
 *
 * Not used, linked or imported by the product library.
 */
#include "sph_oracle.h"

#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* glm::pi<float>() (exts/glm/ext/scalar_constants.inl) */
#define PI_F ((float)3.14159265358979323846264338327950288)

struct Oracle {
    uint32_t n;
    OracleParams p;
    int wide;
    float *pos, *vel, *pred, *dens, *out4;       /* 3n,3n,3n,2n,4n */
    float *vel_press, *vel_visc;                 /* 3n each */
    /* spatialLookup rows (index, hash, key): floats as the reference (:128,:480) or u32 when wide */
    float *lk_f; uint32_t *lk_u;
    uint32_t *start;                             /* startIndices */
    double t[6];
};

static double now_ms(void)
{
    struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

void oracle_default_params(OracleParams* p)
{   /* physicsWorld.h:96-106,145 */
    p->interaction_radius = 0.35f;
    p->sqr_radius = 0.35f * 0.35f;
    p->target_density = 99.7f;
    p->pressure_multiplier = 300.0f;
    p->near_pressure_multiplier = 20.0f;
    p->viscosity_strength = 0.5f;
    p->gravity_scale = 10.0f;
    p->gravity = 0;
    p->bound[0] = p->bound[1] = p->bound[2] = 20.0f;
}

Oracle* oracle_create(int n)
{
    Oracle* o = (Oracle*)calloc(1, sizeof(Oracle));
    size_t N = (size_t)(n > 0 ? n : 1);
    o->n = (uint32_t)n;
    oracle_default_params(&o->p);
    o->pos = (float*)calloc(3 * N, 4);  o->vel = (float*)calloc(3 * N, 4);
    o->pred = (float*)calloc(3 * N, 4); o->dens = (float*)calloc(2 * N, 4);
    o->out4 = (float*)calloc(4 * N, 4);
    o->vel_press = (float*)calloc(3 * N, 4); o->vel_visc = (float*)calloc(3 * N, 4);
    o->lk_f = (float*)calloc(3 * N, 4); o->lk_u = (uint32_t*)calloc(3 * N, 4);
    o->start = (uint32_t*)calloc(N, 4);
    for (size_t i = 0; i < (size_t)n; i++) { o->out4[4 * i + 3] = 0.25f; o->start[i] = (uint32_t)INT_MAX; }
    return o;
}

void oracle_destroy(Oracle* o)
{
    if (!o) return;
    free(o->pos); free(o->vel); free(o->pred); free(o->dens); free(o->out4);
    free(o->vel_press); free(o->vel_visc); free(o->lk_f); free(o->lk_u); free(o->start);
    free(o);
}

void oracle_set_params(Oracle* o, const OracleParams* p) { o->p = *p; }
void oracle_set_wide_lookup(Oracle* o, int wide) { o->wide = wide; }
void oracle_set_threads(int nthreads)
{
#ifdef _OPENMP
    omp_set_num_threads(nthreads > 0 ? nthreads : 1);
#else
    (void)nthreads;
#endif
}
int oracle_num_particles(const Oracle* o) { return (int)o->n; }

void oracle_set_state(Oracle* o, const float* pos3, const float* vel3)
{
    if (pos3) memcpy(o->pos, pos3, (size_t)o->n * 12);
    if (vel3) memcpy(o->vel, vel3, (size_t)o->n * 12);
}

/* injection points for the slab-decomposition model (tests/slab_model.py): ghost rows get their predicted
 * position, density and post-pressure velocity from the owning rank instead of computing them */
void oracle_set_predicted(Oracle* o, const float* pred3) { memcpy(o->pred, pred3, (size_t)o->n * 12); }
void oracle_set_densities(Oracle* o, const float* dens2) { memcpy(o->dens, dens2, (size_t)o->n * 8); }

/* ---- smoothing kernels: kernels.h:25-82, constants recomputed per call (Q17) ---- */
static inline float SmoothingPow2(float dist, float radius)
{   /* kernels.h:25-34 */
    if (dist < radius) {
        float volume = 15 / (2 * PI_F * powf(radius, 5));
        float v = radius - dist;
        return v * v * volume;
    }
    return 0;
}
static inline float SmoothingPow3(float dist, float radius)
{   /* kernels.h:37-46 */
    if (dist < radius) {
        float volume = 15.0f / (PI_F * powf(radius, 6));
        float v = radius - dist;
        return v * v * v * volume;
    }
    return 0;
}
static inline float SmoothingDerivativePow2(float dist, float radius)
{   /* kernels.h:49-58 */
    if (dist <= radius) {
        float scale = 15.0f / (powf(radius, 5) * PI_F);
        float v = (radius - dist);
        return -v * scale;
    }
    return 0;
}
static inline float SmoothingDerivativePow3(float dist, float radius)
{   /* kernels.h:61-70 */
    if (dist <= radius) {
        float scale = 45 / (powf(radius, 6) * PI_F);
        float v = (radius - dist);
        return -v * v * scale;
    }
    return 0;
}
static inline float SmoothingViscoPoly6(float dist, float radius)
{   /* kernels.h:73-82; abs(radius) with float semantics (Q3) */
    if (dist < radius) {
        float scale = 315 / (64 * PI_F * powf(fabsf(radius), 9));
        float v = radius * radius - dist * dist;
        return v * v * v * scale;
    }
    return 0;
}
void oracle_kernels(float dist, float radius, float* out5)
{
    out5[0] = SmoothingPow2(dist, radius);
    out5[1] = SmoothingPow3(dist, radius);
    out5[2] = SmoothingDerivativePow2(dist, radius);
    out5[3] = SmoothingDerivativePow3(dist, radius);
    out5[4] = SmoothingViscoPoly6(dist, radius);
}

/* ---- cell / hash / key: :499-516 ---- */
static inline void PositionToCellCoord(const Oracle* o, const float* pos, float* cell)
{   /* :499-503  floor(pos / r) with true division (Q6), C-cast to int, stored as float */
    const float r = o->p.interaction_radius;
    cell[0] = (float)(int)floorf(pos[0] / r);
    cell[1] = (float)(int)floorf(pos[1] / r);
    cell[2] = (float)(int)floorf(pos[2] / r);
}
static inline uint32_t f2u_wrap(float x)
{   /* (uint32_t)float on baseline x86-64 = 64-bit cvttss2si then truncation: two's-complement wrap (Q5) */
    return (uint32_t)(int64_t)x;
}
static inline uint32_t HashCell(const float* c)
{   /* :505-511 */
    uint32_t a = f2u_wrap(c[0]) * 15823u;
    uint32_t b = f2u_wrap(c[1]) * 9737333u;
    uint32_t cc = f2u_wrap(c[2]) * 440817757u;
    return a + b + cc;
}
static inline uint32_t GetKeyFromHash(uint32_t hash, uint32_t len) { return hash % len; } /* :513-516 */

/* offsets[27], physicsWorld.h:131-143: x outer, y, z inner, each -1..1 */
static inline void offset27(int i, float* off)
{
    off[0] = (float)(i / 9 - 1); off[1] = (float)((i / 3) % 3 - 1); off[2] = (float)(i % 3 - 1);
}

/* lookup row accessors: the reference keeps (index, hash, key) as three floats */
static inline int row_key_ne(const Oracle* o, uint32_t row, uint32_t key)
{   return o->wide ? (o->lk_u[3 * row + 2] != key) : (o->lk_f[3 * row + 2] != (float)key); }
static inline int row_hash_ne(const Oracle* o, uint32_t row, uint32_t hash)
{   return o->wide ? (o->lk_u[3 * row + 1] != hash) : (o->lk_f[3 * row + 1] != (float)hash); }
static inline int row_index_ge(const Oracle* o, uint32_t row, uint32_t n)
{   return o->wide ? (o->lk_u[3 * row] >= n) : (o->lk_f[3 * row] >= (float)n); }
static inline uint32_t row_index(const Oracle* o, uint32_t row)
{   return o->wide ? o->lk_u[3 * row] : (uint32_t)o->lk_f[3 * row]; }

/* ---- S1: :42-48 with CalculateExternalFoce :313-323 ---- */
void oracle_stage_predict(Oracle* o, float dt)
{
    double t0 = now_ms();
    const uint32_t n = o->n;
    float g[3] = { 0, 0, 0 };
    if (o->p.gravity) { g[0] = 0; g[1] = -o->p.gravity_scale; g[2] = 0; }
    #pragma omp parallel for schedule(static)
    for (uint32_t i = 0; i < n; i++) {
        for (int a = 0; a < 3; a++) {
            o->vel[3 * i + a] += g[a] * dt;
            o->pred[3 * i + a] = o->pos[3 * i + a] + o->vel[3 * i + a] * (1.0f / 120.0f);
        }
    }
    o->t[0] = now_ms() - t0;
}

/* ---- S2: UpdateSpatialLookup :466-498 ---- */
static void radix_sort_pairs(uint32_t n, uint32_t* key, uint32_t* val)
{   /* stable LSD radix sort: equal keys keep ascending particle index = the canonical tie order (Q14) */
    uint32_t* k2 = (uint32_t*)malloc((size_t)n * 4); uint32_t* v2 = (uint32_t*)malloc((size_t)n * 4);
    for (int shift = 0; shift < 32; shift += 8) {
        size_t cnt[257]; memset(cnt, 0, sizeof cnt);
        for (uint32_t i = 0; i < n; i++) cnt[((key[i] >> shift) & 255) + 1]++;
        for (int b = 0; b < 256; b++) cnt[b + 1] += cnt[b];
        for (uint32_t i = 0; i < n; i++) { size_t d = cnt[(key[i] >> shift) & 255]++; k2[d] = key[i]; v2[d] = val[i]; }
        uint32_t* t = key; key = k2; k2 = t; t = val; val = v2; v2 = t;
    }
    free(k2); free(v2);   /* 4 passes: data is back in the caller's arrays */
}

void oracle_stage_spatial(Oracle* o, const uint32_t* forced_order)
{
    double t0 = now_ms();
    const uint32_t n = o->n;
    uint32_t* hash = (uint32_t*)malloc((size_t)n * 4);
    uint32_t* key = (uint32_t*)malloc((size_t)n * 4);
    uint32_t* idx = (uint32_t*)malloc((size_t)n * 4);
    #pragma omp parallel for schedule(static)
    for (uint32_t i = 0; i < n; i++) {   /* :473-482 */
        float cell[3];
        PositionToCellCoord(o, &o->pred[3 * i], cell);
        hash[i] = HashCell(cell);
        key[i] = GetKeyFromHash(hash[i], n);
        o->start[i] = (uint32_t)INT_MAX;
    }
    if (forced_order) {
        memcpy(idx, forced_order, (size_t)n * 4);
    } else {               /* :484 std::sort by key; ties canonicalised by ascending index */
        uint32_t* k = (uint32_t*)malloc((size_t)n * 4);
        memcpy(k, key, (size_t)n * 4);
        for (uint32_t i = 0; i < n; i++) idx[i] = i;
        radix_sort_pairs(n, k, idx);
        free(k);
    }
    for (uint32_t s = 0; s < n; s++) {
        uint32_t i = idx[s];
        o->lk_f[3 * s] = (float)i; o->lk_f[3 * s + 1] = (float)hash[i]; o->lk_f[3 * s + 2] = (float)key[i];
        o->lk_u[3 * s] = i;        o->lk_u[3 * s + 1] = hash[i];        o->lk_u[3 * s + 2] = key[i];
    }
    for (uint32_t s = 0; s < n; s++) {   /* :486-496 (key read back from the stored row, float->u32) */
        uint32_t k = o->wide ? o->lk_u[3 * s + 2] : (uint32_t)o->lk_f[3 * s + 2];
        uint32_t kp = s == 0 ? UINT32_MAX : (o->wide ? o->lk_u[3 * (s - 1) + 2] : (uint32_t)o->lk_f[3 * (s - 1) + 2]);
        if (k != kp) o->start[k] = s;
    }
    free(hash); free(key); free(idx);
    o->t[1] = now_ms() - t0;
}

/* ---- S3: CalculateDensity :325-365 ---- */
static void density_of(const Oracle* o, const float* pos, float* out2, uint32_t* count)
{
    const uint32_t n = o->n;
    const float r = o->p.interaction_radius, sqrRadius = o->p.sqr_radius;
    float originCell[3];
    PositionToCellCoord(o, pos, originCell);
    float density = 0, NearDensity = 0;
    uint32_t cnt = 0;
    for (int i = 0; i < 27; i++) {
        float off[3], c[3];
        offset27(i, off);
        c[0] = originCell[0] + off[0]; c[1] = originCell[1] + off[1]; c[2] = originCell[2] + off[2];
        uint32_t hash = HashCell(c);
        uint32_t key = GetKeyFromHash(hash, n);
        uint32_t currIndex = o->start[key];
        while (currIndex < n) {
            uint32_t row = currIndex;
            currIndex++;
            if (row_key_ne(o, row, key)) break;
            if (row_hash_ne(o, row, hash)) continue;
            if (row_index_ge(o, row, n)) break;
            uint32_t j = row_index(o, row);
            const float* pj = &o->pred[3 * j];
            float ox = pj[0] - pos[0], oy = pj[1] - pos[1], oz = pj[2] - pos[2];
            float sqrDist = ox * ox + oy * oy + oz * oz;   /* glm::dot: (x*x + y*y) + z*z */
            if (sqrDist > sqrRadius) continue;
            float dist = sqrtf(sqrDist);
            density += SmoothingPow2(dist, r);
            NearDensity += SmoothingPow3(dist, r);
            cnt++;
        }
    }
    if (out2) { out2[0] = density; out2[1] = NearDensity; }
    if (count) *count = cnt;
}

void oracle_stage_density(Oracle* o)
{   /* updateDensities :304-311 */
    double t0 = now_ms();
    const uint32_t n = o->n;
    #pragma omp parallel for schedule(dynamic, 256)
    for (uint32_t i = 0; i < n; i++) density_of(o, &o->pred[3 * i], &o->dens[2 * i], NULL);
    o->t[2] = now_ms() - t0;
}

void oracle_get_neighbour_counts(const Oracle* o, uint32_t* out)
{
    #pragma omp parallel for schedule(dynamic, 256)
    for (uint32_t i = 0; i < o->n; i++) density_of(o, &o->pred[3 * i], NULL, &out[i]);
}

/* ---- S4: CalculatePressureForce :367-422 ---- */
static void pressure_of(const Oracle* o, uint32_t particleIndex, float deltatime, float* vout, float* scale)
{
    const uint32_t n = o->n;
    const float r = o->p.interaction_radius, sqrRadius = o->p.sqr_radius;
    const float density = o->dens[2 * particleIndex], nearDensity = o->dens[2 * particleIndex + 1];
    const float pressure = (density - o->p.target_density) * o->p.pressure_multiplier;
    const float nearPressure = nearDensity * o->p.near_pressure_multiplier;
    float F[3] = { 0, 0, 0 };
    float sc = 0;
    const float* pos = &o->pred[3 * particleIndex];
    float originCell[3];
    PositionToCellCoord(o, pos, originCell);
    for (int i = 0; i < 27; i++) {
        float off[3], c[3];
        offset27(i, off);
        c[0] = originCell[0] + off[0]; c[1] = originCell[1] + off[1]; c[2] = originCell[2] + off[2];
        uint32_t hash = HashCell(c);
        uint32_t key = GetKeyFromHash(hash, n);
        int64_t currIndex = (int32_t)o->start[key];   /* `int currIndex` compared with uint32 numParticles (:383,:386) */
        while ((uint32_t)currIndex < n) {
            uint32_t row = (uint32_t)currIndex;
            currIndex++;
            if (row_key_ne(o, row, key)) break;
            if (row_hash_ne(o, row, hash)) continue;
            uint32_t j = row_index(o, row);
            if (j == particleIndex) continue;
            const float* pj = &o->pred[3 * j];
            float o3[3] = { pj[0] - pos[0], pj[1] - pos[1], pj[2] - pos[2] };
            float sqrDist = o3[0] * o3[0] + o3[1] * o3[1] + o3[2] * o3[2];
            if (sqrDist > sqrRadius) continue;
            float neighborDensity = o->dens[2 * j], neighborNearDensity = o->dens[2 * j + 1];
            float neighborPressure = (neighborDensity - o->p.target_density) * o->p.pressure_multiplier;
            float neighborNearPressure = neighborNearDensity * o->p.near_pressure_multiplier;
            float sharedPressure = (pressure + neighborPressure) * 0.5f;
            float sharedNearPressure = (nearPressure + neighborNearPressure) * 0.5f;
            float dist = sqrtf(sqrDist);
            float dir[3];
            if (dist > 0) { dir[0] = o3[0] / dist; dir[1] = o3[1] / dist; dir[2] = o3[2] / dist; }
            else { dir[0] = 0; dir[1] = 1; dir[2] = 0; }
            float d2 = SmoothingDerivativePow2(dist, r), d3 = SmoothingDerivativePow3(dist, r);
            for (int a = 0; a < 3; a++) {
                float t1 = dir[a] * d2 * sharedPressure / neighborDensity;
                F[a] += t1;
                float t2 = dir[a] * d3 * sharedNearPressure / neighborNearDensity;
                F[a] += t2;
            }
            /* tolerance scale: same terms with cancellation-free magnitudes (|rho| + rho0 instead of rho - rho0) */
            {
                float magP = ((fabsf(density) + fabsf(o->p.target_density)) + (fabsf(neighborDensity) + fabsf(o->p.target_density)))
                             * fabsf(o->p.pressure_multiplier) * 0.5f;
                sc += fabsf(d2 * magP / neighborDensity) + fabsf(d3 * sharedNearPressure / neighborNearDensity);
            }
        }
    }
    if (vout) for (int a = 0; a < 3; a++) vout[a] = o->vel[3 * particleIndex + a] + (F[a] / density) * deltatime;
    if (scale) *scale = sc / density * deltatime;
}

void oracle_stage_pressure(Oracle* o, float dt)
{
    double t0 = now_ms();
    const uint32_t n = o->n;
    /* each particle writes only its own velocity and reads no other velocity: order-free */
    #pragma omp parallel for schedule(dynamic, 256)
    for (uint32_t i = 0; i < n; i++) pressure_of(o, i, dt, &o->vel_press[3 * i], NULL);
    memcpy(o->vel, o->vel_press, (size_t)n * 12);
    o->t[3] = now_ms() - t0;
}

/* ---- S5: CalculateViscosityForce :424-464 ---- */
static void viscosity_of(const Oracle* o, const float* velsrc, uint32_t particleIndex, float deltatime, float* vout, float* scale)
{
    const uint32_t n = o->n;
    const float r = o->p.interaction_radius, sqrRadius = o->p.sqr_radius;
    const float* pos = &o->pred[3 * particleIndex];
    float originCell[3];
    PositionToCellCoord(o, pos, originCell);
    float F[3] = { 0, 0, 0 };
    float sc = 0;
    const float* velo = &velsrc[3 * particleIndex];
    for (int i = 0; i < 27; i++) {
        float off[3], c[3];
        offset27(i, off);
        c[0] = originCell[0] + off[0]; c[1] = originCell[1] + off[1]; c[2] = originCell[2] + off[2];
        uint32_t hash = HashCell(c);
        uint32_t key = GetKeyFromHash(hash, n);
        int64_t currIndex = (int32_t)o->start[key];
        while ((uint32_t)currIndex < n) {
            uint32_t row = (uint32_t)currIndex;
            currIndex++;
            if (row_key_ne(o, row, key)) break;
            if (row_hash_ne(o, row, hash)) continue;
            uint32_t j = row_index(o, row);
            if (j == particleIndex) continue;
            const float* pj = &o->pred[3 * j];
            float ox = pj[0] - pos[0], oy = pj[1] - pos[1], oz = pj[2] - pos[2];
            float sqrDist = ox * ox + oy * oy + oz * oz;
            if (sqrDist > sqrRadius) continue;
            float dist = sqrtf(sqrDist);
            float influence = SmoothingViscoPoly6(dist, r);
            float dv[3] = { velsrc[3 * j] - velo[0], velsrc[3 * j + 1] - velo[1], velsrc[3 * j + 2] - velo[2] };
            for (int a = 0; a < 3; a++) F[a] += dv[a] * influence;
            sc += sqrtf(dv[0] * dv[0] + dv[1] * dv[1] + dv[2] * dv[2]) * influence;
        }
    }
    if (vout) for (int a = 0; a < 3; a++) vout[a] = velo[a] + F[a] * o->p.viscosity_strength * deltatime;
    if (scale) *scale = sc * o->p.viscosity_strength * deltatime;
}

void oracle_stage_viscosity(Oracle* o, float dt, int jacobi)
{
    double t0 = now_ms();
    const uint32_t n = o->n;
    if (jacobi) {   /* snapshot semantics: what the GPU implements (SURVEY App.A Q11) */
        #pragma omp parallel for schedule(dynamic, 256)
        for (uint32_t i = 0; i < n; i++) viscosity_of(o, o->vel, i, dt, &o->vel_visc[3 * i], NULL);
        memcpy(o->vel, o->vel_visc, (size_t)n * 12);
    } else {        /* in place, index order: what the serial-PSTL reference does */
        for (uint32_t i = 0; i < n; i++) {
            float v[3];
            viscosity_of(o, o->vel, i, dt, v, NULL);
            o->vel[3 * i] = v[0]; o->vel[3 * i + 1] = v[1]; o->vel[3 * i + 2] = v[2];
        }
        memcpy(o->vel_visc, o->vel, (size_t)n * 12);
    }
    o->t[4] = now_ms() - t0;
}

/* ---- S6: :81-108 ---- */
static inline float glm_sign(float x) { return (float)((0.0f < x) - (x < 0.0f)); } /* glm func_common.inl:144-150 */

void oracle_stage_integrate(Oracle* o, float dt)
{
    double t0 = now_ms();
    const uint32_t n = o->n;
    const float dampFactor = 0.95f;
    #pragma omp parallel for schedule(static)
    for (uint32_t i = 0; i < n; i++) {
        for (int a = 0; a < 3; a++) o->pos[3 * i + a] += o->vel[3 * i + a] * dt;
        for (int a = 0; a < 3; a++) {
            const float halfSize = o->p.bound[a] * 0.5f;
            float edgeDst = halfSize - fabsf(o->pos[3 * i + a]);
            if (edgeDst <= 0) {
                o->pos[3 * i + a] = halfSize * glm_sign(o->pos[3 * i + a]);
                o->vel[3 * i + a] *= -1 * dampFactor;
            }
        }
        o->out4[4 * i] = o->pos[3 * i]; o->out4[4 * i + 1] = o->pos[3 * i + 1];
        o->out4[4 * i + 2] = o->pos[3 * i + 2]; o->out4[4 * i + 3] = 0.34f;
    }
    o->t[5] = now_ms() - t0;
}

void oracle_step(Oracle* o, float dt, int jacobi)
{   /* Update :39-111 */
    oracle_stage_predict(o, dt);
    oracle_stage_spatial(o, NULL);
    oracle_stage_density(o);
    oracle_stage_pressure(o, dt);
    oracle_stage_viscosity(o, dt, jacobi);
    oracle_stage_integrate(o, dt);
}

/* ---- InitializeData :112-147 + GridArrangement :518-557 ---- */
void oracle_spawn_grid(Oracle* o)
{
    const uint32_t n = o->n;
    memset(o->pos, 0, (size_t)n * 12); memset(o->vel, 0, (size_t)n * 12);
    memset(o->pred, 0, (size_t)n * 12); memset(o->dens, 0, (size_t)n * 8);
    for (uint32_t i = 0; i < n; i++) { o->out4[4 * i] = o->out4[4 * i + 1] = o->out4[4 * i + 2] = 0; o->out4[4 * i + 3] = 0.25f; }
    int particlesPerAxis = (int)ceil(powf((float)(int)n, (1.0f / 3.0f)));   /* :139 */
    float gap = 0.215f;
    uint32_t i = 0;
    float Total = particlesPerAxis * gap;
    for (int localY = 0; localY < particlesPerAxis && i < n; localY++)
        for (int localX = 0; localX < particlesPerAxis && i < n; localX++)
            for (int localZ = 0; localZ < particlesPerAxis && i < n; localZ++) {
                float XOffset = localX * gap, YOffset = localY * gap, ZOffset = localZ * gap;
                float worldOffsetX = (0 - ((Total - gap) / 2.0f));
                float worldOffsetY = (0 + (Total - gap) / 2.0f);
                float worldOffsetZ = (0 - (Total - gap) / 2.0f);
                float x = worldOffsetX + XOffset, y = worldOffsetY - YOffset, z = worldOffsetZ + ZOffset;
                o->pos[3 * i] = x; o->pos[3 * i + 1] = y; o->pos[3 * i + 2] = z;
                o->pred[3 * i] = x; o->pred[3 * i + 1] = y; o->pred[3 * i + 2] = z;
                o->out4[4 * i] = x; o->out4[4 * i + 1] = y; o->out4[4 * i + 2] = z; o->out4[4 * i + 3] = 0.34f;
                i++;
            }
    oracle_stage_spatial(o, NULL);   /* :144 */
    oracle_stage_density(o);         /* :145 */
}

/* ---- read-back ---- */
void oracle_get_positions(const Oracle* o, float* out3)     { memcpy(out3, o->pos, (size_t)o->n * 12); }
void oracle_get_out_positions(const Oracle* o, float* out4) { memcpy(out4, o->out4, (size_t)o->n * 16); }
void oracle_get_velocities(const Oracle* o, float* out3)    { memcpy(out3, o->vel, (size_t)o->n * 12); }
void oracle_get_predicted(const Oracle* o, float* out3)     { memcpy(out3, o->pred, (size_t)o->n * 12); }
void oracle_get_densities(const Oracle* o, float* out2)     { memcpy(out2, o->dens, (size_t)o->n * 8); }
void oracle_get_vel_after_pressure(const Oracle* o, float* out3)  { memcpy(out3, o->vel_press, (size_t)o->n * 12); }
void oracle_get_vel_after_viscosity(const Oracle* o, float* out3) { memcpy(out3, o->vel_visc, (size_t)o->n * 12); }
void oracle_get_start_indices(const Oracle* o, uint32_t* out)     { memcpy(out, o->start, (size_t)o->n * 4); }
void oracle_get_timings(const Oracle* o, double* out6)            { memcpy(out6, o->t, sizeof o->t); }

void oracle_get_hash_key(const Oracle* o, uint32_t* hash, uint32_t* key, int32_t* cell3)
{
    for (uint32_t i = 0; i < o->n; i++) {
        float cell[3];
        PositionToCellCoord(o, &o->pred[3 * i], cell);
        uint32_t h = HashCell(cell);
        if (hash) hash[i] = h;
        if (key) key[i] = GetKeyFromHash(h, o->n);
        if (cell3) { cell3[3 * i] = (int32_t)cell[0]; cell3[3 * i + 1] = (int32_t)cell[1]; cell3[3 * i + 2] = (int32_t)cell[2]; }
    }
}

void oracle_get_sorted(const Oracle* o, uint32_t* idx, uint32_t* hash_as_stored, uint32_t* key)
{   /* rows as the reference would hold them: u32 values after the float round trip unless wide */
    for (uint32_t s = 0; s < o->n; s++) {
        if (idx) idx[s] = o->wide ? o->lk_u[3 * s] : (uint32_t)o->lk_f[3 * s];
        if (hash_as_stored) hash_as_stored[s] = o->wide ? o->lk_u[3 * s + 1] : (uint32_t)(int64_t)o->lk_f[3 * s + 1];
        if (key) key[s] = o->wide ? o->lk_u[3 * s + 2] : (uint32_t)o->lk_f[3 * s + 2];
    }
}

void oracle_get_force_scales(const Oracle* o, float dt, float* pressure_scale, float* viscosity_scale)
{   /* needs the tables/pred/dens of the step just taken; viscosity scale uses the post-pressure snapshot */
    #pragma omp parallel for schedule(dynamic, 256)
    for (uint32_t i = 0; i < o->n; i++) {
        if (pressure_scale) pressure_of(o, i, dt, NULL, &pressure_scale[i]);
        if (viscosity_scale) viscosity_of(o, o->vel_press, i, dt, NULL, &viscosity_scale[i]);
    }
}
