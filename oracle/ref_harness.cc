// oracle/ref_harness.cc -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A thin C-ABI harness around the *unmodified* reference translation unit
// /root/reference/engine/physics/physicsWorld.cc (compiled where it lies by
// oracle/Makefile; no reference source is copied into this repository).  It is
// built with -fno-access-control so it can reach the private stage functions
// and arrays of Physics::Fluid::FluidSimulation (physicsWorld.h:82-153).
//
// It exists to (1) pin the plain-C restatement oracle/sph_oracle.c, (2) generate
// the golden fixtures under tests/golden/, (3) serve as the "reference" CPU
// baseline in bench.py.  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load the resulting library.
//
// Stage map (all line numbers: engine/physics/physicsWorld.cc):
//   S1 predict            :42-48   (lambda inside Update; restated in ref_stage_predict,
//                                    proven bit-identical to Update() by tests/test_oracle.py)
//   S2 spatial lookup     :466-498 (UpdateSpatialLookup, called directly)
//   S3 densities          :304-311 (updateDensities, called directly)
//   S4 pressure           :367-422 (CalculatePressureForce, called per particle)
//   S5 viscosity          :424-464 (CalculateViscosityForce, called per particle; Jacobi
//                                    variant = call, record, restore -- SURVEY App.A Q11)
//   S6 integrate/collide  :81-108  (lambda inside Update; restated in ref_stage_integrate,
//                                    proven bit-identical to Update() by tests/test_oracle.py)
#include "config.h"
#include "physics/physicsWorld.h"

#include <cstdint>
#include <cstring>
#include <climits>
#include <vector>

using Physics::Fluid::FluidSimulation;

namespace {
FluidSimulation& S() { return FluidSimulation::getInstance(); }
std::vector<glm::vec3> g_vel_after_pressure;
std::vector<glm::vec3> g_vel_after_viscosity;
}

extern "C" {

struct RefParams {
    float interaction_radius;
    float target_density;
    float pressure_multiplier;
    float near_pressure_multiplier;
    float viscosity_strength;
    float gravity_scale;
    int   gravity;
    float bound[3];
};

// InitializeData(n): the reference's own lattice spawn (physicsWorld.cc:112-147).
void ref_initialize_data(int n)
{
    FluidSimulation& s = S();
    // Q16: the lookup table only ever grows; shrink it by hand so a smaller
    // re-initialisation on the process-wide singleton does not sort stale rows.
    s.spatialLookup.clear();
    s.startIndices.clear();
    s.InitializeData(n);
}

// Allocate for n particles WITHOUT running the spawn (the spawn costs a full
// spatial + density pass and places >800k particles outside the default box).
void ref_allocate(int n)
{
    FluidSimulation& s = S();
    s.numParticles = (uint32)n;
    s.pList.resize(n);
    for (int i = 0; i < n; i++) s.pList[i] = i;
    s.positions.assign(n, glm::vec3(0));
    s.OutPositions.assign(n, glm::vec4(0, 0, 0, 0.25f));
    s.velocity.assign(n, glm::vec3(0));
    s.velocity2.assign(n, glm::vec3(0));
    s.predictedPositions.assign(n, glm::vec3(0));
    s.densities.assign(n, glm::vec2(0));
    s.spatialLookup.assign(n, glm::vec3(0));
    s.startIndices.assign(n, (uint32_t)INT_MAX);
}

void ref_set_params(const RefParams* p)
{
    FluidSimulation& s = S();
    s.setInteractionRadius(p->interaction_radius);
    s.setDensityTarget(p->target_density);
    s.setPressureMultiplier(p->pressure_multiplier);
    s.setNearPressureMultiplier(p->near_pressure_multiplier);
    s.setViscosityStrength(p->viscosity_strength);
    s.setGravityScale(p->gravity_scale);
    s.setGravity(p->gravity != 0);
    s.setBound(glm::vec3(p->bound[0], p->bound[1], p->bound[2]));
}

void ref_get_params(RefParams* p)
{
    FluidSimulation& s = S();
    p->interaction_radius = s.getInteractionRadius();
    p->target_density = s.getDensityTarget();
    p->pressure_multiplier = s.getPressureMultiplier();
    p->near_pressure_multiplier = s.getNearPressureMultiplier();
    p->viscosity_strength = s.getViscosityStrength();
    p->gravity_scale = s.getGravityScale();
    p->gravity = s.getGravityStatus() ? 1 : 0;
    glm::vec3 b = s.getBounds();
    p->bound[0] = b.x; p->bound[1] = b.y; p->bound[2] = b.z;
}

float ref_get_sqr_radius() { return S().sqrRadius; }
int ref_num_particles() { return (int)S().numParticles; }

void ref_set_state(const float* pos3, const float* vel3)
{
    FluidSimulation& s = S();
    const size_t n = s.numParticles;
    if (pos3) std::memcpy(s.positions.data(), pos3, n * sizeof(glm::vec3));
    if (vel3) std::memcpy(s.velocity.data(), vel3, n * sizeof(glm::vec3));
}

// ---- verbatim whole step ---------------------------------------------------
void ref_update(float dt) { S().Update(dt); }

// ---- staged step -----------------------------------------------------------
// S1, restating the lambda at physicsWorld.cc:42-48 with the same glm expressions.
void ref_stage_predict(float dt)
{
    FluidSimulation& s = S();
    for (uint32_t i = 0; i < s.numParticles; i++) {
        s.velocity[i] += s.CalculateExternalFoce(s.positions[i], s.velocity[i]) * dt;
        s.predictedPositions[i] = s.positions[i] + s.velocity[i] * (1.0f / 120.0f);
    }
}
void ref_stage_spatial() { S().UpdateSpatialLookup(); }
void ref_stage_density() { S().updateDensities(); }
void ref_stage_pressure(float dt)
{
    FluidSimulation& s = S();
    for (uint32_t i = 0; i < s.numParticles; i++) s.CalculatePressureForce(i, dt);
    g_vel_after_pressure = s.velocity;
}
// jacobi != 0: every particle sees the post-pressure velocity snapshot (the
// semantics the GPU implements); jacobi == 0: in-place, index order = what the
// serial-PSTL reference does.
void ref_stage_viscosity(float dt, int jacobi)
{
    FluidSimulation& s = S();
    const uint32_t n = s.numParticles;
    if (jacobi) {
        std::vector<glm::vec3> out(n);
        for (uint32_t i = 0; i < n; i++) {
            const glm::vec3 saved = s.velocity[i];
            s.CalculateViscosityForce(i, dt);
            out[i] = s.velocity[i];
            s.velocity[i] = saved;
        }
        s.velocity = out;
    } else {
        for (uint32_t i = 0; i < n; i++) s.CalculateViscosityForce(i, dt);
    }
    g_vel_after_viscosity = s.velocity;
}
// S6, restating the lambda at physicsWorld.cc:81-108 with the same glm expressions.
void ref_stage_integrate(float dt)
{
    FluidSimulation& s = S();
    for (uint32_t i = 0; i < s.numParticles; i++) {
        s.positions[i] += s.velocity[i] * dt;
        const float dampFactor = 0.95f;
        const glm::vec3 halfSize = s.BoundScale * 0.5f;
        glm::vec3 edgeDst = halfSize - abs(s.positions[i]);
        if (edgeDst.x <= 0) {
            s.positions[i].x = halfSize.x * glm::sign(s.positions[i].x);
            s.velocity[i].x *= -1 * dampFactor;
        }
        if (edgeDst.y <= 0) {
            s.positions[i].y = halfSize.y * glm::sign(s.positions[i].y);
            s.velocity[i].y *= -1 * dampFactor;
        }
        if (edgeDst.z <= 0) {
            s.positions[i].z = halfSize.z * glm::sign(s.positions[i].z);
            s.velocity[i].z *= -1 * dampFactor;
        }
        s.OutPositions[i] = glm::vec4(s.positions[i], 0.34f);
    }
}
void ref_step_staged(float dt, int jacobi)
{
    ref_stage_predict(dt);
    ref_stage_spatial();
    ref_stage_density();
    ref_stage_pressure(dt);
    ref_stage_viscosity(dt, jacobi);
    ref_stage_integrate(dt);
}

// ---- read-back -------------------------------------------------------------
void ref_get_positions(float* out3)  { std::memcpy(out3, S().positions.data(), S().numParticles * 12); }
void ref_get_out_positions(float* out4) { std::memcpy(out4, S().OutPositions.data(), S().numParticles * 16); }
void ref_get_velocities(float* out3) { std::memcpy(out3, S().velocity.data(), S().numParticles * 12); }
void ref_get_predicted(float* out3)  { std::memcpy(out3, S().predictedPositions.data(), S().numParticles * 12); }
void ref_get_densities(float* out2)  { std::memcpy(out2, S().densities.data(), S().numParticles * 8); }
void ref_get_vel_after_pressure(float* out3)  { std::memcpy(out3, g_vel_after_pressure.data(), g_vel_after_pressure.size() * 12); }
void ref_get_vel_after_viscosity(float* out3) { std::memcpy(out3, g_vel_after_viscosity.data(), g_vel_after_viscosity.size() * 12); }
// raw float rows (index, hash, key) exactly as the reference stores them
void ref_get_lookup_raw(float* out3) { std::memcpy(out3, S().spatialLookup.data(), S().numParticles * 12); }
void ref_get_start_indices(uint32_t* out) { std::memcpy(out, S().startIndices.data(), S().numParticles * 4); }

// Per-particle exact hash / key of the current predicted positions, through the
// reference's own PositionToCellCoord / HashCell / GetKeyFromHash (:499-516).
void ref_get_hash_key(uint32_t* hash, uint32_t* key, int32_t* cell3)
{
    FluidSimulation& s = S();
    for (uint32_t i = 0; i < s.numParticles; i++) {
        glm::vec3 c = s.PositionToCellCoord(s.predictedPositions[i]);
        uint32_t h = s.HashCell(c);
        if (hash) hash[i] = h;
        if (key) key[i] = s.GetKeyFromHash(h, s.numParticles);
        if (cell3) { cell3[3*i] = (int32_t)c.x; cell3[3*i+1] = (int32_t)c.y; cell3[3*i+2] = (int32_t)c.z; }
    }
}

// Neighbour count, density-pass definition (incl. self): the walk of
// CalculateDensity (:325-365) restated with the reference's own helpers and
// tables, counting the candidates that survive all its filters.
void ref_get_neighbour_counts(uint32_t* out)
{
    FluidSimulation& s = S();
    const uint32_t n = s.numParticles;
    for (uint32_t p = 0; p < n; p++) {
        const glm::vec3 pos = s.predictedPositions[p];
        const glm::vec3 originCell = s.PositionToCellCoord(pos);
        uint32_t cnt = 0;
        for (int i = 0; i < 27; i++) {
            uint32_t hash = s.HashCell(originCell + s.offsets[i]);
            uint32_t key = s.GetKeyFromHash(hash, n);
            uint32 currIndex = s.startIndices[key];
            while (currIndex < n) {
                const glm::vec3& index = s.spatialLookup[currIndex];
                currIndex++;
                if (index.z != key) break;
                if (index.y != hash) continue;
                if (index.x >= n) break;
                uint32_t neighborIndex = index.x;
                const glm::vec3 off = s.predictedPositions[neighborIndex] - pos;
                float sqrDist = dot(off, off);
                if (sqrDist > s.sqrRadius) continue;
                cnt++;
            }
        }
        out[p] = cnt;
    }
}

void ref_get_timings(double* out6)
{
    FluidSimulation& s = S();
    out6[0] = s.getElapsedTimeGravity();  out6[1] = s.getElapsedTimeSpatial();
    out6[2] = s.getElapsedTimeDensity();  out6[3] = s.getElapsedTimePressure();
    out6[4] = s.getElapsedTimeViscosity(); out6[5] = s.getElapsedTimePosNColl();
}

// public getters, for the bounds-check behaviour (physicsWorld.cc:149-182)
void ref_getter_probe(uint32_t i, float* out10)
{
    FluidSimulation& s = S();
    glm::vec3 p = s.getPosition(i), v = s.getVelocity(i);
    out10[0] = p.x; out10[1] = p.y; out10[2] = p.z;
    out10[3] = v.x; out10[4] = v.y; out10[5] = v.z;
    out10[6] = s.getDensity(i); out10[7] = s.getNearDensity(i);
    out10[8] = s.getSpeed(i);   out10[9] = s.getSpeedNormalzied(i);
}

// the five smoothing kernels (kernels.h:25-82) for known-answer tests.
// They are included by physicsWorld.cc only, so re-include here.
}  // extern "C"
#include "physics/kernels.h"
extern "C" {
void ref_kernels(float dist, float radius, float* out5)
{
    out5[0] = Physics::kernels::SmoothingPow2(dist, radius);
    out5[1] = Physics::kernels::SmoothingPow3(dist, radius);
    out5[2] = Physics::kernels::SmoothingDerivativePow2(dist, radius);
    out5[3] = Physics::kernels::SmoothingDerivativePow3(dist, radius);
    out5[4] = Physics::kernels::SmoothingViscoPoly6(dist, radius);
}
}
