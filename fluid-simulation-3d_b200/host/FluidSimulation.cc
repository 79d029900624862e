// host/FluidSimulation.cc -- see FluidSimulation.h.  Pure host C++: talks to the GPU only through
// the C ABI of libsph_b200.so.
#include "FluidSimulation.h"

#include <stdexcept>

namespace Physics
{
	namespace Fluid
	{
		FluidSimulation& FluidSimulation::getInstance()
		{
			static FluidSimulation instance;
			return instance;
		}

		FluidSimulation::~FluidSimulation()
		{
			if (pinnedOut) sph_host_unregister(pinnedOut);
			if (pinnedPos) sph_host_unregister(pinnedPos);
			if (ctx) sph_destroy(ctx);
		}

		// The reference cannot fail (void returns, physicsWorld.cc); a GPU backend can, and a silent
		// failure would leave stale particles on screen, so errors throw.
		void FluidSimulation::check(int rc, const char* what)
		{
			if (rc == SPH_OK) return;
			const char* msg = sph_last_error(ctx);
			error = std::string(what) + ": " + (msg ? msg : "unknown error");
			throw std::runtime_error(error);
		}

		void FluidSimulation::ensureContext(uint32_t cap)
		{
			if (ctx && cap <= capacity) return;
			if (ctx) { sph_destroy(ctx); ctx = nullptr; }
			int rc = sph_create(&ctx, device, cap);
			if (rc != SPH_OK) {
				const char* msg = sph_last_error(nullptr);
				error = std::string("sph_create: ") + (msg ? msg : "unknown error");
				ctx = nullptr;
				throw std::runtime_error(error);
			}
			capacity = cap;
			check(sph_set_table_mode(ctx, tableMode), "sph_set_table_mode");
			pushParams();
		}

		void FluidSimulation::pushParams()
		{
			if (ctx) check(sph_set_params(ctx, &params), "sph_set_params");
		}

		void FluidSimulation::setDevice(int cudaDevice) { device = cudaDevice; }
		void FluidSimulation::setTableMode(int m) { tableMode = m; if (ctx) check(sph_set_table_mode(ctx, m), "sph_set_table_mode"); }
		void FluidSimulation::setHostMirrors(bool outPositions, bool positionsToo) { mirrorOut = outPositions; mirrorPos = positionsToo; }

		void FluidSimulation::InitializeData(int particleAmmount, vec3)
		{
			// (the reference ignores Centre too: GridArrangement is called without it, :142)
			if (particleAmmount < 0) particleAmmount = 0;
			numParticles = (uint32)particleAmmount;
			ensureContext(numParticles ? numParticles : 1);
			if (pinnedOut) { sph_host_unregister(pinnedOut); pinnedOut = nullptr; }
			if (pinnedPos) { sph_host_unregister(pinnedPos); pinnedPos = nullptr; }
			positions.assign(numParticles, vec3(0, 0, 0));
			OutPositions.assign(numParticles, vec4(0, 0, 0, 0.25f));
			if (numParticles) {
				// pin the mirrors in place so the per-frame copies run at full PCIe rate without
				// changing the containers' types
				if (sph_host_register(OutPositions.data(), OutPositions.size() * sizeof(vec4)) == SPH_OK) {
					pinnedOut = OutPositions.data(); pinnedOutBytes = OutPositions.size() * sizeof(vec4);
				}
				if (sph_host_register(positions.data(), positions.size() * sizeof(vec3)) == SPH_OK) {
					pinnedPos = positions.data(); pinnedPosBytes = positions.size() * sizeof(vec3);
				}
			}
			check(sph_spawn_grid(ctx, numParticles), "sph_spawn_grid");     // lattice + lookup + densities (:139-145)
			if (numParticles) {
				check(sph_download(ctx, SPH_FIELD_POSITIONS, positions.data(), positions.size() * sizeof(vec3)), "sph_download");
				check(sph_download(ctx, SPH_FIELD_OUT_POSITIONS, OutPositions.data(), OutPositions.size() * sizeof(vec4)), "sph_download");
			}
			cacheValid = false;
			timingsFresh = true;
		}

		void FluidSimulation::Update(float deltatime)
		{
			if (!ctx || numParticles == 0) return;
			check(sph_step(ctx, deltatime), "sph_step");
			// Update() returns with the host-visible buffers complete: the renderer takes
			// &OutPositions[0] right after (fluidSimCPU.cc:58)
			if (mirrorOut)
				check(sph_download(ctx, SPH_FIELD_OUT_POSITIONS, OutPositions.data(), OutPositions.size() * sizeof(vec4)), "sph_download");
			if (mirrorPos)
				check(sph_download(ctx, SPH_FIELD_POSITIONS, positions.data(), positions.size() * sizeof(vec3)), "sph_download");
			if (!mirrorOut && !mirrorPos) check(sph_synchronize(ctx), "sph_synchronize");
			cacheValid = false;
			timingsFresh = false;
		}

		void FluidSimulation::uploadState(const float* velocities3)
		{
			if (!ctx) return;
			std::vector<float> vel;
			if (!velocities3 && numParticles) {
				vel.resize((size_t)numParticles * 3);
				check(sph_download(ctx, SPH_FIELD_VELOCITIES, vel.data(), vel.size() * 4), "sph_download");
				velocities3 = vel.data();
			}
			check(sph_upload_state(ctx, numParticles, numParticles ? &positions[0].x : nullptr, velocities3), "sph_upload_state");
			check(sph_synchronize(ctx), "sph_synchronize");
			cacheValid = false;
		}

		void FluidSimulation::downloadVelocities(std::vector<vec3>& out)
		{
			out.resize(numParticles);
			if (ctx && numParticles) check(sph_download(ctx, SPH_FIELD_VELOCITIES, out.data(), out.size() * sizeof(vec3)), "sph_download");
		}
		void FluidSimulation::downloadDensities(std::vector<float>& out)
		{
			out.resize((size_t)numParticles * 2);
			if (ctx && numParticles) check(sph_download(ctx, SPH_FIELD_DENSITIES, out.data(), out.size() * 4), "sph_download");
		}
		void FluidSimulation::downloadColors(std::vector<vec4>& out)
		{
			out.resize(numParticles);
			if (ctx && numParticles) check(sph_download(ctx, SPH_FIELD_COLORS, out.data(), out.size() * sizeof(vec4)), "sph_download");
		}

		// per-particle getters: bounds-checked, zero when out of range (:151,157,163,168,174,180)
		void FluidSimulation::readParticle(uint32 index)
		{
			if (cacheValid && cachedIndex == index) return;
			for (float& f : cached) f = 0.0f;
			if (ctx && index < numParticles) check(sph_get_particle(ctx, index, cached), "sph_get_particle");
			cachedIndex = index;
			cacheValid = true;
		}
		FluidSimulation::vec3 FluidSimulation::getPosition(uint32 i) { readParticle(i); return vec3(cached[0], cached[1], cached[2]); }
		FluidSimulation::vec3 FluidSimulation::getVelocity(uint32 i) { readParticle(i); return vec3(cached[3], cached[4], cached[5]); }
		float FluidSimulation::getDensity(uint32 i) { readParticle(i); return cached[6]; }
		float FluidSimulation::getNearDensity(uint32 i) { readParticle(i); return cached[7]; }
		float FluidSimulation::getSpeed(uint32 i) { readParticle(i); return cached[8]; }
		float FluidSimulation::getSpeedNormalzied(uint32 i) { readParticle(i); return cached[9]; }

		void FluidSimulation::refreshTimings()
		{
			if (timingsFresh || !ctx) return;
			check(sph_get_timings(ctx, timings), "sph_get_timings");
			timingsFresh = true;
		}
		double FluidSimulation::getElapsedTimeGravity() { refreshTimings(); return timings[0]; }
		double FluidSimulation::getElapsedTimeSpatial() { refreshTimings(); return timings[1]; }
		double FluidSimulation::getElapsedTimeDensity() { refreshTimings(); return timings[2]; }
		double FluidSimulation::getElapsedTimePressure() { refreshTimings(); return timings[3]; }
		double FluidSimulation::getElapsedTimeViscosity() { refreshTimings(); return timings[4]; }
		double FluidSimulation::getElapsedTimePosNColl() { refreshTimings(); return timings[5]; }

		// setters only store a scalar; they take effect on the next Update (:214-302)
		void FluidSimulation::setSimulationTime(float time) { simTime = time; }
		float FluidSimulation::getSimulationTime() { return simTime; }
		void FluidSimulation::setGravity(bool status) { params.gravity = status ? 1 : 0; pushParams(); }
		bool FluidSimulation::getGravityStatus() { return params.gravity != 0; }
		void FluidSimulation::setInteractionRadius(float value) { params.interaction_radius = value; pushParams(); }   // sqr_radius untouched (Q2)
		float FluidSimulation::getInteractionRadius() { return params.interaction_radius; }
		void FluidSimulation::setDensityTarget(float value) { params.target_density = value; pushParams(); }
		float FluidSimulation::getDensityTarget() { return params.target_density; }
		void FluidSimulation::setPressureMultiplier(float value) { params.pressure_multiplier = value; pushParams(); }
		float FluidSimulation::getPressureMultiplier() { return params.pressure_multiplier; }
		void FluidSimulation::setNearPressureMultiplier(float value) { params.near_pressure_multiplier = value; pushParams(); }
		float FluidSimulation::getNearPressureMultiplier() { return params.near_pressure_multiplier; }
		void FluidSimulation::setViscosityStrength(float value) { params.viscosity_strength = value; pushParams(); }
		float FluidSimulation::getViscosityStrength() { return params.viscosity_strength; }
		void FluidSimulation::setGravityScale(float value) { params.gravity_scale = value; pushParams(); }
		float FluidSimulation::getGravityScale() { return params.gravity_scale; }
		void FluidSimulation::setBound(const vec3& value) { params.bound[0] = value.x; params.bound[1] = value.y; params.bound[2] = value.z; pushParams(); }
		FluidSimulation::vec3 FluidSimulation::getBounds() { return vec3(params.bound[0], params.bound[1], params.bound[2]); }
	}
}
