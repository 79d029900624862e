// host/FluidSimulation.cc -- see FluidSimulation.h.  Pure host C++: talks to the GPU only through
// the C ABI of libsph_b200.so.
#ifdef SPH_B200_USE_GLM
#include "config.h"                      // inside the reference tree: its prelude first, as physicsWorld.cc:18 does (glm, uint32)
#endif
#include "FluidSimulation.h"
#include "SlabGroup.h"

#include <cmath>
#include <cstdio>
#include <cstring>
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

		FluidSimulation::FluidSimulation() {}
		FluidSimulation::~FluidSimulation() { shutdown(); }

		void FluidSimulation::shutdown()
		{
			if (pinnedOut) { sph_host_unregister(pinnedOut); pinnedOut = nullptr; }
			if (pinnedPos) { sph_host_unregister(pinnedPos); pinnedPos = nullptr; }
			group.reset();
			if (ctx) { sph_destroy(ctx); ctx = nullptr; }
			capacity = 0;
			numParticles = 0;
			invalidateGetters();
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

		// The reference's setters cannot fail (physicsWorld.cc:246-297) and are driven by UI sliders: a value the device
		// side rejects is NOT committed -- the host copy, the context and every slab keep the parameters in force -- and the
		// reason is kept in lastError() instead of an exception.
		bool FluidSimulation::commitParams(const SphParams& cand)
		{
			const SphParams old = params;
			if (ctx && sph_set_params(ctx, &cand) != SPH_OK) {       // the context restores its old state itself
				const char* msg = sph_last_error(ctx);
				error = std::string("sph_set_params: ") + (msg ? msg : "unknown error");
				return false;
			}
			if (group) {
				try { group->setParams(cand); }                       // all ranks or none
				catch (const std::exception& e) {
					error = e.what();
					if (ctx) sph_set_params(ctx, &old);
					return false;
				}
			}
			params = cand;
			return true;
		}

		bool FluidSimulation::commitExtras(const SphExtras& cand)
		{
			if (ctx && sph_set_extras(ctx, &cand) != SPH_OK) {
				const char* msg = sph_last_error(ctx);
				error = std::string("sph_set_extras: ") + (msg ? msg : "unknown error");
				return false;
			}
			if (group) {
				try { group->setExtras(cand); }
				catch (const std::exception& e) { error = e.what(); if (ctx) sph_set_extras(ctx, &extras); return false; }
			}
			extras = cand;
			return true;
		}
		bool FluidSimulation::setBoundRotation(float qx, float qy, float qz, float qw)
		{
			SphExtras e = extras;
			e.bound_rotation[0] = qx; e.bound_rotation[1] = qy; e.bound_rotation[2] = qz; e.bound_rotation[3] = qw;
			return commitExtras(e);
		}
		bool FluidSimulation::setStickiness(float strength, float distance)
		{
			SphExtras e = extras;
			e.stick_strength = strength; e.stick_distance = distance;
			return commitExtras(e);
		}

		void FluidSimulation::pushParams()
		{
			if (ctx) check(sph_set_params(ctx, &params), "sph_set_params");
			if (ctx) check(sph_set_extras(ctx, &extras), "sph_set_extras");
			if (group) {
				try { group->setParams(params); }
				catch (const std::exception& e) { error = e.what(); throw; }
			}
		}

		void FluidSimulation::setDevice(int cudaDevice) { device = cudaDevice; devices.clear(); }
		void FluidSimulation::setDevices(const std::vector<int>& cudaDevices)
		{
			devices = cudaDevices;
			if (!devices.empty()) device = devices[0];           // the spawn lattice is made on the first device
			if (devices.size() == 1) devices.clear();
			if (!devices.empty() && tableMode != SPH_TABLE_GRID) {
				error = "setDevices: the slab decomposition uses the GRID table (the reference table is global)";
				throw std::runtime_error(error);
			}
		}
		std::vector<uint32> FluidSimulation::particlesPerDevice() const
		{
			if (group) return group->ownedCounts();
			return std::vector<uint32>(1, numParticles);
		}
		void FluidSimulation::multiGpuBreakdown(double out4[4])
		{
			for (int i = 0; i < 4; i++) out4[i] = 0.0;
			if (group) { group->breakdown(out4); group->resetBreakdown(); }
		}
		void FluidSimulation::fetch(int field, void* out, size_t bytes, const char* what)
		{
			if (group) {
				try { group->download(field, out, bytes); }
				catch (const std::exception& e) { error = std::string(what) + ": " + e.what(); throw; }
			} else check(sph_download(ctx, field, out, bytes), what);
		}
		void FluidSimulation::setTableMode(int m)
		{
			if (m != SPH_TABLE_GRID && (group || devices.size() > 1)) {
				error = "setTableMode: the slab decomposition uses the GRID table";
				throw std::runtime_error(error);
			}
			tableMode = m;
			if (ctx) check(sph_set_table_mode(ctx, m), "sph_set_table_mode");
		}
		void FluidSimulation::setHostMirrors(bool outPositions, bool positionsToo) { mirrorOut = outPositions; mirrorPos = positionsToo; }

		void FluidSimulation::InitializeData(int particleAmmount, vec3)
		{
			// (the reference ignores Centre too: GridArrangement is called without it, :142)
			if (particleAmmount < 0) particleAmmount = 0;
			numParticles = (uint32)particleAmmount;
			group.reset();
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
			densitiesValid = true;
			invalidateGetters();
			timingsFresh = true;
			if (devices.size() > 1 && numParticles) {
				// The lattice, its lookup and its densities (:139-145) were just made on the first device; the slabs
				// take the particles from here.  Until the first Update the getters answer from this spawn state.
				const size_t n = numParticles;
				bulkPos.resize(n * 3); bulkVel.assign(n * 3, 0.0f); bulkDens.resize(n * 2);
				check(sph_download(ctx, SPH_FIELD_POSITIONS, bulkPos.data(), bulkPos.size() * 4), "sph_download");
				check(sph_download(ctx, SPH_FIELD_DENSITIES, bulkDens.data(), bulkDens.size() * 4), "sph_download");
				sph_destroy(ctx); ctx = nullptr; capacity = 0;
				try {
					group.reset(new sphb200::SlabGroup(devices, numParticles, params));
					group->setDirectScatter(directMirrors);
					group->upload(numParticles, bulkPos.data(), nullptr);
				} catch (const std::exception& e) { group.reset(); error = e.what(); throw; }
				bulkFresh.store(true, std::memory_order_release);
				updatesSinceRebalance = 0;
			}
		}

		void FluidSimulation::invalidateGetters()
		{
			cacheValid = false;
			getterCalls = 0;
			bulkFresh.store(false, std::memory_order_release);
		}

		void FluidSimulation::setMaxTimestep(float maxDt) { maxTimestep = (maxDt > 0.0f) ? maxDt : 0.0f; }

		void FluidSimulation::Update(float deltatime)
		{
			if ((!ctx && !group) || numParticles == 0) return;
			substeps = 1;
			if (maxTimestep > 0.0f && deltatime > maxTimestep) {
				const float q = std::ceil(deltatime / maxTimestep);
				substeps = (q < 1.0f) ? 1u : (q > 1024.0f ? 1024u : (uint32)q);
			}
			if (group) {
				try {
					group->step(substeps == 1 ? deltatime : deltatime / (float)substeps, substeps);
					if (rebalanceEvery && ++updatesSinceRebalance >= rebalanceEvery) { updatesSinceRebalance = 0; group->rebalance(1); }
				} catch (const std::exception& e) { error = e.what(); throw; }
			}
			else if (substeps == 1) check(sph_step(ctx, deltatime), "sph_step");
			else check(sph_step_n(ctx, deltatime / (float)substeps, substeps), "sph_step_n");
			// Update() returns with the host-visible buffers complete: the renderer takes
			// &OutPositions[0] right after (fluidSimCPU.cc:58)
			if (mirrorOut) fetch(SPH_FIELD_OUT_POSITIONS, OutPositions.data(), OutPositions.size() * sizeof(vec4), "sph_download");
			if (mirrorPos) fetch(SPH_FIELD_POSITIONS, positions.data(), positions.size() * sizeof(vec3), "sph_download");
			if (!mirrorOut && !mirrorPos && ctx) check(sph_synchronize(ctx), "sph_synchronize");
			densitiesValid = true;
			invalidateGetters();
			timingsFresh = false;
		}

		void FluidSimulation::uploadState(const float* velocities3)
		{
			if (!ctx && !group) return;
			std::vector<float> vel;
			if (!velocities3 && numParticles) {
				vel.resize((size_t)numParticles * 3);
				fetch(SPH_FIELD_VELOCITIES, vel.data(), vel.size() * 4, "sph_download");
				velocities3 = vel.data();
			}
			if (group) {
				try { group->upload(numParticles, numParticles ? &positions[0].x : nullptr, velocities3); }
				catch (const std::exception& e) { error = e.what(); throw; }
			} else {
				check(sph_upload_state(ctx, numParticles, numParticles ? &positions[0].x : nullptr, velocities3), "sph_upload_state");
				check(sph_synchronize(ctx), "sph_synchronize");
			}
			densitiesValid = false;                      // the reference's densities would be stale too until the next Update
			invalidateGetters();
		}

		// ---- snapshots: the C ABI's file format (sph_api.cu: sph_save_state), host-side so it serves every device layout
		// header: "SPHB2002", u32 particle count, u32 sizeof(SphParams), SphParams, then pos3[n], vel3[n] ("SPHB2001", the
		// first version, has no size word and is still read)
		static bool writeSnapshotFileImpl(const std::string& path, uint32_t n, const SphParams& p, const float* pos3, const float* vel3)
		{
			FILE* f = fopen(path.c_str(), "wb");
			if (!f) return false;
			const char magic[8] = {'S', 'P', 'H', 'B', '2', '0', '0', '2'};
			const uint32_t psize = (uint32_t)sizeof(SphParams);
			bool ok = fwrite(magic, 1, 8, f) == 8 && fwrite(&n, 4, 1, f) == 1 && fwrite(&psize, 4, 1, f) == 1 && fwrite(&p, sizeof(SphParams), 1, f) == 1;
			ok = ok && (n == 0 || (fwrite(pos3, 12, n, f) == n && fwrite(vel3, 12, n, f) == n));
			return (fclose(f) == 0) && ok;
		}
		static bool readSnapshotFileImpl(const std::string& path, uint32_t& n, SphParams& p, std::vector<float>& pos3, std::vector<float>& vel3)
		{
			FILE* f = fopen(path.c_str(), "rb");
			if (!f) return false;
			char magic[8];
			bool ok = fread(magic, 1, 8, f) == 8 && fread(&n, 4, 1, f) == 1;
			const bool v2 = ok && memcmp(magic, "SPHB2002", 8) == 0;
			ok = ok && (v2 || memcmp(magic, "SPHB2001", 8) == 0);
			uint32_t psize = (uint32_t)sizeof(SphParams);
			if (ok && v2) ok = fread(&psize, 4, 1, f) == 1 && psize == (uint32_t)sizeof(SphParams);   // another ABI's parameter block
			ok = ok && fread(&p, sizeof(SphParams), 1, f) == 1;
			if (ok) {
				// the count comes from the file: it must fit the rest of the file and the reference's int particle count
				const long here = ftell(f);
				ok = here >= 0 && fseek(f, 0, SEEK_END) == 0;
				const long end = ok ? ftell(f) : -1;
				ok = ok && end >= here && n <= 0x7FFFFFFFu && (uint64_t)(end - here) >= (uint64_t)n * 24u && fseek(f, here, SEEK_SET) == 0;
			}
			if (ok) {
				pos3.resize((size_t)n * 3); vel3.resize((size_t)n * 3);
				ok = n == 0 || (fread(pos3.data(), 12, n, f) == n && fread(vel3.data(), 12, n, f) == n);
			}
			fclose(f);
			return ok;
		}

		void FluidSimulation::saveState(const std::string& path)
		{
			std::vector<float> pos((size_t)numParticles * 3), vel((size_t)numParticles * 3);
			if ((ctx || group) && numParticles) {
				fetch(SPH_FIELD_POSITIONS, pos.data(), pos.size() * 4, "sph_download");
				fetch(SPH_FIELD_VELOCITIES, vel.data(), vel.size() * 4, "sph_download");
			}
			if (!writeSnapshotFileImpl(path, numParticles, params, pos.data(), vel.data())) {
				error = "saveState: cannot write " + path;
				throw std::runtime_error(error);
			}
		}

		void FluidSimulation::loadState(const std::string& path)
		{
			uint32_t n = 0;
			SphParams p;
			std::vector<float> pos, vel;
			if (!readSnapshotFileImpl(path, n, p, pos, vel)) {
				error = "loadState: " + path + " is not a complete snapshot file";
				throw std::runtime_error(error);
			}
			params = p;
			pushParams();                                 // (an existing context is reused by InitializeData without a push)
			InitializeData((int)n);                       // sizes the contexts and the mirrors for n particles (and spawns: overwritten next)
			for (uint32_t i = 0; i < n; i++) positions[i] = vec3(pos[3 * (size_t)i], pos[3 * (size_t)i + 1], pos[3 * (size_t)i + 2]);
			uploadState(n ? vel.data() : nullptr);
			for (uint32_t i = 0; i < n; i++) OutPositions[i] = vec4(positions[i].x, positions[i].y, positions[i].z, 0.34f);
		}

		void FluidSimulation::downloadVelocities(std::vector<vec3>& out)
		{
			out.resize(numParticles);
			if ((ctx || group) && numParticles) fetch(SPH_FIELD_VELOCITIES, out.data(), out.size() * sizeof(vec3), "sph_download");
		}
		void FluidSimulation::downloadDensities(std::vector<float>& out)
		{
			out.resize((size_t)numParticles * 2);
			if ((ctx || group) && numParticles) fetch(SPH_FIELD_DENSITIES, out.data(), out.size() * 4, "sph_download");
		}
		void FluidSimulation::downloadColors(std::vector<vec4>& out)
		{
			out.resize(numParticles);
			if ((ctx || group) && numParticles) fetch(SPH_FIELD_COLORS, out.data(), out.size() * sizeof(vec4), "sph_download");
		}

		// per-particle getters: bounds-checked, zero when out of range (:151,157,163,168,174,180)
		void FluidSimulation::readParticle(uint32 index, float out[10])
		{
			for (int k = 0; k < 10; k++) out[k] = 0.0f;
			if ((!ctx && !group) || index >= numParticles) return;
			if (!bulkFresh.load(std::memory_order_acquire)) {
				std::lock_guard<std::mutex> lock(getterMutex);
				if (!bulkFresh.load(std::memory_order_relaxed)) {
					// (several GPUs: a particle lives on whichever slab owns it now, so every read is the bulk read)
					if (ctx && !(cacheValid && cachedIndex == index) && getterCalls < kSingleReads) {
						getterCalls++;
						check(sph_get_particle(ctx, index, cached), "sph_get_particle");
						cachedIndex = index;
						cacheValid = true;
					}
					if (cacheValid && cachedIndex == index) {
						for (int k = 0; k < 10; k++) out[k] = cached[k];
						return;
					}
					// many getters this frame: one bulk read, then every getter is a host read
					const size_t n = numParticles;
					bulkPos.resize(n * 3); bulkVel.resize(n * 3); bulkDens.assign(n * 2, 0.0f);
					fetch(SPH_FIELD_POSITIONS, bulkPos.data(), bulkPos.size() * 4, "sph_download");
					fetch(SPH_FIELD_VELOCITIES, bulkVel.data(), bulkVel.size() * 4, "sph_download");
					if (densitiesValid) fetch(SPH_FIELD_DENSITIES, bulkDens.data(), bulkDens.size() * 4, "sph_download");
					bulkFresh.store(true, std::memory_order_release);
				}
			}
			const size_t i = index;
			for (int k = 0; k < 3; k++) { out[k] = bulkPos[3 * i + k]; out[3 + k] = bulkVel[3 * i + k]; }
			out[6] = bulkDens[2 * i]; out[7] = bulkDens[2 * i + 1];
			// glm::length = sqrt(dot), dot = (x*x + y*y) + z*z (func_geometric.inl:48-55); :172-182
			const float sp = std::sqrt((out[3] * out[3] + out[4] * out[4]) + out[5] * out[5]);
			out[8] = sp;
			out[9] = (sp < 0.0f ? 0.0f : (sp > 1.5f ? 1.5f : sp)) / 1.5f;
		}
		FluidSimulation::vec3 FluidSimulation::getPosition(uint32 i) { float o[10]; readParticle(i, o); return vec3(o[0], o[1], o[2]); }
		FluidSimulation::vec3 FluidSimulation::getVelocity(uint32 i) { float o[10]; readParticle(i, o); return vec3(o[3], o[4], o[5]); }
		float FluidSimulation::getDensity(uint32 i) { float o[10]; readParticle(i, o); return o[6]; }
		float FluidSimulation::getNearDensity(uint32 i) { float o[10]; readParticle(i, o); return o[7]; }
		float FluidSimulation::getSpeed(uint32 i) { float o[10]; readParticle(i, o); return o[8]; }
		float FluidSimulation::getSpeedNormalzied(uint32 i) { float o[10]; readParticle(i, o); return o[9]; }

		void FluidSimulation::refreshTimings()
		{
			if (timingsFresh || (!ctx && !group)) return;
			if (group) {
				try { group->timings(timings); }
				catch (const std::exception& e) { error = e.what(); throw; }
			} else check(sph_get_timings(ctx, timings), "sph_get_timings");
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
		void FluidSimulation::setGravity(bool status) { SphParams p = params; p.gravity = status ? 1 : 0; commitParams(p); }
		bool FluidSimulation::getGravityStatus() { return params.gravity != 0; }
		void FluidSimulation::setInteractionRadius(float value) { SphParams p = params; p.interaction_radius = value; commitParams(p); }   // sqr_radius untouched (Q2)
		float FluidSimulation::getInteractionRadius() { return params.interaction_radius; }
		void FluidSimulation::setDensityTarget(float value) { SphParams p = params; p.target_density = value; commitParams(p); }
		float FluidSimulation::getDensityTarget() { return params.target_density; }
		void FluidSimulation::setPressureMultiplier(float value) { SphParams p = params; p.pressure_multiplier = value; commitParams(p); }
		float FluidSimulation::getPressureMultiplier() { return params.pressure_multiplier; }
		void FluidSimulation::setNearPressureMultiplier(float value) { SphParams p = params; p.near_pressure_multiplier = value; commitParams(p); }
		float FluidSimulation::getNearPressureMultiplier() { return params.near_pressure_multiplier; }
		void FluidSimulation::setViscosityStrength(float value) { SphParams p = params; p.viscosity_strength = value; commitParams(p); }
		float FluidSimulation::getViscosityStrength() { return params.viscosity_strength; }
		void FluidSimulation::setGravityScale(float value) { SphParams p = params; p.gravity_scale = value; commitParams(p); }
		float FluidSimulation::getGravityScale() { return params.gravity_scale; }
		void FluidSimulation::setBound(const vec3& value) { SphParams p = params; p.bound[0] = value.x; p.bound[1] = value.y; p.bound[2] = value.z; commitParams(p); }
		FluidSimulation::vec3 FluidSimulation::getBounds() { return vec3(params.bound[0], params.bound[1], params.bound[2]); }
	}
}

namespace sphb200 {
bool writeSnapshotFile(const std::string& path, uint32_t n, const SphParams& p, const float* pos3, const float* vel3)
{
	return Physics::Fluid::writeSnapshotFileImpl(path, n, p, pos3, vel3);
}
bool readSnapshotFile(const std::string& path, uint32_t& n, SphParams& p, std::vector<float>& pos3, std::vector<float>& vel3)
{
	return Physics::Fluid::readSnapshotFileImpl(path, n, p, pos3, vel3);
}
}
