// host/FluidSimulation.h -- the reference's solver class, re-hosted on the B200 C ABI.
//
// Same namespace, class name, public member functions and public data members as
// Physics::Fluid::FluidSimulation in the reference (engine/physics/physicsWorld.h:29-153), so the
// application code that drives the CPU solver (fluidSimCPU.cc:13,29,38,45,58,106; gameApp.cc:306-410)
// compiles against it unchanged.  All arithmetic happens in libsph_b200.so (include/sph_b200.h);
// this class only owns the host mirrors the renderer reads (`positions`, `OutPositions`).
//
// Build inside the reference tree with -DSPH_B200_USE_GLM (after config.h, which pulls in glm), or
// stand-alone (this repo's tests) where vec3/vec4 are layout-compatible PODs.
#pragma once

#include <atomic>
#include <cstdint>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "sph_b200.h"

namespace sphb200 {
class SlabGroup;
// the snapshot file of sph_save_state / FluidSimulation::saveState, host-side (no device): "SPHB2001", u32 n, SphParams,
// n x pos3, n x vel3 (fp32, particle index order); false on I/O errors or a file that is not a complete snapshot
bool writeSnapshotFile(const std::string& path, uint32_t n, const SphParams& params, const float* pos3, const float* vel3);
bool readSnapshotFile(const std::string& path, uint32_t& n, SphParams& params, std::vector<float>& pos3, std::vector<float>& vel3);
}

#ifdef SPH_B200_USE_GLM
namespace sphb200 { using vec3 = glm::vec3; using vec4 = glm::vec4; }
#else
namespace sphb200 {
struct vec3 { float x, y, z; vec3() : x(0), y(0), z(0) {} vec3(float X, float Y, float Z) : x(X), y(Y), z(Z) {} };
struct vec4 { float x, y, z, w; vec4() : x(0), y(0), z(0), w(0) {} vec4(float X, float Y, float Z, float W) : x(X), y(Y), z(Z), w(W) {} };
}
#endif
#ifndef SPH_B200_HAVE_UINT32
typedef uint32_t uint32;     // engine/config.h:19-39
#endif

namespace Physics
{
	namespace Fluid
	{
		class FluidSimulation
		{
		public:
			using vec3 = sphb200::vec3;
			using vec4 = sphb200::vec4;

			static FluidSimulation& getInstance();                              // physicsWorld.cc:32-37

			void Update(float deltatime);                                        // :39-111
			void InitializeData(int particleAmmount, vec3 Centre = vec3(0, 0, 0)); // :112-147

			vec3 getPosition(uint32 particleIndex);                              // :149-153
			vec3 getVelocity(uint32 particleIndex);                              // :155-159
			float getDensity(uint32 particleIndex);                              // :161-165
			float getNearDensity(uint32 particleIndex);                          // :166-170
			float getSpeed(uint32 particleIndex);                                // :172-176
			float getSpeedNormalzied(uint32 particleIndex);                      // :178-182

			double getElapsedTimeGravity();                                      // :184-212
			double getElapsedTimeSpatial();
			double getElapsedTimeDensity();
			double getElapsedTimePressure();
			double getElapsedTimeViscosity();
			double getElapsedTimePosNColl();

			void setSimulationTime(float time);                                  // :214-302
			float getSimulationTime();
			void setGravity(bool status);
			bool getGravityStatus();
			void setInteractionRadius(float value);
			float getInteractionRadius();
			void setDensityTarget(float value);
			float getDensityTarget();
			void setPressureMultiplier(float value);
			float getPressureMultiplier();
			void setNearPressureMultiplier(float value);
			float getNearPressureMultiplier();
			void setViscosityStrength(float value);
			float getViscosityStrength();
			void setGravityScale(float value);
			float getGravityScale();
			void setBound(const vec3& value);
			// beyond the reference (its TODO list, README.md:38,43; physicsWorld.h:146-147): both off by default
			bool setBoundRotation(float qx, float qy, float qz, float qw);      // unit quaternion; (0,0,0,1) = axis-aligned
			bool setStickiness(float strength, float distance);                  // wall adhesion; strength 0 = off
			const SphExtras& getExtras() const { return extras; }
			vec3 getBounds();

			std::vector<vec3> positions;                                         // physicsWorld.h:79
			std::vector<vec4> OutPositions;                                      // physicsWorld.h:80

			// ---- B200 additions (no counterpart in the reference) ----
			void setDevice(int cudaDevice);            // before InitializeData; default 0
			// Several GPUs of the box (before InitializeData): the particles are slab-decomposed along z, one slab
			// per device, stepped in lockstep from one host thread each with ghost halos and migration over NCCL
			// (SlabGroup.h).  Everything else -- Update, the mirrors, getters, timers (the slowest slab per stage),
			// parameters -- behaves as on one GPU; results agree with the single-GPU step within fp32 summation
			// order.  GRID table only.  One device = setDevice.
			void setDevices(const std::vector<int>& cudaDevices);
			// Multi-GPU: move the slab planes towards the particle-count quantiles every N Updates (0 = never)
			void setRebalanceInterval(uint32 everyNUpdates) { rebalanceEvery = everyNUpdates; }
			// Multi-GPU, opt-in (before InitializeData): the page-locked mirrors (`positions`, `OutPositions`) are written
			// by the GPUs' export kernels directly, row by row at the particle's index, instead of copied per rank and
			// scattered on the host
			void setDirectMirrors(bool on) { directMirrors = on; }
			std::vector<uint32> particlesPerDevice() const;
			// Multi-GPU: host milliseconds since the last call, slowest slab per phase -- step, download (wait + export +
			// copy), scatter into index order, re-balancing; zeros on one GPU
			void multiGpuBreakdown(double out4[4]);
			// Release every device resource now (the destructor does the same, but at static-destruction time the
			// CUDA / NCCL runtimes may already be unloading): call before exit in multi-GPU programs.
			void shutdown();
			void setTableMode(int sphTableMode);       // SPH_TABLE_GRID (default) / SPH_TABLE_REFERENCE_HASH
			// Which host mirrors Update() refreshes: OutPositions (what the renderer uploads,
			// fluidSimCPU.cc:58) is on by default; `positions` costs a second copy and is off by default.
			void setHostMirrors(bool outPositions, bool positionsToo);
			// Push host-side edits of `positions` (and optionally velocities, n*3 floats) to the device.
			void uploadState(const float* velocities3 = nullptr);
			void downloadVelocities(std::vector<vec3>& out);
			void downloadDensities(std::vector<float>& rhoNearRhoPairs);
			void downloadColors(std::vector<vec4>& out);   // FluidSimCPU::updateColors on the device
			// State snapshots (the reference has none: Reset re-spawns).  The file is the C ABI's
			// (sph_save_state: "SPHB2001", u32 n, SphParams, n x pos3, n x vel3, particle index order), written and
			// read here on the host so it works the same on one GPU and on several -- a snapshot taken on one
			// configuration resumes on the other.  loadState replaces parameters, particle count and state.
			void saveState(const std::string& path);
			void loadState(const std::string& path);
			// Sub-stepping (the reference's README notes the pressure "going crazy" below
			// 40 fps): Update(dt) with dt > maxDt runs ceil(dt / maxDt) equal steps.  0 (default) = one step,
			// exactly the reference's behaviour.
			void setMaxTimestep(float maxDt);
			float getMaxTimestep() const { return maxTimestep; }
			uint32 lastSubsteps() const { return substeps; }
			SphContext* context() { return ctx; }
			const std::string& lastError() const { return error; }

		private:
			FluidSimulation();
			FluidSimulation(const FluidSimulation& cpy) = delete;
			~FluidSimulation();

			void ensureContext(uint32_t capacity);
			void pushParams();
			bool commitParams(const SphParams& candidate);
			bool commitExtras(const SphExtras& candidate);
			SphExtras extras = {{0.0f, 0.0f, 0.0f, 1.0f}, 0.0f, 0.0f};   // setters: apply everywhere or nowhere; failure -> lastError(), no throw
			void check(int rc, const char* what);
			void refreshTimings();
			void readParticle(uint32 index, float out10[10]);

			SphContext* ctx = nullptr;
			std::vector<int> devices;                            // more than one entry: slab mode through `group`
			std::unique_ptr<sphb200::SlabGroup> group;
			uint32 rebalanceEvery = 0, updatesSinceRebalance = 0;
			bool directMirrors = false;
			bool multi() const { return group != nullptr; }
			void fetch(int field, void* out, size_t bytes, const char* what);   // sph_download / SlabGroup::download
			SphParams params = defaultParams();
			static SphParams defaultParams() { SphParams p; sph_default_params(&p); return p; }
			uint32 numParticles = 0;
			uint32_t capacity = 0;
			int device = 0;
			int tableMode = SPH_TABLE_GRID;
			bool mirrorOut = true, mirrorPos = false;
			float simTime = 0.0f;
			double timings[6] = {0, 0, 0, 0, 0, 0};
			bool timingsFresh = true;
			// Per-particle getters.  The first few calls of a frame are single-particle device reads; a caller that
			// asks for many (FluidSimCPU::updateColors reads getSpeedNormalzied of EVERY particle from parallel
			// threads, fluidSimCPU.cc:100-106) gets one bulk read into host mirrors, after which the getters are
			// plain, thread-safe host reads until the next Update.
			static constexpr uint32 kSingleReads = 8;
			float cached[10] = {0};
			uint32 cachedIndex = 0xFFFFFFFFu;
			bool cacheValid = false;
			uint32 getterCalls = 0;
			bool densitiesValid = false;
			std::atomic<bool> bulkFresh{false};
			std::mutex getterMutex;
			std::vector<float> bulkPos, bulkVel, bulkDens;
			void invalidateGetters();
			float maxTimestep = 0.0f;
			uint32 substeps = 1;
			void* pinnedOut = nullptr; size_t pinnedOutBytes = 0;
			void* pinnedPos = nullptr; size_t pinnedPosBytes = 0;
			std::string error;
		};
	}
}
