// host/FluidSimB200.h -- the B200 backend as a third FluidSimBase subclass.
//
// The reference application drives a simulation through the five virtuals of FluidSimBase
// (projects/Simulation/code/simulations/fluidSimBase.h:5-14), created by createSimulation(SimType)
// (gameApp.cc:87-92) next to FluidSimCPU and FluidSimGPU.  FluidSimB200 is that subclass for the B200 solver:
//
//   initialize(n)  InitializeData(n) on the device                       (cf. fluidSimCPU.cc:9-33)
//   update(dt)     Update(dt) -> OutPositions in host memory, plus the speed-gradient colours computed ON THE
//                  DEVICE (SPH_FIELD_COLORS) instead of FluidSimCPU::updateColors' host loop (:100-125)
//   reset()        InitializeData(n) again                                (:42-46)
//   cleanup()      nothing to free that the solver singleton does not own (:48-51)
//   render(s, c)   hands the frame -- N vec4 positions (w = 0.34), N vec4 colours -- to a PRESENTER.
//
// Everything OpenGL lives in the presenter, so this file compiles and is tested without a GL context
// (tests/test_host_gpu.py drives it through host_demo).  INTEGRATION.md section 5 shows the presenter a
// maintainer of the reference adds (the two glBufferData uploads and the batched draw of fluidSimCPU.cc:53-97).
//
// Inside the reference tree define SPH_B200_IN_REFERENCE_TREE: the real fluidSimBase.h (Shader,
// RenderUtils::Camera) is used.  Stand-alone, layout-free stand-ins with the same names are declared here.
#pragma once

#include <vector>

#include "FluidSimulation.h"

#ifdef SPH_B200_IN_REFERENCE_TREE
#include "fluidSimBase.h"
#else
class Shader;
namespace RenderUtils { class Camera; }
class FluidSimBase
{
public:
	virtual void update(float dt) = 0;
	virtual void initialize(int particleAmount) = 0;
	virtual void reset() = 0;
	virtual void cleanup() = 0;
	virtual void render(Shader& renderShader, RenderUtils::Camera& cam) = 0;
	virtual ~FluidSimBase() = default;
};
#endif

class FluidSimB200 : public FluidSimBase
{
public:
	using vec4 = sphb200::vec4;
	// One frame: `count` positions (x, y, z, 0.34) and colours (r, g, b, 1) in particle index order, host memory,
	// valid until the next update().  `shader` / `camera` are render()'s own arguments, passed through.
	struct Frame { const vec4* positions; const vec4* colors; int count; };
	typedef void (*Presenter)(const Frame& frame, Shader* shader, RenderUtils::Camera* camera, void* user);

	FluidSimB200() {}
	explicit FluidSimB200(int cudaDevice) : device(cudaDevice) {}

	void initialize(int particleAmount) override;
	void update(float dt) override;
	void reset() override;
	void cleanup() override;
	void render(Shader& renderShader, RenderUtils::Camera& cam) override;

	void setPresenter(Presenter fn, void* user) { presenter = fn; presenterUser = user; }
	void setMaxTimestep(float maxDt);           // sub-stepping, see FluidSimulation::setMaxTimestep
	Frame frame() const;                        // what render() hands to the presenter
	int particleCount() const { return nrParticles; }

private:
	int nrParticles = 0;
	int device = 0;
	std::vector<vec4> colors;
	Presenter presenter = nullptr;
	void* presenterUser = nullptr;
};
