// host/FluidSimB200.cc -- see FluidSimB200.h.
#ifdef SPH_B200_USE_GLM
#include "config.h"                      // inside the reference tree: its prelude first, as fluidSimCPU.cc:1 does
#endif
#include "FluidSimB200.h"

using Physics::Fluid::FluidSimulation;

void FluidSimB200::initialize(int particleAmount)
{
	nrParticles = particleAmount < 0 ? 0 : particleAmount;
	FluidSimulation& sim = FluidSimulation::getInstance();
	sim.setDevice(device);
	sim.InitializeData(nrParticles);                       // fluidSimCPU.cc:13
	// every particle starts in the slow-end colour of the gradient (fluidSimCPU.cc:19-23, Color1)
	colors.assign((size_t)nrParticles, vec4(0.0f, 0.75f, 1.0f, 1.0f));
}

void FluidSimB200::update(float dt)
{
	FluidSimulation& sim = FluidSimulation::getInstance();
	sim.Update(dt);                                        // fluidSimCPU.cc:38: OutPositions are in host memory now
	sim.downloadColors(colors);                            // :39 updateColors, evaluated by the export kernel
}

void FluidSimB200::reset()
{
	FluidSimulation::getInstance().InitializeData(nrParticles);   // fluidSimCPU.cc:45
	colors.assign((size_t)nrParticles, vec4(0.0f, 0.75f, 1.0f, 1.0f));
}

void FluidSimB200::cleanup() {}

void FluidSimB200::setMaxTimestep(float maxDt) { FluidSimulation::getInstance().setMaxTimestep(maxDt); }

FluidSimB200::Frame FluidSimB200::frame() const
{
	const FluidSimulation& sim = FluidSimulation::getInstance();
	Frame f;
	f.count = nrParticles;
	f.positions = nrParticles ? sim.OutPositions.data() : nullptr;
	f.colors = nrParticles ? colors.data() : nullptr;
	return f;
}

void FluidSimB200::render(Shader& renderShader, RenderUtils::Camera& cam)
{
	if (presenter) presenter(frame(), &renderShader, &cam, presenterUser);
}
