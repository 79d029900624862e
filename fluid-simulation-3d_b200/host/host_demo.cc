// host/host_demo.cc -- headless driver of the host class, the way FluidSimCPU drives the reference
// (fluidSimCPU.cc:9-46): InitializeData(n), then Update(dt) per frame.  Prints one line per run that
// tests/test_host_class_gpu.py compares with the same scene run through the C ABI from Python.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "FluidSimulation.h"

static unsigned long long fnv1a(const void* p, size_t n, unsigned long long h = 0xcbf29ce484222325ull)
{
    const unsigned char* b = (const unsigned char*)p;
    for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 0x100000001b3ull; }
    return h;
}

int main(int argc, char** argv)
{
    int n = argc > 1 ? atoi(argv[1]) : 10000;
    int steps = argc > 2 ? atoi(argv[2]) : 3;
    int mode = argc > 3 ? atoi(argv[3]) : SPH_TABLE_GRID;
    auto& sim = Physics::Fluid::FluidSimulation::getInstance();
    try {
        sim.setTableMode(mode);
        sim.setGravity(true);
        sim.setHostMirrors(true, true);
        sim.InitializeData(n);
        printf("init out0=(%.6f %.6f %.6f %.2f) rho0=%.6f\n", sim.OutPositions[0].x, sim.OutPositions[0].y,
               sim.OutPositions[0].z, sim.OutPositions[0].w, sim.getDensity(0));
        for (int s = 0; s < steps; s++) sim.Update(0.016667f);
        printf("n=%d steps=%d pos_fnv=%016llx out_fnv=%016llx\n", n, steps,
               fnv1a(sim.positions.data(), sim.positions.size() * 12),
               fnv1a(sim.OutPositions.data(), sim.OutPositions.size() * 16));
        printf("p0=(%.6f %.6f %.6f) v0=(%.6f %.6f %.6f) rho=%.6f nrho=%.6f oob=%.1f\n",
               sim.getPosition(0).x, sim.getPosition(0).y, sim.getPosition(0).z,
               sim.getVelocity(0).x, sim.getVelocity(0).y, sim.getVelocity(0).z,
               sim.getDensity(0), sim.getNearDensity(0), sim.getDensity((uint32)n));
        printf("timers_ms %.4f %.4f %.4f %.4f %.4f %.4f\n", sim.getElapsedTimeGravity(), sim.getElapsedTimeSpatial(),
               sim.getElapsedTimeDensity(), sim.getElapsedTimePressure(), sim.getElapsedTimeViscosity(),
               sim.getElapsedTimePosNColl());
    } catch (const std::exception& e) {
        fprintf(stderr, "host_demo failed: %s\n", e.what());
        return 2;
    }
    return 0;
}
