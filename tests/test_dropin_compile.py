"""CPU, only where /root/reference exists: the host class compiled the way the reference tree would compile it --
after the reference's own engine/config.h (glm types, uint32 typedef), with -DSPH_B200_USE_GLM -DSPH_B200_HAVE_UINT32 --
and driven with the call patterns of its two callers (fluidSimCPU.cc:13,29,38,45,58,106; gameApp.cc:306-410).  This is
the INTEGRATION.md drop-in claim as a compiler check: same names, same argument and return types (glm::vec3 / vec4),
public `positions` / `OutPositions` vectors of glm vectors.  The program is linked and run: without a device it must
fail loudly in InitializeData, with one it steps."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
HOST = os.path.join(ROOT, "fluid-simulation-3d_b200", "host")

TU = r'''
#include "config.h"                      // the reference's prelude: glm + the uint32 typedef (engine/config.h:12-39)
#include "FluidSimulation.h"             // stands in for physics/physicsWorld.h
#include <cstdio>
#include <type_traits>
#include <stdexcept>

using Sim = Physics::Fluid::FluidSimulation;
static_assert(std::is_same<decltype(Sim::getInstance().positions), std::vector<glm::vec3>>::value, "positions");
static_assert(std::is_same<decltype(Sim::getInstance().OutPositions), std::vector<glm::vec4>>::value, "OutPositions");
static_assert(std::is_same<decltype(Sim::getInstance().getPosition(0u)), glm::vec3>::value, "getPosition");
static_assert(std::is_same<decltype(Sim::getInstance().getBounds()), glm::vec3>::value, "getBounds");
static_assert(std::is_same<decltype(Sim::getInstance().getElapsedTimeDensity()), double>::value, "timers are doubles");
static_assert(!std::is_copy_constructible<Sim>::value, "singleton: copy deleted (physicsWorld.h)");

int main()
{
    try {
        const uint32 nrParticles = 4096;
        // FluidSimCPU::initialize / update / reset (fluidSimCPU.cc:9-46)
        Sim::getInstance().InitializeData(nrParticles);
        const void* upload = &Sim::getInstance().OutPositions[0];            // what glBufferData is handed (:29, :58)
        Sim::getInstance().Update(0.016667f);
        float normalized = Sim::getInstance().getSpeedNormalzied(7);         // updateColors (:106)
        // GameApp::RenderUI (gameApp.cc:306-410)
        float simTime = Sim::getInstance().getSimulationTime();
        double ms = Sim::getInstance().getElapsedTimeGravity() + Sim::getInstance().getElapsedTimeSpatial() +
                    Sim::getInstance().getElapsedTimeDensity() + Sim::getInstance().getElapsedTimePressure() +
                    Sim::getInstance().getElapsedTimeViscosity() + Sim::getInstance().getElapsedTimePosNColl();
        glm::vec3 pos = Sim::getInstance().getPosition(3);
        glm::vec3 vel = Sim::getInstance().getVelocity(3);
        float rho = Sim::getInstance().getDensity(3) + Sim::getInstance().getNearDensity(3) + Sim::getInstance().getSpeed(3);
        bool gravity = Sim::getInstance().getGravityStatus();
        Sim::getInstance().setGravity(!gravity);
        float v = Sim::getInstance().getInteractionRadius();      Sim::getInstance().setInteractionRadius(v);
        v = Sim::getInstance().getDensityTarget();                Sim::getInstance().setDensityTarget(v);
        v = Sim::getInstance().getPressureMultiplier();           Sim::getInstance().setPressureMultiplier(v);
        v = Sim::getInstance().getNearPressureMultiplier();       Sim::getInstance().setNearPressureMultiplier(v);
        v = Sim::getInstance().getViscosityStrength();            Sim::getInstance().setViscosityStrength(v);
        v = Sim::getInstance().getGravityScale();                 Sim::getInstance().setGravityScale(v);
        glm::vec3 bound = Sim::getInstance().getBounds();
        float b[3] = {bound.x, bound.y, bound.z};
        Sim::getInstance().setBound({b[0], b[1], b[2]});                     // brace-initialised, as at gameApp.cc:410
        Sim::getInstance().setSimulationTime(simTime + 1.0f);
        Sim::getInstance().InitializeData(nrParticles);                      // reset (fluidSimCPU.cc:45)
        std::printf("dropin ok %p %f %f %f %f %f\n", upload, normalized, ms, (double)pos.x + vel.y, rho, (double)glm::length(bound));
    } catch (const std::runtime_error& e) {
        std::printf("dropin threw: %s\n", e.what());
        return 2;
    }
    return 0;
}
'''


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "engine")), reason="/root/reference is not present on this box")
def test_host_class_compiles_and_links_the_way_the_reference_tree_would(tmp_path):
    import __graft_entry__ as g
    g.load_package()                                              # makes sure the libraries are built
    src = tmp_path / "dropin.cc"
    src.write_text(TU)
    exe = tmp_path / "dropin"
    cmd = ["g++", "-std=c++20", "-O1", "-Wall", "-Werror", "-DGLM_ENABLE_EXPERIMENTAL", "-DSPH_B200_USE_GLM", "-DSPH_B200_HAVE_UINT32",
           "-I", os.path.join(REF, "engine"), "-I", os.path.join(REF, "exts", "glm"), "-I", os.path.join(REF, "exts", "glm", "glm"),
           "-I", HOST, "-I", os.path.join(ROOT, "include"),
           str(src), os.path.join(HOST, "FluidSimulation.cc"), os.path.join(HOST, "SlabGroup.cc"), "-pthread",
           "-L", os.path.join(ROOT, "fluid-simulation-3d_b200"), "-lsph_b200",
           "-Wl,-rpath," + os.path.join(ROOT, "fluid-simulation-3d_b200"), "-o", str(exe)]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-4000:]
    run = subprocess.run([str(exe)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=120)
    import torch
    if torch.cuda.is_available():
        assert run.returncode == 0 and "dropin ok" in run.stdout, run.stdout
    else:
        assert run.returncode == 2 and "no CUDA device" in run.stdout, run.stdout


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "engine")), reason="/root/reference is not present on this box")
def test_fluidsimbase_backend_compiles_against_the_reference_headers():
    """FluidSimB200 with -DSPH_B200_IN_REFERENCE_TREE: a subclass of the reference's REAL FluidSimBase (fluidSimBase.h:5-14,
    which pulls in its Shader and RenderUtils::Camera and through them glew and glm).  Compile only: linking would need
    the application's render library and an OpenGL context."""
    cmd = ["g++", "-std=c++20", "-fsyntax-only", "-Wall", "-Werror", "-DGLM_ENABLE_EXPERIMENTAL", "-DGLEW_NO_GLU",
           "-DSPH_B200_USE_GLM", "-DSPH_B200_HAVE_UINT32", "-DSPH_B200_IN_REFERENCE_TREE",
           "-I", os.path.join(REF, "engine"), "-I", os.path.join(REF, "exts", "glm"), "-I", os.path.join(REF, "exts", "glm", "glm"),
           "-I", os.path.join(REF, "exts", "glew", "include"), "-I", os.path.join(REF, "projects", "Simulation", "code", "simulations"),
           "-I", HOST, "-I", os.path.join(ROOT, "include"), os.path.join(HOST, "FluidSimB200.cc")]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-4000:]


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "engine")), reason="/root/reference is not present on this box")
@pytest.mark.parametrize("caller", ["projects/Simulation/code/simulations/fluidSimCPU.cc", "projects/Simulation/code/gameApp.cc"])
def test_the_reference_callers_compile_unmodified_against_the_host_class(tmp_path, caller):
    """The two translation units of the reference that use the solver class -- the CPU adapter and the application with
    its ImGui panels -- compiled where they lie, UNMODIFIED, with `physics/physicsWorld.h` resolved to a one-line header
    that includes host/FluidSimulation.h (what INTEGRATION.md section 1 tells a maintainer to do).  Syntax and types
    only: linking them needs the application's render library, GLFW and an OpenGL context."""
    shim = tmp_path / "physics"
    shim.mkdir()
    (shim / "physicsWorld.h").write_text('#include "FluidSimulation.h"\n')
    cmd = ["g++", "-std=c++20", "-fsyntax-only", "-w", "-DGLM_ENABLE_EXPERIMENTAL", "-DGLEW_NO_GLU", "-DGLFW_INCLUDE_NONE",
           "-DSPH_B200_USE_GLM", "-DSPH_B200_HAVE_UINT32",
           "-I", str(tmp_path),                                   # wins over engine/physics/physicsWorld.h
           "-I", os.path.join(REF, "engine"), "-I", os.path.join(REF, "exts", "glm"), "-I", os.path.join(REF, "exts", "glm", "glm"),
           "-I", os.path.join(REF, "exts", "glew", "include"), "-I", os.path.join(REF, "exts", "glfw", "include"),
           "-I", os.path.join(REF, "exts", "imgui"), "-I", os.path.join(REF, "projects", "Simulation", "code"),
           "-I", os.path.join(REF, "projects", "Simulation", "code", "simulations"),
           "-I", HOST, "-I", os.path.join(ROOT, "include"), os.path.join(REF, caller)]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-4000:]
    # and the shim really was the header in use: with the class hidden the same command must fail
    (shim / "physicsWorld.h").write_text("namespace Physics { namespace Fluid { } }\n")
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode != 0 and "FluidSimulation" in r.stdout
