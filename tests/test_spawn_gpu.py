"""-m gpu: the device-side scene spawn (sph_spawn_block) writes the same bits as the host
generator of the synthetic scenes (fluid-simulation-3d_b200/scenes.py), and stepping from it equals stepping from
an upload of the host arrays."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


@pytest.mark.parametrize("which", ["dam_break", "column", "centred_regular"])
def test_device_spawn_equals_host_generator(pkg, which):
    from fluid_simulation_3d_b200 import scenes
    if which == "dam_break":
        sc = scenes.small_dam_break(17)
    elif which == "column":
        sc = scenes.small_column(7, 20, 9)
    else:
        nx, ny, nz, gap, bound = 5, 3, 11, 0.215, (6.0, 5.0, 7.0)
        pos, vel = scenes.block(nx, ny, nz, gap, bound, 3, anchor="center", jitter=0.0)
        sc = dict(pos=pos, vel=vel, n=nx * ny * nz, params=dict(bound=bound),
                  spawn=scenes.device_spawn_args(nx, ny, nz, gap, bound, 3, anchor="center", jitter=0.0))
    sim = pkg.FluidSimulation(sc["n"], **sc["params"])
    sim.spawn_block(**sc["spawn"])
    assert sim.n == sc["n"]
    assert np.array_equal(_bits(sim.download("positions")), _bits(sc["pos"]))
    assert np.array_equal(_bits(sim.download("velocities")), _bits(sc["vel"]))
    # same state, same device order => the step is bit-identical to the uploaded scene's
    ref = pkg.FluidSimulation(sc["n"], **sc["params"])
    ref.upload_state(sc["pos"], sc["vel"])
    for s in (sim, ref):
        s.step(scenes.DT)
    for f in ("positions", "velocities", "densities"):
        assert np.array_equal(_bits(sim.download(f)), _bits(ref.download(f))), f
    sim.close(); ref.close()


def test_device_spawn_argument_errors(pkg):
    sim = pkg.FluidSimulation(100)
    with pytest.raises(pkg.SphError):
        sim.spawn_block(5, 5, 5, 0.215, (0, 0, 0))            # 125 > capacity
    with pytest.raises(pkg.SphError):
        sim.spawn_block(2, 2, 2, 0.0, (0, 0, 0))              # gap must be positive
    with pytest.raises(pkg.SphError):
        sim.spawn_block(2, 2, 2, 0.215, (0, float("nan"), 0))
    sim.spawn_block(0, 4, 4, 0.215, (0, 0, 0))                # an empty block is an empty scene
    assert sim.n == 0
    sim.close()
