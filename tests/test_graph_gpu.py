"""-m gpu: sph_step_n replays the step as a CUDA graph; the results are bit-identical to plain sph_step calls, and a
change of configuration between (or during) calls falls back to plain steps and records again."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


@pytest.mark.parametrize("mode", [0, 1], ids=["grid", "refhash"])
def test_step_n_graph_replay_equals_plain_steps(pkg, mode):
    from fluid_simulation_3d_b200 import scenes
    sc = scenes.small_dam_break(18)
    a = pkg.FluidSimulation(sc["n"], table_mode=mode, **sc["params"])
    b = pkg.FluidSimulation(sc["n"], table_mode=mode, **sc["params"])
    for s in (a, b):
        s.upload_state(sc["pos"], sc["vel"])
    a.step_n(scenes.DT, 12)
    for _ in range(12):
        b.step(scenes.DT)
    assert a.graph_replays() == 10                  # step 0 and the last step run plainly
    assert a.launch_count() == b.launch_count()
    for f in ("positions", "velocities", "densities", "predicted"):
        assert np.array_equal(_bits(a.download(f)), _bits(b.download(f))), f
    # the timers describe the last (plain) step
    assert a.timings().sum() > 0.0
    # a parameter change invalidates the recording: plain step, new recording, same answers as plain stepping
    for s in (a, b):
        s.set_params(viscosity_strength=0.1, gravity_scale=4.0)
    a.step_n(scenes.DT, 6)
    for _ in range(6):
        b.step(scenes.DT)
    assert a.graph_replays() == 14
    assert np.array_equal(_bits(a.download("positions")), _bits(b.download("positions")))
    # a different dt as well; short calls (< 3 steps) never record
    a.step_n(0.004, 2)
    for _ in range(2):
        b.step(0.004)
    assert a.graph_replays() == 14
    assert np.array_equal(_bits(a.download("positions")), _bits(b.download("positions")))
    a.close(); b.close()


def test_step_n_replay_with_growing_neighbour_lists(pkg):
    """a dense column overflows the default list capacity: the replay must hand over to a plain step that grows the
    list, and stay exact meanwhile.  All three force multipliers are zero, so the particles drift ballistically and the
    column stays dense (with forces on it explodes within three steps, and rounding-level differences between the
    list and the table-walk path would be amplified chaotically); every pass still runs over the full lists."""
    from fluid_simulation_3d_b200 import scenes
    sc = scenes.small_column(10, 30, 10)
    params = dict(sc["params"], pressure_multiplier=0.0, near_pressure_multiplier=0.0, viscosity_strength=0.0, gravity=0)
    a = pkg.FluidSimulation(sc["n"], **params)
    b = pkg.FluidSimulation(sc["n"], **params)
    for s in (a, b):
        s.set_neighbour_count_tap(True)
        s.upload_state(sc["pos"], sc["vel"])
    a.step_n(scenes.DT, 8)
    for _ in range(8):
        b.step(scenes.DT)
    assert a.graph_replays() >= 1
    na, nb = a.download("neighbour_count"), b.download("neighbour_count")
    assert nb.max() > 64                            # the default capacity did overflow
    assert np.array_equal(na, nb)
    assert np.array_equal(_bits(a.download("positions")), _bits(b.download("positions")))
    da, db = a.download("densities"), b.download("densities")
    assert np.all(np.abs(da - db) <= 1e-5 * np.maximum(np.abs(db), 1.0))
    a.close(); b.close()
