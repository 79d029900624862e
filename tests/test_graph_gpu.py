"""-m gpu: sph_step / sph_step_n replay the step as a CUDA graph from the second consecutive step of an unchanged
configuration on; the results are bit-identical to plain launches (sph_set_graph_replay(ctx, 0)), the stage timers are
refreshed by a plain step every 16 steps, and a change of configuration falls back to a plain step and records again."""
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
    b.set_graph_replay(False)                       # b: plain launches throughout
    for s in (a, b):
        s.upload_state(sc["pos"], sc["vel"])
    a.step_n(scenes.DT, 12)
    for _ in range(12):
        b.step(scenes.DT)
    assert a.graph_replays() == 11 and b.graph_replays() == 0      # step 0 runs plainly, step 1 records and replays
    assert a.launch_count() == b.launch_count()
    for f in ("positions", "velocities", "densities", "predicted"):
        assert np.array_equal(_bits(a.download(f)), _bits(b.download(f))), f
    # the six stage timers describe the last plain step (step 0 here; every 16th step is a plain one)
    ta, tb = a.timings(), b.timings()
    assert np.all(ta > 0.0) and np.all(tb > 0.0), (ta, tb)
    a.step_n(scenes.DT, 20)
    for _ in range(20):
        b.step(scenes.DT)
    assert a.graph_replays() == 11 + 19             # one of the twenty refreshed the timers through plain launches
    assert np.array_equal(_bits(a.download("positions")), _bits(b.download("positions")))
    # a parameter change invalidates the recording: plain step, new recording, same answers as plain stepping
    for s in (a, b):
        s.set_params(viscosity_strength=0.1, gravity_scale=4.0)
    a.step_n(scenes.DT, 6)
    for _ in range(6):
        b.step(scenes.DT)
    assert a.graph_replays() == 35                  # one plain step, then five replays of the new recording
    assert np.array_equal(_bits(a.download("positions")), _bits(b.download("positions")))
    # a different dt as well, through sph_step this time: plain, then recorded and replayed
    for _ in range(2):
        a.step(0.004)
        b.step(0.004)
    assert a.graph_replays() == 36
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
    b.set_graph_replay(False)
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
    # the lists are long throughout, so by now both contexts have measured that and run the density pass on its deep
    # survivor stack -- the replaying one too: the depth is re-evaluated before replayed steps as well, a flip makes
    # the step run plainly and record again
    a.step_n(scenes.DT, 4)
    for _ in range(4):
        b.step(scenes.DT)
    assert a.density_stack_rows() == 72 and b.density_stack_rows() == 72
    assert np.array_equal(a.download("neighbour_count"), b.download("neighbour_count"))
    assert np.array_equal(_bits(a.download("positions")), _bits(b.download("positions")))
    a.close(); b.close()
