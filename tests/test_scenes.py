"""CPU: the synthetic-input generators: counter-based, range-splittable, in bounds."""
import numpy as np

import __graft_entry__ as g


def test_generators_are_counter_based_and_in_bounds():
    g.load_package()
    from fluid_simulation_3d_b200 import scenes
    sc = scenes.small_dam_break(10)
    ids = np.array([0, 5, 999, 123], np.uint64)
    sub, _, _ = scenes.dam_break(10, 10, 10, 7, ids=ids)
    assert np.array_equal(sub, sc["pos"][ids.astype(np.int64)])
    half = np.array(sc["bound"], np.float32) / 2
    assert np.all(np.abs(sc["pos"]) < half)
    # block sits in the -x / floor corner; id order is y-outer (top first), x, z-inner
    assert sc["pos"][:, 0].max() < 0 and sc["pos"][:, 1].max() < half[1] / 2
    assert sc["pos"][0, 1] > sc["pos"][-1, 1] and sc["pos"][1, 2] > sc["pos"][0, 2]


def test_named_configs_have_the_survey_geometry():
    g.load_package()
    from fluid_simulation_3d_b200 import scenes
    ids = np.arange(4, dtype=np.uint64)
    c2 = scenes.config("C2_dambreak_1M", ids=ids)
    assert c2["n"] == 1_000_000 and np.allclose(c2["bound"], (64.5, 32.25, 21.715))
    c4 = scenes.config("C4_dambreak_64M", ids=ids)
    assert c4["n"] == 64_000_000 and np.allclose(c4["bound"], (258, 129, 86.215))
    c5 = scenes.config("C5_column_8M", ids=ids)
    assert c5["n"] == 8_000_000 and c5["params"]["viscosity_strength"] == 1.0
    assert np.abs(c5["vel"]).max() <= 0.5
