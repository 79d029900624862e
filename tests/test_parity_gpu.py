"""-m gpu: the CUDA step (through the C ABI) against the oracle on identical input states."""
import numpy as np
import pytest

import __graft_entry__ as g
from helpers import check_step, coincident_scene

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def scenes(pkg):
    from fluid_simulation_3d_b200 import scenes as sc
    return sc


MODES = ["grid", "reference_hash"]


def _mode(pkg, name):
    return pkg.TABLE_GRID if name == "grid" else pkg.TABLE_REFERENCE_HASH


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("side", [8, 16, 24])
def test_dam_break_step(pkg, scenes, mode, side):
    out = check_step(pkg, scenes.small_dam_break(side), _mode(pkg, mode), scenes.DT)
    assert 5 < out["mean_neighbours"] < 40


@pytest.mark.parametrize("mode", MODES)
def test_dense_column_step(pkg, scenes, mode):
    out = check_step(pkg, scenes.small_column(14, 40, 14), _mode(pkg, mode), scenes.DT)
    assert out["mean_neighbours"] > 50


@pytest.mark.parametrize("mode", MODES)
def test_moving_state_negative_coords(pkg, scenes, mode):
    """Random velocities, block centred on the origin (all eight sign octants: Q5), some particles
    beyond the walls so the collision branch and the clamped border cells are exercised."""
    rng = np.random.default_rng(3)
    n = 6000
    bound = (6.0, 5.0, 4.0)
    pos = (rng.random((n, 3), dtype=np.float32) - 0.5) * np.array(bound, np.float32) * 1.02
    vel = (rng.random((n, 3), dtype=np.float32) - 0.5) * 8.0
    sc = dict(pos=pos.astype(np.float32), vel=vel.astype(np.float32), n=n,
              params=dict(gravity=1, viscosity_strength=0.7, bound=bound))
    check_step(pkg, sc, _mode(pkg, mode), scenes.DT)


@pytest.mark.parametrize("mode", MODES)
def test_gravity_off_defaults(pkg, scenes, mode):
    sc = scenes.small_dam_break(12)
    sc["params"] = dict(bound=sc["bound"])          # reference defaults: gravity off (physicsWorld.h:105)
    check_step(pkg, sc, _mode(pkg, mode), scenes.DT)


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("radius", [0.25, 0.5])
def test_q2_radius_differs_from_cutoff(pkg, scenes, mode, radius):
    """SURVEY App.A Q2: setInteractionRadius changes the cell size and the kernel support but NOT the
    d^2 cull (sqrRadius is a const member, physicsWorld.h:96).  With r = 0.25 the reference therefore
    misses neighbours that lie within 0.35 but outside its 27 smaller cells -- and so must we."""
    sc = scenes.small_dam_break(14)
    sc["params"] = dict(sc["params"], interaction_radius=radius)
    out = check_step(pkg, sc, _mode(pkg, mode), scenes.DT)
    assert out["mean_neighbours"] > 3


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("n", [1, 3, 33, 1001])
def test_tiny_and_odd_counts(pkg, scenes, mode, n):
    """Odd particle counts exercise the half-filled last record of the pair-interleaved positions; 1 and 3 the
    degenerate launches (one warp, mostly idle lanes)."""
    rng = np.random.default_rng(100 + n)
    bound = (3.0, 3.0, 3.0)
    pos = ((rng.random((n, 3)) - 0.5) * 1.2).astype(np.float32)
    vel = ((rng.random((n, 3)) - 0.5) * 2.0).astype(np.float32)
    sc = dict(pos=pos, vel=vel, n=n, params=dict(gravity=1, viscosity_strength=0.5, bound=bound))
    check_step(pkg, sc, _mode(pkg, mode), scenes.DT)


@pytest.mark.parametrize("mode", MODES)
def test_coincident_particles(pkg, scenes, mode):
    """Particles at exactly the same (predicted) position -- stacked in a wall corner by the box clamp, or spawned on
    top of each other: d == 0 passes every cull, takes the (0, 1, 0) direction fallback of the pressure pass
    (physicsWorld.cc:414) and the peak kernel weights; the pair records, the band re-test and the list replay must
    treat it like the reference."""
    sc = coincident_scene()
    out = check_step(pkg, sc, _mode(pkg, mode), scenes.DT)
    assert out["mean_neighbours"] > 2


def test_crowded_cells_keep_the_canonical_order(pkg):
    """Hundreds of particles in single cells (what a pile-up in a corner looks like): the counting sort restores the
    stable order inside a cell by bisection once a cell holds more than 32 rows (sph_kernels.cu: source_row); the sorted
    order, the neighbour counts and every float stage still match the reference, and no cell is reported non-canonical."""
    rng = np.random.default_rng(77)
    blobs = [np.array([0.1, -0.9, 0.2]), np.array([-1.3, -1.4, 1.2]), np.array([0.8, 0.3, -0.6])]
    pos = np.concatenate([(b + (rng.random((m, 3)) - 0.5) * 0.3) for b, m in zip(blobs, (500, 180, 40))] +
                         [(rng.random((300, 3)) - 0.5) * 2.6]).astype(np.float32)
    vel = ((rng.random(pos.shape) - 0.5) * 1.0).astype(np.float32)
    sc = dict(pos=np.ascontiguousarray(pos), vel=vel, n=len(pos), params=dict(gravity=1, viscosity_strength=0.4, bound=(3.0, 3.0, 3.0)))
    for mode in (pkg.TABLE_GRID, pkg.TABLE_REFERENCE_HASH):
        out = check_step(pkg, sc, mode, float(np.float32(0.016667)))
        assert out["mean_neighbours"] > 100
    sim = pkg.FluidSimulation(sc["n"], **sc["params"])
    sim.upload_state(sc["pos"], sc["vel"])
    sim.step(float(np.float32(0.016667)))
    assert sim.noncanonical_cells() == 0
    sim.close()
