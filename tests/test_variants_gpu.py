"""-m gpu: every candidate-enumeration variant and the neighbour-list overflow path give the same answers."""
import os
import subprocess
import sys

import numpy as np
import pytest

import __graft_entry__ as g
from helpers import check_step

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r"""
import sys
sys.path.insert(0, %r); sys.path.insert(0, %r)
import __graft_entry__ as g
from helpers import check_step
pkg = g.load_package()
from fluid_simulation_3d_b200 import scenes
for mode in (pkg.TABLE_GRID, pkg.TABLE_REFERENCE_HASH):
    check_step(pkg, scenes.small_dam_break(14), mode, scenes.DT)
    check_step(pkg, scenes.small_column(10, 24, 10), mode, scenes.DT)
print("VARIANT_OK")
"""


@pytest.mark.parametrize("env", [{"SPH_GATHER": "v1"}, {"SPH_GATHER": "v2"}, {"SPH_DENSITY": "walk"}, {"SPH_DENSITY": "pair"}, {"SPH_DENSITY": "2"}, {"SPH_SORT": "radix"},
                                 {"SPH_DENSITY": "list"}, {"SPH_VISC_NOW": "1"},
                                 {"SPH_GATHER": "tile"}, {"SPH_GATHER": "tile", "SPH_TILE_CAPN": "128"}, {"SPH_GATHER": "tile", "SPH_XSUB": "1"}, {"SPH_XSUB": "4"}],
                         ids=["walk_every_pass", "packed_two_phase", "list_with_walk_density", "list_with_pair_density", "list_with_pair2_density", "radix_sort_grid",
                              "scalar_list_density_on_grid", "packed_density_viscosity_without_weights",
                              "tile_generation_tma_staged", "tile_smallest_staging_buffer", "tile_whole_cells_xsub1", "xsub4"])
def test_enumeration_variants_match_oracle(env):
    e = dict(os.environ, **env)
    r = subprocess.run([sys.executable, "-c", SCRIPT % (ROOT, os.path.join(ROOT, "tests"))], env=e, cwd=ROOT,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0 and "VARIANT_OK" in r.stdout, r.stdout[-3000:]


@pytest.mark.parametrize("cap", [0, 4, 16])
def test_neighbour_list_overflow_and_disabled(pkg, cap):
    """cap = 0: no list; cap = 4 / 16: most / some particles overflow and fall back to walking the table."""
    from fluid_simulation_3d_b200 import scenes
    import helpers
    orig = pkg.FluidSimulation.__init__

    def patched(self, *a, **k):
        orig(self, *a, **k)
        self.set_neighbour_list_capacity(cap)
    pkg.FluidSimulation.__init__ = patched
    try:
        for mode in (pkg.TABLE_GRID, pkg.TABLE_REFERENCE_HASH):
            helpers.check_step(pkg, scenes.small_dam_break(12), mode, scenes.DT)
    finally:
        pkg.FluidSimulation.__init__ = orig


def test_long_lists_dense_stack(pkg):
    """The packed density kernel picks its survivor-stack depth from the list lengths it measured in the steps before
    (sph_density_stack_rows): the dense column has ~90 neighbours per particle, so the first step runs on the shallow
    stack -- every warp goes through several partial flushes -- and the next one on the deep stack.  Both must match the
    oracle, and each other bit for bit (the stack depth moves the flush points, not the order of the sums)."""
    from fluid_simulation_3d_b200 import scenes
    import helpers
    sc = scenes.small_column(14, 40, 14)
    out = helpers.check_step(pkg, sc, pkg.TABLE_GRID, scenes.DT)            # a fresh context: shallow stack
    assert out["mean_neighbours"] > 50
    sim = pkg.FluidSimulation(sc["n"], device=0, table_mode=pkg.TABLE_GRID, **sc["params"])
    try:
        sim.set_neighbour_list_capacity(192)
        sim.set_neighbour_count_tap(True)
        fields = ("densities", "vel_after_pressure", "vel_after_viscosity", "positions", "velocities", "neighbour_count")
        got = []
        for _ in range(2):
            sim.upload_state(sc["pos"], sc["vel"])
            sim.step(scenes.DT)
            got.append((sim.density_stack_rows(), [sim.download(f) for f in fields]))
        assert got[0][0] == 24 and got[1][0] == 72, (got[0][0], got[1][0])
        for f, a, b in zip(fields, got[0][1], got[1][1]):
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), f + ": deep and shallow stack differ"
        ref, ps, vs = helpers.oracle_step(sc, scenes.DT)
        helpers.assert_close("deep-stack density", got[1][1][0], ref.densities(), 0.0)
        helpers.assert_close("deep-stack velocities", got[1][1][4], ref.velocities(), (ps + vs)[:, None])
        assert np.array_equal(got[1][1][5], ref.neighbour_counts())
        # a sparse state in the same context brings the shallow stack back (hysteresis: below 40 rows per warp)
        sparse = scenes.small_dam_break(14)
        sim.set_params(**sparse["params"])
        for _ in range(3):
            sim.upload_state(sparse["pos"], sparse["vel"])
            sim.step(scenes.DT)
        assert sim.density_stack_rows() == 24
    finally:
        sim.close()


@pytest.mark.parametrize("mode", ["grid", "reference_hash"])
def test_pairs_on_the_cull_boundary(pkg, mode):
    """Isolated pairs whose squared distance sits within a few ulp of sqrRadius (below, on and above it): the packed
    cull's FMA-fused d^2 cannot decide these, the exact predicate `!(d2 > sqrRadius)` (physicsWorld.cc:357) must.
    Neighbour counts are compared bit-exactly with the oracle by check_step."""
    from fluid_simulation_3d_b200 import scenes
    import helpers
    rng = np.random.default_rng(11)
    m = 14
    g3 = np.stack(np.meshgrid(*[np.arange(m)] * 3, indexing="ij"), -1).reshape(-1, 3)
    centre = ((g3 - (m - 1) / 2.0) * 1.25 + (rng.random(g3.shape) - 0.5) * 0.2).astype(np.float32)
    d = rng.normal(size=g3.shape)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    scale = np.float32(0.35) * (1.0 + rng.integers(-6, 7, size=(len(g3), 1)) * 1.0e-7)
    other = (centre.astype(np.float64) + d * scale).astype(np.float32)
    pos = np.concatenate([centre, other]).astype(np.float32)
    # how many of the pairs are neighbours by the reference's own arithmetic: it must be a real mix
    o = other - centre
    d2 = (o[:, 0] * o[:, 0] + o[:, 1] * o[:, 1]) + o[:, 2] * o[:, 2]
    inside = int((~(d2 > np.float32(0.35) * np.float32(0.35))).sum())
    assert len(g3) // 5 < inside < 4 * len(g3) // 5
    sc = dict(pos=pos, vel=np.zeros_like(pos), n=len(pos), params=dict(gravity=0, bound=(20.0, 20.0, 20.0)))
    # integers and densities only: with the partner exactly on the kernel support the reference's pressure and
    # viscosity terms are exactly zero, so their cancellation-free scale (the float tolerance) is zero too
    sim = pkg.FluidSimulation(sc["n"], device=0, table_mode=pkg.TABLE_GRID if mode == "grid" else pkg.TABLE_REFERENCE_HASH, **sc["params"])
    try:
        sim.set_neighbour_count_tap(True)
        sim.upload_state(sc["pos"], sc["vel"])
        sim.step(scenes.DT)
        ref, _, _ = helpers.oracle_step(sc, scenes.DT)
        nc, nc_ref = sim.download("neighbour_count"), ref.neighbour_counts()
        assert np.array_equal(nc, nc_ref), "neighbour counts: %d differ" % int((nc != nc_ref).sum())
        assert int((nc == 2).sum()) == 2 * inside and int((nc == 1).sum()) == 2 * (len(g3) - inside)
        helpers.assert_close("density", sim.download("densities"), ref.densities(), 0.0)
        assert np.all(np.isfinite(sim.download("velocities")))
    finally:
        sim.close()


def test_multi_step_tracks_oracle(pkg, ob):
    """20 consecutive steps: integer outputs stay exact as long as the float state agrees to the last bit
    of the predicate; we assert the trajectory stays within a drift budget and never goes non-finite."""
    from fluid_simulation_3d_b200 import scenes
    sc = scenes.small_dam_break(12)
    sim = pkg.FluidSimulation(sc["n"], **sc["params"])
    sim.upload_state(sc["pos"], sc["vel"])
    o = ob.PortOracle(sc["n"], threads=4, **sc["params"])
    o.set_state(sc["pos"], sc["vel"])
    for s in range(20):
        sim.step(scenes.DT)
        o.step(scenes.DT, jacobi=True)
    ob.PortOracle.lib().oracle_set_threads(1)
    p, q = sim.download("positions"), o.positions()
    assert np.all(np.isfinite(p))
    assert np.abs(p - q).max() < 5e-3, np.abs(p - q).max()
    sim.close()


def test_getters_timers_and_colors(pkg, ob):
    from fluid_simulation_3d_b200 import scenes
    sim = pkg.FluidSimulation(5000, gravity=1)
    sim.spawn_grid(5000)                                    # InitializeData
    o = ob.PortOracle(5000, gravity=1)
    o.spawn_grid()
    assert np.array_equal(sim.download("positions").view(np.uint32), o.positions().view(np.uint32))
    assert np.allclose(sim.download("densities"), o.densities(), rtol=1e-5)
    o4 = sim.download("out_positions")
    assert np.all(o4[:, 3] == np.float32(0.34))
    sim.step(scenes.DT)
    o.step(scenes.DT)
    t = sim.timings()
    assert np.all(t > 0) and t.sum() < 1e3
    one = sim.get_particle(17)
    assert np.allclose(one[:3], sim.download("positions")[17]) and np.allclose(one[3:6], sim.download("velocities")[17])
    assert np.allclose(one[6:8], sim.download("densities")[17])
    assert np.all(sim.get_particle(5000) == 0) and np.all(sim.get_particle(2 ** 31) == 0)   # OOB -> zeros
    v = sim.download("velocities")
    spd = np.clip(np.linalg.norm(v, axis=1), 0, 1.5) / 1.5
    assert np.allclose(sim.download("speed_normalized"), spd, atol=1e-6)
    col = sim.download("colors")
    assert col.shape == (5000, 4) and np.all(col[:, 3] == 1.0) and np.all((col >= 0) & (col <= 1.0 + 1e-6))
    sim.close()


def test_speed_gradient_colours_match_the_reference_restatement(pkg):
    """SPH_FIELD_COLORS / SPH_FIELD_SPEED_NORMALIZED against oracle/colors.py, the numpy restatement of
    getSpeedNormalzied (physicsWorld.cc:178-182) and FluidSimCPU::updateColors (fluidSimCPU.cc:100-125): bit for bit, on
    velocities that cover all three gradient segments, the breakpoints themselves, rest and far beyond the clamp."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import colors
    rng = np.random.default_rng(3)
    n = 20000
    vel = (rng.standard_normal((n, 3)) * rng.choice([0.05, 0.4, 1.0, 4.0], (n, 1))).astype(np.float32)
    vel[0] = 0.0
    vel[1] = (np.float32(0.495), 0.0, 0.0)          # |v| / 1.5 == 0.33f exactly
    vel[2] = (0.0, np.float32(0.99), 0.0)           # ... == 0.66f
    vel[3] = (1.5, 0.0, 0.0)
    vel[4] = (30.0, -40.0, 5.0)
    pos = ((rng.random((n, 3)) - 0.5) * 8).astype(np.float32)
    sim = pkg.FluidSimulation(n)
    sim.upload_state(pos, vel)                      # the getters read the state as uploaded (no step)
    t = sim.download("speed_normalized")
    assert np.array_equal(t.view(np.uint32), colors.speed_normalized(vel).view(np.uint32))
    col = sim.download("colors")
    ref = colors.speed_colors(vel)
    assert np.array_equal(col.view(np.uint32), ref.view(np.uint32)), int((col != ref).any(axis=1).sum())
    segs = [int((t <= colors.B1).sum()), int(((t > colors.B1) & (t <= colors.B2)).sum()), int((t > colors.B2).sum())]
    assert min(segs) > 1000, segs
    sim.close()


def test_empty_and_tiny_inputs(pkg):
    sim = pkg.FluidSimulation(8)
    sim.upload_state(np.zeros((0, 3), np.float32))
    sim.step(0.016667)                                      # zero particles: a no-op, not an error
    assert sim.n == 0
    sim.upload_state(np.array([[0.1, 0.2, 0.3]], np.float32))
    sim.set_neighbour_count_tap(True)
    sim.step(0.016667)
    assert sim.download("neighbour_count")[0] == 1 and sim.download("densities")[0, 0] > 0
    with pytest.raises(pkg.SphError):
        sim.upload_state(np.zeros((9, 3), np.float32))      # above capacity
    sim.close()


def test_host_class_matches_c_abi(pkg):
    demo = os.path.join(ROOT, "fluid-simulation-3d_b200", "host", "host_demo")
    r = subprocess.run([demo, "4096", "3", "0"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0, r.stdout
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("n=4096")][0]
    fnv = dict(kv.split("=") for kv in line.split())
    ob = g.load_oracle()
    sim = pkg.FluidSimulation(4096, gravity=1)
    sim.spawn_grid(4096)
    for _ in range(3):
        sim.step(float(np.float32(0.016667)))
    assert ob.fnv1a64(sim.download("positions")) == fnv["pos_fnv"]
    assert ob.fnv1a64(sim.download("out_positions")) == fnv["out_fnv"]
    sim.close()


def test_snapshot_roundtrip_resumes_bit_exactly(pkg, tmp_path):
    """save -> load into a fresh context -> both continue to identical states (checkpoint / resume)."""
    from fluid_simulation_3d_b200 import scenes
    sc = scenes.small_dam_break(12)
    a = pkg.FluidSimulation(sc["n"], **sc["params"])
    a.upload_state(sc["pos"], sc["vel"])
    for _ in range(3):
        a.step(scenes.DT)
    path = str(tmp_path / "state.sphb")
    a.save_state(path)
    b = pkg.FluidSimulation(sc["n"])
    b.load_state(path)
    assert b.n == sc["n"] and list(b.get_params().bound) == list(a.get_params().bound) and b.get_params().gravity == 1
    assert np.array_equal(a.download("positions").view(np.uint32), b.download("positions").view(np.uint32))
    for _ in range(2):
        a.step(scenes.DT)
        b.step(scenes.DT)
    # a keeps its history-dependent device order, b restarts from index order: same physics, summation order may differ
    assert np.array_equal(a.download("neighbour_count"), b.download("neighbour_count"))
    assert np.abs(a.download("positions") - b.download("positions")).max() < 1e-5
    with pytest.raises(pkg.SphError):
        b.load_state(str(tmp_path / "missing.sphb"))
    a.close(); b.close()


def _quat(axis, angle):
    a = np.asarray(axis, np.float64)
    a = a / np.linalg.norm(a)
    return tuple(float(x) for x in (*(a * np.sin(angle / 2)), np.cos(angle / 2)))


def _rot(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def test_extras_off_is_the_reference_step_and_rotated_bound_contains_the_fluid(pkg):
    """SphExtras (the reference's TODO list): with the identity rotation and zero stickiness the step
    is bit-identical to a context that never heard of them; with the box rotated 30 degrees about z every particle stays
    inside the ROTATED box, S6 equals a numpy restatement of move + clamp + reflection in the box's own axes, and the
    neighbour search (table over the rotated box's bounding box) still agrees with itself in both table modes."""
    from fluid_simulation_3d_b200 import scenes
    sc = scenes.small_dam_break(16)
    dt = scenes.DT
    a = pkg.FluidSimulation(sc["n"], **sc["params"])
    b = pkg.FluidSimulation(sc["n"], **sc["params"])
    b.set_extras()                                           # defaults: off
    assert b.get_extras()["bound_rotation"] == (0.0, 0.0, 0.0, 1.0)
    for s in (a, b):
        s.upload_state(sc["pos"], sc["vel"])
        s.step_n(dt, 5)
    for f in ("positions", "velocities", "densities"):
        assert np.array_equal(a.download(f).view(np.uint32), b.download(f).view(np.uint32)), f
    a.close(); b.close()
    with pytest.raises(pkg.SphError):
        pkg.FluidSimulation(8).set_extras(bound_rotation=(0, 0, 0, 0))
    with pytest.raises(pkg.SphError):
        pkg.FluidSimulation(8).set_extras(stick_strength=1.0, stick_distance=0.0)

    q = _quat((0, 0, 1), np.pi / 6)
    R = _rot(q)
    half = np.array(sc["bound"], np.float64) / 2
    rng = np.random.default_rng(11)
    n = 20000
    local = (rng.random((n, 3)) - 0.5) * 2 * half * 0.999
    pos = (local @ R.T).astype(np.float32)                   # inside the rotated box, up to its walls
    vel = ((rng.random((n, 3)) - 0.5) * 40).astype(np.float32)
    prm = dict(sc["params"], gravity=0, viscosity_strength=0.0, pressure_multiplier=0.0, near_pressure_multiplier=0.0)
    outs = []
    for mode in (pkg.TABLE_GRID, pkg.TABLE_REFERENCE_HASH):
        sim = pkg.FluidSimulation(n, table_mode=mode, **prm)
        sim.set_extras(bound_rotation=q)
        sim.set_neighbour_count_tap(True)
        sim.upload_state(pos, vel)
        sim.step(dt)
        p1, v1 = sim.download("positions").astype(np.float64), sim.download("velocities").astype(np.float64)
        outs.append((sim.download("neighbour_count"), p1))
        # no forces (k = kn = mu = 0, no gravity): the step is S6 alone -- move, clamp, reflect in the box's axes
        l = pos.astype(np.float64) @ R + (vel.astype(np.float64) @ R) * dt
        u = vel.astype(np.float64) @ R
        hit = half[None, :] - np.abs(l) <= 0
        l = np.where(hit, half[None, :] * np.sign(l), l)
        u = np.where(hit, u * -0.95, u)
        assert hit.any() and np.abs(p1 - l @ R.T).max() < 2e-5 and np.abs(v1 - u @ R.T).max() < 2e-4
        assert np.all(np.abs(p1 @ R) <= half[None, :] * (1 + 1e-6) + 1e-5)
        sim.close()
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])


def test_stickiness_pulls_towards_a_near_wall_only(pkg):
    """stick_strength k, stick_distance d0: a particle at distance d < d0 from a wall gets the impulse
    dt * k * d * (1 - d / d0) towards that wall before S6 moves it; farther particles are untouched."""
    bound = (4.0, 4.0, 4.0)
    prm = dict(bound=bound, gravity=0, viscosity_strength=0.0, pressure_multiplier=0.0, near_pressure_multiplier=0.0)
    pos = np.array([[-1.9, 0.0, 0.0], [0.0, 1.7, 0.0], [0.0, 0.0, 0.0], [1.95, -1.95, 0.3]], np.float32)   # far apart: no neighbours
    dt, k, d0 = 0.01, 50.0, 0.4
    sim = pkg.FluidSimulation(4, **prm)
    sim.set_extras(stick_strength=k, stick_distance=d0)
    sim.upload_state(pos, np.zeros_like(pos))
    sim.step(dt)
    v = sim.download("velocities").astype(np.float64)
    imp = lambda d: dt * k * d * (1 - d / d0)
    assert abs(v[0, 0] + imp(0.1)) < 1e-5 and abs(v[0, 1]) < 1e-7 and abs(v[0, 2]) < 1e-7       # towards the -x wall
    assert abs(v[1, 1] - imp(0.3)) < 1e-5                                                        # towards the +y wall
    assert np.all(v[2] == 0)                                                                     # mid-box: nothing
    assert abs(v[3, 0] - imp(0.05)) < 1e-5 and abs(v[3, 1] + imp(0.05)) < 1e-5 and abs(v[3, 2]) < 1e-7   # a corner: both walls
    sim.close()
