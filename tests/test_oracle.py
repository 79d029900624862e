"""CPU: pin the C restatement (oracle/sph_oracle.c) against the golden vectors generated from the
UNMODIFIED reference, and -- where oracle/_ref is built -- against the reference itself, live."""
import glob
import os

import numpy as np
import pytest

import __graft_entry__ as g

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = sorted(glob.glob(os.path.join(HERE, "golden", "*.npz")))


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def params_of(gold):
    p = gold["params"]
    return dict(interaction_radius=p[0], target_density=p[1], pressure_multiplier=p[2], near_pressure_multiplier=p[3],
                viscosity_strength=p[4], gravity_scale=p[5], gravity=int(p[6]), bound=tuple(p[7:10]))


def test_golden_files_present():
    assert len(GOLDEN) >= 4


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_port_matches_golden_bit_for_bit(ob, path):
    """Given the reference's own sorted tie order, every intermediate of the restated step is
    bit-identical to the unmodified reference's."""
    gold = np.load(path)
    n = gold["pos0"].shape[0]
    dt = float(gold["dt"])
    o = ob.PortOracle(n, **params_of(gold))
    o.set_state(gold["pos0"], gold["vel0"])
    o.stage_predict(dt)
    assert np.array_equal(bits(o.predicted()), bits(gold["pred"]))
    o.stage_spatial(forced_order=gold["sorted_idx"])
    h, k, cells = o.hash_key()
    assert np.array_equal(h, gold["hash"]) and np.array_equal(k, gold["key"]) and np.array_equal(cells, gold["cells"])
    si, sh, sk = o.sorted_lookup()
    assert np.array_equal(sk, gold["sorted_key"]) and np.array_equal(sh, gold["sorted_hash"])
    assert np.array_equal(o.start_indices(), gold["start"])
    o.stage_density()
    assert np.array_equal(bits(o.densities()), bits(gold["dens"]))
    assert np.array_equal(o.neighbour_counts(), gold["ncount"])
    o.stage_pressure(dt)
    assert np.array_equal(bits(o.vel_after_pressure()), bits(gold["vel_press"]))
    o.stage_viscosity(dt, jacobi=True)
    assert np.array_equal(bits(o.vel_after_viscosity()), bits(gold["vel_visc"]))
    o.stage_integrate(dt)
    assert np.array_equal(bits(o.positions()), bits(gold["pos1"]))
    assert np.array_equal(bits(o.velocities()), bits(gold["vel1"]))
    assert np.array_equal(bits(o.out_positions()), bits(gold["out1"]))


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_port_own_sort_is_canonical_and_close(ob, path):
    """With its own (stable, index-ordered) sort the port's key sequence and start table are still
    bit-exact; floats differ from the reference only by summation order."""
    gold = np.load(path)
    n = gold["pos0"].shape[0]
    dt = float(gold["dt"])
    o = ob.PortOracle(n, **params_of(gold))
    o.set_state(gold["pos0"], gold["vel0"])
    o.step(dt, jacobi=True)
    si, sh, sk = o.sorted_lookup()
    assert np.array_equal(sk, gold["sorted_key"])
    assert np.array_equal(o.start_indices(), gold["start"])
    canon = gold["sorted_idx"][np.lexsort((gold["sorted_idx"], gold["sorted_key"]))]
    assert np.array_equal(si, canon)
    assert np.array_equal(o.neighbour_counts(), gold["ncount"])
    assert np.allclose(o.densities(), gold["dens"], rtol=2e-6, atol=0)
    assert np.allclose(o.positions(), gold["pos1"], rtol=0, atol=2e-5)


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_fixture_comparison_used_by_the_gpu_tests(ob, path):
    """helpers.compare_to_golden is what tests/test_zgolden_gpu.py holds the CUDA step against; here the restatement
    (own sort, own summation order) goes through it, so the comparison code itself is exercised without a GPU"""
    from helpers import compare_to_golden, golden_params
    gold = np.load(path)
    o = ob.PortOracle(gold["pos0"].shape[0], **golden_params(gold))
    o.set_state(gold["pos0"], gold["vel0"])
    o.step(float(gold["dt"]), jacobi=True)
    h, k, _ = o.hash_key()
    si, _, sk = o.sorted_lookup()
    got = dict(predicted=o.predicted(), hash=h, key=k, neighbour_count=o.neighbour_counts(), sorted_key=sk, sorted_index=si,
               start_indices=o.start_indices(), densities=o.densities(), vel_after_pressure=o.vel_after_pressure(),
               vel_after_viscosity=o.vel_after_viscosity(), positions=o.positions(), velocities=o.velocities(),
               out_positions=o.out_positions())
    worst = compare_to_golden(got.__getitem__, gold, reference_table=True)
    assert max(worst.values()) <= 1e-5
    bad = dict(got, densities=got["densities"] * np.float32(1.0001))            # and it does notice a 1e-4 error
    with pytest.raises(AssertionError):
        compare_to_golden(bad.__getitem__, gold, reference_table=True)


def test_port_in_place_viscosity_matches_verbatim_update(ob):
    """Gauss-Seidel (index-order, in-place) viscosity of the port == the reference's verbatim Update()."""
    gold = np.load(os.path.join(HERE, "golden", "dambreak_12.npz"))
    n = gold["pos0"].shape[0]
    dt = float(gold["dt"])
    o = ob.PortOracle(n, **params_of(gold))
    o.set_state(gold["pos0"], gold["vel0"])
    o.stage_predict(dt)
    o.stage_spatial(forced_order=gold["sorted_idx"])
    o.stage_density(); o.stage_pressure(dt); o.stage_viscosity(dt, jacobi=False); o.stage_integrate(dt)
    assert np.array_equal(bits(o.positions()), bits(gold["upd_pos"]))
    assert np.array_equal(bits(o.velocities()), bits(gold["upd_vel"]))
    assert np.array_equal(bits(o.densities()), bits(gold["upd_dens"]))
    # and the two viscosity semantics really differ (SURVEY App.A Q11)
    assert not np.array_equal(bits(gold["upd_vel"]), bits(gold["vel1"]))


def test_spawn_matches_reference_initialize_data(ob):
    gold = np.load(os.path.join(HERE, "golden", "spawn_1000.npz"))
    o = ob.PortOracle(1000, gravity=1)
    o.spawn_grid()
    assert np.array_equal(bits(o.positions()), bits(gold["pos0"]))
    assert np.allclose(o.densities(), gold["spawn_dens"], rtol=2e-6)


def test_smoothing_kernels_known_answers(ob):
    """kernels.h:25-82 closed forms, incl. the float abs() of Q3 (an int abs() would give inf)."""
    r = 0.35
    k = ob.PortOracle.kernels(0.1, r)
    pi = np.pi
    want = [(r - 0.1) ** 2 * 15 / (2 * pi * r ** 5), (r - 0.1) ** 3 * 15 / (pi * r ** 6), -(r - 0.1) * 15 / (pi * r ** 5),
            -(r - 0.1) ** 2 * 45 / (pi * r ** 6), (r * r - 0.01) ** 3 * 315 / (64 * pi * r ** 9)]
    assert np.allclose(k, want, rtol=2e-6)
    assert abs(k[4] - 28.3026) < 1e-3                       # SURVEY 8(c) probe value
    edge = ob.PortOracle.kernels(np.float32(r), np.float32(r))
    assert edge[0] == 0 and edge[1] == 0 and edge[4] == 0   # dist < radius is strict
    assert np.all(ob.PortOracle.kernels(0.5, r) == 0)


def test_negative_cell_hash_wraps(ob):
    """Q5: (uint32_t) of a negative cell wraps two's-complement; Q6: floor of true division."""
    o = ob.PortOracle(4, bound=(100, 100, 100))
    pos = np.array([[-0.1, -0.36, 0.34], [0.0, 0.35, -0.35], [-7.0, 3.3, -12.2], [0.7, -0.7, 0.0]], np.float32)
    o.set_state(pos, np.zeros_like(pos))
    o.stage_predict(0.0)
    h, k, c = o.hash_key()
    r = np.float32(0.35)
    want_c = np.floor(pos / r).astype(np.int64)
    assert np.array_equal(c, want_c)
    want_h = (want_c[:, 0] * 15823 + want_c[:, 1] * 9737333 + want_c[:, 2] * 440817757) % (1 << 32)
    assert np.array_equal(h.astype(np.int64), want_h)
    assert np.array_equal(k, (want_h % 4).astype(np.uint32))


def test_q13_two_of_the_27_cells_never_share_a_float_hash():
    """the reference would walk a bucket twice (and double-count its particles) if two of the 27
    queried cells agreed in key AND in float(hash).  They cannot: the hashes of two cells of one 3x3x3 block differ by
    at least 15823 (mod 2^32), and a u32 rounded to fp32 moves by at most 128 -- so the GRID table's exact cell walk
    and the reference's bucket walk always visit the same particles."""
    smallest = min(min(d, 2 ** 32 - d)
                   for a in range(-2, 3) for b in range(-2, 3) for c in range(-2, 3) if (a, b, c) != (0, 0, 0)
                   for d in [(a * 15823 + b * 9737333 + c * 440817757) % 2 ** 32])
    assert smallest == 15823
    worst_rounding = max(abs(int(np.float32(h)) - h) for h in (2 ** 32 - 129, 2 ** 31 + 128, 2 ** 31 - 65, 3000000001))
    assert worst_rounding <= 128 and 2 * worst_rounding < smallest


def test_empty_and_single_particle(ob):
    o = ob.PortOracle(1)
    o.set_state(np.zeros((1, 3), np.float32), np.zeros((1, 3), np.float32))
    o.step(0.016667)
    d = o.densities()
    assert d[0, 0] > 0 and o.neighbour_counts()[0] == 1          # density includes self (Q7)
    assert np.all(o.velocities() == 0)


needs_ref = pytest.mark.skipif(not g.load_oracle().have_ref(), reason="oracle/_ref not built (no /root/reference here)")


@needs_ref
def test_live_reference_staged_equals_update_without_viscosity(ob):
    """The harness' restated S1/S6 lambdas are bit-identical to the reference's Update() (mu = 0)."""
    g.load_package()
    from fluid_simulation_3d_b200 import scenes
    sc = scenes.small_dam_break(10)
    p = dict(sc["params"], viscosity_strength=0.0)
    r = ob.RefOracle(sc["n"], **p)
    r.set_state(sc["pos"], sc["vel"])
    r.update(scenes.DT)
    a = (r.positions(), r.velocities(), r.densities(), r.out_positions())
    r.set_state(sc["pos"], sc["vel"])
    r.step(scenes.DT, jacobi=True)
    b = (r.positions(), r.velocities(), r.densities(), r.out_positions())
    for x, y in zip(a, b):
        assert np.array_equal(bits(x), bits(y))
    r.set_state(sc["pos"], sc["vel"])
    r.set_params(viscosity_strength=0.5)
    r.update(scenes.DT)
    c = r.velocities()
    r.set_state(sc["pos"], sc["vel"])
    r.step(scenes.DT, jacobi=False)
    assert np.array_equal(bits(c), bits(r.velocities()))


@needs_ref
def test_live_reference_vs_port_default_scene(ob):
    """Default scene (InitializeData(10000), the reference's own spawn), 3 steps, in-place viscosity."""
    dt = float(np.float32(0.016667))
    r = ob.RefOracle(10000, spawn=True, gravity=1)
    p = ob.PortOracle(10000, gravity=1)
    p.spawn_grid()
    assert np.array_equal(bits(p.positions()), bits(r.positions()))
    for _ in range(3):
        r.update(dt)
        idx, _, _ = r.sorted_lookup()
        p.stage_predict(dt); p.stage_spatial(forced_order=idx); p.stage_density(); p.stage_pressure(dt)
        p.stage_viscosity(dt, jacobi=False); p.stage_integrate(dt)
        assert np.array_equal(bits(p.positions()), bits(r.positions()))
        assert np.array_equal(bits(p.velocities()), bits(r.velocities()))
        assert np.array_equal(bits(p.densities()), bits(r.densities()))


@needs_ref
def test_live_reference_vs_port_coincident_particles(ob):
    """dist == 0 (stacked particles): the restatement follows the unmodified reference bit for bit through the
    (0, 1, 0) direction fallback (physicsWorld.cc:414) -- two steps, the second from the state the first one leaves"""
    from helpers import coincident_scene
    sc = coincident_scene()
    dt = float(np.float32(0.016667))
    r = ob.RefOracle(sc["n"], **sc["params"])
    p = ob.PortOracle(sc["n"], **sc["params"])
    r.set_state(sc["pos"], sc["vel"]); p.set_state(sc["pos"], sc["vel"])
    pred0 = sc["pos"] + sc["vel"] * np.float32(1.0 / 120.0)
    assert len(np.unique(pred0, axis=0)) < sc["n"] - 8            # the scene really stacks particles
    for _ in range(2):
        r.update(dt)
        idx, _, _ = r.sorted_lookup()
        p.stage_predict(dt); p.stage_spatial(forced_order=idx); p.stage_density(); p.stage_pressure(dt)
        p.stage_viscosity(dt, jacobi=False); p.stage_integrate(dt)
        assert np.all(np.isfinite(r.positions())) and np.all(np.isfinite(r.velocities()))
        assert np.array_equal(bits(p.positions()), bits(r.positions()))
        assert np.array_equal(bits(p.velocities()), bits(r.velocities()))
        assert np.array_equal(bits(p.densities()), bits(r.densities()))


@needs_ref
def test_restatement_fuzzed_against_the_live_reference(ob):
    """80 random problems -- particle count 1..1500, box, interaction radius (also != the cut-off, Q2), every solver
    parameter, gravity on/off, dt, positions partly outside the box, fast particles, stacked duplicates -- one viscous
    staged step each: the restatement (given the reference's sorted tie order) equals the unmodified reference bit for
    bit on every buffer, and on hash, key, start table and neighbour counts"""
    rng = np.random.default_rng(2026)
    for case in range(80):
        n = int(rng.integers(1, 1500))
        bound = tuple(float(x) for x in rng.uniform(2.0, 12.0, 3))
        r = float(rng.choice([0.35, 0.35, 0.25, 0.5, float(rng.uniform(0.2, 0.6))]))
        prm = dict(interaction_radius=r, target_density=float(rng.uniform(20, 200)), pressure_multiplier=float(rng.uniform(10, 500)),
                   near_pressure_multiplier=float(rng.uniform(1, 40)), viscosity_strength=float(rng.uniform(0, 1)),
                   gravity_scale=float(rng.uniform(0, 20)), gravity=int(rng.integers(0, 2)), bound=bound)
        half = np.array(bound, np.float32) / 2
        pos = ((rng.random((n, 3)) - 0.5) * 2 * half * rng.choice([0.3, 0.9, 1.2])).astype(np.float32)
        vel = ((rng.random((n, 3)) - 0.5) * rng.choice([0.0, 2.0, 16.0])).astype(np.float32)
        if n > 4 and rng.random() < 0.5:
            d = rng.integers(0, n, max(1, n // 10)); s_ = rng.integers(0, n, len(d))
            pos[d] = pos[s_]; vel[d] = vel[s_]
        dt = float(np.float32(rng.choice([0.016667, 0.005, 0.033])))
        a, p = ob.RefOracle(n, **prm), ob.PortOracle(n, **prm)
        a.set_state(pos, vel); p.set_state(pos, vel)
        a.step(dt, jacobi=True)
        p.stage_predict(dt); p.stage_spatial(forced_order=a.sorted_lookup()[0]); p.stage_density(); p.stage_pressure(dt)
        p.stage_viscosity(dt, jacobi=True); p.stage_integrate(dt)
        what = "case %d (n=%d, r=%g)" % (case, n, r)
        for f in ("predicted", "densities", "vel_after_pressure", "vel_after_viscosity", "positions", "velocities", "out_positions"):
            assert np.array_equal(bits(getattr(p, f)()), bits(getattr(a, f)())), what + ": " + f
        for x, y in zip(p.hash_key(), a.hash_key()):
            assert np.array_equal(x, y), what + ": hash / key / cell"
        assert np.array_equal(p.start_indices(), a.start_indices()), what + ": start table"
        assert np.array_equal(p.neighbour_counts(), a.neighbour_counts()), what + ": neighbour counts"


needs_ref_par = pytest.mark.skipif(not g.load_oracle().have_ref_par(), reason="oracle/_ref/libsph_ref_par.so not built")


@needs_ref
@needs_ref_par
def test_parallel_reference_build_equals_the_serial_one_where_it_is_race_free(ob):
    """oracle/_ref/libsph_ref_par.so is the same unmodified TU with its std::execution::par loops really parallel
    (oracle/pstl_threads).  With mu = 0 nothing in Update() races, so several whole steps are bit-identical to the serial
    build; with viscosity on the in-place update is a data race (Q11) and only the stages before it are compared."""
    g.load_package()
    from fluid_simulation_3d_b200 import scenes
    sc = scenes.small_dam_break(20)
    dt = scenes.DT
    assert ob.RefOracleParallel.set_threads(4) == 4
    p0 = dict(sc["params"], viscosity_strength=0.0)
    a, b = ob.RefOracle(sc["n"], **p0), ob.RefOracleParallel(sc["n"], **p0)
    for o in (a, b):
        o.set_state(sc["pos"], sc["vel"])
        for _ in range(3):
            o.update(dt)
    for f in ("positions", "velocities", "densities", "out_positions", "predicted"):
        assert np.array_equal(bits(getattr(a, f)()), bits(getattr(b, f)())), f
    assert np.array_equal(a.sorted_lookup()[2], b.sorted_lookup()[2]) and np.array_equal(a.start_indices(), b.start_indices())
    for o in (a, b):                                   # viscosity on: everything up to the pressure stage still agrees
        o.set_params(viscosity_strength=0.5)
        o.set_state(sc["pos"], sc["vel"])
        o.update(dt)
    assert np.array_equal(bits(a.densities()), bits(b.densities()))
    assert np.array_equal(a.neighbour_counts(), b.neighbour_counts())
    # after the racy stage only sanity holds: Gauss-Seidel in index order (serial) against whatever interleaving the
    # threads produced moves velocities by up to ~0.1 in one step of this scene
    assert np.all(np.isfinite(b.positions())) and np.abs(a.positions() - b.positions()).max() < 0.05


@needs_ref_par
@pytest.mark.parametrize("side,seed,steps", [(50, 0x51, 2), (100, 0xC2, 1)], ids=["125k", "C2_full_1M"])
def test_restatement_pinned_at_size_against_the_parallel_reference(ob, side, seed, steps):
    """the pin of the restatement at sizes the serial reference makes slow -- 50^3 particles, and the WHOLE C2 workload
    (100^3, seed 0xC2: the state bench.py and tests/test_fullsize_gpu.py use, which closes the chain CUDA step <->
    restatement <-> unmodified reference at full size): whole steps of the
    unmodified reference on every host thread (mu = 0: race-free, bit-identical to its serial build) against the
    restatement given the reference's sorted tie order -- every buffer bit for bit.  At this size many cells share a
    bucket of the `hash % N` table, the regime the 10 k scenes barely touch."""
    g.load_package()
    from fluid_simulation_3d_b200 import scenes
    sc = scenes.small_dam_break(side, seed=seed)
    dt = scenes.DT
    prm = dict(sc["params"], viscosity_strength=0.0)
    ob.RefOracleParallel.set_threads(os.cpu_count() or 1)
    r = ob.RefOracleParallel(sc["n"], **prm)
    p = ob.PortOracle(sc["n"], threads=os.cpu_count() or 1, **prm)
    r.set_state(sc["pos"], sc["vel"]); p.set_state(sc["pos"], sc["vel"])
    try:
        for _ in range(steps):
            r.update(dt)
            idx, _, key = r.sorted_lookup()
            p.stage_predict(dt); p.stage_spatial(forced_order=idx); p.stage_density(); p.stage_pressure(dt)
            p.stage_viscosity(dt, jacobi=False); p.stage_integrate(dt)
            assert np.array_equal(p.sorted_lookup()[2], key) and np.array_equal(p.start_indices(), r.start_indices())
            assert np.array_equal(bits(p.positions()), bits(r.positions()))
            assert np.array_equal(bits(p.velocities()), bits(r.velocities()))
            assert np.array_equal(bits(p.densities()), bits(r.densities()))
        h, k, cells = r.hash_key()
        shared = len(np.unique(np.stack([k, h], 1), axis=0)) - len(np.unique(k))
        assert shared > 100, "meant to exercise buckets shared by several cells (got %d)" % shared
        if side == 100:
            # and the viscous step the GPU is held against: snapshot (Jacobi) viscosity produced from the UNMODIFIED
            # CalculateViscosityForce by call / record / restore (Q11), every stage buffer and the neighbour counts
            r.set_params(viscosity_strength=0.5); p.set_params(viscosity_strength=0.5)
            r.set_state(sc["pos"], sc["vel"]); p.set_state(sc["pos"], sc["vel"])
            r.step(dt, jacobi=True)
            p.stage_predict(dt); p.stage_spatial(forced_order=r.sorted_lookup()[0]); p.stage_density(); p.stage_pressure(dt)
            p.stage_viscosity(dt, jacobi=True); p.stage_integrate(dt)
            for f in ("densities", "vel_after_pressure", "vel_after_viscosity", "positions", "velocities", "out_positions"):
                assert np.array_equal(bits(getattr(p, f)()), bits(getattr(r, f)())), f
            assert np.array_equal(p.neighbour_counts(), r.neighbour_counts())
    finally:
        ob.PortOracle.lib().oracle_set_threads(1)


def test_pstl_threads_stand_in_really_runs_par_loops_in_parallel(tmp_path):
    """oracle/pstl_threads/tbb/tbb.h: libstdc++ picks its parallel PSTL backend when <tbb/tbb.h> is found; the stand-in's
    parallel_for must spread a std::for_each(par) over the OpenMP threads and visit every element exactly once"""
    import subprocess
    root = os.path.dirname(HERE)
    src = tmp_path / "probe.cc"
    src.write_text(r'''
#include <execution>
#include <algorithm>
#include <vector>
#include <cstdio>
#include <omp.h>
int main() {
    std::vector<unsigned> idx(300000);
    for (unsigned i = 0; i < idx.size(); i++) idx[i] = i;
    std::vector<int> tid(idx.size(), -1), hits(idx.size(), 0);
    std::for_each(std::execution::par, idx.begin(), idx.end(), [&](unsigned i) { tid[i] = omp_get_thread_num(); hits[i]++; });
    int mx = 0, bad = 0;
    for (size_t i = 0; i < idx.size(); i++) { mx = tid[i] > mx ? tid[i] : mx; bad += hits[i] != 1; }
    std::vector<unsigned> tiny(7, 1u);                       // below one chunk: runs inline
    unsigned s = 0;
    std::for_each(std::execution::par, tiny.begin(), tiny.end(), [&](unsigned v) { s += v; });
    std::printf("threads=%d bad=%d tiny=%u\n", mx + 1, bad, s);
    return 0;
}
''')
    exe = tmp_path / "probe"
    subprocess.run(["g++", "-std=c++20", "-O2", "-fopenmp", "-I", os.path.join(root, "oracle", "pstl_threads"), str(src), "-o", str(exe)],
                   check=True, timeout=300)
    env = dict(os.environ, OMP_NUM_THREADS="4")
    out = subprocess.run([str(exe)], stdout=subprocess.PIPE, text=True, env=env, timeout=60, check=True).stdout
    assert "bad=0 tiny=7" in out
    assert "threads=4" in out or (os.cpu_count() or 1) < 2, out


@needs_ref
def test_reference_build_reproduces_its_own_fingerprints(ob):
    """The unmodified reference, built by oracle/Makefile, on its own default scene (InitializeData(10000), gravity on,
    dt = 0.016667f): FNV-1a fingerprints of positions and of the sorted key sequence after 1 and 10 verbatim Update()
    calls.  They pin the BUILD (compiler flags, the abs() and `vector>` accommodations, no -march=native): any change that
    alters a single bit of the reference's output shows up here.  (SURVEY 8(c) quotes fingerprints of the same runs from
    a throw-away probe whose byte convention was not recorded; these are re-derived with ob.fnv1a64 over the raw fp32 /
    u32 arrays.)"""
    dt = float(np.float32(0.016667))
    r = ob.RefOracle(10000, spawn=True, gravity=1)
    r.update(dt)
    assert (ob.fnv1a64(r.positions()), ob.fnv1a64(r.sorted_lookup()[2])) == ("1c71b1884d4c581d", "76a76230a85562c0")
    for _ in range(9):
        r.update(dt)
    assert (ob.fnv1a64(r.positions()), ob.fnv1a64(r.sorted_lookup()[2])) == ("2c3ef7a42996212b", "a561be42e741eb31")


@needs_ref
def test_live_reference_getters_bounds(ob):
    r = ob.RefOracle(64, spawn=True)
    assert np.all(r.getter_probe(64) == 0) and np.all(r.getter_probe(2 ** 31) == 0)      # OOB -> zeros
    assert r.getter_probe(0)[6] > 0
    assert np.allclose(ob.RefOracle.kernels(0.1, 0.35), ob.PortOracle.kernels(0.1, 0.35), rtol=0, atol=0)


def test_openmp_port_is_thread_count_invariant(ob):
    g.load_package()
    from fluid_simulation_3d_b200 import scenes
    sc = scenes.small_dam_break(10)
    res = []
    for t in (1, 4):
        o = ob.PortOracle(sc["n"], threads=t, **sc["params"])
        o.set_state(sc["pos"], sc["vel"])
        o.step(scenes.DT, jacobi=True)
        res.append((o.positions(), o.velocities(), o.densities()))
    ob.PortOracle.lib().oracle_set_threads(1)
    for x, y in zip(*res):
        assert np.array_equal(bits(x), bits(y))


def test_colour_restatement_known_answers():
    """oracle/colors.py (getSpeedNormalzied + FluidSimCPU::updateColors) on values worked out by hand from
    fluidSimCPU.cc:100-125 and the gradient stops of fluidSimCPU.h:26-29."""
    import os, sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import colors
    f = np.float32
    v = np.array([[0, 0, 0], [0.3, 0.4, 0], [0.495, 0, 0], [0.99, 0, 0], [1.5, 0, 0], [3, 4, 0]], f)
    t = colors.speed_normalized(v)
    assert t[0] == 0 and t[1] == f(0.5) / f(1.5) and t[2] == f(0.33) and t[3] == f(0.66) and t[4] == 1 and t[5] == 1   # clamp at 1.5
    c = colors.speed_colors(v)
    assert np.array_equal(c[0], [0.0, 0.75, 1.0, 1.0])                    # at rest: Color1
    assert np.array_equal(c[2], [0.0, 1.0, 0.0, 1.0])                     # first breakpoint (inclusive, :113): Color2
    assert np.array_equal(c[3], [1.0, 1.0, 0.0, 1.0])                     # second breakpoint (:116): Color3
    assert np.array_equal(c[4], [1.0, 0.0, 0.0, 1.0]) and np.array_equal(c[5], c[4])   # clamped: Color4
    a = (t[1] - f(0.33)) / (f(0.66) - f(0.33))                            # second segment: Color2 -> Color3
    assert np.array_equal(c[1], np.array([a, (f(1) - a) + a, 0.0, (f(1) - a) + a], f))
    w = colors.speed_colors(np.array([[0.25, 0, 0]], f))[0]               # first segment: only green moves, 0.75 -> 1
    a0 = (f(0.25) / f(1.5)) / f(0.33)
    assert np.array_equal(w, np.array([0.0, (f(1) - a0) * f(0.75) + a0, (f(1) - a0), (f(1) - a0) + a0], f))
