"""Shared parity machinery: run one step on the GPU (through the C ABI) and on the oracle, compare.

Tolerances (BASELINE.json north_star / SURVEY.md 8(a)): integer outputs bit-exact; fp32 outputs
|gpu - ref| <= 1e-5 * max(|ref|, S) with S the stage's cancellation-free term scale, single step.
"""
import numpy as np

import __graft_entry__ as g

REL = 1e-5


def coincident_scene(seed=17):
    """Particles that share an exact position: pairs and triples inside the fluid (equal velocities, so their
    predicted positions coincide too), a few stacked in a wall corner the way the box clamp leaves them, and a small
    cloud around them.  dist == 0 takes the reference's direction fallback (0, 1, 0) in the pressure pass
    (physicsWorld.cc:414) and the full kernel weights everywhere else."""
    rng = np.random.default_rng(seed)
    bound = (3.0, 3.0, 3.0)
    cloud = ((rng.random((90, 3)) - 0.5) * 1.4).astype(np.float32)
    cvel = ((rng.random((90, 3)) - 0.5) * 2.0).astype(np.float32)
    pos, vel = [cloud], [cvel]
    for k in range(6):                                   # pairs and triples at the position (and velocity) of a cloud particle
        reps = 1 + k % 2 + 1
        pos.append(np.repeat(cloud[k:k + 1], reps, axis=0)); vel.append(np.repeat(cvel[k:k + 1], reps, axis=0))
    corner = np.array([[-1.5, -1.5, 1.5]], np.float32)   # exactly on three walls: what the clamp writes (:88-106)
    pos.append(np.repeat(corner, 4, axis=0)); vel.append(np.zeros((4, 3), np.float32))
    pos.append(np.array([[0.4, -1.5, 0.2]] * 3, np.float32)); vel.append(np.array([[0.3, 0.0, -0.2]] * 3, np.float32))
    pos = np.ascontiguousarray(np.concatenate(pos)); vel = np.ascontiguousarray(np.concatenate(vel))
    return dict(pos=pos, vel=vel, n=len(pos), params=dict(gravity=1, viscosity_strength=0.6, bound=bound))


def oracle_pair(n, params):
    """(value oracle, scale oracle).  Values come from the UNMODIFIED reference when oracle/_ref is
    built, else from the C restatement; the term scales always come from the restatement."""
    ob = g.load_oracle()
    port = ob.PortOracle(n, **params)
    ref = ob.RefOracle(n, **params) if ob.have_ref() else port
    return ref, port


def oracle_step(scene, dt, jacobi=True):
    ref, port = oracle_pair(scene["n"], scene["params"])
    for o in ({id(ref): ref, id(port): port}).values():
        o.set_state(scene["pos"], scene["vel"])
        o.step(dt, jacobi=jacobi)
    ps, vs = port.force_scales(dt)
    return ref, ps, vs


def canonical_order(sorted_idx, sorted_key):
    """Sort each equal-key run by ascending particle index (SURVEY App.A Q14)."""
    order = np.lexsort((sorted_idx, sorted_key))
    return sorted_idx[order]


def grid_keys(pred, cells, dims, origin, xsub, r):
    """GRID key of every particle: y, z = the reference's cells; x = the cell subdivided xsub times,
    floor((pred.x / r) * xsub) in fp32 like the device (sph_device.cuh: grid_cell)."""
    d = dims.astype(np.int64)
    tx = pred[:, 0].astype(np.float32) / np.float32(r)
    xf = np.floor(tx * np.float32(xsub)).astype(np.int64)
    gx = np.clip(xf - int(origin[0]) * xsub, 0, d[0] - 1)
    gy = np.clip(cells[:, 1].astype(np.int64) - int(origin[1]), 0, d[1] - 1)
    gz = np.clip(cells[:, 2].astype(np.int64) - int(origin[2]), 0, d[2] - 1)
    return ((gz * d[1] + gy) * d[0] + gx).astype(np.uint32)


def assert_close(name, got, ref, scale, rel=REL):
    got = np.asarray(got, np.float64)
    ref = np.asarray(ref, np.float64)
    tol = rel * np.maximum(np.abs(ref), np.asarray(scale, np.float64))
    err = np.abs(got - ref)
    bad = err > tol
    assert np.all(np.isfinite(got)), name + ": non-finite values"
    if bad.any():
        i = int(np.argmax(err / np.maximum(tol, 1e-300)))
        raise AssertionError("%s: %d/%d outside %g tolerance; worst got=%r ref=%r tol=%g" %
                             (name, int(bad.sum()), bad.size, rel, got.flat[i], ref.flat[i], tol.flat[i]))
    return float((err / np.maximum(tol, 1e-300)).max()) * rel


def check_step(pkg, scene, mode, dt, report=None):
    """One GPU step vs one oracle step on the same state.  Returns a dict of worst relative errors."""
    n = scene["n"]
    sim = pkg.FluidSimulation(n, device=0, table_mode=mode, **scene["params"])
    try:
        sim.set_neighbour_count_tap(True)
        sim.upload_state(scene["pos"], scene["vel"])
        sim.step(dt)
        ref, ps, vs = oracle_step(scene, dt)
        out = {}
        # ---- integers: bit-exact
        pred = sim.download("predicted")
        assert np.array_equal(pred.view(np.uint32), ref.predicted().view(np.uint32)), "predicted positions not bit-exact"
        h, k, cells = ref.hash_key()
        assert np.array_equal(sim.download("hash"), h), "hash"
        assert np.array_equal(sim.download("key"), k), "key"
        nc = sim.download("neighbour_count")
        nc_ref = ref.neighbour_counts()
        assert np.array_equal(nc, nc_ref), "neighbour counts: %d differ" % int((nc != nc_ref).sum())
        out["mean_neighbours"] = float(nc.mean())
        s_idx = sim.download_table("sorted_index")
        s_key = sim.download_table("sorted_key")
        assert np.array_equal(np.sort(s_idx), np.arange(n, dtype=np.uint32)), "sorted index is not a permutation"
        assert np.all(s_key[1:] >= s_key[:-1]), "sorted keys not sorted"
        table = sim.download_table("start_indices")
        if mode == pkg.TABLE_REFERENCE_HASH:
            r_idx, r_hash, r_key = ref.sorted_lookup()
            assert np.array_equal(s_key, r_key), "sorted key sequence"
            assert np.array_equal(table, ref.start_indices()), "start table"
            assert np.array_equal(canonical_order(s_idx, s_key), canonical_order(r_idx, r_key)), "sorted order (canonical ties)"
            assert np.array_equal(s_key, k[s_idx]), "sorted key vs per-particle key"
        else:
            dims, origin = sim.grid()
            gk = grid_keys(pred, cells, dims, origin, sim.grid_x_subdivision(), float(sim.get_params().interaction_radius))
            assert np.array_equal(s_key, gk[s_idx]), "grid key of sorted rows"
            assert np.array_equal(canonical_order(s_idx, s_key), s_idx), "rows of a cell are not in ascending particle index order"
            ncell = int(dims[0]) * int(dims[1]) * int(dims[2])
            assert table.size == ncell + 1 and table[0] == 0 and table[-1] == n
            assert np.array_equal(table, np.searchsorted(s_key, np.arange(ncell + 1, dtype=np.uint64)).astype(np.uint32)), "prefix table"
        # ---- floats
        dref = ref.densities()
        out["density"] = assert_close("density", sim.download("densities"), dref, 0.0)
        vp_ref = ref.vel_after_pressure()
        out["vel_after_pressure"] = assert_close("vel_after_pressure", sim.download("vel_after_pressure"), vp_ref, ps[:, None])
        vv_ref = ref.vel_after_viscosity()
        out["vel_after_viscosity"] = assert_close("vel_after_viscosity", sim.download("vel_after_viscosity"), vv_ref,
                                                  (ps + vs)[:, None])
        pos_ref, vel_ref = ref.positions(), ref.velocities()
        speed = np.abs(vv_ref).max(axis=1, keepdims=True)
        out["positions"] = assert_close("positions", sim.download("positions"), pos_ref, speed * dt + (ps + vs)[:, None] * dt)
        # a wall hit flips the velocity sign; both sides must agree on who hit
        out["velocities"] = assert_close("velocities", sim.download("velocities"), vel_ref, (ps + vs)[:, None])
        o4 = sim.download("out_positions")
        assert np.array_equal(o4[:, :3].view(np.uint32), sim.download("positions").view(np.uint32))
        assert np.all(o4[:, 3] == np.float32(0.34))
        if report is not None:
            report.update(out)
        return out
    finally:
        sim.close()


def golden_params(gold):
    p = gold["params"]
    return dict(interaction_radius=p[0], target_density=p[1], pressure_multiplier=p[2], near_pressure_multiplier=p[3],
                viscosity_strength=p[4], gravity_scale=p[5], gravity=int(p[6]), bound=tuple(p[7:10]))


def golden_scales(gold):
    """cancellation-free term scales of the fixture's step (from the restatement, which test_oracle.py pins to the
    fixture bit for bit)"""
    ob = g.load_oracle()
    n, dt = gold["pos0"].shape[0], float(gold["dt"])
    port = ob.PortOracle(n, **golden_params(gold))
    port.set_state(gold["pos0"], gold["vel0"])
    port.step(dt, jacobi=True)
    return port.force_scales(dt)


def compare_to_golden(get, gold, reference_table):
    """One step's outputs against a committed fixture of the UNMODIFIED reference (tests/golden/*.npz, made by
    make_golden.py): integers bit-exact, floats within REL of the stage scale.  `get(name)` returns the implementation's
    array: predicted, hash, key, neighbour_count, densities, vel_after_pressure, vel_after_viscosity, positions,
    velocities, out_positions and -- with the reference's table (`reference_table`) -- sorted_key, sorted_index,
    start_indices."""
    dt = float(gold["dt"])
    ps, vs = golden_scales(gold)
    assert np.array_equal(np.asarray(get("predicted")).view(np.uint32), gold["pred"].view(np.uint32)), "predicted positions not bit-exact"
    assert np.array_equal(get("hash"), gold["hash"]), "hash"
    assert np.array_equal(get("key"), gold["key"]), "key"
    assert np.array_equal(get("neighbour_count"), gold["ncount"]), "neighbour counts"
    if reference_table:
        s_key, s_idx = get("sorted_key"), get("sorted_index")
        assert np.array_equal(s_key, gold["sorted_key"]), "sorted key sequence"
        assert np.array_equal(get("start_indices"), gold["start"]), "start table"
        assert np.array_equal(canonical_order(s_idx, s_key), canonical_order(gold["sorted_idx"], gold["sorted_key"])), "sorted order (canonical ties)"
    out = {}
    out["density"] = assert_close("density", get("densities"), gold["dens"], 0.0)
    out["vel_after_pressure"] = assert_close("vel_after_pressure", get("vel_after_pressure"), gold["vel_press"], ps[:, None])
    out["vel_after_viscosity"] = assert_close("vel_after_viscosity", get("vel_after_viscosity"), gold["vel_visc"], (ps + vs)[:, None])
    speed = np.abs(gold["vel_visc"]).max(axis=1, keepdims=True)
    out["positions"] = assert_close("positions", get("positions"), gold["pos1"], speed * dt + (ps + vs)[:, None] * dt)
    out["velocities"] = assert_close("velocities", get("velocities"), gold["vel1"], (ps + vs)[:, None])
    o4 = np.asarray(get("out_positions"))
    out["out_positions"] = assert_close("out_positions", o4[:, :3], gold["out1"][:, :3], speed * dt + (ps + vs)[:, None] * dt)
    assert np.all(o4[:, 3] == np.float32(0.34)) and np.all(gold["out1"][:, 3] == np.float32(0.34))
    return out


def check_step_port(pkg, ob, scene, dt, list_capacity=None, label=""):
    """One GPU step against the OpenMP restatement on every host core, for sizes the serial reference build cannot step
    in test time (8 M particles).  The restatement is pinned bit for bit to the unmodified reference by tests/test_oracle.py.
    Integers bit-exact, floats within 1e-5 of the stage scale.  Returns the mean neighbour count."""
    import os
    n = scene["n"]
    threads = max(1, os.cpu_count() or 1)
    o = ob.PortOracle(n, threads=threads, **scene["params"])
    o.set_state(scene["pos"], scene["vel"])
    o.step(dt, jacobi=True)
    ps, vs = o.force_scales(dt)
    ob.PortOracle.lib().oracle_set_threads(1)
    sim = pkg.FluidSimulation(n, **scene["params"])
    try:
        if list_capacity:
            sim.set_neighbour_list_capacity(list_capacity)
        sim.set_neighbour_count_tap(True)
        sim.upload_state(scene["pos"], scene["vel"])
        sim.step(dt)
        h, k, _ = o.hash_key()
        assert np.array_equal(sim.download("predicted").view(np.uint32), o.predicted().view(np.uint32)), label + ": predicted"
        assert np.array_equal(sim.download("hash"), h) and np.array_equal(sim.download("key"), k), label + ": hash / key"
        nc, nc_ref = sim.download("neighbour_count"), o.neighbour_counts()
        assert np.array_equal(nc, nc_ref), "%s: %d neighbour counts differ" % (label, int((nc != nc_ref).sum()))
        assert_close(label + " density", sim.download("densities"), o.densities(), 0.0)
        assert_close(label + " vel_after_pressure", sim.download("vel_after_pressure"), o.vel_after_pressure(), ps[:, None])
        assert_close(label + " vel_after_viscosity", sim.download("vel_after_viscosity"), o.vel_after_viscosity(), (ps + vs)[:, None])
        speed = np.abs(o.vel_after_viscosity()).max(axis=1, keepdims=True)
        assert_close(label + " positions", sim.download("positions"), o.positions(), speed * dt + (ps + vs)[:, None] * dt)
        assert_close(label + " velocities", sim.download("velocities"), o.velocities(), (ps + vs)[:, None])
        return float(nc.mean())
    finally:
        sim.close()
        o.close()
