"""CPU, world_size 2 and 3 over gloo: the slab decomposition protocol reproduces the single-domain step."""
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

import __graft_entry__ as g
import slab_model


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def scene_fast(n=2500, seed=5):
    rng = np.random.default_rng(seed)
    bound = (3.0, 2.5, 6.0)
    pos = ((rng.random((n, 3), dtype=np.float32) - 0.5) * np.array(bound, np.float32) * 0.98).astype(np.float32)
    vel = ((rng.random((n, 3), dtype=np.float32) - 0.5) * 6.0).astype(np.float32)
    return dict(pos=pos, vel=vel, n=n, params=dict(gravity=1, viscosity_strength=0.7, bound=bound))


@pytest.mark.parametrize("world", [2, 3])
def test_slab_protocol_matches_single_domain(world):
    ob = g.load_oracle()
    g.load_package()
    from fluid_simulation_3d_b200 import scenes
    sc = scene_fast()
    steps, dt = 3, scenes.DT
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=slab_model.run_rank, args=(r, world, port, sc, steps, dt, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    n = sc["n"]
    pos = np.zeros((n, 3), np.float32); vel = np.zeros((n, 3), np.float32); dens = np.zeros((n, 2), np.float32)
    nc = np.zeros(n, np.uint32); seen = np.zeros(n, np.int32); migrated = 0
    for ids, p_, v_, d_, c_, mig, layers in out:
        pos[ids], vel[ids], dens[ids], nc[ids] = p_, v_, d_, c_
        seen[ids] += 1
        migrated += mig
    assert np.all(seen == 1), "ownership is not a partition after migration"
    assert migrated > 0, "the scene was meant to exercise migration"
    ref = ob.PortOracle(n, **sc["params"])
    ref.set_state(sc["pos"], sc["vel"])
    for _ in range(steps):
        ref.step(dt, jacobi=True)
    assert np.array_equal(nc, ref.neighbour_counts()), "neighbour counts differ from the single-domain step"
    assert np.allclose(dens, ref.densities(), rtol=2e-5, atol=0)
    assert np.abs(pos - ref.positions()).max() < 2e-5
    assert np.abs(vel - ref.velocities()).max() < 2e-3


def test_slab_protocol_with_rebalancing_matches_single_domain():
    """planes start two layers off the quantiles and are re-balanced after every step (sph_slab_balance_layers cuts
    the all-reduced layer histogram, at most two layers per plane and call): whole layers change owner through the
    ordinary migration path and the result is still the single-domain step"""
    ob = g.load_oracle()
    g.load_package()
    from fluid_simulation_3d_b200 import scenes
    world = 3
    sc = scene_fast()
    steps, dt = 4, scenes.DT
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=slab_model.run_rank, args=(r, world, port, sc, steps, dt, q, 1, 2)) for r in range(world)]
    for p in procs:
        p.start()
    out = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    n = sc["n"]
    pos = np.zeros((n, 3), np.float32); vel = np.zeros((n, 3), np.float32); dens = np.zeros((n, 2), np.float32)
    nc = np.zeros(n, np.uint32); seen = np.zeros(n, np.int32)
    for ids, p_, v_, d_, c_, mig, (layers0, layers1, moves) in out:
        pos[ids], vel[ids], dens[ids], nc[ids] = p_, v_, d_, c_
        seen[ids] += 1
    assert np.all(seen == 1), "ownership is not a partition after re-balancing"
    assert moves > 0 and layers1 != layers0, "the skewed planes were meant to move: %r -> %r" % (layers0, layers1)
    counts = sorted(len(o[0]) for o in out)
    assert counts[-1] - counts[0] < 0.35 * n / world, "still unbalanced after re-balancing: %r" % (counts,)
    ref = ob.PortOracle(n, **sc["params"])
    ref.set_state(sc["pos"], sc["vel"])
    for _ in range(steps):
        ref.step(dt, jacobi=True)
    assert np.array_equal(nc, ref.neighbour_counts()), "neighbour counts differ from the single-domain step"
    assert np.allclose(dens, ref.densities(), rtol=2e-5, atol=0)
    assert np.abs(pos - ref.positions()).max() < 2e-5
    assert np.abs(vel - ref.velocities()).max() < 2e-3


def test_balance_layers_pure_function(pkg):
    """sph_slab_balance_layers: quantile cuts, the three-layer minimum, the per-call shift and row limits, determinism"""
    from fluid_simulation_3d_b200 import slab_driver as sm
    rng = np.random.default_rng(11)
    for trial in range(200):
        R = int(rng.integers(1, 9))
        gz = int(rng.integers(3 * R, 3 * R + 60))
        hist = rng.integers(0, 5000, gz).astype(np.uint32)
        if trial % 3 == 0:
            hist[rng.integers(0, gz, gz // 2)] = 0                  # air layers
        if trial % 7 == 0:
            hist[:] = 0                                              # nothing anywhere
        L = sm.balance_layers(pkg, hist, R)
        assert L[0] == 0 and L[-1] == gz and all(b - a >= 3 for a, b in zip(L, L[1:])), (L, gz, R)
        cum = np.concatenate([[0], np.cumsum(hist.astype(np.int64))])
        for k in range(1, R):
            # a plane the three-layer minimum did not touch sits on the layer boundary nearest to its quantile
            if _free(L, k):
                t = cum[-1] * k / R
                assert abs(cum[L[k]] - t) <= min(abs(cum[L[k] - 1] - t), abs(cum[L[k] + 1] - t))
        # constrained: start from an arbitrary valid partition, walk towards the quantiles
        cuts = np.sort(rng.choice(np.arange(1, gz // 3), R - 1, replace=False)) * 3 if R > 1 else np.zeros(0, np.int64)
        old = [0] + [int(c) for c in cuts] + [gz]
        old[-1] = gz
        if any(b - a < 3 for a, b in zip(old, old[1:])):
            continue
        shift = int(rng.integers(1, 6))
        budget = int(rng.integers(0, 3)) * 4000
        cur = old
        for _ in range(gz):
            new = sm.balance_layers(pkg, hist, R, cur, shift, budget)
            assert new[0] == 0 and new[-1] == gz and all(b - a >= 3 for a, b in zip(new, new[1:]))
            for k in range(1, R):
                assert abs(new[k] - cur[k]) <= min(shift, 3)
                assert cur[k - 1] <= new[k] <= cur[k + 1]                       # single-hop migration
                moved = abs(int(cum[new[k]]) - int(cum[cur[k]]))
                assert budget == 0 or moved <= budget
            if new == cur:
                break
            cur = new
        assert sm.balance_layers(pkg, hist, R, cur, shift, budget) == cur       # a fixed point stays one
        if budget == 0 and cum[-1] > 0:
            # without a row limit the walk ends where no single-layer move is worth making: it would bring the plane
            # nearer to its quantile by less than a quarter of the rows it hands over (the hysteresis that keeps a plane
            # from flipping between the two boundaries of the layer its quantile falls into)
            for k in range(1, R):
                t = cum[-1] * k / R
                for step in (-1, 1):
                    trial_L = list(cur); trial_L[k] += step
                    if all(b - a >= 3 for a, b in zip(trial_L, trial_L[1:])):
                        gain = abs(cum[cur[k]] - t) - abs(cum[trial_L[k]] - t)
                        moved = abs(int(cum[trial_L[k]]) - int(cum[cur[k]]))
                        assert gain < 0.25 * moved + 1e-6 or gain <= 0, (cur, k, step, gain, moved)
    # the flip the hysteresis is there for: the quantile in the middle of a layer, the histogram wobbling around it
    hist = np.array([100] * 10, np.uint32)                  # 2 ranks: quantile at 500 = the boundary 5 exactly
    assert sm.balance_layers(pkg, hist, 2, [0, 5, 10], 1) == [0, 5, 10]
    wob = np.array([100, 100, 100, 150, 110, 90, 90, 90, 90, 80], np.uint32)   # quantile 500 inside layer 4: boundary 4 at 450, 5 at 560
    assert sm.balance_layers(pkg, wob, 2, None) == [0, 4, 10]                  # a free cut takes the nearer boundary,
    assert sm.balance_layers(pkg, wob, 2, [0, 5, 10], 1) == [0, 5, 10]         # a plane already at 5 stays: 10 nearer for 110 rows moved
    assert sm.balance_layers(pkg, wob, 2, [0, 4, 10], 1) == [0, 4, 10]         # and so does one at 4
    far = hist.copy(); far[:3] = 400                        # a real imbalance moves it
    assert sm.balance_layers(pkg, far, 2, [0, 5, 10], 1) == [0, 4, 10]
    # errors: too few layers, a broken old partition
    with pytest.raises(ValueError):
        sm.balance_layers(pkg, np.ones(5, np.uint32), 2)
    with pytest.raises(ValueError):
        sm.balance_layers(pkg, np.ones(12, np.uint32), 2, [0, 2, 12], 1)
    # the numpy helper used for the first cut agrees on a smooth column
    z = np.linspace(-4.9, 4.9, 20000).astype(np.float32)
    gmin_z, gz = -17, 34
    hist = np.bincount(sm.layer_of(z, 0.35, gmin_z, gz), minlength=gz).astype(np.uint32)
    for world in (2, 4, 8):
        assert sm.balance_layers(pkg, hist, world) == sm.choose_layers(z, world, 0.35, gmin_z, gz)


def _free(L, k):
    """whether plane k sits strictly inside what the three-layer minimum allows (so the quantile alone placed it)"""
    return L[k - 1] + 3 < L[k] < L[k + 1] - 3


def test_partition_helpers():
    g.load_package()
    from fluid_simulation_3d_b200 import slab_driver as sm
    z = np.linspace(-4.9, 4.9, 10000).astype(np.float32)
    gmin_z, gz = -17, 34
    for world in (2, 4, 8):
        L = sm.choose_layers(z, world, 0.35, gmin_z, gz)
        assert L[0] == 0 and L[-1] == gz and all(b - a >= 3 for a, b in zip(L, L[1:]))
        own = sm.owner_of(z, L, 0.35, gmin_z, gz)
        cnt = np.bincount(own, minlength=world)
        assert cnt.min() > 0.6 * len(z) / world and cnt.max() < 1.4 * len(z) / world
        planes = sm.planes_from_layers(L, 0.35, gmin_z)
        back = [int(np.floor(np.float32(p) / np.float32(0.35))) - gmin_z for p in planes]
        assert back == L
    ids = sm.lattice_ids_for_rank(3, 2, 5, 1, 3)
    assert sorted(ids.tolist()) == sorted((iy * 3 + ix) * 5 + iz for iy in range(2) for ix in range(3) for iz in (1, 2))
