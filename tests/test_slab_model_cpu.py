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
