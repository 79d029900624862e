"""Multi-rank check of the slab-decomposed step against the single-GPU step (run under torchrun):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py

Every rank steps its slab; rank 0 also runs the whole scene on its own GPU and compares, gathered by
particle id: neighbour counts and hashes bit-exact, floats within 1e-5 of the stage scale per step.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402


def gather_by_id(dist, n_total, ids, arr):
    objs = [None] * dist.get_world_size()
    dist.all_gather_object(objs, (ids, arr))
    out = np.zeros((n_total,) + arr.shape[1:], arr.dtype)
    seen = np.zeros(n_total, np.int32)
    for i, a in objs:
        out[i] = a
        seen[i] += 1
    assert np.all(seen == 1), "ownership is not a partition: %d missing, %d duplicated" % ((seen == 0).sum(), (seen > 1).sum())
    return out


def run_scene(pkg, slabmod, scenes, dist, torch, rank, world, dev, sc, steps, label, skew=0, rebalance=0):
    dt = scenes.DT
    n = sc["n"]
    idb = slabmod.broadcast_id(pkg, dist, torch, rank, dev)
    slab = slabmod.SlabSimulation(pkg, n + 4096, rank, world, dev, idb, **sc["params"])
    gmin_z, gz = int(slab.origin[2]), int(slab.dims[2])
    pred0 = sc["pos"] + sc["vel"] * np.float32(1.0 / 120.0)
    layers = slabmod.choose_layers(pred0[:, 2], world, slab.r, gmin_z, gz)
    if skew:        # deliberately unbalanced start (every inner plane `skew` layers up); sph_comm_rebalance has to walk it back
        layers = [layers[0]] + [min(l + skew, gz - 3 * (world - k)) for k, l in enumerate(layers[1:-1], 1)] + [layers[-1]]
    layers0 = list(layers)
    slab.set_layers(layers)
    assert slab.get_layers() == layers
    own = slabmod.owner_of(sc["pos"][:, 2], layers, slab.r, gmin_z, gz) == rank
    ids = np.nonzero(own)[0].astype(np.uint32)
    owned0 = int(own.sum())
    slab.sim.set_neighbour_count_tap(True)
    slab.upload_owned(ids, sc["pos"][own], sc["vel"][own])
    single = None
    if rank == 0:
        single = pkg.FluidSimulation(n, device=dev, **sc["params"])
        single.set_neighbour_count_tap(True)
        single.upload_state(sc["pos"], sc["vel"])
    worst = {}
    migrated = 0
    for s in range(steps):
        slab.step(dt)
        st = slab.stats()
        migrated += st["migrated_lo"] + st["migrated_hi"]
        fields = {}
        for f in ("neighbour_count", "hash", "densities", "vel_after_pressure", "vel_after_viscosity", "positions", "velocities"):
            i, a = slab.download_owned(f)
            fields[f] = gather_by_id(dist, n, i, a)
        if rank == 0:
            single.step(dt)
            assert np.array_equal(fields["hash"], single.download("hash")), "%s step %d: hash" % (label, s)
            nc = single.download("neighbour_count")
            assert np.array_equal(fields["neighbour_count"], nc), "%s step %d: %d neighbour counts differ" % (
                label, s, int((fields["neighbour_count"] != nc).sum()))
            for f, scale in (("densities", 0.0), ("vel_after_pressure", 1.0), ("vel_after_viscosity", 1.0),
                             ("positions", 1.0), ("velocities", 1.0)):
                ref = single.download(f).astype(np.float64)
                got = fields[f].astype(np.float64)
                tol = 1e-5 * (s + 1) * np.maximum(np.abs(ref), max(scale, 1e-30) * np.abs(ref).max())
                err = np.abs(got - ref)
                assert np.all(err <= tol), "%s step %d: %s worst %g (tol %g)" % (label, s, f, err.max(), tol[np.unravel_index(err.argmax(), err.shape)])
                worst[f] = max(worst.get(f, 0.0), float((err / np.maximum(np.abs(ref), 1e-30 + scale * np.abs(ref).max())).max()))
        if rebalance:
            # COLLECTIVE: planes walk towards the particle-count quantiles, `rebalance` layers per call at most; the
            # histogram it returns is the global one (sums to n) and equal on every rank
            new_layers, hist, changed = slab.rebalance(rebalance)
            assert int(hist.sum()) == n, "%s step %d: layer histogram sums to %d, not %d" % (label, s, int(hist.sum()), n)
            assert all(abs(a - b) <= rebalance for a, b in zip(new_layers, layers)) and changed == (new_layers != layers)
            assert slab.get_layers() == new_layers
            layers = new_layers
    tot = torch.tensor([migrated], device="cuda:%d" % dev)
    dist.all_reduce(tot)
    if rebalance:
        cnt = torch.zeros(world, dtype=torch.int64, device="cuda:%d" % dev)
        cnt[rank] = slab.stats()["owned"]
        dist.all_reduce(cnt)
        c0 = torch.zeros(world, dtype=torch.int64, device="cuda:%d" % dev)
        c0[rank] = owned0
        dist.all_reduce(c0)
        spread0, spread1 = int(c0.max() - c0.min()), int(cnt.max() - cnt.min())
        assert layers != layers0 and spread1 < spread0, "%s: re-balancing did not help: layers %r -> %r, owned %r -> %r" % (
            label, layers0, layers, c0.tolist(), cnt.tolist())
        if rank == 0:
            print("%s: planes %r -> %r, owned per rank %r -> %r" % (label, layers0, layers, c0.tolist(), cnt.tolist()), flush=True)
    if rank == 0:
        print("%s: %d particles, %d ranks, layers %s, %d steps OK; migrations %d; worst rel err %s" % (
            label, n, world, layers, steps, int(tot.item()), {k: "%.1e" % v for k, v in worst.items()}), flush=True)
        single.close()
    slab.close()
    dist.barrier()


def main():
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); dev = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=torch.device("cuda", dev))
    pkg = g.load_package()
    from fluid_simulation_3d_b200 import scenes, slab_driver
    # (1) dam-break block at rest: ghosts, no migration at first
    run_scene(pkg, slab_driver, scenes, dist, torch, rank, world, dev, scenes.small_dam_break(24), 4, "dam_break_24")
    # (2) fast random particles across the whole box: migration every step, wall hits
    rng = np.random.default_rng(5)
    n = 20000
    bound = (5.0, 4.0, 9.0)
    pos = ((rng.random((n, 3), dtype=np.float32) - 0.5) * np.array(bound, np.float32) * 0.98).astype(np.float32)
    vel = ((rng.random((n, 3), dtype=np.float32) - 0.5) * 6.0).astype(np.float32)
    sc = dict(pos=pos, vel=vel, n=n, params=dict(gravity=1, viscosity_strength=0.7, bound=bound))
    run_scene(pkg, slab_driver, scenes, dist, torch, rank, world, dev, sc, 6, "fast_random_20k")
    # (3) dense column
    run_scene(pkg, slab_driver, scenes, dist, torch, rank, world, dev, scenes.small_column(12, 30, 40), 3, "column_12x30x40")
    # (3b) re-balancing: planes start three layers off the quantiles and are moved back (two layers per call at most)
    # after every step, so whole layers change owner through the migration path -- results still equal the single GPU
    run_scene(pkg, slab_driver, scenes, dist, torch, rank, world, dev, scenes.small_dam_break(24, seed=7), 4, "rebalance_dam_break_24",
              skew=3, rebalance=2)
    run_scene(pkg, slab_driver, scenes, dist, torch, rank, world, dev, sc, 5, "rebalance_fast_random_20k", skew=4, rebalance=1)
    # (4) blow-up: particles fast enough to cross more than a whole slab in one step.  Values are not compared
    # (such a particle takes one ballistic step while it is handed on); the protocol must neither fail its
    # halo cross-check nor lose or duplicate particles.
    rng = np.random.default_rng(9)
    n = 12000
    bound = (4.0, 4.0, 9.0)
    pos = ((rng.random((n, 3), dtype=np.float32) - 0.5) * np.array(bound, np.float32) * 0.98).astype(np.float32)
    vel = ((rng.random((n, 3), dtype=np.float32) - 0.5) * 1200.0).astype(np.float32)
    sc = dict(pos=pos, vel=vel, n=n, params=dict(gravity=1, viscosity_strength=0.2, bound=bound))
    idb = slab_driver.broadcast_id(pkg, dist, torch, rank, dev)
    slab = slab_driver.SlabSimulation(pkg, n + 4096, rank, world, dev, idb, **sc["params"])
    gmin_z, gz = int(slab.origin[2]), int(slab.dims[2])
    layers = slab_driver.choose_layers(pos[:, 2], world, slab.r, gmin_z, gz)
    slab.set_layers(layers)
    own = slab_driver.owner_of(pos[:, 2], layers, slab.r, gmin_z, gz) == rank
    slab.upload_owned(np.nonzero(own)[0].astype(np.uint32), pos[own], vel[own])
    for s in range(6):
        slab.step(scenes.DT)
        i, a = slab.download_owned("positions")
        full = gather_by_id(dist, n, i, a)          # asserts every id is owned exactly once
        assert np.all(np.isfinite(full))
    slab.close()
    dist.barrier()
    if rank == 0:
        print("blow_up_12k: 6 steps, ids conserved", flush=True)
    # (5) the pipelined transfer calls in slab mode: same frames through the blocking and the pipelined calls
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_multi_gpu import _slab_frames
    sc = scenes.small_dam_break(20, seed=3)
    res = []
    for pipelined in (False, True):
        idb = slab_driver.broadcast_id(pkg, dist, torch, rank, dev)
        slab = slab_driver.SlabSimulation(pkg, sc["n"] + 4096, rank, world, dev, idb, **sc["params"])
        gmin_z, gz = int(slab.origin[2]), int(slab.dims[2])
        layers = slab_driver.choose_layers(sc["pos"][:, 2], world, slab.r, gmin_z, gz)
        slab.set_layers(layers)
        frames = []
        for k in range(3):
            f = scenes.small_dam_break(20, seed=3 + k)
            own = slab_driver.owner_of(f["pos"][:, 2], layers, slab.r, gmin_z, gz) == rank
            frames.append((np.nonzero(own)[0].astype(np.uint32), np.ascontiguousarray(f["pos"][own]), np.ascontiguousarray(f["vel"][own])))
        res.append(_slab_frames(pkg, slab, frames, scenes.DT, pipelined))
        slab.close()
        dist.barrier()
    for k, ((ia, a), (ib, b)) in enumerate(zip(*res)):
        assert np.array_equal(ia, ib), "pipelined frame %d: ids differ on rank %d" % (k, rank)
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), "pipelined frame %d differs on rank %d" % (k, rank)
    dist.barrier()
    if rank == 0:
        print("pipelined_owned_transfers: 3 frames bit-identical to the blocking calls", flush=True)
    # (6) a capacity error comes back as an error on EVERY rank, never as a hang: the slabs hold their own rows but have
    # no room for the ghost layers, so both ends of the link refuse it (the verdict is computed from the same count
    # messages on either side), complete the step's communication pattern and return the error; it is sticky until the
    # next upload
    sc = scenes.small_dam_break(24)
    idb = slab_driver.broadcast_id(pkg, dist, torch, rank, dev)
    probe = slab_driver.SlabSimulation(pkg, sc["n"], rank, world, dev, idb, **sc["params"])
    gmin_z, gz = int(probe.origin[2]), int(probe.dims[2])
    layers = slab_driver.choose_layers(sc["pos"][:, 2], world, probe.r, gmin_z, gz)
    own = slab_driver.owner_of(sc["pos"][:, 2], layers, probe.r, gmin_z, gz) == rank
    probe.close()
    dist.barrier()
    idb = slab_driver.broadcast_id(pkg, dist, torch, rank, dev)
    slab = slab_driver.SlabSimulation(pkg, int(own.sum()) + 64, rank, world, dev, idb, **sc["params"])
    slab.set_layers(layers)
    slab.upload_owned(np.nonzero(own)[0].astype(np.uint32), sc["pos"][own], sc["vel"][own])
    errors = 0
    for _ in range(3):
        try:
            slab.step(scenes.DT)
        except pkg.SphError as e:
            errors += 1
            assert "capacity" in str(e) or "room" in str(e) or "neighbour" in str(e), str(e)
    assert errors == 3, "rank %d: %d of 3 steps reported the capacity error" % (rank, errors)
    slab.close()
    dist.barrier()
    if rank == 0:
        print("capacity_error: returned by every rank on every step, no rank left waiting", flush=True)
    dist.destroy_process_group()
    if rank == 0:
        print("MGPU_CHECK_OK")


if __name__ == "__main__":
    main()
