"""-m gpu: slab mode.  One-rank slab mode runs on any box; the 2-rank check needs two GPUs and is
launched the way the driver launches bench.py (torch.distributed.run, one rank per GPU)."""
import os
import subprocess
import sys

import numpy as np
import pytest

import __graft_entry__ as g

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_single_rank_slab_equals_plain_step(pkg):
    """nranks = 1: a one-slab context (ids travel with the state, owned up/downloads) steps through the plain path."""
    from fluid_simulation_3d_b200 import scenes, slab_driver
    sc = scenes.small_dam_break(16)
    idb = slab_driver.SlabSimulation.make_id(pkg)
    slab = slab_driver.SlabSimulation(pkg, sc["n"] + 1024, 0, 1, 0, idb, **sc["params"])
    slab.set_layers([0, int(slab.dims[2])])
    slab.sim.set_neighbour_count_tap(True)
    slab.upload_owned(np.arange(sc["n"], dtype=np.uint32), sc["pos"], sc["vel"])
    plain = pkg.FluidSimulation(sc["n"], **sc["params"])
    plain.set_neighbour_count_tap(True)
    plain.upload_state(sc["pos"], sc["vel"])
    for _ in range(3):
        slab.step(scenes.DT)
        plain.step(scenes.DT)
    for f in ("positions", "velocities", "densities", "neighbour_count"):
        ids, a = slab.download_owned(f)
        full = np.zeros_like(a)
        full[ids] = a
        assert np.array_equal(full, plain.download(f)), f
    slab.close(); plain.close()


def test_single_rank_layer_histogram_and_rebalance_call(pkg):
    """sph_comm_rebalance on one rank: the histogram it reads off the step's table is the per-layer count of the
    predicted positions, no plane can move, and it refuses to run before a step"""
    from fluid_simulation_3d_b200 import scenes, slab_driver
    sc = scenes.small_dam_break(18, seed=4)
    idb = slab_driver.SlabSimulation.make_id(pkg)
    slab = slab_driver.SlabSimulation(pkg, sc["n"] + 1024, 0, 1, 0, idb, **sc["params"])
    gz = int(slab.dims[2])
    slab.set_layers([0, gz])
    slab.upload_owned(np.arange(sc["n"], dtype=np.uint32), sc["pos"], sc["vel"])
    with pytest.raises(pkg.SphError):
        slab.rebalance(1)                                   # no table yet
    for _ in range(2):
        slab.step(scenes.DT)
    layers, hist, changed = slab.rebalance(2)
    assert layers == [0, gz] and not changed and slab.get_layers() == [0, gz]
    _, pred = slab.download_owned("predicted")
    want = np.bincount(slab_driver.layer_of(pred[:, 2], slab.r, int(slab.origin[2]), gz), minlength=gz)
    assert np.array_equal(hist, want.astype(np.uint32))
    # the cut a 2- or 4-rank run would make from this histogram
    for world in (2, 4):
        L = slab_driver.balance_layers(pkg, hist, world)
        own = np.diff(np.concatenate([[0], np.cumsum(want)])[L])
        assert own.sum() == sc["n"] and own.max() - own.min() <= 2 * want.max()
    slab.step(scenes.DT)                                    # and the step after a (no-op) re-balance still runs
    slab.upload_owned(np.arange(sc["n"], dtype=np.uint32), sc["pos"], sc["vel"])
    with pytest.raises(pkg.SphError):
        slab.rebalance(1)                                   # a new state invalidates the table
    slab.close()


def _slab_frames(pkg, slab, frames, dt, pipelined):
    """every frame: upload the rank's owned state, step, read OutPositions + ids back; returns them sorted by id"""
    import ctypes as C
    import torch
    L, h = slab.sim.L, slab.sim.h
    cap = slab.sim.capacity
    pins = [(torch.from_numpy(i.astype(np.int32)).pin_memory(), torch.from_numpy(p).pin_memory(), torch.from_numpy(v).pin_memory())
            for i, p, v in frames]
    outs = [(torch.empty((cap, 4), dtype=torch.float32).pin_memory(), torch.empty(cap, dtype=torch.int32).pin_memory()) for _ in range(2)]
    got = []

    def collect(k, n):
        o, i = outs[k & 1]
        ids = i.numpy()[:n].astype(np.int64)
        order = np.argsort(ids)
        got.append((ids[order].copy(), o.numpy()[:n][order].copy()))

    def begin(k):
        i, p, v = pins[k]
        slab.sim._check(L.sph_upload_owned_begin(h, i.numel(), C.c_void_p(i.data_ptr()), C.c_void_p(p.data_ptr()), C.c_void_p(v.data_ptr())))

    counts = []
    if not pipelined:
        for k, (i, p, v) in enumerate(pins):
            slab.sim._check(L.sph_upload_owned(h, i.numel(), C.c_void_p(i.data_ptr()), C.c_void_p(p.data_ptr()), C.c_void_p(v.data_ptr())))
            slab.step(dt)
            cnt = C.c_uint32(0)
            o, ii = outs[k & 1]
            slab.sim._check(L.sph_download_owned(h, pkg.FIELDS["out_positions"], C.c_void_p(ii.data_ptr()), C.c_void_p(o.data_ptr()),
                                                 cap * 16, C.byref(cnt)))
            collect(k, cnt.value)
        return got
    begin(0)
    for k in range(len(pins)):
        slab.sim._check(L.sph_upload_state_commit(h))
        if k + 1 < len(pins):
            begin(k + 1)
        slab.step(dt)
        if k:
            slab.sim._check(L.sph_download_wait(h))
            collect(k - 1, counts[-1])
        cnt = C.c_uint32(0)
        o, ii = outs[k & 1]
        slab.sim._check(L.sph_download_owned_begin(h, pkg.FIELDS["out_positions"], C.c_void_p(ii.data_ptr()), C.c_void_p(o.data_ptr()),
                                                   cap * 16, C.byref(cnt)))
        counts.append(cnt.value)
    slab.sim._check(L.sph_download_wait(h))
    collect(len(pins) - 1, counts[-1])
    return got


def test_single_rank_slab_pipelined_transfers(pkg):
    """the slab-mode pipelined calls (sph_upload_owned_begin, sph_download_owned_begin) against the blocking ones"""
    from fluid_simulation_3d_b200 import scenes, slab_driver
    frames = []
    for k in range(4):
        sc = scenes.small_dam_break(14, seed=21 + k)
        frames.append((np.arange(sc["n"], dtype=np.uint32), np.ascontiguousarray(sc["pos"]), np.ascontiguousarray(sc["vel"])))
    res = []
    for pipelined in (False, True):
        idb = slab_driver.SlabSimulation.make_id(pkg)
        slab = slab_driver.SlabSimulation(pkg, sc["n"] + 1024, 0, 1, 0, idb, **sc["params"])
        slab.set_layers([0, int(slab.dims[2])])
        res.append(_slab_frames(pkg, slab, frames, scenes.DT, pipelined))
        slab.close()
    for k, ((ia, a), (ib, b)) in enumerate(zip(*res)):
        assert np.array_equal(ia, ib) and ia.size == sc["n"], "frame %d ids" % k
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), "frame %d" % k
    # and the state the plain (non-slab) step produces from the last frame's input is the same
    plain = pkg.FluidSimulation(sc["n"], **sc["params"])
    plain.upload_state(frames[-1][1], frames[-1][2])
    plain.step(scenes.DT)
    assert np.array_equal(plain.download("out_positions").view(np.uint32), res[1][-1][1].view(np.uint32))
    plain.close()


def test_two_rank_slab_matches_single_gpu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run: gpurun --gpus 2 -- python -m torch.distributed.run ... tests/mgpu_check.py)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29517", os.path.join(ROOT, "tests", "mgpu_check.py")]
    r = subprocess.run(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0 and "MGPU_CHECK_OK" in r.stdout, r.stdout[-3000:]
