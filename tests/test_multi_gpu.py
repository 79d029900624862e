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
    """nranks = 1: no neighbours, but the whole slab pipeline (classify, pack, ghost-aware sort) runs."""
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


def test_two_rank_slab_matches_single_gpu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run: gpurun --gpus 2 -- python -m torch.distributed.run ... tests/mgpu_check.py)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29517", os.path.join(ROOT, "tests", "mgpu_check.py")]
    r = subprocess.run(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0 and "MGPU_CHECK_OK" in r.stdout, r.stdout[-3000:]
