"""-m gpu: the C++ host layer above the C ABI -- the FluidSimBase backend (FluidSimB200), the
per-particle getters under FluidSimCPU::updateColors-style parallel access, and sub-stepping (8(f) rank 4) -- driven
through host_demo and compared with the same scene run through the C ABI from Python."""
import os
import subprocess

import numpy as np
import pytest

import __graft_entry__ as g

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEMO = os.path.join(ROOT, "fluid-simulation-3d_b200", "host", "host_demo")
DT = float(np.float32(0.016667))


def _demo(n, steps, what):
    r = subprocess.run([DEMO, str(n), str(steps), "0", what], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0, r.stdout
    line = [ln for ln in r.stdout.splitlines() if ln.startswith(what + " n=")][0]
    return dict(kv.split("=") for kv in line.split()[1:]), r.stdout


def test_fluidsimbase_backend_frames_match_c_abi(pkg):
    ob = g.load_oracle()
    n, steps = 5000, 3
    got, out = _demo(n, steps, "adapter")
    init = [ln for ln in out.splitlines() if ln.startswith("adapter init")][0]
    assert "count=%d" % n in init
    sim = pkg.FluidSimulation(n, gravity=1)
    sim.spawn_grid(n)
    assert "pos_fnv=" + ob.fnv1a64(sim.download("out_positions")) in init       # the spawn frame
    for _ in range(steps):                                                       # reset() restarted from the spawn
        sim.step(DT)
    assert got["frames"] == str(steps + 1)
    assert got["out_fnv"] == ob.fnv1a64(sim.download("out_positions"))
    assert got["col_fnv"] == ob.fnv1a64(sim.download("colors"))
    sim.close()


def test_getters_single_reads_and_parallel_bulk_mirror(pkg):
    ob = g.load_oracle()
    n, steps = 6000, 2
    got, _ = _demo(n, steps, "getters")
    sim = pkg.FluidSimulation(n, gravity=1)
    sim.spawn_grid(n)
    for _ in range(steps):
        sim.step(DT)
    assert float(got["single_vs_bulk_worst"]) == 0.0
    pos = sim.download("positions")
    assert got["pos_fnv"] == got["mirror_fnv"] == ob.fnv1a64(pos)
    assert got["dens_fnv"] == ob.fnv1a64(sim.download("densities"))
    ref = float(sim.download("speed_normalized").astype(np.float64).sum())
    assert abs(float(got["speed_sum"]) - ref) <= 1e-4 * max(1.0, ref)
    assert float(got["oob"]) == 0.0                                              # out of range -> 0 (physicsWorld.cc:180)
    sim.close()


def test_setters_never_throw_and_never_commit_a_rejected_value():
    """The reference's setters cannot fail and are driven by UI sliders (gameApp.cc:371-408).  Radius 0.01 in the default
    20-unit box is 2000^3 grid cells: the context steps on the reference's own table instead of refusing it; a radius of
    0 is refused, recorded in lastError(), and neither the host copy nor the device see it; the largest box of the
    slider (30 units) and the default radius again give a finite step."""
    r = subprocess.run([DEMO, "4000", "1", "0", "setters"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0, r.stdout
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("setters ")][0]
    got = dict(kv.split("=") for kv in line.split()[1:])
    assert got["r_small"] == "0.010" and float(got["rho_small"]) > 0.0
    assert got["r_after_bad"] == "0.010" and got["recorded"] == "1"
    assert got["r_final"] == "0.350" and got["bound"] == "30.0" and got["finite"] == "1"


def test_substepping_equals_explicit_smaller_steps(pkg):
    ob = g.load_oracle()
    n, steps = 4096, 2
    got, _ = _demo(n, steps, "substeps")
    assert got["sub"] == "4"
    sim = pkg.FluidSimulation(n, gravity=1)
    sim.spawn_grid(n)
    sub_dt = float(np.float32(0.016667) / np.float32(4))
    for _ in range(steps):
        sim.step_n(sub_dt, 4)
    assert got["pos_fnv"] == ob.fnv1a64(sim.download("positions"))
    sim.close()


def test_host_class_over_two_gpus_matches_one_gpu():
    """setDevices({0, 1}): the reference's class interface, slab-decomposed in ONE process (a host thread per GPU,
    NCCL between them, planes re-balanced every second Update) against the same class on one GPU"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    n, steps = 20000, 4
    r = subprocess.run([DEMO, str(n), str(steps), "0", "multi", "2"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0, r.stdout
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("multi n=")][0]
    got = dict(kv.split("=", 1) for kv in line.split()[1:])
    # fp32 summation order differs between the decompositions: 1e-5 of the box scale per step
    assert float(got["max_pos_diff"]) <= 1e-5 * 10.0 * steps, line
    assert float(got["max_out_diff"]) <= 1e-5 * 10.0 * steps, line
    assert float(got["max_rho_rel"]) <= 1e-5 * steps, line
    a, b = got["spawn_rho"].split("/")
    assert a == b and float(a) > 0.0, line                      # densities valid before the first Update, same lattice
    assert got["getters_ok"] == "1" and got["owned"] == str(n), line
    per = [int(x) for x in got["per_device"].strip("[]").split(",")]
    assert len(per) == 2 and min(per) > 0.3 * n, line
    assert float(got["density_ms"]) > 0.0, line
