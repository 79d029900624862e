"""Scratch timing of the step stages (not the contract bench): python tools/quick_bench.py [config] [steps]"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as g
pkg = g.load_package()
from fluid_simulation_3d_b200 import scenes

name = sys.argv[1] if len(sys.argv) > 1 else "C2_dambreak_1M"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
sc = scenes.config(name)
for mode, mname in ((pkg.TABLE_GRID, "grid"), (pkg.TABLE_REFERENCE_HASH, "refhash")):
    sim = pkg.FluidSimulation(sc["n"], table_mode=mode, **sc["params"])
    sim.upload_state(sc["pos"], sc["vel"])
    for _ in range(5):
        sim.step(scenes.DT)
    sim.synchronize()
    acc = np.zeros(6)
    t0 = time.time()
    for _ in range(steps):
        sim.step(scenes.DT)
        acc += sim.timings()
    sim.synchronize()
    wall = (time.time() - t0) / steps
    acc /= steps
    print("%s %s n=%d stage_ms predict=%.3f spatial=%.3f density=%.3f pressure=%.3f viscosity=%.3f integrate=%.3f sum=%.3f wall=%.3f Mupd/s=%.1f"
          % (name, mname, sc["n"], *acc, acc.sum(), wall * 1e3, sc["n"] / acc.sum() / 1e3))
    sim.set_stage_timing(False)
    sim.synchronize(); t0 = time.time()
    for _ in range(steps):
        sim.step(scenes.DT)
    sim.synchronize()
    wall = (time.time() - t0) / steps
    print("   untimed wall=%.3f ms  Mupd/s=%.1f" % (wall * 1e3, sc["n"] / wall / 1e6))
    d = sim.download("densities")
    print("   dens mean", d.mean(axis=0), "finite", np.isfinite(d).all())
    sim.close()
