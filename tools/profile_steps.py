"""Run a few steps of one config for ncu: python tools/profile_steps.py <config> <mode grid|refhash> <warmup> <steps>"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
pkg = g.load_package()
from fluid_simulation_3d_b200 import scenes
name = sys.argv[1] if len(sys.argv) > 1 else "C2_dambreak_1M"
mode = pkg.TABLE_REFERENCE_HASH if (len(sys.argv) > 2 and sys.argv[2] == "refhash") else pkg.TABLE_GRID
warm = int(sys.argv[3]) if len(sys.argv) > 3 else 3
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
sc = scenes.config(name)
sim = pkg.FluidSimulation(sc["n"], table_mode=mode, **sc["params"])
sim.set_stage_timing(False)
sim.upload_state(sc["pos"], sc["vel"])
for _ in range(warm + steps):
    sim.step(scenes.DT)
sim.synchronize()
print("launches", sim.launch_count())
