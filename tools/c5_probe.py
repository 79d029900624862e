"""How dense is C5 over time? prints mean/max neighbour count and step time for the first steps."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as g
pkg = g.load_package()
from fluid_simulation_3d_b200 import scenes
sc = scenes.config("C5_column_8M")
sim = pkg.FluidSimulation(sc["n"], **sc["params"])
if len(sys.argv) > 1:
    sim.set_neighbour_list_capacity(int(sys.argv[1]))
    sim.upload_state(sc["pos"], sc["vel"]); sim.step(scenes.DT); sim.synchronize()   # allocate the list outside the timed steps
sim.upload_state(sc["pos"], sc["vel"])
for s in range(16):
    sim.step(scenes.DT)
    t = sim.timings()
    if s in (0, 1, 2, 3, 4, 5, 7, 10, 15):
        nc = sim.download("neighbour_count")
        v = sim.download("velocities")
        print("step %2d: neighbours mean %.1f max %d | stage ms %s sum %.2f | |v| max %.1f" % (
            s, nc.mean(), nc.max(), np.round(t, 2), t.sum(), np.abs(v).max()), flush=True)
