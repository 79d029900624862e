"""Scratch: stage timers of a config after it has evolved for many steps (plain launches)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g
pkg = g.load_package()
from fluid_simulation_3d_b200 import scenes
sc = scenes.config(sys.argv[1] if len(sys.argv) > 1 else "C2_dambreak_1M")
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1500
sim = pkg.FluidSimulation(sc["n"], **sc["params"])
sim.set_neighbour_count_tap(True)
sim.upload_state(sc["pos"], sc["vel"])
for chunk in (10, steps):
    sim.step_n(scenes.DT, chunk)
    sim.set_graph_replay(False)
    acc = []
    for _ in range(10):
        sim.step(scenes.DT); acc.append(sim.timings())
    sim.set_graph_replay(True)
    nc = sim.download("neighbour_count")
    t = np.median(np.array(acc), axis=0) * 1e3
    print("after ~%d steps: stage us %s sum %.1f | neighbours mean %.2f max %d p99 %d | occupied fine cells %d"
          % (chunk, [round(float(x), 1) for x in t], float(t.sum()), float(nc.mean()), int(nc.max()), int(np.percentile(nc, 99)),
             int(np.unique(sim.download_table("sorted_key")).size)), flush=True)
