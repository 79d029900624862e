"""Run a config for N replayed steps, then a few plain steps (for ncu: -k regex:k_density_pk with --launch-skip past the replays)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g
pkg = g.load_package()
from fluid_simulation_3d_b200 import scenes
sc = scenes.config(sys.argv[1] if len(sys.argv) > 1 else "C2_dambreak_1M")
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1500
sim = pkg.FluidSimulation(sc["n"], **sc["params"])
sim.set_stage_timing(False)
sim.upload_state(sc["pos"], sc["vel"])
sim.step_n(scenes.DT, steps)
sim.synchronize()
sim.set_graph_replay(False)
import torch
torch.cuda.profiler.start()
for _ in range(2):
    sim.step(scenes.DT)
sim.synchronize()
torch.cuda.profiler.stop()
