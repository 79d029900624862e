"""Scratch: per-call timing of graph-replayed steps, back to back and one by one."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import __graft_entry__ as g
pkg = g.load_package()
from fluid_simulation_3d_b200 import scenes
sc = scenes.config(sys.argv[1] if len(sys.argv) > 1 else "C2_dambreak_1M")
sim = pkg.FluidSimulation(sc["n"], **sc["params"])
stream = torch.cuda.ExternalStream(sim.stream_ptr(), device=0)
sim.upload_state(sc["pos"], sc["vel"])
for _ in range(5): sim.step(scenes.DT)
sim.set_stage_timing(False)
sim.step_n(scenes.DT, 4)
sim.synchronize()
def timed(fn, reps):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream); fn(); b.record(stream); sim.synchronize()
    return a.elapsed_time(b) / reps
print("replays", sim.graph_replays())
print("step_n(20) back to back  ms/step", timed(lambda: sim.step_n(scenes.DT, 20), 20), "replays", sim.graph_replays())
print("20 x step_n(1)           ms/step", timed(lambda: [sim.step_n(scenes.DT, 1) for _ in range(20)], 20), "replays", sim.graph_replays())
one = [timed(lambda: sim.step_n(scenes.DT, 1), 1) for _ in range(10)]
print("step_n(1) one by one     ms", [round(x, 3) for x in one])
sim.set_stage_timing(True)
print("plain step, timers on    ms", [round(timed(lambda: sim.step(scenes.DT), 1), 3) for _ in range(5)], sim.timings())
sim.set_stage_timing(False)
print("plain step, timers off   ms", [round(timed(lambda: sim.step(scenes.DT), 1), 3) for _ in range(5)])
print("step_n(20) again         ms/step", timed(lambda: sim.step_n(scenes.DT, 20), 20))
