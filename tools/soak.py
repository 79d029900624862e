"""Long-run check on a GPU box: step a BASELINE config for many steps (CUDA-graph replay, list growth and all) and, every
`every` steps, take the device state, step it once more on the GPU AND on the CPU restatement, and compare (integers
bit-exact, floats within 1e-5 of the stage scale: tests/helpers.check_step_port).

    python tools/soak.py [config] [steps] [every]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as g
import helpers
pkg = g.load_package(); ob = g.load_oracle()
from fluid_simulation_3d_b200 import scenes
name = sys.argv[1] if len(sys.argv) > 1 else "C2_dambreak_1M"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
every = int(sys.argv[3]) if len(sys.argv) > 3 else 500
sc = scenes.config(name)
sim = pkg.FluidSimulation(sc["n"], **sc["params"])
sim.upload_state(sc["pos"], sc["vel"])
done = 0
while done < steps:
    t0 = time.perf_counter()
    sim.step_n(scenes.DT, every)
    sim.synchronize()
    wall = (time.perf_counter() - t0) / every * 1e3
    done += every
    pos, vel = sim.download("positions"), sim.download("velocities")
    assert np.all(np.isfinite(pos)) and np.all(np.isfinite(vel)), "non-finite state after %d steps" % done
    ev = dict(pos=np.ascontiguousarray(pos), vel=np.ascontiguousarray(vel), n=sc["n"], params=sc["params"])
    mean = helpers.check_step_port(pkg, ob, ev, scenes.DT, label="%s after %d steps" % (name, done))
    print("step %5d: %.3f ms/step wall, |v|max %.2f, mean neighbours %.2f, replays %d, non-canonical cells %d: one more step matches the oracle"
          % (done, wall, float(np.abs(vel).max()), mean, sim.graph_replays(), sim.noncanonical_cells()), flush=True)
print("SOAK_OK %s %d steps" % (name, steps))
