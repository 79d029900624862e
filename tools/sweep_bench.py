"""A/B timing of library variants and environment switches, one process each (scratch tool, not the contract bench):

    python tools/sweep_bench.py C2_dambreak_1M 30 base build/variants/t64.so SPH_XSUB=2 ...

`base` = the in-tree library; a path = SPH_B200_LIB; NAME=VALUE = an environment switch on the in-tree library.
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r"""
import sys, os, json
sys.path.insert(0, %r)
import numpy as np
import __graft_entry__ as g
pkg = g.load_package()
from fluid_simulation_3d_b200 import scenes
name, steps = sys.argv[1], int(sys.argv[2])
sc = scenes.config(name)
sim = pkg.FluidSimulation(sc["n"], **sc["params"])
sim.set_graph_replay(False)          # plain launches: the six stage timers describe every step
sim.upload_state(sc["pos"], sc["vel"])
for _ in range(6):
    sim.step(scenes.DT)
sim.synchronize()
acc = np.zeros(6)
for _ in range(steps):
    sim.step(scenes.DT)
    acc += sim.timings()
acc /= steps
d = sim.download("densities")
print(json.dumps(dict(stage_us=[round(float(x) * 1e3, 1) for x in acc], sum_us=round(float(acc.sum()) * 1e3, 1),
                      dens_mean=float(d[:, 0].astype(np.float64).mean()))))
"""

if __name__ == "__main__":
    name, steps = sys.argv[1], sys.argv[2]
    for v in sys.argv[3:]:
        env = dict(os.environ)
        if v != "base":
            if "=" in v and not v.endswith(".so"):
                for kv in v.split(","):
                    k, val = kv.split("=", 1)
                    env[k] = val
            else:
                env["SPH_B200_LIB"] = os.path.join(ROOT, v)
        r = subprocess.run([sys.executable, "-c", CHILD % ROOT, name, steps], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                           text=True, timeout=600)
        last = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else "(no output)"
        print("%-28s %s" % (v, last), flush=True)
