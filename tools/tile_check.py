"""Scratch GPU check of the tile generation: a few scenes through tests/helpers.check_step, with timings."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as g
from helpers import check_step
pkg = g.load_package()
from fluid_simulation_3d_b200 import scenes

def run(name, sc, mode=None):
    mode = pkg.TABLE_GRID if mode is None else mode
    try:
        out = check_step(pkg, sc, mode, scenes.DT)
        print("OK  ", name, {k: (round(v, 9) if isinstance(v, float) else v) for k, v in out.items()}, flush=True)
    except Exception as e:
        print("FAIL", name, str(e)[:400], flush=True)

run("dam14", scenes.small_dam_break(14))
run("dam30", scenes.small_dam_break(30))
run("column", scenes.small_column(10, 24, 10))
run("dense_column", scenes.small_column(14, 40, 14))
rng = np.random.default_rng(5)
n = 6000
pos = ((rng.random((n, 3)) - 0.5) * np.array([6.0, 4.0, 5.0])).astype(np.float32)
vel = ((rng.random((n, 3)) - 0.5) * 8).astype(np.float32)
run("random6000", dict(pos=pos, vel=vel, n=n, params=dict(gravity=1, bound=(6.0, 4.0, 5.0))))
run("dam50", scenes.small_dam_break(50))
