import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as g
from helpers import check_step
pkg = g.load_package()
from fluid_simulation_3d_b200 import scenes
try:
    print(check_step(pkg, scenes.small_dam_break(12), pkg.TABLE_GRID, scenes.DT))
except Exception as e:
    print("FAIL", str(e)[:300])
