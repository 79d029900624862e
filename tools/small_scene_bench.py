"""Step rate of the reference's own scene sizes (InitializeData lattice, default bounds, gravity on), where the step
is launch-bound: sph_step_n with CUDA-graph replay against plain launches (SPH_GRAPH=0 in a second process).

    python tools/small_scene_bench.py [n ...]      -> one JSON line per particle count
"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def measure(n, steps=2000):
    import __graft_entry__ as g
    pkg = g.load_package()
    sim = pkg.FluidSimulation(n, gravity=1)
    sim.spawn_grid(n)
    sim.set_stage_timing(False)
    dt = 0.004
    sim.step_n(dt, 50)
    sim.synchronize()
    r0 = sim.graph_replays()
    t0 = time.perf_counter()
    sim.step_n(dt, steps)
    sim.synchronize()
    t = time.perf_counter() - t0
    out = dict(n=n, steps=steps, us_per_step=t / steps * 1e6, m_updates_per_s=n * steps / t / 1e6,
               graph_replays=sim.graph_replays() - r0)
    sim.close()
    return out


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--child":
        print(json.dumps(measure(int(sys.argv[2]))))
        sys.exit(0)
    for n in [int(a) for a in sys.argv[1:]] or [10000, 100000]:
        row = {}
        for label, env in (("graph", {}), ("plain", {"SPH_GRAPH": "0"})):
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", str(n)], env=dict(os.environ, **env),
                               stdout=subprocess.PIPE, text=True, timeout=600)
            row[label] = json.loads(r.stdout.strip().splitlines()[-1])
        print(json.dumps({"particles": n, "graph": row["graph"], "plain": row["plain"],
                          "speedup": row["plain"]["us_per_step"] / row["graph"]["us_per_step"]}), flush=True)
