"""The drop-in C++ class over several GPUs of one box, timed end to end (scratch tool for a multi-GPU box):

    python tools/class_multi_bench.py C3_dambreak_8M 20 8

writes the config as a snapshot file, then runs `host_demo classbench` on 1 GPU and on N GPUs, the latter with the staged
host scatter and with the direct (device-written) mirrors."""
import os, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import __graft_entry__ as g
pkg = g.load_package()
from fluid_simulation_3d_b200 import scenes
name = sys.argv[1] if len(sys.argv) > 1 else "C3_dambreak_8M"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 20
ndev = int(sys.argv[3]) if len(sys.argv) > 3 else 2
sc = scenes.config(name)
p = pkg.default_params(**sc["params"])
demo = os.path.join(ROOT, "fluid-simulation-3d_b200", "host", "host_demo")
with tempfile.TemporaryDirectory() as td:
    path = os.path.join(td, "state.bin")
    bench.write_snapshot(path, sc["pos"], sc["vel"], p)
    for nd, direct in ((1, 0), (ndev, 0), (ndev, 1)):
        r = subprocess.run([demo, "0", str(frames), "0", "classbench", path, str(nd), str(direct)], stdout=subprocess.PIPE,
                           stderr=subprocess.STDOUT, text=True, timeout=900)
        print([ln for ln in r.stdout.splitlines() if ln.startswith("classbench")] or r.stdout[-400:], flush=True)
