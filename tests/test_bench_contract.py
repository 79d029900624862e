"""CPU: the reference arm of bench.py (`--impl reference`) runs without a GPU -- it times the reference's own CPU
implementation through oracle/ -- and prints the contract's JSON line; under torchrun only rank 0 works and prints."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env=None, *args):
    env = dict(os.environ)
    env.update(extra_env or {})
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "C1_default_10k",
                        "--steps", "2", "--warmup", "1", *args], cwd=ROOT, env=env, stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    return [ln for ln in r.stdout.splitlines() if ln.startswith("{")]


def test_reference_arm_line():
    lines = _run()
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "M particle-updates/s" and d["unit"] == "M updates/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None and d["dtype"] == "f32"
    assert d["config"]["workload"].startswith("C1_default_10k: 10000 particles") and d["config"]["sample_particles"] == 10000
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and 1 <= cb["cores"] <= (os.cpu_count() or 1) and cb["value"] == d["value"] and cb["sample"]
    if cb["cores"] > 1:                      # the unmodified reference on every host thread (oracle/pstl_threads)
        assert cb["kind"] == "reference" and "host threads" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "M updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_reference_arm_under_torchrun_only_rank0_prints():
    # torchrun exports OMP_NUM_THREADS=1 to every rank; rank 0 runs the arm alone and must still take the whole box
    env = {"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2", "MASTER_ADDR": "127.0.0.1", "MASTER_PORT": "29599", "OMP_NUM_THREADS": "1"}
    assert _run(env, "--gpus", "2") == []
    env["RANK"] = "0"; env["LOCAL_RANK"] = "0"
    lines = _run(env, "--gpus", "2")
    d = json.loads(lines[0])
    assert len(lines) == 1 and d["n_gpus"] == 2 and d["scaling"] == "strong"
    import __graft_entry__ as g
    if g.load_oracle().have_ref_par():
        assert d["cpu_baseline"]["cores"] == (os.cpu_count() or 1)
