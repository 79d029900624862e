"""Random-problem parity sweep of the CUDA step against the oracle (a tool for a GPU box, not part of the test suite):

    python tools/gpu_fuzz.py [cases] [seed]

Each case draws a particle count (1..4000), box, interaction radius (also != the cut-off), every solver
parameter, gravity, positions partly outside the box, fast particles and stacked duplicates, runs ONE step in both table
modes through tests/helpers.check_step (integers bit-exact, floats within 1e-5 of the stage scale) and prints the cases
that fail.  The same generator runs against the unmodified reference on CPU in
tests/test_oracle.py::test_restatement_fuzzed_against_the_live_reference."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as g  # noqa: E402


def draw(rng):
    n = int(rng.integers(1, 4000))
    bound = tuple(float(x) for x in rng.uniform(2.0, 12.0, 3))
    r = float(rng.choice([0.35, 0.35, 0.25, 0.5, float(rng.uniform(0.2, 0.6))]))
    prm = dict(interaction_radius=r, target_density=float(rng.uniform(20, 200)), pressure_multiplier=float(rng.uniform(10, 500)),
               near_pressure_multiplier=float(rng.uniform(1, 40)), viscosity_strength=float(rng.uniform(0, 1)),
               gravity_scale=float(rng.uniform(0, 20)), gravity=int(rng.integers(0, 2)), bound=bound)
    half = np.array(bound, np.float32) / 2
    pos = ((rng.random((n, 3)) - 0.5) * 2 * half * rng.choice([0.3, 0.9, 1.2])).astype(np.float32)
    vel = ((rng.random((n, 3)) - 0.5) * rng.choice([0.0, 2.0, 16.0])).astype(np.float32)
    if n > 4 and rng.random() < 0.5:
        d = rng.integers(0, n, max(1, n // 10))
        s = rng.integers(0, n, len(d))
        pos[d] = pos[s]
        vel[d] = vel[s]
    dt = float(np.float32(rng.choice([0.016667, 0.005, 0.033])))
    return dict(pos=pos, vel=vel, n=n, params=prm), dt


def main():
    cases = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 2026
    pkg = g.load_package()
    from helpers import check_step
    rng = np.random.default_rng(seed)
    failed = 0
    for case in range(cases):
        sc, dt = draw(rng)
        for mode, name in ((pkg.TABLE_GRID, "grid"), (pkg.TABLE_REFERENCE_HASH, "reference_hash")):
            try:
                check_step(pkg, sc, mode, dt)
            except (AssertionError, pkg.SphError) as e:
                failed += 1
                p = sc["params"]
                print("case %d %s: n=%d r=%.4f bound=%s dt=%g -> %s" % (case, name, sc["n"], p["interaction_radius"],
                                                                      tuple(round(b, 3) for b in p["bound"]), dt, str(e)[:300]), flush=True)
    print("gpu_fuzz: %d cases x 2 table modes, %d failures (seed %d)" % (cases, failed, seed))
    return 1 if failed else 0


if __name__ == "__main__":
    sys.exit(main())
