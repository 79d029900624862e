"""Regenerates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libsph_ref.so, built by
oracle/Makefile from /root/reference/engine/physics/physicsWorld.cc).  Run in the build container:

    python tests/golden/make_golden.py

The fixtures pin (a) the C restatement oracle/sph_oracle.c and (b) the CUDA path on boxes where the
reference sources do not exist.  Each file stores inputs, parameters and every intermediate of ONE
step taken with snapshot (Jacobi) viscosity through the reference's own stage functions, plus the
result of the verbatim Update() (in-place viscosity) for the whole-step check.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

ob = g.load_oracle()
pkg_scenes = None


def scenes():
    g.load_package()
    from fluid_simulation_3d_b200 import scenes as s
    return s


def capture(name, pos, vel, params, dt, spawn_n=None):
    n = spawn_n if spawn_n else pos.shape[0]
    r = ob.RefOracle(n, spawn=bool(spawn_n), **params)
    if spawn_n:
        pos, vel = r.positions(), r.velocities()
        spawn_dens = r.densities()
    else:
        r.set_state(pos, vel)
        spawn_dens = np.zeros((0, 2), np.float32)
    # verbatim Update() first (in-place viscosity), then restore and take the staged Jacobi step
    r.update(dt)
    upd_pos, upd_vel, upd_dens = r.positions(), r.velocities(), r.densities()
    r.set_state(pos, vel)
    r.step(dt, jacobi=True)
    h, k, cells = r.hash_key()
    s_idx, s_hash, s_key = r.sorted_lookup()
    out = dict(pos0=pos, vel0=vel, dt=np.float32(dt), spawn_dens=spawn_dens,
               params=np.array([params.get("interaction_radius", 0.35), params.get("target_density", 99.7),
                                params.get("pressure_multiplier", 300.0), params.get("near_pressure_multiplier", 20.0),
                                params.get("viscosity_strength", 0.5), params.get("gravity_scale", 10.0),
                                float(params.get("gravity", 0))] + list(params.get("bound", (20, 20, 20))), np.float64),
               pred=r.predicted(), hash=h, key=k, cells=cells, sorted_idx=s_idx, sorted_hash=s_hash, sorted_key=s_key,
               start=r.start_indices(), ncount=r.neighbour_counts(), dens=r.densities(),
               vel_press=r.vel_after_pressure(), vel_visc=r.vel_after_viscosity(), pos1=r.positions(), vel1=r.velocities(),
               out1=r.out_positions(), upd_pos=upd_pos, upd_vel=upd_vel, upd_dens=upd_dens)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(name, n, "particles ->", os.path.getsize(path) // 1024, "KiB; fnv(pos1)=", ob.fnv1a64(out["pos1"]))


def main():
    if not ob.have_ref():
        raise SystemExit("oracle/_ref/libsph_ref.so missing: run `make -C oracle` where /root/reference exists")
    sc = scenes()
    dt = sc.DT
    # (1) the reference's own spawn, scaled down from the default scene (InitializeData(1000), gravity on)
    capture("spawn_1000", None, None, dict(gravity=1), dt, spawn_n=1000)
    # (2) jittered dam-break block, C2's rule at 12^3
    s = sc.small_dam_break(12)
    capture("dambreak_12", s["pos"], s["vel"], s["params"], dt)
    # (3) dense column (C5's rule), ~100 neighbours, random velocities, mu = 1
    s = sc.small_column(9, 18, 9)
    capture("column_9x18x9", s["pos"], s["vel"], s["params"], dt)
    # (4) all sign octants, fast particles, wall hits
    rng = np.random.default_rng(11)
    n = 1500
    bound = (4.0, 3.0, 3.5)
    pos = ((rng.random((n, 3), dtype=np.float32) - 0.5) * np.array(bound, np.float32) * 1.02).astype(np.float32)
    vel = ((rng.random((n, 3), dtype=np.float32) - 0.5) * 8.0).astype(np.float32)
    capture("octants_1500", pos, vel, dict(gravity=1, viscosity_strength=0.7, bound=bound), dt)


if __name__ == "__main__":
    main()
