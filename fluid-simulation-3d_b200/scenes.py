"""Synthetic particle sets of BASELINE.json's configs.

All generators are counter-based (splitmix64 of seed ^ (3*id + axis)), so any rank can generate any
sub-range of particle ids and CPU / GPU runs see bit-identical inputs.  Particle id <-> lattice site
follows the reference's fill order (GridArrangement, physicsWorld.cc:526-530): y outer (top layer
first), then x, then z inner:  id = (iy*nx + ix)*nz + iz.
"""
import numpy as np

GAP0 = 0.215            # reference spawn gap (physicsWorld.cc:140)
DT = float(np.float32(0.016667))


def splitmix64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15)) & np.uint64(0xFFFFFFFFFFFFFFFF)
    z = x
    z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & np.uint64(0xFFFFFFFFFFFFFFFF)
    z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & np.uint64(0xFFFFFFFFFFFFFFFF)
    return z ^ (z >> np.uint64(31))


def uniform01(seed, ids, axis):
    """float32 in [0,1) from the top 24 bits of splitmix64(seed ^ (3*id + axis))."""
    with np.errstate(over="ignore"):
        ctr = np.uint64(seed) ^ (ids.astype(np.uint64) * np.uint64(3) + np.uint64(axis))
        bits = splitmix64(ctr) >> np.uint64(40)
    return (bits.astype(np.float32) * np.float32(1.0 / (1 << 24))).astype(np.float32)


def block_origin(nx, ny, nz, gap, bound, anchor="corner"):
    """fp64 position of the lattice site with the smallest x, y, z (what sph_spawn_block takes as `origin`)."""
    b = np.asarray(bound, dtype=np.float64)
    ext = np.array([nx, ny, nz], dtype=np.float64) * gap
    if anchor == "corner":
        return -b / 2 + gap / 2
    if anchor == "floor_center":
        return np.array([-ext[0] / 2 + gap / 2, -b[1] / 2 + gap / 2, -ext[2] / 2 + gap / 2])
    return -ext / 2 + gap / 2


def device_spawn_args(nx, ny, nz, gap, bound, seed, anchor="corner", jitter=0.1, vel_amp=0.0):
    """Arguments of FluidSimulation.spawn_block (sph_spawn_block) that reproduce block(...) bit for bit on the device."""
    return dict(nx=nx, ny=ny, nz=nz, gap=float(gap), origin=[float(x) for x in block_origin(nx, ny, nz, gap, bound, anchor)],
                jitter_amp=float(np.float32(2.0 * jitter * gap)) if jitter else 0.0,
                velocity_scale=float(np.float32(2.0 * vel_amp)) if vel_amp else 0.0, seed=int(seed))


def block(nx, ny, nz, gap, bound, seed, anchor="corner", jitter=0.1, vel_amp=0.0, ids=None):
    """Jittered lattice block.  anchor='corner': flush to the -x wall, the floor and the -z wall
    (min corner = -bound/2 + gap/2); 'floor_center': centred in x,z, standing on the floor;
    'center': centred in all axes.  Returns (pos[n,3], vel[n,3]) float32 for the given ids."""
    n = nx * ny * nz
    if ids is None:
        ids = np.arange(n, dtype=np.uint64)
    ids = np.asarray(ids, dtype=np.uint64)
    iz = (ids % np.uint64(nz)).astype(np.int64)
    ix = ((ids // np.uint64(nz)) % np.uint64(nx)).astype(np.int64)
    iy = (ids // np.uint64(nz * nx)).astype(np.int64)
    lo = block_origin(nx, ny, nz, gap, bound, anchor)
    pos = np.empty((ids.size, 3), dtype=np.float32)
    pos[:, 0] = (lo[0] + ix * gap).astype(np.float32)
    pos[:, 1] = (lo[1] + (ny - 1 - iy) * gap).astype(np.float32)   # iy counts from the top layer down
    pos[:, 2] = (lo[2] + iz * gap).astype(np.float32)
    if jitter:
        amp = np.float32(2.0 * jitter * gap)
        for a in range(3):
            pos[:, a] += (uniform01(seed, ids, a) - np.float32(0.5)) * amp
    vel = np.zeros((ids.size, 3), dtype=np.float32)
    if vel_amp:
        for a in range(3):
            vel[:, a] = (uniform01(seed ^ 0x5EED, ids, a) - np.float32(0.5)) * np.float32(2.0 * vel_amp)
    return pos, vel


def dam_break(nx, ny, nz, seed, ids=None):
    """C2/C3/C4 rule: bounds = (3*nx*g0, 1.5*ny*g0, nz*g0 + g0), block in the -x / floor corner."""
    bound = (3 * nx * GAP0, 1.5 * ny * GAP0, nz * GAP0 + GAP0)
    pos, vel = block(nx, ny, nz, GAP0, bound, seed, anchor="corner", ids=ids)
    return pos, vel, bound


CONFIGS = {
    # name: (nx, ny, nz, seed)
    "C2_dambreak_1M": (100, 100, 100, 0xC2),
    "C3_dambreak_8M": (200, 200, 200, 0xC3),
    "C4_dambreak_64M": (400, 400, 400, 0xC4),
    # a quarter of C4 in z: on 2 GPUs each rank holds what a rank of C4 holds on 8 (400 x 400 x 50 lattice sites)
    "C4q_dambreak_16M": (400, 400, 100, 0xC4),
}


def config_meta(name):
    """dict(bound, params, n, dims, spawn) of a named BASELINE config WITHOUT generating the particles: `spawn` are the
    sph_spawn_block arguments that make the same set, bit for bit, on the device"""
    if name in CONFIGS:
        nx, ny, nz, seed = CONFIGS[name]
        bound = (3 * nx * GAP0, 1.5 * ny * GAP0, nz * GAP0 + GAP0)
        return dict(bound=bound, n=nx * ny * nz, dims=(nx, ny, nz), params=dict(gravity=1, viscosity_strength=0.5, bound=bound),
                    spawn=device_spawn_args(nx, ny, nz, GAP0, bound, seed, anchor="corner"))
    if name == "C5_column_8M":
        nx, ny, nz, gap = 100, 800, 100, 0.1216
        bound = (36.5, 146.0, 36.5)
        return dict(bound=bound, n=nx * ny * nz, dims=(nx, ny, nz), params=dict(gravity=1, viscosity_strength=1.0, bound=bound),
                    spawn=device_spawn_args(nx, ny, nz, gap, bound, 0xC5, anchor="floor_center", vel_amp=0.5))
    raise KeyError(name)


def config(name, ids=None):
    """Returns dict(pos, vel, bound, params, n) for a named BASELINE config."""
    if name in CONFIGS:
        nx, ny, nz, seed = CONFIGS[name]
        pos, vel, bound = dam_break(nx, ny, nz, seed, ids=ids)
        return dict(pos=pos, vel=vel, bound=bound, n=nx * ny * nz, dims=(nx, ny, nz),
                    params=dict(gravity=1, viscosity_strength=0.5, bound=bound),
                    spawn=device_spawn_args(nx, ny, nz, GAP0, bound, seed, anchor="corner"))
    if name == "C5_column_8M":
        nx, ny, nz, gap = 100, 800, 100, 0.1216
        bound = (36.5, 146.0, 36.5)
        pos, vel = block(nx, ny, nz, gap, bound, 0xC5, anchor="floor_center", vel_amp=0.5, ids=ids)
        return dict(pos=pos, vel=vel, bound=bound, n=nx * ny * nz, dims=(nx, ny, nz),
                    params=dict(gravity=1, viscosity_strength=1.0, bound=bound),
                    spawn=device_spawn_args(nx, ny, nz, gap, bound, 0xC5, anchor="floor_center", vel_amp=0.5))
    raise KeyError(name)


def small_dam_break(n_side, seed=7, ids=None):
    """Scaled-down C2 (same rule) for parity tests the oracle finishes in seconds."""
    pos, vel, bound = dam_break(n_side, n_side, n_side, seed, ids=ids)
    return dict(pos=pos, vel=vel, bound=bound, n=n_side ** 3, dims=(n_side,) * 3,
                params=dict(gravity=1, viscosity_strength=0.5, bound=bound),
                spawn=device_spawn_args(n_side, n_side, n_side, GAP0, bound, seed, anchor="corner"))


def small_column(nx, ny, nz, seed=0xC5):
    """Scaled-down C5: dense column (~100 neighbours), random velocities, mu = 1."""
    gap = 0.1216
    bound = (max(4.0, nx * gap * 3), max(6.0, ny * gap * 1.5), max(4.0, nz * gap * 3))
    pos, vel = block(nx, ny, nz, gap, bound, seed, anchor="floor_center", vel_amp=0.5)
    return dict(pos=pos, vel=vel, bound=bound, n=nx * ny * nz, dims=(nx, ny, nz),
                params=dict(gravity=1, viscosity_strength=1.0, bound=bound),
                spawn=device_spawn_args(nx, ny, nz, gap, bound, seed, anchor="floor_center", vel_amp=0.5))
