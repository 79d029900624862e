"""fluid-simulation-3d_b200 -- B200-native per-timestep SPH update behind a C ABI.

This Python module is only a ctypes binding of ``libsph_b200.so`` (include/sph_b200.h) for the
tests and bench.py; the product is the shared library and the C++ host class in ``host/``.
There is NO CPU fallback: importing works without a GPU (so the symbol table can be checked),
but creating a simulation without the library or without a CUDA device raises.

The directory name is not a valid Python identifier; load it with ``load_package()`` from
``__graft_entry__`` (tests and bench.py do), which registers it as ``fluid_simulation_3d_b200``.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# SPH_B200_LIB: load another build of the same library (A/B runs of kernel variants; development only)
LIB_PATH = os.environ.get("SPH_B200_LIB") or os.path.join(HERE, "libsph_b200.so")

SPH_OK = 0
TABLE_GRID, TABLE_REFERENCE_HASH = 0, 1

FIELDS = dict(positions=0, out_positions=1, velocities=2, densities=3, predicted=4,
              vel_after_pressure=5, vel_after_viscosity=6, hash=7, key=8, neighbour_count=9,
              speed_normalized=10, colors=11)
_FIELD_SHAPE = {0: (3, np.float32), 1: (4, np.float32), 2: (3, np.float32), 3: (2, np.float32),
                4: (3, np.float32), 5: (3, np.float32), 6: (3, np.float32), 7: (0, np.uint32),
                8: (0, np.uint32), 9: (0, np.uint32), 10: (0, np.float32), 11: (4, np.float32)}
TABLES = dict(sorted_index=0, sorted_key=1, start_indices=2, sorted_hash=3)

# every symbol include/sph_b200.h declares (tests/test_abi.py checks the header against this list)
ABI_SYMBOLS = [
    "sph_create", "sph_destroy", "sph_last_error", "sph_abi_version", "sph_default_params",
    "sph_set_params", "sph_get_params", "sph_set_table_mode", "sph_get_table_mode",
    "sph_set_stage_timing", "sph_set_neighbour_count_tap", "sph_set_neighbour_list_capacity", "sph_spawn_grid", "sph_spawn_block", "sph_upload_state",
    "sph_num_particles", "sph_step", "sph_step_n", "sph_graph_replays", "sph_set_graph_replay", "sph_noncanonical_cells", "sph_density_stack_rows", "sph_set_extras", "sph_get_extras", "sph_synchronize", "sph_refresh_densities",
    "sph_download", "sph_download_table", "sph_get_particle", "sph_get_timings", "sph_launch_count", "sph_stream",
    "sph_get_grid", "sph_grid_x_subdivision", "sph_save_state", "sph_load_state", "sph_host_register", "sph_host_unregister",
    "sph_upload_state_begin", "sph_upload_state_commit", "sph_download_begin", "sph_download_wait",
    "sph_comm_id_bytes", "sph_comm_get_id", "sph_comm_init", "sph_comm_set_planes",
    "sph_upload_owned", "sph_download_owned", "sph_upload_owned_begin", "sph_download_owned_begin", "sph_comm_stats",
    "sph_comm_get_layers", "sph_comm_rebalance", "sph_slab_balance_layers", "sph_slab_link_ok", "sph_download_owned_scatter",
]


class SphParams(C.Structure):
    """Mirror of ``struct SphParams`` (reference members physicsWorld.h:96-106,145)."""
    _fields_ = [("interaction_radius", C.c_float), ("sqr_radius", C.c_float),
                ("target_density", C.c_float), ("pressure_multiplier", C.c_float),
                ("near_pressure_multiplier", C.c_float), ("viscosity_strength", C.c_float),
                ("gravity_scale", C.c_float), ("gravity", C.c_int32), ("bound", C.c_float * 3)]


class SphExtras(C.Structure):
    """Mirror of ``struct SphExtras`` (rotatable bound, wall stickiness: off by default)."""
    _fields_ = [("bound_rotation", C.c_float * 4), ("stick_strength", C.c_float), ("stick_distance", C.c_float)]


class SphBlockSpawn(C.Structure):
    """Mirror of ``struct SphBlockSpawn`` (device-side scene spawn)."""
    _fields_ = [("nx", C.c_uint32), ("ny", C.c_uint32), ("nz", C.c_uint32), ("reserved", C.c_uint32),
                ("gap", C.c_double), ("origin", C.c_double * 3), ("jitter_amp", C.c_float),
                ("velocity_scale", C.c_float), ("seed", C.c_uint64)]


class SphError(RuntimeError):
    pass


_lib = None


def _preload_torch_nccl():
    """PyTorch bundles its own libnccl.so.2; a process must hold exactly one NCCL, and libsph_b200.so
    adopts whichever is already loaded, so make sure it is the one torch will want."""
    import importlib.util
    import sys
    if "torch" in sys.modules:
        return
    for mod in ("nvidia.nccl",):
        try:
            spec = importlib.util.find_spec(mod)
        except (ImportError, ValueError):
            spec = None
        if spec and spec.submodule_search_locations:
            for d in spec.submodule_search_locations:
                path = os.path.join(d, "lib", "libnccl.so.2")
                if os.path.exists(path):
                    try:
                        C.CDLL(path, mode=C.RTLD_GLOBAL)
                    except OSError:
                        pass
                    return


def load_library():
    """Load libsph_b200.so; raises loudly if it has not been built (no fallback of any kind)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SphError("libsph_b200.so is not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(make -C fluid-simulation-3d_b200). There is no CPU fallback.")
    _preload_torch_nccl()
    L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    vp, u32, f32 = C.c_void_p, C.c_uint32, C.c_float
    L.sph_create.argtypes = [C.POINTER(vp), C.c_int, u32]
    L.sph_destroy.argtypes = [vp]
    L.sph_last_error.argtypes = [vp]
    L.sph_last_error.restype = C.c_char_p
    L.sph_default_params.argtypes = [C.POINTER(SphParams)]
    L.sph_default_params.restype = None
    L.sph_set_params.argtypes = [vp, C.POINTER(SphParams)]
    L.sph_get_params.argtypes = [vp, C.POINTER(SphParams)]
    L.sph_set_table_mode.argtypes = [vp, C.c_int]
    L.sph_get_table_mode.argtypes = [vp]
    L.sph_set_stage_timing.argtypes = [vp, C.c_int]
    L.sph_set_neighbour_count_tap.argtypes = [vp, C.c_int]
    L.sph_set_neighbour_list_capacity.argtypes = [vp, u32]
    L.sph_spawn_grid.argtypes = [vp, u32]
    L.sph_spawn_block.argtypes = [vp, C.POINTER(SphBlockSpawn)]
    L.sph_upload_state.argtypes = [vp, u32, vp, vp]
    L.sph_num_particles.argtypes = [vp]
    L.sph_num_particles.restype = u32
    L.sph_step.argtypes = [vp, f32]
    L.sph_step_n.argtypes = [vp, f32, u32]
    L.sph_synchronize.argtypes = [vp]
    L.sph_refresh_densities.argtypes = [vp]
    L.sph_download.argtypes = [vp, C.c_int, vp, C.c_size_t]
    L.sph_download_table.argtypes = [vp, C.c_int, vp, C.c_size_t, C.POINTER(C.c_size_t)]
    L.sph_get_particle.argtypes = [vp, u32, vp]
    L.sph_get_timings.argtypes = [vp, vp]
    L.sph_launch_count.argtypes = [vp]
    L.sph_launch_count.restype = C.c_uint64
    L.sph_graph_replays.argtypes = [vp]
    L.sph_graph_replays.restype = C.c_uint64
    L.sph_set_extras.argtypes = [vp, C.POINTER(SphExtras)]
    L.sph_get_extras.argtypes = [vp, C.POINTER(SphExtras)]
    L.sph_set_graph_replay.argtypes = [vp, C.c_int]
    L.sph_noncanonical_cells.argtypes = [vp]
    L.sph_noncanonical_cells.restype = C.c_uint64
    L.sph_density_stack_rows.argtypes = [vp]
    L.sph_density_stack_rows.restype = C.c_int
    L.sph_stream.argtypes = [vp]
    L.sph_stream.restype = vp
    L.sph_get_grid.argtypes = [vp, vp, vp]
    L.sph_grid_x_subdivision.argtypes = [vp]
    L.sph_save_state.argtypes = [vp, C.c_char_p]
    L.sph_load_state.argtypes = [vp, C.c_char_p]
    L.sph_host_register.argtypes = [vp, C.c_size_t]
    L.sph_host_unregister.argtypes = [vp]
    L.sph_upload_state_begin.argtypes = [vp, u32, vp, vp]
    L.sph_upload_state_commit.argtypes = [vp]
    L.sph_download_begin.argtypes = [vp, C.c_int, vp, C.c_size_t]
    L.sph_download_wait.argtypes = [vp]
    L.sph_comm_id_bytes.restype = C.c_size_t
    L.sph_comm_get_id.argtypes = [vp, C.c_size_t]
    L.sph_comm_init.argtypes = [vp, C.c_int, C.c_int, vp, C.c_size_t]
    L.sph_comm_set_planes.argtypes = [vp, vp]
    L.sph_upload_owned.argtypes = [vp, u32, vp, vp, vp]
    L.sph_download_owned.argtypes = [vp, C.c_int, vp, vp, C.c_size_t, C.POINTER(u32)]
    L.sph_upload_owned_begin.argtypes = [vp, u32, vp, vp, vp]
    L.sph_download_owned_begin.argtypes = [vp, C.c_int, vp, vp, C.c_size_t, C.POINTER(u32)]
    L.sph_comm_stats.argtypes = [vp, vp]
    L.sph_download_owned_scatter.argtypes = [vp, C.c_int, vp, C.c_size_t, C.POINTER(u32)]
    L.sph_comm_get_layers.argtypes = [vp, vp]
    L.sph_comm_rebalance.argtypes = [vp, u32, vp, vp, C.c_size_t, C.POINTER(C.c_int)]
    L.sph_slab_link_ok.argtypes = [vp, vp]
    L.sph_slab_balance_layers.argtypes = [vp, C.c_int32, C.c_int32, vp, u32, C.c_uint64, vp]
    _lib = L
    return L


def default_params(**kw):
    p = SphParams()
    load_library().sph_default_params(C.byref(p))
    return _update_params(p, kw)


def _update_params(p, kw):
    for k, v in kw.items():
        if k == "bound":
            p.bound[:] = [float(x) for x in v]
        elif k == "gravity":
            p.gravity = int(bool(v))
        else:
            if not hasattr(p, k):
                raise KeyError(k)
            setattr(p, k, float(v))
    return p


def _ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


class FluidSimulation:
    """Thin Python face of the C ABI, named after the reference class it stands in for
    (Physics::Fluid::FluidSimulation, engine/physics/physicsWorld.h:29-153)."""

    def __init__(self, capacity, device=0, table_mode=TABLE_GRID, **params):
        self.L = load_library()
        self.h = C.c_void_p()
        rc = self.L.sph_create(C.byref(self.h), int(device), int(capacity))
        if rc != SPH_OK:
            msg = self.L.sph_last_error(None)
            self.h = None
            raise SphError("sph_create failed (%d): %s" % (rc, msg.decode() if msg else "?"))
        self.capacity = int(capacity)
        self.set_table_mode(table_mode)
        if params:
            self.set_params(**params)

    # -- plumbing
    def _check(self, rc):
        if rc != SPH_OK:
            msg = self.L.sph_last_error(self.h)
            raise SphError("libsph_b200 error %d: %s" % (rc, msg.decode() if msg else "?"))

    def close(self):
        if getattr(self, "h", None):
            self.L.sph_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- parameters
    def get_params(self):
        p = SphParams()
        self._check(self.L.sph_get_params(self.h, C.byref(p)))
        return p

    def set_params(self, **kw):
        p = _update_params(self.get_params(), kw)
        self._check(self.L.sph_set_params(self.h, C.byref(p)))

    def set_extras(self, bound_rotation=(0.0, 0.0, 0.0, 1.0), stick_strength=0.0, stick_distance=0.0):
        e = SphExtras()
        e.bound_rotation[:] = [float(x) for x in bound_rotation]
        e.stick_strength, e.stick_distance = float(stick_strength), float(stick_distance)
        self._check(self.L.sph_set_extras(self.h, C.byref(e)))

    def get_extras(self):
        e = SphExtras()
        self._check(self.L.sph_get_extras(self.h, C.byref(e)))
        return dict(bound_rotation=tuple(e.bound_rotation), stick_strength=e.stick_strength, stick_distance=e.stick_distance)

    def set_table_mode(self, mode):
        self._check(self.L.sph_set_table_mode(self.h, int(mode)))

    def set_stage_timing(self, on):
        self._check(self.L.sph_set_stage_timing(self.h, int(on)))

    def set_neighbour_count_tap(self, on):
        self._check(self.L.sph_set_neighbour_count_tap(self.h, int(on)))

    def set_neighbour_list_capacity(self, k):
        self._check(self.L.sph_set_neighbour_list_capacity(self.h, int(k)))

    # -- state
    @property
    def n(self):
        return int(self.L.sph_num_particles(self.h))

    def spawn_grid(self, n):                       # InitializeData
        self._check(self.L.sph_spawn_grid(self.h, int(n)))

    def spawn_block(self, nx, ny, nz, gap, origin, jitter_amp=0.0, velocity_scale=0.0, seed=0):
        """Device-side lattice block (scenes.device_spawn_args builds the arguments of a named scene)."""
        b = SphBlockSpawn(int(nx), int(ny), int(nz), 0, float(gap), (C.c_double * 3)(*[float(x) for x in origin]),
                          float(jitter_amp), float(velocity_scale), int(seed) & 0xFFFFFFFFFFFFFFFF)
        self._check(self.L.sph_spawn_block(self.h, C.byref(b)))

    def upload_state(self, pos, vel=None):
        pos = np.ascontiguousarray(pos, dtype=np.float32).reshape(-1, 3)
        if vel is not None:
            vel = np.ascontiguousarray(vel, dtype=np.float32).reshape(-1, 3)
            assert vel.shape == pos.shape
        self._check(self.L.sph_upload_state(self.h, pos.shape[0], _ptr(pos), _ptr(vel)))

    def upload_state_ptr(self, n, pos_ptr, vel_ptr):
        """Raw host pointers (e.g. pinned torch tensors' data_ptr())."""
        self._check(self.L.sph_upload_state(self.h, int(n), C.c_void_p(pos_ptr), C.c_void_p(vel_ptr) if vel_ptr else None))

    # -- pipelined transfers (raw host pointers; the buffers must outlive the copies)
    def upload_state_begin(self, n, pos_ptr, vel_ptr=None):
        self._check(self.L.sph_upload_state_begin(self.h, int(n), C.c_void_p(pos_ptr), C.c_void_p(vel_ptr) if vel_ptr else None))

    def upload_state_commit(self):
        self._check(self.L.sph_upload_state_commit(self.h))

    def download_begin(self, field, host_ptr, nbytes):
        fid = FIELDS[field] if isinstance(field, str) else int(field)
        self._check(self.L.sph_download_begin(self.h, fid, C.c_void_p(host_ptr), int(nbytes)))

    def download_wait(self):
        self._check(self.L.sph_download_wait(self.h))

    # -- hot path
    def step(self, dt):                            # Update(dt)
        self._check(self.L.sph_step(self.h, float(dt)))

    def step_n(self, dt, nsteps):
        self._check(self.L.sph_step_n(self.h, float(dt), int(nsteps)))

    def refresh_densities(self):
        self._check(self.L.sph_refresh_densities(self.h))

    def synchronize(self):
        self._check(self.L.sph_synchronize(self.h))

    # -- read-back
    def download(self, field, out=None):
        fid = FIELDS[field] if isinstance(field, str) else int(field)
        comps, dt = _FIELD_SHAPE[fid]
        n = self.n
        shape = (n, comps) if comps else (n,)
        if out is None:
            out = np.empty(shape, dtype=dt)
        assert out.dtype == dt and out.size == int(np.prod(shape)) and out.flags.c_contiguous
        self._check(self.L.sph_download(self.h, fid, _ptr(out), out.nbytes))
        return out

    def download_ptr(self, field, host_ptr, nbytes):
        fid = FIELDS[field] if isinstance(field, str) else int(field)
        self._check(self.L.sph_download(self.h, fid, C.c_void_p(host_ptr), int(nbytes)))

    def download_table(self, table):
        tid = TABLES[table] if isinstance(table, str) else int(table)
        ln = C.c_size_t(0)
        self._check(self.L.sph_download_table(self.h, tid, None, 0, C.byref(ln)))
        out = np.empty(ln.value, dtype=np.uint32)
        self._check(self.L.sph_download_table(self.h, tid, _ptr(out), out.nbytes, C.byref(ln)))
        return out

    def get_particle(self, index):
        out = np.zeros(10, np.float32)
        self._check(self.L.sph_get_particle(self.h, int(index) & 0xFFFFFFFF, _ptr(out)))
        return out

    def timings(self):
        out = np.zeros(6, np.float64)
        self._check(self.L.sph_get_timings(self.h, _ptr(out)))
        return out

    def launch_count(self):
        return int(self.L.sph_launch_count(self.h))

    def set_graph_replay(self, on):
        self._check(self.L.sph_set_graph_replay(self.h, 1 if on else 0))

    def noncanonical_cells(self):
        return int(self.L.sph_noncanonical_cells(self.h))

    def density_stack_rows(self):
        return int(self.L.sph_density_stack_rows(self.h))

    def graph_replays(self):
        return int(self.L.sph_graph_replays(self.h))

    def stream_ptr(self):
        return int(self.L.sph_stream(self.h) or 0)

    def grid_x_subdivision(self):
        return int(self.L.sph_grid_x_subdivision(self.h))

    def grid(self):
        d = np.zeros(3, np.int32)
        o = np.zeros(3, np.int32)
        self._check(self.L.sph_get_grid(self.h, _ptr(d), _ptr(o)))
        return d, o

    def save_state(self, path):
        self._check(self.L.sph_save_state(self.h, os.fsencode(path)))

    def load_state(self, path):
        self._check(self.L.sph_load_state(self.h, os.fsencode(path)))

    # convenience mirrors of the reference getters
    def positions(self): return self.download("positions")
    def out_positions(self): return self.download("out_positions")
    def velocities(self): return self.download("velocities")
    def densities(self): return self.download("densities")
