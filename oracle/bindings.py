"""oracle/bindings.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes bindings for the two CPU checkers:

* ``PortOracle``  -> oracle/libsph_oracle.so   (plain-C restatement, oracle/sph_oracle.c)
* ``RefOracle``   -> oracle/_ref/libsph_ref.so (the UNMODIFIED reference TU
  engine/physics/physicsWorld.cc + oracle/ref_harness.cc; process-wide singleton
  because the reference class is a Meyers singleton, physicsWorld.cc:32-37)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, "libsph_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libsph_ref.so")
REF_PAR_SO = os.path.join(HERE, "_ref", "libsph_ref_par.so")      # same TU, std::execution::par on every host thread

DEFAULTS = dict(interaction_radius=0.35, sqr_radius=float(np.float32(0.35) * np.float32(0.35)),
                target_density=99.7, pressure_multiplier=300.0, near_pressure_multiplier=20.0,
                viscosity_strength=0.5, gravity_scale=10.0, gravity=0, bound=(20.0, 20.0, 20.0))


class OracleParams(C.Structure):
    _fields_ = [("interaction_radius", C.c_float), ("sqr_radius", C.c_float),
                ("target_density", C.c_float), ("pressure_multiplier", C.c_float),
                ("near_pressure_multiplier", C.c_float), ("viscosity_strength", C.c_float),
                ("gravity_scale", C.c_float), ("gravity", C.c_int), ("bound", C.c_float * 3)]


class RefParams(C.Structure):
    _fields_ = [("interaction_radius", C.c_float),
                ("target_density", C.c_float), ("pressure_multiplier", C.c_float),
                ("near_pressure_multiplier", C.c_float), ("viscosity_strength", C.c_float),
                ("gravity_scale", C.c_float), ("gravity", C.c_int), ("bound", C.c_float * 3)]


def _fill(struct, kw):
    d = dict(DEFAULTS)
    d.update(kw)
    for name, _ in struct._fields_:
        if name == "bound":
            struct.bound[:] = [float(x) for x in d["bound"]]
        elif name == "gravity":
            struct.gravity = int(bool(d["gravity"]))
        else:
            setattr(struct, name, float(d[name]))
    return struct


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def have_port():
    return os.path.exists(PORT_SO)


def have_ref():
    return os.path.exists(REF_SO)


def have_ref_par():
    return os.path.exists(REF_PAR_SO)


class PortOracle:
    """The C restatement.  API mirrors RefOracle so tests can swap them."""
    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            L = C.CDLL(PORT_SO)
            L.oracle_create.restype = C.c_void_p
            L.oracle_create.argtypes = [C.c_int]
            for name in ("oracle_destroy", "oracle_spawn_grid", "oracle_stage_density"):
                getattr(L, name).argtypes = [C.c_void_p]
            L.oracle_set_params.argtypes = [C.c_void_p, C.POINTER(OracleParams)]
            L.oracle_set_wide_lookup.argtypes = [C.c_void_p, C.c_int]
            L.oracle_set_threads.argtypes = [C.c_int]
            L.oracle_set_state.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
            L.oracle_set_predicted.argtypes = [C.c_void_p, C.c_void_p]
            L.oracle_set_densities.argtypes = [C.c_void_p, C.c_void_p]
            L.oracle_stage_predict.argtypes = [C.c_void_p, C.c_float]
            L.oracle_stage_spatial.argtypes = [C.c_void_p, C.c_void_p]
            L.oracle_stage_pressure.argtypes = [C.c_void_p, C.c_float]
            L.oracle_stage_viscosity.argtypes = [C.c_void_p, C.c_float, C.c_int]
            L.oracle_stage_integrate.argtypes = [C.c_void_p, C.c_float]
            L.oracle_step.argtypes = [C.c_void_p, C.c_float, C.c_int]
            L.oracle_num_particles.argtypes = [C.c_void_p]
            for name in ("positions", "out_positions", "velocities", "predicted", "densities",
                         "vel_after_pressure", "vel_after_viscosity", "start_indices",
                         "neighbour_counts", "timings"):
                getattr(L, "oracle_get_" + name).argtypes = [C.c_void_p, C.c_void_p]
            L.oracle_get_hash_key.argtypes = [C.c_void_p] + [C.c_void_p] * 3
            L.oracle_get_sorted.argtypes = [C.c_void_p] + [C.c_void_p] * 3
            L.oracle_get_force_scales.argtypes = [C.c_void_p, C.c_float, C.c_void_p, C.c_void_p]
            L.oracle_kernels.argtypes = [C.c_float, C.c_float, C.c_void_p]
            cls._lib = L
        return cls._lib

    kind = "port"

    def __init__(self, n, threads=1, wide=False, **params):
        self.L = self.lib()
        self.n = int(n)
        self.h = C.c_void_p(self.L.oracle_create(self.n))
        self.L.oracle_set_threads(int(threads))
        self.L.oracle_set_wide_lookup(self.h, int(wide))
        self.params = dict(DEFAULTS)
        self.set_params(**params)

    def close(self):
        if self.h:
            self.L.oracle_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_params(self, **kw):
        self.params.update(kw)
        s = _fill(OracleParams(), self.params)
        self.L.oracle_set_params(self.h, C.byref(s))

    def set_threads(self, t):
        self.L.oracle_set_threads(int(t))

    def spawn_grid(self):
        self.L.oracle_spawn_grid(self.h)

    def set_state(self, pos=None, vel=None):
        pos = None if pos is None else _f32(pos).reshape(self.n, 3)
        vel = None if vel is None else _f32(vel).reshape(self.n, 3)
        self.L.oracle_set_state(self.h, None if pos is None else _p(pos), None if vel is None else _p(vel))

    def set_predicted(self, pred):
        pred = _f32(pred).reshape(self.n, 3)
        self.L.oracle_set_predicted(self.h, _p(pred))

    def set_densities(self, dens):
        dens = _f32(dens).reshape(self.n, 2)
        self.L.oracle_set_densities(self.h, _p(dens))

    def stage_predict(self, dt): self.L.oracle_stage_predict(self.h, dt)

    def stage_spatial(self, forced_order=None):
        if forced_order is None:
            self.L.oracle_stage_spatial(self.h, None)
        else:
            o = np.ascontiguousarray(forced_order, dtype=np.uint32)
            assert o.shape == (self.n,)
            self.L.oracle_stage_spatial(self.h, _p(o))

    def stage_density(self): self.L.oracle_stage_density(self.h)
    def stage_pressure(self, dt): self.L.oracle_stage_pressure(self.h, dt)
    def stage_viscosity(self, dt, jacobi=True): self.L.oracle_stage_viscosity(self.h, dt, int(jacobi))
    def stage_integrate(self, dt): self.L.oracle_stage_integrate(self.h, dt)
    def step(self, dt, jacobi=True): self.L.oracle_step(self.h, dt, int(jacobi))

    def _get(self, name, shape, dtype=np.float32):
        out = np.empty(shape, dtype=dtype)
        getattr(self.L, "oracle_get_" + name)(self.h, _p(out))
        return out

    def positions(self): return self._get("positions", (self.n, 3))
    def out_positions(self): return self._get("out_positions", (self.n, 4))
    def velocities(self): return self._get("velocities", (self.n, 3))
    def predicted(self): return self._get("predicted", (self.n, 3))
    def densities(self): return self._get("densities", (self.n, 2))
    def vel_after_pressure(self): return self._get("vel_after_pressure", (self.n, 3))
    def vel_after_viscosity(self): return self._get("vel_after_viscosity", (self.n, 3))
    def start_indices(self): return self._get("start_indices", (self.n,), np.uint32)
    def neighbour_counts(self): return self._get("neighbour_counts", (self.n,), np.uint32)
    def timings(self): return self._get("timings", (6,), np.float64)

    def hash_key(self):
        h = np.empty(self.n, np.uint32); k = np.empty(self.n, np.uint32); c = np.empty((self.n, 3), np.int32)
        self.L.oracle_get_hash_key(self.h, _p(h), _p(k), _p(c))
        return h, k, c

    def sorted_lookup(self):
        """(particle index, hash as the reference stores it, key) in sorted sequence."""
        i = np.empty(self.n, np.uint32); h = np.empty(self.n, np.uint32); k = np.empty(self.n, np.uint32)
        self.L.oracle_get_sorted(self.h, _p(i), _p(h), _p(k))
        return i, h, k

    def force_scales(self, dt):
        ps = np.empty(self.n, np.float32); vs = np.empty(self.n, np.float32)
        self.L.oracle_get_force_scales(self.h, dt, _p(ps), _p(vs))
        return ps, vs

    @classmethod
    def kernels(cls, dist, radius):
        out = np.empty(5, np.float32)
        cls.lib().oracle_kernels(dist, radius, _p(out))
        return out


class RefOracle:
    """The unmodified reference class behind oracle/ref_harness.cc.  ONE per process."""
    _lib = None
    kind = "reference"

    SO = REF_SO

    @classmethod
    def lib(cls):
        if cls._lib is None:
            L = C.CDLL(cls.SO)
            L.ref_initialize_data.argtypes = [C.c_int]
            L.ref_allocate.argtypes = [C.c_int]
            L.ref_set_params.argtypes = [C.POINTER(RefParams)]
            L.ref_get_params.argtypes = [C.POINTER(RefParams)]
            L.ref_get_sqr_radius.restype = C.c_float
            L.ref_set_state.argtypes = [C.c_void_p, C.c_void_p]
            L.ref_update.argtypes = [C.c_float]
            L.ref_stage_predict.argtypes = [C.c_float]
            L.ref_stage_pressure.argtypes = [C.c_float]
            L.ref_stage_viscosity.argtypes = [C.c_float, C.c_int]
            L.ref_stage_integrate.argtypes = [C.c_float]
            L.ref_step_staged.argtypes = [C.c_float, C.c_int]
            for name in ("positions", "out_positions", "velocities", "predicted", "densities",
                         "vel_after_pressure", "vel_after_viscosity", "lookup_raw", "start_indices",
                         "neighbour_counts", "timings"):
                getattr(L, "ref_get_" + name).argtypes = [C.c_void_p]
            L.ref_get_hash_key.argtypes = [C.c_void_p] * 3
            L.ref_getter_probe.argtypes = [C.c_uint32, C.c_void_p]
            L.ref_kernels.argtypes = [C.c_float, C.c_float, C.c_void_p]
            cls._lib = L
        return cls._lib

    def __init__(self, n, spawn=False, **params):
        self.L = self.lib()
        self.n = int(n)
        if spawn:
            self.L.ref_initialize_data(self.n)
        else:
            self.L.ref_allocate(self.n)
        self.params = dict(DEFAULTS)
        self.set_params(**params)

    def close(self):
        pass

    def set_params(self, **kw):
        self.params.update(kw)
        if abs(self.params["sqr_radius"] - DEFAULTS["sqr_radius"]) > 0:
            raise ValueError("the reference's sqrRadius is a const member (physicsWorld.h:96)")
        s = _fill(RefParams(), self.params)
        self.L.ref_set_params(C.byref(s))

    def spawn_grid(self):
        self.L.ref_initialize_data(self.n)

    def set_state(self, pos=None, vel=None):
        pos = None if pos is None else _f32(pos).reshape(self.n, 3)
        vel = None if vel is None else _f32(vel).reshape(self.n, 3)
        self.L.ref_set_state(None if pos is None else _p(pos), None if vel is None else _p(vel))

    def update(self, dt): self.L.ref_update(dt)                    # verbatim Update()
    def stage_predict(self, dt): self.L.ref_stage_predict(dt)
    def stage_spatial(self): self.L.ref_stage_spatial()
    def stage_density(self): self.L.ref_stage_density()
    def stage_pressure(self, dt): self.L.ref_stage_pressure(dt)
    def stage_viscosity(self, dt, jacobi=True): self.L.ref_stage_viscosity(dt, int(jacobi))
    def stage_integrate(self, dt): self.L.ref_stage_integrate(dt)
    def step(self, dt, jacobi=True): self.L.ref_step_staged(dt, int(jacobi))

    def _get(self, name, shape, dtype=np.float32):
        out = np.empty(shape, dtype=dtype)
        getattr(self.L, "ref_get_" + name)(_p(out))
        return out

    def positions(self): return self._get("positions", (self.n, 3))
    def out_positions(self): return self._get("out_positions", (self.n, 4))
    def velocities(self): return self._get("velocities", (self.n, 3))
    def predicted(self): return self._get("predicted", (self.n, 3))
    def densities(self): return self._get("densities", (self.n, 2))
    def vel_after_pressure(self): return self._get("vel_after_pressure", (self.n, 3))
    def vel_after_viscosity(self): return self._get("vel_after_viscosity", (self.n, 3))
    def start_indices(self): return self._get("start_indices", (self.n,), np.uint32)
    def neighbour_counts(self): return self._get("neighbour_counts", (self.n,), np.uint32)
    def timings(self): return self._get("timings", (6,), np.float64)

    def hash_key(self):
        h = np.empty(self.n, np.uint32); k = np.empty(self.n, np.uint32); c = np.empty((self.n, 3), np.int32)
        self.L.ref_get_hash_key(_p(h), _p(k), _p(c))
        return h, k, c

    def sorted_lookup(self):
        raw = self._get("lookup_raw", (self.n, 3))
        return (raw[:, 0].astype(np.uint32), raw[:, 1].astype(np.int64).astype(np.uint32),
                raw[:, 2].astype(np.uint32))

    def getter_probe(self, i):
        out = np.empty(10, np.float32)
        self.L.ref_getter_probe(int(i), _p(out))
        return out

    @classmethod
    def kernels(cls, dist, radius):
        out = np.empty(5, np.float32)
        cls.lib().ref_kernels(dist, radius, _p(out))
        return out


def fnv1a64(*arrays):
    """FNV-1a over the raw bytes of the arrays, as in SURVEY.md 8(c)."""
    h = 0xcbf29ce484222325
    for a in arrays:
        for b in np.ascontiguousarray(a).tobytes():
            h = ((h ^ b) * 0x100000001b3) & 0xFFFFFFFFFFFFFFFF
    return "%016x" % h


class RefOracleParallel(RefOracle):
    """The same unmodified reference TU built against oracle/pstl_threads: its std::execution::par loops run on every
    host thread (OpenMP).  For TIMING (bench.py --impl reference) and race-free stages; the in-place viscosity update
    of Update() is a data race under a parallel backend, so values come from RefOracle."""
    _lib = None
    SO = REF_PAR_SO
    kind = "reference"

    @staticmethod
    def _omp():
        import ctypes.util
        return C.CDLL(ctypes.util.find_library("gomp") or "libgomp.so.1")

    @classmethod
    def threads(cls):
        try:
            return int(cls._omp().omp_get_max_threads())
        except OSError:
            return os.cpu_count() or 1

    @classmethod
    def set_threads(cls, n):
        """(torchrun exports OMP_NUM_THREADS=1 to every rank; the reference arm runs on rank 0 alone and takes the box)"""
        try:
            cls._omp().omp_set_num_threads(int(max(1, n)))
        except OSError:
            pass
        return cls.threads()

