"""Host-side driver of the slab-decomposed multi-GPU step (one process per GPU, torch.distributed for
the rendezvous only; every data-path byte moves inside libsph_b200.so over NCCL).

  planes / layers : the box is cut along z at CELL-LAYER granularity; rank k owns the particles whose
                    predicted position lies in global grid layers [L_k, L_k+1)
  SlabSimulation  : per-rank wrapper -- comm init (ncclUniqueId broadcast through torch.distributed),
                    upload of the rank's particles with their global ids, step, gather for parity
"""
import ctypes as C
import json
import os
import time

import numpy as np


def layer_of(z, r, gmin_z, gz):
    """Global z layer of a coordinate, with the device's arithmetic: floor(float32(z) / float32(r))."""
    q = np.floor(np.asarray(z, np.float32) / np.float32(r)).astype(np.int64) - int(gmin_z)
    return np.clip(q, 0, int(gz) - 1)


def choose_layers(z_sample, nranks, r, gmin_z, gz):
    """Cut layers [0, gz) into nranks contiguous groups with ~equal particle counts (>= 3 layers each)."""
    lay = layer_of(z_sample, r, gmin_z, gz)
    hist = np.bincount(lay, minlength=int(gz)).astype(np.float64)
    cum = np.concatenate([[0.0], np.cumsum(hist)])
    total = cum[-1]
    L = [0]
    for k in range(1, nranks):
        target = total * k / nranks
        cut = int(np.searchsorted(cum, target, side="left"))
        if cut > 0 and abs(cum[cut - 1] - target) <= abs(cum[min(cut, gz)] - target):
            cut -= 1
        cut = max(cut, L[-1] + 3)
        cut = min(cut, int(gz) - 3 * (nranks - k))
        L.append(cut)
    L.append(int(gz))
    for k in range(nranks):
        if L[k + 1] - L[k] < 3:
            raise ValueError("cannot give every rank three cell layers: %r" % (L,))
    return L


def balance_layers(pkg, hist, nranks, layers_old=None, max_shift=1, row_budget=0):
    """sph_slab_balance_layers (pure host function of libsph_b200.so): cut the per-layer particle histogram at the
    particle-count quantiles, optionally constrained to stay near ``layers_old`` (re-balancing)."""
    hist = np.ascontiguousarray(hist, np.uint32)
    old = None if layers_old is None else np.ascontiguousarray(layers_old, np.int32)
    new = np.zeros(nranks + 1, np.int32)
    rc = pkg.load_library().sph_slab_balance_layers(C.c_void_p(hist.ctypes.data), hist.size, nranks,
                                                    None if old is None else C.c_void_p(old.ctypes.data), int(max_shift),
                                                    int(row_budget), C.c_void_p(new.ctypes.data))
    if rc != 0:
        raise ValueError("sph_slab_balance_layers: cannot cut %d layers for %d ranks (rc %d)" % (hist.size, nranks, rc))
    return [int(x) for x in new]


def planes_from_layers(L, r, gmin_z):
    """z value inside the first layer of each slab: the library floors it back to the same layer."""
    return np.array([(l + gmin_z + 0.5) * r for l in L], dtype=np.float32)


def owner_of(z, L, r, gmin_z, gz):
    lay = layer_of(z, r, gmin_z, gz)
    return np.searchsorted(np.asarray(L[1:-1]), lay, side="right")


class SlabSimulation:
    """One rank of the slab-decomposed solver."""

    def __init__(self, pkg, capacity, rank, nranks, device, id_bytes, **params):
        self.pkg, self.rank, self.nranks = pkg, rank, nranks
        self.sim = pkg.FluidSimulation(capacity, device=device, table_mode=pkg.TABLE_GRID, **params)
        L = self.sim.L
        buf = (C.c_ubyte * len(id_bytes)).from_buffer_copy(id_bytes)
        self.sim._check(L.sph_comm_init(self.sim.h, rank, nranks, buf, len(id_bytes)))
        self.dims, self.origin = self.sim.grid()
        self.r = float(self.sim.get_params().interaction_radius)

    @staticmethod
    def make_id(pkg):
        L = pkg.load_library()
        nb = int(L.sph_comm_id_bytes())
        buf = (C.c_ubyte * nb)()
        rc = L.sph_comm_get_id(buf, nb)
        if rc != 0:
            raise pkg.SphError("sph_comm_get_id failed: %d" % rc)
        return bytes(buf)

    def set_layers(self, layers):
        self.layers = list(layers)
        planes = planes_from_layers(self.layers, self.r, int(self.origin[2]))
        self.sim._check(self.sim.L.sph_comm_set_planes(self.sim.h, C.c_void_p(planes.ctypes.data)))

    def upload_owned(self, ids, pos, vel=None):
        ids = np.ascontiguousarray(ids, np.uint32)
        pos = np.ascontiguousarray(pos, np.float32).reshape(-1, 3)
        vp = None
        if vel is not None:
            vel = np.ascontiguousarray(vel, np.float32).reshape(-1, 3)
            vp = C.c_void_p(vel.ctypes.data)
        self.sim._check(self.sim.L.sph_upload_owned(self.sim.h, ids.size, C.c_void_p(ids.ctypes.data),
                                                    C.c_void_p(pos.ctypes.data), vp))

    def step(self, dt):
        self.sim.step(dt)

    def download_owned(self, field):
        fid = self.pkg.FIELDS[field]
        comps, dt = self.pkg._FIELD_SHAPE[fid]
        n = self.sim.n
        ids = np.empty(n, np.uint32)
        out = np.empty((n, comps) if comps else (n,), dt)
        cnt = C.c_uint32(0)
        self.sim._check(self.sim.L.sph_download_owned(self.sim.h, fid, C.c_void_p(ids.ctypes.data),
                                                      C.c_void_p(out.ctypes.data), out.nbytes, C.byref(cnt)))
        assert cnt.value == n
        return ids, out

    def rebalance(self, max_shift=1):
        """COLLECTIVE: move the slab planes towards the particle-count quantiles (sph_comm_rebalance).  Returns
        (layers now in force, global per-layer histogram, whether any plane moved)."""
        layers = np.zeros(self.nranks + 1, np.int32)
        hist = np.zeros(int(self.dims[2]), np.uint32)
        changed = C.c_int(0)
        self.sim._check(self.sim.L.sph_comm_rebalance(self.sim.h, int(max_shift), C.c_void_p(layers.ctypes.data),
                                                      C.c_void_p(hist.ctypes.data), hist.size, C.byref(changed)))
        self.layers = [int(x) for x in layers]
        return self.layers, hist, bool(changed.value)

    def get_layers(self):
        layers = np.zeros(self.nranks + 1, np.int32)
        self.sim._check(self.sim.L.sph_comm_get_layers(self.sim.h, C.c_void_p(layers.ctypes.data)))
        return [int(x) for x in layers]

    def stats(self):
        out = np.zeros(5, np.uint32)
        self.sim._check(self.sim.L.sph_comm_stats(self.sim.h, C.c_void_p(out.ctypes.data)))
        return dict(owned=int(out[0]), ghosts_lo=int(out[1]), ghosts_hi=int(out[2]), migrated_lo=int(out[3]),
                    migrated_hi=int(out[4]))

    def close(self):
        self.sim.close()


def broadcast_id(pkg, dist, torch, rank, device):
    """ncclUniqueId from rank 0 to everyone through torch.distributed (rendezvous plumbing only)."""
    nb = int(pkg.load_library().sph_comm_id_bytes())
    if dist.get_backend() == "nccl":
        t = torch.zeros(nb, dtype=torch.uint8, device="cuda:%d" % device)
        if rank == 0:
            t.copy_(torch.frombuffer(bytearray(SlabSimulation.make_id(pkg)), dtype=torch.uint8))
        dist.broadcast(t, 0)
        return bytes(t.cpu().numpy().tobytes())
    obj = [SlabSimulation.make_id(pkg) if rank == 0 else None]
    dist.broadcast_object_list(obj, 0)
    return obj[0]


def gather_by_id(dist, n_total, ids, arr):
    """every rank's (ids, rows) assembled into index order on every rank; ownership must be a partition"""
    objs = [None] * dist.get_world_size()
    dist.all_gather_object(objs, (ids, arr))
    out = np.zeros((n_total,) + arr.shape[1:], arr.dtype)
    seen = np.zeros(n_total, np.int32)
    for i, a in objs:
        out[i] = a
        seen[i] += 1
    if not np.all(seen == 1):
        raise AssertionError("ownership is not a partition: %d missing, %d duplicated" % ((seen == 0).sum(), (seen > 1).sum()))
    return out


PARITY_FIELDS = ("neighbour_count", "hash", "densities", "vel_after_pressure", "vel_after_viscosity", "positions", "velocities")


def parity_gate(pkg, scenes, dist, torch, rank, world, dev, n_side=100, seed=0xC2, steps=2):
    """The slab-decomposed step against the single-GPU step on the same state (n_side^3 <= 2^24 particles of the
    dam-break rule; n_side = 100, seed 0xC2 is BASELINE configs[1]): every rank steps its slab, rank 0 steps the whole
    set on its own GPU, rows are matched by particle id.  Hashes and neighbour counts bit-exact; floats within
    1e-5 * (step + 1) of max(|ref|, max|ref| of the field) -- summation order is the only difference between the two.
    Returns the dict bench.py prints as "parity"; raises AssertionError on a mismatch (on every rank)."""
    sc = scenes.small_dam_break(n_side, seed=seed)
    n, dt = sc["n"], scenes.DT
    idb = broadcast_id(pkg, dist, torch, rank, dev)
    slab = SlabSimulation(pkg, int(n / world * 1.5) + 65536, rank, world, dev, idb, **sc["params"])
    gmin_z, gz = int(slab.origin[2]), int(slab.dims[2])
    layers = choose_layers(sc["pos"][:, 2], world, slab.r, gmin_z, gz)
    slab.set_layers(layers)
    own = owner_of(sc["pos"][:, 2], layers, slab.r, gmin_z, gz) == rank
    slab.sim.set_neighbour_count_tap(True)
    slab.upload_owned(np.nonzero(own)[0].astype(np.uint32), sc["pos"][own], sc["vel"][own])
    single = None
    if rank == 0:
        single = pkg.FluidSimulation(n, device=dev, **sc["params"])
        single.set_neighbour_count_tap(True)
        single.upload_state(sc["pos"], sc["vel"])
    worst, migrated, err = {}, 0, None
    for s in range(steps):
        slab.step(dt)
        st = slab.stats()
        migrated += st["migrated_lo"] + st["migrated_hi"]
        fields = {}
        for f in PARITY_FIELDS:
            i, a = slab.download_owned(f)
            fields[f] = gather_by_id(dist, n, i, a)
        if rank == 0:
            single.step(dt)
            try:
                assert np.array_equal(fields["hash"], single.download("hash")), "step %d: cell hashes differ" % s
                nc = single.download("neighbour_count")
                assert np.array_equal(fields["neighbour_count"], nc), "step %d: %d neighbour counts differ" % (
                    s, int((fields["neighbour_count"] != nc).sum()))
                for f in PARITY_FIELDS[2:]:
                    ref = single.download(f).astype(np.float64)
                    got = fields[f].astype(np.float64)
                    floor = 0.0 if f == "densities" else np.abs(ref).max()
                    scale = np.maximum(np.abs(ref), floor + 1e-30)
                    rel = np.abs(got - ref) / scale
                    worst[f] = max(worst.get(f, 0.0), float(rel.max()))
                    assert rel.max() <= 1e-5 * (s + 1), "step %d: %s off by %.3g of its scale" % (s, f, rel.max())
            except AssertionError as e:
                err = str(e)
        flag = torch.tensor([1 if err else 0], device="cuda:%d" % dev)
        dist.all_reduce(flag)
        if int(flag.item()):
            break
    tot = torch.tensor([migrated], device="cuda:%d" % dev)
    dist.all_reduce(tot)
    if single is not None:
        single.close()
    slab.close()
    dist.barrier()
    res = {"ok": err is None, "against": "single-GPU step of the same state on rank 0's GPU, rows matched by particle id",
           "particles": n, "state": "dam-break rule %d^3, seed 0x%X" % (n_side, seed), "steps": steps, "ranks": world,
           "layers": layers, "migrations": int(tot.item()),
           "integers": "cell hash and neighbour count bit-exact",
           "floats_worst_rel": worst, "float_tolerance": "1e-5 * (step + 1) of max(|ref|, max|ref| of the field); densities: of |ref|"}
    if int(flag.item()):
        if rank == 0:
            res["error"] = err
            print(json.dumps({"parity": res}), flush=True)
        raise AssertionError("multi-GPU parity gate failed" + (": " + err if err else " (see rank 0)"))
    return res


def lattice_ids_for_rank(nx, ny, nz, iz_lo, iz_hi):
    """Global ids of the lattice sites with iz in [iz_lo, iz_hi): id = (iy*nx + ix)*nz + iz."""
    iz = np.arange(max(iz_lo, 0), min(iz_hi, nz), dtype=np.uint64)
    base = (np.arange(nx * ny, dtype=np.uint64) * np.uint64(nz))[:, None]
    return (base + iz[None, :]).reshape(-1)


def generate_rank_particles(scenes, name, layers, rank, r, gmin_z, gz):
    """This rank's share of a lattice config without materialising the whole set: generate the z range of
    sites that can fall into the rank's layers (+1 site of slack for the jitter), keep the exact ones."""
    nx, ny, nz, seed = scenes.CONFIGS[name]
    bound = (3 * nx * scenes.GAP0, 1.5 * ny * scenes.GAP0, nz * scenes.GAP0 + scenes.GAP0)
    z0 = -bound[2] / 2 + scenes.GAP0 / 2
    zlo = (layers[rank] + gmin_z) * r
    zhi = (layers[rank + 1] + gmin_z) * r
    iz_lo = int(np.floor((zlo - z0) / scenes.GAP0)) - 1
    iz_hi = int(np.ceil((zhi - z0) / scenes.GAP0)) + 2
    ids = lattice_ids_for_rank(nx, ny, nz, iz_lo, iz_hi)
    pos, vel, _ = scenes.dam_break(nx, ny, nz, seed, ids=ids)
    own = owner_of(pos[:, 2], layers, r, gmin_z, gz) == rank
    return ids[own].astype(np.uint32), pos[own], vel[own], bound


def bind_to_gpu_numa(torch, dev):
    """Pin this process to the CPUs of the NUMA node its GPU hangs off (before any page-locked buffer is allocated: the
    pages are then local to the GPU's root complex).  Returns a short description for the bench line."""
    try:
        pr = torch.cuda.get_device_properties(dev)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bdf).read().strip())
        if node < 0:
            return "GPU %s reports no NUMA node: not bound" % bdf
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return "NUMA node %d of GPU %s has no CPU this process may use: not bound" % (node, bdf)
        os.sched_setaffinity(0, cpus)
        return "bound to the %d CPUs of NUMA node %d (GPU %s)" % (len(cpus), node, bdf)
    except Exception as ex:
        return "not bound (%s: %s)" % (type(ex).__name__, ex)


def bench_multi(args, pkg, scenes, torch, dist, rank, world, dev, METRIC, A_BYTES, measured_peaks, ClockSampler, short_line=None, step_roofline=None):
    name = args.config or "C4_dambreak_64M"
    numa = bind_to_gpu_numa(torch, dev)
    # (0) parity gate: the slab-decomposed step must equal the single-GPU step before anything is timed
    parity = None
    if not getattr(args, "no_parity", False):
        parity = parity_gate(pkg, scenes, dist, torch, rank, world, dev)       # raises (exit code != 0) on a mismatch
    # (0b) the same workload on ONE GPU (rank 0's), the anchor of the strong-scaling curve
    strong_base = None
    if short_line is not None and not getattr(args, "no_configs", False):
        if rank == 0:
            try:
                q = short_line(pkg, scenes, torch, dev, name, steps=5, warmup=3,
                               flush=None if args.no_flush else torch.empty(256 << 20, dtype=torch.uint8, device="cuda:%d" % dev))
                strong_base = {"workload": name + " on one GPU (rank 0, before the multi-GPU run)", "ms_per_step": q["ms_per_step"],
                               "value": q["value"], "unit": "M updates/s", "stage_ms": q["stage_ms"]}
            except Exception as ex:
                strong_base = {"error": "%s: %s" % (type(ex).__name__, ex)}
            torch.cuda.empty_cache()
        dist.barrier()
    nx, ny, nz, seed = scenes.CONFIGS[name]
    n_total = nx * ny * nz
    bound = (3 * nx * scenes.GAP0, 1.5 * ny * scenes.GAP0, nz * scenes.GAP0 + scenes.GAP0)
    params = dict(gravity=1, viscosity_strength=0.5, bound=bound)
    cap = int(n_total / world * 1.25) + (1 << 20)
    idb = broadcast_id(pkg, dist, torch, rank, dev)
    slab = SlabSimulation(pkg, cap, rank, world, dev, idb, **params)
    gmin_z, gz = int(slab.origin[2]), int(slab.dims[2])
    # lattice: equal z extents hold equal counts; sample one (x,y) column to place the cuts
    col = np.arange(nz, dtype=np.uint64)
    zs, _, _ = scenes.dam_break(nx, ny, nz, seed, ids=col)
    layers = choose_layers(zs[:, 2], world, slab.r, gmin_z, gz)
    slab.set_layers(layers)
    ids, pos, vel, _ = generate_rank_particles(scenes, name, layers, rank, slab.r, gmin_z, gz)
    slab.upload_owned(ids, pos, vel)
    del pos, vel, ids
    dt = scenes.DT
    stream = torch.cuda.ExternalStream(slab.sim.stream_ptr(), device=dev)
    flush = None if args.no_flush else torch.empty(256 << 20, dtype=torch.uint8, device="cuda:%d" % dev)
    rebalance_every = int(getattr(args, "rebalance", 0) or 0)
    plane_moves = 0
    for k in range(args.warmup):
        slab.step(dt)
        if rebalance_every and (k + 1) % rebalance_every == 0:
            plane_moves += int(slab.rebalance(1)[2])
    slab.sim.synchronize()
    torch.cuda.synchronize()
    dist.barrier()
    clocks = ClockSampler(dev)
    clocks.start()
    l0 = slab.sim.launch_count()
    stage = np.zeros(6)
    total_ms = 0.0
    for k in range(args.steps):
        if flush is not None:
            with torch.cuda.stream(stream):
                flush.fill_(1)
        slab.sim.synchronize()
        dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        slab.step(dt)
        if rebalance_every and (k + 1) % rebalance_every == 0:      # part of the job: timed
            plane_moves += int(slab.rebalance(1)[2])
        b.record(stream)
        slab.sim.synchronize()
        stage += slab.sim.timings()
        total_ms += a.elapsed_time(b)
    launches = slab.sim.launch_count() - l0
    clk = clocks.stop()
    # max over ranks of the device time of the timed region
    t = torch.tensor([total_ms], dtype=torch.float64, device="cuda:%d" % dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    st = torch.tensor(stage / args.steps, dtype=torch.float64, device="cuda:%d" % dev)
    dist.all_reduce(st, op=dist.ReduceOp.MAX)
    stage = st.cpu().numpy()
    owned = torch.tensor([slab.sim.n], dtype=torch.int64, device="cuda:%d" % dev)
    dist.all_reduce(owned)
    lt = torch.tensor([launches], dtype=torch.int64, device="cuda:%d" % dev)
    dist.all_reduce(lt)
    value = n_total * args.steps / (total_ms * 1e-3) / 1e6

    # end to end: every step each rank re-uploads its owned state from pinned host memory and reads its
    # OutPositions back (ids included), through the slab ABI
    ids_h, pos_h = slab.download_owned("positions")
    _, vel_h = slab.download_owned("velocities")
    n_own = ids_h.size
    tp = torch.from_numpy(pos_h).pin_memory(); tv = torch.from_numpy(vel_h).pin_memory(); ti = torch.from_numpy(ids_h.astype(np.int32)).pin_memory()
    out_h = torch.empty((cap, 4), dtype=torch.float32).pin_memory()
    ids_out = [torch.empty(cap, dtype=torch.int32).pin_memory(), torch.empty(cap, dtype=torch.int32).pin_memory()]   # row k of a download is OutPositions[ids[k]]
    L = slab.sim.L
    e2e_steps = max(20, args.steps)
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        slab.sim._check(L.sph_upload_owned(slab.sim.h, n_own, C.c_void_p(ti.data_ptr()), C.c_void_p(tp.data_ptr()), C.c_void_p(tv.data_ptr())))
        slab.step(dt)
        cnt = C.c_uint32(0)
        slab.sim._check(L.sph_download_owned(slab.sim.h, pkg.FIELDS["out_positions"], C.c_void_p(ids_out[0].data_ptr()), C.c_void_p(out_h.data_ptr()),
                                             cap * 16, C.byref(cnt)))
    e2e_blocking_s = time.perf_counter() - t0
    t = torch.tensor([e2e_blocking_s], dtype=torch.float64, device="cuda:%d" % dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_blocking_s = float(t.item())

    # the same frames through the pipelined calls: each rank's upload of frame k+1 and download of frame k-1 travel
    # on their own copy streams while frame k (halo exchanges included) is computed
    out_h2 = [out_h, torch.empty((cap, 4), dtype=torch.float32).pin_memory()]
    h = slab.sim.h

    def up_begin():
        slab.sim._check(L.sph_upload_owned_begin(h, n_own, C.c_void_p(ti.data_ptr()), C.c_void_p(tp.data_ptr()), C.c_void_p(tv.data_ptr())))

    def pipelined(frames):
        up_begin()
        for k in range(frames):
            slab.sim._check(L.sph_upload_state_commit(h))
            if k + 1 < frames:
                up_begin()
            slab.step(dt)
            if k:
                slab.sim._check(L.sph_download_wait(h))
            c2 = C.c_uint32(0)
            slab.sim._check(L.sph_download_owned_begin(h, pkg.FIELDS["out_positions"], C.c_void_p(ids_out[k & 1].data_ptr()),
                                                       C.c_void_p(out_h2[k & 1].data_ptr()), cap * 16, C.byref(c2)))
        slab.sim._check(L.sph_download_wait(h))
        slab.sim.synchronize()

    e2e_path = ("per rank and frame: sph_upload_owned_begin/sph_upload_state_commit(pinned ids+pos3+vel3) -> sph_step -> "
                "sph_download_owned_begin/sph_download_wait(OUT_POSITIONS, pinned); copies overlap the neighbouring frames' steps")
    try:
        pipelined(2)
        dist.barrier()
        t0 = time.perf_counter()
        pipelined(e2e_steps)
        e2e_s = time.perf_counter() - t0
        ok = 1.0
    except pkg.SphError as ex:
        e2e_s, ok = e2e_blocking_s, 0.0
        e2e_path = "blocking calls (pipelined calls failed on rank %d: %s)" % (rank, ex)
    t = torch.tensor([e2e_s, -ok], dtype=torch.float64, device="cuda:%d" % dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t[0].item())
    if float(t[1].item()) > -1.0:                # some rank fell back: report the blocking number
        e2e_s = e2e_blocking_s
        if ok:
            e2e_path = "blocking calls (pipelined calls failed on another rank)"
    e2e = n_total * e2e_steps / e2e_s / 1e6
    stats = slab.stats()
    slab.close()
    if rank != 0:
        return None
    peak, peak_src = measured_peaks()
    gather = {"density": stage[2], "pressure": stage[3], "viscosity": stage[4]}
    dom = max(gather, key=gather.get)
    achieved = A_BYTES[dom] * (n_total / world) / (gather[dom] * 1e-3) / 1e9
    names = ["predict_key", "spatial", "density", "pressure", "viscosity", "integrate"]
    extra = {}
    if parity is not None:
        extra["parity"] = parity
    if strong_base is not None:
        extra["strong_base"] = strong_base
        if "ms_per_step" in strong_base:
            extra["efficiency_vs_strong_base"] = strong_base["ms_per_step"] / (world * (total_ms / args.steps))
    return {
        **extra,
        "metric": METRIC, "value": value, "unit": "M updates/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "%s: %d particles, slab-decomposed along z over %d ranks (layers %s), ghost halo x3 + "
                               "migration per step over NCCL" % (name, n_total, world, layers),
                   "particles": n_total, "table": "grid", "owned_total_after": int(owned.item()),
                   "l2": "flushed between timed steps" if flush is not None else "not flushed",
                   "rank0_last_step": stats,
                   "rebalance": ({"every": rebalance_every, "plane_moves": plane_moves, "layers_after": slab.layers}
                                 if rebalance_every else "off (planes at the initial particle-count quantiles)")},
        "stage_ms": {k: float(v) for k, v in zip(names, stage)},
        "roofline": {"bound": "hbm", "kernel": "k_" + dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": None, "peak_source": peak_src, "per": "rank (max over ranks)",
                     "algorithmic_bytes_per_particle": A_BYTES[dom], "kernel_ms": float(gather[dom]),
                     "step": (step_roofline(value, peak, world) if step_roofline else
                              {"achieved": A_BYTES["step"] * value * 1e6 / 1e9 / world, "frac": A_BYTES["step"] * value * 1e6 / 1e9 / world / peak})},
        "e2e": {"value": e2e, "unit": "M updates/s", "h2d_bytes_per_step": int(n_total) * 28, "d2h_bytes_per_step": int(n_total) * 20,
                "result": "per rank (global id, OutPositions row) of every owned particle: row k is OutPositions[ids[k]]", "numa": numa,
                "ms_per_step": e2e_s / e2e_steps * 1e3, "steps": e2e_steps,
                "path": e2e_path,
                "blocking": {"value": n_total * e2e_steps / e2e_blocking_s / 1e6, "ms_per_step": e2e_blocking_s / e2e_steps * 1e3,
                             "path": "per rank: sph_upload_owned -> sph_step -> sph_download_owned(OUT_POSITIONS), one stream, serial"}},
        "gpu_launches": int(lt.item()),
        "clocks": clk,
    }
