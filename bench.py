#!/usr/bin/env python
"""bench.py -- M particle-updates/s of the per-timestep SPH update (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config NAME]

N = 1 : config C2 (1 M-particle box-fill dam break, BASELINE.json configs[1]) on one B200.
N > 1 : config C4 (64 M-particle dam break, configs[3]) slab-decomposed along z over N ranks
        (launched by torch.distributed.run, one rank per GPU); total work fixed => "strong".
A "step" is one FluidSimulation::Update(dt) (engine/physics/physicsWorld.cc:39-111) over all particles.

One JSON line on rank 0:
  value     whole-job M particle-updates/s, state resident in HBM, CUDA events on the solver's stream,
            L2 flushed between timed steps (the flush is outside the event pairs)
  e2e       the same metric through the host-facing call sequence with HOST (pinned) buffers, every frame:
            H2D positions+velocities -> sph_step -> D2H OutPositions, through the pipelined calls
            (sph_upload_state_begin/_commit, sph_download_begin/_wait); e2e.blocking = the serial calls
  roofline  dominant kernel: algorithmic bytes / its CUDA-event duration / measured HBM peak
  cpu_baseline  the unmodified reference (oracle/_ref) timed on this box's host, bounded sample
--impl reference times the reference CPU implementation alone (rank 0 only).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402

METRIC = "M particle-updates/s"
# algorithmic (compulsory) bytes per particle-update, SURVEY.md 8(d) / DESIGN.md
A_BYTES = dict(predict_key=72, sort=52, table=12, reorder=68, density=32, pressure=64, viscosity=56,
               integrate=84, step=440)
# The timed step does not write OutPositions (16 of K8's 84 bytes): the export kernel runs with the download, inside the
# e2e loops.  The whole-step roofline therefore counts 424 bytes per update; the contract's 440 is printed beside it.
A_STEP_TIMED = A_BYTES["step"] - 16


def step_roofline(value_m, peak, gpus=1):
    ach = A_STEP_TIMED * value_m * 1e6 / 1e9 / gpus
    return {"achieved": ach, "frac": ach / peak, "algorithmic_bytes_per_particle": A_STEP_TIMED,
            "note": "SURVEY 8(d) A_step = 440 B minus the 16 B OutPositions write, which this implementation does at download time",
            "frac_with_contract_440": A_BYTES["step"] * value_m * 1e6 / 1e9 / gpus / peak}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe): NVML in a thread of this
    process (5 ms period, so even a 10 ms timed region is covered), `nvidia-smi -lms` as the fallback."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []
        self.nvml = None
        self.samples = []          # (sm_mhz, reasons bitmask)
        self.stop_flag = False

    def _nvml_handle(self):
        import pynvml
        pynvml.nvmlInit()
        try:
            import torch
            uuid = str(torch.cuda.get_device_properties(self.gpu).uuid)
            uuid = uuid if uuid.startswith("GPU-") else "GPU-" + uuid
            try:
                return pynvml, pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
            except TypeError:
                return pynvml, pynvml.nvmlDeviceGetHandleByUUID(uuid)
        except Exception:
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.gpu]) if vis and vis.split(",")[self.gpu].isdigit() else self.gpu
            return pynvml, pynvml.nvmlDeviceGetHandleByIndex(idx)

    def start(self):
        try:
            self.nvml, self.h = self._nvml_handle()
            self.sm_max = float(self.nvml.nvmlDeviceGetMaxClockInfo(self.h, self.nvml.NVML_CLOCK_SM))
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _poll(self):
        n = self.nvml
        get_reasons = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self.stop_flag:
            try:
                self.samples.append((float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)), int(get_reasons(self.h))))
            except Exception:
                pass
            time.sleep(0.005)

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.nvml is not None:
            self.stop_flag = True
            self.t.join(timeout=1.0)
            n = self.nvml
            names = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))
            reasons = sorted({nm for _, r in self.samples for nm, bit in names if r & bit})
            sm = [c for c, _ in self.samples]
            return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.sm_max, "reasons": reasons,
                    "samples": len(sm), "source": "nvml"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi"}


def reference_cpu_class(ob, parallel):
    """(oracle class, kind, threads, description) of the reference CPU step to time: the unmodified reference built
    against oracle/pstl_threads (its std::execution::par loops on every host thread) when `parallel` and present, else
    the same TU on libstdc++'s serial PSTL backend (TBB is not installed), else the C restatement"""
    if parallel and ob.have_ref_par():
        t = ob.RefOracleParallel.set_threads(os.cpu_count() or 1)
        return ob.RefOracleParallel, "reference", t, ("unmodified reference Update(), its std::execution::par loops on %d host "
                                                       "threads (libstdc++ parallel backend over oracle/pstl_threads)" % t)
    if ob.have_ref():
        return ob.RefOracle, "reference", 1, "unmodified reference Update() (serial PSTL: TBB absent)"
    return ob.PortOracle, "port", 1, "C restatement, 1 thread"


def cpu_reference_run(scene_name, steps, warmup, budget_s=25.0, parallel=True):
    """Time the reference CPU Update() on a bounded sample of the workload.  Returns a dict."""
    ob = graft.load_oracle()
    pkg = graft.load_package()
    from fluid_simulation_3d_b200 import scenes
    cls, kind, cores, what = reference_cpu_class(ob, parallel)
    # per-particle cost is flat in N (0.13-0.16 M updates/s per core from 10 k to 1 M), so a sub-block
    # of the same lattice / rule is a faithful sample; size it so (steps + warmup) fit the budget
    rate_guess = 0.14e6 * (1.0 if cores == 1 else 0.6 * min(cores, 32))
    total = max(1, steps + warmup)
    side = int(round((budget_s * rate_guess / total) ** (1.0 / 3.0)))
    full = scenes.CONFIGS[scene_name][0] if scene_name in scenes.CONFIGS else 100
    side = max(16, min(side, full))
    sc = scenes.small_dam_break(side, seed=scenes.CONFIGS.get(scene_name, (0, 0, 0, 7))[3])
    orc = cls(sc["n"], **sc["params"])
    orc.set_state(sc["pos"], sc["vel"])
    step = (lambda: orc.update(scenes.DT)) if kind == "reference" else (lambda: orc.step(scenes.DT, jacobi=True))
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return dict(value=sc["n"] * steps / dt / 1e6, unit=METRIC, cores=cores, kind=kind, ms_per_step=dt / steps * 1e3,
                sample="%d^3 = %d-particle corner block of the %s lattice (same spacing, jitter, bounds rule, dt), "
                       "%d steps after %d warm-up, %s" % (side, sc["n"], scene_name, steps, warmup, what),
                n=sc["n"])


def cpu_reference_c1(steps, warmup, parallel=True):
    """C1: the reference's default scene, whole (10 000 particles): InitializeData(10000), gravity on, Update(dt)."""
    ob = graft.load_oracle()
    cls, kind, cores, what = reference_cpu_class(ob, parallel)
    n = 10000
    orc = cls(n, spawn=True, gravity=1) if kind == "reference" else cls(n, gravity=1)
    if kind == "port":
        orc.spawn_grid()
    dt = float(np.float32(0.016667))
    step = (lambda: orc.update(dt)) if kind == "reference" else (lambda: orc.step(dt, jacobi=True))
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    t = time.perf_counter() - t0
    return dict(value=n * steps / t / 1e6, unit=METRIC, cores=cores, kind=kind, ms_per_step=t / steps * 1e3, n=n,
                sample="the whole C1 scene (InitializeData(10000), default bounds, gravity on), %d steps after %d warm-up, %s"
                       % (steps, warmup, what))


def workload_text(name, n, bound, mu):
    """the `config.workload` string of a single-GPU run; the reference arm names its workload with the same words"""
    if name == C1_NAME:
        f = "%s: %d particles, the reference's InitializeData lattice (gap 0.215, centred), bounds %s, r=0.35, gravity on, mu=%.2f, dt=0.016667"
    elif name == C5_NAME:
        f = ("%s: %d particles, jittered lattice gap 0.1216 (~100 neighbours per particle), a 100 x 800 x 100 column centred on the "
             "floor, velocities uniform in [-0.5, 0.5]^3, bounds %s, r=0.35, gravity on, mu=%.2f, dt=0.016667; the dense spawn "
             "state is restored before every timed step (the column explodes within three steps, in the reference too)")
    else:
        f = "%s: %d particles, jittered lattice gap 0.215 in the -x/floor corner, bounds %s, r=0.35, gravity on, mu=%.2f, dt=0.016667"
    return f % (name, n, tuple(round(float(x), 3) for x in bound), mu)


def reference_arm_workload(name, gpus):
    """(workload text, full particle count) of the configuration the other arm runs, without generating it"""
    graft.load_package()
    from fluid_simulation_3d_b200 import scenes
    if name == C1_NAME:
        return workload_text(name, 10000, (20.0, 20.0, 20.0), 0.5), 10000
    if name in scenes.CONFIGS:
        nx, ny, nz, _ = scenes.CONFIGS[name]
        n = nx * ny * nz
        if gpus > 1:
            return "%s: %d particles, slab-decomposed along z over %d ranks" % (name, n, gpus), n
        return workload_text(name, n, (3 * nx * scenes.GAP0, 1.5 * ny * scenes.GAP0, nz * scenes.GAP0 + scenes.GAP0), 0.5), n
    return name, None


def run_reference_arm(args, rank):
    if rank != 0:
        return
    name = "C2_dambreak_1M" if args.gpus == 1 else "C4_dambreak_64M"
    if args.config:
        name = args.config
    r = cpu_reference_c1(args.steps, args.warmup) if name == C1_NAME else cpu_reference_run(name, args.steps, args.warmup, budget_s=120.0)
    text, n_full = reference_arm_workload(name, args.gpus)
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": "M updates/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": text, "particles": n_full, "sample_particles": r["n"],
                       "note": "the reference CPU step on a bounded sample of this workload (see cpu_baseline.sample)"},
            "cpu_baseline": {"value": r["value"], "unit": "M updates/s", "cores": r["cores"], "kind": r["kind"],
                             "sample": r["sample"], "host_cores": os.cpu_count()},
            "e2e": {"value": r["value"], "unit": "M updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default=None)
    ap.add_argument("--table", default="grid", choices=["grid", "refhash"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-flush", action="store_true", help="do not flush L2 between timed steps")
    ap.add_argument("--rebalance", type=int, default=4, metavar="K",
                    help="multi-GPU: sph_comm_rebalance every K steps (inside the timed region); 0 = planes stay where the initial quantile cut put them")
    ap.add_argument("--no-configs", action="store_true", help="N = 1: skip the short C1 / C3 / C5 / C4 lines of the `configs` block")
    ap.add_argument("--no-parity", action="store_true", help="N > 1: skip the slab-vs-single-GPU parity gate")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference_arm(args, rank)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if world != args.gpus:
        if rank == 0:
            sys.stderr.write("bench.py: --gpus %d but WORLD_SIZE %d; using WORLD_SIZE\n" % (args.gpus, world))
    pkg = graft.load_package()
    from fluid_simulation_3d_b200 import scenes

    if world == 1:
        result = bench_single(args, pkg, scenes, torch, local_rank)
    else:
        from fluid_simulation_3d_b200 import slab_driver
        result = slab_driver.bench_multi(args, pkg, scenes, torch, dist, rank, world, local_rank, METRIC, A_BYTES,
                                         measured_peaks, ClockSampler, short_line=short_line, step_roofline=step_roofline)
    if rank == 0 and result is not None:
        print(json.dumps(result), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


C1_NAME = "C1_default_10k"      # BASELINE.json configs[0]: exactly InitializeData(10000), default bounds, gravity on
C5_NAME = "C5_column_8M"        # configs[4]: the high-density column; every timed step starts from the dense spawn state


def c1_scene(pkg, dev):
    """The reference's own default scene: the lattice of InitializeData(10000), positions read back
    from sph_spawn_grid so that the end-to-end loop has host arrays like every other config."""
    n = 10000
    params = dict(gravity=1)
    sim = pkg.FluidSimulation(n, device=dev, **params)
    sim.spawn_grid(n)
    pos = sim.download("positions").copy()
    sim.close()
    return dict(pos=pos, vel=np.zeros_like(pos), bound=(20.0, 20.0, 20.0), n=n, params=params)


def write_snapshot(path, pos, vel, params_struct):
    """the snapshot file format of sph_save_state / FluidSimulation::saveState (include/sph_b200.h): "SPHB2002", u32 n,
    u32 sizeof(SphParams), SphParams, n x pos3, n x vel3"""
    import ctypes
    raw = bytes(ctypes.string_at(ctypes.addressof(params_struct), ctypes.sizeof(params_struct)))
    with open(path, "wb") as f:
        f.write(b"SPHB2002")
        f.write(np.uint32(len(pos)).tobytes())
        f.write(np.uint32(len(raw)).tobytes())
        f.write(raw)
        f.write(np.ascontiguousarray(pos, np.float32).tobytes())
        f.write(np.ascontiguousarray(vel, np.float32).tobytes())


def class_update_e2e(pkg, sc, sim_params, frames, ndev=1, direct=False):
    """The drop-in class timed the way the application drives it: host_demo (C++, fluid-simulation-3d_b200/host) loads the
    scene through FluidSimulation::loadState and calls FluidSimulation::Update(dt) per frame, each call returning with
    the OutPositions mirror refreshed in host memory (fluidSimCPU.cc:35-40,58)."""
    import tempfile
    demo = os.path.join(ROOT, "fluid-simulation-3d_b200", "host", "host_demo")
    if not os.path.exists(demo):
        return {"error": "host_demo is not built"}
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "state.bin")
        write_snapshot(path, sc["pos"], sc["vel"], sim_params)
        cmd = [demo, "0", str(frames), "0", "classbench", path, str(ndev), "1" if direct else "0"]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("classbench ")]
    if r.returncode != 0 or not line:
        return {"error": "host_demo classbench failed: " + r.stdout[-300:]}
    kv = dict(x.split("=", 1) for x in line[0].split()[1:])
    return {"value": float(kv["updates_per_s"]) / 1e6, "unit": "M updates/s", "ms_per_step": float(kv["update_ms"]), "frames": frames,
            "devices": ndev, "path": "FluidSimulation::Update(dt) per frame, OutPositions mirror refreshed by every call (blocking D2H of "
            "16 B/particle inside Update); state resident on the device between frames, like the reference's"}


def short_line(pkg, scenes, torch, dev, name, steps=5, warmup=3, flush=None):
    """A short device-timed line of another BASELINE config (the `configs` block of the default run, and the same-workload
    1-GPU anchor of the multi-GPU curve): the scene is spawned on the device (sph_spawn_grid / sph_spawn_block, bit-identical
    to the host generators), `warmup` untimed steps, `steps` steps each inside its own CUDA-event pair on the solver's
    stream, L2 flushed between them."""
    peak, _ = measured_peaks()
    dt = scenes.DT
    if name == C1_NAME:
        n, bound, mu = 10000, (20.0, 20.0, 20.0), 0.5
        sim = pkg.FluidSimulation(n, device=dev, gravity=1)
        respawn = lambda: sim.spawn_grid(n)
    else:
        meta = scenes.config_meta(name)
        n, bound, mu = meta["n"], meta["bound"], meta["params"].get("viscosity_strength", 0.5)
        sim = pkg.FluidSimulation(n, device=dev, **meta["params"])
        respawn = lambda: sim.spawn_block(**meta["spawn"])
    dense = name == C5_NAME
    if dense:
        sim.set_neighbour_list_capacity(192)
    stream = torch.cuda.ExternalStream(sim.stream_ptr(), device=dev)

    def loop(k, timers):
        acc, total = [], 0.0
        for _ in range(k):
            if dense:
                respawn()
            if flush is not None:
                with torch.cuda.stream(stream):
                    flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            sim.step(dt)
            b.record(stream)
            if timers:
                acc.append(sim.timings())
            sim.synchronize()
            total += a.elapsed_time(b)
        return total, (np.median(np.array(acc), axis=0) if acc else None)

    respawn()
    sim.set_stage_timing(False)
    for _ in range(warmup):                  # the first step is a plain one, the second records the step's graph
        if dense:
            respawn()
        sim.step(dt)
    sim.synchronize()
    total_ms, _ = loop(steps, False)         # value: replayed steps, as in the headline loop
    sim.set_stage_timing(True)
    sim.set_graph_replay(False)
    _, stage = loop(steps, True)             # stage_ms: plain launches with the six stage timers on (medians)
    sim.set_graph_replay(True)
    mean_nb = float(sim.download("neighbour_count").mean()) if n <= (1 << 24) else None
    sim.close()
    names = ["predict_key", "spatial", "density", "pressure", "viscosity", "integrate"]
    gather = {"density": stage[2], "pressure": stage[3], "viscosity": stage[4]}
    dom = max(gather, key=gather.get)
    ach = A_BYTES[dom] * n / (gather[dom] * 1e-3) / 1e9
    value = n * steps / (total_ms * 1e-3) / 1e6
    return {"workload": workload_text(name, n, bound, mu), "particles": n, "value": value, "unit": "M updates/s",
            "ms_per_step": total_ms / steps, "steps": steps, "warmup": warmup,
            "stage_ms": {k: float(v) for k, v in zip(names, stage)},
            "mean_neighbours_last_step": mean_nb,
            "roofline": {"bound": "hbm", "kernel": "k_" + dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                         "algorithmic_bytes_per_particle": A_BYTES[dom], "kernel_ms": float(gather[dom]),
                         "step": step_roofline(value, peak)}}


def bench_single(args, pkg, scenes, torch, dev):
    name = args.config or "C2_dambreak_1M"
    sc = c1_scene(pkg, dev) if name == C1_NAME else scenes.config(name)
    n = sc["n"]
    mode = pkg.TABLE_GRID if args.table == "grid" else pkg.TABLE_REFERENCE_HASH
    sim = pkg.FluidSimulation(n, device=dev, table_mode=mode, **sc["params"])
    stream = torch.cuda.ExternalStream(sim.stream_ptr(), device=dev)
    dt = scenes.DT
    # pinned host buffers of the reference's own layouts: positions vec3, velocity vec3, OutPositions vec4
    h_pos = torch.from_numpy(sc["pos"]).pin_memory()
    h_vel = torch.from_numpy(sc["vel"]).pin_memory()
    h_out = torch.empty((n, 4), dtype=torch.float32).pin_memory()
    sim.upload_state_ptr(n, h_pos.data_ptr(), h_vel.data_ptr())
    flush = None if args.no_flush else torch.empty(256 << 20, dtype=torch.uint8, device="cuda:%d" % dev)

    def l2_flush():
        if flush is not None:
            with torch.cuda.stream(stream):
                flush.fill_(1)

    dense = name == C5_NAME          # the column explodes within three steps: every step starts from the dense spawn state
    if dense:
        sim.set_neighbour_list_capacity(192)

    def restore():
        if dense:
            sim.spawn_block(**sc["spawn"])         # device-side, bit-identical to the host arrays (tests/test_spawn_gpu.py)

    for _ in range(args.warmup):
        restore()
        sim.step(dt)
    sim.synchronize()
    torch.cuda.synchronize()

    # ---- device-resident timing: one CUDA-event pair per step on the solver's stream, L2 flushed between.  sph_step
    # replays the recorded CUDA graph of the step from the second step of an unchanged configuration on (one
    # cudaGraphLaunch instead of a dozen kernel launches; include/sph_b200.h: sph_set_graph_replay).  Loop A times `value`
    # that way, stage timers off; loop B runs the same number of steps through plain launches with the six stage timers on
    # and gives stage_ms (the recording carries no timers: seven event nodes cost more than the replay saves).
    clocks = ClockSampler(dev)
    clocks.start()

    def timed_loop(steps, timers):
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        acc = []
        for a, b in ev:
            restore()                       # outside the event pair
            l2_flush()
            a.record(stream)
            sim.step(dt)
            b.record(stream)
            if timers:
                acc.append(sim.timings())   # waits for the step's last stage event
        sim.synchronize()
        torch.cuda.synchronize()
        return np.array([a.elapsed_time(b) for a, b in ev]), (np.median(np.array(acc), axis=0) if acc else None)

    sim.set_stage_timing(False)
    restore()
    sim.step_n(dt, 3)                        # plain step, recording, first replay: outside the timed region
    sim.synchronize()
    l0, r0 = sim.launch_count(), sim.graph_replays()
    ms, _ = timed_loop(args.steps, False)
    launches = sim.launch_count() - l0 - (args.steps if dense else 0)      # the restoring spawn kernel is not part of the step
    timed_replays = sim.graph_replays() - r0
    total_ms = float(ms.sum())
    value = n * args.steps / (total_ms * 1e-3) / 1e6
    sim.set_stage_timing(True)
    sim.set_graph_replay(False)
    ms_b, stage = timed_loop(args.steps, True)           # medians: a list that has to grow reallocates inside one of these steps
    plain_ms = float(np.median(ms_b))
    sim.set_graph_replay(True)

    # ---- steady state: K steps back to back, no flush (what a simulation loop sees).  The scene is restarted first:
    # the block keeps collapsing (more neighbours per particle every step), and this loop should see the same stretch
    # of the flow as the timed one.
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sim.set_stage_timing(False)
    if not dense:
        sim.upload_state_ptr(n, h_pos.data_ptr(), h_vel.data_ptr())
        sim.step_n(dt, args.warmup)
    steady_steps = max(args.steps, 200) if n <= 200000 else args.steps    # small scenes: enough steps to time
    if dense:
        steady_steps = 2                         # back to back from the dense state: the second step is already a blow-up
        restore()
    sim.step_n(dt, 4 if not dense else 1)        # plain step + recording outside the timed region
    r0 = sim.graph_replays()
    a.record(stream)
    sim.step_n(dt, steady_steps)
    b.record(stream)
    sim.synchronize()
    steady_ms = a.elapsed_time(b) / steady_steps
    steady_replays = sim.graph_replays() - r0
    sim.set_stage_timing(True)

    # ---- end to end through host buffers: H2D state, step, D2H OutPositions, every step
    for _ in range(2):
        sim.upload_state_ptr(n, h_pos.data_ptr(), h_vel.data_ptr())
        sim.step(dt)
        sim.download_ptr("out_positions", h_out.data_ptr(), n * 16)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        sim.upload_state_ptr(n, h_pos.data_ptr(), h_vel.data_ptr())
        sim.step(dt)
        sim.download_ptr("out_positions", h_out.data_ptr(), n * 16)   # returns when the host buffer is filled
    e2e_blocking_s = time.perf_counter() - t0
    # the same frames through the pipelined calls (include/sph_b200.h): every frame still uploads its 24 B/particle of
    # input and downloads its 16 B/particle of OutPositions inside the timed region, but the upload of frame k+1 and the
    # download of frame k-1 travel on their own copy streams while frame k is computed
    h_out2 = [h_out, torch.empty((n, 4), dtype=torch.float32).pin_memory()]

    def pipelined(frames):
        sim.upload_state_begin(n, h_pos.data_ptr(), h_vel.data_ptr())
        for k in range(frames):
            sim.upload_state_commit()
            if k + 1 < frames:
                sim.upload_state_begin(n, h_pos.data_ptr(), h_vel.data_ptr())
            sim.step(dt)
            if k:
                sim.download_wait()                      # frame k-1's OutPositions are in host memory
            sim.download_begin("out_positions", h_out2[k & 1].data_ptr(), n * 16)
        sim.download_wait()
        sim.synchronize()

    e2e_path = ("per frame: sph_upload_state_begin/_commit(pinned pos3+vel3) -> sph_step -> sph_download_begin/_wait("
                "OUT_POSITIONS, pinned); copies on their own streams overlap the neighbouring frames' steps")
    try:
        pipelined(3)
        t0 = time.perf_counter()
        pipelined(args.steps)
        e2e_s = time.perf_counter() - t0
    except pkg.SphError as ex:                                # report the serial number rather than no number
        e2e_s = e2e_blocking_s
        e2e_path = "sph_upload_state -> sph_step -> sph_download(OUT_POSITIONS), serial (pipelined calls failed: %s)" % ex
    e2e = n * args.steps / e2e_s / 1e6
    clk = clocks.stop()          # sampled across the timed, steady-state and end-to-end loops (all under load)
    params_now = sim.get_params()

    peak, peak_src = measured_peaks()
    names = ["predict_key", "spatial", "density", "pressure", "viscosity", "integrate"]
    gather = {"density": stage[2], "pressure": stage[3], "viscosity": stage[4]}
    dom = max(gather, key=gather.get)
    dom_ms = gather[dom]
    achieved = A_BYTES[dom] * n / (dom_ms * 1e-3) / 1e9
    # the two purely streaming kernels, by the bytes they really move (sph_kernels.cu): predict_key reads pos+vel
    # (32 B) and writes the key (4 B); integrate reads pos+vel (32 B) and writes pos+vel (32 B)
    resident = n * 170 < 100e6
    streaming = {}
    for kname, bytes_pp, t_ms in (("k_predict_key", 36, stage[0]), ("k_integrate", 64, stage[5])):
        gbs = bytes_pp * n / (t_ms * 1e-3) / 1e9
        streaming[kname] = {"bytes_per_particle": bytes_pp, "achieved": gbs, "frac": gbs / peak, "kernel_ms": float(t_ms),
                            "note": "working set fits the 126 MB L2 at this size: served above the HBM peak" if resident else "HBM-resident"}
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic_latest.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get(name, {}).get(dom)
        except Exception:
            traffic = None
    # all three gather kernels side by side: A-figure roofline, live duration, and (from the committed ncu captures of
    # the same kernels) the measured DRAM traffic and the two SM-side roofs that actually bind them
    ncu = {}
    short = {"C2_dambreak_1M": "C2", "C3_dambreak_8M": "C3"}.get(name)
    if short:
        try:
            for rec in json.load(open(os.path.join(ROOT, "profiles", "r02_gather_final_%s.json" % short))):
                for key, frag in (("density", "k_density_pk"), ("pressure", "k_gather_list<0, 1>"), ("viscosity", "k_viscosity_w")):
                    if frag in rec.get("kernel", ""):
                        ncu[key] = rec
        except Exception:
            ncu = {}
    tr_all = {}
    try:
        tr_all = json.load(open(tp)).get(name, {})
    except Exception:
        pass

    def pct(rec, metric):
        try:
            return float(rec[metric].split()[0])
        except Exception:
            return None

    gather_kernels = {}
    for key in ("density", "pressure", "viscosity"):
        t_ms = float(gather[key])
        ach = A_BYTES[key] * n / (t_ms * 1e-3) / 1e9
        row = {"algorithmic_bytes_per_particle": A_BYTES[key], "kernel_ms": t_ms, "achieved": ach, "frac": ach / peak}
        if tr_all.get(key):
            row["traffic"] = tr_all[key]
            row["dram_gbs_live"] = tr_all[key] / (t_ms * 1e-3) / 1e9      # ncu's bytes per launch over the live duration
            row["dram_frac_live"] = row["dram_gbs_live"] / peak
        if key in ncu:
            row["ncu"] = {"l1_data_pipe_pct": pct(ncu[key], "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
                          "issue_active_pct": pct(ncu[key], "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                          "dram_pct": pct(ncu[key], "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                          "source": "profiles/r02_gather_final_%s.json" % short}
        gather_kernels[key] = row
    result = {
        "metric": METRIC, "value": value, "unit": "M updates/s", "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_text(name, n, sc["bound"], sc["params"].get("viscosity_strength", 0.5)),
                   "particles": n, "table": args.table,
                   "l2": "flushed between timed steps (256 MiB fill outside the event pairs)" if flush is not None
                         else "not flushed"},
        "steady_state": {"value": n / (steady_ms * 1e-3) / 1e6, "ms_per_step": steady_ms,
                         "steps": steady_steps, "graph_replays": int(steady_replays),
                         "note": ("sph_step_n: steps back to back (CUDA-graph replay), no L2 flush, stage timers off" if not dense else
                                  "two steps back to back from the dense state: the second one is already the blow-up")},
        "stage_ms": {k: float(v) for k, v in zip(names, stage)},
        "launch": {"timed_steps": "sph_step: %d of %d steps replayed the step's CUDA graph" % (timed_replays, args.steps),
                   "ms_per_step_plain_launches": plain_ms,
                   "note": "stage_ms (medians) and ms_per_step_plain_launches (median) come from a second loop over as many steps through "
                           "plain launches with the six stage timers on (sph_set_graph_replay(ctx, 0))"},
        "roofline": {"bound": "hbm", "kernel": "k_" + dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_particle": A_BYTES[dom], "kernel_ms": float(dom_ms),
                     "binding_roof": "instruction issue (76 % at 1 M, 79 % at 8 M) together with the SM's L1 data pipe "
                                     "(l1tex__data_pipe_lsu_wavefronts 71 % / 78 %), not HBM: every lane streams its own 12-byte candidates "
                                     "and the gather re-reads neighbours from L1/L2 by design (ncu: profiles/r02_gather_final_C2.txt, "
                                     "r02_gather_final_C3.txt; DESIGN.md section 5)",
                     "stages": {k: {"algorithmic_bytes_per_particle": a, "ms": float(t), "achieved": a * n / (float(t) * 1e-3) / 1e9,
                                    "frac": a * n / (float(t) * 1e-3) / 1e9 / peak}
                                for k, a, t in zip(names, (A_BYTES["predict_key"], A_BYTES["sort"] + A_BYTES["table"] + A_BYTES["reorder"],
                                                           A_BYTES["density"], A_BYTES["pressure"], A_BYTES["viscosity"],
                                                           A_BYTES["integrate"]), stage)},
                     "gather_kernels": gather_kernels,
                     "streaming_kernels": streaming,
                     "step": step_roofline(value, peak)},
        "e2e": {"value": e2e, "unit": "M updates/s", "h2d_bytes_per_step": n * 24, "d2h_bytes_per_step": n * 16,
                "ms_per_step": e2e_s / args.steps * 1e3,
                "path": e2e_path,
                "blocking": {"value": n * args.steps / e2e_blocking_s / 1e6, "ms_per_step": e2e_blocking_s / args.steps * 1e3,
                             "path": "sph_upload_state -> sph_step -> sph_download(OUT_POSITIONS), one stream, serial"}},
        "gpu_launches": int(launches),
        "clocks": clk,
    }
    sim.close()
    # the same frames through the C++ drop-in class (what a caller of the reference actually uses)
    try:
        if not dense:                                         # (the C5 column would blow up across consecutive frames)
            result["e2e"]["class_update"] = class_update_e2e(pkg, sc, params_now, max(args.steps, 20))
    except Exception as ex:
        result["e2e"]["class_update"] = {"error": "%s: %s" % (type(ex).__name__, ex)}
    if not args.config and not args.no_configs:
        # the other BASELINE configs, short lines under the driver's eyes; C4 on ONE GPU is the same-workload anchor of the
        # multi-GPU curve (the N > 1 lines run C4 slab-decomposed): efficiency(N) = strong_base.ms_per_step / (N * ms_per_step(N))
        del h_pos, h_vel, h_out, h_out2
        result["configs"] = {}
        for cname in (C1_NAME, "C3_dambreak_8M", C5_NAME, "C4_dambreak_64M"):
            try:
                result["configs"][cname] = short_line(pkg, scenes, torch, dev, cname, flush=flush)
            except Exception as ex:                         # a failing extra line must not cost the headline
                result["configs"][cname] = {"error": "%s: %s" % (type(ex).__name__, ex)}
        c4 = result["configs"].get("C4_dambreak_64M", {})
        if "ms_per_step" in c4:
            result["strong_base"] = {"workload": "C4_dambreak_64M on one GPU", "ms_per_step": c4["ms_per_step"], "value": c4["value"],
                                     "unit": "M updates/s", "particles": c4["particles"]}
    def one_thread(fn, **kw):            # the same reference on libstdc++'s serial backend, beside the all-threads number
        try:
            q = fn(parallel=False, **kw)
            return {"value": q["value"], "unit": "M updates/s", "cores": q["cores"], "kind": q["kind"], "sample": q["sample"]}
        except Exception as e:
            return {"error": str(e)}

    if not args.no_cpu_baseline and name == C1_NAME:
        r = cpu_reference_c1(steps=100, warmup=20)
        result["cpu_baseline"] = {"value": r["value"], "unit": "M updates/s", "cores": r["cores"], "kind": r["kind"],
                                  "sample": r["sample"], "host_cores": os.cpu_count()}
        if r["cores"] > 1:
            result["cpu_baseline"]["one_thread"] = one_thread(cpu_reference_c1, steps=50, warmup=10)
    elif not args.no_cpu_baseline:
        cfg = name if name in scenes.CONFIGS else "C2_dambreak_1M"
        r = cpu_reference_run(cfg, steps=2, warmup=1, budget_s=15.0)
        result["cpu_baseline"] = {"value": r["value"], "unit": "M updates/s", "cores": r["cores"], "kind": r["kind"],
                                  "sample": r["sample"], "host_cores": os.cpu_count()}
        if r["cores"] > 1:
            result["cpu_baseline"]["one_thread"] = one_thread(cpu_reference_run, scene_name=cfg, steps=2, warmup=1, budget_s=10.0)
        # a further, clearly labelled line: the C restatement with OpenMP on every host core
        try:
            ob = graft.load_oracle()
            cores = os.cpu_count() or 1
            o = ob.PortOracle(n, threads=cores, **sc["params"])
            o.set_state(sc["pos"], sc["vel"])
            o.step(dt, jacobi=True)
            t0 = time.perf_counter()
            for _ in range(2):
                o.step(dt, jacobi=True)
            tt = (time.perf_counter() - t0) / 2
            ob.PortOracle.lib().oracle_set_threads(1)
            result["cpu_baseline"]["port_openmp"] = {"value": n / tt / 1e6, "unit": "M updates/s", "cores": cores, "kind": "port",
                                                     "sample": "full %s, 2 steps after 1 warm-up, snapshot viscosity" % name}
            o.close()
        except Exception as e:                                  # never let the extra line break the bench
            result["cpu_baseline"]["port_openmp"] = {"error": str(e)}
    return result


if __name__ == "__main__":
    main()
