"""CPU: the C-ABI library loads and exports every symbol include/sph_b200.h declares; struct layout;
the host C++ class library links against it.  No compute calls (no GPU here)."""
import ctypes as C
import os
import re
import subprocess

import pytest

import __graft_entry__ as g

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "sph_b200.h")


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sph_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_what_the_binding_lists(pkg):
    assert header_functions() == sorted(pkg.ABI_SYMBOLS)


def test_library_exports_every_declared_symbol(pkg):
    lib = pkg.load_library()
    for name in header_functions():
        assert hasattr(lib, name), name
    assert lib.sph_abi_version() == 1


def test_exports_are_plain_c(pkg):
    out = subprocess.run(["nm", "-D", "--defined-only", pkg.LIB_PATH], stdout=subprocess.PIPE, text=True).stdout
    exported = {ln.split()[-1] for ln in out.splitlines() if " T " in ln}
    for name in header_functions():
        assert name in exported, name + " is not an unmangled export"


def test_params_struct_layout_and_defaults(pkg):
    assert C.sizeof(pkg.SphParams) == 44
    p = pkg.default_params()
    assert abs(p.interaction_radius - 0.35) < 1e-7 and abs(p.sqr_radius - 0.1225) < 1e-7
    assert abs(p.target_density - 99.7) < 1e-5 and p.pressure_multiplier == 300 and p.near_pressure_multiplier == 20
    assert p.viscosity_strength == 0.5 and p.gravity_scale == 10 and p.gravity == 0       # gravity off by default
    assert list(p.bound) == [20, 20, 20]


def test_struct_layouts_match_a_c_compiler(pkg, tmp_path):
    """the ctypes mirrors against what gcc lays out from the header itself (a plain-C translation unit: the header
    must stay C, not C++)"""
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "sph_b200.h"\n'
                   'int main(void){printf("%zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(SphParams), offsetof(SphParams, bound),'
                   ' sizeof(SphBlockSpawn), offsetof(SphBlockSpawn, gap), offsetof(SphBlockSpawn, origin),'
                   ' offsetof(SphBlockSpawn, jitter_amp), offsetof(SphBlockSpawn, velocity_scale), offsetof(SphBlockSpawn, seed));return 0;}\n')
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], stdout=subprocess.PIPE, text=True, check=True).stdout.split()]
    B = pkg.SphBlockSpawn
    assert got == [C.sizeof(pkg.SphParams), pkg.SphParams.bound.offset, C.sizeof(B), B.gap.offset, B.origin.offset,
                   B.jitter_amp.offset, B.velocity_scale.offset, B.seed.offset]


def test_create_without_gpu_fails_loudly(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(pkg.SphError) as e:
        pkg.FluidSimulation(16)
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)


def test_product_library_does_not_link_the_oracle(pkg):
    out = subprocess.run(["ldd", pkg.LIB_PATH], stdout=subprocess.PIPE, text=True).stdout
    assert "sph_oracle" not in out and "sph_ref" not in out
    for root, _, files in os.walk(os.path.join(ROOT, "fluid-simulation-3d_b200")):
        for f in files:
            if f.endswith((".cu", ".cc", ".h", ".cuh", ".py")):
                txt = open(os.path.join(root, f), errors="ignore").read()
                assert "oracle/" not in txt.replace("oracle/_ref", "").replace("oracle/", "oracle/") or f == "__init__.py" or \
                    "load_oracle" not in txt, f


def test_host_class_library_links(pkg):
    host = os.path.join(ROOT, "fluid-simulation-3d_b200", "host", "libfluidsim_b200_host.so")
    assert os.path.exists(host), "run __graft_entry__.build()"
    out = subprocess.run(["nm", "-DC", "--defined-only", host], stdout=subprocess.PIPE, text=True).stdout
    for sym in ("Physics::Fluid::FluidSimulation::Update(float)", "Physics::Fluid::FluidSimulation::getInstance()",
                "Physics::Fluid::FluidSimulation::InitializeData(int,", "Physics::Fluid::FluidSimulation::getSpeedNormalzied(",
                "Physics::Fluid::FluidSimulation::setBound(", "Physics::Fluid::FluidSimulation::setMaxTimestep(float)",
                "FluidSimB200::initialize(int)", "FluidSimB200::update(float)", "FluidSimB200::reset()", "FluidSimB200::cleanup()",
                "FluidSimB200::render(Shader&, RenderUtils::Camera&)"):
        assert sym in out, sym


def test_host_driver_fails_loudly_without_gpu():
    """the C++ host layer has no CPU fallback either: without a device InitializeData throws"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    demo = os.path.join(ROOT, "fluid-simulation-3d_b200", "host", "host_demo")
    for what in ("class", "adapter", "multi", "slabgroup"):
        # ("slabgroup" constructs the multi-GPU group directly: its worker threads must be stopped and joined on the
        # failure path, or the process would terminate instead of reporting)
        r = subprocess.run([demo, "100", "1", "0", what], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=60)
        assert r.returncode == 2 and "no CUDA device" in r.stdout, r.stdout


def test_every_entry_point_survives_null_arguments(pkg):
    """the reference's class cannot be misused through a null pointer; a C ABI can.  Every handle-taking entry point
    returns an error status (or a benign value from the pure queries) for a NULL context and NULL buffers -- no compute,
    no device needed"""
    L = pkg.load_library()
    P, B = pkg.SphParams(), pkg.SphBlockSpawn()
    status = {          # must report an error
        "sph_create": lambda: L.sph_create(None, 0, 10),
        "sph_set_params": lambda: L.sph_set_params(None, C.byref(P)), "sph_get_params": lambda: L.sph_get_params(None, C.byref(P)),
        "sph_set_graph_replay": lambda: L.sph_set_graph_replay(None, 1),
        "sph_set_extras": lambda: L.sph_set_extras(None, None), "sph_get_extras": lambda: L.sph_get_extras(None, None),
        "sph_set_table_mode": lambda: L.sph_set_table_mode(None, 0), "sph_set_stage_timing": lambda: L.sph_set_stage_timing(None, 1),
        "sph_set_neighbour_count_tap": lambda: L.sph_set_neighbour_count_tap(None, 1),
        "sph_set_neighbour_list_capacity": lambda: L.sph_set_neighbour_list_capacity(None, 8),
        "sph_spawn_grid": lambda: L.sph_spawn_grid(None, 10), "sph_spawn_block": lambda: L.sph_spawn_block(None, C.byref(B)),
        "sph_upload_state": lambda: L.sph_upload_state(None, 0, None, None), "sph_step": lambda: L.sph_step(None, 0.01),
        "sph_step_n": lambda: L.sph_step_n(None, 0.01, 3), "sph_synchronize": lambda: L.sph_synchronize(None),
        "sph_refresh_densities": lambda: L.sph_refresh_densities(None), "sph_download": lambda: L.sph_download(None, 0, None, 0),
        "sph_download_table": lambda: L.sph_download_table(None, 0, None, 0, None),
        "sph_get_particle": lambda: L.sph_get_particle(None, 0, None), "sph_get_timings": lambda: L.sph_get_timings(None, None),
        "sph_get_grid": lambda: L.sph_get_grid(None, None, None), "sph_save_state": lambda: L.sph_save_state(None, b"/tmp/none"),
        "sph_load_state": lambda: L.sph_load_state(None, b"/tmp/none"), "sph_host_register": lambda: L.sph_host_register(None, 0),
        "sph_host_unregister": lambda: L.sph_host_unregister(None),
        "sph_upload_state_begin": lambda: L.sph_upload_state_begin(None, 0, None, None),
        "sph_upload_state_commit": lambda: L.sph_upload_state_commit(None), "sph_download_begin": lambda: L.sph_download_begin(None, 0, None, 0),
        "sph_download_wait": lambda: L.sph_download_wait(None), "sph_comm_get_id": lambda: L.sph_comm_get_id(None, 0),
        "sph_comm_init": lambda: L.sph_comm_init(None, 0, 1, None, 0), "sph_comm_set_planes": lambda: L.sph_comm_set_planes(None, None),
        "sph_upload_owned": lambda: L.sph_upload_owned(None, 0, None, None, None),
        "sph_download_owned": lambda: L.sph_download_owned(None, 0, None, None, 0, None),
        "sph_upload_owned_begin": lambda: L.sph_upload_owned_begin(None, 0, None, None, None),
        "sph_download_owned_begin": lambda: L.sph_download_owned_begin(None, 0, None, None, 0, None),
        "sph_download_owned_scatter": lambda: L.sph_download_owned_scatter(None, 0, None, 0, None),
        "sph_comm_stats": lambda: L.sph_comm_stats(None, None), "sph_comm_get_layers": lambda: L.sph_comm_get_layers(None, None),
        "sph_comm_rebalance": lambda: L.sph_comm_rebalance(None, 1, None, None, 0, None),
        "sph_slab_balance_layers": lambda: L.sph_slab_balance_layers(None, 0, 0, None, 0, 0, None),
        "sph_slab_link_ok": lambda: L.sph_slab_link_ok(None, None),
    }
    benign = {          # pure queries: a neutral answer
        "sph_destroy": (lambda: L.sph_destroy(None), 0), "sph_get_table_mode": (lambda: L.sph_get_table_mode(None), -1),
        "sph_num_particles": (lambda: L.sph_num_particles(None), 0), "sph_graph_replays": (lambda: L.sph_graph_replays(None), 0),
        "sph_launch_count": (lambda: L.sph_launch_count(None), 0), "sph_stream": (lambda: L.sph_stream(None), None),
        "sph_grid_x_subdivision": (lambda: L.sph_grid_x_subdivision(None), 0),
        "sph_noncanonical_cells": (lambda: L.sph_noncanonical_cells(None), 0),
        "sph_density_stack_rows": (lambda: L.sph_density_stack_rows(None), 0),
    }
    other = {"sph_default_params", "sph_last_error", "sph_abi_version", "sph_comm_id_bytes"}     # take no handle / checked elsewhere
    assert set(status) | set(benign) | other == set(pkg.ABI_SYMBOLS)
    for name, call in status.items():
        assert call() != 0, name + " accepted NULL arguments"
    for name, (call, want) in benign.items():
        assert call() == want, name
    assert L.sph_comm_id_bytes() >= 128 and isinstance(L.sph_last_error(None), bytes)


def test_host_snapshot_file_is_the_documented_format(tmp_path):
    """FluidSimulation::saveState / loadState go through sphb200::writeSnapshotFile / readSnapshotFile (pure host code):
    the file is the C ABI's snapshot format (include/sph_b200.h: "SPHB2002", u32 n, u32 sizeof(SphParams), SphParams,
    n x pos3, n x vel3); a count that does not fit the file is rejected before anything is allocated"""
    import numpy as np
    demo = os.path.join(ROOT, "fluid-simulation-3d_b200", "host", "host_demo")
    path = str(tmp_path / "state.bin")
    n = 777
    r = subprocess.run([demo, str(n), "0", "0", "snapshotio", path], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=60)
    assert r.returncode == 0 and "roundtrip=1 rejects_missing=1" in r.stdout, r.stdout
    raw = open(path, "rb").read()
    assert raw[:8] == b"SPHB2002" and len(raw) == 8 + 4 + 4 + 44 + 24 * n
    assert int(np.frombuffer(raw, np.uint32, 1, 8)[0]) == n and int(np.frombuffer(raw, np.uint32, 1, 12)[0]) == 44
    par = np.frombuffer(raw, np.float32, 11, 16)
    assert abs(par[0] - 0.35) < 1e-7 and par[5] == np.float32(0.75) and np.frombuffer(raw, np.int32, 1, 16 + 28)[0] == 1
    assert list(par[8:11]) == [20.0, 20.0, 7.5]
    pos = np.frombuffer(raw, np.float32, 3 * n, 60)
    vel = np.frombuffer(raw, np.float32, 3 * n, 60 + 12 * n)
    i = np.arange(3 * n, dtype=np.float32)
    assert np.array_equal(pos, np.float32(0.25) * i - np.float32(3.0)) and np.array_equal(vel, np.float32(-0.5) * i)
    # a header that claims more particles than the file holds (or more than an int) must be refused, not allocated
    bad = str(tmp_path / "bad.bin")
    for claim in (n + 1, 0x7FFFFFFF, 0xFFFFFFFF):
        open(bad, "wb").write(raw[:8] + np.uint32(claim).tobytes() + raw[12:])
        r = subprocess.run([demo, "0", "0", "0", "snapshotread", bad], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=60)
        assert r.returncode == 0 and "read=0" in r.stdout, (claim, r.stdout)


def test_slab_link_verdict_is_symmetric_and_catches_every_overflow(pkg):
    """sph_slab_link_ok (pure host): the two ends of a slab link evaluate it with the messages swapped and must agree, or
    one of them posts a receive the other never matches.  Checked on random messages, plus the four ways a link fails:
    a failed rank, a list longer than the sender's exchange buffers, arrivals beyond the receiver's free rows, ghosts
    beyond its free ghost rows."""
    import numpy as np
    L = pkg.load_library()
    ok = lambda a, b: L.sph_slab_link_ok(C.c_void_p(a.ctypes.data), C.c_void_p(b.ctypes.data))
    rng = np.random.default_rng(5)
    seen = set()
    for _ in range(4000):
        a = rng.integers(0, 12, 8).astype(np.uint32); b = rng.integers(0, 12, 8).astype(np.uint32)
        a[3] = rng.integers(0, 8) == 0; b[3] = rng.integers(0, 8) == 0
        va, vb = ok(a, b), ok(b, a)
        assert va == vb and va in (0, 1)
        seen.add(va)
    assert seen == {0, 1}
    good = np.array([5, 7, 2, 0, 100, 50, 40, 0], np.uint32)
    assert ok(good, good) == 1
    for word, value in ((3, 1), (0, 101), (1, 101), (2, 101)):                 # status, lists beyond the exchange buffer
        bad = good.copy(); bad[word] = value
        assert ok(bad, good) == 0 and ok(good, bad) == 0
    tight = good.copy(); tight[5] = 11                                         # 5 + 7 arrivals > 11 free rows
    assert ok(good, tight) == 0 and ok(tight, good) == 0
    tight = good.copy(); tight[6] = 6                                          # 7 ghosts > 6 free ghost rows
    assert ok(good, tight) == 0 and ok(tight, good) == 0


def test_every_chained_launch_waits_in_kernel():
    """launch_chained() (programmatic stream serialisation) lets a kernel start while its predecessor drains; the kernel
    itself must order its memory accesses with chain_prologue() (griddepcontrol.wait) before it touches anything.  A kernel
    launched that way without the wait would race silently, so the sources are checked: every kernel name handed to
    launch_chained has chain_prologue() as the first statement of its body."""
    import glob
    import re
    src = {}
    for path in glob.glob(os.path.join(ROOT, "fluid-simulation-3d_b200", "csrc", "*.cu")):
        src[path] = open(path).read()
    text = "\n".join(src.values())
    names = set(re.findall(r"launch_chained\(\s*(k_\w+)", text))
    assert len(names) >= 10, names
    for name in sorted(names):
        m = re.search(r"__global__[^;{]*?\b" + name + r"\s*\([^{;]*?\)\s*(?://[^\n]*)?\s*\{(.*?)\n\}", text, re.S)
        assert m, "no definition found for " + name
        body = re.sub(r"//[^\n]*", "", m.group(1))               # drop comments
        first = body.strip().split(";")[0].strip()
        assert first == "chain_prologue()", "%s: first statement is %r" % (name, first)
