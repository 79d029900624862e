"""-m gpu: the pipelined transfer calls (sph_upload_state_begin/_commit, sph_download_begin/_wait) give
bit-identical results to the blocking sph_upload_state / sph_download, in the frame loop include/sph_b200.h shows."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _frames(scenes, n_side, frames):
    """a different input state per frame (so a frame that picked up the wrong upload is caught)"""
    out = []
    for k in range(frames):
        sc = scenes.small_dam_break(n_side, seed=11 + k)
        vel = sc["vel"].copy()
        vel[:, 0] += np.float32(0.25 * k)
        out.append((np.ascontiguousarray(sc["pos"]), np.ascontiguousarray(vel)))
    return scenes.small_dam_break(n_side, seed=11), out


def test_pipelined_frame_loop_is_bit_identical_to_blocking_calls(pkg):
    import torch
    from fluid_simulation_3d_b200 import scenes
    sc, frames = _frames(scenes, 20, 5)
    n = sc["n"]
    dt = scenes.DT

    ref_out = []
    a = pkg.FluidSimulation(n, **sc["params"])
    for pos, vel in frames:
        a.upload_state(pos, vel)
        a.step(dt)
        ref_out.append(a.download("out_positions").copy())
    a.close()

    b = pkg.FluidSimulation(n, **sc["params"])
    pinned = [(torch.from_numpy(p).pin_memory(), torch.from_numpy(v).pin_memory()) for p, v in frames]
    outs = [torch.empty((n, 4), dtype=torch.float32).pin_memory() for _ in range(2)]
    got = []
    b.upload_state_begin(n, pinned[0][0].data_ptr(), pinned[0][1].data_ptr())
    for k in range(len(frames)):
        b.upload_state_commit()
        if k + 1 < len(frames):
            b.upload_state_begin(n, pinned[k + 1][0].data_ptr(), pinned[k + 1][1].data_ptr())
        b.step(dt)
        if k:
            b.download_wait()
            got.append(outs[(k - 1) & 1].numpy().copy())
        b.download_begin("out_positions", outs[k & 1].data_ptr(), n * 16)
    b.download_wait()
    got.append(outs[(len(frames) - 1) & 1].numpy().copy())
    for k, (r, g) in enumerate(zip(ref_out, got)):
        assert np.array_equal(r.view(np.uint32), g.view(np.uint32)), "frame %d differs" % k
    # the blocking calls still work on the same context afterwards, and see the same state
    assert np.array_equal(b.download("out_positions").view(np.uint32), ref_out[-1].view(np.uint32))
    b.close()


def test_pipeline_call_order_errors_and_other_fields(pkg):
    from fluid_simulation_3d_b200 import scenes
    sc = scenes.small_dam_break(10)
    n = sc["n"]
    sim = pkg.FluidSimulation(n, **sc["params"])
    with pytest.raises(pkg.SphError):
        sim.upload_state_commit()                       # nothing pending
    with pytest.raises(pkg.SphError):
        sim.download_wait()                             # nothing pending
    pos = np.ascontiguousarray(sc["pos"])
    sim.upload_state_begin(n, pos.ctypes.data, None)    # pageable memory, no velocities: still correct
    with pytest.raises(pkg.SphError):
        sim.upload_state_begin(n, pos.ctypes.data, None)   # one upload at a time
    sim.upload_state_commit()
    sim.step(scenes.DT)
    dens = np.empty((n, 2), np.float32)
    with pytest.raises(pkg.SphError):
        sim.download_begin("densities", dens.ctypes.data, 8)   # buffer too small
    sim.download_begin("densities", dens.ctypes.data, dens.nbytes)
    with pytest.raises(pkg.SphError):
        sim.download_begin("densities", dens.ctypes.data, dens.nbytes)   # one download at a time
    sim.download_wait()
    assert np.array_equal(dens.view(np.uint32), sim.download("densities").view(np.uint32))
    # an empty upload through the pipeline is a no-op state, like the blocking call
    sim.upload_state_begin(0, 0, None)
    sim.upload_state_commit()
    assert sim.n == 0
    sim.step(scenes.DT)
    sim.close()
