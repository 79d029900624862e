"""CPU model of the slab-decomposed step (the protocol sph_multi.cu implements), run with one process per
slab over torch.distributed/gloo.  The physics of every stage is the oracle's; this file only models the
decomposition: ownership by the z layer of the PREDICTED position, migration, one ghost layer per side,
and the three halos (predicted positions, densities, post-pressure velocities).

It validates the algorithm and the host-side partition helpers (slab_driver.choose_layers / owner_of)
without a GPU; the GPU implementation itself is checked against the single-GPU step by tests/mgpu_check.py.
"""
import numpy as np
import torch
import torch.distributed as dist


def _send(arrs, dst):
    """send a list of numpy arrays (shapes known to the receiver only through the header)."""
    hdr = torch.tensor([a.shape[0] for a in arrs], dtype=torch.int64)
    dist.send(hdr, dst)
    for a in arrs:
        if a.shape[0]:
            dist.send(torch.from_numpy(np.ascontiguousarray(a)), dst)


def _recv(templates, src):
    hdr = torch.zeros(len(templates), dtype=torch.int64)
    dist.recv(hdr, src)
    out = []
    for n, (tail, dt) in zip(hdr.tolist(), templates):
        a = np.zeros((n,) + tail, dt)
        if n:
            t = torch.from_numpy(a)
            dist.recv(t, src)
        out.append(a)
    return out


def _exchange(rank, world, to_lo, to_hi, templates):
    """neighbour exchange without deadlock: even ranks send first."""
    from_lo = from_hi = None
    for phase in (0, 1):
        if (rank % 2) == phase:
            if rank > 0:
                _send(to_lo, rank - 1)
            if rank < world - 1:
                _send(to_hi, rank + 1)
        else:
            if rank < world - 1:
                from_hi = _recv(templates, rank + 1)
            if rank > 0:
                from_lo = _recv(templates, rank - 1)
    return from_lo, from_hi


class SlabModel:
    def __init__(self, ob, slabmod, rank, world, layers, r, gmin_z, gz, params):
        self.ob, self.sm, self.rank, self.world = ob, slabmod, rank, world
        self.L, self.r, self.gmin_z, self.gz, self.params = layers, r, gmin_z, gz, params
        self.ids = np.zeros(0, np.uint32)
        self.pos = np.zeros((0, 3), np.float32)
        self.vel = np.zeros((0, 3), np.float32)

    def upload(self, ids, pos, vel):
        self.ids, self.pos, self.vel = ids.astype(np.uint32), pos.astype(np.float32), vel.astype(np.float32)

    def step(self, dt):
        ob, rank, world = self.ob, self.rank, self.world
        own_lo, own_hi = self.L[rank], self.L[rank + 1]
        f3, f2, u1 = ((3,), np.float32), ((2,), np.float32), ((), np.uint32)
        # (1) predict owned rows (oracle S1), classify by the layer of the predicted z
        a = ob.PortOracle(max(len(self.ids), 1), **self.params)
        if len(self.ids):
            a.set_state(self.pos, self.vel)
            a.stage_predict(dt)
            pred, vel1 = a.predicted(), a.velocities()
        else:
            pred, vel1 = np.zeros((0, 3), np.float32), np.zeros((0, 3), np.float32)
        lay = self.sm.layer_of(pred[:, 2], self.r, self.gmin_z, self.gz) if len(self.ids) else np.zeros(0, np.int64)
        mig_lo = (lay < own_lo) & (rank > 0)
        mig_hi = (lay >= own_hi) & (rank < world - 1)
        stay = ~(mig_lo | mig_hi)
        ghost_lo = stay & (lay == own_lo) & (rank > 0)
        ghost_hi = stay & (lay == own_hi - 1) & (rank < world - 1)
        keep_lo = mig_lo & (lay == own_lo - 1)
        keep_hi = mig_hi & (lay == own_hi)
        # (2) one exchange round: migrants (raw state) + ghosts (ids + predicted positions)
        pack = lambda m, g: [self.ids[m], self.pos[m], self.vel[m], self.ids[g], pred[g]]
        from_lo, from_hi = _exchange(rank, world, pack(mig_lo, ghost_lo), pack(mig_hi, ghost_hi), [u1, f3, f3, u1, f3])
        arr_ids, arr_pos, arr_vel, g_ids, g_pred = [], [], [], [], []
        for src in (from_lo, from_hi):
            if src is not None:
                arr_ids.append(src[0]); arr_pos.append(src[1]); arr_vel.append(src[2]); g_ids.append(src[3]); g_pred.append(src[4])
        g_ids += [self.ids[keep_lo], self.ids[keep_hi]]
        g_pred += [pred[keep_lo], pred[keep_hi]]
        arr_ids = np.concatenate(arr_ids + [np.zeros(0, np.uint32)]); arr_pos = np.concatenate(arr_pos + [np.zeros((0, 3), np.float32)])
        arr_vel = np.concatenate(arr_vel + [np.zeros((0, 3), np.float32)])
        g_ids = np.concatenate(g_ids); g_pred = np.concatenate(g_pred)
        self.stats = dict(migrated=int(mig_lo.sum() + mig_hi.sum()), ghosts=int(len(g_ids)))
        # arrivals are predicted here, from their raw state, exactly like resident rows
        if len(arr_ids):
            b = ob.PortOracle(len(arr_ids), **self.params)
            b.set_state(arr_pos, arr_vel)
            b.stage_predict(dt)
            arr_pred, arr_vel1 = b.predicted(), b.velocities()
        else:
            arr_pred, arr_vel1 = np.zeros((0, 3), np.float32), np.zeros((0, 3), np.float32)
        o_ids = np.concatenate([self.ids[stay], arr_ids])
        o_pos = np.concatenate([self.pos[stay], arr_pos])
        o_vel = np.concatenate([vel1[stay], arr_vel1])          # velocities AFTER S1
        o_pred = np.concatenate([pred[stay], arr_pred])
        n_own, n_all = len(o_ids), len(o_ids) + len(g_ids)
        # (3) local problem = owned + ghost rows; the oracle's own table (hash % n_local) is private to the rank
        c = ob.PortOracle(max(n_all, 1), **self.params)
        all_pos = np.concatenate([o_pos, g_pred]); all_vel = np.concatenate([o_vel, np.zeros((len(g_ids), 3), np.float32)])
        all_pred = np.concatenate([o_pred, g_pred])
        c.set_state(all_pos, all_vel)
        c.set_predicted(all_pred)
        c.stage_spatial()
        c.stage_density()
        dens = c.densities()
        self.ncount = c.neighbour_counts()[:n_own]
        o_lay = self.sm.layer_of(o_pred[:, 2], self.r, self.gmin_z, self.gz)
        self.owned_layers = o_lay                       # what sph_comm_rebalance reads off the step's table
        b_lo = (o_lay == own_lo) & (rank > 0)
        b_hi = (o_lay == own_hi - 1) & (rank < world - 1)
        # (4) halo of densities, matched by id
        from_lo, from_hi = _exchange(rank, world, [o_ids[b_lo], dens[:n_own][b_lo]], [o_ids[b_hi], dens[:n_own][b_hi]], [u1, f2])
        pos_of = {int(i): k for k, i in enumerate(g_ids)}
        for src in (from_lo, from_hi):
            if src is not None:
                for i, d in zip(src[0], src[1]):
                    dens[n_own + pos_of[int(i)]] = d
        assert len(pos_of) == len(g_ids), "duplicate ghost ids"
        c.set_densities(dens)
        c.stage_pressure(dt)
        velp = c.velocities()
        # (5) halo of post-pressure velocities
        from_lo, from_hi = _exchange(rank, world, [o_ids[b_lo], velp[:n_own][b_lo]], [o_ids[b_hi], velp[:n_own][b_hi]], [u1, f3])
        for src in (from_lo, from_hi):
            if src is not None:
                for i, v in zip(src[0], src[1]):
                    velp[n_own + pos_of[int(i)]] = v
        c.set_state(None, velp)
        c.stage_viscosity(dt, jacobi=True)
        c.stage_integrate(dt)
        self.ids, self.pos, self.vel = o_ids, c.positions()[:n_own], c.velocities()[:n_own]
        self.dens = c.densities()[:n_own]


    def rebalance(self, pkg, max_shift):
        """model of sph_comm_rebalance: global per-layer histogram of the last step's owned rows (sum over ranks),
        cut by the library's own pure host function, new layers in force from the next step on"""
        hist = torch.from_numpy(np.bincount(self.owned_layers, minlength=self.gz).astype(np.int64))
        dist.all_reduce(hist)
        new = self.sm.balance_layers(pkg, hist.numpy().astype(np.uint32), self.world, self.L, max_shift)
        changed = list(new) != list(self.L)
        self.L = list(new)
        return changed


def run_rank(rank, world, port, scene, steps, dt, q, rebalance_every=0, skew=0):
    import os
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import __graft_entry__ as g
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    ob = g.load_oracle()
    pkg = g.load_package()
    from fluid_simulation_3d_b200 import slab_driver as sm
    r = 0.35
    half = np.float32(scene["params"]["bound"][2]) * np.float32(0.5)
    qz = int(np.floor(float(half) / r))
    gmin_z, gz = -qz - 3, (qz + 2) - (-qz - 3) + 1           # same geometry rule as the library (update_grid_geometry)
    pred0 = scene["pos"] + scene["vel"] * np.float32(1.0 / 120.0)
    layers = sm.choose_layers(pred0[:, 2], world, r, gmin_z, gz)
    if skew:                                                 # deliberately unbalanced start: every inner plane `skew` layers up
        layers = [layers[0]] + [min(l + skew, gz - 3 * (world - k)) for k, l in enumerate(layers[1:-1], 1)] + [layers[-1]]
    layers0 = list(layers)
    own = sm.owner_of(scene["pos"][:, 2], layers, r, gmin_z, gz) == rank
    m = SlabModel(ob, sm, rank, world, layers, r, gmin_z, gz, scene["params"])
    m.upload(np.nonzero(own)[0], scene["pos"][own], scene["vel"][own])
    migrated = 0
    moves = 0
    for k in range(steps):
        m.step(dt)
        migrated += m.stats["migrated"]
        if rebalance_every and (k + 1) % rebalance_every == 0 and k + 1 < steps:
            moves += int(m.rebalance(pkg, 2))
    out = [None] * world
    dist.all_gather_object(out, (m.ids, m.pos, m.vel, m.dens, m.ncount, migrated, (layers0, list(m.L), moves)))
    if rank == 0:
        q.put(out)
    dist.barrier()
    dist.destroy_process_group()
