"""Concurrent PCIe ceiling of a multi-GPU box: every rank copies page-locked host memory to its GPU and back at the
same time (two streams), all ranks together -- what the multi-GPU end-to-end loop of bench.py can reach at best.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port 29520 tools/pcie_ceiling.py [MiB]

Prints one JSON line on rank 0: per-rank and aggregate GB/s for H2D alone, D2H alone and both directions at once."""
import json, os, sys, time
import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); dev = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(dev)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", dev))
mib = int(sys.argv[1]) if len(sys.argv) > 1 else 256
n = mib << 20
h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); h_in.fill_(1)
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device="cuda"); d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
reps = 10

def run(h2d, d2h):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())

for _ in range(2):
    run(True, True)
res = {}
for name, a, b in (("h2d", True, False), ("d2h", False, True), ("both", True, True)):
    t = run(a, b)
    per_dir = n * reps / t / 1e9
    res[name] = {"per_gpu_per_direction_gbs": per_dir, "aggregate_gbs": per_dir * world * ((1 if a else 0) + (1 if b else 0))}
if rank == 0:
    print(json.dumps({"gpus": world, "buffer_mib": mib, **res}), flush=True)
if world > 1:
    dist.destroy_process_group()
