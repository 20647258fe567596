#!/usr/bin/env python3
"""Per-call time of the three halo exchanges at 256^3 per GPU (torchrun, one rank per GPU)."""
import os, sys, time
import torch, torch.distributed as dist
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import ludwig_b200 as lb

rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
sim = lb.Lb200((n, n, n), nhalo=2, have_phi=True, device=local, cart_size=world, cart_rank=rank)
if world > 1:
    ids = [sim.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    sim.nccl_init(ids[0], world, rank)
for name, fn in (("phi_halo", sim.phi_halo), ("u_halo", sim.hydro_u_halo), ("lb_halo", sim.lb_halo)):
    for _ in range(5):
        fn()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        fn()
    dt = (time.perf_counter() - t0) / 20
    if rank == 0:
        print(f"world={world} {name}: {dt*1e3:.3f} ms per call (synchronous)", flush=True)
sim.close()
if world > 1:
    dist.destroy_process_group()
