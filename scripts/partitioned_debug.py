"""torchrun --nproc-per-node N scripts/partitioned_debug.py [level] [l_max] [chunk] [nchunks] [kernel_select]
One process per GPU, the bench's workload; steps in chunks with a synchronize (which reports an in-kernel wait that gave up) after
each, prints per-chunk device time. For finding where a partitioned run stalls, and for per-step times at N GPUs."""
import os, sys, time
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import geodesicodis_b200 as odis

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
level = int(sys.argv[1]) if len(sys.argv) > 1 else 9
l_max = int(sys.argv[2]) if len(sys.argv) > 2 else 2
chunk = int(sys.argv[3]) if len(sys.argv) > 3 else 100
nchunks = int(sys.argv[4]) if len(sys.argv) > 4 else 15
sel = int(sys.argv[5]) if len(sys.argv) > 5 else 0
pos, fr, cen = odis.generate_grid(level)
r = 252.1e3 - 23e3
mesh = odis.Mesh.from_arrays(pos, fr, cen, r)
dmin = float(mesh.tables["face_node_dist"].min())
prm = dict(g=0.113, h=38e3, alpha=1e-7, dt=0.2 * dmin / np.sqrt(0.113 * 38e3), radius=r, omega=5.307e-5, love_reduct=0.9, ecc=0.0047,
           obl=0.0, shell_thickness=23e3, semimajor_axis=0.0, potential=5, friction=0, surface=2, init_load=0, reorder=1, kernel_select=sel)
s = odis.Solver(mesh, prm, device=local, rank=rank, world=world)
blobs = [None] * world
dist.all_gather_object(blobs, s.halo_blob())
s.halo_connect(blobs)
dist.barrier()
if l_max >= 2:
    f = 0.1 * np.ones(l_max + 1); f[:2] = 0.0
    s.enable_self_gravity(l_max, f)
env = " ".join(f"{k}={v}" for k, v in sorted(os.environ.items()) if k.startswith("ODIS_B200_") and k != "ODIS_B200_LIB")
ok, times = True, []
for c in range(nchunks):
    dist.barrier(); torch.cuda.synchronize()
    try:
        ms = s.step_timed(chunk)
        times.append(ms / chunk * 1e3)
    except Exception as e:                       # noqa: BLE001
        print(f"rank {rank}: chunk {c} (steps {c * chunk}..{(c + 1) * chunk}) FAILED: {e}", flush=True)
        ok = False
        break
t = torch.tensor([min(times) if times else -1.0, 1.0 if ok else 0.0], device="cuda", dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MIN if not times else dist.ReduceOp.MAX) if False else None
best = torch.tensor([min(times) if times else 1e9], device="cuda", dtype=torch.float64)
dist.all_reduce(best, op=dist.ReduceOp.MAX)
okt = torch.tensor([1.0 if ok else 0.0], device="cuda", dtype=torch.float64)
dist.all_reduce(okt, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"N={world} level {level} l_max {l_max} select {sel} [{env}]: ok={bool(okt.item() > 0.5)} chunks done {len(times)}/{nchunks}, "
          f"best chunk {best.item():.2f} us/step ({1e6 / best.item():.0f} steps/s), launches/step {s.launches / max(1, s.iter):.1f}", flush=True)
try:
    s.close()
except Exception:                                # noqa: BLE001
    pass
dist.barrier()
dist.destroy_process_group()
