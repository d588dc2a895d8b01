"""torchrun --nproc-per-node N scripts/multigpu_check.py [level]: one process per GPU, CUDA-IPC halo exchange;
checks the partitioned run bit-for-bit against a single-GPU run of the same problem (done on every rank)."""
import os, sys, time
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import geodesicodis_b200 as odis

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
level = int(sys.argv[1]) if len(sys.argv) > 1 else 7
pos, fr, cen = odis.generate_grid(level)
r = 252.1e3
mesh = odis.Mesh.from_arrays(pos, fr, cen, r)
dmin = float(mesh.tables["face_node_dist"].min())
prm = dict(g=0.113, h=38e3, alpha=1e-6, dt=0.1 * dmin / np.sqrt(0.113 * 38e3), radius=r, omega=5.307e-5, love_reduct=1.0, ecc=0.0047, obl=0.001,
           shell_thickness=0.0, semimajor_axis=0.0, potential=8, friction=0, surface=0, init_load=0, reorder=1)
rng = np.random.default_rng(3)
v0, e0 = rng.uniform(-1, 1, mesh.n_edges) * 1e-2, rng.uniform(-1, 1, mesh.n_cells)
part = odis.Solver(mesh, prm, device=local, rank=rank, world=world)
blobs = [None] * world
dist.all_gather_object(blobs, part.halo_blob())
part.halo_connect(blobs)
dist.barrier()
part.set_state(v0, e0)
nsteps = 100
part.step(nsteps)
def whole(fid):
    t = torch.from_numpy(part.field(fid)).cuda(); dist.all_reduce(t); return t.cpu().numpy()
v, eta = whole(0), whole(1)
series = torch.from_numpy(part.dissipation_series()).cuda(); dist.all_reduce(series); series = series.cpu().numpy()
ref = odis.Solver(mesh, prm, device=local)
ref.set_state(v0, e0); ref.step(nsteps)
ok = np.array_equal(v, ref.field(0)) and np.array_equal(eta, ref.field(1)) and np.allclose(series, ref.dissipation_series(), rtol=1e-12, atol=0)
dist.barrier(); torch.cuda.synchronize()
ms = part.step_timed(500)
t = torch.tensor([ms], device="cuda", dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX)
ms1 = ref.step_timed(500)
print(f"rank {rank}/{world}: level {level} cells {mesh.n_cells} partition {part.partition()} bit-identical={ok} "
      f"partitioned {500 / t.item() * 1e3:.0f} steps/s vs single GPU {500 / ms1 * 1e3:.0f} steps/s", flush=True)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
