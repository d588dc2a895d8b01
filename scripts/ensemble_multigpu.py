"""BASELINE config 5: a 256-member sweep (16 ocean thicknesses x 16 drag coefficients) on the 40,962-cell grid,
32 members per GPU, self-gravity to degree 8 as FP64 tensor-core GEMMs. Members are independent: replicas only, no data-path
communication (torch.distributed carries the barrier and the max-over-ranks time).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 scripts/ensemble_multigpu.py [level] [members_per_gpu] [l_max]

Prints one JSON line on rank 0 (a reported number; bench.py's line is the benchmark)."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import geodesicodis_b200 as odis

level = int(sys.argv[1]) if len(sys.argv) > 1 else 7
per_gpu = int(sys.argv[2]) if len(sys.argv) > 2 else 32
l_max = int(sys.argv[3]) if len(sys.argv) > 3 else 8
rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
have_cuda = torch.cuda.is_available()          # False only under the test emulation (ODIS_B200_LIB)
if have_cuda:
    torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
pos, fr, cen = odis.generate_grid(level)
r = 252.1e3 - 23e3
mesh = odis.Mesh.from_arrays(pos, fr, cen, r)
dmin = float(mesh.tables["face_node_dist"].min())
M = per_gpu * world
side = int(round(np.sqrt(M)))
hs, alphas = np.logspace(3, 5, side), np.logspace(-11, -6, max(M // side, 1))           # 1-100 km, 1e-11-1e-6
base = dict(g=0.113, h=38e3, alpha=1e-7, dt=0.2 * dmin / np.sqrt(0.113 * 100e3), radius=r, omega=5.307e-5, love_reduct=0.9, ecc=0.0047,
            obl=0.0, shell_thickness=23e3, semimajor_axis=0.0, potential=5, friction=0, surface=2, init_load=0, reorder=1)
mine = range(rank * per_gpu, (rank + 1) * per_gpu)
plist = [dict(base, h=float(hs[m % side]), alpha=float(alphas[(m // side) % len(alphas)])) for m in mine]
ens = odis.Ensemble(mesh, plist, device=local)
if l_max >= 2:
    ens.enable_self_gravity(l_max, 0.1 * np.ones(l_max + 1))
ens.step(40)
steps = 400
if world > 1:
    dist.barrier()
if have_cuda:
    torch.cuda.synchronize()
ms = torch.tensor([ens.step_timed(steps)], device="cuda" if have_cuda else "cpu")
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    dist.barrier()
if rank == 0:
    t = float(ms.item()) * 1e-3
    info = ens.info()
    print(json.dumps({"config": f"{M}-member sweep, {mesh.n_cells} cells, {per_gpu} members per GPU x {world} GPUs, self-gravity degree {l_max}",
                      "batched_steps_per_s_per_gpu": round(steps / t, 1), "member_steps_per_s_total": round(M * steps / t, 1),
                      "us_per_member_step": round(t / steps / per_gpu * 1e6, 3),
                      "algorithmic_GBps_per_gpu": round(info["algorithmic_bytes_per_step"] * steps / t / 1e9, 1), "scaling": "replicas only"}), flush=True)
ens.close()
if world > 1:
    dist.destroy_process_group()
