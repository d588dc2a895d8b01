"""Per-CTA time stamps of the two staged kernels (variant library built with -DODIS_TRACE), single GPU or partitioned:
    [torchrun --nproc-per-node N] python scripts/halo_trace.py [level] [l_max] [outdir]
Writes <outdir>/trace_n<N>_l<level>_sg<l_max>_rank<r>.npy: [2 kernels][16 launches][296 CTAs][8 events][2: globaltimer ns, clock64].
Events (odis_kernels_pipe.cu): edge 0 entry, 1 first tile's rows arrived, 2 first tile stored, 3 first tile counted done (boundary: fence + flag),
4 / 5 group 0 / 1 through their tiles, 6 CTA done; cell 0 entry, 1 first rows arrived, 2,3 / 4,5 group 0 / 1 before, after the flag wait, 6 / 7 groups done."""
import ctypes, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
lib_path = os.path.join(ROOT, "geodesicodis_b200", "libodis_b200_trace.so")
os.environ["ODIS_B200_LIB"] = lib_path          # before the package is imported: geodesicodis_b200._lib reads it at import
from geodesicodis_b200.build import build_variant
if not os.path.exists(lib_path):
    build_variant("trace", ["ODIS_TRACE"])
import torch
import geodesicodis_b200 as odis

rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
level = int(sys.argv[1]) if len(sys.argv) > 1 else 9
l_max = int(sys.argv[2]) if len(sys.argv) > 2 else 0
outdir = sys.argv[3] if len(sys.argv) > 3 else "gpurun_out"
pos, fr, cen = odis.generate_grid(level)
r = 252.1e3 - 23e3
mesh = odis.Mesh.from_arrays(pos, fr, cen, r)
dmin = float(mesh.tables["face_node_dist"].min())
prm = dict(g=0.113, h=38e3, alpha=1e-7, dt=0.2 * dmin / np.sqrt(0.113 * 38e3), radius=r, omega=5.307e-5, love_reduct=0.9, ecc=0.0047,
           obl=0.0, shell_thickness=23e3, semimajor_axis=0.0, potential=5, friction=0, surface=2, init_load=0, reorder=1)
s = odis.Solver(mesh, prm, device=local, rank=rank, world=world)
if world > 1:
    blobs = [None] * world
    dist.all_gather_object(blobs, s.halo_blob())
    s.halo_connect(blobs)
    dist.barrier()
if l_max >= 2:
    f = 0.1 * np.ones(l_max + 1); f[:2] = 0.0
    s.enable_self_gravity(l_max, f)
SLOTS, CTAS = 16, 296
buf = torch.zeros(2 * SLOTS * CTAS * 8 * 2, dtype=torch.int64, device="cuda")
s.step(120)
s.synchronize()
torch.cuda.synchronize()
if dist is not None:
    dist.barrier()
    torch.cuda.synchronize()
from geodesicodis_b200 import _lib as odis_lib
assert os.path.samefile(odis_lib.LIB_PATH, lib_path), odis_lib.LIB_PATH
lib = ctypes.CDLL(lib_path)
lib.odis_debug_trace_enable.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32]
assert lib.odis_debug_trace_enable(buf.data_ptr(), SLOTS, CTAS) == 0
ms = s.step_timed(240)
s.synchronize()
assert lib.odis_debug_trace_enable(None, SLOTS, CTAS) == 0
torch.cuda.synchronize()
a = buf.cpu().numpy().reshape(2, SLOTS, CTAS, 8, 2)
os.makedirs(outdir, exist_ok=True)
np.save(os.path.join(outdir, f"trace_n{world}_l{level}_sg{l_max}_rank{rank}.npy"), a)
part = s.partition()
print(f"rank {rank}/{world} level {level} l_max {l_max}: {ms / 240 * 1e3:.2f} us/step, own cells {part['own_cells']} ghost {part['ghost_cells']} peers {part['n_peers']}", flush=True)
s.close()
if dist is not None:
    dist.barrier()
    dist.destroy_process_group()
