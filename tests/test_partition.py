"""Host-side logic of the multi-GPU path: the domain decomposition and its halo send/receive lists.
Runs on CPU; the cross-rank agreement is also checked between two real processes over gloo."""
import os
import socket

import numpy as np
import pytest


def check_plans(plans, mesh):
    W = len(plans)
    for p in plans:
        off_e = off_c = 0
        assert list(p["peer_rank"]) == sorted(p["peer_rank"]) and p["rank"] not in p["peer_rank"]
        for k, q in enumerate(p["peer_rank"]):
            ne, nc, re_, rc_ = p["peer_counts"][k]
            Q = plans[q]
            # the slot the sender writes to holds, in the receiver's numbering, exactly that cell/edge — and it is a ghost slot
            assert np.array_equal(Q["local_edge_ref"][p["send_edge_slot"][off_e:off_e + ne]], p["send_edge_ref"][off_e:off_e + ne])
            assert np.array_equal(Q["local_cell_ref"][p["send_cell_slot"][off_c:off_c + nc]], p["send_cell_ref"][off_c:off_c + nc])
            assert (p["send_edge_slot"][off_e:off_e + ne] >= Q["own_edges"]).all() and (p["send_cell_slot"][off_c:off_c + nc] >= Q["own_cells"]).all()
            kq = list(Q["peer_rank"]).index(p["rank"])
            assert Q["peer_counts"][kq][2] == ne and Q["peer_counts"][kq][3] == nc          # receiver expects what the sender sends
            off_e += ne
            off_c += nc
        # every ghost is filled by exactly one neighbour
        assert p["peer_counts"][:, 2].sum() == len(p["local_edge_ref"]) - p["own_edges"]
        assert p["peer_counts"][:, 3].sum() == len(p["local_cell_ref"]) - p["own_cells"]
    own_c = np.concatenate([p["local_cell_ref"][:p["own_cells"]] for p in plans])
    own_e = np.concatenate([p["local_edge_ref"][:p["own_edges"]] for p in plans])
    assert np.array_equal(np.sort(own_c), np.arange(mesh.n_cells)) and np.array_equal(np.sort(own_e), np.arange(mesh.n_edges))
    sizes = [p["own_cells"] for p in plans]
    assert max(sizes) - min(sizes) <= 1                                                       # balanced to one cell
    # the halo closes the stencils: every edge of an own cell, and every edge of both cells of an own edge, is held locally
    T = mesh.tables
    for p in plans:
        held_e = set(p["local_edge_ref"].tolist())
        held_c = set(p["local_cell_ref"].tolist())
        own_edges = p["local_edge_ref"][:p["own_edges"]]
        cells = T["face_nodes"][own_edges].ravel()
        assert set(cells.tolist()) <= held_c
        need = T["faces"][np.unique(np.concatenate([cells, p["local_cell_ref"][:p["own_cells"]]]))].ravel()
        assert set(need[need >= 0].tolist()) <= held_e


@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
def test_partition_plans_are_consistent(odis, world):
    pos, fr, cen = odis.generate_grid(5)
    mesh = odis.Mesh.from_arrays(pos, fr, cen, 1.0e6)
    plans = [odis.partition_plan(mesh, r, world) for r in range(world)]
    check_plans(plans, mesh)
    if world == 1:
        assert plans[0]["own_cells"] == mesh.n_cells and len(plans[0]["peer_rank"]) == 0


def test_halo_is_small(odis):
    """Contiguous ranges of the space-filling curve are compact patches: halo ~ perimeter, not area."""
    pos, fr, cen = odis.generate_grid(7)                     # 40,962 cells
    mesh = odis.Mesh.from_arrays(pos, fr, cen, 1.0e6)
    for r in range(8):
        p = odis.partition_plan(mesh, r, 8)
        ghosts = len(p["local_cell_ref"]) - p["own_cells"]
        assert ghosts < 12 * np.sqrt(p["own_cells"]), (r, ghosts)
        assert len(p["peer_rank"]) <= 8


def _gloo_worker(rank, world, port, level, q):
    import torch.distributed as dist
    import geodesicodis_b200 as odis
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        pos, fr, cen = odis.generate_grid(level)
        mesh = odis.Mesh.from_arrays(pos, fr, cen, 1.0e6)
        mine = odis.partition_plan(mesh, rank, world)
        plans = [None] * world
        dist.all_gather_object(plans, mine)                   # every rank derived its plan independently
        check_plans(plans, mesh)
        q.put((rank, "ok", mine["own_cells"]))
    except Exception as e:  # noqa: BLE001
        q.put((rank, repr(e), 0))
    finally:
        dist.destroy_process_group()


def test_two_processes_agree_over_gloo():
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, 4, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res
    assert sum(r[2] for r in res) == 642
