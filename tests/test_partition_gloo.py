"""CPU, world_size 2 and 4 over gloo: the host logic of the multi-GPU path (cell partition + ghost layer,
dof ownership, halo lists).  The numerical part is emulated with the oracle: a rank's local matrix
times a consistent local vector must reproduce the global SpMV on owned rows, and one halo exchange
driven by the send/recv lists must make the ghost entries right."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle.mesh import unit_cube_mesh, SimplexMesh
from oracle.mpet import MPETOracle

PARAMS = dict(J=2, E=2.2, nu=0.4545, alpha=(0.5, 0.5), c=(1.0, 1.0), K=(1.0, 1.0), S=((0, 1.0), (1.0, 0)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _global_node_index(o_global, keys, nvg):
    """global scalar-node index of rank-independent node keys (vertex id, or nvg + lo*nvg + hi)."""
    ev = o_global.space.edge_vertices
    gkeys = np.concatenate([np.arange(nvg), nvg + ev[:, 0] * nvg + ev[:, 1]])
    order = np.argsort(gkeys)
    pos = np.searchsorted(gkeys[order], keys)
    assert np.all(gkeys[order][pos] == keys)
    return order[pos]


def _worker(rank, world, port, n, results, mode="slab"):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from waterscapes_b200.parallel import (extract_local, box_slab, box_brick, brick_grid, rcb_partition, Partition,
                                               node_global_keys)
        from waterscapes_b200.mpet.dolfin_shim import FunctionSpace, UnitCubeMesh
        g = unit_cube_mesh(n, jitter=0.2 if mode == "rcb" else 0.0)
        J = PARAMS["J"]
        if mode == "slab":
            # cell partition into z-slabs by cell index (cells are ordered layer by layer, 6 n^2 per layer)
            layer = np.arange(g.num_cells) // (6 * n * n)
            bounds = [(n * r) // world for r in range(world + 1)]
            cell_owner = np.searchsorted(np.asarray(bounds[1:]), layer, side="right").astype(np.int64)
            local = extract_local(g.coords, g.cells, cell_owner, rank)
            slab = box_slab((0, 0, 0), (1, 1, 1), n, n, n, rank, world)
            assert np.array_equal(slab.global_vertex[slab.cells], local.global_vertex[local.cells])
            assert np.allclose(slab.coordinates, local.coordinates)
            assert np.array_equal(slab.cell_owner, local.cell_owner)
            # artificial (cut) facets carry no boundary: both constructions agree
            fs, fl = slab.exterior_facets(), local.exterior_facets()
            assert np.array_equal(fs["vertices"], fl["vertices"])
            brick = box_brick((0, 0, 0), (1, 1, 1), n, n, n, rank, world, grid=(1, 1, world))
            assert np.array_equal(brick.global_vertex[brick.cells], slab.global_vertex[slab.cells])
        elif mode == "brick":
            # 3-D bricks: corner ghosts are touched by up to 8 ranks
            assert brick_grid(8) == (2, 2, 2) and brick_grid(4) == (1, 2, 2) and brick_grid(2) == (1, 1, 2)
            local = box_brick((0, 0, 0), (1, 1, 1), n, n, n, rank, world)
            own_cells = torch.tensor([int((local.cell_owner == rank).sum())], dtype=torch.int64)
            dist.all_reduce(own_cells)
            assert int(own_cells) == g.num_cells                # the bricks tile the box
        else:
            # unstructured route: recursive coordinate bisection of the centroids of a jittered mesh
            cell_owner = rcb_partition(g.coords[g.cells].mean(axis=1), world)
            assert np.bincount(cell_owner, minlength=world).min() >= g.num_cells // world - 1
            local = extract_local(g.coords, g.cells, cell_owner, rank)
        fl = local.exterior_facets()
        # none of the remaining local-exterior facets lies inside the cube
        xm = local.coordinates[fl["vertices"]].mean(axis=1)
        on_bnd = (np.abs(xm) < 1e-12).any(axis=1) | (np.abs(xm - 1) < 1e-12).any(axis=1)
        assert on_bnd.all()

        space = FunctionSpace.from_host(local, J)
        part = Partition(rank, world).build(space, local)
        # partition of unity over dofs
        cnt = torch.tensor([part.n_owned_dofs()], dtype=torch.int64)
        dist.all_reduce(cnt)
        og = MPETOracle(g, PARAMS, dt=0.1, theta=0.5)
        assert int(cnt) == og.space.N

        # local oracle on the local mesh (same lexicographic local edge numbering as FunctionSpace.from_host)
        ol = MPETOracle(SimplexMesh(local.coordinates, local.cells), PARAMS, dt=0.1, theta=0.5)
        assert np.array_equal(ol.space.edge_vertices, space.edge_vertices())
        keys = node_global_keys(space, local)
        gnode = _global_node_index(og, keys, g.num_vertices)
        N2g, Nvg, N2, Nv = og.space.N2, og.space.Nv, space.N2, space.Nv
        gdof = np.concatenate([k * N2g + gnode for k in range(3)] +
                              [3 * N2g + i * Nvg + gnode[:Nv] for i in range(J)])
        Ag = og.assemble_lhs()
        Al = ol.assemble_lhs()
        x = np.random.default_rng(0).standard_normal(og.space.N)
        yg = Ag @ x
        yl = Al @ x[gdof]
        owned = part.owned_dofs.astype(bool)
        # owned rows are complete without any assembly collective
        assert np.allclose(yl[owned], yg[gdof][owned], rtol=1e-12, atol=1e-12)
        assert not np.allclose(yl[~owned], yg[gdof][~owned])          # ghost rows are not (as designed)
        # forward halo exchange with the send/recv lists
        for q in part.neighbours:
            sbuf = torch.from_numpy(yl[part.send[q]].copy())
            rbuf = torch.empty(part.recv[q].shape[0], dtype=torch.float64)
            if rank < q:
                dist.send(sbuf, q); dist.recv(rbuf, q)
            else:
                dist.recv(rbuf, q); dist.send(sbuf, q)
            yl[part.recv[q]] = rbuf.numpy()
        assert np.allclose(yl, yg[gdof], rtol=1e-12, atol=1e-12)
        # every ghost dof is received exactly once
        recv_all = np.concatenate([part.recv[q] for q in part.neighbours])
        assert np.unique(recv_all).size == recv_all.size == (~owned).sum()
        results[rank] = True
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n,mode", [(2, 6, "slab"), (4, 12, "slab"), (4, 6, "brick"), (8, 6, "brick"), (3, 5, "rcb")])
def test_partition_and_halo_lists(world, n, mode):
    port = _free_port()
    mgr = mp.get_context("spawn").Manager()      # no fork of a process that already runs OpenMP threads
    results = mgr.dict()
    mp.spawn(_worker, args=(world, port, n, results, mode), nprocs=world, join=True)
    assert all(results.get(r) for r in range(world))
