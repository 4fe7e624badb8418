"""Host side of the multi-GPU path: mesh partition by cells with a one-cell ghost layer, dof ownership,
halo index lists.  One process per GPU; torch.distributed is the plumbing that moves the
ncclUniqueId and the (small) interface key lists; all per-iteration traffic is NCCL inside the library.

Ownership rule: a node (vertex or edge) is owned by the lowest rank that owns a cell containing it.
Every rank keeps all cells that touch a node of its own cells (its cells + one ghost layer), so
  * the rows of its OWNED dofs are assembled completely from local cells (no assembly collective),
  * min(cell owner) over the local cells of a node is the global owner for every node of its own cells.
"""
import ctypes as C

import numpy as np
import torch

from .mpet.dolfin_shim import Mesh


class LocalMesh(Mesh):
    """A rank's part of a partitioned mesh: own cells + ghost layer.  ``global_vertex`` maps local
    vertex ids to global ones (monotone, so 'vertices ascending per cell' holds in both numberings);
    ``cell_owner`` gives the owning rank of every local cell; ``artificial`` flags local-exterior
    facets that are interior to the global mesh (partition cuts) -- they carry no boundary condition."""

    def __init__(self, coordinates, cells, global_vertex, cell_owner, n_global_vertices, artificial_fn=None):
        super().__init__(coordinates, cells)
        self.global_vertex = np.asarray(global_vertex, dtype=np.int64)
        self.cell_owner = np.asarray(cell_owner, dtype=np.int64)
        self.n_global_vertices = int(n_global_vertices)
        self._artificial_fn = artificial_fn

    def exterior_facets(self):
        if self._facets is None:
            F = super().exterior_facets()
            if self._artificial_fn is not None:
                keep = ~self._artificial_fn(self, F)
                F = {k: v[keep] for k, v in F.items()}
            self._facets = F
        return self._facets


def extract_local(coordinates, cells, cell_owner, rank):
    """General partition: local mesh of ``rank`` from a GLOBAL mesh and a cell->rank map (tests, small
    meshes).  Ghost layer = all cells sharing a vertex with an own cell."""
    cells = np.sort(np.asarray(cells, dtype=np.int64), axis=1)
    cell_owner = np.asarray(cell_owner, dtype=np.int64)
    nvg = coordinates.shape[0]
    own = cell_owner == rank
    touched = np.zeros(nvg, dtype=bool)
    touched[cells[own].ravel()] = True
    keep = touched[cells].any(axis=1)
    lc = cells[keep]
    gv = np.unique(lc)
    g2l = np.full(nvg, -1, dtype=np.int64)
    g2l[gv] = np.arange(gv.size)
    # global facet multiplicity decides which local-exterior facets are real boundary
    tri = np.stack([cells[:, [1, 2, 3]], cells[:, [0, 2, 3]], cells[:, [0, 1, 3]], cells[:, [0, 1, 2]]], axis=1).reshape(-1, 3)
    # two-stage keys ((a, b) -> id, then (id, c)): a single (a*nv + b)*nv + c product overflows int64 beyond ~2 M vertices
    pair = tri[:, 0] * nvg + tri[:, 1]
    upair, pid = np.unique(pair, return_inverse=True)
    key = pid.astype(np.int64) * nvg + tri[:, 2]
    uk, cnt = np.unique(key, return_counts=True)
    boundary_keys = uk[cnt == 1]

    def artificial(mesh, F):
        fv = mesh.global_vertex[F["vertices"]]
        pr = fv[:, 0] * nvg + fv[:, 1]
        pos = np.searchsorted(upair, pr)
        pos = np.minimum(pos, upair.shape[0] - 1)
        assert np.all(upair[pos] == pr)
        return ~np.isin(pos.astype(np.int64) * nvg + fv[:, 2], boundary_keys)

    return LocalMesh(coordinates[gv], g2l[lc], gv, cell_owner[keep], nvg, artificial)


def box_slab(p0, p1, nx, ny, nz, rank, nranks):
    """Slab partition of BoxMesh(p0, p1, nx, ny, nz) along z, generated directly per rank (the global
    mesh is never materialised): own cell layers [z0, z1) plus one ghost layer on each side."""
    assert nz >= 3 * nranks, "every slab needs at least three cell layers"
    bounds = [(nz * r) // nranks for r in range(nranks + 1)]
    z0, z1 = bounds[rank], bounds[rank + 1]
    lo, hi = max(z0 - 1, 0), min(z1 + 1, nz)
    mx, my = nx + 1, ny + 1
    p0 = np.asarray(p0, dtype=float)
    p1 = np.asarray(p1, dtype=float)
    izs = np.arange(lo, hi + 1)
    iz, iy, ix = np.meshgrid(izs, np.arange(my), np.arange(mx), indexing="ij")
    coords = np.stack([p0[0] + (p1[0] - p0[0]) * ix.ravel() / nx,
                       p0[1] + (p1[1] - p0[1]) * iy.ravel() / ny,
                       p0[2] + (p1[2] - p0[2]) * iz.ravel() / nz], axis=1)
    gv = (iz.ravel().astype(np.int64) * my + iy.ravel()) * mx + ix.ravel()
    kz, ky, kx = np.meshgrid(np.arange(hi - lo), np.arange(ny), np.arange(nx), indexing="ij")
    v0 = (kz * my * mx + ky * mx + kx).ravel().astype(np.int64)
    v1, v2 = v0 + 1, v0 + mx
    v3 = v1 + mx
    v4, v5, v6, v7 = v0 + mx * my, v1 + mx * my, v2 + mx * my, v3 + mx * my
    tets = np.stack([np.stack(t, axis=1) for t in
                     [(v0, v1, v3, v7), (v0, v1, v7, v5), (v0, v5, v7, v4),
                      (v0, v3, v2, v7), (v0, v6, v4, v7), (v0, v2, v6, v7)]], axis=1).reshape(-1, 4)
    layer = np.repeat(kz.ravel() + lo, 6)
    owner = np.searchsorted(np.asarray(bounds[1:]), layer, side="right")
    zlo = p0[2] + (p1[2] - p0[2]) * lo / nz if lo > 0 else None
    zhi = p0[2] + (p1[2] - p0[2]) * hi / nz if hi < nz else None

    def artificial(mesh, F):
        z = mesh.coordinates[F["vertices"]][:, :, 2]
        out = np.zeros(z.shape[0], dtype=bool)
        if zlo is not None:
            out |= np.all(z == zlo, axis=1)
        if zhi is not None:
            out |= np.all(z == zhi, axis=1)
        return out

    return LocalMesh(coords, tets, gv, owner, mx * my * (nz + 1), artificial)


def brick_grid(nranks):
    """(px, py, pz) with px*py*pz == nranks, as cubic as possible, the longest side along z (SURVEY.md 8e:
    "contiguous bricks"): 2 -> (1,1,2), 4 -> (1,2,2), 8 -> (2,2,2)."""
    best = None
    for px in range(1, nranks + 1):
        if nranks % px:
            continue
        for py in range(px, nranks // px + 1):
            if (nranks // px) % py:
                continue
            pz = nranks // (px * py)
            if pz < py:
                continue
            surf = px * py + py * pz + px * pz
            if best is None or surf < best[0]:
                best = (surf, (px, py, pz))
    return best[1]


def box_brick(p0, p1, nx, ny, nz, rank, nranks, grid=None):
    """Brick partition of BoxMesh(p0, p1, nx, ny, nz), generated directly per rank: the process grid
    (px, py, pz) cuts the cube ranges of every axis evenly; rank = (bz*py + by)*px + bx owns the cubes of its
    brick and holds one ghost layer of cubes around it (<= 26 neighbours).  box_slab is the (1, 1, P) case."""
    px, py, pz = grid or brick_grid(nranks)
    assert px * py * pz == nranks
    dims = (nx, ny, nz)
    parts = (px, py, pz)
    for d in range(3):
        assert parts[d] == 1 or dims[d] >= 3 * parts[d], "every brick needs at least three cell layers per cut axis"
    b = (rank % px, (rank // px) % py, rank // (px * py))
    bounds = [[(dims[d] * r) // parts[d] for r in range(parts[d] + 1)] for d in range(3)]
    own = [(bounds[d][b[d]], bounds[d][b[d] + 1]) for d in range(3)]
    ext = [(max(own[d][0] - 1, 0), min(own[d][1] + 1, dims[d])) for d in range(3)]
    mx, my = nx + 1, ny + 1
    p0 = np.asarray(p0, dtype=float)
    p1 = np.asarray(p1, dtype=float)
    ixs, iys, izs = (np.arange(ext[d][0], ext[d][1] + 1) for d in range(3))
    iz, iy, ix = np.meshgrid(izs, iys, ixs, indexing="ij")
    coords = np.stack([p0[0] + (p1[0] - p0[0]) * ix.ravel() / nx,
                       p0[1] + (p1[1] - p0[1]) * iy.ravel() / ny,
                       p0[2] + (p1[2] - p0[2]) * iz.ravel() / nz], axis=1)
    gv = (iz.ravel().astype(np.int64) * my + iy.ravel()) * mx + ix.ravel()
    lx, ly, lz = (ext[d][1] - ext[d][0] for d in range(3))
    kz, ky, kx = np.meshgrid(np.arange(lz), np.arange(ly), np.arange(lx), indexing="ij")
    sx, sy = lx + 1, (lx + 1) * (ly + 1)
    v0 = (kz * sy + ky * sx + kx).ravel().astype(np.int64)
    v1, v2 = v0 + 1, v0 + sx
    v3 = v1 + sx
    v4, v5, v6, v7 = v0 + sy, v1 + sy, v2 + sy, v3 + sy
    tets = np.stack([np.stack(t, axis=1) for t in
                     [(v0, v1, v3, v7), (v0, v1, v7, v5), (v0, v5, v7, v4),
                      (v0, v3, v2, v7), (v0, v6, v4, v7), (v0, v2, v6, v7)]], axis=1).reshape(-1, 4)
    brick_of = [np.searchsorted(np.asarray(bounds[d][1:]), (k.ravel() + ext[d][0]), side="right")
                for d, k in ((0, kx), (1, ky), (2, kz))]
    owner = np.repeat((brick_of[2] * py + brick_of[1]) * px + brick_of[0], 6)
    cuts = []          # (axis, coordinate) of the local-exterior planes that are interior to the global box
    for d in range(3):
        if ext[d][0] > 0:
            cuts.append((d, p0[d] + (p1[d] - p0[d]) * ext[d][0] / dims[d]))
        if ext[d][1] < dims[d]:
            cuts.append((d, p0[d] + (p1[d] - p0[d]) * ext[d][1] / dims[d]))

    def artificial(mesh, F):
        xf = mesh.coordinates[F["vertices"]]
        out = np.zeros(xf.shape[0], dtype=bool)
        for d, c in cuts:
            out |= np.all(xf[:, :, d] == c, axis=1)
        return out

    return LocalMesh(coords, tets, gv, owner, mx * my * (nz + 1), artificial)


def rcb_partition(points, nparts):
    """Recursive coordinate bisection of cell centroids (unstructured meshes): returns the part
    of every point.  Splits the longest extent at the weighted median, parts sized proportionally when nparts
    is not a power of two; deterministic (stable sorts)."""
    points = np.asarray(points, dtype=float)
    part = np.zeros(points.shape[0], dtype=np.int64)

    def split(idx, first, count):
        if count == 1:
            part[idx] = first
            return
        left = count // 2
        ext = points[idx].max(axis=0) - points[idx].min(axis=0)
        axis = int(np.argmax(ext))
        order = idx[np.argsort(points[idx, axis], kind="stable")]
        k = (order.shape[0] * left) // count
        split(order[:k], first, left)
        split(order[k:], first + left, count - left)

    split(np.arange(points.shape[0]), 0, int(nparts))
    return part


def node_owners(space, mesh):
    """Owner rank of every scalar P2 node (vertices, then edges) = min owner over its local cells."""
    cells = mesh.cells.astype(np.int64)
    co = mesh.cell_owner
    big = np.iinfo(np.int64).max
    own = np.full(space.N2, big, dtype=np.int64)
    np.minimum.at(own, cells.ravel(), np.repeat(co, 4))
    for a, b in ((2, 3), (1, 3), (1, 2), (0, 3), (0, 2), (0, 1)):
        e = space.Nv + space.edge_index(cells[:, a], cells[:, b])
        np.minimum.at(own, e, co)
    return own


def own_cell_node_mask(space, mesh, rank):
    """Nodes (vertices, then edges) touched by at least one cell this rank OWNS."""
    cells = mesh.cells.astype(np.int64)
    mine = mesh.cell_owner == rank
    touched = np.zeros(space.N2, dtype=bool)
    touched[cells[mine].ravel()] = True
    for a, b in ((2, 3), (1, 3), (1, 2), (0, 3), (0, 2), (0, 1)):
        touched[space.Nv + space.edge_index(cells[mine, a], cells[mine, b])] = True
    return touched


def resolve_node_owners(space, mesh, rank, nranks, group=None):
    """True owner of every local node = lowest rank owning a cell that contains it.

    min(cell owner) over the LOCAL cells (node_owners) is right for the nodes of a rank's own cells, but not
    for the outer nodes of its ghost layer when a third, lower rank also touches them (brick corners, any
    unstructured partition with >= 3 ranks).  Every rank therefore publishes the keys of the nodes of its OWN
    cells that are shared with a cell of another rank (its interface set I_r, O(N^(2/3)) keys); the owner of a
    local node X is min{r : X in I_r}, or the local minimum when X is in no interface set (then every cell that
    contains X belongs to one rank, and this rank holds one of them)."""
    import torch.distributed as dist
    own = node_owners(space, mesh)
    keys = node_global_keys(space, mesh)
    mine = own_cell_node_mask(space, mesh, rank)
    cells = mesh.cells.astype(np.int64)
    other = mesh.cell_owner != rank
    shared = np.zeros(space.N2, dtype=bool)
    shared[cells[other].ravel()] = True
    for a, b in ((2, 3), (1, 3), (1, 2), (0, 3), (0, 2), (0, 1)):
        shared[space.Nv + space.edge_index(cells[other, a], cells[other, b])] = True
    iface = np.sort(keys[mine & shared])
    sets = [None] * nranks
    dist.all_gather_object(sets, iface, group=group)
    big = np.iinfo(np.int64).max
    best = np.full(space.N2, big, dtype=np.int64)
    for r in range(nranks):
        Ir = sets[r]
        if Ir is None or len(Ir) == 0:
            continue
        pos = np.minimum(np.searchsorted(Ir, keys), len(Ir) - 1)
        hit = Ir[pos] == keys
        best[hit] = np.minimum(best[hit], r)
    return np.where(best < big, best, own)


def node_global_keys(space, mesh):
    """Rank-independent integer key of every local P2 node."""
    gv = mesh.global_vertex
    nvg = mesh.n_global_vertices
    ev = space.edge_vertices()
    return np.concatenate([gv, nvg + gv[ev[:, 0]] * nvg + gv[ev[:, 1]]])


class Partition:
    """Ownership + halo plan of one rank; ``setup`` pushes it into the engine."""

    def __init__(self, rank, nranks, group=None):
        self.rank, self.nranks, self.group = rank, nranks, group

    def build(self, space, mesh):
        import torch.distributed as dist
        rank = self.rank
        own = resolve_node_owners(space, mesh, rank, self.nranks, self.group)
        keys = node_global_keys(space, mesh)
        self.node_owner = own
        self.owned_nodes = own == rank
        # what I need from each owner: keys of my ghost nodes, grouped by owner, ascending key
        need = {}
        for q in np.unique(own[own != rank]):
            sel = np.nonzero(own == q)[0]
            order = np.argsort(keys[sel])
            need[int(q)] = (sel[order], keys[sel][order])
        requests = [None] * self.nranks
        dist.all_gather_object(requests, {q: k for q, (_, k) in need.items()}, group=self.group)
        order = np.argsort(keys)
        skeys = keys[order]
        send_nodes = {}
        for q in range(self.nranks):
            if q != rank and requests[q] and rank in requests[q]:
                want = requests[q][rank]
                pos = np.searchsorted(skeys, want)
                assert np.all(skeys[pos] == want), "neighbour asks for a node this rank does not hold"
                nodes = order[pos]
                assert np.all(own[nodes] == rank), "ownership disagreement between ranks"
                send_nodes[q] = nodes
        self.neighbours = sorted(set(need) | set(send_nodes))
        Nv, N2, J = space.Nv, space.N2, space.J

        def dofs_of(nodes):
            verts = nodes[nodes < Nv]
            parts = [k * N2 + nodes for k in range(3)] + [3 * N2 + i * Nv + verts for i in range(J)]
            return np.concatenate(parts).astype(np.int32)

        empty = np.zeros(0, dtype=np.int64)
        self.send_nodes = {q: send_nodes.get(q, empty).astype(np.int32) for q in self.neighbours}
        self.recv_nodes = {q: (need[q][0] if q in need else empty).astype(np.int32) for q in self.neighbours}
        # dof-level lists in the UFC numbering (host-side checks; the library derives its own from the nodes)
        self.send = {q: dofs_of(send_nodes.get(q, np.zeros(0, dtype=np.int64))) for q in self.neighbours}
        self.recv = {q: dofs_of(need[q][0] if q in need else np.zeros(0, dtype=np.int64)) for q in self.neighbours}
        owned = np.zeros(space.N, dtype=np.uint8)
        on = self.owned_nodes
        for k in range(3):
            owned[k * N2:(k + 1) * N2] = on
        for i in range(J):
            owned[3 * N2 + i * Nv: 3 * N2 + (i + 1) * Nv] = on[:Nv]
        self.owned_dofs = owned
        return self

    def setup(self, engine, space, mesh):
        import torch.distributed as dist
        self.build(space, mesh)
        uid = [None]
        if self.rank == 0:
            uid[0] = engine.nccl_unique_id()
        dist.broadcast_object_list(uid, src=0, group=self.group)
        engine.attach_comm(uid[0], self.rank, self.nranks)
        engine.set_halo(self.neighbours, [self.send_nodes[q] for q in self.neighbours],
                        [self.recv_nodes[q] for q in self.neighbours], self.owned_nodes.astype(np.uint8))
        return self

    def n_owned_dofs(self):
        return int(self.owned_dofs.sum())
