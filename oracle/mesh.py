"""Oracle meshes (test infrastructure; see oracle/__init__.py).

Conventions restated here are DOLFIN 2019.1's [EXT, SURVEY.md section 8c items 1-2]:

* ``UnitCubeMesh(n)`` == ``BoxMesh``: vertex id ``iz*(n+1)^2 + iy*(n+1) + ix``,
  six tets per cube, cell vertices sorted ascending (``mesh.order()``).
* ``UnitSquareMesh(n, n)`` with the default ``"right"`` diagonal: two
  triangles ``(v0, v1, v3)``, ``(v0, v2, v3)`` per square
  (used by ``src/mpet/test/test_convergence_mpetsolver.py:125``).
* edges are numbered in lexicographic order of their sorted vertex pair
  (DOLFIN's key-matching ``TopologyComputation``); an external numbering can be
  supplied instead.
* the reference's only mesh fixture, ``src/mpet/test/donut2D.h5``, is read with
  numpy at its fixed HDF5 dataset offsets (no h5py in this image).
"""
import numpy as np


class SimplexMesh:
    """coords f64[N_v, d], cells i32[N_c, d+1] (each row sorted ascending)."""

    def __init__(self, coords, cells):
        self.coords = np.ascontiguousarray(coords, dtype=np.float64)
        cells = np.sort(np.asarray(cells, dtype=np.int64), axis=1)
        self.cells = np.ascontiguousarray(cells)
        self.dim = self.coords.shape[1]
        assert self.cells.shape[1] == self.dim + 1
        self._edges = None
        self._facets = None

    @property
    def num_vertices(self):
        return self.coords.shape[0]

    @property
    def num_cells(self):
        return self.cells.shape[0]

    # UFC local edge -> (local vertex a, local vertex b), a < b
    @staticmethod
    def local_edges(dim):
        if dim == 2:
            return np.array([[1, 2], [0, 2], [0, 1]])
        return np.array([[2, 3], [1, 3], [1, 2], [0, 3], [0, 2], [0, 1]])

    def edges(self):
        """Return (edge_vertices i64[N_e,2], cell_edges i64[N_c, n_le])."""
        if self._edges is None:
            le = self.local_edges(self.dim)
            ev = self.cells[:, le]                      # [N_c, n_le, 2] (already lo<hi)
            key = ev[..., 0] * self.num_vertices + ev[..., 1]
            uniq, inv = np.unique(key.ravel(), return_inverse=True)
            edge_vertices = np.stack([uniq // self.num_vertices,
                                      uniq % self.num_vertices], axis=1)
            self._edges = (edge_vertices, inv.reshape(key.shape))
        return self._edges

    def exterior_facets(self):
        """Facets that belong to exactly one cell.

        Returns dict with
          vertices i64[N_f, d]  (sorted), cell i64[N_f], local i64[N_f]
        where ``local`` is the UFC local facet number (= index of the opposite
        vertex).  Facets are ordered by (cell, local).
        """
        if self._facets is None:
            d = self.dim
            nc = self.num_cells
            allf = []
            for opp in range(d + 1):
                keep = [k for k in range(d + 1) if k != opp]
                allf.append(self.cells[:, keep])
            allf = np.stack(allf, axis=1)               # [N_c, d+1, d]
            flat = allf.reshape(-1, d)
            nv = self.num_vertices
            assert nv ** d < 2 ** 62
            k = flat[:, 0].astype(np.int64)
            for j in range(1, d):
                k = k * nv + flat[:, j]
            uniq, inv, cnt = np.unique(k, return_inverse=True, return_counts=True)
            ext = np.nonzero(cnt[inv] == 1)[0]
            self._facets = dict(vertices=flat[ext], cell=ext // (d + 1),
                                local=ext % (d + 1))
        return self._facets

    def cell_volumes(self):
        x = self.coords[self.cells]
        J = np.swapaxes(x[:, 1:, :] - x[:, :1, :], 1, 2)
        fact = 2.0 if self.dim == 2 else 6.0
        return np.abs(np.linalg.det(J)) / fact


def unit_cube_mesh(n, jitter=0.0, seed=1234):
    """DOLFIN ``UnitCubeMesh(n, n, n)`` [EXT]; optional interior-vertex jitter
    ``jitter/n * U(-1,1)^3`` (SURVEY.md section 8d parity variant)."""
    m = n + 1
    iz, iy, ix = np.meshgrid(np.arange(m), np.arange(m), np.arange(m), indexing="ij")
    coords = np.stack([ix.ravel(), iy.ravel(), iz.ravel()], axis=1) / float(n)
    cells = []
    for kz in range(n):
        for ky in range(n):
            for kx in range(n):
                v0 = kz * m * m + ky * m + kx
                v1 = v0 + 1
                v2 = v0 + m
                v3 = v1 + m
                v4, v5, v6, v7 = v0 + m * m, v1 + m * m, v2 + m * m, v3 + m * m
                cells += [(v0, v1, v3, v7), (v0, v1, v7, v5), (v0, v5, v7, v4),
                          (v0, v3, v2, v7), (v0, v6, v4, v7), (v0, v2, v6, v7)]
    if jitter:
        rng = np.random.default_rng(seed)
        interior = np.all((coords > 1e-12) & (coords < 1 - 1e-12), axis=1)
        coords = coords.copy()
        coords[interior] += jitter / n * rng.uniform(-1, 1, size=(interior.sum(), 3))
    return SimplexMesh(coords, np.array(cells))


def unit_square_mesh(n):
    """DOLFIN ``UnitSquareMesh(n, n)`` with diagonal "right" [EXT]."""
    m = n + 1
    iy, ix = np.meshgrid(np.arange(m), np.arange(m), indexing="ij")
    coords = np.stack([ix.ravel(), iy.ravel()], axis=1) / float(n)
    cells = []
    for ky in range(n):
        for kx in range(n):
            v0 = ky * m + kx
            v1 = v0 + 1
            v2 = v0 + m
            v3 = v1 + m
            cells += [(v0, v1, v3), (v0, v2, v3)]
    return SimplexMesh(coords, np.array(cells))


def read_donut_h5(path):
    """Read ``/mesh/coordinates`` f8[784,2] and ``/mesh/topology`` i8[1449,3] of the
    reference fixture ``src/mpet/test/donut2D.h5`` (contiguous datasets at byte
    offsets 0x980 and 0x4280; checked by the caller via Euler characteristic and
    area)."""
    raw = open(path, "rb").read()
    assert raw[:8] == b"\x89HDF\r\n\x1a\n" and len(raw) == 63392
    coords = np.frombuffer(raw, dtype="<f8", count=784 * 2, offset=0x980).reshape(784, 2)
    topo = np.frombuffer(raw, dtype="<i8", count=1449 * 3, offset=0x4280).reshape(1449, 3)
    return SimplexMesh(coords.copy(), topo.copy())
