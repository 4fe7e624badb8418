"""Oracle finite elements, quadrature and dof maps (test infrastructure).

Restates [EXT: FIAT / UFC conventions that DOLFIN 2019.1 uses, SURVEY.md 8a1, 8c5]:

* Lagrange P_k on the UFC reference simplex (vertices 0, e_1, .., e_d); P1 and
  P2 node order = vertices, then edge midpoints in UFC local-edge order.
* UFC global numbering of the mixed space built by
  ``MPETSolver.create_function_spaces`` (``mpetsolver.py:100-132``):
  component k of the P2 vector at ``k*N2 + node``, pressure i at
  ``d*N2 + i*N_v + vertex``; the optional Real (Lagrange multiplier) dofs
  follow.  An external permutation (e.g. a recorded DOLFIN numbering) can be
  applied on top.
* collapsed Gauss-Jacobi quadrature, exact to a requested degree (FFC picks
  the degree as the sum of the integrand's polynomial degrees; any exact rule
  gives the same integrals up to round-off).
"""
import itertools
import numpy as np
from scipy.special import roots_jacobi

from .mesh import SimplexMesh


# --------------------------------------------------------------------------- quadrature
def simplex_quadrature(dim, degree):
    """Points [n_q, dim] / weights [n_q] on the reference simplex, exact for
    polynomials of total degree <= ``degree``."""
    n = max(1, (degree + 2) // 2)
    rules = []
    for a in range(dim - 1, -1, -1):
        x, w = roots_jacobi(n, a, 0)
        rules.append(((x + 1) / 2, w / 2 ** (a + 1)))
    pts, wts = [], []
    for idx in itertools.product(range(n), repeat=dim):
        u = [rules[k][0][idx[k]] for k in range(dim)]
        w = np.prod([rules[k][1][idx[k]] for k in range(dim)])
        x = []
        scale = 1.0
        for k in range(dim):
            x.append(u[k] * scale)
            scale *= (1 - u[k])
        pts.append(x)
        wts.append(w)
    return np.array(pts), np.array(wts)


# --------------------------------------------------------------------------- Lagrange
def lagrange_nodes(dim, degree):
    """Nodes of P_k.  degree 0: barycentre.  degree 1, 2: UFC order (vertices,
    edge midpoints).  degree >= 3: lattice order (only used cell-wise)."""
    if degree == 0:
        return np.full((1, dim), 1.0 / (dim + 1))
    verts = np.vstack([np.zeros(dim), np.eye(dim)])
    if degree == 1:
        return verts
    if degree == 2:
        le = SimplexMesh.local_edges(dim)
        return np.vstack([verts, 0.5 * (verts[le[:, 0]] + verts[le[:, 1]])])
    pts = [np.array(i, dtype=float) / degree
           for i in itertools.product(range(degree + 1), repeat=dim) if sum(i) <= degree]
    return np.array(pts)


def _monomials(dim, degree):
    return [e for e in itertools.product(range(degree + 1), repeat=dim) if sum(e) <= degree]


def tabulate(dim, degree, pts):
    """Return (values [n_pts, n_basis], grads [n_pts, n_basis, dim]) of the P_k
    nodal basis on the reference simplex."""
    pts = np.atleast_2d(pts)
    expo = _monomials(dim, degree)
    nodes = lagrange_nodes(dim, degree)

    def vander(x):
        return np.stack([np.prod(x ** np.array(e), axis=1) for e in expo], axis=1)

    C = np.linalg.inv(vander(nodes))           # columns = basis coefficients
    vals = vander(pts) @ C
    grads = np.zeros((pts.shape[0], len(expo), dim))
    for k in range(dim):
        dV = np.zeros((pts.shape[0], len(expo)))
        for j, e in enumerate(expo):
            if e[k] == 0:
                continue
            ee = np.array(e)
            ee[k] -= 1
            dV[:, j] = e[k] * np.prod(pts ** ee, axis=1)
        grads[:, :, k] = dV @ C
    return vals, grads


# --------------------------------------------------------------------------- dof map
class MixedSpace:
    """[P2]^d x [P1]^A (x R^nreal) on a SimplexMesh -- ``mpetsolver.py:100-132``."""

    def __init__(self, mesh, A, nreal=0, perm=None):
        self.mesh = mesh
        self.A = A
        d = mesh.dim
        self.dim = d
        edge_vertices, cell_edges = mesh.edges()
        self.edge_vertices = edge_vertices
        self.Nv = mesh.num_vertices
        self.Ne = edge_vertices.shape[0]
        self.N2 = self.Nv + self.Ne
        self.n2 = (d + 1) + cell_edges.shape[1]           # local scalar P2 dofs
        self.n1 = d + 1
        self.nloc = d * self.n2 + A * self.n1
        self.nfe = d * self.N2 + A * self.Nv              # field dofs
        self.nreal = nreal
        self.N = self.nfe + nreal
        # scalar P2 node map per cell: vertices then edges
        self.cell_nodes2 = np.concatenate([mesh.cells, self.Nv + cell_edges], axis=1)
        blocks = [k * self.N2 + self.cell_nodes2 for k in range(d)]
        blocks += [d * self.N2 + i * self.Nv + mesh.cells for i in range(A)]
        cd = np.concatenate(blocks, axis=1)
        self.perm = None
        if perm is not None:
            self.perm = np.asarray(perm, dtype=np.int64)
            cd = self.perm[cd]
        self.cell_dofs = cd

    def _p(self, idx):
        return idx if self.perm is None else self.perm[idx]

    def u_dofs(self, k):
        return self._p(k * self.N2 + np.arange(self.N2))

    def p_dofs(self, i):
        return self._p(self.dim * self.N2 + i * self.Nv + np.arange(self.Nv))

    def real_dofs(self):
        return self.nfe + np.arange(self.nreal)

    def node2_coords(self):
        """Coordinates of the scalar P2 nodes (vertices, then edge midpoints)."""
        x = self.mesh.coords
        ev = self.edge_vertices
        return np.vstack([x, 0.5 * (x[ev[:, 0]] + x[ev[:, 1]])])

    def facet_nodes2(self, facets):
        """Scalar P2 nodes (vertices + edge midpoints) of each facet: [N_f, n]."""
        fv = facets["vertices"]
        d = self.dim
        pairs = [(a, b) for a in range(d) for b in range(a + 1, d)]
        keys_all = self.edge_vertices[:, 0] * self.Nv + self.edge_vertices[:, 1]
        cols = [fv]
        for a, b in pairs:
            key = fv[:, a] * self.Nv + fv[:, b]
            e = np.searchsorted(keys_all, key)
            assert np.all(keys_all[e] == key)
            cols.append(self.Nv + e[:, None])
        return np.concatenate(cols, axis=1)
