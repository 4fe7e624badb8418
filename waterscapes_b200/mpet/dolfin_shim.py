"""Minimal stand-ins for the DOLFIN objects the reference's mpet API touches on the hot path.

The reference does ``from dolfin import *`` (src/mpet/mpet/mpetsolver.py:7) and hands DOLFIN objects
across its API: ``Mesh``, ``Constant``, ``Expression``, facet ``MeshFunction``, ``FacetNormal``,
``CompiledSubDomain``, ``Function`` (``.vector()``, ``.split()``, ``.assign``), ``DirichletBC``,
``Parameters``, ``assemble``.  DOLFIN is not a dependency here: these classes carry exactly the data
the B200 path needs (host numpy for O(N^(2/3)) boundary data, torch CUDA tensors for fields).
They are data holders -- no finite-element arithmetic happens in this file.
"""
import math
import sys

import numpy as np
import torch

INVALID = sys.maxsize   # mpetproblem.py:153


def info(msg):
    pass


def warning(msg):
    sys.stderr.write("*** Warning: %s\n" % msg)


# ----------------------------------------------------------------------------------------- mesh
class Mesh:
    """Tetrahedral mesh: coordinates f64[Nv,3], cells i32[Nc,4] (vertices ascending per cell)."""

    def __init__(self, coordinates=None, cells=None):
        self._facets = None
        if coordinates is None:         # ``mesh = Mesh(); HDF5File(...).read(mesh, "/mesh", False)`` (hdf5.py)
            self.coordinates = np.zeros((0, 3))
            self.cells = np.zeros((0, 4), dtype=np.int32)
            return
        self.coordinates = np.ascontiguousarray(coordinates, dtype=np.float64)
        self.cells = np.ascontiguousarray(np.sort(np.asarray(cells), axis=1), dtype=np.int32)
        assert self.coordinates.shape[1] == 3 and self.cells.shape[1] == 4, "the B200 path is 3-D (tets)"
        self._facets = None

    def geometry(self):
        return self

    def topology(self):
        return self

    def dim(self):
        return 3

    def num_vertices(self):
        return self.coordinates.shape[0]

    def num_cells(self):
        return self.cells.shape[0]

    def ufl_cell(self):
        return "tetrahedron"

    def exterior_facets(self):
        """dict(vertices i64[Nf,3] ascending, cell i64[Nf], local i64[Nf]) ordered by (cell, local)."""
        if self._facets is None:
            c = self.cells.astype(np.int64)
            nv = self.num_vertices()
            tri = np.stack([c[:, [1, 2, 3]], c[:, [0, 2, 3]], c[:, [0, 1, 3]], c[:, [0, 1, 2]]], axis=1)
            flat = tri.reshape(-1, 3)
            # two-stage key (pair id, third vertex): an nv^3 product overflows int64 beyond ~2.09 M vertices
            _, pid = np.unique(flat[:, 0] * nv + flat[:, 1], return_inverse=True)
            key = pid.astype(np.int64).ravel() * nv + flat[:, 2]
            order = np.argsort(key, kind="stable")
            ks = key[order]
            single = np.ones(ks.shape[0], dtype=bool)
            same = ks[1:] == ks[:-1]
            single[1:] &= ~same
            single[:-1] &= ~same
            ext = np.sort(order[single])
            self._facets = dict(vertices=flat[ext], cell=ext // 4, local=ext % 4)
        return self._facets


def BoxMesh(p0, p1, nx, ny, nz):
    """DOLFIN BoxMesh: vertex id iz*(nx+1)(ny+1) + iy*(nx+1) + ix, six tets per brick [EXT]."""
    mx, my, mz = nx + 1, ny + 1, nz + 1
    iz, iy, ix = np.meshgrid(np.arange(mz), np.arange(my), np.arange(mx), indexing="ij")
    p0 = np.asarray(p0, dtype=float)
    p1 = np.asarray(p1, dtype=float)
    coords = np.stack([p0[0] + (p1[0] - p0[0]) * ix.ravel() / nx,
                       p0[1] + (p1[1] - p0[1]) * iy.ravel() / ny,
                       p0[2] + (p1[2] - p0[2]) * iz.ravel() / nz], axis=1)
    kz, ky, kx = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    v0 = (kz * my * mx + ky * mx + kx).ravel().astype(np.int64)
    v1, v2 = v0 + 1, v0 + mx
    v3 = v1 + mx
    v4, v5, v6, v7 = v0 + mx * my, v1 + mx * my, v2 + mx * my, v3 + mx * my
    tets = np.stack([np.stack(t, axis=1) for t in
                     [(v0, v1, v3, v7), (v0, v1, v7, v5), (v0, v5, v7, v4),
                      (v0, v3, v2, v7), (v0, v6, v4, v7), (v0, v2, v6, v7)]], axis=1).reshape(-1, 4)
    return Mesh(coords, tets)


def UnitCubeMesh(nx, ny=None, nz=None):
    ny = nx if ny is None else ny
    nz = nx if nz is None else nz
    return BoxMesh((0.0, 0.0, 0.0), (1.0, 1.0, 1.0), nx, ny, nz)


# ----------------------------------------------------------------------------------------- data
class Constant:
    def __init__(self, value):
        self._v = np.asarray(value, dtype=float)

    def assign(self, value):
        self._v = np.asarray(float(value) if np.ndim(value) == 0 else value, dtype=float)

    def __float__(self):
        return float(self._v)

    def values(self):
        return np.atleast_1d(self._v)

    def value_shape(self):
        return self._v.shape

    def __mul__(self, other):
        if isinstance(other, FacetNormal):
            return NormalProduct(self)
        return NotImplemented

    def eval_points(self, x):
        return np.broadcast_to(self._v, (x.shape[0],) + self._v.shape).copy()

    @property
    def is_constant(self):
        return True


class CellFunction:
    """A DG0 coefficient: one value per cell (``Function(FunctionSpace(mesh, "DG", 0))`` in the reference's
    sandbox/biot-robin/three_fields_precond.py:263-271).  Accepted for the permeabilities ``params["K"][i]``."""

    def __init__(self, mesh, values):
        self.mesh = mesh
        self.values = np.ascontiguousarray(values, dtype=float).reshape(-1)
        assert self.values.shape[0] == mesh.num_cells()

    def vector(self):
        return self.values

    def __float__(self):
        raise TypeError("a DG0 coefficient has no single value")


_NS = {k: getattr(np, k) for k in ("sin", "cos", "tan", "exp", "log", "sqrt", "tanh", "sinh", "cosh",
                                   "arctan", "arcsin", "arccos", "abs")}
_NS.update(pi=math.pi, pow=np.power, fabs=np.abs, atan=np.arctan, asin=np.arcsin, acos=np.arccos,
           DOLFIN_PI=math.pi, M_PI=math.pi)


class _X:
    """`x[i]` of a DOLFIN C++ expression string."""

    def __init__(self, pts):
        self.pts = pts

    def __getitem__(self, i):
        return self.pts[:, i]


class Expression:
    """``Expression("sin(x[0])*t", t=time, degree=2)`` -- the C++ string is evaluated with numpy on
    arrays of points; user parameters may be numbers or ``Constant`` objects (read at call time, so a
    shared time Constant works as in the reference).  A Python callable ``f(x, **params)`` is also
    accepted in place of the string."""

    def __init__(self, code, degree=None, **params):
        self.code = code
        self.degree = degree
        self.params = params

    def value_shape(self):
        c = self.code
        if isinstance(c, (tuple, list)):
            if isinstance(c[0], (tuple, list)):
                return (len(c), len(c[0]))
            return (len(c),)
        return ()

    @property
    def is_constant(self):
        return False

    def _params(self):
        return {k: (float(v) if isinstance(v, Constant) else v) for k, v in self.params.items()}

    def _eval_one(self, code, x, ns):
        if callable(code):
            return np.broadcast_to(np.asarray(code(x, **self._params()), dtype=float), (x.shape[0],)).copy()
        src = code.replace("&&", " and ").replace("||", " or ")
        val = eval(src, {"__builtins__": {}}, ns)
        return np.broadcast_to(np.asarray(val, dtype=float), (x.shape[0],)).copy()

    def eval_points(self, x):
        ns = dict(_NS)
        ns.update(self._params())
        ns["x"] = _X(x)
        c = self.code
        if callable(c) and not isinstance(c, (tuple, list)):
            return np.asarray(c(x, **self._params()), dtype=float)
        if isinstance(c, (tuple, list)):
            if isinstance(c[0], (tuple, list)):
                return np.stack([np.stack([self._eval_one(e, x, ns) for e in row], axis=1) for row in c], axis=1)
            return np.stack([self._eval_one(e, x, ns) for e in c], axis=1)
        return self._eval_one(c, x, ns)

    def __mul__(self, other):
        if isinstance(other, FacetNormal):
            return NormalProduct(self)
        return NotImplemented


class FacetNormal:
    def __init__(self, mesh):
        self.mesh = mesh


class NormalProduct:
    """``expr * n``: scalar expr -> expr n ; matrix expr -> expr . n  (test_donut.py:46,
    test_convergence_mpetsolver.py:137-138)."""

    def __init__(self, expr):
        self.expr = expr

    @property
    def is_constant(self):
        return False


class MeshFunction:
    """Facet markers (``MeshFunction("size_t", mesh, 2)``), stored over EXTERIOR facets only --
    the only facets the MPET forms integrate over (ds) or constrain."""

    def __init__(self, value_type, mesh, dim, value=None):
        assert dim == 2, "only facet markers are used by the mpet API"
        self.mesh = mesh
        self.array_ = np.full(mesh.exterior_facets()["cell"].shape[0], INVALID if value is None else value,
                              dtype=np.int64)

    def set_all(self, v):
        self.array_[:] = v

    def array(self):
        return self.array_


DOLFIN_EPS = 3.0e-16


def near(a, b, eps=DOLFIN_EPS):
    """DOLFIN's ``near(a, b, eps)``; works on numbers and on arrays of coordinates."""
    return np.abs(np.asarray(a) - b) < eps


class _Points(np.ndarray):
    """Coordinates of npts points [npts, 3] handed to a user-written ``SubDomain.inside(x, on_boundary)``.
    DOLFIN calls ``inside`` with ONE point, so reference-style code writes ``x[0]`` for the first COORDINATE;
    an integer index therefore selects a column here (the coordinate of all points), not a row."""

    def __getitem__(self, i):
        if isinstance(i, (int, np.integer)):
            return np.asarray(self)[:, i]
        return np.asarray(self)[i]


class SubDomain:
    """Boundary predicate.  ``inside(x, on_boundary)`` as in DOLFIN: ``x[k]`` is the k-th coordinate.  It is
    evaluated on all points of the boundary facets at once (``x[k]`` is then an array); an ``inside`` that only
    works on one point (``and`` / ``or`` / ``if`` on coordinates) is detected and called point by point."""

    def inside(self, x, on_boundary):
        raise NotImplementedError

    def _inside_points(self, pts):
        pts = np.ascontiguousarray(pts, dtype=float)
        try:
            r = np.asarray(self.inside(pts.view(_Points), True))
            if r.dtype != object and r.shape in ((), (pts.shape[0],)):
                return np.broadcast_to(r.astype(bool), (pts.shape[0],))
        except (ValueError, TypeError):      # "truth value of an array is ambiguous": scalar-style predicate
            pass
        return np.fromiter((bool(self.inside(p, True)) for p in pts), dtype=bool, count=pts.shape[0])

    def mark(self, markers, value):
        mesh = markers.mesh
        fv = mesh.exterior_facets()["vertices"]
        xs = mesh.coordinates[fv]                                 # [Nf, 3, 3]
        ok = np.ones(fv.shape[0], dtype=bool)
        for k in range(3):                                        # all vertices and the midpoint inside
            ok &= self._inside_points(xs[:, k])
        ok &= self._inside_points(xs.mean(axis=1))
        markers.array_[ok] = value


class CompiledSubDomain(SubDomain):
    """``CompiledSubDomain("on_boundary && near(x[0], 1.0)")`` evaluated with numpy."""

    def __init__(self, code, **params):
        self.code = code
        self.params = params

    def inside(self, x, on_boundary):
        ns = dict(_NS)
        ns.update(self.params)
        ns["x"] = _X(x)
        ns["on_boundary"] = np.full(x.shape[0], bool(on_boundary))
        ns["near"] = lambda a, b, eps=3e-16: np.abs(a - b) < eps
        ns["DOLFIN_EPS"] = 3e-16
        src = self.code.replace("&&", "&").replace("||", "|")
        src = _parenthesise_comparisons(src)
        return np.broadcast_to(np.asarray(eval(src, {"__builtins__": {}}, ns), dtype=bool), (x.shape[0],))


def _parenthesise_comparisons(src):
    # "a < b & c > d" must become "(a < b) & (c > d)" for numpy's operator precedence
    parts = []
    for disj in src.split("|"):
        parts.append(" & ".join("(" + t.strip() + ")" for t in disj.split("&")))
    return " | ".join("(" + p + ")" for p in parts)


# ----------------------------------------------------------------------------------------- params
class Parameters(dict):
    """DOLFIN ``Parameters``: add / update / [] (mpetsolver.py:86-98)."""

    def __init__(self, name="parameters"):
        super().__init__()
        self.name = name

    def add(self, key, value):
        self[key] = value

    def update(self, other):
        for k, v in dict(other).items():
            self[k] = v


parameters = {"form_compiler": Parameters("form_compiler")}   # accepted and ignored (test_donut.py:12-14)


# ----------------------------------------------------------------------------------------- fields
class Vector:
    """``Function.vector()``: the dof vector, resident on the GPU."""

    def __init__(self, tensor):
        self.t = tensor

    def size(self):
        return self.t.numel()

    def get_local(self):
        return self.t.detach().cpu().numpy()

    def set_local(self, arr):
        self.t.copy_(torch.as_tensor(np.asarray(arr, dtype=float), device=self.t.device))

    def __getitem__(self, idx):
        return self.get_local()[idx]

    def __setitem__(self, idx, value):
        if isinstance(idx, slice) and idx == slice(None):
            if np.ndim(value) == 0:
                self.t.fill_(float(value))
            else:
                self.set_local(value)
        else:
            a = self.get_local()
            a[idx] = value
            self.set_local(a)

    def norm(self, kind="l2"):
        if kind == "l2":
            return float(torch.linalg.vector_norm(self.t))
        if kind == "linf":
            return float(self.t.abs().max())
        raise ValueError(kind)

    def axpy(self, a, other):
        self.t.add_(other.t, alpha=float(a))

    def copy(self):
        return Vector(self.t.clone())


class SubSpace:
    def __init__(self, space, index):
        self.space, self.index = space, index

    def collapse(self):
        return self

    def sub(self, i):
        return self


class FunctionSpace:
    """Layout of [P2]^3 x [P1]^J (UFC numbering, include/mpet_b200.h)."""

    def __init__(self, mesh, J, engine, nreal=0):
        self.mesh, self.J, self.engine = mesh, J, engine
        self.nreal = int(nreal)          # Real-space (Lagrange multiplier) dofs appended behind the N finite-element dofs
        s = engine.sizes
        self.Nv, self.Ne, self.N2, self.N = s["Nv"], s["Ne"], s["N2"], s["N"]
        self._edges = None
        self._x2 = None

    @classmethod
    def from_host(cls, mesh, J):
        """Same layout computed on the host (no engine): edges numbered lexicographically by their
        sorted vertex pair, exactly what graph.cu builds on the device.  Used by host-only logic/tests."""
        self = cls.__new__(cls)
        self.mesh, self.J, self.engine = mesh, J, None
        self.nreal = 0
        c = mesh.cells.astype(np.int64)
        nv = mesh.num_vertices()
        pairs = np.concatenate([c[:, [a, b]] for a, b in ((2, 3), (1, 3), (1, 2), (0, 3), (0, 2), (0, 1))])
        keys = np.unique(pairs[:, 0] * nv + pairs[:, 1])
        self._edges = np.stack([keys // nv, keys % nv], axis=1)
        self.Nv, self.Ne = nv, keys.shape[0]
        self.N2 = self.Nv + self.Ne
        self.N = 3 * self.N2 + J * self.Nv
        self._x2 = None
        return self

    def dim(self):
        return self.N

    def sub(self, i):
        return SubSpace(self, i)

    def num_sub_spaces(self):
        return 1 + self.J

    def edge_vertices(self):
        if self._edges is None:
            self._edges = self.engine.edges().cpu().numpy().astype(np.int64)
        return self._edges

    def node2_coordinates(self):
        """Coordinates of the scalar P2 nodes: vertices, then edge midpoints."""
        if self._x2 is None:
            ev = self.edge_vertices()
            x = self.mesh.coordinates
            self._x2 = np.vstack([x, 0.5 * (x[ev[:, 0]] + x[ev[:, 1]])])
        return self._x2

    def edge_index(self, lo, hi):
        ev = self.edge_vertices()
        keys = ev[:, 0] * self.Nv + ev[:, 1]          # lexicographic numbering == sorted keys
        k = np.asarray(lo, dtype=np.int64) * self.Nv + np.asarray(hi, dtype=np.int64)
        e = np.searchsorted(keys, k)
        assert np.all(keys[np.minimum(e, keys.shape[0] - 1)] == k), "edge not in mesh"
        return e

    def sub_range(self, i):
        """dof range of sub-space i (0 = displacement, i >= 1 = pressure i)."""
        if i == 0:
            return 0, 3 * self.N2
        return 3 * self.N2 + (i - 1) * self.Nv, 3 * self.N2 + i * self.Nv


class SubFunction:
    """One field of a split Function (host copy when deepcopy=True)."""

    def __init__(self, space, index, values):
        self.space, self.index, self.values = space, index, values   # u: [N2,3]; p: [Nv]

    def vector(self):
        return self.values

    def compute_vertex_values(self):
        return self.values[: self.space.Nv]

    def __call__(self, point):
        return _point_eval(self.space, self.index, self.values, np.asarray(point, dtype=float))


class FunctionRef:
    def __init__(self, fn, index):
        self.fn, self.index = fn, index


class Function:
    def __init__(self, space, tensor=None):
        self.space = space
        dev = space.engine.device
        self.x = tensor if tensor is not None else torch.zeros(space.N + space.nreal, dtype=torch.float64, device=dev)

    def function_space(self):
        return self.space

    def vector(self):
        return Vector(self.x)

    def assign(self, other):
        self.x.copy_(other.x)

    def copy(self, deepcopy=True):
        return Function(self.space, self.x.clone())

    def sub(self, i):
        return FunctionRef(self, i)

    def __len__(self):
        sp = self.space
        return 1 + sp.J + ((1 if sp.nreal >= 6 else 0) + (sp.nreal - 6 if sp.nreal >= 6 else sp.nreal) if sp.nreal else 0)

    def split(self, deepcopy=True):
        a = self.x.detach().cpu().numpy()
        sp = self.space
        out = [SubFunction(sp, 0, a[: 3 * sp.N2].reshape(3, sp.N2).T.copy())]
        for i in range(sp.J):
            lo, hi = sp.sub_range(i + 1)
            out.append(SubFunction(sp, i + 1, a[lo:hi].copy()))
        if sp.nreal:      # (u, p_1..p_J, r, p_null...) as in the reference's mixed space (mpetsolver.py:121-128)
            tail = a[sp.N:sp.N + sp.nreal]
            nz = 6 if sp.nreal >= 6 else 0
            if nz:
                out.append(SubFunction(sp, -1, tail[:nz].copy()))
            for k in range(nz, sp.nreal):
                out.append(SubFunction(sp, -1, tail[k:k + 1].copy()))
        return tuple(out)

    def set_sub(self, index, data):
        """Nodal interpolation of ``data`` (Expression / Constant / array) into sub-space index."""
        sp = self.space
        lo, hi = sp.sub_range(index)
        if isinstance(data, np.ndarray):
            vals = data
        else:
            pts = sp.node2_coordinates() if index == 0 else sp.mesh.coordinates
            vals = data.eval_points(pts)
        if index == 0:
            vals = np.asarray(vals, dtype=float).T.reshape(-1)
        self.x[lo:hi] = torch.as_tensor(np.ascontiguousarray(vals, dtype=float), device=self.x.device)


def interpolate(expr, space):
    """``interpolate(expr, VP.sub(i).collapse())`` -> nodal values (host)."""
    assert isinstance(space, SubSpace)
    sp = space.space
    pts = sp.node2_coordinates() if space.index == 0 else sp.mesh.coordinates
    return SubFunction(sp, space.index, np.asarray(expr.eval_points(pts), dtype=float))


def assign(target, source):
    """``assign(up_.sub(i), interpolate(...))`` (test_convergence_mpetsolver.py:158-163)."""
    assert isinstance(target, FunctionRef) and isinstance(source, SubFunction)
    target.fn.set_sub(target.index, source.values)


def _point_eval(space, index, values, pt):
    mesh = space.mesh
    x = mesh.coordinates[mesh.cells]
    J = np.swapaxes(x[:, 1:, :] - x[:, :1, :], 1, 2)
    X = np.linalg.solve(J, (pt - x[:, 0])[:, :, None])[:, :, 0]
    inside = np.nonzero((X.min(axis=1) > -1e-10) & (X.sum(axis=1) < 1 + 1e-10))[0]
    if inside.size == 0:
        raise ValueError("point outside the mesh")
    c = inside[0]
    lam = np.r_[1 - X[c].sum(), X[c]]
    cv = mesh.cells[c].astype(np.int64)
    if index > 0:
        return float(lam @ values[cv])
    le = [(2, 3), (1, 3), (1, 2), (0, 3), (0, 2), (0, 1)]
    phi = [lam[i] * (2 * lam[i] - 1) for i in range(4)] + [4 * lam[a] * lam[b] for a, b in le]
    nodes = list(cv) + [space.Nv + int(space.edge_index(cv[a], cv[b])) for a, b in le]
    return np.asarray(phi) @ values[nodes]


class DirichletBC:
    """``DirichletBC(VP.sub(k), value, markers, 0)`` (mpetsolver.py:74,81): topological search over
    the facets carrying the marker; ``get_boundary_values()`` evaluates the data at the dof points."""

    def __init__(self, subspace, value, markers, marker_id):
        self.subspace, self.value, self.markers, self.marker_id = subspace, value, markers, marker_id
        self._dofs = None
        self._pts = None

    def dofs(self):
        if self._dofs is None:
            sp = self.subspace.space
            F = sp.mesh.exterior_facets()
            sel = np.nonzero(self.markers.array() == self.marker_id)[0]
            fv = F["vertices"][sel]
            if self.subspace.index == 0:
                nodes = [fv.ravel()]
                for a, b in ((0, 1), (0, 2), (1, 2)):
                    if fv.shape[0]:
                        nodes.append(sp.Nv + sp.edge_index(fv[:, a], fv[:, b]))
                nodes = np.unique(np.concatenate(nodes)) if fv.shape[0] else np.zeros(0, dtype=np.int64)
                self._pts = sp.node2_coordinates()[nodes]
                self._dofs = np.concatenate([k * sp.N2 + nodes for k in range(3)]).astype(np.int64)
            else:
                verts = np.unique(fv) if fv.shape[0] else np.zeros(0, dtype=np.int64)
                self._pts = sp.mesh.coordinates[verts]
                lo, _ = sp.sub_range(self.subspace.index)
                self._dofs = (lo + verts).astype(np.int64)
        return self._dofs

    def values(self):
        d = self.dofs()
        if d.size == 0:
            return np.zeros(0)
        v = np.asarray(self.value.eval_points(self._pts), dtype=float)
        return v.T.reshape(-1) if self.subspace.index == 0 else v.reshape(-1)

    def get_boundary_values(self):
        return dict(zip(self.dofs().tolist(), self.values().tolist()))

    def apply(self, obj):
        obj._apply_dirichlet(self)
