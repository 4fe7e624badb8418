"""MPETProblem: the reference's problem container, unchanged surface
(src/mpet/mpet/mpetproblem.py:8-24 Lame conversions, :103-180 class)."""
import numpy as np

from .dolfin_shim import Constant, MeshFunction, INVALID


def convert_to_E_nu(mu, lmbda):
    """mpetproblem.py:8-11."""
    E = mu * (3 * lmbda + 2 * mu) / (lmbda + mu)
    nu = lmbda / (2 * (lmbda + mu))
    return (E, nu)


def convert_to_mu_lmbda(E, nu):
    """mpetproblem.py:13-16."""
    mu = E / (2.0 * ((1.0 + nu)))
    lmbda = nu * E / ((1.0 - 2.0 * nu) * (1.0 + nu))
    return (mu, lmbda)


class CellTensorField(object):
    """Per-cell evaluation of a tensor built from the gradient of a split P2 displacement (host side,
    post-processing only): ``cell_values()`` -> [Nc, 3, 3] at the cell barycentres."""

    def __init__(self, u, two_mu, lmbda):
        self.u, self.two_mu, self.lmbda = u, two_mu, lmbda

    def cell_values(self):
        sp = self.u.space
        mesh = sp.mesh
        cells = mesh.cells.astype(np.int64)
        x = mesh.coordinates[cells]                                     # [Nc, 4, 3]
        J = np.swapaxes(x[:, 1:, :] - x[:, :1, :], 1, 2)                # d x_i / d X_j
        Jinv = np.linalg.inv(J)
        lam = np.full(4, 0.25)                                          # barycentre
        gl = np.array([[-1.0, -1.0, -1.0], [1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0]])
        le = [(2, 3), (1, 3), (1, 2), (0, 3), (0, 2), (0, 1)]           # UFC local edges
        dphi = [(4 * lam[i] - 1) * gl[i] for i in range(4)] + [4 * (lam[a] * gl[b] + lam[b] * gl[a]) for a, b in le]
        dphi = np.asarray(dphi)                                         # [10, 3] reference gradients
        nodes = np.concatenate([cells] + [sp.Nv + sp.edge_index(cells[:, a], cells[:, b])[:, None]
                                          for a, b in le], axis=1)       # [Nc, 10]
        g = np.einsum("ak,ckm->cam", dphi, Jinv)                        # physical gradients
        grad_u = np.einsum("cai,cam->cim", self.u.values[nodes], g)     # d u_i / d x_m
        eps = 0.5 * (grad_u + np.swapaxes(grad_u, 1, 2))
        div = np.trace(grad_u, axis1=1, axis2=2)
        return self.two_mu * eps + self.lmbda * div[:, None, None] * np.eye(3)[None]


def elastic_stress(u, E, nu):
    """mpetproblem.py:18-24: sigma(u) = 2 mu eps(u) + lambda div(u) I for a split displacement ``u``
    (``up.split()[0]``); evaluated per cell on the host (post-processing, not on the timestep path)."""
    (mu, lmbda) = convert_to_mu_lmbda(float(E), float(nu))
    return CellTensorField(u, 2.0 * mu, lmbda)


class MPETProblem(object):
    """Data of one MPET problem: mesh, shared time Constant, material parameters
    (required keys J, E, nu, alpha, K, S, c), forcing / boundary data and facet markers
    (0 = Dirichlet, 1 = Neumann, 2 = Robin; sys.maxsize = unmarked)."""

    def __init__(self, mesh, time, params=None):
        self.mesh = mesh
        self.time = time
        self.params = self.default_parameters()
        if params is not None:
            self.params.update(params)
        Js = range(int(self.params["J"]))
        gdim = self.mesh.geometry().dim()

        self.f = Constant((0.0,) * gdim)
        self.s = Constant((0.0,) * gdim)
        self.g = [Constant(0.0) for i in Js]
        self.I = [Constant(0.0) for i in Js]
        self.beta = [Constant(0.0) for i in Js]
        self.p_robin = [Constant(0.0) for i in Js]
        self.u_bar = Constant((0.0,) * gdim)
        self.p_bar = [Constant(0.0) for i in Js]

        tdim = mesh.topology().dim()
        markers = MeshFunction("size_t", mesh, tdim - 1)
        markers.set_all(INVALID)
        self.momentum_boundary_markers = markers
        self.continuity_boundary_markers = []
        for i in Js:
            markers = MeshFunction("size_t", mesh, tdim - 1)
            markers.set_all(INVALID)
            self.continuity_boundary_markers += [markers]

        self.u_has_nullspace = False
        self.p_has_nullspace = list(False for i in Js)

    @classmethod
    def default_parameters(cls):
        "Same defaults as the reference (including its 'X' key; callers always pass S)."
        return {"J": 1.0, "rho": Constant(1.0), "nu": Constant(0.479), "E": Constant(1500),
                "alpha": (Constant(1.0),), "K": (Constant(1.0),), "X": ((Constant(1.0),),),
                "c": (Constant(1.0),)}
