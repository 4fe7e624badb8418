"""MPETProblem: the reference's problem container, unchanged surface
(src/mpet/mpet/mpetproblem.py:8-24 Lame conversions, :103-180 class)."""
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
