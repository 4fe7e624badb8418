"""All-host-core variants of the oracle's sparse operators (TEST INFRASTRUCTURE, see oracle/__init__.py).

``OmpCsr`` wraps a scipy CSR matrix so that ``A @ x`` runs the OpenMP row loop of oracle/csrc/cpu_kernels.c
(the restatement of PETSc's MatMult on the reference path) instead of scipy's single-threaded csr_matvec.
``parallelise(M)`` swaps the matrices of a BlockAMG hierarchy in place.  Results are bit-identical to the
scipy path (same per-row summation order); only the timing changes."""
import ctypes as C
import os

import numpy as np

from . import build as _build

_lib = None


def lib():
    global _lib
    if _lib is None:
        path = _build.LIB if os.path.exists(_build.LIB) else _build.build()
        L = C.CDLL(path)
        L.oracle_csr_spmm.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.oracle_csr_spmm.restype = None
        L.oracle_num_threads.restype = C.c_int
        L.oracle_assemble_lhs_tet.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                              C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                              C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_assemble_lhs_tet.restype = None
        _lib = L
    return _lib


def num_threads():
    return int(lib().oracle_num_threads())


def use_all_cores():
    """Run the OpenMP loops on every core this process may use (torchrun sets OMP_NUM_THREADS=1 for its ranks;
    the CPU timing arm runs on rank 0 alone).  Returns the thread count."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    lib().oracle_set_num_threads(C.c_int(n))
    return num_threads()


class OmpCsr:
    def __init__(self, A, with_transpose=False):
        A = A.tocsr()
        A.sort_indices()
        assert A.indices.dtype == np.int32 and A.indptr.dtype == np.int32
        self.shape = A.shape
        self._rp = np.ascontiguousarray(A.indptr)
        self._ci = np.ascontiguousarray(A.indices)
        self._v = np.ascontiguousarray(A.data, dtype=np.float64)
        self._T = OmpCsr(A.T.tocsr()) if with_transpose else None
        self._scipy = A

    @property
    def T(self):
        if self._T is None:
            self._T = OmpCsr(self._scipy.T.tocsr())
        return self._T

    def diagonal(self):
        return self._scipy.diagonal()

    def __matmul__(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        w = 1 if x.ndim == 1 else x.shape[1]
        assert x.shape[0] == self.shape[1] and w <= 8
        y = np.empty((self.shape[0],) if x.ndim == 1 else (self.shape[0], w))
        lib().oracle_csr_spmm(self.shape[0], self._rp.ctypes.data, self._ci.ctypes.data, self._v.ctypes.data,
                              x.ctypes.data, y.ctypes.data, w)
        return y


def parallelise(block_amg):
    """Swap every level matrix of a oracle.krylov.BlockAMG for its OpenMP twin (in place)."""
    for h in [block_amg.u] + list(block_amg.p):
        for L in h.levels:
            L.A = OmpCsr(L.A)
            if L.P is not None:
                L.P = OmpCsr(L.P, with_transpose=True)
    return block_amg


class LhsAssembler:
    """``assemble(a)`` of the standard formulation as an OpenMP cell loop (oracle/csrc/cpu_kernels.c), on the
    oracle's own pattern.  Same numbers as ``MPETOracle.assemble_lhs`` without Robin / nullspace terms (checked in
    tests/test_oracle_krylov.py); exists so that the CPU timing arm is not dominated by numpy's assembly."""

    def __init__(self, o):
        import scipy.sparse as sp
        from .fem import simplex_quadrature, tabulate
        from .mpet import convert_to_mu_lmbda
        assert o.d == 3 and o.P_SHIFT == 0 and o.space.nreal == 0 and o.space.perm is None
        self.o = o
        self.rowptr, self.cols = o.pattern()
        self.rowptr = np.ascontiguousarray(self.rowptr, dtype=np.int64)
        self.cols = np.ascontiguousarray(self.cols, dtype=np.int32)
        pts, wts = simplex_quadrature(3, 2)
        _, dN2 = tabulate(3, 2, pts)
        N1, dN1 = tabulate(3, 1, pts)
        self.nq = len(wts)
        self.dN2 = np.ascontiguousarray(dN2, dtype=np.float64)          # [nq, 10, 3]
        self.N1 = np.ascontiguousarray(N1, dtype=np.float64)            # [nq, 4]
        self.dN1 = np.ascontiguousarray(dN1[0], dtype=np.float64)       # [4, 3] (constant)
        self.wq = np.ascontiguousarray(wts, dtype=np.float64)
        self.cells = np.ascontiguousarray(o.mesh.cells, dtype=np.int64)
        self.cell_dofs = np.ascontiguousarray(o.space.cell_dofs, dtype=np.int64)
        self.coords = np.ascontiguousarray(o.mesh.coords, dtype=np.float64)
        self._sp = sp
        self._lame = convert_to_mu_lmbda

    def __call__(self):
        o = self.o
        mu, lmbda = self._lame(o.E, o.nu)
        coef = np.ascontiguousarray(np.concatenate([[mu, lmbda, o.dt * o.theta], o.alpha, o.c, o.K,
                                                    np.asarray(o.S, dtype=float).ravel()]), dtype=np.float64)
        vals = np.zeros(self.cols.shape[0])
        p = lambda a: a.ctypes.data
        lib().oracle_assemble_lhs_tet(self.cells.shape[0], p(self.cells), p(self.cell_dofs), o.A, p(self.coords),
                                      self.nq, p(self.dN2), p(self.N1), p(self.dN1), p(self.wq), p(coef),
                                      p(self.rowptr), p(self.cols), p(vals))
        N = o.space.N
        return self._sp.csr_matrix((vals, self.cols, self.rowptr.astype(np.int32)), shape=(N, N))
