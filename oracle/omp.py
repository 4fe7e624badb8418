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
        _lib = L
    return _lib


def num_threads():
    return int(lib().oracle_num_threads())


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
