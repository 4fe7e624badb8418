"""Thin object wrapper over the C-ABI (one Engine == one mpet_ctx == one GPU).

All arrays are torch CUDA tensors; only raw pointers, sizes and the current stream cross the
boundary.  Nothing here computes: every method is one library call.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


class Engine:
    def __init__(self, device=0):
        if not torch.cuda.is_available():
            raise _lib.MpetLibraryError("waterscapes_b200 needs a CUDA device (B200, sm_100a); "
                                        "there is no CPU fallback")
        self.lib = _lib.load()
        self.device = torch.device("cuda", device)
        torch.cuda.set_device(self.device)
        self._ctx = C.c_void_p()
        rc = self.lib.mpet_create(int(device), C.byref(self._ctx))
        if rc != 0:
            msg = self.lib.mpet_last_error(self._ctx) if self._ctx else b"mpet_create failed"
            raise _lib.MpetLibraryError(msg.decode())
        self.sizes = None

    # ------------------------------------------------------------------ plumbing
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _ck(self, rc):
        _lib.check(self._ctx, rc)

    def close(self):
        if self._ctx:
            self.lib.mpet_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _dev(self, a, dtype):
        if isinstance(a, torch.Tensor):
            return a.to(device=self.device, dtype=dtype).contiguous()
        return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype, device=self.device)

    # ------------------------------------------------------------------ space
    def set_mesh(self, coords, cells, n_networks):
        coords = self._dev(coords, torch.float64)
        cells = self._dev(cells, torch.int32)
        assert coords.shape[1] == 3 and cells.shape[1] == 4
        self._ck(self.lib.mpet_set_mesh(self._ctx, _ptr(coords), _ptr(cells), coords.shape[0],
                                        cells.shape[0], int(n_networks), self._stream()))
        s = (C.c_int64 * 10)()
        self._ck(self.lib.mpet_get_sizes(self._ctx, s))
        names = ["Nv", "Ne", "N2", "Nc", "N", "nnz", "nnz22", "nnz21", "nnz11", "A"]
        self.sizes = dict(zip(names, [int(v) for v in s]))
        return self.sizes

    def edges(self):
        out = torch.empty((self.sizes["Ne"], 2), dtype=torch.int32, device=self.device)
        self._ck(self.lib.mpet_get_edges(self._ctx, _ptr(out), self._stream()))
        return out

    def cell_dofs(self):
        nloc = 30 + 4 * self.sizes["A"]
        out = torch.empty((self.sizes["Nc"], nloc), dtype=torch.int32, device=self.device)
        self._ck(self.lib.mpet_get_cell_dofs(self._ctx, _ptr(out), self._stream()))
        return out

    def pattern(self):
        rowptr = torch.empty(self.sizes["N"] + 1, dtype=torch.int64, device=self.device)
        cols = torch.empty(self.sizes["nnz"], dtype=torch.int32, device=self.device)
        self._ck(self.lib.mpet_get_pattern(self._ctx, _ptr(rowptr), _ptr(cols), self._stream()))
        return rowptr, cols

    # ------------------------------------------------------------------ coefficients / assembly
    def set_params(self, E, nu, alpha, K, S, c, dt, theta):
        A = self.sizes["A"]
        arr = lambda v, n: (C.c_double * max(n, 1))(*[float(x) for x in np.asarray(v, dtype=float).ravel()])
        self._ck(self.lib.mpet_set_params(self._ctx, float(E), float(nu), arr(alpha, A), arr(K, A),
                                          arr(S, A * A), arr(c, A), float(dt), float(theta)))

    def set_params_total_pressure(self, E, nu, alpha, K, S, c, dt, theta):
        """Total-pressure formulation: the mesh was set with J + 1 P1 fields (field 0 = total pressure)."""
        J = self.sizes["A"] - 1
        arr = lambda v, n: (C.c_double * max(n, 1))(*[float(x) for x in np.asarray(v, dtype=float).ravel()])
        self._ck(self.lib.mpet_set_params_total_pressure(self._ctx, float(E), float(nu), arr(alpha, J), arr(K, J),
                                                         arr(S, J * J), arr(c, J), float(dt), float(theta)))

    def set_cell_coefficient(self, field, values):
        """DG0 permeability of P1 field ``field``: one value per cell, or None for the constant."""
        if values is None:
            self._ck(self.lib.mpet_set_cell_coefficient(self._ctx, int(field), C.c_void_p(0), self._stream()))
            return
        v = self._dev(values, torch.float64)
        assert v.numel() == self.sizes["Nc"]
        self._ck(self.lib.mpet_set_cell_coefficient(self._ctx, int(field), _ptr(v), self._stream()))

    def set_dof_permutation(self, ext_of_contract):
        """Caller's dof numbering: entry c = the caller's index of contract (UFC) dof c; None = contract numbering.
        Afterwards dof vectors and dof indices crossing the ABI are in the caller's numbering."""
        if ext_of_contract is None:
            self._ck(self.lib.mpet_set_dof_permutation(self._ctx, C.c_void_p(0), self._stream()))
            return
        perm = self._dev(ext_of_contract, torch.int32)
        assert perm.numel() == self.sizes["N"]
        self._ck(self.lib.mpet_set_dof_permutation(self._ctx, _ptr(perm), self._stream()))

    def assemble_lhs(self):
        self._ck(self.lib.mpet_assemble_lhs(self._ctx, self._stream()))

    def add_entries(self, rows, cols, vals):
        rows = self._dev(rows, torch.int32)
        cols = self._dev(cols, torch.int32)
        vals = self._dev(vals, torch.float64)
        self._ck(self.lib.mpet_add_entries(self._ctx, _ptr(rows), _ptr(cols), _ptr(vals), rows.numel(),
                                           self._stream()))

    def assemble_prec(self):
        self._ck(self.lib.mpet_assemble_prec(self._ctx, self._stream()))

    def values(self, which=0):
        out = torch.empty(self.sizes["nnz"], dtype=torch.float64, device=self.device)
        self._ck(self.lib.mpet_get_values(self._ctx, int(which), _ptr(out), self._stream()))
        return out

    # ------------------------------------------------------------------ Dirichlet / rhs
    def set_dirichlet_dofs(self, dofs):
        dofs = self._dev(dofs, torch.int32)
        self._ck(self.lib.mpet_set_dirichlet_dofs(self._ctx, _ptr(dofs), dofs.numel(), self._stream()))

    def set_dirichlet_values(self, vals):
        vals = self._dev(vals, torch.float64)
        self._ck(self.lib.mpet_set_dirichlet_values(self._ctx, _ptr(vals), self._stream()))

    def rhs_prev(self, up_prev, b):
        self._ck(self.lib.mpet_rhs_prev(self._ctx, _ptr(up_prev), _ptr(b), self._stream()))

    def mass_apply(self, space, scale, x, y):
        self._ck(self.lib.mpet_mass_apply(self._ctx, int(space), float(scale), _ptr(x), _ptr(y), self._stream()))

    def lumped(self, space):
        n = self.sizes["N2"] if space == 2 else self.sizes["Nv"]
        w = torch.empty(n, dtype=torch.float64, device=self.device)
        self._ck(self.lib.mpet_lumped(self._ctx, int(space), _ptr(w), self._stream()))
        return w

    def apply_dirichlet_rhs(self, b):
        self._ck(self.lib.mpet_apply_dirichlet_rhs(self._ctx, _ptr(b), self._stream()))

    # ------------------------------------------------------------------ sparse kernels
    def spmv(self, x, y):
        self._ck(self.lib.mpet_spmv(self._ctx, _ptr(x), _ptr(y), self._stream()))

    def csr_spmv(self, rowptr, cols, vals, x, y, beta=0.0):
        self._ck(self.lib.mpet_csr_spmv(self._ctx, rowptr.numel() - 1, _ptr(rowptr), _ptr(cols), _ptr(vals),
                                        _ptr(x), _ptr(y), float(beta), self._stream()))

    # ------------------------------------------------------------------ Krylov
    def krylov_setup(self, method="minres", pc="amg", rtol=1e-5, atol=1e-50, maxit=10000, restart=30,
                     reference_norm="b"):
        m = {"minres": 0, "gmres": 1}[method]
        p = {"none": 0, "jacobi": 1, "amg": 2}[pc]
        self._ck(self.lib.mpet_krylov_setup(self._ctx, m, p, float(rtol), float(atol), int(maxit), int(restart)))
        self._ck(self.lib.mpet_krylov_reference_norm(self._ctx, {"b": 0, "min_b_r0": 1}[reference_norm]))

    def set_border(self, columns):
        """columns: [nb, N] device tensor (or None / empty to remove the border)."""
        if columns is None or columns.numel() == 0:
            self._ck(self.lib.mpet_set_border(self._ctx, 0, C.c_void_p(0), self._stream()))
            self.nb = 0
            return
        cols = self._dev(columns, torch.float64)
        assert cols.shape[1] == self.sizes["N"]
        self._ck(self.lib.mpet_set_border(self._ctx, int(cols.shape[0]), _ptr(cols), self._stream()))
        self.nb = int(cols.shape[0])

    def set_prec_shift(self, shift_u, shift_p):
        A = self.sizes["A"]
        arr = (C.c_double * max(A, 1))(*[float(v) for v in shift_p])
        self._ck(self.lib.mpet_set_prec_shift(self._ctx, float(shift_u), arr))

    def pc_setup(self):
        self._ck(self.lib.mpet_pc_setup(self._ctx, self._stream()))

    def solve(self, b, x):
        info = (C.c_double * 8)()
        self._ck(self.lib.mpet_solve(self._ctx, _ptr(b), _ptr(x), info, self._stream()))
        return dict(niter=int(info[0]), converged=bool(info[1]), rel_res=float(info[2]), res0=float(info[3]),
                    breakdown=bool(info[4]), reason=int(info[5]), refnorm=float(info[6]), bnorm=float(info[7]))

    def pc_apply(self, r, z):
        self._ck(self.lib.mpet_pc_apply(self._ctx, _ptr(r), _ptr(z), self._stream()))

    # ------------------------------------------------------------------ multi-GPU
    def nccl_unique_id(self):
        buf = (C.c_char * 128)()
        if self.lib.mpet_nccl_unique_id(C.cast(buf, C.c_void_p)) != 0:
            raise _lib.MpetLibraryError("ncclGetUniqueId failed (is libnccl.so.2 loadable?)")
        return bytes(buf.raw)

    def attach_comm(self, uid_bytes, rank, nranks):
        buf = C.create_string_buffer(uid_bytes, 128)
        self._ck(self.lib.mpet_attach_comm(self._ctx, C.cast(buf, C.c_void_p), int(rank), int(nranks)))

    def set_halo(self, neighbours, send_lists, recv_lists, owned):
        n = len(neighbours)
        ranks = (C.c_int * max(n, 1))(*[int(q) for q in neighbours])
        soff = np.concatenate([[0], np.cumsum([len(a) for a in send_lists])]).astype(np.int64)
        roff = np.concatenate([[0], np.cumsum([len(a) for a in recv_lists])]).astype(np.int64)
        cat = lambda ls: np.concatenate(ls).astype(np.int32) if ls and sum(len(a) for a in ls) else np.zeros(1, np.int32)
        sd = self._dev(cat(send_lists), torch.int32)
        rd = self._dev(cat(recv_lists), torch.int32)
        od = self._dev(np.asarray(owned, dtype=np.uint8), torch.uint8)
        self._ck(self.lib.mpet_set_halo(self._ctx, n, ranks, (C.c_int64 * (n + 1))(*soff.tolist()), _ptr(sd),
                                        (C.c_int64 * (n + 1))(*roff.tolist()), _ptr(rd), _ptr(od), self._stream()))

    # ------------------------------------------------------------------ instrumentation
    def launch_count(self, reset=False):
        return int(self.lib.mpet_launch_count(self._ctx, int(reset)))

    def profile(self, enable=-1):
        """Read (and optionally reset/switch) the library's CUDA-event profile."""
        out = (C.c_double * 16)()
        self._ck(self.lib.mpet_profile(self._ctx, int(enable), out))
        names = ["spmv", "pc", "vec", "assemble", "rhs", "comm", "pc_u", "pc_p"]
        return {n: dict(ms=float(out[i]), count=int(out[8 + i])) for i, n in enumerate(names)}

    def pc_bytes(self):
        return int(self.lib.mpet_pc_bytes(self._ctx))

    def comm_kind(self):
        return {0: "single GPU", 1: "NCCL send/recv", 2: "peer-memory (NVLink stores + flags)"}[
            int(self.lib.mpet_comm_kind(self._ctx))]

    def device_bytes(self):
        return int(self.lib.mpet_device_bytes(self._ctx))
