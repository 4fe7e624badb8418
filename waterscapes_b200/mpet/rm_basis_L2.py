"""Rigid motions of a 3-D body, orthonormal in L2 -- host-side twin of the reference's
``rigid_motions(mesh)`` (src/mpet/mpet/rm_basis_L2.py:10-74): translations along the principal axes of the
rotational Gram matrix scaled by 1/sqrt(volume), rotations ``cross(x - c, v_k) / sqrt(w_k)`` about the centre of
mass c, (w_k, v_k) the eigenpairs of R_ij = int cross(x - c, e_i) . cross(x - c, e_j) dx.  The integrals (degree
<= 2 on affine tets) are exact with the 4-point rule.  Returns six callables Z_i(x[npts, 3]) -> [npts, 3]; the
Lagrange multipliers that use them enter the solver as a dense border (include/mpet_b200.h: mpet_set_border)."""
import numpy as np

_A, _B = 0.1381966011250105, 0.5854101966249685      # degree-2 rule on the reference tet
_QP = np.array([[_A, _A, _A], [_B, _A, _A], [_A, _B, _A], [_A, _A, _B]])


def rigid_motions(mesh):
    x = mesh.coordinates[mesh.cells.astype(np.int64)]                    # [c, 4, 3]
    J = x[:, 1:, :] - x[:, :1, :]
    vol = np.abs(np.linalg.det(J)) / 6.0
    volume = vol.sum()
    cen = (x.mean(axis=1) * vol[:, None]).sum(axis=0) / volume           # exact: x is affine per cell
    lam = np.concatenate([1.0 - _QP.sum(axis=1, keepdims=True), _QP], axis=1)
    xq = np.einsum("qv,cvd->cqd", lam, x) - cen
    wq = np.repeat(vol[:, None] / 4.0, 4, axis=1)
    eye = np.eye(3)
    R = np.zeros((3, 3))
    for i in range(3):
        ci = np.cross(xq, eye[i])
        for j in range(i, 3):
            R[i, j] = R[j, i] = np.einsum("cq,cqd,cqd->", wq, ci, np.cross(xq, eye[j]))
    eigw, eigv = np.linalg.eigh(R)
    eigv = eigv.T
    s = np.sqrt(volume)
    translations = [lambda p, v=v: np.tile(v / s, (p.shape[0], 1)) for v in eigv]
    rotations = [lambda p, v=v, w=w: np.cross(p - cen, v) / np.sqrt(w) for v, w in zip(eigv, eigw)]
    return translations + rotations
