"""CPU oracle for the MPET assemble+solve hot path.

TEST INFRASTRUCTURE ONLY.  This package is a plain numpy/scipy restatement of
what the reference (meg-simula/waterscapes, ``src/mpet/mpet/mpetsolver.py``)
asks DOLFIN/FFC/PETSc to compute on the per-timestep path.  It is imported
only by ``tests/``, by ``__graft_entry__.smoke()`` and by the ``cpu_baseline``
/ ``--impl reference`` legs of ``bench.py`` -- never by the product package
``waterscapes_b200`` (which fails loudly when its CUDA library is missing).

Pinning status
--------------
The arithmetic of the reference path lives in un-vendored third-party packages
(FEniCS DOLFIN/FFC/UFL/FIAT "2019.1 or later", ``src/mpet/INSTALL.rst:5-8``;
PETSc + MUMPS + hypre) that are not installed here and cannot be installed
(no network), so the reference itself cannot be executed to produce vectors.
What the reference's OWN tests pin for this path is solution-level only, and
the oracle is checked against every one of them (``tests/test_oracle_pins.py``):

* ``src/mpet/test/test_donut.py:16-71``  -- exact constant solution p == -t on
  the committed copy of the reference's ``donut2D.h5`` geometry
  (``|p(0,50)+0.2| < 1e-8`` and ``| ||p||_L2/sqrt(vol) - 0.2 | < 1e-10``),
  including the rigid-motion Lagrange multipliers of ``rm_basis_L2.py``.
* ``src/mpet/test/test_convergence_mpetsolver.py:103-245`` -- MMS convergence
  rates on UnitSquareMesh (thresholds 1.70 / 1.70 / 1.70 / 0.95).

CSR pattern, dof maps, matrix entries and iteration counts are pinned by NO
test or fixture of the reference: for those rows the oracle is
**parity unpinned** against DOLFIN and is anchored on the
published UFC/FIAT conventions instead (cited per function).
"""
