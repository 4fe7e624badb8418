/*
 * mpet_b200.h -- C-ABI of libmpet_b200.so: the B200-native (sm_100a) replacement for the
 * per-timestep assemble + Krylov-solve path of waterscapes' `mpet` solver.
 *
 * The reference has no FFI of its own; its boundary for this path is the DOLFIN/PETSc Python API
 * used by src/mpet/mpet/mpetsolver.py.  Each entry point below names the reference call it replaces
 * (paths relative to the reference checkout, `mpetsolver.py` = src/mpet/mpet/mpetsolver.py).
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error; mpet_last_error(ctx) gives the message.
 *     No C++ exception crosses this boundary.
 *   - "dev" pointers are DEVICE pointers owned by the caller (torch tensors: tensor.data_ptr());
 *     "host" pointers are ordinary host memory.  The library never frees caller memory.
 *   - `stream` is a cudaStream_t passed as void* (torch.cuda.current_stream().cuda_stream).
 *     Calls are asynchronous w.r.t. the host unless stated otherwise.
 *   - one ctx per GPU and per host thread; not thread-safe.
 *   - all floating-point data is fp64; dof / column indices are int32, CSR row pointers int64.
 *
 * Dof numbering (UFC numbering of MixedElement([P2^3] + [P1]*A), mpetsolver.py:100-132):
 *   displacement component k, scalar P2 node n  ->  k*N2 + n      (n < Nv: vertex, else Nv + edge)
 *   pressure i, vertex v                        ->  3*N2 + i*Nv + v
 *   edges are numbered lexicographically by (lo vertex, hi vertex).
 */
#ifndef MPET_B200_H
#define MPET_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mpet_ctx mpet_ctx;

/* ---- lifetime ------------------------------------------------------------------------------- */
int mpet_create(int device, mpet_ctx** out);
void mpet_destroy(mpet_ctx* ctx);
const char* mpet_last_error(mpet_ctx* ctx);
int mpet_abi_version(void);

/* ---- function space + sparsity graph (built once, on device) ----------------------------------
 * Replaces FunctionSpace(mesh, MixedElement(...)) (mpetsolver.py:126-130: dof map) and DOLFIN's
 * SparsityPatternBuilder + PETScMatrix::init inside assemble() (mpetsolver.py:335,412,496).
 * coords_dev f64[nv*3], cells_dev i32[nc*4] (each cell's vertices ascending).  Synchronous. */
int mpet_set_mesh(mpet_ctx* ctx, const double* coords_dev, const int32_t* cells_dev,
                  int64_t nv, int64_t nc, int n_networks, void* stream);

/* sizes[0..9] = Nv, Ne, N2, Nc, N, nnz, nnz22, nnz21, nnz11, n_networks */
int mpet_get_sizes(mpet_ctx* ctx, int64_t* sizes_host);
/* edge_vertices i32[Ne*2]; cell_dofs i32[Nc*(30+4A)] in UFC local order (device outputs) */
int mpet_get_edges(mpet_ctx* ctx, int32_t* edge_vertices_dev, void* stream);
int mpet_get_cell_dofs(mpet_ctx* ctx, int32_t* cell_dofs_dev, void* stream);
/* CSR pattern of the block system: rowptr i64[N+1], cols i32[nnz] (ascending per row) */
int mpet_get_pattern(mpet_ctx* ctx, int64_t* rowptr_dev, int32_t* cols_dev, void* stream);

/* ---- coefficients ------------------------------------------------------------------------------
 * Replaces the Constants captured by create_variational_forms (mpetsolver.py:183-189) and
 * solver.dt.assign / params["dt"], params["theta"] (mpetsolver.py:137,179,329-330).
 * alpha, K, c: host f64[A]; S: host f64[A*A] row-major. */
int mpet_set_params(mpet_ctx* ctx, double E, double nu, const double* alpha_host,
                    const double* K_host, const double* S_host, const double* c_host,
                    double dt, double theta);
/* Total-pressure formulation (mpettotalpressuresolver.py:267-283: unknowns u, p0 = total pressure,
 * p_1..p_J): the context must have been created with n_networks = J + 1 P1 fields (field 0 is p0);
 * alpha, K, c: host f64[J]; S: host f64[J*J].  Every other entry point (assembly, right-hand side,
 * preconditioner, Krylov solve) is shared with the standard formulation. */
int mpet_set_params_total_pressure(mpet_ctx* ctx, double E, double nu, const double* alpha_host,
                                   const double* K_host, const double* S_host, const double* c_host,
                                   double dt, double theta);

/* Spatially varying permeability: K of P1 field `field` as a DG0 function, one value per cell,
 * as in the reference's sandbox/biot-robin/three_fields_precond.py:263-271 (`K = Function(DG0)`).  The scalar K
 * passed to mpet_set_params for that field then acts as a multiplier (pass 1.0).  values_dev f64[Nc] (copied) or
 * NULL to return to the constant.  Assemble (lhs and prec) again afterwards. */
int mpet_set_cell_coefficient(mpet_ctx* ctx, int field, const double* values_dev, void* stream);

/* ---- external dof numbering (`mpet_dofmap(ctx, perm*)`) ---------------------------------------
 * DOLFIN re-orders the dofs of its FunctionSpace (parameters["reorder_dofs_serial"], SCOTCH/Boost: not reproducible
 * outside DOLFIN), so `up.vector()` of mpetsolver.py:131-132 is not in the UFC numbering this library's contract
 * uses.  ext_of_contract_dev i32[N] (device, copied): entry c = the CALLER's index of contract dof c -- for DOLFIN
 * built once from `V.sub(k).dofmap().dofs()` / `vertex_to_dof_map(V.sub(3+i).collapse())` and the edge table of
 * mpet_get_edges.  Must be a bijection of 0..N-1 (checked).  NULL returns to the contract numbering.
 * Afterwards every dof VECTOR crossing the boundary is in the caller's numbering -- b and x of mpet_solve (the nb
 * multipliers of a bordered system stay behind the N dofs), up_prev and b of mpet_rhs_prev, x and y of mpet_spmv,
 * b of mpet_apply_dirichlet_rhs, r and z of mpet_pc_apply, the columns of mpet_set_border -- and so is every dof
 * INDEX: mpet_set_dirichlet_dofs, rows / cols of mpet_add_entries.  The exports that describe the matrix itself
 * (mpet_get_cell_dofs, mpet_get_pattern, mpet_get_values) stay in the contract numbering; the node vectors of
 * mpet_mass_apply / mpet_lumped are per scalar space and unaffected.  Cost: one gather and one scatter of N doubles
 * per vector argument (0.05 ms at 10 M dofs). */
int mpet_set_dof_permutation(mpet_ctx* ctx, const int32_t* ext_of_contract_dev, void* stream);

/* ---- matrix assembly ---------------------------------------------------------------------------
 * mpet_assemble_lhs  : A = assemble(a)                       (mpetsolver.py:335,412,496; form :196-201,260)
 * mpet_add_entries   : A.axpy(1.0, assemble(a_robin[i]))     (mpetsolver.py:336-338; form :252-253)
 *                      rows/cols i32[n], vals f64[n] on device, (row,col) pairs unique
 * mpet_assemble_prec : P = assemble(prec)                    (mpetsolver.py:264-277,502; well-formed
 *                      analogue mpettotalpressuresolver.py:276-283): scalar P2 block mu*(grad,grad)
 *                      shared by the 3 displacement components + one P1 block per network.
 * mpet_get_values    : copy values to caller memory (device), laid out on mpet_get_pattern's CSR.
 *                      which = 0: A as assembled (no BCs)
 *                              1: A after bc.apply(A)            (rows zeroed, unit diagonal)
 *                              2: A after apply_symmetric(bc, A) (bc_symmetric.py:11-22)
 *                              3: block-diagonal P expanded on the full pattern, symmetric BCs applied
 */
int mpet_assemble_lhs(mpet_ctx* ctx, void* stream);
int mpet_add_entries(mpet_ctx* ctx, const int32_t* rows_dev, const int32_t* cols_dev,
                     const double* vals_dev, int64_t n, void* stream);
int mpet_assemble_prec(mpet_ctx* ctx, void* stream);
int mpet_get_values(mpet_ctx* ctx, int which, double* vals_dev, void* stream);

/* ---- Dirichlet conditions ----------------------------------------------------------------------
 * Replaces create_dirichlet_bcs (mpetsolver.py:63-84) / DirichletBC.get_boundary_values:
 * the dof set is fixed; values change every step.  dofs i32[n] (device), unique. */
int mpet_set_dirichlet_dofs(mpet_ctx* ctx, const int32_t* dofs_dev, int64_t n, void* stream);
int mpet_set_dirichlet_values(mpet_ctx* ctx, const double* vals_dev, void* stream);

/* ---- right-hand side ---------------------------------------------------------------------------
 * mpet_rhs_prev : b  = assemble(L)   previous-state load   (mpetsolver.py:356,433,528; L = rhs(F) :261)
 * mpet_mass_apply: y += scale * M x  with the consistent mass of the chosen scalar space
 *                  space = 2: P2 (x, y are N2 long)   -> int f_k phi_a   (L0, mpetsolver.py:233)
 *                  space = 1: P1 (x, y are Nv long)   -> int g_i psi_m   (L1, mpetsolver.py:249)
 * mpet_lumped   : w[n] = int phi_n dx for the scalar space (constant-coefficient loads)
 * mpet_apply_dirichlet_rhs : bc.apply(b): b[dof] = value   (mpetsolver.py:375-376,452-453)
 */
int mpet_rhs_prev(mpet_ctx* ctx, const double* up_prev_dev, double* b_dev, void* stream);
int mpet_mass_apply(mpet_ctx* ctx, int space, double scale, const double* x_dev, double* y_dev,
                    void* stream);
int mpet_lumped(mpet_ctx* ctx, int space, double* w_dev, void* stream);
int mpet_apply_dirichlet_rhs(mpet_ctx* ctx, double* b_dev, void* stream);

/* ---- sparse kernels exposed for measurement and tests ------------------------------------------
 * y = A x on the assembled block matrix (no BC masking); what PETSc MatMult does inside KSP. */
int mpet_spmv(mpet_ctx* ctx, const double* x_dev, double* y_dev, void* stream);
/* generic CSR SpMV on caller-owned arrays (tests / facet operators) */
int mpet_csr_spmv(mpet_ctx* ctx, int64_t nrows, const int64_t* rowptr_dev, const int32_t* cols_dev,
                  const double* vals_dev, const double* x_dev, double* y_dev, double beta,
                  void* stream);

/* ---- Krylov solve ------------------------------------------------------------------------------
 * Replaces PETScKrylovSolver("minres","hypre_amg") + set_operators(A, P) + solve(x, b)
 * (mpetsolver.py:507,553,556; mpettotalpressuresolver.py:448,492-493) and, for parity with the
 * default path, LUSolver(A,"mumps").solve (mpetsolver.py:347,379,422,456) when run to tight rtol.
 * method: 0 = MINRES (symmetric S), 1 = restarted GMRES (any S).
 * pc    : 0 = none, 1 = Jacobi, 2 = block-diagonal AMG V-cycle (needs mpet_assemble_prec first).
 * mpet_pc_setup builds the AMG hierarchies from the current P (once per dt).
 * mpet_solve: x_dev holds the initial guess on entry (Dirichlet entries are overwritten with the
 * boundary values), the solution on exit.  Convergence: ||r||_B <= max(rtol*||b||_B, atol) with b the
 * right-hand side after apply_symmetric (bc_symmetric.py:11-22) and B the preconditioner -- PETSc's default
 * test for a nonzero initial guess (KSPConvergedDefault: reference norm = preconditioned norm of b, not of
 * the initial residual); MINRES also stops when the residual norm fell by < 1 % over 256 iterations
 * (stagnation at round-off).  Synchronous.  info_host f64[8]: [0] iterations, [1] converged flag,
 * [2] final ||r||_B / ||b||_B, [3] initial residual norm, [4] breakdown flag (indefinite preconditioner),
 * [5] reason (2 tolerance reached, -3 iteration limit, -4 breakdown, -5 stagnation), [6] the reference norm of the
 * test (mpet_krylov_reference_norm), [7] ||b||_B. */
int mpet_krylov_setup(mpet_ctx* ctx, int method, int pc, double rtol, double atol, int maxit,
                      int restart);
/* reference norm of the convergence test: 0 = ||b||_B (PETSc default, KSPConvergedDefault), 1 = min(||b||_B,
 * ||r0||_B) (KSPConvergedDefaultSetUMIRNorm) -- used by the LUSolver stand-in, whose answer must not depend on how
 * large the boundary data is compared with the update of one time step */
int mpet_krylov_reference_norm(mpet_ctx* ctx, int mode);
int mpet_pc_setup(mpet_ctx* ctx, void* stream);
int mpet_solve(mpet_ctx* ctx, const double* b_dev, double* x_dev, double* info_host, void* stream);
/* apply the preconditioner once: z = M^-1 r (tests) */
int mpet_pc_apply(mpet_ctx* ctx, const double* r_dev, double* z_dev, void* stream);

/* ---- nullspace handling: Lagrange multipliers as a dense border -------------------------
 * Replaces the Real-space components the reference appends to its mixed space when the displacement is only
 * determined up to rigid motions (u_has_nullspace) or a pressure up to a constant (p_has_nullspace):
 * mpetsolver.py:115-124 (spaces), :203-215 (r_i (Z_i, u) + z_i (Z_i, v), p_i q_null + p_null q_i), with the rigid
 * motions Z_i of rm_basis_L2.py:10-74.  The nb <= 16 dense rows/columns stay OUTSIDE the CSR matrix: the solver
 * iterates on K = [A C; C^T 0] with column i of C = columns_dev[i*N .. (i+1)*N) (e.g. the mass matrix applied to the
 * nodal values of Z_i).  With a border, every vector passed to mpet_solve has N + nb entries, multipliers last.
 * mpet_set_prec_shift adds shift_u (u, v) and shift_p[i] (p_i, q_i) to the blocks of the preconditioner (the
 * reference's commented-out `inner(u, v)*dx` term, mpetsolver.py:276-277, and `p[k]*q[k]*dx`, :274): without
 * essential boundary conditions those blocks are singular.  Call it before mpet_assemble_prec. */
int mpet_set_border(mpet_ctx* ctx, int nb, const double* columns_dev, void* stream);
int mpet_set_prec_shift(mpet_ctx* ctx, double shift_u, const double* shift_p_host);

/* ---- multi-GPU (one process per GPU; rows owned by rank, ghost entries exchanged over NVLink) ---
 * Replaces what DOLFIN + PETSc do under mpirun (src/mpet/utils/jobscript.sh:43): VecScatter halo
 * updates inside MatMult and MPI_Allreduce in the Krylov dot products.  Every rank calls mpet_set_mesh
 * with its OWN cells plus a one-cell ghost layer (no assembly collective is needed), then
 *   mpet_nccl_unique_id : rank 0 creates the 128-byte ncclUniqueId; the host plumbing broadcasts it
 *   mpet_attach_comm    : ncclCommInitRank on this context's device
 *   mpet_set_halo       : neighbour ranks (host, ascending) and, per neighbour q, the LOCAL scalar P2
 *                         node ids (vertex v -> v, edge e -> Nv + e) whose dofs this rank sends to q /
 *                         receives from q, as two concatenated device arrays with host offsets
 *                         [n_neighbours+1]; both sides list a shared node in the same order.
 *                         owned_nodes_dev u8[N2]: 1 where this rank owns the node (hence all its dofs).
 * After that mpet_solve (MINRES or GMRES) refreshes ghost entries after every SpMV, counts each dof once in the
 * dot products and all-reduces them.  The AMG V-cycles run distributed on the mesh-defined levels (P2, P1:
 * halo exchange after every smoothing step) and replicated below (restricted residual all-gathered), so the
 * preconditioner is the same operator for any number of GPUs. */
int mpet_nccl_unique_id(void* out128_host);
int mpet_attach_comm(mpet_ctx* ctx, const void* nccl_uid_host, int rank, int nranks);
int mpet_set_halo(mpet_ctx* ctx, int n_neighbours, const int* ranks_host, const int64_t* send_off_host,
                  const int32_t* send_nodes_dev, const int64_t* recv_off_host, const int32_t* recv_nodes_dev,
                  const uint8_t* owned_nodes_dev, void* stream);

/* ---- instrumentation ---------------------------------------------------------------------------
 * number of kernels this library launched since the last reset (bench.py's "gpu_launches") */
int64_t mpet_launch_count(mpet_ctx* ctx, int reset);
/* CUDA-event profile of the library's own launches (bench.py's live roofline measurement).
 * enable: 1/0 = reset counters and switch on/off, -1 = leave unchanged.  out_host (nullable) f64[16]:
 * [0..7] device ms per category (0 SpMV of A, 1 preconditioner, 3 assemble_lhs, 4 rhs_prev,
 * 5 halo exchanges + collectives on the main stream; categories 1 and 5 overlap),
 * [8..15] launch-group counts.  Synchronises the device. */
int mpet_profile(mpet_ctx* ctx, int enable, double* out_host);
int64_t mpet_device_bytes(mpet_ctx* ctx);
/* algorithmic bytes of the last preconditioner application (sum over all levels and passes: matrix stream,
 * gathered vector entries and row-wise operands once each); bench.py's second roofline entry */
int64_t mpet_pc_bytes(mpet_ctx* ctx);
/* how ghost entries travel: 0 single GPU, 1 NCCL send/recv + all-reduce, 2 peer-memory kernels over NVLink */
int mpet_comm_kind(mpet_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* MPET_B200_H */
