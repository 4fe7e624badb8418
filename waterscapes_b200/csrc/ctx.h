// Internal context of libmpet_b200 (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include <exception>
#include <stdexcept>
#include <cstdio>

#include "../../include/mpet_b200.h"
#include "staged.h"

#define MPET_MAX_NETWORKS 8

struct MpetError : public std::runtime_error {
    using std::runtime_error::runtime_error;
};

#define CUDA_CHECK(call)                                                                   \
    do {                                                                                   \
        cudaError_t e__ = (call);                                                          \
        if (e__ != cudaSuccess) {                                                          \
            char buf__[512];                                                               \
            snprintf(buf__, sizeof buf__, "%s:%d: %s -> %s", __FILE__, __LINE__, #call,    \
                     cudaGetErrorString(e__));                                             \
            throw MpetError(buf__);                                                        \
        }                                                                                  \
    } while (0)

#define MPET_REQUIRE(cond, msg)                                                            \
    do {                                                                                   \
        if (!(cond)) {                                                                     \
            char buf__[512];                                                               \
            snprintf(buf__, sizeof buf__, "%s:%d: %s", __FILE__, __LINE__, msg);           \
            throw MpetError(buf__);                                                        \
        }                                                                                  \
    } while (0)

// A scalar CSR graph between two node sets (P2 nodes = "2", P1 vertices = "1").
struct NodeGraph {
    int64_t nrows = 0, ncols = 0, nnz = 0;
    int32_t* rowptr = nullptr;   // [nrows+1]
    int32_t* col = nullptr;      // [nnz] ascending per row
    int32_t* gptr = nullptr;     // [nnz+1] offsets into glist (contributions per entry)
    uint32_t* glist = nullptr;   // [ncontrib] cell * npairs + local pair id, ascending cell per entry
    int64_t ncontrib = 0;
};

// Generic device CSR matrix (AMG levels, facet operators).
struct DevCsr {
    int64_t nrows = 0, ncols = 0, nnz = 0;
    int32_t* rowptr = nullptr;
    int32_t* col = nullptr;
    double* val = nullptr;
};

struct AmgLevel {
    DevCsr A;            // operator on this level (Dirichlet rows = identity)
    DevCsr P;            // prolongation from the next coarser level (nrows = this level)
    DevCsr R;            // restriction = P^T
    double* dinv = nullptr;   // inverse diagonal
    double* x = nullptr;      // [nrhs * n] work vectors
    double* b = nullptr;
    double* r = nullptr;
    double* t = nullptr;
    double lambda_max = 0.0;  // estimate of rho(D^-1 A)
    int halo_plan = -1;       // >= 0: level vectors are distributed (owned + ghost rows), refreshed with this plan
    bool is_mesh_level = false;   // finest level of a P1-field hierarchy (smoothing degree kP1Degree; coarser: kP1DegreeCoarse)
    SpmmPlan planA, planP, planR;   // staged-kernel chunk tables (nchunks == 0: not staged)
};

struct AmgHierarchy {
    std::vector<AmgLevel> levels;
    int nrhs = 1;
    double* coarse_inv = nullptr;  // dense inverse on the coarsest level
    int64_t coarse_n = 0;
    // polynomial mode (well-conditioned block): the whole "cycle" is poly_degree Chebyshev steps on [poly_lo, poly_hi]
    int poly_degree = 0;
    int poly_degree_light = 0;     // ... in light mode (ctx->pc_light)
    double* cyc_r = nullptr;       // work vectors of multi-cycle applications (development aid)
    double* cyc_e = nullptr;
    double poly_lo = 0.0, poly_hi = 0.0;
    int transition = -1;           // multi-GPU: index of the last distributed level (coarser ones are replicated)
    int64_t gather_count = 0;      // rows per rank in the padded replicated numbering
    double* gather_send = nullptr; // [gather_count * W] local slot of the all-gathered coarse right-hand side
};

// CUDA-event profiler: per-category device time of the launches made inside the library
enum { PROF_SPMV = 0, PROF_PC = 1, PROF_VEC = 2, PROF_ASM = 3, PROF_RHS = 4, PROF_COMM = 5, PROF_PC_U = 6, PROF_PC_P = 7, PROF_NCAT = 8 };
struct Prof {
    bool on = false;
    struct Span { int cat; cudaEvent_t a, b; };
    std::vector<Span> pending;
    std::vector<cudaEvent_t> pool;
    double ms[PROF_NCAT] = {0};
    int64_t cnt[PROF_NCAT] = {0};
};

// Coefficient tables of the generic Taylor-Hood block system [P2]^3 x [P1]^NP that the kernels assemble.
// Both formulations of the reference reduce to it (filled by mpet_set_params / mpet_set_params_total_pressure):
//   (u_k,a ; u_l,b)  = uu_mu * (delta_kl int grad phi_a . grad phi_b + int d_l phi_a d_k phi_b) + uu_lam * int d_k phi_a d_l phi_b
//   (u_k,a ; p_i,m)  = cup[i] * int psi_m d_k phi_a          (p_i,m ; u_l,b) = cpu[i] * int psi_m d_l phi_b
//   (p_i,m ; p_j,n)  = cm[i][j] * M_mn + delta_ij * cl[i] * L_mn          (M: P1 mass, L: P1 stiffness)
//   previous-state load of row p_i:  sum_j rm[i][j] M p_j^- + rl[i] L p_i^- + ru[i] * (the assembled (p_i ; u) block) u^-
//   preconditioner: p_mu * int grad phi_a . grad phi_b per displacement component; pm[i] M + pk[i] L per P1 field
struct BlockCoefs {
    double uu_mu = 0, uu_lam = 0, p_mu = 0;
    double cup[MPET_MAX_NETWORKS] = {}, cpu[MPET_MAX_NETWORKS] = {};
    double cm[MPET_MAX_NETWORKS * MPET_MAX_NETWORKS] = {}, cl[MPET_MAX_NETWORKS] = {};
    double rm[MPET_MAX_NETWORKS * MPET_MAX_NETWORKS] = {}, rl[MPET_MAX_NETWORKS] = {}, ru[MPET_MAX_NETWORKS] = {};
    double pm[MPET_MAX_NETWORKS] = {}, pk[MPET_MAX_NETWORKS] = {};
};

struct mpet_ctx {
    int device = 0;
    Prof prof;
    std::string err;
    int64_t launches = 0;
    int64_t bytes = 0;
    int64_t pc_bytes_acc = 0;     // algorithmic bytes of the preconditioner application in flight (amg.cu)
    int64_t pc_bytes_last = 0;    // ... of the last complete application (mpet_pc_bytes)
    std::vector<void*> allocs;
    std::vector<void*> amg_allocs;          // freed and rebuilt by every mpet_pc_setup
    std::vector<void*>* arena = nullptr;    // when set, dev_alloc records here instead of allocs
    int sm_count = 148;

    // mesh / space
    int64_t Nv = 0, Ne = 0, N2 = 0, Nc = 0, N = 0, nnz = 0;
    int64_t Nint = 0;            // 4*N2 + A*Nv: length of solver-internal vectors (layout.cuh)
    int A = 0;                   // networks
    int nloc = 0;
    double* coords = nullptr;    // [Nv*3] (library copy)
    int32_t* cells = nullptr;    // [Nc*4]
    int32_t* edge_v = nullptr;   // [Ne*2]
    int32_t* cell_nodes = nullptr;  // [Nc*10] scalar P2 nodes of each cell
    double* geom = nullptr;      // [Nc*10]: Jinv (row-major 3x3: Jinv[k][m] = dX_k/dx_m), |detJ|
    NodeGraph g22, g21, g12, g11;
    int32_t* t21to12 = nullptr;  // [nnz21] position of the transposed entry in g12
    int64_t* rowptr = nullptr;   // [N+1] block-system CSR
    int32_t* cols = nullptr;     // [nnz]
    double* vals = nullptr;      // [nnz] A as assembled (never modified by BCs)
    // scalar operators kept beside A
    double* m11 = nullptr;       // [nnz11] P1 mass
    double* l11 = nullptr;       // [nnz11] P1 stiffness
    double* m22 = nullptr;       // [nnz22] P2 mass (lazy)
    double* k22 = nullptr;       // [nnz22] P2 stiffness (grad,grad) (prec, lazy)
    double* pp11 = nullptr;      // [A*nnz11] preconditioner pressure blocks
    // DG0 (cell-wise constant) permeability of a P1 field: per-cell values and the stiffness
    // weighted with them; fields without one use l11 times their scalar K
    double* kcell[MPET_MAX_NETWORKS] = {};   // [Nc]
    double* l11w[MPET_MAX_NETWORKS] = {};    // [nnz11]
    bool cell_coef_dirty = false;

    // coefficients
    double E = 0, nu = 0, mu = 0, lmbda = 0, dt = 0, theta = 1;
    double alpha[MPET_MAX_NETWORKS], K[MPET_MAX_NETWORKS], c[MPET_MAX_NETWORKS];
    double S[MPET_MAX_NETWORKS * MPET_MAX_NETWORKS];
    BlockCoefs coef;             // what the kernels read (either formulation)
    int formulation = 0;         // 0: standard (mpetsolver.py), 1: total pressure (mpettotalpressuresolver.py)
    bool params_set = false, lhs_ready = false, prec_ready = false;

    // Dirichlet
    int64_t n_bc = 0;
    int32_t* bc_dofs = nullptr;
    double* bc_vals = nullptr;
    uint8_t* bc_mask = nullptr;  // [N] 1 on Dirichlet rows (API numbering)
    uint8_t* bc_mask_int = nullptr;   // [Nint] same in the solver-internal layout
    int32_t* bc_dofs_int = nullptr;   // [n_bc] internal indices of the Dirichlet dofs
    int* d_missing = nullptr;         // mpet_add_entries: count of entries outside the pattern
    double* scratch_int[2] = {nullptr, nullptr};   // [Nint] layout-conversion buffers
    BlockPlan plan_u, plan_p;         // chunk tables of the staged block SpMV
    bool staged_ok = false;

    // Lagrange multipliers bordering the system (nullspace handling)
    int nb = 0;                      // number of multipliers (<= 16)
    // external dof numbering at the ABI (mpet_set_dof_permutation): perm[c] = caller's index of contract dof c
    int32_t* perm = nullptr;
    int32_t* perm_inv = nullptr;
    double* perm_a = nullptr;        // [N + 16] scratch vectors in the contract numbering
    double* perm_b = nullptr;
    double* border = nullptr;        // [nb][Nint] columns c_i in the solver-internal layout
    double* border_scale = nullptr;  // [16] c_i . B c_i
    bool border_scaled = false;
    double prec_shift_u = 0.0;       // + shift * (u, v) in the displacement block of the preconditioner
    double prec_shift_p[MPET_MAX_NETWORKS] = {};   // + shift_i * (p_i, q_i)

    // Krylov
    int method = 0, pc = 2, maxit = 10000, restart = 30;
    double rtol = 1e-5, atol = 1e-50;
    bool pc_light = false;       // light P1-field cycles (amg.cu): set per solve from rtol
    int norm_mode = 0;           // reference norm of the convergence test: 0 = |b|_B (PETSc default), 1 = min(|b|_B, |r0|_B)
    struct KrylovWork* kw = nullptr;
    struct DistState* dist = nullptr;           // NCCL communicator + halo plan (dist.cu)
    AmgHierarchy* amg_u = nullptr;              // scalar P2 block, 3 right-hand sides
    AmgHierarchy* amg_p[MPET_MAX_NETWORKS] = {};  // one per P1 field
    double* jac_dinv = nullptr;                 // Jacobi preconditioner (pc = 1)
    // the per-network V-cycles are independent of the displacement V-cycle: they run on their own streams
    cudaStream_t pc_stream[MPET_MAX_NETWORKS] = {};
    cudaEvent_t pc_fork = nullptr, pc_join[MPET_MAX_NETWORKS] = {};
    int pc_streams_ready = 0;
    cudaStream_t solve_stream = nullptr;     // used when the caller hands us the (uncapturable) legacy default stream
    cudaEvent_t solve_event = nullptr;
    int dist_lane = 0;           // which lane (flags + staging areas / communicator) dist_halo and dist_allgather use (dist.cu)
    bool prof_suspended = false; // inside a stream capture: no profiler events
    int64_t graph_epoch = 0;     // bumped whenever pointers baked into captured launches change (graphs are re-captured)
};

// ---- helpers --------------------------------------------------------------------------------
template <typename T>
T* dev_alloc(mpet_ctx* ctx, int64_t n) {
    if (n <= 0) n = 1;
    void* p = nullptr;
    // 64 bytes of slack: the bulk-copy kernels read 16-byte aligned supersets of array slices
    CUDA_CHECK(cudaMalloc(&p, (size_t)n * sizeof(T) + 64));
    (ctx->arena ? *ctx->arena : ctx->allocs).push_back(p);
    ctx->bytes += n * (int64_t)sizeof(T);
    return (T*)p;
}
void dev_free(mpet_ctx* ctx, void* p);

#define MPET_TRY(ctx) try {
#define MPET_CATCH(ctx)                                                    \
    }                                                                      \
    catch (const std::exception& e__) {                                    \
        if (ctx) (ctx)->err = e__.what();                                  \
        return -1;                                                         \
    }                                                                      \
    return 0;

inline cudaStream_t as_stream(void* s) { return (cudaStream_t)s; }

inline int grid_for(int64_t work, int block) {
    int64_t g = (work + block - 1) / block;
    if (g < 1) g = 1;
    if (g > 2147483647LL) g = 2147483647LL;
    return (int)g;
}

#define LAUNCH_CHECK(ctx)             \
    do {                              \
        (ctx)->launches++;            \
        CUDA_CHECK(cudaGetLastError()); \
    } while (0)

// profiler (api.cu)
cudaEvent_t prof_begin(mpet_ctx* ctx, cudaStream_t st);
void prof_end(mpet_ctx* ctx, int cat, cudaEvent_t a, cudaStream_t st);
void prof_collect(mpet_ctx* ctx);

// graph.cu
void build_space(mpet_ctx* ctx, cudaStream_t st);
// assemble.cu
void init_reference_tables();
void compute_geometry(mpet_ctx* ctx, cudaStream_t st);
void assemble_lhs(mpet_ctx* ctx, cudaStream_t st);
void assemble_prec(mpet_ctx* ctx, cudaStream_t st);
void ensure_m22(mpet_ctx* ctx, cudaStream_t st);
// amg.cu
void amg_setup(mpet_ctx* ctx, cudaStream_t st);
void amg_apply(mpet_ctx* ctx, const double* r, double* z, const int* done, cudaStream_t st);
void amg_free(mpet_ctx* ctx);
// staged.cu
void staged_build_block_plans(mpet_ctx* ctx);
bool staged_block_spmv(mpet_ctx* ctx, const double* x, double* y, const uint8_t* mask, const int* done, cudaStream_t st);
bool staged_build_spmm_plan(mpet_ctx* ctx, const std::vector<int32_t>& rp, SpmmPlan& plan);
void staged_spmm(mpet_ctx* ctx, int W, int epi, const SpmmPlan& P, const DevCsr& M, const double* x, const double* b,
                 double* out, double* d, const double* dinv, double c1, double c2, const int* done, cudaStream_t st);
// dist.cu (multi-GPU)
// P1WA: the A vertex fields of the merged pressure hierarchy (entry i*Nv + v)
enum { DIST_PLAN_KRYLOV = 0, DIST_PLAN_P2W4 = 1, DIST_PLAN_P1W4 = 2, DIST_PLAN_P1W1 = 3, DIST_PLAN_P1WA = 4, DIST_NPLANS = 5 };
bool dist_active(mpet_ctx* ctx);
int dist_comm_kind(mpet_ctx* ctx);      // 0: single GPU, 1: NCCL send/recv + all-reduce, 2: peer-memory kernels over NVLink
bool dist_has_lane1(mpet_ctx* ctx);
int dist_rank(mpet_ctx* ctx);
int dist_nranks(mpet_ctx* ctx);
const uint8_t* dist_owned_mask(mpet_ctx* ctx);
const std::vector<uint8_t>& dist_own_nodes(mpet_ctx* ctx);
void dist_halo(mpet_ctx* ctx, int plan, double* v, bool reverse, const int* done, cudaStream_t st);
void dist_allreduce_sum(mpet_ctx* ctx, double* dev_scalars, int count, cudaStream_t st);
void dist_allreduce_max(mpet_ctx* ctx, double* dev_scalars, int count, cudaStream_t st);
void dist_allgather(mpet_ctx* ctx, const double* send, double* recv, int64_t count, const int* done, cudaStream_t st);
void dist_reserve_gather(mpet_ctx* ctx, int64_t doubles, cudaStream_t st);
bool dist_reduce_partials(mpet_ctx* ctx, const double* partials, int nparts, int count, double* out, const int* done,
                          cudaStream_t st);
void dist_check(mpet_ctx* ctx);
void dist_bcast_bytes(mpet_ctx* ctx, void* buf, int64_t nbytes, int root, cudaStream_t st);
// spmv.cu
void csr_spmv(mpet_ctx* ctx, int64_t nrows, const int64_t* rowptr, const int32_t* cols,
              const double* vals, const double* x, double* y, double beta, const uint8_t* rowmask,
              cudaStream_t st, const int* done = nullptr);
void block_spmv(mpet_ctx* ctx, const double* x_int, double* y_int, const uint8_t* mask_int, const int* done,
                cudaStream_t st);
void to_internal(mpet_ctx* ctx, const double* src_api, double* dst_int, cudaStream_t st);
void to_api(mpet_ctx* ctx, const double* src_int, double* dst_api, cudaStream_t st);
void csr32_spmm(mpet_ctx* ctx, const DevCsr& M, const double* x, int64_t ldx, double* y, int64_t ldy,
                int nrhs, double alpha, double beta, cudaStream_t st);
