// Right-hand-side and Dirichlet kernels.
//
//  * k_rhs_prev     : b = assemble(L), L = rhs(F)  (mpetsolver.py:198-201,261,356,433,528) evaluated
//                     as a sparse operator acting on the previous state instead of a cell loop:
//                     pressure row (i,v) = ru[i] * [A's own "pu" coupling block] * u_prev
//                                        + sum_j rm[i][j] M p_j_prev + rl[i] L p_i_prev ;  momentum rows = 0
//                     (BlockCoefs in ctx.h; standard formulation: ru = 1, rm[i][i] = -c_i + dt(1-theta) sum_j S_ij,
//                     rl[i] = dt(1-theta) K_i, rm[i][j] = -dt(1-theta) S_ij).
//  * k_bc_values    : export-time application of bc.apply(A) / apply_symmetric(bc, A)
//                     (mpetsolver.py:343-344,418-419; bc_symmetric.py:11-22).  The solver itself never
//                     modifies A: it masks rows on the fly (spmv.cu).
//  * k_scatter_vals : bc.apply(b)  (mpetsolver.py:375-376,452-453).
#include "ctx.h"
#include "layout.cuh"

namespace {

struct PrevCoefs {
    double c[2 * MPET_MAX_NETWORKS + MPET_MAX_NETWORKS * MPET_MAX_NETWORKS];
    const double* lw[MPET_MAX_NETWORKS];      // DG0-weighted stiffness of a field (null: l11)
};

__global__ void __launch_bounds__(256)
k_rhs_prev(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ cols,
           const double* __restrict__ vals, const int32_t* __restrict__ rp12,
           const int32_t* __restrict__ rp11, const int32_t* __restrict__ col11,
           const double* __restrict__ m11, const double* __restrict__ l11, int64_t n2, int64_t nv, int A,
           const PrevCoefs K /* by value: no per-step allocation */,
           const double* __restrict__ up, double* __restrict__ b) {
    const double* coef = K.c;       // [A] ru, [A] rl, [A*A] rm
    int64_t w = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (w >= (int64_t)A * nv) return;
    int i = (int)(w / nv);
    int64_t v = w - (int64_t)i * nv;
    int64_t row = 3 * n2 + w;
    int64_t s = rowptr[row];
    int32_t d12 = rp12[v + 1] - rp12[v];
    double sum = 0.0;
    const double ru = coef[i];
    if (ru != 0.0) {
        for (int32_t t = lane; t < 3 * d12; t += 32) sum += vals[s + t] * up[cols[s + t]];
        sum *= ru;
    }
    const double* pbase = up + 3 * n2;
    for (int32_t e = rp11[v] + lane; e < rp11[v + 1]; e += 32) {
        int32_t vp = col11[e];
        double M = m11[e], L = K.lw[i] ? K.lw[i][e] : l11[e];
        double acc = coef[A + i] * L * pbase[(int64_t)i * nv + vp];
        for (int j = 0; j < A; ++j) acc += coef[2 * A + i * A + j] * M * pbase[(int64_t)j * nv + vp];
        sum += acc;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0) b[row] = sum;
}

__global__ void __launch_bounds__(256)
k_bc_values(int64_t nrows, const int64_t* __restrict__ rowptr, const int32_t* __restrict__ cols,
            const uint8_t* __restrict__ mask, int mode, double* __restrict__ vals) {
    int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (row >= nrows) return;
    bool rd = mask[row] != 0;
    for (int64_t t = rowptr[row] + lane; t < rowptr[row + 1]; t += 32) {
        int32_t c = cols[t];
        bool cd = mask[c] != 0;
        if (rd) vals[t] = (c == row) ? 1.0 : 0.0;
        else if (mode == 2 && cd) vals[t] = 0.0;
    }
}

// expand the block-diagonal preconditioner onto the block-system pattern
__global__ void __launch_bounds__(256)
k_expand_prec(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ rp22,
              const int32_t* __restrict__ rp21, const int32_t* __restrict__ rp12,
              const int32_t* __restrict__ rp11, const double* __restrict__ k22,
              const double* __restrict__ pp11, int64_t nnz11, double mu, int64_t n2, int64_t nv, int A,
              double* __restrict__ out) {
    int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    int64_t N = 3 * n2 + (int64_t)A * nv;
    if (row >= N) return;
    int64_t s = rowptr[row], e = rowptr[row + 1];
    for (int64_t t = s + lane; t < e; t += 32) out[t] = 0.0;
    __syncwarp();
    if (row < 3 * n2) {
        int k = (int)(row / n2);
        int64_t a = row - k * n2;
        int32_t d22 = rp22[a + 1] - rp22[a];
        for (int32_t j = lane; j < d22; j += 32) out[s + (int64_t)k * d22 + j] = mu * k22[rp22[a] + j];
    } else {
        int64_t w = row - 3 * n2;
        int i = (int)(w / nv);
        int64_t v = w - (int64_t)i * nv;
        int32_t d12 = rp12[v + 1] - rp12[v];
        int32_t d11 = rp11[v + 1] - rp11[v];
        for (int32_t j = lane; j < d11; j += 32)
            out[s + 3 * (int64_t)d12 + (int64_t)i * d11 + j] = pp11[(int64_t)i * nnz11 + rp11[v] + j];
    }
}

__global__ void k_scatter_vals(const int32_t* __restrict__ idx, const double* __restrict__ v, int64_t n,
                               double* __restrict__ out) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) out[idx[i]] = v[i];
}

__global__ void k_set_mask(const int32_t* __restrict__ idx, int64_t n, int64_t n2, uint8_t* __restrict__ mask,
                           uint8_t* __restrict__ mask_int, int32_t* __restrict__ idx_int) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    mask[idx[i]] = 1;
    int64_t j = api_to_internal(idx[i], n2);
    mask_int[j] = 1;
    idx_int[i] = (int32_t)j;
}

// A[row, col] += val for unique (row, col) pairs that exist in the pattern (binary search per entry)
__global__ void k_add_entries(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ cols,
                              const int32_t* __restrict__ r, const int32_t* __restrict__ c,
                              const double* __restrict__ v, int64_t n, double* __restrict__ vals,
                              int* __restrict__ missing) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    int64_t lo = rowptr[r[i]], hi = rowptr[r[i] + 1];
    int32_t target = c[i];
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (cols[mid] < target) lo = mid + 1; else hi = mid;
    }
    if (lo < rowptr[r[i] + 1] && cols[lo] == target) vals[lo] += v[i];
    else atomicAdd(missing, 1);
}

}  // namespace

void rhs_prev(mpet_ctx* ctx, const double* up, double* b, cudaStream_t st) {
    MPET_REQUIRE(ctx->lhs_ready, "mpet_assemble_lhs must run before mpet_rhs_prev");
    const int A = ctx->A;
    CUDA_CHECK(cudaMemsetAsync(b, 0, sizeof(double) * 3 * ctx->N2, st));
    if (A == 0) return;
    PrevCoefs K;
    double* coef = K.c;
    for (int f = 0; f < MPET_MAX_NETWORKS; ++f) K.lw[f] = ctx->kcell[f] ? ctx->l11w[f] : nullptr;
    for (int i = 0; i < A; ++i) {
        coef[i] = ctx->coef.ru[i];
        coef[A + i] = ctx->coef.rl[i];
        for (int j = 0; j < A; ++j) coef[2 * A + i * A + j] = ctx->coef.rm[i * A + j];
    }
    k_rhs_prev<<<grid_for((int64_t)A * ctx->Nv * 32, 256), 256, 0, st>>>(
        ctx->rowptr, ctx->cols, ctx->vals, ctx->g12.rowptr, ctx->g11.rowptr, ctx->g11.col, ctx->m11,
        ctx->l11, ctx->N2, ctx->Nv, A, K, up, b);
    LAUNCH_CHECK(ctx);
}

void export_values(mpet_ctx* ctx, int which, double* out, cudaStream_t st) {
    if (which == 3) {
        MPET_REQUIRE(ctx->prec_ready, "mpet_assemble_prec must run first");
        k_expand_prec<<<grid_for(ctx->N * 32, 256), 256, 0, st>>>(
            ctx->rowptr, ctx->g22.rowptr, ctx->g21.rowptr, ctx->g12.rowptr, ctx->g11.rowptr, ctx->k22,
            ctx->pp11, ctx->g11.nnz, ctx->coef.p_mu, ctx->N2, ctx->Nv, ctx->A, out);
        LAUNCH_CHECK(ctx);
    } else {
        MPET_REQUIRE(ctx->lhs_ready, "mpet_assemble_lhs must run first");
        CUDA_CHECK(cudaMemcpyAsync(out, ctx->vals, sizeof(double) * ctx->nnz, cudaMemcpyDeviceToDevice, st));
    }
    if (which >= 1 && ctx->bc_mask) {
        int mode = (which == 1) ? 1 : 2;
        k_bc_values<<<grid_for(ctx->N * 32, 256), 256, 0, st>>>(ctx->N, ctx->rowptr, ctx->cols, ctx->bc_mask,
                                                                 mode, out);
        LAUNCH_CHECK(ctx);
    }
}

void set_dirichlet_dofs(mpet_ctx* ctx, const int32_t* dofs, int64_t n, cudaStream_t st) {
    ctx->graph_epoch++;               // mask pointers / presence are baked into captured launches
    if (ctx->bc_dofs) { dev_free(ctx, ctx->bc_dofs); dev_free(ctx, ctx->bc_vals); dev_free(ctx, ctx->bc_dofs_int); }
    if (!ctx->bc_mask) ctx->bc_mask = dev_alloc<uint8_t>(ctx, ctx->N);
    if (!ctx->bc_mask_int) {      // + the pad behind the multipliers of a bordered system (never masked)
        ctx->bc_mask_int = dev_alloc<uint8_t>(ctx, ctx->Nint + 64);
        CUDA_CHECK(cudaMemsetAsync(ctx->bc_mask_int, 0, ctx->Nint + 64, st));
    }
    ctx->bc_dofs_int = dev_alloc<int32_t>(ctx, n);
    CUDA_CHECK(cudaMemsetAsync(ctx->bc_mask_int, 0, ctx->Nint, st));
    ctx->n_bc = n;
    ctx->bc_dofs = dev_alloc<int32_t>(ctx, n);
    ctx->bc_vals = dev_alloc<double>(ctx, n);
    CUDA_CHECK(cudaMemcpyAsync(ctx->bc_dofs, dofs, sizeof(int32_t) * n, cudaMemcpyDeviceToDevice, st));
    CUDA_CHECK(cudaMemsetAsync(ctx->bc_vals, 0, sizeof(double) * (n > 0 ? n : 1), st));
    CUDA_CHECK(cudaMemsetAsync(ctx->bc_mask, 0, ctx->N, st));
    if (n > 0) {
        k_set_mask<<<grid_for(n, 256), 256, 0, st>>>(ctx->bc_dofs, n, ctx->N2, ctx->bc_mask, ctx->bc_mask_int,
                                                     ctx->bc_dofs_int);
        LAUNCH_CHECK(ctx);
    }
}

void scatter_bc_values_int(mpet_ctx* ctx, double* out_int, cudaStream_t st) {
    if (ctx->n_bc == 0) return;
    k_scatter_vals<<<grid_for(ctx->n_bc, 256), 256, 0, st>>>(ctx->bc_dofs_int, ctx->bc_vals, ctx->n_bc, out_int);
    LAUNCH_CHECK(ctx);
}

void scatter_bc_values(mpet_ctx* ctx, double* out, cudaStream_t st) {
    if (ctx->n_bc == 0) return;
    k_scatter_vals<<<grid_for(ctx->n_bc, 256), 256, 0, st>>>(ctx->bc_dofs, ctx->bc_vals, ctx->n_bc, out);
    LAUNCH_CHECK(ctx);
}

void add_entries(mpet_ctx* ctx, const int32_t* r, const int32_t* c, const double* v, int64_t n,
                 cudaStream_t st) {
    MPET_REQUIRE(ctx->lhs_ready, "mpet_assemble_lhs must run before mpet_add_entries");
    if (n == 0) return;
    if (!ctx->d_missing) ctx->d_missing = dev_alloc<int>(ctx, 1);
    int* d_missing = ctx->d_missing;
    CUDA_CHECK(cudaMemsetAsync(d_missing, 0, sizeof(int), st));
    k_add_entries<<<grid_for(n, 256), 256, 0, st>>>(ctx->rowptr, ctx->cols, r, c, v, n, ctx->vals, d_missing);
    LAUNCH_CHECK(ctx);
    int missing = 0;
    CUDA_CHECK(cudaMemcpyAsync(&missing, d_missing, sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
    MPET_REQUIRE(missing == 0, "mpet_add_entries: entry outside the sparsity pattern");
}
