// CSR sparse matrix-vector kernels (fp64 values, int32 columns), HBM-bandwidth-bound.
//
// Replaces PETSc MatMult inside KSPSolve (PETScKrylovSolver.solve, mpetsolver.py:556) and the
// matrix-vector products of BoomerAMG's V-cycle.
//
// k_spmv_vec<LANES>: LANES threads cooperate on one row (LANES = 32 for the block system whose rows
// hold 100-250 entries; 8/16 for the short rows of the scalar / AMG-level matrices: "segment
// selection" by mean row length).  Values and columns are streamed with 16-byte / 8-byte
// L1::no_allocate loads on an aligned body (row starts are peeled to even offsets) so that the x
// vector, gathered through the read-only path, keeps L1/L2 to itself; partial sums are combined with
// warp shuffles in a fixed order (deterministic).
#include "ctx.h"

namespace {

__device__ __forceinline__ double2 ld_stream_f64x2(const double* p) {
    double2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ int2 ld_stream_s32x2(const int32_t* p) {
    int2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.s32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ double ld_stream_f64(const double* p) {
    double r;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ int32_t ld_stream_s32(const int32_t* p) {
    int32_t r;
    asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}

template <int LANES>
__device__ __forceinline__ double group_reduce(double v) {
#pragma unroll
    for (int o = LANES / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o, LANES);
    return v;
}

// y = A x (+ beta y).  rowmask (optional): rows flagged 1 return x[row] (identity rows of the
// symmetrically eliminated Dirichlet dofs; see DESIGN.md "Dirichlet conditions").
template <int LANES, typename RP>
__global__ void __launch_bounds__(256)
k_spmv_vec(int64_t nrows, const RP* __restrict__ rowptr, const int32_t* __restrict__ cols,
           const double* __restrict__ vals, const double* __restrict__ x, double* __restrict__ y,
           double beta, const uint8_t* __restrict__ rowmask, const int* __restrict__ done) {
    if (done && *done) return;
    const int lane = threadIdx.x & (LANES - 1);
    int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / LANES;
    if (row >= nrows) return;   // whole groups exit together (blockDim % LANES == 0)
    if (rowmask && rowmask[row]) {
        if (lane == 0) y[row] = x[row];
        return;
    }
    const int64_t s = rowptr[row], e = rowptr[row + 1];
    const int64_t s2 = (s + 1) & ~(int64_t)1, e2 = e & ~(int64_t)1;
    double sum = 0.0;
    if (lane == 0 && s < s2 && s < e) sum += ld_stream_f64(vals + s) * __ldg(x + ld_stream_s32(cols + s));
    if (lane == (1 % LANES) && e2 < e && e2 >= s2)
        sum += ld_stream_f64(vals + e2) * __ldg(x + ld_stream_s32(cols + e2));
    int64_t i = s2 + 2 * lane;
    // two independent 24-byte requests in flight per lane
    for (; i + 2 * LANES < e2; i += 4 * LANES) {
        double2 v0 = ld_stream_f64x2(vals + i);
        int2 c0 = ld_stream_s32x2(cols + i);
        double2 v1 = ld_stream_f64x2(vals + i + 2 * LANES);
        int2 c1 = ld_stream_s32x2(cols + i + 2 * LANES);
        sum += v0.x * __ldg(x + c0.x);
        sum += v0.y * __ldg(x + c0.y);
        sum += v1.x * __ldg(x + c1.x);
        sum += v1.y * __ldg(x + c1.y);
    }
    if (i < e2) {
        double2 v0 = ld_stream_f64x2(vals + i);
        int2 c0 = ld_stream_s32x2(cols + i);
        sum += v0.x * __ldg(x + c0.x);
        sum += v0.y * __ldg(x + c0.y);
    }
    sum = group_reduce<LANES>(sum);
    if (lane == 0) y[row] = (beta == 0.0) ? sum : sum + beta * y[row];
}

// Y = alpha * M X + beta * Y for NRHS column vectors (leading dimensions ldx, ldy): the matrix is
// streamed once for all right-hand sides (the three displacement components share one scalar block).
template <int LANES, int NRHS>
__global__ void __launch_bounds__(256)
k_spmm32(int64_t nrows, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ cols,
         const double* __restrict__ vals, const double* __restrict__ x, int64_t ldx,
         double* __restrict__ y, int64_t ldy, double alpha, double beta) {
    const int lane = threadIdx.x & (LANES - 1);
    int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / LANES;
    if (row >= nrows) return;
    const int32_t s = rowptr[row], e = rowptr[row + 1];
    double sum[NRHS];
#pragma unroll
    for (int k = 0; k < NRHS; ++k) sum[k] = 0.0;
    for (int32_t i = s + lane; i < e; i += LANES) {
        double v = ld_stream_f64(vals + i);
        int32_t c = ld_stream_s32(cols + i);
#pragma unroll
        for (int k = 0; k < NRHS; ++k) sum[k] += v * __ldg(x + k * ldx + c);
    }
#pragma unroll
    for (int k = 0; k < NRHS; ++k) {
        double r = group_reduce<LANES>(sum[k]);
        if (lane == 0) {
            double* yp = y + k * ldy + row;
            *yp = (beta == 0.0) ? alpha * r : alpha * r + beta * (*yp);
        }
    }
}

template <int NRHS>
void launch_spmm(mpet_ctx* ctx, const DevCsr& M, const double* x, int64_t ldx, double* y, int64_t ldy,
                 double alpha, double beta, cudaStream_t st) {
    double mean = M.nrows > 0 ? (double)M.nnz / (double)M.nrows : 0.0;
    const int threads = 256;
    if (mean > 24) {
        k_spmm32<32, NRHS><<<grid_for(M.nrows * 32, threads), threads, 0, st>>>(
            M.nrows, M.rowptr, M.col, M.val, x, ldx, y, ldy, alpha, beta);
    } else if (mean > 10) {
        k_spmm32<16, NRHS><<<grid_for(M.nrows * 16, threads), threads, 0, st>>>(
            M.nrows, M.rowptr, M.col, M.val, x, ldx, y, ldy, alpha, beta);
    } else if (mean > 3) {
        k_spmm32<8, NRHS><<<grid_for(M.nrows * 8, threads), threads, 0, st>>>(
            M.nrows, M.rowptr, M.col, M.val, x, ldx, y, ldy, alpha, beta);
    } else {
        k_spmm32<2, NRHS><<<grid_for(M.nrows * 2, threads), threads, 0, st>>>(
            M.nrows, M.rowptr, M.col, M.val, x, ldx, y, ldy, alpha, beta);
    }
    LAUNCH_CHECK(ctx);
}

}  // namespace

void csr_spmv(mpet_ctx* ctx, int64_t nrows, const int64_t* rowptr, const int32_t* cols,
              const double* vals, const double* x, double* y, double beta, const uint8_t* rowmask,
              cudaStream_t st, const int* done) {
    const int threads = 256;
    k_spmv_vec<32, int64_t><<<grid_for(nrows * 32, threads), threads, 0, st>>>(nrows, rowptr, cols, vals, x,
                                                                                y, beta, rowmask, done);
    LAUNCH_CHECK(ctx);
}

void csr32_spmm(mpet_ctx* ctx, const DevCsr& M, const double* x, int64_t ldx, double* y, int64_t ldy,
                int nrhs, double alpha, double beta, cudaStream_t st) {
    if (M.nrows == 0) return;
    if (nrhs == 1) launch_spmm<1>(ctx, M, x, ldx, y, ldy, alpha, beta, st);
    else if (nrhs == 3) launch_spmm<3>(ctx, M, x, ldx, y, ldy, alpha, beta, st);
    else MPET_REQUIRE(false, "csr32_spmm: nrhs must be 1 or 3");
}
