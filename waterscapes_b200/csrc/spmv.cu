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
#include "layout.cuh"

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


// ------------------------------------------------------------------------------ node-blocked SpMV
// y = A x on the block system, vectors in the solver-internal layout (layout.cuh).  The kernels run
// over the SCALAR node graphs: a P2 node pair (a, b) owns the 3x3 values A[(k,a),(l,b)], which sit in
// the ordinary CSR value array at rowptr[k*N2+a] + l*deg22(a) + j.  One warp handles node a (its three
// rows): lane j streams the nine values (coalesced across lanes for each (k,l)), reads ONE column
// index and gathers the neighbour's displacement with ONE 256-bit load.  Bytes per node pair:
// 72 (values) + 4 (index) instead of 108 for scalar CSR, and 1 gather wavefront instead of 9.
template <int NA>
__global__ void __launch_bounds__(256)
k_spmv_block_u(int64_t n2, int64_t nv, const int32_t* __restrict__ rp22, const int32_t* __restrict__ col22,
               const int32_t* __restrict__ rp21, const int32_t* __restrict__ col21,
               const int64_t* __restrict__ rowptr, const double* __restrict__ vals,
               const double* __restrict__ x, double* __restrict__ y, const uint8_t* __restrict__ mask,
               const int* __restrict__ done) {
    if (done && *done) return;
    int64_t a = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (a >= n2) return;
    const int32_t e0 = rp22[a], d22 = rp22[a + 1] - e0;
    const int32_t f0 = rp21[a], d21 = rp21[a + 1] - f0;
    const double* v0 = vals + rowptr[a];
    const double* v1 = vals + rowptr[n2 + a];
    const double* v2 = vals + rowptr[2 * n2 + a];
    double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0;
    for (int32_t j = lane; j < d22; j += 32) {
        const int32_t c = ldg_stream_s32(col22 + e0 + j);
        // nine independent streaming loads are in flight before the dependent gather
        const double a00 = ldg_stream_f64(v0 + j), a01 = ldg_stream_f64(v0 + d22 + j), a02 = ldg_stream_f64(v0 + 2 * d22 + j);
        const double a10 = ldg_stream_f64(v1 + j), a11 = ldg_stream_f64(v1 + d22 + j), a12 = ldg_stream_f64(v1 + 2 * d22 + j);
        const double a20 = ldg_stream_f64(v2 + j), a21 = ldg_stream_f64(v2 + d22 + j), a22 = ldg_stream_f64(v2 + 2 * d22 + j);
        const d4 xb = ld256_gather(x + 4 * (int64_t)c);
        acc0 += a00 * xb.x + a01 * xb.y + a02 * xb.z;
        acc1 += a10 * xb.x + a11 * xb.y + a12 * xb.z;
        acc2 += a20 * xb.x + a21 * xb.y + a22 * xb.z;
    }
    if (NA > 0) {
        const double* p = x + 4 * n2;
        const int64_t off = 3 * (int64_t)d22;
        for (int32_t j = lane; j < d21; j += 32) {
            const int32_t v = ldg_stream_s32(col21 + f0 + j);
#pragma unroll
            for (int i = 0; i < NA; ++i) {
                const double pv = __ldg(p + (int64_t)i * nv + v);
                const int64_t o = off + (int64_t)i * d21 + j;
                acc0 += ldg_stream_f64(v0 + o) * pv;
                acc1 += ldg_stream_f64(v1 + o) * pv;
                acc2 += ldg_stream_f64(v2 + o) * pv;
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        acc0 += __shfl_xor_sync(0xffffffffu, acc0, o);
        acc1 += __shfl_xor_sync(0xffffffffu, acc1, o);
        acc2 += __shfl_xor_sync(0xffffffffu, acc2, o);
    }
    if (lane == 0) {
        d4 out = {acc0, acc1, acc2, 0.0};
        if (mask) {
            const uint32_t m = *reinterpret_cast<const uint32_t*>(mask + 4 * a);
            if (m) {
                const d4 xa = ld256(x + 4 * a);
                if (m & 0x000000ffu) out.x = xa.x;
                if (m & 0x0000ff00u) out.y = xa.y;
                if (m & 0x00ff0000u) out.z = xa.z;
            }
        }
        st256(y + 4 * a, out);
    }
}

// pressure rows: one warp per vertex handles the rows of all NA networks
template <int NA>
__global__ void __launch_bounds__(256)
k_spmv_block_p(int64_t n2, int64_t nv, const int32_t* __restrict__ rp12, const int32_t* __restrict__ col12,
               const int32_t* __restrict__ rp11, const int32_t* __restrict__ col11,
               const int64_t* __restrict__ rowptr, const double* __restrict__ vals,
               const double* __restrict__ x, double* __restrict__ y, const uint8_t* __restrict__ mask,
               const int* __restrict__ done) {
    if (done && *done) return;
    int64_t v = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (v >= nv) return;
    const int32_t g0 = rp12[v], d12 = rp12[v + 1] - g0;
    const int32_t h0 = rp11[v], d11 = rp11[v + 1] - h0;
    const double* vb[NA];
    double acc[NA];
#pragma unroll
    for (int i = 0; i < NA; ++i) {
        vb[i] = vals + rowptr[3 * n2 + (int64_t)i * nv + v];
        acc[i] = 0.0;
    }
    for (int32_t j = lane; j < d12; j += 32) {
        const int32_t a = ldg_stream_s32(col12 + g0 + j);
        double w[NA][3];
#pragma unroll
        for (int i = 0; i < NA; ++i)
#pragma unroll
            for (int k = 0; k < 3; ++k) w[i][k] = ldg_stream_f64(vb[i] + (int64_t)k * d12 + j);
        const d4 xa = ld256_gather(x + 4 * (int64_t)a);
#pragma unroll
        for (int i = 0; i < NA; ++i) acc[i] += w[i][0] * xa.x + w[i][1] * xa.y + w[i][2] * xa.z;
    }
    const double* p = x + 4 * n2;
    const int64_t off = 3 * (int64_t)d12;
    for (int32_t j = lane; j < d11; j += 32) {
        const int32_t vp = ldg_stream_s32(col11 + h0 + j);
#pragma unroll
        for (int jj = 0; jj < NA; ++jj) {
            const double pj = __ldg(p + (int64_t)jj * nv + vp);
#pragma unroll
            for (int i = 0; i < NA; ++i) acc[i] += ldg_stream_f64(vb[i] + off + (int64_t)jj * d11 + j) * pj;
        }
    }
#pragma unroll
    for (int i = 0; i < NA; ++i) {
        double r = acc[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
        if (lane == 0) {
            const int64_t idx = 4 * n2 + (int64_t)i * nv + v;
            y[idx] = (mask && mask[idx]) ? x[idx] : r;
        }
    }
}

__global__ void k_to_internal(int64_t n2, int64_t npress, const double* __restrict__ src, double* __restrict__ dst) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= 4 * n2 + npress) return;
    if (i < 4 * n2) {
        int64_t a = i >> 2;
        int k = (int)(i & 3);
        dst[i] = k < 3 ? src[k * n2 + a] : 0.0;
    } else {
        dst[i] = src[i - n2];
    }
}

__global__ void k_to_api(int64_t n2, int64_t n, const double* __restrict__ src, double* __restrict__ dst) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[api_to_internal(i, n2)];
}

template <int NA>
void launch_block(mpet_ctx* ctx, const double* x, double* y, const uint8_t* mask, const int* done, cudaStream_t st) {
    const int th = 256;
    k_spmv_block_u<NA><<<grid_for(ctx->N2 * 32, th), th, 0, st>>>(ctx->N2, ctx->Nv, ctx->g22.rowptr, ctx->g22.col,
                                                                  ctx->g21.rowptr, ctx->g21.col, ctx->rowptr,
                                                                  ctx->vals, x, y, mask, done);
    LAUNCH_CHECK(ctx);
    if (NA > 0) {
        k_spmv_block_p<(NA > 0 ? NA : 1)><<<grid_for(ctx->Nv * 32, th), th, 0, st>>>(
            ctx->N2, ctx->Nv, ctx->g12.rowptr, ctx->g12.col, ctx->g11.rowptr, ctx->g11.col, ctx->rowptr, ctx->vals,
            x, y, mask, done);
        LAUNCH_CHECK(ctx);
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

void block_spmv(mpet_ctx* ctx, const double* x_int, double* y_int, const uint8_t* mask_int, const int* done,
                cudaStream_t st) {
    if (staged_block_spmv(ctx, x_int, y_int, mask_int, done, st)) return;
    switch (ctx->A) {
        case 0: launch_block<0>(ctx, x_int, y_int, mask_int, done, st); break;
        case 1: launch_block<1>(ctx, x_int, y_int, mask_int, done, st); break;
        case 2: launch_block<2>(ctx, x_int, y_int, mask_int, done, st); break;
        case 3: launch_block<3>(ctx, x_int, y_int, mask_int, done, st); break;
        case 4: launch_block<4>(ctx, x_int, y_int, mask_int, done, st); break;
        case 5: launch_block<5>(ctx, x_int, y_int, mask_int, done, st); break;
        case 6: launch_block<6>(ctx, x_int, y_int, mask_int, done, st); break;
        case 7: launch_block<7>(ctx, x_int, y_int, mask_int, done, st); break;
        default: launch_block<8>(ctx, x_int, y_int, mask_int, done, st); break;
    }
}

void to_internal(mpet_ctx* ctx, const double* src_api, double* dst_int, cudaStream_t st) {
    int64_t n = 4 * ctx->N2 + (int64_t)ctx->A * ctx->Nv;
    k_to_internal<<<grid_for(n, 256), 256, 0, st>>>(ctx->N2, (int64_t)ctx->A * ctx->Nv, src_api, dst_int);
    LAUNCH_CHECK(ctx);
}

void to_api(mpet_ctx* ctx, const double* src_int, double* dst_api, cudaStream_t st) {
    k_to_api<<<grid_for(ctx->N, 256), 256, 0, st>>>(ctx->N2, ctx->N, src_int, dst_api);
    LAUNCH_CHECK(ctx);
}
