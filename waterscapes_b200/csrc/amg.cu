// Block-diagonal AMG preconditioner: one V-cycle per field block, applied entirely on device.
//
// Replaces PC "hypre_amg" (one BoomerAMG V-cycle on the assembled block-diagonal matrix P,
// mpetsolver.py:502-507, mpettotalpressuresolver.py:443-448).  P's structure is exploited instead of
// treating it as one big matrix:
//   * displacement: the three components share ONE scalar P2 operator mu*(grad,grad)
//     (mpetsolver.py:268-269) -> one hierarchy applied to 3 right-hand sides at once (the matrix is
//     streamed once per SpMM).  Level 0 -> 1 is p-coarsening P2 -> P1 (P1 is a subspace of P2, so the
//     Galerkin operator is the P1 stiffness matrix, assembled directly by assemble.cu); below that,
//     smoothed aggregation.
//   * each network pressure: its own P1 operator (c_i + dt theta sum S_ij) M + dt theta K_i L
//     (mpetsolver.py:270-271), smoothed aggregation below.
// Smoother: Chebyshev polynomial in D^-1 A (symmetric, no triangular solves: GPU-friendly stand-in
// for BoomerAMG's hybrid Gauss-Seidel), fused with the SpMM that produces its residual.  With equal
// pre/post smoothing and a zero initial guess the V-cycle is a fixed SPD operator, as MINRES needs.
// The aggregation set-up (strength graph, aggregates, smoothed prolongator, Galerkin product) runs
// once per (mesh, dt) on the host in this translation unit on the small P1-sized matrices and is
// uploaded; everything per-iteration is device code.
#include "ctx.h"
#include "layout.cuh"
#include <algorithm>
#include <cmath>
#include <numeric>
#include <cstdlib>

namespace {

// =============================================================================== device kernels
// Vectors of a hierarchy are "W-wide": row r owns W consecutive doubles.  W = 4 for the displacement
// block (three components + zero pad, solver-internal layout of layout.cuh: one 256-bit gather per
// matrix entry serves all three right-hand sides), W = 1 for a pressure block.
enum { EPI_RESID = 0, EPI_CHEB = 1 };

template <int W> struct Acc;
template <> struct Acc<1> {
    double v;
    __device__ __forceinline__ void zero() { v = 0.0; }
    __device__ __forceinline__ void fma_gather(double a, const double* x, int64_t c) { v += a * __ldg(x + c); }
    template <int LANES> __device__ __forceinline__ void reduce() {
#pragma unroll
        for (int o = LANES / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o, LANES);
    }
};
template <> struct Acc<4> {
    d4 v;
    __device__ __forceinline__ void zero() { v.x = v.y = v.z = v.w = 0.0; }
    __device__ __forceinline__ void fma_gather(double a, const double* x, int64_t c) {
        const d4 g = ld256_gather(x + 4 * c);
        v.x += a * g.x; v.y += a * g.y; v.z += a * g.z; v.w += a * g.w;
    }
    template <int LANES> __device__ __forceinline__ void reduce() {
#pragma unroll
        for (int o = LANES / 2; o > 0; o >>= 1) {
            v.x += __shfl_xor_sync(0xffffffffu, v.x, o, LANES);
            v.y += __shfl_xor_sync(0xffffffffu, v.y, o, LANES);
            v.z += __shfl_xor_sync(0xffffffffu, v.z, o, LANES);
            v.w += __shfl_xor_sync(0xffffffffu, v.w, o, LANES);
        }
    }
};

// Fused SpMM + epilogue:
//   EPI_RESID: out = b - A x
//   EPI_CHEB : r = b - A x ; d = c1*d + c2*dinv*r ; out = x + d          (out must not alias x)
template <int LANES, int W, int EPI>
__global__ void __launch_bounds__(256)
k_spmm_epi(int64_t nrows, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ cols,
           const double* __restrict__ vals, const double* __restrict__ x, const double* __restrict__ b,
           double* __restrict__ out, double* __restrict__ d, const double* __restrict__ dinv, double c1,
           double c2, const int* __restrict__ done) {
    if (done && *done) return;
    const int lane = threadIdx.x & (LANES - 1);
    int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / LANES;
    if (row >= nrows) return;
    const int32_t s = rowptr[row], e = rowptr[row + 1];
    // the row-wise operands of the epilogue do not depend on the SpMM: issue their loads first so that
    // they travel together with the matrix stream instead of after the reduction
    double eb1 = 0, ed1 = 0, ex1 = 0, edi = 0;
    d4 eb4 = {0, 0, 0, 0}, ed4 = {0, 0, 0, 0}, ex4 = {0, 0, 0, 0};
    if (lane == 0) {
        if constexpr (W == 1) {
            eb1 = b[row];
            if (EPI == EPI_CHEB) { edi = dinv[row]; ex1 = x[row]; if (c1 != 0.0) ed1 = d[row]; }
        } else {
            eb4 = ld256(b + 4 * row);
            if (EPI == EPI_CHEB) { edi = dinv[row]; ex4 = ld256(x + 4 * row); if (c1 != 0.0) ed4 = ld256(d + 4 * row); }
        }
    }
    Acc<W> acc;
    acc.zero();
    for (int32_t i = s + lane; i < e; i += LANES) {
        const double v = ldg_stream_f64(vals + i);
        const int32_t c = ldg_stream_s32(cols + i);
        acc.fma_gather(v, x, c);
    }
    acc.template reduce<LANES>();
    if (lane != 0) return;
    if constexpr (W == 1) {
        const double res = eb1 - acc.v;
        if (EPI == EPI_RESID) {
            out[row] = res;
        } else {
            double dn = c2 * edi * res + c1 * ed1;
            d[row] = dn;
            out[row] = ex1 + dn;
        }
    } else {
        const d4 a = acc.v;
        d4 res = {eb4.x - a.x, eb4.y - a.y, eb4.z - a.z, eb4.w - a.w};
        if (EPI == EPI_RESID) {
            st256(out + 4 * row, res);
        } else {
            const double s2 = c2 * edi;
            d4 dn = {s2 * res.x + c1 * ed4.x, s2 * res.y + c1 * ed4.y, s2 * res.z + c1 * ed4.z, s2 * res.w + c1 * ed4.w};
            st256(d + 4 * row, dn);
            d4 o = {ex4.x + dn.x, ex4.y + dn.y, ex4.z + dn.z, ex4.w + dn.w};
            st256(out + 4 * row, o);
        }
    }
}

// first Chebyshev step from a zero guess: d = c2 * dinv * b ; x = d
template <int W>
__global__ void k_cheb_first(int64_t n, const double* __restrict__ b, const double* __restrict__ dinv,
                             double c2, double* __restrict__ d, double* __restrict__ x,
                             const int* __restrict__ done) {
    if (done && *done) return;
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n * W) return;
    double v = c2 * dinv[i / W] * b[i];
    d[i] = v;
    x[i] = v;
}

template <int W>
__global__ void k_dense_apply(int n, const double* __restrict__ inv, const double* __restrict__ b,
                              double* __restrict__ x, const int* __restrict__ done) {
    if (done && *done) return;
    int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (row >= n) return;
    double sum[W];
#pragma unroll
    for (int k = 0; k < W; ++k) sum[k] = 0.0;
    for (int j = lane; j < n; j += 32) {
        double a = inv[(int64_t)row * n + j];
#pragma unroll
        for (int k = 0; k < W; ++k) sum[k] += a * b[(int64_t)j * W + k];
    }
#pragma unroll
    for (int k = 0; k < W; ++k) {
        double r = sum[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
        if (lane == 0) x[(int64_t)row * W + k] = r;
    }
}

// y = M x (+ y) on W-wide vectors (restriction / prolongation / power iteration)
template <int LANES, int W>
__global__ void __launch_bounds__(256)
k_spmm_plain(int64_t nrows, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ cols,
             const double* __restrict__ vals, const double* __restrict__ x, double* __restrict__ y,
             double beta, const int* __restrict__ done) {
    if (done && *done) return;
    const int lane = threadIdx.x & (LANES - 1);
    int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / LANES;
    if (row >= nrows) return;
    const int32_t s = rowptr[row], e = rowptr[row + 1];
    Acc<W> acc;
    acc.zero();
    for (int32_t i = s + lane; i < e; i += LANES) acc.fma_gather(vals[i], x, cols[i]);
    acc.template reduce<LANES>();
    if (lane != 0) return;
    if constexpr (W == 1) {
        double r = acc.v;
        y[row] = (beta == 0.0) ? r : r + beta * y[row];
    } else {
        d4 r = acc.v;
        if (beta != 0.0) {
            const d4 yy = ld256(y + 4 * row);
            r.x += beta * yy.x; r.y += beta * yy.y; r.z += beta * yy.z; r.w += beta * yy.w;
        }
        st256(y + 4 * row, r);
    }
}

// scalar-block matrix with symmetric Dirichlet elimination: out = scale*in, rows/cols of masked
// nodes zeroed, unit diagonal (apply_symmetric(bc, P), mpetsolver.py:503-504)
__global__ void __launch_bounds__(256)
k_masked_copy(int64_t nrows, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ cols,
              const double* __restrict__ in, double scale, const uint8_t* __restrict__ mask,
              double* __restrict__ out, double* __restrict__ dinv, const double* __restrict__ in2, double scale2) {
    int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 3;
    int lane = threadIdx.x & 7;
    if (row >= nrows) return;
    bool rd = mask && mask[row];
    for (int32_t t = rowptr[row] + lane; t < rowptr[row + 1]; t += 8) {
        int32_t c = cols[t];
        double v = scale * in[t];
        if (in2) v += scale2 * in2[t];
        if (rd) v = (c == row) ? 1.0 : 0.0;
        else if (mask && mask[c]) v = 0.0;
        out[t] = v;
        if (c == row) dinv[row] = 1.0 / v;
    }
}

// pattern of diag(B_0, .., B_{A-1}) for A blocks that share the vertex graph g11
__global__ void k_blockdiag_pattern(const int32_t* __restrict__ rp11, const int32_t* __restrict__ col11, int64_t nv,
                                    int64_t nnz11, int A, int32_t* __restrict__ rp, int32_t* __restrict__ col) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t <= (int64_t)A * nv) {
        const int64_t i = t / nv, v = t - i * nv;
        rp[t] = (int32_t)((i < A) ? i * nnz11 + rp11[v] : (int64_t)A * nnz11);
    }
    if (t < (int64_t)A * nnz11) {
        const int64_t i = t / nnz11, e = t - i * nnz11;
        col[t] = (int32_t)(i * nv + col11[e]);
    }
}

// ---- Lanczos in the D inner product (setup only): deterministic single-block reductions
__global__ void __launch_bounds__(1024)
k_lz_dot(int64_t n, const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ dinv,
         const uint8_t* __restrict__ own, double* __restrict__ out) {
    __shared__ double sh[1024];
    double s = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += 1024)
        if (!own || own[i]) s += x[i] * y[i] / dinv[i];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = 512; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = sh[0];
}
__global__ void k_lz_start(int64_t n, double* __restrict__ v, double* __restrict__ vprev) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    v[i] = 1.0 + 0.5 * sin(0.37 * (double)i + 0.1 * (double)(i % 7));
    vprev[i] = 0.0;
}
__global__ void k_lz_scale(int64_t n, double* __restrict__ v, double s) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) v[i] *= s;
}
// w = dinv * w - alpha * v - beta * vprev
__global__ void k_lz_update(int64_t n, double* __restrict__ w, const double* __restrict__ dinv, const double* __restrict__ v,
                            const double* __restrict__ vprev, double alpha, double beta, int scale_first) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    double t = scale_first ? dinv[i] * w[i] : w[i];
    w[i] = t - alpha * v[i] - beta * vprev[i];
}
// vprev = v ; v = w / beta
__global__ void k_lz_next(int64_t n, const double* __restrict__ w, double inv_beta, double* __restrict__ v,
                          double* __restrict__ vprev) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    vprev[i] = v[i];
    v[i] = w[i] * inv_beta;
}

__global__ void k_expand4(int64_t n, const double* __restrict__ in, double* __restrict__ out) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < 4 * n) out[i] = in[i >> 2];
}
__global__ void k_take4(int64_t n, const double* __restrict__ in, double* __restrict__ out) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[4 * i];
}

__global__ void k_add_inplace(int64_t n, const double* __restrict__ e, double* __restrict__ x, const int* __restrict__ done) {
    if (done && *done) return;
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) x[i] += e[i];
}

__global__ void k_gershgorin(int64_t nrows, const int32_t* __restrict__ rowptr, const double* __restrict__ vals,
                             const double* __restrict__ dinv, double* __restrict__ out) {
    int64_t row = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (row >= nrows) return;
    double s = 0;
    for (int32_t t = rowptr[row]; t < rowptr[row + 1]; ++t) s += fabs(vals[t]);
    out[row] = s * fabs(dinv[row]);
}

__global__ void k_scale_by(int64_t n, const double* __restrict__ dinv, double* __restrict__ y) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) y[i] *= dinv[i];
}

// =============================================================================== host CSR tools
struct HostCsr {
    int64_t nrows = 0, ncols = 0;
    std::vector<int32_t> rp, ci;
    std::vector<double> v;
    int64_t nnz() const { return (int64_t)ci.size(); }
};

HostCsr download(const DevCsr& M) {
    HostCsr H;
    H.nrows = M.nrows; H.ncols = M.ncols;
    H.rp.resize(M.nrows + 1); H.ci.resize(M.nnz); H.v.resize(M.nnz);
    CUDA_CHECK(cudaMemcpy(H.rp.data(), M.rowptr, sizeof(int32_t) * (M.nrows + 1), cudaMemcpyDeviceToHost));
    CUDA_CHECK(cudaMemcpy(H.ci.data(), M.col, sizeof(int32_t) * M.nnz, cudaMemcpyDeviceToHost));
    CUDA_CHECK(cudaMemcpy(H.v.data(), M.val, sizeof(double) * M.nnz, cudaMemcpyDeviceToHost));
    return H;
}

DevCsr upload(mpet_ctx* ctx, const HostCsr& H) {
    DevCsr M;
    M.nrows = H.nrows; M.ncols = H.ncols; M.nnz = H.nnz();
    M.rowptr = dev_alloc<int32_t>(ctx, H.nrows + 1);
    M.col = dev_alloc<int32_t>(ctx, M.nnz);
    M.val = dev_alloc<double>(ctx, M.nnz);
    CUDA_CHECK(cudaMemcpy(M.rowptr, H.rp.data(), sizeof(int32_t) * (H.nrows + 1), cudaMemcpyHostToDevice));
    if (M.nnz) {
        CUDA_CHECK(cudaMemcpy(M.col, H.ci.data(), sizeof(int32_t) * M.nnz, cudaMemcpyHostToDevice));
        CUDA_CHECK(cudaMemcpy(M.val, H.v.data(), sizeof(double) * M.nnz, cudaMemcpyHostToDevice));
    }
    return M;
}

HostCsr transpose(const HostCsr& A) {
    HostCsr T;
    T.nrows = A.ncols; T.ncols = A.nrows;
    T.rp.assign(A.ncols + 1, 0);
    for (int32_t c : A.ci) T.rp[c + 1]++;
    for (int64_t i = 0; i < A.ncols; ++i) T.rp[i + 1] += T.rp[i];
    T.ci.resize(A.nnz()); T.v.resize(A.nnz());
    std::vector<int32_t> pos(T.rp.begin(), T.rp.end() - 1);
    for (int64_t r = 0; r < A.nrows; ++r)
        for (int32_t t = A.rp[r]; t < A.rp[r + 1]; ++t) {
            int32_t p = pos[A.ci[t]]++;
            T.ci[p] = (int32_t)r;
            T.v[p] = A.v[t];
        }
    return T;
}

// C = A * B (Gustavson, dense accumulator), columns sorted per row
HostCsr spgemm(const HostCsr& A, const HostCsr& B) {
    HostCsr Cm;
    Cm.nrows = A.nrows; Cm.ncols = B.ncols;
    Cm.rp.assign(A.nrows + 1, 0);
    std::vector<double> acc(B.ncols, 0.0);
    std::vector<int32_t> marker(B.ncols, -1), list;
    for (int64_t r = 0; r < A.nrows; ++r) {
        list.clear();
        for (int32_t t = A.rp[r]; t < A.rp[r + 1]; ++t) {
            int32_t k = A.ci[t];
            double av = A.v[t];
            for (int32_t u = B.rp[k]; u < B.rp[k + 1]; ++u) {
                int32_t c = B.ci[u];
                if (marker[c] != (int32_t)r) { marker[c] = (int32_t)r; list.push_back(c); acc[c] = 0.0; }
                acc[c] += av * B.v[u];
            }
        }
        std::sort(list.begin(), list.end());
        for (int32_t c : list) { Cm.ci.push_back(c); Cm.v.push_back(acc[c]); }
        Cm.rp[r + 1] = (int32_t)Cm.ci.size();
    }
    return Cm;
}

// Greedy aggregation (three passes) on the strength graph |a_ij| >= theta sqrt(a_ii a_jj).
// agg[i] = aggregate of row i or -1 for identity (eliminated Dirichlet) rows.
void aggregate(const HostCsr& A, double theta, std::vector<int32_t>& agg, int64_t& nagg, std::vector<char>& isolated,
               std::vector<double>& diag) {
    const int64_t n = A.nrows;
    diag.assign(n, 1.0);
    for (int64_t r = 0; r < n; ++r)
        for (int32_t t = A.rp[r]; t < A.rp[r + 1]; ++t)
            if (A.ci[t] == r) diag[r] = A.v[t];
    std::vector<int32_t> srp(n + 1, 0), sci;
    std::vector<double> sval;
    isolated.assign(n, 0);
    for (int64_t r = 0; r < n; ++r) {
        for (int32_t t = A.rp[r]; t < A.rp[r + 1]; ++t) {
            int32_t c = A.ci[t];
            if (c == r || A.v[t] == 0.0) continue;
            if (std::fabs(A.v[t]) >= theta * std::sqrt(std::fabs(diag[r] * diag[c]))) {
                sci.push_back(c);
                sval.push_back(std::fabs(A.v[t]) / std::sqrt(std::fabs(diag[r] * diag[c])));
            }
        }
        srp[r + 1] = (int32_t)sci.size();
        bool any_off = false;
        for (int32_t t = A.rp[r]; t < A.rp[r + 1]; ++t)
            if (A.ci[t] != r && A.v[t] != 0.0) { any_off = true; break; }
        isolated[r] = any_off ? 0 : 1;        // identity rows (eliminated Dirichlet dofs) stay out
    }
    agg.assign(n, -1);
    nagg = 0;
    for (int64_t r = 0; r < n; ++r) {          // pass 1: seeds whose whole strong neighbourhood is free
        if (agg[r] != -1 || isolated[r]) continue;
        bool free_nb = true;
        for (int32_t t = srp[r]; t < srp[r + 1]; ++t)
            if (agg[sci[t]] != -1) { free_nb = false; break; }
        if (!free_nb || srp[r + 1] == srp[r]) continue;
        agg[r] = (int32_t)nagg;
        for (int32_t t = srp[r]; t < srp[r + 1]; ++t) agg[sci[t]] = (int32_t)nagg;
        nagg++;
    }
    std::vector<int32_t> agg2(agg);
    for (int64_t r = 0; r < n; ++r) {          // pass 2: attach leftovers to the strongest neighbouring aggregate
        if (agg[r] != -1 || isolated[r]) continue;
        double best = -1; int32_t who = -1;
        for (int32_t t = srp[r]; t < srp[r + 1]; ++t)
            if (agg[sci[t]] != -1 && sval[t] > best) { best = sval[t]; who = agg[sci[t]]; }
        if (who >= 0) agg2[r] = who;
    }
    agg.swap(agg2);
    for (int64_t r = 0; r < n; ++r) {          // pass 3: the rest forms new aggregates
        if (agg[r] != -1 || isolated[r]) continue;
        agg[r] = (int32_t)nagg;
        for (int32_t t = srp[r]; t < srp[r + 1]; ++t)
            if (agg[sci[t]] == -1 && !isolated[sci[t]]) agg[sci[t]] = (int32_t)nagg;
        nagg++;
    }
}

// Smoothed-aggregation prolongator for an SPD matrix with possible identity (Dirichlet) rows.
// Returns P (nrows x naggregates); naggregates == 0 means "cannot coarsen".
HostCsr sa_prolongator(const HostCsr& A, double theta, int64_t& nagg) {
    const int64_t n = A.nrows;
    std::vector<int32_t> agg;
    std::vector<char> isolated;
    std::vector<double> diag;
    aggregate(A, theta, agg, nagg, isolated, diag);
    HostCsr P;
    if (nagg == 0) return P;
    // tentative prolongator (piecewise constant), then one damped-Jacobi smoothing step
    HostCsr T;
    T.nrows = n; T.ncols = nagg;
    T.rp.assign(n + 1, 0);
    for (int64_t r = 0; r < n; ++r) {
        if (agg[r] >= 0) { T.ci.push_back(agg[r]); T.v.push_back(1.0); }
        T.rp[r + 1] = (int32_t)T.ci.size();
    }
    // rho(D^-1 A) by power iteration
    std::vector<double> x(n), y(n);
    for (int64_t i = 0; i < n; ++i) x[i] = 1.0 + 0.37 * std::sin(1.7 * (double)i);
    double rho = 1.0;
    for (int it = 0; it < 15; ++it) {
        double nx = 0;
        for (int64_t i = 0; i < n; ++i) nx += x[i] * x[i];
        nx = std::sqrt(nx);
        for (int64_t i = 0; i < n; ++i) x[i] /= nx;
        for (int64_t r = 0; r < n; ++r) {
            double s = 0;
            for (int32_t t = A.rp[r]; t < A.rp[r + 1]; ++t) s += A.v[t] * x[A.ci[t]];
            y[r] = s / diag[r];
        }
        double ny = 0;
        for (int64_t i = 0; i < n; ++i) ny += y[i] * y[i];
        rho = std::sqrt(ny);
        x.swap(y);
    }
    rho *= 1.1;
    const double omega = (4.0 / 3.0) / rho;
    // P = (I - omega D^-1 A) T
    HostCsr AT = spgemm(A, T);
    P.nrows = n; P.ncols = nagg;
    P.rp.assign(n + 1, 0);
    std::vector<std::pair<int32_t, double>> rowbuf;
    for (int64_t r = 0; r < n; ++r) {
        if (!isolated[r]) {
            rowbuf.clear();
            for (int32_t t = AT.rp[r]; t < AT.rp[r + 1]; ++t)
                rowbuf.emplace_back(AT.ci[t], -omega / diag[r] * AT.v[t]);
            if (agg[r] >= 0) rowbuf.emplace_back(agg[r], 1.0);
            std::sort(rowbuf.begin(), rowbuf.end(),
                      [](const std::pair<int32_t, double>& a, const std::pair<int32_t, double>& b) { return a.first < b.first; });
            for (size_t q = 0; q < rowbuf.size(); ++q) {
                if (!P.ci.empty() && (int32_t)P.ci.size() > P.rp[r] && P.ci.back() == rowbuf[q].first)
                    P.v.back() += rowbuf[q].second;
                else { P.ci.push_back(rowbuf[q].first); P.v.push_back(rowbuf[q].second); }
            }
        }
        P.rp[r + 1] = (int32_t)P.ci.size();
    }
    return P;
}

std::vector<double> dense_inverse(const HostCsr& A) {
    const int n = (int)A.nrows;
    std::vector<double> M((size_t)n * n, 0.0), I((size_t)n * n, 0.0);
    double tr = 0;
    for (int r = 0; r < n; ++r) {
        for (int32_t t = A.rp[r]; t < A.rp[r + 1]; ++t) M[(size_t)r * n + A.ci[t]] += A.v[t];
        I[(size_t)r * n + r] = 1.0;
        tr += std::fabs(M[(size_t)r * n + r]);
    }
    const double tiny = 1e-13 * (tr / std::max(1, n));
    for (int k = 0; k < n; ++k) {   // Gauss-Jordan with partial pivoting
        int piv = k;
        for (int r = k + 1; r < n; ++r)
            if (std::fabs(M[(size_t)r * n + k]) > std::fabs(M[(size_t)piv * n + k])) piv = r;
        if (piv != k)
            for (int c = 0; c < n; ++c) {
                std::swap(M[(size_t)k * n + c], M[(size_t)piv * n + c]);
                std::swap(I[(size_t)k * n + c], I[(size_t)piv * n + c]);
            }
        double p = M[(size_t)k * n + k];
        if (std::fabs(p) < tiny) p = (p < 0 ? -tiny : tiny);   // singular block (pure Neumann): regularise
        double ip = 1.0 / p;
        for (int c = 0; c < n; ++c) { M[(size_t)k * n + c] *= ip; I[(size_t)k * n + c] *= ip; }
        for (int r = 0; r < n; ++r) {
            if (r == k) continue;
            double f = M[(size_t)r * n + k];
            if (f == 0.0) continue;
            for (int c = 0; c < n; ++c) {
                M[(size_t)r * n + c] -= f * M[(size_t)k * n + c];
                I[(size_t)r * n + c] -= f * I[(size_t)k * n + c];
            }
        }
    }
    return I;
}

// lump off-diagonal entries with |a_ij| < tol * sqrt(|a_ii a_jj|) into the diagonal (symmetric criterion, row
// sums preserved); oracle/krylov.py:drop_small is the same pass
HostCsr drop_small(const HostCsr& A, double tol) {
    if (tol <= 0) return A;
    const int64_t n = A.nrows;
    std::vector<double> d(n, 0.0);
    for (int64_t r = 0; r < n; ++r)
        for (int32_t t = A.rp[r]; t < A.rp[r + 1]; ++t)
            if (A.ci[t] == r) d[r] = A.v[t];
    HostCsr B;
    B.nrows = A.nrows; B.ncols = A.ncols;
    B.rp.assign(n + 1, 0);
    B.ci.reserve(A.ci.size() / 2);
    B.v.reserve(A.v.size() / 2);
    for (int64_t r = 0; r < n; ++r) {
        double lump = 0.0;
        int64_t dpos = -1;
        for (int32_t t = A.rp[r]; t < A.rp[r + 1]; ++t) {
            const int32_t c = A.ci[t];
            if (c != r && std::fabs(A.v[t]) < tol * std::sqrt(std::fabs(d[r] * d[c]))) { lump += A.v[t]; continue; }
            if (c == r) dpos = (int64_t)B.ci.size();
            B.ci.push_back(c);
            B.v.push_back(A.v[t]);
        }
        if (dpos >= 0) B.v[dpos] += lump;
        B.rp[r + 1] = (int32_t)B.ci.size();
    }
    return B;
}

// =============================================================================== hierarchy
const int kChebDegree = 2;
// well-conditioned blocks (cond(D^-1 A) <= kPolyKappaMax, e.g. mass-dominated network blocks) are inverted by a
// Chebyshev polynomial on the WHOLE spectrum instead of a V-cycle (oracle/krylov.py: same constants)
// Threshold: a degree-k polynomial costs k passes over the P1 matrix and no coarse levels / all-gathers; up to
// cond ~ 32 (degree <= 28 for the 1e-4 target) that is no more than the 2-3 strong cycles of degree 4 it replaces.
// It was 12 until the 8-GPU weak-scaling run (144^3 cubes): h had shrunk enough for the second network of cfg5 to
// cross 12 (spectrum [0.136, 1.91]), its block fell back to ONE light V-cycle and MINRES needed 154-181 iterations
// instead of 58-67 (profiles/r02_multi_gpu.md).
const double kPolyKappaMax = 32.0;
const double kPolyTarget = 1.0e-4;
const double kPolyTargetLight = 1.0e-4;    // light mode (loose tolerances): residual-polynomial target of the polynomial fields
const int kPolyMaxDegree = 28;
const int kLanczosSteps = 40;
// Galerkin operators of the aggregated levels: entries below kDropTol * sqrt(a_ii a_jj) are lumped into the diagonal.
// Smoothed aggregation fills in quickly (cfg5: 851 entries per row two levels below the mesh, more entries than the
// mesh level itself), and on multi-GPU these levels are replicated on every rank.  Oracle experiments (cfg5
// family at the bench's mesh width): 0.01 leaves the MINRES iteration count unchanged (343/339/347 vs 344/-/346
// at n = 16/20/24) with 2.7x fewer entries on the first aggregated level and 10x below; 0.03 is already fragile.
const double kDropTol = 0.01;
// P1-field blocks that keep a hierarchy are applied as kP1Cycles stationary cycles with degree-kP1Degree Chebyshev
// smoothing.  MINRES on the MPET system is very sensitive to the accuracy of these (cheap) blocks; measured on
// cfg5 (iterations / ms per step): 1 cycle, degree 2: 549 / 2892; degree 4: 469 / 2522; 2 cycles: 406 / 2214;
// 3 cycles: 350 / 1989; 2 cycles, degree 4: 341 / 1942; 4 cycles: 340 / 1995 (plateau = exact block solves).
const int kP1Cycles = 2;
const int kP1Degree = 4;
const int kP1DegreeCoarse = 4;
const int kP1DegreeLight = 2;
// multi-GPU: the transition to the replicated levels uses plain aggregation (weaker coarse correction); measured
// at N = 2 (91^3): 2 cycles 541-600 iterations / 4417 ms, 3 cycles 473 / 3981, 4 cycles 462 / 4241
const int kP1CyclesDist = 3;
const double kChebRatio = 4.0;    // smooth the upper [lambda_max / ratio, lambda_max] of D^-1 A
const int64_t kCoarseMax = 300;
const int kMaxLevels = 12;

// Algorithmic bytes of one pass (bench.py's preconditioner roofline): the matrix stream once (8-byte value +
// 4-byte column index per entry, row pointers), every gathered vector entry once, the row-wise operands once.
inline int64_t spmm_bytes(const DevCsr& M, int W, int row_streams) {
    return 12 * M.nnz + 4 * (M.nrows + 1) + 8 * (int64_t)W * M.ncols + 8 * (int64_t)W * M.nrows * row_streams;
}

template <int W, int EPI>
void launch_epi(mpet_ctx* ctx, const DevCsr& M, const SpmmPlan& plan, const double* x, const double* b, double* out,
                double* d, const double* dinv, double c1, double c2, const int* done, cudaStream_t st) {
    // RESID: read b, write out.  CHEB: read b, (d), dinv; write d, out
    ctx->pc_bytes_acc += EPI == EPI_RESID ? spmm_bytes(M, W, 2) : spmm_bytes(M, W, 3 + (c1 != 0.0)) + 8 * M.nrows;
    if (plan.nchunks > 0) {
        staged_spmm(ctx, W, EPI, plan, M, x, b, out, d, dinv, c1, c2, done, st);
        return;
    }
    double mean = M.nrows > 0 ? (double)M.nnz / (double)M.nrows : 0.0;
    const int th = 256;
    if (mean > 48)
        k_spmm_epi<32, W, EPI><<<grid_for(M.nrows * 32, th), th, 0, st>>>(M.nrows, M.rowptr, M.col, M.val, x, b, out, d, dinv, c1, c2, done);
    else if (mean > 20)
        k_spmm_epi<16, W, EPI><<<grid_for(M.nrows * 16, th), th, 0, st>>>(M.nrows, M.rowptr, M.col, M.val, x, b, out, d, dinv, c1, c2, done);
    else
        k_spmm_epi<8, W, EPI><<<grid_for(M.nrows * 8, th), th, 0, st>>>(M.nrows, M.rowptr, M.col, M.val, x, b, out, d, dinv, c1, c2, done);
    LAUNCH_CHECK(ctx);
}

template <int W>
void launch_plain(mpet_ctx* ctx, const DevCsr& M, const SpmmPlan& plan, const double* x, double* y, double beta,
                  const int* done, cudaStream_t st) {
    ctx->pc_bytes_acc += spmm_bytes(M, W, 1 + (beta != 0.0));
    if (plan.nchunks > 0) {
        staged_spmm(ctx, W, 2, plan, M, x, nullptr, y, nullptr, nullptr, beta, 0.0, done, st);
        return;
    }
    double mean = M.nrows > 0 ? (double)M.nnz / (double)M.nrows : 0.0;
    const int th = 256;
    if (mean > 10)
        k_spmm_plain<16, W><<<grid_for(M.nrows * 16, th), th, 0, st>>>(M.nrows, M.rowptr, M.col, M.val, x, y, beta, done);
    else if (mean > 3)
        k_spmm_plain<8, W><<<grid_for(M.nrows * 8, th), th, 0, st>>>(M.nrows, M.rowptr, M.col, M.val, x, y, beta, done);
    else
        k_spmm_plain<2, W><<<grid_for(M.nrows * 2, th), th, 0, st>>>(M.nrows, M.rowptr, M.col, M.val, x, y, beta, done);
    LAUNCH_CHECK(ctx);
}

// Chebyshev smoothing on level L: x_out = S(b, x_in); x_in == nullptr means zero initial guess.
// development aid: MPET_ALL_HALOS=1 restores the exchange after every level operation
static bool skip_redundant_halos() {
    static const bool v = []() { const char* e = getenv("MPET_ALL_HALOS"); return !(e && e[0] == '1'); }();
    return v;
}

// "Light" P1-field cycles (one cycle, degree-2 smoothing) for loose tolerances: MINRES' first ~100 iterations are
// insensitive to how accurately the (cheap, latency-bound) P1 blocks are inverted -- measured on cfg5 at the bench
// tolerance: 62/68/71/71 iterations with 1 cycle of degree 2 exactly as with 2 or 3 cycles of degree 4, 5.5 % less time
// per step on one GPU and ~2 ms less per iteration on two --, while tight tolerances (the LU stand-in) need the strong
// cycles (r01: 549 iterations with 1 cycle of degree 2, 341 with 2 cycles of degree 4).  krylov.cu picks the mode
// from the relative tolerance; oracle/krylov.py BlockAMG(light=...) is the same rule.
static int g_p_degree(const mpet_ctx* ctx) {
    static const int d = []() { const char* e = getenv("MPET_P_DEGREE"); return e ? std::max(1, atoi(e)) : 0; }();
    return d ? d : (ctx->pc_light ? kP1DegreeLight : kP1Degree);
}
// displacement cycle: degree-2 Chebyshev smoothing, degree 1 (damped Jacobi) in light mode -- measured on cfg5 at the
// bench tolerance: 63/69/72/73 iterations instead of 62/68/71/71 with HALF the P2-level passes (2 instead of 4 per
// application): 288 instead of 355 ms per step
static int g_u_degree(const mpet_ctx* ctx) {
    static const int d = []() { const char* e = getenv("MPET_U_DEGREE"); return e ? std::max(1, atoi(e)) : 0; }();
    return d ? d : (ctx->pc_light ? 1 : kChebDegree);
}
// smoothing degree on the aggregated (non-mesh) levels of the P1-field hierarchies
static int g_p_degree_coarse(const mpet_ctx* ctx) {
    static const int d = []() { const char* e = getenv("MPET_P_DEGREE_COARSE"); return e ? std::max(1, atoi(e)) : 0; }();
    return d ? d : (ctx->pc_light ? kP1DegreeLight : kP1DegreeCoarse);
}

template <int W>
void chebyshev(mpet_ctx* ctx, AmgLevel& L, const double* b, const double* x_in, double* x_out, const int* done,
               cudaStream_t st, double lo = 0.0, double hi = 0.0, int degree = 0) {
    const int64_t n = L.A.nrows;
    const double lmax = degree > 0 ? hi : L.lambda_max, lmin = degree > 0 ? lo : lmax / kChebRatio;
    const double theta = 0.5 * (lmax + lmin), delta = 0.5 * (lmax - lmin);
    const double sigma = theta / delta;
    double rho_old = 1.0 / sigma;
    double* buf[2] = {L.x, L.t};
    const double* cur = x_in;
    const int steps = degree > 0 ? degree : (W == 1 ? (L.is_mesh_level ? g_p_degree(ctx) : g_p_degree_coarse(ctx)) : g_u_degree(ctx));
    for (int k = 0; k < steps; ++k) {
        bool last = (k == steps - 1);
        double* dst = last ? x_out : buf[k & 1];
        bool bounce = last && cur == x_out;    // the SpMM cannot run in place
        if (bounce) dst = buf[k & 1];
        if (k == 0) {
            if (cur == nullptr) {
                k_cheb_first<W><<<grid_for(n * W, 256), 256, 0, st>>>(n, b, L.dinv, 1.0 / theta, L.r, dst, done);
                LAUNCH_CHECK(ctx);
                ctx->pc_bytes_acc += 8 * n * (3 * W + 1);
            } else {
                launch_epi<W, EPI_CHEB>(ctx, L.A, L.planA, cur, b, dst, L.r, L.dinv, 0.0, 1.0 / theta, done, st);
            }
        } else {
            double rho = 1.0 / (2.0 * sigma - rho_old);
            launch_epi<W, EPI_CHEB>(ctx, L.A, L.planA, cur, b, dst, L.r, L.dinv, rho * rho_old, 2.0 * rho / delta, done, st);
            rho_old = rho;
        }
        if (bounce)
            CUDA_CHECK(cudaMemcpyAsync(x_out, dst, sizeof(double) * n * W, cudaMemcpyDeviceToDevice, st));
        // distributed level: owned rows are right, ghost rows are refreshed from their owners.  The first step
        // from a zero guess is diagonal (x = c * dinv * b): its ghost rows are already right because b's ghosts
        // are valid and dinv's ghosts were taken from their owners at set-up (sync_ghost_diagonal)
        const bool diagonal_step = (k == 0 && x_in == nullptr && skip_redundant_halos());
        if (L.halo_plan >= 0 && !diagonal_step) dist_halo(ctx, L.halo_plan, bounce ? x_out : dst, false, done, st);
        cur = dst;
    }
}

template <int W>
void vcycle(mpet_ctx* ctx, AmgHierarchy& H, int lev, const double* b, double* x, const int* done, cudaStream_t st) {
    AmgLevel& L = H.levels[lev];
    const int64_t n = L.A.nrows;
    if (H.poly_degree > 0) {       // polynomial mode: no hierarchy
        chebyshev<W>(ctx, L, b, nullptr, x, done, st, H.poly_lo, H.poly_hi,
                     ctx->pc_light ? H.poly_degree_light : H.poly_degree);
        return;
    }
    if (lev == (int)H.levels.size() - 1) {
        if (H.coarse_inv) {
            k_dense_apply<W><<<grid_for((int64_t)n * 32, 256), 256, 0, st>>>((int)n, H.coarse_inv, b, x, done);
            LAUNCH_CHECK(ctx);
            ctx->pc_bytes_acc += 8 * n * n + 16 * n * W;
        } else {
            chebyshev<W>(ctx, L, b, nullptr, x, done, st);
        }
        return;
    }
    AmgLevel& C = H.levels[lev + 1];
    const int64_t nc = C.A.nrows;
    chebyshev<W>(ctx, L, b, nullptr, x, done, st);                                            // pre-smooth
    launch_epi<W, EPI_RESID>(ctx, L.A, L.planA, x, b, L.t, nullptr, nullptr, 0, 0, done, st);  // residual
    if (L.halo_plan >= 0) dist_halo(ctx, L.halo_plan, L.t, false, done, st);
    if (lev == H.transition) {
        // restrict onto this rank's aggregates, all-gather the padded slots: coarser levels are replicated
        launch_plain<W>(ctx, C.R, C.planR, L.t, H.gather_send, 0.0, done, st);
        dist_allgather(ctx, H.gather_send, C.b, H.gather_count * W, done, st);
    } else {
        launch_plain<W>(ctx, C.R, C.planR, L.t, C.b, 0.0, done, st);                           // restrict
        if (C.halo_plan >= 0) dist_halo(ctx, C.halo_plan, C.b, false, done, st);
    }
    double* xc = C.x + (int64_t)W * nc;   // second half of the coarse x buffer holds the coarse solution
    vcycle<W>(ctx, H, lev + 1, C.b, xc, done, st);
    launch_plain<W>(ctx, C.P, C.planP, xc, x, 1.0, done, st);                                  // prolong + correct
    // the prolongators carry the rows of the ghost nodes too and the coarse solution is valid on ghosts / replicated:
    // x stays consistent without an exchange
    if (L.halo_plan >= 0 && !skip_redundant_halos()) dist_halo(ctx, L.halo_plan, x, false, done, st);
    chebyshev<W>(ctx, L, b, x, x, done, st);                                                   // post-smooth
}

void alloc_level_work(mpet_ctx* ctx, AmgLevel& L, int nrhs) {
    int64_t n = L.A.nrows;
    L.x = dev_alloc<double>(ctx, 2 * n * nrhs);   // [0,nrhs*n): ping buffer, [nrhs*n, 2*nrhs*n): coarse solution
    L.t = dev_alloc<double>(ctx, n * nrhs);
    L.r = dev_alloc<double>(ctx, n * nrhs);       // Chebyshev direction d
    L.b = dev_alloc<double>(ctx, n * nrhs);
}

double estimate_lambda_max(mpet_ctx* ctx, AmgLevel& L, cudaStream_t st) {
    // Gershgorin bound on D^-1 A caps a power-iteration estimate
    const int64_t n = L.A.nrows;
    double* tmp = nullptr;
    CUDA_CHECK(cudaMalloc(&tmp, sizeof(double) * n));
    k_gershgorin<<<grid_for(n, 256), 256, 0, st>>>(n, L.A.rowptr, L.A.val, L.dinv, tmp);
    LAUNCH_CHECK(ctx);
    std::vector<double> h(n);
    CUDA_CHECK(cudaMemcpyAsync(h.data(), tmp, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
    double gersh = 0;
    for (double v : h) gersh = std::max(gersh, v);
    // power iteration on device
    std::vector<double> x0(n);
    for (int64_t i = 0; i < n; ++i) x0[i] = 1.0 + 0.5 * std::sin(0.37 * (double)i + 0.1 * (double)(i % 7));
    double* x = nullptr; double* y = nullptr;
    CUDA_CHECK(cudaMalloc(&x, sizeof(double) * n));
    CUDA_CHECK(cudaMalloc(&y, sizeof(double) * n));
    CUDA_CHECK(cudaMemcpy(x, x0.data(), sizeof(double) * n, cudaMemcpyHostToDevice));
    double lam = 0;
    for (int it = 0; it < 20; ++it) {
        launch_plain<1>(ctx, L.A, SpmmPlan(), x, y, 0.0, nullptr, st);
        k_scale_by<<<grid_for(n, 256), 256, 0, st>>>(n, L.dinv, y);
        LAUNCH_CHECK(ctx);
        CUDA_CHECK(cudaMemcpyAsync(h.data(), y, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
        CUDA_CHECK(cudaStreamSynchronize(st));
        double ny = 0;
        for (double v : h) ny += v * v;
        ny = std::sqrt(ny);
        lam = ny;   // x is normalised
        if (ny == 0) break;
        for (double& v : h) v /= ny;
        CUDA_CHECK(cudaMemcpy(x, h.data(), sizeof(double) * n, cudaMemcpyHostToDevice));
    }
    cudaFree(tmp); cudaFree(x); cudaFree(y);
    double est = 1.1 * lam;
    if (est <= 0 || est > gersh) est = gersh;
    if (L.halo_plan >= 0 && dist_active(ctx)) {     // every rank must run the same smoother
        double* dv = nullptr;
        CUDA_CHECK(cudaMalloc(&dv, sizeof(double)));
        CUDA_CHECK(cudaMemcpy(dv, &est, sizeof(double), cudaMemcpyHostToDevice));
        dist_allreduce_max(ctx, dv, 1, st);
        CUDA_CHECK(cudaMemcpyAsync(&est, dv, sizeof(double), cudaMemcpyDeviceToHost, st));
        CUDA_CHECK(cudaStreamSynchronize(st));
        cudaFree(dv);
    }
    return est;
}

// smallest / largest eigenvalue of a symmetric tridiagonal matrix (diagonal a, off-diagonal b) by Sturm bisection
static void tridiag_extremes(const std::vector<double>& a, const std::vector<double>& b, double& emin, double& emax) {
    const int m = (int)a.size();
    double lo = a[0], hi = a[0];
    for (int i = 0; i < m; ++i) {
        const double r = (i > 0 ? std::fabs(b[i - 1]) : 0.0) + (i + 1 < m ? std::fabs(b[i]) : 0.0);
        lo = std::min(lo, a[i] - r);
        hi = std::max(hi, a[i] + r);
    }
    auto count_below = [&](double x) {      // number of eigenvalues < x
        int cnt = 0;
        double q = a[0] - x;
        if (q < 0) ++cnt;
        for (int i = 1; i < m; ++i) {
            const double den = (q == 0.0) ? 1e-300 : q;
            q = a[i] - x - b[i - 1] * b[i - 1] / den;
            if (q < 0) ++cnt;
        }
        return cnt;
    };
    auto kth = [&](int k) {                 // k-th smallest eigenvalue, k = 0 .. m-1
        double l = lo, h = hi;
        for (int it = 0; it < 200; ++it) {
            const double mid = 0.5 * (l + h);
            if (count_below(mid) <= k) l = mid; else h = mid;
            if (h - l <= 1e-15 * std::max(1.0, std::fabs(h))) break;
        }
        return 0.5 * (l + h);
    };
    emin = kth(0);
    emax = kth(m - 1);
}

// Extreme Ritz values of D^-1 A after kLanczosSteps Lanczos steps in the D inner product (deterministic start
// vector, fixed-order reductions; multi-GPU: ghost refresh before every product, owned rows in the sums,
// all-reduced).  oracle/krylov.py:lanczos_bounds is the same recurrence.
void lanczos_bounds(mpet_ctx* ctx, AmgLevel& L, const uint8_t* own_dev, double& lmin, double& lmax, cudaStream_t st) {
    const int64_t n = L.A.nrows;
    double *v = nullptr, *vp = nullptr, *w = nullptr, *sc = nullptr;
    CUDA_CHECK(cudaMalloc(&v, sizeof(double) * n));
    CUDA_CHECK(cudaMalloc(&vp, sizeof(double) * n));
    CUDA_CHECK(cudaMalloc(&w, sizeof(double) * n));
    CUDA_CHECK(cudaMalloc(&sc, sizeof(double)));
    auto dot = [&](const double* x, const double* y) {
        k_lz_dot<<<1, 1024, 0, st>>>(n, x, y, L.dinv, own_dev, sc);
        LAUNCH_CHECK(ctx);
        dist_allreduce_sum(ctx, sc, 1, st);
        double h = 0;
        CUDA_CHECK(cudaMemcpyAsync(&h, sc, sizeof(double), cudaMemcpyDeviceToHost, st));
        CUDA_CHECK(cudaStreamSynchronize(st));
        return h;
    };
    const int g = grid_for(n, 256);
    k_lz_start<<<g, 256, 0, st>>>(n, v, vp);
    LAUNCH_CHECK(ctx);
    if (L.halo_plan >= 0) dist_halo(ctx, L.halo_plan, v, false, nullptr, st);
    k_lz_scale<<<g, 256, 0, st>>>(n, v, 1.0 / std::sqrt(dot(v, v)));
    LAUNCH_CHECK(ctx);
    std::vector<double> al, be;
    double beta = 0.0;
    for (int j = 0; j < kLanczosSteps && j < n; ++j) {
        launch_plain<1>(ctx, L.A, SpmmPlan(), v, w, 0.0, nullptr, st);
        k_lz_update<<<g, 256, 0, st>>>(n, w, L.dinv, v, vp, 0.0, 0.0, 1);          // w = D^-1 A v
        LAUNCH_CHECK(ctx);
        const double a = dot(w, v);
        k_lz_update<<<g, 256, 0, st>>>(n, w, L.dinv, v, vp, a, beta, 0);            // w -= a v + beta v_prev
        LAUNCH_CHECK(ctx);
        al.push_back(a);
        const double b2 = dot(w, w);
        if (b2 <= 1e-28 * std::max(1.0, a * a)) break;
        beta = std::sqrt(b2);
        be.push_back(beta);
        k_lz_next<<<g, 256, 0, st>>>(n, w, 1.0 / beta, v, vp);
        LAUNCH_CHECK(ctx);
        if (L.halo_plan >= 0) dist_halo(ctx, L.halo_plan, v, false, nullptr, st);
    }
    CUDA_CHECK(cudaStreamSynchronize(st));
    be.resize(al.size() > 0 ? al.size() - 1 : 0);
    tridiag_extremes(al, be, lmin, lmax);
    cudaFree(v); cudaFree(vp); cudaFree(w); cudaFree(sc);
}

int poly_degree_for(double lo, double hi, double target = kPolyTarget) {
    const double kappa = hi / lo;
    const double sigma = (std::sqrt(kappa) - 1.0) / (std::sqrt(kappa) + 1.0);
    int k = 1;
    while (k < kPolyMaxDegree && 2.0 * std::pow(sigma, k) / (1.0 + std::pow(sigma, 2 * k)) > target) ++k;
    return k;
}

void compute_dinv_host(mpet_ctx* ctx, AmgLevel& L, const HostCsr& H) {
    std::vector<double> d(H.nrows, 1.0);
    for (int64_t r = 0; r < H.nrows; ++r)
        for (int32_t t = H.rp[r]; t < H.rp[r + 1]; ++t)
            if (H.ci[t] == r && H.v[t] != 0.0) d[r] = 1.0 / H.v[t];
    L.dinv = dev_alloc<double>(ctx, H.nrows);
    CUDA_CHECK(cudaMemcpy(L.dinv, d.data(), sizeof(double) * H.nrows, cudaMemcpyHostToDevice));
}

// extend H below its current last level (whose A is on device) by smoothed aggregation
void compute_dinv_host(mpet_ctx* ctx, AmgLevel& L, const HostCsr& H);

template <typename T>
void bcast_vector(mpet_ctx* ctx, std::vector<T>& v, int64_t count, int root, cudaStream_t st) {
    v.resize(count);
    if (count == 0) return;
    T* d = nullptr;
    CUDA_CHECK(cudaMalloc(&d, sizeof(T) * count));
    if (dist_rank(ctx) == root) CUDA_CHECK(cudaMemcpy(d, v.data(), sizeof(T) * count, cudaMemcpyHostToDevice));
    dist_bcast_bytes(ctx, d, (int64_t)sizeof(T) * count, root, st);
    CUDA_CHECK(cudaMemcpyAsync(v.data(), d, sizeof(T) * count, cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
    cudaFree(d);
}

// Multi-GPU: from the last mesh-defined (distributed) level to the first replicated one.  Every rank
// aggregates its OWNED vertices (plain aggregation, so the prolongator needs no neighbour data), learns
// the global aggregate id of its ghost vertices through one halo exchange, forms its rows of the
// Galerkin matrix from its complete owned rows, and the row blocks are broadcast to everybody.  The
// replicated numbering is padded: rank r's aggregates are r*maxcount .. r*maxcount+nagg_r-1, padding
// rows are identity rows.
HostCsr distributed_transition(mpet_ctx* ctx, AmgHierarchy& H, int W, cudaStream_t st) {
    AmgLevel& F = H.levels.back();
    HostCsr A = download(F.A);
    const int64_t n = A.nrows;
    // level vectors are nblocks vertex fields (1, or the A fields of the merged pressure hierarchy)
    const int64_t nblocks = n / ctx->Nv;
    MPET_REQUIRE(nblocks * ctx->Nv == n, "distributed transition level is not a multiple of the vertex count");
    const int halo_w1 = nblocks == 1 ? DIST_PLAN_P1W1 : DIST_PLAN_P1WA;
    std::vector<uint8_t> ownn((size_t)n);
    {
        const std::vector<uint8_t>& on = dist_own_nodes(ctx);   // vertices are the first Nv nodes
        for (int64_t i = 0; i < n; ++i) ownn[i] = on[i % ctx->Nv];
    }
    const int rank = dist_rank(ctx), nr = dist_nranks(ctx);
    std::vector<int32_t> l2o(n, -1), o2l;
    for (int64_t i = 0; i < n; ++i)
        if (ownn[i]) { l2o[i] = (int32_t)o2l.size(); o2l.push_back((int32_t)i); }
    HostCsr Aoo;
    Aoo.nrows = Aoo.ncols = (int64_t)o2l.size();
    Aoo.rp.assign(Aoo.nrows + 1, 0);
    for (int64_t o = 0; o < Aoo.nrows; ++o) {
        const int32_t i = o2l[o];
        for (int32_t t = A.rp[i]; t < A.rp[i + 1]; ++t)
            if (l2o[A.ci[t]] >= 0) { Aoo.ci.push_back(l2o[A.ci[t]]); Aoo.v.push_back(A.v[t]); }
        Aoo.rp[o + 1] = (int32_t)Aoo.ci.size();
    }
    std::vector<int32_t> agg;
    std::vector<char> isolated;
    std::vector<double> diag;
    int64_t nagg = 0;
    aggregate(Aoo, 0.08, agg, nagg, isolated, diag);
    // padded slot size = max aggregate count over ranks
    double* dcnt = nullptr;
    CUDA_CHECK(cudaMalloc(&dcnt, sizeof(double)));
    double cnt = (double)nagg;
    CUDA_CHECK(cudaMemcpy(dcnt, &cnt, sizeof(double), cudaMemcpyHostToDevice));
    dist_allreduce_max(ctx, dcnt, 1, st);
    CUDA_CHECK(cudaMemcpyAsync(&cnt, dcnt, sizeof(double), cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
    cudaFree(dcnt);
    const int64_t maxcount = (int64_t)cnt;
    MPET_REQUIRE(maxcount > 0 && maxcount * nr < 2147483647LL, "bad aggregate counts");
    H.gather_count = maxcount;
    // global (padded) aggregate id of every local vertex: owned from the aggregation, ghosts via halo
    std::vector<double> g(n, -1.0);
    for (int64_t o = 0; o < Aoo.nrows; ++o)
        if (agg[o] >= 0) g[o2l[o]] = (double)(rank * maxcount + agg[o]);
    double* dg = nullptr;
    CUDA_CHECK(cudaMalloc(&dg, sizeof(double) * n));
    CUDA_CHECK(cudaMemcpy(dg, g.data(), sizeof(double) * n, cudaMemcpyHostToDevice));
    dist_halo(ctx, halo_w1, dg, false, nullptr, st);
    CUDA_CHECK(cudaMemcpyAsync(g.data(), dg, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
    cudaFree(dg);
    // my rows of the Galerkin matrix (unit prolongator weights)
    struct Trip { int32_t r, c; double v; };
    std::vector<Trip> trips;
    for (int64_t o = 0; o < Aoo.nrows; ++o) {
        if (agg[o] < 0) continue;
        const int32_t i = o2l[o];
        for (int32_t t = A.rp[i]; t < A.rp[i + 1]; ++t) {
            const double gj = g[A.ci[t]];
            if (gj < 0 || A.v[t] == 0.0) continue;
            trips.push_back({agg[o], (int32_t)gj, A.v[t]});
        }
    }
    std::sort(trips.begin(), trips.end(), [](const Trip& a, const Trip& b) { return a.r != b.r ? a.r < b.r : a.c < b.c; });
    std::vector<int32_t> my_rp(nagg + 1, 0), my_ci;
    std::vector<double> my_v;
    for (size_t t = 0; t < trips.size();) {
        size_t u = t;
        double sum = 0;
        while (u < trips.size() && trips[u].r == trips[t].r && trips[u].c == trips[t].c) sum += trips[u++].v;
        my_ci.push_back(trips[t].c);
        my_v.push_back(sum);
        my_rp[trips[t].r + 1] = (int32_t)my_ci.size();
        t = u;
    }
    for (int64_t r = 0; r < nagg; ++r) my_rp[r + 1] = std::max(my_rp[r + 1], my_rp[r]);
    // replicate: every rank broadcasts its row block
    HostCsr Ac;
    Ac.nrows = Ac.ncols = maxcount * nr;
    Ac.rp.assign(Ac.nrows + 1, 0);
    for (int root = 0; root < nr; ++root) {
        std::vector<int64_t> hdr(2);
        if (root == rank) { hdr[0] = nagg; hdr[1] = (int64_t)my_ci.size(); }
        bcast_vector(ctx, hdr, 2, root, st);
        std::vector<int32_t> rp, ci;
        std::vector<double> vv;
        if (root == rank) { rp = my_rp; ci = my_ci; vv = my_v; }
        bcast_vector(ctx, rp, hdr[0] + 1, root, st);
        bcast_vector(ctx, ci, hdr[1], root, st);
        bcast_vector(ctx, vv, hdr[1], root, st);
        for (int64_t r = 0; r < maxcount; ++r) {
            const int64_t grow = root * maxcount + r;
            if (r < hdr[0]) {
                for (int32_t t = rp[r]; t < rp[r + 1]; ++t) { Ac.ci.push_back(ci[t]); Ac.v.push_back(vv[t]); }
            } else {
                Ac.ci.push_back((int32_t)grow);      // padding: identity row
                Ac.v.push_back(1.0);
            }
            Ac.rp[grow + 1] = (int32_t)Ac.ci.size();
        }
    }
    // transfer operators of this rank: P (local vertices x replicated rows), R (my padded slot x local vertices)
    HostCsr P, R;
    P.nrows = n; P.ncols = Ac.nrows;
    P.rp.assign(n + 1, 0);
    for (int64_t i = 0; i < n; ++i) {
        if (g[i] >= 0) { P.ci.push_back((int32_t)g[i]); P.v.push_back(1.0); }      // ghost rows too: no exchange after it
        P.rp[i + 1] = (int32_t)P.ci.size();
    }
    R.nrows = maxcount; R.ncols = n;
    std::vector<std::vector<int32_t>> members(maxcount);
    for (int64_t o = 0; o < Aoo.nrows; ++o)
        if (agg[o] >= 0) members[agg[o]].push_back(o2l[o]);
    R.rp.assign(maxcount + 1, 0);
    for (int64_t r = 0; r < maxcount; ++r) {
        for (int32_t m : members[r]) { R.ci.push_back(m); R.v.push_back(1.0); }
        R.rp[r + 1] = (int32_t)R.ci.size();
    }
    AmgLevel L;
    L.A = upload(ctx, Ac);
    L.P = upload(ctx, P);
    L.R = upload(ctx, R);
    compute_dinv_host(ctx, L, Ac);
    H.levels.push_back(L);
    H.gather_send = dev_alloc<double>(ctx, maxcount * W);
    dist_reserve_gather(ctx, maxcount * W * nr, st);      // collective: same size on every rank
    return Ac;
}

void extend_by_aggregation(mpet_ctx* ctx, AmgHierarchy& H, cudaStream_t st) {
    CUDA_CHECK(cudaStreamSynchronize(st));
    HostCsr A;
    if (dist_active(ctx)) {
        H.transition = (int)H.levels.size() - 1;
        A = distributed_transition(ctx, H, H.nrhs, st);
    } else {
        A = download(H.levels.back().A);
    }
    double theta = 0.08;
    while (A.nrows > kCoarseMax && (int)H.levels.size() < kMaxLevels) {
        int64_t nagg = 0;
        HostCsr P = sa_prolongator(A, theta, nagg);
        if (nagg == 0 || nagg > 0.8 * A.nrows) break;
        HostCsr R = transpose(P);
        HostCsr Ac = drop_small(spgemm(R, spgemm(A, P)), kDropTol);
        AmgLevel L;
        L.A = upload(ctx, Ac);
        L.P = upload(ctx, P);
        L.R = upload(ctx, R);
        compute_dinv_host(ctx, L, Ac);
        H.levels.push_back(L);
        A = std::move(Ac);
        theta *= 0.5;
    }
    if (A.nrows <= 700) {
        std::vector<double> inv = dense_inverse(A);
        H.coarse_n = A.nrows;
        H.coarse_inv = dev_alloc<double>(ctx, (int64_t)inv.size());
        CUDA_CHECK(cudaMemcpy(H.coarse_inv, inv.data(), sizeof(double) * inv.size(), cudaMemcpyHostToDevice));
    }
}

void make_plan(mpet_ctx* ctx, const DevCsr& M, SpmmPlan& plan) {
    if (M.nrows < 4000 || M.nnz == 0) return;      // small levels: launch latency dominates either way
    std::vector<int32_t> rp(M.nrows + 1);
    CUDA_CHECK(cudaMemcpy(rp.data(), M.rowptr, sizeof(int32_t) * (M.nrows + 1), cudaMemcpyDeviceToHost));
    if (!staged_build_spmm_plan(ctx, rp, plan)) plan = SpmmPlan();
}

void finish_hierarchy(mpet_ctx* ctx, AmgHierarchy& H, cudaStream_t st) {
    if (getenv("MPET_AMG_VERBOSE")) {
        fprintf(stderr, "[mpet amg] hierarchy W=%d:", H.nrhs);
        for (auto& L : H.levels) fprintf(stderr, " %lld rows/%lld nnz", (long long)L.A.nrows, (long long)L.A.nnz);
        fprintf(stderr, " | coarse %s (%lld)\n", H.coarse_inv ? "dense inverse" : "Chebyshev", (long long)H.levels.back().A.nrows);
    }
    if (H.poly_degree == 0) {     // work vectors of multi-cycle applications (owned by the hierarchy's arena)
        H.cyc_r = dev_alloc<double>(ctx, H.levels[0].A.nrows * H.nrhs);
        H.cyc_e = dev_alloc<double>(ctx, H.levels[0].A.nrows * H.nrhs);
    }
    for (auto& L : H.levels) {
        alloc_level_work(ctx, L, H.nrhs);
        L.lambda_max = estimate_lambda_max(ctx, L, st);
        make_plan(ctx, L.A, L.planA);
        if (L.P.nrows > 0) { make_plan(ctx, L.P, L.planP); make_plan(ctx, L.R, L.planR); }
    }
}

DevCsr masked_block(mpet_ctx* ctx, const NodeGraph& g, const double* vals, double scale, const uint8_t* mask,
                    double** dinv_out, cudaStream_t st, const double* vals2 = nullptr, double scale2 = 0.0) {
    DevCsr M;
    M.nrows = g.nrows; M.ncols = g.ncols; M.nnz = g.nnz;
    M.rowptr = g.rowptr; M.col = g.col;
    M.val = dev_alloc<double>(ctx, g.nnz);
    *dinv_out = dev_alloc<double>(ctx, g.nrows);
    k_masked_copy<<<grid_for(g.nrows * 8, 256), 256, 0, st>>>(g.nrows, g.rowptr, g.col, vals, scale, mask, M.val, *dinv_out,
                                                              scale2 != 0.0 ? vals2 : nullptr, scale2);
    LAUNCH_CHECK(ctx);
    return M;
}

// multi-GPU: the local matrix misses part of the ghost rows, so their diagonal is wrong; take it from the owners
// (scalar_plan: DIST_PLAN_P1W1 for vertex levels, -1 for the scalar P2 level, which goes through a 4-wide copy)
void sync_ghost_diagonal(mpet_ctx* ctx, AmgLevel& L, bool p2_level, cudaStream_t st) {
    if (!dist_active(ctx)) return;
    const int64_t n = L.A.nrows;
    if (!p2_level) {
        dist_halo(ctx, DIST_PLAN_P1W1, L.dinv, false, nullptr, st);
    } else {
        double* tmp = nullptr;
        CUDA_CHECK(cudaMalloc(&tmp, sizeof(double) * 4 * n));
        k_expand4<<<grid_for(4 * n, 256), 256, 0, st>>>(n, L.dinv, tmp);
        LAUNCH_CHECK(ctx);
        dist_halo(ctx, DIST_PLAN_P2W4, tmp, false, nullptr, st);
        k_take4<<<grid_for(n, 256), 256, 0, st>>>(n, tmp, L.dinv);
        LAUNCH_CHECK(ctx);
        CUDA_CHECK(cudaStreamSynchronize(st));
        cudaFree(tmp);
    }
}

// P2 -> P1 embedding (rows: P2 nodes, cols: vertices); Dirichlet rows/cols dropped
HostCsr p2_to_p1(int64_t nv, int64_t ne, const std::vector<int32_t>& edge_v, const std::vector<uint8_t>& mask2) {
    HostCsr P;
    P.nrows = nv + ne; P.ncols = nv;
    P.rp.assign(P.nrows + 1, 0);
    for (int64_t r = 0; r < P.nrows; ++r) {
        if (!mask2[r]) {
            if (r < nv) { P.ci.push_back((int32_t)r); P.v.push_back(1.0); }
            else {
                int32_t a = edge_v[2 * (r - nv)], b = edge_v[2 * (r - nv) + 1];
                if (!mask2[a]) { P.ci.push_back(a); P.v.push_back(0.5); }
                if (!mask2[b]) { P.ci.push_back(b); P.v.push_back(0.5); }
            }
        }
        P.rp[r + 1] = (int32_t)P.ci.size();
    }
    return P;
}

}  // namespace

// ---------------------------------------------------------------------------------- public
void amg_free(mpet_ctx* ctx) {
    for (void* p : ctx->amg_allocs) cudaFree(p);
    ctx->amg_allocs.clear();
    delete ctx->amg_u;
    ctx->amg_u = nullptr;
    for (int i = 0; i < MPET_MAX_NETWORKS; ++i) { delete ctx->amg_p[i]; ctx->amg_p[i] = nullptr; }
}

void amg_setup(mpet_ctx* ctx, cudaStream_t st) {
    MPET_REQUIRE(ctx->prec_ready, "mpet_assemble_prec must run before mpet_pc_setup");
    amg_free(ctx);                    // a new dt / new Dirichlet set rebuilds the hierarchies
    ctx->graph_epoch++;               // captured iterations hold pointers into the old hierarchies
    ctx->border_scaled = false;       // c_i . B c_i of the multiplier block belongs to the old hierarchy
    struct ArenaGuard {
        mpet_ctx* c;
        explicit ArenaGuard(mpet_ctx* c_) : c(c_) { c->arena = &c->amg_allocs; }
        ~ArenaGuard() { c->arena = nullptr; }
    } guard(ctx);
    const int64_t n2 = ctx->N2, nv = ctx->Nv;
    const int A = ctx->A;
    if (!ctx->bc_mask) {
        ctx->bc_mask = dev_alloc<uint8_t>(ctx, ctx->N);
        CUDA_CHECK(cudaMemsetAsync(ctx->bc_mask, 0, ctx->N, st));
    }
    // the three displacement components must share their Dirichlet set to share one hierarchy
    std::vector<uint8_t> hmask(ctx->N);
    CUDA_CHECK(cudaMemcpyAsync(hmask.data(), ctx->bc_mask, ctx->N, cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
    for (int k = 1; k < 3; ++k)
        MPET_REQUIRE(std::equal(hmask.begin(), hmask.begin() + n2, hmask.begin() + k * n2),
                     "displacement components carry different Dirichlet sets (not supported by the shared block)");
    std::vector<uint8_t> mask2(hmask.begin(), hmask.begin() + n2);
    std::vector<int32_t> edge_v(2 * ctx->Ne);
    CUDA_CHECK(cudaMemcpy(edge_v.data(), ctx->edge_v, sizeof(int32_t) * 2 * ctx->Ne, cudaMemcpyDeviceToHost));

    // ---- displacement block
    {
        AmgHierarchy* H = new AmgHierarchy();
        H->nrhs = 4;   // W = 4: three components + pad
        AmgLevel L0;
        // + shift (u, v): the displacement block without essential boundary conditions (u_has_nullspace) is singular
        const double shift = ctx->prec_shift_u;
        if (shift != 0.0) ensure_m22(ctx, st);
        L0.A = masked_block(ctx, ctx->g22, ctx->k22, ctx->coef.p_mu, ctx->bc_mask, &L0.dinv, st, ctx->m22, shift);
        if (dist_active(ctx)) L0.halo_plan = DIST_PLAN_P2W4;
        sync_ghost_diagonal(ctx, L0, true, st);
        H->levels.push_back(L0);
        AmgLevel L1;
        L1.A = masked_block(ctx, ctx->g11, ctx->l11, ctx->coef.p_mu, ctx->bc_mask /* vertices are the first Nv nodes */,
                            &L1.dinv, st, ctx->m11, shift);
        if (dist_active(ctx)) L1.halo_plan = DIST_PLAN_P1W4;
        sync_ghost_diagonal(ctx, L1, false, st);
        HostCsr P = p2_to_p1(nv, ctx->Ne, edge_v, mask2);
        L1.P = upload(ctx, P);
        L1.R = upload(ctx, transpose(P));
        H->levels.push_back(L1);
        extend_by_aggregation(ctx, *H, st);
        finish_hierarchy(ctx, *H, st);
        ctx->amg_u = H;
    }
    // ---- P1 fields: one hierarchy per field.  (Merging the fields into one block-diagonal hierarchy was tried:
    // one third of the launches, but the shared Chebyshev interval / prolongator damping cost 11 % more MINRES
    // iterations on cfg5 -- 614 vs 552 -- for the same time per iteration.)
    uint8_t* own_dev = nullptr;
    if (dist_active(ctx) && A > 0) {
        const std::vector<uint8_t>& on = dist_own_nodes(ctx);
        own_dev = dev_alloc<uint8_t>(ctx, nv);
        CUDA_CHECK(cudaMemcpy(own_dev, on.data(), (size_t)nv, cudaMemcpyHostToDevice));
    }
    for (int i = 0; i < A; ++i) {
        AmgHierarchy* H = new AmgHierarchy();
        H->nrhs = 1;
        AmgLevel L0;
        L0.is_mesh_level = true;
        L0.A = masked_block(ctx, ctx->g11, ctx->pp11 + (int64_t)i * ctx->g11.nnz, 1.0,
                            ctx->bc_mask + 3 * n2 + (int64_t)i * nv, &L0.dinv, st);
        if (dist_active(ctx)) L0.halo_plan = DIST_PLAN_P1W1;
        sync_ghost_diagonal(ctx, L0, false, st);
        H->levels.push_back(L0);
        // spectrum of D^-1 A: a small condition number (mass-dominated field) is inverted far more accurately --
        // and cheaper -- by a Chebyshev polynomial on the whole spectrum than by a V-cycle, and MINRES on this
        // system is very sensitive to the accuracy of the P1 blocks.
        double lmin = 0, lmax = 0;
        lanczos_bounds(ctx, H->levels[0], own_dev, lmin, lmax, st);
        const double lo = 0.97 * lmin, hi = 1.02 * lmax;
        static const double kappa_max = []() { const char* e = getenv("MPET_POLY_KAPPA_MAX"); return e ? atof(e) : kPolyKappaMax; }();
        if (lmin > 0 && hi / lo <= kappa_max && !getenv("MPET_NO_POLY")) {
            H->poly_degree = poly_degree_for(lo, hi);
            static const double light_target = []() { const char* e = getenv("MPET_POLY_TARGET_LIGHT"); return e ? atof(e) : kPolyTargetLight; }();
            H->poly_degree_light = poly_degree_for(lo, hi, light_target);
            H->poly_lo = lo;
            H->poly_hi = hi;
        } else {
            extend_by_aggregation(ctx, *H, st);
        }
        if (getenv("MPET_AMG_VERBOSE"))
            fprintf(stderr, "[mpet amg] P1 field %d: D^-1 A spectrum [%.4f, %.4f] -> %s (degree %d)\n", i, lmin, lmax,
                    H->poly_degree ? "Chebyshev polynomial" : "V-cycle", H->poly_degree);
        finish_hierarchy(ctx, *H, st);
        ctx->amg_p[i] = H;
    }
    CUDA_CHECK(cudaStreamSynchronize(st));
}

// k cycles per application = the symmetric stationary iteration x <- x + B (b - A x) started from zero
// (kP1Cycles for the P1 fields, 1 for the displacement block; development aid: MPET_P_CYCLES / MPET_U_CYCLES)
template <int W>
static void cycles(mpet_ctx* ctx, AmgHierarchy& H, int ncycles, const double* b, double* x, const int* done,
                   cudaStream_t st) {
    vcycle<W>(ctx, H, 0, b, x, done, st);
    AmgLevel& L = H.levels[0];
    for (int c = 1; c < ncycles; ++c) {
        launch_epi<W, EPI_RESID>(ctx, L.A, L.planA, x, b, H.cyc_r, nullptr, nullptr, 0, 0, done, st);
        if (L.halo_plan >= 0) dist_halo(ctx, L.halo_plan, H.cyc_r, false, done, st);
        vcycle<W>(ctx, H, 0, H.cyc_r, H.cyc_e, done, st);
        k_add_inplace<<<grid_for(L.A.nrows * W, 256), 256, 0, st>>>(L.A.nrows * W, H.cyc_e, x, done);
        LAUNCH_CHECK(ctx);
        ctx->pc_bytes_acc += 24 * L.A.nrows * W;
    }
}

// r, z in the solver-internal layout (layout.cuh).  The block-diagonal preconditioner is A + 1 independent
// cycles; on one GPU the per-field cycles (dozens of small launches each) run on their own high-priority
// streams beside the displacement cycle and join before returning.  Multi-GPU keeps one stream: the halo
// exchanges of all cycles share one NCCL communicator and must be issued in one order on every rank.
static void amg_apply_impl(mpet_ctx* ctx, const double* r, double* z, const int* done, cudaStream_t st);

void amg_apply(mpet_ctx* ctx, const double* r, double* z, const int* done, cudaStream_t st) {
    ctx->pc_bytes_acc = 0;
    amg_apply_impl(ctx, r, z, done, st);
    ctx->pc_bytes_last = ctx->pc_bytes_acc;
}

static void amg_apply_impl(mpet_ctx* ctx, const double* r, double* z, const int* done, cudaStream_t st) {
    const int64_t n2 = ctx->N2, nv = ctx->Nv;
    static const bool want_streams = []() { const char* e = getenv("MPET_PC_STREAMS"); return !(e && e[0] == '0'); }();
    static const int p_cycles = []() { const char* e = getenv("MPET_P_CYCLES"); return e ? std::max(1, atoi(e)) : 0; }();
    static const int u_cycles = []() { const char* e = getenv("MPET_U_CYCLES"); return e ? std::max(1, atoi(e)) : 1; }();
    const bool fork = want_streams && ctx->A > 0 && !dist_active(ctx);
    const int p_cyc = p_cycles > 0 ? p_cycles : (ctx->pc_light ? 1 : (dist_active(ctx) ? kP1CyclesDist : kP1Cycles));
    if (fork) {
        if (ctx->pc_streams_ready < ctx->A) {
            int lo = 0, hi = 0;
            CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&lo, &hi));      // hi = numerically lowest = highest priority
            if (!ctx->pc_fork) CUDA_CHECK(cudaEventCreateWithFlags(&ctx->pc_fork, cudaEventDisableTiming));
            for (int i = ctx->pc_streams_ready; i < ctx->A; ++i) {
                CUDA_CHECK(cudaStreamCreateWithPriority(&ctx->pc_stream[i], cudaStreamNonBlocking, hi));
                CUDA_CHECK(cudaEventCreateWithFlags(&ctx->pc_join[i], cudaEventDisableTiming));
            }
            ctx->pc_streams_ready = ctx->A;
        }
        CUDA_CHECK(cudaEventRecord(ctx->pc_fork, st));
        for (int i = 0; i < ctx->A; ++i) {
            const int64_t off = 4 * n2 + (int64_t)i * nv;
            AmgHierarchy& H = *ctx->amg_p[i];
            CUDA_CHECK(cudaStreamWaitEvent(ctx->pc_stream[i], ctx->pc_fork, 0));
            cycles<1>(ctx, H, H.poly_degree ? 1 : p_cyc, r + off, z + off, done, ctx->pc_stream[i]);
            CUDA_CHECK(cudaEventRecord(ctx->pc_join[i], ctx->pc_stream[i]));
        }
        cycles<4>(ctx, *ctx->amg_u, u_cycles, r, z, done, st);
        for (int i = 0; i < ctx->A; ++i) CUDA_CHECK(cudaStreamWaitEvent(st, ctx->pc_join[i], 0));
        return;
    }
    if (want_streams && ctx->A > 0 && dist_has_lane1(ctx)) {
        // multi-GPU: all P1-field cycles on ONE side stream with their own communicator (lane 1); their many small
        // halo exchanges then overlap the displacement cycle instead of queueing behind it
        if (ctx->pc_streams_ready < 1) {
            int lo = 0, hi = 0;
            CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
            if (!ctx->pc_fork) CUDA_CHECK(cudaEventCreateWithFlags(&ctx->pc_fork, cudaEventDisableTiming));
            CUDA_CHECK(cudaStreamCreateWithPriority(&ctx->pc_stream[0], cudaStreamNonBlocking, hi));
            CUDA_CHECK(cudaEventCreateWithFlags(&ctx->pc_join[0], cudaEventDisableTiming));
            ctx->pc_streams_ready = 1;
        }
        cudaStream_t sp = ctx->pc_stream[0];
        CUDA_CHECK(cudaEventRecord(ctx->pc_fork, st));
        CUDA_CHECK(cudaStreamWaitEvent(sp, ctx->pc_fork, 0));
        ctx->dist_lane = 1;
        for (int i = 0; i < ctx->A; ++i) {
            const int64_t off = 4 * n2 + (int64_t)i * nv;
            AmgHierarchy& H = *ctx->amg_p[i];
            cycles<1>(ctx, H, H.poly_degree ? 1 : p_cyc, r + off, z + off, done, sp);
        }
        ctx->dist_lane = 0;
        CUDA_CHECK(cudaEventRecord(ctx->pc_join[0], sp));
        cycles<4>(ctx, *ctx->amg_u, u_cycles, r, z, done, st);
        CUDA_CHECK(cudaStreamWaitEvent(st, ctx->pc_join[0], 0));
        return;
    }
    // serial order (MPET_PC_STREAMS=0): also the measurement mode of the per-block times (mpet_profile 6, 7)
    cudaEvent_t pe = prof_begin(ctx, st);
    cycles<4>(ctx, *ctx->amg_u, u_cycles, r, z, done, st);
    prof_end(ctx, PROF_PC_U, pe, st);
    pe = prof_begin(ctx, st);
    for (int i = 0; i < ctx->A; ++i) {
        const int64_t off = 4 * n2 + (int64_t)i * nv;
        AmgHierarchy& H = *ctx->amg_p[i];
        cycles<1>(ctx, H, H.poly_degree ? 1 : p_cyc, r + off, z + off, done, st);
    }
    prof_end(ctx, PROF_PC_P, pe, st);
}

int amg_num_levels(mpet_ctx* ctx, int block) {
    AmgHierarchy* H = block == 0 ? ctx->amg_u : ctx->amg_p[block - 1];
    return H ? (int)H->levels.size() : 0;
}
