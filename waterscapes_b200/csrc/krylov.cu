// Preconditioned Krylov solvers on device: MINRES (symmetric S) and restarted GMRES (any S).
//
// Replaces PETScKrylovSolver("minres", "hypre_amg").solve (mpetsolver.py:507,553-556;
// mpettotalpressuresolver.py:448,492-493) and -- run to a tight tolerance -- the LUSolver path
// (mpetsolver.py:347,379,422,456).  Algorithm = PETSc's KSPMINRES recurrence (one operator and one
// preconditioner application per iteration, convergence on the preconditioned norm |eta|).
//
// Dirichlet conditions are imposed the way apply_symmetric(bc, A, b) does (bc_symmetric.py:11-22)
// without ever touching A: the initial guess carries the boundary values, the initial residual is
// zero on Dirichlet rows, and the operator returns identity on those rows (spmv.cu row mask).  All
// Krylov vectors then stay exactly zero there, so columns need no masking.
//
// All scalars (Lanczos coefficients, Givens rotations, convergence flag) live in device memory and
// are updated by single-thread kernels, so the host enqueues iterations without synchronising; once
// the `done` flag is set every later kernel returns immediately.  Reductions are two-stage with a
// fixed block count and fixed summation order (bit-reproducible).
#include "ctx.h"
#include <cmath>
#include <vector>

void scatter_bc_values_int(mpet_ctx* ctx, double* out_int, cudaStream_t st);   // rhs.cu
// multi-GPU hooks: see ctx.h (dist.cu)


namespace {

const int kRedBlocks = 1184;   // 148 SMs x 8
const int kBorderPad = 16;     // entries appended to every Krylov vector when multipliers border the system
const int kRedThreads = 256;
const int kGmresMaxRestart = 120, kGmresLd = kGmresMaxRestart + 1;

enum {
    S_BETA = 0, S_BETA_OLD, S_ETA, S_C, S_C_OLD, S_S, S_S_OLD, S_ALPHA, S_RHO1, S_RHO2, S_RHO3, S_CETA,
    S_NORM, S_NORM0, S_TOL, S_DP, S_BNORM, S_STAG_REF, S_BTRUE, S_INV, S_COUNT = 32
};
enum { F_DONE = 0, F_CONV, F_ITERS, F_MAXIT, F_BREAKDOWN, F_PENDING, F_STAG, F_JCOUNT, F_COUNT = 8 };
// stagnation exit: less than 1 % reduction of the (monotone) MINRES residual norm over kStagWindow iterations
const int kStagWindow = 256;
const double kStagFactor = 0.99;

__device__ __forceinline__ double block_reduce_sum(double v, double* sm) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) sm[w] = v;
    __syncthreads();
    int nw = blockDim.x >> 5;
    v = (threadIdx.x < nw) ? sm[threadIdx.x] : 0.0;
    if (w == 0) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    }
    __syncthreads();
    return v;   // valid in thread 0
}

__global__ void __launch_bounds__(kRedThreads)
k_dot_partial(const double* __restrict__ a, const double* __restrict__ b, int64_t n,
              double* __restrict__ partials, const uint8_t* __restrict__ own, const int* __restrict__ done) {
    if (done && *done) return;
    __shared__ double sm[32];
    double s = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        if (!own || own[i]) s += a[i] * b[i];      // multi-GPU: every dof is counted by its owner only
    s = block_reduce_sum(s, sm);
    if (threadIdx.x == 0) partials[blockIdx.x] = s;
}

// sum of kRedBlocks partials in a fixed order by one block
__device__ double final_sum(const double* __restrict__ partials, int nparts, double* sm) {
    double s = 0.0;
    for (int i = threadIdx.x; i < nparts; i += blockDim.x) s += partials[i];
    return block_reduce_sum(s, sm);
}

__global__ void __launch_bounds__(kRedThreads)
k_final_store(const double* __restrict__ partials, int nparts, double* __restrict__ out,
              const int* __restrict__ done) {
    if (done && *done) return;
    __shared__ double sm[32];
    double s = final_sum(partials, nparts, sm);
    if (threadIdx.x == 0) *out = s;
}

// r = mask ? (dvals ? dvals : 0) : b - y
__global__ void k_residual(const double* __restrict__ b, const double* __restrict__ y,
                           const uint8_t* __restrict__ mask, int64_t n, double* __restrict__ r,
                           const double* __restrict__ dvals) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) r[i] = (mask && mask[i]) ? (dvals ? dvals[i] : 0.0) : b[i] - y[i];
}

__global__ void k_jacobi(const double* __restrict__ dinv, const double* __restrict__ r, int64_t n,
                         double* __restrict__ z, const int* __restrict__ done) {
    if (done && *done) return;
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) z[i] = dinv[i] * r[i];
}

__global__ void k_abs_diag_inv(int64_t n, const int64_t* __restrict__ rowptr, const int32_t* __restrict__ cols,
                               const double* __restrict__ vals, const uint8_t* __restrict__ mask,
                               double* __restrict__ dinv) {
    int64_t row = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (row >= n) return;
    if (mask && mask[row]) { dinv[row] = 1.0; return; }
    int64_t lo = rowptr[row], hi = rowptr[row + 1];
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (cols[mid] < (int32_t)row) lo = mid + 1; else hi = mid;
    }
    double d = vals[lo];
    dinv[row] = (d != 0.0) ? 1.0 / fabs(d) : 1.0;
}

// ---------------------------------------------------------------------------------- MINRES
__global__ void __launch_bounds__(kRedThreads)
k_minres_init(const double* __restrict__ red, double rtol, double atol, int maxit, int norm_mode,
              double* __restrict__ sc, int* __restrict__ fl) {
    if (threadIdx.x != 0) return;
    const double dp = red[0];       // r0 . B r0
    const double dpb = red[1];      // b . B b  (b after the symmetric Dirichlet elimination)
    for (int i = 0; i < S_COUNT; ++i) sc[i] = 0.0;
    for (int i = 0; i < F_COUNT; ++i) fl[i] = 0;
    fl[F_MAXIT] = maxit;
    if (dp < 0.0) { fl[F_BREAKDOWN] = 1; fl[F_DONE] = 1; return; }   // indefinite preconditioner
    double beta = sqrt(dp);
    sc[S_BETA] = beta; sc[S_ETA] = beta; sc[S_C] = 1.0; sc[S_C_OLD] = 1.0;
    sc[S_NORM] = beta; sc[S_NORM0] = beta;
    // PETSc's default test (KSPConvergedDefault) with a nonzero initial guess: the reference norm is the
    // preconditioned norm of the RIGHT-HAND SIDE, not of the initial residual; a zero rhs falls back to |r0|
    double bnorm = dpb > 0.0 ? sqrt(dpb) : beta;
    sc[S_BTRUE] = bnorm;
    if (norm_mode == 1) bnorm = fmin(bnorm, beta);    // KSPConvergedDefaultSetUMIRNorm: min(|b|, |r0|)
    sc[S_BNORM] = bnorm;
    sc[S_STAG_REF] = 1e300;
    sc[S_TOL] = fmax(rtol * bnorm, atol);
    if (beta <= sc[S_TOL] || beta == 0.0) { fl[F_CONV] = 1; fl[F_DONE] = 1; }
}

__global__ void k_minres_start(int64_t n, const double* __restrict__ sc, const double* __restrict__ r,
                               const double* __restrict__ z, double* __restrict__ v, double* __restrict__ u,
                               double* __restrict__ v_old, double* __restrict__ u_old, double* __restrict__ w1,
                               double* __restrict__ w2, const int* __restrict__ done) {
    if (*done) return;
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    double ib = 1.0 / sc[S_BETA];
    v[i] = r[i] * ib; u[i] = z[i] * ib;
    v_old[i] = 0.0; u_old[i] = 0.0; w1[i] = 0.0; w2[i] = 0.0;
}

__global__ void __launch_bounds__(kRedThreads)
k_minres_alpha(const double* __restrict__ red, double* __restrict__ sc, const int* __restrict__ fl) {
    if (fl[F_DONE]) return;
    if (threadIdx.x == 0) sc[S_ALPHA] = *red;
}

// r -= alpha v + beta v_old ; z -= alpha u + beta u_old ; partial(r.z)
__global__ void __launch_bounds__(kRedThreads)
k_minres_update(int64_t n, const double* __restrict__ sc, const double* __restrict__ v,
                const double* __restrict__ v_old, const double* __restrict__ u,
                const double* __restrict__ u_old, double* __restrict__ r, double* __restrict__ z,
                double* __restrict__ partials, const uint8_t* __restrict__ own, const int* __restrict__ done) {
    if (*done) return;
    __shared__ double sm[32];
    const double alpha = sc[S_ALPHA], beta = sc[S_BETA];
    double s = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double ri = r[i] - alpha * v[i] - beta * v_old[i];
        double zi = z[i] - alpha * u[i] - beta * u_old[i];
        r[i] = ri; z[i] = zi;
        if (!own || own[i]) s += ri * zi;
    }
    s = block_reduce_sum(s, sm);
    if (threadIdx.x == 0) partials[blockIdx.x] = s;
}

__global__ void __launch_bounds__(kRedThreads)
k_minres_rotate(const double* __restrict__ red, double* __restrict__ sc, int* __restrict__ fl) {
    if (fl[F_DONE]) return;
    if (threadIdx.x != 0) return;
    double dp = *red;
    if (dp < 0.0) {
        // tolerate round-off-sized negatives (exact convergence), flag real indefiniteness
        if (-dp > 1e-12 * sc[S_NORM0] * sc[S_NORM0]) fl[F_BREAKDOWN] = 1;
        dp = 0.0;
    }
    double beta_old = sc[S_BETA];
    double beta = sqrt(dp);
    double alpha = sc[S_ALPHA];
    double c_oold = sc[S_C_OLD], c_old = sc[S_C], s_oold = sc[S_S_OLD], s_old = sc[S_S];
    double rho0 = c_old * alpha - c_oold * s_old * beta_old;
    double rho1 = sqrt(rho0 * rho0 + beta * beta);
    double rho2 = s_old * alpha + c_oold * c_old * beta_old;
    double rho3 = s_oold * beta_old;
    if (rho1 == 0.0) { fl[F_BREAKDOWN] = 1; fl[F_CONV] = 0; fl[F_DONE] = 1; return; }
    double c = rho0 / rho1, s = beta / rho1;
    double eta = sc[S_ETA];
    sc[S_C_OLD] = c_old; sc[S_S_OLD] = s_old; sc[S_C] = c; sc[S_S] = s;
    sc[S_RHO1] = rho1; sc[S_RHO2] = rho2; sc[S_RHO3] = rho3;
    sc[S_CETA] = c * eta;
    sc[S_ETA] = -s * eta;
    sc[S_BETA_OLD] = beta_old; sc[S_BETA] = beta;
    sc[S_NORM] = fabs(sc[S_ETA]);
    sc[S_DP] = dp;
    int it = fl[F_ITERS] + 1;
    fl[F_ITERS] = it;
    if (sc[S_NORM] <= sc[S_TOL]) fl[F_CONV] = 1;
    if (it % kStagWindow == 0) {
        if (sc[S_NORM] > kStagFactor * sc[S_STAG_REF]) fl[F_STAG] = 1;
        sc[S_STAG_REF] = sc[S_NORM];
    }
    if (fl[F_CONV] || it >= fl[F_MAXIT] || beta == 0.0 || fl[F_BREAKDOWN] || fl[F_STAG]) fl[F_PENDING] = 1;
}

// w = (u - rho2 w1 - rho3 w2) / rho1 ; x += c*eta*w ; shift Lanczos vectors
__global__ void k_minres_finalize(int64_t n, const double* __restrict__ sc, double* __restrict__ x,
                                  double* __restrict__ w1, double* __restrict__ w2, double* __restrict__ v,
                                  double* __restrict__ v_old, double* __restrict__ u, double* __restrict__ u_old,
                                  const double* __restrict__ r, const double* __restrict__ z,
                                  const int* __restrict__ done) {
    if (*done) return;
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double irho1 = 1.0 / sc[S_RHO1], rho2 = sc[S_RHO2], rho3 = sc[S_RHO3], ceta = sc[S_CETA];
    const double beta = sc[S_BETA];
    const double ib = beta != 0.0 ? 1.0 / beta : 0.0;
    double ui = u[i], w1i = w1[i];
    double wn = (ui - rho2 * w1i - rho3 * w2[i]) * irho1;
    w2[i] = w1i; w1[i] = wn;
    x[i] += ceta * wn;
    v_old[i] = v[i]; v[i] = r[i] * ib;
    u_old[i] = ui;   u[i] = z[i] * ib;
}

__global__ void k_commit(int* __restrict__ fl) {
    if (fl[F_PENDING]) fl[F_DONE] = 1;
}

// ---------------------------------------------------------------------------------- GMRES helpers
template <int NV>
__global__ void __launch_bounds__(kRedThreads)
k_multidot(const double* __restrict__ V, int64_t ldv, const double* __restrict__ w, int64_t n,
           double* __restrict__ partials /* [NV][gridDim] */, const uint8_t* __restrict__ own) {
    __shared__ double sm[32];
    double s[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) s[k] = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        if (own && !own[i]) continue;              // multi-GPU: every dof is counted by its owner only
        double wi = w[i];
#pragma unroll
        for (int k = 0; k < NV; ++k) s[k] += V[k * ldv + i] * wi;
    }
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        double r = block_reduce_sum(s[k], sm);
        if (threadIdx.x == 0) partials[k * gridDim.x + blockIdx.x] = r;
    }
}

__global__ void __launch_bounds__(kRedThreads)
k_multifinal(const double* __restrict__ partials, int nparts, double* __restrict__ out, int accumulate) {
    __shared__ double sm[32];
    double s = final_sum(partials + (int64_t)blockIdx.x * nparts, nparts, sm);
    if (threadIdx.x == 0) out[blockIdx.x] = accumulate ? out[blockIdx.x] + s : s;
}

// w -= sum_k h[k] V[k]   (h on device)
__global__ void k_multi_axpy(const double* __restrict__ V, int64_t ldv, int nv, const double* __restrict__ h,
                             double sign, int64_t n, double* __restrict__ w) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    double acc = w[i];
    for (int k = 0; k < nv; ++k) acc += sign * h[k] * V[k * ldv + i];
    w[i] = acc;
}

// dst = src * (*scale), scale on the device (1 / ||.|| of the vector being normalised)
__global__ void k_scale_copy(const double* __restrict__ src, const double* __restrict__ scale, int64_t n,
                             double* __restrict__ dst, const int* __restrict__ done) {
    if (done && *done) return;
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[i] * *scale;
}

// x += sum_{k < *count} y[k] V[k]   (end of a GMRES cycle; the column count of the cycle lives on the device)
__global__ void k_multi_axpy_count(const double* __restrict__ V, int64_t ldv, const int* __restrict__ count,
                                   const double* __restrict__ y, int64_t n, double* __restrict__ x) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int nv = *count;
    double acc = x[i];
    for (int k = 0; k < nv; ++k) acc += y[k] * V[k * ldv + i];
    x[i] = acc;
}

// GMRES scalar recurrences on the device (one thread; the Hessenberg matrix has at most 120 columns), so that
// the host only polls the flags every few iterations, like MINRES.
__global__ void k_gmres_init(const double* __restrict__ red, int maxit, double* __restrict__ sc, int* __restrict__ fl) {
    if (threadIdx.x != 0) return;
    for (int i = 0; i < S_COUNT; ++i) sc[i] = 0.0;
    for (int i = 0; i < F_COUNT; ++i) fl[i] = 0;
    sc[S_BNORM] = sqrt(fmax(red[0], 0.0));        // ||B b||_2: PETSc's default reference norm (left preconditioning)
    sc[S_NORM0] = -1.0;
    fl[F_MAXIT] = maxit;
}

// start of a cycle: beta = ||B r||_2 (red[0] = its square); g = beta e_0
__global__ void k_gmres_begin(const double* __restrict__ red, double rtol, double atol, int norm_mode, int m,
                              double* __restrict__ sc, int* __restrict__ fl, double* __restrict__ g) {
    if (threadIdx.x != 0 || fl[F_DONE]) return;
    const double beta = sqrt(fmax(red[0], 0.0));
    if (sc[S_NORM0] < 0.0) {
        sc[S_NORM0] = beta;
        double bn = sc[S_BNORM];
        if (!(bn > 0.0)) bn = beta;
        sc[S_BTRUE] = bn;
        if (norm_mode == 1) bn = fmin(bn, beta);
        sc[S_BNORM] = bn;
        sc[S_TOL] = fmax(rtol * bn, atol);
    }
    sc[S_NORM] = beta;
    fl[F_JCOUNT] = 0;
    if (beta <= sc[S_TOL] || beta == 0.0) { fl[F_CONV] = 1; fl[F_DONE] = 1; return; }
    if (fl[F_ITERS] >= fl[F_MAXIT]) { fl[F_DONE] = 1; return; }
    g[0] = beta;
    for (int i = 1; i <= m; ++i) g[i] = 0.0;
    sc[S_INV] = 1.0 / beta;
}

// column j of the Hessenberg matrix: h = h1 + h2 (the two Gram-Schmidt passes), h[j+1] = ||w|| (red[0] = its
// square); previous Givens rotations, the new one, the residual norm estimate |g[j+1]| and the convergence test.
// R (upper triangular after the rotations) is column-major with leading dimension ldh.
__global__ void k_gmres_rotate(int j, int ldh, const double* __restrict__ h1, const double* __restrict__ h2,
                               const double* __restrict__ red, double* __restrict__ R, double* __restrict__ cs,
                               double* __restrict__ sn, double* __restrict__ g, double* __restrict__ sc,
                               int* __restrict__ fl) {
    if (threadIdx.x != 0 || fl[F_DONE]) return;
    double* col = R + (int64_t)j * ldh;
    for (int i = 0; i <= j; ++i) col[i] = h1[i] + h2[i];
    const double hn = sqrt(fmax(red[0], 0.0));
    double below = hn;
    for (int i = 0; i < j; ++i) {
        const double t = cs[i] * col[i] + sn[i] * col[i + 1];
        col[i + 1] = -sn[i] * col[i] + cs[i] * col[i + 1];
        col[i] = t;
    }
    const double den = hypot(col[j], below);
    cs[j] = den > 0.0 ? col[j] / den : 1.0;
    sn[j] = den > 0.0 ? below / den : 0.0;
    col[j] = den;
    g[j + 1] = -sn[j] * g[j];
    g[j] = cs[j] * g[j];
    fl[F_ITERS] += 1;
    fl[F_JCOUNT] = j + 1;
    const double norm = fabs(g[j + 1]);
    sc[S_NORM] = norm;
    sc[S_INV] = hn > 0.0 ? 1.0 / hn : 0.0;
    if (norm <= sc[S_TOL]) { fl[F_CONV] = 1; fl[F_DONE] = 1; }
    else if (hn == 0.0) { fl[F_BREAKDOWN] = 1; fl[F_DONE] = 1; }
    else if (fl[F_ITERS] >= fl[F_MAXIT]) fl[F_DONE] = 1;
}

// end of a cycle: back substitution R y = g over the columns completed in this cycle
__global__ void k_gmres_solve(int ldh, const double* __restrict__ R, const double* __restrict__ g,
                              const int* __restrict__ fl, double* __restrict__ y) {
    if (threadIdx.x != 0) return;
    const int jj = fl[F_JCOUNT];
    for (int i = jj - 1; i >= 0; --i) {
        double s = g[i];
        for (int q = i + 1; q < jj; ++q) s -= R[(int64_t)q * ldh + i] * y[q];
        y[i] = s / R[(int64_t)i * ldh + i];
    }
}

// ---------------------------------------------------------------------------------- bordered system (multipliers)
// out[j] += sum_i mu_i c_i[j] on the non-Dirichlet rows j < n;  mu_i = in[n + i]
__global__ void k_border_axpy(int64_t n, int nb, const double* __restrict__ C, const double* __restrict__ in,
                              double* __restrict__ out, const uint8_t* __restrict__ mask, const int* __restrict__ done) {
    if (done && *done) return;
    int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (j >= n || (mask && mask[j])) return;
    double acc = out[j];
    for (int i = 0; i < nb; ++i) acc += in[n + i] * C[(int64_t)i * n + j];
    out[j] = acc;
}

// partials[i][block] of c_i . in[0..n)
__global__ void __launch_bounds__(kRedThreads)
k_border_dots(int64_t n, int nb, const double* __restrict__ C, const double* __restrict__ in,
              double* __restrict__ partials, const int* __restrict__ done) {
    if (done && *done) return;
    __shared__ double sm[32];
    for (int i = 0; i < nb; ++i) {
        double s = 0.0;
        for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x)
            s += C[(int64_t)i * n + j] * in[j];
        s = block_reduce_sum(s, sm);
        if (threadIdx.x == 0) partials[i * gridDim.x + blockIdx.x] = s;
    }
}

// out[n + i] = c_i . in   (fixed summation order); the pad entries behind the multipliers stay zero
__global__ void __launch_bounds__(kRedThreads)
k_border_final(int64_t n, int nb, int npad, const double* __restrict__ partials, int nparts, double* __restrict__ out,
               const int* __restrict__ done) {
    if (done && *done) return;
    __shared__ double sm[32];
    for (int i = 0; i < npad; ++i) {
        double s = 0.0;
        if (i < nb) s = final_sum(partials + (int64_t)i * nparts, nparts, sm);
        if (threadIdx.x == 0) out[n + i] = s;
        __syncthreads();
    }
}

// z[n + i] = r[n + i] / s_i : the multiplier block of the preconditioner (Schur-complement scale s_i = c_i . B c_i)
__global__ void k_border_pc(int64_t n, int nb, int npad, const double* __restrict__ scale, const double* __restrict__ r,
                            double* __restrict__ z, const int* __restrict__ done) {
    if (done && *done) return;
    int i = threadIdx.x;
    if (i < npad) z[n + i] = i < nb ? r[n + i] / scale[i] : 0.0;
}

__global__ void k_copy_tail(int nb, int npad, const double* __restrict__ src, double* __restrict__ dst) {
    int i = threadIdx.x;
    if (i < npad) dst[i] = i < nb ? src[i] : 0.0;
}

}  // namespace

struct KrylovWork {
    int64_t n = 0;
    double *r = nullptr, *z = nullptr, *v = nullptr, *v_old = nullptr, *u = nullptr, *u_old = nullptr,
           *w1 = nullptr, *w2 = nullptr, *xi = nullptr, *bi = nullptr, *partials = nullptr, *sc = nullptr, *red = nullptr;
    int* fl = nullptr;
    double* basis = nullptr;   // GMRES: (restart + 1) vectors
    int basis_m = 0;
    double* hdev = nullptr;
    double* gm = nullptr;      // GMRES: R (121 x 120, column-major), then cs, sn, g (128 each)
    double* h_sc = nullptr;    // pinned mirrors
    int* h_fl = nullptr;
    // CUDA graphs of the MINRES iteration body: [0..3] its four segments (SpMV | halo + dot | preconditioner |
    // vector updates), launched one by one when the profiler brackets them with events, [4] the whole iteration
    cudaGraphExec_t graph[5] = {};
    int64_t graph_launches[5] = {};
    int64_t graph_key = -1;
};

static void drop_graphs(KrylovWork* k) {
    for (auto& g : k->graph)
        if (g) { cudaGraphExecDestroy(g); g = nullptr; }
    k->graph_key = -1;
}

static KrylovWork* get_work(mpet_ctx* ctx) {
    // all Krylov vectors live in the solver-internal layout (layout.cuh), followed -- when Lagrange multipliers
    // border the system -- by kBorderPad entries that hold the multipliers
    const int64_t need = ctx->Nint + (ctx->nb > 0 ? kBorderPad : 0);
    if (ctx->kw && ctx->kw->n == need) return ctx->kw;
    if (ctx->kw) {
        KrylovWork* o = ctx->kw;
        drop_graphs(o);
        double* old[] = {o->r, o->z, o->v, o->v_old, o->u, o->u_old, o->w1, o->w2, o->xi, o->bi, o->basis};
        for (double* q : old)
            if (q) dev_free(ctx, q);
        o->basis = nullptr;
        o->basis_m = 0;
    }
    KrylovWork* k = ctx->kw ? ctx->kw : new KrylovWork();
    const bool fresh = ctx->kw == nullptr;
    k->n = need;
    double** vecs[] = {&k->r, &k->z, &k->v, &k->v_old, &k->u, &k->u_old, &k->w1, &k->w2, &k->xi, &k->bi};
    for (auto p : vecs) {
        *p = dev_alloc<double>(ctx, need);
        CUDA_CHECK(cudaMemset(*p, 0, sizeof(double) * need));
    }
    if (!fresh) return k;
    k->partials = dev_alloc<double>(ctx, (int64_t)kRedBlocks * 8);
    k->sc = dev_alloc<double>(ctx, S_COUNT);
    k->red = dev_alloc<double>(ctx, 8);
    k->fl = dev_alloc<int>(ctx, F_COUNT);
    k->hdev = dev_alloc<double>(ctx, 256);
    k->gm = dev_alloc<double>(ctx, kGmresLd * kGmresMaxRestart + 3 * 128);
    CUDA_CHECK(cudaMallocHost(&k->h_sc, sizeof(double) * S_COUNT));
    CUDA_CHECK(cudaMallocHost(&k->h_fl, sizeof(int) * F_COUNT));
    ctx->kw = k;
    return k;
}

void krylov_free(mpet_ctx* ctx) {
    if (!ctx->kw) return;
    drop_graphs(ctx->kw);
    cudaFreeHost(ctx->kw->h_sc);
    cudaFreeHost(ctx->kw->h_fl);
    delete ctx->kw;
    ctx->kw = nullptr;
}

void pc_setup(mpet_ctx* ctx, cudaStream_t st) {
    MPET_REQUIRE(ctx->lhs_ready, "mpet_assemble_lhs must run before mpet_pc_setup");
    if (ctx->pc == 1) {
        if (!ctx->jac_dinv) ctx->jac_dinv = dev_alloc<double>(ctx, ctx->Nint);
        double* tmp = nullptr;
        CUDA_CHECK(cudaMalloc(&tmp, sizeof(double) * ctx->N));
        k_abs_diag_inv<<<grid_for(ctx->N, 256), 256, 0, st>>>(ctx->N, ctx->rowptr, ctx->cols, ctx->vals,
                                                              ctx->bc_mask, tmp);
        LAUNCH_CHECK(ctx);
        to_internal(ctx, tmp, ctx->jac_dinv, st);
        CUDA_CHECK(cudaStreamSynchronize(st));
        cudaFree(tmp);
    } else if (ctx->pc == 2) {
        amg_setup(ctx, st);
    }
}

static void pc_apply_flag(mpet_ctx* ctx, const double* r, double* z, const int* done, cudaStream_t st) {
    const int64_t n = ctx->Nint;   // r, z: solver-internal layout
    if (ctx->pc == 0) {
        CUDA_CHECK(cudaMemcpyAsync(z, r, sizeof(double) * n, cudaMemcpyDeviceToDevice, st));
    } else if (ctx->pc == 1) {
        MPET_REQUIRE(ctx->jac_dinv, "mpet_pc_setup must run before solving with pc = Jacobi");
        k_jacobi<<<grid_for(n, 256), 256, 0, st>>>(ctx->jac_dinv, r, n, z, done);
        LAUNCH_CHECK(ctx);
    } else {
        MPET_REQUIRE(ctx->amg_u, "mpet_pc_setup must run before solving with pc = AMG");
        amg_apply(ctx, r, z, done, st);
    }
}

static void pc_apply_dist(mpet_ctx* ctx, const double* r, double* z, const int* done, cudaStream_t st) {
    // the V-cycles exchange their own halos level by level (amg.cu); one more refresh covers Jacobi / none
    pc_apply_flag(ctx, r, z, done, st);
    if (dist_active(ctx) && ctx->pc != 2) dist_halo(ctx, DIST_PLAN_KRYLOV, z, false, done, st);
    if (ctx->nb > 0) {      // multiplier block: diagonal, Schur-complement scale (border_setup)
        k_border_pc<<<1, 32, 0, st>>>(ctx->Nint, ctx->nb, kBorderPad, ctx->border_scale, r, z, done);
        LAUNCH_CHECK(ctx);
    }
}

// y = K x for the (possibly bordered) system  K = [A C; C^T 0]  (mpetsolver.py:203-215: the Real-space Lagrange
// multipliers of the rigid motions / pressure constants are 6 + k dense rows and columns kept OUTSIDE the CSR
// matrix).  Dirichlet rows are identity rows; Krylov vectors vanish there, so the dots run over
// all entries while the column update skips the masked rows.
static void apply_operator(mpet_ctx* ctx, KrylovWork* k, const double* x, double* y, const uint8_t* mask,
                           const int* done, cudaStream_t st) {
    block_spmv(ctx, x, y, mask, done, st);
    if (ctx->nb == 0) return;
    const int64_t n = ctx->Nint;
    k_border_axpy<<<grid_for(n, 256), 256, 0, st>>>(n, ctx->nb, ctx->border, x, y, mask, done);
    LAUNCH_CHECK(ctx);
    const int G = 296;
    k_border_dots<<<G, kRedThreads, 0, st>>>(n, ctx->nb, ctx->border, x, k->partials, done);
    LAUNCH_CHECK(ctx);
    k_border_final<<<1, kRedThreads, 0, st>>>(n, ctx->nb, kBorderPad, k->partials, G, y, done);
    LAUNCH_CHECK(ctx);
}

static void ensure_scratch(mpet_ctx* ctx) {
    for (int i = 0; i < 2; ++i)
        if (!ctx->scratch_int[i]) ctx->scratch_int[i] = dev_alloc<double>(ctx, ctx->Nint);
}

// API-layout entry points (tests / measurement): convert, run the production kernels, convert back
// loose tolerances take the light P1-field cycles (amg.cu); MPET_PC_LIGHT=0/1 forces the mode
static bool light_tolerance(const mpet_ctx* ctx) {
    static const int forced = []() { const char* e = getenv("MPET_PC_LIGHT"); return e ? (e[0] == '1' ? 1 : 0) : -1; }();
    return forced >= 0 ? forced == 1 : ctx->rtol >= 1e-8;
}

void pc_apply(mpet_ctx* ctx, const double* r, double* z, cudaStream_t st) {
    ctx->pc_light = light_tolerance(ctx);
    ensure_scratch(ctx);
    to_internal(ctx, r, ctx->scratch_int[0], st);
    pc_apply_flag(ctx, ctx->scratch_int[0], ctx->scratch_int[1], nullptr, st);
    to_api(ctx, ctx->scratch_int[1], z, st);
}

void spmv_api(mpet_ctx* ctx, const double* x, double* y, cudaStream_t st) {
    ensure_scratch(ctx);
    to_internal(ctx, x, ctx->scratch_int[0], st);
    block_spmv(ctx, ctx->scratch_int[0], ctx->scratch_int[1], nullptr, nullptr, st);
    to_api(ctx, ctx->scratch_int[1], y, st);
}


// partials -> k->red[0] (fixed order), then the NCCL all-reduce over ranks when a communicator is attached
static void finish_reduction(mpet_ctx* ctx, KrylovWork* k, const int* done, cudaStream_t st) {
    if (dist_reduce_partials(ctx, k->partials, kRedBlocks, 1, k->red, done, st)) return;   // one fused launch
    k_final_store<<<1, kRedThreads, 0, st>>>(k->partials, kRedBlocks, k->red, done);
    LAUNCH_CHECK(ctx);
    dist_allreduce_sum(ctx, k->red, 1, st);
}

static void dot_to(mpet_ctx* ctx, KrylovWork* k, const double* a, const double* b, const int* done, cudaStream_t st) {
    k_dot_partial<<<kRedBlocks, kRedThreads, 0, st>>>(a, b, k->n, k->partials, dist_owned_mask(ctx), done);
    LAUNCH_CHECK(ctx);
    finish_reduction(ctx, k, done, st);
}

static void pc_apply_dist(mpet_ctx* ctx, const double* r, double* z, const int* done, cudaStream_t st);


// xi (internal layout) carries the Dirichlet values; r = b - A x on free rows, 0 on Dirichlet rows
static void initial_residual(mpet_ctx* ctx, KrylovWork* k, cudaStream_t st) {
    if (ctx->n_bc > 0) scatter_bc_values_int(ctx, k->xi, st);
    dist_halo(ctx, DIST_PLAN_KRYLOV, k->xi, false, nullptr, st);      // ghosts of x and b come from their owners
    dist_halo(ctx, DIST_PLAN_KRYLOV, k->bi, false, nullptr, st);
    apply_operator(ctx, k, k->xi, k->v, nullptr, nullptr, st);
    k_residual<<<grid_for(k->n, 256), 256, 0, st>>>(k->bi, k->v, ctx->n_bc > 0 ? ctx->bc_mask_int : nullptr, k->n, k->r,
                                                    nullptr);
    LAUNCH_CHECK(ctx);
    dist_halo(ctx, DIST_PLAN_KRYLOV, k->r, false, nullptr, st);       // ghost rows of the local matrix are incomplete
}

// k->r = right-hand side after apply_symmetric(bc, A, b) (bc_symmetric.py:11-22): b - A[:, D] x_D on the free rows,
// the boundary values on the Dirichlet rows.  PETSc measures convergence against ITS preconditioned norm when the
// initial guess is nonzero (KSPConvergedDefault [EXT]); bi must be set, k->w1 / k->v are scratch.
static void eliminated_rhs(mpet_ctx* ctx, KrylovWork* k, cudaStream_t st) {
    if (ctx->n_bc > 0) {
        CUDA_CHECK(cudaMemsetAsync(k->w1, 0, sizeof(double) * k->n, st));
        scatter_bc_values_int(ctx, k->w1, st);
        dist_halo(ctx, DIST_PLAN_KRYLOV, k->w1, false, nullptr, st);
        dist_halo(ctx, DIST_PLAN_KRYLOV, k->bi, false, nullptr, st);
        apply_operator(ctx, k, k->w1, k->v, nullptr, nullptr, st);
        k_residual<<<grid_for(k->n, 256), 256, 0, st>>>(k->bi, k->v, ctx->bc_mask_int, k->n, k->r, k->w1);
        LAUNCH_CHECK(ctx);
    } else {
        dist_halo(ctx, DIST_PLAN_KRYLOV, k->bi, false, nullptr, st);
        CUDA_CHECK(cudaMemcpyAsync(k->r, k->bi, sizeof(double) * k->n, cudaMemcpyDeviceToDevice, st));
    }
    dist_halo(ctx, DIST_PLAN_KRYLOV, k->r, false, nullptr, st);
}

// API vectors carry the multipliers behind the N finite-element dofs (as the reference's mixed space does)
static void load_system(mpet_ctx* ctx, KrylovWork* k, const double* b, const double* x, cudaStream_t st) {
    to_internal(ctx, b, k->bi, st);
    to_internal(ctx, x, k->xi, st);
    if (ctx->nb > 0) {
        k_copy_tail<<<1, 32, 0, st>>>(ctx->nb, kBorderPad, b + ctx->N, k->bi + ctx->Nint);
        LAUNCH_CHECK(ctx);
        k_copy_tail<<<1, 32, 0, st>>>(ctx->nb, kBorderPad, x + ctx->N, k->xi + ctx->Nint);
        LAUNCH_CHECK(ctx);
    }
}

static void store_solution(mpet_ctx* ctx, KrylovWork* k, double* x, cudaStream_t st) {
    to_api(ctx, k->xi, x, st);
    if (ctx->nb > 0)
        CUDA_CHECK(cudaMemcpyAsync(x + ctx->N, k->xi + ctx->Nint, sizeof(double) * ctx->nb, cudaMemcpyDeviceToDevice, st));
}

// Schur-complement scale of the multiplier block of the preconditioner: s_i = c_i . B c_i (once per hierarchy)
static void border_setup(mpet_ctx* ctx, KrylovWork* k, cudaStream_t st) {
    if (ctx->nb == 0 || ctx->border_scaled) return;
    MPET_REQUIRE(!dist_active(ctx), "Lagrange-multiplier borders are single-GPU for now");
    if (!ctx->border_scale) ctx->border_scale = dev_alloc<double>(ctx, kBorderPad);
    std::vector<double> s(kBorderPad, 1.0);
    CUDA_CHECK(cudaMemcpyAsync(ctx->border_scale, s.data(), sizeof(double) * kBorderPad, cudaMemcpyHostToDevice, st));
    const int64_t n = ctx->Nint;
    for (int i = 0; i < ctx->nb; ++i) {
        CUDA_CHECK(cudaMemsetAsync(k->v, 0, sizeof(double) * k->n, st));
        CUDA_CHECK(cudaMemcpyAsync(k->v, ctx->border + (int64_t)i * n, sizeof(double) * n, cudaMemcpyDeviceToDevice, st));
        if (ctx->n_bc > 0) {       // the columns act on the free rows only
            k_residual<<<grid_for(n, 256), 256, 0, st>>>(k->v, k->u_old, ctx->bc_mask_int, n, k->v, nullptr);
            LAUNCH_CHECK(ctx);
        }
        pc_apply_flag(ctx, k->v, k->u, nullptr, st);
        k_dot_partial<<<kRedBlocks, kRedThreads, 0, st>>>(k->v, k->u, n, k->partials, nullptr, nullptr);
        LAUNCH_CHECK(ctx);
        k_final_store<<<1, kRedThreads, 0, st>>>(k->partials, kRedBlocks, k->red, nullptr);
        LAUNCH_CHECK(ctx);
        CUDA_CHECK(cudaMemcpyAsync(&s[i], k->red, sizeof(double), cudaMemcpyDeviceToHost, st));
        CUDA_CHECK(cudaStreamSynchronize(st));
        MPET_REQUIRE(s[i] > 0.0, "a multiplier column is zero or the preconditioner is not positive on it");
    }
    CUDA_CHECK(cudaMemcpyAsync(ctx->border_scale, s.data(), sizeof(double) * kBorderPad, cudaMemcpyHostToDevice, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
    ctx->border_scaled = true;
}

static void minres(mpet_ctx* ctx, const double* b, double* x, double* info, cudaStream_t st) {
    KrylovWork* k = get_work(ctx);
    const int64_t n = k->n;
    const uint8_t* mask = ctx->n_bc > 0 ? ctx->bc_mask_int : nullptr;
    CUDA_CHECK(cudaMemsetAsync(k->u_old, 0, sizeof(double) * k->n, st));
    border_setup(ctx, k, st);
    load_system(ctx, k, b, x, st);
    eliminated_rhs(ctx, k, st);                              // reference norm sqrt(b . B b) -> red[1]
    pc_apply_dist(ctx, k->r, k->z, nullptr, st);
    dot_to(ctx, k, k->r, k->z, nullptr, st);
    CUDA_CHECK(cudaMemcpyAsync(k->red + 1, k->red, sizeof(double), cudaMemcpyDeviceToDevice, st));
    initial_residual(ctx, k, st);
    pc_apply_dist(ctx, k->r, k->z, nullptr, st);
    dot_to(ctx, k, k->r, k->z, nullptr, st);
    k_minres_init<<<1, 32, 0, st>>>(k->red, ctx->rtol, ctx->atol, ctx->maxit, ctx->norm_mode, k->sc, k->fl);
    LAUNCH_CHECK(ctx);
    const int* done = k->fl + F_DONE;
    k_minres_start<<<grid_for(n, 256), 256, 0, st>>>(n, k->sc, k->r, k->z, k->v, k->u, k->v_old, k->u_old, k->w1,
                                                     k->w2, done);
    LAUNCH_CHECK(ctx);
    const int check_every = 5;
    int enq = 0;
    // the four segments of one iteration (fixed pointers, scalars on the device: every iteration enqueues the same work)
    auto seg = [&](int which) {
        switch (which) {
            case 0:
                apply_operator(ctx, k, k->u, k->r, mask, done, st);
                break;
            case 1:
                dist_halo(ctx, DIST_PLAN_KRYLOV, k->r, false, done, st);
                dot_to(ctx, k, k->r, k->u, done, st);
                k_minres_alpha<<<1, 32, 0, st>>>(k->red, k->sc, k->fl);
                LAUNCH_CHECK(ctx);
                break;
            case 2:
                pc_apply_dist(ctx, k->r, k->z, done, st);
                break;
            default:
                k_minres_update<<<kRedBlocks, kRedThreads, 0, st>>>(n, k->sc, k->v, k->v_old, k->u, k->u_old, k->r,
                                                                    k->z, k->partials, dist_owned_mask(ctx), done);
                LAUNCH_CHECK(ctx);
                finish_reduction(ctx, k, done, st);
                k_minres_rotate<<<1, 32, 0, st>>>(k->red, k->sc, k->fl);
                LAUNCH_CHECK(ctx);
                k_minres_finalize<<<grid_for(n, 256), 256, 0, st>>>(n, k->sc, k->xi, k->w1, k->w2, k->v, k->v_old, k->u,
                                                                    k->u_old, k->r, k->z, done);
                LAUNCH_CHECK(ctx);
                k_commit<<<1, 1, 0, st>>>(k->fl);
                LAUNCH_CHECK(ctx);
                break;
        }
    };
    // Graph capture: ~160 launches per iteration collapse into 1 (or 4, when the profiler times SpMV and the
    // preconditioner) graph launches.  The first iteration of every solve runs eagerly (lazy one-time kernel
    // attributes, side streams); graphs are re-captured when anything baked into the launches changed
    // (hierarchy, Dirichlet set, peer arena: ctx->graph_epoch).  NCCL halos are not captured.
    static const bool want_graphs = []() { const char* e = getenv("MPET_GRAPHS"); return !(e && e[0] == '0'); }();
    const bool can_graph = want_graphs && (!dist_active(ctx) || dist_comm_kind(ctx) == 2);
    const int64_t key = ctx->graph_epoch * 16 + ctx->pc * 2 + (ctx->pc_light ? 1 : 0);
    if (k->graph_key != key) drop_graphs(k);
    auto capture = [&](int slot) {
        ctx->prof_suspended = true;
        const int64_t l0 = ctx->launches;
        CUDA_CHECK(cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed));
        try {
            if (slot < 4) seg(slot);
            else for (int w = 0; w < 4; ++w) seg(w);
        } catch (...) {
            cudaGraph_t junk = nullptr;
            cudaStreamEndCapture(st, &junk);
            if (junk) cudaGraphDestroy(junk);
            ctx->prof_suspended = false;
            throw;
        }
        cudaGraph_t g = nullptr;
        CUDA_CHECK(cudaStreamEndCapture(st, &g));
        ctx->prof_suspended = false;
        k->graph_launches[slot] = ctx->launches - l0;
        ctx->launches = l0;
        CUDA_CHECK(cudaGraphInstantiate(&k->graph[slot], g, 0));
        cudaGraphDestroy(g);
        k->graph_key = key;
    };
    auto run = [&](int slot) {
        if (!k->graph[slot]) capture(slot);
        CUDA_CHECK(cudaGraphLaunch(k->graph[slot], st));
        ctx->launches += k->graph_launches[slot];
    };
    while (true) {
        for (int q = 0; q < check_every && enq < ctx->maxit; ++q, ++enq) {
            const bool eager = !can_graph || enq == 0;
            if (!eager && !ctx->prof.on) { run(4); continue; }
            cudaEvent_t pe = prof_begin(ctx, st);
            if (eager) seg(0); else run(0);
            prof_end(ctx, PROF_SPMV, pe, st);
            if (eager) seg(1); else run(1);
            pe = prof_begin(ctx, st);
            if (eager) seg(2); else run(2);
            prof_end(ctx, PROF_PC, pe, st);
            if (eager) seg(3); else run(3);
        }
        CUDA_CHECK(cudaMemcpyAsync(k->h_fl, k->fl, sizeof(int) * F_COUNT, cudaMemcpyDeviceToHost, st));
        CUDA_CHECK(cudaMemcpyAsync(k->h_sc, k->sc, sizeof(double) * S_COUNT, cudaMemcpyDeviceToHost, st));
        CUDA_CHECK(cudaStreamSynchronize(st));
        if (k->h_fl[F_DONE] || enq >= ctx->maxit) break;
    }
    store_solution(ctx, k, x, st);
    CUDA_CHECK(cudaStreamSynchronize(st));
    dist_check(ctx);
    prof_collect(ctx);
    info[0] = (double)k->h_fl[F_ITERS];
    info[1] = (double)k->h_fl[F_CONV];
    info[2] = k->h_sc[S_BNORM] > 0 ? k->h_sc[S_NORM] / k->h_sc[S_BNORM] : 0.0;
    info[3] = k->h_sc[S_NORM0];
    info[4] = (double)k->h_fl[F_BREAKDOWN];
    // KSPConvergedReason-style code: 2 rtol/atol reached, -3 iteration limit, -4 breakdown, -5 stagnation
    info[5] = k->h_fl[F_CONV] ? 2.0 : (k->h_fl[F_BREAKDOWN] ? -4.0 : (k->h_fl[F_STAG] ? -5.0 : -3.0));
    info[6] = k->h_sc[S_BNORM];
    info[7] = k->h_sc[S_BTRUE];
}

// ---------------------------------------------------------------------------------- GMRES(m)
static void multidot(mpet_ctx* ctx, KrylovWork* k, const double* V, int64_t ldv, int nv, const double* w,
                     double* hdev, int accumulate, cudaStream_t st) {
    const int G = 296;
    const uint8_t* own = dist_owned_mask(ctx);
    for (int k0 = 0; k0 < nv; k0 += 4) {
        int c = std::min(4, nv - k0);
        const double* Vk = V + (int64_t)k0 * ldv;
        switch (c) {
            case 1: k_multidot<1><<<G, kRedThreads, 0, st>>>(Vk, ldv, w, k->n, k->partials, own); break;
            case 2: k_multidot<2><<<G, kRedThreads, 0, st>>>(Vk, ldv, w, k->n, k->partials, own); break;
            case 3: k_multidot<3><<<G, kRedThreads, 0, st>>>(Vk, ldv, w, k->n, k->partials, own); break;
            default: k_multidot<4><<<G, kRedThreads, 0, st>>>(Vk, ldv, w, k->n, k->partials, own); break;
        }
        LAUNCH_CHECK(ctx);
        k_multifinal<<<c, kRedThreads, 0, st>>>(k->partials, G, hdev + k0, accumulate);
        LAUNCH_CHECK(ctx);
    }
    dist_allreduce_sum(ctx, hdev, nv, st);       // no-op on one GPU (accumulate is only used with 0)
}

static void gmres(mpet_ctx* ctx, const double* b, double* x, double* info, cudaStream_t st) {
    // Left-preconditioned GMRES(m), classical Gram-Schmidt in two passes (CGS2).  Everything scalar -- Hessenberg
    // column, Givens rotations, residual estimate, convergence / iteration-limit test, the triangular solve at the
    // end of a cycle -- runs in single-thread kernels on the device; once the DONE flag is up all queued kernels
    // return immediately and the host polls the flags every few iterations, like MINRES.
    KrylovWork* k = get_work(ctx);
    const int64_t n = k->n;
    const int m = ctx->restart;
    MPET_REQUIRE(m >= 1 && m <= kGmresMaxRestart, "GMRES restart must be in 1..120");
    const int64_t ldv = (n + 3) & ~(int64_t)3;   // 32-byte aligned basis vectors (256-bit loads in block_spmv)
    if (k->basis_m < m) {
        k->basis = dev_alloc<double>(ctx, (int64_t)(m + 1) * ldv);
        k->basis_m = m;
    }
    const uint8_t* mask = ctx->n_bc > 0 ? ctx->bc_mask_int : nullptr;
    double* V = k->basis;
    double *R = k->gm, *cs = k->gm + (int64_t)kGmresLd * kGmresMaxRestart, *sn = cs + 128, *g = sn + 128;
    const int* done = k->fl + F_DONE;
    const int gridn = grid_for(n, 256);
    CUDA_CHECK(cudaMemsetAsync(k->u_old, 0, sizeof(double) * k->n, st));
    border_setup(ctx, k, st);
    load_system(ctx, k, b, x, st);
    // PETSc default with a nonzero initial guess: tolerance relative to ||B b||_2 (left preconditioning)
    eliminated_rhs(ctx, k, st);
    pc_apply_dist(ctx, k->r, k->z, nullptr, st);
    dot_to(ctx, k, k->z, k->z, nullptr, st);
    k_gmres_init<<<1, 32, 0, st>>>(k->red, ctx->maxit, k->sc, k->fl);
    LAUNCH_CHECK(ctx);
    auto poll = [&]() {
        CUDA_CHECK(cudaMemcpyAsync(k->h_fl, k->fl, sizeof(int) * F_COUNT, cudaMemcpyDeviceToHost, st));
        CUDA_CHECK(cudaMemcpyAsync(k->h_sc, k->sc, sizeof(double) * S_COUNT, cudaMemcpyDeviceToHost, st));
        CUDA_CHECK(cudaStreamSynchronize(st));
        return k->h_fl[F_DONE] != 0;
    };
    const int check_every = 5;
    int enq = 0;
    while (true) {
        initial_residual(ctx, k, st);
        pc_apply_dist(ctx, k->r, k->z, nullptr, st);
        dot_to(ctx, k, k->z, k->z, nullptr, st);          // k->red[0]: summed over owned dofs and all ranks
        k_gmres_begin<<<1, 32, 0, st>>>(k->red, ctx->rtol, ctx->atol, ctx->norm_mode, m, k->sc, k->fl, g);
        LAUNCH_CHECK(ctx);
        if (poll()) break;                                // converged (or out of iterations) at the start of a cycle
        k_scale_copy<<<gridn, 256, 0, st>>>(k->z, k->sc + S_INV, n, V, done);
        LAUNCH_CHECK(ctx);
        bool stop = false;
        for (int j = 0; j < m && enq < ctx->maxit && !stop; ++j) {
            double* w = V + (int64_t)(j + 1) * ldv;
            cudaEvent_t pe = prof_begin(ctx, st);
            apply_operator(ctx, k, V + (int64_t)j * ldv, k->r, mask, done, st);
            prof_end(ctx, PROF_SPMV, pe, st);
            dist_halo(ctx, DIST_PLAN_KRYLOV, k->r, false, done, st);      // ghost rows of the local matrix are incomplete
            pe = prof_begin(ctx, st);
            pc_apply_dist(ctx, k->r, w, done, st);
            prof_end(ctx, PROF_PC, pe, st);
            multidot(ctx, k, V, ldv, j + 1, w, k->hdev, 0, st);
            k_multi_axpy<<<gridn, 256, 0, st>>>(V, ldv, j + 1, k->hdev, -1.0, n, w);
            LAUNCH_CHECK(ctx);
            multidot(ctx, k, V, ldv, j + 1, w, k->hdev + 128, 0, st);
            k_multi_axpy<<<gridn, 256, 0, st>>>(V, ldv, j + 1, k->hdev + 128, -1.0, n, w);
            LAUNCH_CHECK(ctx);
            dot_to(ctx, k, w, w, nullptr, st);
            k_gmres_rotate<<<1, 32, 0, st>>>(j, kGmresLd, k->hdev, k->hdev + 128, k->red, R, cs, sn, g, k->sc, k->fl);
            LAUNCH_CHECK(ctx);
            k_scale_copy<<<gridn, 256, 0, st>>>(w, k->sc + S_INV, n, w, done);
            LAUNCH_CHECK(ctx);
            ++enq;
            if ((j + 1) % check_every == 0 || j + 1 == m || enq >= ctx->maxit) stop = poll();
        }
        // x += V y with R y = g over the columns this cycle completed (count on the device)
        k_gmres_solve<<<1, 32, 0, st>>>(kGmresLd, R, g, k->fl, k->hdev);
        LAUNCH_CHECK(ctx);
        k_multi_axpy_count<<<gridn, 256, 0, st>>>(V, ldv, k->fl + F_JCOUNT, k->hdev, n, k->xi);
        LAUNCH_CHECK(ctx);
        if (stop || enq >= ctx->maxit) break;
    }
    store_solution(ctx, k, x, st);
    poll();
    dist_check(ctx);
    prof_collect(ctx);
    info[0] = (double)k->h_fl[F_ITERS];
    info[1] = (double)k->h_fl[F_CONV];
    info[2] = k->h_sc[S_BNORM] > 0 ? k->h_sc[S_NORM] / k->h_sc[S_BNORM] : 0.0;
    info[3] = k->h_sc[S_NORM0] < 0 ? 0.0 : k->h_sc[S_NORM0];
    info[4] = (double)k->h_fl[F_BREAKDOWN];
    info[5] = k->h_fl[F_CONV] ? 2.0 : (k->h_fl[F_BREAKDOWN] ? -4.0 : -3.0);
    info[6] = k->h_sc[S_BNORM];
    info[7] = k->h_sc[S_BTRUE];
}

void krylov_solve(mpet_ctx* ctx, const double* b, double* x, double* info, cudaStream_t st) {
    MPET_REQUIRE(ctx->lhs_ready, "mpet_assemble_lhs must run before mpet_solve");
    for (int i = 0; i < 8; ++i) info[i] = 0.0;
    ctx->pc_light = light_tolerance(ctx);
    if (st == nullptr || st == cudaStreamLegacy || st == cudaStreamPerThread) {
        // the legacy default stream cannot be captured into a CUDA graph: solve on a stream of our own, ordered
        // after everything the caller has queued (the solve is synchronous, so nothing has to be ordered after it)
        if (!ctx->solve_stream) {
            CUDA_CHECK(cudaStreamCreateWithFlags(&ctx->solve_stream, cudaStreamNonBlocking));
            CUDA_CHECK(cudaEventCreateWithFlags(&ctx->solve_event, cudaEventDisableTiming));
        }
        CUDA_CHECK(cudaEventRecord(ctx->solve_event, st));
        CUDA_CHECK(cudaStreamWaitEvent(ctx->solve_stream, ctx->solve_event, 0));
        st = ctx->solve_stream;
    }
    if (ctx->method == 0) minres(ctx, b, x, info, st);
    else gmres(ctx, b, x, info, st);
}

