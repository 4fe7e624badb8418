// Batched fp64 element kernels + deterministic gather assembly for the MPET block system.
//
// Replaces assemble(a) / assemble(prec) and the FFC-generated tabulate_tensor kernels
// (mpetsolver.py:196-201,260,264-277,335,412,496,502).  Element-level formulas: SURVEY.md appendix A.
//
// Layout/algorithm: instead of forming the (30+4A)^2 element matrix per tet and scattering it, the
// kernels run over the ENTRIES of the scalar node graphs (graph.cu).  For one entry (node a, node b)
// the thread walks the entry's gather list (the tets containing both nodes, ascending cell id),
// contracts the reference tables staged in shared memory with the tet's inverse Jacobian, and sums in
// that fixed order.  One P2xP2 entry yields the 3x3 tensor G = int grad(phi_a) (x) grad(phi_b), from
// which all nine displacement-block values follow; one P2xP1 entry yields D_k = int d_k(phi_a) psi_m
// for the 2*3*A coupling values; one P1xP1 entry yields mass + stiffness for the A^2 pressure values.
// Every CSR value is written exactly once (no atomics, bit-reproducible); element matrices are never
// materialised in HBM.
#include "ctx.h"
#include <map>
#include <algorithm>
#include <array>
#include <cmath>

namespace {

__constant__ double c_R22[100 * 9];   // int_ref dX_m(phi_a) dX_n(phi_b)
__constant__ double c_Q21[40 * 3];    // int_ref dX_j(phi_a) psi_m
__constant__ double c_M22[100];       // int_ref phi_a phi_b
__constant__ double c_M11[16];        // int_ref psi_m psi_n
__constant__ double c_GL[12];         // grad_X lambda_m

// ---- host: exact integrals of polynomials in barycentric coordinates --------------------------
typedef std::map<std::array<int, 4>, double> Poly;

Poly pmul(const Poly& a, const Poly& b) {
    Poly r;
    for (auto& x : a)
        for (auto& y : b) {
            std::array<int, 4> e;
            for (int i = 0; i < 4; ++i) e[i] = x.first[i] + y.first[i];
            r[e] += x.second * y.second;
        }
    return r;
}
Poly padd(const Poly& a, const Poly& b, double sb = 1.0) {
    Poly r = a;
    for (auto& y : b) r[y.first] += sb * y.second;
    return r;
}
Poly pscale(const Poly& a, double s) {
    Poly r;
    for (auto& x : a) r[x.first] = s * x.second;
    return r;
}
Poly lam(int i) {
    std::array<int, 4> e = {0, 0, 0, 0};
    e[i] = 1;
    return Poly{{e, 1.0}};
}
Poly pconst(double c) { return Poly{{{0, 0, 0, 0}, c}}; }
double fact(int n) { double f = 1; for (int i = 2; i <= n; ++i) f *= i; return f; }
double pint(const Poly& p) {  // int over the reference tet: a! b! c! d! / (a+b+c+d+3)!
    double s = 0;
    for (auto& x : p) {
        int t = 0; double num = 1;
        for (int i = 0; i < 4; ++i) { t += x.first[i]; num *= fact(x.first[i]); }
        s += x.second * num / fact(t + 3);
    }
    return s;
}

const int h_ledge[6][2] = {{2, 3}, {1, 3}, {1, 2}, {0, 3}, {0, 2}, {0, 1}};
const double h_GL[4][3] = {{-1, -1, -1}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}};

// P2 basis a as polynomial in lambda, and its reference gradient components
Poly phi2(int a) {
    if (a < 4) return padd(pscale(pmul(lam(a), lam(a)), 2.0), lam(a), -1.0);
    return pscale(pmul(lam(h_ledge[a - 4][0]), lam(h_ledge[a - 4][1])), 4.0);
}
Poly dphi2(int a, int m) {
    if (a < 4) return pscale(padd(pscale(lam(a), 4.0), pconst(1.0), -1.0), h_GL[a][m]);
    int p = h_ledge[a - 4][0], q = h_ledge[a - 4][1];
    return padd(pscale(lam(p), 4.0 * h_GL[q][m]), pscale(lam(q), 4.0 * h_GL[p][m]));
}

bool g_tables_ready = false;

// ---- device ------------------------------------------------------------------------------------
// coefficient tables travel as a kernel argument (by value): no per-assembly cudaMalloc / copy / sync
struct AsmCoefs {
    double cup[MPET_MAX_NETWORKS], cpu[MPET_MAX_NETWORKS], cl[MPET_MAX_NETWORKS];
    double cm[MPET_MAX_NETWORKS * MPET_MAX_NETWORKS];
};
// DG0 permeability: per-cell values kc[f] (null: constant) and where the weighted stiffness of field f is kept
struct CellCoefs {
    const double* kc[MPET_MAX_NETWORKS];
    double* lw[MPET_MAX_NETWORKS];
};

__global__ void k_geometry(const double* __restrict__ coords, const int32_t* __restrict__ cells,
                           int64_t nc, double* __restrict__ geom) {
    int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= nc) return;
    double x[4][3];
#pragma unroll
    for (int v = 0; v < 4; ++v) {
        int64_t id = cells[c * 4 + v];
#pragma unroll
        for (int d = 0; d < 3; ++d) x[v][d] = coords[id * 3 + d];
    }
    double J[3][3];   // J[i][j] = d x_i / d X_j
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) J[i][j] = x[j + 1][i] - x[0][i];
    double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1];
    double c01 = J[1][2] * J[2][0] - J[1][0] * J[2][2];
    double c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
    double det = J[0][0] * c00 + J[0][1] * c01 + J[0][2] * c02;
    double id = 1.0 / det;
    double* g = geom + c * 10;
    // Jinv[k][m] = dX_k / dx_m = adj(J)[k][m] / det
    g[0] = c00 * id;
    g[1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) * id;
    g[2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) * id;
    g[3] = c01 * id;
    g[4] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) * id;
    g[5] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) * id;
    g[6] = c02 * id;
    g[7] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) * id;
    g[8] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * id;
    g[9] = fabs(det);
}

__device__ __forceinline__ void load_geom(const double* __restrict__ geom, uint32_t c, double* Ji,
                                          double& det) {
    const double2* g = reinterpret_cast<const double2*>(geom + (size_t)c * 10);
    double2 a = __ldg(g), b = __ldg(g + 1), cc = __ldg(g + 2), d = __ldg(g + 3), e = __ldg(g + 4);
    Ji[0] = a.x; Ji[1] = a.y; Ji[2] = b.x; Ji[3] = b.y; Ji[4] = cc.x; Ji[5] = cc.y;
    Ji[6] = d.x; Ji[7] = d.y; Ji[8] = e.x; det = e.y;
}

// One warp per P2 node row; lanes stride over the row's entries.
// BLOCKS: write the nine displacement-block values of A.  STIFF: write tr(G) (scalar stiffness).
// MASS: write the scalar P2 mass.
template <bool BLOCKS, bool STIFF, bool MASS>
__global__ void __launch_bounds__(256)
k_asm22(const int32_t* __restrict__ rp22, const int32_t* __restrict__ gptr,
        const uint32_t* __restrict__ glist, const double* __restrict__ geom,
        const int64_t* __restrict__ rowptr, int64_t n2, double mu, double lmbda,
        double* __restrict__ vals, double* __restrict__ k22, double* __restrict__ m22) {
    __shared__ double sR[900];
    __shared__ double sM[100];
    for (int i = threadIdx.x; i < 900; i += blockDim.x) sR[i] = c_R22[i];
    for (int i = threadIdx.x; i < 100; i += blockDim.x) sM[i] = c_M22[i];
    __syncthreads();
    // persistent CTAs: the reference tables are staged once per CTA, warps stride over the node rows
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int lane = threadIdx.x & 31;
    for (int64_t a = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5; a < n2; a += nwarps) {
        const int32_t e0 = rp22[a], d22 = rp22[a + 1] - e0;
        int64_t base[3];
        if (BLOCKS) {
#pragma unroll
            for (int k = 0; k < 3; ++k) base[k] = rowptr[k * n2 + a];
        }
        for (int j = lane; j < d22; j += 32) {
            const int32_t e = e0 + j;
            double G[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
            double mass = 0.0;
            int32_t g = gptr[e];
            const int32_t g1 = gptr[e + 1];
            uint32_t id = g < g1 ? glist[g] : 0u;            // next contribution id is fetched one step ahead
            for (; g < g1; ++g) {
                const uint32_t idn = (g + 1 < g1) ? glist[g + 1] : 0u;
                const uint32_t c = id / 100u;
                const int p = (int)(id - c * 100u);
                double Ji[9], det;
                load_geom(geom, c, Ji, det);
                if (BLOCKS || STIFF) {
                    const double* R = sR + p * 9;
                    double T[9];
#pragma unroll
                    for (int k = 0; k < 3; ++k)
#pragma unroll
                        for (int n = 0; n < 3; ++n)
                            T[k * 3 + n] = R[k * 3] * Ji[n] + R[k * 3 + 1] * Ji[3 + n] + R[k * 3 + 2] * Ji[6 + n];
#pragma unroll
                    for (int m = 0; m < 3; ++m)
#pragma unroll
                        for (int n = 0; n < 3; ++n)
                            G[m * 3 + n] += det * (Ji[m] * T[n] + Ji[3 + m] * T[3 + n] + Ji[6 + m] * T[6 + n]);
                }
                if (MASS) mass += det * sM[p];
                id = idn;
            }
            const double tr = G[0] + G[4] + G[8];
            if (BLOCKS) {
#pragma unroll
                for (int k = 0; k < 3; ++k)
#pragma unroll
                    for (int l = 0; l < 3; ++l) {
                        double v = mu * G[l * 3 + k] + lmbda * G[k * 3 + l];
                        if (k == l) v += mu * tr;
                        vals[base[k] + (int64_t)l * d22 + j] = v;
                    }
            }
            if (STIFF) k22[e] = tr;
            if (MASS) m22[e] = mass;
        }
    }
}

__device__ __forceinline__ int64_t row_of_entry(const int32_t* __restrict__ rp, int64_t nrows, int32_t e) {
    int64_t lo = 0, hi = nrows;      // largest r with rp[r] <= e
    while (hi - lo > 1) {
        int64_t mid = (lo + hi) >> 1;
        if (rp[mid] <= e) lo = mid; else hi = mid;
    }
    return lo;
}

// One thread per P2xP1 entry (node a, vertex v): D_k = int d_k(phi_a) psi_m; writes the "up" and
// "pu" coupling values of every network.
__global__ void __launch_bounds__(256)
k_asm21(const int32_t* __restrict__ rp21, const int32_t* __restrict__ col21,
        const int32_t* __restrict__ gptr, const uint32_t* __restrict__ glist,
        const double* __restrict__ geom, const int32_t* __restrict__ rp22,
        const int32_t* __restrict__ rp12, const int32_t* __restrict__ t21to12,
        const int64_t* __restrict__ rowptr, int64_t n2, int64_t nv, int64_t nnz21, int A,
        const AsmCoefs K, double* __restrict__ vals) {
    const double* cup = K.cup;
    const double* cpu = K.cpu;
    __shared__ double sQ[120];
    for (int i = threadIdx.x; i < 120; i += blockDim.x) sQ[i] = c_Q21[i];
    __syncthreads();
    int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= nnz21) return;
    int64_t a = row_of_entry(rp21, n2, (int32_t)e);
    int32_t v = col21[e];
    double D[3] = {0, 0, 0};
    int32_t g1 = gptr[e + 1];
    for (int32_t g = gptr[e]; g < g1; ++g) {
        uint32_t id = glist[g];
        uint32_t c = id / 40u;
        int p = (int)(id - c * 40u);
        double Ji[9], det;
        load_geom(geom, c, Ji, det);
        const double* Q = sQ + p * 3;
#pragma unroll
        for (int k = 0; k < 3; ++k) D[k] += det * (Q[0] * Ji[k] + Q[1] * Ji[3 + k] + Q[2] * Ji[6 + k]);
    }
    int32_t d22 = rp22[a + 1] - rp22[a];
    int32_t d21 = rp21[a + 1] - rp21[a];
    int32_t j = (int32_t)e - rp21[a];
    int32_t d12 = rp12[v + 1] - rp12[v];
    int32_t jt = t21to12[e] - rp12[v];
    for (int i = 0; i < A; ++i) {
        const double au = cup[i], ap = cpu[i];
        int64_t prow = rowptr[3 * n2 + i * nv + v];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            vals[rowptr[k * n2 + a] + 3 * (int64_t)d22 + (int64_t)i * d21 + j] = au * D[k];
            vals[prow + (int64_t)k * d12 + jt] = ap * D[k];
        }
    }
}

// One thread per P1xP1 entry: mass and stiffness; writes the A^2 pressure-block values
//   (i,j): cm[i][j] M + delta_ij cl[i] L      (BlockCoefs, ctx.h)
__global__ void __launch_bounds__(256)
k_asm11(const int32_t* __restrict__ rp11, const int32_t* __restrict__ gptr,
        const uint32_t* __restrict__ glist, const double* __restrict__ geom,
        const int32_t* __restrict__ rp12, const int64_t* __restrict__ rowptr, int64_t n2, int64_t nv,
        int64_t nnz11, int A, const AsmCoefs K, const CellCoefs CC, double* __restrict__ vals,
        double* __restrict__ m11, double* __restrict__ l11) {
    const double* cm = K.cm;
    const double* cl = K.cl;
    int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= nnz11) return;
    int64_t v = row_of_entry(rp11, nv, (int32_t)e);
    double M = 0, L = 0;
    double Lw[MPET_MAX_NETWORKS];
#pragma unroll
    for (int f = 0; f < MPET_MAX_NETWORKS; ++f) Lw[f] = 0.0;
    int32_t g1 = gptr[e + 1];
    for (int32_t g = gptr[e]; g < g1; ++g) {
        uint32_t id = glist[g];
        uint32_t c = id / 16u;
        int p = (int)(id - c * 16u);
        int m = p >> 2, n = p & 3;
        double Ji[9], det;
        load_geom(geom, c, Ji, det);
        M += det * c_M11[p];
        double s = 0;
#pragma unroll
        for (int x = 0; x < 3; ++x) {
            double gm = c_GL[m * 3] * Ji[x] + c_GL[m * 3 + 1] * Ji[3 + x] + c_GL[m * 3 + 2] * Ji[6 + x];
            double gn = c_GL[n * 3] * Ji[x] + c_GL[n * 3 + 1] * Ji[3 + x] + c_GL[n * 3 + 2] * Ji[6 + x];
            s += gm * gn;
        }
        const double lc = det * s * (1.0 / 6.0);
        L += lc;
#pragma unroll
        for (int f = 0; f < MPET_MAX_NETWORKS; ++f)
            if (f < A && CC.kc[f]) Lw[f] += CC.kc[f][c] * lc;        // same fixed order: deterministic
    }
    m11[e] = M;
    l11[e] = L;
    for (int f = 0; f < A; ++f)
        if (CC.kc[f]) CC.lw[f][e] = Lw[f];
    if (vals == nullptr) return;
    int32_t d12 = rp12[v + 1] - rp12[v];
    int32_t d11 = rp11[v + 1] - rp11[v];
    int32_t jj = (int32_t)e - rp11[v];
    for (int i = 0; i < A; ++i) {
        int64_t base = rowptr[3 * n2 + i * nv + v] + 3 * (int64_t)d12 + jj;
        for (int j = 0; j < A; ++j) {
            double val = cm[i * A + j] * M;
            if (i == j) val += cl[i] * (CC.kc[i] ? Lw[i] : L);
            vals[base + (int64_t)j * d11] = val;
        }
    }
}

__global__ void k_prec11(const double* __restrict__ m11, const double* __restrict__ l11, int64_t nnz11,
                         int A, const AsmCoefs K, const CellCoefs CC, double* __restrict__ pp11) {
    const double* cm = K.cup;      // slots re-used: pm in cup, pk in cpu
    const double* ck = K.cpu;
    int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= nnz11) return;
    double M = m11[e], L = l11[e];
    for (int i = 0; i < A; ++i) pp11[i * nnz11 + e] = cm[i] * M + ck[i] * (CC.kc[i] ? CC.lw[i][e] : L);
}

}  // namespace

void init_reference_tables() {
    if (g_tables_ready) return;
    double R[900], Q[120], M22[100], M11[16], GL[12];
    for (int a = 0; a < 10; ++a)
        for (int b = 0; b < 10; ++b) {
            for (int m = 0; m < 3; ++m)
                for (int n = 0; n < 3; ++n)
                    R[(a * 10 + b) * 9 + m * 3 + n] = pint(pmul(dphi2(a, m), dphi2(b, n)));
            M22[a * 10 + b] = pint(pmul(phi2(a), phi2(b)));
        }
    for (int a = 0; a < 10; ++a)
        for (int m = 0; m < 4; ++m)
            for (int j = 0; j < 3; ++j) Q[(a * 4 + m) * 3 + j] = pint(pmul(dphi2(a, j), lam(m)));
    for (int m = 0; m < 4; ++m)
        for (int n = 0; n < 4; ++n) M11[m * 4 + n] = pint(pmul(lam(m), lam(n)));
    for (int m = 0; m < 4; ++m)
        for (int j = 0; j < 3; ++j) GL[m * 3 + j] = h_GL[m][j];
    CUDA_CHECK(cudaMemcpyToSymbol(c_R22, R, sizeof R));
    CUDA_CHECK(cudaMemcpyToSymbol(c_Q21, Q, sizeof Q));
    CUDA_CHECK(cudaMemcpyToSymbol(c_M22, M22, sizeof M22));
    CUDA_CHECK(cudaMemcpyToSymbol(c_M11, M11, sizeof M11));
    CUDA_CHECK(cudaMemcpyToSymbol(c_GL, GL, sizeof GL));
    g_tables_ready = true;
}

void compute_geometry(mpet_ctx* ctx, cudaStream_t st) {
    init_reference_tables();
    ctx->geom = dev_alloc<double>(ctx, ctx->Nc * 10);
    k_geometry<<<grid_for(ctx->Nc, 128), 128, 0, st>>>(ctx->coords, ctx->cells, ctx->Nc, ctx->geom);
    LAUNCH_CHECK(ctx);
}

// persistent grid of k_asm22: 4 CTAs of 8 warps per SM (or fewer when there are fewer rows)
static int asm22_grid(mpet_ctx* ctx, int64_t n2) {
    return (int)std::min<int64_t>(grid_for(n2 * 32, 256), (int64_t)ctx->sm_count * 4);
}

static CellCoefs cell_coefs(mpet_ctx* ctx) {
    CellCoefs CC;
    for (int f = 0; f < MPET_MAX_NETWORKS; ++f) { CC.kc[f] = ctx->kcell[f]; CC.lw[f] = ctx->l11w[f]; }
    return CC;
}

void assemble_lhs(mpet_ctx* ctx, cudaStream_t st) {
    MPET_REQUIRE(ctx->params_set, "mpet_set_params must be called before assembly");
    const int A = ctx->A;
    const int64_t n2 = ctx->N2, nv = ctx->Nv;
    const BlockCoefs& C = ctx->coef;
    AsmCoefs K;
    for (int i = 0; i < MPET_MAX_NETWORKS; ++i) { K.cup[i] = C.cup[i]; K.cpu[i] = C.cpu[i]; K.cl[i] = C.cl[i]; }
    for (int i = 0; i < MPET_MAX_NETWORKS * MPET_MAX_NETWORKS; ++i) K.cm[i] = C.cm[i];

    const int threads = 256;
    if (ctx->k22) {
        k_asm22<true, true, false><<<asm22_grid(ctx, n2), threads, 0, st>>>(
            ctx->g22.rowptr, ctx->g22.gptr, ctx->g22.glist, ctx->geom, ctx->rowptr, n2, C.uu_mu, C.uu_lam,
            ctx->vals, ctx->k22, nullptr);
    } else {
        k_asm22<true, false, false><<<asm22_grid(ctx, n2), threads, 0, st>>>(
            ctx->g22.rowptr, ctx->g22.gptr, ctx->g22.glist, ctx->geom, ctx->rowptr, n2, C.uu_mu, C.uu_lam,
            ctx->vals, nullptr, nullptr);
    }
    LAUNCH_CHECK(ctx);
    if (A > 0) {
        k_asm21<<<grid_for(ctx->g21.nnz, threads), threads, 0, st>>>(
            ctx->g21.rowptr, ctx->g21.col, ctx->g21.gptr, ctx->g21.glist, ctx->geom, ctx->g22.rowptr,
            ctx->g12.rowptr, ctx->t21to12, ctx->rowptr, n2, nv, ctx->g21.nnz, A, K, ctx->vals);
        LAUNCH_CHECK(ctx);
        k_asm11<<<grid_for(ctx->g11.nnz, threads), threads, 0, st>>>(
            ctx->g11.rowptr, ctx->g11.gptr, ctx->g11.glist, ctx->geom, ctx->g12.rowptr, ctx->rowptr, n2, nv,
            ctx->g11.nnz, A, K, cell_coefs(ctx), ctx->vals, ctx->m11, ctx->l11);
        LAUNCH_CHECK(ctx);
    }
    ctx->lhs_ready = true;
}

void ensure_m22(mpet_ctx* ctx, cudaStream_t st) {
    if (ctx->m22) return;
    ctx->m22 = dev_alloc<double>(ctx, ctx->g22.nnz);
    k_asm22<false, false, true><<<asm22_grid(ctx, ctx->N2), 256, 0, st>>>(
        ctx->g22.rowptr, ctx->g22.gptr, ctx->g22.glist, ctx->geom, ctx->rowptr, ctx->N2, 0, 0, nullptr,
        nullptr, ctx->m22);
    LAUNCH_CHECK(ctx);
}

void assemble_prec(mpet_ctx* ctx, cudaStream_t st) {
    MPET_REQUIRE(ctx->params_set, "mpet_set_params must be called before assembly");
    const int A = ctx->A;
    if (!ctx->k22) {
        ctx->k22 = dev_alloc<double>(ctx, ctx->g22.nnz);
        k_asm22<false, true, false><<<asm22_grid(ctx, ctx->N2), 256, 0, st>>>(
            ctx->g22.rowptr, ctx->g22.gptr, ctx->g22.glist, ctx->geom, ctx->rowptr, ctx->N2, 0, 0, nullptr,
            ctx->k22, nullptr);
        LAUNCH_CHECK(ctx);
    }
    if (A == 0) { ctx->prec_ready = true; return; }
    if (!ctx->lhs_ready || ctx->cell_coef_dirty) {   // m11 / l11 (/ weighted stiffness) not yet computed
        k_asm11<<<grid_for(ctx->g11.nnz, 256), 256, 0, st>>>(
            ctx->g11.rowptr, ctx->g11.gptr, ctx->g11.glist, ctx->geom, ctx->g12.rowptr, ctx->rowptr, ctx->N2,
            ctx->Nv, ctx->g11.nnz, A, AsmCoefs(), cell_coefs(ctx), nullptr, ctx->m11, ctx->l11);
        LAUNCH_CHECK(ctx);
    }
    if (!ctx->pp11) ctx->pp11 = dev_alloc<double>(ctx, (int64_t)A * ctx->g11.nnz);
    AsmCoefs K;                           // mpetsolver.py:270-271 / mpettotalpressuresolver.py:279-282
    for (int i = 0; i < MPET_MAX_NETWORKS; ++i) { K.cup[i] = ctx->coef.pm[i] + ctx->prec_shift_p[i]; K.cpu[i] = ctx->coef.pk[i]; }
    k_prec11<<<grid_for(ctx->g11.nnz, 256), 256, 0, st>>>(ctx->m11, ctx->l11, ctx->g11.nnz, A, K, cell_coefs(ctx), ctx->pp11);
    LAUNCH_CHECK(ctx);
    ctx->cell_coef_dirty = false;
    ctx->prec_ready = true;
}
