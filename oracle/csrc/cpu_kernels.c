/* CPU kernels of the oracle's all-host-core timing arm (TEST INFRASTRUCTURE, see oracle/__init__.py).
 *
 * What PETSc does on the reference path inside KSPSolve with an AIJ matrix (MatMult_SeqAIJ /
 * MatMult_MPIAIJ: one pass over the CSR rows, [EXT] petsc/src/mat/impls/aij/seq/aij.c) restated as an
 * OpenMP row loop so that bench.py's `--impl reference` / `cpu_baseline` legs can use every host core
 * the way `mpirun -n <cores>` would (src/mpet/utils/jobscript.sh:43).  Numerically identical to
 * scipy's csr_matvec (same summation order within a row).
 *
 * Built by oracle/build.py:  gcc -O3 -march=native -fopenmp -shared -fPIC
 */
#include <stdint.h>
#include <stddef.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* Y[n x w] = A X[ncols x w]  (row-major dense blocks, w right-hand sides) */
void oracle_csr_spmm(int64_t n, const int32_t* rowptr, const int32_t* cols, const double* vals,
                     const double* x, double* y, int w) {
    if (w == 1) {
#pragma omp parallel for schedule(static)
        for (int64_t r = 0; r < n; ++r) {
            double s = 0.0;
            for (int32_t e = rowptr[r]; e < rowptr[r + 1]; ++e) s += vals[e] * x[cols[e]];
            y[r] = s;
        }
        return;
    }
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < n; ++r) {
        double s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int32_t e = rowptr[r]; e < rowptr[r + 1]; ++e) {
            const double v = vals[e];
            const double* xr = x + (size_t)cols[e] * w;
            for (int k = 0; k < w; ++k) s[k] += v * xr[k];
        }
        for (int k = 0; k < w; ++k) y[(size_t)r * w + k] = s[k];
    }
}

/* z = a*x + b*y (+ c*w when w != NULL): the MINRES vector updates */
void oracle_axpbypcz(int64_t n, double a, const double* x, double b, const double* y, double c,
                     const double* w, double* z) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) z[i] = a * x[i] + b * y[i] + (w ? c * w[i] : 0.0);
}

double oracle_dot(int64_t n, const double* x, const double* y) {
    double s = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : s)
    for (int64_t i = 0; i < n; ++i) s += x[i] * y[i];
    return s;
}

/* torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU arm runs on rank 0 alone and takes the cores back */
void oracle_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ---------------------------------------------------------------------------------------------------
 * Cell loop of `assemble(a)` for the standard formulation on tetrahedra (mpetsolver.py:196-201,260,335),
 * the way DOLFIN runs it: per cell an element matrix from tabulated basis gradients at the quadrature
 * points (what FFC's generated tabulate_tensor does), then MatSetValues-style insertion (binary search of
 * the column in the CSR row, ADD_VALUES).  OpenMP over cells with atomic adds = what `mpirun -n <cores>` +
 * MatAssemblyEnd achieve.  Restates oracle/mpet.py:element_matrix_lhs; checked against it to round-off
 * (tests/test_oracle_krylov.py).  Used only by the CPU timing arms of bench.py.
 *
 * dN2 [nq][10][3], N1 [nq][4], dN1 [4][3] (constant), wq [nq]: reference tables from oracle/fem.py.
 * coef: mu, lambda, dt*theta, then alpha[A], c[A], K[A], S[A*A].  */
static inline int64_t find_col(const int32_t* cols, int64_t lo, int64_t hi, int32_t c) {
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (cols[mid] < c) lo = mid + 1; else hi = mid;
    }
    return lo;
}

void oracle_assemble_lhs_tet(int64_t nc, const int64_t* cells, const int64_t* cell_dofs, int A,
                             const double* coords, int nq, const double* dN2, const double* N1,
                             const double* dN1, const double* wq, const double* coef,
                             const int64_t* rowptr, const int32_t* cols, double* vals) {
    const int nloc = 30 + 4 * A;
    const double mu = coef[0], lmbda = coef[1], dth = coef[2];
    const double* alpha = coef + 3;
    const double* cc = alpha + A;
    const double* K = cc + A;
    const double* S = K + A;
#pragma omp parallel
    {
        double Ae[62 * 62];
#pragma omp for schedule(static)
        for (int64_t c = 0; c < nc; ++c) {
            const int64_t* cv = cells + 4 * c;
            double J[3][3], Ji[3][3];
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j) J[i][j] = coords[3 * cv[j + 1] + i] - coords[3 * cv[0] + i];
            const double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1];
            const double c01 = J[1][2] * J[2][0] - J[1][0] * J[2][2];
            const double c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
            const double det = J[0][0] * c00 + J[0][1] * c01 + J[0][2] * c02;
            const double id = 1.0 / det;
            Ji[0][0] = c00 * id; Ji[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) * id; Ji[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) * id;
            Ji[1][0] = c01 * id; Ji[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) * id; Ji[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) * id;
            Ji[2][0] = c02 * id; Ji[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) * id; Ji[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * id;
            const double adet = det < 0 ? -det : det;
            for (int i = 0; i < nloc * nloc; ++i) Ae[i] = 0.0;
            double g1[4][3];                               /* physical P1 gradients (constant) */
            for (int m = 0; m < 4; ++m)
                for (int x = 0; x < 3; ++x)
                    g1[m][x] = dN1[m * 3] * Ji[0][x] + dN1[m * 3 + 1] * Ji[1][x] + dN1[m * 3 + 2] * Ji[2][x];
            for (int q = 0; q < nq; ++q) {
                const double w = wq[q] * adet;
                double g2[10][3];
                for (int a = 0; a < 10; ++a)
                    for (int x = 0; x < 3; ++x)
                        g2[a][x] = dN2[(q * 10 + a) * 3] * Ji[0][x] + dN2[(q * 10 + a) * 3 + 1] * Ji[1][x] +
                                   dN2[(q * 10 + a) * 3 + 2] * Ji[2][x];
                for (int a = 0; a < 10; ++a)
                    for (int b = 0; b < 10; ++b) {
                        const double gg = g2[a][0] * g2[b][0] + g2[a][1] * g2[b][1] + g2[a][2] * g2[b][2];
                        for (int k = 0; k < 3; ++k)
                            for (int l = 0; l < 3; ++l) {
                                /* mu (delta_kl grad.grad + d_l phi_a d_k phi_b) + lambda d_k phi_a d_l phi_b */
                                double v = mu * g2[a][l] * g2[b][k] + lmbda * g2[a][k] * g2[b][l];
                                if (k == l) v += mu * gg;
                                Ae[(k * 10 + a) * nloc + l * 10 + b] += w * v;
                            }
                    }
                for (int i = 0; i < A; ++i) {
                    const int pi = 30 + 4 * i;
                    for (int a = 0; a < 10; ++a)
                        for (int m = 0; m < 4; ++m)
                            for (int k = 0; k < 3; ++k) {
                                const double v = -alpha[i] * w * g2[a][k] * N1[q * 4 + m];
                                Ae[(k * 10 + a) * nloc + pi + m] += v;
                                Ae[(pi + m) * nloc + k * 10 + a] += v;
                            }
                    double offsum = 0.0;
                    for (int j = 0; j < A; ++j) if (j != i) offsum += S[i * A + j];
                    for (int m = 0; m < 4; ++m)
                        for (int n = 0; n < 4; ++n) {
                            const double Mmn = w * N1[q * 4 + m] * N1[q * 4 + n];
                            const double Lmn = w * (g1[m][0] * g1[n][0] + g1[m][1] * g1[n][1] + g1[m][2] * g1[n][2]);
                            Ae[(pi + m) * nloc + pi + n] += -cc[i] * Mmn - dth * K[i] * Lmn - dth * offsum * Mmn;
                            for (int j = 0; j < A; ++j)
                                if (j != i) Ae[(pi + m) * nloc + 30 + 4 * j + n] += dth * S[i * A + j] * Mmn;
                        }
                }
            }
            const int64_t* cd = cell_dofs + (int64_t)nloc * c;
            for (int r = 0; r < nloc; ++r) {
                const int64_t row = cd[r], lo = rowptr[row], hi = rowptr[row + 1];
                for (int s = 0; s < nloc; ++s) {
                    const int64_t t = find_col(cols, lo, hi, (int32_t)cd[s]);
#pragma omp atomic
                    vals[t] += Ae[r * nloc + s];
                }
            }
        }
    }
}
