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

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
