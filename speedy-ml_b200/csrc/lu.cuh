// lu.cuh -- dgesv on the device: blocked LU with partial pivoting + the two triangular solves.
//
// mldivide (src/mod_linalg.f90:109-151) is "call dgesv(n, nrhs, A, lda, ipiv, B, ldb, info)"; fit_chunk_hybrid / _ml
// reach it with the regularised Gram.  The engine's fit path factorises that symmetric positive-definite system by
// Cholesky (chol.cuh); this file is the general route -- sml_mldivide, and the per-region fallback when a Cholesky
// pivot is not positive -- with dgesv's semantics: row interchanges (first maximal |a| in the column, like idamax),
// unit-lower L and U stored over A, info = index of the first exactly-zero pivot (then B is not touched).
// Right-looking, panel width 32:
//   k_lu_panel  one CTA factorises the panel columns over all remaining rows (pivot search, swap inside the panel,
//               scale, rank-1 updates inside the panel) and records ipiv;
//   k_lu_laswp  applies the panel's interchanges to the columns left and right of it;
//   k_lu_trsm   U12 = L11^-1 A12 (unit lower, thread per column);
//   k_lu_gemm   A22 -= L21 U12, 64 x 64 tiles, K = 32 (FP64 FMA; this is the fallback, the DMMA path is chol.cuh).
// Solve: interchanges on B, forward substitution (unit L), back substitution (U), one CTA per right-hand side.
#pragma once
#include <cuda_runtime.h>

namespace sml {

constexpr int LU_NB = 32;

__global__ void __launch_bounds__(1024, 1)
k_lu_panel(double *__restrict__ A, int lda, int n, int j0, int nb, int *__restrict__ ipiv, int *__restrict__ info)
{
    __shared__ double s_val[32];
    __shared__ int s_idx[32];
    __shared__ int s_piv;
    __shared__ double s_pivval;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nt = blockDim.x;
    for (int jj = 0; jj < nb; ++jj) {
        const int j = j0 + jj;
        double *col = A + (size_t)lda * j;
        // pivot: first index of the largest |a| in rows j..n-1
        double best = -1.0;
        int bi = n;
        for (int i = j + tid; i < n; i += nt) {
            const double v = fabs(col[i]);
            if (v > best || (v == best && i < bi)) { best = v; bi = i; }
        }
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_down_sync(0xffffffffu, best, o);
            const int oi = __shfl_down_sync(0xffffffffu, bi, o);
            if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
        }
        if (lane == 0) { s_val[warp] = best; s_idx[warp] = bi; }
        __syncthreads();
        if (warp == 0) {
            best = (lane < nt / 32) ? s_val[lane] : -1.0;
            bi = (lane < nt / 32) ? s_idx[lane] : n;
            for (int o = 16; o > 0; o >>= 1) {
                const double ov = __shfl_down_sync(0xffffffffu, best, o);
                const int oi = __shfl_down_sync(0xffffffffu, bi, o);
                if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
            }
            if (lane == 0) {
                if (bi >= n) bi = j;   // a column of NaNs: keep the diagonal (the NaNs propagate, as in LAPACK)
                s_piv = bi;
                s_pivval = col[bi];
                ipiv[j] = bi;
                if (col[bi] == 0.0 && *info == 0) *info = j + 1;   // exactly singular: U(j,j) = 0 (dgetf2)
            }
        }
        __syncthreads();
        const int p = s_piv;
        const double piv = s_pivval;
        // interchange inside the panel
        if (p != j && tid < nb) {
            double *c = A + (size_t)lda * (j0 + tid);
            const double t = c[j];
            c[j] = c[p];
            c[p] = t;
        }
        __syncthreads();
        if (piv != 0.0) {
            for (int i = j + 1 + tid; i < n; i += nt) col[i] = col[i] / piv;
        }
        __syncthreads();
        // rank-1 update of the remaining panel columns
        const int rem = nb - jj - 1;
        if (rem > 0) {
            const int rows = n - j - 1;
            for (long long e = tid; e < (long long)rows * rem; e += nt) {
                const int i = j + 1 + (int)(e % rows), c = j + 1 + (int)(e / rows);
                A[(size_t)lda * c + i] -= col[i] * A[(size_t)lda * c + j];
            }
        }
        __syncthreads();
    }
}

// interchanges of panel [j0, j0+nb) applied to columns [c_lo, c_hi) of M (leading dimension ldm), thread per column
__global__ void k_lu_laswp(double *__restrict__ M, int ldm, int c_lo, int c_hi, const int *__restrict__ ipiv, int j0, int nb)
{
    const int c = c_lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= c_hi) return;
    double *col = M + (size_t)ldm * c;
    for (int jj = 0; jj < nb; ++jj) {
        const int j = j0 + jj, p = ipiv[j];
        if (p != j) {
            const double t = col[j];
            col[j] = col[p];
            col[p] = t;
        }
    }
}

// U12 = L11^-1 A12 for the columns right of the panel: unit-lower forward substitution, thread per column
__global__ void k_lu_trsm(double *__restrict__ A, int lda, int n, int j0, int nb)
{
    __shared__ double L[LU_NB][LU_NB + 1];
    for (int e = threadIdx.x; e < nb * nb; e += blockDim.x) L[e % nb][e / nb] = A[(size_t)lda * (j0 + e / nb) + j0 + e % nb];
    __syncthreads();
    const int c = j0 + nb + blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    double *col = A + (size_t)lda * c + j0;
    double x[LU_NB];
#pragma unroll
    for (int i = 0; i < LU_NB; ++i) {
        double v = (i < nb) ? col[i] : 0.0;
#pragma unroll
        for (int m = 0; m < i; ++m) v = fma(-L[i][m], x[m], v);
        x[i] = v;
    }
#pragma unroll
    for (int i = 0; i < LU_NB; ++i)
        if (i < nb) col[i] = x[i];
}

// A22 -= L21 * U12 : CTA tile 64 x 64, 256 threads x (4 x 4), K = nb <= 32
__global__ void __launch_bounds__(256)
k_lu_gemm(double *__restrict__ A, int lda, int n, int j0, int nb)
{
    __shared__ double sL[LU_NB][64 + 1];   // L21 tile: [k][row]
    __shared__ double sU[LU_NB][64 + 1];   // U12 tile: [k][col]
    const int r0 = j0 + nb + blockIdx.x * 64, c0 = j0 + nb + blockIdx.y * 64;
    const int tid = threadIdx.x;
    for (int e = tid; e < nb * 64; e += 256) {
        const int k = e / 64, i = e % 64;
        sL[k][i] = (r0 + i < n) ? A[(size_t)lda * (j0 + k) + r0 + i] : 0.0;
    }
    for (int e = tid; e < nb * 64; e += 256) {
        const int c = e / nb, k = e % nb;
        sU[k][c] = (c0 + c < n) ? A[(size_t)lda * (c0 + c) + j0 + k] : 0.0;
    }
    __syncthreads();
    const int ti = (tid % 16) * 4, tc = (tid / 16) * 4;
    double acc[4][4] = {};
    for (int k = 0; k < nb; ++k) {
        double a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = sL[k][ti + i];
#pragma unroll
        for (int c = 0; c < 4; ++c) b[c] = sU[k][tc + c];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[i][c] = fma(a[i], b[c], acc[i][c]);
    }
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (r0 + ti + i < n && c0 + tc + c < n) A[(size_t)lda * (c0 + tc + c) + r0 + ti + i] -= acc[i][c];
}

// L y = b (unit lower) then U x = y for one right-hand side per CTA; x lives in shared memory
__global__ void __launch_bounds__(1024, 1)
k_lu_solve(const double *__restrict__ A, int lda, int n, double *__restrict__ B, int ldb)
{
    extern __shared__ double xs[];
    double *b = B + (size_t)ldb * blockIdx.x;
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int i = tid; i < n; i += nt) xs[i] = b[i];
    __syncthreads();
    for (int j = 0; j < n; ++j) {
        const double xj = xs[j];
        const double *col = A + (size_t)lda * j;
        for (int i = j + 1 + tid; i < n; i += nt) xs[i] = fma(-col[i], xj, xs[i]);
        __syncthreads();
    }
    for (int j = n - 1; j >= 0; --j) {
        const double *col = A + (size_t)lda * j;
        if (tid == 0) xs[j] = xs[j] / col[j];
        __syncthreads();
        const double xj = xs[j];
        for (int i = tid; i < j; i += nt) xs[i] = fma(-col[i], xj, xs[i]);
        __syncthreads();
    }
    for (int i = tid; i < n; i += nt) b[i] = xs[i];
}

}  // namespace sml
