// lu.cuh -- dgesv on the device: blocked LU with partial pivoting + the two triangular solves.
//
// mldivide (src/mod_linalg.f90:109-151) is "call dgesv(n, nrhs, A, lda, ipiv, B, ldb, info)"; fit_chunk_hybrid / _ml
// reach it with the regularised Gram.  The engine's fit path factorises that symmetric positive-definite system by
// Cholesky (chol.cuh); this file is the general route -- sml_mldivide, and the per-region fallback when a Cholesky
// pivot is not positive -- with dgesv's semantics: row interchanges (first maximal |a| in the column, like idamax),
// unit-lower L and U stored over A, info = index of the first exactly-zero pivot (then B is not touched).
// Right-looking, panel width 32:
//   k_lu_panel  a cooperative launch factorises the panel: its rows are split over the CTAs (each block in shared
//               memory), one grid barrier per column carries the pivot candidates and the two interchanged rows;
//   k_lu_laswp  applies the panel's interchanges to the columns left and right of it;
//   k_lu_trsm   U12 = L11^-1 A12 (unit lower, thread per column);
//   k_lu_gemm   A22 -= L21 U12 on the FP64 tensor cores (DMMA), 128 x 64 tiles, K = 32.
// Solve: interchanges on B, forward substitution (unit L), back substitution (U), one CTA per right-hand side,
// blocked by the panel width.
#pragma once
#include "train.cuh"   // dmma884
#include <cuda_runtime.h>

namespace sml {

constexpr int LU_NB = 32;

// ---------------------------------------------------------------------------------------------
// k_lu_panel: the panel's rows are split over the CTAs of one cooperative launch (all co-resident), each keeping its
// row block of the panel in shared memory.  Per column ONE grid barrier: before it every CTA publishes its local pivot
// candidate (|a|, row index) together with that row's nb panel values, and the owner of row j publishes row j; after
// it every CTA derives the same pivot (largest |a|, first index on ties -- idamax), has both rows of the interchange,
// and scales / rank-1-updates its own rows.  The exchange buffers are double-buffered by column parity (a fast CTA
// may publish column j+1 while a slow one still reads column j).
//   xch layout per parity: cand_val[G] | cand_row[G][LU_NB] | jrow[LU_NB] (doubles), cand_idx[G] (ints) behind them.
// Measured at n = 5892: 22.8 ms for the 184 panels (3.8 us per column).  A variant with two block barriers per column
// (only warp 0 publishes and waits at the grid barrier; interchange, scaling and the rank-1 update folded into one pass
// that also yields the next column's candidates) measured 38.4 ms and was dropped.
// ---------------------------------------------------------------------------------------------
constexpr int LU_PANEL_THREADS = 256;

__device__ __forceinline__ unsigned ld_acquire_gpu_u32(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// all CTAs of the (cooperative) launch; *ctr counts arrivals monotonically, target = G * (barriers passed + 1)
__device__ __forceinline__ void lu_grid_barrier(unsigned *ctr, unsigned target)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(ctr, 1u);
        while (ld_acquire_gpu_u32(ctr) < target) {
        }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(LU_PANEL_THREADS, 1)
k_lu_panel(double *__restrict__ A, int lda, int n, int j0, int nb, int rows_per, int *__restrict__ ipiv, int *__restrict__ info,
           double *__restrict__ xch, unsigned *__restrict__ ctr, unsigned ctr_base)
{
    extern __shared__ double pan[];           // [nb][rows_per + 1]: this CTA's rows of the panel, column-major
    __shared__ double s_val[LU_PANEL_THREADS / 32];
    __shared__ int s_idx[LU_PANEL_THREADS / 32];
    __shared__ double s_prow[LU_NB], s_jrow[LU_NB];
    __shared__ int s_piv;
    const int G = gridDim.x, cta = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ldp = rows_per + 1;
    const int g0 = j0 + cta * rows_per;                      // first global row of this CTA
    const int nr = max(0, min(rows_per, n - g0));            // rows held
    for (int e = tid; e < nr * nb; e += LU_PANEL_THREADS) {
        const int c = e / nr, i = e % nr;
        pan[c * ldp + i] = A[(size_t)lda * (j0 + c) + g0 + i];
    }
    __syncthreads();
    const size_t xstride = (size_t)G + (size_t)G * LU_NB + LU_NB + (G + 1) / 2;   // doubles per parity
    for (int jj = 0; jj < nb; ++jj) {
        const int j = j0 + jj;
        double *xb = xch + (size_t)(jj & 1) * xstride;
        double *cand_val = xb, *cand_row = xb + G, *jrow = cand_row + (size_t)G * LU_NB;
        int *cand_idx = reinterpret_cast<int *>(jrow + LU_NB);
        // ---- local candidate over rows >= j
        double best = -1.0;
        int bi = n;
        const double *col = pan + jj * ldp;
        for (int i = tid; i < nr; i += LU_PANEL_THREADS) {
            const int gi = g0 + i;
            if (gi < j) continue;
            const double v = fabs(col[i]);
            if (v > best || (v == best && gi < bi)) { best = v; bi = gi; }
        }
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_down_sync(0xffffffffu, best, o);
            const int oi = __shfl_down_sync(0xffffffffu, bi, o);
            if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
        }
        if (lane == 0) { s_val[warp] = best; s_idx[warp] = bi; }
        __syncthreads();
        if (warp == 0) {
            best = (lane < LU_PANEL_THREADS / 32) ? s_val[lane] : -1.0;
            bi = (lane < LU_PANEL_THREADS / 32) ? s_idx[lane] : n;
            for (int o = 16; o > 0; o >>= 1) {
                const double ov = __shfl_down_sync(0xffffffffu, best, o);
                const int oi = __shfl_down_sync(0xffffffffu, bi, o);
                if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
            }
            if (lane == 0) {
                cand_val[cta] = best;
                cand_idx[cta] = bi;
                s_piv = bi;
            }
        }
        __syncthreads();
        {
            const int lb = s_piv;   // local best row (n if this CTA has no row >= j)
            if (tid < nb && lb < n) cand_row[(size_t)cta * LU_NB + tid] = pan[tid * ldp + (lb - g0)];
            if (tid < nb && j >= g0 && j < g0 + nr) jrow[tid] = pan[tid * ldp + (j - g0)];
        }
        lu_grid_barrier(ctr, ctr_base + (unsigned)G * (unsigned)(jj + 1));
        // ---- global pivot, identical in every CTA
        if (warp == 0) {
            best = -1.0;
            bi = n;
            int bc = 0;
            for (int c = lane; c < G; c += 32) {
                const double v = __ldcg(cand_val + c);   // exchange buffers: L2 reads, never a stale L1 line
                const int ci = __ldcg(cand_idx + c);
                if (v > best || (v == best && ci < bi)) { best = v; bi = ci; bc = c; }
            }
            for (int o = 16; o > 0; o >>= 1) {
                const double ov = __shfl_down_sync(0xffffffffu, best, o);
                const int oi = __shfl_down_sync(0xffffffffu, bi, o);
                const int oc = __shfl_down_sync(0xffffffffu, bc, o);
                if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; bc = oc; }
            }
            bi = __shfl_sync(0xffffffffu, bi, 0);
            bc = __shfl_sync(0xffffffffu, bc, 0);
            if (bi >= n) {   // a column of NaNs: keep the diagonal (the NaNs propagate, as in LAPACK)
                bi = j;
                bc = (j - j0) / rows_per;
            }
            if (lane < nb) {
                s_prow[lane] = (bi == j) ? __ldcg(jrow + lane) : __ldcg(cand_row + (size_t)bc * LU_NB + lane);
                s_jrow[lane] = __ldcg(jrow + lane);
            }
            if (lane == 0) s_piv = bi;
        }
        __syncthreads();
        const int p = s_piv;
        const double piv = s_prow[jj];
        if (cta == 0 && tid == 0) {
            ipiv[j] = p;
            if (piv == 0.0 && *info == 0) *info = j + 1;   // exactly singular: U(j,j) = 0 (dgetf2)
        }
        // ---- interchange inside the panel (rows j and p), then scale and rank-1 update of the local rows
        if (p != j && tid < nb) {
            if (j >= g0 && j < g0 + nr) pan[tid * ldp + (j - g0)] = s_prow[tid];
            if (p >= g0 && p < g0 + nr) pan[tid * ldp + (p - g0)] = s_jrow[tid];
        }
        __syncthreads();
        const int i_lo = max(0, j + 1 - g0);   // local rows strictly below the diagonal
        if (piv != 0.0) {
            double *cj = pan + jj * ldp;
            for (int i = i_lo + tid; i < nr; i += LU_PANEL_THREADS) cj[i] = cj[i] / piv;
        }
        __syncthreads();
        const int rem = nb - jj - 1, rows = nr - i_lo;
        if (rem > 0 && rows > 0) {
            const double *cj = pan + jj * ldp;
            for (int e = tid; e < rows * rem; e += LU_PANEL_THREADS) {
                const int i = i_lo + e % rows, c = jj + 1 + e / rows;
                pan[c * ldp + i] -= cj[i] * s_prow[c];
            }
        }
        __syncthreads();
    }
    for (int e = tid; e < nr * nb; e += LU_PANEL_THREADS) {
        const int c = e / nr, i = e % nr;
        A[(size_t)lda * (j0 + c) + g0 + i] = pan[c * ldp + i];
    }
}

// interchanges of panel [j0, j0+nb) applied to columns [c_lo, c_hi) of M (leading dimension ldm), thread per column.
// The nb sequential row swaps touch at most 2 nb rows and compose to one permutation that is the same for every
// column: warp 0 of each CTA derives it once (row r receives the element that was in row src[r]); every thread
// then issues all its loads, then all its stores -- two dependent memory phases instead of nb.
__global__ void __launch_bounds__(128)
k_lu_laswp(double *__restrict__ M, int ldm, int c_lo, int c_hi, const int *__restrict__ ipiv, int j0, int nb)
{
    // entries 0..nb-1 are the panel's own rows j0..j0+nb-1 (direct index); pivot rows below the panel get entries
    // nb.. as they appear -- warp 0 looks them up with one ballot per interchange
    __shared__ int s_row[2 * LU_NB], s_src[2 * LU_NB];
    __shared__ int s_cnt;
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        const int myp = (lane < nb) ? ipiv[j0 + lane] : 0;
        if (lane < nb) { s_row[lane] = j0 + lane; s_src[lane] = j0 + lane; }
        int orow = -1;   // the row held by entry nb + lane
        int ocnt = 0;
        __syncwarp();
        for (int jj = 0; jj < nb; ++jj) {
            const int p = __shfl_sync(0xffffffffu, myp, jj);
            if (p == j0 + jj) continue;
            int ap;
            if (p < j0 + nb) {
                ap = p - j0;
            } else {
                const unsigned m = __ballot_sync(0xffffffffu, orow == p);
                if (m) {
                    ap = nb + __ffs(m) - 1;
                } else {
                    ap = nb + ocnt;
                    if (lane == ocnt) orow = p;
                    if (lane == 0) { s_row[ap] = p; s_src[ap] = p; }
                    ++ocnt;
                }
            }
            __syncwarp();
            if (lane == 0) {
                const int t = s_src[jj];
                s_src[jj] = s_src[ap];
                s_src[ap] = t;
            }
            __syncwarp();
        }
        if (lane == 0) s_cnt = nb + ocnt;
    }
    __syncthreads();
    const int c = c_lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= c_hi) return;
    double *col = M + (size_t)ldm * c;
    const int cnt = s_cnt;
    double v[2 * LU_NB];
#pragma unroll
    for (int e = 0; e < 2 * LU_NB; ++e)
        if (e < cnt) v[e] = col[s_src[e]];
#pragma unroll
    for (int e = 0; e < 2 * LU_NB; ++e)
        if (e < cnt) col[s_row[e]] = v[e];
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

// A22 -= L21 * U12 on the FP64 tensor cores (mma.sync m8n8k4 -> DMMA): CTA tile 128 x 64, 4 warps (2 x 2) of 64 x 32 =
// 8 x 4 DMMA tiles, K = nb <= 32 zero-padded to 32.  L21 is staged k-major (rows contiguous, as it lies in memory), U12
// transposed into the same form; the accumulators start from A22 (loaded first, in flight while the operands are
// staged) and the L operand is negated, so the tile leaves as A22 - L21 U12 with one read and one write of A22.  128 threads x 164 registers: 3 CTAs per SM cover each other's load / store phases.
constexpr int LG_BM = 128, LG_BN = 64, LG_LDL = 132, LG_LDU = 68;
constexpr size_t LG_SMEM = sizeof(double) * (size_t)LU_NB * (LG_LDL + LG_LDU);

__global__ void __launch_bounds__(128, 3)
k_lu_gemm(double *__restrict__ A, int lda, int n, int j0, int nb)
{
    extern __shared__ __align__(16) double lg_smem[];
    double *sL = lg_smem;                    // [LU_NB][LG_LDL]: sL[k][i] = -L21(r0 + i, k)
    double *sU = lg_smem + LU_NB * LG_LDL;   // [LU_NB][LG_LDU]: sU[k][c] =  U12(k, c0 + c)
    const int r0 = j0 + nb + blockIdx.x * LG_BM, c0 = j0 + nb + blockIdx.y * LG_BN;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int e = tid; e < LU_NB * LG_BM; e += 128) {
        const int k = e / LG_BM, i = e % LG_BM;
        sL[k * LG_LDL + i] = (k < nb && r0 + i < n) ? -A[(size_t)lda * (j0 + k) + r0 + i] : 0.0;
    }
    for (int e = tid; e < LU_NB * LG_BN; e += 128) {
        const int c = e / LU_NB, k = e % LU_NB;
        sU[k * LG_LDU + c] = (k < nb && c0 + c < n) ? A[(size_t)lda * (c0 + c) + j0 + k] : 0.0;
    }
    const int wm = warp >> 1, wn = warp & 1;
    const int g = lane >> 2, q = lane & 3;
    // accumulators <- A22 tile, issued before the barrier so that these loads fly while the operand tiles are staged
    // (thread holds C[g][2q], C[g][2q+1] of every 8 x 8 tile)
    double acc[8][4][2];
    const int ib = r0 + wm * 64 + g;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        const int c = c0 + wn * 32 + b * 8 + 2 * q;
        const double *col0 = A + (size_t)lda * c + ib, *col1 = col0 + lda;
        const bool ok0 = c < n, ok1 = c + 1 < n;
#pragma unroll
        for (int a = 0; a < 8; ++a) {
            const bool okr = ib + a * 8 < n;
            acc[a][b][0] = (okr && ok0) ? col0[a * 8] : 0.0;
            acc[a][b][1] = (okr && ok1) ? col1[a * 8] : 0.0;
        }
    }
    __syncthreads();
#pragma unroll
    for (int k4 = 0; k4 < LU_NB; k4 += 4) {
        double af[8], bf[4];
        const double *pa = sL + (k4 + q) * LG_LDL + wm * 64 + g;
        const double *pb = sU + (k4 + q) * LG_LDU + wn * 32 + g;
#pragma unroll
        for (int a = 0; a < 8; ++a) af[a] = pa[a * 8];
#pragma unroll
        for (int b = 0; b < 4; ++b) bf[b] = pb[b * 8];
#pragma unroll
        for (int a = 0; a < 8; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) dmma884(acc[a][b][0], acc[a][b][1], af[a], bf[b]);
    }
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        const int c = c0 + wn * 32 + b * 8 + 2 * q;
        double *col0 = A + (size_t)lda * c + ib, *col1 = col0 + lda;
        const bool ok0 = c < n, ok1 = c + 1 < n;
#pragma unroll
        for (int a = 0; a < 8; ++a) {
            if (ib + a * 8 < n) {
                if (ok0) col0[a * 8] = acc[a][b][0];
                if (ok1) col1[a * 8] = acc[a][b][1];
            }
        }
    }
}

// A22 -= L21 * U12 with plain FP64 FMAs: CTA tile 64 x 64, 256 threads x (4 x 4), K = nb <= 32 (SML_LU_GEMM=fma; A/B)
__global__ void __launch_bounds__(256)
k_lu_gemm_fma(double *__restrict__ A, int lda, int n, int j0, int nb)
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


// L y = b (unit lower) then U x = y for one right-hand side per CTA; x lives in shared memory.  Blocked by LU_NB
// columns: the diagonal block is solved by one warp from a shared-memory copy, then all threads apply the block's LU_NB
// columns to the rest of x -- two block barriers per LU_NB columns instead of one per column.  Each x(i) receives its
// updates in the order of the column-by-column substitution (j ascending for L, descending for U).
__global__ void __launch_bounds__(1024, 1)
k_lu_solve(const double *__restrict__ A, int lda, int n, double *__restrict__ B, int ldb)
{
    extern __shared__ double xs[];
    __shared__ double blk[LU_NB][LU_NB + 1];   // blk[i][k] = A(jb + i, jb + k)
    double *b = B + (size_t)ldb * blockIdx.x;
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < n; i += nt) xs[i] = b[i];
    __syncthreads();
    // ---- forward: unit lower triangle
    for (int jb = 0; jb < n; jb += LU_NB) {
        const int nb = min(LU_NB, n - jb);
        for (int e = tid; e < nb * nb; e += nt) blk[e % nb][e / nb] = A[(size_t)lda * (jb + e / nb) + jb + e % nb];
        __syncthreads();
        if (warp == 0) {
            double x = (lane < nb) ? xs[jb + lane] : 0.0;
            for (int k = 0; k < nb; ++k) {
                const double xk = __shfl_sync(0xffffffffu, x, k);
                if (lane > k && lane < nb) x = fma(-blk[lane][k], xk, x);
            }
            if (lane < nb) xs[jb + lane] = x;
        }
        __syncthreads();
        for (int i = jb + nb + tid; i < n; i += nt) {
            double acc = xs[i];
            const double *row = A + (size_t)lda * jb + i;
            for (int k = 0; k < nb; ++k) acc = fma(-row[(size_t)lda * k], xs[jb + k], acc);
            xs[i] = acc;
        }
        __syncthreads();
    }
    // ---- backward: upper triangle with its diagonal
    for (int jb = ((n - 1) / LU_NB) * LU_NB; jb >= 0; jb -= LU_NB) {
        const int nb = min(LU_NB, n - jb);
        for (int e = tid; e < nb * nb; e += nt) blk[e % nb][e / nb] = A[(size_t)lda * (jb + e / nb) + jb + e % nb];
        __syncthreads();
        if (warp == 0) {
            double x = (lane < nb) ? xs[jb + lane] : 0.0;
            for (int k = nb - 1; k >= 0; --k) {
                if (lane == k) x = x / blk[k][k];
                const double xk = __shfl_sync(0xffffffffu, x, k);
                if (lane < k) x = fma(-blk[lane][k], xk, x);
            }
            if (lane < nb) xs[jb + lane] = x;
        }
        __syncthreads();
        for (int i = tid; i < jb; i += nt) {
            double acc = xs[i];
            const double *row = A + (size_t)lda * jb + i;
            for (int k = nb - 1; k >= 0; --k) acc = fma(-row[(size_t)lda * k], xs[jb + k], acc);
            xs[i] = acc;
        }
        __syncthreads();
    }
    for (int i = tid; i < n; i += nt) b[i] = xs[i];
}

}  // namespace sml
