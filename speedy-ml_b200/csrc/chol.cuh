// chol.cuh -- batched blocked Cholesky solve of the ridge systems (fit_chunk_hybrid / fit_chunk_ml,
// src/mod_reservoir.f90:1177-1334 -> mldivide -> dgesv, src/mod_linalg.f90:109-151).
//
// The reference hands dgesv the regularised Gram matrix A = R R^T + ridge (symmetric positive definite whenever
// the ridge is > 0) and B = (Y R^T)^T, and keeps W_out = X^T.  Here every region of the training wave is solved at
// once, in place in its augmented accumulator Gaug (train.cuh): the lower triangle holds A and, in rows N..N+P,
// Y R^T.  Running a left-looking blocked Cholesky over the first N columns of that augmented matrix yields
//     L (N x N)            with A = L L^T
//     Z = (Y R^T) L^-T     in rows N..N+P  (the panel solve covers them like any other row below the diagonal)
// and a blocked back substitution W L = Z, again only on rows N..N+P, leaves W_out = (Y R^T) A^-1 exactly where
// Y R^T was.  All O(N^3) work runs through one FP64 tensor-core (DMMA) tile-GEMM with TMA-fed operands:
//     UPDATE      A[i-tile, blk k] -= L[i-tile, 0:j0] L[blk k, 0:j0]^T              (K = j0)
//     TRSM        L[rows below, blk k]  = A[rows below, blk k] Linv_k^T             (K = nb)
//     BACK_TRI    W[:, blk k]  = W[:, blk k] Linv_k                                 (K = nb)
//     BACK_UPDATE W[:, blk j] -= W[:, blk k] L[blk k, blk j]   for every j < k      (K = nb; right-looking, so
//                 each step is k*ceil(P/128) independent tiles instead of one long-K tile)
// with the 128 x 128 diagonal blocks factorised and inverted in shared memory by k_chol_diag.
// BACK_UPDATE needs L^T with the K index along a column; after the factorisation L is mirrored into the upper
// triangle once (k_train_ridge_mirror without ridge terms).  Until then the upper triangle still holds the
// untouched copy of A, so a region whose Cholesky meets a non-positive pivot (ridge 0 / rank-deficient Gram) is
// restored from it and solved by LU with partial pivoting exactly as dgesv would -- same info semantics.
#pragma once
#include "train.cuh"

namespace sml {

constexpr int CH_NB = 128;                       // panel width = tile size
constexpr int CH_LINV = 2 * CH_NB * CH_NB;       // doubles per panel: Linv (row c, col m) then LinvT
constexpr int CH_SB = 32;                        // sub-block of the diagonal-block kernel
constexpr int CH_LD = CH_NB + 1;
constexpr int CH_TLD = 3 * CH_SB + 1;
enum CholOp { CH_UPDATE = 0, CH_TRSM = 1, CH_BACK_UPDATE = 2, CH_BACK_TRI = 3 };

// C(128x128) = sum_kk Aop[i, kk] * Bop[j, kk], kk in [0, kvalid), kvalid % 4 == 0.
// Aop element (i, kk) at A[kk*lda + i], Bop element (j, kk) at B[kk*ldb + j]; rows beyond rowsA / rowsB are not
// loaded (their accumulators are never stored).  Same pipeline as k_syrk_dmma: one TMA producer warp, 8 DMMA warps.
__global__ void __launch_bounds__(SY_THREADS, 1)
k_chol_gemm(const TrainRegionDev *__restrict__ T, int op, int k)
{
    extern __shared__ __align__(128) unsigned char sy_smem[];
    double *stage0 = reinterpret_cast<double *>(sy_smem);
    uint64_t *full = reinterpret_cast<uint64_t *>(sy_smem + (size_t)SY_STAGES * SY_STAGE_DOUBLES * 8);
    uint64_t *empty = full + SY_STAGES;

    const TrainRegionDev &t = T[blockIdx.y];
    if (*t.chol_info != 0) return;  // this region left the Cholesky path (LU fallback)
    const int ld = t.ld, N = t.R.n + t.R.S, P = t.R.P;
    const int j0 = k * CH_NB;
    if (j0 >= N) return;
    const int nb = min(CH_NB, N - j0);
    double *G = t.gram;
    const double *linv = t.linv + (size_t)k * CH_LINV;

    // ---- operand / output geometry of this tile
    const double *A, *B;
    int lda, ldb, rowsA, rowsB, kvalid;
    bool same = false, lower_only = false, transposed = false;
    int mode;             // 0: C -= acc (RED), 1: C = acc (store)
    double *C;            // output element (i, j) at C[j*ld + i]  (transposed: C[i*ld + j])
    int ni, nj;           // valid output extent
    if (op == CH_UPDATE) {
        const int ti = k + blockIdx.x, i0 = ti * CH_NB;
        if (i0 >= ld) return;
        A = G + i0; lda = ld; rowsA = min(CH_NB, ld - i0);
        B = G + j0; ldb = ld; rowsB = min(CH_NB, ld - j0);
        kvalid = j0;
        same = blockIdx.x == 0;
        lower_only = same;
        mode = 0;
        C = G + (size_t)j0 * ld + i0;
        ni = rowsA; nj = nb;
    } else if (op == CH_TRSM) {
        const int i0 = j0 + nb + blockIdx.x * CH_NB;
        if (i0 >= ld) return;
        A = G + (size_t)j0 * ld + i0; lda = ld; rowsA = min(CH_NB, ld - i0);
        B = linv; ldb = CH_NB; rowsB = CH_NB;
        kvalid = nb;
        mode = 1;
        C = G + (size_t)j0 * ld + i0;
        ni = rowsA; nj = nb;
    } else {
        const int ptiles = (P + CH_NB - 1) / CH_NB;
        const int p0 = (blockIdx.x % ptiles) * CH_NB;
        const int np = min(CH_NB, P - p0);
        transposed = true;
        if (op == CH_BACK_UPDATE) {
            // block column jt < k of W receives the contribution of the just-solved block k:
            //   W[p, c] -= sum_{m in blk k} W[p, m] L[m, c]; L^T lives in the upper triangle: U[c, m] at G[m*ld + c]
            const int jt = blockIdx.x / ptiles;
            if (jt >= k) return;
            const int c0 = jt * CH_NB;
            A = G + (size_t)j0 * ld + c0; lda = ld; rowsA = CH_NB;        // rows c of block jt, K index m - j0
            B = G + (size_t)j0 * ld + N + p0; ldb = ld; rowsB = np;       // rows p, same K index
            kvalid = nb;
            mode = 0;
            C = G + (size_t)c0 * ld + N + p0;   // W[p0 + j, c0 + i] at C[i*ld + j]
            ni = CH_NB; nj = np;
        } else {
            if (blockIdx.x >= (unsigned)ptiles) return;
            C = G + (size_t)j0 * ld + N + p0;   // W[p0 + j, j0 + i] at C[i*ld + j]
            ni = nb; nj = np;
            A = linv + CH_NB * CH_NB; lda = CH_NB; rowsA = CH_NB;  // LinvT: element (c, c') = Linv[c', c]
            B = G + (size_t)j0 * ld + N + p0; ldb = ld; rowsB = np;
            kvalid = nb;
            mode = 1;
        }
    }
    // TMA needs 16-byte multiples: round the loaded row counts up to even (the pad rows exist: ld % 16 == 0 and
    // the scratch blocks are full 128 x 128)
    rowsA = (rowsA + 1) & ~1;
    rowsB = (rowsB + 1) & ~1;
    const int nchunks = (kvalid + SY_BK - 1) / SY_BK;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int s = 0; s < SY_STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], SY_CONS_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async;" ::: "memory");
    }
    __syncthreads();

    if (warp == SY_CONS_WARPS) {
        if (lane == 0) {
            for (int kc = 0; kc < nchunks; ++kc) {
                const int s = kc % SY_STAGES;
                const int nk = min(SY_BK, kvalid - kc * SY_BK);
                mbar_wait(&empty[s], ((kc / SY_STAGES) & 1) ^ 1);
                mbar_expect_tx(&full[s], (uint32_t)nk * (rowsA + (same ? 0 : rowsB)) * 8u);
                double *sA = stage0 + (size_t)s * SY_STAGE_DOUBLES;
                double *sB = sA + SY_BK * SY_LDS;
                for (int kk = 0; kk < nk; ++kk) {
                    const size_t col = (size_t)(kc * SY_BK + kk);
                    tma_load_1d(sA + kk * SY_LDS, A + col * lda, rowsA * 8u, &full[s]);
                    if (!same) tma_load_1d(sB + kk * SY_LDS, B + col * ldb, rowsB * 8u, &full[s]);
                }
            }
        }
        return;
    }

    const int wm = warp >> 2, wn = warp & 3;
    const int g = lane >> 2, q = lane & 3;
    const int ma = max(0, min(8, (rowsA - wm * 64 + 7) >> 3));
    const int nbt = max(0, min(4, (rowsB - wn * 32 + 7) >> 3));
    double acc[8][4][2];
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;

    for (int kc = 0; kc < nchunks; ++kc) {
        const int s = kc % SY_STAGES;
        mbar_wait(&full[s], (kc / SY_STAGES) & 1);
        const double *sA = stage0 + (size_t)s * SY_STAGE_DOUBLES;
        const double *sB = same ? sA : sA + SY_BK * SY_LDS;
        const int nk = min(SY_BK, kvalid - kc * SY_BK);  // multiple of 4
        if (ma == 8 && nbt == 4 && nk == SY_BK) {
#pragma unroll
            for (int k4 = 0; k4 < SY_BK; k4 += 4) {
                double af[8], bf[4];
                const double *pa = sA + (k4 + q) * SY_LDS + wm * 64 + g;
                const double *pb = sB + (k4 + q) * SY_LDS + wn * 32 + g;
#pragma unroll
                for (int a = 0; a < 8; ++a) af[a] = pa[a * 8];
#pragma unroll
                for (int b = 0; b < 4; ++b) bf[b] = pb[b * 8];
#pragma unroll
                for (int a = 0; a < 8; ++a)
#pragma unroll
                    for (int b = 0; b < 4; ++b) dmma884(acc[a][b][0], acc[a][b][1], af[a], bf[b]);
            }
        } else if (ma > 0 && nbt > 0) {
            for (int k4 = 0; k4 < nk; k4 += 4) {
                const double *pa = sA + (k4 + q) * SY_LDS + wm * 64 + g;
                const double *pb = sB + (k4 + q) * SY_LDS + wn * 32 + g;
                double bf[4];
#pragma unroll
                for (int b = 0; b < 4; ++b) bf[b] = (b < nbt) ? pb[b * 8] : 0.0;
#pragma unroll
                for (int a = 0; a < 8; ++a) {
                    if (a < ma) {
                        const double af = pa[a * 8];
#pragma unroll
                        for (int b = 0; b < 4; ++b)
                            if (b < nbt) dmma884(acc[a][b][0], acc[a][b][1], af, bf[b]);
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
    }

    // epilogue.  Every output element has exactly one owner CTA per launch (no split-K), so the RED is deterministic.
#pragma unroll
    for (int a = 0; a < 8; ++a) {
        const int i = wm * 64 + a * 8 + g;
        if (i >= ni) continue;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int j = wn * 32 + b * 8 + 2 * q + e;
                if (j >= nj) continue;
                if (lower_only && i < j) continue;
                double *dst = transposed ? C + (size_t)i * ld + j : C + (size_t)j * ld + i;
                if (mode == 0) red_add_f64(dst, -acc[a][b][e]);
                else *dst = acc[a][b][e];
            }
        }
    }
}

// Diagonal block of panel k: Cholesky of the nb x nb block in shared memory, then its inverse, both blocked by 32.
//   factor : per 32-wide sub-panel -- one warp factorises the 32 x 32 diagonal block in registers (lane = row,
//            columns exchanged by shuffles), threads-per-row solve the rows below it, all threads apply the rank-32
//            update to the trailing block;
//   invert : X = L^-1.  Warp b inverts diagonal sub-block b (lane = column, forward substitution in registers);
//            the off-diagonal blocks follow by block rows, X_ij = -X_ii (sum_m L_im X_mj), with all threads.
// One array holds both: the lower triangle is L, the strictly upper triangle collects X^T, diag(X) sits in dinv.
// Writes L back (lower part only: the upper triangle of Gaug still holds A), and the zero-padded 128 x 128 blocks
// Linv (element (c, m) at c + 128 m) and LinvT.
// A non-positive or non-finite pivot sets info = 1-based column (dpotrf convention) and the region drops out.
// Cholesky of one 32 x 32 diagonal sub-block by one warp, entirely in registers: lane i holds row i, the column
// being eliminated travels by shuffles.  Rows/columns beyond w are padded with the identity.  Returns the 1-based
// column of the first non-positive / non-finite pivot, or 0 (then the factor is written back).
__device__ __noinline__ int chol32_warp(double *__restrict__ Sb, int LD, int w, int lane)
{
    double a[CH_SB];
    const int i = lane;
#pragma unroll
    for (int c = 0; c < CH_SB; ++c) a[c] = (i < w && c < w && c <= i) ? Sb[i * LD + c] : ((i == c) ? 1.0 : 0.0);
    int failcol = 0;
#pragma unroll
    for (int j = 0; j < CH_SB; ++j) {
        const double d = __shfl_sync(0xffffffffu, a[j], j);
        if (failcol == 0 && (!(d > 0.0) || !isfinite(d))) failcol = j + 1;
        const double r = sqrt(d);
        const double l = (i == j) ? r : a[j] / r;   // lanes above the diagonal carry zeros
        a[j] = l;
#pragma unroll
        for (int c = j + 1; c < CH_SB; ++c) {
            const double lc = __shfl_sync(0xffffffffu, l, c);
            if (i >= c) a[c] = fma(-l, lc, a[c]);
        }
    }
    if (failcol) return failcol;
#pragma unroll
    for (int c = 0; c < CH_SB; ++c)
        if (i < w && c < w && c <= i) Sb[i * LD + c] = a[c];
    return 0;
}

__global__ void __launch_bounds__(256, 1)
k_chol_diag(const TrainRegionDev *__restrict__ T, int k)
{
    extern __shared__ __align__(16) double ch_s[];
    constexpr int LD = CH_LD;
    double *S = ch_s;                      // [128][129]
    double *dinv = ch_s + CH_NB * LD;      // [128]
    double *Tm = dinv + CH_NB;             // [32][97] scratch of the inversion
    __shared__ int s_fail;
    const TrainRegionDev &t = T[blockIdx.x];
    if (*t.chol_info != 0) return;
    const int ld = t.ld, N = t.R.n + t.R.S;
    const int j0 = k * CH_NB;
    if (j0 >= N) return;
    const int nb = min(CH_NB, N - j0);
    double *G = t.gram + (size_t)j0 * ld + j0;
    const int tid = threadIdx.x, nt = blockDim.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) s_fail = 0;
    for (int e = tid; e < nb * nb; e += nt) {
        const int i = e % nb, c = e / nb;
        S[i * LD + c] = (i >= c) ? G[(size_t)c * ld + i] : 0.0;
    }
    __syncthreads();

    // ---------------- blocked factorisation
    for (int b0 = 0; b0 < nb; b0 += CH_SB) {
        const int w = min(CH_SB, nb - b0);
        if (warp == 0) {
            const int failcol = chol32_warp(S + b0 * LD + b0, LD, w, lane);
            if (failcol && lane == 0) s_fail = b0 + failcol;
        }
        __syncthreads();
        if (s_fail) {
            if (tid == 0) *t.chol_info = j0 + s_fail;
            return;
        }
        const int below = nb - b0 - w;   // rows under the sub-block; > 0 only when w == 32
        if (below > 0) {
            if (tid < below) {
                const int r = b0 + CH_SB + tid;
                double x[CH_SB];
#pragma unroll
                for (int c = 0; c < CH_SB; ++c) {
                    double v = S[r * LD + b0 + c];
#pragma unroll
                    for (int m = 0; m < c; ++m) v = fma(-x[m], S[(b0 + c) * LD + b0 + m], v);
                    x[c] = v / S[(b0 + c) * LD + b0 + c];
                }
#pragma unroll
                for (int c = 0; c < CH_SB; ++c) S[r * LD + b0 + c] = x[c];
            }
            __syncthreads();
            const int t0 = b0 + CH_SB;
            for (int e = tid; e < below * below; e += nt) {
                const int i = t0 + e % below, c = t0 + e / below;
                if (i < c) continue;
                const double *ri = S + i * LD + b0, *rc = S + c * LD + b0;
                double v = S[i * LD + c];
#pragma unroll 8
                for (int m = 0; m < CH_SB; ++m) v = fma(-ri[m], rc[m], v);
                S[i * LD + c] = v;
            }
            __syncthreads();
        }
    }

    // ---------------- inverse of the diagonal sub-blocks: warp b, lane = column
    if (warp < (nb + CH_SB - 1) / CH_SB) {
        const int b0 = warp * CH_SB, w = min(CH_SB, nb - b0), c = lane;
        double x[CH_SB];
        // branch-free so that x[] stays in registers: x[m] = 0 for m < c makes the full-length dot product exact
#pragma unroll
        for (int i = 0; i < CH_SB; ++i) {
            const bool in = i < w;
            double acc = 0.0;
#pragma unroll
            for (int m = 0; m < i; ++m) acc = fma(in ? S[(b0 + i) * LD + b0 + m] : 0.0, x[m], acc);
            const double dii = in ? S[(b0 + i) * LD + b0 + i] : 1.0;
            x[i] = (i == c) ? 1.0 / dii : ((i > c) ? -acc / dii : 0.0);
        }
        __syncwarp();
        if (c < w) {
#pragma unroll
            for (int i = 0; i < CH_SB; ++i) {
                if (i == c) dinv[b0 + c] = x[i];                                // static index keeps x[] in registers
                else if (i > c && i < w) S[(b0 + c) * LD + b0 + i] = x[i];      // X[b0+i][b0+c] kept transposed
            }
        }
    }
    __syncthreads();
    // X[m][c] for m >= c (all already computed): dinv on the diagonal, else the transposed store
#define CH_X(m, c) (((m) == (c)) ? dinv[(c)] : S[(c) * LD + (m)])
    for (int r0 = CH_SB; r0 < nb; r0 += CH_SB) {
        const int w = min(CH_SB, nb - r0);
        // Tm[i][c] = sum_{m=c}^{r0-1} L[r0+i][m] X[m][c],  i < w, c < r0
        for (int e = tid; e < w * r0; e += nt) {
            const int i = e % w, c = e / w;
            const double *Lrow = S + (r0 + i) * LD;
            double acc = Lrow[c] * dinv[c];
            for (int m = c + 1; m < r0; ++m) acc = fma(Lrow[m], S[c * LD + m], acc);
            Tm[i * CH_TLD + c] = acc;
        }
        __syncthreads();
        // X[r0+i][c] = -sum_{m<=i} X_ii[i][m] Tm[m][c]
        for (int e = tid; e < w * r0; e += nt) {
            const int i = e % w, c = e / w;
            double acc = dinv[r0 + i] * Tm[i * CH_TLD + c];
            for (int m = 0; m < i; ++m) acc = fma(S[(r0 + m) * LD + r0 + i], Tm[m * CH_TLD + c], acc);
            S[c * LD + r0 + i] = -acc;
        }
        __syncthreads();
    }

    double *linv = t.linv + (size_t)k * CH_LINV, *linvT = linv + CH_NB * CH_NB;
    for (int e = tid; e < CH_NB * CH_NB; e += nt) {
        const int r = e % CH_NB, m = e / CH_NB;  // element (row r, col m)
        double x = 0.0, xt = 0.0;
        if (r < nb && m < nb) {
            if (r == m) x = xt = dinv[r];
            else if (r > m) x = CH_X(r, m);      // X[r][m]
            else xt = CH_X(m, r);                // X^T[r][m] = X[m][r]
        }
        linv[e] = x;
        linvT[e] = xt;
    }
#undef CH_X
    for (int e = tid; e < nb * nb; e += nt) {
        const int i = e % nb, c = e / nb;
        if (i >= c) G[(size_t)c * ld + i] = S[i * LD + c];
    }
}

constexpr size_t CH_DIAG_SMEM = (size_t)(CH_NB * CH_LD + CH_NB + CH_SB * CH_TLD) * 8;

// diagonal of the regularised A, saved before the factorisation overwrites it (LU fallback)
__global__ void k_chol_save_diag(const TrainRegionDev *__restrict__ T)
{
    const TrainRegionDev &t = T[blockIdx.y];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < t.ld) t.dsave[i] = t.gram[(size_t)t.ld * i + i];
}

// a region whose Cholesky failed: lower triangle <- upper triangle (the untouched copy of A), diagonal <- dsave.
// grid (ceil(ld/32), ceil(ld/32)) for ONE region
__global__ void k_chol_restore(double *__restrict__ G, const double *__restrict__ dsave, int ld)
{
    __shared__ double tile[32][33];
    const int bi = blockIdx.x, bj = blockIdx.y;
    if (bi < bj) return;
    const int tx = threadIdx.x, ty = threadIdx.y;
    for (int r = ty; r < 32; r += blockDim.y) {
        const int i = bj * 32 + tx, j = bi * 32 + r;  // upper element (i, j), i <= j
        tile[r][tx] = (i < ld && j < ld) ? G[(size_t)ld * j + i] : 0.0;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += blockDim.y) {
        const int i = bi * 32 + tx, j = bj * 32 + r;  // lower element (i, j) <- upper (j, i)
        if (i < ld && j < ld) {
            if (i > j) G[(size_t)ld * j + i] = tile[tx][r];
            else if (i == j) G[(size_t)ld * j + i] = dsave[i];
        }
    }
}

// reservoir%wout = transpose(b_trans) (src/mod_reservoir.f90:1313): after the back substitution W_out (P x N) sits in
// rows N..N+P of Gaug
__global__ void k_chol_store_wout(const TrainRegionDev *__restrict__ T, int region_in_wave, double *__restrict__ wout, int ldw)
{
    const TrainRegionDev &t = T[region_in_wave];
    const int N = t.R.n + t.R.S, P = t.R.P;
    const int j = blockIdx.x;
    if (j >= N) return;
    for (int p = threadIdx.x; p < P; p += blockDim.x) wout[(size_t)ldw * j + p] = t.gram[(size_t)t.ld * j + N + p];
}

}  // namespace sml
