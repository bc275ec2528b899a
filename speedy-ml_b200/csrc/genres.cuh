// genres.cuh -- reservoir construction on the device: the spectral radius the reference gets
// from ARPACK (sparse_eigen, src/mod_linalg.f90:220-514, called by gen_res, src/mod_reservoir.f90:182-212) and
// the rescale  vals = (vals / eig) * radius  (:191-196), for every local region at once.
//
// The adjacency is entry-wise non-negative (vals ~ U[0,1), makesparse :180-218), so its largest-magnitude
// eigenvalue -- what dnaupd/dneupd with which='LM' return and maxval(d) picks (:246,511) -- is the Perron root,
// and power iteration from a positive start vector converges to it monotonically in direction.  Each iteration
// is one batched ELL SpMV (thread per row, as in the state update) plus a fixed-order norm reduction, so the
// result is deterministic.
#pragma once
#include "kernels.cuh"

namespace sml {

constexpr int GR_BLOCK = 256;

// y = A x for every region; per-block sums of y^2 into part[region][block].   grid (ceil(n_max/256), nregions)
__global__ void k_eig_spmv(const RegionDev *__restrict__ regs, const double *__restrict__ x, double *__restrict__ y,
                           double *__restrict__ part, int nblk)
{
    __shared__ double red[GR_BLOCK];
    const RegionDev &R = regs[blockIdx.y];
    const int row = blockIdx.x * GR_BLOCK + threadIdx.x;
    double v = 0.0;
    if (row < R.n) {
        const double *xo = x + R.x_off;
        const int *ec = R.ell_col + row;
        const double *ev = R.ell_val + row;
        for (int s = 0; s < R.ell_w; ++s) v = fma(ev[(size_t)s * R.n], xo[ec[(size_t)s * R.n]], v);
        y[R.x_off + row] = v;
    }
    red[threadIdx.x] = v * v;
    __syncthreads();
    for (int o = GR_BLOCK / 2; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) part[(size_t)blockIdx.y * nblk + blockIdx.x] = red[0];
}

// lambda = ||y|| (with ||x|| = 1), x <- y / lambda; lam[region] = {current, previous}.  One block per region.
__global__ void k_eig_normalize(const RegionDev *__restrict__ regs, const double *__restrict__ y, double *__restrict__ x,
                                const double *__restrict__ part, int nblk, double *__restrict__ lam)
{
    __shared__ double s_norm;
    const RegionDev &R = regs[blockIdx.x];
    if (R.n == 0) return;
    if (threadIdx.x == 0) {
        double s = 0.0;
        const int nb = (R.n + GR_BLOCK - 1) / GR_BLOCK;
        for (int b = 0; b < nb; ++b) s += part[(size_t)blockIdx.x * nblk + b];
        s_norm = sqrt(s);
        lam[2 * blockIdx.x + 1] = lam[2 * blockIdx.x];
        lam[2 * blockIdx.x] = s_norm;
    }
    __syncthreads();
    const double inv = s_norm > 0.0 ? 1.0 / s_norm : 0.0;
    for (int i = threadIdx.x; i < R.n; i += blockDim.x) x[R.x_off + i] = y[R.x_off + i] * inv;
}

__global__ void k_eig_init(const RegionDev *__restrict__ regs, double *__restrict__ x)
{
    const RegionDev &R = regs[blockIdx.x];
    if (R.n == 0) return;
    const double v = 1.0 / sqrt((double)R.n);
    for (int i = threadIdx.x; i < R.n; i += blockDim.x) x[R.x_off + i] = v;
}

// vals = vals * factor[region]  (gen_res :191-196 with factor = radius / eig; rounding: one multiply per entry)
__global__ void k_adj_scale(const RegionDev *__restrict__ regs, const double *__restrict__ factor)
{
    const RegionDev &R = regs[blockIdx.y];
    const size_t total = (size_t)R.ell_w * R.n;
    double *ev = const_cast<double *>(R.ell_val);
    const double f = factor[blockIdx.y];
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) ev[i] *= f;
}


// ---------------------------------------------------------------------------------------------
// makesparse + the W_in build on the device: the random structure of the reservoir
// (src/mod_linalg.f90:180-218, shuffle src/mod_utilities.f90:1569-1596, W_in src/mod_reservoir.f90:262-283).
// The Fortran random_number stream cannot be reproduced, so the draws come from the engine's counter-based generator
// (mix64, train.cuh) keyed by (seed, region, stream, index) -- the C and NumPy oracles restate the same streams, so the
// structure is comparable bit for bit.  Streams: 0 vals, 1 + 2*round row shuffle, 2 + 2*round column shuffle, GEN_STREAM_WIN.
// ---------------------------------------------------------------------------------------------
constexpr int GEN_STREAM_WIN = 4096;

struct GenDesc {
    int local;            // local region index (RegionDev table)
    int region;           // region id (keys the generator)
    int n, k, D, W;       // W = ELL width = number of shuffle rounds
    unsigned long long seed;
    double sigma;
    int *coo_rows, *coo_cols;    // [k] 1-based, makesparse's entry order
    double *coo_vals;            // [k]
    int *ell_col;                // [W][n] slot-major, zero-initialised (padding = (col 0, val 0.0))
    double *ell_val;
    double *winc;                // [n]
    int *wcol;                   // [n]
};

__device__ __forceinline__ unsigned long long counter_bits(unsigned long long seed, int region, int stream, long long index)
{
    const unsigned long long h1 = mix64(seed ^ ((unsigned long long)(unsigned)region << 32) ^ (unsigned long long)(unsigned)stream);
    return mix64(h1 + (unsigned long long)index);
}

// One CTA per (shuffle, region): blockIdx.x = 2*round + (0 rows | 1 cols).  The k-shuffle is a sequential chain (each pick
// depends on the swaps before it), so ONE thread walks it -- in shared memory, with the draws precomputed by the
// whole block -- while the launch runs thousands of such chains side by side.  `this = a*(n - n_chosen) + 1` is evaluated
// in single precision like the reference's default-real `a`, and clamped to the live range (see the oracle).
__global__ void __launch_bounds__(256)
k_makesparse_shuffle(const GenDesc *__restrict__ gd, int n_cap)
{
    extern __shared__ int ms_smem[];
    int *choices = ms_smem;                                   // [n]
    int *out = ms_smem + n_cap;                               // [n]
    float *draws = reinterpret_cast<float *>(ms_smem + 2 * n_cap);   // [n]
    const GenDesc g = gd[blockIdx.y];
    const int round = blockIdx.x >> 1, which = blockIdx.x & 1;
    const int n = g.n;
    int rounds_full, leftover;
    if (g.k > n) { rounds_full = g.k / n; leftover = g.k % n; }
    else { rounds_full = 0; leftover = g.k; }
    int size;
    if (round < rounds_full) size = n;
    else if (round == rounds_full && leftover > 0) size = leftover;
    else return;
    const int stream = 1 + 2 * round + which;
    for (int i = threadIdx.x; i < n; i += blockDim.x) choices[i] = i + 1;
    for (int i = threadIdx.x; i < size; i += blockDim.x)
        draws[i] = (float)(counter_bits(g.seed, g.region, stream, i) >> 40) * (1.0f / 16777216.0f);
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 0; i < size; ++i) {
            const int live = n - i;
            int pick = (int)__fadd_rn(__fmul_rn(draws[i], (float)live), 1.0f);
            pick = min(pick, live);
            const int tmp = choices[pick - 1];
            out[i] = tmp;
            choices[pick - 1] = choices[live - 1];
            choices[live - 1] = tmp;
        }
    }
    __syncthreads();
    int *dst = (which ? g.coo_cols : g.coo_rows) + (size_t)round * n;
    for (int i = threadIdx.x; i < size; i += blockDim.x) dst[i] = out[i];
}

// vals = random_number, COO -> slot-major ELL (entry e belongs to shuffle round e / n, in which every row occurs at
// most once, so its slot IS the round: the per-row entry order of the COO is kept), and the W_in build.
// grid (ceil(max(k, n) / 256), regions)
__global__ void __launch_bounds__(256)
k_makesparse_fill(const GenDesc *__restrict__ gd)
{
    const GenDesc g = gd[blockIdx.y];
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < g.k) {
        const double v = (double)(counter_bits(g.seed, g.region, 0, e) >> 11) * (1.0 / 9007199254740992.0);
        g.coo_vals[e] = v;
        const int slot = e / g.n, r = g.coo_rows[e] - 1;
        g.ell_col[(size_t)slot * g.n + r] = g.coo_cols[e] - 1;
        g.ell_val[(size_t)slot * g.n + r] = v;
    }
    if (e < g.n) {
        const int q = g.n / g.D;
        const int col = e / q;
        if (col < g.D) {
            const double rnd = (double)(counter_bits(g.seed, g.region, GEN_STREAM_WIN, e) >> 11) * (1.0 / 9007199254740992.0);
            const double ip = __dadd_rn(-1.0, __dmul_rn(2.0, rnd));
            g.winc[e] = __dmul_rn(g.sigma, ip);
            g.wcol[e] = col;
        } else {
            g.winc[e] = 0.0;
            g.wcol[e] = 0;
        }
    }
}

// the rescale of gen_res applied to the COO copy as well, so that sml_region_coo_get returns reservoir%vals
__global__ void k_coo_scale(const GenDesc *__restrict__ gd, const int *__restrict__ gen_of_local, const double *__restrict__ factor)
{
    const int gi = gen_of_local[blockIdx.y];
    if (gi < 0) return;
    const GenDesc g = gd[gi];
    const double f = factor[blockIdx.y];
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < g.k; e += gridDim.x * blockDim.x) g.coo_vals[e] *= f;
}

}  // namespace sml
