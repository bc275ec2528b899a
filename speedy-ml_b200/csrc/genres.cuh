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

}  // namespace sml
