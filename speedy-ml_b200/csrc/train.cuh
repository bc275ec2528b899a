// train.cuh -- training path: state generation, FP64 tensor-core (DMMA) Gram accumulation, ridge terms.
//
// One augmented slab per region holds, column by column (one column per kept reservoir state):
//     rows [0, S)        imperfect-model (SPEEDY) forecast valid at the target time
//     rows [S, S+n)      x~ : the state with even 1-based entries squared
//     rows [N, N+P)      the target (truth at t+1, region interior rows of the input series), N = S+n
// and ONE symmetric accumulation  Gaug += slab * slab^T  on the lower triangle gives both reference
// accumulators at once:  states_x_states_aug = Gaug[0:N,0:N]  and  states_x_trainingdata_aug = Gaug[N:N+P,0:N]
// (chunking_matmul, src/mod_reservoir.f90:1645-1701: R*R^T via DGEMM and target*aug^T via matmul).
#pragma once
#include "kernels.cuh"
#include <cuda_runtime.h>
#include <map>
#include <vector>

namespace sml {

struct TrainRegionDev {
    RegionDev R;              // the reservoir's weights (ELL adjacency, W_in, ...)
    double *slab;             // augmented state slab, ROW-BLOCK-major and padded (slab_at): block b = rows 128b..128b+127,
                              // inside a block column k is a run of 132 doubles (128 rows + 4 pad) -- exactly the padded
                              // shared-memory operand layout of k_syrk_dmma, so a K-chunk of a tile operand is ONE bulk copy
    double *gram;             // [ld][ld] column-major, lower triangle accumulated
    double *xa, *xb;          // training state ping-pong [n]
    const double *td;         // [D][ncols] this phase's (pre-noised) input series
    const double *im;         // [S][ncols] imperfect-model series (hybrid) or null
    const int *target_map;    // [P] rows of the input vector that form the target
    int ld;                   // padded N+P (multiple of 16)
    int ks_total;             // slab columns per row block (KS, or 2*KS with two buffers in overlap mode)
    // ridge solve (chol.cuh)
    double *linv;             // [ceil(N/128)][2][128*128]: inverse of every diagonal block of L, and its transpose
    double *dsave;            // [ld] diagonal of the regularised A (restored if the region falls back to LU)
    int *chol_info;           // 0: on the Cholesky path; > 0: non-positive pivot at that column; < 0: not eligible
    int region;               // reservoir%assigned_region (keys the input-noise generator)
    int precip_off, precip_len;  // rows of the input vector that hold precip (noised in linear space), -1 / 0 if none
    long long pack_off;       // the region's tiles in the kind's tile-major adjacency pack (k_sync_pack), k_train_stategen_ring
};

// element (row i, column k) of a region's slab; SLAB_LD == SY_LDS (static_assert below)
constexpr int SLAB_RB = 128, SLAB_LD = 132;
__host__ __device__ __forceinline__ size_t slab_at(int ks_total, int i, int k)
{
    return ((size_t)(i >> 7) * ks_total + k) * SLAB_LD + (i & 127);
}
__host__ __device__ __forceinline__ size_t slab_doubles(int ld, int ks_total)
{
    return (size_t)((ld + SLAB_RB - 1) / SLAB_RB) * ks_total * SLAB_LD;
}

// Device-resident global training series (sml_train_global_series): column t is the conditioned global state at time
// t in the layout of the exchange buffers, G = [w4d | w2d | precip | sst | tisr] and F = [f4d | f2d].  With it the
// training kernels tile and standardise every region's input, imperfect-model and target columns on the fly --
// the arithmetic of tile_4d_and_logp_to_local_state_input + standardize_state_vec_input (src/res_domain.f90:1081-1125,
// 1211-1268) exactly as k_build_inputs does it for the forecast -- instead of reading per-region series.
struct GlobalSeries {
    const double *G = nullptr;   // [ncols_total][g_len]; null: per-region series (TrainRegionDev::td / im)
    const double *F = nullptr;   // [ncols_total][f_len]
    long long g_len = 0, f_len = 0;
    int first = 0, stride = 1;   // phase column c <-> global column first + stride * c
    // multiplicative input noise of the training state generation (gaussian_noise_1d_function(_precip),
    // src/mod_utilities.f90:1387-1464; reservoir%noisemag = 0.2, src/res_domain.f90:1607-1616); 0: off
    double noisemag = 0.0, precip_eps = 0.001;
    unsigned long long seed = 0;
};

// Counter-based standard normal: the draw for (seed, region, global column, input element) is a pure function of those
// four numbers (splitmix64 finaliser twice, Box-Muller), so every row that reads input element c sees the same noise
// and a run is reproducible.  This is the engine's generator; the reference's random_number stream cannot be reproduced.
__device__ __forceinline__ unsigned long long mix64(unsigned long long z)
{
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
__device__ __forceinline__ double counter_gauss(unsigned long long seed, int region, int col, int c)
{
    const unsigned long long k = mix64(seed ^ ((unsigned long long)(unsigned)region << 42) ^ ((unsigned long long)(unsigned)col << 21) ^
                                       (unsigned long long)(unsigned)c);
    const unsigned long long a = mix64(k), b = mix64(k ^ 0xD1B54A32D192ED03ULL);
    const double u1 = ((double)(a >> 11) + 1.0) * (1.0 / 9007199254740993.0);   // (0, 1)
    const double u2 = (double)(b >> 11) * (1.0 / 9007199254740992.0);           // [0, 1)
    return sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
}

// the input element the state update sees: standardised value, then the multiplicative noise -- precip in linear space
__device__ __forceinline__ double noised_value(const GlobalSeries &gs, const RegionDev &R, int region, int precip_off,
                                               int precip_len, int col, int c, double v, double g)
{
    const double nm = gs.noisemag;
    if (c >= precip_off && c < precip_off + precip_len) {
        const int ms = R.fb_ms[c];
        double t = __dadd_rn(__dmul_rn(v, R.std[ms]), R.mean[ms]);
        t = __dmul_rn(gs.precip_eps, exp(t) - 1.0);
        t = __dadd_rn(t, __dmul_rn(__dmul_rn(g, nm), t));
        t = fabs(t);
        t = log(1.0 + t / gs.precip_eps);
        return __ddiv_rn(__dsub_rn(t, R.mean[ms]), R.std[ms]);
    }
    return __dadd_rn(v, __dmul_rn(__dmul_rn(g, nm), v));
}

__device__ __forceinline__ double series_input(const GlobalSeries &gs, const RegionDev &R, int col, int c)
{
    double v = gs.G[(size_t)(gs.first + gs.stride * col) * gs.g_len + R.fb_src[c]];
    const int ms = R.fb_ms[c];
    if (ms >= 0) v = __ddiv_rn(__dsub_rn(v, R.mean[ms]), R.std[ms]);
    return v;
}

// input element c of phase column col as the state update of training sees it: tiled + standardised, then noised
__device__ __forceinline__ double train_input(const GlobalSeries &gs, const TrainRegionDev &t, int col, int c)
{
    double v = series_input(gs, t.R, col, c);
    if (gs.noisemag != 0.0) {
        const double g = counter_gauss(gs.seed, t.region, gs.first + gs.stride * col, c);
        v = noised_value(gs, t.R, t.region, t.precip_off, t.precip_len, col, c, v, g);
    }
    return v;
}

// inspection kernel for the tests: clean / gaussian / noised value of every input element of one region and column
__global__ void k_train_noise_sample(const TrainRegionDev *__restrict__ T, int wave_index, int col, GlobalSeries gs,
                                     double *__restrict__ clean, double *__restrict__ gauss, double *__restrict__ noisy)
{
    const TrainRegionDev &t = T[wave_index];
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < t.R.D; c += gridDim.x * blockDim.x) {
        clean[c] = series_input(gs, t.R, col, c);
        gauss[c] = counter_gauss(gs.seed, t.region, gs.first + gs.stride * col, c);
        noisy[c] = train_input(gs, t, col, c);
    }
}

// one reservoir step for every region of the wave: reads input column in_col, writes the new state to the
// other ping-pong buffer and (out_col >= 0) x~ into slab column out_col.  gather_col >= 0 takes the SpMV
// operand from slab column gather_col instead of the state -- the ML-only restart of the reference, which
// feeds the squared copy states(:,batch_size) back (src/mod_reservoir.f90:1034; slab ocean :933).
// reservoir_layer_chunking_hybrid/_ml, src/mod_reservoir.f90:963-1175.   grid (ceil(n_max/256), nwave)
__global__ void k_train_update(const TrainRegionDev *__restrict__ T, int parity, int in_col, int out_col, int gather_col,
                               GlobalSeries gs)
{
    const TrainRegionDev &t = T[blockIdx.y];
    const RegionDev &R = t.R;
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= R.n) return;
    const double *xo = parity ? t.xb : t.xa;
    double *xn = parity ? t.xa : t.xb;
    const bool from_slab = gather_col >= 0;   // ML-only restart: the SpMV operand is x~ of the previous column
    const double *u = gs.G ? nullptr : t.td + (size_t)R.D * in_col;
    const int n = R.n;
    const int *__restrict__ ec = R.ell_col + row;
    const double *__restrict__ ev = R.ell_val + row;
    double acc = 0.0;
    for (int s = 0; s < R.ell_w; ++s) {
        const int c = __ldg(ec + (size_t)s * n);
        acc = fma(__ldg(ev + (size_t)s * n), from_slab ? t.slab[slab_at(t.ks_total, R.S + c, gather_col)] : xo[c], acc);
    }
    double tw;
    if (R.win_mode == 0) {
        const int wc = __ldg(R.wcol + row);
        tw = __dmul_rn(__ldg(R.winc + row), gs.G ? train_input(gs, t, in_col, wc) : u[wc]);
    } else {
        tw = 0.0;
        for (int i = 0; i < R.D; ++i)
            tw = fma(R.win_dense[(size_t)i * n + row], gs.G ? train_input(gs, t, in_col, i) : u[i], tw);
    }
    const double xt = tanh(__dadd_rn(acc, tw));
    const double xv = __dadd_rn(__dmul_rn(1.0 - R.leak, xo[row]), __dmul_rn(R.leak, xt));
    xn[row] = xv;
    if (out_col >= 0) t.slab[slab_at(t.ks_total, R.S + row, out_col)] = (row & 1) ? __dmul_rn(xv, xv) : xv;
}

// ---------------------------------------------------------------------------------------------
// k_train_stategen: the time loop of reservoir_layer_chunking_hybrid/_ml (src/mod_reservoir.f90:963-1175) INSIDE the
// kernel -- one CTA per region of the wave keeps the region's state vector in shared memory and advances it nsteps
// time steps, two block barriers per step, instead of one launch per time step over the whole wave (k_train_update).
// Per step: (a) the step's input vector is staged in shared memory (tiled + standardised + noised on the fly from the
// resident global series, or read from the per-region series; each input element is evaluated once, not once per
// row), (b) every thread computes its rows -- ELL slots streamed from global memory (L2-resident across steps), the x
// gathers from shared memory -- and parks the new values in the region's scratch vector xb, writing x~ into the slab
// column when the state is kept, (c) after a barrier each thread moves its own rows from xb into the shared vector.
// The state lives in xa between launches.  Arithmetic and order per row are those of k_train_update, so the states
// are bit-identical.  The ML-only restart (SpMV operand = the squared copy states(:,batch_size), :1034; slab ocean
// :933) is evaluated from the shared vector: the squared copy of the previous state is x~(x).
//   c_first..c_end-1 : slab-column range produced by this launch (state index s = s0 + c, input column discard+s-1);
//                      store_first: column c_first-1... see train_run_phase; out_base < 0: discard steps, nothing kept.
// grid (nwave); block 256 (fits beside a resident Gram CTA: 64 registers) or 512.
// ---------------------------------------------------------------------------------------------
constexpr int SG_MAX_THREADS = 512;
// G: ELL slots fetched per group for the thread's two rows (4*G loads in flight); <3, 512>: 64 registers, fits beside a
// resident Gram CTA at 256 threads; <6, 384>: up to 85 registers, twice the loads in flight when the kernel runs alone
template <int G, int MAXT>
__global__ void __launch_bounds__(MAXT, 2)
k_train_stategen(const TrainRegionDev *__restrict__ T, int in_col0, int nsteps, int out_col0, int store_first,
                 int s_first, int restart_period, int xs_cap, GlobalSeries gs)
{
    extern __shared__ __align__(16) double sg_smem[];
    const TrainRegionDev &t = T[blockIdx.x];
    const RegionDev &R = t.R;
    const int n = R.n, D = R.D, W = R.ell_w, S = R.S;
    const int tid = threadIdx.x, nt = blockDim.x;
    double *xs = sg_smem, *us = sg_smem + xs_cap;
    double *__restrict__ xa = t.xa, *__restrict__ xb = t.xb;
    for (int i = tid; i < n; i += nt) xs[i] = xa[i];
    __syncthreads();
    if (store_first) {   // states(:,1) = x after the discard loop: no update, only the x~ copy
        for (int i = tid; i < n; i += nt) {
            const double xv = xs[i];
            t.slab[slab_at(t.ks_total, S + i, out_col0 - 1)] = (i & 1) ? __dmul_rn(xv, xv) : xv;
        }
    }
    const int *__restrict__ ecol = R.ell_col;
    const double *__restrict__ eval = R.ell_val;
    const bool compact = R.win_mode == 0;
    const double leak = R.leak;
    for (int k = 0; k < nsteps; ++k) {
        const int in_col = in_col0 + k;
        // (a) this step's input vector
        if (gs.G) {
            for (int i = tid; i < D; i += nt) us[i] = train_input(gs, t, in_col, i);
        } else {
            const double *u = t.td + (size_t)D * in_col;
            for (int i = tid; i < D; i += nt) us[i] = u[i];
        }
        __syncthreads();
        // (b) rows of this thread
        const bool restart = restart_period > 0 && ((s_first + k) % restart_period) == 0;
        const bool keep = out_col0 >= 0;
        const int ocol = out_col0 + k;
        for (int base = tid; base < n; base += 2 * nt) {
            const int r0 = base, r1 = base + nt;
            const bool ok1 = r1 < n;
            const int q1 = ok1 ? r1 : r0;
            double acc0 = 0.0, acc1 = 0.0;
            for (int s = 0; s < W; s += G) {
                int c0[G], c1[G];
                double v0[G], v1[G];
#pragma unroll
                for (int j = 0; j < G; ++j) {
                    const bool in = s + j < W;
                    c0[j] = in ? __ldg(ecol + (size_t)(s + j) * n + r0) : 0;
                    v0[j] = in ? __ldg(eval + (size_t)(s + j) * n + r0) : 0.0;
                    c1[j] = in ? __ldg(ecol + (size_t)(s + j) * n + q1) : 0;
                    v1[j] = in ? __ldg(eval + (size_t)(s + j) * n + q1) : 0.0;
                }
#pragma unroll
                for (int j = 0; j < G; ++j) {
                    if (s + j < W) {
                        double g0 = xs[c0[j]], g1 = xs[c1[j]];
                        if (restart) {
                            if (c0[j] & 1) g0 = __dmul_rn(g0, g0);
                            if (c1[j] & 1) g1 = __dmul_rn(g1, g1);
                        }
                        acc0 = fma(v0[j], g0, acc0);
                        acc1 = fma(v1[j], g1, acc1);
                    }
                }
            }
            double tw0, tw1;
            if (compact) {
                tw0 = __dmul_rn(__ldg(R.winc + r0), us[__ldg(R.wcol + r0)]);
                tw1 = __dmul_rn(__ldg(R.winc + q1), us[__ldg(R.wcol + q1)]);
            } else {
                tw0 = tw1 = 0.0;
                for (int i = 0; i < D; ++i) {
                    tw0 = fma(R.win_dense[(size_t)i * n + r0], us[i], tw0);
                    tw1 = fma(R.win_dense[(size_t)i * n + q1], us[i], tw1);
                }
            }
            const double xv0 = __dadd_rn(__dmul_rn(1.0 - leak, xs[r0]), __dmul_rn(leak, tanh(__dadd_rn(acc0, tw0))));
            const double xv1 = __dadd_rn(__dmul_rn(1.0 - leak, xs[q1]), __dmul_rn(leak, tanh(__dadd_rn(acc1, tw1))));
            xb[r0] = xv0;
            if (keep) t.slab[slab_at(t.ks_total, S + r0, ocol)] = (r0 & 1) ? __dmul_rn(xv0, xv0) : xv0;
            if (ok1) {
                xb[r1] = xv1;
                if (keep) t.slab[slab_at(t.ks_total, S + r1, ocol)] = (r1 & 1) ? __dmul_rn(xv1, xv1) : xv1;
            }
        }
        __syncthreads();   // every gather of the old state is done
        // (c) own rows: xb -> shared (written by this very thread above)
        for (int i = tid; i < n; i += nt) xs[i] = xb[i];
        // the next step's barrier after (a) orders these writes before the next gathers
    }
    __syncthreads();
    for (int i = tid; i < n; i += nt) xa[i] = xs[i];
}

// the two halves of train_input for a kernel that wants the load in flight early: the raw series value, and
// standardisation + noise applied to it later.  finish(raw(...)) == train_input(...) bit for bit.
__device__ __forceinline__ double train_input_raw(const GlobalSeries &gs, const TrainRegionDev &t, int col, int c)
{
    if (gs.G) return gs.G[(size_t)(gs.first + gs.stride * col) * gs.g_len + t.R.fb_src[c]];
    return t.td[(size_t)t.R.D * col + c];
}
__device__ __forceinline__ double train_input_finish(const GlobalSeries &gs, const TrainRegionDev &t, int col, int c, double v)
{
    if (!gs.G) return v;
    const int ms = t.R.fb_ms[c];
    if (ms >= 0) v = __ddiv_rn(__dsub_rn(v, t.R.mean[ms]), t.R.std[ms]);
    if (gs.noisemag != 0.0) {
        const double g = counter_gauss(gs.seed, t.region, gs.first + gs.stride * col, c);
        v = noised_value(gs, t.R, t.region, t.precip_off, t.precip_len, col, c, v, g);
    }
    return v;
}

// ---------------------------------------------------------------------------------------------
// k_train_stategen_ring: the same time loop on the ring of k_sync_persist (kernels.cuh).  k_train_stategen fetches the
// ELL slots with ordinary loads -- a thread issues a group, waits an L2 round trip, gathers, issues the next: 26 us per
// step for a wave of 144 (2.9 TB/s although the adjacency is L2-resident).  Here ONE CTA per region and per SM keeps the
// state in shared memory (two copies, one consumer barrier per step), a TMA producer warp streams the region's
// tile-major adjacency pack (one bulk copy per tile) through the ring, and the consumers take one row per thread from
// shared memory only; the step's input vector is evaluated once per element by the first D consumer threads -- its
// raw loads are issued before the step's tiles and finished (standardise + noise) after them, so their latency hides
// behind the step.  x~ goes to the slab column as before.  Per-row arithmetic and order are those of k_train_update:
// bit-identical accumulators.  Needs compact W_in, n % 4 == 0, D <= 2 * consumer threads.
// grid (nwave): one wave region per CTA; waves larger than the SM count run in rounds.
// ---------------------------------------------------------------------------------------------
template <int VPT>
__global__ void __launch_bounds__(SP_MAX_THREADS, 1)
k_train_stategen_ring(const TrainRegionDev *__restrict__ T, int in_col0, int nsteps, int out_col0, int store_first, int s_first,
                      int restart_period, int xs_cap, int us_cap, int w_max, int nstages, int tr, int ngroups,
                      const unsigned char *__restrict__ pack, GlobalSeries gs)
{
    extern __shared__ __align__(128) unsigned char sr_smem[];
    const TrainRegionDev &t = T[blockIdx.x];
    const RegionDev &R = t.R;
    const int n = R.n, D = R.D, W = R.ell_w, S = R.S;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ncons = ngroups * tr;
    double *xs = reinterpret_cast<double *>(sr_smem), *us = xs + 2 * (size_t)xs_cap;
    unsigned char *ring = reinterpret_cast<unsigned char *>(us + 2 * (size_t)us_cap);
    const int tile_stride = tr * sp_row_bytes(w_max);
    uint64_t *full = reinterpret_cast<uint64_t *>(ring + (size_t)nstages * tile_stride);
    uint64_t *empty = full + nstages;
    const int ntiles = (n + tr - 1) / tr;
    if (tid == 0) {
        for (int s = 0; s < nstages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], tr / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async;" ::: "memory");
    }
    __syncthreads();

    if (warp == ncons / 32) {
        // ---------------- producer: nsteps passes over the region's tiles ----------------
        const unsigned char *__restrict__ tiles = pack + t.pack_off;
        int s = 0;
        uint32_t par = 1;
        for (int k = 0; k < nsteps; ++k)
            for (int j = 0; j < ntiles; ++j) {
                mbar_wait(&empty[s], par);
                if (lane == 0) {
                    mbar_expect_tx(&full[s], (uint32_t)tile_stride);
                    tma_load_1d(ring + (size_t)s * tile_stride, tiles + (size_t)j * tile_stride, (uint32_t)tile_stride, &full[s]);
                }
                if (++s == nstages) { s = 0; par ^= 1u; }
            }
        return;
    }

    // ---------------- consumers ----------------
    const int gi = tid / tr, gt = tid - gi * tr;
    const double leak = R.leak;
    double *__restrict__ xa = t.xa;
    for (int i = tid; i < n; i += ncons) xs[i] = xa[i];
    for (int i = tid; i < D && nsteps > 0; i += ncons) us[i] = train_input_finish(gs, t, in_col0, i, train_input_raw(gs, t, in_col0, i));
    asm volatile("bar.sync 1, %0;" ::"r"(ncons) : "memory");
    if (store_first) {   // states(:,1) = x after the discard loop: no update, only the x~ copy
        for (int i = tid; i < n; i += ncons) {
            const double xv = xs[i];
            t.slab[slab_at(t.ks_total, S + i, out_col0 - 1)] = (i & 1) ? __dmul_rn(xv, xv) : xv;
        }
    }
    int s = gi, j = gi;
    uint32_t par = 0;
    const int i0 = tid, i1 = tid + ncons;   // the input elements this thread evaluates for the next step
    for (int k = 0; k < nsteps; ++k) {
        const double *__restrict__ xr = xs + (size_t)(k & 1) * xs_cap;
        double *__restrict__ xw = xs + (size_t)((k + 1) & 1) * xs_cap;
        const double *__restrict__ uu = us + (size_t)(k & 1) * us_cap;
        double *__restrict__ un = us + (size_t)((k + 1) & 1) * us_cap;
        const bool more = k + 1 < nsteps;
        const int ncol = in_col0 + k + 1;
        double raw0 = 0.0, raw1 = 0.0;
        if (more) {
            if (i0 < D) raw0 = train_input_raw(gs, t, ncol, i0);
            if (i1 < D) raw1 = train_input_raw(gs, t, ncol, i1);
        }
        const bool restart = restart_period > 0 && ((s_first + k) % restart_period) == 0;
        const bool keep = out_col0 >= 0;
        const int ocol = out_col0 + k;
        for (; j < ntiles; j += ngroups) {
            mbar_wait(&full[s], par);
            const unsigned char *tile = ring + (size_t)s * tile_stride;
            const int row = j * tr + gt;
            if (row < n) {
                const double xv = restart ? sp_row<VPT, true>(tile, gt, tr, w_max, W, xr, uu, xr[row], leak)
                                          : sp_row<VPT, false>(tile, gt, tr, w_max, W, xr, uu, xr[row], leak);
                xw[row] = xv;
                if (keep) t.slab[slab_at(t.ks_total, S + row, ocol)] = (row & 1) ? __dmul_rn(xv, xv) : xv;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
            s += ngroups;
            if (s >= nstages) { s -= nstages; par ^= 1u; }
        }
        j -= ntiles;
        if (more) {
            if (i0 < D) un[i0] = train_input_finish(gs, t, ncol, i0, raw0);
            if (i1 < D) un[i1] = train_input_finish(gs, t, ncol, i1, raw1);
        }
        asm volatile("bar.sync 1, %0;" ::"r"(ncons) : "memory");
    }
    const double *__restrict__ xf = xs + (size_t)(nsteps & 1) * xs_cap;
    for (int i = tid; i < n; i += ncons) xa[i] = xf[i];
}

// copy the current state (no update) into slab column out_col: states(:,1) = x after the discard loop
__global__ void k_train_store_state(const TrainRegionDev *__restrict__ T, int parity, int out_col)
{
    const TrainRegionDev &t = T[blockIdx.y];
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= t.R.n) return;
    const double xv = (parity ? t.xb : t.xa)[row];
    t.slab[slab_at(t.ks_total, t.R.S + row, out_col)] = (row & 1) ? __dmul_rn(xv, xv) : xv;
}

// imperfect-model rows and target rows of slab columns [0, ncols): series column = first_series_col + c.
// chunking_matmul :1668 (imperfect), tile_full_input_to_target_data2d src/res_domain.f90:602-651 (target).
// Columns [ncols, kpad) are zeroed entirely (K padding of the tensor-core kernel).  col_base selects the slab
// buffer (0 or ks: the slab is double-buffered when state generation overlaps the Gram).  grid (kpad, nwave)
__global__ void k_train_fill(const TrainRegionDev *__restrict__ T, int first_series_col, int ncols, int kpad, int col_base,
                             GlobalSeries gs)
{
    const TrainRegionDev &t = T[blockIdx.y];
    const RegionDev &R = t.R;
    const int c = blockIdx.x;
    const int kst = t.ks_total, kc = col_base + c;
    auto at = [&](int i) -> double & { return t.slab[slab_at(kst, i, kc)]; };
    const int N = R.S + R.n;
    if (c >= ncols) {
        for (int i = threadIdx.x; i < t.ld; i += blockDim.x) at(i) = 0.0;
        return;
    }
    const int sc = first_series_col + c;
    if (gs.G) {
        // imperfect model: tile_4d_and_logp_full_grid_to_local_res_vec + standardize_state_vec_res of the forecast grid
        const double *Fc = gs.F ? gs.F + (size_t)(gs.first + gs.stride * sc) * gs.f_len : nullptr;
        for (int i = threadIdx.x; i < R.S; i += blockDim.x) {
            double v = Fc[R.lm_src[i]];
            const int ms = R.lm_ms[i];
            if (ms >= 0) v = __ddiv_rn(__dsub_rn(v, R.mean[ms]), R.std[ms]);
            at(i) = v;
        }
        for (int p = threadIdx.x; p < R.P; p += blockDim.x) at(N + p) = series_input(gs, R, sc, t.target_map[p]);
    } else {
        for (int i = threadIdx.x; i < R.S; i += blockDim.x) at(i) = t.im[(size_t)R.S * sc + i];
        for (int p = threadIdx.x; p < R.P; p += blockDim.x) at(N + p) = t.td[(size_t)R.D * sc + t.target_map[p]];
    }
    for (int i = N + R.P + threadIdx.x; i < t.ld; i += blockDim.x) at(i) = 0.0;
}

// ---------------------------------------------------------------------------------------------
// k_syrk_dmma: Gaug(lower) += slab * slab^T with FP64 tensor-core MMA (mma.sync m8n8k4 f64 -> DMMA).
//   CTA = one 128x128 tile (ti >= tj) of one region; K is walked in chunks of 16 columns.
//   warp 8 (producer): TMA bulk copies -- each slab column restricted to the tile's rows is one
//     contiguous run (column-major), so a chunk is 16 (+16) cp.async.bulk of <= 1 KB into padded rows.
//   warps 0..7 (2 x 4): each owns a 64x32 sub-tile = 8 x 4 DMMA tiles, 64 FP64 accumulators per thread.
// ---------------------------------------------------------------------------------------------
constexpr int SY_BM = 128, SY_BK = 16, SY_LDS = 132, SY_STAGES = 4;
constexpr int SY_CONS_WARPS = 8;
constexpr int SY_THREADS = (SY_CONS_WARPS + 1) * 32;
constexpr int SY_STAGE_DOUBLES = 2 * SY_BK * SY_LDS;
static_assert(SY_LDS == SLAB_LD && SY_BM == SLAB_RB, "the slab is stored in the Gram kernel's padded operand layout");
constexpr size_t SY_SMEM = (size_t)SY_STAGES * SY_STAGE_DOUBLES * 8 + 2 * SY_STAGES * 8;
constexpr int SY_STAGES_DEEP = 6;   // A/B: SML_SYRK_STAGES=6 (203 KB of shared memory)
constexpr size_t SY_SMEM_DEEP = (size_t)SY_STAGES_DEEP * SY_STAGE_DOUBLES * 8 + 2 * SY_STAGES_DEEP * 8;

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// fire-and-forget global FP64 reduction (RED.E.ADD.F64): no return value, so no scoreboard wait
__device__ __forceinline__ void red_add_f64(double *p, double v)
{
    asm volatile("red.global.add.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}

// WM x WN consumer warps, each owning a (128/WM) x (128/WN) sub-tile; <2,4> = 8 warps with 64 accumulators per thread
// (fewest shared-memory loads per DMMA), <4,4> = 16 warps with 32 (more warps per scheduler to cover the DMMA
// issue bubbles).  Launch with (WM*WN + 1) * 32 threads.
template <int WM, int WN, int ST = SY_STAGES>
__global__ void __launch_bounds__((WM * WN + 1) * 32, 1)
k_syrk_dmma(const TrainRegionDev *__restrict__ T, const int2 *__restrict__ tiles, int kpad, int col_base)
{
    constexpr int NCW = WM * WN;               // consumer warps
    constexpr int TR = SY_BM / WM, TC = SY_BM / WN;   // warp sub-tile
    constexpr int TA = TR / 8, TB = TC / 8;    // 8 x 8 DMMA tiles per warp
    extern __shared__ __align__(128) unsigned char sy_smem[];
    double *stage0 = reinterpret_cast<double *>(sy_smem);
    uint64_t *full = reinterpret_cast<uint64_t *>(sy_smem + (size_t)ST * SY_STAGE_DOUBLES * 8);
    uint64_t *empty = full + ST;

    const TrainRegionDev &t = T[blockIdx.y];
    const int2 tile = tiles[blockIdx.x];
    const int ld = t.ld;
    const int i0 = tile.x * SY_BM, j0 = tile.y * SY_BM;
    if (i0 >= ld) return;  // tile list is built for the widest region of the wave
    const bool diag = tile.x == tile.y;
    const int rowsA = min(SY_BM, ld - i0), rowsB = min(SY_BM, ld - j0);
    const int nchunks = kpad / SY_BK;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int s = 0; s < ST; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], NCW);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async;" ::: "memory");
    }
    __syncthreads();

    if (warp == NCW) {
        if (lane == 0) {
            // the slab is stored row-block-major with 132-double padded columns (slab_at): the K-chunk of a tile operand
            // is ONE contiguous run that lands in the padded shared-memory layout as it is.  (Column-major, a chunk was
            // 16 + 16 copies of 1 KB; at ~140 cycles of issue each the producer lane needed longer than the 4096 cycles
            // the tensor pipe takes for the chunk.)
            constexpr uint32_t OPB = (uint32_t)SY_BK * SY_LDS * 8u;   // bytes of one operand chunk
            const uint32_t bytes = diag ? OPB : 2u * OPB;
            const int kst = t.ks_total;
            for (int kc = 0; kc < nchunks; ++kc) {
                const int s = kc % ST;
                mbar_wait(&empty[s], ((kc / ST) & 1) ^ 1);
                mbar_expect_tx(&full[s], bytes);
                double *sA = stage0 + (size_t)s * SY_STAGE_DOUBLES;
                double *sB = sA + SY_BK * SY_LDS;
                const int k0 = col_base + kc * SY_BK;
                tma_load_1d(sA, t.slab + ((size_t)tile.x * kst + k0) * SLAB_LD, OPB, &full[s]);
                if (!diag) tma_load_1d(sB, t.slab + ((size_t)tile.y * kst + k0) * SLAB_LD, OPB, &full[s]);
            }
        }
        return;
    }

    const int wm = warp / WN, wn = warp % WN;
    const int g = lane >> 2, q = lane & 3;         // DMMA group / thread-in-group
    // 8x8 DMMA tiles of this warp that hold real rows / columns (edge tiles of the padded matrix are slivers)
    const int ma = max(0, min(TA, (rowsA - wm * TR + 7) >> 3));
    const int nb = max(0, min(TB, (rowsB - wn * TC + 7) >> 3));
    double acc[TA][TB][2];
#pragma unroll
    for (int a = 0; a < TA; ++a)
#pragma unroll
        for (int b = 0; b < TB; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;

    for (int kc = 0; kc < nchunks; ++kc) {
        const int s = kc % ST;
        mbar_wait(&full[s], (kc / ST) & 1);
        const double *sA = stage0 + (size_t)s * SY_STAGE_DOUBLES;
        const double *sB = diag ? sA : sA + SY_BK * SY_LDS;
        if (ma == TA && nb == TB) {
#pragma unroll
            for (int k4 = 0; k4 < SY_BK; k4 += 4) {
                double af[TA], bf[TB];
                const double *pa = sA + (k4 + q) * SY_LDS + wm * TR + g;
                const double *pb = sB + (k4 + q) * SY_LDS + wn * TC + g;
#pragma unroll
                for (int a = 0; a < TA; ++a) af[a] = pa[a * 8];
#pragma unroll
                for (int b = 0; b < TB; ++b) bf[b] = pb[b * 8];
#pragma unroll
                for (int a = 0; a < TA; ++a)
#pragma unroll
                    for (int b = 0; b < TB; ++b) dmma884(acc[a][b][0], acc[a][b][1], af[a], bf[b]);
            }
        } else if (ma > 0 && nb > 0) {
            // sliver tile: only the DMMA tiles with real rows/columns (warp-uniform predicates)
#pragma unroll
            for (int k4 = 0; k4 < SY_BK; k4 += 4) {
                const double *pa = sA + (k4 + q) * SY_LDS + wm * TR + g;
                const double *pb = sB + (k4 + q) * SY_LDS + wn * TC + g;
                double bf[TB];
#pragma unroll
                for (int b = 0; b < TB; ++b) bf[b] = pb[b * 8];
#pragma unroll
                for (int a = 0; a < TA; ++a) {
                    if (a < ma) {
                        const double af = pa[a * 8];
#pragma unroll
                        for (int b = 0; b < TB; ++b)
                            if (b < nb) dmma884(acc[a][b][0], acc[a][b][1], af, bf[b]);
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
    }

    // epilogue: Gaug[i, j] += acc ; thread holds C[g][2q], C[g][2q+1] of every 8x8 tile.  Every element of
    // the lower triangle is owned by exactly one CTA per launch, so a fire-and-forget reduction
    // (RED.ADD.F64) is deterministic and never stalls on the read of Gaug.
    const int rows_total = ld;
#pragma unroll
    for (int a = 0; a < TA; ++a) {
        const int i = i0 + wm * TR + a * 8 + g;
        if (i >= rows_total) continue;
#pragma unroll
        for (int b = 0; b < TB; ++b) {
            const int j = j0 + wn * TC + b * 8 + 2 * q;
            if (j < rows_total) red_add_f64(&t.gram[(size_t)ld * j + i], acc[a][b][0]);
            if (j + 1 < rows_total) red_add_f64(&t.gram[(size_t)ld * (j + 1) + i], acc[a][b][1]);
        }
    }
}

// ridge terms on the diagonal and the prior on the right-hand side (fit_chunk_hybrid
// src/mod_reservoir.f90:1261-1291,1308-1310; fit_chunk_ml :1203-1205), then mirror the lower triangle
// (incl. the target block) into the upper one so that A = Gaug[0:N,0:N] is a full matrix and
// B = Gaug[0:N, N:N+P] = states_x_trainingdata_aug^T.   grid: (ceil(ld/32), ceil(ld/32))
__global__ void k_train_ridge_mirror(double *__restrict__ G, int ld, int N, int S, int P, double add_model,
                                     double add_res, double prior_add, int ml_first_n)
{
    __shared__ double tile[32][33];
    const int bi = blockIdx.x, bj = blockIdx.y;
    if (bi < bj) return;  // lower-triangle blocks only
    const int tx = threadIdx.x, ty = threadIdx.y;
    for (int r = ty; r < 32; r += blockDim.y) {
        const int i = bi * 32 + tx, j = bj * 32 + r;
        double v = 0.0;
        if (i < ld && j < ld) {
            v = G[(size_t)ld * j + i];
            if (i == j && i < N) {
                const double add = (ml_first_n >= 0) ? (i < ml_first_n ? add_res : 0.0) : (i < S ? add_model : add_res);
                v += add;
                G[(size_t)ld * j + i] = v;
            }
            // prior(i,i) on states_x_trainingdata_aug(p = i, column i), i < S: element Gaug[N+i, i]
            if (i >= N && i - N == j && j < S && j < P && prior_add != 0.0) {
                v += prior_add;
                G[(size_t)ld * j + i] = v;
            }
        }
        tile[r][tx] = v;  // tile[j_local][i_local]
    }
    __syncthreads();
    if (bi == bj) {
        for (int r = ty; r < 32; r += blockDim.y) {
            const int i = bi * 32 + r, j = bj * 32 + tx;  // upper element (i < j) <- lower (j, i)
            if (i < j && j < ld) G[(size_t)ld * j + i] = tile[r][tx];
        }
    } else {
        for (int r = ty; r < 32; r += blockDim.y) {
            const int i = bj * 32 + tx, j = bi * 32 + r;  // transposed block
            if (i < ld && j < ld) G[(size_t)ld * j + i] = tile[tx][r];
        }
    }
}

// W_out(P, N) <- X^T where X (N x P, ld) is the solution sitting in Gaug[0:N, N:N+P]; wout has leading dim ldw
__global__ void k_train_store_wout(const double *__restrict__ X, int ldx, int N, int P, double *__restrict__ wout, int ldw)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;  // feature index
    if (j >= N) return;
    for (int p = 0; p < P; ++p) wout[(size_t)ldw * j + p] = X[(size_t)ldx * p + j];
}

struct TrainRegionHost {
    int local = -1, region = -1;
    TrainRegionDev dev{};
    std::vector<std::pair<void *, size_t>> allocs;
};

// device buffers of finished waves, reused by the next sml_train_begin (a wave is ~350 MB per region: cudaMalloc /
// cudaFree of thousands of such blocks cost more than the solve)
struct TrainPool {
    std::multimap<size_t, void *> free_blocks;
    void *take(size_t bytes)
    {
        auto it = free_blocks.find(bytes);
        if (it == free_blocks.end()) return nullptr;
        void *p = it->second;
        free_blocks.erase(it);
        return p;
    }
    // the smallest kept block that holds `bytes` (a wave's arena fits any wave that is not larger)
    void *take_at_least(size_t bytes, size_t *got)
    {
        auto it = free_blocks.lower_bound(bytes);
        if (it == free_blocks.end()) return nullptr;
        void *p = it->second;
        *got = it->first;
        free_blocks.erase(it);
        return p;
    }
    void drop_all()
    {
        for (auto &b : free_blocks) cudaFree(b.second);
        free_blocks.clear();
    }
};

struct TrainState {
    bool active = false;
    int kind = 0, batch_size = 0;
    bool hybrid = true;
    std::vector<TrainRegionHost> regs;
    TrainRegionDev *d_regs = nullptr;
    int2 *d_tiles = nullptr;
    int ntiles = 0;
    int ld_max = 0, n_max = 0, D_max = 0, ks = 2048;
    double *d_series_td = nullptr, *d_series_im = nullptr;
    size_t series_td_cap = 0, series_im_cap = 0;
    double gram_flops_useful = 0.0;   // N(N+1)K + 2PNK summed over feeds
    double gram_ms = 0.0;             // CUDA-event time of the Gram kernels
    double stategen_ms = 0.0;
    double solve_ms = 0.0;
    int solved_by_cholesky = 0;
    // overlap mode: the Gram of slab buffer b runs on its own stream while the state generation fills buffer b^1
    bool overlap = false;
    void *arena = nullptr;                    // ONE device allocation per wave, carved up by sml_train_begin
    size_t arena_bytes = 0;
    int stategen_route = 0;                   // SML_TRAIN_STATEGEN: 0 auto (by wave size), 1 'steps' (k_train_update per time step), 2 'kernel' (k_train_stategen)
    unsigned slab_seq = 0;                    // slabs produced so far in this wave; buffer = slab_seq & 1
    cudaEvent_t ev_gram[2] = {nullptr, nullptr};   // end of the last Gram that read buffer b (owned by spans)
    struct Span { cudaEvent_t a, b; int what; };   // what: 0 state generation, 1 Gram; resolved at the next sync point
    std::vector<Span> spans;
    std::vector<TrainRegionDev> uploaded;     // what d_regs holds (re-uploaded only when it changes)
};

// Per-region standardisation constants from the resident global series (get_training_data, src/mod_reservoir.f90:
// 413-470): for every (variable, level) slot the mean and the population standard deviation over the region's halo
// block and all columns of the window --
//   kind 0 (atmosphere levels, logp, tisr; standardize_data_5d_logp_tisr, src/mod_utilities.f90:1144-1193):
//          mean = sum/N, std = sqrt(sum((x-mean)^2)/N)
//   kind 1 (precip; standardize_data_3d :894-912): std = sqrt((sum(x^2) - sum(x)^2/N)/N)
//   kind 2 (SST; standardize_sst_data_3d :853-892): as kind 0 if sum(x^2)-sum(x)^2/N > 0 and std > 0.2, else
//          mean = std = 0 and the region gets no SST input (any_change = .False.)
// One block per (slot, region); fixed-order tree reductions, so the result is deterministic.
struct StatSlot {
    int first_cell;   // index into the region's cell list
    int ncells;
    long long base;   // offset of the slot's 2-D field (or of (var, level)) in a G column; element = base + cell*cstride
    int cstride;
    int kind;
};

__global__ void __launch_bounds__(256)
k_cond_stats(const double *__restrict__ Gs, long long g_len, int first, int stride, int ncols,
             const StatSlot *__restrict__ slots, int L, const int *__restrict__ cells, double *__restrict__ mean_out,
             double *__restrict__ std_out, int *__restrict__ sst_flag)
{
    __shared__ double r1[256], r2[256];
    __shared__ double s_mean;
    const int region = blockIdx.y, l = blockIdx.x;
    const StatSlot sl = slots[(size_t)region * L + l];
    const int *cl = cells + sl.first_cell;
    const long long total = (long long)sl.ncells * ncols;
    double a = 0.0, b = 0.0;
    for (long long e = threadIdx.x; e < total; e += 256) {
        const int c = (int)(e / sl.ncells), k = (int)(e % sl.ncells);
        const double x = Gs[(size_t)(first + stride * c) * g_len + sl.base + (long long)cl[k] * sl.cstride];
        a += x;
        b = fma(x, x, b);
    }
    r1[threadIdx.x] = a;
    r2[threadIdx.x] = b;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) {
            r1[threadIdx.x] += r1[threadIdx.x + o];
            r2[threadIdx.x] += r2[threadIdx.x + o];
        }
        __syncthreads();
    }
    const double S1 = r1[0], Sq = r2[0], N = (double)total;
    if (threadIdx.x == 0) s_mean = S1 / N;
    __syncthreads();
    const double mean = s_mean;
    double v = 0.0;
    for (long long e = threadIdx.x; e < total; e += 256) {
        const int c = (int)(e / sl.ncells), k = (int)(e % sl.ncells);
        const double x = Gs[(size_t)(first + stride * c) * g_len + sl.base + (long long)cl[k] * sl.cstride];
        const double d = x - mean;
        v = fma(d, d, v);
    }
    __syncthreads();
    r1[threadIdx.x] = v;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) r1[threadIdx.x] += r1[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        double m = mean, sd;
        const double onepass = Sq - S1 * S1 / N;
        if (sl.kind == 1) {
            sd = sqrt(onepass / N);
        } else {
            sd = sqrt(r1[0] / N);
            if (sl.kind == 2) {
                const bool keep = onepass > 0.0 && sd > 0.2;
                if (!keep) m = sd = 0.0;
                sst_flag[region] = keep ? 1 : 0;
            }
        }
        mean_out[(size_t)region * L + l] = m;
        std_out[(size_t)region * L + l] = sd;
    }
}

// get_training_data's conditioning of the raw (hourly, physical-unit) series, in place on the resident copy
// (src/mod_reservoir.f90:362-395):
//   specific humidity: *1000 (g/kg), floored at 1e-6 (:365-368);   TISR: negative -> 0 (:373-375)
//   precip: negative -> 0 (:380-382); total over the hybrid time step, total_precip_over_a_period
//           (src/mod_utilities.f90:1688-1729: sum(copy(t-period : t)), i.e. period+1 hourly values, sum(copy(1:t)) while
//           t - period < 1); then log(1 + p/precip_epsilon) (:387)
//   SST: floored at 272 K (:390-394)
// One thread per grid cell walks the time axis; the windowed precip sum runs from the last column backwards so that
// the original values it needs are still in place.
__global__ void k_condition_series(double *__restrict__ Gs, long long g_len, int ncols, int period, double eps,
                                   long long off_w2d, long long off_precip, long long off_sst, long long off_tisr,
                                   int do_precip, int do_sst)
{
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= g_len) return;
    if (e < off_w2d) {
        if ((e & 3) == 3)  // var index is the fastest one: q is variable 4
            for (int t = 0; t < ncols; ++t) {
                double v = Gs[(size_t)t * g_len + e] * 1000.0;
                if (v < 0.000001) v = 0.000001;
                Gs[(size_t)t * g_len + e] = v;
            }
    } else if (e >= off_tisr) {
        for (int t = 0; t < ncols; ++t)
            if (Gs[(size_t)t * g_len + e] < 0.0) Gs[(size_t)t * g_len + e] = 0.0;
    } else if (e >= off_sst) {
        if (do_sst)
            for (int t = 0; t < ncols; ++t)
                if (Gs[(size_t)t * g_len + e] < 272.0) Gs[(size_t)t * g_len + e] = 272.0;
    } else if (e >= off_precip) {
        if (do_precip) {
            for (int t = 0; t < ncols; ++t)
                if (Gs[(size_t)t * g_len + e] < 0.0) Gs[(size_t)t * g_len + e] = 0.0;
            for (int t = ncols - 1; t >= 0; --t) {
                const int t0 = (t - period < 0) ? 0 : t - period;   // 1-based: t-period < 1 -> 1..t, else t-period..t
                double s = 0.0;
                for (int k = t0; k <= t; ++k) s += Gs[(size_t)k * g_len + e];   // Fortran sum: first to last
                Gs[(size_t)t * g_len + e] = log(1.0 + s / eps);
            }
        }
    }
}

// rolling_average_over_a_period_2d (src/mod_utilities.f90:1773-1815), which get_training_data_from_atmo applies to the
// atmosphere rows of every slab-ocean reservoir's training series (src/mod_slab_ocean_reservoir.f90:398, :452):
//   t - period < 1 (1-based):  out(i,t) = sum(copy(i,1:t)) / t
//   else                    :  out(i,t) = sum(copy(i,t-period:t)) / period   -- period+1 values over period, as written --
//                              kept only if |sum| > 1e-7 (keep_small: the 2-D variant; the 3-D variant :1731 has no test)
// Every window is summed first to last like the Fortran intrinsic (no sliding update: that would change the rounding).
// src, dst: [t_len][nrows] column-major copies (row fastest).   grid (t_len, ceil(nrows/128))
__global__ void k_rolling_average(const double *__restrict__ src, double *__restrict__ dst, int nrows, int t_len, int period,
                                  int keep_small)
{
    const int t = blockIdx.x, i = blockIdx.y * blockDim.x + threadIdx.x;
    if (i >= nrows) return;
    const bool head = (t + 1) - period < 1;
    const int lo = head ? 0 : t - period;
    double s = 0.0;
    for (int k = lo; k <= t; ++k) s = __dadd_rn(s, src[(size_t)nrows * k + i]);
    double out;
    if (head) out = s / (double)(t + 1);
    else if (keep_small && !(fabs(s) > 0.0000001)) out = src[(size_t)nrows * t + i];
    else out = s / (double)period;
    dst[(size_t)nrows * t + i] = out;
}

// global training series, kept across waves (sml_train_global_series / sml_train_global_release)
struct TrainGlobal {
    double *d_G = nullptr, *d_F = nullptr;
    int ncols_total = 0;
    bool conditioned = false;
    double noisemag = 0.0, precip_eps = 0.001;   // sml_train_set_noise
    unsigned long long seed = 0;
};

inline void train_release(TrainState &t, TrainPool *pool = nullptr)
{
    for (auto &r : t.regs)
        for (auto &a : r.allocs) {
            if (pool) pool->free_blocks.emplace(a.second, a.first);
            else cudaFree(a.first);
        }
    t.regs.clear();
    if (t.arena) {
        if (pool) pool->free_blocks.emplace(t.arena_bytes, t.arena);
        else cudaFree(t.arena);
        t.arena = nullptr;
        t.arena_bytes = 0;
    }
    for (auto &sp : t.spans) { cudaEventDestroy(sp.a); cudaEventDestroy(sp.b); }
    t.spans.clear();
    t.ev_gram[0] = t.ev_gram[1] = nullptr;
    t.uploaded.clear();
    t.slab_seq = 0;
    cudaFree(t.d_regs); t.d_regs = nullptr;
    cudaFree(t.d_tiles); t.d_tiles = nullptr;
    cudaFree(t.d_series_td); t.d_series_td = nullptr; t.series_td_cap = 0;
    cudaFree(t.d_series_im); t.d_series_im = nullptr; t.series_im_cap = 0;
    t.active = false;
}

}  // namespace sml
