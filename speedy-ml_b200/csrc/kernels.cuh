// kernels.cuh -- sm_100a kernels of the forecast hot path.
//
//  k_step            fused state update (ELL SpMV + compact W_in + tanh + leak + even-square) and
//                    W_out readout partials; W_out tiles arrive by TMA bulk copies (cp.async.bulk,
//                    mbarrier complete_tx) through a multi-stage shared-memory ring.
//  k_win_dense       fallback W_in*u for regions whose W_in is not one-non-zero-per-row.
//  k_readout_finish  fixed-order reduction of the partials + un-standardise -> outvec slab, pushed into every
//                    rank's gathered buffer over NVLink (fused all-gather) when peers are attached.
//  k_pack_grids      outvec slabs of all regions -> global grids with the exchange clamps, and the wholegrid_sst
//                    assembly (ocean tiles / 272 K / land mask / floor); waits for the peers' step flags.
//  k_build_inputs    global grids -> every region's feedback and local_model (gather + standardise).
//
// Reference statements each kernel reproduces are cited at the kernel.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sml {

struct RegionDev {
    int n, D, P, S;
    int ldw;        // leading dimension of the device W_out copy: P rounded up to even
    int ell_w;      // ELL width (max entries per row; every row padded with (col 0, val 0.0))
    int win_mode;   // 0: compact one-per-row W_in, 1: dense fallback (temp pool)
    int item0, nitems;
    int L;          // mean/std length; slot L holds the SST feedback mean/std
    double leak;
    const int *ell_col;      // [ell_w][n] slot-major, 0-based
    const double *ell_val;   // [ell_w][n]
    const double *winc;      // [n]
    const int *wcol;         // [n]
    const double *win_dense; // [n*D] or null
    const double *wout;      // [ldw*(S+n)]
    const double *mean;      // [L+1]
    const double *std;       // [L+1]
    long long x_off, fb_off, lm_off, out_off;
    const int *fb_src, *fb_ms;   // [D]
    const int *lm_src, *lm_ms;   // [S]
    const int *out_ms;           // [P]
};

struct StepItem {
    int reg;     // local region index
    int row0;    // first state row of the chunk
    int nrows;
    int col0;    // first W_out column (0 for the chunk that also carries the S model columns)
    int ncols;
    int xs_off;  // where the chunk's x~ starts in the feature tile (S for the first chunk, else 0)
};

constexpr int NCONS = 544;            // consumer threads (17 warps); 2*NCONS = 1088 = 8 * 136
constexpr int NTHREADS = NCONS + 32;  // + one TMA producer warp
constexpr int NCONS_WARPS = NCONS / 32;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) {
    }
}
// TMA 1-D bulk copy global -> shared, completion counted in bytes on the mbarrier
__device__ __forceinline__ void tma_load_1d(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void consumer_bar()
{
    asm volatile("bar.sync 1, %0;" ::"n"(NCONS) : "memory");
}

// ---------------------------------------------------------------------------------------------
// peer exchange over NVLink (one process per GPU, buffers mapped with CUDA IPC): every rank owns an exchange
// block  [2][R*P] doubles (the all-gathered outvecs, double-buffered by step parity) + MAX_PEERS step flags.
// The readout-finish kernel stores each region's outvec straight into EVERY rank's block and, once all local
// regions are out, publishes the step number in every rank's flag slot; the pack kernel of each rank waits for
// all slots to reach the step.  This replaces the per-step NCCL all-gather (1152 x 136 doubles in total).
// ---------------------------------------------------------------------------------------------
constexpr int MAX_PEERS = 8;
struct PeerTable {
    int world, rank;
    double *gathered[MAX_PEERS];             // rank k's [2][R*P] buffer as mapped into this process
    unsigned long long *flags[MAX_PEERS];    // rank k's flags[MAX_PEERS]; slot r is written by rank r
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ double ld_cg_f64(const double *p)
{
    double v;
    asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}

// ---------------------------------------------------------------------------------------------
// state update of one row:  y = A x (ELL, entries in COO order so duplicates sum in entry order),
// temp = W_in u (one product), x <- (1-leak) x + leak tanh(y + temp)
// src/mod_reservoir.f90:1444-1448 (predict), :1373-1377 (synchronize)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double update_row(const RegionDev &R, int row, const double *__restrict__ xo,
                                             const double *__restrict__ u, const double *__restrict__ temp_pool)
{
    const int n = R.n;
    const int *__restrict__ ec = R.ell_col + row;
    const double *__restrict__ ev = R.ell_val + row;
    double acc = 0.0;
    int s = 0;
    const int W = R.ell_w;
    for (; s + 6 <= W; s += 6) {
        int c[6];
        double v[6], xv[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            c[i] = __ldg(ec + (size_t)(s + i) * n);
            v[i] = __ldg(ev + (size_t)(s + i) * n);
        }
#pragma unroll
        for (int i = 0; i < 6; ++i) xv[i] = xo[c[i]];
#pragma unroll
        for (int i = 0; i < 6; ++i) acc = fma(v[i], xv[i], acc);
    }
    for (; s < W; ++s) {
        const int c = __ldg(ec + (size_t)s * n);
        const double v = __ldg(ev + (size_t)s * n);
        acc = fma(v, xo[c], acc);
    }
    double t;
    if (R.win_mode == 0) t = __dmul_rn(__ldg(R.winc + row), u[__ldg(R.wcol + row)]);
    else t = temp_pool[R.x_off + row];
    const double xt = tanh(__dadd_rn(acc, t));
    return __dadd_rn(__dmul_rn(1.0 - R.leak, xo[row]), __dmul_rn(R.leak, xt));
}

// ---------------------------------------------------------------------------------------------
// k_step: one CTA per (region, row chunk).
//   warps 0..16 : state update of the chunk's rows -> x_new (global) and x~ (shared), then consume
//                 W_out column tiles from the shared-memory ring: thread (rp, cs) owns rows 2rp,2rp+1
//                 and columns cs, cs+cpi, ... of every tile.
//   warp 17     : lane 0 streams W_out[:, col0 : col0+ncols] with TMA bulk copies, STAGES deep.
// do_readout == 0 gives the update-only step used by synchronize.
// Readout statement: outvec = matmul(wout, [local_model ; x~])   src/mod_reservoir.f90:1450-1456
// ---------------------------------------------------------------------------------------------
template <int STAGES>
__global__ void __launch_bounds__(NTHREADS, 2)
k_step(const RegionDev *__restrict__ regs, const StepItem *__restrict__ items, const int *__restrict__ order,
       int item_base, const double *__restrict__ x_old, double *__restrict__ x_new, const double *__restrict__ u_pool, const long long *__restrict__ u_offs, int u_t,
       const double *__restrict__ lm_pool, const double *__restrict__ temp_pool, double *__restrict__ partials,
       int ldw_max, int stage_cols, int stage_bytes, int xs_cap, int do_readout)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *xs = reinterpret_cast<double *>(smem_raw + (size_t)STAGES * stage_bytes);
    double *red = xs + xs_cap;
    uint64_t *full = reinterpret_cast<uint64_t *>(red + 2 * NCONS);
    uint64_t *empty = full + STAGES;

    const int item = order ? order[blockIdx.x] : (int)blockIdx.x;  // optional launch-order permutation
    const StepItem it = items[item];
    const RegionDev R = regs[it.reg];
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int ldw = R.ldw;
    const int nst = do_readout ? (it.ncols + stage_cols - 1) / stage_cols : 0;

    if (do_readout) {
        if (tid == 0) {
            for (int s = 0; s < STAGES; ++s) {
                mbar_init(&full[s], 1);
                mbar_init(&empty[s], NCONS_WARPS);
            }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("fence.proxy.async;" ::: "memory");
        }
        __syncthreads();
    }

    if (warp == NCONS_WARPS) {
        // ---------------- TMA producer ----------------
        if (lane == 0) {
            const double *src = R.wout + (size_t)it.col0 * ldw;
            for (int k = 0; k < nst; ++k) {
                const int s = k % STAGES;
                const uint32_t par = ((k / STAGES) & 1) ^ 1;
                mbar_wait(&empty[s], par);  // passes at once on the first lap
                const int nc = min(stage_cols, it.ncols - k * stage_cols);
                const uint32_t bytes = (uint32_t)nc * ldw * 8u;
                mbar_expect_tx(&full[s], bytes);
                tma_load_1d(smem_raw + (size_t)s * stage_bytes, src + (size_t)k * stage_cols * ldw, bytes, &full[s]);
            }
        }
        return;
    }

    // ---------------- consumers: state update of this chunk ----------------
    const double *xo = x_old + R.x_off;
    double *xn = x_new + R.x_off;
    const double *u = u_pool + u_offs[it.reg] + (long long)u_t * R.D;
    if (do_readout && it.xs_off > 0) {
        const double *lm = lm_pool + R.lm_off;
        for (int i = tid; i < it.xs_off; i += NCONS) xs[i] = lm[i];
    }
    for (int j = tid; j < it.nrows; j += NCONS) {
        const int row = it.row0 + j;
        const double xv = update_row(R, row, xo, u, temp_pool);
        xn[row] = xv;
        if (do_readout) xs[it.xs_off + j] = (row & 1) ? __dmul_rn(xv, xv) : xv;  // even 1-based index squared
    }
    if (!do_readout) return;
    consumer_bar();

    // ---------------- consumers: W_out tiles ----------------
    const int HP = ldw >> 1;            // row pairs
    const int cpi = NCONS / HP;         // columns processed per sweep of the consumer threads
    const bool active = tid < cpi * HP;
    const int rp = tid % HP, cs = tid / HP;
    double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
    for (int k = 0; k < nst; ++k) {
        const int s = k % STAGES;
        mbar_wait(&full[s], (k / STAGES) & 1);
        const int nc = min(stage_cols, it.ncols - k * stage_cols);
        const double *sb = reinterpret_cast<const double *>(smem_raw + (size_t)s * stage_bytes) + 2 * rp;
        const double *xk = xs + k * stage_cols;
        if (active) {
            int c = cs;
            for (; c + cpi < nc; c += 2 * cpi) {
                const double2 w0 = *reinterpret_cast<const double2 *>(sb + (size_t)c * ldw);
                const double2 w1 = *reinterpret_cast<const double2 *>(sb + (size_t)(c + cpi) * ldw);
                const double x0 = xk[c], x1 = xk[c + cpi];
                a0 = fma(w0.x, x0, a0);
                a1 = fma(w0.y, x0, a1);
                b0 = fma(w1.x, x1, b0);
                b1 = fma(w1.y, x1, b1);
            }
            if (c < nc) {
                const double2 w0 = *reinterpret_cast<const double2 *>(sb + (size_t)c * ldw);
                const double x0 = xk[c];
                a0 = fma(w0.x, x0, a0);
                a1 = fma(w0.y, x0, a1);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
    }
    if (active) {
        red[cs * ldw + 2 * rp] = a0 + b0;
        red[cs * ldw + 2 * rp + 1] = a1 + b1;
    }
    consumer_bar();
    for (int p = tid; p < ldw; p += NCONS) {
        double sum = 0.0;
        for (int g = 0; g < cpi; ++g) sum += red[g * ldw + p];
        partials[(size_t)(item_base + item) * ldw_max + p] = sum;
    }
}

// ---------------------------------------------------------------------------------------------
// k_update: the state update alone (synchronize, src/mod_reservoir.f90:1354-1381; slab :1237-1266).
// The fused kernel hides this phase behind the W_out stream; on its own it is a latency problem (four dependent
// global loads per row), so this variant keeps more loads in flight: 256-thread CTAs at high occupancy, RPT rows
// per thread (rows 256 apart, so every load stays coalesced), ELL slots fetched three at a time for all of the
// thread's rows before the x gathers.  Accumulation order per row is the entry order, as in update_row.
// grid (ceil(n_max / (256*RPT)), regions); region_list selects regions (one region's synchronize) or is null.
// ---------------------------------------------------------------------------------------------
template <int RPT>
__global__ void __launch_bounds__(256)
k_update(const RegionDev *__restrict__ regs, const int *__restrict__ region_list, const double *__restrict__ x_old,
         double *__restrict__ x_new, const double *__restrict__ u_pool, const long long *__restrict__ u_offs, int u_t,
         const double *__restrict__ temp_pool)
{
    const int reg = region_list ? region_list[blockIdx.y] : (int)blockIdx.y;
    const RegionDev &R = regs[reg];
    const int n = R.n;
    const int base = blockIdx.x * (256 * RPT) + threadIdx.x;
    if (base >= n) return;
    const double *__restrict__ xo = x_old + R.x_off;
    const double *__restrict__ u = u_pool + u_offs[reg] + (long long)u_t * R.D;
    const int *__restrict__ ecol = R.ell_col;
    const double *__restrict__ eval = R.ell_val;
    const int W = R.ell_w;
    int row[RPT];
    bool ok[RPT];
    double acc[RPT];
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
        const int r = base + i * 256;
        ok[i] = r < n;
        row[i] = ok[i] ? r : base;  // out-of-range lanes recompute a valid row and drop the result
        acc[i] = 0.0;
    }
    for (int s = 0; s < W; s += 3) {
        int c[RPT][3];
        double v[RPT][3], xv[RPT][3];
#pragma unroll
        for (int j = 0; j < 3; ++j)
#pragma unroll
            for (int i = 0; i < RPT; ++i) {
                const bool in = s + j < W;
                c[i][j] = in ? __ldg(ecol + (size_t)(s + j) * n + row[i]) : 0;
                v[i][j] = in ? __ldg(eval + (size_t)(s + j) * n + row[i]) : 0.0;
            }
#pragma unroll
        for (int j = 0; j < 3; ++j)
#pragma unroll
            for (int i = 0; i < RPT; ++i) xv[i][j] = xo[c[i][j]];
#pragma unroll
        for (int j = 0; j < 3; ++j)
#pragma unroll
            for (int i = 0; i < RPT; ++i)
                if (s + j < W) acc[i] = fma(v[i][j], xv[i][j], acc[i]);
    }
    double t[RPT], xr[RPT];
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
        if (R.win_mode == 0) t[i] = __dmul_rn(__ldg(R.winc + row[i]), u[__ldg(R.wcol + row[i])]);
        else t[i] = temp_pool[R.x_off + row[i]];
        xr[i] = xo[row[i]];
    }
    double *__restrict__ xn = x_new + R.x_off;
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
        const double xt = tanh(__dadd_rn(acc[i], t[i]));
        const double xv2 = __dadd_rn(__dmul_rn(1.0 - R.leak, xr[i]), __dmul_rn(R.leak, xt));
        if (ok[i]) xn[row[i]] = xv2;
    }
}

// ---------------------------------------------------------------------------------------------
// k_update_sx: the same update with the region's state vector (and input vector) staged in shared memory, so that
// the x gathers of the SpMV -- random within the region's 46 KB vector -- are shared-memory reads instead of L2
// round trips and only the coalesced ELL / W_in streams go to global memory.  One CTA per (region, row split), 512
// threads, 2 CTAs per SM (64 registers), n*8 + D*8 bytes of dynamic shared memory.  x arrives by one TMA bulk copy (the x pool is padded to 256 B per region) while every thread
// already has its first ELL group in flight; RPT rows per thread per sweep.
// Accumulation order per row is the entry order, as in update_row.   grid (nsplit, regions)
// Measured (tools/ab_update.py, 1152 regions, m = 6000): 0.796 of the HBM roof at degree 6, 0.917 at degree 24,
// against 0.723 / 0.778 for k_update<4>; 144 regions (one of 8 ranks, 2 row splits): 0.60 against 0.47.
// ---------------------------------------------------------------------------------------------
constexpr int UPD_SX_THREADS = 512;
template <int RPT>
__global__ void __launch_bounds__(UPD_SX_THREADS, 2)
k_update_sx(const RegionDev *__restrict__ regs, const int *__restrict__ region_list, const double *__restrict__ x_old,
            double *__restrict__ x_new, const double *__restrict__ u_pool, const long long *__restrict__ u_offs, int u_t,
            const double *__restrict__ temp_pool, int nsplit, int xs_cap)
{
    extern __shared__ __align__(128) double sx_smem[];
    const int reg = region_list ? region_list[blockIdx.y] : (int)blockIdx.y;
    const RegionDev &R = regs[reg];
    const int n = R.n, D = R.D, W = R.ell_w;
    const int per = (((n + nsplit - 1) / nsplit) + 31) & ~31;
    const int r0 = blockIdx.x * per, r1 = min(n, r0 + per);
    if (r0 >= n) return;
    const int tid = threadIdx.x, nt = blockDim.x;
    const double *__restrict__ xo = x_old + R.x_off;
    const double *__restrict__ u = u_pool + u_offs[reg] + (long long)u_t * D;
    double *xs = sx_smem, *us = sx_smem + xs_cap;
    uint64_t *bar = reinterpret_cast<uint64_t *>(us + ((D + 1) & ~1));
    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async;" ::: "memory");
        const uint32_t bytes = (uint32_t)((n + 1) & ~1) * 8u;   // inside the region's padded slot of the x pool
        mbar_expect_tx(bar, bytes);
        tma_load_1d(xs, xo, bytes, bar);
    }
    const bool compact = R.win_mode == 0;
    if (compact)
        for (int i = tid; i < D; i += nt) us[i] = u[i];
    __syncthreads();   // the barrier is initialised and u is staged; x may still be in flight
    const int *__restrict__ ecol = R.ell_col;
    const double *__restrict__ eval = R.ell_val;
    double *__restrict__ xn = x_new + R.x_off;
    const double leak = R.leak;
    bool x_ready = false;
    for (int base = r0 + tid; base < r1; base += nt * RPT) {
        int row[RPT];
        bool ok[RPT];
        double acc[RPT];
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
            const int r = base + i * nt;
            ok[i] = r < r1;
            row[i] = ok[i] ? r : base;
            acc[i] = 0.0;
        }
        double wv[RPT];
        int wc[RPT];
        if (compact) {
#pragma unroll
            for (int i = 0; i < RPT; ++i) {
                wv[i] = __ldg(R.winc + row[i]);
                wc[i] = __ldg(R.wcol + row[i]);
            }
        }
        for (int s = 0; s < W; s += 3) {
            int c[RPT][3];
            double v[RPT][3];
#pragma unroll
            for (int j = 0; j < 3; ++j)
#pragma unroll
                for (int i = 0; i < RPT; ++i) {
                    const bool in = s + j < W;
                    c[i][j] = in ? __ldg(ecol + (size_t)(s + j) * n + row[i]) : 0;
                    v[i][j] = in ? __ldg(eval + (size_t)(s + j) * n + row[i]) : 0.0;
                }
            if (!x_ready) {   // first group of the first sweep: its loads are in flight while x lands
                mbar_wait(bar, 0);
                x_ready = true;
            }
#pragma unroll
            for (int j = 0; j < 3; ++j)
#pragma unroll
                for (int i = 0; i < RPT; ++i)
                    if (s + j < W) acc[i] = fma(v[i][j], xs[c[i][j]], acc[i]);
        }
        if (!x_ready) {   // W == 0
            mbar_wait(bar, 0);
            x_ready = true;
        }
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
            const double t = compact ? __dmul_rn(wv[i], us[wc[i]]) : temp_pool[R.x_off + row[i]];
            const double xt = tanh(__dadd_rn(acc[i], t));
            const double xv2 = __dadd_rn(__dmul_rn(1.0 - leak, xs[row[i]]), __dmul_rn(leak, xt));
            if (ok[i]) xn[row[i]] = xv2;
        }
    }
}

// dense W_in fallback: temp = matmul(win, u) for regions with win_mode == 1 (src/mod_reservoir.f90:1445).
// grid: (ceil(n_max/256), nregions)
__global__ void k_win_dense(const RegionDev *__restrict__ regs, const double *__restrict__ u_pool,
                            const long long *__restrict__ u_offs, int u_t, double *__restrict__ temp_pool)
{
    const RegionDev R = regs[blockIdx.y];
    if (R.win_mode != 1) return;
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= R.n) return;
    const double *u = u_pool + u_offs[blockIdx.y] + (long long)u_t * R.D;
    double acc = 0.0;
    for (int i = 0; i < R.D; ++i) acc = fma(R.win_dense[(size_t)i * R.n + row], u[i], acc);
    temp_pool[R.x_off + row] = acc;
}

// partials of a region summed in item order, then unstandardize_state_vec_res (src/res_domain.f90:1424-1475):
// v*std then +mean, two roundings (src/mod_utilities.f90:799-829).  One block per region.
// model_part != 0 (split-order readout of the overlapped step): the partials hold only W_out[:, S:]*x~ (the
// reference's v_ml, src/mod_reservoir.f90:1460) and this kernel adds v_p = W_out[:, 0:S]*local_model (:1459)
// once the host model's forecast has arrived.
constexpr int FIN_GROUPS = 4;     // column groups of the model-part GEMV
constexpr int FIN_PMAX = 160;     // threads per group (outputs are strided over them: any chunk_size_prediction)

__global__ void __launch_bounds__(FIN_GROUPS *FIN_PMAX)
k_readout_finish(const RegionDev *__restrict__ regs, const double *__restrict__ partials, int ldw_max,
                 double *__restrict__ out_pool, int unstandardize, int model_part,
                 const double *__restrict__ lm_pool, PeerTable pt, unsigned long long seq,
                 long long peer_off, unsigned int *__restrict__ done_counter, double *__restrict__ vp_pool,
                 double *__restrict__ vml_pool)
{
    extern __shared__ double s_vp[];   // [FIN_GROUPS][pstride], overlapped mode only
    const RegionDev R = regs[blockIdx.x];
    const int pstride = (R.P + 1) & ~1;
    const int grp = threadIdx.x / FIN_PMAX, p0 = threadIdx.x % FIN_PMAX;
    if (model_part) {
        // v_p = W_out[:, 0:S] * local_model: group g takes columns g, g+4, ...; 4 independent FMA chains per thread
        // keep the loads in flight; the groups are combined in fixed order below (deterministic)
        const double *lm = lm_pool + R.lm_off;
        for (int p = p0; p < R.P; p += FIN_PMAX) {
            const double *w = R.wout + p;
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
            int j = grp;
            for (; j + 3 * FIN_GROUPS < R.S; j += 4 * FIN_GROUPS) {
                a0 = fma(w[(size_t)j * R.ldw], lm[j], a0);
                a1 = fma(w[(size_t)(j + FIN_GROUPS) * R.ldw], lm[j + FIN_GROUPS], a1);
                a2 = fma(w[(size_t)(j + 2 * FIN_GROUPS) * R.ldw], lm[j + 2 * FIN_GROUPS], a2);
                a3 = fma(w[(size_t)(j + 3 * FIN_GROUPS) * R.ldw], lm[j + 3 * FIN_GROUPS], a3);
            }
            for (; j < R.S; j += FIN_GROUPS) a0 = fma(w[(size_t)j * R.ldw], lm[j], a0);
            s_vp[grp * pstride + p] = (a0 + a1) + (a2 + a3);
        }
        __syncthreads();
    }
    if (grp == 0) {
        for (int p = p0; p < R.P; p += FIN_PMAX) {
            double v = 0.0;
            if (model_part) v = (s_vp[p] + s_vp[pstride + p]) + (s_vp[2 * pstride + p] + s_vp[3 * pstride + p]);
            if (vp_pool) {
                // reservoir%v_p / reservoir%v_ml (outvec_component_contribs, src/mod_reservoir.f90:1458-1461): the two
                // halves of the readout, in standardised units as the reference keeps them
                double vml = 0.0;
                for (int c = 0; c < R.nitems; ++c) vml += partials[(size_t)(R.item0 + c) * ldw_max + p];
                vp_pool[R.out_off + p] = v;
                vml_pool[R.out_off + p] = vml;
            }
            for (int c = 0; c < R.nitems; ++c) v += partials[(size_t)(R.item0 + c) * ldw_max + p];
            if (unstandardize) {
                const int ms = R.out_ms[p];
                if (ms >= 0) v = __dadd_rn(__dmul_rn(v, R.std[ms]), R.mean[ms]);
            }
            out_pool[R.out_off + p] = v;
            // fused all-gather: the outvec goes straight into every rank's gathered buffer (peer stores over NVLink)
            if (pt.world > 1) {
                const long long dst = peer_off + R.out_off + p;
                for (int k = 0; k < pt.world; ++k) pt.gathered[k][dst] = v;
            }
        }
    }
    if (pt.world > 1) {
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned int old = atomicAdd(done_counter, 1u);
            if (old == gridDim.x - 1) {  // every local region's outvec is on its way: publish the step
                *done_counter = 0;
                __threadfence_system();
                for (int k = 0; k < pt.world; ++k) st_release_sys(pt.flags[k] + pt.rank, seq);
            }
        }
    }
}

// k_pack_grids: the root's grid assembly of sendrecievegrid in one launch.
//  blocks [0, nsc): tile_full_grid_with_local_state_vec_res1d for every region of the model
//     (src/res_domain.f90:791-826) + the clamps: q < 1e-6 -> 1e-6 (src/mpires.f90:460-462), precip < 1e-5 -> 0
//     (:486-490).  With peers, first wait until every rank has published this step's outvecs.
//  blocks [nsc, ..): wholegrid_sst (src/mpires.f90:288-290, 315-328, 470-484).  mode 0: every cell takes the first
//     fx*fy outputs of its region's ocean reservoir (tile_full_2d_grid_with_local_res, src/res_domain.f90:828-850);
//     regions without one hold 272.0 in their slab row (:323-326, 383); mode 1: prescribed field; mode -1: no SST.
struct PackArgs {
    const double *gathered;
    const int *out_dst;
    int total;
    double *G;
    long long precip_lo, precip_hi, w4d_hi, sst_off;
    int nsc;
    // peers
    const unsigned long long *my_flags;
    int world;
    unsigned long long seq;
    int *err;
    // sst
    const double *base, *mask, *prescribed, *ocean_out;
    const int *cell_region, *cell_slot;
    int ocean_P, sst_mode;
};

__global__ void k_pack_grids(PackArgs a)
{
    if ((int)blockIdx.x < a.nsc) {
        if (a.world > 1) {
            if ((int)threadIdx.x < a.world) {
                const long long t0 = clock64();
                while (ld_acquire_sys(a.my_flags + threadIdx.x) < a.seq) {
                    if (clock64() - t0 > 20000000000LL) {  // ~10 s: a peer died; report instead of hanging
                        *a.err = 1;
                        break;
                    }
                }
            }
            __syncthreads();
        }
        const int i = blockIdx.x * blockDim.x + threadIdx.x;
        if (i >= a.total) return;
        const int dst = a.out_dst[i];
        double v = ld_cg_f64(a.gathered + i);
        if (dst < a.w4d_hi) {
            if ((dst & 3) == 3 && v < 0.000001) v = 0.000001;
        } else if (dst >= a.precip_lo && dst < a.precip_hi) {
            if (v < 0.00001) v = 0.0;
        }
        a.G[dst] = v;
        return;
    }
    const int e = (blockIdx.x - a.nsc) * blockDim.x + threadIdx.x;
    if (e >= 96 * 48 || a.sst_mode < 0) return;
    double v;
    if (a.sst_mode == 1) v = a.prescribed[e];
    else v = a.ocean_out[(size_t)a.cell_region[e] * a.ocean_P + a.cell_slot[e]];
    if (a.mask[e] > 0.0) v = a.base[e];
    if (v < 272.0) v = 272.0;
    a.G[a.sst_off + e] = v;
}

// feedback / local_model from the global buffers (src/mpires.f90:581-604, 749-775):
// gather by precomputed maps, then (x - mean) / std with the slot's constants (two roundings,
// src/mod_utilities.f90:1307-1329).  grid: (nregions), threads stride over D then S.
__global__ void k_build_inputs(const RegionDev *__restrict__ regs, const double *__restrict__ G,
                               const double *__restrict__ F, double *__restrict__ fb_pool,
                               double *__restrict__ lm_pool, int do_model, int do_feedback)
{
    const RegionDev R = regs[blockIdx.x];
    if (do_feedback)
        for (int d = threadIdx.x; d < R.D; d += blockDim.x) {
            double v = G[R.fb_src[d]];
            const int ms = R.fb_ms[d];
            if (ms >= 0) v = __ddiv_rn(__dsub_rn(v, R.mean[ms]), R.std[ms]);
            fb_pool[R.fb_off + d] = v;
        }
    if (do_model)
        for (int s = threadIdx.x; s < R.S; s += blockDim.x) {
            double v = F[R.lm_src[s]];
            const int ms = R.lm_ms[s];
            if (ms >= 0) v = __ddiv_rn(__dsub_rn(v, R.mean[ms]), R.std[ms]);
            lm_pool[R.lm_off + s] = v;
        }
}

// ocean reservoir feedback (src/mpires.f90:594-600, 776-781; intended semantics):
//   ring(:, slot) = the atmosphere reservoir's standardised lowest-level + logp feedback
//   feedback(1:A) = sum(ring, dim=2) / nslots   (slot order, divides by nslots even while slots are zero)
//   feedback(sst) = (halo'd tile of wholegrid_sst - mean_sst) / std_sst ;  TISR / OHTC slots untouched.
struct OceanFb {
    long long atmo_fb_off;  // atmosphere feedback of the same region + atmo_slice0
    long long fb_off;       // ocean feedback
    long long ring_off;     // ring [nslots][A]
    int A, ixy;
    const int *sst_src;     // [ixy] offsets into G
    double sst_mean, sst_std;
};

__global__ void k_build_ocean_inputs(const OceanFb *__restrict__ O, const double *__restrict__ G,
                                     const double *__restrict__ atmo_fb, double *__restrict__ ocean_fb,
                                     double *__restrict__ ring, int slot, int nslots)
{
    const OceanFb o = O[blockIdx.x];
    double *rg = ring + o.ring_off;
    for (int e = threadIdx.x; e < o.A; e += blockDim.x) {
        rg[(size_t)slot * o.A + e] = atmo_fb[o.atmo_fb_off + e];
        double s = 0.0;
        for (int k = 0; k < nslots; ++k) s = __dadd_rn(s, rg[(size_t)k * o.A + e]);
        ocean_fb[o.fb_off + e] = __ddiv_rn(s, (double)nslots);
    }
    for (int e = threadIdx.x; e < o.ixy; e += blockDim.x)
        ocean_fb[o.fb_off + o.A + e] = __ddiv_rn(__dsub_rn(G[o.sst_src[e]], o.sst_mean), o.sst_std);
}

}  // namespace sml
