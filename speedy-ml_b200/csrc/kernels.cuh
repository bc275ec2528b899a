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
    int part0, nparts;  // fixed row blocks of the persistent step kernel: the region's partial outvecs
    int L;          // mean/std length; slot L holds the SST feedback mean/std
    int lm_self;    // hybrid slab-ocean reservoir (predict_slab, src/mod_slab_ocean_reservoir.f90:1303): after every readout
                    // local_model <- the STANDARDISED outvec (the reservoir's own prediction is its next "imperfect model")
    int ell_stream; // 1: the ELL / W_in streams of the fused step are loaded evict-first (ld.global.cs) so that they do
                    //    not push the region's state vector -- the target of the random gathers -- out of L1
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

// Programmatic dependent launch (PDL): the four kernels of a step are launched with programmatic stream serialization, so
// a kernel's CTAs may be scheduled while its predecessor is still draining.  pdl_launch() lets the successor be staged as
// early as possible; pdl_wait() -- the first thing every such kernel does before touching memory its predecessor writes or
// reads -- returns once the predecessor has completed and its writes are visible.  Launched the ordinary way both are no-ops.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

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
// peer exchange over NVLink (one process per GPU, blocks mapped with CUDA IPC).  Every rank owns ONE exchange block:
//     atmo   [2][R*P]        the all-gathered atmosphere outvecs, double-buffered by step parity
//     ocean  [2][R*P_ocean]  the all-gathered ocean outvecs, double-buffered by ocean-step parity
//     fcst   [F_TOTAL + 96*48 + 8]  landing buffer of the root's host-model forecast, TISR field and run_speedy flag
//     flags  [XF_KINDS][MAX_PEERS] u64: slot [kind][r] is written by rank r (kind 2: only the root, slot 0)
// The readout-finish kernel stores each region's outvec straight into EVERY rank's block and, once all local regions
// are out, publishes the step number in every rank's flag slot; the pack kernel of each rank waits for all slots to
// reach the step (src/mpires.f90:346-454 gather without the root).  The forecast goes the other way: the root's
// push kernel writes it into every rank's landing buffer (src/mpires.f90:606-739 scatter, :744 bcast) and the
// consumers' wait kernel spins on the forecast flag.  No host collective on the data path.
// ---------------------------------------------------------------------------------------------
constexpr int MAX_PEERS = 8;
constexpr int XF_ATMO = 0, XF_OCEAN = 1, XF_FCST = 2, XF_KINDS = 3;
struct PeerTable {
    int world, rank;
    char *base[MAX_PEERS];                   // rank k's exchange block as mapped into this process
    long long off_atmo, off_ocean, off_fcst, off_flags;   // byte offsets of the sections (same on every rank)
    __host__ __device__ double *atmo(int k) const { return reinterpret_cast<double *>(base[k] + off_atmo); }
    __host__ __device__ double *ocean(int k) const { return reinterpret_cast<double *>(base[k] + off_ocean); }
    __host__ __device__ double *fcst(int k) const { return reinterpret_cast<double *>(base[k] + off_fcst); }
    __host__ __device__ unsigned long long *flag(int k, int kind, int slot) const
    {
        return reinterpret_cast<unsigned long long *>(base[k] + off_flags) + kind * MAX_PEERS + slot;
    }
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
// L2 residency control (round 2).  At small shards (N = 8: 144 regions per GPU) the adjacency, W_in and state vectors of
// the whole shard are 90 MB -- they would live in the 126 MB L2 from one step to the next if the 920 MB W_out stream
// did not flush them every step.  So the W_out bulk copies carry an evict-first policy and the ELL / W_in loads an
// evict-last one (createpolicy + .L2::cache_hint); RegionDev::ell_stream == 2 selects it (sml_finalize decides by size).
__device__ __forceinline__ uint64_t l2_policy_evict_first()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void tma_load_1d_hint(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar, uint64_t policy)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
                 : "memory");
}
// MODE 0: ld.global.nc, 1: evict-first streaming load (ld.global.cs), 2: L2 evict-last through a cache-hint policy
template <int MODE>
__device__ __forceinline__ int ld_stream(const int *p, uint64_t pol)
{
    if (MODE == 2) {
        int v;
        asm volatile("ld.global.nc.L2::cache_hint.b32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
        return v;
    }
    return MODE == 1 ? __ldcs(p) : __ldg(p);
}
template <int MODE>
__device__ __forceinline__ double ld_stream(const double *p, uint64_t pol)
{
    if (MODE == 2) {
        double v;
        asm volatile("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
        return v;
    }
    return MODE == 1 ? __ldcs(p) : __ldg(p);
}

// state-vector gathers / stores with the evict-last policy (MODE 2): x is re-read by the next step's gathers
template <int MODE>
__device__ __forceinline__ double ld_state(const double *p, uint64_t pol)
{
    if (MODE == 2) {
        double v;
        asm volatile("ld.global.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol) : "memory");
        return v;
    }
    return *p;
}
__device__ __forceinline__ void st_keep_f64(double *p, double v, uint64_t pol)
{
    asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(p), "d"(v), "l"(pol) : "memory");
}

template <int MODE>
__device__ __forceinline__ double update_row_t(const RegionDev &R, int row, const double *__restrict__ xo,
                                               const double *__restrict__ u, const double *__restrict__ temp_pool, uint64_t pol)
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
            c[i] = ld_stream<MODE>(ec + (size_t)(s + i) * n, pol);
            v[i] = ld_stream<MODE>(ev + (size_t)(s + i) * n, pol);
        }
#pragma unroll
        for (int i = 0; i < 6; ++i) xv[i] = ld_state<MODE>(xo + c[i], pol);
#pragma unroll
        for (int i = 0; i < 6; ++i) acc = fma(v[i], xv[i], acc);
    }
    for (; s < W; ++s) {
        const int c = ld_stream<MODE>(ec + (size_t)s * n, pol);
        const double v = ld_stream<MODE>(ev + (size_t)s * n, pol);
        acc = fma(v, ld_state<MODE>(xo + c, pol), acc);
    }
    double t;
    if (R.win_mode == 0) t = __dmul_rn(ld_stream<MODE>(R.winc + row, pol), u[ld_stream<MODE>(R.wcol + row, pol)]);
    else t = temp_pool[R.x_off + row];
    const double xt = tanh(__dadd_rn(acc, t));
    return __dadd_rn(__dmul_rn(1.0 - R.leak, xo[row]), __dmul_rn(R.leak, xt));
}

__device__ __forceinline__ double update_row(const RegionDev &R, int row, const double *__restrict__ xo,
                                             const double *__restrict__ u, const double *__restrict__ temp_pool, uint64_t pol)
{
    if (R.ell_stream == 2) return update_row_t<2>(R, row, xo, u, temp_pool, pol);
    return R.ell_stream ? update_row_t<1>(R, row, xo, u, temp_pool, pol) : update_row_t<0>(R, row, xo, u, temp_pool, pol);
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

    pdl_launch();
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
    pdl_wait();   // the plan tables and W_out are constants; the state, inputs and partials belong to the predecessors

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
    const uint64_t pol_last = l2_policy_evict_last();   // only read when RegionDev::ell_stream == 2
    for (int j = tid; j < it.nrows; j += NCONS) {
        const int row = it.row0 + j;
        const double xv = update_row(R, row, xo, u, temp_pool, pol_last);
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
// k_step_persist: the same fused step as k_step, as a PERSISTENT kernel: one CTA per slot (2 per SM), each slot owning
// a contiguous, statically balanced run of work -- no tail wave, no per-item CTA start-up, and the TMA producer warp
// keeps streaming the next item's W_out columns while the consumers reduce / update.
//   * every region is cut into fixed row blocks of part_rows rows (independent of the rank count); a slot's items are
//     runs of whole blocks of one region.  Each block yields its own partial outvec, so the summation structure of a
//     region's readout does not depend on how the model is sharded: results are bit-identical for any numprocs.
//   * consumer thread (rp, cs): lanes of a warp hold the cpi column groups of 32/cpi row pairs; column c of a block
//     goes to group c mod cpi (two alternating accumulators), and a block's partial is reduced over the groups with
//     warp shuffles -- no shared-memory reduction, no block barrier per partial.
//   * a stage is one TMA bulk copy of stage_cols whole W_out columns when the unpadded column stride is already
//     bank-conflict-free for the lane layout (ldw = 136: yes); otherwise one copy per column into slots of ldp doubles.
// Item layout in xs: [local_model (S, only with_model) | x~ of the item's rows].
// ---------------------------------------------------------------------------------------------
struct StepSeg {
    int reg;          // local region index
    int row0, nrows;  // state rows of the item; row0 is a multiple of part_rows
    int part0;        // index of the partial of the item's first row block
    int with_model;   // the item also carries the S local_model columns (first item of a region, fused order)
};

__global__ void __launch_bounds__(NTHREADS, 2)
k_step_persist(const RegionDev *__restrict__ regs, const StepSeg *__restrict__ segs, const int2 *__restrict__ slots,
               const double *__restrict__ x_old, double *__restrict__ x_new, const double *__restrict__ u_pool,
               const long long *__restrict__ u_offs, int u_t, const double *__restrict__ lm_pool,
               const double *__restrict__ temp_pool, double *__restrict__ partials, int ldw_max, int stage_cols, int ldp,
               int xs_cap, int part_rows, int cpi, int nstages)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int stage_bytes = stage_cols * ldp * 8;
    double *xs = reinterpret_cast<double *>(smem_raw + (size_t)nstages * stage_bytes);
    uint64_t *full = reinterpret_cast<uint64_t *>(xs + xs_cap);
    uint64_t *empty = full + nstages;

    pdl_launch();
    const int2 slot = slots[blockIdx.x];   // (first item, item count)
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < nstages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], NCONS_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async;" ::: "memory");
    }
    __syncthreads();

    if (warp == NCONS_WARPS) {
        // ---------------- TMA producer warp: the tile sequence of every item of the slot, back to back ----------------
        // It does NOT wait for the predecessor kernel (PDL): the plan tables and W_out are constants of the step, so the
        // ring is already full when the consumers are released
        int s = 0;
        uint32_t par = 1;   // parity to wait for on empty[s]: passes at once on the first lap
        const uint64_t pol_first = l2_policy_evict_first();
        for (int i = 0; i < slot.y; ++i) {
            const StepSeg sg = segs[slot.x + i];
            const int ldw = regs[sg.reg].ldw, S = regs[sg.reg].S;
            const bool keep_l2 = regs[sg.reg].ell_stream == 2;   // W_out must not flush the L2-resident adjacency
            const double *wout = regs[sg.reg].wout;
            const uint32_t colbytes = (uint32_t)ldw * 8u;
            // segment A: the S local_model columns (fused order only); segment B: the item's state rows
            for (int seg = sg.with_model ? 0 : 1; seg < 2; ++seg) {
                const int ncol = seg == 0 ? S : sg.nrows;
                const double *src = wout + (size_t)(seg == 0 ? 0 : S + sg.row0) * ldw;
                for (int c0 = 0; c0 < ncol; c0 += stage_cols) {
                    const int nc = min(stage_cols, ncol - c0);
                    mbar_wait(&empty[s], par);
                    if (lane == 0) mbar_expect_tx(&full[s], (uint32_t)nc * colbytes);
                    __syncwarp();
                    unsigned char *dst = smem_raw + (size_t)s * stage_bytes;
                    if (ldp == ldw) {   // unpadded tile: the nc columns are one contiguous block -- ONE bulk copy
                        if (lane == 0) {
                            if (keep_l2) tma_load_1d_hint(dst, src + (size_t)c0 * ldw, (uint32_t)nc * colbytes, &full[s], pol_first);
                            else tma_load_1d(dst, src + (size_t)c0 * ldw, (uint32_t)nc * colbytes, &full[s]);
                        }
                    } else {
                        for (int c = lane; c < nc; c += 32)
                            tma_load_1d(dst + (size_t)c * ldp * 8, src + (size_t)(c0 + c) * ldw, colbytes, &full[s]);
                    }
                    if (++s == nstages) { s = 0; par ^= 1; }
                }
            }
        }
        return;
    }

    // ---------------- consumers ----------------
    pdl_wait();   // state, inputs, local_model and the partials belong to the predecessor kernels of the stream
    // lane -> (row pair within the warp, column group): the 8 lanes of one 128-bit load phase are rpw consecutive row
    // pairs (contiguous 16-byte pieces of a column) x 8/rpw column groups, which hit 32 distinct banks when the column
    // stride is 2*ldp = 16 (mod 32) words -- true for the UNPADDED W_out of the standard tiling (ldw = 136)
    const int rpw = 32 / cpi;                     // row pairs per warp
    const int cs = lane / rpw;
    const int rp = warp * rpw + (lane % rpw);
    int s = 0;
    uint32_t par = 0;
    const uint64_t pol_last = l2_policy_evict_last();
    for (int i = 0; i < slot.y; ++i) {
        const StepSeg sg = segs[slot.x + i];
        const RegionDev R = regs[sg.reg];
        const int xs_off = sg.with_model ? R.S : 0;
        // state update of the item's rows -> x_new (global) and x~ (shared).  An item is at most one sweep of the
        // consumer threads, so this phase is ONE latency chain -- about what the ring's prefetch covers
        {
            const double *xo = x_old + R.x_off;
            double *xn = x_new + R.x_off;
            const double *u = u_pool + u_offs[sg.reg] + (long long)u_t * R.D;
            if (sg.with_model) {
                const double *lm = lm_pool + R.lm_off;
                for (int j = tid; j < R.S; j += NCONS) xs[j] = lm[j];
            }
            for (int j = tid; j < sg.nrows; j += NCONS) {
                const int row = sg.row0 + j;
                const double xv = update_row(R, row, xo, u, temp_pool, pol_last);
                if (R.ell_stream == 2) st_keep_f64(xn + row, xv, pol_last);
                else xn[row] = xv;
                xs[xs_off + j] = (row & 1) ? __dmul_rn(xv, xv) : xv;  // even 1-based index squared
            }
        }
        consumer_bar();

        const bool active = rp < (R.ldw >> 1);
        double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
        int part = sg.part0;
        for (int seg = sg.with_model ? 0 : 1; seg < 2; ++seg) {
            const int ncol = seg == 0 ? R.S : sg.nrows;
            const double *xseg = seg == 0 ? xs : xs + xs_off;
            int to_flush = part_rows;             // columns left in the current row block (segment B only)
            for (int c0 = 0; c0 < ncol; c0 += stage_cols) {
                const int nc = min(stage_cols, ncol - c0);
                const double *xk = xseg + c0;
                mbar_wait(&full[s], par);
                if (active) {
                    const double *sb = reinterpret_cast<const double *>(smem_raw + (size_t)s * stage_bytes) + 2 * rp;
                    int c = cs;
                    for (; c + cpi < nc; c += 2 * cpi) {
                        const double2 w0 = *reinterpret_cast<const double2 *>(sb + (size_t)c * ldp);
                        const double2 w1 = *reinterpret_cast<const double2 *>(sb + (size_t)(c + cpi) * ldp);
                        const double x0 = xk[c], x1 = xk[c + cpi];
                        a0 = fma(w0.x, x0, a0);
                        a1 = fma(w0.y, x0, a1);
                        b0 = fma(w1.x, x1, b0);
                        b1 = fma(w1.y, x1, b1);
                    }
                    if (c < nc) {
                        const double2 w0 = *reinterpret_cast<const double2 *>(sb + (size_t)c * ldp);
                        const double x0 = xk[c];
                        a0 = fma(w0.x, x0, a0);
                        a1 = fma(w0.y, x0, a1);
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[s]);
                if (++s == nstages) { s = 0; par ^= 1; }
                if (seg == 1) {
                    to_flush -= nc;
                    if (to_flush == 0 || c0 + nc == ncol) {
                        // the row block is complete: reduce over the column groups (lanes cs = 0..cpi-1), butterfly order
                        double r0 = a0 + b0, r1 = a1 + b1;
                        for (int m = rpw; m < 32; m <<= 1) {
                            r0 += __shfl_xor_sync(0xffffffffu, r0, m);
                            r1 += __shfl_xor_sync(0xffffffffu, r1, m);
                        }
                        if (active && cs == 0)
                            *reinterpret_cast<double2 *>(partials + (size_t)part * ldw_max + 2 * rp) = make_double2(r0, r1);
                        a0 = a1 = b0 = b1 = 0.0;
                        ++part;
                        to_flush = part_rows;
                    }
                }
            }
        }
        consumer_bar();   // xs is rewritten by the next item's update
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

// ---------------------------------------------------------------------------------------------
// k_update_ring: the state update alone with the adjacency STREAMED by a producer warp (round 2).
// k_update_sx still stalls on its own ELL loads: a thread issues a group of loads, waits a full HBM round trip, gathers,
// issues the next group -- ncu showed 12 of 18 stall cycles per issue on that long scoreboard, and nothing is in flight
// while the warps compute.  Here the loads are decoupled from the arithmetic:
//   * ONE CTA per (region, row split) and per SM: 960 consumer threads + 1 TMA producer warp, the region's whole state
//     vector staged once in shared memory (one TMA bulk copy), the input vector beside it;
//   * the producer warp streams row tiles of `tr` rows (up to 960: one row per consumer thread) through an
//     nstages-deep ring: per tile 2W+2 bulk copies (the W column-index slots, the W value slots, W_in value and column
//     of the tile's rows -- each contiguous in the slot-major ELL), completion counted in bytes on the stage's
//     mbarrier.  Tiles are LARGE on purpose: issuing one bulk copy costs the producer warp ~140 cycles whatever its
//     size (measured: 1 KB copies capped both this kernel and the persistent step kernel near 3 TB/s), so each copy
//     must carry several KB; with tr = 960 and W = 6 a tile is 80 KB in 14 copies;
//   * the consumers take the tiles in order, one row per thread, everything from shared memory except the coalesced
//     x_new store.  Accumulation order per row is the entry order, as in update_row: bit-identical states.
// grid (nsplit, regions); dynamic shared memory xs_cap*8 + us_cap*8 + nstages*tile_stride + barriers.
// ---------------------------------------------------------------------------------------------
constexpr int UR_CONS = 960, UR_THREADS = UR_CONS + 32;   // 992 threads
static_assert(UR_THREADS <= 1024 && UR_CONS % 32 == 0, "one CTA: at most 1024 threads");
__global__ void __launch_bounds__(UR_THREADS, 1)
k_update_ring(const RegionDev *__restrict__ regs, const int *__restrict__ region_list, const double *__restrict__ x_old,
              double *__restrict__ x_new, const double *__restrict__ u_pool, const long long *__restrict__ u_offs, int u_t,
              const double *__restrict__ temp_pool, int nsplit, int xs_cap, int us_cap, int w_max, int nstages, int tr)
{
    extern __shared__ __align__(128) unsigned char ur_smem[];
    const int reg = region_list ? region_list[blockIdx.y] : (int)blockIdx.y;
    const RegionDev &R = regs[reg];
    const int n = R.n, D = R.D, W = R.ell_w;
    const int per = (((n + nsplit - 1) / nsplit) + 31) / 32 * 32;
    const int r0 = blockIdx.x * per, r1 = min(n, r0 + per);
    if (r0 >= n) return;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    double *xs = reinterpret_cast<double *>(ur_smem), *us = xs + xs_cap;
    unsigned char *ring = reinterpret_cast<unsigned char *>(us + us_cap);
    const int tile_stride = tr * (12 * w_max + 12);
    const int off_val = w_max * tr * 4, off_winc = off_val + w_max * tr * 8, off_wcol = off_winc + tr * 8;
    uint64_t *full = reinterpret_cast<uint64_t *>(ring + (size_t)nstages * tile_stride);
    uint64_t *empty = full + nstages, *xbar = empty + nstages;
    const int ntiles = (r1 - r0 + tr - 1) / tr;
    const bool compact = R.win_mode == 0;
    const double *__restrict__ xo = x_old + R.x_off;
    if (tid == 0) {
        for (int s = 0; s < nstages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], UR_CONS / 32);
        }
        mbar_init(xbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async;" ::: "memory");
        const uint32_t bytes = (uint32_t)((n + 1) & ~1) * 8u;   // inside the region's padded slot of the x pool
        mbar_expect_tx(xbar, bytes);
        tma_load_1d(xs, xo, bytes, xbar);
    }
    if (compact) {
        const double *__restrict__ u = u_pool + u_offs[reg] + (long long)u_t * D;
        for (int i = tid; i < D; i += UR_THREADS) us[i] = u[i];
    }
    __syncthreads();

    if (warp == UR_CONS / 32) {
        // ---------------- producer: the tiles of this CTA's rows, in order ----------------
        const int ncopy = 2 * W + (compact ? 2 : 0);
        int s = 0;
        uint32_t par = 1;
        for (int k = 0; k < ntiles; ++k) {
            const int t0 = r0 + k * tr, rows = min(tr, r1 - t0);
            mbar_wait(&empty[s], par);
            if (lane == 0) mbar_expect_tx(&full[s], (uint32_t)rows * (uint32_t)(12 * W + (compact ? 12 : 0)));
            __syncwarp();
            unsigned char *dst = ring + (size_t)s * tile_stride;
            for (int c = lane; c < ncopy; c += 32) {
                if (c < W) tma_load_1d(dst + (size_t)c * tr * 4, R.ell_col + (size_t)c * n + t0, (uint32_t)rows * 4u, &full[s]);
                else if (c < 2 * W) tma_load_1d(dst + off_val + (size_t)(c - W) * tr * 8, R.ell_val + (size_t)(c - W) * n + t0, (uint32_t)rows * 8u, &full[s]);
                else if (c == 2 * W) tma_load_1d(dst + off_winc, R.winc + t0, (uint32_t)rows * 8u, &full[s]);
                else tma_load_1d(dst + off_wcol, R.wcol + t0, (uint32_t)rows * 4u, &full[s]);
            }
            if (++s == nstages) { s = 0; par ^= 1; }
        }
        return;
    }

    // ---------------- consumers: every tile in order, one row per thread ----------------
    double *__restrict__ xn = x_new + R.x_off;
    const double leak = R.leak;
    mbar_wait(xbar, 0);
    int s = 0;
    uint32_t par = 0;
    for (int k = 0; k < ntiles; ++k) {
        mbar_wait(&full[s], par);
        const unsigned char *tile = ring + (size_t)s * tile_stride;
        const int row = r0 + k * tr + tid;
        if (tid < tr && row < r1) {
            const int *tc = reinterpret_cast<const int *>(tile) + tid;
            const double *tv = reinterpret_cast<const double *>(tile + off_val) + tid;
            double acc = 0.0;
            int sl = 0;
            for (; sl + 3 <= W; sl += 3) {
                const int c0 = tc[(size_t)sl * tr], c1 = tc[(size_t)(sl + 1) * tr], c2 = tc[(size_t)(sl + 2) * tr];
                const double v0 = tv[(size_t)sl * tr], v1 = tv[(size_t)(sl + 1) * tr], v2 = tv[(size_t)(sl + 2) * tr];
                const double x0 = xs[c0], x1 = xs[c1], x2 = xs[c2];
                acc = fma(v0, x0, acc);
                acc = fma(v1, x1, acc);
                acc = fma(v2, x2, acc);
            }
            for (; sl < W; ++sl) acc = fma(tv[(size_t)sl * tr], xs[tc[(size_t)sl * tr]], acc);
            double tin;
            if (compact) {
                const double wv = reinterpret_cast<const double *>(tile + off_winc)[tid];
                const int wc = reinterpret_cast<const int *>(tile + off_wcol)[tid];
                tin = __dmul_rn(wv, us[wc]);
            } else {
                tin = temp_pool[R.x_off + row];
            }
            const double xt = tanh(__dadd_rn(acc, tin));
            xn[row] = __dadd_rn(__dmul_rn(1.0 - leak, xs[row]), __dmul_rn(leak, xt));
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
        if (++s == nstages) { s = 0; par ^= 1; }
    }
}

// ---------------------------------------------------------------------------------------------
// k_sync_persist: the WHOLE time loop of synchronize in one launch (src/mod_reservoir.f90:1354-1381, the slab-ocean
// twin src/mod_slab_ocean_reservoir.f90:1237-1266):   do i = 1, T:  x <- (1-leak) x + leak tanh(A x + W_in u(:, i)).
// A step-per-launch kernel streams every region's adjacency from HBM once per step (k_update_ring: 0.80 of the HBM roof
// and no further).  Regions do not interact during the spin-up, so the loop order can be turned inside out: ONE CTA per
// SM takes a region and runs all T steps of it before the next region.  Then
//   * the state vector never leaves shared memory (two copies, read one / write the other, one consumer barrier per step);
//   * the region's adjacency (484 KB at the headline config) is re-read every step by the TMA producer warp, but only
//     one region per SM is live at any time: 148 x 484 KB = 72 MB stays resident in the 126 MB L2, so steps 2..T stream from
//     L2, not from HBM -- cache blocking over time, which a launch per step cannot have at 1152 regions (558 MB per step);
//   * the step's input vector u(:, t) arrives by one bulk copy into a two-deep buffer of its own;
//   * the tiles come from a TILE-MAJOR copy of the adjacency (k_sync_pack): one bulk copy per tile.  Fed from the
//     slot-major ELL (2W+2 copies of 1.5-3 KB per tile) the first version ran at exactly the speed of the step launches
//     -- 17.6 us per region-step, the producer warp's ~140 cycles per copy x 14 copies x 15 tiles.
// Ring: nstages slots of `tr` rows, `ngroups` consumer groups of `tr` threads (one row per thread); tile g (counted over
// the whole launch) lives in slot g % nstages and belongs to group g % ngroups; nstages % ngroups == 0, so a slot is
// always drained by the same group and a parity wait can never alias an older phase.
// Per-row arithmetic and order are those of update_row: the states are bit-identical to T launches of any of the
// update kernels above.  The state pool is updated in place.  Needs every listed region in compact-W_in form with
// n % 4 == 0 and an even D (bulk copies are 16-byte multiples); the host falls back to step launches otherwise.
// grid min(regions, SMs); block ngroups*tr + 32; dynamic shared memory 2*xs_cap*8 + 2*us_cap*8 + nstages*tile + barriers.
// ---------------------------------------------------------------------------------------------
constexpr int SP_MAX_THREADS = 1024;
// Tile-major copy of every region's adjacency and compact W_in for k_sync_persist.  Tile k of region r is ONE contiguous
// block of tr rows, laid out for 128-bit shared-memory loads (thread j of the tile's group owns row j):
//     [VP][tr] double2   values of slots (2p, 2p+1)                         VP = ceil(w_max / 2)
//     [CP][tr] 8 x u16   column indices of slots 8q .. 8q+7; the LAST u16 of the LAST slab is the row's W_in column
//                                                                           CP = ceil((w_max + 1) / 8)
//     [tr]     double    the row's W_in value
// (rows past n and slots past the region's width are zero).  72 bytes per row at width 6 against 84 in the slot-major
// ELL: 16-bit indices suffice for n, D <= 65535.  One bulk copy per tile; 5 shared-memory loads per row instead of 14.
// grid (tiles of the largest region, local regions); rebuilt when the ring geometry or the adjacency changes.
__host__ __device__ inline int sp_vp(int w_max) { return (w_max + 1) / 2; }
__host__ __device__ inline int sp_cp(int w_max) { return (w_max + 1 + 7) / 8; }
__host__ __device__ inline int sp_row_bytes(int w_max) { return 16 * sp_vp(w_max) + 16 * sp_cp(w_max) + 8; }

__global__ void __launch_bounds__(256)
k_sync_pack(const RegionDev *__restrict__ regs, unsigned char *__restrict__ pack, const long long *__restrict__ pack_off,
            int tr, int w_max)
{
    const RegionDev &R = regs[blockIdx.y];
    const int n = R.n, W = R.ell_w;
    const int t0 = blockIdx.x * tr;
    if (n <= 0 || t0 >= n || R.winc == nullptr) return;
    const int VP = sp_vp(w_max), CP = sp_cp(w_max);
    const size_t tile_stride = (size_t)tr * sp_row_bytes(w_max);
    unsigned char *dst = pack + pack_off[blockIdx.y] + (size_t)blockIdx.x * tile_stride;
    double *pv = reinterpret_cast<double *>(dst);
    unsigned short *pc = reinterpret_cast<unsigned short *>(dst + (size_t)VP * tr * 16);
    double *pw = reinterpret_cast<double *>(dst + (size_t)(VP + CP) * tr * 16);
    for (int j = threadIdx.x; j < tr; j += blockDim.x) {
        const int row = t0 + j;
        const bool ok = row < n;
        for (int s = 0; s < 2 * VP; ++s) {
            const bool in = ok && s < W;
            pv[((size_t)(s >> 1) * tr + j) * 2 + (s & 1)] = in ? R.ell_val[(size_t)s * n + row] : 0.0;
        }
        for (int s = 0; s < 8 * CP; ++s) {
            const bool in = ok && s < W;
            unsigned short c = in ? (unsigned short)R.ell_col[(size_t)s * n + row] : (unsigned short)0;
            if (s == 8 * CP - 1) c = ok ? (unsigned short)R.wcol[row] : (unsigned short)0;
            pc[((size_t)(s >> 3) * tr + j) * 8 + (s & 7)] = c;
        }
        pw[j] = ok ? R.winc[row] : 0.0;
    }
}

// 128-bit shared-memory loads, spelled out: left to the compiler, a double2 / uint4 access whose halves are used under
// different predicates is split into 64- or 32-bit loads, and with a 16-byte lane stride those are 2- and 4-way bank
// conflicts (ncu: 43 excess wavefronts per warp-row instead of 25).  volatile: never hoisted above the mbarrier wait.
__device__ __forceinline__ double2 lds128_f64(const void *p)
{
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(smem_u32(p)));
    return v;
}
__device__ __forceinline__ uint4 lds128_u32(const void *p)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(smem_u32(p)));
    return v;
}

// one row of a packed tile: y = sum over the region's W slots in slot order (FMA chain, as update_row), then the W_in
// product, tanh and the leak.  VPT > 0: compile-time number of value pairs with ONE column slab (w_max <= min(2 VPT, 7));
// VPT == 0: any width.  RS: the ML-only restart of training (src/mod_reservoir.f90:1034) -- the SpMV operand is the
// squared copy of the state, i.e. gathered values of odd 0-based columns are squared first.
template <int VPT, bool RS = false>
__device__ __forceinline__ double sp_row(const unsigned char *__restrict__ tile, int gt, int tr, int w_max, int W,
                                         const double *__restrict__ xr, const double *__restrict__ uu, double x_own, double leak)
{
    const int VP = VPT > 0 ? VPT : sp_vp(w_max), CP = VPT > 0 ? 1 : sp_cp(w_max);
    const double2 *tv = reinterpret_cast<const double2 *>(tile) + gt;
    const uint4 *tc = reinterpret_cast<const uint4 *>(tile + (size_t)VP * tr * 16) + gt;
    const double wv = reinterpret_cast<const double *>(tile + (size_t)(VP + CP) * tr * 16)[gt];
    double acc = 0.0;
    unsigned wc;
    if constexpr (VPT > 0) {
        const uint4 cw = lds128_u32(tc);
        const unsigned cwa[4] = {cw.x, cw.y, cw.z, cw.w};
        double2 v[VPT];
#pragma unroll
        for (int p = 0; p < VPT; ++p) v[p] = lds128_f64(tv + (size_t)p * tr);
        double xa[VPT], xb[VPT];
#pragma unroll
        for (int p = 0; p < VPT; ++p) {
            const unsigned ca = cwa[p] & 0xffffu, cb = p == 3 ? 0u : cwa[p] >> 16;   // pair 3's upper half is the W_in column
            xa[p] = xr[ca];
            xb[p] = xr[cb];
            if (RS) {
                if (ca & 1u) xa[p] = __dmul_rn(xa[p], xa[p]);
                if (cb & 1u) xb[p] = __dmul_rn(xb[p], xb[p]);
            }
        }
#pragma unroll
        for (int p = 0; p < VPT; ++p) {
            if (2 * p < W) acc = fma(v[p].x, xa[p], acc);
            if (2 * p + 1 < W) acc = fma(v[p].y, xb[p], acc);
        }
        wc = cw.w >> 16;
    } else {
        uint4 cw = lds128_u32(tc);
        for (int p = 0; p < VP; ++p) {
            if ((p & 3) == 0 && p) cw = lds128_u32(tc + (size_t)(p >> 2) * tr);
            const unsigned pair = (p & 3) == 0 ? cw.x : (p & 3) == 1 ? cw.y : (p & 3) == 2 ? cw.z : cw.w;
            const double2 v = lds128_f64(tv + (size_t)p * tr);
            const unsigned ca = pair & 0xffffu, cb = pair >> 16;
            double ga = xr[ca], gb = xr[cb];
            if (RS) {
                if (ca & 1u) ga = __dmul_rn(ga, ga);
                if (cb & 1u) gb = __dmul_rn(gb, gb);
            }
            if (2 * p < W) acc = fma(v.x, ga, acc);
            if (2 * p + 1 < W) acc = fma(v.y, gb, acc);
        }
        if (((VP - 1) >> 2) != CP - 1) cw = lds128_u32(tc + (size_t)(CP - 1) * tr);   // the last slab was not the last one read
        wc = cw.w >> 16;
    }
    const double xt = tanh(__dadd_rn(acc, __dmul_rn(wv, uu[wc])));
    return __dadd_rn(__dmul_rn(1.0 - leak, x_own), __dmul_rn(leak, xt));
}

template <int VPT>
__global__ void __launch_bounds__(SP_MAX_THREADS, 1)
k_sync_persist(const RegionDev *__restrict__ regs, const int *__restrict__ region_list, int nreg, double *__restrict__ x_pool,
               const double *__restrict__ u_pool, const long long *__restrict__ u_offs, int T, int xs_cap, int us_cap,
               int w_max, int nstages, int tr, int ngroups, const unsigned char *__restrict__ pack,
               const long long *__restrict__ pack_off)
{
    extern __shared__ __align__(128) unsigned char sp_smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ncons = ngroups * tr;           // consumer threads; the last warp of the block is the producer
    double *xs = reinterpret_cast<double *>(sp_smem), *us = xs + 2 * (size_t)xs_cap;
    unsigned char *ring = reinterpret_cast<unsigned char *>(us + 2 * (size_t)us_cap);
    const int tile_stride = tr * sp_row_bytes(w_max);
    uint64_t *full = reinterpret_cast<uint64_t *>(ring + (size_t)nstages * tile_stride);
    uint64_t *empty = full + nstages, *ufull = empty + nstages, *uempty = ufull + 2, *xbar = uempty + 2;
    if (tid == 0) {
        for (int s = 0; s < nstages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], tr / 32);     // the warps of the one group that drains the slot
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&ufull[b], 1);
            mbar_init(&uempty[b], ncons / 32); // every consumer warp, once per step
        }
        mbar_init(xbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async;" ::: "memory");
    }
    __syncthreads();

    if (warp == ncons / 32) {
        // ---------------- producer: for every region of this CTA, T passes over its row tiles ----------------
        int s = 0;
        uint32_t par = 1;
        unsigned ug = 0;
        for (int ri = blockIdx.x; ri < nreg; ri += gridDim.x) {
            const int reg = region_list ? region_list[ri] : ri;
            const RegionDev &R = regs[reg];
            const int n = R.n, D = R.D;
            const int ntiles = (n + tr - 1) / tr;
            const double *__restrict__ u0 = u_pool + u_offs[reg];
            const unsigned char *__restrict__ tiles = pack + pack_off[reg];
            for (int t = 0; t < T; ++t) {
                const unsigned b = ug & 1u;
                mbar_wait(&uempty[b], ((ug >> 1) & 1u) ^ 1u);   // the consumers are done with step t-2
                if (lane == 0) {
                    mbar_expect_tx(&ufull[b], (uint32_t)D * 8u);
                    tma_load_1d(us + (size_t)b * us_cap, u0 + (long long)t * D, (uint32_t)D * 8u, &ufull[b]);
                }
                ++ug;
                for (int k = 0; k < ntiles; ++k) {
                    mbar_wait(&empty[s], par);
                    if (lane == 0) {
                        // ONE bulk copy per tile: issuing a copy costs the warp ~140 cycles whatever its size, and the
                        // 2W+2 slot-major pieces per tile of the first version made that issue rate the bound
                        mbar_expect_tx(&full[s], (uint32_t)tile_stride);
                        tma_load_1d(ring + (size_t)s * tile_stride, tiles + (size_t)k * tile_stride, (uint32_t)tile_stride, &full[s]);
                    }
                    if (++s == nstages) { s = 0; par ^= 1u; }
                }
            }
        }
        return;
    }

    // ---------------- consumers: group gi takes every ngroups-th tile of the launch, one row per thread ----------------
    const int gi = tid / tr, gt = tid - gi * tr;
    int s = gi;              // slot of this group's next tile; it advances by ngroups per tile (nstages % ngroups == 0)
    uint32_t par = 0;
    int k = gi;              // index of that tile inside its step; carried over step and region boundaries
    unsigned ug = 0, rg = 0;
    for (int ri = blockIdx.x; ri < nreg; ri += gridDim.x, ++rg) {
        const RegionDev &R = regs[region_list ? region_list[ri] : ri];
        const int n = R.n, W = R.ell_w;
        const int ntiles = (n + tr - 1) / tr;
        const double leak = R.leak;
        double *__restrict__ xg = x_pool + R.x_off;
        if (tid == 0) {
            // the previous region's write-back read xs through the generic proxy; order it before the bulk copy
            asm volatile("fence.proxy.async;" ::: "memory");
            const uint32_t bytes = (uint32_t)((n + 1) & ~1) * 8u;   // inside the region's padded slot of the x pool
            mbar_expect_tx(xbar, bytes);
            tma_load_1d(xs, xg, bytes, xbar);
        }
        mbar_wait(xbar, rg & 1u);
        for (int t = 0; t < T; ++t) {
            const double *__restrict__ xr = xs + (size_t)(t & 1) * xs_cap;
            double *__restrict__ xw = xs + (size_t)((t + 1) & 1) * xs_cap;
            const unsigned b = ug & 1u;
            mbar_wait(&ufull[b], (ug >> 1) & 1u);
            const double *__restrict__ uu = us + (size_t)b * us_cap;
            for (; k < ntiles; k += ngroups) {
                mbar_wait(&full[s], par);
                const unsigned char *tile = ring + (size_t)s * tile_stride;
                const int row = k * tr + gt;
                if (row < n) xw[row] = sp_row<VPT>(tile, gt, tr, w_max, W, xr, uu, xr[row], leak);
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[s]);
                s += ngroups;
                if (s >= nstages) { s -= nstages; par ^= 1u; }
            }
            k -= ntiles;     // the group's first tile of the next step (of this or the next region)
            ++ug;
            __syncwarp();
            if (lane == 0) mbar_arrive(&uempty[b]);
            // every row of x(t+1) is written and every read of x(t) is over before the buffers swap roles
            asm volatile("bar.sync 1, %0;" ::"r"(ncons) : "memory");
        }
        const double *__restrict__ xf = xs + (size_t)(T & 1) * xs_cap;
        for (int i = tid; i < n; i += ncons) xg[i] = xf[i];
        asm volatile("bar.sync 1, %0;" ::"r"(ncons) : "memory");   // xs[0] is free for the next region's bulk copy
    }
}

// dense W_in fallback: temp = matmul(win, u) for regions with win_mode == 1 (src/mod_reservoir.f90:1445).
// grid: (ceil(n_max/256), nregions)
__global__ void k_win_dense(const RegionDev *__restrict__ regs, const int *__restrict__ region_list,
                            const double *__restrict__ u_pool, const long long *__restrict__ u_offs, int u_t,
                            double *__restrict__ temp_pool)
{
    const int reg = region_list ? region_list[blockIdx.y] : (int)blockIdx.y;
    const RegionDev R = regs[reg];
    if (R.win_mode != 1) return;
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= R.n) return;
    const double *u = u_pool + u_offs[reg] + (long long)u_t * R.D;
    double acc = 0.0;
    for (int i = 0; i < R.D; ++i) acc = fma(R.win_dense[(size_t)i * R.n + row], u[i], acc);
    temp_pool[R.x_off + row] = acc;
}

// partials of a region summed in item order, then unstandardize_state_vec_res (src/res_domain.f90:1424-1475):
// v*std then +mean, two roundings (src/mod_utilities.f90:799-829).  One block per region.
// model_part != 0 (split-order readout of the overlapped step): the partials hold only W_out[:, S:]*x~ (the
// reference's v_ml, src/mod_reservoir.f90:1460) and this kernel adds v_p = W_out[:, 0:S]*local_model (:1459)
// once the host model's forecast has arrived.
// use_parts != 0: the partials are those of the persistent step kernel (one per fixed row block of the region,
// RegionDev::part0 / nparts) instead of one per classic step item.
// Fused all-gather (pt.world > 1): the outvec is staged in shared memory and each warp pushes it to one rank's
// gathered buffer with coalesced 16-byte stores over NVLink; ONE system fence per CTA, and the last CTA publishes
// the step in every rank's flag slot, one lane per rank (the release stores overlap instead of queueing).
constexpr int FIN_GROUPS = 4;     // column groups of the model-part GEMV
constexpr int FIN_PMAX = 160;     // threads per group (outputs are strided over them: any chunk_size_prediction)

__global__ void __launch_bounds__(FIN_GROUPS *FIN_PMAX, 2)
k_readout_finish(const RegionDev *__restrict__ regs, const double *__restrict__ partials, int ldw_max,
                 double *__restrict__ out_pool, int unstandardize, int model_part,
                 const double *__restrict__ lm_pool, PeerTable pt, unsigned long long seq,
                 long long peer_off, unsigned int *__restrict__ done_counter, double *__restrict__ vp_pool,
                 double *__restrict__ vml_pool, int use_parts, double *__restrict__ lm_out)
{
    extern __shared__ __align__(16) double s_fin[];   // [pstride] outvec staging, then [FIN_GROUPS][pstride] (overlapped mode)
    __shared__ int s_last;
    pdl_launch();
    const RegionDev R = regs[blockIdx.x];   // the region table is constant after sml_finalize
    pdl_wait();
    const int pstride = (R.P + 1) & ~1;
    double *s_out = s_fin, *s_vp = s_fin + pstride;
    const int grp = threadIdx.x / FIN_PMAX, p0 = threadIdx.x % FIN_PMAX;
    const int first = use_parts ? R.part0 : R.item0, cnt = use_parts ? R.nparts : R.nitems;
    if (model_part) {
        // v_p = W_out[:, 0:S] * local_model: group g takes columns g, g+4, ...; 4 independent FMA chains per thread
        // keep the loads in flight; the groups are combined in fixed order below (deterministic)
        const double *lm = lm_pool + R.lm_off;
        // 8 independent FMA chains per thread: eight W_out loads in flight each (ncu, round 2: with four the kernel sat at
        // 26 % of the DRAM roof, latency-bound on its 197 MB); the chains are combined in fixed order (deterministic)
        for (int p = p0; p < R.P; p += FIN_PMAX) {
            const double *w = R.wout + p;
            double a[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) a[u] = 0.0;
            int j = grp;
            for (; j + 7 * FIN_GROUPS < R.S; j += 8 * FIN_GROUPS) {
                double wv[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) wv[u] = __ldg(w + (size_t)(j + u * FIN_GROUPS) * R.ldw);
#pragma unroll
                for (int u = 0; u < 8; ++u) a[u] = fma(wv[u], lm[j + u * FIN_GROUPS], a[u]);
            }
            for (; j < R.S; j += FIN_GROUPS) a[0] = fma(__ldg(w + (size_t)j * R.ldw), lm[j], a[0]);
            s_vp[grp * pstride + p] = ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
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
                for (int c = 0; c < cnt; ++c) vml += partials[(size_t)(first + c) * ldw_max + p];
                vp_pool[R.out_off + p] = v;
                vml_pool[R.out_off + p] = vml;
            }
            for (int c = 0; c < cnt; ++c) v += partials[(size_t)(first + c) * ldw_max + p];
            if (R.lm_self && lm_out) lm_out[R.lm_off + p] = v;   // before the un-standardisation, as the reference does
            if (unstandardize) {
                const int ms = R.out_ms[p];
                if (ms >= 0) v = __dadd_rn(__dmul_rn(v, R.std[ms]), R.mean[ms]);
            }
            out_pool[R.out_off + p] = v;
            s_out[p] = v;
        }
    }
    if (pt.world > 1) {
        __syncthreads();
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
        const long long dst0 = peer_off + R.out_off;
        const bool vec = ((R.P & 1) == 0) && ((dst0 & 1) == 0);
        for (int k = warp; k < pt.world; k += nwarps) {
            double *dst = pt.atmo(k) + dst0;
            if (vec) {
                for (int i = lane; i < (R.P >> 1); i += 32)
                    reinterpret_cast<double2 *>(dst)[i] = reinterpret_cast<const double2 *>(s_out)[i];
            } else {
                for (int i = lane; i < R.P; i += 32) dst[i] = s_out[i];
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence_system();   // cumulative: covers the block's peer stores ordered before the barrier
            const unsigned int old = atomicAdd(done_counter, 1u);
            s_last = (old == gridDim.x - 1) ? 1 : 0;
            if (s_last) {
                *done_counter = 0;
                __threadfence();
            }
        }
        __syncthreads();
        if (s_last && (int)threadIdx.x < pt.world) {   // every local region's outvec is on its way: publish the step
            __threadfence_system();
            st_release_sys(pt.flag(threadIdx.x, XF_ATMO, pt.rank), seq);
        }
    }
}

// k_peer_push: `count` doubles from src into the `section` of every rank's exchange block at dst_off, then the
// sequence number into flag [section][flag_slot] of every rank.  Used for the ocean reservoirs' outvec slab (the
// all-gather after predict_slab_ml, src/mpires.f90:375-454) and for the root's forecast block (scatter + bcast,
// :606-744).  grid (nchunk, world); block (c, k) copies chunk c to rank k with 16-byte stores.
__global__ void __launch_bounds__(256)
k_peer_push(const double *__restrict__ src, long long count, PeerTable pt, int section, long long dst_off, int skip_self,
            unsigned long long seq, int flag_slot, unsigned int *__restrict__ done_counter)
{
    __shared__ int s_last;
    const int k = blockIdx.y;
    if (!(skip_self && k == pt.rank)) {
        double *dst = (section == XF_OCEAN ? pt.ocean(k) : pt.fcst(k)) + dst_off;
        long long per = (count + gridDim.x - 1) / gridDim.x;
        per = (per + 1) & ~1LL;
        const long long i0 = (long long)blockIdx.x * per, i1 = min(count, i0 + per);
        const bool vec = ((reinterpret_cast<unsigned long long>(src) | reinterpret_cast<unsigned long long>(dst)) & 15ULL) == 0;
        if (vec) {
            const long long n2 = (i1 - i0) >> 1;
            const double2 *s2 = reinterpret_cast<const double2 *>(src + i0);
            double2 *d2 = reinterpret_cast<double2 *>(dst + i0);
            for (long long i = threadIdx.x; i < n2; i += blockDim.x) d2[i] = s2[i];
            if (((i1 - i0) & 1) && threadIdx.x == 0) dst[i1 - 1] = src[i1 - 1];
        } else {
            for (long long i = i0 + threadIdx.x; i < i1; i += blockDim.x) dst[i] = src[i];
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        const unsigned int old = atomicAdd(done_counter, 1u);
        s_last = (old == gridDim.x * gridDim.y - 1) ? 1 : 0;
        if (s_last) {
            *done_counter = 0;
            __threadfence();
        }
    }
    __syncthreads();
    if (s_last && (int)threadIdx.x < pt.world) {
        __threadfence_system();
        st_release_sys(pt.flag(threadIdx.x, section, flag_slot), seq);
    }
}

// k_wait_flag: the consumer side of the forecast push: wait until the root has published sequence number seq, then
// (optionally) move the TISR field that came with the forecast into G.  One block; it gives up after ~10 s and
// raises *err instead of hanging (sml_peer_check reports it).
__global__ void __launch_bounds__(256)
k_wait_flag(const unsigned long long *__restrict__ flag, unsigned long long seq, int *__restrict__ err,
            const double *__restrict__ copy_src, double *__restrict__ copy_dst, int copy_n)
{
    if (threadIdx.x == 0) {
        const long long t0 = clock64();
        while (ld_acquire_sys(flag) < seq) {
            if (clock64() - t0 > 20000000000LL) {
                *err = 2;
                break;
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < copy_n; i += blockDim.x) copy_dst[i] = ld_cg_f64(copy_src + i);
}

// k_pack_grids: the root's grid assembly of sendrecievegrid in one launch.
//  blocks [0, nsc): tile_full_grid_with_local_state_vec_res1d for every region of the model
//     (src/res_domain.f90:791-826) + the clamps: q < 1e-6 -> 1e-6 (src/mpires.f90:460-462), precip < 1e-5 -> 0
//     (:486-490).  With peers, first wait until every rank has published this step's outvecs.
//  blocks [nsc, ..): wholegrid_sst (src/mpires.f90:288-290, 315-328, 470-484).  mode 0: every cell takes the first
//     fx*fy outputs of its region's ocean reservoir (tile_full_2d_grid_with_local_res, src/res_domain.f90:828-850);
//     regions without one hold 272.0 in their slab row (:323-326, 383); mode 1: prescribed field; mode -1: no SST.
//  Failure detection: the assembled atmosphere grid is checked on the way -- non-finite values and the bounds SPEEDY's
//     own input check applies before it agrees to run (src/ppo_iogrid.f90:562-577: u in [-150,150], v in [-120,120],
//     T in [160,330], q in [-6,30]) set bits in *status (sticky; SML_GRID_* in the header).
struct PackArgs {
    const double *gathered;
    const int *out_dst;
    int total;
    double *G;
    long long precip_lo, precip_hi, w4d_hi, sst_off;
    int nsc;
    // peers
    const unsigned long long *my_flags;   // this rank's flags [XF_KINDS][MAX_PEERS]
    int world;
    unsigned long long seq, ocean_seq;
    int *err;
    int *status;
    // sst
    const double *base, *mask, *prescribed, *ocean_out;
    const int *cell_region, *cell_slot;
    int ocean_P, sst_mode;
};

__global__ void k_pack_grids(PackArgs a)
{
    pdl_launch();
    pdl_wait();
    if (a.world > 1) {
        if ((int)threadIdx.x < 2 * a.world) {
            const int kind = threadIdx.x / a.world, r = threadIdx.x % a.world;
            const unsigned long long want = kind ? a.ocean_seq : a.seq;
            const unsigned long long *f = a.my_flags + kind * MAX_PEERS + r;
            const long long t0 = clock64();
            while (want > 0 && ld_acquire_sys(f) < want) {
                if (clock64() - t0 > 20000000000LL) {  // ~10 s: a peer died; report instead of hanging
                    *a.err = 1;
                    break;
                }
            }
        }
        __syncthreads();
    }
    if ((int)blockIdx.x < a.nsc) {
        const int i = blockIdx.x * blockDim.x + threadIdx.x;
        if (i >= a.total) return;
        const int dst = a.out_dst[i];
        double v = ld_cg_f64(a.gathered + i);
        int bad = 0;
        if (dst < a.w4d_hi) {
            const int var = dst & 3;   // wholegrid4d(var, x, y, z): 0 T, 1 u, 2 v, 3 q
            if (var == 3 && v < 0.000001) v = 0.000001;
            if (var == 0) bad = (v < 160.0 || v > 330.0) ? 8 : 0;
            else if (var == 1) bad = (v < -150.0 || v > 150.0) ? 2 : 0;
            else if (var == 2) bad = (v < -120.0 || v > 120.0) ? 4 : 0;
            else bad = (v < -6.0 || v > 30.0) ? 16 : 0;
        } else if (dst >= a.precip_lo && dst < a.precip_hi) {
            if (v < 0.00001) v = 0.0;
        }
        if (!(fabs(v) <= 1.7976931348623157e308)) bad = 1;   // NaN or Inf
        if (bad) atomicOr(a.status, bad);
        a.G[dst] = v;
        return;
    }
    const int e = (blockIdx.x - a.nsc) * blockDim.x + threadIdx.x;
    if (e >= 96 * 48 || a.sst_mode < 0) return;
    double v;
    if (a.sst_mode == 1) v = a.prescribed[e];
    else v = ld_cg_f64(a.ocean_out + (size_t)a.cell_region[e] * a.ocean_P + a.cell_slot[e]);
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
    pdl_launch();
    const RegionDev R = regs[blockIdx.x];
    pdl_wait();
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

// k_exchange_fused: k_pack_grids and k_build_inputs as ONE cooperative launch for the device-resident step (ML-only
// runs, and the `value` leg of bench.py): every CTA scatters its share of the gathered outvecs into G (after the wait
// for the peers' flags), a grid barrier makes G complete, then the CTAs rebuild the local regions' feedback and
// local_model.  Saves one launch boundary and one kernel ramp per step -- at 144 regions per GPU the four small kernels
// and their gaps were a fifth of the step.  Same element-wise arithmetic as the two kernels it replaces: bit-identical.
__device__ __forceinline__ unsigned ld_acquire_gpu_u32x(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(512)
k_exchange_fused(PackArgs a, const RegionDev *__restrict__ regs, int nregions, const double *__restrict__ F,
                 double *__restrict__ fb_pool, double *__restrict__ lm_pool, int do_model, unsigned *__restrict__ ctr,
                 unsigned target)
{
    if (a.world > 1) {
        if ((int)threadIdx.x < 2 * a.world) {
            const int kind = threadIdx.x / a.world, r = threadIdx.x % a.world;
            const unsigned long long want = kind ? a.ocean_seq : a.seq;
            const unsigned long long *f = a.my_flags + kind * MAX_PEERS + r;
            const long long t0 = clock64();
            while (want > 0 && ld_acquire_sys(f) < want) {
                if (clock64() - t0 > 20000000000LL) {
                    *a.err = 1;
                    break;
                }
            }
        }
        __syncthreads();
    }
    const int nthreads = gridDim.x * blockDim.x, gtid = blockIdx.x * blockDim.x + threadIdx.x;
    for (int i = gtid; i < a.total; i += nthreads) {
        const int dst = a.out_dst[i];
        double v = ld_cg_f64(a.gathered + i);
        int bad = 0;
        if (dst < a.w4d_hi) {
            const int var = dst & 3;
            if (var == 3 && v < 0.000001) v = 0.000001;
            if (var == 0) bad = (v < 160.0 || v > 330.0) ? 8 : 0;
            else if (var == 1) bad = (v < -150.0 || v > 150.0) ? 2 : 0;
            else if (var == 2) bad = (v < -120.0 || v > 120.0) ? 4 : 0;
            else bad = (v < -6.0 || v > 30.0) ? 16 : 0;
        } else if (dst >= a.precip_lo && dst < a.precip_hi) {
            if (v < 0.00001) v = 0.0;
        }
        if (!(fabs(v) <= 1.7976931348623157e308)) bad = 1;
        if (bad) atomicOr(a.status, bad);
        a.G[dst] = v;
    }
    if (a.sst_mode >= 0)
        for (int e = gtid; e < 96 * 48; e += nthreads) {
            double v;
            if (a.sst_mode == 1) v = a.prescribed[e];
            else v = ld_cg_f64(a.ocean_out + (size_t)a.cell_region[e] * a.ocean_P + a.cell_slot[e]);
            if (a.mask[e] > 0.0) v = a.base[e];
            if (v < 272.0) v = 272.0;
            a.G[a.sst_off + e] = v;
        }
    // ---- grid barrier: G is complete (monotonic arrival counter, release / acquire at gpu scope)
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(ctr, 1u);
        while (ld_acquire_gpu_u32x(ctr) < target) {
        }
    }
    __syncthreads();
    // ---- feedback / local_model of the local regions (k_build_inputs)
    for (int reg = blockIdx.x; reg < nregions; reg += gridDim.x) {
        const RegionDev R = regs[reg];
        for (int d = threadIdx.x; d < R.D; d += blockDim.x) {
            double v = ld_cg_f64(a.G + R.fb_src[d]);
            const int ms = R.fb_ms[d];
            if (ms >= 0) v = __ddiv_rn(__dsub_rn(v, R.mean[ms]), R.std[ms]);
            fb_pool[R.fb_off + d] = v;
        }
        if (do_model)
            for (int sidx = threadIdx.x; sidx < R.S; sidx += blockDim.x) {
                double v = F[R.lm_src[sidx]];
                const int ms = R.lm_ms[sidx];
                if (ms >= 0) v = __ddiv_rn(__dsub_rn(v, R.mean[ms]), R.std[ms]);
                lm_pool[R.lm_off + sidx] = v;
            }
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
