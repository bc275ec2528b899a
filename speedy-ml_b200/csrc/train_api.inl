// train_api.inl -- C entry points of the training path (included at the end of engine.cu).
//
// sml_train_begin / sml_train_feed / sml_train_solve / sml_train_end replace, for a wave of regions,
// train_reservoir's inner sequence (src/mod_reservoir.f90:287-316):
//   initialize_chunk_training -> reservoir_layer_chunking_hybrid|_ml per phase (state generation,
//   chunking_matmul per batch) -> fit_chunk_hybrid|_ml (ridge terms, mldivide).
// Batches only decide which states are kept (floor(TL/batch_size)*batch_size) and, on the ML-only path,
// where the squared copy is fed back; the Gram itself is accumulated over larger K slabs (summation
// order is the only difference).

namespace {

// state generation alone on the SMs (serial schedule): whether the <6, 384> instantiation of k_train_stategen is used
constexpr bool SG_WIDE_DEFAULT = false;   // measured slower (0.91 vs 0.76 s for 384 regions)
constexpr int SG_MIN_WAVE = 64;   // smallest wave for which the in-kernel time loop is the default route
constexpr int SG_RING_MIN_WAVE = 24;   // smallest wave for which the ring kernel (one CTA per region and per SM) is the default

// device buffer freed on every exit path of its scope
struct ScopedDev {
    void *p = nullptr;
    ~ScopedDev() { cudaFree(p); }
};

// LU with partial pivoting + solve, the dgesv the reference calls (src/mod_linalg.f90:145); A, B on device (lu.cuh)
int device_dgesv(sml_engine *h, double *dA, int lda, double *dB, int ldb, int n, int nrhs, int *info_out)
{
    ScopedDev b_ipiv, b_info, b_xch;
    // CTAs of the cooperative panel launch: one per SM at most (SML_LU_GMAX lowers it: test hook for the many-rows-per-CTA path)
    const int Gmax = std::max(1, getenv("SML_LU_GMAX") ? std::min(atoi(getenv("SML_LU_GMAX")), h->num_sms) : h->num_sms);
    const size_t xstride = (size_t)Gmax + (size_t)Gmax * LU_NB + LU_NB + (Gmax + 1) / 2;
    CK(h, cudaMalloc(&b_ipiv.p, sizeof(int) * (size_t)std::max(n, 1)));
    CK(h, cudaMalloc(&b_info.p, 2 * sizeof(int)));
    CK(h, cudaMalloc(&b_xch.p, sizeof(double) * 2 * xstride));
    int *ipiv = static_cast<int *>(b_ipiv.p), *dinfo = static_cast<int *>(b_info.p);
    double *xch = static_cast<double *>(b_xch.p);
    CK(h, cudaMemsetAsync(dinfo, 0, 2 * sizeof(int), h->stream));
    unsigned *ctr = reinterpret_cast<unsigned *>(dinfo + 1);
    // SML_LU_TIMING=1: CUDA-event split of the factorisation (panel / interchanges + trsm + gemm) and the solve on stderr
    const bool timing = getenv("SML_LU_TIMING") && atoi(getenv("SML_LU_TIMING")) != 0;
    float t_panel = 0.f, t_trail = 0.f, t_solve = 0.f;
    cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
    if (timing) for (auto &e : ev) cudaEventCreate(&e);
    unsigned ctr_base = 0;
    const bool gemm_fma = getenv("SML_LU_GEMM") && std::string(getenv("SML_LU_GEMM")) == "fma";   // A/B switch
    CK(h, cudaFuncSetAttribute(k_lu_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LG_SMEM));
    for (int j0 = 0; j0 < n; j0 += LU_NB) {
        const int nb = std::min(LU_NB, n - j0);
        const int rows = n - j0;
        int rows_per = std::max(64, (rows + Gmax - 1) / Gmax);
        int G = (rows + rows_per - 1) / rows_per;
        const size_t smem = sizeof(double) * (size_t)nb * (rows_per + 1);
        if (smem > 200 * 1024) { FAIL(h, "mldivide: n = %d is too large for the panel kernel", n); }
        if (smem > 48 * 1024) CK(h, cudaFuncSetAttribute(k_lu_panel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (timing) cudaEventRecord(ev[0], h->stream);
        {
            double *A_ = dA; int lda_ = lda, n_ = n, j0_ = j0, nb_ = nb, rp_ = rows_per;
            void *args[] = {&A_, &lda_, &n_, &j0_, &nb_, &rp_, &ipiv, &dinfo, &xch, &ctr, &ctr_base};
            cudaError_t e = cudaLaunchCooperativeKernel((const void *)k_lu_panel, dim3(G), dim3(LU_PANEL_THREADS), args, smem, h->stream);
            if (e != cudaSuccess) { FAIL(h, "mldivide: cooperative panel launch failed: %s", cudaGetErrorString(e)); }
        }
        ctr_base += (unsigned)G * (unsigned)nb;
        if (timing) cudaEventRecord(ev[1], h->stream);
        if (j0 > 0) k_lu_laswp<<<(j0 + 127) / 128, 128, 0, h->stream>>>(dA, lda, 0, j0, ipiv, j0, nb);
        const int rest = n - j0 - nb;
        if (rest > 0) {
            k_lu_laswp<<<(rest + 127) / 128, 128, 0, h->stream>>>(dA, lda, j0 + nb, n, ipiv, j0, nb);
            k_lu_trsm<<<(rest + 127) / 128, 128, 0, h->stream>>>(dA, lda, n, j0, nb);
            if (gemm_fma) k_lu_gemm_fma<<<dim3((rest + 63) / 64, (rest + 63) / 64), 256, 0, h->stream>>>(dA, lda, n, j0, nb);
            else k_lu_gemm<<<dim3((rest + LG_BM - 1) / LG_BM, (rest + LG_BN - 1) / LG_BN), 128, LG_SMEM, h->stream>>>(dA, lda, n, j0, nb);
        }
        h->launches += rest > 0 ? 5 : 2;
        if (timing) {
            cudaEventRecord(ev[2], h->stream);
            cudaEventSynchronize(ev[2]);
            float a = 0.f, b = 0.f;
            cudaEventElapsedTime(&a, ev[0], ev[1]);
            cudaEventElapsedTime(&b, ev[1], ev[2]);
            t_panel += a;
            t_trail += b;
        }
    }
    if (cudaGetLastError() != cudaSuccess) { FAIL(h, "mldivide: factorisation launch failed"); }
    int info = 0;
    CK(h, cudaMemcpyAsync(&info, dinfo, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    if (info == 0 && nrhs > 0) {
        // dgetrs: P applied to B, then L, then U
        if (timing) cudaEventRecord(ev[0], h->stream);
        for (int j0 = 0; j0 < n; j0 += LU_NB)
            k_lu_laswp<<<(nrhs + 127) / 128, 128, 0, h->stream>>>(dB, ldb, 0, nrhs, ipiv, j0, std::min(LU_NB, n - j0));
        const size_t smem = sizeof(double) * (size_t)n;
        if (smem > 200 * 1024) {
            FAIL(h, "mldivide: n = %d exceeds the solve kernel's shared-memory vector", n);
        }
        CK(h, cudaFuncSetAttribute(k_lu_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_lu_solve<<<nrhs, 1024, smem, h->stream>>>(dA, lda, n, dB, ldb);
        h->launches += (n + LU_NB - 1) / LU_NB + 1;
        if (timing) cudaEventRecord(ev[1], h->stream);
        CK(h, cudaGetLastError());
        CK(h, cudaStreamSynchronize(h->stream));
        if (timing) cudaEventElapsedTime(&t_solve, ev[0], ev[1]);
    }
    if (timing) {
        fprintf(stderr, "[sml lu] n=%d nrhs=%d panel %.3f ms, interchanges+trsm+gemm %.3f ms, solve %.3f ms\n", n, nrhs, t_panel, t_trail, t_solve);
        for (auto &e : ev) cudaEventDestroy(e);
    }
    *info_out = info;
    return 0;
}

}  // namespace

namespace {
// issue-rate probe of the FP64 tensor pipe: back-to-back DMMA m8n8k4 from registers, 16 independent accumulators per
// warp, 8 warps per CTA, 2 CTAs per SM -- the roof the Gram kernel is reported against, measured in the same run
__global__ void __launch_bounds__(256)
k_dmma_probe(double *out, int iters)
{
    double c[16][2];
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i][0] = c[i][1] = 0.0;
    const double a = threadIdx.x * 1e-3, b = threadIdx.x * 2e-3 + 1.0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
}  // namespace

extern "C" {

// FP64 tensor-core (DMMA) issue peak of this GPU in TFLOP/s, measured now (about 10 ms): best of 3 timed launches
int sml_dmma_probe(sml_engine *h, double *tflops)
{
    if (!h || !tflops) return -1;
    CK(h, cudaSetDevice(h->p.device));
    const int ctas = 2 * h->num_sms, iters = 4000;
    double *out = nullptr;
    CK(h, cudaMalloc(&out, sizeof(double) * ctas * 256));
    cudaEvent_t e0, e1;
    CK(h, cudaEventCreate(&e0));
    CK(h, cudaEventCreate(&e1));
    k_dmma_probe<<<ctas, 256, 0, h->stream>>>(out, 50);
    double best = 0.0;
    for (int rep = 0; rep < 3; ++rep) {
        CK(h, cudaEventRecord(e0, h->stream));
        k_dmma_probe<<<ctas, 256, 0, h->stream>>>(out, iters);
        CK(h, cudaEventRecord(e1, h->stream));
        CK(h, cudaEventSynchronize(e1));
        float ms = 0.f;
        CK(h, cudaEventElapsedTime(&ms, e0, e1));
        const double flops = (double)ctas * 8 * iters * 16 * (2.0 * 8 * 8 * 4);
        best = std::max(best, flops / (ms * 1e-3) / 1e12);
    }
    h->launches += 4;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    CK(h, cudaGetLastError());
    *tflops = best;
    return 0;
}

int sml_train_begin(sml_engine *h, int kind, const int32_t *regions, int nregions, int batch_size)
{
    if (check_ready(h, kind)) return -1;
    CK(h, cudaSetDevice(h->p.device));
    TrainState &T = h->train;
    if (T.active) FAIL(h, "sml_train_begin while a training wave is active (call sml_train_end)");
    if (nregions <= 0 || batch_size < 2) FAIL(h, "bad training wave (nregions %d, batch_size %d)", nregions, batch_size);
    KindState &K = h->kinds[kind];
    {
        // two entries for one region would make two Gram accumulations write the same W_out
        std::vector<int32_t> sorted(regions, regions + nregions);
        std::sort(sorted.begin(), sorted.end());
        for (int i = 1; i < nregions; ++i)
            if (sorted[i] == sorted[i - 1]) FAIL(h, "sml_train_begin: region %d is listed twice in the wave", sorted[i]);
    }
    T = TrainState{};
    T.kind = kind;
    T.batch_size = batch_size;
    // every early return below gives the arena back to the pool instead of leaking it (the next begin resets T)
    struct BeginGuard {
        sml_engine *h;
        bool armed = true;
        ~BeginGuard() { if (armed) train_release(h->train, &h->train_pool); }
    } guard{h};
    if (const char *s = getenv("SML_TRAIN_SLAB")) T.ks = std::max(16, atoi(s) / 16 * 16);
    {
        int ov = h->train_overlap;
        if (ov < 0) {
            const char *e = getenv("SML_TRAIN_OVERLAP");
            ov = (e && atoi(e) != 0) ? 1 : 0;   // off unless asked for: at full wave size it does not pay
        }
        T.overlap = ov != 0;
        const char *sg = getenv("SML_TRAIN_STATEGEN");
        T.stategen_route = !sg ? 0 : (std::string(sg) == "steps" ? 1 : (std::string(sg) == "kernel" ? 2 : (std::string(sg) == "ring" ? 3 : 0)));
        if (T.overlap && !h->train_gram_stream) {
            int lo = 0, hi = 0;   // lowest priority: the small state-generation launches must not queue behind Gram CTAs
            CK(h, cudaDeviceGetStreamPriorityRange(&lo, &hi));
            CK(h, cudaStreamCreateWithPriority(&h->train_gram_stream, cudaStreamNonBlocking, lo));
        }
    }
    std::vector<TrainRegionDev> devs;
    const bool tb_timing = getenv("SML_TRAIN_TIMING") && atoi(getenv("SML_TRAIN_TIMING")) != 0;   // host clock split on stderr
    auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double t_alloc = 0.0, t_memset = 0.0, t_map = 0.0;
    const double tb0 = now();
    // ---- one device allocation for the whole wave: its regions' blocks are carved out of an arena.  (1536 separate
    //      cudaMalloc calls for a wave of 192 cost 2.9-5.1 s in a process that already holds the model; one call of the
    //      same 68 GB costs milliseconds -- tools/malloc_probe.py, gpurun_out/train_hostwall*.log.)
    auto al256 = [](size_t b) { return (b + 255) / 256 * 256; };
    size_t arena_need = 0;
    for (int i = 0; i < nregions; ++i) {
        int li;
        if (local_of(h, kind, regions[i], &li)) return -1;
        const RegionDev &R = K.regs[li].dev;
        const size_t N = (size_t)R.n + R.S, ld = (N + R.P + 15) / 16 * 16;
        arena_need += al256(ld * ld * 8) + al256(slab_doubles((int)ld, T.ks * (T.overlap ? 2 : 1)) * 8) + 2 * al256((size_t)R.n * 8) +
                      al256(((N + CH_NB - 1) / CH_NB) * CH_LINV * 8) + al256(ld * 8) + al256(sizeof(int)) + al256(sizeof(int) * R.P);
    }
    {
        size_t got = 0;
        void *a = h->train_pool.take_at_least(arena_need, &got);
        if (!a) {
            cudaError_t e = cudaMalloc(&a, arena_need);
            if (e != cudaSuccess) {   // smaller arenas of earlier waves may be hoarding the memory: give them back and retry
                cudaGetLastError();
                h->train_pool.drop_all();
                e = cudaMalloc(&a, arena_need);
            }
            if (e != cudaSuccess) {
                h->err = std::string("training wave does not fit in HBM (cudaMalloc: ") + cudaGetErrorString(e) + "); use fewer regions per wave";
                return -1;
            }
            got = arena_need;
        }
        T.arena = a;
        T.arena_bytes = got;
    }
    size_t arena_off = 0;
    std::vector<std::pair<void *, const std::vector<int32_t> *>> tmap_copies;
    for (int i = 0; i < nregions; ++i) {
        int li;
        if (local_of(h, kind, regions[i], &li)) return -1;
        TrainRegionHost tr;
        tr.local = li;
        tr.region = regions[i];
        TrainRegionDev &d = tr.dev;
        d.R = K.regs[li].dev;
        d.region = regions[i];
        d.precip_off = -1;
        d.precip_len = 0;
        if (kind == SML_ATMO && K.regs[li].sizes.precip_off >= 0) {
            d.precip_off = K.regs[li].sizes.precip_off;
            d.precip_len = K.regs[li].sizes.logp_off > 0 ? K.regs[li].sizes.precip_off - K.regs[li].sizes.logp_off : 0;  // = ixy
        }
        const int N = d.R.n + d.R.S;
        d.ld = (N + d.R.P + 15) / 16 * 16;
        T.ld_max = std::max(T.ld_max, d.ld);
        T.n_max = std::max(T.n_max, d.R.n);
        T.D_max = std::max(T.D_max, d.R.D);
        T.hybrid = d.R.S > 0;
        auto alloc = [&](size_t bytes, void **p) -> int {
            if (arena_off + al256(bytes) > T.arena_bytes) {
                h->err = "internal: training arena overflow";
                return -1;
            }
            *p = static_cast<char *>(T.arena) + arena_off;
            arena_off += al256(bytes);
            return 0;
        };
        void *p = nullptr;
        bool bad = false;
        double tq = now();
        bad = bad || alloc((size_t)d.ld * d.ld * 8, &p); d.gram = (double *)p;
        t_alloc += now() - tq; tq = now();
        if (!bad) CK(h, cudaMemsetAsync(d.gram, 0, (size_t)d.ld * d.ld * 8, h->stream));
        t_memset += now() - tq; tq = now();
        d.ks_total = T.ks * (T.overlap ? 2 : 1);
        bad = bad || alloc(slab_doubles(d.ld, d.ks_total) * 8, &p); d.slab = (double *)p;
        bad = bad || alloc((size_t)d.R.n * 8, &p); d.xa = (double *)p;
        bad = bad || alloc((size_t)d.R.n * 8, &p); d.xb = (double *)p;
        bad = bad || alloc((size_t)((N + CH_NB - 1) / CH_NB) * CH_LINV * 8, &p); d.linv = (double *)p;
        bad = bad || alloc((size_t)d.ld * 8, &p); d.dsave = (double *)p;
        bad = bad || alloc(sizeof(int), &p); d.chol_info = (int *)p;
        // target rows (tile_full_input_to_target_data2d / _ocean_model), flattened at upload
        const std::vector<int32_t> &tmap = K.regs[li].target_map;
        if ((int)tmap.size() != d.R.P) FAIL(h, "internal: target map size");
        bad = bad || alloc(sizeof(int) * d.R.P, &p);
        t_alloc += now() - tq; tq = now();
        if (!bad) tmap_copies.push_back({p, &tmap});
        t_map += now() - tq;
        d.target_map = (const int *)p;
        T.regs.push_back(tr);
        if (bad) return -1;
        devs.push_back(d);
    }
    {
        // the wave's target maps go up with one pass of asynchronous copies and a single synchronisation
        const double tq = now();
        for (auto &c : tmap_copies)
            CK(h, cudaMemcpyAsync(c.first, c.second->data(), sizeof(int) * c.second->size(), cudaMemcpyHostToDevice, h->stream));
        CK(h, cudaStreamSynchronize(h->stream));
        t_map += now() - tq;
    }
    CK(h, cudaMalloc(&T.d_regs, sizeof(TrainRegionDev) * devs.size()));
    // lower-triangle tile list, row-major so that neighbouring CTAs share slab row blocks in L2
    std::vector<int2> tiles;
    const int nt = (T.ld_max + SY_BM - 1) / SY_BM;
    for (int ti = 0; ti < nt; ++ti)
        for (int tj = 0; tj <= ti; ++tj) tiles.push_back(make_int2(ti, tj));
    T.ntiles = (int)tiles.size();
    CK(h, cudaMalloc(&T.d_tiles, sizeof(int2) * tiles.size()));
    CK(h, cudaMemcpy(T.d_tiles, tiles.data(), sizeof(int2) * tiles.size(), cudaMemcpyHostToDevice));
    CK(h, cudaFuncSetAttribute(k_syrk_dmma<2, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SY_SMEM));
    CK(h, cudaFuncSetAttribute(k_syrk_dmma<4, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SY_SMEM));
    CK(h, cudaFuncSetAttribute(k_syrk_dmma<2, 4, SY_STAGES_DEEP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SY_SMEM_DEEP));
    CK(h, cudaFuncSetAttribute(k_chol_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SY_SMEM));
    CK(h, cudaFuncSetAttribute(k_chol_diag, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CH_DIAG_SMEM));
    CK(h, cudaStreamSynchronize(h->stream));
    if (tb_timing)
        fprintf(stderr, "[sml train_begin] %d regions: %.3f s (allocation %.3f, memset enqueue %.3f, target map + sync %.3f)\n", nregions,
                now() - tb0, t_alloc, t_memset, t_map);
    T.active = true;
    guard.armed = false;
    return 0;
}

// wait for everything the training wave has in flight (both streams) and fold the recorded event spans into the timers
static int train_sync(sml_engine *h)
{
    TrainState &T = h->train;
    CK(h, cudaStreamSynchronize(h->stream));
    if (h->train_gram_stream) CK(h, cudaStreamSynchronize(h->train_gram_stream));
    for (auto &sp : T.spans) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, sp.a, sp.b);
        (sp.what ? T.gram_ms : T.stategen_ms) += ms;
        cudaEventDestroy(sp.a);
        cudaEventDestroy(sp.b);
    }
    T.spans.clear();
    T.ev_gram[0] = T.ev_gram[1] = nullptr;
    return 0;
}

// state generation + Gram accumulation of one phase for the wave; inputs come either from the per-region series
// already placed in T.regs[i].dev.td / im, or from the device-resident global series gs.
// Overlap mode (sml_train_set_overlap(1) / SML_TRAIN_OVERLAP=1; off by default): the slab is double-buffered and the Gram of buffer b runs on its own low-priority stream
// while the state generation -- a chain of small latency-bound launches -- fills buffer b^1, also across phase
// boundaries (the call returns with the last Gram still in flight; train_sync at solve / gram_get / stats / end).
// The arithmetic and its order are those of the serial schedule, so the accumulators are bit-identical.
static int train_run_phase(sml_engine *h, int ncols, int discard_cols, const GlobalSeries &gs)
{
    TrainState &T = h->train;
    const int nw = (int)T.regs.size();
    std::vector<TrainRegionDev> devs(nw);
    for (int i = 0; i < nw; ++i) {
        TrainRegionDev &d = T.regs[i].dev;
        CK(h, cudaMemsetAsync(d.xa, 0, (size_t)d.R.n * 8, h->stream));  // x = 0 at the start of every phase (:1091)
        devs[i] = d;
    }
    if (T.uploaded.size() != devs.size() || memcmp(T.uploaded.data(), devs.data(), sizeof(TrainRegionDev) * nw) != 0) {
        if (train_sync(h)) return -1;   // a Gram still in flight reads d_regs
        CK(h, cudaMemcpyAsync(T.d_regs, devs.data(), sizeof(TrainRegionDev) * nw, cudaMemcpyHostToDevice, h->stream));
        CK(h, cudaStreamSynchronize(h->stream));
        T.uploaded = devs;
    }
    cudaStream_t S = h->stream;
    cudaStream_t GS = T.overlap ? h->train_gram_stream : h->stream;

    const dim3 ugrid((T.n_max + 255) / 256, nw);
    int parity = 0;
    cudaEvent_t d0 = nullptr, d1 = nullptr;
    if (discard_cols > 0) {
        cudaEventCreate(&d0); cudaEventCreate(&d1);
        CK(h, cudaEventRecord(d0, S));
    }
    // the time loop runs inside k_train_stategen (one CTA per region, state in shared memory); 256 threads fit beside
    // a resident Gram CTA (64 registers x 256 of the 17 K the Gram leaves per SM), 512 are faster when it runs alone
    const int sg_xs_cap = (T.n_max + 1) & ~1;
    const size_t sg_smem = sizeof(double) * ((size_t)sg_xs_cap + ((T.D_max + 1) & ~1));
    // one CTA per region: pays once the wave fills the machine (32 us per step at 192 regions against 42-48 us for the
    // per-step launches; 18 us against 7.5 us at 16 regions, where the per-step launches spread the rows over all SMs)
    bool persistent = sg_smem <= 90 * 1024 && (T.stategen_route == 2 || (T.stategen_route == 0 && nw >= SG_MIN_WAVE));
    // serial schedule, A/B: SML_TRAIN_SG_GROUP=6 -> 6 ELL slots per group and 384 threads (more loads in flight per thread)
    const bool sg_wide = !T.overlap && (getenv("SML_TRAIN_SG_GROUP") ? atoi(getenv("SML_TRAIN_SG_GROUP")) == 6 : SG_WIDE_DEFAULT);
    const int sg_cap = sg_wide ? 384 : SG_MAX_THREADS;
    const int sg_threads = getenv("SML_TRAIN_SG_THREADS") ? std::max(64, std::min(sg_cap, atoi(getenv("SML_TRAIN_SG_THREADS")) / 32 * 32))
                                                           : (T.overlap ? 256 : sg_cap);
    // k_train_stategen_ring: the time loop on the TMA ring of k_sync_persist, fed by the kind's tile-major adjacency pack.
    // Needs the whole SM (serial schedule only), compact W_in and 16-byte-multiple sizes in every region of the wave.
    SyncPlan ring_plan;
    bool ring = false;
    if (!T.overlap && (T.stategen_route == 3 || (T.stategen_route == 0 && nw >= SG_RING_MIN_WAVE))) {
        KindState &K = h->kinds[T.kind];
        ring = !K.any_dense;
        for (auto &r : T.regs) ring = ring && sync_region_ok(r.dev.R);
        if (ring) {
            const int ok = sync_plan_and_pack(h, K, ring_plan);
            if (ok < 0) return -1;
            ring = ok == 1 && T.D_max <= 2 * ring_plan.ngroups * ring_plan.tr;
        }
        if (ring) {
            bool changed = false;
            for (int i = 0; i < nw; ++i) {
                const long long off = K.sync_pack_off[T.regs[i].local];
                if (T.regs[i].dev.pack_off != off) { T.regs[i].dev.pack_off = off; changed = true; }
            }
            if (changed) {
                if (train_sync(h)) return -1;
                for (int i = 0; i < nw; ++i) devs[i] = T.regs[i].dev;
                CK(h, cudaMemcpyAsync(T.d_regs, devs.data(), sizeof(TrainRegionDev) * nw, cudaMemcpyHostToDevice, h->stream));
                CK(h, cudaStreamSynchronize(h->stream));
                T.uploaded = devs;
            }
        }
    }
    const unsigned char *ring_pack = ring ? h->kinds[T.kind].d_sync_pack : nullptr;
    auto launch_ring = [&](auto kern, int in_col0, int nsteps, int out_col0, int store_first, int s_first, int restart_period) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ring_plan.smem);
        kern<<<nw, ring_plan.ngroups * ring_plan.tr + 32, ring_plan.smem, S>>>(T.d_regs, in_col0, nsteps, out_col0, store_first, s_first,
                                                                              restart_period, ring_plan.xs_cap, ring_plan.us_cap,
                                                                              ring_plan.w_max, ring_plan.nst, ring_plan.tr,
                                                                              ring_plan.ngroups, ring_pack, gs);
    };
    auto launch_stategen = [&](int in_col0, int nsteps, int out_col0, int store_first, int s_first, int restart_period) {
        if (ring) {
            const bool generic = getenv("SML_SYNC_GENERIC") != nullptr;
            if (ring_plan.w_max <= 4 && !generic) launch_ring(k_train_stategen_ring<2>, in_col0, nsteps, out_col0, store_first, s_first, restart_period);
            else if (ring_plan.w_max <= 6 && !generic) launch_ring(k_train_stategen_ring<3>, in_col0, nsteps, out_col0, store_first, s_first, restart_period);
            else if (ring_plan.w_max <= 7 && !generic) launch_ring(k_train_stategen_ring<4>, in_col0, nsteps, out_col0, store_first, s_first, restart_period);
            else launch_ring(k_train_stategen_ring<0>, in_col0, nsteps, out_col0, store_first, s_first, restart_period);
            h->launches++;
            return;
        }
        if (sg_wide)
            k_train_stategen<6, 384><<<nw, sg_threads, sg_smem, S>>>(T.d_regs, in_col0, nsteps, out_col0, store_first, s_first,
                                                                      restart_period, sg_xs_cap, gs);
        else
            k_train_stategen<3, SG_MAX_THREADS><<<nw, sg_threads, sg_smem, S>>>(T.d_regs, in_col0, nsteps, out_col0, store_first, s_first,
                                                                                 restart_period, sg_xs_cap, gs);
        h->launches++;
    };
    if (ring) persistent = true;   // same launch structure: one launch per discard loop / slab
    h->train_last_route = ring ? 2 : persistent ? 1 : 0;
    if (persistent && !ring) {
        CK(h, cudaFuncSetAttribute(k_train_stategen<3, SG_MAX_THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sg_smem));
        CK(h, cudaFuncSetAttribute(k_train_stategen<6, 384>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sg_smem));
    }
    // discard loop (:1093-1106)
    if (persistent) {
        if (discard_cols > 0) launch_stategen(0, discard_cols, -1, 0, 0, 0);
    } else
    for (int i = 0; i < discard_cols; ++i) {
        k_train_update<<<ugrid, 256, 0, S>>>(T.d_regs, parity, i, -1, -1, gs);
        parity ^= 1;
        h->launches++;
    }
    if (discard_cols > 0) {
        CK(h, cudaEventRecord(d1, S));
        T.spans.push_back({d0, d1, 0});
    }
    const int TL = ncols - discard_cols;
    const int bs = T.batch_size;
    const int kept = (TL / bs) * bs;  // states 1..kept enter the Gram
    // state s (1-based) pairs with series column discard+s (1-based) = discard+s-1 (0-based); it is produced
    // from state s-1 with input column discard+s-1 (1-based) = discard+s-2 (0-based)
    int prev_base = 0;
    for (int s0 = 0; s0 < kept; s0 += T.ks) {
        const int nc = std::min(T.ks, kept - s0);
        const int kpad = (nc + SY_BK - 1) / SY_BK * SY_BK;
        const int buf = T.overlap ? (int)(T.slab_seq & 1u) : 0;
        const int base = buf * T.ks;   // first slab column of this buffer
        T.slab_seq++;
        cudaEvent_t e0, e1, g0, g1;
        cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&g0); cudaEventCreate(&g1);
        if (T.overlap && T.ev_gram[buf]) CK(h, cudaStreamWaitEvent(S, T.ev_gram[buf], 0));  // the Gram that read this buffer
        CK(h, cudaEventRecord(e0, S));
        if (persistent) {
            // columns base .. base+nc-1 hold states s0 .. s0+nc-1; state 0 is the stored x after the discard loop, state s > 0
            // is produced from input column discard+s-1; the ML-only paths restart every batch from the squared copy
            const int first = (s0 == 0) ? 1 : 0;
            launch_stategen(discard_cols + s0 + first - 1, nc - first, base + first, first, s0 + first, T.hybrid ? 0 : bs);
        } else
        for (int c = 0; c < nc; ++c) {
            const int s = s0 + c;  // 0-based state index
            if (s == 0) {
                k_train_store_state<<<ugrid, 256, 0, S>>>(T.d_regs, parity, base);
            } else {
                // ML-only paths restart every batch from the squared copy (SpMV operand only)
                // (at a slab boundary the previous slab was full, its last column is still intact)
                int gather = -1;
                if (!T.hybrid && (s % bs) == 0) gather = (c > 0) ? base + c - 1 : prev_base + T.ks - 1;
                k_train_update<<<ugrid, 256, 0, S>>>(T.d_regs, parity, discard_cols + s - 1, base + c, gather, gs);
                parity ^= 1;
            }
            h->launches++;
        }
        k_train_fill<<<dim3(kpad, nw), 128, 0, S>>>(T.d_regs, discard_cols + s0, nc, kpad, base, gs);
        h->launches++;
        CK(h, cudaEventRecord(e1, S));
        T.spans.push_back({e0, e1, 0});
        if (T.overlap) CK(h, cudaStreamWaitEvent(GS, e1, 0));
        CK(h, cudaEventRecord(g0, GS));
        // warp layout of the Gram kernel: A/B switch SML_SYRK_WARPS=16 -> 4 x 4 warps of 32 x 32, default 2 x 4 of 64 x 32
        static const bool w16 = getenv("SML_SYRK_WARPS") && atoi(getenv("SML_SYRK_WARPS")) == 16;
        const bool deep = getenv("SML_SYRK_STAGES") && atoi(getenv("SML_SYRK_STAGES")) == SY_STAGES_DEEP;   // A/B switch
        if (deep) k_syrk_dmma<2, 4, SY_STAGES_DEEP><<<dim3(T.ntiles, nw), SY_THREADS, SY_SMEM_DEEP, GS>>>(T.d_regs, T.d_tiles, kpad, base);
        else if (w16) k_syrk_dmma<4, 4><<<dim3(T.ntiles, nw), 17 * 32, SY_SMEM, GS>>>(T.d_regs, T.d_tiles, kpad, base);
        else k_syrk_dmma<2, 4><<<dim3(T.ntiles, nw), SY_THREADS, SY_SMEM, GS>>>(T.d_regs, T.d_tiles, kpad, base);
        h->launches++;
        CK(h, cudaEventRecord(g1, GS));
        T.spans.push_back({g0, g1, 1});
        T.ev_gram[buf] = g1;
        CK(h, cudaGetLastError());
        prev_base = base;
        for (auto &r : T.regs) {
            const double N = r.dev.R.n + r.dev.R.S, P = r.dev.R.P;
            T.gram_flops_useful += (N * (N + 1.0) + 2.0 * P * N) * nc;
        }
    }
    if (!T.overlap) return train_sync(h);
    // the host series of sml_train_feed may be reused by the caller as soon as the call returns, and the next call
    // overwrites the device copy: the state-generation stream has consumed both once it is idle
    CK(h, cudaStreamSynchronize(S));
    return 0;
}

int sml_train_feed(sml_engine *h, const double *td, const int64_t *td_off, const double *im, const int64_t *im_off,
                   int ncols, int discard_cols)
{
    if (!h) return -1;
    TrainState &T = h->train;
    if (!T.active) FAIL(h, "sml_train_feed without sml_train_begin");
    CK(h, cudaSetDevice(h->p.device));
    if (!td || !td_off) FAIL(h, "trainingdata is required");
    if (T.hybrid && (!im || !im_off)) FAIL(h, "hybrid training needs the imperfect-model series");
    if (discard_cols < 0 || ncols - discard_cols < T.batch_size) FAIL(h, "phase too short: %d columns, discard %d, batch %d", ncols, discard_cols, T.batch_size);
    const int nw = (int)T.regs.size();
    // upload this phase's series
    size_t td_total = 0, im_total = 0;
    for (auto &r : T.regs) {
        td_total += (size_t)r.dev.R.D * ncols;
        im_total += (size_t)r.dev.R.S * ncols;
    }
    if (td_total > T.series_td_cap) {
        cudaFree(T.d_series_td);
        T.d_series_td = nullptr;
        CK(h, cudaMalloc(&T.d_series_td, td_total * 8));
        T.series_td_cap = td_total;
    }
    if (im_total > T.series_im_cap) {
        cudaFree(T.d_series_im);
        T.d_series_im = nullptr;
        CK(h, cudaMalloc(&T.d_series_im, im_total * 8));
        T.series_im_cap = im_total;
    }
    size_t pt = 0, pi = 0;
    for (int i = 0; i < nw; ++i) {
        TrainRegionDev &d = T.regs[i].dev;
        const size_t tsz = (size_t)d.R.D * ncols, isz = (size_t)d.R.S * ncols;
        CK(h, cudaMemcpyAsync(T.d_series_td + pt, td + td_off[i], tsz * 8, cudaMemcpyHostToDevice, h->stream));
        d.td = T.d_series_td + pt;
        pt += tsz;
        if (isz) {
            CK(h, cudaMemcpyAsync(T.d_series_im + pi, im + im_off[i], isz * 8, cudaMemcpyHostToDevice, h->stream));
            d.im = T.d_series_im + pi;
            pi += isz;
        } else {
            d.im = nullptr;
        }
    }
    return train_run_phase(h, ncols, discard_cols, GlobalSeries{});
}

// The conditioned global series of the whole training period, uploaded ONCE and kept on the device across waves
// and phases (12 000 six-hourly states of the T30 grid are 30 GB: G 1.33 MB + F 1.22 MB per column).
int sml_train_global_series(sml_engine *h, const double *G_series, const double *F_series, int ncols_total)
{
    if (!h) return -1;  // needs no uploaded region: the statistics that parameterise the uploads come from this series
    if (!G_series || ncols_total < 1) FAIL(h, "sml_train_global_series: bad arguments");
    if (!h->p.ml_only && !F_series) FAIL(h, "hybrid training needs the forecast series F");
    CK(h, cudaSetDevice(h->p.device));
    TrainGlobal &TG = h->train_global;
    cudaFree(TG.d_G); cudaFree(TG.d_F);
    TG = TrainGlobal{};
    h->train_pool.drop_all();  // blocks kept from finished waves must not stand in the way
    if (cudaMalloc(&TG.d_G, sizeof(double) * (size_t)G_TOTAL * ncols_total) != cudaSuccess)
        FAIL(h, "the global series does not fit in HBM (%d columns)", ncols_total);
    CK(h, cudaMemcpy(TG.d_G, G_series, sizeof(double) * (size_t)G_TOTAL * ncols_total, cudaMemcpyHostToDevice));
    if (F_series) {
        if (cudaMalloc(&TG.d_F, sizeof(double) * (size_t)F_TOTAL * ncols_total) != cudaSuccess)
            FAIL(h, "the global forecast series does not fit in HBM (%d columns)", ncols_total);
        CK(h, cudaMemcpy(TG.d_F, F_series, sizeof(double) * (size_t)F_TOTAL * ncols_total, cudaMemcpyHostToDevice));
    }
    TG.ncols_total = ncols_total;
    return 0;
}

// give the device blocks kept from finished waves back to the allocator (they are reused by the next
// sml_train_begin otherwise; sml_destroy frees them in any case)
int sml_train_trim(sml_engine *h)
{
    if (!h) return -1;
    CK(h, cudaSetDevice(h->p.device));
    h->train_pool.drop_all();
    return 0;
}

// multiplicative Gaussian input noise of the state generation when feeding from the resident series
// (gaussian_noise_1d_function / _precip, src/mod_utilities.f90:1387-1464): u*(1 + noisemag*g), precip noised in linear
// space.  noisemag = 0 switches it off (parity runs).  The generator is counter-based (train.cuh), keyed by seed.
int sml_train_set_noise(sml_engine *h, double noisemag, unsigned long long seed, double precip_epsilon)
{
    if (!h) return -1;
    if (noisemag < 0.0 || !(precip_epsilon > 0.0)) FAIL(h, "sml_train_set_noise: bad arguments");
    h->train_global.noisemag = noisemag;
    h->train_global.seed = seed;
    h->train_global.precip_eps = precip_epsilon;
    return 0;
}

// what the state update would read for one region of the current wave at phase column col: the standardised input
// vector, the N(0,1) draws and the noised vector (test / inspection hook)
int sml_train_noise_sample(sml_engine *h, int region, int first_col, int stride, int col, double *clean, double *gauss,
                           double *noisy)
{
    if (!h) return -1;
    TrainState &T = h->train;
    TrainGlobal &TG = h->train_global;
    if (!T.active || !TG.d_G) FAIL(h, "sml_train_noise_sample needs an active wave and a resident series");
    int wi = -1;
    for (size_t i = 0; i < T.regs.size(); ++i)
        if (T.regs[i].region == region) wi = (int)i;
    if (wi < 0) FAIL(h, "region %d is not in the training wave", region);
    if (first_col < 0 || stride < 1 || col < 0 || first_col + (long long)stride * col >= TG.ncols_total) FAIL(h, "column out of range");
    CK(h, cudaSetDevice(h->p.device));
    const int D = T.regs[wi].dev.R.D;
    std::vector<TrainRegionDev> devs(T.regs.size());
    for (size_t i = 0; i < T.regs.size(); ++i) devs[i] = T.regs[i].dev;
    if (train_sync(h)) return -1;
    CK(h, cudaMemcpyAsync(T.d_regs, devs.data(), sizeof(TrainRegionDev) * devs.size(), cudaMemcpyHostToDevice, h->stream));
    T.uploaded = devs;
    double *d = nullptr;
    CK(h, cudaMalloc(&d, sizeof(double) * 3 * D));
    GlobalSeries gs;
    gs.G = TG.d_G; gs.F = TG.d_F; gs.g_len = G_TOTAL; gs.f_len = F_TOTAL; gs.first = first_col; gs.stride = stride;
    gs.noisemag = TG.noisemag; gs.precip_eps = TG.precip_eps; gs.seed = TG.seed;
    k_train_noise_sample<<<(D + 127) / 128, 128, 0, h->stream>>>(T.d_regs, wi, col, gs, d, d + D, d + 2 * D);
    h->launches++;
    CK(h, cudaGetLastError());
    CK(h, cudaMemcpyAsync(clean, d, sizeof(double) * D, cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaMemcpyAsync(gauss, d + D, sizeof(double) * D, cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaMemcpyAsync(noisy, d + 2 * D, sizeof(double) * D, cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    cudaFree(d);
    return 0;
}

// get_training_data's conditioning (unit conversion, floors, precip accumulation + log transform) applied in place to the
// resident RAW series; call once, before sml_conditioning_stats / sml_train_feed_global
int sml_condition_series(sml_engine *h, int period, double precip_epsilon)
{
    if (!h) return -1;
    TrainGlobal &TG = h->train_global;
    if (!TG.d_G) FAIL(h, "sml_train_global_series has not been called");
    if (TG.conditioned) FAIL(h, "the resident series has already been conditioned");
    if (period < 1 || !(precip_epsilon > 0.0)) FAIL(h, "sml_condition_series: bad period / precip_epsilon");
    CK(h, cudaSetDevice(h->p.device));
    k_condition_series<<<(unsigned)((G_TOTAL + 255) / 256), 256, 0, h->stream>>>(TG.d_G, G_TOTAL, TG.ncols_total, period,
                                                                                 precip_epsilon, G_W2D, G_PRECIP, G_SST, G_TISR,
                                                                                 h->p.precip_bool, h->p.slab_ocean_model_bool);
    h->launches++;
    CK(h, cudaGetLastError());
    CK(h, cudaStreamSynchronize(h->stream));
    TG.conditioned = true;
    return 0;
}

// mean/std of every local region from the resident series, window = columns first_col + stride*c, c < ncols.
// mean / std: [nloc][L] with L = 32 + logp + tisr (+ precip) (+ sst), the slot order of grid%mean
// (src/mod_reservoir.f90:414-436); sst_bool_input[nloc]: standardize_sst_data_3d's any_change (1 when there is no SST slot).
int sml_conditioning_stats(sml_engine *h, int first_col, int stride, int ncols, double *mean, double *std,
                           int32_t *sst_bool_input)
{
    if (!h) return -1;
    TrainGlobal &TG = h->train_global;
    if (!TG.d_G) FAIL(h, "sml_train_global_series has not been called");
    if (first_col < 0 || stride < 1 || ncols < 1 || first_col + (long long)stride * (ncols - 1) >= TG.ncols_total)
        FAIL(h, "window (first %d, stride %d, %d columns) runs past the %d resident columns", first_col, stride, ncols, TG.ncols_total);
    CK(h, cudaSetDevice(h->p.device));
    const int nloc = (int)h->local_ids.size();
    const bool precip = h->p.precip_bool, sst = h->p.slab_ocean_model_bool;
    const int L = NVAR * ZG + 2 + (precip ? 1 : 0) + (sst ? 1 : 0);
    std::vector<StatSlot> slots((size_t)nloc * L);
    std::vector<int> cells;
    for (int i = 0; i < nloc; ++i) {
        RegionGeom g = make_geom(h->tiling, h->local_ids[i], h->p.overlap);
        const int first_cell = (int)cells.size();
        for (int ly = 0; ly < g.iyc; ++ly)
            for (int lx = 0; lx < g.ixc; ++lx) cells.push_back((int)off2(g.gx[lx], g.iys - 1 + ly));
        const int nc = g.ixc * g.iyc;
        int l = 0;
        for (int v = 0; v < NVAR; ++v)
            for (int z = 0; z < ZG; ++z)
                slots[(size_t)i * L + l++] = StatSlot{first_cell, nc, G_W4D + v + (long long)NVAR * XG * YG * z, NVAR, 0};
        slots[(size_t)i * L + l++] = StatSlot{first_cell, nc, G_W2D, 1, 0};    // logp
        slots[(size_t)i * L + l++] = StatSlot{first_cell, nc, G_TISR, 1, 0};   // tisr
        if (precip) slots[(size_t)i * L + l++] = StatSlot{first_cell, nc, G_PRECIP, 1, 1};
        if (sst) slots[(size_t)i * L + l++] = StatSlot{first_cell, nc, G_SST, 1, 2};
    }
    StatSlot *d_slots = nullptr;
    int *d_cells = nullptr, *d_flag = nullptr;
    double *d_mean = nullptr, *d_std = nullptr;
    CK(h, cudaMalloc(&d_slots, sizeof(StatSlot) * slots.size()));
    CK(h, cudaMalloc(&d_cells, sizeof(int) * cells.size()));
    CK(h, cudaMalloc(&d_flag, sizeof(int) * nloc));
    CK(h, cudaMalloc(&d_mean, sizeof(double) * slots.size()));
    CK(h, cudaMalloc(&d_std, sizeof(double) * slots.size()));
    CK(h, cudaMemcpyAsync(d_slots, slots.data(), sizeof(StatSlot) * slots.size(), cudaMemcpyHostToDevice, h->stream));
    CK(h, cudaMemcpyAsync(d_cells, cells.data(), sizeof(int) * cells.size(), cudaMemcpyHostToDevice, h->stream));
    std::vector<int> ones(nloc, 1);
    CK(h, cudaMemcpyAsync(d_flag, ones.data(), sizeof(int) * nloc, cudaMemcpyHostToDevice, h->stream));
    k_cond_stats<<<dim3(L, nloc), 256, 0, h->stream>>>(TG.d_G, G_TOTAL, first_col, stride, ncols, d_slots, L, d_cells, d_mean,
                                                       d_std, d_flag);
    h->launches++;
    CK(h, cudaGetLastError());
    CK(h, cudaMemcpyAsync(mean, d_mean, sizeof(double) * slots.size(), cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaMemcpyAsync(std, d_std, sizeof(double) * slots.size(), cudaMemcpyDeviceToHost, h->stream));
    std::vector<int> flags(nloc, 1);
    CK(h, cudaMemcpyAsync(flags.data(), d_flag, sizeof(int) * nloc, cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    if (sst_bool_input) std::copy(flags.begin(), flags.end(), sst_bool_input);
    cudaFree(d_slots); cudaFree(d_cells); cudaFree(d_flag); cudaFree(d_mean); cudaFree(d_std);
    return L;
}

// rolling_average_over_a_period_2d(grid, period) on a caller-owned host array, in place: grid(i, t) at grid[ld*t + i],
// i < nrows (pass the address of the first averaged row: trainingdata(grid%atmo3d_start, 1), ld = reservoir_numinputs)
int sml_rolling_average_2d(sml_engine *h, double *grid, int ld, int nrows, int t_len, int period, int keep_small)
{
    if (!h) return -1;
    if (!grid || nrows < 1 || t_len < 1 || ld < nrows || period < 1) FAIL(h, "sml_rolling_average_2d: bad arguments");
    CK(h, cudaSetDevice(h->p.device));
    double *d = nullptr;
    const size_t cnt = (size_t)nrows * t_len;
    CK(h, cudaMalloc(&d, sizeof(double) * 2 * cnt));
    cudaError_t e = cudaMemcpy2DAsync(d, sizeof(double) * nrows, grid, sizeof(double) * ld, sizeof(double) * nrows, t_len,
                                      cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) {
        k_rolling_average<<<dim3(t_len, (nrows + 127) / 128), 128, 0, h->stream>>>(d, d + cnt, nrows, t_len, period, keep_small);
        h->launches++;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess)
        e = cudaMemcpy2DAsync(grid, sizeof(double) * ld, d + cnt, sizeof(double) * nrows, sizeof(double) * nrows, t_len,
                              cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(d);
    if (e != cudaSuccess) FAIL(h, "sml_rolling_average_2d: %s", cudaGetErrorString(e));
    return 0;
}

int sml_train_global_release(sml_engine *h)
{
    if (!h) return -1;
    cudaFree(h->train_global.d_G); cudaFree(h->train_global.d_F);
    h->train_global = TrainGlobal{};
    return 0;
}

// one phase from the resident global series: phase column c is global column first_col + stride*c
// (trainingdata(:, i::timestep), src/mod_reservoir.f90:289-301)
int sml_train_feed_global(sml_engine *h, int first_col, int stride, int ncols, int discard_cols)
{
    if (!h) return -1;
    TrainState &T = h->train;
    if (!T.active) FAIL(h, "sml_train_feed_global without sml_train_begin");
    if (T.kind != SML_ATMO) FAIL(h, "the global series feeds atmosphere reservoirs (the ocean inputs are time-averaged on the host)");
    TrainGlobal &TG = h->train_global;
    if (!TG.d_G) FAIL(h, "sml_train_global_series has not been called");
    if (T.hybrid && !TG.d_F) FAIL(h, "hybrid training needs the forecast series F");
    if (first_col < 0 || stride < 1 || ncols < 1 || first_col + (long long)stride * (ncols - 1) >= TG.ncols_total)
        FAIL(h, "phase (first %d, stride %d, %d columns) runs past the %d resident columns", first_col, stride, ncols, TG.ncols_total);
    if (discard_cols < 0 || ncols - discard_cols < T.batch_size) FAIL(h, "phase too short: %d columns, discard %d, batch %d", ncols, discard_cols, T.batch_size);
    CK(h, cudaSetDevice(h->p.device));
    GlobalSeries gs;
    gs.G = TG.d_G; gs.F = TG.d_F;
    gs.g_len = G_TOTAL; gs.f_len = F_TOTAL;
    gs.first = first_col; gs.stride = stride;
    gs.noisemag = TG.noisemag; gs.precip_eps = TG.precip_eps; gs.seed = TG.seed;
    return train_run_phase(h, ncols, discard_cols, gs);
}

int sml_train_solve(sml_engine *h, double beta_res, double beta_model, int using_prior, double prior_val,
                    int32_t *info_per_region)
{
    if (!h) return -1;
    TrainState &T = h->train;
    if (!T.active) FAIL(h, "sml_train_solve without sml_train_begin");
    CK(h, cudaSetDevice(h->p.device));
    KindState &K = h->kinds[T.kind];
    const int nw = (int)T.regs.size();
    // SML_SOLVER=lu forces dgesv-style LU with partial pivoting for every region (the fallback path)
    const char *solver_env = getenv("SML_SOLVER");
    const bool force_lu = solver_env && std::string(solver_env) == "lu";
    if (train_sync(h)) return -1;   // the last Gram may still be running on its own stream
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    CK(h, cudaEventRecord(e0, h->stream));

    // ---- ridge terms, prior, lower -> upper mirror (the upper triangle keeps A for the LU fallback)
    std::vector<int> init_info(nw, 0);
    int kb_max = 0, P_max = 0;
    for (int i = 0; i < nw; ++i) {
        TrainRegionDev &d = T.regs[i].dev;
        const int N = d.R.n + d.R.S, S = d.R.S, P = d.R.P, ld = d.ld;
        double add_model, add_res, prior_add = 0.0;
        int ml_first_n = -1;
        if (T.hybrid) {
            // fit_chunk_hybrid :1275-1291: beta**2 with a prior, plain beta without
            add_model = using_prior ? beta_model * beta_model : beta_model;
            add_res = using_prior ? beta_res * beta_res : beta_res;
            if (using_prior) prior_add = prior_val * beta_model * beta_model;
        } else {
            // fit_chunk_ml :1203-1205: + beta_res on the first n diagonals only
            add_model = 0.0;
            add_res = beta_res;
            ml_first_n = d.R.n;
        }
        const dim3 mg((ld + 31) / 32, (ld + 31) / 32);
        k_train_ridge_mirror<<<mg, dim3(32, 8), 0, h->stream>>>(d.gram, ld, N, S, P, add_model, add_res, prior_add, ml_first_n);
        h->launches++;
        // the tile kernels address operands in 16-byte units: N must be a multiple of 4 (every reservoir size the
        // reference derives is); anything else goes straight to LU
        init_info[i] = (force_lu || (N % 4) != 0) ? -1 : 0;
        CK(h, cudaMemcpyAsync(d.chol_info, &init_info[i], sizeof(int), cudaMemcpyHostToDevice, h->stream));
        kb_max = std::max(kb_max, (N + CH_NB - 1) / CH_NB);
        P_max = std::max(P_max, P);
    }
    CK(h, cudaGetLastError());

    // ---- batched left-looking Cholesky of the first N columns of every Gaug of the wave (chol.cuh)
    if (!force_lu) {
        k_chol_save_diag<<<dim3((T.ld_max + 255) / 256, nw), 256, 0, h->stream>>>(T.d_regs);
        h->launches++;
        const int nt = (T.ld_max + CH_NB - 1) / CH_NB;
        for (int k = 0; k < kb_max; ++k) {
            if (k > 0) {
                k_chol_gemm<<<dim3(nt - k, nw), SY_THREADS, SY_SMEM, h->stream>>>(T.d_regs, CH_UPDATE, k);
                h->launches++;
            }
            k_chol_diag<<<nw, 256, CH_DIAG_SMEM, h->stream>>>(T.d_regs, k);
            k_chol_gemm<<<dim3(nt - k, nw), SY_THREADS, SY_SMEM, h->stream>>>(T.d_regs, CH_TRSM, k);
            h->launches += 2;
        }
        CK(h, cudaGetLastError());
    }
    std::vector<int> chol(nw, -1);
    for (int i = 0; i < nw; ++i)
        CK(h, cudaMemcpyAsync(&chol[i], T.regs[i].dev.chol_info, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));

    // ---- regions on the Cholesky path: L^T into the upper triangle, then the back substitution W L = Z
    bool any_chol = false;
    for (int i = 0; i < nw; ++i) {
        if (chol[i] != 0) continue;
        any_chol = true;
        TrainRegionDev &d = T.regs[i].dev;
        const dim3 mg((d.ld + 31) / 32, (d.ld + 31) / 32);
        k_train_ridge_mirror<<<mg, dim3(32, 8), 0, h->stream>>>(d.gram, d.ld, d.R.n + d.R.S, d.R.S, d.R.P, 0.0, 0.0, 0.0, -1);
        h->launches++;
    }
    if (any_chol) {
        const int ptiles = (P_max + CH_NB - 1) / CH_NB;
        for (int k = kb_max - 1; k >= 0; --k) {
            k_chol_gemm<<<dim3(ptiles, nw), SY_THREADS, SY_SMEM, h->stream>>>(T.d_regs, CH_BACK_TRI, k);
            h->launches++;
            if (k > 0) {
                k_chol_gemm<<<dim3(ptiles * k, nw), SY_THREADS, SY_SMEM, h->stream>>>(T.d_regs, CH_BACK_UPDATE, k);
                h->launches++;
            }
        }
        CK(h, cudaGetLastError());
    }
    for (int i = 0; i < nw; ++i) {
        TrainRegionDev &d = T.regs[i].dev;
        const int N = d.R.n + d.R.S, P = d.R.P, ld = d.ld;
        double *wout = const_cast<double *>(K.regs[T.regs[i].local].dev.wout);
        int info = 0;
        if (chol[i] == 0) {
            k_chol_store_wout<<<N, 64, 0, h->stream>>>(T.d_regs, i, wout, d.R.ldw);
            h->launches++;
        } else {
            // not positive definite (ridge 0 / rank-deficient Gram) or not eligible: restore A and do what dgesv does
            if (chol[i] > 0) {
                const dim3 mg((ld + 31) / 32, (ld + 31) / 32);
                k_chol_restore<<<mg, dim3(32, 8), 0, h->stream>>>(d.gram, d.dsave, ld);
                h->launches++;
            }
            if (device_dgesv(h, d.gram, ld, d.gram + (size_t)ld * N, ld, N, P, &info)) return -1;
            if (info == 0) {
                // reservoir%wout = transpose(b_trans) (:1313)
                k_train_store_wout<<<(N + 127) / 128, 128, 0, h->stream>>>(d.gram + (size_t)ld * N, ld, N, P, wout, d.R.ldw);
                h->launches++;
            }
            // 'something went wrong with dgesv' is print-and-continue in the reference (src/mod_linalg.f90:147-150):
            // W_out is left untouched and info is reported
        }
        if (info_per_region) info_per_region[i] = info;
        T.solved_by_cholesky += chol[i] == 0 ? 1 : 0;
    }
    CK(h, cudaEventRecord(e1, h->stream));
    CK(h, cudaEventSynchronize(e1));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    T.solve_ms += ms;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    CK(h, cudaGetLastError());
    return 0;
}

// how many regions of the current wave the Cholesky path solved (the rest went through LU)
int sml_train_solver_stats(sml_engine *h, int *by_cholesky)
{
    if (!h) return -1;
    *by_cholesky = h->train.solved_by_cholesky;
    return 0;
}

int sml_train_gram_get(sml_engine *h, int region, double *sxs, double *sxt)
{
    if (!h) return -1;
    TrainState &T = h->train;
    if (!T.active) FAIL(h, "no active training wave");
    if (train_sync(h)) return -1;
    for (auto &r : T.regs) {
        if (r.region != region) continue;
        const TrainRegionDev &d = r.dev;
        const int N = d.R.n + d.R.S, P = d.R.P, ld = d.ld;
        std::vector<double> g((size_t)ld * ld);
        CK(h, cudaMemcpyAsync(g.data(), d.gram, g.size() * 8, cudaMemcpyDeviceToHost, h->stream));
        CK(h, cudaStreamSynchronize(h->stream));
        if (sxs)
            for (int j = 0; j < N; ++j)
                for (int i = 0; i < N; ++i) sxs[(size_t)N * j + i] = (i >= j) ? g[(size_t)ld * j + i] : g[(size_t)ld * i + j];
        if (sxt)
            for (int j = 0; j < N; ++j)
                for (int p = 0; p < P; ++p) sxt[(size_t)P * j + p] = g[(size_t)ld * j + N + p];
        return 0;
    }
    FAIL(h, "region %d is not in the training wave", region);
}

int sml_train_stategen_route(const sml_engine *h) { return h ? h->train_last_route : -1; }

int sml_train_stats(sml_engine *h, double *gram_flops_useful, double *gram_ms, double *stategen_ms, double *solve_ms)
{
    if (!h) return -1;
    if (h->train.active && train_sync(h)) return -1;
    *gram_flops_useful = h->train.gram_flops_useful;
    *gram_ms = h->train.gram_ms;
    *stategen_ms = h->train.stategen_ms;
    *solve_ms = h->train.solve_ms;
    return 0;
}

// 1: state generation overlaps the Gram of the previous slab, 0: serial schedule (default); takes effect at the next
// sml_train_begin
int sml_train_set_overlap(sml_engine *h, int on)
{
    if (!h) return -1;
    h->train_overlap = on ? 1 : 0;
    return 0;
}

int sml_train_end(sml_engine *h)
{
    if (!h) return -1;
    CK(h, cudaSetDevice(h->p.device));
    if (train_sync(h)) return -1;
    train_release(h->train, &h->train_pool);  // the next wave reuses the blocks
    return 0;
}

int sml_mldivide(sml_engine *h, double *A, int lda, double *B, int ldb, int n, int nrhs)
{
    if (!h) return -1;
    if (n <= 0 || nrhs <= 0 || lda < n || ldb < n) FAIL(h, "mldivide: bad dimensions");
    CK(h, cudaSetDevice(h->p.device));
    double *dA = nullptr, *dB = nullptr;
    CK(h, cudaMalloc(&dA, (size_t)n * n * 8));
    CK(h, cudaMalloc(&dB, (size_t)n * nrhs * 8));
    CK(h, cudaMemcpy2DAsync(dA, (size_t)n * 8, A, (size_t)lda * 8, (size_t)n * 8, n, cudaMemcpyHostToDevice, h->stream));
    CK(h, cudaMemcpy2DAsync(dB, (size_t)n * 8, B, (size_t)ldb * 8, (size_t)n * 8, nrhs, cudaMemcpyHostToDevice, h->stream));
    int info = 0;
    int rc = device_dgesv(h, dA, n, dB, n, n, nrhs, &info);
    if (rc == 0) {
        // like dgesv, A returns its LU factors and B the solution (only meaningful when info == 0)
        cudaMemcpy2DAsync(A, (size_t)lda * 8, dA, (size_t)n * 8, (size_t)n * 8, n, cudaMemcpyDeviceToHost, h->stream);
        if (info == 0)
            cudaMemcpy2DAsync(B, (size_t)ldb * 8, dB, (size_t)n * 8, (size_t)n * 8, nrhs, cudaMemcpyDeviceToHost, h->stream);
        cudaStreamSynchronize(h->stream);
    }
    cudaFree(dA); cudaFree(dB);
    return rc ? -1 : info;
}

}  // extern "C"
