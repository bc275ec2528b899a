// engine.cu -- host side of the B200 reservoir engine behind include/speedyml_engine.h.
//
// Data layout in HBM (per rank = per GPU):
//   per region, allocated at upload (256 B aligned by cudaMalloc):
//     ELL adjacency  ell_col[W][n] int32 + ell_val[W][n] f64, slot-major (thread-per-row loads coalesce)
//     W_in compact   winc[n] f64 + wcol[n] int32   (dense n*D copy only when the one-per-row test fails)
//     W_out          ldw*(S+n) f64 column-major, ldw = P rounded up to even (TMA tiles stay 16 B aligned)
//     mean/std       L+1 f64 each (slot L: SST feedback constants)
//     gather maps    fb_src/fb_ms[D], lm_src/lm_ms[S], out_ms[P] int32
//   pooled at finalize: x ping/pong, feedback, local_model, outvec slab [nloc][P], readout partials,
//     the global buffers G = [w4d|w2d|precip|sst|tisr], F = [f4d|f2d] and the scatter table out_dst[R][P].
// No CPU fallback anywhere: every entry point that computes launches the kernels in kernels.cuh.
#include "../../include/speedyml_engine.h"
#include "kernels.cuh"
#include <chrono>
#include "resdomain.hpp"
#include "train.cuh"
#include "chol.cuh"
#include "lu.cuh"
#include "genres.cuh"
#include "ncfile.hpp"

#include <algorithm>
#include <cmath>
#include <functional>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

using namespace sml;

namespace {

std::string g_create_err;

constexpr int STAGES = 4;
// the forecast block the root hands to every rank: [forecast_4d | forecast_2d | tisr grid | run_speedy + padding]
constexpr long long FCST_TISR = F_TOTAL, FCST_FLAG = F_TOTAL + XG * YG, FCST_DOUBLES = FCST_FLAG + 8;
constexpr int STAGE_BYTES_TARGET = 20 * 1088;  // 20 columns of a 136-row W_out
constexpr int PROFILE_RING = 4096;

struct HostRegion {
    bool uploaded = false;
    int region = -1;
    RegionDev dev{};
    std::vector<void *> allocs;
    RegionSizes sizes{};
    OceanSizes osizes{};
    bool sst_in = false;
    std::vector<int32_t> target_map;   // rows of the input vector that form the training target
    const int *d_sst_src = nullptr;    // ocean: offsets of the halo'd SST tile in G
    double sst_mean = 0.0, sst_std = 1.0;
    int gen_index = -1;                // >= 0: the adjacency and W_in are generated on the device (K.gen[gen_index])
    int k = 0;
};

struct KindState {
    bool any = false;
    int P = -1, ldw = 0;
    std::vector<HostRegion> regs;
    RegionDev *d_regs = nullptr;
    StepItem *d_items = nullptr;
    StepItem *d_items_split = nullptr;  // same chunks without the S local_model columns (overlapped step)
    std::vector<StepItem> items;
    int nitems = 0;
    double *d_x[2] = {nullptr, nullptr};
    int cur = 0;
    double *d_fb = nullptr, *d_lm = nullptr, *d_out = nullptr, *d_partials = nullptr, *d_temp = nullptr;
    long long *d_fb_offs = nullptr;
    long long x_total = 0, fb_total = 0, lm_total = 0;
    int stage_cols = 0, stage_bytes = 0, xs_cap = 0, n_max = 0, chunk_rows = 0, D_max = 0;
    int *d_order = nullptr;  // launch order of the items: largest first
    // persistent step kernel (k_step_persist): fixed row blocks ("parts") per region, items = runs of whole blocks,
    // slots = statically balanced contiguous runs of items, one CTA each
    bool persist = false;
    bool l2_keep = false;   // evict-last adjacency / evict-first W_out in the persistent step kernel
    StepSeg *d_segs = nullptr, *d_segs_split = nullptr;
    int2 *d_slots = nullptr;
    int nslots = 0, nparts = 0, part_rows = 0, p_stage_cols = 0, p_ldp = 0, p_xs_cap = 0, p_cpi = 0, p_stages = 0;
    size_t p_smem_bytes = 0;
    int *d_one_region = nullptr;  // region index for one region's synchronize
    int one_region = -1;
    int *d_sync_list = nullptr;   // regions of one k_sync_persist launch
    size_t sync_list_cap = 0;
    // tile-major copy of the adjacency + compact W_in for k_sync_persist (k_sync_pack), built lazily
    unsigned char *d_sync_pack = nullptr;
    long long *d_sync_pack_off = nullptr;
    size_t sync_pack_cap = 0;
    int pack_tr = 0, pack_w = 0;
    bool pack_valid = false;
    std::vector<long long> sync_pack_off;   // host copy of d_sync_pack_off
    size_t smem_bytes = 0;
    bool any_dense = false;
    int64_t alg_bytes = 0, alg_bytes_update = 0;
    // reservoirs constructed on the device (sml_region_generate): descriptors, run by sml_finalize
    std::vector<GenDesc> gen;
    GenDesc *d_gen = nullptr;
    int *d_gen_of_local = nullptr;
    // synchronize input staging
    double *d_in = nullptr;
    size_t d_in_cap = 0;
    long long *d_in_offs = nullptr;
};

}  // namespace

// device memory of the uploaded weights: bump-allocated from a few large chunks instead of ~14 cudaMalloc calls per
// region (16 k allocations for the whole model made set-up allocation-bound; the training path learnt the same lesson)
struct DevArena {
    static constexpr size_t CHUNK = 512ull << 20;
    std::vector<void *> chunks;
    char *cur = nullptr;
    size_t left = 0;
    size_t total = 0;
    cudaError_t alloc(size_t bytes, void **out)
    {
        bytes = (std::max<size_t>(bytes, 1) + 255) / 256 * 256;
        if (bytes > left) {
            const size_t want = std::max(bytes, CHUNK);
            void *c = nullptr;
            cudaError_t e = cudaMalloc(&c, want);
            if (e != cudaSuccess) return e;
            chunks.push_back(c);
            total += want;
            if (bytes >= CHUNK) {   // an oversized array gets a chunk of its own; the open chunk stays open
                *out = c;
                return cudaSuccess;
            }
            cur = static_cast<char *>(c);
            left = want;
        }
        *out = cur;
        cur += bytes;
        left -= bytes;
        return cudaSuccess;
    }
    void release()
    {
        for (void *c : chunks) cudaFree(c);
        chunks.clear();
        cur = nullptr;
        left = 0;
        total = 0;
    }
};

struct sml_engine {
    sml_params p{};
    std::string err;
    DevArena arena;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    Tiling tiling;
    std::vector<int32_t> local_ids;
    std::unordered_map<int, int> local_index;
    KindState kinds[2];
    bool finalized = false;
    int P_atmo = 0;
    // exchange
    double *d_G = nullptr, *d_F = nullptr, *d_gathered = nullptr;
    double *d_base_sst = nullptr, *d_mask = nullptr, *d_prescribed = nullptr;
    int *d_out_dst = nullptr;
    int *d_cell_region = nullptr, *d_cell_slot = nullptr;
    double *d_ocean_gathered = nullptr;
    OceanFb *d_ocean_fb = nullptr;     // one entry per local ocean reservoir
    double *d_ocean_ring = nullptr;    // averaged_atmo_input_vec of every local ocean reservoir
    size_t ocean_ring_doubles = 0;
    int n_ocean_fb = 0, ocean_slots = 27, P_ocean = 0;
    double *h_pin_G = nullptr, *h_pin_F = nullptr;
    bool sst_static_set = false, sst_prescribed_set = false;
    // profiling
    std::vector<cudaEvent_t> ev;  // ring of (start, step done, finish done) triples
    int ev_used = 0;              // triples recorded since the last read
    bool profile = false;
    struct PhaseTimer {
        std::vector<cudaEvent_t> ev;  // (start, stop) pairs
        int used = 0;
    } pt_pack, pt_unpack;
    double sync_ms_sum = 0.0;   // update-only launches of sml_synchronize while profiling (tools/sweep.py)
    long long sync_steps = 0;
    int64_t launches = 0;
    int train_last_route = -1;   // state-generation route of the last training phase: 0 steps, 1 kernel, 2 ring
    TrainState train;
    TrainGlobal train_global;
    TrainPool train_pool;
    int num_sms = 148;
    cudaStream_t train_gram_stream = nullptr;   // the Gram kernels' stream when state generation overlaps them
    int train_overlap = -1;                     // -1: SML_TRAIN_OVERLAP decides (default off), 0 / 1: sml_train_set_overlap
    // overlapped step: the state update and the W_out[:, S:]*x~ partials of the NEXT
    // predict run while the host model works on this step's grids; the S model columns are added when its
    // forecast arrives
    bool overlap = false;
    bool ahead_pending = false;  // ML part of the next predict launched, model part still missing
    bool ahead_done = false;     // next predict complete: the next sml_predict(ATMO) only consumes it
    bool tisr_fresh = false;     // sml_set_tisr has been called for the step being exchanged
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_pack = nullptr, ev_d2h = nullptr, ev_h2d = nullptr;
    double *h_pin_tisr = nullptr;
    // outvec_component_contribs (src/mod_reservoir.f90:1458-1461): v_p and v_ml of every local atmosphere region
    bool contribs = false;
    double *d_vp = nullptr, *d_vml = nullptr;
    // peer exchange over NVLink (kernels.cuh PeerTable): this rank's exchange block
    //   [2][R*P] atmosphere outvecs | [2][R*P_ocean] ocean outvecs | forecast landing [F | tisr | run_speedy] | flags
    void *d_xchg = nullptr;
    size_t xchg_bytes = 0;
    long long xo_atmo = 0, xo_ocean = 0, xo_fcst = 0, xo_flags = 0;   // byte offsets of the sections
    PeerTable peers{};                 // world == 0/1: not attached
    std::vector<void *> peer_mapped;   // cudaIpcOpenMemHandle results to close
    unsigned long long peer_seq = 0;   // atmosphere predicts published so far
    unsigned long long ocean_seq = 0;  // ocean slabs published so far
    unsigned long long fcst_seq = 0;   // forecasts pushed (root) / expected (others) so far
    bool ocean_publish_pending = true; // the seeded ocean outvecs have not been pushed yet (start_prediction_slab)
    unsigned int *d_done = nullptr;    // [4] completion counters of the pushing kernels
    int *d_peer_err = nullptr;
    // failure detection: sticky status bits of the assembled grid (k_pack_grids) and the root's run_speedy flag
    int *d_status = nullptr;
    int *h_status = nullptr;           // pinned
    int run_speedy = 1;                // what the root passed to sml_set_run_speedy (travels with the forecast)
    double *h_pin_flag = nullptr;      // pinned read-back of the run_speedy slot on the other ranks
    double setup_upload_s = 0.0;       // host wall clock spent inside sml_region_upload (bench: setup split)
    unsigned xch_calls = 0;            // fused exchange launches since the arrival counter was last zeroed
    // ranks other than the root never wait for the host model; a ring of blocking events keeps them at most a few
    // steps ahead of the device so that they neither fill the launch queue nor burn a core spinning in it
    cudaEvent_t ev_ahead[4] = {nullptr, nullptr, nullptr, nullptr};
    unsigned long long consumer_steps = 0;
};

#define FAIL(h, ...)                                      \
    do {                                                  \
        char _b[512];                                     \
        snprintf(_b, sizeof(_b), __VA_ARGS__);            \
        (h)->err = _b;                                    \
        return -1;                                        \
    } while (0)

#define CK(h, call)                                                                              \
    do {                                                                                         \
        cudaError_t _e = (call);                                                                 \
        if (_e != cudaSuccess) FAIL(h, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

namespace {

// one array of a region's weights: space from the rank's arena, copy enqueued on the engine stream.  The caller
// synchronises ONCE per region (sml_region_upload) before the host staging vectors go out of scope.
template <typename T>
int dev_upload(sml_engine *h, HostRegion *hr, const T *src, size_t count, const T **out)
{
    (void)hr;
    void *d = nullptr;
    CK(h, h->arena.alloc(std::max<size_t>(count, 1) * sizeof(T), &d));
    if (count) CK(h, cudaMemcpyAsync(d, src, count * sizeof(T), cudaMemcpyHostToDevice, h->stream));
    *out = static_cast<const T *>(d);
    return 0;
}

// Launch with programmatic stream serialization (PDL): the kernel may be scheduled while its predecessor in the stream is
// still draining; it calls pdl_wait() before touching dependent memory (kernels.cuh).  SML_PDL=0 launches the ordinary way.
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args)
{
    static const bool pdl = !(getenv("SML_PDL") && atoi(getenv("SML_PDL")) == 0);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

int check_ready(sml_engine *h, int kind)
{
    if (!h) return -1;
    if (kind != SML_ATMO && kind != SML_OCEAN) FAIL(h, "bad kind %d", kind);
    if (!h->finalized) FAIL(h, "sml_finalize has not been called");
    if (!h->kinds[kind].any) FAIL(h, "no reservoirs of kind %d uploaded", kind);
    return 0;
}

int local_of(sml_engine *h, int kind, int region, int *li)
{
    auto it = h->local_index.find(region);
    if (it == h->local_index.end()) FAIL(h, "region %d is not owned by rank %d", region, h->p.irank);
    if (!h->kinds[kind].regs[it->second].uploaded) FAIL(h, "region %d kind %d has no reservoir", region, kind);
    *li = it->second;
    return 0;
}

}  // namespace

extern "C" {

const char *sml_last_error(const sml_engine *h) { return h ? h->err.c_str() : g_create_err.c_str(); }

int sml_create(sml_engine **out, const sml_params *p)
{
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        g_create_err = std::string("no CUDA device: ") + cudaGetErrorString(e) + " (this engine has no CPU fallback)";
        return -1;
    }
    if (p->device < 0 || p->device >= ndev) {
        g_create_err = "device ordinal out of range";
        return -1;
    }
    if (p->numprocs < 1 || p->irank < 0 || p->irank >= p->numprocs) {
        g_create_err = "bad irank/numprocs";
        return -1;
    }
    Tiling t = make_tiling(p->number_of_regions);
    if (!t.ok || t.ntx * t.nty != p->number_of_regions) {
        g_create_err = "number_of_regions does not tile the 96x48 grid (domaindecomposition would not exit)";
        return -1;
    }
    if (t.fx + 2 * p->overlap > XG || p->overlap < 0) {
        g_create_err = "overlap too large for the tiling";
        return -1;
    }
    if (p->numprocs > 1 && p->number_of_regions % p->numprocs != 0) {
        g_create_err = "multi-rank runs need number_of_regions divisible by numprocs (contiguous slabs)";
        return -1;
    }
    if ((e = cudaSetDevice(p->device)) != cudaSuccess) {
        g_create_err = std::string("cudaSetDevice: ") + cudaGetErrorString(e);
        return -1;
    }
    sml_engine *h = new sml_engine();
    h->p = *p;
    h->tiling = t;
    if ((e = cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking)) != cudaSuccess) {
        g_create_err = std::string("cudaStreamCreate: ") + cudaGetErrorString(e);
        delete h;
        return -1;
    }
    h->stream = h->own_stream;
    if (cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, p->device) != cudaSuccess || h->num_sms < 1) h->num_sms = 148;
    h->local_ids = regions_of_rank(p->irank, p->numprocs, p->number_of_regions);
    for (size_t i = 0; i < h->local_ids.size(); ++i) h->local_index[h->local_ids[i]] = (int)i;
    for (int k = 0; k < 2; ++k) h->kinds[k].regs.resize(h->local_ids.size());
    *out = h;
    return 0;
}

// which update-only kernel sml_synchronize launches when SML_UPDATE_KERNEL is not set (profiles/round1_summary.md section 3)
constexpr bool SML_UPDATE_SX_DEFAULT = true;
// round 2: the TMA-fed ELL ring kernel (k_update_ring) when the region's state vector and a ring fit in shared memory
constexpr const char *SML_UPDATE_DEFAULT = "ring";
// evict-first loads for the ELL / W_in streams of the fused step kernels (keeps the gathered state vector in L1)
constexpr int SML_ELL_STREAM_DEFAULT = 0;
// which fused step kernel sml_predict launches when SML_STEP_KERNEL is not set: k_step_persist (persistent, statically
// balanced slots) or the classic one-CTA-per-item k_step
constexpr bool SML_STEP_PERSIST_DEFAULT = true;

static void free_kind(KindState &K)
{
    for (auto &r : K.regs)
        for (void *p : r.allocs) cudaFree(p);
    cudaFree(K.d_regs); cudaFree(K.d_items); cudaFree(K.d_items_split); cudaFree(K.d_order); cudaFree(K.d_one_region); cudaFree(K.d_x[0]); cudaFree(K.d_x[1]); cudaFree(K.d_fb);
    cudaFree(K.d_lm); cudaFree(K.d_out); cudaFree(K.d_partials); cudaFree(K.d_temp); cudaFree(K.d_fb_offs);
    cudaFree(K.d_in); cudaFree(K.d_in_offs); cudaFree(K.d_sync_list); cudaFree(K.d_sync_pack); cudaFree(K.d_sync_pack_off);
    cudaFree(K.d_segs); cudaFree(K.d_segs_split); cudaFree(K.d_slots);
    cudaFree(K.d_gen); cudaFree(K.d_gen_of_local);
}

int sml_destroy(sml_engine *h)
{
    if (!h) return 0;
    cudaSetDevice(h->p.device);
    cudaDeviceSynchronize();
    train_release(h->train);
    h->train_pool.drop_all();
    cudaFree(h->train_global.d_G); cudaFree(h->train_global.d_F);
    for (int k = 0; k < 2; ++k) free_kind(h->kinds[k]);
    h->arena.release();
    if (h->d_xchg) h->d_F = nullptr;   // the forecast landing buffer lives inside the exchange block
    cudaFree(h->d_status); cudaFreeHost(h->h_status); cudaFreeHost(h->h_pin_flag);
    cudaFree(h->d_G); cudaFree(h->d_F); cudaFree(h->d_gathered); cudaFree(h->d_base_sst); cudaFree(h->d_mask);
    cudaFree(h->d_prescribed); cudaFree(h->d_out_dst); cudaFree(h->d_cell_region); cudaFree(h->d_cell_slot);
    cudaFree(h->d_ocean_gathered); cudaFree(h->d_ocean_fb); cudaFree(h->d_ocean_ring);
    for (void *m : h->peer_mapped) cudaIpcCloseMemHandle(m);
    cudaFree(h->d_xchg); cudaFree(h->d_done); cudaFree(h->d_peer_err); cudaFree(h->d_vp); cudaFree(h->d_vml);
    cudaFreeHost(h->h_pin_G); cudaFreeHost(h->h_pin_F); cudaFreeHost(h->h_pin_tisr);
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    if (h->train_gram_stream) cudaStreamDestroy(h->train_gram_stream);
    if (h->ev_pack) cudaEventDestroy(h->ev_pack);
    if (h->ev_d2h) cudaEventDestroy(h->ev_d2h);
    if (h->ev_h2d) cudaEventDestroy(h->ev_h2d);
    for (cudaEvent_t e : h->ev_ahead)
        if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : h->ev) cudaEventDestroy(e);
    for (cudaEvent_t e : h->pt_pack.ev) cudaEventDestroy(e);
    for (cudaEvent_t e : h->pt_unpack.ev) cudaEventDestroy(e);
    cudaStreamDestroy(h->own_stream);
    delete h;
    return 0;
}

int sml_set_stream(sml_engine *h, void *s)
{
    h->stream = s ? (cudaStream_t)s : h->own_stream;
    return 0;
}
int sml_synchronize_stream(sml_engine *h)
{
    CK(h, cudaStreamSynchronize(h->stream));
    return 0;
}
int sml_num_local_regions(const sml_engine *h) { return (int)h->local_ids.size(); }
int sml_local_region_ids(const sml_engine *h, int32_t *ids)
{
    std::copy(h->local_ids.begin(), h->local_ids.end(), ids);
    return 0;
}

/* ------------------------------------------------------------------ index arithmetic */
int sml_domaindecomposition(int numregions, int *fx, int *fy)
{
    Tiling t = make_tiling(numregions);
    if (!t.ok) return -1;
    *fx = t.fx; *fy = t.fy;
    return 0;
}
int sml_getxyresextent(int R, int region, int *xs, int *xe, int *ys, int *ye, int *xc, int *yc)
{
    Tiling t = make_tiling(R);
    if (!t.ok) return -1;
    RegionGeom g = make_geom(t, region, 0);
    *xs = g.xs; *xe = g.xe; *ys = g.ys; *ye = g.ye; *xc = t.fx; *yc = t.fy;
    return 0;
}
int sml_getoverlapindices(int R, int region, int ov, int *ixs, int *ixe, int *iys, int *iye, int *ixc, int *iyc,
                          int *pole, int *periodic)
{
    Tiling t = make_tiling(R);
    if (!t.ok) return -1;
    RegionGeom g = make_geom(t, region, ov);
    *ixs = g.ixs; *ixe = g.ixe; *iys = g.iys; *iye = g.iye; *ixc = g.ixc; *iyc = g.iyc;
    *pole = g.pole; *periodic = g.periodic;
    return 0;
}
int sml_get_trainingdataindices(int R, int region, int ov, int *xs, int *xe, int *ys, int *ye)
{
    Tiling t = make_tiling(R);
    if (!t.ok) return -1;
    RegionGeom g = make_geom(t, region, ov);
    *xs = g.tdx0 + 1; *xe = g.tdx0 + t.fx; *ys = g.tdy0 + 1; *ye = g.tdy0 + t.fy;
    return 0;
}
int sml_processor_decomposition(int irank, int numprocs, int nregions, int32_t *idx, int *count)
{
    if (numprocs < 1 || irank < 0 || irank >= numprocs) return -1;
    auto v = regions_of_rank(irank, numprocs, nregions);
    std::copy(v.begin(), v.end(), idx);
    *count = (int)v.size();
    return 0;
}
int sml_region_dims(int R, int region, int ov, int m, double deg, int precip, int sst_bool, int sst_in, int ml_only,
                    int *n, int *k, int *D, int *P, int *S, int *L)
{
    Tiling t = make_tiling(R);
    if (!t.ok) return -1;
    RegionGeom g = make_geom(t, region, ov);
    RegionSizes s = make_sizes(t, g, m, deg, precip, sst_bool, sst_in, ml_only);
    *n = s.n; *k = s.k; *D = s.D; *P = s.P; *S = s.S; *L = s.L;
    return 0;
}
int sml_ocean_region_dims(int R, int region, int ov, int m, double deg, int *n, int *k, int *D, int *P, int *A)
{
    Tiling t = make_tiling(R);
    if (!t.ok) return -1;
    RegionGeom g = make_geom(t, region, ov);
    OceanSizes s = make_ocean_sizes(t, g, m, deg);
    *n = s.n; *k = s.k; *D = s.D; *P = s.P; *A = s.A;
    return 0;
}
int sml_ocean_region_maps(int R, int region, int ov, int32_t *sst_map, int32_t *target_map, int *atmo_slice0)
{
    Tiling t = make_tiling(R);
    if (!t.ok) return -1;
    RegionGeom g = make_geom(t, region, ov);
    OceanSizes s = make_ocean_sizes(t, g, 4000, 6.0);
    OceanMaps m = make_ocean_maps(t, g, s);
    if (sst_map) std::copy(m.sst_src.begin(), m.sst_src.end(), sst_map);
    if (target_map) std::copy(m.target_map.begin(), m.target_map.end(), target_map);
    if (atmo_slice0) *atmo_slice0 = s.atmo_slice0;
    return 0;
}
int sml_region_maps(int R, int region, int ov, int precip, int sst_in, int32_t *input_map, int32_t *input_ms,
                    int32_t *output_map, int32_t *output_ms, int32_t *model_map, int32_t *model_ms,
                    int32_t *target_map)
{
    Tiling t = make_tiling(R);
    if (!t.ok) return -1;
    RegionGeom g = make_geom(t, region, ov);
    RegionSizes s = make_sizes(t, g, 6000, 6.0, precip, true, sst_in, false);
    RegionMaps m = make_maps(t, g, s, precip, sst_in);
    if (input_map) std::copy(m.input_map.begin(), m.input_map.end(), input_map);
    if (input_ms) std::copy(m.input_ms.begin(), m.input_ms.end(), input_ms);
    if (output_map) std::copy(m.output_map.begin(), m.output_map.end(), output_map);
    if (output_ms) std::copy(m.output_ms.begin(), m.output_ms.end(), output_ms);
    if (model_map) std::copy(m.model_map.begin(), m.model_map.end(), model_map);
    if (model_ms) std::copy(m.model_ms.begin(), m.model_ms.end(), model_ms);
    if (target_map) std::copy(m.target_map.begin(), m.target_map.end(), target_map);
    return 0;
}
/* vertical localisation: the res_domain.f90 functions for num_vert_levels in {1, 2, 4, 8} and the sizes / maps of one
 * vertical slab (the engine itself steps num_vert_levels == 1, the reference's configuration) */
int sml_get_z_res_extent(int nvl, int level, int *zs, int *ze, int *zchunk)
{
    VertGeom v = make_vert(nvl, level, 0);
    if (!v.ok) return -1;
    *zs = v.zs; *ze = v.ze; *zchunk = v.zc;
    return 0;
}
int sml_getoverlapindices_vert(int nvl, int level, int vov, int *izs, int *ize, int *izc, int *top, int *bottom)
{
    VertGeom v = make_vert(nvl, level, vov);
    if (!v.ok) return -1;
    *izs = v.izs; *ize = v.ize; *izc = v.izc; *top = v.top; *bottom = v.bottom;
    return 0;
}
int sml_get_trainingdataindices_vert(int nvl, int level, int vov, int *zs, int *ze)
{
    VertGeom v = make_vert(nvl, level, vov);
    if (!v.ok) return -1;
    *zs = v.tdzs; *ze = v.tdze;
    return 0;
}
int sml_region_dims_vert(int R, int region, int ov, int nvl, int level, int vov, int m, double deg, int precip, int sst_bool,
                         int sst_in, int ml_only, int *n, int *k, int *D, int *P, int *S, int *L)
{
    Tiling t = make_tiling(R);
    VertGeom v = make_vert(nvl, level, vov);
    if (!t.ok || !v.ok) return -1;
    RegionGeom g = make_geom(t, region, ov);
    RegionSizes s = make_sizes_vert(t, g, v, m, deg, precip, sst_bool, sst_in, ml_only);
    *n = s.n; *k = s.k; *D = s.D; *P = s.P; *S = s.S; *L = s.L;
    return 0;
}
int sml_region_maps_vert(int R, int region, int ov, int nvl, int level, int vov, int precip, int sst_in,
                         int32_t *input_map, int32_t *input_ms, int32_t *output_map, int32_t *output_ms,
                         int32_t *model_map, int32_t *model_ms, int32_t *target_map)
{
    Tiling t = make_tiling(R);
    VertGeom v = make_vert(nvl, level, vov);
    if (!t.ok || !v.ok) return -1;
    RegionGeom g = make_geom(t, region, ov);
    RegionSizes s = make_sizes_vert(t, g, v, 6000, 6.0, precip, true, sst_in, false);
    RegionMaps m = make_maps_vert(t, g, v, s, precip, sst_in);
    if (input_map) std::copy(m.input_map.begin(), m.input_map.end(), input_map);
    if (input_ms) std::copy(m.input_ms.begin(), m.input_ms.end(), input_ms);
    if (output_map) std::copy(m.output_map.begin(), m.output_map.end(), output_map);
    if (output_ms) std::copy(m.output_ms.begin(), m.output_ms.end(), output_ms);
    if (model_map) std::copy(m.model_map.begin(), m.model_map.end(), model_map);
    if (model_ms) std::copy(m.model_ms.begin(), m.model_ms.end(), model_ms);
    if (target_map) std::copy(m.target_map.begin(), m.target_map.end(), target_map);
    return 0;
}
int sml_global_layout(int64_t off[5], int64_t *g_total, int64_t *f_total)
{
    off[0] = G_W4D; off[1] = G_W2D; off[2] = G_PRECIP; off[3] = G_SST; off[4] = G_TISR;
    *g_total = G_TOTAL;
    *f_total = F_TOTAL;
    return 0;
}

/* ------------------------------------------------------------------ upload */
static int region_install(sml_engine *h, const sml_region_weights *w, bool generate, unsigned long long seed, double sigma);

int sml_region_upload(sml_engine *h, const sml_region_weights *w) { return region_install(h, w, false, 0, 0.0); }

// gen_res's makesparse and train_reservoir's W_in build on the device (src/mod_linalg.f90:180-218,
// src/mod_reservoir.f90:262-283): the region is installed like an uploaded one, but its adjacency structure, its
// (unscaled) values and its W_in are drawn by the engine's counter-based generator when sml_finalize runs.
int sml_region_generate(sml_engine *h, const sml_region_weights *w, unsigned long long seed, double sigma)
{
    return region_install(h, w, true, seed, sigma);
}

static int region_install(sml_engine *h, const sml_region_weights *w, bool generate, unsigned long long seed, double sigma)
{
    if (!h || !w) return -1;
    if (h->finalized) FAIL(h, "upload after sml_finalize");
    if (w->kind != SML_ATMO && w->kind != SML_OCEAN) FAIL(h, "bad kind %d", w->kind);
    CK(h, cudaSetDevice(h->p.device));
    const auto t_up0 = std::chrono::steady_clock::now();
    auto it = h->local_index.find(w->region);
    if (it == h->local_index.end()) FAIL(h, "region %d is not owned by rank %d", w->region, h->p.irank);
    KindState &K = h->kinds[w->kind];
    HostRegion &hr = K.regs[it->second];
    if (hr.uploaded) FAIL(h, "region %d kind %d uploaded twice", w->region, w->kind);
    const int n = w->n, k = w->k, D = w->D, P = w->P, S = w->S, L = w->L;
    if (n <= 0 || k < 0 || D <= 0 || P <= 0 || S < 0 || L <= 0) FAIL(h, "bad dimensions for region %d", w->region);
    if (!w->mean || !w->std) FAIL(h, "null mean / std for region %d", w->region);
    if (!generate) {
        if (!w->rows || !w->cols || !w->vals) FAIL(h, "null weight array for region %d", w->region);
        if ((w->win_dense != nullptr) == (w->win_compact != nullptr))
            FAIL(h, "exactly one of win_dense / win_compact must be given (region %d)", w->region);
        if (w->win_compact && !w->win_col) FAIL(h, "win_compact needs win_col (region %d)", w->region);
    } else {
        if (k <= 0) FAIL(h, "region %d: k must be positive to generate an adjacency", w->region);
        if (n % D != 0) FAIL(h, "region %d: n = %d is not a multiple of reservoir_numinputs %d (allocate_res_new gives n = nodes_per_input * D)", w->region, n, D);
    }

    RegionGeom g = make_geom(h->tiling, w->region, h->p.overlap);
    RegionMaps maps;
    if (w->kind == SML_ATMO) {
        // the vector layouts are fixed by the tiling: D, P, S, L must be what allocate_res_new derives
        RegionSizes s = make_sizes(h->tiling, g, 6000, 6.0, h->p.precip_bool, h->p.slab_ocean_model_bool,
                                   w->sst_bool_input, h->p.ml_only);
        if (s.D != D || s.P != P || s.S != S || s.L != L)
            FAIL(h, "region %d: D/P/S/L = %d/%d/%d/%d but the tiling gives %d/%d/%d/%d", w->region, D, P, S, L, s.D,
                 s.P, s.S, s.L);
        hr.sizes = s;
        hr.sst_in = w->sst_bool_input && h->p.slab_ocean_model_bool;
        maps = make_maps(h->tiling, g, s, h->p.precip_bool, hr.sst_in);
        hr.target_map = maps.target_map;
    } else {
        // res%reservoir_special: sizes of initialize_slab_ocean_model, always ML-only
        if (!h->p.slab_ocean_model_bool) FAIL(h, "ocean reservoir uploaded but slab_ocean_model_bool is off");
        OceanSizes s = make_ocean_sizes(h->tiling, g, 4000, 6.0);
        // S = 0: predict_slab_ml (ml_only_ocean, the reference's setting); S = P: the hybrid predict_slab, whose feature
        // vector carries chunk_size_prediction "model" entries = its own previous standardised prediction
        if (s.D != D || s.P != P || (S != 0 && S != P))
            FAIL(h, "ocean region %d: D/P/S = %d/%d/%d but the tiling gives %d/%d/(0 or %d)", w->region, D, P, S, s.D, s.P, s.P);
        if (n % D != 0) FAIL(h, "ocean region %d: n = %d is not a multiple of reservoir_numinputs %d", w->region, n, D);
        if (w->sst_std == 0.0) FAIL(h, "ocean region %d: sst_std must be the SST slot of grid_special%%std", w->region);
        hr.osizes = s;
        OceanMaps om = make_ocean_maps(h->tiling, g, s);
        hr.target_map = om.target_map;
        hr.sst_mean = w->sst_mean;
        hr.sst_std = w->sst_std;
        if (dev_upload(h, &hr, om.sst_src.data(), om.sst_src.size(), &hr.d_sst_src)) return -1;
        // predict_slab_ml un-standardises every output with the SST constants (:1354): slot L
        maps.output_ms.assign(P, L);
    }
    const int ldw = P + (P & 1);
    if (ldw / 2 > NCONS) FAIL(h, "chunk_size_prediction %d too large for the readout kernel", P);
    if (K.P >= 0 && K.P != P) FAIL(h, "all reservoirs of a kind must share chunk_size_prediction");
    K.P = P;
    K.ldw = ldw;

    RegionDev &d = hr.dev;
    d = RegionDev{};
    std::vector<double> wp;   // lives until the synchronisation at the end of this function
    hr.k = k;
    if (generate) {
        // ---- device construction: reserve the arrays, describe the job; sml_finalize launches it for all regions at once
        const int W = k > n ? k / n + (k % n != 0 ? 1 : 0) : 1;   // one ELL slot per shuffle round (each row at most once per round)
        d.n = n; d.D = D; d.P = P; d.S = S; d.ldw = ldw; d.ell_w = W; d.L = L; d.leak = w->leakage;
        d.win_mode = 0;
        GenDesc g{};
        g.local = it->second; g.region = w->region; g.n = n; g.k = k; g.D = D; g.W = W; g.seed = seed; g.sigma = sigma;
        void *p = nullptr;
        CK(h, h->arena.alloc(sizeof(int) * (size_t)W * n, &p)); g.ell_col = (int *)p;
        CK(h, cudaMemsetAsync(p, 0, sizeof(int) * (size_t)W * n, h->stream));
        CK(h, h->arena.alloc(sizeof(double) * (size_t)W * n, &p)); g.ell_val = (double *)p;
        CK(h, cudaMemsetAsync(p, 0, sizeof(double) * (size_t)W * n, h->stream));
        CK(h, h->arena.alloc(sizeof(double) * (size_t)n, &p)); g.winc = (double *)p;
        CK(h, h->arena.alloc(sizeof(int) * (size_t)n, &p)); g.wcol = (int *)p;
        CK(h, h->arena.alloc(sizeof(int) * (size_t)k, &p)); g.coo_rows = (int *)p;
        CK(h, h->arena.alloc(sizeof(int) * (size_t)k, &p)); g.coo_cols = (int *)p;
        CK(h, h->arena.alloc(sizeof(double) * (size_t)k, &p)); g.coo_vals = (double *)p;
        d.ell_col = g.ell_col; d.ell_val = g.ell_val; d.winc = g.winc; d.wcol = g.wcol;
        hr.gen_index = (int)K.gen.size();
        K.gen.push_back(g);
    } else {
    // --- adjacency: COO (1-based, duplicates kept, entry order preserved per row) -> ELL, slot-major
    std::vector<int> cnt(n, 0);
    for (int e = 0; e < k; ++e) {
        const int r = w->rows[e], c = w->cols[e];
        if (r < 1 || r > n || c < 1 || c > n)  // mkl_sparse_d_create_coo would fail and mklsparse stop
            FAIL(h, "region %d: COO entry %d (%d,%d) outside 1..%d", w->region, e, r, c, n);
        cnt[r - 1]++;
    }
    int W = 1;
    for (int r = 0; r < n; ++r) W = std::max(W, cnt[r]);
    std::vector<int> ecol((size_t)W * n, 0);
    std::vector<double> eval((size_t)W * n, 0.0);
    std::fill(cnt.begin(), cnt.end(), 0);
    for (int e = 0; e < k; ++e) {
        const int r = w->rows[e] - 1;
        const int s = cnt[r]++;
        ecol[(size_t)s * n + r] = w->cols[e] - 1;
        eval[(size_t)s * n + r] = w->vals[e];
    }
    d.n = n; d.D = D; d.P = P; d.S = S; d.ldw = ldw; d.ell_w = W; d.L = L; d.leak = w->leakage;
    if (dev_upload(h, &hr, ecol.data(), ecol.size(), &d.ell_col)) return -1;
    if (dev_upload(h, &hr, eval.data(), eval.size(), &d.ell_val)) return -1;

    // --- W_in: accept the dense n x D matrix, verify the one-non-zero-per-row structure, compress
    std::vector<double> winc(n, 0.0);
    std::vector<int> wcol(n, 0);
    d.win_mode = 0;
    if (w->win_compact) {
        for (int j = 0; j < n; ++j) {
            if (w->win_col[j] < 0 || w->win_col[j] >= D) FAIL(h, "region %d: win_col[%d] out of range", w->region, j);
            winc[j] = w->win_compact[j];
            wcol[j] = w->win_col[j];
        }
    } else {
        std::vector<int> nnz(n, 0);
        for (int i = 0; i < D; ++i) {
            const double *col = w->win_dense + (size_t)i * n;
            for (int j = 0; j < n; ++j)
                if (col[j] != 0.0) {
                    nnz[j]++;
                    winc[j] = col[j];
                    wcol[j] = i;
                }
        }
        for (int j = 0; j < n; ++j)
            if (nnz[j] > 1) { d.win_mode = 1; break; }
        if (d.win_mode == 1) {
            if (dev_upload(h, &hr, w->win_dense, (size_t)n * D, &d.win_dense)) return -1;
            K.any_dense = true;
        }
    }
    if (dev_upload(h, &hr, winc.data(), winc.size(), &d.winc)) return -1;
    if (dev_upload(h, &hr, wcol.data(), wcol.size(), &d.wcol)) return -1;
    }   // uploaded (not generated) adjacency and W_in

    // --- W_out, padded to an even leading dimension
    const size_t N = (size_t)n + S;
    if (ldw == P && w->wout) {
        if (dev_upload(h, &hr, w->wout, (size_t)P * N, &d.wout)) return -1;
    } else {
        wp.assign((size_t)ldw * N, 0.0);
        if (w->wout)
            for (size_t j = 0; j < N; ++j) std::memcpy(&wp[j * ldw], w->wout + j * P, sizeof(double) * P);
        if (dev_upload(h, &hr, wp.data(), wp.size(), &d.wout)) return -1;
    }

    // --- mean/std (+ slot L for the SST feedback)
    std::vector<double> mean(w->mean, w->mean + L), sd(w->std, w->std + L);
    mean.push_back(w->sst_mean);
    sd.push_back(w->sst_std != 0.0 ? w->sst_std : 1.0);
    if (dev_upload(h, &hr, mean.data(), mean.size(), &d.mean)) return -1;
    if (dev_upload(h, &hr, sd.data(), sd.size(), &d.std)) return -1;

    // --- maps
    if (dev_upload(h, &hr, maps.input_map.data(), maps.input_map.size(), &d.fb_src)) return -1;
    if (dev_upload(h, &hr, maps.input_ms.data(), maps.input_ms.size(), &d.fb_ms)) return -1;
    if (dev_upload(h, &hr, maps.model_map.data(), maps.model_map.size(), &d.lm_src)) return -1;
    if (dev_upload(h, &hr, maps.model_ms.data(), maps.model_ms.size(), &d.lm_ms)) return -1;
    if (dev_upload(h, &hr, maps.output_ms.data(), maps.output_ms.size(), &d.out_ms)) return -1;
    CK(h, cudaStreamSynchronize(h->stream));   // one synchronisation per region: the engine owns copies from here on

    if (w->kind == SML_OCEAN && S > 0) {
        d.lm_self = 1;
        d.leak = 1.0;   // predict_slab has no leak term: x = tanh(A x + W_in u) (src/mod_slab_ocean_reservoir.f90:1294)
    }
    hr.region = w->region;
    hr.uploaded = true;
    K.any = true;
    h->setup_upload_s += std::chrono::duration<double>(std::chrono::steady_clock::now() - t_up0).count();
    return 0;
}

/* ------------------------------------------------------------------ weight container (read_trained_res) */
// the seven variables of write_trained_res' NetCDF-classic file, widened to FP64 (src/mod_io.f90:2938-2983)
namespace {
struct TrainedRes {
    int n = 0, D = 0, P = 0, N = 0, k = 0, L = 0;
    std::vector<double> win, wout, vals, mean, std;
    std::vector<int32_t> rows, cols;
};
void load_trained_res(const char *path, TrainedRes &t, bool dims_only)
{
    NcClassicFile f(path);
    const auto &vw = f.var("win");   // Fortran win(n, D) -> file shape (D, n)
    const auto &vo = f.var("wout");  // Fortran wout(P, n+S) -> file shape (n+S, P)
    if (vw.shape.size() != 2 || vo.shape.size() != 2) throw std::runtime_error("win / wout must be two-dimensional");
    t.D = (int)vw.shape[0]; t.n = (int)vw.shape[1];
    t.N = (int)vo.shape[0]; t.P = (int)vo.shape[1];
    t.k = (int)f.var("rows").count();
    t.L = (int)f.var("mean").count();
    if (f.var("cols").count() != t.k || f.var("vals").count() != t.k || f.var("std").count() != t.L)
        throw std::runtime_error("rows/cols/vals or mean/std lengths disagree");
    if (t.N < t.n) throw std::runtime_error("wout has fewer columns than the reservoir has nodes");
    if (dims_only) return;
    t.win = f.read_real("win");
    t.wout = f.read_real("wout");
    t.rows = f.read_int("rows");
    t.cols = f.read_int("cols");
    t.vals = f.read_real("vals");
    t.mean = f.read_real("mean");
    t.std = f.read_real("std");
}
}  // namespace

extern "C" int sml_trained_res_dims(const char *path, int *n, int *k, int *D, int *P, int *S, int *L)
{
    try {
        TrainedRes t;
        load_trained_res(path, t, true);
        *n = t.n; *k = t.k; *D = t.D; *P = t.P; *S = t.N - t.n; *L = t.L;
        return 0;
    } catch (const std::exception &e) {
        g_create_err = e.what();
        return -1;
    }
}

// read_trained_res + mklsparse for one region straight from its file (trained_reservoir_prediction,
// src/mod_reservoir.f90:1783-1886); sst_mean/sst_std default to the last mean/std slot (the SST slot)
extern "C" int sml_region_upload_file(sml_engine *h, const char *path, int region, int kind, int sst_bool_input,
                                      double leakage)
{
    if (!h || !path) return -1;
    TrainedRes t;
    try {
        load_trained_res(path, t, false);
    } catch (const std::exception &e) {
        FAIL(h, "%s: %s", path, e.what());
    }
    sml_region_weights w{};
    w.region = region;
    w.kind = kind;
    w.n = t.n; w.k = t.k; w.D = t.D; w.P = t.P; w.S = t.N - t.n; w.L = t.L;
    w.sst_bool_input = sst_bool_input;
    w.leakage = leakage;
    if (t.L <= 0 || t.mean.empty() || t.std.empty()) FAIL(h, "%s: mean / std are empty", path);
    w.sst_mean = t.mean.back();
    w.sst_std = t.std.back();
    w.rows = t.rows.data(); w.cols = t.cols.data(); w.vals = t.vals.data();
    w.win_dense = t.win.data();
    w.wout = t.wout.data();
    w.mean = t.mean.data(); w.std = t.std.data();
    return sml_region_upload(h, &w);
}

// ---- plan of the persistent step kernel (kernels.cuh k_step_persist) -----------------------------------------
// Every region is cut into fixed row blocks of part_rows rows -- independent of the rank count, so the summation
// structure of a region's readout (one partial per block, summed in block order by k_readout_finish) is the same for
// any sharding.  The blocks of all local regions, in region order, are dealt to nslots = 2 x SMs slots in contiguous
// runs of equal cost (cost = W_out columns streamed); inside a slot, consecutive blocks of one region form an item
// of at most max_item_rows rows (the x~ tile in shared memory).
static int build_persistent_plan(sml_engine *h, KindState &K, int S_max)
{
    const char *sk = getenv("SML_STEP_KERNEL");
    K.persist = sk ? std::string(sk) == "persist" : SML_STEP_PERSIST_DEFAULT;
    const int HP = K.ldw / 2;
    // column groups per warp: the largest power of two <= 32 whose row pairs still fit the 17 consumer warps
    int cpi = 32;
    while (cpi > 1 && (HP + (32 / cpi) - 1) / (32 / cpi) > NCONS_WARPS) cpi >>= 1;
    if ((HP + (32 / cpi) - 1) / (32 / cpi) > NCONS_WARPS) {
        K.persist = false;   // chunk_size_prediction too large for the shuffle layout: classic kernel
        return 0;
    }
    K.p_cpi = cpi;
    // column stride in shared memory: the 8 lanes of one 128-bit load phase must hit 32 distinct banks.
    // Lane -> (row pair r = lane % rpw, column group c = lane / rpw); word address = c * 2*ldp + 4 * r.  ldp = ldw (no
    // padding, ONE bulk copy per stage) whenever that is already conflict-free -- e.g. ldw = 136 with 4 row pairs per warp.
    const int rpw = 32 / cpi;
    int ldp = K.ldw;
    if (getenv("SML_PERSIST_PAD")) ldp += 2 * atoi(getenv("SML_PERSIST_PAD"));   // A/B: force per-column copies
    for (;; ldp += 2) {
        bool ok = true;
        unsigned seen = 0;
        for (int l = 0; l < 8 && ok; ++l) {
            const int r = l % rpw, c = l / rpw;
            const int bank4 = ((c * 2 * ldp + 4 * r) % 32) / 4;   // which group of 4 banks the 16-byte access starts in
            if (seen & (1u << bank4)) ok = false;
            seen |= 1u << bank4;
        }
        if (ok || ldp > K.ldw + 64) break;
    }
    K.p_ldp = ldp;
    const int unit = std::max(16, 2 * cpi);   // tile and block sizes are multiples of this (fixed column -> group map)
    int stage_cols = std::max(unit, (STAGE_BYTES_TARGET / (ldp * 8)) / unit * unit);
    if (const char *e = getenv("SML_PERSIST_STAGE_COLS")) stage_cols = std::max(unit, atoi(e) / unit * unit);
    K.p_stage_cols = stage_cols;
    // fixed row blocks: the granularity of the slot balance and of the partial sums.  Items are runs of whole blocks of
    // at most item_rows rows -- ONE sweep of the 544 consumer threads, so an item's update phase is a single latency
    // chain (~3 us), about what the ring's prefetch covers; longer items stall the CTA's W_out stream for the excess
    // (measured: 1152-row items, three sweeps, ran at 0.82 of the roof where the classic kernel reaches 1.00)
    int part_rows = std::max(stage_cols, 272 / stage_cols * stage_cols);
    if (const char *e = getenv("SML_PART_ROWS")) part_rows = std::max(stage_cols, atoi(e) / stage_cols * stage_cols);
    K.part_rows = part_rows;
    // two sweeps per item (1088 rows) at 4 ring stages measured best at both ends: 0.985 of the roof at 1152 regions per
    // GPU (one-sweep items: 0.959; the classic kernel: 0.98-0.99) and 0.90 at 144 (classic: 0.85)
    int item_rows = 2 * NCONS;
    if (const char *e = getenv("SML_ITEM_ROWS")) item_rows = atoi(e);
    const int max_item_rows = std::max(part_rows, item_rows / part_rows * part_rows);
    K.p_stages = 4;   // 3 / 4 / 5 stages at 144 regions: 0.889 / 0.903 / 0.866 (a deeper ring takes L1 from the x gathers)
    if (const char *e = getenv("SML_PERSIST_STAGES")) K.p_stages = std::max(2, std::min(8, atoi(e)));
    const bool stagger = !(getenv("SML_PERSIST_STAGGER") && atoi(getenv("SML_PERSIST_STAGGER")) == 0);
    struct Part { int reg, row0, nrows, part; long long cost; };
    std::vector<Part> parts;
    const int nloc = (int)K.regs.size();
    long long total_cost = 0;
    for (int i = 0; i < nloc; ++i) {
        HostRegion &hr = K.regs[i];
        if (!hr.uploaded) continue;
        RegionDev &d = hr.dev;
        d.part0 = (int)parts.size();
        for (int r0 = 0; r0 < d.n; r0 += part_rows) {
            Part pt{i, r0, std::min(part_rows, d.n - r0), (int)parts.size(), 0};
            pt.cost = pt.nrows + (r0 == 0 ? d.S : 0);
            total_cost += pt.cost;
            parts.push_back(pt);
        }
        d.nparts = (int)parts.size() - d.part0;
    }
    K.nparts = (int)parts.size();
    int nslots = 2 * h->num_sms;
    if (const char *e = getenv("SML_STEP_SLOTS")) nslots = std::max(1, atoi(e));
    nslots = std::max(1, std::min(nslots, K.nparts));
    K.nslots = nslots;
    std::vector<StepSeg> segs;
    std::vector<int2> slots(nslots, make_int2(0, 0));
    long long cum = 0;
    int cur_slot = -1;
    for (const Part &pt : parts) {
        // the slot is chosen by the block's midpoint on the cost axis: contiguous runs, balanced to within one block
        int sl = (int)(((cum + pt.cost / 2) * (long long)nslots) / std::max<long long>(1, total_cost));
        sl = std::min(std::max(sl, std::max(cur_slot, 0)), nslots - 1);
        cum += pt.cost;
        const bool new_slot = sl != cur_slot;
        if (new_slot) {
            cur_slot = sl;
            slots[sl].x = (int)segs.size();
        }
        // the two CTAs of an SM start half an item apart: every other slot opens with a one-block item
        const bool first_of_slot_short = stagger && (sl & 1) && slots[sl].y == 1 && segs.size() == (size_t)slots[sl].x + 1;
        if (!new_slot && !first_of_slot_short && !segs.empty() && segs.back().reg == pt.reg &&
            segs.back().row0 + segs.back().nrows == pt.row0 && segs.back().nrows + pt.nrows <= max_item_rows) {
            segs.back().nrows += pt.nrows;
        } else {
            StepSeg sg{pt.reg, pt.row0, pt.nrows, pt.part, pt.row0 == 0 ? 1 : 0};
            segs.push_back(sg);
            slots[sl].y++;
        }
    }
    K.p_xs_cap = (S_max + max_item_rows + 1) & ~1;
    K.p_smem_bytes = (size_t)K.p_stages * stage_cols * ldp * 8 + (size_t)K.p_xs_cap * 8 + 2 * (size_t)K.p_stages * 8;
    if (K.p_smem_bytes > 112 * 1024) {   // two CTAs per SM or nothing
        K.persist = false;
        K.p_smem_bytes = 0;
        return 0;
    }
    std::vector<StepSeg> split = segs;
    for (StepSeg &sg : split) sg.with_model = 0;
    CK(h, cudaMalloc(&K.d_segs, sizeof(StepSeg) * std::max<size_t>(1, segs.size())));
    CK(h, cudaMemcpy(K.d_segs, segs.data(), sizeof(StepSeg) * segs.size(), cudaMemcpyHostToDevice));
    CK(h, cudaMalloc(&K.d_segs_split, sizeof(StepSeg) * std::max<size_t>(1, split.size())));
    CK(h, cudaMemcpy(K.d_segs_split, split.data(), sizeof(StepSeg) * split.size(), cudaMemcpyHostToDevice));
    CK(h, cudaMalloc(&K.d_slots, sizeof(int2) * nslots));
    CK(h, cudaMemcpy(K.d_slots, slots.data(), sizeof(int2) * nslots, cudaMemcpyHostToDevice));
    return 0;
}

static int finalize_kind(sml_engine *h, int kind)
{
    KindState &K = h->kinds[kind];
    const int nloc = (int)K.regs.size();
    if (!K.any) {
        if (kind == SML_OCEAN && h->p.slab_ocean_model_bool) {
            // no ocean reservoir on this rank: every region reports 272.0 (src/mpires.f90:323-326)
            K.P = 2 * h->tiling.fx * h->tiling.fy;
            K.ldw = K.P;
            std::vector<double> init((size_t)nloc * K.P, 272.0);
            CK(h, cudaMalloc(&K.d_out, sizeof(double) * init.size()));
            CK(h, cudaMemcpy(K.d_out, init.data(), sizeof(double) * init.size(), cudaMemcpyHostToDevice));
        }
        return 0;
    }
    if (!K.gen.empty()) {
        // ---- makesparse + W_in for every generated region of the kind, two launches for all of them
        int n_cap = 0, w_gen = 1, ek = 0;
        std::vector<int> gen_of_local(nloc, -1);
        for (size_t i = 0; i < K.gen.size(); ++i) {
            n_cap = std::max(n_cap, K.gen[i].n);
            w_gen = std::max(w_gen, K.gen[i].W);
            ek = std::max(ek, std::max(K.gen[i].k, K.gen[i].n));
            gen_of_local[K.gen[i].local] = (int)i;
        }
        CK(h, cudaMalloc(&K.d_gen, sizeof(GenDesc) * K.gen.size()));
        CK(h, cudaMemcpy(K.d_gen, K.gen.data(), sizeof(GenDesc) * K.gen.size(), cudaMemcpyHostToDevice));
        CK(h, cudaMalloc(&K.d_gen_of_local, sizeof(int) * nloc));
        CK(h, cudaMemcpy(K.d_gen_of_local, gen_of_local.data(), sizeof(int) * nloc, cudaMemcpyHostToDevice));
        const size_t ms_smem = sizeof(int) * 3 * (size_t)n_cap;
        if (ms_smem > 227 * 1024) FAIL(h, "reservoir of %d nodes is too large for the device k-shuffle", n_cap);
        CK(h, cudaFuncSetAttribute(k_makesparse_shuffle, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ms_smem));
        k_makesparse_shuffle<<<dim3(2 * w_gen, (unsigned)K.gen.size()), 256, ms_smem, h->stream>>>(K.d_gen, n_cap);
        k_makesparse_fill<<<dim3((ek + 255) / 256, (unsigned)K.gen.size()), 256, 0, h->stream>>>(K.d_gen);
        h->launches += 2;
        CK(h, cudaGetLastError());
        CK(h, cudaStreamSynchronize(h->stream));
    }
    // rows per CTA of the step kernel: 8-9 chunks per region.  Measured on B200 (profiles/round1_summary.md): 360..1027
    // rows differ by < 3 % at every shard size (1152..144 regions per GPU), 720 is best or within noise of the best
    int chunk_rows = 720;
    if (const char *s = getenv("SML_CHUNK_ROWS")) chunk_rows = std::max(64, atoi(s));
    K.chunk_rows = chunk_rows;
    long long xo = 0, fo = 0, lo = 0;
    int S_max = 0, max_rows = 0;
    K.items.clear();
    K.alg_bytes = 0;
    K.alg_bytes_update = 0;
    std::vector<long long> fb_offs(nloc, 0);
    for (int i = 0; i < nloc; ++i) {
        HostRegion &hr = K.regs[i];
        if (!hr.uploaded) {
            if (kind == SML_ATMO) FAIL(h, "atmosphere reservoir of region %d was never uploaded", h->local_ids[i]);
            hr.dev = RegionDev{};  // region without an ocean reservoir
            continue;
        }
        RegionDev &d = hr.dev;
        d.ell_stream = getenv("SML_ELL_STREAM") ? atoi(getenv("SML_ELL_STREAM")) : SML_ELL_STREAM_DEFAULT;
        d.x_off = xo; d.fb_off = fo; d.lm_off = lo; d.out_off = (long long)i * K.P;
        fb_offs[i] = fo;
        xo += (d.n + 31) / 32 * 32;
        fo += (d.D + 31) / 32 * 32;
        lo += (d.S + 31) / 32 * 32;
        S_max = std::max(S_max, d.S);
        K.n_max = std::max(K.n_max, d.n);
        K.D_max = std::max(K.D_max, d.D);
        const int nch = (d.n + chunk_rows - 1) / chunk_rows;
        const int rows_per = (d.n + nch - 1) / nch;
        d.item0 = (int)K.items.size();
        for (int c = 0, r0 = 0; r0 < d.n; ++c, r0 += rows_per) {
            StepItem it;
            it.reg = i;
            it.row0 = r0;
            it.nrows = std::min(rows_per, d.n - r0);
            it.col0 = (c == 0) ? 0 : d.S + r0;
            it.ncols = (c == 0) ? d.S + it.nrows : it.nrows;
            it.xs_off = (c == 0) ? d.S : 0;
            max_rows = std::max(max_rows, it.nrows);
            K.items.push_back(it);
        }
        d.nitems = (int)K.items.size() - d.item0;
        // algorithmic bytes per region-step: ELL adjacency + x read + x write +
        // compact W_in + feedback + W_out + local_model + outvec + mean/std
        K.alg_bytes_update += (int64_t)d.ell_w * d.n * 12 + 8LL * d.n * 2 + 12LL * d.n + 8LL * d.D;   // the synchronize step
        K.alg_bytes += (int64_t)d.ell_w * d.n * 12 + 8LL * d.n * 2 + 12LL * d.n + 8LL * d.D +
                       8LL * d.P * ((int64_t)d.n + d.S) + 8LL * d.S + 8LL * d.P + 16LL * d.L;
    }
    K.x_total = xo; K.fb_total = fo; K.lm_total = lo;
    K.nitems = (int)K.items.size();
    K.stage_cols = std::max(1, STAGE_BYTES_TARGET / (K.ldw * 8));
    if (const char *s = getenv("SML_STAGE_COLS")) K.stage_cols = std::max(1, atoi(s));
    K.stage_bytes = K.stage_cols * K.ldw * 8;
    K.xs_cap = (S_max + max_rows + 1) & ~1;
    K.smem_bytes = (size_t)STAGES * K.stage_bytes + ((size_t)K.xs_cap + 2 * NCONS) * 8 + 2 * STAGES * 8;
    if (K.smem_bytes > 227 * 1024) FAIL(h, "step kernel needs %zu B of shared memory", K.smem_bytes);

    if (build_persistent_plan(h, K, S_max)) return -1;   // fills part0 / nparts of every RegionDev
    {
        // L2 residency (kernels.cuh, l2_policy_*): when the shard's adjacency, W_in and state vectors fit in L2 beside the
        // W_out stream, the persistent step kernel loads them evict-last and streams W_out evict-first.
        // SML_L2_KEEP=0/1 forces it off/on; default: always on with the persistent kernel (measured: -4.9 % at 144 regions
        // per GPU, -2.8 % at 288, -2.7 % at 1152, where only the state vectors and the partials can stay); SML_L2_KEEP_MB caps it
        size_t keep_bytes = 0;
        for (int i = 0; i < nloc; ++i)
            if (K.regs[i].uploaded) {
                const RegionDev &d = K.regs[i].dev;
                keep_bytes += (size_t)d.n * (12 * (size_t)d.ell_w + 12 + 16) + 8 * (size_t)d.D;
            }
        const char *lk = getenv("SML_L2_KEEP");
        const size_t cap_mb = getenv("SML_L2_KEEP_MB") ? (size_t)atoi(getenv("SML_L2_KEEP_MB")) : ((size_t)1 << 30);
        const bool keep = K.persist && (lk ? atoi(lk) != 0 : keep_bytes <= cap_mb * 1024 * 1024);
        K.l2_keep = keep;
        // (a persisting carve-out of L2 -- cudaLimitPersistingL2CacheSize -- for the evict-last lines was measured too and
        //  is WORSE: 0.1672 -> 0.1760 ms at 144 regions, 1.2548 -> 1.2987 ms at 1152; the policies alone are the default)
        if (keep)
            for (int i = 0; i < nloc; ++i)
                if (K.regs[i].uploaded && getenv("SML_ELL_STREAM") == nullptr) K.regs[i].dev.ell_stream = 2;
    }
    std::vector<RegionDev> regs(nloc);
    for (int i = 0; i < nloc; ++i) regs[i] = K.regs[i].dev;
    CK(h, cudaMalloc(&K.d_regs, sizeof(RegionDev) * nloc));
    CK(h, cudaMemcpy(K.d_regs, regs.data(), sizeof(RegionDev) * nloc, cudaMemcpyHostToDevice));
    CK(h, cudaMalloc(&K.d_items, sizeof(StepItem) * std::max(1, K.nitems)));
    CK(h, cudaMemcpy(K.d_items, K.items.data(), sizeof(StepItem) * K.nitems, cudaMemcpyHostToDevice));
    {
        // optional launch order (SML_LPT): largest items first so that the last CTAs to start are the cheapest
        std::vector<int> order(K.nitems);
        for (int i = 0; i < K.nitems; ++i) order[i] = i;
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return K.items[a].ncols > K.items[b].ncols; });
        CK(h, cudaMalloc(&K.d_order, sizeof(int) * std::max(1, K.nitems)));
        CK(h, cudaMemcpy(K.d_order, order.data(), sizeof(int) * K.nitems, cudaMemcpyHostToDevice));
    }
    {
        std::vector<StepItem> split = K.items;
        for (StepItem &it : split) {
            it.col0 = K.regs[it.reg].dev.S + it.row0;
            it.ncols = it.nrows;
            it.xs_off = 0;
        }
        CK(h, cudaMalloc(&K.d_items_split, sizeof(StepItem) * std::max(1, K.nitems)));
        CK(h, cudaMemcpy(K.d_items_split, split.data(), sizeof(StepItem) * K.nitems, cudaMemcpyHostToDevice));
    }
    CK(h, cudaMalloc(&K.d_fb_offs, sizeof(long long) * nloc));
    CK(h, cudaMemcpy(K.d_fb_offs, fb_offs.data(), sizeof(long long) * nloc, cudaMemcpyHostToDevice));
    for (int b = 0; b < 2; ++b) {
        CK(h, cudaMalloc(&K.d_x[b], sizeof(double) * std::max<long long>(1, xo)));
        CK(h, cudaMemset(K.d_x[b], 0, sizeof(double) * std::max<long long>(1, xo)));
    }
    CK(h, cudaMalloc(&K.d_fb, sizeof(double) * std::max<long long>(1, fo)));
    CK(h, cudaMemset(K.d_fb, 0, sizeof(double) * std::max<long long>(1, fo)));
    CK(h, cudaMalloc(&K.d_lm, sizeof(double) * std::max<long long>(1, lo)));
    CK(h, cudaMemset(K.d_lm, 0, sizeof(double) * std::max<long long>(1, lo)));
    {
        // ocean slab rows of regions without an ocean reservoir stay at 272.0 (src/mpires.f90:323-326)
        std::vector<double> init((size_t)nloc * K.P, kind == SML_OCEAN ? 272.0 : 0.0);
        CK(h, cudaMalloc(&K.d_out, sizeof(double) * init.size()));
        CK(h, cudaMemcpy(K.d_out, init.data(), sizeof(double) * init.size(), cudaMemcpyHostToDevice));
    }
    CK(h, cudaMalloc(&K.d_partials, sizeof(double) * (size_t)std::max(std::max(1, K.nitems), K.nparts) * K.ldw));
    if (K.any_dense) {
        CK(h, cudaMalloc(&K.d_temp, sizeof(double) * std::max<long long>(1, xo)));
        CK(h, cudaMemset(K.d_temp, 0, sizeof(double) * std::max<long long>(1, xo)));
    }
    return 0;
}

int sml_finalize(sml_engine *h)
{
    if (!h) return -1;
    if (h->finalized) FAIL(h, "sml_finalize called twice");
    CK(h, cudaSetDevice(h->p.device));
    for (int k = 0; k < 2; ++k)
        if (finalize_kind(h, k)) return -1;
    if (!h->kinds[SML_ATMO].any) FAIL(h, "no atmosphere reservoirs uploaded");
    {
        // the dynamic shared-memory limit is a property of the kernel, not of a launch: take the larger kind
        // ... and of the process, not of an engine: several engines may live side by side (tests, one process driving
        // two GPUs), so the limits only ever grow
        static size_t smem_max = 0, psmem_max = 0;
        smem_max = std::max(smem_max, std::max(h->kinds[SML_ATMO].smem_bytes, h->kinds[SML_OCEAN].smem_bytes));
        psmem_max = std::max(psmem_max, std::max<size_t>(1024, std::max(h->kinds[SML_ATMO].p_smem_bytes, h->kinds[SML_OCEAN].p_smem_bytes)));
        CK(h, cudaFuncSetAttribute(k_step<STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
        CK(h, cudaFuncSetAttribute(k_step_persist, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psmem_max));
    }
    const int R = h->p.number_of_regions, P = h->kinds[SML_ATMO].P;
    h->P_atmo = P;
    h->P_ocean = h->kinds[SML_OCEAN].P > 0 ? h->kinds[SML_OCEAN].P : 2 * h->tiling.fx * h->tiling.fy;
    if (h->kinds[SML_OCEAN].any) {
        // feedback assembly of the ocean reservoirs: ring of timestep_slab/timestep - 1 slots
        // (src/mod_slab_ocean_reservoir.f90:810), zeroed (:811)
        KindState &KA = h->kinds[SML_ATMO], &KO = h->kinds[SML_OCEAN];
        if (h->p.timestep <= 0 || h->p.timestep_slab / h->p.timestep - 1 < 1) FAIL(h, "bad timestep / timestep_slab");
        h->ocean_slots = h->p.timestep_slab / h->p.timestep - 1;
        std::vector<OceanFb> fbs;
        size_t ring = 0;
        for (size_t i = 0; i < KO.regs.size(); ++i) {
            HostRegion &ho = KO.regs[i];
            if (!ho.uploaded) continue;
            const RegionDev &da = KA.regs[i].dev;
            if (ho.osizes.atmo_slice0 + ho.osizes.A > da.D) FAIL(h, "internal: atmosphere slice of ocean region out of range");
            OceanFb f{};
            f.atmo_fb_off = da.fb_off + ho.osizes.atmo_slice0;
            f.fb_off = ho.dev.fb_off;
            f.ring_off = (long long)ring;
            f.A = ho.osizes.A;
            f.ixy = ho.osizes.ixy;
            f.sst_src = ho.d_sst_src;
            f.sst_mean = ho.sst_mean;
            f.sst_std = ho.sst_std;
            ring += (size_t)f.A * h->ocean_slots;
            fbs.push_back(f);
        }
        h->n_ocean_fb = (int)fbs.size();
        h->ocean_ring_doubles = ring;
        CK(h, cudaMalloc(&h->d_ocean_fb, sizeof(OceanFb) * fbs.size()));
        CK(h, cudaMemcpy(h->d_ocean_fb, fbs.data(), sizeof(OceanFb) * fbs.size(), cudaMemcpyHostToDevice));
        CK(h, cudaMalloc(&h->d_ocean_ring, sizeof(double) * ring));
        CK(h, cudaMemset(h->d_ocean_ring, 0, sizeof(double) * ring));
    }
    // scatter table for ALL regions of the model (every rank rebuilds the whole grid after the all-gather)
    std::vector<int> out_dst((size_t)R * P);
    std::vector<int> cell_region(XG * YG), cell_slot(XG * YG);
    for (int r = 0; r < R; ++r) {
        RegionGeom g = make_geom(h->tiling, r, h->p.overlap);
        RegionSizes s = make_sizes(h->tiling, g, 6000, 6.0, h->p.precip_bool, h->p.slab_ocean_model_bool, false,
                                   h->p.ml_only);
        RegionMaps m = make_maps(h->tiling, g, s, h->p.precip_bool, false);
        if (s.P != P) FAIL(h, "internal: P mismatch");
        std::copy(m.output_map.begin(), m.output_map.end(), out_dst.begin() + (size_t)r * P);
        for (int ry = 0; ry < h->tiling.fy; ++ry)
            for (int rx = 0; rx < h->tiling.fx; ++rx) {
                const int e = (int)off2(g.xs - 1 + rx, g.ys - 1 + ry);
                cell_region[e] = r;
                cell_slot[e] = rx + h->tiling.fx * ry;
            }
    }
    CK(h, cudaMalloc(&h->d_out_dst, sizeof(int) * out_dst.size()));
    CK(h, cudaMemcpy(h->d_out_dst, out_dst.data(), sizeof(int) * out_dst.size(), cudaMemcpyHostToDevice));
    CK(h, cudaMalloc(&h->d_cell_region, sizeof(int) * XG * YG));
    CK(h, cudaMemcpy(h->d_cell_region, cell_region.data(), sizeof(int) * XG * YG, cudaMemcpyHostToDevice));
    CK(h, cudaMalloc(&h->d_cell_slot, sizeof(int) * XG * YG));
    CK(h, cudaMemcpy(h->d_cell_slot, cell_slot.data(), sizeof(int) * XG * YG, cudaMemcpyHostToDevice));
    CK(h, cudaMalloc(&h->d_G, sizeof(double) * G_TOTAL));
    CK(h, cudaMemset(h->d_G, 0, sizeof(double) * G_TOTAL));
    if (h->p.numprocs > 1) {
        // the exchange block every peer maps (kernels.cuh PeerTable); the forecast buffer F is its landing section
        auto al = [](size_t b) { return (b + 255) / 256 * 256; };
        h->xo_atmo = 0;
        h->xo_ocean = (long long)al(sizeof(double) * 2 * (size_t)R * P);
        h->xo_fcst = h->xo_ocean + (long long)al(sizeof(double) * 2 * (size_t)R * h->P_ocean);
        h->xo_flags = h->xo_fcst + (long long)al(sizeof(double) * FCST_DOUBLES);
        h->xchg_bytes = (size_t)h->xo_flags + al(sizeof(unsigned long long) * XF_KINDS * MAX_PEERS);
        CK(h, cudaMalloc(&h->d_xchg, h->xchg_bytes));
        CK(h, cudaMemset(h->d_xchg, 0, h->xchg_bytes));
        {
            std::vector<double> init(2 * (size_t)R * h->P_ocean, 272.0);   // src/mpires.f90:323-326
            CK(h, cudaMemcpy((char *)h->d_xchg + h->xo_ocean, init.data(), sizeof(double) * init.size(), cudaMemcpyHostToDevice));
        }
        h->d_F = (double *)((char *)h->d_xchg + h->xo_fcst);
    } else {
        CK(h, cudaMalloc(&h->d_F, sizeof(double) * FCST_DOUBLES));
        CK(h, cudaMemset(h->d_F, 0, sizeof(double) * FCST_DOUBLES));
    }
    CK(h, cudaMalloc(&h->d_gathered, sizeof(double) * (size_t)R * P));
    CK(h, cudaMemset(h->d_gathered, 0, sizeof(double) * (size_t)R * P));
    CK(h, cudaMalloc(&h->d_base_sst, sizeof(double) * XG * YG));
    CK(h, cudaMemset(h->d_base_sst, 0, sizeof(double) * XG * YG));
    CK(h, cudaMalloc(&h->d_mask, sizeof(double) * XG * YG));
    CK(h, cudaMemset(h->d_mask, 0, sizeof(double) * XG * YG));
    CK(h, cudaMalloc(&h->d_prescribed, sizeof(double) * XG * YG));
    CK(h, cudaMemset(h->d_prescribed, 0, sizeof(double) * XG * YG));
    {
        std::vector<double> init((size_t)R * h->P_ocean, 272.0);
        CK(h, cudaMalloc(&h->d_ocean_gathered, sizeof(double) * init.size()));
        CK(h, cudaMemcpy(h->d_ocean_gathered, init.data(), sizeof(double) * init.size(), cudaMemcpyHostToDevice));
    }
    CK(h, cudaMallocHost(&h->h_pin_G, sizeof(double) * G_TOTAL));
    CK(h, cudaMallocHost(&h->h_pin_F, sizeof(double) * FCST_DOUBLES));
    std::memset(h->h_pin_F, 0, sizeof(double) * FCST_DOUBLES);
    h->h_pin_F[FCST_FLAG] = 1.0;
    CK(h, cudaMallocHost(&h->h_pin_tisr, sizeof(double) * XG * YG));
    CK(h, cudaMalloc(&h->d_done, 4 * sizeof(unsigned int)));
    CK(h, cudaMemset(h->d_done, 0, 4 * sizeof(unsigned int)));
    CK(h, cudaMalloc(&h->d_peer_err, sizeof(int)));
    CK(h, cudaMemset(h->d_peer_err, 0, sizeof(int)));
    CK(h, cudaMalloc(&h->d_status, sizeof(int)));
    CK(h, cudaMemset(h->d_status, 0, sizeof(int)));
    CK(h, cudaMallocHost(&h->h_status, sizeof(int)));
    *h->h_status = 0;
    CK(h, cudaMallocHost(&h->h_pin_flag, sizeof(double)));
    *h->h_pin_flag = 1.0;
    CK(h, cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    CK(h, cudaEventCreateWithFlags(&h->ev_pack, cudaEventDisableTiming));
    CK(h, cudaEventCreateWithFlags(&h->ev_d2h, cudaEventDisableTiming));
    CK(h, cudaEventCreateWithFlags(&h->ev_h2d, cudaEventDisableTiming));
    for (int k = 0; k < 2; ++k) h->kinds[k].pack_valid = false;
    h->finalized = true;
    return 0;
}

/* ------------------------------------------------------------------ state access */
static int copy_vec(sml_engine *h, int kind, int region, double *host, const double *chost, int which)
{
    if (check_ready(h, kind)) return -1;
    int li;
    if (local_of(h, kind, region, &li)) return -1;
    KindState &K = h->kinds[kind];
    const RegionDev &d = K.regs[li].dev;
    double *dp = nullptr;
    size_t cnt = 0;
    switch (which) {
    case 0: dp = K.d_x[K.cur] + d.x_off; cnt = d.n; break;
    case 1: dp = K.d_fb + d.fb_off; cnt = d.D; break;
    case 2: dp = K.d_lm + d.lm_off; cnt = d.S; break;
    case 3: dp = K.d_out + d.out_off; cnt = d.P; break;
    }
    CK(h, cudaSetDevice(h->p.device));
    if (cnt == 0) return 0;
    if (chost) CK(h, cudaMemcpyAsync(dp, chost, cnt * 8, cudaMemcpyHostToDevice, h->stream));
    else CK(h, cudaMemcpyAsync(host, dp, cnt * 8, cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    return 0;
}
int sml_state_set(sml_engine *h, int kind, int region, const double *x) { return copy_vec(h, kind, region, nullptr, x, 0); }
int sml_state_get(sml_engine *h, int kind, int region, double *x) { return copy_vec(h, kind, region, x, nullptr, 0); }
int sml_feedback_set(sml_engine *h, int kind, int region, const double *v) { return copy_vec(h, kind, region, nullptr, v, 1); }
int sml_feedback_get(sml_engine *h, int kind, int region, double *v) { return copy_vec(h, kind, region, v, nullptr, 1); }
int sml_local_model_set(sml_engine *h, int kind, int region, const double *v) { return copy_vec(h, kind, region, nullptr, v, 2); }
int sml_local_model_get(sml_engine *h, int kind, int region, double *v) { return copy_vec(h, kind, region, v, nullptr, 2); }
int sml_outvec_get(sml_engine *h, int kind, int region, double *v) { return copy_vec(h, kind, region, v, nullptr, 3); }
int sml_outvec_set(sml_engine *h, int kind, int region, const double *v) { return copy_vec(h, kind, region, nullptr, v, 3); }

// every local region's outvec in ONE copy: slab[i*P + p] for local region i (rows of regions without a reservoir of
// the kind hold what the exchange uses for them); the batched form of the per-region reads of reservoir%outvec
int sml_outvec_get_all(sml_engine *h, int kind, double *slab)
{
    if (check_ready(h, kind)) return -1;
    if (!slab) return -1;
    KindState &K = h->kinds[kind];
    CK(h, cudaSetDevice(h->p.device));
    CK(h, cudaMemcpyAsync(slab, K.d_out, sizeof(double) * K.regs.size() * K.P, cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    return 0;
}

int sml_wout_get(sml_engine *h, int kind, int region, double *wout)
{
    if (check_ready(h, kind)) return -1;
    int li;
    if (local_of(h, kind, region, &li)) return -1;
    const RegionDev &d = h->kinds[kind].regs[li].dev;
    CK(h, cudaMemcpy2DAsync(wout, (size_t)d.P * 8, d.wout, (size_t)d.ldw * 8, (size_t)d.P * 8, (size_t)d.n + d.S,
                            cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    return 0;
}
int sml_wout_set(sml_engine *h, int kind, int region, const double *wout)
{
    if (check_ready(h, kind)) return -1;
    int li;
    if (local_of(h, kind, region, &li)) return -1;
    const RegionDev &d = h->kinds[kind].regs[li].dev;
    CK(h, cudaMemcpy2DAsync(const_cast<double *>(d.wout), (size_t)d.ldw * 8, wout, (size_t)d.P * 8, (size_t)d.P * 8,
                            (size_t)d.n + d.S, cudaMemcpyHostToDevice, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    return 0;
}

/* ------------------------------------------------------------------ step launches */
static int launch_step(sml_engine *h, KindState &K, const StepItem *d_items, int nitems, const double *u_pool,
                       const long long *u_offs, int u_t, int do_readout, int cur_override = -1)
{
    if (nitems == 0) return 0;
    // which buffer of the state ping-pong pair is read (the other one is written)
    struct CurGuard {
        KindState &K; int saved;
        ~CurGuard() { K.cur = saved; }
    } cur_guard{K, K.cur};
    if (cur_override >= 0) K.cur = cur_override;
    const bool all = (d_items == K.d_items || d_items == K.d_items_split) && nitems == K.nitems;
    int nreg = (int)K.regs.size();
    const int *list = nullptr;
    if (!all) {
        // one region's synchronize: every kernel of the step is restricted to it (the other regions' input slots in
        // K.d_in are not theirs, so the dense-W_in kernel must not touch them either)
        if (!K.d_one_region) CK(h, cudaMalloc(&K.d_one_region, sizeof(int)));
        const int reg = K.items[(int)(d_items - K.d_items)].reg;
        if (reg != K.one_region) {
            CK(h, cudaMemcpyAsync(K.d_one_region, &reg, sizeof(int), cudaMemcpyHostToDevice, h->stream));
            CK(h, cudaStreamSynchronize(h->stream));
            K.one_region = reg;
        }
        list = K.d_one_region;
        nreg = 1;
    }
    if (K.any_dense) {
        dim3 grid((K.n_max + 255) / 256, (unsigned)nreg);
        k_win_dense<<<grid, 256, 0, h->stream>>>(K.d_regs, list, u_pool, u_offs, u_t, K.d_temp);
        h->launches++;
    }
    if (!do_readout) {
        // update-only step (synchronize): the dedicated latency-tolerant kernel, all regions or one region's slice
        // rows per thread: measured 0.200 / 0.206 / 0.151 ms for 1 / 2 / 4 at the bench config (profiles/round1_summary.md)
        const int rpt = getenv("SML_UPDATE_RPT") ? atoi(getenv("SML_UPDATE_RPT")) : 4;  // A/B switch
        const int RPT = (rpt == 1 || rpt == 2 || rpt == 8) ? rpt : 4;
        // SML_UPDATE_KERNEL=ring: TMA-fed ELL ring with a producer warp (k_update_ring, the default when it fits);
        //                   =sx: state vector staged in shared memory (k_update_sx); =global: gathers from L2 (k_update)
        const char *uk = getenv("SML_UPDATE_KERNEL");
        {
            const std::string want = uk ? uk : SML_UPDATE_DEFAULT;
            const int xs_cap_r = (K.n_max + 1) & ~1, us_cap_r = (K.D_max + 1) & ~1;
            int w_max = 1;
            bool aligned = true;
            for (const HostRegion &hr : K.regs)
                if (hr.uploaded) {
                    w_max = std::max(w_max, hr.dev.ell_w);
                    if (hr.dev.n % 4 != 0) aligned = false;   // the tile copies need 16-byte multiples
                }
            // tile rows: as many as fit twice beside the state vector, at most one per consumer thread.  Large tiles keep
            // the bulk copies long (>= several KB each): the producer's issue cost per copy does not depend on its size
            const size_t fixed = sizeof(double) * ((size_t)xs_cap_r + us_cap_r) + 64;
            // 3 stages of ~640 rows measured better than 2 of 960 (0.794 / 0.719 of the roof at 1152 / 144 regions against
            // 0.758 / 0.651): a deeper ring matters more than the longest possible copies
            int nst = getenv("SML_UPDATE_STAGES") ? atoi(getenv("SML_UPDATE_STAGES")) : 3;
            nst = std::max(2, std::min(nst, 8));
            const size_t row_bytes = 12 * (size_t)w_max + 12;
            const size_t budget = 227 * 1024 - 64;
            int tr = fixed + 64 < budget ? (int)((budget - fixed - 16 * (size_t)nst) / ((size_t)nst * row_bytes)) : 0;
            tr = std::min(tr, UR_CONS) / 32 * 32;
            if (getenv("SML_UPDATE_TILE_ROWS")) tr = std::min(tr, std::max(32, atoi(getenv("SML_UPDATE_TILE_ROWS")) / 32 * 32));
            const size_t tile_stride = (size_t)tr * row_bytes;
            const size_t ring_smem = fixed + (size_t)nst * (tile_stride + 16) + 16;
            if (want == "ring" && aligned && tr >= 256 && ring_smem <= 227 * 1024) {
                // one CTA per SM: split a region's rows only when there are fewer regions than SMs
                int nsplit = getenv("SML_UPDATE_SPLIT") ? atoi(getenv("SML_UPDATE_SPLIT")) : h->num_sms / std::max(1, nreg);
                nsplit = std::max(1, std::min(nsplit, 16));
                CK(h, cudaFuncSetAttribute(k_update_ring, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ring_smem));
                k_update_ring<<<dim3((unsigned)nsplit, (unsigned)nreg), UR_THREADS, ring_smem, h->stream>>>(
                    K.d_regs, list, K.d_x[K.cur], K.d_x[K.cur ^ 1], u_pool, u_offs, u_t, K.d_temp, nsplit, xs_cap_r, us_cap_r, w_max, nst, tr);
                h->launches++;
                CK(h, cudaGetLastError());
                return 0;
            }
        }
        const int xs_cap = (K.n_max + 1) & ~1;
        const size_t sx_smem = sizeof(double) * ((size_t)xs_cap + ((K.D_max + 1) & ~1)) + 16;
        const bool use_sx = (uk ? std::string(uk) != "global" : SML_UPDATE_SX_DEFAULT) && sx_smem <= 110 * 1024;
        if (use_sx) {
            // about one wave of 2 CTAs per SM when there are few regions (each split stages the whole x: L2 hits after
            // the first); measured best: 1 split at 1152 regions, 2 at 144 (tools/ab_update.py)
            int nsplit = getenv("SML_UPDATE_SPLIT") ? atoi(getenv("SML_UPDATE_SPLIT")) : (2 * h->num_sms + nreg / 2) / nreg;
            nsplit = std::max(1, std::min(nsplit, 16));
            const int sxr = getenv("SML_UPDATE_RPT") ? rpt : 2;   // rows per thread per sweep: 2 measured best (64 registers)
            const int SXR = (sxr == 1 || sxr == 4) ? sxr : 2;
            // block size: 512 measured best (0.796 of the roof against 0.781 / 0.776 / 0.706 for 480 / 448 / 384 threads,
            // although 480 divides the 5760 rows evenly); SML_UPDATE_THREADS is the A/B switch
            int nt = getenv("SML_UPDATE_THREADS") ? atoi(getenv("SML_UPDATE_THREADS")) / 32 * 32 : UPD_SX_THREADS;
            if (nt < 64 || nt > UPD_SX_THREADS) nt = UPD_SX_THREADS;
            dim3 grid((unsigned)nsplit, (unsigned)nreg);
            if (SXR == 2) {
                CK(h, cudaFuncSetAttribute(k_update_sx<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sx_smem));
                k_update_sx<2><<<grid, nt, sx_smem, h->stream>>>(K.d_regs, list, K.d_x[K.cur], K.d_x[K.cur ^ 1], u_pool, u_offs, u_t, K.d_temp, nsplit, xs_cap);
            } else if (SXR == 1) {
                CK(h, cudaFuncSetAttribute(k_update_sx<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sx_smem));
                k_update_sx<1><<<grid, nt, sx_smem, h->stream>>>(K.d_regs, list, K.d_x[K.cur], K.d_x[K.cur ^ 1], u_pool, u_offs, u_t, K.d_temp, nsplit, xs_cap);
            } else {
                CK(h, cudaFuncSetAttribute(k_update_sx<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sx_smem));
                k_update_sx<4><<<grid, nt, sx_smem, h->stream>>>(K.d_regs, list, K.d_x[K.cur], K.d_x[K.cur ^ 1], u_pool, u_offs, u_t, K.d_temp, nsplit, xs_cap);
            }
            h->launches++;
            CK(h, cudaGetLastError());
            return 0;
        }
        dim3 grid((K.n_max + 256 * RPT - 1) / (256 * RPT), (unsigned)nreg);
        if (RPT == 1)
            k_update<1><<<grid, 256, 0, h->stream>>>(K.d_regs, list, K.d_x[K.cur], K.d_x[K.cur ^ 1], u_pool, u_offs, u_t, K.d_temp);
        else if (RPT == 2)
            k_update<2><<<grid, 256, 0, h->stream>>>(K.d_regs, list, K.d_x[K.cur], K.d_x[K.cur ^ 1], u_pool, u_offs, u_t, K.d_temp);
        else if (RPT == 8)
            k_update<8><<<grid, 256, 0, h->stream>>>(K.d_regs, list, K.d_x[K.cur], K.d_x[K.cur ^ 1], u_pool, u_offs, u_t, K.d_temp);
        else
            k_update<4><<<grid, 256, 0, h->stream>>>(K.d_regs, list, K.d_x[K.cur], K.d_x[K.cur ^ 1], u_pool, u_offs, u_t, K.d_temp);
        h->launches++;
        CK(h, cudaGetLastError());
        return 0;
    }
    if (K.persist && all) {
        // persistent kernel: one CTA per slot, each with its statically balanced run of items
        const StepSeg *segs = (d_items == K.d_items_split) ? K.d_segs_split : K.d_segs;
        CK(h, launch_pdl(k_step_persist, dim3(K.nslots), dim3(NTHREADS), K.p_smem_bytes, h->stream,
                         K.d_regs, segs, K.d_slots, K.d_x[K.cur], K.d_x[K.cur ^ 1], u_pool, u_offs, u_t, K.d_lm, K.d_temp, K.d_partials,
                         K.ldw, K.p_stage_cols, K.p_ldp, K.p_xs_cap, K.part_rows, K.p_cpi, K.p_stages));
        h->launches++;
        CK(h, cudaGetLastError());
        return 0;
    }
    const size_t smem = K.smem_bytes;
    // the full item list is launched largest-first; partial lists (one region's synchronize) in place.
    // partials are indexed by the item's position in the FULL list (k_readout_finish reads item0 + c)
    // CTAs are launched in region-major item order: neighbouring CTAs stream neighbouring W_out columns.  A
    // largest-first (LPT) order was measured 1.8 % slower (1.2794 vs 1.2559 ms per launch, same box) and is off
    // unless SML_LPT is set
    static const bool lpt = getenv("SML_LPT") != nullptr;
    const bool full = (d_items == K.d_items || d_items == K.d_items_split) && nitems == K.nitems;
    const int item_base = full ? 0 : (int)(d_items - K.d_items);
    CK(h, launch_pdl(k_step<STAGES>, dim3(nitems), dim3(NTHREADS), smem, h->stream, K.d_regs, d_items,
                     (full && lpt) ? K.d_order : nullptr, item_base, K.d_x[K.cur], K.d_x[K.cur ^ 1], u_pool, u_offs, u_t, K.d_lm,
                     K.d_temp, K.d_partials, K.ldw, K.stage_cols, K.stage_bytes, K.xs_cap, do_readout));
    h->launches++;
    CK(h, cudaGetLastError());
    return 0;
}

// CUDA-event bracket of one phase of the step (only while profiling)
static cudaEvent_t *phase_events(sml_engine *h, sml_engine::PhaseTimer &t)
{
    if (!h->profile || t.used >= PROFILE_RING) return nullptr;
    while ((int)t.ev.size() < 2 * (t.used + 1)) {
        cudaEvent_t e;
        if (cudaEventCreate(&e) != cudaSuccess) return nullptr;
        t.ev.push_back(e);
    }
    return &t.ev[2 * t.used++];
}

// partials -> outvec slab (+ the model columns in the overlapped mode); for the atmosphere kind with peers attached
// the same kernel pushes the outvecs into every rank's gathered buffer and publishes the step
static int launch_finish(sml_engine *h, KindState &K, int model_part)
{
    PeerTable pt{};
    unsigned long long seq = 0;
    long long peer_off = 0;
    const bool atmo = &K == &h->kinds[SML_ATMO];
    if (atmo && h->peers.world > 1) {
        pt = h->peers;
        seq = ++h->peer_seq;
        const long long RP = (long long)h->p.number_of_regions * K.P;
        // rank-major == region order (contiguous shards): this rank's rows start at local_ids[0]
        peer_off = (long long)(seq & 1) * RP + (long long)h->local_ids[0] * K.P;
    }
    // sequential mode: one group of threads sums the partials; overlapped mode: 4 groups share the model columns
    const int threads = model_part ? FIN_GROUPS * FIN_PMAX : FIN_PMAX;
    const size_t fsmem = sizeof(double) * ((K.P + 1) & ~1) * (model_part ? FIN_GROUPS + 1 : 1);
    CK(h, launch_pdl(k_readout_finish, dim3((unsigned)K.regs.size()), dim3(threads), fsmem, h->stream, K.d_regs, K.d_partials, K.ldw,
                     K.d_out, 1, model_part, K.d_lm, pt, seq, peer_off, h->d_done,
                     (atmo && h->contribs && model_part) ? h->d_vp : nullptr,
                     (atmo && h->contribs && model_part) ? h->d_vml : nullptr, K.persist ? 1 : 0,
                     (!atmo && !model_part) ? K.d_lm : nullptr));
    h->launches++;
    CK(h, cudaGetLastError());
    if (!atmo && h->peers.world > 1) h->ocean_publish_pending = true;   // pushed by the next grid assembly
    return 0;
}

// the ocean reservoirs' outvec slab [nloc][P_ocean] into every rank's gathered copy (the all-gather after
// predict_slab_ml, src/mpires.f90:375-454): one push kernel over NVLink, double-buffered by ocean-step parity
static int push_ocean_slab(sml_engine *h)
{
    KindState &KO = h->kinds[SML_OCEAN];
    const long long Po = h->P_ocean, nloc = (long long)KO.regs.size();
    const unsigned long long seq = ++h->ocean_seq;
    const long long dst_off = (long long)(seq & 1) * h->p.number_of_regions * Po + (long long)h->local_ids[0] * Po;
    k_peer_push<<<dim3(1, h->peers.world), 256, 0, h->stream>>>(KO.d_out, nloc * Po, h->peers, XF_OCEAN, dst_off, 0, seq,
                                                                h->p.irank, h->d_done + 1);
    h->launches++;
    CK(h, cudaGetLastError());
    h->ocean_publish_pending = false;
    return 0;
}

int sml_predict(sml_engine *h, int kind)
{
    if (h && h->finalized && kind == SML_OCEAN && h->p.slab_ocean_model_bool && !h->kinds[SML_OCEAN].any) {
        // a rank whose regions are all land: nothing to step, but the ocean slabs are still re-published with the
        // other ranks' (every rank pushes after every ocean step)
        if (h->peers.world > 1) h->ocean_publish_pending = true;
        return 0;
    }
    if (check_ready(h, kind)) return -1;
    CK(h, cudaSetDevice(h->p.device));
    KindState &K = h->kinds[kind];
    if (kind == SML_ATMO) {
        if (h->ahead_pending) FAIL(h, "overlapped step: sml_step_exchange_end / sml_step_unpack_device has not closed the previous step");
        if (h->ahead_done) {  // this predict was computed while the host model ran
            h->ahead_done = false;
            return 0;
        }
    }
    cudaEvent_t *ev = nullptr;
    if (h->profile && h->ev_used < PROFILE_RING) {
        while ((int)h->ev.size() < 3 * (h->ev_used + 1)) {
            cudaEvent_t e;
            CK(h, cudaEventCreate(&e));
            h->ev.push_back(e);
        }
        ev = &h->ev[3 * h->ev_used++];
    }
    if (ev) CK(h, cudaEventRecord(ev[0], h->stream));
    // the finish stays a separate launch: folding it into the step kernel's tail (last chunk of a region reduces)
    // was measured 2.6 % slower -- the fence + atomic round trip idles every CTA slot for ~1 us per chunk
    // with outvec_component_contribs the readout runs in split order (v_p + v_ml) so that both halves exist
    const bool split = kind == SML_ATMO && h->contribs;
    if (launch_step(h, K, split ? K.d_items_split : K.d_items, K.nitems, K.d_fb, K.d_fb_offs, 0, 1)) return -1;
    K.cur ^= 1;
    if (ev) CK(h, cudaEventRecord(ev[1], h->stream));
    if (launch_finish(h, K, split ? 1 : 0)) return -1;
    if (ev) CK(h, cudaEventRecord(ev[2], h->stream));
    return 0;
}

// Ring geometry of the time-loop kernels (k_sync_persist, k_train_stategen_ring) and the tile-major pack they stream.
// The geometry follows the whole kind (not the regions of one call), so that one pack serves every call.
struct SyncPlan {
    int ngroups = 0, nst = 0, tr = 0, w_max = 1, xs_cap = 0, us_cap = 0;
    size_t tile_stride = 0, smem = 0;
};
// -> 1: plan made and the pack is current; 0: the kind does not qualify (too large for shared memory, > 16-bit indices); -1: error
static int sync_plan_and_pack(sml_engine *h, KindState &K, SyncPlan &P)
{
    int w_max = 1;
    const int n_max = K.n_max, D_max = K.D_max;
    for (const HostRegion &hr : K.regs)
        if (hr.uploaded) w_max = std::max(w_max, hr.dev.ell_w);
    // plan: ngroups consumer groups of tr threads, nstages ring slots of tr rows (nstages a multiple of ngroups)
    const int xs_cap = (n_max + 1) & ~1, us_cap = (D_max + 1) & ~1;
    const size_t fixed = sizeof(double) * 2 * ((size_t)xs_cap + us_cap);
    if (n_max > 65535 || D_max > 65535) return 0;   // the tile-major pack stores 16-bit column indices
    const size_t row_bytes = (size_t)sp_row_bytes(w_max);
    const size_t budget = 227 * 1024 - 128;
    int ngroups = getenv("SML_SYNC_GROUPS") ? atoi(getenv("SML_SYNC_GROUPS")) : 2;
    ngroups = std::max(1, std::min(ngroups, 4));
    int nst = getenv("SML_SYNC_STAGES") ? atoi(getenv("SML_SYNC_STAGES")) : 2 * ngroups;
    nst = std::max(ngroups, std::min(nst, 12)) / ngroups * ngroups;
    if (fixed + 4096 > budget) return 0;
    const size_t bar_bytes = 8 * (2 * (size_t)nst + 5);
    int tr = (int)((budget - fixed - bar_bytes) / ((size_t)nst * row_bytes));
    tr = std::min(tr, (SP_MAX_THREADS - 32) / ngroups) / 32 * 32;
    if (getenv("SML_SYNC_TILE_ROWS")) tr = std::min(tr, std::max(32, atoi(getenv("SML_SYNC_TILE_ROWS")) / 32 * 32));
    if (tr < 64) return 0;
    {   // even out the tiles of the largest region: same tile count, no short last tile
        const int nt = (n_max + tr - 1) / tr;
        tr = std::min(tr, ((n_max + nt - 1) / nt + 31) / 32 * 32);
    }
    const size_t tile_stride = (size_t)tr * row_bytes;
    const size_t smem = fixed + (size_t)nst * tile_stride + bar_bytes;
    if (smem > 227 * 1024 || tile_stride >= (1u << 20)) return 0;
    // large reservoirs (m = 12000: two state vectors take 190 KB) leave room for a ring of ~100-row tiles only, i.e. ~200
    // consumer threads: below ~400 threads the time loop is slower than the step launches (measured: 576 threads 36 ms,
    // 768+ threads 24 ms, step launches 49 ms at m = 6000), so such shards keep the step-per-launch kernels
    if (!getenv("SML_SYNC_TILE_ROWS") && n_max >= 4096 && ngroups * tr < 384) return 0;
    P.ngroups = ngroups; P.nst = nst; P.tr = tr; P.w_max = w_max; P.xs_cap = xs_cap; P.us_cap = us_cap;
    P.tile_stride = tile_stride; P.smem = smem;
    // the tile-major pack of this geometry (all regions of the kind)
    if (!K.pack_valid || K.pack_tr != tr || K.pack_w != w_max) {
        const int nloc = (int)K.regs.size();
        K.sync_pack_off.assign(nloc, 0);
        size_t total = 0;
        for (int i = 0; i < nloc; ++i) {
            if (!K.regs[i].uploaded) continue;
            K.sync_pack_off[i] = (long long)total;
            total += (size_t)((K.regs[i].dev.n + tr - 1) / tr) * tile_stride;
        }
        if (total > K.sync_pack_cap) {
            cudaFree(K.d_sync_pack);
            K.d_sync_pack = nullptr;
            K.sync_pack_cap = 0;
            CK(h, cudaMalloc(&K.d_sync_pack, total));
            K.sync_pack_cap = total;
        }
        if (!K.d_sync_pack_off) CK(h, cudaMalloc(&K.d_sync_pack_off, sizeof(long long) * nloc));
        CK(h, cudaMemcpyAsync(K.d_sync_pack_off, K.sync_pack_off.data(), sizeof(long long) * nloc, cudaMemcpyHostToDevice, h->stream));
        k_sync_pack<<<dim3((unsigned)((n_max + tr - 1) / tr), (unsigned)nloc), 256, 0, h->stream>>>(K.d_regs, K.d_sync_pack,
                                                                                                     K.d_sync_pack_off, tr, w_max);
        h->launches++;
        CK(h, cudaGetLastError());
        CK(h, cudaStreamSynchronize(h->stream));
        K.pack_valid = true;
        K.pack_tr = tr;
        K.pack_w = w_max;
    }
    return 1;
}
static bool sync_region_ok(const RegionDev &d)
{
    // compact W_in, bulk copies in 16-byte multiples
    return d.win_mode == 0 && d.n % 4 == 0 && d.D % 2 == 0 && d.n > 0 && d.D > 0;
}

// The whole time loop of synchronize in ONE launch (k_sync_persist): a CTA per SM runs all `length` steps of a region
// before it takes the next one, the state vector stays in shared memory and the adjacency is re-read from L2.
// Returns 1 when launched, 0 when the shard does not qualify (dense W_in, odd sizes, no room in shared memory, or
// SML_SYNC_KERNEL=steps) and the caller launches step by step, -1 on error.  The state pool K.d_x[K.cur] is updated in place.
static int launch_sync_persist(sml_engine *h, KindState &K, int first, int last, int length)
{
    const char *sk = getenv("SML_SYNC_KERNEL");
    if (sk && std::string(sk) == "steps") return 0;
    if (K.any_dense) return 0;
    std::vector<int> list;
    for (int i = first; i < last; ++i) {
        const HostRegion &hr = K.regs[i];
        if (!hr.uploaded) continue;
        if (!sync_region_ok(hr.dev)) return 0;
        list.push_back(i);
    }
    if (list.empty()) return 1;
    SyncPlan P;
    const int ok = sync_plan_and_pack(h, K, P);
    if (ok <= 0) return ok;
    if (K.sync_list_cap < list.size()) {
        cudaFree(K.d_sync_list);
        K.d_sync_list = nullptr;
        K.sync_list_cap = 0;
        CK(h, cudaMalloc(&K.d_sync_list, sizeof(int) * K.regs.size()));
        K.sync_list_cap = K.regs.size();
    }
    CK(h, cudaMemcpyAsync(K.d_sync_list, list.data(), sizeof(int) * list.size(), cudaMemcpyHostToDevice, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));   // `list` is a local
    int ctas = getenv("SML_SYNC_CTAS") ? atoi(getenv("SML_SYNC_CTAS")) : h->num_sms;
    ctas = std::max(1, std::min({ctas, (int)list.size(), 4 * h->num_sms}));
    // compile-time width for the common cases (one column slab, 2 to 4 value pairs), any width otherwise
    auto launch = [&](auto kern) -> cudaError_t {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P.smem);
        if (e != cudaSuccess) return e;
        kern<<<ctas, P.ngroups * P.tr + 32, P.smem, h->stream>>>(K.d_regs, K.d_sync_list, (int)list.size(), K.d_x[K.cur], K.d_in,
                                                                 K.d_in_offs, length, P.xs_cap, P.us_cap, P.w_max, P.nst, P.tr,
                                                                 P.ngroups, K.d_sync_pack, K.d_sync_pack_off);
        return cudaSuccess;
    };
    const bool generic = getenv("SML_SYNC_GENERIC") != nullptr;   // test hook: the any-width instantiation
    if (P.w_max <= 4 && !generic) CK(h, launch(k_sync_persist<2>));
    else if (P.w_max <= 6 && !generic) CK(h, launch(k_sync_persist<3>));
    else if (P.w_max <= 7 && !generic) CK(h, launch(k_sync_persist<4>));
    else CK(h, launch(k_sync_persist<0>));
    h->launches++;
    CK(h, cudaGetLastError());
    return 1;
}

int sml_synchronize(sml_engine *h, int kind, int region, const double *inputs, int ld, int length,
                    const int64_t *offsets)
{
    if (check_ready(h, kind)) return -1;
    if (length <= 0) return 0;
    CK(h, cudaSetDevice(h->p.device));
    KindState &K = h->kinds[kind];
    const int nloc = (int)K.regs.size();
    std::vector<long long> offs(nloc, 0);
    size_t total = 0;
    int first = 0, last = nloc;
    if (region != SML_ALL_REGIONS) {
        int li;
        if (local_of(h, kind, region, &li)) return -1;
        first = li;
        last = li + 1;
        if (ld < K.regs[li].dev.D) FAIL(h, "synchronize: ld %d < reservoir_numinputs %d", ld, K.regs[li].dev.D);
    } else if (!offsets) {
        FAIL(h, "synchronize(ALL) needs the per-region offsets");
    }
    for (int i = first; i < last; ++i) {
        if (!K.regs[i].uploaded) continue;
        offs[i] = (long long)total;
        total += (size_t)K.regs[i].dev.D * length;
    }
    if (total > K.d_in_cap) {
        cudaFree(K.d_in);
        K.d_in = nullptr;
        CK(h, cudaMalloc(&K.d_in, total * 8));
        K.d_in_cap = total;
    }
    if (!K.d_in_offs) CK(h, cudaMalloc(&K.d_in_offs, sizeof(long long) * nloc));
    for (int i = first; i < last; ++i) {
        if (!K.regs[i].uploaded) continue;
        const int D = K.regs[i].dev.D;
        if (region != SML_ALL_REGIONS)
            CK(h, cudaMemcpy2DAsync(K.d_in + offs[i], (size_t)D * 8, inputs, (size_t)ld * 8, (size_t)D * 8, length,
                                    cudaMemcpyHostToDevice, h->stream));
        else
            CK(h, cudaMemcpyAsync(K.d_in + offs[i], inputs + offsets[i], (size_t)D * length * 8,
                                  cudaMemcpyHostToDevice, h->stream));
    }
    CK(h, cudaMemcpyAsync(K.d_in_offs, offs.data(), sizeof(long long) * nloc, cudaMemcpyHostToDevice, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    const StepItem *items = K.d_items;
    int nitems = K.nitems;
    if (region != SML_ALL_REGIONS) {
        items = K.d_items + K.regs[first].dev.item0;
        nitems = K.regs[first].dev.nitems;
    }
    cudaEvent_t se0 = nullptr, se1 = nullptr;
    if (h->profile) {
        CK(h, cudaEventCreate(&se0));
        CK(h, cudaEventCreate(&se1));
        CK(h, cudaEventRecord(se0, h->stream));
    }
    int persisted = launch_sync_persist(h, K, first, last, length);
    if (persisted < 0) return -1;
    for (int t = 0; t < length && !persisted; ++t) {
        if (region != SML_ALL_REGIONS) {
            // the untouched regions keep their state: only this region's slice of the ping-pong pair alternates, and
            // ONE copy at the end brings it back to the current buffer when the step count is odd
            if (launch_step(h, K, items, nitems, K.d_in, K.d_in_offs, t, 0, K.cur ^ (t & 1))) return -1;
        } else {
            if (launch_step(h, K, items, nitems, K.d_in, K.d_in_offs, t, 0)) return -1;
            K.cur ^= 1;
        }
    }
    if (!persisted && region != SML_ALL_REGIONS && (length & 1)) {
        const RegionDev &d = K.regs[first].dev;
        CK(h, cudaMemcpyAsync(K.d_x[K.cur] + d.x_off, K.d_x[K.cur ^ 1] + d.x_off, (size_t)d.n * 8,
                              cudaMemcpyDeviceToDevice, h->stream));
    }
    if (se0) CK(h, cudaEventRecord(se1, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    if (se0) {
        float ms = 0.f;
        CK(h, cudaEventElapsedTime(&ms, se0, se1));
        h->sync_ms_sum += ms;
        h->sync_steps += length;
        cudaEventDestroy(se0);
        cudaEventDestroy(se1);
    }
    return 0;
}

/* ------------------------------------------------------------------ exchange */
int sml_set_sst_static(sml_engine *h, const double *base, const double *mask)
{
    if (!h || !h->finalized || !base || !mask) return -1;
    CK(h, cudaSetDevice(h->p.device));
    // ordered on the engine's stream: a grid assembly still in flight reads the old fields
    CK(h, cudaMemcpyAsync(h->d_base_sst, base, sizeof(double) * XG * YG, cudaMemcpyHostToDevice, h->stream));
    CK(h, cudaMemcpyAsync(h->d_mask, mask, sizeof(double) * XG * YG, cudaMemcpyHostToDevice, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));   // caller-owned arrays: the engine holds copies when this returns
    h->sst_static_set = true;
    return 0;
}
int sml_set_sst_prescribed(sml_engine *h, const double *sst)
{
    if (!h || !h->finalized || !sst) return -1;
    CK(h, cudaSetDevice(h->p.device));
    CK(h, cudaMemcpyAsync(h->d_prescribed, sst, sizeof(double) * XG * YG, cudaMemcpyHostToDevice, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    h->sst_prescribed_set = true;
    return 0;
}

int sml_exchange_buffers(sml_engine *h, void **slab, int64_t *slab_count, void **gathered, int64_t *gathered_count,
                         void **gbuf, int64_t *g_count, void **fbuf, int64_t *f_count)
{
    if (check_ready(h, SML_ATMO)) return -1;
    KindState &K = h->kinds[SML_ATMO];
    *slab = K.d_out;
    *slab_count = (int64_t)K.regs.size() * K.P;
    *gathered = h->d_gathered;
    *gathered_count = (int64_t)h->p.number_of_regions * K.P;
    *gbuf = h->d_G;
    *g_count = G_TOTAL;
    *fbuf = h->d_F;
    *f_count = F_TOTAL;
    return 0;
}

int sml_ocean_exchange_buffers(sml_engine *h, void **slab, int64_t *slab_count, void **gathered,
                               int64_t *gathered_count)
{
    if (check_ready(h, SML_ATMO)) return -1;
    KindState &K = h->kinds[SML_OCEAN];
    if (!K.d_out) FAIL(h, "slab_ocean_model_bool is off: there is no ocean slab");
    *slab = K.d_out;
    *slab_count = (int64_t)K.regs.size() * h->P_ocean;
    *gathered = h->d_ocean_gathered;
    *gathered_count = (int64_t)h->p.number_of_regions * h->P_ocean;
    return 0;
}

int sml_ocean_ring_reset(sml_engine *h)
{
    if (check_ready(h, SML_ATMO)) return -1;
    if (h->d_ocean_ring)
        CK(h, cudaMemsetAsync(h->d_ocean_ring, 0, sizeof(double) * h->ocean_ring_doubles, h->stream));
    return 0;
}

static int build_pack_args(sml_engine *h, PackArgs &a);

int sml_step_pack_device(sml_engine *h, int timestep)
{
    (void)timestep;
    if (check_ready(h, SML_ATMO)) return -1;
    CK(h, cudaSetDevice(h->p.device));
    PackArgs a{};
    if (build_pack_args(h, a)) return -1;
    const int nsst = (a.sst_mode >= 0) ? (XG * YG + 255) / 256 : 0;
    cudaEvent_t *pe = phase_events(h, h->pt_pack);
    if (pe) CK(h, cudaEventRecord(pe[0], h->stream));
    CK(h, launch_pdl(k_pack_grids, dim3(a.nsc + nsst), dim3(256), 0, h->stream, a));
    if (pe) CK(h, cudaEventRecord(pe[1], h->stream));
    h->launches++;
    CK(h, cudaGetLastError());
    return 0;
}

// pack + unpack of the device-resident step as ONE cooperative launch (kernels.cuh k_exchange_fused); equivalent to
// sml_step_pack_device followed by sml_step_unpack_device(timestep) in the sequential mode, bit for bit
int sml_step_exchange_device(sml_engine *h, int timestep)
{
    if (check_ready(h, SML_ATMO)) return -1;
    CK(h, cudaSetDevice(h->p.device));
    if (h->overlap && h->ahead_pending) FAIL(h, "sml_step_exchange_device is the sequential device step; close the overlapped step first");
    KindState &K = h->kinds[SML_ATMO];
    // measured (profiles/round2_summary.md): 18.6 us against 19.7 us for the two kernels at 144 regions per GPU, but 36 us
    // against 23 us at 1152 (148 CTAs walk 1152 regions) -- so the separate kernels stay the default, this is the A/B switch
    static const bool fused = getenv("SML_FUSED_EXCHANGE") && atoi(getenv("SML_FUSED_EXCHANGE")) == 1;
    if (!fused) {
        if (sml_step_pack_device(h, timestep)) return -1;
        return sml_step_unpack_device(h, timestep);
    }
    if (h->n_ocean_fb > 0 && timestep < 1)
        FAIL(h, "timestep must be the 1-based hybrid step (it selects the ring slot mod(timestep-1,%d)+1)", h->ocean_slots);
    PackArgs a{};
    if (build_pack_args(h, a)) return -1;
    if (h->xch_calls >= (1u << 20)) {   // keep the arrival counter far from wrapping
        CK(h, cudaMemsetAsync(h->d_done + 3, 0, sizeof(unsigned int), h->stream));
        h->xch_calls = 0;
    }
    const unsigned grid = (unsigned)h->num_sms;
    unsigned target = grid * (++h->xch_calls);
    const RegionDev *regs = K.d_regs;
    int nreg = (int)K.regs.size(), do_model = h->p.ml_only ? 0 : 1;
    const double *F = h->d_F;
    double *fb = K.d_fb, *lm = K.d_lm;
    unsigned *ctr = h->d_done + 3;
    void *args[] = {&a, &regs, &nreg, &F, &fb, &lm, &do_model, &ctr, &target};
    cudaEvent_t *pe = phase_events(h, h->pt_pack);
    cudaEvent_t *ue = phase_events(h, h->pt_unpack);
    if (pe) CK(h, cudaEventRecord(pe[0], h->stream));
    CK(h, cudaLaunchCooperativeKernel((const void *)k_exchange_fused, dim3(grid), dim3(512), args, 0, h->stream));
    if (pe) CK(h, cudaEventRecord(pe[1], h->stream));
    if (ue) {   // the fused kernel is booked under "pack"; "unpack" reads zero
        CK(h, cudaEventRecord(ue[0], h->stream));
        CK(h, cudaEventRecord(ue[1], h->stream));
    }
    h->launches++;
    if (h->n_ocean_fb > 0) {
        k_build_ocean_inputs<<<h->n_ocean_fb, 128, 0, h->stream>>>(h->d_ocean_fb, h->d_G, K.d_fb, h->kinds[SML_OCEAN].d_fb,
                                                                   h->d_ocean_ring, (timestep - 1) % h->ocean_slots, h->ocean_slots);
        h->launches++;
    }
    CK(h, cudaGetLastError());
    return 0;
}

static int build_pack_args(sml_engine *h, PackArgs &a)
{
    KindState &K = h->kinds[SML_ATMO];
    a.total = h->p.number_of_regions * K.P;
    a.out_dst = h->d_out_dst;
    a.G = h->d_G;
    a.precip_lo = G_PRECIP; a.precip_hi = G_SST; a.w4d_hi = G_W2D; a.sst_off = G_SST;
    a.nsc = (a.total + 255) / 256;
    a.err = h->d_peer_err;
    a.status = h->d_status;
    const bool peers = h->peers.world > 1;
    const bool ocean_live = h->p.slab_ocean_model_bool && !h->p.sst_prescribed;
    if (peers && ocean_live && h->ocean_publish_pending)
        if (push_ocean_slab(h)) return -1;
    if (h->p.numprocs == 1) {
        a.gathered = K.d_out;
    } else if (peers) {
        // fused all-gather: wait for every rank's flag of this step, read this step's half of the exchange block
        a.gathered = h->peers.atmo(h->p.irank) + (size_t)(h->peer_seq & 1) * a.total;
        a.my_flags = h->peers.flag(h->p.irank, 0, 0);
        a.world = h->peers.world;
        a.seq = h->peer_seq;
        a.ocean_seq = ocean_live ? h->ocean_seq : 0;
    } else {
        a.gathered = h->d_gathered;  // filled by the host's collective (NCCL all-gather)
    }
    a.sst_mode = -1;
    if (h->p.slab_ocean_model_bool) {
        if (!h->sst_static_set) FAIL(h, "sml_set_sst_static (base_sst_grid, sea_mask) has not been called");
        a.sst_mode = h->p.sst_prescribed ? 1 : 0;
        if (a.sst_mode == 1 && !h->sst_prescribed_set) FAIL(h, "sst_prescribed is on but sml_set_sst_prescribed was never called");
        a.base = h->d_base_sst; a.mask = h->d_mask; a.prescribed = h->d_prescribed;
        a.cell_region = h->d_cell_region; a.cell_slot = h->d_cell_slot;
        // single rank: the ocean outvec slab already holds every region; with peers: the pushed copy of the latest
        // ocean step; otherwise the host all-gathers it
        a.ocean_out = (h->p.numprocs == 1) ? h->kinds[SML_OCEAN].d_out
                      : peers ? h->peers.ocean(h->p.irank) + (size_t)(h->ocean_seq & 1) * h->p.number_of_regions * h->P_ocean
                              : h->d_ocean_gathered;
        a.ocean_P = h->P_ocean;
    }
    return 0;
}

// ---- peer exchange set-up: every rank exports the IPC handle of its exchange block, the host passes all of them
// around (any transport: torch.distributed all_gather_object, MPI_Allgather of 64 bytes) and attaches them
int sml_peer_export(sml_engine *h, void *handle64)
{
    if (check_ready(h, SML_ATMO)) return -1;
    if (!h->d_xchg) FAIL(h, "peer exchange needs numprocs > 1");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    CK(h, cudaSetDevice(h->p.device));
    cudaIpcMemHandle_t hd;
    CK(h, cudaIpcGetMemHandle(&hd, h->d_xchg));
    std::memcpy(handle64, &hd, 64);
    return 0;
}

int sml_peer_attach(sml_engine *h, const void *handles, int count)
{
    if (check_ready(h, SML_ATMO)) return -1;
    if (!h->d_xchg) FAIL(h, "peer exchange needs numprocs > 1");
    if (count != h->p.numprocs) FAIL(h, "sml_peer_attach: %d handles for %d ranks", count, h->p.numprocs);
    if (count > MAX_PEERS) FAIL(h, "peer exchange supports up to %d ranks of one node", MAX_PEERS);
    if (h->peers.world > 1) FAIL(h, "peers are already attached");
    CK(h, cudaSetDevice(h->p.device));
    PeerTable pt{};
    pt.world = count;
    pt.rank = h->p.irank;
    pt.off_atmo = h->xo_atmo; pt.off_ocean = h->xo_ocean; pt.off_fcst = h->xo_fcst; pt.off_flags = h->xo_flags;
    for (int k = 0; k < count; ++k) {
        void *base = nullptr;
        if (k == h->p.irank) {
            base = h->d_xchg;
        } else {
            cudaIpcMemHandle_t hd;
            std::memcpy(&hd, (const char *)handles + 64 * (size_t)k, 64);
            cudaError_t e = cudaIpcOpenMemHandle(&base, hd, cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) {
                for (void *m : h->peer_mapped) cudaIpcCloseMemHandle(m);
                h->peer_mapped.clear();
                FAIL(h, "cudaIpcOpenMemHandle(rank %d): %s -- ranks must be on one node with peer access; use the host "
                        "collective (sml_exchange_buffers) instead", k, cudaGetErrorString(e));
            }
            h->peer_mapped.push_back(base);
        }
        pt.base[k] = (char *)base;
    }
    CK(h, cudaStreamSynchronize(h->stream));
    h->peers = pt;
    h->peer_seq = 0;
    h->ocean_seq = 0;
    h->fcst_seq = 0;
    h->ocean_publish_pending = h->p.slab_ocean_model_bool && !h->p.sst_prescribed;
    return 0;
}

// ---- sml_comm_bootstrap: the whole multi-rank set-up behind one call.  The host supplies ONE primitive -- an
// all-gather of a fixed-size byte block per rank (MPI_Allgather on the reference's mpi_world, torch.distributed, a
// shared-memory rendezvous) -- and the library does the rest: checks that every rank was built for the same model,
// exchanges the IPC handles of the exchange blocks and attaches them.  From then on sml_predict /
// sml_step_exchange_begin / sml_step_exchange_end are complete at numprocs > 1 with no host collective on the data
// path: outvec all-gather, ocean-slab gather and the distribution of the root's forecast (src/mpires.f90:346-454,
// :606-739, :744) are peer stores over NVLink from the engine's own kernels.
int sml_comm_bootstrap(sml_engine *h, sml_allgather_fn allgather, void *ctx)
{
    if (check_ready(h, SML_ATMO)) return -1;
    if (h->p.numprocs == 1) return 0;   // nothing to connect
    if (!allgather) FAIL(h, "sml_comm_bootstrap: the all-gather callback is required");
    const int W = h->p.numprocs;
    struct Hello {
        int32_t rank, world, regions, P, P_ocean, slab, prescribed, pad;
        long long xchg_bytes;
        unsigned char handle[64];
    } mine{};
    mine.rank = h->p.irank; mine.world = W; mine.regions = h->p.number_of_regions; mine.P = h->P_atmo;
    mine.P_ocean = h->P_ocean; mine.slab = h->p.slab_ocean_model_bool; mine.prescribed = h->p.sst_prescribed;
    mine.xchg_bytes = (long long)h->xchg_bytes;
    if (sml_peer_export(h, mine.handle)) return -1;
    std::vector<Hello> all(W);
    if (allgather(ctx, &mine, all.data(), (int)sizeof(Hello)) != 0) FAIL(h, "sml_comm_bootstrap: the host all-gather failed");
    std::vector<unsigned char> handles(64 * (size_t)W);
    for (int k = 0; k < W; ++k) {
        const Hello &o = all[k];
        if (o.rank != k || o.world != W) FAIL(h, "sml_comm_bootstrap: slot %d holds rank %d of %d (the all-gather must be in rank order)", k, o.rank, o.world);
        if (o.regions != mine.regions || o.P != mine.P || o.P_ocean != mine.P_ocean || o.slab != mine.slab ||
            o.prescribed != mine.prescribed || o.xchg_bytes != mine.xchg_bytes)
            FAIL(h, "sml_comm_bootstrap: rank %d was built for a different model configuration", k);
        std::memcpy(&handles[64 * (size_t)k], o.handle, 64);
    }
    // second round: nobody may push into a block before its owner has zeroed and mapped everything; a rank whose
    // attach failed still takes part so that the others do not hang in the all-gather
    int32_t ok = sml_peer_attach(h, handles.data(), W) == 0 ? 1 : 0;
    const std::string attach_err = h->err;
    std::vector<int32_t> oks(W, 0);
    if (allgather(ctx, &ok, oks.data(), (int)sizeof(int32_t)) != 0) FAIL(h, "sml_comm_bootstrap: the host all-gather failed");
    if (!ok) { h->err = attach_err; return -1; }
    for (int k = 0; k < W; ++k)
        if (oks[k] != 1) FAIL(h, "sml_comm_bootstrap: rank %d did not attach", k);
    return 0;
}

int sml_peer_attached(const sml_engine *h) { return h && h->peers.world > 1 ? 1 : 0; }

// a peer that never published its step shows up here (the pack kernel gives up after ~10 s instead of hanging)
int sml_peer_check(sml_engine *h)
{
    if (check_ready(h, SML_ATMO)) return -1;
    int err = 0;
    CK(h, cudaMemcpyAsync(&err, h->d_peer_err, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    if (err == 1) FAIL(h, "peer exchange timed out: a rank did not publish its outvecs");
    if (err) FAIL(h, "peer exchange timed out: the root never published its forecast");
    return 0;
}

int sml_step_predict_ahead(sml_engine *h, int timestep)
{
    if (check_ready(h, SML_ATMO)) return -1;
    CK(h, cudaSetDevice(h->p.device));
    if (!h->overlap) FAIL(h, "sml_step_predict_ahead needs sml_set_overlap(h, 1)");
    if (h->ahead_pending || h->ahead_done) FAIL(h, "overlapped step: the previous look-ahead predict was never consumed");
    if (!h->tisr_fresh) FAIL(h, "overlapped step: call sml_set_tisr with the date's TISR field before the exchange begins");
    KindState &K = h->kinds[SML_ATMO];
    if (h->n_ocean_fb > 0 && timestep < 1) FAIL(h, "timestep must be the 1-based hybrid step");
    CK(h, launch_pdl(k_build_inputs, dim3((unsigned)K.regs.size()), dim3(256), 0, h->stream, K.d_regs, h->d_G, h->d_F, K.d_fb, K.d_lm, 0, 1));
    h->launches++;
    if (h->n_ocean_fb > 0) {
        k_build_ocean_inputs<<<h->n_ocean_fb, 128, 0, h->stream>>>(h->d_ocean_fb, h->d_G, K.d_fb,
                                                                   h->kinds[SML_OCEAN].d_fb, h->d_ocean_ring,
                                                                   (timestep - 1) % h->ocean_slots, h->ocean_slots);
        h->launches++;
    }
    cudaEvent_t *ev = nullptr;
    if (h->profile && h->ev_used < PROFILE_RING) {
        while ((int)h->ev.size() < 3 * (h->ev_used + 1)) {
            cudaEvent_t e;
            CK(h, cudaEventCreate(&e));
            h->ev.push_back(e);
        }
        ev = &h->ev[3 * h->ev_used++];
    }
    if (ev) CK(h, cudaEventRecord(ev[0], h->stream));
    if (launch_step(h, K, K.d_items_split, K.nitems, K.d_fb, K.d_fb_offs, 0, 1)) return -1;
    K.cur ^= 1;
    if (ev) {
        CK(h, cudaEventRecord(ev[1], h->stream));
        CK(h, cudaEventRecord(ev[2], h->stream));
    }
    h->tisr_fresh = false;
    h->ahead_pending = true;
    return 0;
}

// pack (+ the look-ahead predict in the overlapped mode) and, when the caller wants the grids, their copy-out through
// the pinned staging together with the grid-status word.  Returns 1 when the assembled grid holds a non-finite value
// (the grids are delivered all the same), 0 otherwise; without a copy-out nothing is waited for.
static int begin_common(sml_engine *h, int timestep, bool copy_out)
{
    if (sml_step_pack_device(h, timestep)) return -1;
    if (h->overlap) {
        // D2H of the grids on the copy stream, the look-ahead predict on the main stream; the host only
        // waits for the copy
        if (copy_out) {
            CK(h, cudaEventRecord(h->ev_pack, h->stream));
            CK(h, cudaStreamWaitEvent(h->copy_stream, h->ev_pack, 0));
            CK(h, cudaMemcpyAsync(h->h_pin_G, h->d_G, sizeof(double) * G_TISR, cudaMemcpyDeviceToHost, h->copy_stream));
            CK(h, cudaMemcpyAsync(h->h_status, h->d_status, sizeof(int), cudaMemcpyDeviceToHost, h->copy_stream));
            CK(h, cudaEventRecord(h->ev_d2h, h->copy_stream));
        }
        if (sml_step_predict_ahead(h, timestep)) return -1;
        if (copy_out) CK(h, cudaEventSynchronize(h->ev_d2h));
    } else if (copy_out) {
        CK(h, cudaMemcpyAsync(h->h_pin_G, h->d_G, sizeof(double) * G_TISR, cudaMemcpyDeviceToHost, h->stream));
        CK(h, cudaMemcpyAsync(h->h_status, h->d_status, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        CK(h, cudaStreamSynchronize(h->stream));
    }
    return (copy_out && (*h->h_status & SML_GRID_NONFINITE)) ? 1 : 0;
}

int sml_step_exchange_begin(sml_engine *h, int timestep, double *w4d, double *w2d, double *wprecip, double *wsst)
{
    const bool copy_out = w4d || w2d || wprecip || wsst;
    const int rc = begin_common(h, timestep, copy_out);
    if (rc < 0) return rc;
    if (copy_out) {
        if (w4d) std::memcpy(w4d, h->h_pin_G + G_W4D, sizeof(double) * G_W2D);
        if (w2d) std::memcpy(w2d, h->h_pin_G + G_W2D, sizeof(double) * XG * YG);
        if (wprecip) std::memcpy(wprecip, h->h_pin_G + G_PRECIP, sizeof(double) * XG * YG);
        if (wsst) std::memcpy(wsst, h->h_pin_G + G_SST, sizeof(double) * XG * YG);
    }
    return rc;
}

// zero-copy variant for hosts that can consume the grids in place: the same as sml_step_exchange_begin, but instead of
// copying into caller arrays it returns pointers into the engine's pinned staging (valid until the next begin)
int sml_step_exchange_begin_view(sml_engine *h, int timestep, const double **w4d, const double **w2d,
                                 const double **wprecip, const double **wsst)
{
    if (!w4d || !w2d || !wprecip || !wsst) return -1;
    const int rc = begin_common(h, timestep, true);
    if (rc < 0) return rc;
    *w4d = h->h_pin_G + G_W4D;
    *w2d = h->h_pin_G + G_W2D;
    *wprecip = h->h_pin_G + G_PRECIP;
    *wsst = h->h_pin_G + G_SST;
    return rc;
}

// the grids of the last assembly as they stand on the device -- on ANY rank, since every rank rebuilds the whole grid
// (the reference only has them on the root)
int sml_grids_get(sml_engine *h, double *w4d, double *w2d, double *wprecip, double *wsst)
{
    if (check_ready(h, SML_ATMO)) return -1;
    CK(h, cudaSetDevice(h->p.device));
    if (w4d) CK(h, cudaMemcpyAsync(w4d, h->d_G + G_W4D, sizeof(double) * G_W2D, cudaMemcpyDeviceToHost, h->stream));
    if (w2d) CK(h, cudaMemcpyAsync(w2d, h->d_G + G_W2D, sizeof(double) * XG * YG, cudaMemcpyDeviceToHost, h->stream));
    if (wprecip) CK(h, cudaMemcpyAsync(wprecip, h->d_G + G_PRECIP, sizeof(double) * XG * YG, cudaMemcpyDeviceToHost, h->stream));
    if (wsst) CK(h, cudaMemcpyAsync(wsst, h->d_G + G_SST, sizeof(double) * XG * YG, cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    return 0;
}

// ---- failure detection: the grid-status word of k_pack_grids and the run_speedy flag
// sticky bits since the last reset; synchronises the engine stream (after a begin with copy-out it is current anyway)
int sml_grid_status(sml_engine *h, int *bits)
{
    if (check_ready(h, SML_ATMO)) return -1;
    CK(h, cudaSetDevice(h->p.device));
    CK(h, cudaMemcpyAsync(h->h_status, h->d_status, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    if (bits) *bits = *h->h_status;
    return 0;
}
int sml_grid_status_reset(sml_engine *h)
{
    if (check_ready(h, SML_ATMO)) return -1;
    CK(h, cudaSetDevice(h->p.device));
    CK(h, cudaMemsetAsync(h->d_status, 0, sizeof(int), h->stream));
    *h->h_status = 0;
    return 0;
}
// model_parameters%run_speedy: set on the root after run_model (src/mpires.f90:1655-1659), broadcast to every rank
// with the forecast (MPI_Bcast, :744); the step loop leaves when it is false (src/parallelmain.f90:269-271)
int sml_set_run_speedy(sml_engine *h, int run_speedy)
{
    if (!h) return -1;
    h->run_speedy = run_speedy ? 1 : 0;
    return 0;
}
int sml_run_speedy(sml_engine *h, int *run_speedy)
{
    if (check_ready(h, SML_ATMO)) return -1;
    if (!run_speedy) return -1;
    if (h->peers.world > 1 && h->p.irank != 0) {
        // what the root sent with the last forecast this rank has consumed
        CK(h, cudaSetDevice(h->p.device));
        CK(h, cudaMemcpyAsync(h->h_pin_flag, h->d_F + FCST_FLAG, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        CK(h, cudaStreamSynchronize(h->stream));
        *run_speedy = (h->fcst_seq == 0 || *h->h_pin_flag != 0.0) ? 1 : 0;
    } else {
        *run_speedy = h->run_speedy;
    }
    return 0;
}

// the pinned staging sml_step_exchange_end uploads from: a host model that writes its forecast (and the TISR field)
// straight into these arrays and passes the same pointers to sml_step_exchange_end saves the intermediate copy
int sml_forecast_staging(sml_engine *h, double **forecast_4d, double **forecast_2d, double **tisr_grid)
{
    if (check_ready(h, SML_ATMO)) return -1;
    if (forecast_4d) *forecast_4d = h->h_pin_F + F_F4D;
    if (forecast_2d) *forecast_2d = h->h_pin_F + F_F2D;
    if (tisr_grid) *tisr_grid = h->h_pin_F + FCST_TISR;
    return 0;
}

int sml_step_unpack_device(sml_engine *h, int timestep)
{
    if (check_ready(h, SML_ATMO)) return -1;
    CK(h, cudaSetDevice(h->p.device));
    KindState &K = h->kinds[SML_ATMO];
    if (h->overlap && h->ahead_pending) {
        // the forecast is in F: local_model, then v_p = W_out[:, 0:S]*local_model joins the partials
        CK(h, launch_pdl(k_build_inputs, dim3((unsigned)K.regs.size()), dim3(256), 0, h->stream, K.d_regs, h->d_G, h->d_F, K.d_fb,
                         K.d_lm, h->p.ml_only ? 0 : 1, 0));
        h->launches++;
        if (launch_finish(h, K, 1)) return -1;
        h->ahead_pending = false;
        h->ahead_done = true;
        return 0;
    }
    if (h->n_ocean_fb > 0 && timestep < 1)
        FAIL(h, "timestep must be the 1-based hybrid step (it selects the ring slot mod(timestep-1,%d)+1)", h->ocean_slots);
    cudaEvent_t *pe = phase_events(h, h->pt_unpack);
    if (pe) CK(h, cudaEventRecord(pe[0], h->stream));
    CK(h, launch_pdl(k_build_inputs, dim3((unsigned)K.regs.size()), dim3(256), 0, h->stream, K.d_regs, h->d_G, h->d_F, K.d_fb, K.d_lm,
                     h->p.ml_only ? 0 : 1, 1));
    if (pe) CK(h, cudaEventRecord(pe[1], h->stream));
    h->launches++;
    if (h->n_ocean_fb > 0) {
        k_build_ocean_inputs<<<h->n_ocean_fb, 128, 0, h->stream>>>(h->d_ocean_fb, h->d_G, K.d_fb,
                                                                   h->kinds[SML_OCEAN].d_fb, h->d_ocean_ring,
                                                                   (timestep - 1) % h->ocean_slots, h->ocean_slots);
        h->launches++;
    }
    CK(h, cudaGetLastError());
    return 0;
}

int sml_step_exchange_end(sml_engine *h, int timestep, const double *f4d, const double *f2d, const double *tisr)
{
    if (check_ready(h, SML_ATMO)) return -1;
    CK(h, cudaSetDevice(h->p.device));
    const bool ahead = h->overlap && h->ahead_pending;
    const bool peers = h->peers.world > 1;
    const bool root = h->p.irank == 0;
    if (h->p.numprocs > 1 && !peers && !root)
        FAIL(h, "sml_step_exchange_end on rank %d needs sml_comm_bootstrap (or the device-only pieces with a host collective)", h->p.irank);
    if (peers && !root) {
        // ---- consumer rank: the forecast block comes from the root over NVLink (scatter + bcast of
        // src/mpires.f90:606-744); wait for it on the device -- the host never blocks here -- and rebuild the inputs
        const unsigned long long seq = ++h->fcst_seq;
        const bool own_tisr = tisr != nullptr && !ahead;
        if (own_tisr) {   // every rank holds the TISR table (get_tisr_by_date, :751-753): its own copy goes up
            CK(h, cudaStreamSynchronize(h->stream));   // the pinned staging may still feed the previous upload
            std::memcpy(h->h_pin_F + FCST_TISR, tisr, sizeof(double) * XG * YG);
            CK(h, cudaMemcpyAsync(h->d_G + G_TISR, h->h_pin_F + FCST_TISR, sizeof(double) * XG * YG, cudaMemcpyHostToDevice, h->stream));
        }
        const bool take_tisr = !own_tisr && !ahead;   // else: the root's field, which travels with the forecast
        k_wait_flag<<<1, 256, 0, h->stream>>>(h->peers.flag(h->p.irank, XF_FCST, 0), seq, h->d_peer_err,
                                              take_tisr ? h->d_F + FCST_TISR : nullptr, h->d_G + G_TISR, take_tisr ? XG * YG : 0);
        h->launches++;
        CK(h, cudaGetLastError());
        if (sml_step_unpack_device(h, timestep)) return -1;
        // bounded run-ahead: sleep (blocking event, no spinning) until the step enqueued two calls ago is done
        const int slot = (int)(h->consumer_steps & 3);
        if (!h->ev_ahead[slot]) CK(h, cudaEventCreateWithFlags(&h->ev_ahead[slot], cudaEventDisableTiming | cudaEventBlockingSync));
        CK(h, cudaEventRecord(h->ev_ahead[slot], h->stream));
        ++h->consumer_steps;
        const int old_slot = (int)((h->consumer_steps + 1) & 3);   // recorded three calls ago
        if (h->consumer_steps >= 3 && h->ev_ahead[old_slot]) CK(h, cudaEventSynchronize(h->ev_ahead[old_slot]));
        return 0;
    }
    if (!tisr && !ahead) FAIL(h, "tisr_grid is required");
    // the forecast goes up on the copy stream in the overlapped mode (the main stream is busy with the look-ahead
    // predict); its pinned staging is free again: the last copy from it finished before this step's pack ran
    cudaStream_t up = ahead ? h->copy_stream : h->stream;
    if (!h->p.ml_only) {
        if (!f4d || !f2d) FAIL(h, "hybrid mode needs forecast_4d and forecast_2d");
        if (f4d != h->h_pin_F + F_F4D) std::memcpy(h->h_pin_F + F_F4D, f4d, sizeof(double) * G_W2D);
        if (f2d != h->h_pin_F + F_F2D) std::memcpy(h->h_pin_F + F_F2D, f2d, sizeof(double) * XG * YG);
    }
    if (tisr && tisr != h->h_pin_F + FCST_TISR) std::memcpy(h->h_pin_F + FCST_TISR, tisr, sizeof(double) * XG * YG);
    if (peers) {
        // ---- root of a multi-rank run: ONE upload of [forecast | tisr | run_speedy] into the landing section of its
        // own exchange block, then one kernel pushes the block into every other rank's landing section and publishes
        // the forecast sequence number (no NCCL broadcast on the critical path)
        h->h_pin_F[FCST_FLAG] = (double)h->run_speedy;
        CK(h, cudaMemcpyAsync(h->d_F, h->h_pin_F, sizeof(double) * FCST_DOUBLES, cudaMemcpyHostToDevice, up));
        if (tisr && !ahead)
            CK(h, cudaMemcpyAsync(h->d_G + G_TISR, h->d_F + FCST_TISR, sizeof(double) * XG * YG, cudaMemcpyDeviceToDevice, up));
        const unsigned long long seq = ++h->fcst_seq;
        k_peer_push<<<dim3(16, h->peers.world), 256, 0, up>>>(h->d_F, FCST_DOUBLES, h->peers, XF_FCST, 0, 1, seq, 0, h->d_done + 2);
        h->launches++;
        CK(h, cudaGetLastError());
    } else {
        if (!h->p.ml_only) CK(h, cudaMemcpyAsync(h->d_F, h->h_pin_F, sizeof(double) * F_TOTAL, cudaMemcpyHostToDevice, up));
        if (!ahead)
            CK(h, cudaMemcpyAsync(h->d_G + G_TISR, h->h_pin_F + FCST_TISR, sizeof(double) * XG * YG, cudaMemcpyHostToDevice, up));
    }
    if (ahead) {
        CK(h, cudaEventRecord(h->ev_h2d, h->copy_stream));
        CK(h, cudaStreamWaitEvent(h->stream, h->ev_h2d, 0));
        return sml_step_unpack_device(h, timestep);
    }
    if (sml_step_unpack_device(h, timestep)) return -1;
    CK(h, cudaStreamSynchronize(h->stream));  // pinned staging is reused by the next call
    return 0;
}

// model_parameters%outvec_component_contribs (src/mod_reservoir.f90:1458-1461): keep v_p = wout(:,1:S)*local_model and
// v_ml = wout(:,S+1:)*x~ of every local atmosphere region (standardised units, as the reference stores them).  The
// readout then runs in split order, outvec = unstandardise(v_p + v_ml): equal to the fused order within 1e-13.
int sml_set_contribs(sml_engine *h, int on)
{
    if (check_ready(h, SML_ATMO)) return -1;
    CK(h, cudaSetDevice(h->p.device));
    KindState &K = h->kinds[SML_ATMO];
    if (on && !h->d_vp) {
        const size_t bytes = sizeof(double) * K.regs.size() * K.P;
        CK(h, cudaMalloc(&h->d_vp, bytes));
        CK(h, cudaMalloc(&h->d_vml, bytes));
        CK(h, cudaMemset(h->d_vp, 0, bytes));
        CK(h, cudaMemset(h->d_vml, 0, bytes));
    }
    h->contribs = on != 0;
    return 0;
}

int sml_contribs_get(sml_engine *h, int region, double *v_p, double *v_ml)
{
    if (check_ready(h, SML_ATMO)) return -1;
    if (!h->contribs) FAIL(h, "sml_set_contribs(h, 1) has not been called");
    int li;
    if (local_of(h, SML_ATMO, region, &li)) return -1;
    const RegionDev &d = h->kinds[SML_ATMO].regs[li].dev;
    CK(h, cudaMemcpyAsync(v_p, h->d_vp + d.out_off, sizeof(double) * d.P, cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaMemcpyAsync(v_ml, h->d_vml + d.out_off, sizeof(double) * d.P, cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    return 0;
}

int sml_set_overlap(sml_engine *h, int on)
{
    if (check_ready(h, SML_ATMO)) return -1;
    if (h->ahead_pending) FAIL(h, "cannot switch modes between sml_step_exchange_begin and _end");
    CK(h, cudaStreamSynchronize(h->stream));
    h->overlap = on != 0;
    h->ahead_done = false;
    h->tisr_fresh = false;
    return 0;
}

// get_tisr_by_date's field for the step being exchanged (src/mpires.f90:751-753, 1676-1708): in the overlapped mode
// the feedback is rebuilt before the host model returns, so TISR has to be on the device when the exchange begins
int sml_set_tisr(sml_engine *h, const double *tisr)
{
    if (check_ready(h, SML_ATMO)) return -1;
    if (!tisr) FAIL(h, "tisr_grid is required");
    CK(h, cudaSetDevice(h->p.device));
    std::memcpy(h->h_pin_tisr, tisr, sizeof(double) * XG * YG);
    CK(h, cudaMemcpyAsync(h->d_G + G_TISR, h->h_pin_tisr, sizeof(double) * XG * YG, cudaMemcpyHostToDevice, h->stream));
    h->tisr_fresh = true;
    return 0;
}

/* ------------------------------------------------------------------ reservoir construction (gen_res) */
// sparse_eigen for every local region of the kind: power iteration until the estimate moves by less than tol
// (relative) in every region, at most maxit iterations.  eigs[nloc] in local order (0 where there is no reservoir).
int sml_sparse_eigen(sml_engine *h, int kind, int maxit, double tol, double *eigs, int *iterations)
{
    if (check_ready(h, kind)) return -1;
    if (maxit < 1 || !eigs) FAIL(h, "sml_sparse_eigen: bad arguments");
    CK(h, cudaSetDevice(h->p.device));
    KindState &K = h->kinds[kind];
    const int nloc = (int)K.regs.size();
    const int nblk = (K.n_max + GR_BLOCK - 1) / GR_BLOCK;
    double *x = nullptr, *y = nullptr, *part = nullptr, *lam = nullptr;
    CK(h, cudaMalloc(&x, sizeof(double) * std::max<long long>(1, K.x_total)));
    CK(h, cudaMalloc(&y, sizeof(double) * std::max<long long>(1, K.x_total)));
    CK(h, cudaMalloc(&part, sizeof(double) * (size_t)nloc * nblk));
    CK(h, cudaMalloc(&lam, sizeof(double) * 2 * nloc));
    CK(h, cudaMemsetAsync(lam, 0, sizeof(double) * 2 * nloc, h->stream));
    k_eig_init<<<nloc, 256, 0, h->stream>>>(K.d_regs, x);
    h->launches++;
    std::vector<double> hl(2 * nloc, 0.0);
    int it = 0;
    bool done = false;
    while (it < maxit && !done) {
        const int burst = std::min(8, maxit - it);  // check convergence every few iterations (one D2H each)
        for (int b = 0; b < burst; ++b) {
            k_eig_spmv<<<dim3(nblk, nloc), GR_BLOCK, 0, h->stream>>>(K.d_regs, x, y, part, nblk);
            k_eig_normalize<<<nloc, 256, 0, h->stream>>>(K.d_regs, y, x, part, nblk, lam);
            h->launches += 2;
        }
        it += burst;
        CK(h, cudaMemcpyAsync(hl.data(), lam, sizeof(double) * 2 * nloc, cudaMemcpyDeviceToHost, h->stream));
        CK(h, cudaStreamSynchronize(h->stream));
        done = true;
        for (int i = 0; i < nloc; ++i)
            if (K.regs[i].uploaded && std::fabs(hl[2 * i] - hl[2 * i + 1]) > tol * std::fabs(hl[2 * i])) done = false;
    }
    for (int i = 0; i < nloc; ++i) eigs[i] = K.regs[i].uploaded ? hl[2 * i] : 0.0;
    if (iterations) *iterations = it;
    cudaFree(x); cudaFree(y); cudaFree(part); cudaFree(lam);
    CK(h, cudaGetLastError());
    return done ? 0 : 1;  // 1: not converged within maxit (eigs hold the last estimates)
}

// reservoir%vals = (reservoir%vals / eigs) * radius on the device copy (src/mod_reservoir.f90:193-195): the
// caller passes factor[i] = radius / eigs[i] and applies the same factor to its host copy of vals
int sml_adjacency_scale(sml_engine *h, int kind, const double *factor)
{
    if (check_ready(h, kind)) return -1;
    if (!factor) FAIL(h, "sml_adjacency_scale: factor is required");
    CK(h, cudaSetDevice(h->p.device));
    KindState &K = h->kinds[kind];
    const int nloc = (int)K.regs.size();
    double *df = nullptr;
    CK(h, cudaMalloc(&df, sizeof(double) * nloc));
    CK(h, cudaMemcpyAsync(df, factor, sizeof(double) * nloc, cudaMemcpyHostToDevice, h->stream));
    k_adj_scale<<<dim3(32, nloc), 256, 0, h->stream>>>(K.d_regs, df);
    h->launches++;
    K.pack_valid = false;   // the spin-up kernel's tile-major copy holds the old values
    if (K.d_gen) {   // generated regions keep their COO on the device: reservoir%vals is rescaled there as well
        k_coo_scale<<<dim3(32, nloc), 256, 0, h->stream>>>(K.d_gen, K.d_gen_of_local, df);
        h->launches++;
    }
    CK(h, cudaStreamSynchronize(h->stream));
    cudaFree(df);
    CK(h, cudaGetLastError());
    return 0;
}

// what makesparse left in reservoir%rows / cols / vals (and, after sml_adjacency_scale, gen_res's rescaled vals) for a
// region constructed by sml_region_generate: k entries each, 1-based, in makesparse's entry order
int sml_region_coo_get(sml_engine *h, int kind, int region, int32_t *rows, int32_t *cols, double *vals)
{
    if (check_ready(h, kind)) return -1;
    int li;
    if (local_of(h, kind, region, &li)) return -1;
    KindState &K = h->kinds[kind];
    const HostRegion &hr = K.regs[li];
    if (hr.gen_index < 0) FAIL(h, "region %d was uploaded, not generated: the host already holds its COO arrays", region);
    const GenDesc &g = K.gen[hr.gen_index];
    CK(h, cudaSetDevice(h->p.device));
    if (rows) CK(h, cudaMemcpyAsync(rows, g.coo_rows, sizeof(int) * (size_t)g.k, cudaMemcpyDeviceToHost, h->stream));
    if (cols) CK(h, cudaMemcpyAsync(cols, g.coo_cols, sizeof(int) * (size_t)g.k, cudaMemcpyDeviceToHost, h->stream));
    if (vals) CK(h, cudaMemcpyAsync(vals, g.coo_vals, sizeof(double) * (size_t)g.k, cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    return 0;
}

// reservoir%win in the one-non-zero-per-row form: value and 0-based column of every row
int sml_region_win_get(sml_engine *h, int kind, int region, double *win_compact, int32_t *win_col)
{
    if (check_ready(h, kind)) return -1;
    int li;
    if (local_of(h, kind, region, &li)) return -1;
    const RegionDev &d = h->kinds[kind].regs[li].dev;
    if (d.win_mode != 0) FAIL(h, "region %d keeps a dense W_in", region);
    CK(h, cudaSetDevice(h->p.device));
    if (win_compact) CK(h, cudaMemcpyAsync(win_compact, d.winc, sizeof(double) * (size_t)d.n, cudaMemcpyDeviceToHost, h->stream));
    if (win_col) CK(h, cudaMemcpyAsync(win_col, d.wcol, sizeof(int) * (size_t)d.n, cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    return 0;
}

/* ------------------------------------------------------------------ measurement */
int sml_profile(sml_engine *h, int on)
{
    if (!h) return -1;
    h->profile = on != 0;
    h->ev_used = 0;
    h->pt_pack.used = 0;
    h->pt_unpack.used = 0;
    return 0;
}
// CUDA-event time of the pack (grid assembly, incl. the wait for the peers' outvecs) and unpack (feedback rebuild)
// kernels since the last call, for the share-of-step table in profiles/
int sml_phase_times(sml_engine *h, double *pack_ms_sum, double *unpack_ms_sum, int *count)
{
    if (!h) return -1;
    *pack_ms_sum = *unpack_ms_sum = 0.0;
    *count = std::min(h->pt_pack.used, h->pt_unpack.used);
    CK(h, cudaStreamSynchronize(h->stream));
    for (int i = 0; i < h->pt_pack.used; ++i) {
        float a = 0.f;
        CK(h, cudaEventElapsedTime(&a, h->pt_pack.ev[2 * i], h->pt_pack.ev[2 * i + 1]));
        *pack_ms_sum += a;
    }
    for (int i = 0; i < h->pt_unpack.used; ++i) {
        float a = 0.f;
        CK(h, cudaEventElapsedTime(&a, h->pt_unpack.ev[2 * i], h->pt_unpack.ev[2 * i + 1]));
        *unpack_ms_sum += a;
    }
    h->pt_pack.used = 0;
    h->pt_unpack.used = 0;
    return 0;
}
int sml_kernel_times(sml_engine *h, double *step_ms_sum, double *finish_ms_sum, int *count)
{
    if (!h) return -1;
    *step_ms_sum = 0.0;
    *finish_ms_sum = 0.0;
    *count = h->ev_used;
    if (h->ev_used == 0) return 0;
    CK(h, cudaEventSynchronize(h->ev[3 * (h->ev_used - 1) + 2]));
    for (int i = 0; i < h->ev_used; ++i) {
        float a = 0.f, b = 0.f;
        CK(h, cudaEventElapsedTime(&a, h->ev[3 * i], h->ev[3 * i + 1]));
        CK(h, cudaEventElapsedTime(&b, h->ev[3 * i + 1], h->ev[3 * i + 2]));
        *step_ms_sum += a;
        *finish_ms_sum += b;
    }
    h->ev_used = 0;
    return 0;
}
// CUDA-event time of the update-only step launches issued by sml_synchronize while profiling, and their count
int sml_sync_times(sml_engine *h, double *ms_sum, int64_t *steps)
{
    if (!h) return -1;
    *ms_sum = h->sync_ms_sum;
    *steps = h->sync_steps;
    h->sync_ms_sum = 0.0;
    h->sync_steps = 0;
    return 0;
}
int64_t sml_kernel_launch_count(const sml_engine *h) { return h ? h->launches : 0; }
int sml_step_plan(const sml_engine *h, int kind, int *kernel, int *slots, int *part_rows, int *parts)
{
    if (!h || kind < 0 || kind > 1) return -1;
    const KindState &K = h->kinds[kind];
    if (kernel) *kernel = K.persist ? 1 : 0;
    if (slots) *slots = K.persist ? K.nslots : K.nitems;
    if (part_rows) *part_rows = K.persist ? K.part_rows : K.chunk_rows;
    if (parts) *parts = K.persist ? K.nparts : K.nitems;
    return 0;
}
int sml_step_l2_keep(const sml_engine *h, int kind) { return (h && kind >= 0 && kind < 2 && h->kinds[kind].l2_keep) ? 1 : 0; }
int sml_setup_stats(const sml_engine *h, double *upload_seconds, int64_t *arena_bytes, int *arena_chunks)
{
    if (!h) return -1;
    if (upload_seconds) *upload_seconds = h->setup_upload_s;
    if (arena_bytes) *arena_bytes = (int64_t)h->arena.total;
    if (arena_chunks) *arena_chunks = (int)h->arena.chunks.size();
    return 0;
}
int sml_step_chunk_rows(const sml_engine *h, int kind) { return (h && kind >= 0 && kind < 2) ? h->kinds[kind].chunk_rows : 0; }
int64_t sml_predict_algorithmic_bytes(const sml_engine *h, int kind)
{
    if (!h || kind < 0 || kind > 1) return 0;
    return h->kinds[kind].alg_bytes;
}

int64_t sml_update_algorithmic_bytes(const sml_engine *h, int kind)
{
    if (!h || kind < 0 || kind > 1) return -1;
    return h->kinds[kind].alg_bytes_update;
}

}  // extern "C"

#include "train_api.inl"
