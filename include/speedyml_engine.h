/*
 * speedyml_engine.h -- C ABI of the B200-native SPEEDY-ML local-reservoir engine.
 *
 * The reference has no FFI: its hot path is reached through Fortran module procedures on derived
 * types (mod_reservoir / mod_linalg / mpires / resdomain).  Each entry point below names the
 * reference procedure it replaces (paths relative to the reference tree); the thin ISO_C_BINDING
 * layer that keeps those Fortran names and forwards here is speedy-ml_b200/fortran/speedyml_gpu.f90,
 * and INTEGRATION.md shows where the reference calls change.
 *
 * Conventions (kept from the reference):
 *   - all floating point data is FP64, column-major, caller-owned; the engine copies;
 *   - COO indices are 1-based int32 exactly as mklsparse receives them; region ids are 0-based
 *     (reservoir%assigned_region); every other index this API returns is 1-based like res_domain.f90;
 *   - one host thread calls in program order (as each MPI rank does); calls block until results the
 *     caller can read are in its arrays;
 *   - return value: 0 ok, <0 engine error (sml_last_error), >0 LAPACK-style info where stated.
 *     The library never aborts and has NO CPU fallback: without a usable CUDA device sml_create fails.
 *
 * Scope: num_vert_levels == 1 (the reference's configuration, src/mod_reservoir.f90:57).
 */
#ifndef SPEEDYML_ENGINE_H
#define SPEEDYML_ENGINE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SML_XGRID 96
#define SML_YGRID 48
#define SML_ZGRID 8

/* reservoir kinds: res%reservoir(i,1) and res%reservoir_special(i,1) (src/parallelmain.f90:56-61) */
#define SML_ATMO 0
#define SML_OCEAN 1

#define SML_ALL_REGIONS (-1)

/* model_parameters_type subset (src/mod_utilities.f90:371-509; values src/mod_reservoir.f90:12-77) */
typedef struct sml_params {
    int32_t number_of_regions;     /* 1152 */
    int32_t overlap;               /* 1 */
    int32_t precip_bool;           /* 1 */
    int32_t slab_ocean_model_bool; /* 1: SST grid is exchanged and fed back */
    int32_t ml_only;               /* 0: hybrid (chunk_size_speedy > 0) */
    int32_t irank, numprocs;       /* region sharding: processor_decomposition(irank, numprocs) */
    int32_t device;                /* CUDA device ordinal for this rank */
    int32_t timestep, timestep_slab; /* hours: 6, 168 */
    int32_t sst_prescribed;        /* 1: wholegrid_sst comes from the host every step (no ocean reservoirs) */
    int32_t reserved[5];
} sml_params;

/* what read_trained_res / allocate_res_new hold for one reservoir (src/mod_io.f90:2938-2983,
 * src/mod_reservoir.f90:80-180).  Exactly one of win_dense / win_compact must be non-NULL. */
typedef struct sml_region_weights {
    int32_t region;          /* reservoir%assigned_region, 0-based */
    int32_t kind;            /* SML_ATMO / SML_OCEAN */
    int32_t n, k;            /* reservoir%n, reservoir%k */
    int32_t D, P, S;         /* reservoir_numinputs, chunk_size_prediction, chunk_size_speedy */
    int32_t L;               /* length of mean/std */
    int32_t sst_bool_input;  /* atmosphere: SST slot present in the input vector */
    int32_t reserved0;
    double leakage;          /* reservoir%leakage */
    double sst_mean, sst_std;/* grid_special%mean/std(sst_mean_std_idx) used for the SST feedback slot */
    const int32_t *rows;     /* [k] 1-based */
    const int32_t *cols;     /* [k] 1-based */
    const double *vals;      /* [k] */
    const double *win_dense; /* [n*D] column-major win(n,D), or NULL */
    const double *win_compact; /* [n]: the single non-zero of each row, or NULL */
    const int32_t *win_col;  /* [n] 0-based column of that non-zero (with win_compact) */
    const double *wout;      /* [P*(n+S)] column-major wout(P,n+S); NULL = zeros (to be trained) */
    const double *mean;      /* [L] */
    const double *std;       /* [L] */
} sml_region_weights;

typedef struct sml_engine sml_engine;

/* ---- life cycle ---- */
int sml_create(sml_engine **h, const sml_params *p);
int sml_destroy(sml_engine *h);
const char *sml_last_error(const sml_engine *h); /* also valid with h == NULL for sml_create failures */
int sml_set_stream(sml_engine *h, void *cuda_stream); /* run on the caller's stream (NULL: engine's own) */
int sml_synchronize_stream(sml_engine *h);
int sml_num_local_regions(const sml_engine *h);
int sml_local_region_ids(const sml_engine *h, int32_t *ids); /* processor_decomposition, src/res_domain.f90:31-62 */

/* ---- res_domain.f90 index arithmetic (host, integers; bit-exact contract) ---- */
int sml_domaindecomposition(int numregions, int *factorx, int *factory);           /* :258-280 */
int sml_getxyresextent(int num_regions, int region, int *xs, int *xe, int *ys, int *ye,
                       int *xchunk, int *ychunk);                                  /* :123-141 */
int sml_getoverlapindices(int num_regions, int region, int overlap, int *ixs, int *ixe, int *iys,
                          int *iye, int *ixc, int *iyc, int *pole, int *periodic);  /* :155-204 */
int sml_get_trainingdataindices(int num_regions, int region, int overlap, int *xs, int *xe,
                                int *ys, int *ye);                                 /* :547-574 */
int sml_processor_decomposition(int irank, int numprocs, int number_of_regions,
                                int32_t *region_indices, int *count);              /* :31-62 */
/* sizes of allocate_res_new (src/mod_reservoir.f90:155-173) for an atmosphere reservoir */
int sml_region_dims(int num_regions, int region, int overlap, int m, double deg, int precip_bool,
                    int sst_bool, int sst_bool_input, int ml_only, int *n, int *k, int *D, int *P,
                    int *S, int *L);
/* sizes of initialize_slab_ocean_model (src/mod_slab_ocean_reservoir.f90:9-133) for the ocean reservoir of a
 * region (m = 4000, deg = 6 there): n, k, D = reservoir_numinputs, P = chunk_size_prediction (S is always 0) and
 * A = logp_end, the length of the time-averaged atmosphere part of its input vector */
int sml_ocean_region_dims(int num_regions, int region, int overlap, int m, double deg, int *n, int *k,
                          int *D, int *P, int *A);
/* ocean maps: sst_map[D/8] offsets in G of the halo'd SST tile (tile_4d_and_logp_to_local_state_input_slab,
 * src/res_domain.f90:1127-1152); target_map[P] rows of the ocean input vector that form the target
 * (tile_full_input_to_target_data2d_ocean_model :691-728); atmo_slice0 = 0-based start of
 * atmo_training_data_idx in the atmosphere input vector (src/mod_slab_ocean_reservoir.f90:1621-1625) */
int sml_ocean_region_maps(int num_regions, int region, int overlap, int32_t *sst_map, int32_t *target_map,
                          int *atmo_slice0);
/* flattened 0-based gather/scatter maps into the global buffers (layout: sml_global_layout):
 *   input_map[D]  : feedback element -> offset in G  (tile_4d_and_logp_to_local_state_input :1081-1125,
 *                   tileoverlapgrid2d for sst/tisr)         input_ms[D] : 0-based mean/std slot, L = sst
 *   output_map[P] : outvec element   -> offset in G  (tile_full_grid_with_local_state_vec_res1d :791-826)
 *   model_map[S]  : local_model elem -> offset in F  (tile_4d_and_logp_full_grid_to_local_res_vec :1022-1053)
 *   target_map[P] : target row -> input-vector row   (tile_full_input_to_target_data2d :602-651) */
int sml_region_maps(int num_regions, int region, int overlap, int precip_bool, int sst_bool_input,
                    int32_t *input_map, int32_t *input_ms, int32_t *output_map, int32_t *output_ms,
                    int32_t *model_map, int32_t *model_ms, int32_t *target_map);
/* vertical localisation, num_vert_levels in {1, 2, 4, 8}: get_z_res_extent (src/res_domain.f90:143-153),
 * getoverlapindices_vert (:206-256), get_trainingdataindices_vert (:576-600) and the sizes / flattened maps of ONE
 * vertical slab (only the bottom slab carries logp, precip and SST, src/mod_reservoir.f90:1793-1818).  Host integer
 * arithmetic, bit-exact against the oracle.  The engine's kernels step num_vert_levels == 1, the reference's
 * configuration (src/mod_reservoir.f90:57); these entries are what a multi-level host would build its tables from. */
int sml_get_z_res_extent(int num_vert_levels, int vert_level, int *zs, int *ze, int *zchunk);
int sml_getoverlapindices_vert(int num_vert_levels, int vert_level, int vert_overlap, int *izs, int *ize, int *izc,
                               int *top, int *bottom);
int sml_get_trainingdataindices_vert(int num_vert_levels, int vert_level, int vert_overlap, int *zs, int *ze);
int sml_region_dims_vert(int num_regions, int region, int overlap, int num_vert_levels, int vert_level, int vert_overlap,
                         int m, double deg, int precip_bool, int sst_bool, int sst_bool_input, int ml_only, int *n, int *k,
                         int *D, int *P, int *S, int *L);
int sml_region_maps_vert(int num_regions, int region, int overlap, int num_vert_levels, int vert_level, int vert_overlap,
                         int precip_bool, int sst_bool_input, int32_t *input_map, int32_t *input_ms, int32_t *output_map,
                         int32_t *output_ms, int32_t *model_map, int32_t *model_ms, int32_t *target_map);
/* offsets (in doubles) of wholegrid4d, wholegrid2d, wholegrid_precip, wholegrid_sst, tisr in G and the
 * total; F = [forecast_4d | forecast_2d] */
int sml_global_layout(int64_t off[5], int64_t *g_total, int64_t *f_total);

/* ---- weights: mklsparse (src/mod_linalg.f90:10-25) + trained_reservoir_prediction
 *      (src/mod_reservoir.f90:1783-1886); slab twin src/mod_slab_ocean_reservoir.f90:1561-1652 ---- */
int sml_region_upload(sml_engine *h, const sml_region_weights *w);
/* the same straight from the file write_trained_res produces (src/mod_reservoir.f90:1703-1738; read_trained_res
 * src/mod_io.f90:2938-2983): one NetCDF-classic container per region with win, wout (float), rows, cols (int), vals,
 * mean, std (float), read without any NetCDF library and widened to FP64.  sml_trained_res_dims only reads the header
 * (errors: sml_last_error(NULL)). */
int sml_trained_res_dims(const char *path, int *n, int *k, int *D, int *P, int *S, int *L);
int sml_region_upload_file(sml_engine *h, const char *path, int region, int kind, int sst_bool_input, double leakage);
int sml_finalize(sml_engine *h); /* after the last upload: builds the batched step plan */

/* ---- reservoir construction on the device (gen_res, src/mod_reservoir.f90:182-212): after the unscaled adjacency
 * of every region has been uploaded (makesparse -> mklsparse, src/mod_linalg.f90:180-218) and sml_finalize,
 *   sml_sparse_eigen   replaces sparse_eigen (ARPACK dnaupd/dneupd, src/mod_linalg.f90:220-514): the largest-
 *                      magnitude eigenvalue of every local adjacency -- the Perron root of a non-negative matrix --
 *                      by batched power iteration; eigs[nloc] in local order; returns 1 if maxit was reached
 *   sml_adjacency_scale applies vals *= factor[i] (factor = radius / eig, :193-195) to the device copy; the host
 *                      applies the same factor to reservoir%vals ---- */
int sml_sparse_eigen(sml_engine *h, int kind, int maxit, double tol, double *eigs, int *iterations);
/* the rest of gen_res on the device: makesparse (src/mod_linalg.f90:180-218: vals = random_number, rows / cols = rounds of
 * the k-shuffle, src/mod_utilities.f90:1569-1596) and the W_in build of train_reservoir (src/mod_reservoir.f90:262-283:
 * rows (i-1)q+1..iq of column i = sigma*(-1 + 2*rand), q = n / reservoir_numinputs).  Call INSTEAD of sml_region_upload
 * with the sizes of allocate_res_new in w (rows, cols, vals and the win_* pointers are ignored; wout may be NULL);
 * sml_finalize then constructs every such region in two launches, after which sml_sparse_eigen / sml_adjacency_scale
 * complete gen_res.  The draws come from the engine's counter-based generator keyed by (seed, region, stream, index) --
 * reproducible, restated in oracle/ from the same streams, but not the Fortran random_number stream.
 * sml_region_coo_get returns reservoir%rows / cols / vals (1-based, makesparse's entry order; vals rescaled once
 * sml_adjacency_scale has run), sml_region_win_get the W_in value and 0-based column of every row -- what
 * write_trained_res needs on the host. */
int sml_region_generate(sml_engine *h, const sml_region_weights *w, unsigned long long seed, double sigma);
int sml_region_coo_get(sml_engine *h, int kind, int region, int32_t *rows, int32_t *cols, double *vals);
int sml_region_win_get(sml_engine *h, int kind, int region, double *win_compact, int32_t *win_col);
int sml_adjacency_scale(sml_engine *h, int kind, const double *factor);

/* ---- per-region state (reservoir%current_state / saved_state / feedback / local_model / outvec) ---- */
int sml_state_set(sml_engine *h, int kind, int region, const double *x);
int sml_state_get(sml_engine *h, int kind, int region, double *x);
int sml_feedback_set(sml_engine *h, int kind, int region, const double *feedback);
int sml_feedback_get(sml_engine *h, int kind, int region, double *feedback);
int sml_local_model_set(sml_engine *h, int kind, int region, const double *local_model);
int sml_local_model_get(sml_engine *h, int kind, int region, double *local_model);
int sml_outvec_get(sml_engine *h, int kind, int region, double *outvec);
/* start_prediction_slab seeds reservoir%outvec with the last observed SST/OHTC tile
 * (src/mod_slab_ocean_reservoir.f90:853-859); the exchange uses it until the first ocean step */
int sml_outvec_set(sml_engine *h, int kind, int region, const double *outvec);
/* all local outvecs of a kind in one copy: slab[i*chunk_size_prediction + p], i = local region order */
int sml_outvec_get_all(sml_engine *h, int kind, double *slab);
int sml_wout_get(sml_engine *h, int kind, int region, double *wout);
int sml_wout_set(sml_engine *h, int kind, int region, const double *wout);

/* ---- synchronize (src/mod_reservoir.f90:1354-1381, synchronize_print :1383-1416; slab :1237-1266)
 * input(ld, length) column-major for ONE region, or with region == SML_ALL_REGIONS the local regions'
 * series back to back (region i's block starts at inputs + offsets[i], leading dimension D_i). ---- */
int sml_synchronize(sml_engine *h, int kind, int region, const double *inputs, int ld, int length,
                    const int64_t *offsets);

/* ---- predict / predict_ml (src/mod_reservoir.f90:1418-1535) for every local region of the kind: state
 * update + readout + un-standardise; outvec stays on the device.  kind == SML_OCEAN is predict_slab_ml
 * (src/mod_slab_ocean_reservoir.f90:1318-1363; every output * std(sst) + mean(sst)); the caller decides when
 * (mod(t*timestep, timestep_slab) == 0, src/parallelmain.f90:238).  The ocean feedback vector is rebuilt by
 * every sml_step_exchange_end / sml_step_unpack_device(timestep): atmosphere part = mean of the
 * timestep_slab/timestep-1 slot ring (slot mod(timestep-1, slots)+1), SST part = standardised halo tile of
 * wholegrid_sst, TISR and OHTC parts as sml_feedback_set left them (src/mpires.f90:594-600,776-781; the
 * intended semantics of SURVEY.md Appendix C, where the source itself has shape hazards) ---- */
int sml_predict(sml_engine *h, int kind);

/* model_parameters%outvec_component_contribs (src/mod_reservoir.f90:1458-1461): when on, every atmosphere predict also
 * keeps reservoir%v_p = wout(:,1:chunk_size_speedy)*local_model and reservoir%v_ml = wout(:,chunk_size_speedy+1:)*x~
 * (standardised units).  The readout then runs in split order (outvec = unstandardise(v_p + v_ml)). */
int sml_set_contribs(sml_engine *h, int on);
int sml_contribs_get(sml_engine *h, int region, double *v_p, double *v_ml);

/* ---- sendrecievegrid (src/mpires.f90:218-804), split where the root calls run_model (:565-569).
 * begin: gathers the outvecs into the global grids, applies the clamps (:456-490) and returns them
 *        (rank 0 arrays may be NULL to skip the copy-out): wholegrid4d[4*96*48*8], wholegrid2d[96*48],
 *        wholegrid_precip[96*48], wholegrid_sst[96*48].
 * end:   takes run_model's forecast (forecast_4d, forecast_2d), the date's global TISR field
 *        (get_tisr_by_date :1676-1708) and rebuilds every local region's feedback and local_model
 *        (:581-604, :749-791). ---- */
int sml_step_exchange_begin(sml_engine *h, int timestep, double *wholegrid4d, double *wholegrid2d,
                            double *wholegrid_precip, double *wholegrid_sst);
int sml_step_exchange_end(sml_engine *h, int timestep, const double *forecast_4d,
                          const double *forecast_2d, const double *tisr_grid);
/* ---- overlapped step (optional).  feedback(t) depends only on the gathered outvec grid,
 * TISR and SST -- not on run_model's forecast -- and local_model(t) enters only the first S columns of the next
 * readout (the reference computes that split itself as v_p / v_ml, src/mod_reservoir.f90:1458-1461).  With
 * sml_set_overlap(h, 1):
 *   sml_set_tisr(h, tisr)            the date's TISR field, before the exchange begins
 *   sml_step_exchange_begin(...)     additionally rebuilds every feedback vector and launches the NEXT predict's
 *                                    state update + W_out[:, S:]*x~ while only the grid copy-out is waited for
 *   (host: NetCDF output, run_model)
 *   sml_step_exchange_end(...)       uploads the forecast, builds local_model, adds W_out[:, 0:S]*local_model,
 *                                    un-standardises: outvec of step t+1 is complete (tisr may be NULL)
 *   sml_predict(h, SML_ATMO)         of step t+1 then only consumes that result.
 * Summation order differs from the fused readout (v_p + v_ml instead of one dot product): equal within 1e-13,
 * not bit-identical; reservoir state/feedback are bit-identical.  sml_state_get after exchange_end returns x(t+1).
 * Multi-rank callers use the device-only pieces: sml_step_pack_device, sml_step_predict_ahead,
 * (broadcast of F), sml_step_unpack_device. ---- */
int sml_set_overlap(sml_engine *h, int on);
int sml_set_tisr(sml_engine *h, const double *tisr_grid);
int sml_step_predict_ahead(sml_engine *h, int timestep);
/* zero-copy variants for hosts that can work in place: sml_step_exchange_begin_view returns pointers into the engine's
 * pinned staging instead of copying into caller arrays (valid until the next begin); sml_forecast_staging returns the
 * pinned arrays sml_step_exchange_end uploads from -- pass the same pointers to sml_step_exchange_end and the
 * intermediate copy is skipped.  The device<->host transfers themselves are unchanged. */
int sml_step_exchange_begin_view(sml_engine *h, int timestep, const double **wholegrid4d, const double **wholegrid2d,
                                 const double **wholegrid_precip, const double **wholegrid_sst);
int sml_forecast_staging(sml_engine *h, double **forecast_4d, double **forecast_2d, double **tisr_grid);
/* the grids of the last assembly, read back from the device on ANY rank (every rank rebuilds the whole grid; the
 * reference holds them on the root only).  NULL arrays are skipped.  Synchronises the engine's stream. */
int sml_grids_get(sml_engine *h, double *wholegrid4d, double *wholegrid2d, double *wholegrid_precip,
                  double *wholegrid_sst);
/* ---- failure detection.  sml_step_exchange_begin / _begin_view return 1 (grids still delivered) when the assembled
 * grid holds a non-finite value.  The status word also carries the bounds SPEEDY's own input check applies before it
 * agrees to run (src/ppo_iogrid.f90:562-577), evaluated on the grid handed to run_model; bits are sticky until
 * sml_grid_status_reset.  sml_grid_status synchronises the engine's stream. */
#define SML_GRID_NONFINITE 1
#define SML_GRID_U_RANGE 2   /* u outside [-150, 150] */
#define SML_GRID_V_RANGE 4   /* v outside [-120, 120] */
#define SML_GRID_T_RANGE 8   /* T outside [160, 330] */
#define SML_GRID_Q_RANGE 16  /* q outside [-6, 30] */
int sml_grid_status(sml_engine *h, int *bits);
int sml_grid_status_reset(sml_engine *h);
/* model_parameters%run_speedy: the root sets it after run_model (src/mpires.f90:1655-1659) BEFORE
 * sml_step_exchange_end; it travels to every rank with the forecast (MPI_Bcast, :744) and the step loop leaves when
 * it is false (src/parallelmain.f90:269-271).  sml_run_speedy on a non-root rank synchronises that rank's stream. */
int sml_set_run_speedy(sml_engine *h, int run_speedy);
int sml_run_speedy(sml_engine *h, int *run_speedy);
/* static fields of the exchange: base_sst_grid and sea_mask (src/mod_reservoir.f90:847-887) */
int sml_set_sst_static(sml_engine *h, const double *base_sst_grid, const double *sea_mask);
/* sst_prescribed == 1: the SST field (96x48) the next exchanges start from instead of ocean-reservoir
 * output; the land mask and the 272 K floor still apply (the role full_sst plays in get_sst_by_date,
 * src/mpires.f90:1710-1757) */
int sml_set_sst_prescribed(sml_engine *h, const double *sst_grid);

/* multi-rank plumbing (one process per GPU): device pointers of the exchange buffers so the host
 * can run the collective (NCCL all-gather of the outvec slabs; broadcast of F) on them. */
int sml_exchange_buffers(sml_engine *h, void **outvec_slab, int64_t *slab_count, void **gathered,
                         int64_t *gathered_count, void **gbuf, int64_t *g_count, void **fbuf,
                         int64_t *f_count);
/* ---- fused all-gather over NVLink (ranks of ONE node, one process per GPU).  After sml_finalize every rank
 * exports the CUDA IPC handle of its exchange block (64 bytes), the host passes all handles to all ranks (any
 * transport) and attaches them in rank order.  From then on the readout-finish kernel of sml_predict(h, SML_ATMO)
 * stores every outvec straight into every rank's gathered buffer (peer stores) and publishes a step flag; the
 * pack kernel of sml_step_exchange_begin / sml_step_pack_device waits for all flags.  No host collective is
 * needed for the atmosphere slabs any more (sml_exchange_buffers' gathered buffer is then unused).  Requires
 * number_of_regions divisible by numprocs (contiguous shards) and numprocs <= 8.  sml_peer_check reports a rank
 * that never published (the wait gives up after ~10 s instead of hanging). ---- */
int sml_peer_export(sml_engine *h, void *handle64);
/* the same set-up behind one call (replaces the MPI communicator the reference's sendrecievegrid uses,
 * src/mpires.f90:218-804): the host supplies ONE primitive, an all-gather of `bytes_per_rank` bytes from every rank
 * into recv in rank order (MPI_Allgather(send, n, MPI_BYTE, recv, n, MPI_BYTE, mpi_world) on the reference's
 * communicator; torch.distributed; a shared-memory rendezvous) returning 0 on success.  The library exchanges and
 * attaches the IPC handles itself and checks that every rank was built for the same model.  Afterwards the WHOLE
 * multi-rank step lives behind sml_predict / sml_step_exchange_begin / sml_step_exchange_end:
 *   - atmosphere outvecs: pushed into every rank's gathered copy by the readout-finish kernel (gather, :346-454);
 *   - ocean outvecs: pushed after every ocean step, and once for the seeded values (also after sml_ocean_ring_reset);
 *   - forecast: sml_step_exchange_end on rank 0 uploads [forecast_4d | forecast_2d | tisr | run_speedy] once and one
 *     kernel pushes the block into every rank's landing buffer (scatter :606-739, bcast :744); on the other ranks
 *     sml_step_exchange_end takes no forecast (arguments may be NULL; a non-NULL tisr_grid is that rank's own
 *     get_tisr_by_date field), waits for the block ON THE DEVICE and never blocks the host;
 *   - sml_step_exchange_begin with NULL arrays (ranks other than the root) only enqueues the grid assembly.
 * There is no host collective on the data path.  Ranks of ONE node, numprocs <= 8. */
typedef int (*sml_allgather_fn)(void *ctx, const void *send, void *recv, int bytes_per_rank);
int sml_comm_bootstrap(sml_engine *h, sml_allgather_fn allgather, void *ctx);
int sml_peer_attach(sml_engine *h, const void *handles /* numprocs x 64 bytes, rank order */, int count);
int sml_peer_attached(const sml_engine *h);
int sml_peer_check(sml_engine *h);
/* the same for the ocean reservoirs' outvec slab [nloc][P_ocean] (rows of regions without an ocean reservoir
 * hold 272.0, src/mpires.f90:323-326): all-gather it after every sml_predict(h, SML_OCEAN) when numprocs > 1 */
int sml_ocean_exchange_buffers(sml_engine *h, void **ocean_slab, int64_t *slab_count, void **ocean_gathered,
                               int64_t *gathered_count);
/* averaged_atmo_input_vec = 0 (initialize_prediction_slab, src/mod_slab_ocean_reservoir.f90:810-811) */
int sml_ocean_ring_reset(sml_engine *h);
/* device-only halves of begin/end for callers that keep the grids on the device */
int sml_step_pack_device(sml_engine *h, int timestep);                 /* gathered -> G (+clamps) */
int sml_step_unpack_device(sml_engine *h, int timestep);               /* G,F -> feedback, local_model */
/* the two as ONE cooperative launch (grid barrier between the scatter and the gather): the device-resident step of
 * ML-only runs.  Bit-identical to sml_step_pack_device + sml_step_unpack_device in the sequential mode. */
int sml_step_exchange_device(sml_engine *h, int timestep);

/* ---- training: reservoir_layer_chunking_hybrid/_ml + chunking_matmul(_ml)
 *      (src/mod_reservoir.f90:963-1175,1594-1701), fit_chunk_hybrid/_ml (:1177-1334) ---- */
int sml_train_begin(sml_engine *h, int kind, const int32_t *regions, int nregions, int batch_size);
/* one phase (trainingdata(:, i::timestep)) for the regions of sml_train_begin, series back to back:
 * trainingdata block of region i at td + td_off[i] (ld D_i), imperfect at im + im_off[i] (ld S_i);
 * inputs are pre-noised. */
int sml_train_feed(sml_engine *h, const double *td, const int64_t *td_off, const double *im,
                   const int64_t *im_off, int ncols, int discard_cols);
/* training data on the device: instead of per-region series, upload the conditioned GLOBAL series of
 * the training period once -- column t = [wholegrid4d | wholegrid2d(logp) | precip | sst | tisr] in the layout of
 * sml_global_layout (physical units after get_training_data's conditioning, src/mod_reservoir.f90:362-389: q in g/kg
 * floored at 1e-6, precip log(1 + p/eps), SST floored at 272) and, for hybrid training, F = [forecast_4d | forecast_2d],
 * the SPEEDY forecast valid at t.  sml_train_feed_global then tiles and standardises every region's input, imperfect-
 * model and target columns on the fly with the arithmetic of the forecast exchange (tile_4d_and_logp_to_local_state_input
 * + standardize_state_vec_input, src/res_domain.f90:1081-1125,1211-1268).  Inputs are noise-free (SURVEY.md 8c quirk 7:
 * per-region pre-noised series go through sml_train_feed).  phase column c = global column first_col + stride*c. */
int sml_train_global_series(sml_engine *h, const double *G_series, const double *F_series, int ncols_total);
int sml_train_feed_global(sml_engine *h, int first_col, int stride, int ncols, int discard_cols);
int sml_train_global_release(sml_engine *h);
/* multiplicative Gaussian input noise for sml_train_feed_global (gaussian_noise_1d_function / _precip,
 * src/mod_utilities.f90:1387-1464; reservoir%noisemag = 0.2): u*(1 + noisemag*g), precip rows noised in linear space and
 * transformed back; targets and the imperfect model stay noise-free as in the reference.  The N(0,1) draws come from a
 * counter-based generator keyed by (seed, region, column, element) -- reproducible, but not the Fortran random_number
 * stream.  noisemag = 0 (default) switches it off.  sml_train_noise_sample returns, for one region of the current wave,
 * the clean / Gaussian / noised input vector [D] of a phase column (inspection hook used by the tests). */
int sml_train_set_noise(sml_engine *h, double noisemag, unsigned long long seed, double precip_epsilon);
int sml_train_noise_sample(sml_engine *h, int region, int first_col, int stride, int col, double *clean, double *gauss,
                           double *noisy);
/* when the uploaded series is RAW (hourly, physical units as read from the reanalysis): apply get_training_data's
 * conditioning in place on the device (src/mod_reservoir.f90:362-395) -- q*1000 floored at 1e-6, TISR and precip
 * floored at 0, total_precip_over_a_period(period) (src/mod_utilities.f90:1688-1729) then log(1 + p/precip_epsilon),
 * SST floored at 272 K.  Once per upload. */
int sml_condition_series(sml_engine *h, int period, double precip_epsilon);
/* grid%mean / grid%std of every local region from the resident series (get_training_data, src/mod_reservoir.f90:413-470;
 * formulas: standardize_data_5d_logp_tisr src/mod_utilities.f90:1144-1193 two-pass population std, standardize_data_3d
 * :894-912 for precip, standardize_sst_data_3d :853-892 with its std > 0.2 gate -> sst_bool_input).  The series may be
 * uploaded before any region is (the constants are inputs of sml_region_upload).  Returns L, the slot count; mean and
 * std are [nloc][L] in local region order. */
int sml_conditioning_stats(sml_engine *h, int first_col, int stride, int ncols, double *mean, double *std,
                           int32_t *sst_bool_input);
int sml_train_solve(sml_engine *h, double beta_res, double beta_model, int using_prior,
                    double prior_val, int32_t *info_per_region);
/* sml_train_solve factorises the regularised Gram (symmetric positive definite for ridge > 0) of every region of
 * the wave by a batched blocked Cholesky on the FP64 tensor cores and falls back, per region, to LU with partial
 * pivoting -- what dgesv does -- when a pivot is not positive; info keeps dgesv's meaning either way.
 * by_cholesky: regions of the current wave that stayed on the Cholesky path. */
int sml_train_solver_stats(sml_engine *h, int *by_cholesky);
int sml_train_gram_get(sml_engine *h, int region, double *states_x_states_aug,
                       double *states_x_trainingdata_aug);
int sml_train_end(sml_engine *h);
/* scheduling of a wave's feeds: on = 0 (default) is the serial schedule -- per slab, state generation then Gram.
 * on = 1 (or SML_TRAIN_OVERLAP=1 in the environment) double-buffers the state slab and runs the Gram of one slab on its
 * own stream while the state generation -- the sequential part, reservoir_layer_chunking_hybrid's time loop -- fills
 * the other, across phases too; sml_train_solve / _gram_get / _stats / _end wait for both.  Same arithmetic in the same
 * order either way: the accumulators are bit-identical.  Measured: it pays for small waves (96 regions: 6.5 -> 5.5 s)
 * and not at full wave size (192 regions: 24.2 vs 23.6 s: the latency-bound state generation gets a quarter of its
 * threads beside a resident Gram CTA and slows the Gram by 9 %).  Takes effect at the next sml_train_begin. */
int sml_train_set_overlap(sml_engine *h, int on);
/* sml_train_end keeps the wave's device blocks for the next sml_train_begin (allocating ~350 MB per region anew for
 * every wave costs more than the solve); sml_train_trim returns them to the allocator */
int sml_train_trim(sml_engine *h);
/* measurement: useful Gram flops N(N+1)K + 2PNK accumulated by sml_train_feed and the CUDA-event time (ms) of
 * the Gram kernels, the state generation and the solves of the current wave */
/* how the last training phase generated its states: 0 one launch per time step (k_train_update), 1 the time loop inside
 * k_train_stategen, 2 the same loop on the TMA ring (k_train_stategen_ring, the default for waves of >= 24 regions); -1 none yet */
int sml_train_stategen_route(const sml_engine *h);
int sml_train_stats(sml_engine *h, double *gram_flops_useful, double *gram_ms, double *stategen_ms,
                    double *solve_ms);
/* the roof the Gram figure is reported against: FP64 tensor-core (DMMA m8n8k4) issue peak of this GPU in TFLOP/s,
 * measured on the spot by back-to-back register-resident MMAs (about 10 ms) */
int sml_dmma_probe(sml_engine *h, double *tflops);

/* rolling_average_over_a_period_2d(grid, period) (src/mod_utilities.f90:1773-1815), the time smoothing
 * get_training_data_from_atmo / get_prediction_data_from_atmo apply to the atmosphere rows of a slab-ocean reservoir's
 * input series (src/mod_slab_ocean_reservoir.f90:398, :452).  In place on a caller-owned host array: grid(i,t) at
 * grid[ld*t + i], i < nrows, t < t_len.  As written in the reference: sum(copy(i,1:t))/t while t-period < 1, else
 * sum(copy(i,t-period:t))/period (period+1 values), the latter kept only when |sum| > 1e-7 (keep_small = 1; 0 gives the
 * 3-D variant :1731-1771, which has no such test).  Windows are summed first to last. */
int sml_rolling_average_2d(sml_engine *h, double *grid, int ld, int nrows, int t_len, int period, int keep_small);

/* mldivide (src/mod_linalg.f90:109-151): solves A X = B in place of B; A(n,n) lda, B(n,nrhs) ldb.
 * returns dgesv's info (>0: singular, B is not the solution) */
int sml_mldivide(sml_engine *h, double *A, int lda, double *B, int ldb, int n, int nrhs);

/* ---- measurement hooks (bench.py): with profiling on, every sml_predict brackets its step kernel and
 * its finish kernel with CUDA events on the launching stream; sml_kernel_times sums and resets them ---- */
int sml_profile(sml_engine *h, int on);
int sml_kernel_times(sml_engine *h, double *step_ms_sum, double *finish_ms_sum, int *count);
int sml_phase_times(sml_engine *h, double *pack_ms_sum, double *unpack_ms_sum, int *count);
int sml_sync_times(sml_engine *h, double *update_ms_sum, int64_t *steps); /* update-only launches of sml_synchronize */
int sml_step_chunk_rows(const sml_engine *h, int kind); /* rows per CTA the step plan chose */
/* the step plan in use: kernel 0 = k_step (one CTA per item), 1 = k_step_persist; slots = persistent CTAs,
 * part_rows = rows per fixed row block, parts = partial outvecs per launch */
int sml_step_plan(const sml_engine *h, int kind, int *kernel, int *slots, int *part_rows, int *parts);
/* 1 when the persistent step kernel keeps the shard's adjacency, W_in and state vectors L2-resident across steps (evict-last
 * loads, W_out streamed evict-first): chosen at sml_finalize when they fit (SML_L2_KEEP=0/1 forces it) */
int sml_step_l2_keep(const sml_engine *h, int kind);
/* set-up cost: host seconds spent inside sml_region_upload so far, bytes of the weight arena, device allocations made for it */
int sml_setup_stats(const sml_engine *h, double *upload_seconds, int64_t *arena_bytes, int *arena_chunks);
int64_t sml_kernel_launch_count(const sml_engine *h);
/* algorithmic bytes one sml_predict(kind) moves */
int64_t sml_predict_algorithmic_bytes(const sml_engine *h, int kind);
/* the state-update part of it: what one step of sml_synchronize(kind, all regions) moves */
int64_t sml_update_algorithmic_bytes(const sml_engine *h, int kind);

#ifdef __cplusplus
}
#endif
#endif
