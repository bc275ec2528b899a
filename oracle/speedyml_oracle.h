/*
 * speedyml_oracle.h -- CPU restatement of the SPEEDY-ML local-reservoir hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is the checker for the CUDA engine; only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load it.  The product path (speedy-ml_b200/) never links or calls it.
 *
 * PARITY UNPINNED: the reference (Fortran 90 + MPI + MKL + ARPACK + NetCDF) cannot
 * be compiled here and its own tests hold no runnable golden vector for this path
 * (SURVEY.md section 8c).  The only reference-held known answers are
 *   - tests/mod_unit_test.f90:63-96  (288 regions, region 145 -> 4x4 tile, x 49..52)
 *   - tests/mod_unit_test.f90:16-47  (pinv(diag(1..10)) = diag(1/i))
 * both of which tests/test_oracle_index.py / test_oracle_linalg.py check.  Fidelity otherwise
 * rests on this C restatement agreeing with the independent NumPy restatement
 * oracle/oracle_np.py, both written from the reference text cited per function.
 *
 * All indices in this API are 1-based exactly as in the Fortran source unless a
 * comment says otherwise; arrays are column-major (first index fastest).
 */
#ifndef SPEEDYML_ORACLE_H
#define SPEEDYML_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_XGRID 96 /* src/mod_utilities.f90:18 */
#define ORC_YGRID 48 /* src/mod_utilities.f90:19 */
#define ORC_ZGRID 8  /* src/mod_utilities.f90:20 */

/* grid_type subset (src/mod_utilities.f90:32-167) */
typedef struct orc_grid {
    int number_of_regions, region, overlap, num_vert_levels, level_index, vert_overlap;
    int res_xstart, res_xend, res_ystart, res_yend, resxchunk, resychunk;
    int res_zstart, res_zend, reszchunk;
    int input_xstart, input_xend, input_ystart, input_yend, inputxchunk, inputychunk;
    int input_zstart, input_zend, inputzchunk;
    int tdata_xstart, tdata_xend, tdata_ystart, tdata_yend, tdata_zstart, tdata_zend;
    int pole, periodicboundary, top, bottom;
    int atmo3d_start, atmo3d_end, logp_start, logp_end, precip_start, precip_end;
    int sst_start, sst_end, tisr_start, tisr_end, predict_start, predict_end;
    int logp_mean_std_idx, tisr_mean_std_idx, precip_mean_std_idx, sst_mean_std_idx;
    int mean_std_length;
    /* slab-ocean grid_special only (src/mod_slab_ocean_reservoir.f90:1598-1640) */
    int ohtc_start, ohtc_end, ohtc_mean_std_idx, is_ocean;
} orc_grid;

/* reservoir_type sizing subset (src/mod_utilities.f90:169-369) */
typedef struct orc_dims {
    int local_predictvars, local_heightlevels_input, local_heightlevels_res;
    int logp_bool, precip_bool, precip_input_bool, sst_bool, sst_bool_input, tisr_input_bool;
    int logp_size_input, sst_size_input, precip_size_input, tisr_size_input;
    int logp_size_res, precip_size_res, sst_size_res, tisr_size_res;
    int chunk_size, chunk_size_prediction, chunk_size_speedy, locality;
    int m, n, k, reservoir_numinputs, nodes_per_input, ml_only;
    double deg, density, leakage;
} orc_dims;

/* ---- res_domain.f90 index arithmetic ---- */
int  orc_domaindecomposition(int numregions, int *factorx, int *factory);
void orc_getworkerlower_leftcorner(int region_num, int factory, int *row, int *col);
int  orc_getxyresextent(int num_regions, int region_num, int *xs, int *xe, int *ys, int *ye,
                        int *xchunk, int *ychunk);
void orc_get_z_res_extent(int num_vert_levels, int vert_level, int *zs, int *ze, int *zchunk);
int  orc_getoverlapindices(int numregions, int region_num, int overlap, int *ixs, int *ixe,
                           int *iys, int *iye, int *ixc, int *iyc, int *pole, int *periodic);
int  orc_getoverlapindices_vert(int num_vert_levels, int vert_level, int vert_overlap, int *izs,
                                int *ize, int *izc, int *top, int *bottom);
int  orc_get_trainingdataindices(int num_regions, int region_num, int overlap, int *xs, int *xe,
                                 int *ys, int *ye);
void orc_get_trainingdataindices_vert(int num_vert_levels, int vert_level, int vert_overlap,
                                      int *zs, int *ze);
int  orc_processor_decomposition(int irank, int numprocs, int number_of_regions, int *region_indices);
int  orc_find_closest_divisor(int target, int number);

/* initializedomain + allocate_res_new + trained_reservoir_prediction sizing */
int  orc_setup_region(int num_regions, int region, int overlap, int num_vert_levels, int vert_level,
                      int vert_overlap, int m, double deg, int precip_bool, int slab_ocean_model_bool,
                      int sst_bool_input, int ml_only, orc_grid *g, orc_dims *d);

/* slab-ocean reservoir sizing: initialize_slab_ocean_model (src/mod_slab_ocean_reservoir.f90:9-133) and the
 * offsets of trained_ocean_reservoir_prediction (:1561-1652); grid_special is a copy of the bottom-level
 * atmosphere grid with the vector offsets overwritten */
int  orc_setup_ocean_region(int num_regions, int region, int overlap, int m, double deg, int precip_bool,
                            orc_grid *g, orc_dims *d);

/* ---- tilers (array copies exactly as the reference slices them) ---- */
int  orc_tileoverlapgrid4d(const double *grid4d, int nvars, int numregions, int region, int overlap,
                           int num_vert_levels, int vert_level, int vert_overlap, double *localgrid);
int  orc_tileoverlapgrid2d(const double *grid2d, int numregions, int region, int overlap,
                           double *localgrid);
int  orc_tile_4d_and_logp_to_local_state_input(int numregions, int region, int overlap,
                           int num_vert_levels, int vert_level, int vert_overlap, int precip_bool,
                           const double *grid4d, const double *grid2d, const double *precip_grid,
                           double *inputvec);
void orc_tile_full_grid_with_local_state_vec_res1d(int numregions, int region, int num_vert_levels,
                           int vert_level, int precip_bool, const double *statevec, int length,
                           double *wholegrid4d, double *wholegrid2d, double *wholegrid_precip);
void orc_tile_full_2d_grid_with_local_res(int numregions, int region, const double *statevec,
                           double *wholegrid2d);
void orc_tile_4d_and_logp_full_grid_to_local_res_vec(int numregions, int region, int num_vert_levels,
                           int vert_level, const double *grid4d, const double *grid2d, double *statevec);
void orc_tile_full_input_to_target_data2d(const orc_grid *g, const orc_dims *d, const double *statevec,
                           int ld, int ncols, double *tiled /* P x ncols */);
void orc_standardize_state_vec_input(const orc_grid *g, const orc_dims *d, const double *mean,
                           const double *std, double *state_vec);
void orc_standardize_state_vec_res(const orc_grid *g, const orc_dims *d, const double *mean,
                           const double *std, double *state_vec);
void orc_unstandardize_state_vec_res(const orc_grid *g, const orc_dims *d, const double *mean,
                           const double *std, double *state_vec);

/* ---- linear algebra (mod_linalg.f90) ---- */
void orc_coo_mv(int n, int k, const int *rows, const int *cols, const double *vals, const double *x,
                double *y);
int  orc_dgesv(int n, int nrhs, double *A, int lda, int *ipiv, double *B, int ldb);
int  orc_mldivide(double *A, int n, int m, double *B, int l, int k);

/* ---- reservoir object ---- */
typedef struct orc_region orc_region;
orc_region *orc_region_new(const orc_grid *g, const orc_dims *d);
void orc_region_free(orc_region *r);
int  orc_region_set_weights(orc_region *r, const int *rows, const int *cols, const double *vals,
                            const double *win, const double *wout, const double *mean,
                            const double *std, int mean_std_length);
void orc_region_set_leakage(orc_region *r, double leakage);
int  orc_region_set_win_compact(orc_region *r, const double *winc, const int *wcol);
/* reservoir construction with the counter-based generator shared with the engine (makesparse src/mod_linalg.f90:180-218,
 * shuffle src/mod_utilities.f90:1569-1596, W_in src/mod_reservoir.f90:262-283) */
#define ORC_STREAM_WIN 4096
unsigned long long orc_counter_bits(unsigned long long seed, int region, int stream, long long index);
void orc_shuffle(int n, int returnsize, unsigned long long seed, int region, int stream, int *shufflereturn);
void orc_makesparse(int n, int k, unsigned long long seed, int region, int *rows, int *cols, double *vals);
void orc_gen_win(int n, int D, double sigma, unsigned long long seed, int region, double *winc, int *wcol);
/* compact W_in -> the dense win(n, D) the reference stores; predict runs the dense GEMV again (bench.py CPU arm) */
int  orc_region_densify_win(orc_region *r);
double *orc_region_ptr(orc_region *r, const char *field); /* x feedback local_model outvec wout ... */
const orc_grid *orc_region_grid(const orc_region *r);
const orc_dims *orc_region_dims(const orc_region *r);

void orc_synchronize(orc_region *r, const double *input, int ld, double *x, int length);
void orc_predict(orc_region *r, double *x);      /* hybrid: uses r->feedback, r->local_model */
void orc_predict_ml(orc_region *r, double *x);
void orc_predict_all(orc_region **regs, int nreg, int ml_only, int nthreads);
/* predict_slab_ml (src/mod_slab_ocean_reservoir.f90:1318-1363): all outputs * std(sst) + mean(sst) */
void orc_predict_slab_ml(orc_region *r, double *x);
void orc_predict_slab(orc_region *r, double *x);   /* hybrid ocean reservoir, src/mod_slab_ocean_reservoir.f90:1268-1316 */
/* ocean feedback of one hybrid step, intended semantics of SURVEY.md Appendix C (src/mpires.f90:594-600,
 * 776-781): ring(:, mod(timestep-1,nslots)+1) = atmosphere feedback(atmo3d_end-4*ixy+1 : logp_end) (already
 * standardised); feedback(1:logp_end) = sum(ring,dim=2)/nslots; feedback(sst_start:sst_end) = standardised
 * halo'd tile of wholegrid_sst; the TISR and OHTC slots are left as they are.  ring is (logp_end, nslots). */
void orc_ocean_feedback(orc_region *ocean, const orc_region *atmo, double *ring, int nslots, int timestep,
                        const double *wholegrid_sst);

/* sendrecievegrid split at the host-model call (src/mpires.f90:218-804) */
void orc_step_gather(orc_region **regs, int nreg, int precip_bool, int ocean_model,
                     const double *base_sst_grid, const double *sea_mask,
                     const double *ocean_outvec /* nreg x 4 or NULL */, const int *has_ocean,
                     double *wholegrid4d, double *wholegrid2d, double *wholegrid_precip,
                     double *wholegrid_sst);
void orc_step_scatter(orc_region **regs, int nreg, int precip_bool, int ocean_model, int ml_only,
                      const double *wholegrid4d, const double *wholegrid2d,
                      const double *wholegrid_precip, const double *wholegrid_sst,
                      const double *forecast_4d, const double *forecast_2d,
                      const double *tisr_grid, const double *sst_mean, const double *sst_std,
                      int nthreads);
void orc_run_model_clamp(double *grid4d); /* q floor before/after the host model */
void orc_host_stub(const double *grid4d, const double *grid2d, const double *clim4d,
                   const double *clim2d, double *forecast_4d, double *forecast_2d);

/* rolling_average_over_a_period_2d (src/mod_utilities.f90:1773-1815); keep_small = 0: the 3-D variant :1731-1771 */
void orc_rolling_average_2d(double *grid, int ld, int nrows, int t_len, int period, int keep_small);

/* ---- training (src/mod_reservoir.f90:1067-1334,1561-1701) ---- */
int  orc_train_init(orc_region *r, int batch_size);
void orc_train_phase_hybrid(orc_region *r, const double *trainingdata, int ld_t,
                            const double *imperfect, int ld_i, int ncols, int discard_cols);
void orc_train_phase_ml(orc_region *r, const double *trainingdata, int ld_t, int ncols,
                        int discard_cols);
int  orc_fit_chunk_hybrid(orc_region *r, double beta_res, double beta_model, int using_prior,
                          double prior_val);
int  orc_fit_chunk_ml(orc_region *r, double beta_res);
void orc_train_free(orc_region *r);

#ifdef __cplusplus
}
#endif
#endif
