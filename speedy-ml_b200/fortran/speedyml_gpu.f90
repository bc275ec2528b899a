!> speedyml_gpu.f90 -- ISO_C_BINDING layer between the SPEEDY-ML Fortran host and libspeedyml_b200.so
!>
!> Keeps the names and argument meaning of the reference procedures on the hot path
!>   mklsparse (src/mod_linalg.f90:10), synchronize (src/mod_reservoir.f90:1354), predict / predict_ml
!>   (:1418, :1491), predict_slab_ml (src/mod_slab_ocean_reservoir.f90:1318), sendrecievegrid
!>   (src/mpires.f90:218), chunking / fit (src/mod_reservoir.f90:963-1334) and mldivide (src/mod_linalg.f90:109)
!> and forwards them to the C ABI declared in include/speedyml_engine.h.  INTEGRATION.md lists the call
!> sites that change.  The derived types are the reference's own (mod_utilities).
!>
!> STATUS: source only.  The image this engine is developed in has no Fortran compiler (gcc lacks f951) and no
!> MPI/MKL/NetCDF, so this file has never been compiled; the same C entry points are exercised from Python
!> (speedy-ml_b200/engine.py) by tests/.  Build on a host with the reference toolchain:
!>     mpif90 -c speedyml_gpu.f90 ; link imp.exe with -lspeedyml_b200
!>
!> Execution model: the device holds every local region's weights and state.  predict() enqueues nothing per
!> region; the first predict call of a hybrid step launches ONE batched kernel for all local regions, later
!> calls in the same step only copy that region's outvec back (reservoir%outvec stays valid for host code such
!> as the diagnostics writers).  sendrecievegrid() closes the step.
module speedyml_gpu
  use, intrinsic :: iso_c_binding
  use mod_utilities, only : dp, main_type, reservoir_type, grid_type, model_parameters_type, xgrid, ygrid, zgrid
  implicit none
  private
  public :: gpu_engine_create, gpu_engine_finalize, gpu_engine_destroy
  public :: mklsparse, synchronize, predict, predict_ml, predict_slab_ml, sendrecievegrid
  public :: gpu_train_begin, gpu_train_phase, gpu_fit_chunk, gpu_train_end, mldivide
  public :: gen_res, read_trained_res, gpu_train_global_series, gpu_train_phase_global, gpu_set_overlap, gpu_set_tisr
  public :: rolling_average_over_a_period_2d

  integer(c_int), parameter :: SML_ATMO = 0, SML_OCEAN = 1, SML_ALL_REGIONS = -1

  type, bind(C) :: sml_params
     integer(c_int32_t) :: number_of_regions, overlap, precip_bool, slab_ocean_model_bool, ml_only
     integer(c_int32_t) :: irank, numprocs, device, timestep, timestep_slab, sst_prescribed
     integer(c_int32_t) :: reserved(5)
  end type

  type, bind(C) :: sml_region_weights
     integer(c_int32_t) :: region, kind, n, k, D, P, S, L, sst_bool_input, reserved0
     real(c_double)     :: leakage, sst_mean, sst_std
     type(c_ptr)        :: rows, cols, vals, win_dense, win_compact, win_col, wout, mean, std
  end type

  type(c_ptr), save :: h = c_null_ptr          !< the engine of this MPI rank (one rank per GPU)
  integer, save     :: step_predicted = -1     !< hybrid step whose batched predict has been launched
  integer, save     :: current_step = 0

  interface
     integer(c_int) function sml_create(h, p) bind(C, name='sml_create')
       import :: c_ptr, c_int, sml_params
       type(c_ptr), intent(out) :: h
       type(sml_params), intent(in) :: p
     end function
     integer(c_int) function sml_destroy(h) bind(C, name='sml_destroy')
       import :: c_ptr, c_int
       type(c_ptr), value :: h
     end function
     type(c_ptr) function sml_last_error(h) bind(C, name='sml_last_error')
       import :: c_ptr
       type(c_ptr), value :: h
     end function
     integer(c_int) function sml_region_upload(h, w) bind(C, name='sml_region_upload')
       import :: c_ptr, c_int, sml_region_weights
       type(c_ptr), value :: h
       type(sml_region_weights), intent(in) :: w
     end function
     integer(c_int) function sml_finalize(h) bind(C, name='sml_finalize')
       import :: c_ptr, c_int
       type(c_ptr), value :: h
     end function
     integer(c_int) function sml_state_set(h, kind, region, x) bind(C, name='sml_state_set')
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: h
       integer(c_int), value :: kind, region
       real(c_double), intent(in) :: x(*)
     end function
     integer(c_int) function sml_state_get(h, kind, region, x) bind(C, name='sml_state_get')
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: h
       integer(c_int), value :: kind, region
       real(c_double), intent(out) :: x(*)
     end function
     integer(c_int) function sml_feedback_set(h, kind, region, v) bind(C, name='sml_feedback_set')
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: h
       integer(c_int), value :: kind, region
       real(c_double), intent(in) :: v(*)
     end function
     integer(c_int) function sml_feedback_get(h, kind, region, v) bind(C, name='sml_feedback_get')
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: h
       integer(c_int), value :: kind, region
       real(c_double), intent(out) :: v(*)
     end function
     integer(c_int) function sml_local_model_set(h, kind, region, v) bind(C, name='sml_local_model_set')
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: h
       integer(c_int), value :: kind, region
       real(c_double), intent(in) :: v(*)
     end function
     integer(c_int) function sml_local_model_get(h, kind, region, v) bind(C, name='sml_local_model_get')
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: h
       integer(c_int), value :: kind, region
       real(c_double), intent(out) :: v(*)
     end function
     integer(c_int) function sml_outvec_get(h, kind, region, v) bind(C, name='sml_outvec_get')
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: h
       integer(c_int), value :: kind, region
       real(c_double), intent(out) :: v(*)
     end function
     integer(c_int) function sml_outvec_set(h, kind, region, v) bind(C, name='sml_outvec_set')
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: h
       integer(c_int), value :: kind, region
       real(c_double), intent(in) :: v(*)
     end function
     integer(c_int) function sml_wout_get(h, kind, region, w) bind(C, name='sml_wout_get')
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: h
       integer(c_int), value :: kind, region
       real(c_double), intent(out) :: w(*)
     end function
     integer(c_int) function sml_synchronize(h, kind, region, inputs, ld, length, offsets) bind(C, name='sml_synchronize')
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: h
       integer(c_int), value :: kind, region, ld, length
       real(c_double), intent(in) :: inputs(*)
       type(c_ptr), value :: offsets
     end function
     integer(c_int) function sml_predict(h, kind) bind(C, name='sml_predict')
       import :: c_ptr, c_int
       type(c_ptr), value :: h
       integer(c_int), value :: kind
     end function
     integer(c_int) function sml_step_exchange_begin(h, timestep, w4d, w2d, wprecip, wsst) bind(C, name='sml_step_exchange_begin')
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: h
       integer(c_int), value :: timestep
       real(c_double), intent(out) :: w4d(*), w2d(*), wprecip(*), wsst(*)
     end function
     integer(c_int) function sml_step_exchange_end(h, timestep, f4d, f2d, tisr) bind(C, name='sml_step_exchange_end')
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: h
       integer(c_int), value :: timestep
       real(c_double), intent(in) :: f4d(*), f2d(*), tisr(*)
     end function
     integer(c_int) function sml_set_sst_static(h, base_sst, sea_mask) bind(C, name='sml_set_sst_static')
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: h
       real(c_double), intent(in) :: base_sst(*), sea_mask(*)
     end function
     integer(c_int) function sml_train_begin(h, kind, regions, nregions, batch_size) bind(C, name='sml_train_begin')
       import :: c_ptr, c_int, c_int32_t
       type(c_ptr), value :: h
       integer(c_int), value :: kind, nregions, batch_size
       integer(c_int32_t), intent(in) :: regions(*)
     end function
     integer(c_int) function sml_train_feed(h, td, td_off, im, im_off, ncols, discard_cols) bind(C, name='sml_train_feed')
       import :: c_ptr, c_int, c_double, c_int64_t
       type(c_ptr), value :: h
       real(c_double), intent(in) :: td(*), im(*)
       integer(c_int64_t), intent(in) :: td_off(*), im_off(*)
       integer(c_int), value :: ncols, discard_cols
     end function
     integer(c_int) function sml_train_solve(h, beta_res, beta_model, using_prior, prior_val, info) bind(C, name='sml_train_solve')
       import :: c_ptr, c_int, c_double, c_int32_t
       type(c_ptr), value :: h
       real(c_double), value :: beta_res, beta_model, prior_val
       integer(c_int), value :: using_prior
       integer(c_int32_t), intent(out) :: info(*)
     end function
     integer(c_int) function sml_train_end(h) bind(C, name='sml_train_end')
       import :: c_ptr, c_int
       type(c_ptr), value :: h
     end function
     integer(c_int) function sml_sparse_eigen(h, kind, maxit, tol, eigs, iterations) bind(C, name='sml_sparse_eigen')
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: h
       integer(c_int), value :: kind, maxit
       real(c_double), value :: tol
       real(c_double), intent(out) :: eigs(*)
       integer(c_int), intent(out) :: iterations
     end function
     integer(c_int) function sml_adjacency_scale(h, kind, factor) bind(C, name='sml_adjacency_scale')
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: h
       integer(c_int), value :: kind
       real(c_double), intent(in) :: factor(*)
     end function
     integer(c_int) function sml_region_upload_file(h, path, region, kind, sst_bool_input, leakage) bind(C, name='sml_region_upload_file')
       import :: c_ptr, c_int, c_double, c_char
       type(c_ptr), value :: h
       character(kind=c_char), intent(in) :: path(*)
       integer(c_int), value :: region, kind, sst_bool_input
       real(c_double), value :: leakage
     end function
     integer(c_int) function sml_train_global_series(h, g_series, f_series, ncols_total) bind(C, name='sml_train_global_series')
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: h
       real(c_double), intent(in) :: g_series(*), f_series(*)
       integer(c_int), value :: ncols_total
     end function
     integer(c_int) function sml_train_feed_global(h, first_col, stride, ncols, discard_cols) bind(C, name='sml_train_feed_global')
       import :: c_ptr, c_int
       type(c_ptr), value :: h
       integer(c_int), value :: first_col, stride, ncols, discard_cols
     end function
     integer(c_int) function sml_set_overlap(h, on) bind(C, name='sml_set_overlap')
       import :: c_ptr, c_int
       type(c_ptr), value :: h
       integer(c_int), value :: on
     end function
     integer(c_int) function sml_set_tisr(h, tisr) bind(C, name='sml_set_tisr')
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: h
       real(c_double), intent(in) :: tisr(*)
     end function
     integer(c_int) function sml_mldivide(h, A, lda, B, ldb, n, nrhs) bind(C, name='sml_mldivide')
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: h
       real(c_double), intent(inout) :: A(*), B(*)
       integer(c_int), value :: lda, ldb, n, nrhs
     end function
     integer(c_int) function sml_rolling_average_2d(h, grid, ld, nrows, t_len, period, keep_small) &
          bind(C, name='sml_rolling_average_2d')
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: h
       real(c_double), intent(inout) :: grid(*)
       integer(c_int), value :: ld, nrows, t_len, period, keep_small
     end function
  end interface

contains

  !> print-and-stop on engine errors, like mklsparse's stat check (src/mod_linalg.f90:18-22)
  subroutine ck(rc, where)
    integer(c_int), intent(in) :: rc
    character(len=*), intent(in) :: where
    character(kind=c_char), pointer :: msg(:)
    integer :: i
    if (rc >= 0) return
    call c_f_pointer(sml_last_error(h), msg, [512])
    i = 1
    do while (i < 512 .and. msg(i) /= c_null_char)
       i = i + 1
    end do
    print *, 'speedyml_gpu: ', where, ' failed: ', msg(1:i-1)
    stop
  end subroutine

  !> once per rank, after MPI start-up (replaces nothing; call before trained_reservoir_prediction)
  subroutine gpu_engine_create(model_parameters, device)
    type(model_parameters_type), intent(in) :: model_parameters
    integer, intent(in) :: device
    type(sml_params) :: p
    p%number_of_regions = model_parameters%number_of_regions
    p%overlap = model_parameters%overlap
    p%precip_bool = merge(1, 0, model_parameters%precip_bool)
    p%slab_ocean_model_bool = merge(1, 0, model_parameters%slab_ocean_model_bool)
    p%ml_only = merge(1, 0, model_parameters%ml_only)
    p%irank = model_parameters%irank
    p%numprocs = model_parameters%numprocs
    p%device = device
    p%timestep = model_parameters%timestep
    p%timestep_slab = model_parameters%timestep_slab
    p%sst_prescribed = 0
    p%reserved = 0
    call ck(sml_create(h, p), 'sml_create')
  end subroutine

  !> mklsparse(reservoir): the reference builds the MKL COO handle here; the engine takes the whole reservoir
  !> (adjacency, W_in, W_out) because that is the moment all of them exist (src/mod_reservoir.f90:1852).
  !> grid supplies mean/std (and, for the ocean reservoir, the SST slot); kind defaults to the atmosphere.
  subroutine mklsparse(reservoir, grid, kind)
    type(reservoir_type), intent(inout), target :: reservoir
    type(grid_type), intent(in), target, optional :: grid
    integer, intent(in), optional :: kind
    type(sml_region_weights) :: w
    real(dp), target, save :: zero(1) = 0.0_dp, one(1) = 1.0_dp
    w%region = reservoir%assigned_region
    w%kind = SML_ATMO
    if (present(kind)) w%kind = kind
    w%n = reservoir%n
    w%k = reservoir%k
    w%D = reservoir%reservoir_numinputs
    w%P = reservoir%chunk_size_prediction
    w%S = reservoir%chunk_size_speedy
    w%sst_bool_input = merge(1, 0, reservoir%sst_bool_input)
    w%reserved0 = 0
    w%leakage = reservoir%leakage
    w%rows = c_loc(reservoir%rows)
    w%cols = c_loc(reservoir%cols)
    w%vals = c_loc(reservoir%vals)
    w%win_dense = c_loc(reservoir%win)      ! the engine verifies "one non-zero per row" and compresses
    w%win_compact = c_null_ptr
    w%win_col = c_null_ptr
    w%wout = c_null_ptr
    if (allocated(reservoir%wout)) w%wout = c_loc(reservoir%wout)
    if (present(grid)) then
       w%L = size(grid%mean)
       w%mean = c_loc(grid%mean)
       w%std = c_loc(grid%std)
       w%sst_mean = 0.0_dp
       w%sst_std = 1.0_dp
       if (grid%sst_mean_std_idx > 0) then
          w%sst_mean = grid%mean(grid%sst_mean_std_idx)
          w%sst_std = grid%std(grid%sst_mean_std_idx)
       end if
    else
       w%L = 1
       w%mean = c_loc(zero)
       w%std = c_loc(one)
       w%sst_mean = 0.0_dp
       w%sst_std = 1.0_dp
    end if
    call ck(sml_region_upload(h, w), 'mklsparse/sml_region_upload')
  end subroutine

  !> after the last mklsparse of the rank (end of the trained_reservoir_prediction loop, src/parallelmain.f90:160-185)
  subroutine gpu_engine_finalize(model_parameters)
    type(model_parameters_type), intent(in) :: model_parameters
    call ck(sml_finalize(h), 'sml_finalize')
    if (model_parameters%slab_ocean_model_bool) then
       call ck(sml_set_sst_static(h, model_parameters%base_sst_grid, model_parameters%sea_mask), 'sml_set_sst_static')
    end if
  end subroutine

  subroutine gpu_engine_destroy()
    integer(c_int) :: rc
    rc = sml_destroy(h)
    h = c_null_ptr
  end subroutine

  !> synchronize(reservoir,input,x,length), src/mod_reservoir.f90:1354-1381 (slab twin :1237-1266 with kind)
  subroutine synchronize(reservoir, input, x, length, kind)
    type(reservoir_type), intent(inout) :: reservoir
    real(kind=dp), intent(in)    :: input(:,:)
    real(kind=dp), intent(inout) :: x(:)
    integer, intent(in)          :: length
    integer, intent(in), optional :: kind
    integer(c_int) :: k
    k = SML_ATMO
    if (present(kind)) k = kind
    call ck(sml_state_set(h, k, reservoir%assigned_region, x), 'sml_state_set')
    call ck(sml_synchronize(h, k, reservoir%assigned_region, input, size(input, 1), length, c_null_ptr), 'sml_synchronize')
    call ck(sml_state_get(h, k, reservoir%assigned_region, x), 'sml_state_get')
  end subroutine

  !> predict(reservoir,model_parameters,grid,x,local_model_in), src/mod_reservoir.f90:1418-1489.
  !> The first call of a hybrid step runs the batched kernel for every local region; x (current_state) lives
  !> on the device between steps and is only copied back on request (sml_state_get).
  subroutine predict(reservoir, model_parameters, grid, x, local_model_in)
    type(reservoir_type), intent(inout)     :: reservoir
    type(model_parameters_type), intent(in) :: model_parameters
    type(grid_type), intent(in)             :: grid
    real(kind=dp), intent(inout) :: x(:)
    real(kind=dp), intent(inout) :: local_model_in(:)
    if (step_predicted /= current_step) then
       call ck(sml_predict(h, SML_ATMO), 'sml_predict')
       step_predicted = current_step
    end if
    call ck(sml_outvec_get(h, SML_ATMO, reservoir%assigned_region, reservoir%outvec), 'sml_outvec_get')
  end subroutine

  !> predict_ml, src/mod_reservoir.f90:1491-1535: same entry, the engine was created with ml_only
  subroutine predict_ml(reservoir, model_parameters, grid, x)
    type(reservoir_type), intent(inout)     :: reservoir
    type(model_parameters_type), intent(in) :: model_parameters
    type(grid_type), intent(in)             :: grid
    real(kind=dp), intent(inout) :: x(:)
    real(kind=dp) :: none(1)
    call predict(reservoir, model_parameters, grid, x, none)
  end subroutine

  !> predict_slab_ml, src/mod_slab_ocean_reservoir.f90:1318-1363; the caller keeps the schedule test
  !> mod(t*timestep, timestep_slab) == 0 (src/parallelmain.f90:238).  One batched launch per ocean step.
  subroutine predict_slab_ml(reservoir, model_parameters, grid, x)
    type(reservoir_type), intent(inout)     :: reservoir
    type(model_parameters_type), intent(in) :: model_parameters
    type(grid_type), intent(in)             :: grid
    real(kind=dp), intent(inout) :: x(:)
    integer, save :: ocean_step_predicted = -1
    if (ocean_step_predicted /= current_step) then
       call ck(sml_predict(h, SML_OCEAN), 'sml_predict(ocean)')
       ocean_step_predicted = current_step
    end if
    call ck(sml_outvec_get(h, SML_OCEAN, reservoir%assigned_region, reservoir%outvec), 'sml_outvec_get(ocean)')
  end subroutine

  !> sendrecievegrid(res,timestep,ocean_model), src/mpires.f90:218-804.  The MPI star through the root, the
  !> clamps (:456-490) and the feedback / local_model rebuild (:581-604,749-791) happen on the devices; the
  !> root still owns NetCDF output and run_model (:565-569), passed in as procedure arguments so this module
  !> does not depend on mod_io / speedy_res_interface.  tisr_grid: get_tisr_by_date's field for timestep-1.
  !> With more than one rank the outvec slabs are all-gathered by NCCL inside the library's host layer
  !> (see INTEGRATION.md: the Python/torch.distributed driver does it today; an MPI_Allgather on the device
  !> pointers of sml_exchange_buffers is the Fortran equivalent with a CUDA-aware MPI).
  subroutine sendrecievegrid(res, timestep, ocean_model, run_model, write_prediction, tisr_grid)
    type(main_type), intent(inout) :: res
    integer, intent(in) :: timestep
    logical, intent(in) :: ocean_model
    real(kind=dp), intent(in) :: tisr_grid(:,:)
    interface
       subroutine run_model(model_parameters, timestep, grid4d, grid2d, sst_grid, speedy_grid4d, speedy_grid2d)
         import :: model_parameters_type, dp
         type(model_parameters_type), intent(inout) :: model_parameters
         integer, intent(in) :: timestep
         real(kind=dp), intent(inout) :: grid4d(:,:,:,:), grid2d(:,:), sst_grid(:,:)
         real(kind=dp), intent(out)   :: speedy_grid4d(:,:,:,:), speedy_grid2d(:,:)
       end subroutine
       subroutine write_prediction(res, timestep, grid4d, grid2d, precip_grid, sst_grid)
         import :: main_type, dp
         type(main_type), intent(inout) :: res
         integer, intent(in) :: timestep
         real(kind=dp), intent(in) :: grid4d(:,:,:,:), grid2d(:,:), precip_grid(:,:), sst_grid(:,:)
       end subroutine
    end interface
    real(kind=dp), allocatable :: wholegrid4d(:,:,:,:), wholegrid2d(:,:), wholegrid_precip(:,:), wholegrid_sst(:,:)
    real(kind=dp), allocatable :: forecast_4d(:,:,:,:), forecast_2d(:,:)
    integer :: nv
    nv = res%model_parameters%full_predictvars
    allocate(wholegrid4d(nv, xgrid, ygrid, zgrid), wholegrid2d(xgrid, ygrid))
    allocate(wholegrid_precip(xgrid, ygrid), wholegrid_sst(xgrid, ygrid))
    allocate(forecast_4d(nv, xgrid, ygrid, zgrid), forecast_2d(xgrid, ygrid))
    call ck(sml_step_exchange_begin(h, timestep, wholegrid4d, wholegrid2d, wholegrid_precip, wholegrid_sst), &
            'sml_step_exchange_begin')
    if (res%model_parameters%irank == 0) then
       call write_prediction(res, timestep, wholegrid4d, wholegrid2d, wholegrid_precip, wholegrid_sst)
       if (.not. res%model_parameters%ml_only) then
          call run_model(res%model_parameters, timestep, wholegrid4d, wholegrid2d, wholegrid_sst, forecast_4d, forecast_2d)
       end if
    end if
    call ck(sml_step_exchange_end(h, timestep, forecast_4d, forecast_2d, tisr_grid), 'sml_step_exchange_end')
    current_step = timestep + 1
  end subroutine

  !> train_reservoir's inner sequence for a wave of regions (src/mod_reservoir.f90:287-316):
  !>   initialize_chunk_training           -> gpu_train_begin
  !>   reservoir_layer_chunking_hybrid/_ml -> gpu_train_phase (one call per phase trainingdata(:, i::timestep))
  !>   fit_chunk_hybrid/_ml                -> gpu_fit_chunk (ridge terms + dgesv; reservoir%wout filled)
  subroutine gpu_train_begin(regions, batch_size, kind)
    integer(c_int32_t), intent(in) :: regions(:)
    integer, intent(in) :: batch_size
    integer, intent(in), optional :: kind
    integer(c_int) :: k
    k = SML_ATMO
    if (present(kind)) k = kind
    call ck(sml_train_begin(h, k, regions, size(regions), batch_size), 'sml_train_begin')
  end subroutine

  !> trainingdata / imperfect_model: the phase's (pre-noised) series of the wave's regions back to back,
  !> offsets in doubles; discard_cols = discardlength/timestep (:1093)
  subroutine gpu_train_phase(trainingdata, td_off, imperfect_model, im_off, ncols, discard_cols)
    real(kind=dp), intent(in) :: trainingdata(:), imperfect_model(:)
    integer(c_int64_t), intent(in) :: td_off(:), im_off(:)
    integer, intent(in) :: ncols, discard_cols
    call ck(sml_train_feed(h, trainingdata, td_off, imperfect_model, im_off, ncols, discard_cols), 'sml_train_feed')
  end subroutine

  !> fit_chunk_hybrid (src/mod_reservoir.f90:1235-1334) / fit_chunk_ml (:1177-1233) for every region of the wave;
  !> info(i) is dgesv's info ('something went wrong with dgesv' is print-and-continue, src/mod_linalg.f90:147-150)
  subroutine gpu_fit_chunk(reservoirs, using_prior, info)
    type(reservoir_type), intent(inout) :: reservoirs(:)
    logical, intent(in) :: using_prior
    integer(c_int32_t), intent(out) :: info(:)
    integer :: i
    call ck(sml_train_solve(h, reservoirs(1)%beta_res, reservoirs(1)%beta_model, merge(1, 0, using_prior), &
                            reservoirs(1)%prior_val, info), 'sml_train_solve')
    do i = 1, size(reservoirs)
       if (info(i) /= 0) then
          print *, 'something went wrong with dgesv info = ', info(i)
          print *, 'B is not the solution'
       else
          call ck(sml_wout_get(h, SML_ATMO, reservoirs(i)%assigned_region, reservoirs(i)%wout), 'sml_wout_get')
       end if
    end do
  end subroutine

  subroutine gpu_train_end()
    call ck(sml_train_end(h), 'sml_train_end')
  end subroutine

  !> gen_res(reservoir), src/mod_reservoir.f90:182-212, for every local reservoir at once: after makesparse + mklsparse
  !> of the unscaled adjacency and gpu_engine_finalize, the spectral radii come from the device (replaces the ARPACK
  !> sparse_eigen, src/mod_linalg.f90:220-514) and vals = (vals/eig)*radius is applied on both sides
  subroutine gen_res(reservoirs, kind)
    type(reservoir_type), intent(inout) :: reservoirs(:)
    integer, intent(in), optional :: kind
    real(kind=dp), allocatable :: eigs(:), factor(:)
    integer(c_int) :: k, iters, rc
    integer :: i
    k = SML_ATMO
    if (present(kind)) k = kind
    allocate(eigs(size(reservoirs)), factor(size(reservoirs)))
    rc = sml_sparse_eigen(h, k, 500, 1.0d-13, eigs, iters)
    call ck(rc, 'sml_sparse_eigen')
    if (rc == 1) print *, 'sparse_eigen did not converge in', iters, 'iterations'
    do i = 1, size(reservoirs)
       factor(i) = 1.0_dp
       if (eigs(i) > 0.0_dp) factor(i) = reservoirs(i)%radius / eigs(i)
       reservoirs(i)%vals = reservoirs(i)%vals * factor(i)
    end do
    call ck(sml_adjacency_scale(h, k, factor), 'sml_adjacency_scale')
  end subroutine

  !> read_trained_res + mklsparse in one call (src/mod_io.f90:2938-2983, src/mod_reservoir.f90:1852): the engine reads
  !> the NetCDF-classic file write_trained_res produced itself and uploads it (float32 -> real(dp) widening)
  subroutine read_trained_res(reservoir, filename, kind)
    type(reservoir_type), intent(in) :: reservoir
    character(len=*), intent(in) :: filename
    integer, intent(in), optional :: kind
    integer(c_int) :: k
    k = SML_ATMO
    if (present(kind)) k = kind
    call ck(sml_region_upload_file(h, trim(filename)//c_null_char, reservoir%assigned_region, k, &
                                   merge(1, 0, reservoir%sst_bool_input), reservoir%leakage), 'sml_region_upload_file')
  end subroutine

  !> the conditioned global series of the training period, resident on the device for all waves and phases;
  !> g_series(g_total, ncols), f_series(f_total, ncols) in the layout of sml_global_layout
  subroutine gpu_train_global_series(g_series, f_series)
    real(kind=dp), intent(in) :: g_series(:,:), f_series(:,:)
    call ck(sml_train_global_series(h, g_series, f_series, size(g_series, 2)), 'sml_train_global_series')
  end subroutine

  !> one phase trainingdata(:, i::timestep) from the resident series (first_col is 0-based)
  subroutine gpu_train_phase_global(first_col, stride, ncols, discard_cols)
    integer, intent(in) :: first_col, stride, ncols, discard_cols
    call ck(sml_train_feed_global(h, first_col, stride, ncols, discard_cols), 'sml_train_feed_global')
  end subroutine

  !> overlapped step: the next predict's state update and x~ readout run while run_model works (call gpu_set_tisr
  !> with get_tisr_by_date's field before every sendrecievegrid)
  subroutine gpu_set_overlap(on)
    logical, intent(in) :: on
    call ck(sml_set_overlap(h, merge(1, 0, on)), 'sml_set_overlap')
  end subroutine
  subroutine gpu_set_tisr(tisr_grid)
    real(kind=dp), intent(in) :: tisr_grid(:,:)
    call ck(sml_set_tisr(h, tisr_grid), 'sml_set_tisr')
  end subroutine

  !> mldivide(A,B), src/mod_linalg.f90:109-151: A*X = B, B becomes X when info == 0
  subroutine mldivide(A, B)
    real(kind=dp), intent(inout) :: A(:,:), B(:,:)
    integer :: n, l, info
    n = size(A, 1)
    l = size(B, 1)
    if (n /= l) then
       print *, 'Column of A is not the same size of column of B. Cant compute solution returning A and B unchanged'
       return
    end if
    info = sml_mldivide(h, A, max(1, n), B, max(1, n), n, size(B, 2))
    if (info < 0) call ck(info, 'sml_mldivide')
    if (info /= 0) then
       print *, 'something went wrong with dgesv info = ', info
       print *, 'B is not the solution'
    end if
  end subroutine

  !> rolling_average_over_a_period_2d(grid,period), src/mod_utilities.f90:1773-1815 (called on a row section of
  !> reservoir%trainingdata, src/mod_slab_ocean_reservoir.f90:398,452: a non-contiguous actual is copied in/out by
  !> the compiler, so the dummy is contiguous here)
  subroutine rolling_average_over_a_period_2d(grid, period)
    real(kind=dp), intent(inout), contiguous :: grid(:,:)
    integer, intent(in) :: period
    call ck(sml_rolling_average_2d(h, grid, size(grid, 1), size(grid, 1), size(grid, 2), period, 1), &
            'sml_rolling_average_2d')
  end subroutine

end module speedyml_gpu
