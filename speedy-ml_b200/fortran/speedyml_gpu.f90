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
!> region; the first predict call of a hybrid step launches ONE batched kernel for all local regions and reads
!> every local outvec back in ONE copy, later calls in the same step only hand out that region's row
!> (reservoir%outvec stays valid for host code such as the diagnostics writers).  sendrecievegrid() closes the step.
!>
!> Multi-rank (numprocs > 1, one rank per GPU of one node): call gpu_comm_bootstrap(mpi_world) once after
!> gpu_engine_finalize.  It hands the engine ONE host primitive -- MPI_Allgather of a few bytes on the reference's
!> communicator -- and the engine connects the ranks' exchange blocks over NVLink.  From then on sendrecievegrid is
!> complete on every rank: the gather of the outvecs (src/mpires.f90:346-454), the scatter of the forecast (:606-739)
!> and the broadcast of run_speedy (:744) are done by the engine's kernels; ranks other than the root enqueue their
!> part and never wait for the host model.
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
  public :: gpu_comm_bootstrap, gpu_synchronize_all, gpu_grid_status

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
  integer, save     :: my_rank = 0, my_numprocs = 1
  integer, save     :: bootstrap_comm = -1     !< MPI communicator handed to gpu_comm_bootstrap
  integer, allocatable, save :: local_of(:)    !< region id -> 1-based local index (0: not on this rank)
  real(dp), allocatable, save :: outvec_cache(:,:), ocean_outvec_cache(:,:)   !< (chunk_size_prediction, local regions)

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
       import :: c_ptr, c_int
       type(c_ptr), value :: h
       integer(c_int), value :: timestep
       type(c_ptr), value :: w4d, w2d, wprecip, wsst       ! c_loc(array), or c_null_ptr on ranks that take no grids
     end function
     integer(c_int) function sml_step_exchange_end(h, timestep, f4d, f2d, tisr) bind(C, name='sml_step_exchange_end')
       import :: c_ptr, c_int
       type(c_ptr), value :: h
       integer(c_int), value :: timestep
       type(c_ptr), value :: f4d, f2d, tisr                 ! c_null_ptr on ranks other than the root
     end function
     integer(c_int) function sml_comm_bootstrap(h, allgather, ctx) bind(C, name='sml_comm_bootstrap')
       import :: c_ptr, c_int, c_funptr
       type(c_ptr), value :: h
       type(c_funptr), value :: allgather
       type(c_ptr), value :: ctx
     end function
     integer(c_int) function sml_grid_status(h, bits) bind(C, name='sml_grid_status')
       import :: c_ptr, c_int
       type(c_ptr), value :: h
       integer(c_int), intent(out) :: bits
     end function
     integer(c_int) function sml_set_run_speedy(h, run_speedy) bind(C, name='sml_set_run_speedy')
       import :: c_ptr, c_int
       type(c_ptr), value :: h
       integer(c_int), value :: run_speedy
     end function
     integer(c_int) function sml_run_speedy(h, run_speedy) bind(C, name='sml_run_speedy')
       import :: c_ptr, c_int
       type(c_ptr), value :: h
       integer(c_int), intent(out) :: run_speedy
     end function
     integer(c_int) function sml_outvec_get_all(h, kind, slab) bind(C, name='sml_outvec_get_all')
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: h
       integer(c_int), value :: kind
       real(c_double), intent(out) :: slab(*)
     end function
     integer(c_int) function sml_num_local_regions(h) bind(C, name='sml_num_local_regions')
       import :: c_ptr, c_int
       type(c_ptr), value :: h
     end function
     integer(c_int) function sml_local_region_ids(h, ids) bind(C, name='sml_local_region_ids')
       import :: c_ptr, c_int
       type(c_ptr), value :: h
       integer(c_int), intent(out) :: ids(*)
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
    my_rank = model_parameters%irank
    my_numprocs = model_parameters%numprocs
    block
      integer(c_int), allocatable :: ids(:)
      integer :: i, nloc
      nloc = sml_num_local_regions(h)
      allocate(ids(nloc))
      call ck(sml_local_region_ids(h, ids), 'sml_local_region_ids')
      if (allocated(local_of)) deallocate(local_of)
      allocate(local_of(0:model_parameters%number_of_regions-1))
      local_of = 0
      do i = 1, nloc
         local_of(ids(i)) = i
      end do
    end block
  end subroutine

  !> the one host primitive sml_comm_bootstrap needs, on the reference's communicator (mpi_res%mpi_world)
  integer(c_int) function allgather_cb(ctx, send, recv, nbytes) bind(C)
    use mpi
    type(c_ptr), value :: ctx, send, recv
    integer(c_int), value :: nbytes
    character(kind=c_char), pointer :: s(:), r(:)
    integer :: ierr
    call c_f_pointer(send, s, [nbytes])
    call c_f_pointer(recv, r, [nbytes * my_numprocs])
    call MPI_Allgather(s, nbytes, MPI_BYTE, r, nbytes, MPI_BYTE, bootstrap_comm, ierr)
    allgather_cb = ierr
  end function

  !> multi-rank runs: once, after gpu_engine_finalize, on every rank (replaces nothing; the reference's ranks meet in
  !> every MPI call of sendrecievegrid, these meet here once).  Ranks of ONE node, numprocs <= 8.
  subroutine gpu_comm_bootstrap(mpi_world)
    integer, intent(in) :: mpi_world
    if (my_numprocs == 1) return
    bootstrap_comm = mpi_world
    call ck(sml_comm_bootstrap(h, c_funloc(allgather_cb), c_null_ptr), 'sml_comm_bootstrap')
  end subroutine

  !> sticky status bits of the assembled grids (non-finite values, SPEEDY's input bounds of src/ppo_iogrid.f90:562-577)
  integer function gpu_grid_status()
    integer(c_int) :: bits
    call ck(sml_grid_status(h, bits), 'sml_grid_status')
    gpu_grid_status = bits
  end function

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

  !> The same for EVERY local region in one call -- the region loops of initialize_prediction / start_prediction
  !> (src/mod_reservoir.f90:818-824, :951) collapse into one batched launch per time step instead of `length` launches
  !> and two state copies per region.  inputs holds the regions' (reservoir_numinputs, length) series back to back in
  !> local region order, offsets(i) the 0-based position of region i's block; states start from and return to
  !> reservoirs(i)%saved_state.
  subroutine gpu_synchronize_all(reservoirs, inputs, offsets, length, kind)
    type(reservoir_type), intent(inout) :: reservoirs(:)
    real(kind=dp), intent(in) :: inputs(:)
    integer(c_int64_t), intent(in), target :: offsets(:)
    integer, intent(in) :: length
    integer, intent(in), optional :: kind
    integer(c_int) :: k
    integer :: i
    k = SML_ATMO
    if (present(kind)) k = kind
    do i = 1, size(reservoirs)
       call ck(sml_state_set(h, k, reservoirs(i)%assigned_region, reservoirs(i)%saved_state), 'sml_state_set')
    end do
    call ck(sml_synchronize(h, k, SML_ALL_REGIONS, inputs, 0, length, c_loc(offsets)), 'sml_synchronize(all)')
    do i = 1, size(reservoirs)
       call ck(sml_state_get(h, k, reservoirs(i)%assigned_region, reservoirs(i)%saved_state), 'sml_state_get')
    end do
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
       ! ONE read-back for all local regions (a blocking 1 KB copy per region would cost more than the step)
       if (.not. allocated(outvec_cache)) allocate(outvec_cache(reservoir%chunk_size_prediction, sml_num_local_regions(h)))
       call ck(sml_outvec_get_all(h, SML_ATMO, outvec_cache), 'sml_outvec_get_all')
    end if
    reservoir%outvec = outvec_cache(:, local_of(reservoir%assigned_region))
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
       if (.not. allocated(ocean_outvec_cache)) &
            allocate(ocean_outvec_cache(reservoir%chunk_size_prediction, sml_num_local_regions(h)))
       call ck(sml_outvec_get_all(h, SML_OCEAN, ocean_outvec_cache), 'sml_outvec_get_all(ocean)')
    end if
    reservoir%outvec = ocean_outvec_cache(:, local_of(reservoir%assigned_region))
  end subroutine

  !> sendrecievegrid(res,timestep,ocean_model), src/mpires.f90:218-804.  The MPI star through the root, the
  !> clamps (:456-490) and the feedback / local_model rebuild (:581-604,749-791) happen on the devices; the
  !> root still owns NetCDF output and run_model (:565-569), passed in as procedure arguments so this module
  !> does not depend on mod_io / speedy_res_interface.  tisr_grid: get_tisr_by_date's field for timestep-1.
  !> Complete at numprocs > 1 (after gpu_comm_bootstrap): the root copies the assembled grids out, runs the host
  !> model and hands [forecast | tisr | run_speedy] to the engine, which pushes the block into every rank's landing
  !> buffer; the other ranks only enqueue the grid assembly and the device-side wait for that block.  run_speedy
  !> (MPI_Bcast at :744) comes back on every rank when check_run_speedy is present and true -- that read is the only
  !> point where a non-root rank waits for the root.
  subroutine sendrecievegrid(res, timestep, ocean_model, run_model, write_prediction, tisr_grid, check_run_speedy)
    type(main_type), intent(inout) :: res
    integer, intent(in) :: timestep
    logical, intent(in) :: ocean_model
    real(kind=dp), intent(in), target :: tisr_grid(:,:)
    logical, intent(in), optional :: check_run_speedy
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
    real(kind=dp), allocatable, target :: wholegrid4d(:,:,:,:), wholegrid2d(:,:), wholegrid_precip(:,:), wholegrid_sst(:,:)
    real(kind=dp), allocatable, target :: forecast_4d(:,:,:,:), forecast_2d(:,:)
    integer :: nv
    integer(c_int) :: rc, flag
    if (my_numprocs > 1 .and. my_rank /= 0) then
       ! enqueue only: grid assembly (waits for every rank's outvecs on the device), then the wait for the root's forecast
       call ck(sml_step_exchange_begin(h, timestep, c_null_ptr, c_null_ptr, c_null_ptr, c_null_ptr), 'sml_step_exchange_begin')
       call ck(sml_step_exchange_end(h, timestep, c_null_ptr, c_null_ptr, c_null_ptr), 'sml_step_exchange_end')
    else
       nv = res%model_parameters%full_predictvars
       allocate(wholegrid4d(nv, xgrid, ygrid, zgrid), wholegrid2d(xgrid, ygrid))
       allocate(wholegrid_precip(xgrid, ygrid), wholegrid_sst(xgrid, ygrid))
       allocate(forecast_4d(nv, xgrid, ygrid, zgrid), forecast_2d(xgrid, ygrid))
       rc = sml_step_exchange_begin(h, timestep, c_loc(wholegrid4d), c_loc(wholegrid2d), c_loc(wholegrid_precip), &
                                    c_loc(wholegrid_sst))
       call ck(rc, 'sml_step_exchange_begin')
       if (rc > 0) then   ! a non-finite value in the assembled grid: SPEEDY must not be run on it
          print *, 'hybrid grid is not finite, stopping hybrid prediction'
          res%model_parameters%run_speedy = .False.
       end if
       call write_prediction(res, timestep, wholegrid4d, wholegrid2d, wholegrid_precip, wholegrid_sst)
       forecast_4d = 0.0_dp
       forecast_2d = 0.0_dp
       if (.not. res%model_parameters%ml_only .and. res%model_parameters%run_speedy) then
          call run_model(res%model_parameters, timestep, wholegrid4d, wholegrid2d, wholegrid_sst, forecast_4d, forecast_2d)
       end if
       call ck(sml_set_run_speedy(h, merge(1_c_int, 0_c_int, res%model_parameters%run_speedy)), 'sml_set_run_speedy')
       call ck(sml_step_exchange_end(h, timestep, c_loc(forecast_4d), c_loc(forecast_2d), c_loc(tisr_grid)), &
               'sml_step_exchange_end')
    end if
    if (present(check_run_speedy)) then
       if (check_run_speedy) then
          call ck(sml_run_speedy(h, flag), 'sml_run_speedy')
          res%model_parameters%run_speedy = flag /= 0
       end if
    end if
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
