!> dump_golden.f90 -- golden vectors of the SPEEDY-ML reservoir hot path, produced BY THE REFERENCE ITSELF.
!>
!> Purpose.  The engine in this repository is checked against a CPU restatement of the reference (oracle/), because the
!> reference cannot be built where the engine is developed (no Fortran compiler, MPI, MKL, ARPACK or NetCDF there).  This
!> program closes the loop on a host that HAS the reference's toolchain: it links against the reference's own objects
!> (mod_utilities, resdomain, mod_linalg, mod_reservoir ...), calls the reference procedures on small seeded cases and
!> writes inputs AND outputs to one flat binary.  Copy that file to tests/golden/reference_v1.bin and
!> tests/test_reference_fixtures.py checks the oracle (and, on a GPU box, the engine) against it -- at which point the
!> oracle is pinned by the reference instead of by its own second restatement.
!>
!> Build (see Makefile.fragment): compile inside the reference's src/ after `make` has produced the module files, e.g.
!>     mpif90 $(COMOTT) -c dump_golden.f90 && mpif90 -o dump_golden dump_golden.o <reference objects except parallelmain.o> $(COMOTT)
!>     mpirun -np 1 ./dump_golden reference_v1.bin
!>
!> File format (tests/reffix_io.py reads it): int32 1 (endianness probe: the reference's gfortran flags include
!> -fconvert=swap), then records until end of file:
!>     character(len=24) name | int32 kind (1 = int32, 2 = real64) | int32 count | data(count)
!> Arrays are written in Fortran (column-major) order.
!>
!> Cases (all with noisemag = 0: the random_number stream is the one thing that cannot be reproduced):
!>   idx_*      res_domain index functions for regions 0, 23, 555, 1128, 1151 of the 1152-region tiling and 145 of 288
!>   pd_*       processor_decomposition(irank, numprocs) for (0,8), (3,8), (5,7)
!>   p_*        one region (555, m = 600): mklsparse, synchronize over 5 inputs, then 3 predict calls (state, outvec)
!>   t_*        the same region without precipitation: initialize_chunk_training, reservoir_layer_chunking_hybrid (one
!>              phase), fit_chunk_hybrid -> states_x_states_aug, states_x_trainingdata_aug, wout
!>   md_*       mldivide on a 40 x 40 system with 3 right-hand sides
program dump_golden
  use mod_utilities, only : dp, reservoir_type, grid_type, model_parameters_type
  use resdomain, only : initializedomain, getxyresextent, getoverlapindices, get_trainingdataindices, &
                        processor_decomposition
  use mod_linalg, only : mklsparse, mldivide
  use mod_reservoir, only : synchronize, predict, initialize_chunk_training, reservoir_layer_chunking_hybrid, &
                            fit_chunk_hybrid
  implicit none

  integer, parameter :: lun = 77
  character(len=256) :: outfile
  integer(kind=8) :: lcg_state = 20251018_8

  call get_command_argument(1, outfile)
  if (len_trim(outfile) == 0) outfile = 'reference_v1.bin'
  open(unit=lun, file=trim(outfile), access='stream', form='unformatted', status='replace')
  write(lun) 1

  call dump_index_tables()
  call dump_predict_case()
  call dump_training_case()
  call dump_mldivide_case()

  close(lun)
  print *, 'wrote ', trim(outfile)

contains

  !> 48-bit linear congruential generator (drand48's constants): the inputs are written to the file, so the consumer
  !> never needs to reproduce it
  real(kind=dp) function lcg()
    lcg_state = iand(lcg_state * 25214903917_8 + 11_8, 281474976710655_8)
    lcg = real(lcg_state, kind=dp) / 281474976710656.0_dp
  end function

  subroutine put_i(name, a)
    character(len=*), intent(in) :: name
    integer, intent(in) :: a(:)
    character(len=24) :: nm
    nm = name
    write(lun) nm, 1, size(a), a
  end subroutine

  subroutine put_r(name, a)
    character(len=*), intent(in) :: name
    real(kind=dp), intent(in) :: a(:)
    character(len=24) :: nm
    nm = name
    write(lun) nm, 2, size(a), a
  end subroutine

  subroutine dump_index_tables()
    integer :: regions(5), r, i, v(18), nreg
    logical :: pole, periodic
    character(len=24) :: nm
    type(model_parameters_type) :: mp
    regions = (/ 0, 23, 555, 1128, 1151 /)
    do i = 1, 6
       if (i <= 5) then
          r = regions(i); nreg = 1152
       else
          r = 145; nreg = 288
       end if
       v = 0
       ! v(1:6)  getxyresextent: xstart, xend, ystart, yend, xchunk, ychunk
       call getxyresextent(nreg, r, v(1), v(2), v(3), v(4), v(5), v(6))
       ! v(7:14) getoverlapindices: xstart, xend, ystart, yend, xchunk, ychunk, pole, periodic
       call getoverlapindices(nreg, r, 1, v(7), v(8), v(9), v(10), v(11), v(12), pole, periodic, .false.)
       v(13) = merge(1, 0, pole)
       v(14) = merge(1, 0, periodic)
       ! v(15:18) get_trainingdataindices: xstart, xend, ystart, yend
       call get_trainingdataindices(nreg, r, 1, v(15), v(16), v(17), v(18))
       write(nm, '(a,i0,a,i0)') 'idx_', nreg, '_', r
       call put_i(trim(nm), v)
    end do
    ! processor_decomposition(model_parameters): region_indices of (irank, numprocs)
    mp%number_of_regions = 1152
    mp%irank = 0; mp%numprocs = 8
    call processor_decomposition(mp);  call put_i('pd_0_8', mp%region_indices);  deallocate(mp%region_indices)
    mp%irank = 3; mp%numprocs = 8
    call processor_decomposition(mp);  call put_i('pd_3_8', mp%region_indices);  deallocate(mp%region_indices)
    mp%irank = 5; mp%numprocs = 7
    call processor_decomposition(mp);  call put_i('pd_5_7', mp%region_indices);  deallocate(mp%region_indices)
  end subroutine

  !> the fields allocate_res_new / trained_reservoir_prediction would set for one bottom-level atmosphere reservoir
  subroutine make_region(reservoir, grid, model_parameters, region, m, precip)
    type(reservoir_type), intent(inout) :: reservoir
    type(grid_type), intent(inout) :: grid
    type(model_parameters_type), intent(inout) :: model_parameters
    integer, intent(in) :: region, m
    logical, intent(in) :: precip
    integer :: ixy, rxy, q, i, j, counter, e, L
    model_parameters%number_of_regions = 1152
    model_parameters%overlap = 1
    model_parameters%num_vert_levels = 1
    model_parameters%vert_loc_overlap = 0
    model_parameters%ml_only = .false.
    model_parameters%precip_bool = precip
    model_parameters%slab_ocean_model_bool = .false.
    model_parameters%outvec_component_contribs = .false.
    model_parameters%full_predictvars = 4
    model_parameters%timestep = 6
    model_parameters%using_prior = .true.
    call initializedomain(1152, region, 1, 1, 1, 0, grid)
    reservoir%assigned_region = region
    reservoir%local_predictvars = 4
    reservoir%local_heightlevels_input = grid%inputzchunk
    reservoir%local_heightlevels_res = grid%reszchunk
    reservoir%logp_bool = .true.
    grid%logp_bool = .true.
    reservoir%tisr_input_bool = .true.
    reservoir%precip_bool = precip
    reservoir%precip_input_bool = precip
    reservoir%sst_bool = .false.
    reservoir%sst_bool_input = .false.
    reservoir%sst_bool_prediction = .false.
    ixy = grid%inputxchunk * grid%inputychunk
    rxy = grid%resxchunk * grid%resychunk
    reservoir%chunk_size_prediction = rxy * 4 * grid%reszchunk + rxy + merge(rxy, 0, precip)
    reservoir%chunk_size = reservoir%chunk_size_prediction
    reservoir%chunk_size_speedy = rxy * 4 * grid%reszchunk + rxy
    reservoir%reservoir_numinputs = ixy * grid%inputzchunk * 4 + ixy + merge(ixy, 0, precip) + ixy
    reservoir%m = m
    q = nint(dble(m) / dble(reservoir%reservoir_numinputs))
    reservoir%n = q * reservoir%reservoir_numinputs
    reservoir%deg = 6
    reservoir%density = reservoir%deg / real(m, kind=dp)
    reservoir%k = reservoir%density * reservoir%n * reservoir%n
    reservoir%leakage = 1.0_dp
    reservoir%noisemag = 0.0_dp
    reservoir%beta_res = 0.001_dp
    reservoir%beta_model = 1.0_dp
    reservoir%prior_val = 0.0_dp
    ! mean / std: one slot per (variable, level), then logp, tisr, (precip)
    L = 4 * grid%inputzchunk
    L = L + 1; grid%logp_mean_std_idx = L
    L = L + 1; grid%tisr_mean_std_idx = L
    if (precip) then
       L = L + 1; grid%precip_mean_std_idx = L
    end if
    allocate(grid%mean(L), grid%std(L))
    do i = 1, L
       grid%mean(i) = 10.0_dp * (lcg() - 0.5_dp)
       grid%std(i) = 0.5_dp + lcg()
    end do
    allocate(reservoir%rows(reservoir%k), reservoir%cols(reservoir%k), reservoir%vals(reservoir%k))
    allocate(reservoir%win(reservoir%n, reservoir%reservoir_numinputs))
    allocate(reservoir%wout(reservoir%chunk_size_prediction, reservoir%n + reservoir%chunk_size_speedy))
    allocate(reservoir%feedback(reservoir%reservoir_numinputs), reservoir%local_model(reservoir%chunk_size_speedy))
    allocate(reservoir%outvec(reservoir%chunk_size_prediction))
    ! adjacency: rounds of cyclic shifts (distinct rows per round, like makesparse), values scaled to keep |x| < 1
    counter = reservoir%k / reservoir%n
    do e = 1, reservoir%k
       i = mod(e - 1, reservoir%n)
       j = (e - 1) / reservoir%n
       reservoir%rows(e) = i + 1
       reservoir%cols(e) = mod(i * 7 + 13 * j + int(lcg() * reservoir%n), reservoir%n) + 1
       reservoir%vals(e) = 0.2_dp * lcg()
    end do
    reservoir%win = 0.0_dp
    do i = 1, reservoir%reservoir_numinputs
       do j = (i - 1) * q + 1, i * q
          reservoir%win(j, i) = 0.5_dp * (-1.0_dp + 2.0_dp * lcg())
       end do
    end do
    do j = 1, size(reservoir%wout, 2)
       do i = 1, size(reservoir%wout, 1)
          reservoir%wout(i, j) = (lcg() - 0.5_dp) / sqrt(real(size(reservoir%wout, 2), kind=dp))
       end do
    end do
  end subroutine

  subroutine dump_predict_case()
    type(reservoir_type) :: reservoir
    type(grid_type) :: grid
    type(model_parameters_type) :: model_parameters
    real(kind=dp), allocatable :: x(:), inputs(:,:), lm(:)
    integer :: i, t
    character(len=24) :: nm
    call make_region(reservoir, grid, model_parameters, 555, 600, .true.)
    call put_i('p_dims', (/ reservoir%n, reservoir%k, reservoir%reservoir_numinputs, reservoir%chunk_size_prediction, &
                            reservoir%chunk_size_speedy, size(grid%mean) /))
    call put_i('p_rows', reservoir%rows); call put_i('p_cols', reservoir%cols); call put_r('p_vals', reservoir%vals)
    call put_r('p_win', reshape(reservoir%win, (/ size(reservoir%win) /)))
    call put_r('p_wout', reshape(reservoir%wout, (/ size(reservoir%wout) /)))
    call put_r('p_mean', grid%mean); call put_r('p_std', grid%std)
    call mklsparse(reservoir)
    allocate(x(reservoir%n), inputs(reservoir%reservoir_numinputs, 5), lm(reservoir%chunk_size_speedy))
    do i = 1, reservoir%n
       x(i) = 0.2_dp * (lcg() - 0.5_dp)
    end do
    do t = 1, 5
       do i = 1, reservoir%reservoir_numinputs
          inputs(i, t) = 2.0_dp * (lcg() - 0.5_dp)
       end do
    end do
    call put_r('p_x0', x)
    call put_r('p_sync_inputs', reshape(inputs, (/ size(inputs) /)))
    call synchronize(reservoir, inputs, x, 5)
    call put_r('p_x_sync', x)
    do t = 1, 3
       do i = 1, reservoir%reservoir_numinputs
          reservoir%feedback(i) = 2.0_dp * (lcg() - 0.5_dp)
       end do
       do i = 1, reservoir%chunk_size_speedy
          reservoir%local_model(i) = 2.0_dp * (lcg() - 0.5_dp)
       end do
       lm = reservoir%local_model
       write(nm, '(a,i0)') 'p_feedback_', t;    call put_r(trim(nm), reservoir%feedback)
       write(nm, '(a,i0)') 'p_local_model_', t; call put_r(trim(nm), lm)
       call predict(reservoir, model_parameters, grid, x, lm)
       write(nm, '(a,i0)') 'p_x_', t;      call put_r(trim(nm), x)
       write(nm, '(a,i0)') 'p_outvec_', t; call put_r(trim(nm), reservoir%outvec)
    end do
  end subroutine

  subroutine dump_training_case()
    type(reservoir_type) :: reservoir
    type(grid_type) :: grid
    type(model_parameters_type) :: model_parameters
    real(kind=dp), allocatable :: td(:,:), im(:,:)
    integer :: i, t, ncols
    call make_region(reservoir, grid, model_parameters, 555, 600, .false.)
    ! 20 batches of 7 columns after 3 discarded ones: traininglength / discardlength are in hours, one column per timestep
    ncols = 143
    model_parameters%discardlength = 3 * model_parameters%timestep
    model_parameters%traininglength = ncols * model_parameters%timestep
    allocate(td(reservoir%reservoir_numinputs, ncols), im(reservoir%chunk_size_speedy, ncols))
    do t = 1, ncols
       do i = 1, reservoir%reservoir_numinputs
          td(i, t) = 2.0_dp * (lcg() - 0.5_dp)
       end do
       do i = 1, reservoir%chunk_size_speedy
          im(i, t) = 2.0_dp * (lcg() - 0.5_dp)
       end do
    end do
    call put_i('t_dims', (/ reservoir%n, reservoir%k, reservoir%reservoir_numinputs, reservoir%chunk_size_prediction, &
                            reservoir%chunk_size_speedy, size(grid%mean), ncols, model_parameters%discardlength / model_parameters%timestep /))
    call put_i('t_rows', reservoir%rows); call put_i('t_cols', reservoir%cols); call put_r('t_vals', reservoir%vals)
    call put_r('t_win', reshape(reservoir%win, (/ size(reservoir%win) /)))
    call put_r('t_mean', grid%mean); call put_r('t_std', grid%std)
    call put_r('t_trainingdata', reshape(td, (/ size(td) /)))
    call put_r('t_imperfect', reshape(im, (/ size(im) /)))
    call mklsparse(reservoir)
    call initialize_chunk_training(reservoir, model_parameters)
    call put_i('t_batch_size', (/ reservoir%batch_size /))
    call reservoir_layer_chunking_hybrid(reservoir, model_parameters, grid, td, im)
    call put_r('t_states_x_states', reshape(reservoir%states_x_states_aug, (/ size(reservoir%states_x_states_aug) /)))
    call put_r('t_states_x_tdata', reshape(reservoir%states_x_trainingdata_aug, (/ size(reservoir%states_x_trainingdata_aug) /)))
    call put_r('t_betas', (/ reservoir%beta_res, reservoir%beta_model, reservoir%prior_val /))
    call fit_chunk_hybrid(reservoir, model_parameters, grid)
    call put_r('t_wout', reshape(reservoir%wout, (/ size(reservoir%wout) /)))
  end subroutine

  subroutine dump_mldivide_case()
    real(kind=dp) :: A(40, 40), B(40, 3), A0(40, 40), B0(40, 3)
    integer :: i, j
    do j = 1, 40
       do i = 1, 40
          A(i, j) = lcg() - 0.5_dp
       end do
       A(j, j) = A(j, j) + 4.0_dp
    end do
    do j = 1, 3
       do i = 1, 40
          B(i, j) = lcg() - 0.5_dp
       end do
    end do
    A0 = A; B0 = B
    call put_r('md_A', reshape(A0, (/ 1600 /)))
    call put_r('md_B', reshape(B0, (/ 120 /)))
    call mldivide(A, B)
    call put_r('md_X', reshape(B, (/ 120 /)))
  end subroutine

end program dump_golden
