!> nekstab_b200_shim.f90 -- replacement bodies for the hot-path routines of nekStab that route to the CUDA library.
!!
!! Replaces, with identical names and argument lists, the routines of core/krylov_subspace.f (:24-258),
!! core/matvec.f (:64-159) and update_hessenberg_matrix (core/krylov_decomposition.f:116-202).  Everything above them
!! (arnoldi_factorization, krylov_schur, schur_condensation, ts_gmres, newton_krylov, the LAPACK wrappers, outpost_ks)
!! stays the reference's own Fortran.  `type(krylov_vector)` gains one integer member, `slot`, the handle of the
!! device-resident copy; its host arrays become a lazily synchronised mirror (nsb_b200_pull / nsb_b200_push at the
!! field-access sites listed in SURVEY.md 8b: core/eigensolvers.f:226-232,267-273,282,554-564; core/newton_krylov.f:56-75).
!! NOT compiled in this repository's build container (no Fortran compiler); kept declarative and small on purpose.
      subroutine nsb_b200_check(ierr, where)
      implicit none
      integer ierr
      character(len=*) where
      if (ierr .ne. 0) then
         write(6,*) 'nekstab_b200 error in ', where   ! message text: nsb_last_error()
         call nek_end                                   ! the reference's only error convention (core/krylov_subspace.f:53)
      endif
      end subroutine

!     Called once from nekStab_init (core/usr_extra.f:72) after bm1s and the sponge are set.
      subroutine nsb_b200_setup
      use nekstab_b200_c
      use krylov_subspace
      implicit none
      include 'SIZE'
      include 'TOTAL'
      character(kind=c_char) :: id(128)
      integer(c_long_long) :: glo(lx1*ly1*lz1*lelv)
      common /nsb_glo/ glo                              ! filled by the case from Nek's glo_num (setvert2d/3d)
      integer ierr
      if (np .gt. 1) then
         if (nid .eq. 0) ierr = nsb_comm_unique_id(id)
         call bcast(id, 128)                            ! Nek5000's MPI_Bcast wrapper (core/matvec.f:18 uses the same)
         call nsb_b200_check(nsb_comm_init(nid, np, id, mod(nid, 8)), 'nsb_comm_init')
      endif
      call nsb_b200_check(nsb_init(ldim, lx1, lxd, lx2, nelv, int(nelgv, c_long_long), xm1, ym1, zm1, &
                                   v1mask, v2mask, v3mask, glo, mod(nid, 8)), 'nsb_init')
      call nsb_b200_check(nsb_set_params(param(2), param(1), param(22), param(21), 0, 0), 'nsb_set_params')
      call nsb_b200_check(nsb_set_weights(bm1s), 'nsb_set_weights')
      if (spng_str .ne. 0) call nsb_b200_check(nsb_set_sponge(spng_fun), 'nsb_set_sponge')
      call nsb_b200_check(nsb_set_ifvcor(merge(1, 0, ifvcor), -1), 'nsb_set_ifvcor')
      ! [PRESSURE] preconditioner = semg_xxt in every shipped .par (1cyl.par:28): the multilevel Schwarz class (kind 1);
      ! kind 0 keeps Jacobi.  nagg = 0: automatic number of coarse aggregates.
      call nsb_b200_check(nsb_set_pressure_preconditioner(1, 0), 'nsb_set_pressure_preconditioner')
      call nsb_b200_check(nsb_vec_alloc(k_dim + 8), 'nsb_vec_alloc')
      end subroutine

      subroutine krylov_inner_product(alpha, p, q)       ! core/krylov_subspace.f:24
      use nekstab_b200_c
      use krylov_subspace
      implicit none
      type(krylov_vector), intent(in) :: p, q
      real, intent(out) :: alpha
      call nsb_b200_check(nsb_vec_inner_product(p%slot, q%slot, alpha), 'krylov_inner_product')
      if (uparam(1) .eq. 2.1) alpha = alpha + p%time * q%time
      end subroutine

      subroutine krylov_norm(alpha, p)                    ! :58
      use nekstab_b200_c
      use krylov_subspace
      implicit none
      type(krylov_vector), intent(in) :: p
      real, intent(out) :: alpha
      call nsb_b200_check(nsb_vec_norm(p%slot, alpha), 'krylov_norm')
      if (uparam(1) .eq. 2.1) alpha = sqrt(alpha**2 + p%time**2)
      end subroutine

      subroutine krylov_normalize(p, alpha)               ! :71
      use nekstab_b200_c
      use krylov_subspace
      implicit none
      type(krylov_vector), intent(inout) :: p
      real, intent(out) :: alpha
      call nsb_b200_check(nsb_vec_normalize(p%slot, alpha), 'krylov_normalize')
      p%time = p%time / alpha
      end subroutine

      subroutine krylov_cmult(p, alpha)                   ! :90
      use nekstab_b200_c
      use krylov_subspace
      implicit none
      type(krylov_vector) :: p
      real alpha
      call nsb_b200_check(nsb_vec_cmult(p%slot, alpha), 'krylov_cmult')
      p%time = p%time * alpha
      end subroutine

      subroutine krylov_add2(p, q)                        ! :116
      use nekstab_b200_c
      use krylov_subspace
      implicit none
      type(krylov_vector) :: p, q
      call nsb_b200_check(nsb_vec_add2(p%slot, q%slot), 'krylov_add2')
      p%time = p%time + q%time
      end subroutine

      subroutine krylov_sub2(p, q)                        ! :142
      use nekstab_b200_c
      use krylov_subspace
      implicit none
      type(krylov_vector) :: p, q
      call nsb_b200_check(nsb_vec_sub2(p%slot, q%slot), 'krylov_sub2')
      p%time = p%time - q%time
      end subroutine

      subroutine krylov_zero(p)                           ! :166
      use nekstab_b200_c
      use krylov_subspace
      implicit none
      type(krylov_vector) :: p
      call nsb_b200_check(nsb_vec_zero(p%slot), 'krylov_zero')
      p%time = 0.0d0
      end subroutine

      subroutine krylov_copy(p, q)                        ! :190
      use nekstab_b200_c
      use krylov_subspace
      implicit none
      type(krylov_vector) :: p, q
      call nsb_b200_check(nsb_vec_copy(p%slot, q%slot), 'krylov_copy')
      p%time = q%time
      end subroutine

      subroutine krylov_matmul(dq, Q, yvec, k)            ! :214
      use nekstab_b200_c
      use krylov_subspace
      implicit none
      integer :: k
      type(krylov_vector) :: dq
      type(krylov_vector), dimension(k) :: Q
      real, dimension(k) :: yvec
      call nsb_b200_check(nsb_basis_gemv(k, Q(1)%slot, yvec, dq%slot), 'krylov_matmul')   ! Q(1:k) are consecutive slots
      dq%time = dot_product(Q(1:k)%time, yvec(1:k))
      end subroutine

      subroutine update_hessenberg_matrix(H, f, q, k)     ! core/krylov_decomposition.f:116
      use nekstab_b200_c
      use krylov_subspace
      implicit none
      integer, intent(in) :: k
      real, dimension(k+1, k), intent(inout) :: H
      type(krylov_vector), dimension(k) :: q
      type(krylov_vector) :: f
      call nsb_b200_check(nsb_orthonormalize(k, q(1)%slot, f%slot, H(1, k)), 'update_hessenberg_matrix')
      end subroutine

      subroutine matvec(f, q)                             ! core/matvec.f:64
      use nekstab_b200_c
      use krylov_subspace
      implicit none
      include 'SIZE'
      include 'TOTAL'
      type(krylov_vector) :: q, f
      logical, save :: init = .false.
      integer(c_int) :: mode, nst
      real(c_double) :: ddt, ct
      if (.not. init) then                                ! prepare_linearized_solver, core/matvec.f:1-52,115-118
         call nsb_b200_check(nsb_set_baseflow(ubase, vbase, wbase), 'nsb_set_baseflow')
         call nsb_b200_check(nsb_prepare_linearized_solver(param(10), param(26), ddt, nst, ct), 'prepare_linearized_solver')
         dt = ddt; nsteps = nst; ctarg = ct; param(12) = -abs(dt)
         init = .true.
      endif
      mode = -1
      if (uparam(1) .ge. 3.0 .and. uparam(1) .lt. 3.2) then; evop = 'd'; mode = NSB_DIRECT; endif
      if (uparam(1) .ge. 3.2 .and. uparam(1) .lt. 3.3) then; evop = 'a'; mode = NSB_ADJOINT; endif
      if (uparam(1) .ge. 3.3 .and. uparam(1) .lt. 3.4) then; evop = 'p'; mode = NSB_DIRECT_ADJOINT; endif
      if (floor(uparam(1)) .eq. 4) mode = NSB_FORCE_SENS
      if (floor(uparam(1)) .eq. 2) then; evop = 'n'; mode = NSB_NEWTON; init = .false.; endif
      call nsb_b200_check(nsb_matvec(mode, q%slot, f%slot), 'matvec')
      f%time = 0.0d0
      end subroutine

      subroutine nonlinear_forward_map(f, q)             ! core/newton_krylov.f:336-378
      use nekstab_b200_c
      use krylov_subspace
      implicit none
      include 'SIZE'
      include 'TOTAL'
      type(krylov_vector) :: f, q
      integer(c_int) :: nst
      real(c_double) :: ddt, ct
      ! newton_krylov re-prepares the solver on every iterate (core/newton_krylov.f:69): CFL of the current q
      call nsb_b200_check(nsb_prepare_solver_from_slot(q%slot, param(10), param(26), ddt, nst, ct), 'prepare_linearized_solver')
      dt = ddt; nsteps = nst; ctarg = ct; param(12) = -abs(dt)
      call nsb_b200_check(nsb_nonlinear_forward_map(q%slot, f%slot), 'nonlinear_forward_map')   ! f = phi_T(q) - q ; ubase <- q
      f%time = 0.0d0
      end subroutine

!     Host mirror <-> device slot, to be called at the field-access sites listed in SURVEY.md 8b (seeding from vxp.., load_files,
!     outpost of KRY / mode files, arnoldi_checkpoint): the Fortran members vx,vy,vz,pr stay the I/O view of the vector.
      subroutine krylov_to_device(p)
      use nekstab_b200_c
      use krylov_subspace
      implicit none
      type(krylov_vector) :: p
      call nsb_b200_check(nsb_vec_upload(p%slot, p%vx, p%vy, p%vz, p%pr), 'krylov_to_device')
      end subroutine

      subroutine krylov_to_host(p)
      use nekstab_b200_c
      use krylov_subspace
      implicit none
      type(krylov_vector) :: p
      call nsb_b200_check(nsb_vec_download(p%slot, p%vx, p%vy, p%vz, p%pr), 'krylov_to_host')
      end subroutine
