!     nekstab_b200_shim.f -- replacement bodies for the hot-path routines of nekStab that route to the CUDA library.
!
!     FIXED-FORM source like the reference's core/*.f (it includes Nek5000's fixed-form SIZE and TOTAL); compile with the
!     reference's own flags (core/compiler.sh: -fdefault-real-8 / -r8, -ffixed-line-length-none, -mcmodel=large).
!     Replaces, with identical names and argument lists, the routines of core/krylov_subspace.f (:24-258),
!     core/matvec.f (:64-159), update_hessenberg_matrix (core/krylov_decomposition.f:116-202) and nonlinear_forward_map
!     (core/newton_krylov.f:336-378).  Everything above them (arnoldi_factorization, krylov_schur, schur_condensation,
!     ts_gmres, newton_krylov, the LAPACK wrappers, outpost_ks) stays the reference's own Fortran.  type(krylov_vector)
!     gains one integer member, slot, the handle of the device-resident copy; its host arrays become a lazily
!     synchronised mirror (krylov_to_host / krylov_to_device at the field-access sites listed in SURVEY.md 8b).
!     NOT compiled in this repository's build image (no Fortran compiler); tests/test_fortran_bindings.py checks the
!     source form, that every nsb_* call is bound by nekstab_b200_c.f90 and that the bindings match the C header.

      subroutine nsb_b200_check(ierr, where)
      use nekstab_b200_c
      implicit none
      include 'SIZE'
      include 'TOTAL'
      integer ierr
      character(len=*) where
      if (ierr .ne. 0) then
         if (nid .eq. 0) write(6,*) 'nekstab_b200 error in ', where, ': ', nsb_error_message()
         call nek_end
      endif
      end subroutine nsb_b200_check

!     Called once from nekStab_init (core/usr_extra.f:72) after bm1s and the sponge are set.
      subroutine nsb_b200_setup
      use nekstab_b200_c
      use krylov_subspace
      implicit none
      include 'SIZE'
      include 'TOTAL'
      character(kind=c_char) :: id(128)
      integer(c_long_long) :: glo(lx1*ly1*lz1*lelv), ngv
!     Nek5000's element-vertex table (filled by get_vert in setup_topo) [UPSTREAM connect1.f / navier8.f]
      integer vertex
      common /ivrtx/ vertex((2**ldim)*lelt)
      integer ierr, ldev
      real nsb_ri
      common /nsb_buoy/ nsb_ri
      ldev = mod(nid, 8)
!     global GLL node numbers of the velocity mesh, exactly what setupds hands to gs_setup: Nek5000's own set_vert
!     (navier8.f -> setvert2d / setvert3d).  No Nek5000 patch is needed; INTEGRATION.md 1.3 shows the 5-line
!     alternative (copy glo_num out of setupds) for a fork whose set_vert differs.
      call set_vert(glo, ngv, lx1, nelv, vertex, .false.)
      if (np .gt. 1) then
         if (nid .eq. 0) ierr = nsb_comm_unique_id(id)
         call bcast(id, 128)
         call nsb_b200_check(nsb_comm_init(nid, np, id, ldev), 'nsb_comm_init')
      endif
      call nsb_b200_check(nsb_init(ldim, lx1, lxd, lx2, nelv, int(nelgv, c_long_long), xm1, ym1, zm1,
     $     v1mask, v2mask, v3mask, glo, ldev), 'nsb_init')
      call nsb_b200_check(nsb_set_params(param(2), param(1), param(22), param(21), 0, 0), 'nsb_set_params')
      call nsb_b200_check(nsb_set_weights(bm1s), 'nsb_set_weights')
      if (spng_str .ne. 0) call nsb_b200_check(nsb_set_sponge(spng_fun), 'nsb_set_sponge')
      call nsb_b200_check(nsb_set_ifvcor(merge(1, 0, ifvcor), -1), 'nsb_set_ifvcor')
!     [PRESSURE] preconditioner = semg_xxt in every shipped .par (1cyl.par:28): the multilevel Schwarz class (kind 1);
!     kind 0 keeps Jacobi.  nagg = 0: automatic number of coarse aggregates.
      call nsb_b200_check(nsb_set_pressure_preconditioner(1, 0), 'nsb_set_pressure_preconditioner')
!     Scalar transport (ifheat, ldimt = 1): theta joins every device vector; conductivity = param(8), rhocp = param(7), Nek5000's
!     tmask; the buoyancy coefficient is the case's own (userf: ffy = temp * uparam(6) in examples/cylinder/baseflow/newton_dyn_temp,
!     temp * Pr * Ra in examples/thersyphon) -- set nsb_ri accordingly before this call.  Must precede nsb_vec_alloc (layout change).
      if (ifheat) call nsb_b200_check(nsb_set_scalar(1, param(8), param(7), tmask(1,1,1,1,1), nsb_ri, 1), 'nsb_set_scalar')
!     Newton-GMRES for UPOs: time component in the inner product, orbit storage and border vectors on the device
      if (uparam(1) .eq. 2.1) call nsb_b200_check(nsb_set_upo(1), 'nsb_set_upo')
      call nsb_b200_check(nsb_vec_alloc(k_dim + 8), 'nsb_vec_alloc')
      end subroutine nsb_b200_setup

      subroutine krylov_inner_product(alpha, p, q)
!     core/krylov_subspace.f:24 ; uparam(1) = 2.1 (UPO): after nsb_set_upo(1) (nsb_b200_setup) the library adds the
!     time term of :48-50 itself -- every slot carries q%time, kept equal to the host member by the routines below
      use nekstab_b200_c
      use krylov_subspace
      implicit none
      include 'SIZE'
      include 'TOTAL'
      type(krylov_vector), intent(in) :: p, q
      real, intent(out) :: alpha
      call nsb_b200_check(nsb_vec_inner_product(p%slot, q%slot, alpha), 'krylov_inner_product')
      end subroutine krylov_inner_product

      subroutine krylov_norm(alpha, p)
!     core/krylov_subspace.f:58
      use nekstab_b200_c
      use krylov_subspace
      implicit none
      include 'SIZE'
      include 'TOTAL'
      type(krylov_vector), intent(in) :: p
      real, intent(out) :: alpha
      call nsb_b200_check(nsb_vec_norm(p%slot, alpha), 'krylov_norm')
      end subroutine krylov_norm

      subroutine krylov_normalize(p, alpha)
!     core/krylov_subspace.f:71
      use nekstab_b200_c
      use krylov_subspace
      implicit none
      type(krylov_vector), intent(inout) :: p
      real, intent(out) :: alpha
      call krylov_norm(alpha, p)
      call krylov_cmult(p, 1.0d0 / alpha)
      end subroutine krylov_normalize

      subroutine krylov_cmult(p, alpha)
!     core/krylov_subspace.f:90
      use nekstab_b200_c
      use krylov_subspace
      implicit none
      type(krylov_vector) :: p
      real alpha
      call nsb_b200_check(nsb_vec_cmult(p%slot, alpha), 'krylov_cmult')
      p%time = p%time * alpha
      end subroutine krylov_cmult

      subroutine krylov_add2(p, q)
!     core/krylov_subspace.f:116
      use nekstab_b200_c
      use krylov_subspace
      implicit none
      type(krylov_vector) :: p, q
      call nsb_b200_check(nsb_vec_add2(p%slot, q%slot), 'krylov_add2')
      p%time = p%time + q%time
      end subroutine krylov_add2

      subroutine krylov_sub2(p, q)
!     core/krylov_subspace.f:142
      use nekstab_b200_c
      use krylov_subspace
      implicit none
      type(krylov_vector) :: p, q
      call nsb_b200_check(nsb_vec_sub2(p%slot, q%slot), 'krylov_sub2')
      p%time = p%time - q%time
      end subroutine krylov_sub2

      subroutine krylov_zero(p)
!     core/krylov_subspace.f:166
      use nekstab_b200_c
      use krylov_subspace
      implicit none
      type(krylov_vector) :: p
      call nsb_b200_check(nsb_vec_zero(p%slot), 'krylov_zero')
      p%time = 0.0d0
      end subroutine krylov_zero

      subroutine krylov_copy(p, q)
!     core/krylov_subspace.f:190
      use nekstab_b200_c
      use krylov_subspace
      implicit none
      type(krylov_vector) :: p, q
      call nsb_b200_check(nsb_vec_copy(p%slot, q%slot), 'krylov_copy')
      p%time = q%time
      end subroutine krylov_copy

      subroutine krylov_matmul(dq, Q, yvec, k)
!     core/krylov_subspace.f:214 ; Q(1:k) occupy consecutive device slots
      use nekstab_b200_c
      use krylov_subspace
      implicit none
      integer :: k
      type(krylov_vector) :: dq
      type(krylov_vector), dimension(k) :: Q
      real, dimension(k) :: yvec
      call nsb_b200_check(nsb_basis_gemv(k, Q(1)%slot, yvec, dq%slot), 'krylov_matmul')
      dq%time = dot_product(Q(1:k)%time, yvec(1:k))
      end subroutine krylov_matmul

      subroutine update_hessenberg_matrix(H, f, q, k)
!     core/krylov_decomposition.f:116 ; the new column H(1:k+1,k) comes back from the device (CGS2/DGKS)
      use nekstab_b200_c
      use krylov_subspace
      implicit none
      integer, intent(in) :: k
      real, dimension(k+1, k), intent(inout) :: H
      type(krylov_vector), dimension(k) :: q
      type(krylov_vector) :: f
      real(c_double) :: tt
      call nsb_b200_check(nsb_orthonormalize(k, q(1)%slot, f%slot, H(1, k)), 'update_hessenberg_matrix')
      call nsb_b200_check(nsb_vec_get_time(f%slot, tt), 'update_hessenberg_matrix')
      f%time = tt
      end subroutine update_hessenberg_matrix

!     The per-step host hook: the reference calls nekstab_usrchk() before every nek_advance (core/matvec.f:221,304).
!     Registered once with nsb_set_step_callback(c_funloc(nsb_b200_step_hook), c_null_ptr) when the case's
!     nekstab_usrchk does more than set defaults at istep = 0.
      subroutine nsb_b200_step_hook(istep_c, time_c, user) bind(C)
      use iso_c_binding
      implicit none
      include 'SIZE'
      include 'TOTAL'
      integer(c_int), value :: istep_c
      real(c_double), value :: time_c
      type(c_ptr), value :: user
      istep = istep_c
      time = time_c
      call nekstab_usrchk
      end subroutine nsb_b200_step_hook

      subroutine matvec(f, q)
!     core/matvec.f:64-159
      use nekstab_b200_c
      use krylov_subspace
      implicit none
      include 'SIZE'
      include 'TOTAL'
      type(krylov_vector) :: q, f
      logical, save :: init = .false.
      integer(c_int) :: mode, nst
      real(c_double) :: ddt, ct
      if (.not. init) then
!        prepare_linearized_solver, core/matvec.f:1-52,115-118
         call nsb_b200_check(nsb_set_baseflow(ubase, vbase, wbase), 'nsb_set_baseflow')
         if (ifheat) call nsb_b200_check(nsb_set_scalar_base(tbase), 'nsb_set_scalar_base')
!        Floquet (uparam(1) = 3.11 / 3.21, core/matvec.f:192,278): ifbase co-evolution + orbit storage on the device;
!        the co-evolving base flow feels the DNS branch of nekStab_forcing (core/utils.f:166-171)
         if (uparam(1) .eq. 3.11 .or. uparam(1) .eq. 3.21) then
            if (spng_str .ne. 0) call nsb_b200_check(nsb_set_dns_sponge(spng_str, spng_vr(1,1), spng_vr(1,2),
     $           spng_vr(1,ndim)), 'nsb_set_dns_sponge')
            call nsb_b200_check(nsb_set_floquet(1, pr), 'nsb_set_floquet')
         endif
         call nsb_b200_check(nsb_prepare_linearized_solver(param(10), param(26), ddt, nst, ct),
     $        'prepare_linearized_solver')
         dt = ddt
         nsteps = nst
         ctarg = ct
         param(12) = -abs(dt)
         init = .true.
      endif
      mode = -1
      if (uparam(1) .ge. 3.0 .and. uparam(1) .lt. 3.2) then
         evop = 'd'
         mode = NSB_DIRECT
      endif
      if (uparam(1) .ge. 3.2 .and. uparam(1) .lt. 3.3) then
         evop = 'a'
         mode = NSB_ADJOINT
      endif
      if (uparam(1) .ge. 3.3 .and. uparam(1) .lt. 3.4) then
         evop = 'p'
         mode = NSB_DIRECT_ADJOINT
      endif
      if (floor(uparam(1)) .eq. 4) mode = NSB_FORCE_SENS
      if (floor(uparam(1)) .eq. 2) then
         evop = 'n'
         mode = NSB_NEWTON
         init = .false.
      endif
!     newton_linearized_map for UPOs (core/matvec.f:407-419): the border terms bvec(fc_nwt) * q%time and
!     f%time = <bvec(ic_nwt), q> are evaluated on the device (nsb_set_upo); f%time = 0 otherwise (:421)
      call nsb_b200_check(nsb_vec_set_time(q%slot, real(q%time, c_double)), 'matvec')
      call nsb_b200_check(nsb_matvec(mode, q%slot, f%slot), 'matvec')
      f%time = 0.0d0
      if (mode .eq. NSB_NEWTON) then
         call nsb_b200_check(nsb_vec_get_time(f%slot, ddt), 'matvec')
         f%time = ddt
      endif
      end subroutine matvec

      subroutine nonlinear_forward_map(f, q)
!     core/newton_krylov.f:336-378 ; newton_krylov re-prepares the solver on every iterate (:69): CFL of the current q
      use nekstab_b200_c
      use krylov_subspace
      implicit none
      include 'SIZE'
      include 'TOTAL'
      type(krylov_vector) :: f, q
      integer(c_int) :: nst
      real(c_double) :: ddt, ct
      call nsb_b200_check(nsb_prepare_solver_from_slot(q%slot, param(10), param(26), ddt, nst, ct),
     $     'prepare_linearized_solver')
      dt = ddt
      nsteps = nst
      ctarg = ct
      param(12) = -abs(dt)
!     f = phi_T(q) - q ; ubase <- q
      call nsb_b200_check(nsb_nonlinear_forward_map(q%slot, f%slot), 'nonlinear_forward_map')
      f%time = 0.0d0
      end subroutine nonlinear_forward_map

!     Host mirror <-> device slot, to be called at the field-access sites listed in SURVEY.md 8b (seeding from vxp..,
!     load_files, outpost of KRY / mode files, arnoldi_checkpoint): vx,vy,vz,pr stay the I/O view of the vector.
      subroutine krylov_to_device(p)
      use nekstab_b200_c
      use krylov_subspace
      implicit none
      include 'SIZE'
      include 'TOTAL'
      type(krylov_vector) :: p
      call nsb_b200_check(nsb_vec_upload(p%slot, p%vx, p%vy, p%vz, p%pr), 'krylov_to_device')
      if (ifheat) call nsb_b200_check(nsb_vec_upload_scalar(p%slot, p%theta(1,1)), 'krylov_to_device')
      call nsb_b200_check(nsb_vec_set_time(p%slot, real(p%time, c_double)), 'krylov_to_device')
      end subroutine krylov_to_device

      subroutine krylov_to_host(p)
      use nekstab_b200_c
      use krylov_subspace
      implicit none
      include 'SIZE'
      include 'TOTAL'
      type(krylov_vector) :: p
      real(c_double) :: tt
      call nsb_b200_check(nsb_vec_download(p%slot, p%vx, p%vy, p%vz, p%pr), 'krylov_to_host')
      if (ifheat) call nsb_b200_check(nsb_vec_download_scalar(p%slot, p%theta(1,1)), 'krylov_to_host')
      call nsb_b200_check(nsb_vec_get_time(p%slot, tt), 'krylov_to_host')
      p%time = tt
      end subroutine krylov_to_host
