!> nekstab_b200_c.f90 -- ISO_C_BINDING interfaces to libnekstab_b200.so (include/nekstab_b200.h).
!!
!! Drop-in layer for nekStab: compile this file and nekstab_b200_shim.f90 with the case (add both objects to the USR
!! list of core/nekStab.sh:13-25) and link -lnekstab_b200.  NOT compiled in the build container of this repository
!! (no Fortran compiler there); the same C entry points are exercised through ctypes by tests/.
      module nekstab_b200_c
      use iso_c_binding
      implicit none

      integer(c_int), parameter :: NSB_DIRECT = 1, NSB_ADJOINT = 2, NSB_DIRECT_ADJOINT = 3
      integer(c_int), parameter :: NSB_NEWTON = 4, NSB_FORCE_SENS = 5

      interface
      integer(c_int) function nsb_comm_unique_id(id) bind(C, name='nsb_comm_unique_id')
        import :: c_int, c_char
        character(kind=c_char) :: id(128)
      end function
      integer(c_int) function nsb_comm_init(rank, nranks, id, device) bind(C, name='nsb_comm_init')
        import :: c_int, c_char
        integer(c_int), value :: rank, nranks, device
        character(kind=c_char) :: id(128)
      end function
      integer(c_int) function nsb_init(ldim, lx1, lxd, lx2, nelv, nelgv, xm1, ym1, zm1, v1mask, v2mask, v3mask, &
                                       glo_num, device) bind(C, name='nsb_init')
        import :: c_int, c_long_long, c_double
        integer(c_int), value :: ldim, lx1, lxd, lx2, nelv, device
        integer(c_long_long), value :: nelgv
        real(c_double) :: xm1(*), ym1(*), zm1(*), v1mask(*), v2mask(*), v3mask(*)
        integer(c_long_long) :: glo_num(*)
      end function
      integer(c_int) function nsb_finalize() bind(C, name='nsb_finalize')
        import :: c_int
      end function
      integer(c_int) function nsb_set_params(visc, dens, tolv, tolp, maxv, maxp) bind(C, name='nsb_set_params')
        import :: c_int, c_double
        real(c_double), value :: visc, dens, tolv, tolp
        integer(c_int), value :: maxv, maxp
      end function
      integer(c_int) function nsb_set_weights(bm1s) bind(C, name='nsb_set_weights')
        import :: c_int, c_double
        real(c_double) :: bm1s(*)
      end function
      integer(c_int) function nsb_set_baseflow(u, v, w) bind(C, name='nsb_set_baseflow')
        import :: c_int, c_double
        real(c_double) :: u(*), v(*), w(*)
      end function
      integer(c_int) function nsb_set_sponge(spng) bind(C, name='nsb_set_sponge')
        import :: c_int, c_double
        real(c_double) :: spng(*)
      end function
      integer(c_int) function nsb_set_ifvcor(direct, adjoint) bind(C, name='nsb_set_ifvcor')
        import :: c_int
        integer(c_int), value :: direct, adjoint
      end function
      integer(c_int) function nsb_set_pressure_preconditioner(kind, nagg) bind(C, name='nsb_set_pressure_preconditioner')
         import :: c_int
         integer(c_int), value :: kind, nagg
      end function
      integer(c_int) function nsb_set_projection(mxprev) bind(C, name='nsb_set_projection')
         import :: c_int
         integer(c_int), value :: mxprev
      end function
      integer(c_int) function nsb_set_adjoint_masks(m1, m2, m3) bind(C, name='nsb_set_adjoint_masks')
        import :: c_int, c_double
        real(c_double) :: m1(*), m2(*), m3(*)
      end function
      integer(c_int) function nsb_prepare_linearized_solver(endt, cfl, dt, nsteps, ctarg) &
          bind(C, name='nsb_prepare_linearized_solver')
        import :: c_int, c_double
        real(c_double), value :: endt, cfl
        real(c_double) :: dt, ctarg
        integer(c_int) :: nsteps
      end function
      integer(c_int) function nsb_vec_alloc(nslots) bind(C, name='nsb_vec_alloc')
        import :: c_int
        integer(c_int), value :: nslots
      end function
      integer(c_int) function nsb_vec_upload(slot, vx, vy, vz, pr) bind(C, name='nsb_vec_upload')
        import :: c_int, c_double
        integer(c_int), value :: slot
        real(c_double) :: vx(*), vy(*), vz(*), pr(*)
      end function
      integer(c_int) function nsb_vec_download(slot, vx, vy, vz, pr) bind(C, name='nsb_vec_download')
        import :: c_int, c_double
        integer(c_int), value :: slot
        real(c_double) :: vx(*), vy(*), vz(*), pr(*)
      end function
      integer(c_int) function nsb_vec_copy(dst, src) bind(C, name='nsb_vec_copy')
        import :: c_int
        integer(c_int), value :: dst, src
      end function
      integer(c_int) function nsb_vec_zero(slot) bind(C, name='nsb_vec_zero')
        import :: c_int
        integer(c_int), value :: slot
      end function
      integer(c_int) function nsb_vec_cmult(slot, alpha) bind(C, name='nsb_vec_cmult')
        import :: c_int, c_double
        integer(c_int), value :: slot
        real(c_double), value :: alpha
      end function
      integer(c_int) function nsb_vec_add2(p, q) bind(C, name='nsb_vec_add2')
        import :: c_int
        integer(c_int), value :: p, q
      end function
      integer(c_int) function nsb_vec_sub2(p, q) bind(C, name='nsb_vec_sub2')
        import :: c_int
        integer(c_int), value :: p, q
      end function
      integer(c_int) function nsb_vec_inner_product(p, q, alpha) bind(C, name='nsb_vec_inner_product')
        import :: c_int, c_double
        integer(c_int), value :: p, q
        real(c_double) :: alpha
      end function
      integer(c_int) function nsb_vec_normalize(p, alpha) bind(C, name='nsb_vec_normalize')
        import :: c_int, c_double
        integer(c_int), value :: p
        real(c_double) :: alpha
      end function
      integer(c_int) function nsb_basis_gemv(k, first, y, slot_out) bind(C, name='nsb_basis_gemv')
        import :: c_int, c_double
        integer(c_int), value :: k, first, slot_out
        real(c_double) :: y(*)
      end function
      integer(c_int) function nsb_basis_rotate(k, first, S, lds) bind(C, name='nsb_basis_rotate')
        import :: c_int, c_double
        integer(c_int), value :: k, first, lds
        real(c_double) :: S(lds, *)
      end function
      integer(c_int) function nsb_orthonormalize(k, first, slot_f, hcol) bind(C, name='nsb_orthonormalize')
        import :: c_int, c_double
        integer(c_int), value :: k, first, slot_f
        real(c_double) :: hcol(*)
      end function
      integer(c_int) function nsb_matvec(mode, slot_in, slot_out) bind(C, name='nsb_matvec')
        import :: c_int
        integer(c_int), value :: mode, slot_in, slot_out
      end function
      integer(c_int) function nsb_vec_norm(p, alpha) bind(C, name='nsb_vec_norm')
        import :: c_int, c_double
        integer(c_int), value :: p
        real(c_double) :: alpha
      end function
      integer(c_int) function nsb_nonlinear_forward_map(slot_q, slot_f) bind(C, name='nsb_nonlinear_forward_map')
        import :: c_int
        integer(c_int), value :: slot_q, slot_f
      end function
      integer(c_int) function nsb_prepare_solver_from_slot(slot, endt, cfl, dt, nsteps, ctarg) &
          bind(C, name='nsb_prepare_solver_from_slot')
        import :: c_int, c_double
        integer(c_int), value :: slot
        real(c_double), value :: endt, cfl
        real(c_double) :: dt, ctarg
        integer(c_int) :: nsteps
      end function
      integer(c_int) function nsb_basis_gemv_complex(k, first, yre, yim, slot_re, slot_im) bind(C, name='nsb_basis_gemv_complex')
        import :: c_int, c_double
        integer(c_int), value :: k, first, slot_re, slot_im
        real(c_double) :: yre(*), yim(*)
      end function
      end interface
      end module nekstab_b200_c
