module diva_gpu_bindings

  ! ISO_C_BINDING interfaces to libufe_diva.so (include/ufe_diva.h): the B200-native replacement of
  ! the DIVA / SSA velocity solve.  Intended location in the reference tree:
  !   src/UFEMISM/ice_dynamics/conservation_of_momentum/SSA_DIVA/diva_gpu_bindings.f90
  ! Call sites it serves (INTEGRATION.md):
  !   L0  solve_linearised_SSA_DIVA.f90:159      call solve_matrix_equation_CSR_PETSc( ...)
  !   L1  DIVA_main.f90:189-192, SSA_main.f90:178 call solve_SSA_DIVA_linearised( ...)
  !   L2  conservation_of_momentum_main.f90:142   call solve_DIVA( ...)
  ! NOTE: the build image of this repository has no Fortran compiler; this file is compile-checked
  ! only where gfortran exists (gfortran -c diva_gpu_bindings.f90).

  use, intrinsic :: iso_c_binding

  implicit none

  integer(c_int), parameter :: UFE_OK = 0
  integer(c_int), parameter :: UFE_FLAG_PICARD_MAXIT = 1, UFE_FLAG_KRYLOV_MAXIT = 2, UFE_FLAG_KRYLOV_DIVERGED = 4
  integer(c_int), parameter :: UFE_BC_INFINITE = 1, UFE_BC_ZERO = 2, UFE_BC_PERIODIC_ISMIP_HOM = 3, UFE_BC_INFINITE_SSA_ICESTREAM = 4
  integer(c_int), parameter :: UFE_KRYLOV_BICGSTAB = 0, UFE_KRYLOV_GMRES = 1
  integer(c_int), parameter :: UFE_PC_JACOBI = 0, UFE_PC_BJACOBI2 = 1, UFE_PC_BJACOBI_LU = 2, UFE_PC_AUTO = 3

  type, bind(C) :: ufe_csr
    integer(c_int32_t) :: m, n, m_loc, n_loc, i1, i2, j1, j2, nnz
    type(c_ptr)        :: ptr, ind, val
  end type ufe_csr

  type, bind(C) :: ufe_mesh
    integer(c_int32_t) :: nV, nTri, nC_mem, nz
    real(c_double)     :: xmin, xmax, ymin, ymax
    type(c_ptr)        :: V, Tri, TriC, C, nC, iTri, niTri, VBI, TriBI, TriGC, zeta
    type(c_ptr)        :: M_a_b(3), M_b_a(3), M2_b_b(5)
  end type ufe_mesh

  type, bind(C) :: ufe_config
    integer(c_int32_t) :: do_include_SSADIVA_crossterms
    real(c_double)     :: visc_it_norm_dUV_tol
    integer(c_int32_t) :: visc_it_nit
    real(c_double)     :: visc_it_relax, visc_eff_min, vel_max, stress_balance_PETSc_rtol, stress_balance_PETSc_abstol
    integer(c_int32_t) :: BC_u(4), BC_v(4)
    integer(c_int32_t) :: choice_sliding_law, choice_idealised_sliding_law
    real(c_double)     :: slid_Weertman_m, slid_Budd_q_plastic, slid_Budd_u_threshold, slid_ZI_p, slid_ZI_ut
    integer(c_int32_t) :: do_GL_subgrid_friction, do_subgrid_friction_on_A_grid
    real(c_double)     :: subgrid_friction_exponent_on_B_grid, slid_beta_max, slid_delta_v, Hi_min
    real(c_double)     :: Glens_flow_law_exponent, Glens_flow_law_epsilon_sq_0
    integer(c_int32_t) :: choice_ice_rheology_Glen
    real(c_double)     :: uniform_Glens_flow_factor
    integer(c_int32_t) :: choice_enhancement_factor_transition
    real(c_double)     :: m_enh_sheet, m_enh_shelf
    real(c_double)     :: refgeo_idealised_SSA_icestream_Hi, refgeo_idealised_SSA_icestream_dhdx
    real(c_double)     :: refgeo_idealised_SSA_icestream_L, refgeo_idealised_SSA_icestream_m
    real(c_double)     :: refgeo_idealised_ISMIP_HOM_L
    integer(c_int32_t) :: krylov_method, krylov_pc, krylov_maxits, krylov_guess_nonzero, krylov_pc_lag, krylov_pc_strip_only
  end type ufe_config

  type, bind(C) :: ufe_ice_inputs
    type(c_ptr) :: Hi, Hs, Hib, SL, fraction_gr, fraction_gr_b, effective_pressure
    type(c_ptr) :: mask_grounded_ice, mask_floating_ice, mask_icefree_land
    type(c_ptr) :: Ti, till_friction_angle, alpha_sq, beta_sq
    type(c_ptr) :: BC_prescr_mask_b, BC_prescr_u_b, BC_prescr_v_b
  end type ufe_ice_inputs

  type, bind(C) :: ufe_diva_state
    type(c_ptr) :: u_vav_b, v_vav_b, tau_bx_b, tau_by_b, eta_3D_b, u_base_b, v_base_b
    type(c_ptr) :: u_3D_b, v_3D_b, du_dx_a, du_dy_a, dv_dx_a, dv_dy_a, du_dz_3D_a, dv_dz_3D_a, eta_3D_a
    type(c_ptr) :: basal_friction_coefficient_a
  end type ufe_diva_state

  type, bind(C) :: ufe_ssa_state
    type(c_ptr) :: u_b, v_b, basal_friction_coefficient_a
  end type ufe_ssa_state

  type, bind(C) :: ufe_solve_info
    integer(c_int32_t) :: n_visc_its, n_Axb_its, flags
    real(c_double)     :: L2_uv, visc_it_relax_applied, Glens_flow_law_epsilon_sq_0_applied
    real(c_double)     :: ms_total, ms_closures, ms_assembly, ms_krylov, ms_h2d, ms_d2h
    integer(c_int64_t) :: gpu_launches
    integer(c_int32_t) :: krylov_pc_used, reserved
  end type ufe_solve_info

  type, bind(C) :: ufe_comm
    integer(c_int32_t) :: rank, nranks, device
    type(c_ptr)        :: nccl_unique_id
  end type ufe_comm

  ! ---- ice-thickness path (SURVEY.md 8f rank 2) -------------------------------------------------
  ! type_mesh members read by calc_ice_flux_divergence_matrix_upwind / map_velocities_from_b_to_c_2D
  type, bind(C) :: ufe_mesh_edges
    integer(c_int32_t) :: nE
    type(c_ptr)        :: VE, ETri, A, Cw, D_x, D_y, D      ! c_loc( mesh%VE), c_loc( mesh%ETri), ...
  end type ufe_mesh_edges

  type, bind(C) :: ufe_thickness_config
    real(c_double)     :: dHi_semiimplicit_fs, dHi_PETSc_rtol, dHi_PETSc_abstol
    integer(c_int32_t) :: BC_H(4)                           ! north, east, south, west: 1 'infinite', 2 'zero'
    real(c_double)     :: dt_ice_max, dt_ice_min, Hi_min
    integer(c_int32_t) :: krylov_method, krylov_maxits
  end type ufe_thickness_config

  type, bind(C) :: ufe_thickness_fields
    type(c_ptr) :: Hi, Hb, SL, u_vav_b, v_vav_b, SMB, BMB, LMB, fraction_margin, dHi_dt_target
    type(c_ptr) :: mask_noice, BC_prescr_mask, BC_prescr_Hi
    type(c_ptr) :: AMB, dHi_dt, Hi_tplusdt, divQ
  end type ufe_thickness_fields

  ! dummy arguments / type_ice_model members read by calc_vertical_velocities (vertical_velocities.f90:18-210)
  type, bind(C) :: ufe_vertical_velocity_inputs
    type(c_ptr) :: Hi, Hib, dHb_dt, dHi_dt, BMB, mask_grounded_ice, mask_floating_ice
    type(c_ptr) :: dzeta_dx_ak, dzeta_dy_ak, dzeta_dz_ak
  end type ufe_vertical_velocity_inputs

  interface

    function ufe_last_error_string() bind(C, name='ufe_last_error_string') result(s)
      import :: c_ptr
      type(c_ptr) :: s
    end function ufe_last_error_string

    integer(c_int) function ufe_comm_get_unique_id( id_out) bind(C, name='ufe_comm_get_unique_id')
      import :: c_int, c_char
      character(kind=c_char), intent(out) :: id_out(128)
    end function ufe_comm_get_unique_id

    subroutine ufe_partition_list( ntot, i, n, i1, i2) bind(C, name='ufe_partition_list')
      import :: c_int32_t
      integer(c_int32_t), value       :: ntot, i, n
      integer(c_int32_t), intent(out) :: i1, i2
    end subroutine ufe_partition_list

    ! L0: replaces solve_matrix_equation_CSR_PETSc (src/UPSY/basic/petsc_basic.f90:32-64)
    integer(c_int) function ufe_krylov_solve( A, b, x, rtol, abstol, method, maxits, guess_nonzero, n_its, flags) &
        bind(C, name='ufe_krylov_solve')
      import :: c_int, c_int32_t, c_double, ufe_csr
      type(ufe_csr),      intent(in)    :: A
      real(c_double),     intent(in)    :: b(*)
      real(c_double),     intent(inout) :: x(*)
      real(c_double),     value         :: rtol, abstol
      integer(c_int32_t), value         :: method, maxits, guess_nonzero
      integer(c_int32_t), intent(out)   :: n_its, flags
    end function ufe_krylov_solve

    ! L0 as the reference calls it: every process passes ITS rows of A_CSR (i1..i2) and its slices of bb / xx; collective
    ! over the ranks of the handle; method / preconditioner / maxits from the handle's ufe_config
    integer(c_int) function ufe_solve_matrix_equation_CSR( handle, A, bb, xx, rtol, abstol, n_Axb_its, flags) &
        bind(C, name='ufe_solve_matrix_equation_CSR')
      import :: c_int, c_int32_t, c_double, c_ptr, ufe_csr
      type(c_ptr),        value         :: handle
      type(ufe_csr),      intent(in)    :: A
      real(c_double),     intent(in)    :: bb(*)
      real(c_double),     intent(inout) :: xx(*)
      real(c_double),     value         :: rtol, abstol
      integer(c_int32_t), intent(out)   :: n_Axb_its, flags
    end function ufe_solve_matrix_equation_CSR

    integer(c_int) function ufe_last_l0_preconditioner( handle) bind(C, name='ufe_last_l0_preconditioner')
      import :: c_int, c_ptr
      type(c_ptr), value :: handle
    end function ufe_last_l0_preconditioner

    ! ABI check: c_sizeof( info) of type(ufe_solve_info) must equal this (96)
    integer(c_int) function ufe_sizeof_solve_info() bind(C, name='ufe_sizeof_solve_info')
      import :: c_int
    end function ufe_sizeof_solve_info

    ! exact multifrontal nested-dissection solver for the b-grid (u,v) systems (csrc/ufe_nd.cu, ufe_nd_numeric.cu):
    ! analyse + create once per mesh, factor + solve per Picard iteration; also reachable as cfg%krylov_pc = 4 (nd_lu)
    integer(c_int) function ufe_nd_analyse( nT, gcx, gcy, bptr, bind_, leaf_triangles, tree) bind(C, name='ufe_nd_analyse')
      import :: c_int, c_int32_t, c_double, c_ptr
      integer(c_int32_t), value       :: nT, leaf_triangles
      real(c_double),     intent(in)  :: gcx(*), gcy(*)          ! mesh%TriGC(:,1), mesh%TriGC(:,2)
      integer(c_int32_t), intent(in)  :: bptr(*), bind_(*)       ! 0-based block pattern over triangles
      type(c_ptr),        intent(out) :: tree
    end function ufe_nd_analyse

    subroutine ufe_nd_tree_free( tree) bind(C, name='ufe_nd_tree_free')
      import :: c_ptr
      type(c_ptr), value :: tree
    end subroutine ufe_nd_tree_free

    integer(c_int) function ufe_nd_solver_create( tree, N, ptr, ind, solver) bind(C, name='ufe_nd_solver_create')
      import :: c_int, c_int32_t, c_ptr
      type(c_ptr),        value       :: tree
      integer(c_int32_t), value       :: N
      integer(c_int32_t), intent(in)  :: ptr(*), ind(*)          ! A_CSR%ptr - 1, A_CSR%ind - 1
      type(c_ptr),        intent(out) :: solver
    end function ufe_nd_solver_create

    integer(c_int) function ufe_nd_solver_factor( solver, val) bind(C, name='ufe_nd_solver_factor')
      import :: c_int, c_double, c_ptr
      type(c_ptr),    value      :: solver
      real(c_double), intent(in) :: val(*)                       ! A_CSR%val
    end function ufe_nd_solver_factor

    integer(c_int) function ufe_nd_solver_solve( solver, b, x, n_refine, relres) bind(C, name='ufe_nd_solver_solve')
      import :: c_int, c_int32_t, c_double, c_ptr
      type(c_ptr),        value       :: solver
      real(c_double),     intent(in)  :: b(*)
      real(c_double),     intent(out) :: x(*)
      integer(c_int32_t), value       :: n_refine
      real(c_double),     intent(out) :: relres
    end function ufe_nd_solver_solve

    subroutine ufe_nd_solver_free( solver) bind(C, name='ufe_nd_solver_free')
      import :: c_ptr
      type(c_ptr), value :: solver
    end subroutine ufe_nd_solver_free

    ! multiply_CSR_matrix_with_vector_1D / _2D (CSR_matrix_vector_multiplication.f90:198,336)
    integer(c_int) function ufe_spmv( A, x, y, nlayers) bind(C, name='ufe_spmv')
      import :: c_int, c_int32_t, c_double, ufe_csr
      type(ufe_csr),      intent(in)  :: A
      real(c_double),     intent(in)  :: x(*)
      real(c_double),     intent(out) :: y(*)
      integer(c_int32_t), value       :: nlayers
    end function ufe_spmv

    ! lifecycle: initialise_DIVA_solver / allocate_DIVA_solver (DIVA_main.f90:37,752)
    integer(c_int) function ufe_diva_create( mesh, cfg, comm, handle) bind(C, name='ufe_diva_create')
      import :: c_int, c_ptr, ufe_mesh, ufe_config
      type(ufe_mesh),   intent(in)  :: mesh
      type(ufe_config), intent(in)  :: cfg
      type(c_ptr),      value       :: comm        ! c_loc( ufe_comm) or c_null_ptr (one rank)
      type(c_ptr),      intent(out) :: handle
    end function ufe_diva_create

    integer(c_int) function ufe_diva_destroy( handle) bind(C, name='ufe_diva_destroy')
      import :: c_int, c_ptr
      type(c_ptr), value :: handle
    end function ufe_diva_destroy

    integer(c_int) function ufe_diva_set_config( handle, cfg) bind(C, name='ufe_diva_set_config')
      import :: c_int, c_ptr, ufe_config
      type(c_ptr),      value      :: handle
      type(ufe_config), intent(in) :: cfg
    end function ufe_diva_set_config

    ! L2: replaces solve_DIVA (DIVA_main.f90:88-262) and solve_SSA (SSA_main.f90:87-242)
    integer(c_int) function ufe_diva_solve( handle, ice, state, info) bind(C, name='ufe_diva_solve')
      import :: c_int, c_ptr, ufe_ice_inputs, ufe_diva_state, ufe_solve_info
      type(c_ptr),          value         :: handle
      type(ufe_ice_inputs), intent(in)    :: ice
      type(ufe_diva_state), intent(inout) :: state
      type(ufe_solve_info), intent(out)   :: info
    end function ufe_diva_solve

    integer(c_int) function ufe_ssa_solve( handle, ice, state, info) bind(C, name='ufe_ssa_solve')
      import :: c_int, c_ptr, ufe_ice_inputs, ufe_ssa_state, ufe_solve_info
      type(c_ptr),          value         :: handle
      type(ufe_ice_inputs), intent(in)    :: ice
      type(ufe_ssa_state),  intent(inout) :: state
      type(ufe_solve_info), intent(out)   :: info
    end function ufe_ssa_solve

    ! L1: replaces solve_SSA_DIVA_linearised (solve_linearised_SSA_DIVA.f90:23-178)
    integer(c_int) function ufe_ssa_diva_linearised( handle, u_b, v_b, N_b, dN_dx_b, dN_dy_b, basal_friction_coefficient_b, &
        tau_dx_b, tau_dy_b, u_b_prev, v_b_prev, PETSc_rtol, PETSc_abstol, n_Axb_its, BC_prescr_mask_b, BC_prescr_u_b, BC_prescr_v_b) &
        bind(C, name='ufe_ssa_diva_linearised')
      import :: c_int, c_int32_t, c_double, c_ptr
      type(c_ptr),        value         :: handle
      real(c_double),     intent(inout) :: u_b(*), v_b(*)
      real(c_double),     intent(in)    :: N_b(*), dN_dx_b(*), dN_dy_b(*), basal_friction_coefficient_b(*), tau_dx_b(*), tau_dy_b(*)
      real(c_double),     intent(out)   :: u_b_prev(*), v_b_prev(*)
      real(c_double),     value         :: PETSc_rtol, PETSc_abstol
      integer(c_int32_t), intent(out)   :: n_Axb_its
      type(c_ptr),        value         :: BC_prescr_mask_b, BC_prescr_u_b, BC_prescr_v_b   ! c_null_ptr when absent
    end function ufe_ssa_diva_linearised

    ! upload mesh%VE, ETri, A, Cw, D_x, D_y, D once per mesh
    integer(c_int) function ufe_mesh_set_edges( handle, edges) bind(C, name='ufe_mesh_set_edges')
      import :: c_int, c_ptr, ufe_mesh_edges
      type(c_ptr),          value      :: handle
      type(ufe_mesh_edges), intent(in) :: edges
    end function ufe_mesh_set_edges

    ! replaces calc_dHi_dt (conservation_of_mass_main.f90:22-109); method: 0 'none', 1 'explicit', 2 'semi-implicit'
    integer(c_int) function ufe_calc_dHi_dt( handle, cfg, method, fields, dt, n_Axb_its, flags) bind(C, name='ufe_calc_dHi_dt')
      import :: c_int, c_int32_t, c_ptr, c_double, ufe_thickness_config, ufe_thickness_fields
      type(c_ptr),                value         :: handle
      type(ufe_thickness_config), intent(in)    :: cfg
      integer(c_int32_t),         value         :: method
      type(ufe_thickness_fields), intent(inout) :: fields
      real(c_double),             intent(inout) :: dt
      integer(c_int32_t),         intent(out)   :: n_Axb_its, flags
    end function ufe_calc_dHi_dt

    ! replaces calc_dHi_dt_explicit (conservation_of_mass_explicit.f90:23-138)
    integer(c_int) function ufe_calc_dHi_dt_explicit( handle, cfg, fields, dt) bind(C, name='ufe_calc_dHi_dt_explicit')
      import :: c_int, c_ptr, c_double, ufe_thickness_config, ufe_thickness_fields
      type(c_ptr),                value         :: handle
      type(ufe_thickness_config), intent(in)    :: cfg
      type(ufe_thickness_fields), intent(inout) :: fields
      real(c_double),             intent(inout) :: dt
    end function ufe_calc_dHi_dt_explicit

    ! replaces calc_dHi_dt_semiimplicit (conservation_of_mass_semiimplicit.f90:24-173)
    integer(c_int) function ufe_calc_dHi_dt_semiimplicit( handle, cfg, fields, dt, n_Axb_its, flags) &
        bind(C, name='ufe_calc_dHi_dt_semiimplicit')
      import :: c_int, c_int32_t, c_ptr, c_double, ufe_thickness_config, ufe_thickness_fields
      type(c_ptr),                value         :: handle
      type(ufe_thickness_config), intent(in)    :: cfg
      type(ufe_thickness_fields), intent(inout) :: fields
      real(c_double),             value         :: dt
      integer(c_int32_t),         intent(out)   :: n_Axb_its, flags
    end function ufe_calc_dHi_dt_semiimplicit

    ! replaces calc_vertical_velocities (vertical_velocities.f90:18-210); w_3D (nV,nz) out
    integer(c_int) function ufe_calc_vertical_velocities( handle, inputs, w_3D) bind(C, name='ufe_calc_vertical_velocities')
      import :: c_int, c_ptr, c_double, ufe_vertical_velocity_inputs
      type(c_ptr),                        value       :: handle
      type(ufe_vertical_velocity_inputs), intent(in)  :: inputs
      real(c_double),                     intent(out) :: w_3D(*)
    end function ufe_calc_vertical_velocities

  end interface

contains

  subroutine ufe_check( ierr, routine_name)
    ! Maps a non-zero status to the reference's crash() (control_resources_and_error_messaging.f90:377)
    integer(c_int),   intent(in) :: ierr
    character(len=*), intent(in) :: routine_name
    character(kind=c_char), pointer :: cmsg(:)
    character(len=1024) :: msg
    integer :: i
    if (ierr == UFE_OK) return
    call c_f_pointer( ufe_last_error_string(), cmsg, [1024])
    msg = ''
    do i = 1, 1024
      if (cmsg( i) == c_null_char) exit
      msg( i:i) = cmsg( i)
    end do
    ! call crash( trim( routine_name) // ': ' // trim( msg))
    error stop trim( routine_name) // ': ' // trim( msg)
  end subroutine ufe_check

end module diva_gpu_bindings
