! rxn_b200_shim.F90 — ISO_C_BINDING module that binds the C ABI of include/rxn_b200.h 1:1.
!
! This is the file a PFLOTRAN maintainer adds to src/pflotran/ (INTEGRATION.md shows the call-site
! patches in reactive_transport.F90).  It holds NO chemistry: it passes c_loc() of the reference's
! own compressed tables (reaction_type, reaction_aux.F90:142-335; mineral_type,
! reaction_mineral_aux.F90:77-128; surface_complexation_type, reaction_surf_complex_aux.F90:68-128)
! and of the PETSc Vec arrays to the batched entry points.  Species ids stay 1-based and arrays keep
! their Fortran memory order (see RxnSpecList in the header), so nothing is repacked on the host.
!
! NOT COMPILED IN THIS REPOSITORY: the build image has no Fortran compiler.  tests/test_abi.py checks
! that every `bind(C, name=...)` below names a symbol librxn_b200.so exports, and that the derived
! types have the field order of the C structs.
module Rxn_B200_module

  use, intrinsic :: iso_c_binding

  implicit none

  private

  integer(c_int), parameter, public :: RXN_OK = 0, RXN_ERR_INVALID = 1, RXN_ERR_UNSUPPORTED = 2, &
                                       RXN_ERR_CUDA = 3, RXN_ERR_NO_DEVICE = 4, RXN_ERR_CELL_FAILED = 5
  integer(c_int), parameter, public :: RXN_DT_AS_WRITTEN = 0, RXN_DT_CONSISTENT = 1
  integer(c_int), parameter, public :: RXN_EXIT_RESIDUAL = 1, RXN_EXIT_REL_CHANGE = 2, &
                                       RXN_FLAG_CAPPED = 256, RXN_FLAG_LU_ZERO_ROW = 512, &
                                       RXN_FLAG_ACT_DIVERGED = 1024, RXN_FLAG_NONFINITE = 2048, &
                                       RXN_FLAG_INACTIVE = 4096
  integer(c_int), parameter, public :: RXN_LOGK_FIXED = 0, RXN_LOGK_FIT5 = 1, RXN_LOGK_HPT = 2
  ! RxnField (order of the C enum)
  integer(c_int), parameter, public :: RXN_F_PRI_MOLAL = 0, RXN_F_TOTAL = 1, RXN_F_SEC_MOLAL = 2, &
    RXN_F_PRI_ACT_COEF = 3, RXN_F_SEC_ACT_COEF = 4, RXN_F_LN_ACT_H2O = 5, RXN_F_TOTAL_SORB_EQ = 6, &
    RXN_F_FREE_SITE_CONC = 7, RXN_F_EQSRFCPLX_CONC = 8, RXN_F_KINMR_TOTAL_SORB = 9, &
    RXN_F_EQIONX_REF_CATION_SORBED_CONC = 10, RXN_F_EQIONX_CONC = 11, RXN_F_MNRL_VOLFRAC = 12, &
    RXN_F_MNRL_AREA = 13, RXN_F_MNRL_RATE = 14, RXN_F_DEN_KG = 15, RXN_F_SAT = 16, RXN_F_TEMP = 17, &
    RXN_F_PRES = 18, RXN_F_VOLUME = 19, RXN_F_POROSITY = 20, RXN_F_SOIL_PARTICLE_DENSITY = 21, &
    RXN_F_DTOTAL = 22, RXN_F_DTOTAL_SORB_EQ = 23, RXN_F_KINSRFCPLX_CONC = 24, RXN_F_KINSRFCPLX_CONC_KP1 = 25, &
    RXN_F_KINSRFCPLX_FREE_SITE_CONC = 26, RXN_F_IMMOBILE = 27

  ! struct RxnSpecList
  type, bind(C), public :: rxn_spec_list_type
    type(c_ptr) :: id         ! specid(0:m,n)
    type(c_ptr) :: stoich     ! stoich(0:m,n) or stoich(m,n)
    type(c_ptr) :: h2oid
    type(c_ptr) :: h2ostoich
    type(c_ptr) :: logK
    type(c_ptr) :: logKcoef
    integer(c_int32_t) :: id_ld
    integer(c_int32_t) :: stoich_ld
    integer(c_int32_t) :: stoich_off   ! 0: stoich(0:m,n), 1: stoich(m,n)
    integer(c_int32_t) :: n
  end type rxn_spec_list_type

  ! struct RxnTablesDesc (field order = include/rxn_b200.h)
  type, bind(C), public :: rxn_tables_desc_type
    integer(c_int32_t) :: struct_size
    integer(c_int32_t) :: naqcomp
    integer(c_int32_t) :: ncomp
    integer(c_int32_t) :: logK_mode
    integer(c_int32_t) :: num_logK_coef
    integer(c_int32_t) :: use_log_formulation
    integer(c_int32_t) :: act_coef_update_frequency
    integer(c_int32_t) :: act_coef_update_algorithm
    integer(c_int32_t) :: use_activity_h2o
    integer(c_int32_t) :: h2o_aq_id
    integer(c_int32_t) :: h_ion_id
    integer(c_int32_t) :: reserved0
    real(c_double) :: debyeA, debyeB, debyeBdot
    real(c_double) :: max_dlnC, max_relative_change_tolerance, max_residual_tolerance
    type(c_ptr) :: primary_spec_Z
    type(c_ptr) :: primary_spec_a0
    type(rxn_spec_list_type) :: eqcplx
    type(c_ptr) :: eqcplx_Z
    type(c_ptr) :: eqcplx_a0
    type(rxn_spec_list_type) :: kinmnrl
    type(c_ptr) :: kinmnrl_rate_constant
    type(c_ptr) :: kinmnrl_activation_energy
    type(c_ptr) :: kinmnrl_molar_vol
    type(c_ptr) :: kinmnrl_affinity_threshold
    type(c_ptr) :: kinmnrl_rate_limiter
    type(c_ptr) :: kinmnrl_Temkin_const        ! c_null_ptr <=> .not.associated()
    type(c_ptr) :: kinmnrl_min_scale_factor
    type(c_ptr) :: kinmnrl_affinity_power
    type(c_ptr) :: kinmnrl_num_prefactors
    type(c_ptr) :: kinmnrl_pref_rate
    type(c_ptr) :: kinmnrl_pref_activation_energy
    type(c_ptr) :: kinmnrl_prefactor_id
    type(c_ptr) :: kinmnrl_pref_alpha
    type(c_ptr) :: kinmnrl_pref_beta
    type(c_ptr) :: kinmnrl_pref_atten_coef
    integer(c_int32_t) :: max_num_prefactors
    integer(c_int32_t) :: max_num_prefactor_species
    type(rxn_spec_list_type) :: mnrl
    type(rxn_spec_list_type) :: paseq
    type(rxn_spec_list_type) :: srfcplx
    type(c_ptr) :: srfcplx_free_site_stoich
    type(c_ptr) :: srfcplx_Z
    integer(c_int32_t) :: nsrfcplxrxn
    integer(c_int32_t) :: srfcplxrxn_to_complex_ld
    type(c_ptr) :: srfcplxrxn_to_surf
    type(c_ptr) :: srfcplxrxn_surf_type
    type(c_ptr) :: srfcplxrxn_to_complex
    type(c_ptr) :: srfcplxrxn_stoich_flag
    type(c_ptr) :: srfcplxrxn_site_density
    integer(c_int32_t) :: neqsrfcplxrxn
    integer(c_int32_t) :: nkinmrsrfcplxrxn
    type(c_ptr) :: eqsrfcplxrxn_to_srfcplxrxn
    type(c_ptr) :: kinmrsrfcplxrxn_to_srfcplxrxn
    type(c_ptr) :: kinmr_nrate
    type(c_ptr) :: kinmr_rate
    type(c_ptr) :: kinmr_frac
    integer(c_int32_t) :: kinmr_ld
    integer(c_int32_t) :: nkinsrfcplxrxn
    integer(c_int32_t) :: neqionxrxn
    integer(c_int32_t) :: eqionx_ld
    type(c_ptr) :: eqionx_rxn_cationid
    type(c_ptr) :: eqionx_rxn_k
    type(c_ptr) :: eqionx_rxn_CEC
    type(c_ptr) :: eqionx_rxn_Z_flag
    type(c_ptr) :: eqionx_rxn_to_surf
    integer(c_int32_t) :: neqkdrxn
    integer(c_int32_t) :: reserved1
    type(c_ptr) :: eqkdspecid
    type(c_ptr) :: eqkdtype
    type(c_ptr) :: eqkdmineral
    type(c_ptr) :: eqkddistcoef
    type(c_ptr) :: eqkdlangmuirb
    type(c_ptr) :: eqkdfreundlichn
    integer(c_int32_t) :: nactive_gas, nimmobile, ncoll, ngeneral_rxn, nradiodecay_rxn, nmicrobial_rxn, &
                          nimmobile_decay_rxn, has_sandbox, has_clm, has_solid_solution, co2_flow_mode, &
                          numerical_derivatives
    ! general reactions (reaction%general*), radioactive decay (reaction%radiodecay*), kinetic surface complexation
    integer(c_int32_t) :: general_ld
    integer(c_int32_t) :: radiodecay_ld
    type(c_ptr) :: generalspecid
    type(c_ptr) :: generalstoich
    type(c_ptr) :: generalforwardspecid
    type(c_ptr) :: generalforwardstoich
    type(c_ptr) :: generalbackwardspecid
    type(c_ptr) :: generalbackwardstoich
    type(c_ptr) :: general_kf
    type(c_ptr) :: general_kr
    type(c_ptr) :: radiodecayspecid
    type(c_ptr) :: radiodecaystoich
    type(c_ptr) :: radiodecayforwardspecid
    type(c_ptr) :: radiodecay_kf
    type(c_ptr) :: kinsrfcplxrxn_to_srfcplxrxn
    type(c_ptr) :: kinsrfcplx_forward_rate
    type(c_ptr) :: kinsrfcplx_backward_rate
    integer(c_int32_t) :: kinsrfcplx_ld
    integer(c_int32_t) :: reserved2
    ! immobile decay (reaction%immobile%decay*), microbial reactions (reaction%microbial%*)
    type(c_ptr) :: immobile_decayspecid
    type(c_ptr) :: immobile_decay_rate_constant
    integer(c_int32_t) :: microbial_ld
    integer(c_int32_t) :: microbial_monod_ld
    integer(c_int32_t) :: microbial_inhibition_ld
    integer(c_int32_t) :: nmicrobial_monod, nmicrobial_inhibition, reserved3
    type(c_ptr) :: microbial_specid
    type(c_ptr) :: microbial_stoich
    type(c_ptr) :: microbial_rate_constant
    type(c_ptr) :: microbial_activation_energy
    type(c_ptr) :: microbial_biomassid
    type(c_ptr) :: microbial_biomass_yield
    type(c_ptr) :: microbial_monodid
    type(c_ptr) :: microbial_inhibitionid
    type(c_ptr) :: microbial_monod_specid
    type(c_ptr) :: microbial_monod_K
    type(c_ptr) :: microbial_monod_Cth
    type(c_ptr) :: microbial_inhibition_type
    type(c_ptr) :: microbial_inhibition_specid
    type(c_ptr) :: microbial_inhibition_C
    type(c_ptr) :: microbial_inhibition_C2
  end type rxn_tables_desc_type

  public :: rxn_tables_create, rxn_tables_destroy, rxn_state_create, rxn_state_destroy, &
            rxn_state_ncells, rxn_state_materialize, rxn_field_rows, rxn_state_upload, &
            rxn_state_download, rxn_state_broadcast, rxn_set_cell_scalars, rxn_react_batch, &
            rxn_update_auxvars_batch, rxn_fixed_accum_batch, rxn_residual_blocks_batch, &
            rxn_jacobian_blocks_batch, rxn_residual_jacobian_blocks_batch, &
            rxn_update_kinetic_state_batch, rxn_equilibrate_constraint_batch, rxn_update_auxvars_batch_device, &
            rxn_residual_jacobian_blocks_batch_device, rxn_connset_create, rxn_connset_destroy, &
            rxn_connset_structure, rxn_connset_device_structure, rxn_connset_flux_coefs, rxn_flux_residual_batch, &
            rxn_flux_jacobian_batch, rxn_flux_residual_batch_device, rxn_flux_jacobian_batch_device, &
            rxn_couplerset_create, rxn_couplerset_destroy, rxn_couplerset_bc_coefs, rxn_couplerset_ss_coefs, &
            rxn_couplerset_set_totals, rxn_couplerset_totals_from_state, rxn_coupler_residual_batch, &
            rxn_coupler_jacobian_batch, rxn_coupler_residual_batch_device, rxn_coupler_jacobian_batch_device, &
            rxn_last_kernel_ms, rxn_last_error

  interface

    ! replaces: reading `reaction` inside every per-cell call
    integer(c_int) function rxn_tables_create(desc, device, tables) bind(C, name='rxn_tables_create')
      import :: c_int, c_ptr, rxn_tables_desc_type
      type(rxn_tables_desc_type), intent(in) :: desc
      integer(c_int), value :: device
      type(c_ptr), intent(out) :: tables
    end function

    integer(c_int) function rxn_tables_destroy(tables) bind(C, name='rxn_tables_destroy')
      import :: c_int, c_ptr
      type(c_ptr), value :: tables
    end function

    ! replaces: RTAuxVarInit per ghosted cell (reactive_transport_aux.F90:213-400)
    integer(c_int) function rxn_state_create(tables, ncells_ghosted, state) bind(C, name='rxn_state_create')
      import :: c_int, c_int64_t, c_ptr
      type(c_ptr), value :: tables
      integer(c_int64_t), value :: ncells_ghosted
      type(c_ptr), intent(out) :: state
    end function

    integer(c_int) function rxn_state_destroy(state) bind(C, name='rxn_state_destroy')
      import :: c_int, c_ptr
      type(c_ptr), value :: state
    end function

    integer(c_int64_t) function rxn_state_ncells(state) bind(C, name='rxn_state_ncells')
      import :: c_int64_t, c_ptr
      type(c_ptr), value :: state
    end function

    integer(c_int) function rxn_state_materialize(state, field) bind(C, name='rxn_state_materialize')
      import :: c_int, c_ptr
      type(c_ptr), value :: state
      integer(c_int), value :: field
    end function

    integer(c_int32_t) function rxn_field_rows(tables, field) bind(C, name='rxn_field_rows')
      import :: c_int32_t, c_int, c_ptr
      type(c_ptr), value :: tables
      integer(c_int), value :: field
    end function

    ! replaces: direct rt_auxvar field access (PatchGetVariable patch.F90:3529-4788, checkpoint pm_rt.F90:1159-1305)
    integer(c_int) function rxn_state_upload(state, field, host, row_stride, cell_stride) bind(C, name='rxn_state_upload')
      import :: c_int, c_int64_t, c_ptr
      type(c_ptr), value :: state
      integer(c_int), value :: field
      type(c_ptr), value :: host
      integer(c_int64_t), value :: row_stride, cell_stride
    end function

    integer(c_int) function rxn_state_download(state, field, host, row_stride, cell_stride) bind(C, name='rxn_state_download')
      import :: c_int, c_int64_t, c_ptr
      type(c_ptr), value :: state
      integer(c_int), value :: field
      type(c_ptr), value :: host
      integer(c_int64_t), value :: row_stride, cell_stride
    end function

    ! replaces: CondControlAssignTranInitCond region fill (condition_control.F90:498-949)
    integer(c_int) function rxn_state_broadcast(state, field, row_values) bind(C, name='rxn_state_broadcast')
      import :: c_int, c_ptr
      type(c_ptr), value :: state
      integer(c_int), value :: field
      type(c_ptr), value :: row_values
    end function

    ! replaces: global_auxvar / material_auxvar reads and the imat<=0 skip (reactive_transport.F90:1699)
    integer(c_int) function rxn_set_cell_scalars(state, den_kg, sat, temp, pres, volume, porosity, &
                                                 soil_particle_density, active) bind(C, name='rxn_set_cell_scalars')
      import :: c_int, c_ptr
      type(c_ptr), value :: state, den_kg, sat, temp, pres, volume, porosity, soil_particle_density, active
    end function

    ! replaces: the RTReact cell loop (reactive_transport.F90:1697-1724)
    integer(c_int) function rxn_react_batch(state, tran_xx, l2g, nlocal, dt, dt_mode, iters_out, flags_out) &
        bind(C, name='rxn_react_batch')
      import :: c_int, c_int64_t, c_double, c_ptr
      type(c_ptr), value :: state, tran_xx, l2g
      integer(c_int64_t), value :: nlocal
      real(c_double), value :: dt
      integer(c_int), value :: dt_mode
      type(c_ptr), value :: iters_out, flags_out
    end function

    ! replaces: RTUpdateAuxVars cells part (:3790-3846) + RTUpdateActivityCoefficients (:3620-3700)
    integer(c_int) function rxn_update_auxvars_batch(state, xx_loc, update_act_coefs) bind(C, name='rxn_update_auxvars_batch')
      import :: c_int, c_ptr
      type(c_ptr), value :: state, xx_loc
      integer(c_int), value :: update_act_coefs
    end function

    ! replaces: RTUpdateFixedAccumulation (:786-843)
    integer(c_int) function rxn_fixed_accum_batch(state, xx, l2g, nlocal, accum_out) bind(C, name='rxn_fixed_accum_batch')
      import :: c_int, c_int64_t, c_ptr
      type(c_ptr), value :: state, xx, l2g
      integer(c_int64_t), value :: nlocal
      type(c_ptr), value :: accum_out
    end function

    ! replaces: RTResidualNonFlux accumulation + reaction loops (:2545-2586, 2735-2758)
    integer(c_int) function rxn_residual_blocks_batch(state, l2g, nlocal, dt, res_out) bind(C, name='rxn_residual_blocks_batch')
      import :: c_int, c_int64_t, c_double, c_ptr
      type(c_ptr), value :: state, l2g
      integer(c_int64_t), value :: nlocal
      real(c_double), value :: dt
      type(c_ptr), value :: res_out
    end function

    ! replaces: RTJacobianNonFlux loops (:3342-3389, 3445-3465)
    integer(c_int) function rxn_jacobian_blocks_batch(state, l2g, nlocal, dt, jac_out) bind(C, name='rxn_jacobian_blocks_batch')
      import :: c_int, c_int64_t, c_double, c_ptr
      type(c_ptr), value :: state, l2g
      integer(c_int64_t), value :: nlocal
      real(c_double), value :: dt
      type(c_ptr), value :: jac_out
    end function

    integer(c_int) function rxn_residual_jacobian_blocks_batch(state, l2g, nlocal, dt, res_out, jac_out) &
        bind(C, name='rxn_residual_jacobian_blocks_batch')
      import :: c_int, c_int64_t, c_double, c_ptr
      type(c_ptr), value :: state, l2g
      integer(c_int64_t), value :: nlocal
      real(c_double), value :: dt
      type(c_ptr), value :: res_out, jac_out
    end function

    ! replaces: RTUpdateKineticState loop (:692-705)
    integer(c_int) function rxn_update_kinetic_state_batch(state, dt) bind(C, name='rxn_update_kinetic_state_batch')
      import :: c_int, c_double, c_ptr
      type(c_ptr), value :: state
      real(c_double), value :: dt
    end function

    ! replaces: ReactionEquilibrateConstraint (reaction.F90:1308-2046) cell by cell
    ! (CondControlAssignTranInitCond condition_control.F90:725-741, PatchInitCouplerConstraints patch.F90:3346-3461)
    integer(c_int) function rxn_equilibrate_constraint_batch(state, constraint_type, constraint_conc, conc_stride, &
                                                             constraint_id, free_ion_guess, use_prev_soln_as_guess, &
                                                             initialize_with_molality, l2g, nlocal, basis_molarity_out, &
                                                             iters_out, status_out) &
        bind(C, name='rxn_equilibrate_constraint_batch')
      import :: c_int, c_int32_t, c_int64_t, c_ptr
      type(c_ptr), value :: state
      type(c_ptr), value :: constraint_type      ! integer(c_int32_t) (naqcomp): aq_species_constraint%constraint_type
      type(c_ptr), value :: constraint_conc      ! real(c_double) (naqcomp) or (conc_stride, nlocal)
      integer(c_int64_t), value :: conc_stride   ! 0: one constraint for all cells
      type(c_ptr), value :: constraint_id        ! integer(c_int32_t) (naqcomp): constraint_spec_id (1-based mineral / gas id)
      type(c_ptr), value :: free_ion_guess       ! real(c_double) (naqcomp) or c_null_ptr
      integer(c_int), value :: use_prev_soln_as_guess, initialize_with_molality
      type(c_ptr), value :: l2g                  ! integer(c_int32_t) (nlocal) ghosted id - 1, or c_null_ptr
      integer(c_int64_t), value :: nlocal
      type(c_ptr), value :: basis_molarity_out, iters_out, status_out
    end function

    ! device-resident variants (PETSc VECCUDA / MATAIJCUSPARSE arrays: VecCUDAGetArray / MatSeqAIJCUSPARSE pointers)
    integer(c_int) function rxn_update_auxvars_batch_device(state, d_xx_loc, update_act_coefs) &
        bind(C, name='rxn_update_auxvars_batch_device')
      import :: c_int, c_ptr
      type(c_ptr), value :: state, d_xx_loc
      integer(c_int), value :: update_act_coefs
    end function

    integer(c_int) function rxn_residual_jacobian_blocks_batch_device(state, d_l2g, nlocal, dt, d_res, d_jac) &
        bind(C, name='rxn_residual_jacobian_blocks_batch_device')
      import :: c_int, c_int64_t, c_double, c_ptr
      type(c_ptr), value :: state, d_l2g
      integer(c_int64_t), value :: nlocal
      real(c_double), value :: dt
      type(c_ptr), value :: d_res, d_jac
    end function

    ! flux side (RTResidualFlux / RTJacobianFlux interior loops, TFluxCoef / TFlux / TFluxDerivative): the connection set is
    ! grid%internal_connection_set_list flattened in loop order with 0-based ghosted ids; ghost_to_local = grid%nG2L - 1
    integer(c_int) function rxn_connset_create(state, nconn, id_up, id_dn, ghost_to_local, nlocal, active, connset) &
        bind(C, name='rxn_connset_create')
      import :: c_int, c_int64_t, c_ptr
      type(c_ptr), value :: state
      integer(c_int64_t), value :: nconn, nlocal
      type(c_ptr), value :: id_up, id_dn, ghost_to_local, active
      type(c_ptr) :: connset
    end function

    integer(c_int) function rxn_connset_destroy(connset) bind(C, name='rxn_connset_destroy')
      import :: c_int, c_ptr
      type(c_ptr), value :: connset
    end function

    integer(c_int) function rxn_connset_structure(connset, nnz_blocks, row_ptr, col) bind(C, name='rxn_connset_structure')
      import :: c_int, c_int64_t, c_ptr
      type(c_ptr), value :: connset
      integer(c_int64_t) :: nnz_blocks
      type(c_ptr), value :: row_ptr, col
    end function

    integer(c_int) function rxn_connset_device_structure(connset, d_row_ptr, d_col) bind(C, name='rxn_connset_device_structure')
      import :: c_int, c_ptr
      type(c_ptr), value :: connset
      type(c_ptr) :: d_row_ptr, d_col
    end function

    integer(c_int) function rxn_connset_flux_coefs(connset, area, velocity, disp_over_dist, fraction_upwind, use_upwinding) &
        bind(C, name='rxn_connset_flux_coefs')
      import :: c_int, c_ptr
      type(c_ptr), value :: connset, area, velocity, disp_over_dist, fraction_upwind
      integer(c_int), value :: use_upwinding
    end function

    integer(c_int) function rxn_flux_residual_batch(state, connset, res_out) bind(C, name='rxn_flux_residual_batch')
      import :: c_int, c_ptr
      type(c_ptr), value :: state, connset, res_out
    end function

    integer(c_int) function rxn_flux_jacobian_batch(state, connset, val_out) bind(C, name='rxn_flux_jacobian_batch')
      import :: c_int, c_ptr
      type(c_ptr), value :: state, connset, val_out
    end function

    integer(c_int) function rxn_flux_residual_batch_device(state, connset, d_res) bind(C, name='rxn_flux_residual_batch_device')
      import :: c_int, c_ptr
      type(c_ptr), value :: state, connset, d_res
    end function

    integer(c_int) function rxn_flux_jacobian_batch_device(state, connset, d_val) bind(C, name='rxn_flux_jacobian_batch_device')
      import :: c_int, c_ptr
      type(c_ptr), value :: state, connset, d_val
    end function

    ! boundary conditions (RTResidualFlux :2347-2430, RTJacobianFlux :3176-3240) and source/sinks (RTResidualNonFlux :2623-2672,
    ! RTJacobianNonFlux :3394-3436): kind 0 = patch%boundary_condition_list, 1 = patch%source_sink_list, flattened in loop order
    ! (sum_connection); id_dn = 0-based ghosted id of the cell of each connection
    integer(c_int) function rxn_couplerset_create(state, kind, nconn, id_dn, ghost_to_local, nlocal, active, couplerset) &
        bind(C, name='rxn_couplerset_create')
      import :: c_int, c_int64_t, c_ptr
      type(c_ptr), value :: state
      integer(c_int), value :: kind
      integer(c_int64_t), value :: nconn, nlocal
      type(c_ptr), value :: id_dn, ghost_to_local, active
      type(c_ptr) :: couplerset
    end function

    integer(c_int) function rxn_couplerset_destroy(couplerset) bind(C, name='rxn_couplerset_destroy')
      import :: c_int, c_ptr
      type(c_ptr), value :: couplerset
    end function

    ! area = connection%area, velocity = patch%boundary_velocities(1,:), disp_over_dist = patch%boundary_tran_coefs(:,1,:)
    integer(c_int) function rxn_couplerset_bc_coefs(couplerset, area, velocity, disp_over_dist, use_upwinding) &
        bind(C, name='rxn_couplerset_bc_coefs')
      import :: c_int, c_ptr
      type(c_ptr), value :: couplerset, area, velocity, disp_over_dist
      integer(c_int), value :: use_upwinding
    end function

    ! qsrc = patch%ss_flow_vol_fluxes(1,:), tran_src_sink_type = source_sink%tran_condition%itype per connection
    integer(c_int) function rxn_couplerset_ss_coefs(couplerset, qsrc, tran_src_sink_type) bind(C, name='rxn_couplerset_ss_coefs')
      import :: c_int, c_ptr
      type(c_ptr), value :: couplerset, qsrc, tran_src_sink_type
    end function

    integer(c_int) function rxn_couplerset_set_totals(couplerset, total) bind(C, name='rxn_couplerset_set_totals')
      import :: c_int, c_ptr
      type(c_ptr), value :: couplerset, total
    end function

    ! bc_state: the state that replaces rt_auxvars_bc(:) (one cell per boundary connection)
    integer(c_int) function rxn_couplerset_totals_from_state(couplerset, bc_state) bind(C, name='rxn_couplerset_totals_from_state')
      import :: c_int, c_ptr
      type(c_ptr), value :: couplerset, bc_state
    end function

    integer(c_int) function rxn_coupler_residual_batch(state, couplerset, res_inout, flux_out) bind(C, name='rxn_coupler_residual_batch')
      import :: c_int, c_ptr
      type(c_ptr), value :: state, couplerset, res_inout, flux_out
    end function

    integer(c_int) function rxn_coupler_jacobian_batch(state, connset, couplerset, val_inout) bind(C, name='rxn_coupler_jacobian_batch')
      import :: c_int, c_ptr
      type(c_ptr), value :: state, connset, couplerset, val_inout
    end function

    integer(c_int) function rxn_coupler_residual_batch_device(state, couplerset, d_res, d_flux_out) &
        bind(C, name='rxn_coupler_residual_batch_device')
      import :: c_int, c_ptr
      type(c_ptr), value :: state, couplerset, d_res, d_flux_out
    end function

    integer(c_int) function rxn_coupler_jacobian_batch_device(state, connset, couplerset, d_val) &
        bind(C, name='rxn_coupler_jacobian_batch_device')
      import :: c_int, c_ptr
      type(c_ptr), value :: state, connset, couplerset, d_val
    end function

    real(c_float) function rxn_last_kernel_ms(state) bind(C, name='rxn_last_kernel_ms')
      import :: c_float, c_ptr
      type(c_ptr), value :: state
    end function

    ! replaces: option%io_buffer
    integer(c_int) function rxn_last_error(buf, len) bind(C, name='rxn_last_error')
      import :: c_int, c_int32_t, c_char
      character(kind=c_char), intent(out) :: buf(*)
      integer(c_int32_t), value :: len
    end function

  end interface

end module Rxn_B200_module
