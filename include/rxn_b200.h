/*
 * rxn_b200.h — C ABI of the B200-native batched reaction path for PFLOTRAN.
 *
 * The reference (petsc/pflotran) has no C or plugin ABI for this path: the
 * boundary is the set of public Fortran module procedures of Reaction_module
 * (src/pflotran/reaction.F90:38-70) called one cell at a time from the loops in
 * src/pflotran/reactive_transport.F90.  Each entry point below is the BATCHED
 * replacement of one of those loops; the citation on each declaration names the
 * reference loop / routine it replaces.  A thin `use iso_c_binding` module
 * (fortran/rxn_b200_shim.F90) binds these names 1:1.
 *
 * Conventions
 *   - plain C, no exceptions, no aborts: every function returns an RxnStatus;
 *     rxn_last_error() returns the message (reference: option%io_buffer +
 *     printErrMsg, src/pflotran/option.F90:630-677).
 *   - all reals are IEEE double (PetscReal), all indices int32 (PetscInt),
 *     species ids are 1-BASED exactly as stored by the reference, so Fortran can
 *     pass `c_loc(reaction%eqcplxspecid)` etc. without repacking.
 *   - array shapes are the Fortran shapes in Fortran memory order; see
 *     RxnSpecList.
 *   - host buffers belong to the caller and are only touched during the call.
 *   - one handle is bound to one CUDA device; calls are synchronous at return.
 */
#ifndef RXN_B200_H
#define RXN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum RxnStatus {
  RXN_OK = 0,
  RXN_ERR_INVALID = 1,      /* bad argument / inconsistent descriptor            */
  RXN_ERR_UNSUPPORTED = 2,  /* tables enable a reaction type outside the path    */
  RXN_ERR_CUDA = 3,         /* CUDA runtime error (message has the detail)       */
  RXN_ERR_NO_DEVICE = 4,    /* no usable GPU: there is NO CPU fallback           */
  RXN_ERR_CELL_FAILED = 5   /* >=1 cell raised a flag the reference would abort on */
} RxnStatus;

/* RReact dt handling (reference src/pflotran/reaction.F90:3347,3362,3424-3426
 * vs :5189-5191): AS_WRITTEN keeps residual = accum - fixed_accum [mol] with J/dt;
 * DT_CONSISTENT uses residual = (accum - fixed_accum)/dt as the live global-implicit
 * caller does (src/pflotran/reactive_transport.F90:2564). */
typedef enum RxnDtMode { RXN_DT_AS_WRITTEN = 0, RXN_DT_CONSISTENT = 1 } RxnDtMode;

/* per-cell exit reason / flags written by rxn_react_batch (flags_out) */
enum {
  RXN_EXIT_RESIDUAL   = 1,   /* max|residual| < max_residual_tolerance  (reaction.F90:3443) */
  RXN_EXIT_REL_CHANGE = 2,   /* max rel. change < tolerance             (reaction.F90:3476) */
  RXN_FLAG_CAPPED       = 1 << 8,   /* iteration cap hit (reference would spin, :3478-3496) */
  RXN_FLAG_LU_ZERO_ROW  = 1 << 9,   /* ludcmp all-zero row (reference MPI_Abort, utility.F90:423) */
  RXN_FLAG_ACT_DIVERGED = 1 << 10,  /* activity-coef Newton > 50 its (reaction.F90:3864) */
  RXN_FLAG_NONFINITE    = 1 << 11,
  RXN_FLAG_INACTIVE     = 1 << 12   /* cell skipped: imat<=0 (reactive_transport.F90:1699) */
};

/* reaction_aux.F90:23-27 */
enum { RXN_ACT_COEF_FREQUENCY_OFF = 0, RXN_ACT_COEF_FREQUENCY_TIMESTEP = 1,
       RXN_ACT_COEF_FREQUENCY_NEWTON_ITER = 2,
       RXN_ACT_COEF_ALGORITHM_LAG = 3, RXN_ACT_COEF_ALGORITHM_NEWTON = 4 };
/* reaction_surf_complex_aux.F90:19-22 */
enum { RXN_NULL_SURFACE = 0, RXN_COLLOID_SURFACE = 1, RXN_MINERAL_SURFACE = 2, RXN_ROCK_SURFACE = 3 };
/* pflotran_constants.F90:94-96 */
enum { RXN_SORPTION_LINEAR = 1, RXN_SORPTION_LANGMUIR = 2, RXN_SORPTION_FREUNDLICH = 3 };
/* logK evaluation: 0 tables hold logK at the (isothermal) reference temperature;
 * 1 per-cell 5-term fit (reaction_aux.F90:1461-1488); 2 per-cell 17-term hpt form (:1529-1571) */
enum { RXN_LOGK_FIXED = 0, RXN_LOGK_FIT5 = 1, RXN_LOGK_HPT = 2 };

/*
 * A reaction list in the reference's compressed form:
 *   Fortran  specid(0:m,n)  -> id[r*id_ld + 0] = nspec, id[r*id_ld + i] = species i (1-based id)
 *   Fortran  stoich(0:m,n)  -> stoich_off = 0 : stoich of species i at stoich[r*stoich_ld + i]
 *   Fortran  stoich(m,n)    -> stoich_off = 1 : stoich of species i at stoich[r*stoich_ld + i - 1]
 * h2oid[r] > 0 means H2O takes part with stoichiometry h2ostoich[r].
 */
typedef struct RxnSpecList {
  const int32_t *id;
  const double *stoich;
  const int32_t *h2oid;
  const double *h2ostoich;
  const double *logK;      /* [n]                     */
  const double *logKcoef;  /* [n][num_logK_coef] or NULL */
  int32_t id_ld;
  int32_t stoich_ld;
  int32_t stoich_off;
  int32_t n;
} RxnSpecList;

/* Flat view of reaction_type (reference src/pflotran/reaction_aux.F90:142-335),
 * mineral_type (reaction_mineral_aux.F90:77-128) and surface_complexation_type
 * (reaction_surf_complex_aux.F90:68-128).  Pointers may be NULL when the count is 0. */
typedef struct RxnTablesDesc {
  int32_t struct_size;            /* = sizeof(RxnTablesDesc) */
  int32_t naqcomp;
  int32_t ncomp;                  /* = naqcomp + nimmobile (reaction%ncomp; immobile dofs follow the aqueous ones, offset_immobile = naqcomp; no colloid dofs) */
  int32_t logK_mode;              /* RXN_LOGK_*            */
  int32_t num_logK_coef;
  int32_t use_log_formulation;
  int32_t act_coef_update_frequency;
  int32_t act_coef_update_algorithm;
  int32_t use_activity_h2o;
  int32_t h2o_aq_id;              /* species_idx%h2o_aq_id, 0 if none */
  int32_t h_ion_id;               /* species_idx%h_ion_id (>0 primary, <0 complex, 0 none) */
  int32_t reserved0;
  double debyeA, debyeB, debyeBdot;
  double max_dlnC, max_relative_change_tolerance, max_residual_tolerance;
  const double *primary_spec_Z;   /* [naqcomp] */
  const double *primary_spec_a0;  /* [naqcomp] */

  /* aqueous complexes: eqcplxspecid(0:m,n), eqcplxstoich(0:m,n) -> stoich_off 0 */
  RxnSpecList eqcplx;
  const double *eqcplx_Z;
  const double *eqcplx_a0;

  /* kinetic minerals: kinmnrlspecid(0:m,n), kinmnrlstoich(m,n) -> stoich_off 1 */
  RxnSpecList kinmnrl;
  const double *kinmnrl_rate_constant;
  const double *kinmnrl_activation_energy;
  const double *kinmnrl_molar_vol;
  const double *kinmnrl_affinity_threshold;
  const double *kinmnrl_rate_limiter;
  const double *kinmnrl_Temkin_const;      /* NULL <=> not associated() */
  const double *kinmnrl_min_scale_factor;  /* NULL <=> not associated() */
  const double *kinmnrl_affinity_power;    /* NULL <=> not associated() */
  const int32_t *kinmnrl_num_prefactors;   /* [nkin] */
  const double *kinmnrl_pref_rate;               /* (maxpref, nkin)              */
  const double *kinmnrl_pref_activation_energy;  /* (maxpref, nkin)              */
  const int32_t *kinmnrl_prefactor_id;           /* (0:maxprefspec, maxpref, nkin) */
  const double *kinmnrl_pref_alpha;              /* (maxprefspec, maxpref, nkin)   */
  const double *kinmnrl_pref_beta;
  const double *kinmnrl_pref_atten_coef;
  int32_t max_num_prefactors;
  int32_t max_num_prefactor_species;

  /* all minerals + passive gases: used only by constraint equilibration */
  RxnSpecList mnrl;     /* mnrlstoich(m,n) -> stoich_off 1 */
  RxnSpecList paseq;    /* paseqstoich(0:m,n) -> stoich_off 0 */

  /* surface complexation */
  RxnSpecList srfcplx;  /* srfcplxstoich(m,n) -> stoich_off 1 */
  const double *srfcplx_free_site_stoich;
  const double *srfcplx_Z;
  int32_t nsrfcplxrxn;
  int32_t srfcplxrxn_to_complex_ld;           /* srfcplxrxn_to_complex(0:mc, nrxn) */
  const int32_t *srfcplxrxn_to_surf;
  const int32_t *srfcplxrxn_surf_type;
  const int32_t *srfcplxrxn_to_complex;
  const int32_t *srfcplxrxn_stoich_flag;
  const double *srfcplxrxn_site_density;
  int32_t neqsrfcplxrxn;
  int32_t nkinmrsrfcplxrxn;
  const int32_t *eqsrfcplxrxn_to_srfcplxrxn;
  const int32_t *kinmrsrfcplxrxn_to_srfcplxrxn;
  const int32_t *kinmr_nrate;                 /* (0:nkinmr): [0] = max */
  const double *kinmr_rate;                   /* (maxrate, nkinmr) */
  const double *kinmr_frac;                   /* (maxrate, nkinmr) */
  int32_t kinmr_ld;                           /* = maxrate */
  int32_t nkinsrfcplxrxn;                     /* 0 or 1: tables at the end of this struct */

  /* ion exchange: eqionx_rxn_cationid(0:mc,n), eqionx_rxn_k(mc,n) */
  int32_t neqionxrxn;
  int32_t eqionx_ld;                          /* = mc (ids use mc+1) */
  const int32_t *eqionx_rxn_cationid;
  const double *eqionx_rxn_k;
  const double *eqionx_rxn_CEC;
  const int32_t *eqionx_rxn_Z_flag;
  const int32_t *eqionx_rxn_to_surf;

  /* KD isotherms */
  int32_t neqkdrxn;
  int32_t reserved1;
  const int32_t *eqkdspecid;
  const int32_t *eqkdtype;
  const int32_t *eqkdmineral;
  const double *eqkddistcoef;
  const double *eqkdlangmuirb;
  const double *eqkdfreundlichn;

  /* reaction types outside the path: must be 0, else RXN_ERR_UNSUPPORTED
   * (active gas/RTotalGas, colloids, sandbox, CLM, solid solution, CO2 flow modes -> RTotalCO2, numerical Jacobian).
   * ngeneral_rxn, nradiodecay_rxn, nmicrobial_rxn, nimmobile_decay_rxn and nimmobile are COUNTS of supported reactions /
   * immobile species (tables below). */
  int32_t nactive_gas, nimmobile, ncoll, ngeneral_rxn, nradiodecay_rxn, nmicrobial_rxn,
          nimmobile_decay_rxn, has_sandbox, has_clm, has_solid_solution, co2_flow_mode,
          numerical_derivatives;

  /* general forward/backward-rate reactions, RGeneral (reaction.F90:4694-4831; tables reaction_database.F90:3011-3135) */
  int32_t general_ld;                         /* = m: generalspecid(0:m,n), generalstoich(m,n); forward/backward alike */
  int32_t radiodecay_ld;                      /* = m: radiodecayspecid(0:m,n), radiodecaystoich(m,n) */
  const int32_t *generalspecid;
  const double *generalstoich;
  const int32_t *generalforwardspecid;
  const double *generalforwardstoich;
  const int32_t *generalbackwardspecid;
  const double *generalbackwardstoich;
  const double *general_kf;                   /* [ngeneral_rxn] */
  const double *general_kr;
  /* radioactive decay with one reactant, RRadioactiveDecay (reaction.F90:4607-4690; tables reaction_database.F90:2915-3006) */
  const int32_t *radiodecayspecid;
  const double *radiodecaystoich;
  const int32_t *radiodecayforwardspecid;     /* [nradiodecay_rxn] */
  const double *radiodecay_kf;                /* [nradiodecay_rxn] 1/s */
  /* kinetic surface complexation, RKineticSurfCplx (reaction_surf_complex.F90:938-1137): nkinsrfcplxrxn (above) must be 0 or 1
   * - the reference allocates the per-cell concentrations for one kinetic reaction (reactive_transport_aux.F90:284-290) -
   * and that reaction must be surface complexation reaction 1 on mineral surface 1 (the reference indexes the site arrays
   * with the mineral id and the rate tables with the global complex id). */
  const int32_t *kinsrfcplxrxn_to_srfcplxrxn; /* [nkinsrfcplxrxn] */
  const double *kinsrfcplx_forward_rate;      /* (kinsrfcplx_ld, nkinsrfcplxrxn) */
  const double *kinsrfcplx_backward_rate;
  int32_t kinsrfcplx_ld;
  int32_t reserved2;
  /* immobile species (reaction_immobile_aux.F90:29-60; rt_auxvar%immobile [mol/m^3 bulk], dofs naqcomp+1 .. ncomp) and their
   * first-order decay, RImmobileDecay (reaction_immobile.F90:240-293) */
  const int32_t *immobile_decayspecid;        /* [nimmobile_decay_rxn] 1-based immobile species id */
  const double *immobile_decay_rate_constant; /* [nimmobile_decay_rxn] 1/s */
  /* microbial reactions, RMicrobial (reaction_microbial.F90:236-450; tables reaction_microbial_aux.F90:62-82, built at
   * reaction_database.F90:3126-3333).  Species ids run over the ncomp dofs (an immobile species i is naqcomp + i). */
  int32_t microbial_ld;                       /* = m: specid(0:m,n), stoich(m,n) */
  int32_t microbial_monod_ld;                 /* = m: monodid(0:m,n) */
  int32_t microbial_inhibition_ld;            /* = m: inhibitionid(0:m,n) */
  int32_t nmicrobial_monod, nmicrobial_inhibition, reserved3;
  const int32_t *microbial_specid;
  const double *microbial_stoich;
  const double *microbial_rate_constant;      /* [nmicrobial_rxn] */
  const double *microbial_activation_energy;  /* [nmicrobial_rxn] J/mol, or NULL (allocated only if some reaction sets one) */
  const int32_t *microbial_biomassid;         /* [nmicrobial_rxn] 1-based immobile species id, 0 = no biomass term */
  const double *microbial_biomass_yield;      /* [nmicrobial_rxn] */
  const int32_t *microbial_monodid;           /* 1-based ids into the monod_* arrays */
  const int32_t *microbial_inhibitionid;
  const int32_t *microbial_monod_specid;      /* [nmicrobial_monod] primary species */
  const double *microbial_monod_K;
  const double *microbial_monod_Cth;
  const int32_t *microbial_inhibition_type;   /* [nmicrobial_inhibition] RXN_INHIBITION_* */
  const int32_t *microbial_inhibition_specid;
  const double *microbial_inhibition_C;
  const double *microbial_inhibition_C2;
} RxnTablesDesc;

/* reaction_microbial_aux.F90:13-16 */
enum { RXN_INHIBITION_THRESHOLD = 1, RXN_INHIBITION_THERMODYNAMIC = 2, RXN_INHIBITION_MONOD = 3, RXN_INHIBITION_INVERSE_MONOD = 4 };

/* Per-cell state fields (reactive_transport_auxvar_type, reference
 * src/pflotran/reactive_transport_aux.F90:18-73; global_auxvar_type global_aux.F90:11-33;
 * material_auxvar_type material_aux.F90:26-43).  Device layout is SoA FP64,
 * field[row][cell], cell index fastest, leading dimension padded. */
typedef enum RxnField {
  RXN_F_PRI_MOLAL = 0,        /* rows = naqcomp                          */
  RXN_F_TOTAL,                /* naqcomp   total(:,1)                    */
  RXN_F_SEC_MOLAL,            /* neqcplx                                 */
  RXN_F_PRI_ACT_COEF,         /* naqcomp                                 */
  RXN_F_SEC_ACT_COEF,         /* neqcplx                                 */
  RXN_F_LN_ACT_H2O,           /* 1                                       */
  RXN_F_TOTAL_SORB_EQ,        /* naqcomp                                 */
  RXN_F_FREE_SITE_CONC,       /* nsrfcplxrxn  srfcplxrxn_free_site_conc  */
  RXN_F_EQSRFCPLX_CONC,       /* nsrfcplx                                */
  RXN_F_KINMR_TOTAL_SORB,     /* nkinmr*(maxrate+1)*naqcomp: row = (rxn*(maxrate+1)+rate)*naq+comp */
  RXN_F_EQIONX_REF_CATION_SORBED_CONC, /* neqionxrxn                      */
  RXN_F_EQIONX_CONC,          /* neqionxrxn*eqionx_ld: row = rxn*ld + cation */
  RXN_F_MNRL_VOLFRAC,         /* nkinmnrl                                */
  RXN_F_MNRL_AREA,            /* nkinmnrl                                */
  RXN_F_MNRL_RATE,            /* nkinmnrl                                */
  RXN_F_DEN_KG, RXN_F_SAT, RXN_F_TEMP, RXN_F_PRES,         /* 1 each      */
  RXN_F_VOLUME, RXN_F_POROSITY, RXN_F_SOIL_PARTICLE_DENSITY,
  RXN_F_DTOTAL,               /* naqcomp^2, column-major: row = j*naq + i (GI entry points only) */
  RXN_F_DTOTAL_SORB_EQ,       /* naqcomp^2, column-major                                         */
  RXN_F_KINSRFCPLX_CONC,      /* nkinsrfcplx (complexes of the kinetic surface complexation reaction): S^k     */
  RXN_F_KINSRFCPLX_CONC_KP1,  /* nkinsrfcplx: S^{k+1}, becomes S^k in rxn_update_kinetic_state_batch (reaction.F90:5411-5419) */
  RXN_F_KINSRFCPLX_FREE_SITE_CONC, /* nkinsrfcplxrxn */
  RXN_F_IMMOBILE,             /* nimmobile: rt_auxvar%immobile [mol/m^3 bulk] (reactive_transport_aux.F90:56) */
  RXN_F_COUNT
} RxnField;

typedef struct RxnTables RxnTables;   /* opaque */
typedef struct RxnState RxnState;     /* opaque */

/* replaces: reading `reaction` (reaction_type) inside every per-cell call.
 * Copies and repacks the tables to the device; immutable afterwards. */
int rxn_tables_create(const RxnTablesDesc *desc, int device, RxnTables **out);
int rxn_tables_destroy(RxnTables *t);

/* replaces: RTAuxVarInit per ghosted cell (reactive_transport_aux.F90:213-400),
 * allocation loop reactive_transport.F90:271-274.  Initial values as the reference:
 * act coefs = 1, free_site_conc = 1e-9, eqionx_ref_cation_sorbed_conc = 1e-9, rest 0. */
int rxn_state_create(const RxnTables *t, int64_t ncells_ghosted, RxnState **out);
int rxn_state_destroy(RxnState *s);
int64_t rxn_state_ncells(const RxnState *s);
/* DTOTAL / DTOTAL_SORB_EQ (naq^2 rows each) are not kept in HBM unless asked for: the
 * flux-side consumers (TFluxDerivative, transport.F90:368-626) need them, RReact does not.
 * After this call rxn_update_auxvars_batch / rxn_react_batch also store that field. */
int rxn_state_materialize(RxnState *s, int field);
/* RReact kernel: 0 = automatic (the tensor-memory kernel when the Newton matrix fits a TMEM lane, N <= 15; else the
 * resident-lane kernel; else thread per cell), 1 = one thread per cell with per-thread arrays in local memory,
 * 3 = on-chip kernels only (tensor memory / resident lane; fails if the tables do not fit them).  2 was round 1's cooperative
 * kernel and is rejected.  Benchmark / test control only; results are the same path. */
int rxn_set_react_kernel(RxnState *s, int which);
/* constraint types of ReactionEquilibrateConstraint (transport_constraint.F90:217-330, reaction_aux.F90 CONSTRAINT_*) */
enum { RXN_CONSTRAINT_NULL = 0, RXN_CONSTRAINT_FREE = 1, RXN_CONSTRAINT_TOTAL = 2, RXN_CONSTRAINT_LOG = 3, RXN_CONSTRAINT_PH = 4,
       RXN_CONSTRAINT_MINERAL = 5, RXN_CONSTRAINT_GAS = 6, RXN_CONSTRAINT_CHARGE_BAL = 7, RXN_CONSTRAINT_TOTAL_SORB = 9 };
/* per-cell status of rxn_equilibrate_constraint_batch: 0 converged; the others are the reference's fatal errors */
enum { RXN_EQ_OK = 0, RXN_EQ_NO_H_ION = 2, RXN_EQ_BAD_CONSTRAINT = 3, RXN_EQ_LU_ZERO_ROW = 4, RXN_EQ_ZERO_CONCENTRATION = 5,
       RXN_EQ_NOT_CONVERGED = 6 };

/* one-line description of the kernel rxn_react_batch would launch for this state (shape, shared memory) */
int rxn_react_kernel_info(const RxnState *s, char *buf, int32_t len);
int32_t rxn_field_rows(const RxnTables *t, int field);

/* replaces: ReactionEquilibrateConstraint (reaction.F90:1308-2046) applied cell by cell: initial / boundary condition
 * speciation (CondControlAssignTranInitCond condition_control.F90:725-741, PatchInitCouplerConstraints patch.F90:3346-3461).
 * One constraint for all cells (conc_stride = 0) or one concentration row per cell (conc_stride = naqcomp, dataset-driven
 * conditions); constraint_type / constraint_id [naqcomp] (RXN_CONSTRAINT_*, 1-based mineral / gas ids); free_ion_guess
 * [naqcomp] or NULL.  The cell state (den_kg, temp, mineral volume fractions, ...) must be set; on return it holds the
 * equilibrated pri_molal, activity coefficients, totals, sorbed state and the multirate sorbed totals.
 * basis_molarity_out [nlocal][naqcomp] (or NULL), iters_out / status_out [nlocal] (or NULL). */
int rxn_equilibrate_constraint_batch(RxnState *s, const int32_t *constraint_type, const double *constraint_conc, int64_t conc_stride,
                                     const int32_t *constraint_id, const double *free_ion_guess, int use_prev_soln_as_guess,
                                     int initialize_with_molality, const int32_t *l2g, int64_t nlocal, double *basis_molarity_out,
                                     int32_t *iters_out, int32_t *status_out);

/* replaces: direct field access by PatchGetVariable (patch.F90:3529-4788), checkpoint
 * (pm_rt.F90:1159-1305) and CondControlAssignTranInitCond (condition_control.F90:498-949).
 * host element (row r, cell c) lives at host[r*row_stride + c*cell_stride]. */
int rxn_state_upload(RxnState *s, int field, const double *host, int64_t row_stride,
                     int64_t cell_stride);
int rxn_state_download(const RxnState *s, int field, double *host, int64_t row_stride,
                       int64_t cell_stride);

/* replaces: assigning one constraint's equilibrated state to every cell of a region
 * (CondControlAssignTranInitCond, condition_control.F90:498-949): field(row, c) = row_values[row]
 * for all cells c. */
int rxn_state_broadcast(RxnState *s, int field, const double *row_values);

/* replaces: R2 reads + the imat<=0 / nG2L<0 skips (reactive_transport.F90:1699, 3791-3794).
 * Any pointer may be NULL (field left unchanged). active: 1 = compute, 0 = skip. */
int rxn_set_cell_scalars(RxnState *s, const double *den_kg, const double *sat, const double *temp,
                         const double *pres, const double *volume, const double *porosity,
                         const double *soil_particle_density, const uint8_t *active);

/* replaces: the RTReact cell loop (reactive_transport.F90:1697-1724) = RReact per cell
 * (reaction.F90:3322-3511).  tran_xx: AoS nlocal x ncomp, in: transported totals [mol/L],
 * out: free-ion molalities.  l2g: ghosted (state) index of local cell i, 0-based, or NULL
 * for identity.  iters_out/flags_out: int32[nlocal] or NULL.
 * Immobile dofs (ncomp > naqcomp): columns naqcomp.. of a row hold the immobile concentrations [mol/m^3 bulk], in and out.  Two
 * deviations from the reference's (untested) RReact there: the immobile values are written back to the cell's own row
 * (reactive_transport.F90:1712-1716 drops the cell offset), and with LOG_FORMULATION the immobile columns of the Newton matrix
 * are scaled by the immobile concentrations (reaction.F90:3445 hands RSolve pri_molal(naqcomp) as conc(ncomp)).  The
 * global-implicit entry points below have no deviation (pinned by the ABCD_microbial gold files). */
int rxn_react_batch(RxnState *s, double *tran_xx, const int32_t *l2g, int64_t nlocal, double dt,
                    int dt_mode, int32_t *iters_out, int32_t *flags_out);
/* same, with tran_xx / iters / flags already resident on the state's device (no PCIe). */
int rxn_react_batch_device(RxnState *s, double *d_tran_xx, const int32_t *d_l2g, int64_t nlocal,
                           double dt, int dt_mode, int32_t *d_iters, int32_t *d_flags);

/* replaces: RTUpdateAuxVars cells part (reactive_transport.F90:3790-3846) and
 * RTUpdateActivityCoefficients (:3620-3700).  xx_loc: AoS nghosted x ncomp free-ion
 * molalities (NULL: keep the state's pri_molal). */
int rxn_update_auxvars_batch(RxnState *s, const double *xx_loc, int update_act_coefs);

/* replaces: RTUpdateFixedAccumulation (reactive_transport.F90:786-843).
 * accum_out: AoS nlocal x ncomp [mol]. */
int rxn_fixed_accum_batch(RxnState *s, const double *xx, const int32_t *l2g, int64_t nlocal,
                          double *accum_out);

/* replaces: accumulation + reaction loops of RTResidualNonFlux (reactive_transport.F90:
 * 2545-2586, 2735-2758): res_out[cell] = (RTAccumulation + RAccumulationSorb)/dt + RReaction.
 * The caller subtracts its fixed accumulation / dt.  AoS nlocal x ncomp. */
int rxn_residual_blocks_batch(RxnState *s, const int32_t *l2g, int64_t nlocal, double dt,
                              double *res_out);
/* replaces: RTJacobianNonFlux loops (reactive_transport.F90:3342-3389, 3445-3465):
 * jac_out[cell] = RTAccumulationDerivative + RAccumulationSorbDerivative + RReactionDerivative,
 * ncomp x ncomp column-major per block, ready for MatSetValuesBlockedLocal
 * (MAT_ROW_ORIENTED = FALSE, discretization.F90:857). */
int rxn_jacobian_blocks_batch(RxnState *s, const int32_t *l2g, int64_t nlocal, double dt,
                              double *jac_out);

/* both of the above in one launch (the reference evaluates RTResidual and RTJacobian on the
 * same iterate); either output may be NULL. */
int rxn_residual_jacobian_blocks_batch(RxnState *s, const int32_t *l2g, int64_t nlocal, double dt,
                                       double *res_out, double *jac_out);

/* replaces: RTUpdateKineticState loop (reactive_transport.F90:692-705) =
 * RUpdateKineticState per cell (reaction.F90:5320-5429). */
int rxn_update_kinetic_state_batch(RxnState *s, double dt);

/* Device-resident variants of the global-implicit entry points (PETSc VECCUDA / MATAIJCUSPARSE arrays stay
 * on the GPU, no PCIe round trip per Newton iteration).  Pointers are device pointers owned by the caller
 * (rxn_device_alloc or the caller's own allocations on the state's device); same semantics otherwise. */
int rxn_update_auxvars_batch_device(RxnState *s, const double *d_xx_loc, int update_act_coefs);
int rxn_residual_jacobian_blocks_batch_device(RxnState *s, const int32_t *d_l2g, int64_t nlocal, double dt, double *d_res,
                                              double *d_jac);

/* ---- Flux side of the global-implicit transport residual / Jacobian ------------------------------
 * replaces: the interior-connection loops of RTResidualFlux (reactive_transport.F90:2252-2310) and RTJacobianFlux
 * (:3094-3140) with TFluxCoef (transport.F90:756-819), TFlux (:368-439) and TFluxDerivative (:529-622), liquid phase.
 * They consume total / dtotal of the state where rxn_update_auxvars_batch left them (RXN_F_DTOTAL materialised).
 *
 * A connection set is the reference's connection list (grid%internal_connection_set_list flattened in loop order):
 * id_up / id_dn ghosted cell ids (0-based), ghost_to_local = grid%nG2L - 1 (< 0 for ghost cells, NULL = identity),
 * active = imat > 0 per ghosted cell (NULL = all; a connection with an inactive side is skipped).  The library turns it
 * into the row view (per local cell: its connections in connection order), which is also the block-CSR structure of
 * the flux Jacobian: slot row_ptr[r] is the diagonal block of local row r, slots row_ptr[r]+1 .. row_ptr[r+1]-1 its
 * connections in connection order; col = ghosted id of the column cell; blocks n x n column-major, as
 * MatSetValuesBlockedLocal receives Jup / Jdn.  A set is bound to the state it was created on (other states are rejected);
 * destroy it before or after that state, but do not pass it to a call once the state is gone. */
typedef struct RxnConnSet RxnConnSet;
int rxn_connset_create(RxnState *s, int64_t nconn, const int32_t *id_up, const int32_t *id_dn, const int32_t *ghost_to_local,
                       int64_t nlocal, const uint8_t *active, RxnConnSet **out);
int rxn_connset_destroy(RxnConnSet *c);
/* block-CSR structure: *nnz_blocks = row_ptr[nlocal]; row_ptr (nlocal+1) and col (nnz_blocks) may be NULL */
int rxn_connset_structure(const RxnConnSet *c, int64_t *nnz_blocks, int32_t *row_ptr, int32_t *col);
int rxn_connset_device_structure(const RxnConnSet *c, const int32_t **d_row_ptr, const int32_t **d_col);
/* TFluxCoef for every connection, on the device.  Host arrays: area (connection%area), velocity
 * (patch%internal_velocities(1,:)), disp_over_dist (patch%internal_tran_coefs(:,1,:), nconn x naqcomp, component fastest),
 * fraction_upwind (connection%dist(-1,:); may be NULL with use_upwinding).  Call again when the flow field changes. */
int rxn_connset_flux_coefs(RxnConnSet *c, const double *area, const double *velocity, const double *disp_over_dist,
                           const double *fraction_upwind, int use_upwinding);
/* res_out: AoS nlocal x ncomp, r_p after the interior-flux loop (starts from r_p = 0, reactive_transport.F90:2249) */
int rxn_flux_residual_batch(RxnState *s, RxnConnSet *c, double *res_out);
/* val_out: nnz_blocks x ncomp x ncomp, the flux Jacobian in the block-CSR structure above (starts from zero) */
int rxn_flux_jacobian_batch(RxnState *s, RxnConnSet *c, double *val_out);
/* same with the outputs resident on the state's device (PETSc VECCUDA array / MATSEQBAIJ value array on the GPU) */
int rxn_flux_residual_batch_device(RxnState *s, RxnConnSet *c, double *d_res);
int rxn_flux_jacobian_batch_device(RxnState *s, RxnConnSet *c, double *d_val);

/* ---- Boundary conditions and source/sinks ("coupler" connections; SURVEY.md 8f.3 remainder) -------------------------
 * replaces: the boundary-connection loops of RTResidualFlux (reactive_transport.F90:2347-2430) and RTJacobianFlux
 * (:3176-3240), and the source/sink loops of RTResidualNonFlux (:2623-2672) and RTJacobianNonFlux (:3394-3436).
 *
 * A coupler set is the flattened connection list of patch%boundary_condition_list (kind RXN_COUPLER_BOUNDARY) or of
 * patch%source_sink_list (RXN_COUPLER_SRC_SINK) in loop order (sum_connection): id_dn = ghosted id (0-based) of the cell of
 * each connection, ghost_to_local / active as for rxn_connset_create.  Every connection carries an EXTERNAL TOTAL vector:
 * rt_auxvars_bc(sum_connection)%total for a boundary connection - the reference keeps a second auxvar array for the
 * boundary faces (reactive_transport.F90:3851-4030); here that is a second RxnState with one cell per connection, updated
 * with rxn_update_auxvars_batch / rxn_equilibrate_constraint_batch and handed over by rxn_couplerset_totals_from_state -
 * or tran_condition%cur_constraint_coupler%rt_auxvar%total for a source/sink (rxn_couplerset_set_totals).
 *   boundary   : Res = coef_up total_ext + coef_dn total_cell ; r_p -= Res ; diagonal block -= dtotal_cell(i,:) coef_dn(i)
 *   source/sink: Res = coef_in total_cell + coef_out total_ext ; r_p += Res ; diagonal block += coef_in dtotal_cell
 * The residual / Jacobian calls ADD onto the arrays the interior-flux (or accumulation) calls produced, in connection
 * order, as the reference's loops do after the interior loop. */
typedef struct RxnCouplerSet RxnCouplerSet;
enum { RXN_COUPLER_BOUNDARY = 0, RXN_COUPLER_SRC_SINK = 1 };
/* tran_condition%itype values TSrcSinkCoef distinguishes (pflotran_constants.F90:139,144); anything else: volumetric rate */
enum { RXN_SS_MASS_RATE = 7, RXN_SS_EQUILIBRIUM = 12 };
int rxn_couplerset_create(RxnState *s, int kind, int64_t nconn, const int32_t *id_dn, const int32_t *ghost_to_local, int64_t nlocal,
                          const uint8_t *active, RxnCouplerSet **out);
int rxn_couplerset_destroy(RxnCouplerSet *b);
/* boundary: TFluxCoef (transport.F90:756-819) with fraction_upwind = 0.5 as reactive_transport.F90:2369-2373.  Host arrays:
 * area (connection%area), velocity (patch%boundary_velocities(1,:)), disp_over_dist (patch%boundary_tran_coefs(:,1,:),
 * nconn x naqcomp, component fastest) */
int rxn_couplerset_bc_coefs(RxnCouplerSet *b, const double *area, const double *velocity, const double *disp_over_dist,
                            int use_upwinding);
/* source/sink: TSrcSinkCoef (transport.F90:901-954); qsrc = patch%ss_flow_vol_fluxes(1,:), type = tran_condition%itype */
int rxn_couplerset_ss_coefs(RxnCouplerSet *b, const double *qsrc, const int32_t *tran_src_sink_type);
/* external totals: host array nconn x naqcomp (component fastest), or the TOTAL field of a state with >= nconn cells on the
 * same device (cell c of that state = connection c) */
int rxn_couplerset_set_totals(RxnCouplerSet *b, const double *total);
int rxn_couplerset_totals_from_state(RxnCouplerSet *b, const RxnState *bc_state);
/* res_inout: AoS nlocal x ncomp (r_p); flux_out: optional nconn x ncomp (patch%boundary_tran_fluxes = -Res resp.
 * patch%ss_tran_fluxes = Res; rows of connections on inactive cells are left untouched) */
int rxn_coupler_residual_batch(RxnState *s, RxnCouplerSet *b, double *res_inout, double *flux_out);
/* Jacobian: adds into the diagonal blocks.  With a connection set: val_inout = its block-CSR value array (nnz_blocks x
 * ncomp x ncomp; the diagonal block of local row r is slot row_ptr[r]); with c = NULL: val_inout = nlocal x ncomp x ncomp
 * diagonal blocks (the layout of rxn_jacobian_blocks_batch). */
int rxn_coupler_jacobian_batch(RxnState *s, RxnConnSet *c, RxnCouplerSet *b, double *val_inout);
/* same on arrays resident on the state's device */
int rxn_coupler_residual_batch_device(RxnState *s, RxnCouplerSet *b, double *d_res, double *d_flux_out);
int rxn_coupler_jacobian_batch_device(RxnState *s, RxnConnSet *c, RxnCouplerSet *b, double *d_val);

/* timing of the last batched kernel sequence on the handle's stream, in ms (CUDA events). */
float rxn_last_kernel_ms(const RxnState *s);
/* CUDA-event bracket on the handle's stream around any sequence of calls (bench.py) */
int rxn_timer_start(RxnState *s);
int rxn_timer_stop(RxnState *s, float *ms);
/* measured FP64 FMA throughput of the handle's device in TFLOP/s (2 flop per DFMA): the
 * roofline denominator for this FP64 CUDA-core path, which MEASURED_PEAKS.json does not hold */
int rxn_probe_fp64(RxnState *s, double *tflops);
/* number of kernels this library has launched since load (bench.py: gpu_launches). */
int64_t rxn_launch_count(void);

/* raw device pointer of a field (row-major [rows][ld]); for zero-copy callers and benchmarks */
int rxn_state_device_ptr(RxnState *s, int field, double **d_ptr, int64_t *ld);
int rxn_device_alloc(RxnState *s, int64_t bytes, void **d_ptr);
int rxn_device_free(RxnState *s, void *d_ptr);
int rxn_device_copy(RxnState *s, void *dst, const void *src, int64_t bytes, int kind /*0 h2d,1 d2h,2 d2d*/);
int rxn_device_sync(RxnState *s);
/* pinned host memory for the caller's tran_xx / result buffers (optional, faster PCIe) */
int rxn_host_alloc(int64_t bytes, void **h_ptr);
int rxn_host_free(void *h_ptr);

/* replaces: option%io_buffer */
int rxn_last_error(char *buf, int32_t len);
const char *rxn_version(void);

#ifdef __cplusplus
}
#endif
#endif /* RXN_B200_H */
