/* vfs_b200.h — C ABI of the B200-native momentum RHS + LES path of VFS-Wind.
 *
 * Drop-in boundary.  Every entry point replaces one reference function;
 * citations are relative to the reference tree (Source/...):
 *
 *   vfs_form_metrics        <- FormMetrics(UserCtx*)                         metrics.c:12
 *   vfs_contra2cart         <- Contra2Cart(UserCtx*) / Contra2Cart_2         rhs.c:35,65
 *   vfs_ib_bc               <- IB_BC(UserCtx*)                               momentum.c:2016
 *   vfs_formfunction2       <- Formfunction_2(UserCtx*, Vec Rhs, double)     momentum.c:454
 *   vfs_formfunction_snes   <- FormFunction_SNES(SNES, Vec X, Vec F, void*)  momentum.c:2237
 *   vfs_les_cs              <- Compute_Smagorinsky_Constant_1(UserCtx*,Vec,Vec)  les.c:75
 *   vfs_les_nut             <- Compute_eddy_viscosity_LES(UserCtx*)          les.c:1143
 *   vfs_halo_exchange       <- DAGlobalToLocal / DALocalToLocal (k direction, between ranks)
 *   vfs_pressure_gradient   <- Pressure_Gradient(UserCtx*, Vec dP)           momentum.c:203
 *   vfs_update_pressure     <- UpdatePressure(UserCtx*)                      poisson.c:3137
 *   vfs_projection          <- Projection(UserCtx*)                          poisson.c:2700
 *   vfs_calc_f_eul / vfs_calc_u_lagr <- Calc_F_eul / Calc_U_lagr            rotor_model.c:3668,2937
 *   vfs_convection / vfs_viscous <- Convection / Viscous (legacy explicit path)   rhs.c:751,1071
 *   vfs_cylinder_forces     <- the lA_cyl / lFp*_cyl / lFv*_cyl sums inside Formfunction_2   momentum.c:570-579,822-849
 *   vfs_momentum_solve      <- SNESSolve in Implicit_MatrixFree              implicitsolver.c:4203-4299
 *
 * Plain C, POD only, no torch / PETSc types.  All numerics are FP64.  One vfs_ctx per GPU / rank;
 * the context owns every device buffer; host arrays are borrowed for the duration of a call.
 * Host arrays use the reference's DA "global Vec" layout restricted to this rank's k-slab:
 * [nzl][my][mx][dof] doubles, i fastest, dof interleaved (Cmpnts{x,y,z}, variables.h:44).
 *
 * There is NO CPU fallback: every compute entry point launches sm_100a kernels and returns
 * VFS_ERR_CUDA if no device is usable.
 */
#ifndef VFS_B200_H
#define VFS_B200_H
#ifdef __cplusplus
extern "C" {
#endif

#define VFS_OK 0
#define VFS_ERR_ARG (-1)
#define VFS_ERR_CUDA (-2)
#define VFS_ERR_UNSUPPORTED (-3)
#define VFS_ERR_HALO (-4)

/* Fields a caller can upload / download / exchange.  dof 3 fields are Cmpnts. */
enum vfs_field {
  VFS_COOR = 0,     /* node coordinates (dof 3), input of vfs_form_metrics          */
  VFS_CSI, VFS_ETA, VFS_ZET, /* centre metrics lCsi/lEta/lZet (dof 3)                */
  VFS_AJ,           /* lAj  (dof 1)                                                  */
  VFS_NVERT,        /* lNvert (dof 1; 0 fluid, 1 IB node, 3 solid)                   */
  VFS_UCONT,        /* lUcont  contravariant fluxes (dof 3)                          */
  VFS_UCAT,         /* Ucat / lUcat Cartesian velocity (dof 3), persistent in/out    */
  VFS_UCAT_OLD,     /* lUcat_old (dof 3), slip-wall ghost rule rhs.c:466             */
  VFS_UCONT_O,      /* Ucont_o (dof 3)                                               */
  VFS_UCONT_RM1,    /* Ucont_rm1 (dof 3), BDF2 only                                  */
  VFS_RHS_O,        /* RHS_o (dof 3)                                                 */
  VFS_DP,           /* dP (dof 3)                                                    */
  VFS_F_EUL,        /* F_eul (dof 3) actuator forcing                                */
  VFS_RHS,          /* Rhs (dof 3) output                                            */
  VFS_CS,           /* lCs (dof 1) output of vfs_les_cs                              */
  VFS_NU_T,         /* lNu_t (dof 1) output of vfs_les_nut                           */
  VFS_USTAR,        /* lUstar (dof 1) wall-model friction velocity                   */
  VFS_CONV,         /* Conv (dof 3) output of vfs_convection (download only)           */
  VFS_VISC,         /* Visc (dof 3) output of vfs_viscous (download only)              */
  VFS_P,            /* P / lP (dof 1) pressure, input of vfs_pressure_gradient         */
  VFS_PHI,          /* Phi / lPhi (dof 1) pressure correction from the Poisson solver, input of vfs_update_pressure / vfs_projection */
  VFS_NFIELDS_PUBLIC
};

typedef struct vfs_params {
  int mx, my, mz;        /* DA node counts (IM+1, JM+1, KM+1), init.c:157-160                */
  int kofs, nzl;         /* this rank owns global planes k in [kofs, kofs+nzl)               */
  int rank, nranks;      /* k-slab decomposition                                             */
  int device;            /* CUDA device ordinal                                              */
  int ii_periodic, jj_periodic, kk_periodic; /* DA-wrap periodicity (main.c:284-288)         */
  int bctype[6];         /* bcs.dat (init.c:503-510)                                         */
  int les;               /* 0 off, 1 constant Cs, 2 dynamic                                  */
  int second_order, laplacian, immersed, clark, central;
  int testfilter_ik;
  int viscosity_wallmodel, wallfunction;
  int rotor_model, nacelle_model, IB_delta;   /* any non-zero => Rhs += F_eul (momentum.c:2329) */
  int ti, tistart, rstart_flg;
  int levelset, rans, inviscid, skew, movefsi, rotatefsi; /* levelset, rans, movefsi, rotatefsi must be 0 (out of scope);
                          * inviscid: WENO3 convection, no viscous term (momentum.c:754,1624,1663); skew: skew-symmetric
                          * convection (:789-800,1638-1651); clark (above): mixed model in the viscous flux (:904-923)
                          * and in the dynamic procedure (les.c:497-556,656).  These three run the one-thread-per-face
                          * kernels instead of the marching ones.                                                    */
  int i_periodic, j_periodic, k_periodic;     /* legacy periodicity (explicit index remaps on a non-periodic DA): same images as
                                               * ii/jj/kk, one difference in IB_BC (momentum.c:2206-2211); k_periodic single rank only */
  int i_homo_filter, j_homo_filter, k_homo_filter; /* Cs from LM, MM averaged over homogeneous directions, les.c:798-965 */
  double ren, dt, max_cs;
  double roughness_size; /* -roughness (main.c:319,1734): k_s of the rough-wall log law, bctype -2            */
  /* levelset_weno: 0, or 5 = WENO3 convection everywhere (momentum.c:754,1015,1301 test it even without levelset);
   * 1-4 key on the level-set field and are rejected, as are freesurface_wallmodel / air_flow_levelset (they change the
   * wall model and the boundary fluxes).  `central` only matters together with rans, which is rejected. */
  int levelset_weno, freesurface_wallmodel, air_flow_levelset;
} vfs_params;

typedef struct vfs_ctx vfs_ctx;

/* k-halo callback.  Called (only when nranks > 1) whenever the path needs the k-ghost planes of
 * `nfields` scalar planes refreshed from the neighbouring ranks.  ids are INTERNAL scalar ids;
 * use vfs_scalar_ptr() to obtain each scalar's device base pointer and vfs_layout() for the
 * plane geometry.  Must return 0 on success, after the exchange has been enqueued on (or
 * synchronised with) the context's stream. */
typedef int (*vfs_halo_fn)(void *user, int nfields, const int *scalar_ids);

int vfs_create(const vfs_params *p, vfs_ctx **out);
/* helpers for host glue that does not link the CUDA runtime itself: page-locked staging memory, visible devices */
void *vfs_host_alloc(unsigned long bytes);
void vfs_host_free(void *p);
int vfs_device_count(void);
int vfs_destroy(vfs_ctx *c);
const char *vfs_last_error(vfs_ctx *c);          /* c may be NULL: last create() error      */
int vfs_set_params(vfs_ctx *c, const vfs_params *p); /* update run-time switches (ti, dt, ...) */
int vfs_set_stream(vfs_ctx *c, void *cuda_stream);   /* run on a caller stream (cudaStream_t) */
int vfs_set_halo_callback(vfs_ctx *c, vfs_halo_fn fn, void *user);
/* In-library halo layer (the default on GPUs): NCCL point-to-point between k-neighbours, issued on the
 * context's stream straight from / into the padded arrays (replaces DAGlobalToLocal/DALocalToLocal
 * between ranks, init.c:131-160).  Rank 0 obtains a 128-byte ncclUniqueId with vfs_nccl_unique_id and
 * distributes it by any means (MPI_Bcast in the reference's host code, torch.distributed in bench.py);
 * every rank then calls vfs_nccl_init (collective).  When set, it takes precedence over the callback. */
int vfs_nccl_unique_id(char *out128);
int vfs_nccl_init(vfs_ctx *c, const char *id128);
/* number of k-halo exchanges performed by this context (and bytes sent, if bytes != NULL) */
long vfs_halo_count(vfs_ctx *c, long *bytes);
/* inside a halo callback: how many ghost planes below (lo) / above (hi) the slab the exchange in progress must fill
 * (the fields are read no deeper before their next refresh; G, the full ghost width, is always valid) */
int vfs_halo_layers(vfs_ctx *c, int *lo, int *hi);
int vfs_sync(vfs_ctx *c);

/* layout[0..7] = G (ghost width), pitch, ny (=my+2G), nzt (=nzl+2G), plane doubles (=ny*pitch),
 * scalar doubles (=nzt*plane), number of internal scalars, dof-offset convention (=1). */
int vfs_layout(vfs_ctx *c, long *layout8);
int vfs_field_scalar_id(vfs_ctx *c, int field, int comp);  /* public field -> internal scalar  */
void *vfs_scalar_ptr(vfs_ctx *c, int scalar_id);           /* device pointer of padded scalar  */

/* host <-> device, host layout [nzl][my][mx][dof]; `host` may also be a device pointer (unified addressing) */
int vfs_upload(vfs_ctx *c, int field, const double *host);
int vfs_download(vfs_ctx *c, int field, double *host);
/* refresh ghosts (i/j wrap + k halo) of one public field, as DAGlobalToLocal would */
int vfs_halo_exchange(vfs_ctx *c, int field);

int vfs_form_metrics(vfs_ctx *c);
int vfs_contra2cart(vfs_ctx *c);
int vfs_ib_bc(vfs_ctx *c);
int vfs_les_cs(vfs_ctx *c);
int vfs_les_nut(vfs_ctx *c);
/* Rhs (device field `rhs_field`, VFS_RHS or VFS_RHS_O) += scale * R(Ucont, Ucat) */
int vfs_formfunction2(vfs_ctx *c, int rhs_field, double scale);
/* Asynchronous download: the field is packed on the context's stream and copied to `host` (pinned memory for a
 * truly asynchronous copy) on a separate copy stream, so the transfer overlaps whatever is queued next on the
 * context (e.g. the host->device copy and kernels of vfs_formfunction_snes: PCIe is full duplex).  `host` is valid
 * after vfs_download_wait (or vfs_destroy).  At most one scalar (dof 1) or vector field per slot, slot = 0 or 1. */
int vfs_download_async(vfs_ctx *c, int field, double *host, int slot);
int vfs_download_wait(vfs_ctx *c);
/* Legacy explicit-solver pieces, Convection(UserCtx*,Vec Ucont,Vec Ucat,Vec Conv) Source/rhs.c:751 and
 * Viscous(UserCtx*,Vec,Vec,Vec Visc) Source/rhs.c:1071 (callers timeadvancing1.c:75-76): QUICK
 * flux-difference convection and the (nu + nu_t) viscous term of the current VFS_UCONT / VFS_UCAT
 * (+ VFS_NU_T when les) into VFS_CONV / VFS_VISC.  Like the reference they ignore periodicity. */
/* Pressure_Gradient(UserCtx*, Vec dP) Source/momentum.c:203-439: dP (VFS_DP) = the contravariant
 * pressure gradient of VFS_P on the i/j/k faces, IB-aware one-sided differences, zero on masked faces; P's ghosts and
 * periodic boundary nodes are refreshed first, as the reference does (:247-286).  `k_forcing` is the scalar the
 * reference adds to dP/dzeta per unit dz in k-periodic runs (:399-411): mean_pressure_gradient when dpdz_set, else
 * (mean_k_flux - inlet_flux) / dt / mean_k_area unless inletprofile == 17; pass 0 otherwise. */
int vfs_pressure_gradient(vfs_ctx *c, double k_forcing);
/* UpdatePressure(UserCtx*) Source/poisson.c:3137-3296 and Projection(UserCtx*) Source/poisson.c:2700-3052, called back to
 * back after the Poisson solve (solvers.c:662-663).  VFS_PHI = the solver's pressure correction (owned
 * values; ghosts are refreshed here).  vfs_update_pressure: P += Phi at fluid cells, 0 at solid cells, periodic boundary
 * nodes of P and Phi <- their images, ghosts of both refreshed.  vfs_projection: VFS_UCONT -= dt * st * (contravariant
 * gradient of Phi) on every unmasked face (nvert sum < poisson_threshold; one-sided / zero tangential differences next to
 * masked cells and non-periodic ends) and the component-wise periodic copies.  The reference's Projection ends with
 * Contra2Cart (:3049): call vfs_contra2cart next.  `st` = user->st; time_coeff() = 1 and the density factors are 1 on
 * this path (no levelset / RANS). */
int vfs_update_pressure(vfs_ctx *c);
int vfs_projection(vfs_ctx *c, double st, double poisson_threshold);
/* The body-fitted-cylinder diagnostics Formfunction_2 accumulates when bctype[0] == 11 and bctype[1] == 1
 * (Source/momentum.c:570-579, 822-849): out7 = this rank's lA_cyl, lA_cyl_x, lA_cyl_z, lFpx_cyl, lFpz_cyl, lFvx_cyl,
 * lFvz_cyl over the wall faces i = mx-2, from the current VFS_UCAT (as the last residual evaluation left it) and VFS_P.
 * All zero for other boundary types.  The reference adds them over ranks itself (main.c:1269-1277). */
int vfs_cylinder_forces(vfs_ctx *c, double *out7);
int vfs_convection(vfs_ctx *c);
int vfs_viscous(vfs_ctx *c);
/* F = residual(X); X, F host arrays [nzl][my][mx][3] (pinned or pageable) */
int vfs_formfunction_snes(vfs_ctx *c, const double *x_host, double *f_host);
/* same with X already in VFS_UCONT and F left in VFS_RHS (device resident Krylov vectors) */
int vfs_formfunction_snes_dev(vfs_ctx *c);
/* one cell-update unit: vfs_contra2cart + vfs_les_cs + vfs_les_nut + vfs_formfunction_snes_dev */
int vfs_rhs_les_fused(vfs_ctx *c);

/* Device-resident implicit momentum solve: replaces the SNESSolve of Implicit_MatrixFree
 * (Source/implicitsolver.c:4203-4299) — SNES trust region + Eisenstat-Walker v3, matrix-free Jacobian by forward
 * differences of FormFunction_SNES, restarted GMRES without preconditioner.  In/out: VFS_UCONT on the device
 * (U = Ucont before, Ucont = U after, :4297,4307); the Krylov basis and all work vectors stay in HBM, dot products
 * are summed over the ranks with ncclAllReduce (needs vfs_nccl_init when nranks > 1).  PETSc 3.1, where the
 * reference's solver lives, is not part of the reference tree: its published algorithms are restated
 * (vfs-wind_b200/csrc/vfs_solver.h). */
typedef struct vfs_solver_params {
  int max_newton;                       /* SNESSetTolerances maxit = 50            implicitsolver.c:4254 */
  int restart;                          /* KSPGMRES restart (PETSc default 30)                            */
  int max_krylov;                       /* KSPSetTolerances maxits = 1000          implicitsolver.c:4279 */
  double snes_atol, snes_rtol, snes_stol;   /* SNESSetTolerances(PETSC_DEFAULT, imp_free_tol, ...) :4254 */
  double ksp_rtol, ksp_atol, ksp_dtol;  /* KSPSetTolerances(imp_free_tol, ...) :4277; rtol superseded by Eisenstat-Walker */
  int use_ew;                           /* SNESKSPSetUseEW + version 3             implicitsolver.c:4251-4252 */
  int trust_region;                     /* 1 = SNESTR (:4247); 0 = full Newton steps                       */
} vfs_solver_params;
typedef struct vfs_solver_info {
  int newton_iterations, krylov_iterations, residual_evals;
  int reason;                           /* SNESConvergedReason numbering: 2 atol, 3 rtol, 4 step, 7 tr delta, < 0 diverged */
  double fnorm0, fnorm, xnorm, delta;
  int n_history;                        /* entries of fnorm_history (|F| before the first and after every Newton step) */
  double fnorm_history[17];
  int ksp_its_history[16];
} vfs_solver_info;
int vfs_solver_defaults(vfs_solver_params *p);
int vfs_momentum_solve(vfs_ctx *c, const vfs_solver_params *p, vfs_solver_info *info);
/* free the Krylov basis and work vectors (restart + 8 vectors of nzl*my*mx*3 doubles, kept between solves) */
int vfs_momentum_release(vfs_ctx *c);

/* Actuator forcing.  vfs_actuator = the IBMNodes fields (Source/variables.h:159,188,192) the two
 * functions read, one per turbine / nacelle object; index windows are GLOBAL node indices, upper bounds exclusive.
 *   vfs_calc_f_eul  <- Calc_F_eul(UserCtx*, IBMNodes*, FSInfo*, n, dh, df)   Source/rotor_model.c:3668-3960
 *       VFS_F_EUL (+)= the elements' forces spread with the delta function `df` (0: 2h hat, 7: Gaussian of half width
 *       halfwidth_dfunc, else the smoothed 4h function), then zeroed near solid cells and domain boundaries and its
 *       ghosts refreshed.  accumulate = 0 zeroes F_eul first (VecSet(lF_eul, 0), solvers.c:484).
 *   vfs_calc_u_lagr <- Calc_U_lagr(UserCtx*, IBMNodes*, FSInfo*, n)          Source/rotor_model.c:2937-3031
 *       U_lagr_{x,y,z}[l] = sum over the element's window of VFS_UCAT * dfunc_s3h^3, summed over the ranks
 *       (MPI_Allreduce there, ncclAllReduce here); the periodic-turbine bookkeeping that follows (:3034-3150) is
 *       host logic and stays with the caller. */
typedef struct vfs_actuator {
  int n_elmt;
  const double *cent_x, *cent_y, *cent_z, *dA;
  const double *F_lagr_x, *F_lagr_y, *F_lagr_z;
  double *U_lagr_x, *U_lagr_y, *U_lagr_z;
  const int *i_min, *i_max, *j_min, *j_max, *k_min, *k_max;
} vfs_actuator;
int vfs_calc_f_eul(vfs_ctx *c, int n_objects, const vfs_actuator *objs, int df, double halfwidth_dfunc, int forcewidthfixed, const double *dh_fixed3, int accumulate);
int vfs_calc_u_lagr(vfs_ctx *c, int n_objects, const vfs_actuator *objs);

/* number of kernels launched by this context since creation (bench `gpu_launches`) */
long vfs_launch_count(vfs_ctx *c);
/* CUDA-event timing (ms, on the context's stream) of the most recent execution of a kernel
 * group; 0 if it has not run.  Valid after vfs_sync / any synchronous entry point. */
enum vfs_timer { VFS_T_TOTAL = 0, VFS_T_C2C, VFS_T_FLUX, VFS_T_FP, VFS_T_PROJECT, VFS_T_LES1, VFS_T_LES2, VFS_T_LES3, VFS_T_NUT, VFS_T_COUNT };
double vfs_last_ms(vfs_ctx *c, int which);
/* tuning switches (defaults are the measured-best settings; every setting gives the same fields to rounding):
 *   0  kernel family: 1 = TMA-staged marching kernels (default), 0 = one-thread-per-cell staged kernels (the literal
 *      restatement), 2 = experimental fully fused residual kernel
 *   1  replay vfs_rhs_les_fused / vfs_formfunction_snes_dev as a CUDA graph (default 0; per-kernel timers are not
 *      updated while replaying; with nranks > 1 only together with vfs_nccl_init)
 *   2  LES pass-2 tile height 16 / 12 (default) / 8        3  resident blocks per SM of the one-ring flux kernel
 *   4  LES pass-1 form: 1 = TMA tile march (default), 0 / 2 / 3 = block-program variants
 *   6  mask-free fast paths for warps far from any nvert != 0 (default 1)
 *   7  face-flux kernel: 1 = two TMA rings, ucat and metric planes (default), 0 = one ring
 *   8  single-rank ghost-refresh sequences as one launch (default 1)
 *   9  overlap the k-face-flux and Fp halo exchanges with interior planes (default 1; nranks > 1 with vfs_nccl_init)
 *  11  asynchronous compute-only entry points (vfs_contra2cart, vfs_ib_bc, vfs_les_cs, vfs_les_nut return once their work
 *      is queued; default 0): a following vfs_formfunction_snes then copies X while those kernels still run
 *  15  thread-block shape of the one-thread-per-node kernels (tuning)    16  LES pass 1 replayed on ghost planes between ranks (default 1)
 *  17  FpCell on pairs of cells with 16-byte loads (default 0: measured slower than 8-byte loads, 0.82 vs 0.70 ms)
 *  14  exchange only the ghost layers each refresh is read at (default 1; 0 = always the full ghost width G)
 *  12  Fp evaluated inside the projection kernel instead of FpCell + Fp planes (default 0: bitwise equal, measured slower)
 *  18  vfs_rhs_les_fused, single rank: the residual's Contra2Cart + IB_BC on a second stream beside LES pass 3 / nu_t
 *      (default 0: bitwise equal, measured gain 0.3 %)
 *  19  resident blocks per SM the LES pass-3 kernel is compiled for: 4 (default, 32 registers), 3 (40) or 2 (48 registers)
 *  20  resident 256-thread blocks per SM the projection kernel is compiled for: 6 (default), 8, or 0 = no cap
 *  21  vfs_rhs_les_fused: nu_t written by LES pass 3 instead of a separate pass (default 1; bitwise equal) */
int vfs_set_option(vfs_ctx *c, int key, int value);

#ifdef __cplusplus
}
#endif
#endif
