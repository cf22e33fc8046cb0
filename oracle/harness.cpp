// Driver glue around the UNMODIFIED reference sources (compiled from /root/reference/Source by
// oracle/build_ref.py) so Python tests can run the reference's own hot-path functions on explicit
// arrays.  TEST INFRASTRUCTURE ONLY: the checker for parity tests and the "reference" CPU
// baseline; never linked or imported by the product path.
//
// Mirrors what MG_Initial (Source/init.c:212, Vec allocation at :591-804) and Flow_Solver
// (Source/solvers.c:430-436) set up before the hot path is entered.
#include "variables.h"

extern "C" DA shim_da_create(int, int, int, int, int, int, int);
extern "C" void shim_da_set_cda(DA, DA);
extern "C" double *shim_vec_data(Vec);
extern "C" long shim_vec_size(Vec);
extern "C" int shim_vec_is_local(Vec);
extern "C" int shim_vec_dof(Vec);

// poisson.c:23-31 is the only thing the hot path needs from poisson.c (which needs HYPRE).
#ifndef VFS_REF_HAS_POISSON      /* the reference build links poisson.c itself (Projection, UpdatePressure); the glue builds do not */
double time_coeff() { if (levelset || rans) { if (ti == tistart) return 1.; else return 1.5; } else return 1.; }
#endif

PetscErrorCode FormFunction_SNES(SNES snes, Vec Ucont, Vec Rhs, void *ptr);

#define GLOBAL3(X) X(Ucont) X(Ucat) X(Ucont_o) X(Ucont_rm1) X(RHS_o) X(dP) X(F_eul) X(Cent) X(GridSpace) X(Rhs)
#define GLOBAL1(X) X(Nvert) X(P) X(Phi)
#define LOCAL3(X) X(lCsi) X(lEta) X(lZet) X(lICsi) X(lIEta) X(lIZet) X(lJCsi) X(lJEta) X(lJZet) X(lKCsi) X(lKEta) X(lKZet) \
  X(lGridSpace) X(lCent) X(lUcont) X(lUcat) X(lUcat_old) X(lUcont_o) X(lUcont_rm1) X(Fp) X(Div1) X(Div2) X(Div3) X(Visc1) X(Visc2) X(Visc3) X(lF_eul)
#define LOCAL1(X) X(lAj) X(lIAj) X(lJAj) X(lKAj) X(lP) X(lPhi) X(lNvert) X(lNvert_o) X(lNu_t) X(lCs) X(lUstar)

extern "C" {

UserCtx *ref_create(int mx, int my, int mz) {
  UserCtx *u = (UserCtx *)calloc(1, sizeof(UserCtx));
  u->da = shim_da_create(mx, my, mz, 1, ii_periodic, jj_periodic, kk_periodic);
  u->fda = shim_da_create(mx, my, mz, 3, ii_periodic, jj_periodic, kk_periodic);
  u->fda2 = shim_da_create(mx, my, mz, 2, ii_periodic, jj_periodic, kk_periodic);
  shim_da_set_cda(u->da, u->fda); shim_da_set_cda(u->fda, u->fda);
  DAGetLocalInfo(u->da, &u->info);
  u->IM = mx - 1; u->JM = my - 1; u->KM = mz - 1;
#define X(n) DACreateGlobalVector(u->fda, &u->n);
  GLOBAL3(X)
#undef X
#define X(n) DACreateGlobalVector(u->da, &u->n);
  GLOBAL1(X)
#undef X
#define X(n) DACreateLocalVector(u->fda, &u->n);
  LOCAL3(X)
#undef X
#define X(n) DACreateLocalVector(u->da, &u->n);
  LOCAL1(X)
#undef X
  return u;
}

Vec ref_vec(UserCtx *u, const char *name) {
#define X(n) if (!strcmp(name, #n)) return u->n;
  GLOBAL3(X) GLOBAL1(X) LOCAL3(X) LOCAL1(X)
#undef X
  if (!strcmp(name, "coords")) { Vec c; DAGetGhostedCoordinates(u->da, &c); return c; }
  return 0;
}
double *ref_vec_data(Vec v) { return shim_vec_data(v); }
long ref_vec_size(Vec v) { return shim_vec_size(v); }
int ref_vec_is_local(Vec v) { return shim_vec_is_local(v); }
int ref_vec_dof(Vec v) { return shim_vec_dof(v); }
void ref_local_info(UserCtx *u, int *out) { DALocalInfo i; DAGetLocalInfo(u->da, &i);
  out[0] = i.gxs; out[1] = i.gys; out[2] = i.gzs; out[3] = i.gxm; out[4] = i.gym; out[5] = i.gzm; }
void ref_set_scalars(UserCtx *u, double ren, double dt, const int *bctype) {
  u->ren = ren; u->dt = dt; for (int q = 0; q < 6; q++) u->bctype[q] = bctype[q]; }
void ref_global_to_local(UserCtx *u, const char *g, const char *l) {
  Vec G = ref_vec(u, g), L = ref_vec(u, l); DA d = shim_vec_dof(G) == 3 ? u->fda : u->da;
  DAGlobalToLocalBegin(d, G, INSERT_VALUES, L); DAGlobalToLocalEnd(d, G, INSERT_VALUES, L); }

int ref_FormMetrics(UserCtx *u) {
  // FormMetrics destroys its global work vectors on exit (metrics.c:1065-1084): recreate each time.
  DACreateGlobalVector(u->fda, &u->Csi); DACreateGlobalVector(u->fda, &u->Eta); DACreateGlobalVector(u->fda, &u->Zet);
  DACreateGlobalVector(u->fda, &u->ICsi); DACreateGlobalVector(u->fda, &u->IEta); DACreateGlobalVector(u->fda, &u->IZet);
  DACreateGlobalVector(u->fda, &u->JCsi); DACreateGlobalVector(u->fda, &u->JEta); DACreateGlobalVector(u->fda, &u->JZet);
  DACreateGlobalVector(u->fda, &u->KCsi); DACreateGlobalVector(u->fda, &u->KEta); DACreateGlobalVector(u->fda, &u->KZet);
  DACreateGlobalVector(u->da, &u->Aj); DACreateGlobalVector(u->da, &u->IAj);
  DACreateGlobalVector(u->da, &u->JAj); DACreateGlobalVector(u->da, &u->KAj);
  return FormMetrics(u);
}
void ref_Contra2Cart(UserCtx *u) { Contra2Cart(u); }
void ref_IB_BC(UserCtx *u) { IB_BC(u); }
int ref_Formfunction_2(UserCtx *u, Vec rhs, double scale) { return Formfunction_2(u, rhs, scale); }
int ref_FormFunction_SNES(UserCtx *u, Vec x, Vec f) { return FormFunction_SNES((SNES)0, x, f, (void *)u); }
void ref_Compute_Smagorinsky_Constant_1(UserCtx *u) { Compute_Smagorinsky_Constant_1(u, u->lUcont, u->lUcat); }
void ref_Compute_eddy_viscosity_LES(UserCtx *u) { Compute_eddy_viscosity_LES(u); }
void ref_Pressure_Gradient(UserCtx *u, Vec dp, double mean_k_flux, double mean_k_area) { u->mean_k_flux = mean_k_flux; u->mean_k_area = mean_k_area; Pressure_Gradient(u, dp); }
// actuator forcing (rotor_model.c:3668, 2937): IBMNodes with just the arrays the two functions read
IBMNodes *ref_actuator_new(int n_elmt) {
  IBMNodes *b = (IBMNodes *)calloc(1, sizeof(IBMNodes));
  b->n_elmt = n_elmt;
  double **d[] = {&b->cent_x, &b->cent_y, &b->cent_z, &b->dA, &b->F_lagr_x, &b->F_lagr_y, &b->F_lagr_z, &b->U_lagr_x, &b->U_lagr_y, &b->U_lagr_z};
  for (unsigned q = 0; q < sizeof(d) / sizeof(d[0]); q++) *d[q] = (double *)calloc(n_elmt, sizeof(double));
  int **w[] = {&b->i_min, &b->i_max, &b->j_min, &b->j_max, &b->k_min, &b->k_max};
  for (unsigned q = 0; q < 6; q++) *w[q] = (int *)calloc(n_elmt, sizeof(int));
  return b;
}
double *ref_actuator_d(IBMNodes *b, int which) {
  double *d[] = {b->cent_x, b->cent_y, b->cent_z, b->dA, b->F_lagr_x, b->F_lagr_y, b->F_lagr_z, b->U_lagr_x, b->U_lagr_y, b->U_lagr_z};
  return d[which];
}
int *ref_actuator_i(IBMNodes *b, int which) { int *w[] = {b->i_min, b->i_max, b->j_min, b->j_max, b->k_min, b->k_max}; return w[which]; }
int ref_Calc_F_eul(UserCtx *u, IBMNodes *b, int df) { FSInfo *f = (FSInfo *)calloc(1, sizeof(FSInfo)); int r = Calc_F_eul(u, b, f, 1, 1.0, df); free(f); return r; }
int ref_Calc_U_lagr(UserCtx *u, IBMNodes *b) { FSInfo *f = (FSInfo *)calloc(1, sizeof(FSInfo)); int r = Calc_U_lagr(u, b, f, 1); free(f); return r; }
// pressure update and projection step after the Poisson solve (poisson.c:3137, 2700; solvers.c:662-663)
int ref_UpdatePressure(UserCtx *u) { return UpdatePressure(u); }
int ref_Projection(UserCtx *u, double st) { u->st = st; return Projection(u); }
// body-fitted-cylinder diagnostics Formfunction_2 leaves in the context (momentum.c:570-579, 822-849)
void ref_cylinder_forces(UserCtx *u, double *out7) {
  out7[0] = u->lA_cyl; out7[1] = u->lA_cyl_x; out7[2] = u->lA_cyl_z; out7[3] = u->lFpx_cyl; out7[4] = u->lFpz_cyl; out7[5] = u->lFvx_cyl; out7[6] = u->lFvz_cyl;
}
// several objects at once (a turbine array): bodies[] from ref_actuator_new, centres xyz_c[3 * nobj] -> FSInfo.x_c/y_c/z_c
int ref_Calc_U_lagr_multi(UserCtx *u, IBMNodes **bodies, int nobj, const double *xyz_c) {
  IBMNodes *ibm = (IBMNodes *)calloc(nobj, sizeof(IBMNodes));
  FSInfo *f = (FSInfo *)calloc(nobj, sizeof(FSInfo));
  for (int b = 0; b < nobj; b++) { ibm[b] = *bodies[b]; f[b].x_c = xyz_c[3 * b]; f[b].y_c = xyz_c[3 * b + 1]; f[b].z_c = xyz_c[3 * b + 2]; }
  int r = Calc_U_lagr(u, ibm, f, nobj);
  free(ibm); free(f);
  return r;
}
int ref_Convection(UserCtx *u, Vec conv) { return Convection(u, u->lUcont, u->lUcat, conv); }
int ref_Viscous(UserCtx *u, Vec visc) { return Viscous(u, u->lUcont, u->lUcat, visc); }
Vec ref_vec_new(UserCtx *u, int dof, int local) { Vec v; DA d = dof == 3 ? u->fda : u->da;
  if (local) DACreateLocalVector(d, &v); else DACreateGlobalVector(d, &v); return v; }
void ref_vec_free(Vec v) { VecDestroy(v); }
}
