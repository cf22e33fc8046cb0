// vfs_petsc_glue.cpp — host-side drop-in for the reference's momentum RHS + LES entry points.
//
// Compiled INSIDE the reference tree (it includes the reference's own variables.h and PETSc 3.1
// headers) in place of the corresponding function bodies; every function keeps the reference's
// name, signature and side effects on the UserCtx Vecs, and forwards the arithmetic to the
// sm_100a kernels through the C ABI of include/vfs_b200.h:
//
//   FormMetrics                      Source/metrics.c:12      -> vfs_form_metrics
//   Contra2Cart / Contra2Cart_2      Source/rhs.c:35,65       -> vfs_contra2cart
//   IB_BC                            Source/momentum.c:2016   -> vfs_ib_bc
//   Formfunction_2                   Source/momentum.c:454    -> vfs_formfunction2
//   FormFunction_SNES                Source/momentum.c:2237   -> vfs_formfunction_snes
//   Compute_Smagorinsky_Constant_1   Source/les.c:75          -> vfs_les_cs
//   Compute_eddy_viscosity_LES       Source/les.c:1143        -> vfs_les_nut
//   Convection / Viscous (legacy)    Source/rhs.c:751,1071    -> vfs_convection / vfs_viscous
//
// Data contract: the UserCtx Vecs stay the source of truth on the host (the rest of VFS-Wind —
// Poisson solve, IBM, turbine models, I/O — keeps reading them), so each entry point uploads the
// Vecs the reference function reads and downloads the Vecs it writes.  Constant inputs (metrics,
// Nvert, Ucont_o, RHS_o, dP, F_eul, lUcat_old) are uploaded only when vfs_glue_invalidate() has
// been called since their last upload (the time loop calls it once per step), so a Krylov
// iteration moves exactly X down and F up.  Single rank (one GPU) in this round; the k-slab
// multi-GPU path is driven through the C ABI + halo callback directly.
//
// This file is product code: it contains no arithmetic of the path and never calls the oracle.
#include "variables.h"
#include "vfs_b200.h"
#include <map>
#include <vector>

extern PetscInt les, second_order, immersed, inviscid, movefsi, rotatefsi, rotor_model, nacelle_model, IB_delta, wallfunction, ti, tistart;
extern int laplacian, clark, central, testfilter_ik, viscosity_wallmodel, levelset, rans, skew;
extern int i_periodic, j_periodic, k_periodic, ii_periodic, jj_periodic, kk_periodic, i_homo_filter, j_homo_filter, k_homo_filter;
extern PetscReal max_cs;
extern double roughness_size;
extern PetscTruth rstart_flg;

struct GlueState { vfs_ctx *ctx; bool const_valid; std::vector<double> buf; };
static std::map<UserCtx *, GlueState> g_state;

extern "C" void vfs_glue_invalidate(UserCtx *user) { std::map<UserCtx *, GlueState>::iterator it = g_state.find(user); if (it != g_state.end()) it->second.const_valid = false; }
extern "C" void vfs_glue_release(UserCtx *user) { std::map<UserCtx *, GlueState>::iterator it = g_state.find(user); if (it != g_state.end()) { vfs_destroy(it->second.ctx); g_state.erase(it); } }

static void fill_params(UserCtx *user, vfs_params *p) {
  memset(p, 0, sizeof(*p));
  DALocalInfo info = user->info;
  p->mx = info.mx; p->my = info.my; p->mz = info.mz; p->kofs = 0; p->nzl = info.mz; p->rank = 0; p->nranks = 1; p->device = 0;
  p->ii_periodic = ii_periodic; p->jj_periodic = jj_periodic; p->kk_periodic = kk_periodic;
  for (int q = 0; q < 6; q++) p->bctype[q] = user->bctype[q];
  p->les = les; p->second_order = second_order; p->laplacian = laplacian; p->immersed = immersed; p->clark = clark; p->central = central;
  p->testfilter_ik = testfilter_ik; p->viscosity_wallmodel = viscosity_wallmodel; p->wallfunction = wallfunction;
  p->rotor_model = rotor_model; p->nacelle_model = nacelle_model; p->IB_delta = IB_delta;
  p->ti = ti; p->tistart = tistart; p->rstart_flg = rstart_flg;
  p->levelset = levelset; p->rans = rans; p->inviscid = inviscid; p->skew = skew; p->movefsi = movefsi; p->rotatefsi = rotatefsi;
  p->i_periodic = i_periodic; p->j_periodic = j_periodic; p->k_periodic = k_periodic;
  p->i_homo_filter = i_homo_filter; p->j_homo_filter = j_homo_filter; p->k_homo_filter = k_homo_filter;
  p->ren = user->ren; p->dt = user->dt; p->max_cs = max_cs; p->roughness_size = roughness_size;
}

static GlueState *state(UserCtx *user) {
  std::map<UserCtx *, GlueState>::iterator it = g_state.find(user);
  vfs_params p; fill_params(user, &p);
  if (it == g_state.end()) {
    GlueState s; s.ctx = 0; s.const_valid = false;
    int r = vfs_create(&p, &s.ctx);
    if (r) { PetscPrintf(PETSC_COMM_WORLD, "vfs_b200: cannot create device context (%d): %s\n", r, vfs_last_error(0)); exit(1); }   // no CPU fallback
    s.buf.resize((size_t)p.mx * p.my * p.mz * 3);
    it = g_state.insert(std::make_pair(user, s)).first;
  } else if (vfs_set_params(it->second.ctx, &p)) {
    PetscPrintf(PETSC_COMM_WORLD, "vfs_b200: %s\n", vfs_last_error(it->second.ctx)); exit(1);
  }
  return &it->second;
}
static void ck(GlueState *s, int r, const char *what) { if (r) { PetscPrintf(PETSC_COMM_WORLD, "vfs_b200: %s failed (%d): %s\n", what, r, vfs_last_error(s->ctx)); exit(1); } }

// Vec (local or global, dof 1 or 3) -> owned block -> device field
static void push(UserCtx *user, GlueState *s, Vec v, int dof, int field) {
  DALocalInfo info = user->info; const int mx = info.mx, my = info.my, mz = info.mz;
  double *b = s->buf.data();
  if (dof == 3) {
    Cmpnts ***a; DAVecGetArray(user->fda, v, &a);
    for (int k = 0; k < mz; k++) for (int j = 0; j < my; j++) for (int i = 0; i < mx; i++) { size_t q = (((size_t)k * my + j) * mx + i) * 3; b[q] = a[k][j][i].x; b[q + 1] = a[k][j][i].y; b[q + 2] = a[k][j][i].z; }
    DAVecRestoreArray(user->fda, v, &a);
  } else {
    PetscReal ***a; DAVecGetArray(user->da, v, &a);
    for (int k = 0; k < mz; k++) for (int j = 0; j < my; j++) for (int i = 0; i < mx; i++) b[((size_t)k * my + j) * mx + i] = a[k][j][i];
    DAVecRestoreArray(user->da, v, &a);
  }
  ck(s, vfs_upload(s->ctx, field, b), "vfs_upload");
}
// device field -> owned block of Vec v; if `local` is given it is refreshed from v (ghosts included)
static void pull(UserCtx *user, GlueState *s, int field, int dof, Vec v, bool v_is_local) {
  DALocalInfo info = user->info; const int mx = info.mx, my = info.my, mz = info.mz;
  double *b = s->buf.data();
  ck(s, vfs_download(s->ctx, field, b), "vfs_download");
  if (dof == 3) {
    Cmpnts ***a; DAVecGetArray(user->fda, v, &a);
    for (int k = 0; k < mz; k++) for (int j = 0; j < my; j++) for (int i = 0; i < mx; i++) { size_t q = (((size_t)k * my + j) * mx + i) * 3; a[k][j][i].x = b[q]; a[k][j][i].y = b[q + 1]; a[k][j][i].z = b[q + 2]; }
    DAVecRestoreArray(user->fda, v, &a);
    if (v_is_local) { DALocalToLocalBegin(user->fda, v, INSERT_VALUES, v); DALocalToLocalEnd(user->fda, v, INSERT_VALUES, v); }
  } else {
    PetscReal ***a; DAVecGetArray(user->da, v, &a);
    for (int k = 0; k < mz; k++) for (int j = 0; j < my; j++) for (int i = 0; i < mx; i++) a[k][j][i] = b[((size_t)k * my + j) * mx + i];
    DAVecRestoreArray(user->da, v, &a);
    if (v_is_local) { DALocalToLocalBegin(user->da, v, INSERT_VALUES, v); DALocalToLocalEnd(user->da, v, INSERT_VALUES, v); }
  }
}
static void push_constants(UserCtx *user, GlueState *s) {
  if (s->const_valid) return;
  push(user, s, user->lCsi, 3, VFS_CSI); push(user, s, user->lEta, 3, VFS_ETA); push(user, s, user->lZet, 3, VFS_ZET);
  push(user, s, user->lAj, 1, VFS_AJ); push(user, s, user->lNvert, 1, VFS_NVERT);
  push(user, s, user->lUcat_old, 3, VFS_UCAT_OLD);
  push(user, s, user->Ucat, 3, VFS_UCAT);                 // persistent in/out state (IBM cells, inflow ghosts)
  push(user, s, user->Ucont_o, 3, VFS_UCONT_O);
  if (user->Ucont_rm1) push(user, s, user->Ucont_rm1, 3, VFS_UCONT_RM1);
  push(user, s, user->RHS_o, 3, VFS_RHS_O); push(user, s, user->dP, 3, VFS_DP);
  if ((rotor_model || nacelle_model || IB_delta) && user->F_eul) push(user, s, user->F_eul, 3, VFS_F_EUL);
  if (les) { push(user, s, user->lNu_t, 1, VFS_NU_T); push(user, s, user->lCs, 1, VFS_CS); }
  s->const_valid = true;
}

// face metrics for the host code that still reads them (Pressure_Gradient, Projection, ...):
// NEWMETRIC definition, metrics.c:589-592,697-700,805-808
static void host_face_metrics(UserCtx *user) {
  DALocalInfo info = user->info; const int mx = info.mx, my = info.my, mz = info.mz;
  Cmpnts ***c[3], ***f[3][3]; PetscReal ***aj, ***faj[3];
  Vec cen[3] = {user->lCsi, user->lEta, user->lZet};
  Vec fv[3][3] = {{user->lICsi, user->lIEta, user->lIZet}, {user->lJCsi, user->lJEta, user->lJZet}, {user->lKCsi, user->lKEta, user->lKZet}};
  Vec fa[3] = {user->lIAj, user->lJAj, user->lKAj};
  for (int m = 0; m < 3; m++) DAVecGetArray(user->fda, cen[m], &c[m]);
  DAVecGetArray(user->da, user->lAj, &aj);
  for (int D = 0; D < 3; D++) { for (int m = 0; m < 3; m++) DAVecGetArray(user->fda, fv[D][m], &f[D][m]); DAVecGetArray(user->da, fa[D], &faj[D]); }
  for (int D = 0; D < 3; D++) {
    const int di = D == 0, dj = D == 1, dk = D == 2;
    for (int k = dk ? 0 : 1; k < mz - 1; k++) for (int j = dj ? 0 : 1; j < my - 1; j++) for (int i = di ? 0 : 1; i < mx - 1; i++) {
      for (int m = 0; m < 3; m++) {
        Cmpnts a = c[m][k][j][i], b = c[m][k + dk][j + dj][i + di];
        f[D][m][k][j][i].x = 0.5 * a.x + 0.5 * b.x; f[D][m][k][j][i].y = 0.5 * a.y + 0.5 * b.y; f[D][m][k][j][i].z = 0.5 * a.z + 0.5 * b.z;
      }
      faj[D][k][j][i] = 2. / (1. / aj[k][j][i] + 1. / aj[k + dk][j + dj][i + di]);
    }
  }
  for (int D = 0; D < 3; D++) { for (int m = 0; m < 3; m++) DAVecRestoreArray(user->fda, fv[D][m], &f[D][m]); DAVecRestoreArray(user->da, fa[D], &faj[D]); }
  DAVecRestoreArray(user->da, user->lAj, &aj);
  for (int m = 0; m < 3; m++) DAVecRestoreArray(user->fda, cen[m], &c[m]);
  for (int D = 0; D < 3; D++) {
    for (int m = 0; m < 3; m++) { DALocalToLocalBegin(user->fda, fv[D][m], INSERT_VALUES, fv[D][m]); DALocalToLocalEnd(user->fda, fv[D][m], INSERT_VALUES, fv[D][m]); }
    DALocalToLocalBegin(user->da, fa[D], INSERT_VALUES, fa[D]); DALocalToLocalEnd(user->da, fa[D], INSERT_VALUES, fa[D]);
  }
}

PetscErrorCode FormMetrics(UserCtx *user) {
  GlueState *s = state(user);
  Vec coords; DAGetGhostedCoordinates(user->da, &coords);
  push(user, s, coords, 3, VFS_COOR);
  ck(s, vfs_form_metrics(s->ctx), "vfs_form_metrics");
  pull(user, s, VFS_CSI, 3, user->lCsi, true); pull(user, s, VFS_ETA, 3, user->lEta, true); pull(user, s, VFS_ZET, 3, user->lZet, true);
  pull(user, s, VFS_AJ, 1, user->lAj, true);
  host_face_metrics(user);
  s->const_valid = false;
  return 0;
}

void Contra2Cart_2(UserCtx *user) {
  GlueState *s = state(user);
  push_constants(user, s);
  push(user, s, user->lUcont, 3, VFS_UCONT);
  ck(s, vfs_contra2cart(s->ctx), "vfs_contra2cart");
  if (ii_periodic || jj_periodic || kk_periodic) pull(user, s, VFS_UCONT, 3, user->lUcont, true);   // rhs.c:129-156 rewrites lUcont's periodic nodes
  pull(user, s, VFS_UCAT, 3, user->Ucat, false);
  DAGlobalToLocalBegin(user->fda, user->Ucat, INSERT_VALUES, user->lUcat); DAGlobalToLocalEnd(user->fda, user->Ucat, INSERT_VALUES, user->lUcat);
  for (int q = 0; q < 4; q++) if (user->bctype[q] == -1 || user->bctype[q] == -2) { pull(user, s, VFS_USTAR, 1, user->lUstar, false); break; }   // rhs.c:336,371,401,435
}
void Contra2Cart(UserCtx *user) { Contra2Cart_2(user); }

void IB_BC(UserCtx *user) {
  GlueState *s = state(user);
  push_constants(user, s);
  push(user, s, user->lUcont, 3, VFS_UCONT); push(user, s, user->lUcat, 3, VFS_UCAT);
  ck(s, vfs_ib_bc(s->ctx), "vfs_ib_bc");
  pull(user, s, VFS_UCONT, 3, user->lUcont, true);
  // first time step without immersed bodies: IB_BC marks the first cells of wall-function sides nvert = 1
  // (momentum.c:2048-2074) in lNvert and Nvert
  bool wallfn = false;
  for (int q = 0; q < 6; q++) wallfn = wallfn || user->bctype[q] == -1 || user->bctype[q] == -2;
  if (wallfn && !immersed && ti == tistart) {
    pull(user, s, VFS_NVERT, 1, user->lNvert, true);
    DALocalToGlobal(user->da, user->lNvert, INSERT_VALUES, user->Nvert);
  }
}

void Compute_Smagorinsky_Constant_1(UserCtx *user, Vec Ucont, Vec Ucat) {
  GlueState *s = state(user);
  push_constants(user, s);
  push(user, s, Ucat, 3, VFS_UCAT);
  ck(s, vfs_les_cs(s->ctx), "vfs_les_cs");
  pull(user, s, VFS_CS, 1, user->lCs, true);
}
void Compute_eddy_viscosity_LES(UserCtx *user) {
  GlueState *s = state(user);
  push_constants(user, s);
  push(user, s, user->lUcat, 3, VFS_UCAT); push(user, s, user->lCs, 1, VFS_CS);
  ck(s, vfs_les_nut(s->ctx), "vfs_les_nut");
  pull(user, s, VFS_NU_T, 1, user->lNu_t, true);
}

// legacy explicit-solver terms (callers: timeadvancing1.c:75-76 FormFunctionSNES, Prediction)
PetscErrorCode Convection(UserCtx *user, Vec Ucont, Vec Ucat, Vec Conv) {
  GlueState *s = state(user);
  push_constants(user, s);
  push(user, s, Ucont, 3, VFS_UCONT); push(user, s, Ucat, 3, VFS_UCAT);
  ck(s, vfs_convection(s->ctx), "vfs_convection");
  pull(user, s, VFS_CONV, 3, Conv, false);
  return 0;
}
PetscErrorCode Viscous(UserCtx *user, Vec Ucont, Vec Ucat, Vec Visc) {
  GlueState *s = state(user);
  push_constants(user, s);
  push(user, s, Ucont, 3, VFS_UCONT); push(user, s, Ucat, 3, VFS_UCAT);
  if (les) push(user, s, user->lNu_t, 1, VFS_NU_T);
  ck(s, vfs_viscous(s->ctx), "vfs_viscous");
  pull(user, s, VFS_VISC, 3, Visc, false);
  return 0;
}

PetscErrorCode Formfunction_2(UserCtx *user, Vec Rhs, double scale) {
  GlueState *s = state(user);
  push_constants(user, s);
  push(user, s, user->lUcont, 3, VFS_UCONT); push(user, s, user->lUcat, 3, VFS_UCAT);
  if (les) push(user, s, user->lNu_t, 1, VFS_NU_T);
  push(user, s, Rhs, 3, VFS_RHS);
  ck(s, vfs_formfunction2(s->ctx, VFS_RHS, scale), "vfs_formfunction2");
  pull(user, s, VFS_RHS, 3, Rhs, false);
  if (viscosity_wallmodel && les) pull(user, s, VFS_USTAR, 1, user->lUstar, false);      // momentum.c:1150
  return 0;
}

PetscErrorCode FormFunction_SNES(SNES snes, Vec Ucont, Vec Rhs, void *ptr) {
  UserCtx *user = (UserCtx *)ptr;
  GlueState *s = state(user);
  push_constants(user, s);
  PetscScalar *x, *f;
  VecGetArray(Ucont, &x); VecGetArray(Rhs, &f);
  ck(s, vfs_formfunction_snes(s->ctx, x, f), "vfs_formfunction_snes");      // X down, F up: the only per-Krylov-iteration traffic
  VecRestoreArray(Ucont, &x); VecRestoreArray(Rhs, &f);
  return 0;
}
// Side effects of the reference's FormFunction_SNES on user->Ucont/lUcont/Ucat/lUcat are
// materialised on the host on demand (after the SNES solve the caller runs Contra2Cart anyway,
// implicitsolver.c:4444); call this to mirror them explicitly.
extern "C" void vfs_glue_sync_state(UserCtx *user) {
  GlueState *s = state(user);
  pull(user, s, VFS_UCONT, 3, user->lUcont, true);
  pull(user, s, VFS_UCAT, 3, user->Ucat, false);
  DAGlobalToLocalBegin(user->fda, user->Ucat, INSERT_VALUES, user->lUcat); DAGlobalToLocalEnd(user->fda, user->Ucat, INSERT_VALUES, user->lUcat);
  for (int q = 0; q < 4; q++) if (user->bctype[q] == -1 || user->bctype[q] == -2) { pull(user, s, VFS_USTAR, 1, user->lUstar, false); break; }   // rhs.c:336,371,401,435
}
