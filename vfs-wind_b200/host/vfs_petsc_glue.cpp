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
//   Pressure_Gradient                Source/momentum.c:203    -> vfs_pressure_gradient
//   Calc_F_eul / Calc_U_lagr         Source/rotor_model.c:3668,2937 -> vfs_calc_f_eul / vfs_calc_u_lagr
//   SNESSolve in Implicit_MatrixFree Source/implicitsolver.c:4299 -> vfs_glue_snes_solve -> vfs_momentum_solve
//
// Data contract: the UserCtx Vecs stay the source of truth on the host (the rest of VFS-Wind —
// Poisson solve, IBM, turbine models, I/O — keeps reading them), so each entry point uploads the
// Vecs the reference function reads and downloads the Vecs it writes.  Constant inputs (metrics,
// Nvert, Ucont_o, RHS_o, dP, F_eul, lUcat_old) are uploaded only when vfs_glue_invalidate() has
// been called since their last upload (the time loop calls it once per step), so a Krylov
// iteration moves exactly X down and F up.  MPI runs: one rank per GPU, the DA decomposed along k only
// (1 x 1 x P process grid, init.c:131-160 with -da_processors_x 1 -da_processors_y 1): every rank
// creates a context for its k-slab [info.zs, info.zs + info.zm) and the ncclUniqueId of the in-library
// halo layer travels by MPI_Bcast.
//
// This file is product code: it contains no arithmetic of the path and never calls the oracle.
#include "variables.h"
#include "vfs_b200.h"
#include <map>
#include <vector>

extern PetscInt les, second_order, immersed, inviscid, movefsi, rotatefsi, rotor_model, nacelle_model, IB_delta, wallfunction, ti, tistart;
extern int levelset_weno, freesurface_wallmodel, air_flow_levelset;
extern int laplacian, clark, central, testfilter_ik, viscosity_wallmodel, levelset, rans, skew;
extern int i_periodic, j_periodic, k_periodic, ii_periodic, jj_periodic, kk_periodic, i_homo_filter, j_homo_filter, k_homo_filter;
extern PetscReal max_cs;
extern double mean_pressure_gradient, inlet_flux;
extern PetscInt inletprofile;
extern PetscTruth dpdz_set;
extern PetscInt forcewidthfixed, ii_periodicWT, jj_periodicWT, kk_periodicWT;
extern PetscReal dhi_fixed, dhj_fixed, dhk_fixed, halfwidth_dfunc;
extern double roughness_size;
extern PetscTruth rstart_flg;

struct GlueState { vfs_ctx *ctx; bool const_valid, metrics_valid; double *buf; size_t buf_doubles; int zs, zm; };
static std::map<UserCtx *, GlueState> g_state;
static std::map<UserCtx *, Vec> g_last_x;      // the X of the most recent FormFunction_SNES (owned by the caller's SNES)
static inline bool any_periodic() { return ii_periodic || jj_periodic || kk_periodic || i_periodic || j_periodic || k_periodic; }
static int g_eager = 0;       // 1: FormFunction_SNES mirrors its side effects on the host Vecs at every call (see vfs_glue_sync_state)

extern "C" void vfs_glue_invalidate(UserCtx *user) { std::map<UserCtx *, GlueState>::iterator it = g_state.find(user); if (it != g_state.end()) it->second.const_valid = false; }
// the grid moved or the host rewrote the centre metrics itself: they go down again with the next call (FormMetrics through
// the glue leaves them on the device, so a fixed grid never ships its ten metric scalars per step)
extern "C" void vfs_glue_invalidate_grid(UserCtx *user) { std::map<UserCtx *, GlueState>::iterator it = g_state.find(user); if (it != g_state.end()) it->second.const_valid = it->second.metrics_valid = false; }
extern "C" void vfs_glue_release(UserCtx *user) {
  std::map<UserCtx *, GlueState>::iterator it = g_state.find(user);
  if (it != g_state.end()) { vfs_destroy(it->second.ctx); vfs_host_free(it->second.buf); g_state.erase(it); }
  g_last_x.erase(user);
}
extern "C" void vfs_glue_set_eager(int on) { g_eager = on; }

static void fill_params(UserCtx *user, vfs_params *p) {
  memset(p, 0, sizeof(*p));
  DALocalInfo info = user->info;
  int rank = 0, size = 1;
  MPI_Comm_rank(PETSC_COMM_WORLD, &rank); MPI_Comm_size(PETSC_COMM_WORLD, &size);
  if (info.xm != info.mx || info.ym != info.my) {
    PetscPrintf(PETSC_COMM_WORLD, "vfs_b200: the DA must be decomposed along k only (run with -da_processors_x 1 -da_processors_y 1)\n"); exit(1);
  }
  p->mx = info.mx; p->my = info.my; p->mz = info.mz; p->kofs = info.zs; p->nzl = info.zm; p->rank = rank; p->nranks = size;
  const int ndev = vfs_device_count();
  p->device = ndev > 0 ? rank % ndev : 0;
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
  // switches that change this path in the reference and are NOT built: the library rejects them (no silent central scheme)
  p->levelset_weno = levelset_weno; p->freesurface_wallmodel = freesurface_wallmodel; p->air_flow_levelset = air_flow_levelset;
}

static GlueState *state(UserCtx *user) {
  std::map<UserCtx *, GlueState>::iterator it = g_state.find(user);
  vfs_params p; fill_params(user, &p);
  if (it == g_state.end()) {
    GlueState s; s.ctx = 0; s.const_valid = false; s.metrics_valid = false; s.zs = p.kofs; s.zm = p.nzl;
    int r = vfs_create(&p, &s.ctx);
    if (r) { PetscPrintf(PETSC_COMM_WORLD, "vfs_b200: cannot create device context (%d): %s\n", r, vfs_last_error(0)); exit(1); }   // no CPU fallback
    if (p.nranks > 1) {         // in-library NCCL halo layer: rank 0's id to everybody
      char id[128]; memset(id, 0, sizeof(id));
      if (p.rank == 0 && vfs_nccl_unique_id(id)) { PetscPrintf(PETSC_COMM_WORLD, "vfs_b200: %s\n", vfs_last_error(0)); exit(1); }
      MPI_Bcast(id, 128, MPI_CHAR, 0, PETSC_COMM_WORLD);
      if (vfs_nccl_init(s.ctx, id)) { PetscPrintf(PETSC_COMM_SELF, "vfs_b200: %s\n", vfs_last_error(s.ctx)); exit(1); }
    }
    s.buf_doubles = (size_t)p.mx * p.my * p.nzl * 3;
    s.buf = (double *)vfs_host_alloc(s.buf_doubles * sizeof(double));       // pinned: the copies run at PCIe speed
    if (!s.buf) { PetscPrintf(PETSC_COMM_WORLD, "vfs_b200: cannot allocate the pinned staging buffer\n"); exit(1); }
    it = g_state.insert(std::make_pair(user, s)).first;
  } else if (vfs_set_params(it->second.ctx, &p)) {
    PetscPrintf(PETSC_COMM_WORLD, "vfs_b200: %s\n", vfs_last_error(it->second.ctx)); exit(1);
  }
  return &it->second;
}
static void ck(GlueState *s, int r, const char *what) { if (r) { PetscPrintf(PETSC_COMM_WORLD, "vfs_b200: %s failed (%d): %s\n", what, r, vfs_last_error(s->ctx)); exit(1); } }

// Vec (local or global, dof 1 or 3) -> the rank's owned block [zm][my][mx][dof] -> device field.  Rows are contiguous
// in both layouts (a ghosted local Vec only has a longer row pitch): one memcpy per row.
static void push(UserCtx *user, GlueState *s, Vec v, int dof, int field) {
  DALocalInfo info = user->info; const int mx = info.mx, my = info.my, zs = info.zs, ze = info.zs + info.zm;
  double *b = s->buf;
  const size_t row = (size_t)mx * dof;
  if (dof == 3) {
    Cmpnts ***a; DAVecGetArray(user->fda, v, &a);
    for (int k = zs; k < ze; k++) for (int j = 0; j < my; j++) memcpy(b + ((size_t)(k - zs) * my + j) * row, &a[k][j][0].x, row * sizeof(double));
    DAVecRestoreArray(user->fda, v, &a);
  } else {
    PetscReal ***a; DAVecGetArray(user->da, v, &a);
    for (int k = zs; k < ze; k++) for (int j = 0; j < my; j++) memcpy(b + ((size_t)(k - zs) * my + j) * row, &a[k][j][0], row * sizeof(double));
    DAVecRestoreArray(user->da, v, &a);
  }
  ck(s, vfs_upload(s->ctx, field, b), "vfs_upload");
}
// device field -> owned block of Vec v; a local v has its ghosts refreshed afterwards
static void pull(UserCtx *user, GlueState *s, int field, int dof, Vec v, bool v_is_local) {
  DALocalInfo info = user->info; const int mx = info.mx, my = info.my, zs = info.zs, ze = info.zs + info.zm;
  double *b = s->buf;
  const size_t row = (size_t)mx * dof;
  ck(s, vfs_download(s->ctx, field, b), "vfs_download");
  if (dof == 3) {
    Cmpnts ***a; DAVecGetArray(user->fda, v, &a);
    for (int k = zs; k < ze; k++) for (int j = 0; j < my; j++) memcpy(&a[k][j][0].x, b + ((size_t)(k - zs) * my + j) * row, row * sizeof(double));
    DAVecRestoreArray(user->fda, v, &a);
    if (v_is_local) { DALocalToLocalBegin(user->fda, v, INSERT_VALUES, v); DALocalToLocalEnd(user->fda, v, INSERT_VALUES, v); }
  } else {
    PetscReal ***a; DAVecGetArray(user->da, v, &a);
    for (int k = zs; k < ze; k++) for (int j = 0; j < my; j++) memcpy(&a[k][j][0], b + ((size_t)(k - zs) * my + j) * row, row * sizeof(double));
    DAVecRestoreArray(user->da, v, &a);
    if (v_is_local) { DALocalToLocalBegin(user->da, v, INSERT_VALUES, v); DALocalToLocalEnd(user->da, v, INSERT_VALUES, v); }
  }
}
static void push_constants(UserCtx *user, GlueState *s) {
  if (s->const_valid) return;
  if (!s->metrics_valid) {      // only when the device does not hold them yet (no FormMetrics through the glue since the context exists)
    push(user, s, user->lCsi, 3, VFS_CSI); push(user, s, user->lEta, 3, VFS_ETA); push(user, s, user->lZet, 3, VFS_ZET);
    push(user, s, user->lAj, 1, VFS_AJ);
    s->metrics_valid = true;
  }
  push(user, s, user->lNvert, 1, VFS_NVERT);
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

// Side outputs of FormMetrics that host code outside the path reads (bcs.c:3018, fsi.c:158, ibm search): the cell
// centres Cent / lCent (metrics.c:89-107: mean of the cell's eight corner nodes) and the grid spacings GridSpace /
// lGridSpace (metrics.c:420-497,1059-1060: distance between the centres of opposite faces), once per grid, on the host.
static void host_cent_gridspace(UserCtx *user) {
  DALocalInfo info = user->info; const int mx = info.mx, my = info.my, mz = info.mz;
  const int lxs = 1, lxe = mx - 1, lys = 1, lye = my - 1, lzs = info.zs == 0 ? 1 : info.zs, lze = info.zs + info.zm == mz ? mz - 1 : info.zs + info.zm;
  Vec Coor; DAGetGhostedCoordinates(user->da, &Coor);
  Cmpnts ***coor, ***cent, ***gs;
  DAVecGetArray(user->fda, Coor, &coor); DAVecGetArray(user->fda, user->Cent, &cent); DAVecGetArray(user->fda, user->GridSpace, &gs);
  for (int k = lzs; k < lze; k++) for (int j = lys; j < lye; j++) for (int i = lxs; i < lxe; i++) {
    // corner nodes of cell (i,j,k): (i-a, j-b, k-c), a,b,c in {0,1}; summed in the reference's order
    const Cmpnts c000 = coor[k][j][i], c010 = coor[k][j - 1][i], c001 = coor[k - 1][j][i], c011 = coor[k - 1][j - 1][i];
    const Cmpnts c100 = coor[k][j][i - 1], c110 = coor[k][j - 1][i - 1], c101 = coor[k - 1][j][i - 1], c111 = coor[k - 1][j - 1][i - 1];
    cent[k][j][i].x = 0.125 * (c000.x + c010.x + c001.x + c011.x + c100.x + c110.x + c101.x + c111.x);
    cent[k][j][i].y = 0.125 * (c000.y + c010.y + c001.y + c011.y + c100.y + c110.y + c101.y + c111.y);
    cent[k][j][i].z = 0.125 * (c000.z + c010.z + c001.z + c011.z + c100.z + c110.z + c101.z + c111.z);
    // face centres: +i face (nodes with index i), -i face (i-1), and twins; node order as in metrics.c
#define FC(A, B, C, D, m) (0.25 * ((A).m + (B).m + (C).m + (D).m))
#define DIST(A, B, C, D, E, F, G, H) sqrt((FC(A, B, C, D, x) - FC(E, F, G, H, x)) * (FC(A, B, C, D, x) - FC(E, F, G, H, x)) + \
                                          (FC(A, B, C, D, y) - FC(E, F, G, H, y)) * (FC(A, B, C, D, y) - FC(E, F, G, H, y)) + \
                                          (FC(A, B, C, D, z) - FC(E, F, G, H, z)) * (FC(A, B, C, D, z) - FC(E, F, G, H, z)))
    gs[k][j][i].x = DIST(c000, c010, c011, c001, c100, c110, c111, c101);
    gs[k][j][i].y = DIST(c000, c100, c001, c101, c010, c110, c011, c111);
    gs[k][j][i].z = DIST(c000, c100, c010, c110, c001, c101, c011, c111);
#undef DIST
#undef FC
  }
  DAVecRestoreArray(user->fda, Coor, &coor); DAVecRestoreArray(user->fda, user->Cent, &cent); DAVecRestoreArray(user->fda, user->GridSpace, &gs);
  DAGlobalToLocalBegin(user->fda, user->Cent, INSERT_VALUES, user->lCent); DAGlobalToLocalEnd(user->fda, user->Cent, INSERT_VALUES, user->lCent);
  DAGlobalToLocalBegin(user->fda, user->GridSpace, INSERT_VALUES, user->lGridSpace); DAGlobalToLocalEnd(user->fda, user->GridSpace, INSERT_VALUES, user->lGridSpace);
}

PetscErrorCode FormMetrics(UserCtx *user) {
  GlueState *s = state(user);
  Vec coords; DAGetGhostedCoordinates(user->da, &coords);
  push(user, s, coords, 3, VFS_COOR);
  ck(s, vfs_form_metrics(s->ctx), "vfs_form_metrics");
  pull(user, s, VFS_CSI, 3, user->lCsi, true); pull(user, s, VFS_ETA, 3, user->lEta, true); pull(user, s, VFS_ZET, 3, user->lZet, true);
  pull(user, s, VFS_AJ, 1, user->lAj, true);
  host_face_metrics(user);
  host_cent_gridspace(user);
  s->const_valid = false;
  s->metrics_valid = true;      // the device holds what the host Vecs now hold
  return 0;
}

void Contra2Cart_2(UserCtx *user) {
  GlueState *s = state(user);
  push_constants(user, s);
  push(user, s, user->lUcont, 3, VFS_UCONT);
  // The reference updates user->Ucat IN PLACE: fluid interior cells are recomputed, IB / solid cells and boundary nodes
  // no rule touches keep what the host wrote since the last call (FormBCS bcs.c:2404, ibm_interpolation_advanced
  // ibm.c:3673 are followed by Contra2Cart, implicitsolver.c:4406-4444) — so the current host Ucat goes down first.
  push(user, s, user->Ucat, 3, VFS_UCAT);
  ck(s, vfs_contra2cart(s->ctx), "vfs_contra2cart");
  if (any_periodic()) pull(user, s, VFS_UCONT, 3, user->lUcont, true);   // rhs.c:129-156 rewrites lUcont's periodic nodes
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

// momentum.c:203-439.  Side effects kept: lP's ghosts / periodic boundary nodes and P are refreshed (:247-286).
void Pressure_Gradient(UserCtx *user, Vec dP) {
  GlueState *s = state(user);
  push_constants(user, s);
  push(user, s, user->P, 1, VFS_P);
  double kf = 0;                                                       // :399-411
  if (k_periodic || kk_periodic) {
    if (dpdz_set) kf = mean_pressure_gradient;
    else if (inletprofile != 17) kf = (user->mean_k_flux - inlet_flux) / user->dt / user->mean_k_area;
  }
  ck(s, vfs_pressure_gradient(s->ctx, kf), "vfs_pressure_gradient");
  pull(user, s, VFS_DP, 3, dP, false);
  if (any_periodic()) {
    pull(user, s, VFS_P, 1, user->lP, true);
    DALocalToGlobal(user->da, user->lP, INSERT_VALUES, user->P);
  } else { DAGlobalToLocalBegin(user->da, user->P, INSERT_VALUES, user->lP); DAGlobalToLocalEnd(user->da, user->P, INSERT_VALUES, user->lP); }
}

// poisson.c:3137-3296: P += Phi, periodic boundary nodes of P / Phi, ghosts of lP / lPhi
PetscErrorCode UpdatePressure(UserCtx *user) {
  GlueState *s = state(user);
  push_constants(user, s);
  push(user, s, user->P, 1, VFS_P); push(user, s, user->Phi, 1, VFS_PHI);
  ck(s, vfs_update_pressure(s->ctx), "vfs_update_pressure");
  pull(user, s, VFS_P, 1, user->P, false); pull(user, s, VFS_PHI, 1, user->Phi, false);
  DAGlobalToLocalBegin(user->da, user->P, INSERT_VALUES, user->lP); DAGlobalToLocalEnd(user->da, user->P, INSERT_VALUES, user->lP);
  DAGlobalToLocalBegin(user->da, user->Phi, INSERT_VALUES, user->lPhi); DAGlobalToLocalEnd(user->da, user->Phi, INSERT_VALUES, user->lPhi);
  DAGlobalToLocalBegin(user->fda, user->Ucont, INSERT_VALUES, user->lUcont); DAGlobalToLocalEnd(user->fda, user->Ucont, INSERT_VALUES, user->lUcont);   // :3167-3168
  return 0;
}
// poisson.c:2700-3052: Ucont -= dt st grad(Phi), periodic copies, Contra2Cart (Ucat / lUcat of the corrected field)
PetscErrorCode Projection(UserCtx *user) {
  extern PetscReal poisson_threshold;
  GlueState *s = state(user);
  push_constants(user, s);
  push(user, s, user->lPhi, 1, VFS_PHI);
  push(user, s, user->Ucont, 3, VFS_UCONT);
  push(user, s, user->Ucat, 3, VFS_UCAT);      // Contra2Cart updates Ucat in place (see Contra2Cart_2 above)
  ck(s, vfs_projection(s->ctx, user->st, poisson_threshold), "vfs_projection");
  pull(user, s, VFS_UCONT, 3, user->Ucont, false);
  DAGlobalToLocalBegin(user->fda, user->Ucont, INSERT_VALUES, user->lUcont); DAGlobalToLocalEnd(user->fda, user->Ucont, INSERT_VALUES, user->lUcont);
  ck(s, vfs_contra2cart(s->ctx), "vfs_contra2cart");                    // :3049, on the fields already on the device
  if (any_periodic()) pull(user, s, VFS_UCONT, 3, user->lUcont, true);   // rhs.c:129-156
  pull(user, s, VFS_UCAT, 3, user->Ucat, false);
  DAGlobalToLocalBegin(user->fda, user->Ucat, INSERT_VALUES, user->lUcat); DAGlobalToLocalEnd(user->fda, user->Ucat, INSERT_VALUES, user->lUcat);
  for (int q = 0; q < 4; q++) if (user->bctype[q] == -1 || user->bctype[q] == -2) { pull(user, s, VFS_USTAR, 1, user->lUstar, false); break; }
  return 0;
}

// rotor_model.c:3668-3960 / 2937-3150: the IBMNodes arrays go down as they are, F_eul accumulates onto the host's lF_eul
static std::vector<vfs_actuator> actuators(IBMNodes *ibm, int n) {
  std::vector<vfs_actuator> a(n);
  for (int b = 0; b < n; b++) {
    a[b].n_elmt = ibm[b].n_elmt;
    a[b].cent_x = ibm[b].cent_x; a[b].cent_y = ibm[b].cent_y; a[b].cent_z = ibm[b].cent_z; a[b].dA = ibm[b].dA;
    a[b].F_lagr_x = ibm[b].F_lagr_x; a[b].F_lagr_y = ibm[b].F_lagr_y; a[b].F_lagr_z = ibm[b].F_lagr_z;
    a[b].U_lagr_x = ibm[b].U_lagr_x; a[b].U_lagr_y = ibm[b].U_lagr_y; a[b].U_lagr_z = ibm[b].U_lagr_z;
    a[b].i_min = ibm[b].i_min; a[b].i_max = ibm[b].i_max; a[b].j_min = ibm[b].j_min; a[b].j_max = ibm[b].j_max; a[b].k_min = ibm[b].k_min; a[b].k_max = ibm[b].k_max;
  }
  return a;
}
PetscErrorCode Calc_F_eul(UserCtx *user, IBMNodes *ibm, FSInfo *fsi, PetscInt NumberOfObjects, double dh, int df) {
  GlueState *s = state(user);
  push_constants(user, s);
  push(user, s, user->lF_eul, 3, VFS_F_EUL);
  std::vector<vfs_actuator> a = actuators(ibm, NumberOfObjects);
  const double dhf[3] = {dhi_fixed, dhj_fixed, dhk_fixed};
  ck(s, vfs_calc_f_eul(s->ctx, NumberOfObjects, a.data(), df, halfwidth_dfunc, forcewidthfixed, dhf, 1), "vfs_calc_f_eul");
  pull(user, s, VFS_F_EUL, 3, user->lF_eul, true);
  DALocalToGlobal(user->fda, user->lF_eul, INSERT_VALUES, user->F_eul);
  return 0;
}
// Host bookkeeping that follows the interpolation in the reference (rotor_model.c:3062-3138), on the IBMNodes arrays only:
// in a periodic turbine array (Nx_WT x Ny_WT x Nz_WT objects, spacing Sx/Sy/Sz_WT) the first and the last object of a
// periodic direction are images of each other — the first receives the sum of both, then the last a copy of the first —
// and a moving frame adds its velocity.
static void ulagr_postprocess(IBMNodes *ibm, FSInfo *fsi, int nobj) {
  extern PetscInt MoveFrame, Nx_WT, Ny_WT, Nz_WT;
  extern PetscReal u_frame, v_frame, w_frame, Sx_WT, Sy_WT, Sz_WT;
  if (ii_periodicWT || jj_periodicWT || kk_periodicWT) {
    double cmin[3] = {1.0e6, 1.0e6, 1.0e6};
    for (int b = 0; b < nobj; b++) { cmin[0] = PetscMin(cmin[0], fsi[b].x_c); cmin[1] = PetscMin(cmin[1], fsi[b].y_c); cmin[2] = PetscMin(cmin[2], fsi[b].z_c); }
    const int n[3] = {(int)Nx_WT, (int)Ny_WT, (int)Nz_WT};
    std::vector<int> ind((size_t)n[0] * n[1] * n[2], 0);
    const double fac[3] = {1.0 / Sx_WT, 1.0 / Sy_WT, 1.0 / Sz_WT};
    for (int b = 0; b < nobj; b++) {
      const int ii = (int)((fsi[b].x_c - cmin[0] + 1.e-9) * fac[0]), jj = (int)((fsi[b].y_c - cmin[1] + 1.e-9) * fac[1]), kk = (int)((fsi[b].z_c - cmin[2] + 1.e-9) * fac[2]);
      PetscPrintf(PETSC_COMM_WORLD, "ibi ii jj kk %d %d %d %d \n", b, ii, jj, kk);
      ind[((size_t)kk * n[1] + jj) * n[0] + ii] = b;
    }
    const int per[3] = {(int)ii_periodicWT, (int)jj_periodicWT, (int)kk_periodicWT};
    // pass 0: first += last (directions in the order i, j, k at every grid position); pass 1: last = first
    for (int pass = 0; pass < 2; pass++)
      for (int k = 0; k < n[2]; k++) for (int j = 0; j < n[1]; j++) for (int i = 0; i < n[0]; i++) {
        const int c[3] = {i, j, k};
        for (int D = 0; D < 3; D++) {
          if (!per[D] || c[D] != (pass == 0 ? 0 : n[D] - 1)) continue;
          int o[3] = {i, j, k}; o[D] = pass == 0 ? n[D] - 1 : 0;
          IBMNodes &me = ibm[ind[((size_t)k * n[1] + j) * n[0] + i]], &other = ibm[ind[((size_t)o[2] * n[1] + o[1]) * n[0] + o[0]]];
          for (int l = 0; l < me.n_elmt; l++) {
            if (pass == 0) { me.U_lagr_x[l] = me.U_lagr_x[l] + other.U_lagr_x[l]; me.U_lagr_y[l] = me.U_lagr_y[l] + other.U_lagr_y[l]; me.U_lagr_z[l] = me.U_lagr_z[l] + other.U_lagr_z[l]; }
            else { me.U_lagr_x[l] = other.U_lagr_x[l]; me.U_lagr_y[l] = other.U_lagr_y[l]; me.U_lagr_z[l] = other.U_lagr_z[l]; }
          }
        }
      }
  }
  if (MoveFrame)
    for (int b = 0; b < nobj; b++) for (int l = 0; l < ibm[b].n_elmt; l++) { ibm[b].U_lagr_x[l] += u_frame; ibm[b].U_lagr_y[l] += v_frame; ibm[b].U_lagr_z[l] += w_frame; }
}
PetscErrorCode Calc_U_lagr(UserCtx *user, IBMNodes *ibm, FSInfo *fsi, int NumberOfObjects) {
  GlueState *s = state(user);
  push_constants(user, s);
  push(user, s, user->lUcat, 3, VFS_UCAT);
  std::vector<vfs_actuator> a = actuators(ibm, NumberOfObjects);
  ck(s, vfs_calc_u_lagr(s->ctx, NumberOfObjects, a.data()), "vfs_calc_u_lagr");      // interpolation + the sum over ranks
  ulagr_postprocess(ibm, fsi, NumberOfObjects);
  return 0;
}

// body-fitted cylinder runs (bctype[0] == 11 with a wall at i = mx-2): the wall areas and pressure / viscous force sums
// Formfunction_2 leaves in the context for main.c:1269-1277 (momentum.c:570-579, 822-849)
static void cylinder_diag(UserCtx *user, GlueState *s) {
  if (user->bctype[0] != 11 || user->bctype[1] != 1) return;
  push(user, s, user->lP, 1, VFS_P);
  double o[7];
  ck(s, vfs_cylinder_forces(s->ctx, o), "vfs_cylinder_forces");
  user->lA_cyl = o[0]; user->lA_cyl_x = o[1]; user->lA_cyl_z = o[2]; user->lFpx_cyl = o[3]; user->lFpz_cyl = o[4]; user->lFvx_cyl = o[5]; user->lFvz_cyl = o[6];
}

PetscErrorCode Formfunction_2(UserCtx *user, Vec Rhs, double scale) {
  GlueState *s = state(user);
  push_constants(user, s);
  push(user, s, user->lUcont, 3, VFS_UCONT); push(user, s, user->lUcat, 3, VFS_UCAT);
  if (les) push(user, s, user->lNu_t, 1, VFS_NU_T);
  // RHS_o = Formfunction_2(U_o) is assembled once per step AFTER the per-step constants went down (solvers.c:629): when the
  // target is user->RHS_o the result is accumulated in the device's RHS_o field itself, so the residual evaluations that
  // follow read the fresh one without another upload
  const int field = Rhs == user->RHS_o ? VFS_RHS_O : VFS_RHS;
  push(user, s, Rhs, 3, field);
  ck(s, vfs_formfunction2(s->ctx, field, scale), "vfs_formfunction2");
  pull(user, s, field, 3, Rhs, false);
  if (viscosity_wallmodel && les) pull(user, s, VFS_USTAR, 1, user->lUstar, false);      // momentum.c:1150
  cylinder_diag(user, s);
  return 0;
}

// Side effects of the reference's FormFunction_SNES on the UserCtx Vecs (momentum.c:2240-2295: Ucont <- X with the
// wall-normal fluxes zeroed, lUcont, Ucat / lUcat through Contra2Cart_2, lUstar, and on the first step IB_BC's nvert
// marking in lNvert / Nvert).  The reference reads them right after SNESSolve (outflow BCs, VecMax(Ucat),
// implicitsolver.c:4360-4376).  They live on the device during the Krylov iterations; this call materialises them on
// the host.  INTEGRATION.md: call it once after SNESSolve — or switch vfs_glue_set_eager(1) on, and FormFunction_SNES
// does it at every evaluation (a strict drop-in at three times the PCIe traffic).
// user->Ucont <- X with the wall-normal fluxes zeroed (momentum.c:2240, 2264-2289), on the host: the global Ucont does
// NOT receive the periodic boundary-node rewrites that Contra2Cart_2 / IB_BC apply to lUcont (the device holds lUcont)
static void host_ucont_from_x(UserCtx *user, Vec X) {
  VecCopy(X, user->Ucont);
  DALocalInfo info = user->info; const int mx = info.mx, my = info.my, mz = info.mz, zs = info.zs, ze = info.zs + info.zm;
  const int *bc = user->bctype;
  Cmpnts ***u; DAVecGetArray(user->fda, user->Ucont, &u);
  for (int k = zs; k < ze; k++) for (int j = 0; j < my; j++) {
    const bool jin = j != 0 && j != my - 1, kin = k != 0 && k != mz - 1;
    if (bc[0] == 1 || (bc[0] == 10 && jin && kin)) u[k][j][0].x = 0;
    if (bc[1] == 1 || (bc[1] == 10 && jin && kin)) u[k][j][mx - 2].x = 0;
  }
  for (int k = zs; k < ze; k++) for (int i = 0; i < mx; i++) {
    const bool iin = i != 0 && i != mx - 1, kin = k != 0 && k != mz - 1;
    if (bc[2] == 1 || bc[2] == 12 || (bc[2] == 10 && iin && kin)) u[k][0][i].y = 0;
    if (bc[3] == 1 || bc[3] == 2 || bc[3] == 12 || ((bc[3] == 10 || bc[3] == -10) && iin && kin)) u[k][my - 2][i].y = 0;
  }
  for (int j = 0; j < my; j++) for (int i = 0; i < mx; i++) {
    if (bc[4] == 1 && zs == 0) u[0][j][i].z = 0;
    if (bc[5] == 1 && mz - 2 >= zs && mz - 2 < ze) u[mz - 2][j][i].z = 0;
  }
  DAVecRestoreArray(user->fda, user->Ucont, &u);
}

extern "C" void vfs_glue_sync_state(UserCtx *user) {
  GlueState *s = state(user);
  std::map<UserCtx *, Vec>::iterator lx = g_last_x.find(user);
  if (lx != g_last_x.end() && lx->second) host_ucont_from_x(user, lx->second);
  pull(user, s, VFS_UCONT, 3, user->lUcont, true);       // lUcont: periodic boundary nodes rewritten by Contra2Cart_2 / IB_BC
  pull(user, s, VFS_UCAT, 3, user->Ucat, false);
  DAGlobalToLocalBegin(user->fda, user->Ucat, INSERT_VALUES, user->lUcat); DAGlobalToLocalEnd(user->fda, user->Ucat, INSERT_VALUES, user->lUcat);
  bool wallfn = false;
  for (int q = 0; q < 6; q++) wallfn = wallfn || user->bctype[q] == -1 || user->bctype[q] == -2;
  if (wallfn || (viscosity_wallmodel && les)) pull(user, s, VFS_USTAR, 1, user->lUstar, false);   // rhs.c:336,371,401,435; momentum.c:1150
  if (wallfn && !immersed && ti == tistart) {            // momentum.c:2048-2074
    pull(user, s, VFS_NVERT, 1, user->lNvert, true);
    DALocalToGlobal(user->da, user->lNvert, INSERT_VALUES, user->Nvert);
  }
  cylinder_diag(user, s);
}

PetscErrorCode FormFunction_SNES(SNES snes, Vec Ucont, Vec Rhs, void *ptr) {
  UserCtx *user = (UserCtx *)ptr;
  GlueState *s = state(user);
  push_constants(user, s);
  PetscScalar *x, *f;
  VecGetArray(Ucont, &x); VecGetArray(Rhs, &f);
  ck(s, vfs_formfunction_snes(s->ctx, x, f), "vfs_formfunction_snes");      // X down, F up: the only per-Krylov-iteration traffic
  VecRestoreArray(Ucont, &x); VecRestoreArray(Rhs, &f);
  g_last_x[user] = Ucont;
  if (g_eager) vfs_glue_sync_state(user);
  return 0;
}

// The one-line replacement of `SNESSolve(user->snes, PETSC_NULL, U)` in Implicit_MatrixFree (implicitsolver.c:4299):
// the whole Newton-Krylov solve on the device (vfs_momentum_solve), U in/out, side effects mirrored on the host Vecs.
// p == NULL: the reference's settings (vfs_solver_defaults with snes_rtol = imp_free_tol).  Returns the final |F|.
extern "C" double vfs_glue_snes_solve(UserCtx *user, Vec U, const vfs_solver_params *p, vfs_solver_info *info_out) {
  GlueState *s = state(user);
  push_constants(user, s);
  push(user, s, U, 3, VFS_UCONT);
  vfs_solver_params sp;
  if (p) sp = *p; else { extern double imp_free_tol; vfs_solver_defaults(&sp); sp.snes_rtol = imp_free_tol; sp.ksp_rtol = imp_free_tol; }
  vfs_solver_info info;
  ck(s, vfs_momentum_solve(s->ctx, &sp, &info), "vfs_momentum_solve");
  pull(user, s, VFS_UCONT, 3, U, false);
  // FormFunction_SNES's side effects as of the last evaluation the reference's SNES would have made (the accepted iterate)
  ck(s, vfs_formfunction_snes_dev(s->ctx), "vfs_formfunction_snes_dev");
  g_last_x[user] = U;
  vfs_glue_sync_state(user);
  if (info_out) *info_out = info;
  return info.fnorm;
}
