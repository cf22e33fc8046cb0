// Single-rank implementation of the PETSc-3.1 / MPI subset declared in petsc_shim.h.
// TEST INFRASTRUCTURE ONLY (parity checker + CPU baseline); see petsc_shim.h.
// DA model: one rank owns the whole mx*my*mz box; local (ghosted) vectors carry a width-3
// ghost ring only in directions declared periodic (PETSc clips ghosts at non-periodic ends).
#include "petsc_shim.h"
#include <stdarg.h>
#include <time.h>
#include <vector>

struct _p_DA {
  int mx, my, mz, dof, sw;
  int wx, wy, wz;          // wrap flags
  DAPeriodicType pt;
  DA cda;                  // dof-3 companion ("coordinate DA" == user->fda)
  Vec coords;              // ghosted coordinates (local vector on cda)
};
struct _p_Vec {
  DA da; int is_local; long n; double *a;
};

extern "C" DA shim_da_create(int mx, int my, int mz, int dof, int wx, int wy, int wz) {
  DA d = (DA)calloc(1, sizeof(_p_DA));
  d->mx = mx; d->my = my; d->mz = mz; d->dof = dof; d->sw = 3; d->wx = wx; d->wy = wy; d->wz = wz;
  if (wx && wy && wz) d->pt = DA_XYZPERIODIC; else if (wx && wy) d->pt = DA_XYPERIODIC;
  else if (wy && wz) d->pt = DA_YZPERIODIC; else if (wx && wz) d->pt = DA_XZPERIODIC;
  else if (wx) d->pt = DA_XPERIODIC; else if (wy) d->pt = DA_YPERIODIC;
  else if (wz) d->pt = DA_ZPERIODIC; else d->pt = DA_NONPERIODIC;
  return d;
}
extern "C" void shim_da_set_cda(DA d, DA cda) { d->cda = cda; }

static inline void ext(DA d, int loc, int &gxs, int &gys, int &gzs, int &gxm, int &gym, int &gzm) {
  gxs = (loc && d->wx) ? -d->sw : 0; gxm = d->mx + ((loc && d->wx) ? 2 * d->sw : 0);
  gys = (loc && d->wy) ? -d->sw : 0; gym = d->my + ((loc && d->wy) ? 2 * d->sw : 0);
  gzs = (loc && d->wz) ? -d->sw : 0; gzm = d->mz + ((loc && d->wz) ? 2 * d->sw : 0);
}
static Vec vec_new(DA d, int loc) {
  int a, b, c, l, m, n; ext(d, loc, a, b, c, l, m, n);
  Vec v = (Vec)calloc(1, sizeof(_p_Vec));
  v->da = d; v->is_local = loc; v->n = (long)l * m * n * d->dof;
  v->a = (double *)calloc(v->n, sizeof(double));
  return v;
}
extern "C" double *shim_vec_data(Vec v) { return v->a; }
extern "C" long shim_vec_size(Vec v) { return v->n; }
extern "C" int shim_vec_is_local(Vec v) { return v->is_local; }
extern "C" int shim_vec_dof(Vec v) { return v->da->dof; }

PetscErrorCode DACreateGlobalVector(DA d, Vec *v) { *v = vec_new(d, 0); return 0; }
PetscErrorCode DACreateLocalVector(DA d, Vec *v) { *v = vec_new(d, 1); return 0; }
PetscErrorCode DAGetLocalVector(DA d, Vec *v) { *v = vec_new(d, 1); return 0; }
PetscErrorCode DARestoreLocalVector(DA, Vec *v) { VecDestroy(*v); *v = 0; return 0; }
PetscErrorCode VecDuplicate(Vec x, Vec *y) { *y = vec_new(x->da, x->is_local); return 0; }
PetscErrorCode VecDestroy(Vec v) { if (v) { free(v->a); free(v); } return 0; }

PetscErrorCode DAGetLocalInfo(DA d, DALocalInfo *i) {
  memset(i, 0, sizeof(*i));
  i->dim = 3; i->dof = d->dof; i->sw = d->sw; i->mx = d->mx; i->my = d->my; i->mz = d->mz;
  i->xs = i->ys = i->zs = 0; i->xm = d->mx; i->ym = d->my; i->zm = d->mz;
  ext(d, 1, i->gxs, i->gys, i->gzs, i->gxm, i->gym, i->gzm);
  i->pt = d->pt; i->st = DA_STENCIL_BOX; i->da = d;
  return 0;
}
PetscErrorCode DAGetCoordinateDA(DA d, DA *c) { *c = d->cda; return 0; }
PetscErrorCode DAGetGhostedCoordinates(DA d, Vec *c) {
  if (!d->coords) d->coords = vec_new(d->cda ? d->cda : d, 1);
  *c = d->coords; return 0;
}
PetscErrorCode DAGetCoordinates(DA d, Vec *c) { return DAGetGhostedCoordinates(d, c); }

// T*** tables: one allocation holding gzm plane pointers followed by gzm*gym row pointers.
PetscErrorCode DAVecGetArray(DA d, Vec v, void *out) {
  int gxs, gys, gzs, gxm, gym, gzm; ext(d, v->is_local, gxs, gys, gzs, gxm, gym, gzm);
  if ((long)gxm * gym * gzm * d->dof != v->n) { fprintf(stderr, "shim: DAVecGetArray size mismatch\n"); abort(); }
  char ***planes = (char ***)malloc(sizeof(char **) * gzm + sizeof(char *) * (size_t)gzm * gym);
  char **rows = (char **)(planes + gzm);
  size_t rowb = (size_t)gxm * d->dof * sizeof(double);
  for (int k = 0; k < gzm; k++) {
    planes[k] = rows + (size_t)k * gym - gys;
    for (int j = 0; j < gym; j++)
      rows[(size_t)k * gym + j] = (char *)v->a + ((size_t)k * gym + j) * rowb - (ptrdiff_t)gxs * d->dof * (ptrdiff_t)sizeof(double);
  }
  *(char ****)out = planes - gzs;
  return 0;
}
PetscErrorCode DAVecRestoreArray(DA d, Vec v, void *out) {
  int gxs, gys, gzs, gxm, gym, gzm; ext(d, v->is_local, gxs, gys, gzs, gxm, gym, gzm);
  char ***p = *(char ****)out; free(p + gzs); *(char ****)out = 0; return 0;
}
static inline int wrapi(int i, int m) { return i < 0 ? i + m : (i >= m ? i - m : i); }
static void fill(DA d, const double *src, int src_local, double *dst, int dst_local, int ghosts_only) {
  int sxs, sys, szs, sxm, sym, szm; ext(d, src_local, sxs, sys, szs, sxm, sym, szm);
  int gxs, gys, gzs, gxm, gym, gzm; ext(d, dst_local, gxs, gys, gzs, gxm, gym, gzm);
  int dof = d->dof;
  for (int k = gzs; k < gzs + gzm; k++) for (int j = gys; j < gys + gym; j++) for (int i = gxs; i < gxs + gxm; i++) {
    int in = (i >= 0 && i < d->mx && j >= 0 && j < d->my && k >= 0 && k < d->mz);
    if (ghosts_only && in) continue;
    int a = wrapi(i, d->mx), b = wrapi(j, d->my), c = wrapi(k, d->mz);
    const double *s = src + (((size_t)(c - szs) * sym + (b - sys)) * sxm + (a - sxs)) * dof;
    double *t = dst + (((size_t)(k - gzs) * gym + (j - gys)) * gxm + (i - gxs)) * dof;
    for (int q = 0; q < dof; q++) t[q] = s[q];
  }
}
PetscErrorCode DAGlobalToLocalBegin(DA d, Vec g, InsertMode, Vec l) { fill(d, g->a, g->is_local, l->a, l->is_local, 0); return 0; }
PetscErrorCode DAGlobalToLocalEnd(DA, Vec, InsertMode, Vec) { return 0; }
PetscErrorCode DALocalToLocalBegin(DA d, Vec a, InsertMode, Vec b) {
  if (a == b) fill(d, a->a, 1, b->a, 1, 1);       // refresh ghosts from owned values
  else fill(d, a->a, a->is_local, b->a, b->is_local, 0);
  return 0;
}
PetscErrorCode DALocalToLocalEnd(DA, Vec, InsertMode, Vec) { return 0; }
PetscErrorCode DALocalToGlobal(DA d, Vec l, InsertMode, Vec g) { fill(d, l->a, l->is_local, g->a, g->is_local, 0); return 0; }

PetscErrorCode VecSet(Vec v, PetscScalar s) { for (long i = 0; i < v->n; i++) v->a[i] = s; return 0; }
PetscErrorCode VecCopy(Vec x, Vec y) { if (x->n != y->n) abort(); memcpy(y->a, x->a, sizeof(double) * x->n); return 0; }
PetscErrorCode VecAXPY(Vec y, PetscScalar al, Vec x) { if (x->n != y->n) abort(); for (long i = 0; i < y->n; i++) y->a[i] += al * x->a[i]; return 0; }
PetscErrorCode VecWAXPY(Vec w, PetscScalar al, Vec x, Vec y) { for (long i = 0; i < w->n; i++) w->a[i] = al * x->a[i] + y->a[i]; return 0; }
PetscErrorCode VecScale(Vec v, PetscScalar s) { for (long i = 0; i < v->n; i++) v->a[i] *= s; return 0; }
PetscErrorCode VecMax(Vec v, PetscInt *p, PetscReal *r) { long b = 0; for (long i = 1; i < v->n; i++) if (v->a[i] > v->a[b]) b = i; if (p) *p = (int)b; *r = v->a[b]; return 0; }
PetscErrorCode VecMin(Vec v, PetscInt *p, PetscReal *r) { long b = 0; for (long i = 1; i < v->n; i++) if (v->a[i] < v->a[b]) b = i; if (p) *p = (int)b; *r = v->a[b]; return 0; }
PetscErrorCode VecNorm(Vec v, NormType t, PetscReal *r) {
  double s = 0; if (t == NORM_INFINITY) { for (long i = 0; i < v->n; i++) s = fmax(s, fabs(v->a[i])); *r = s; }
  else if (t == NORM_1) { for (long i = 0; i < v->n; i++) s += fabs(v->a[i]); *r = s; }
  else { for (long i = 0; i < v->n; i++) s += v->a[i] * v->a[i]; *r = sqrt(s); } return 0;
}
PetscErrorCode VecAssemblyBegin(Vec) { return 0; }
PetscErrorCode VecAssemblyEnd(Vec) { return 0; }
PetscErrorCode VecGetArray(Vec v, PetscScalar **a) { *a = v->a; return 0; }
PetscErrorCode VecRestoreArray(Vec, PetscScalar **) { return 0; }
PetscErrorCode VecGetSize(Vec v, PetscInt *n) { *n = (int)v->n; return 0; }
PetscErrorCode PetscGlobalMax(PetscReal *l, PetscReal *g, MPI_Comm) { *g = *l; return 0; }
PetscErrorCode PetscGlobalMin(PetscReal *l, PetscReal *g, MPI_Comm) { *g = *l; return 0; }
PetscErrorCode PetscGlobalSum(PetscScalar *l, PetscScalar *g, MPI_Comm) { *g = *l; return 0; }
extern "C" int shim_verbose = 0;
PetscErrorCode PetscPrintf(MPI_Comm, const char *f, ...) { if (shim_verbose) { va_list ap; va_start(ap, f); vprintf(f, ap); va_end(ap); } return 0; }
PetscErrorCode PetscFPrintf(MPI_Comm, FILE *fp, const char *f, ...) { va_list ap; va_start(ap, f); vfprintf(fp, f, ap); va_end(ap); return 0; }
PetscErrorCode PetscBarrier(void *) { return 0; }
PetscErrorCode PetscGetTime(PetscLogDouble *t) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); *t = ts.tv_sec + 1e-9 * ts.tv_nsec; return 0; }
PetscErrorCode PetscOptionsGetReal(const char *, const char *, PetscReal *, PetscTruth *f) { if (f) *f = PETSC_FALSE; return 0; }
PetscErrorCode PetscOptionsGetInt(const char *, const char *, PetscInt *, PetscTruth *f) { if (f) *f = PETSC_FALSE; return 0; }
PetscErrorCode PetscMalloc(size_t n, void *p) { *(void **)p = malloc(n); return 0; }
PetscErrorCode PetscFree(void *p) { free(p); return 0; }
int MPI_Comm_rank(MPI_Comm, int *r) { *r = 0; return 0; }
int MPI_Comm_size(MPI_Comm, int *s) { *s = 1; return 0; }
static size_t tsize(MPI_Datatype t) { return t == MPI_INT ? sizeof(int) : sizeof(double); }
int MPI_Allreduce(void *s, void *r, int n, MPI_Datatype t, MPI_Op, MPI_Comm) { if (s != r) memcpy(r, s, n * tsize(t)); return 0; }
int MPI_Reduce(void *s, void *r, int n, MPI_Datatype t, MPI_Op, int, MPI_Comm) { if (s != r) memcpy(r, s, n * tsize(t)); return 0; }
int MPI_Bcast(void *, int, MPI_Datatype, int, MPI_Comm) { return 0; }
int MPI_Barrier(MPI_Comm) { return 0; }
