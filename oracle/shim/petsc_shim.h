/* Single-rank stand-in for the PETSc 3.1 / MPI / HYPRE declarations the VFS-Wind hot-path
 * sources use.  TEST INFRASTRUCTURE ONLY: it exists so the unmodified reference sources under
 * /root/reference/Source can be compiled into oracle/_ref/libvfsref.so (the parity checker and
 * CPU baseline).  Nothing in the product path includes or links this.
 * Semantics follow SURVEY.md section 8(c): one rank, DA ghost width 3, wrap only when periodic. */
#ifndef VFS_PETSC_SHIM_H
#define VFS_PETSC_SHIM_H
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

typedef int PetscInt;
typedef double PetscReal;
typedef double PetscScalar;
typedef int PetscErrorCode;
typedef int PetscMPIInt;
typedef enum { PETSC_FALSE, PETSC_TRUE } PetscTruth;
typedef double PetscLogDouble;
#define PETSC_NULL 0
#define PETSC_DEFAULT (-2)
#define PETSC_DECIDE (-1)
#define PetscMax(a,b) (((a)<(b)) ? (b) : (a))
#define PetscMin(a,b) (((a)<(b)) ? (a) : (b))
#define CHKERRQ(e) do { if (e) return (e); } while (0)

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
#define MPI_COMM_WORLD 0
#define PETSC_COMM_WORLD 0
#define PETSC_COMM_SELF 1
#define MPI_DOUBLE 1
#define MPI_INT 2
#define MPI_CHAR 3
#define MPI_SUM 1
#define MPI_MAX 2
#define MPI_MIN 3
#define MPIU_SCALAR MPI_DOUBLE
#define MPIU_REAL MPI_DOUBLE
#define MPIU_INT MPI_INT

struct _p_DA; typedef struct _p_DA *DA;
struct _p_Vec; typedef struct _p_Vec *Vec;
typedef struct _p_Mat *Mat;
typedef struct _p_KSP *KSP;
typedef struct _p_SNES *SNES;
typedef struct _p_PC *PC;
typedef struct _p_AO *AO;
typedef struct _p_IS *IS;
typedef struct _p_MatNullSpace *MatNullSpace;
typedef struct _p_PetscViewer *PetscViewer;
typedef struct _p_VecScatter *VecScatter;
typedef const char *SNESType; typedef const char *KSPType; typedef const char *PCType;
typedef enum { DA_STENCIL_STAR, DA_STENCIL_BOX } DAStencilType;
typedef enum { DA_NONPERIODIC, DA_XPERIODIC, DA_YPERIODIC, DA_XYPERIODIC, DA_XYZPERIODIC,
               DA_XZPERIODIC, DA_YZPERIODIC, DA_ZPERIODIC } DAPeriodicType;
typedef enum { INSERT_VALUES = 1, ADD_VALUES = 2 } InsertMode;
typedef enum { NORM_1 = 0, NORM_2 = 1, NORM_FROBENIUS = 2, NORM_INFINITY = 3 } NormType;
typedef enum { SAME_NONZERO_PATTERN, DIFFERENT_NONZERO_PATTERN, SAME_PRECONDITIONER } MatStructure;
#define SNESTR "tr"
#define SNESLS "ls"
#define KSPGMRES "gmres"
#define PCNONE "none"

typedef struct {
  PetscInt dim, dof, sw;
  PetscInt mx, my, mz;
  PetscInt xs, ys, zs;
  PetscInt xm, ym, zm;
  PetscInt gxs, gys, gzs;
  PetscInt gxm, gym, gzm;
  DAPeriodicType pt;
  DAStencilType st;
  DA da;
} DALocalInfo;

/* --- implemented in shim.cpp --- */
PetscErrorCode DAVecGetArray(DA, Vec, void *);
PetscErrorCode DAVecRestoreArray(DA, Vec, void *);
PetscErrorCode DAGetLocalInfo(DA, DALocalInfo *);
PetscErrorCode DAGlobalToLocalBegin(DA, Vec, InsertMode, Vec);
PetscErrorCode DAGlobalToLocalEnd(DA, Vec, InsertMode, Vec);
PetscErrorCode DALocalToLocalBegin(DA, Vec, InsertMode, Vec);
PetscErrorCode DALocalToLocalEnd(DA, Vec, InsertMode, Vec);
PetscErrorCode DALocalToGlobal(DA, Vec, InsertMode, Vec);
PetscErrorCode DAGetGhostedCoordinates(DA, Vec *);
PetscErrorCode DAGetCoordinates(DA, Vec *);
PetscErrorCode DAGetCoordinateDA(DA, DA *);
PetscErrorCode DAGetLocalVector(DA, Vec *);
PetscErrorCode DARestoreLocalVector(DA, Vec *);
PetscErrorCode DACreateGlobalVector(DA, Vec *);
PetscErrorCode DACreateLocalVector(DA, Vec *);
PetscErrorCode VecDuplicate(Vec, Vec *);
PetscErrorCode VecDestroy(Vec);
PetscErrorCode VecSet(Vec, PetscScalar);
PetscErrorCode VecCopy(Vec, Vec);
PetscErrorCode VecAXPY(Vec, PetscScalar, Vec);
PetscErrorCode VecWAXPY(Vec, PetscScalar, Vec, Vec);
PetscErrorCode VecScale(Vec, PetscScalar);
PetscErrorCode VecMax(Vec, PetscInt *, PetscReal *);
PetscErrorCode VecMin(Vec, PetscInt *, PetscReal *);
PetscErrorCode VecNorm(Vec, NormType, PetscReal *);
PetscErrorCode VecAssemblyBegin(Vec);
PetscErrorCode VecAssemblyEnd(Vec);
PetscErrorCode VecGetArray(Vec, PetscScalar **);
PetscErrorCode VecRestoreArray(Vec, PetscScalar **);
PetscErrorCode VecGetSize(Vec, PetscInt *);
PetscErrorCode PetscGlobalMax(PetscReal *, PetscReal *, MPI_Comm);
PetscErrorCode PetscGlobalMin(PetscReal *, PetscReal *, MPI_Comm);
PetscErrorCode PetscGlobalSum(PetscScalar *, PetscScalar *, MPI_Comm);
PetscErrorCode PetscPrintf(MPI_Comm, const char *, ...);
PetscErrorCode PetscFPrintf(MPI_Comm, FILE *, const char *, ...);
PetscErrorCode PetscBarrier(void *);
PetscErrorCode PetscGetTime(PetscLogDouble *);
PetscErrorCode PetscOptionsGetReal(const char *, const char *, PetscReal *, PetscTruth *);
PetscErrorCode PetscOptionsGetInt(const char *, const char *, PetscInt *, PetscTruth *);
PetscErrorCode PetscMalloc(size_t, void *);
PetscErrorCode PetscFree(void *);
int MPI_Comm_rank(MPI_Comm, int *);
int MPI_Comm_size(MPI_Comm, int *);
int MPI_Allreduce(void *, void *, int, MPI_Datatype, MPI_Op, MPI_Comm);
int MPI_Reduce(void *, void *, int, MPI_Datatype, MPI_Op, int, MPI_Comm);
int MPI_Bcast(void *, int, MPI_Datatype, int, MPI_Comm);
int MPI_Barrier(MPI_Comm);

/* --- declared only; abort() stubs generated at link time (solver callers, out of scope) --- */
PetscErrorCode SNESCreate(MPI_Comm, SNES *);
PetscErrorCode SNESDestroy(SNES);
PetscErrorCode SNESSetFunction(SNES, Vec, PetscErrorCode (*)(SNES, Vec, Vec, void *), void *);
PetscErrorCode SNESSetJacobian(SNES, Mat, Mat, PetscErrorCode (*)(SNES, Vec, Mat *, Mat *, MatStructure *, void *), void *);
PetscErrorCode SNESSetType(SNES, SNESType);
PetscErrorCode SNESSetTolerances(SNES, PetscReal, PetscReal, PetscReal, PetscInt, PetscInt);
PetscErrorCode SNESSetMaxLinearSolveFailures(SNES, PetscInt);
PetscErrorCode SNESSetMaxNonlinearStepFailures(SNES, PetscInt);
PetscErrorCode SNESKSPSetUseEW(SNES, PetscTruth);
PetscErrorCode SNESKSPSetParametersEW(SNES, PetscInt, PetscReal, PetscReal, PetscReal, PetscReal, PetscReal, PetscReal);
PetscErrorCode SNESGetKSP(SNES, KSP *);
PetscErrorCode SNESMonitorSet(SNES, PetscErrorCode (*)(SNES, PetscInt, PetscReal, void *), void *, PetscErrorCode (*)(void *));
PetscErrorCode SNESSolve(SNES, Vec, Vec);
PetscErrorCode SNESGetFunctionNorm(SNES, PetscReal *);
PetscErrorCode MatCreateSNESMF(SNES, Mat *);
PetscErrorCode MatMFFDComputeJacobian(SNES, Vec, Mat *, Mat *, MatStructure *, void *);
PetscErrorCode MatDestroy(Mat);
PetscErrorCode KSPSetType(KSP, KSPType);
PetscErrorCode KSPGetPC(KSP, PC *);
PetscErrorCode KSPSetTolerances(KSP, PetscReal, PetscReal, PetscReal, PetscInt);
PetscErrorCode KSPGMRESSetPreAllocateVectors(KSP);
PetscErrorCode PCSetType(PC, PCType);

/* --- poisson.c: only Projection / UpdatePressure (poisson.c:2700, 3137) are driven by the harness.  The Poisson matrix
 * assembly and the KSP / multigrid set-up in the same file never run here; the PETSc calls they make are declared as
 * catch-all stubs that abort if ever reached. --- */
#ifdef __cplusplus
#include <stdlib.h>
#define VFS_SHIM_STUB(name) template <class... A> inline PetscErrorCode name(A...) { abort(); return 0; }
VFS_SHIM_STUB(DACreateNaturalVector) VFS_SHIM_STUB(DAGetGlobalIndices) VFS_SHIM_STUB(DAGlobalToNaturalBegin) VFS_SHIM_STUB(DAGlobalToNaturalEnd)
VFS_SHIM_STUB(KSPAppendOptionsPrefix) VFS_SHIM_STUB(KSPBuildResidual) VFS_SHIM_STUB(KSPCreate) VFS_SHIM_STUB(KSPDestroy) VFS_SHIM_STUB(KSPGMRESSetRestart)
VFS_SHIM_STUB(KSPMonitorSet) VFS_SHIM_STUB(KSPSetFromOptions) VFS_SHIM_STUB(KSPSetInitialGuessNonzero) VFS_SHIM_STUB(KSPSetNullSpace) VFS_SHIM_STUB(KSPSetOperators)
VFS_SHIM_STUB(KSPSetUp) VFS_SHIM_STUB(KSPSolve) VFS_SHIM_STUB(MatAssemblyBegin) VFS_SHIM_STUB(MatAssemblyEnd) VFS_SHIM_STUB(MatCreate) VFS_SHIM_STUB(MatCreateShell)
VFS_SHIM_STUB(MatGetVecs) VFS_SHIM_STUB(MatMPIAIJSetPreallocation) VFS_SHIM_STUB(MatMult) VFS_SHIM_STUB(MatNullSpaceCreate) VFS_SHIM_STUB(MatNullSpaceDestroy)
VFS_SHIM_STUB(MatNullSpaceSetFunction) VFS_SHIM_STUB(MatSetFromOptions) VFS_SHIM_STUB(MatSetSizes) VFS_SHIM_STUB(MatSetType) VFS_SHIM_STUB(MatSetValues)
VFS_SHIM_STUB(MatShellGetContext) VFS_SHIM_STUB(MatShellSetOperation) VFS_SHIM_STUB(MatZeroEntries) VFS_SHIM_STUB(PCBJacobiGetSubKSP) VFS_SHIM_STUB(PCFactorSetShiftAmount)
VFS_SHIM_STUB(PCFactorSetShiftType) VFS_SHIM_STUB(PCHYPRESetType) VFS_SHIM_STUB(PCMGGetCoarseSolve) VFS_SHIM_STUB(PCMGGetSmoother) VFS_SHIM_STUB(PCMGSetCycleType)
VFS_SHIM_STUB(PCMGSetInterpolation) VFS_SHIM_STUB(PCMGSetLevels) VFS_SHIM_STUB(PCMGSetResidual) VFS_SHIM_STUB(PCMGSetRestriction) VFS_SHIM_STUB(PCMGSetRhs)
VFS_SHIM_STUB(PCMGSetType) VFS_SHIM_STUB(PCSetFromOptions) VFS_SHIM_STUB(PCSetOperators) VFS_SHIM_STUB(PCSetUp) VFS_SHIM_STUB(PetscOptionsInsertString)
VFS_SHIM_STUB(VecCreateMPI) VFS_SHIM_STUB(VecGetArray3d) VFS_SHIM_STUB(VecGetLocalSize) VFS_SHIM_STUB(VecRestoreArray3d) VFS_SHIM_STUB(VecScatterBegin)
VFS_SHIM_STUB(VecScatterCreateToZero) VFS_SHIM_STUB(VecScatterDestroy) VFS_SHIM_STUB(VecScatterEnd) VFS_SHIM_STUB(VecSetValue) VFS_SHIM_STUB(VecShift) VFS_SHIM_STUB(VecSum)
inline PetscErrorCode KSPMonitorTrueResidualNorm(KSP, PetscInt, PetscReal, void *) { abort(); return 0; }
inline PetscErrorCode PCMGDefaultResidual(Mat, Vec, Vec, Vec) { abort(); return 0; }
#endif
#define KSPFGMRES "fgmres"
#define MATMPIAIJ "mpiaij"
#define PCBJACOBI "bjacobi"
#define PCHYPRE "hypre"
#define PCMG "mg"
enum { MATOP_MULT = 3, MATOP_MULT_ADD = 4, MAT_FINAL_ASSEMBLY = 0, MAT_SHIFT_NONZERO = 1, PC_MG_CYCLE_V = 1, PC_MG_MULTIPLICATIVE = 0, PETSC_DETERMINE = -1, SCATTER_FORWARD = 0 };

/* HYPRE handles referenced by prototypes in variables.h */
typedef struct hypre_s1 *HYPRE_IJMatrix; typedef struct hypre_s2 *HYPRE_IJVector;
typedef struct hypre_s3 *HYPRE_ParCSRMatrix; typedef struct hypre_s4 *HYPRE_ParVector;
typedef struct hypre_s5 *HYPRE_Solver;
#endif
