// vfs_common.h — device-side grid/field model shared by all kernels.
//
// HBM layout: every scalar plane set is a padded 3-D array
//   [nzl + 2G][my + 2G][pitch]   (k slowest, i fastest, FP64)
// G = 4 ghost layers on every side.  Vector fields (Cmpnts in the reference) are split into three
// such scalars (SoA), so loads along i are unit-stride and coalesced.  Logical index (i,j,k) with
// k LOCAL to the rank's slab (global k = k + kofs) maps to ((k+G)*ny + (j+G))*pitch + (i+G).
// Ghosts hold DA-wrap images in periodic directions / neighbour-rank planes in k, exactly like
// the reference's ghosted local Vecs (width 3 there, init.c:131-160), so reference index
// arithmetic such as ucat[k][j][-3] carries over literally.
#ifndef VFS_COMMON_H
#define VFS_COMMON_H
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define VFS_HD __host__ __device__ __forceinline__
#else
#define VFS_HD inline
#endif

#define VFS_G 4

// warp-uniform "does any active lane see x": the mask-free fast paths branch on it so a warp never
// diverges; the host emulation (one "thread" at a time) just tests its own value
#if defined(__CUDA_ARCH__)
#define VFS_WARP_ANY(x) (__any_sync(__activemask(), (x)) != 0)
#else
#define VFS_WARP_ANY(x) (x)
#endif

// internal scalar ids ------------------------------------------------------------------------
enum {
  S_X = 0, S_Y, S_Z,                                  // node coordinates
  S_CSI0, S_CSI1, S_CSI2, S_ETA0, S_ETA1, S_ETA2, S_ZET0, S_ZET1, S_ZET2, S_AJ, S_NV,
  S_UC0, S_UC1, S_UC2,                                // ucont
  S_U0, S_U1, S_U2,                                   // ucat
  S_UO0, S_UO1, S_UO2,                                // ucat_old
  S_UCO0, S_UCO1, S_UCO2,                             // ucont_o
  S_RO0, S_RO1, S_RO2,                                // rhs_o
  S_DP0, S_DP1, S_DP2,                                // dP
  S_FE0, S_FE1, S_FE2,                                // F_eul
  S_R0, S_R1, S_R2,                                   // rhs
  S_CS, S_NUT, S_USTAR,
  // work (never cross the ABI)
  S_FC1, S_FC1b, S_FC1c, S_FC2, S_FC2b, S_FC2c, S_FC3, S_FC3b, S_FC3c,   // convective face fluxes (Div1-3)
  S_FV1, S_FV1b, S_FV1c, S_FV2, S_FV2b, S_FV2c, S_FV3, S_FV3b, S_FV3c,   // viscous+SGS face fluxes (Visc1-3)
  S_FP0, S_FP1, S_FP2,                                                   // Fp
  S_SABS,                                                                // LES: |S|
  S_UF0, S_UF1, S_UF2,                                                   // LES: test-filtered ucat
  S_LM, S_MM,
  // LES per-node derived quantities entering the test filters (contiguous: w, U(3), |S|S_ij(6))
  S_LW, S_LU0, S_LU1, S_LU2, S_LSS0, S_LSS1, S_LSS2, S_LSS3, S_LSS4, S_LSS5,
  S_IAJ,                                                                 // 1/aj (cell volume), filter weight
  // LES factors that depend on the grid and the nvert mask only (LesGeo, vfs_les_kernels.h):
  // 1/sum(s*w), test_filter^2, filter^2
  S_LFINV, S_LTF2, S_LF2,
  S_WM,                                                                  // wall-model nu_t of the j = 0 faces (plane j = 0 only)
  S_P,                                                                   // pressure (input of Pressure_Gradient, momentum.c:203)
  // rarely used scalars: allocated on first use, outside the main pool (9 of 93 scalars = 11 GB on a 2048 x 1024 x 64 slab)
  S_TAIL0,
  S_UCM0 = S_TAIL0, S_UCM1, S_UCM2,                                      // ucont_rm1 (BDF2 assembly only)
  S_CONV0, S_CONV1, S_CONV2, S_VISC0, S_VISC1, S_VISC2,                  // legacy Convection / Viscous results (rhs.c:751,1071)
  S_ADV1, S_ADV1b, S_ADV1c, S_ADV2, S_ADV2b, S_ADV2c, S_ADV3, S_ADV3b, S_ADV3c,   // advective half of the skew-symmetric form (Adv1-3, momentum.c:607-615)
  S_PHI,                                                                 // pressure correction (Phi / lPhi), input of UpdatePressure / Projection (poisson.c:3137, 2700)
  S_GR0, S_GR1, S_GR2, S_GR3, S_GR4, S_GR5, S_GR6, S_GR7, S_GR8,         // clark: du_a/dx_b at the cell centres (lSx, lSy, lSz of les.c:181-300), a-major
  S_COUNT
};

struct VfsDev {
  // geometry
  int mx, my, mz, nzl, kofs;
  int pitch, ny, nzt;
  long sj, sk;            // strides in doubles: sj = pitch, sk = ny*pitch
  long org;               // offset of logical (0,0,0)
  // switches (vfs_params)
  int perx, pery, perz;
  int bc[6];
  int les, second_order, laplacian, immersed, clark, testfilter_ik, visc_wm, wallfunction, has_feul;
  int ti, tistart, rstart_flg, bdf2, single_rank;
  int legx, legy, legz;     // the direction is periodic through the legacy i/j/k_periodic switches (explicit index remaps in the reference, no DA wrap): same ghost images, one difference in IB_BC (IbBcBoundary)
  int weno, skew, inviscid; // variants of the staged flux path only (face_flux_core<.., X = true>): WENO3 convection (inviscid or levelset_weno == 5), skew-symmetric form, no viscous term
  int homo;               // homogeneous-direction averaging of LM/MM (les.c:798-965): 0 off, 1 = i and k, 2 = i, 3 = j, 4 = k
  double ren, dt, max_cs, roughness;
  double *s[S_COUNT];
  // near[p] != 0: some node within +-2 of p (any direction) has nvert != 0 (NearSolid, vfs_c2c_kernels.h).
  // Warps whose nodes are all "far" run mask-free specialisations of the stencil code (same arithmetic:
  // every nvert comparison is known to be false), see VFS_WARP_ANY.
  const unsigned char *near;
  VFS_HD long idx(int i, int j, int k) const { return org + (long)k * sk + (long)j * sj + i; }
  // global k of local plane k; a ghost plane across the periodic seam maps to the plane it images
  // (-1 -> mz-1, mz -> 0, ...), which is what the owner of that plane evaluates its boundary logic with
  VFS_HD int kglob(int k) const { int kg = k + kofs; if (perz) { if (kg < 0) kg += mz; else if (kg >= mz) kg -= mz; } return kg; }
};

struct Box { int i0, i1, j0, j1, k0, k1; };

// 3-component helpers -------------------------------------------------------------------------
struct V3 { double x, y, z; };
VFS_HD V3 ld3(const VfsDev &d, int s0, long p) { V3 v; v.x = d.s[s0][p]; v.y = d.s[s0 + 1][p]; v.z = d.s[s0 + 2][p]; return v; }
VFS_HD void st3(const VfsDev &d, int s0, long p, const V3 &v) { d.s[s0][p] = v.x; d.s[s0 + 1][p] = v.y; d.s[s0 + 2][p] = v.z; }
VFS_HD V3 mk3(double x, double y, double z) { V3 v; v.x = x; v.y = y; v.z = z; return v; }
VFS_HD double dot3(const V3 &a, const V3 &b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

#endif
