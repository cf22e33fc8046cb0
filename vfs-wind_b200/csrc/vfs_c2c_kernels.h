// vfs_c2c_kernels.h — Contra2Cart_2 (Source/rhs.c:65-749) and IB_BC (Source/momentum.c:2016-2235).
#ifndef VFS_C2C_KERNELS_H
#define VFS_C2C_KERNELS_H
#include "vfs_common.h"

// rhs.c:158-247: interior cells with nvert < 0.1: q = face-averaged fluxes, solve
// [csi;eta;zet] u = q by Cramer's rule.  Other cells keep their value (IBM state).
struct C2CInterior {
  VfsDev d;
  VFS_HD void operator()(int i, int j, int k) const {
    long p = d.idx(i, j, k);
    { const int kg = d.kglob(k); if (kg < 1 || kg > d.mz - 2) return; }      // see C2CInteriorFix
    if (!(d.s[S_NV][p] < 0.1)) return;
    const double m00 = d.s[S_CSI0][p], m01 = d.s[S_CSI1][p], m02 = d.s[S_CSI2][p];
    const double m10 = d.s[S_ETA0][p], m11 = d.s[S_ETA1][p], m12 = d.s[S_ETA2][p];
    const double m20 = d.s[S_ZET0][p], m21 = d.s[S_ZET1][p], m22 = d.s[S_ZET2][p];
    const double q0 = 0.5 * (d.s[S_UC0][p - 1] + d.s[S_UC0][p]);
    const double q1 = 0.5 * (d.s[S_UC1][p - d.sj] + d.s[S_UC1][p]);
    const double q2 = 0.5 * (d.s[S_UC2][p - d.sk] + d.s[S_UC2][p]);
    const double det = m00 * (m11 * m22 - m12 * m21) - m01 * (m10 * m22 - m12 * m20) + m02 * (m10 * m21 - m11 * m20);
    const double det0 = q0 * (m11 * m22 - m12 * m21) - q1 * (m01 * m22 - m02 * m21) + q2 * (m01 * m12 - m02 * m11);
    const double det1 = -q0 * (m10 * m22 - m12 * m20) + q1 * (m00 * m22 - m02 * m20) - q2 * (m00 * m12 - m02 * m10);
    const double det2 = q0 * (m10 * m21 - m11 * m20) - q1 * (m00 * m21 - m01 * m20) + q2 * (m00 * m11 - m01 * m10);
    d.s[S_U0][p] = det0 / det; d.s[S_U1][p] = det1 / det; d.s[S_U2][p] = det2 / det;
  }
};

// rhs2.c:595-611 inverse of [csi;eta;zet]; column `col` (0: x_csi.., 1: x_eta.., 2: x_zet..)
VFS_HD V3 cov_column(const VfsDev &d, long p, int col) {
  const double a11 = d.s[S_CSI0][p], a12 = d.s[S_CSI1][p], a13 = d.s[S_CSI2][p];
  const double a21 = d.s[S_ETA0][p], a22 = d.s[S_ETA1][p], a23 = d.s[S_ETA2][p];
  const double a31 = d.s[S_ZET0][p], a32 = d.s[S_ZET1][p], a33 = d.s[S_ZET2][p];
  const double det = a11 * (a33 * a22 - a32 * a23) - a21 * (a33 * a12 - a32 * a13) + a31 * (a23 * a12 - a22 * a13);
  if (col == 0) return mk3((a33 * a22 - a32 * a23) / det, -(a33 * a21 - a31 * a23) / det, (a32 * a21 - a31 * a22) / det);
  if (col == 1) return mk3(-(a33 * a12 - a32 * a13) / det, (a33 * a11 - a31 * a13) / det, -(a32 * a11 - a31 * a12) / det);
  return mk3((a23 * a12 - a22 * a13) / det, -(a23 * a11 - a21 * a13) / det, (a22 * a11 - a21 * a12) / det);
}

// slip wall (bctype 10): reflect the OLD interior velocity about the wall normal (rhs.c:454-535)
VFS_HD V3 slip_ghost(const VfsDev &d, long pn, int col, double sgn) {
  V3 n = cov_column(d, pn, col);
  double nx = sgn * n.x, ny = sgn * n.y, nz = sgn * n.z;
  double sum = sqrt(nx * nx + ny * ny + nz * nz);
  nx /= sum, ny /= sum, nz /= sum;
  V3 uo = ld3(d, S_UO0, pn);
  double un = uo.x * nx + uo.y * ny + uo.z * nz;
  V3 r = mk3(uo.x - 2 * un * nx, uo.y - 2 * un * ny, uo.z - 2 * un * nz);
  if (d.s[S_NV][pn] > 0.1) r = mk3(0, 0, 0);
  return r;
}

// rhs.c:302-682 restricted to the domain-boundary nodes (i, j or global k equal to 0 or m-1).
// Every rule reads interior neighbours only (the reference reads the lUcat snapshot taken
// before this loop), so running it in place over boundary nodes is race-free.
// Wall-function boundary types (-1/-2) are rejected at vfs_create.
struct C2CGhostRules {
  VfsDev d;
  int jtop;      // -1: every boundary node; 0: all but the j = my-1 plane; 1: the j = my-1 plane only (see below)
  VFS_HD void operator()(int i, int j, int k) const {
    const int mx = d.mx, my = d.my, mz = d.mz, kg = d.kglob(k);
    if (!(i == 0 || i == mx - 1 || j == 0 || j == my - 1 || kg == 0 || kg == mz - 1)) return;
    if (jtop >= 0 && (j == my - 1) != (jtop == 1)) return;
    long p = d.idx(i, j, k);
    const int *bc = d.bc;
    // neighbours of an edge/corner node are boundary nodes themselves: read them from the
    // snapshot (S_FP0..2) taken before this kernel, as the reference reads its lUcat copy
    const int nb = (i == 0 || i == mx - 1) + (j == 0 || j == my - 1) + (kg == 0 || kg == mz - 1);
    const int SU = nb >= 2 ? S_FP0 : S_U0;
    if ((int)(d.s[S_NV][p] + 0.1) == 3) { st3(d, S_U0, p, mk3(0, 0, 0)); return; }
    bool w = false; V3 u = mk3(0, 0, 0);
    // the mirror rules 13 / 14 are the only ones that read the array being written (ucat[k][j-1][i], rhs.c:443-452),
    // i.e. the j-1 node AFTER its own rules (solid / wall-function / ghost rules): the host runs the j = my-1 plane
    // in a second pass, after the interior kernels, when one of them is active
    if (bc[3] == 13 && j == my - 1) { V3 a = ld3(d, S_U0, p - d.sj); u = mk3(a.x, -a.y, a.z); w = true; }
    if (bc[3] == 14 && j == my - 1) { V3 a = ld3(d, S_U0, p - d.sj); u = mk3(a.x, a.y, -a.z); w = true; }
    if (bc[0] == 10 && i == 0 && j != 0 && kg != 0) { u = slip_ghost(d, p + 1, 0, -1.); w = true; }
    if (bc[1] == 10 && i == mx - 1 && j != 0 && kg != 0) { u = slip_ghost(d, p - 1, 0, 1.); w = true; }
    if (bc[2] == 10 && j == 0 && i != 0 && kg != 0) { u = slip_ghost(d, p + d.sj, 1, -1.); w = true; }
    if ((bc[3] == 10 || bc[3] == -10) && j == my - 1 && i != 0 && kg != 0) { u = slip_ghost(d, p - d.sj, 1, 1.); w = true; }
    bool solid_flag = false;
    if (i == 0 && bc[0] == 1 && j != 0 && kg != 0) { V3 a = ld3(d, SU, p + 1); u = mk3(-a.x, -a.y, -a.z); w = true; solid_flag = true; }
    if (i == mx - 1 && bc[1] == 1 && j != 0 && kg != 0) { V3 a = ld3(d, SU, p - 1); u = mk3(-a.x, -a.y, -a.z); w = true; solid_flag = true; }
    if (j == 0 && bc[2] == 1 && i != 0 && kg != 0) { V3 a = ld3(d, SU, p + d.sj); u = mk3(-a.x, -a.y, -a.z); w = true; solid_flag = true; }
    if (j == my - 1 && bc[3] == 1 && i != 0 && kg != 0) { V3 a = ld3(d, SU, p - d.sj); u = mk3(-a.x, -a.y, -a.z); w = true; solid_flag = true; }
    if (kg == 0 && bc[4] == 1 && i != 0 && j != 0) { V3 a = ld3(d, SU, p + d.sk); u = mk3(-a.x, -a.y, -a.z); w = true; solid_flag = true; }
    if (kg == mz - 1 && bc[5] == 1 && i != 0 && j != 0) { V3 a = ld3(d, SU, p - d.sk); u = mk3(-a.x, -a.y, -a.z); w = true; solid_flag = true; }
    if (j == my - 1 && bc[3] == 2) {   // cavity lid
      V3 a = solid_flag ? ld3(d, S_UO0, p - d.sj) : ld3(d, SU, p - d.sj);
      u = mk3(2.0 - a.x, -a.y, -a.z); w = true;
    }
    if (j == 0 && bc[2] == 12) { V3 a = ld3(d, SU, p + d.sj); u = mk3(-a.x, -a.y, 2.0 - a.z); if (solid_flag) u = mk3(0, 0, 0); w = true; }
    if (j == my - 1 && bc[3] == 12) { V3 a = ld3(d, SU, p - d.sj); u = mk3(-a.x, -a.y, 2.0 - a.z); if (solid_flag) u = mk3(0, 0, 0); w = true; }
    // body-fitted cylinder, inflow half of the i = 0 side (rhs.c:626-634): cells whose centre lies at z <= 0 mirror to w = 1
    // (no bctype-1 rule can have fired at these nodes — interior j and k, i = 0 — so the reference's solid_flag is false here)
    if (bc[0] == 11 && i == 0 && j != 0 && j != my - 1 && kg != 0 && kg != mz - 1) {
      const double *Z = d.s[S_Z];
      const double zc = (Z[p + 1] + Z[p + 1 - d.sk] + Z[p + 1 - d.sj] + Z[p + 1 - d.sk - d.sj]) * 0.25;
      if (zc <= 0) { V3 a = ld3(d, SU, p + 1); u = mk3(-a.x, -a.y, 2.0 - a.z); w = true; }
    }
    if (bc[3] == 4 && j == my - 1 && i != 0 && i != mx - 1 && kg != 0 && kg != mz - 1) { u = ld3(d, SU, p - d.sj); w = true; }
    if (bc[5] == 4 && kg == mz - 1 && i != 0 && i != mx - 1 && j != 0 && j != my - 1) {
      if (d.s[S_NV][p - d.sk] > 0.1 || solid_flag) { u = mk3(0, 0, 0); w = true; }
    }
    // the corner zeroing of rhs.c:675-681 also hits the k = mz-1 boundary plane (guards: j != my-2, k != 0, k != mz-2)
    if (j != my - 2 && kg != 0 && kg != mz - 2 &&
        ((bc[0] <= 1 && bc[2] <= 1 && i == 1 && j == 1) || (bc[1] <= 1 && bc[2] <= 1 && i == mx - 2 && j == 1) ||
         (bc[0] <= 1 && bc[3] <= 1 && i == 1 && j == my - 2) || (bc[1] <= 1 && bc[3] <= 1 && i == mx - 2 && j == my - 2))) { u = mk3(0, 0, 0); w = true; }
    if (w) st3(d, S_U0, p, u);
  }
};

// interior-node part of the same loop: solid cells -> 0 (rhs.c:305-308) and the "101220" corner
// zeroing (rhs.c:676-681, reproduced with its j!=my-2 && k!=mz-2 guard).
struct C2CInteriorFix {
  VfsDev d;
  VFS_HD void operator()(int i, int j, int k) const {
    const int mx = d.mx, my = d.my, mz = d.mz, kg = d.kglob(k);
    if (kg < 1 || kg > mz - 2) return;          // image of a boundary plane inside an extended k range (see contra2cart)
    long p = d.idx(i, j, k);
    if ((int)(d.s[S_NV][p] + 0.1) == 3) { st3(d, S_U0, p, mk3(0, 0, 0)); return; }
    if (j != my - 2 && kg != mz - 2) {
      const int *bc = d.bc;
      bool z = false;
      if (bc[0] <= 1 && bc[2] <= 1 && i == 1 && j == 1) z = true;
      if (bc[1] <= 1 && bc[2] <= 1 && i == mx - 2 && j == 1) z = true;
      if (bc[0] <= 1 && bc[3] <= 1 && i == 1 && j == my - 2) z = true;
      if (bc[1] <= 1 && bc[3] <= 1 && i == mx - 2 && j == my - 2) z = true;
      if (z) st3(d, S_U0, p, mk3(0, 0, 0));
    }
  }
};

// ---- IB_BC -----------------------------------------------------------------------------------
// face metric = 0.5*centre + 0.5*centre (NEWMETRIC), see vfs_metrics_kernels.h
VFS_HD V3 face3(const VfsDev &d, int s0, long p, long q) {
  return mk3(0.5 * d.s[s0][p] + 0.5 * d.s[s0][q], 0.5 * d.s[s0 + 1][p] + 0.5 * d.s[s0 + 1][q], 0.5 * d.s[s0 + 2][p] + 0.5 * d.s[s0 + 2][q]);
}
VFS_HD double ib_face_flux(const VfsDev &d, long p, long q, int smet) {
  V3 a = ld3(d, S_U0, p), b = ld3(d, S_U0, q);
  double ucx = (a.x + b.x) * 0.5, ucy = (a.y + b.y) * 0.5, ucz = (a.z + b.z) * 0.5;
  V3 m = face3(d, smet, p, q);
  return ucx * m.x + ucy * m.y + ucz * m.z;
}
// momentum.c:2109-2166: faces touching an IB node ((int)nvert == 1) get the interpolated flux
struct IbBcFaces {
  VfsDev d;
  VFS_HD void operator()(int i, int j, int k) const {
    long p = d.idx(i, j, k);
    const double f = d.immersed == 3 ? 0.0 : 1.0;
    const double *nv = d.s[S_NV];
    const bool me = ((int)nv[p] == 1);
    if (me || (int)nv[p + 1] == 1) d.s[S_UC0][p] = ib_face_flux(d, p, p + 1, S_CSI0) * f;
    if (me || (int)nv[p + d.sj] == 1) d.s[S_UC1][p] = ib_face_flux(d, p, p + d.sj, S_ETA0) * f;
    if (me || (int)nv[p + d.sk] == 1) d.s[S_UC2][p] = ib_face_flux(d, p, p + d.sk, S_ZET0) * f;
    if (d.immersed == 3) {
      if (nv[p + 1] + nv[p] > 1.1) d.s[S_UC0][p] = 0;
      if (nv[p + d.sj] + nv[p] > 1.1) d.s[S_UC1][p] = 0;
      if (nv[p + d.sk] + nv[p] > 1.1) d.s[S_UC2][p] = 0;
    }
  }
};
// momentum.c:2193-2222: slip-wall normal flux zeroing and the component-wise periodic copies
// (which read the ghosts as they were BEFORE this call's final DALocalToLocal).
struct IbBcBoundary {
  VfsDev d;
  VFS_HD void operator()(int i, int j, int k) const {
    const int mx = d.mx, my = d.my, mz = d.mz, kg = k + d.kofs;
    long p = d.idx(i, j, k);
    if (d.bc[0] == 10 && i == 0) d.s[S_UC0][p] = 0;
    if (d.bc[1] == 10 && i == mx - 2) d.s[S_UC0][p] = 0;
    if (d.bc[2] == 10 && j == 0) d.s[S_UC1][p] = 0;
    if (d.bc[3] == 10 && j == my - 2) d.s[S_UC1][p] = 0;
    if (d.bc[4] == 10 && kg == 0) d.s[S_UC2][p] = 0;
    if (d.bc[5] == 10 && kg == mz - 2) d.s[S_UC2][p] = 0;
    // ii/jj/kk_periodic read the DA ghost (index -2 / m+1: the image as of the last refresh, i.e. BEFORE this call's changes);
    // the legacy i/j/k_periodic lines read the interior node itself (m-2 / 1: its value NOW, after IbBcFaces), :2206-2211
    if (d.perx && i == 0) d.s[S_UC0][p] = d.s[S_UC0][d.idx(d.legx ? mx - 2 : -2, j, k)];
    if (d.perx && i == mx - 1) d.s[S_UC0][p] = d.s[S_UC0][d.idx(d.legx ? 1 : mx + 1, j, k)];
    if (d.pery && j == 0) d.s[S_UC1][p] = d.s[S_UC1][d.idx(i, d.legy ? my - 2 : -2, k)];
    if (d.pery && j == my - 1) d.s[S_UC1][p] = d.s[S_UC1][d.idx(i, d.legy ? 1 : my + 1, k)];
    if (d.perz && kg == 0) d.s[S_UC2][p] = d.s[S_UC2][d.idx(i, j, d.legz ? k + (mz - 2) : k - 2)];
    if (d.perz && kg == mz - 1) d.s[S_UC2][p] = d.s[S_UC2][d.idx(i, j, d.legz ? k - (mz - 2) : k + 2)];
  }
};

#endif
