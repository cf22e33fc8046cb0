// vfs_actuator_kernels.h — actuator forcing on the background grid:
//   Calc_F_eul   Source/rotor_model.c:3668-3960   spread the Lagrangian forces of the turbine / nacelle elements onto
//                                                 the contravariant F_eul with a smoothed delta function
//   Calc_U_lagr  Source/rotor_model.c:2937-3150   interpolate the Cartesian velocity to the Lagrangian elements
// The reference scatters (elements outer, the cells of each element's index window inner).  Here Calc_F_eul is a
// GATHER: one thread per cell of the windows' bounding box walks the element list in the reference's order
// (object, then element) and adds the contributions of the windows that contain the cell — the same terms in the
// same order per cell, no atomics, deterministic.  Calc_U_lagr is one block per element reducing over its window.
#ifndef VFS_ACTUATOR_KERNELS_H
#define VFS_ACTUATOR_KERNELS_H
#include "vfs_common.h"

// the IBMNodes fields of all objects, concatenated on the device (13 doubles/ints per element)
struct ActElems {
  int n;                          // total number of elements (all objects, reference order)
  const double *cx, *cy, *cz, *dA, *fx, *fy, *fz;
  const int *i0, *i1, *j0, *j1, *k0, *k1;      // index windows, GLOBAL node indices, upper bounds exclusive
};

VFS_HD double dfunc_2h(double r) { return fabs(r) < 1.0 ? 1.0 - fabs(r) : 0.0; }                     // rotor_model.c:5130
VFS_HD double dfunc_s4h(double r) {                                                                    // rotor_model.c:5140
  const double a = fabs(r);
  if (a <= 0.5) return 3.0 / 8.0 + 3.14159265 / 32.0 - pow(r, 2) / 4.0;
  if (a >= 0.5 && a <= 1.5) return 1.0 / 4.0 + (1.0 - a) * sqrt(-2.0 + 8.0 * a - 4.0 * pow(r, 2)) / 8.0 - asin(sqrt(2.0) * (a - 1.0)) / 8.0;
  if (a >= 1.5 && a <= 2.5)
    return 17.0 / 16.0 - 3.14159265 / 64.0 - 3.0 * a / 4.0 + pow(r, 2) / 8.0 + (a - 2.0) * sqrt(-14.0 + 16.0 * a - 4.0 * pow(r, 2)) / 16.0 + asin(sqrt(2.0) * (a - 2.0)) / 16.0;
  return 0.0;
}
VFS_HD double dfunc_exp(double r, double n) { return exp(-(r / n) * (r / n)) / (pow(n, 1) * pow(3.1415926, 0.5)); }   // rotor_model.c:5230
VFS_HD double dfunc_s3h(double r) {                                                                    // rotor_model.c:5117
  const double a = fabs(r);
  if (a <= 1.0) return 17.0 / 48.0 + sqrt(3.0) * 3.14159265 / 108.0 + a / 4.0 - r * r / 4.0 + (1.0 - 2.0 * a) * sqrt(-12.0 * r * r + 12.0 * a + 1.0) / 16.0 - sqrt(3.0) * asin(sqrt(3.0) * (2.0 * a - 1.0) / 2.0) / 12.0;
  if (a >= 1.0 && a <= 2.0) return 55.0 / 48.0 - sqrt(3.0) * 3.14159265 / 108.0 - 13.0 * a / 12.0 + r * r / 4.0 + (2.0 * a - 3.0) * sqrt(-12.0 * r * r + 36.0 * a - 23.0) / 48.0 + sqrt(3.0) * asin(sqrt(3.0) * (2.0 * a - 3.0) / 2.0) / 36.0;
  return 0.0;
}

// per-cell geometry shared by both functions: unit normals of the coordinate lines (Calculate_normal, rhs2.c:614-646:
// the normalised columns of [csi;eta;zet]^-1, Calculate_Covariant_metrics rhs2.c:595-611) and the cell widths
struct ActCell { double ni[3], nj[3], nk[3], dhx, dhy, dhz; };
VFS_HD void act_cell(const VfsDev &d, long p, ActCell &C, int widthfixed, const double *dhf) {
  const double a11 = d.s[S_CSI0][p], a12 = d.s[S_CSI1][p], a13 = d.s[S_CSI2][p];
  const double a21 = d.s[S_ETA0][p], a22 = d.s[S_ETA1][p], a23 = d.s[S_ETA2][p];
  const double a31 = d.s[S_ZET0][p], a32 = d.s[S_ZET1][p], a33 = d.s[S_ZET2][p];
  const double det = a11 * (a33 * a22 - a32 * a23) - a21 * (a33 * a12 - a32 * a13) + a31 * (a23 * a12 - a22 * a13);
  const double G00 = (a33 * a22 - a32 * a23) / det, G01 = -(a33 * a12 - a32 * a13) / det, G02 = (a23 * a12 - a22 * a13) / det;
  const double G10 = -(a33 * a21 - a31 * a23) / det, G11 = (a33 * a11 - a31 * a13) / det, G12 = -(a23 * a11 - a21 * a13) / det;
  const double G20 = (a32 * a21 - a31 * a22) / det, G21 = -(a32 * a11 - a31 * a12) / det, G22 = (a22 * a11 - a21 * a12) / det;
  const double si = sqrt(G00 * G00 + G10 * G10 + G20 * G20), sj = sqrt(G01 * G01 + G11 * G11 + G21 * G21), sk = sqrt(G02 * G02 + G12 * G12 + G22 * G22);
  C.ni[0] = G00 / si; C.ni[1] = G10 / si; C.ni[2] = G20 / si;
  C.nj[0] = G01 / sj; C.nj[1] = G11 / sj; C.nj[2] = G21 / sj;
  C.nk[0] = G02 / sk; C.nk[1] = G12 / sk; C.nk[2] = G22 / sk;
  if (widthfixed) { C.dhx = dhf[0]; C.dhy = dhf[1]; C.dhz = dhf[2]; }
  else {
    const double aj = d.s[S_AJ][p];
    C.dhx = 1.0 / aj / sqrt(a11 * a11 + a12 * a12 + a13 * a13);
    C.dhy = 1.0 / aj / sqrt(a21 * a21 + a22 * a22 + a23 * a23);
    C.dhz = 1.0 / aj / sqrt(a31 * a31 + a32 * a32 + a33 * a33);
  }
}

// rotor_model.c:3731-3829: F_eul of one cell += sum over the elements whose window holds the cell
struct FEulGather {
  VfsDev d; ActElems E; int df, widthfixed; double halfwidth, dhf[3];
  VFS_HD double d1(double r) const { return df == 0 ? dfunc_2h(r) : (df == 7 ? dfunc_exp(r, halfwidth) : dfunc_s4h(r)); }     // :3785-3799
  VFS_HD void operator()(int i, int j, int k) const {
    const int kg = k + d.kofs;
    const long p = d.idx(i, j, k);
    const double *X = d.s[S_X], *Y = d.s[S_Y], *Z = d.s[S_Z];
    bool any = false;
    for (int l = 0; l < E.n && !any; l++) any = i >= E.i0[l] && i < E.i1[l] && j >= E.j0[l] && j < E.j1[l] && kg >= E.k0[l] && kg < E.k1[l];
    if (!any) return;
    const long pj = p - d.sj, pk = p - d.sk, pi = p - 1;
    // centres of the cell's +i, +j, +k faces (node order as in the reference)
    const double xi = (X[p] + X[pk] + X[pj] + X[pk - d.sj]) * 0.25, yi = (Y[p] + Y[pk] + Y[pj] + Y[pk - d.sj]) * 0.25, zi = (Z[p] + Z[pk] + Z[pj] + Z[pk - d.sj]) * 0.25;
    const double xj = (X[p] + X[pk] + X[pi] + X[pk - 1]) * 0.25, yj = (Y[p] + Y[pk] + Y[pi] + Y[pk - 1]) * 0.25, zj = (Z[p] + Z[pk] + Z[pi] + Z[pk - 1]) * 0.25;
    const double xk = (X[p] + X[pj] + X[pi] + X[pj - 1]) * 0.25, yk = (Y[p] + Y[pj] + Y[pi] + Y[pj - 1]) * 0.25, zk = (Z[p] + Z[pj] + Z[pi] + Z[pj - 1]) * 0.25;
    ActCell C; act_cell(d, p, C, widthfixed, dhf);
    const double vol_eul = 1.0 / (C.dhx * C.dhy * C.dhz);
    const double c0 = d.s[S_CSI0][p], c1 = d.s[S_CSI1][p], c2 = d.s[S_CSI2][p];
    const double e0 = d.s[S_ETA0][p], e1 = d.s[S_ETA1][p], e2 = d.s[S_ETA2][p];
    const double z0 = d.s[S_ZET0][p], z1 = d.s[S_ZET1][p], z2 = d.s[S_ZET2][p];
    double f0 = d.s[S_FE0][p], f1 = d.s[S_FE1][p], f2 = d.s[S_FE2][p];
    for (int l = 0; l < E.n; l++) {
      if (!(i >= E.i0[l] && i < E.i1[l] && j >= E.j0[l] && j < E.j1[l] && kg >= E.k0[l] && kg < E.k1[l])) continue;
      const double cx = E.cx[l], cy = E.cy[l], cz = E.cz[l];
      double w[3];
      const double fc[3][3] = {{xi - cx, yi - cy, zi - cz}, {xj - cx, yj - cy, zj - cz}, {xk - cx, yk - cy, zk - cz}};
#pragma unroll
      for (int q = 0; q < 3; q++) {
        const double r1 = fabs(fc[q][0] * C.ni[0] + fc[q][1] * C.ni[1] + fc[q][2] * C.ni[2]) / C.dhx;
        const double r2 = fabs(fc[q][0] * C.nj[0] + fc[q][1] * C.nj[1] + fc[q][2] * C.nj[2]) / C.dhy;
        const double r3 = fabs(fc[q][0] * C.nk[0] + fc[q][1] * C.nk[1] + fc[q][2] * C.nk[2]) / C.dhz;
        w[q] = vol_eul * d1(r1) * d1(r2) * d1(r3);
      }
      const double Fx = E.fx[l], Fy = E.fy[l], Fz = E.fz[l], dA = E.dA[l];
      f0 += Fx * w[0] * dA * c0 + Fy * w[0] * dA * c1 + Fz * w[0] * dA * c2;
      f1 += Fx * w[1] * dA * e0 + Fy * w[1] * dA * e1 + Fz * w[1] * dA * e2;
      f2 += Fx * w[2] * dA * z0 + Fy * w[2] * dA * z1 + Fz * w[2] * dA * z2;
    }
    d.s[S_FE0][p] = f0; d.s[S_FE1][p] = f1; d.s[S_FE2][p] = f2;
  }
};

// rotor_model.c:3832-3903: no forcing in cells whose 27-neighbourhood holds a solid node, nor in the two node layers
// next to every domain boundary.  Two passes in the reference's effect: every interior cell decides for itself, the
// boundary nodes are cleared by their interior neighbour (same j,k / i,k / i,j line).
struct FEulMask {
  VfsDev d;
  VFS_HD void operator()(int i, int j, int k) const {
    const int kg = k + d.kofs;
    const long p = d.idx(i, j, k);
    const bool iin = i >= 1 && i <= d.mx - 2, jin = j >= 1 && j <= d.my - 2, kin = kg >= 1 && kg <= d.mz - 2;
    bool zero = false;
    if (iin && jin && kin) {
      double s = 0;
      for (int c = -1; c <= 1; c++) for (int b = -1; b <= 1; b++) for (int a = -1; a <= 1; a++) s += d.s[S_NV][p + c * d.sk + b * d.sj + a];
      zero = s > 2.9 || i == 1 || i == d.mx - 2 || j == 1 || j == d.my - 2 || kg == 1 || kg == d.mz - 2;
    } else {
      // a boundary node is cleared when the interior cell next to it along its boundary direction exists
      const int nb = (iin ? 0 : 1) + (jin ? 0 : 1) + (kin ? 0 : 1);
      zero = nb == 1;
    }
    if (zero) { d.s[S_FE0][p] = 0; d.s[S_FE1][p] = 0; d.s[S_FE2][p] = 0; }
  }
};

// rotor_model.c:2984-3031: U_lagr of element l = sum over its window of ucat * dfunc_s3h^3; the rank's share (owned
// interior cells), one block per element, fixed-order tree reduction
struct ULagrArgs { VfsDev d; ActElems E; double *out; };      // out[3 * l + c]
VFS_HD void ulagr_cell(const VfsDev &d, const ActElems &E, int l, int i, int j, int k, double acc[3]) {
  const long p = d.idx(i, j, k);
  const double *X = d.s[S_X], *Y = d.s[S_Y], *Z = d.s[S_Z];
  const long q[8] = {p, p - d.sj, p - d.sk, p - d.sk - d.sj, p - 1, p - d.sj - 1, p - d.sk - 1, p - d.sk - d.sj - 1};
  double xc = 0, yc = 0, zc = 0;
  for (int n = 0; n < 8; n++) { xc += X[q[n]]; yc += Y[q[n]]; zc += Z[q[n]]; }
  xc *= 0.125; yc *= 0.125; zc *= 0.125;
  ActCell C; act_cell(d, p, C, 0, nullptr);
  const double rx = xc - E.cx[l], ry = yc - E.cy[l], rz = zc - E.cz[l];
  const double r1 = (rx * C.ni[0] + ry * C.ni[1] + rz * C.ni[2]) / C.dhx, r2 = (rx * C.nj[0] + ry * C.nj[1] + rz * C.nj[2]) / C.dhy, r3 = (rx * C.nk[0] + ry * C.nk[1] + rz * C.nk[2]) / C.dhz;
  const double w = dfunc_s3h(r1) * dfunc_s3h(r2) * dfunc_s3h(r3);
  acc[0] += d.s[S_U0][p] * w; acc[1] += d.s[S_U1][p] * w; acc[2] += d.s[S_U2][p] * w;
}
#ifndef VFS_EMU
__global__ void __launch_bounds__(256) k_ulagr(ULagrArgs A) {
  const int l = blockIdx.x;
  const VfsDev &d = A.d;
  // window clipped to this rank's owned cells
  const int i0 = max(A.E.i0[l], 0), i1 = min(A.E.i1[l], d.mx), j0 = max(A.E.j0[l], 0), j1 = min(A.E.j1[l], d.my);
  const int k0 = max(A.E.k0[l] - d.kofs, 0), k1 = min(A.E.k1[l] - d.kofs, d.nzl);
  const int ni = max(i1 - i0, 0), nj = max(j1 - j0, 0), nk = max(k1 - k0, 0);
  const long n = (long)ni * nj * nk;
  double acc[3] = {0, 0, 0};
  for (long t = threadIdx.x; t < n; t += 256) {
    const int i = i0 + (int)(t % ni), j = j0 + (int)((t / ni) % nj), k = k0 + (int)(t / ((long)ni * nj));
    ulagr_cell(d, A.E, l, i, j, k, acc);
  }
  __shared__ double sm[3][256];
  for (int c = 0; c < 3; c++) sm[c][threadIdx.x] = acc[c];
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) { if (threadIdx.x < s) for (int c = 0; c < 3; c++) sm[c][threadIdx.x] += sm[c][threadIdx.x + s]; __syncthreads(); }
  if (threadIdx.x < 3) A.out[3 * l + threadIdx.x] = sm[threadIdx.x][0];
}
#endif
#endif
