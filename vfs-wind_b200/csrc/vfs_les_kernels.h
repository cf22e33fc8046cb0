// vfs_les_kernels.h — dynamic Smagorinsky constant and eddy viscosity
// (Source/les.c:75-1141 Compute_Smagorinsky_Constant_1, :1143-1361 Compute_eddy_viscosity_LES;
// helpers Source/k-omega.c:313-618 Compute_du_center/Compute_du_dxyz, Source/rhs2.c:499-523
// integrate_testfilter_simpson, :419-441 integrate_testfilter_ik, :595-611 covariant metrics).
//
// The 27-point Simpson filters are accumulated on the fly in the reference's summation order
// (k-offset outer, j, i inner; term (s*w)*v, rhs2.c:510-520) so no 3x3x3 stack arrays are needed.
// Clark terms (les.c:420-428) are never used when clark=0 and are not computed.
#ifndef VFS_LES_KERNELS_H
#define VFS_LES_KERNELS_H
#include "vfs_common.h"
#include "vfs_rhs_kernels.h"

// pow(x, 1./3.) (les.c:462-468,1208): cbrt on the device (<= 1 ulp apart; the CUDA
// generic pow costs ~10x more), literal pow in the host emulation
#if defined(__CUDA_ARCH__)
#define VFS_CBRT(x) cbrt(x)
#else
#define VFS_CBRT(x) pow((x), 1. / 3.)
#endif

// Accessors: u(a,di,dj,dk), nv(di,dj,dk), aj(di,dj,dk) at node offsets relative to the cell.
// GlobalAccS reads velocity-like field S0..S0+2 from the global padded arrays; the tiled kernels
// provide the same interface over TMA-staged shared-memory planes.
template <int S0> struct GlobalAccS {
  const VfsDev &d; long p;
  VFS_HD double u(int a, int di, int dj, int dk) const { return d.s[S0 + a][p + di + dj * d.sj + dk * d.sk]; }
  VFS_HD double nv(int di, int dj, int dk) const { return d.s[S_NV][p + di + dj * d.sj + dk * d.sk]; }
  VFS_HD double iaj(int di, int dj, int dk) const { return d.s[S_IAJ][p + di + dj * d.sj + dk * d.sk]; }
  // the cell's own centre metrics (s = 0..8: csi, eta, zet), aj, and LES grid factors (q = 0..2: S_LFINV, S_LTF2, S_LF2)
  VFS_HD double met(int s) const { return d.s[S_CSI0 + s][p]; }
  VFS_HD double aj() const { return d.s[S_AJ][p]; }
  VFS_HD double geo(int q) const { return d.s[S_LFINV + q][p]; }
  VFS_HD unsigned char nearv(const VfsDev &dd, long pp) const { return dd.near[pp]; }
};

// centre difference of component a along direction T (k-omega.c:318-430); lowc = 1 for i/j, 0 for k
// 
template <int T, class Acc> VFS_HD double dcen(const Acc &A, int a, int c, int m, int per, int lowc) {
  constexpr int ti = (T == 0), tj = (T == 1), tk = (T == 2);
  // one-sided (u0 - um), (up - u0) or centred (up - um)/2 (k-omega.c:318-430) as (P - M) * c:
  // operands fetched unconditionally (independent loads, one memory round trip), chosen by selects;
  // bitwise the reference's value (one subtraction, exact scaling)
  const double up = A.u(a, ti, tj, tk), u0 = A.u(a, 0, 0, 0), um = A.u(a, -ti, -tj, -tk);
  const bool hi = A.nv(ti, tj, tk) > VFS_SOLID || (!per && c == m - 2);
  const bool lo = !hi && (A.nv(-ti, -tj, -tk) > VFS_SOLID || (!per && c == lowc));
  return ((hi ? u0 : up) - (lo ? u0 : um)) * ((hi || lo) ? 1.0 : 0.5);
}

// velocity gradient at a cell centre: g[a][b] = d u_a / d x_b  (k-omega.c:605-618)
// PLAIN = true: the caller knows that no stencil switch fires (no nvert != 0 nearby, not next to a
// domain end), so every difference is the centred one — the value the selects would produce.
template <bool PLAIN = false, class Acc> VFS_HD void grad_center_a(const VfsDev &d, const Acc &A, int i, int j, int kg, long, double g[3][3]) {
  const double ajc = A.aj();
  const double c0 = A.met(0), c1 = A.met(1), c2 = A.met(2);
  const double e0 = A.met(3), e1 = A.met(4), e2 = A.met(5);
  const double z0 = A.met(6), z1 = A.met(7), z2 = A.met(8);
#pragma unroll
  for (int a = 0; a < 3; a++) {
    double dc, de, dz;
    if (PLAIN) {
      dc = (A.u(a, 1, 0, 0) - A.u(a, -1, 0, 0)) * 0.5; de = (A.u(a, 0, 1, 0) - A.u(a, 0, -1, 0)) * 0.5; dz = (A.u(a, 0, 0, 1) - A.u(a, 0, 0, -1)) * 0.5;
    } else {
      dc = dcen<0>(A, a, i, d.mx, d.perx, 1); de = dcen<1>(A, a, j, d.my, d.pery, 1); dz = dcen<2>(A, a, kg, d.mz, d.perz, 0);
    }
    g[a][0] = (dc * c0 + de * e0 + dz * z0) * ajc;
    g[a][1] = (dc * c1 + de * e1 + dz * z1) * ajc;
    g[a][2] = (dc * c2 + de * e2 + dz * z2) * ajc;
  }
}
// warp-uniform choice between the two (VfsDev::near, VFS_WARP_ANY)
template <class Acc> VFS_HD void grad_center_auto(const VfsDev &d, const Acc &A, int i, int j, int kg, long p, double g[3][3]) {
  const bool special = A.nearv(d, p) != 0 || i <= 1 || i >= d.mx - 2 || j <= 1 || j >= d.my - 2 || kg <= 1 || kg >= d.mz - 2;
  if (VFS_WARP_ANY(special)) grad_center_a<false>(d, A, i, j, kg, p, g);
  else grad_center_a<true>(d, A, i, j, kg, p, g);
}
VFS_HD void grad_center(const VfsDev &d, int su, int i, int j, int kg, long p, double g[3][3]) {
  if (su == S_U0) { GlobalAccS<S_U0> A = {d, p}; grad_center_a(d, A, i, j, kg, p, g); }
  else { GlobalAccS<S_UF0> A = {d, p}; grad_center_a(d, A, i, j, kg, p, g); }
}

VFS_HD double sabs_of(const double g[3][3]) {
  const double Sxx = 0.5 * (g[0][0] + g[0][0]), Sxy = 0.5 * (g[0][1] + g[1][0]), Sxz = 0.5 * (g[0][2] + g[2][0]);
  const double Syx = Sxy, Syy = 0.5 * (g[1][1] + g[1][1]), Syz = 0.5 * (g[1][2] + g[2][1]);
  const double Szx = Sxz, Szy = Syz, Szz = 0.5 * (g[2][2] + g[2][2]);
  return sqrt(2.0 * (Sxx * Sxx + Sxy * Sxy + Sxz * Sxz + Syx * Syx + Syy * Syy + Syz * Syz + Szx * Szx + Szy * Szy + Szz * Szz));
}

VFS_HD double simpson_w(int r, int q, int pp) {
  double s = 1.0;
  if (r == 0) s *= 4.;
  if (q == 0) s *= 4.;
  if (pp == 0) s *= 4.;
  return s;
}

// 1/aj = cell volume: the weight of every test filter (get_weight, les.c:31-40; les.c:737).  It only
// changes with the grid, so it is computed once per metrics upload instead of 27 times per cell.
struct InvAj {
  VfsDev d;
  VFS_HD void operator()(int i, int j, int k) const { long p = d.idx(i, j, k); d.s[S_IAJ][p] = 1. / d.s[S_AJ][p]; }
};

// per-node quantities that pass 2 filters over the 27-point neighbourhood (les.c:354-439):
// w = 1/aj (0 where nvert > 0.1), contravariant U = [csi;eta;zet] u, and |S| S_ij.  Written once
// per node instead of being recomputed 27 times by every neighbour.
VFS_HD void les_derive_store(const VfsDev &d, long n, const double g[3][3], double S) {
  d.s[S_LW][n] = (d.s[S_NV][n] > 0.1) ? 0. : d.s[S_IAJ][n];
  const double u0 = d.s[S_U0][n], u1 = d.s[S_U1][n], u2 = d.s[S_U2][n];
  d.s[S_LU0][n] = u0 * d.s[S_CSI0][n] + u1 * d.s[S_CSI1][n] + u2 * d.s[S_CSI2][n];
  d.s[S_LU1][n] = u0 * d.s[S_ETA0][n] + u1 * d.s[S_ETA1][n] + u2 * d.s[S_ETA2][n];
  d.s[S_LU2][n] = u0 * d.s[S_ZET0][n] + u1 * d.s[S_ZET1][n] + u2 * d.s[S_ZET2][n];
  d.s[S_LSS0][n] = (0.5 * (g[0][0] + g[0][0])) * S; d.s[S_LSS1][n] = (0.5 * (g[0][1] + g[1][0])) * S; d.s[S_LSS2][n] = (0.5 * (g[0][2] + g[2][0])) * S;
  d.s[S_LSS3][n] = (0.5 * (g[1][1] + g[1][1])) * S; d.s[S_LSS4][n] = (0.5 * (g[1][2] + g[2][1])) * S; d.s[S_LSS5][n] = (0.5 * (g[2][2] + g[2][2])) * S;
}
// same with the node's centre metrics already in registers
VFS_HD void les_derive_store_m(const VfsDev &d, long n, const double g[3][3], double S, const double *m, double nv, double iaj, double u0, double u1, double u2) {
  d.s[S_LW][n] = (nv > 0.1) ? 0. : iaj;
  d.s[S_LU0][n] = u0 * m[0] + u1 * m[1] + u2 * m[2];
  d.s[S_LU1][n] = u0 * m[3] + u1 * m[4] + u2 * m[5];
  d.s[S_LU2][n] = u0 * m[6] + u1 * m[7] + u2 * m[8];
  d.s[S_LSS0][n] = (0.5 * (g[0][0] + g[0][0])) * S; d.s[S_LSS1][n] = (0.5 * (g[0][1] + g[1][0])) * S; d.s[S_LSS2][n] = (0.5 * (g[0][2] + g[2][0])) * S;
  d.s[S_LSS3][n] = (0.5 * (g[1][1] + g[1][1])) * S; d.s[S_LSS4][n] = (0.5 * (g[1][2] + g[2][1])) * S; d.s[S_LSS5][n] = (0.5 * (g[2][2] + g[2][2])) * S;
}
// domain-boundary nodes: grad u and |S| are zero there in the reference (VecSet, les.c:183-186)
// dynamic_only: only the part that changes with the velocity (U = [csi;eta;zet] u) is rewritten; the weight w (a function
// of nvert / aj) and the zero |S|S_ij of the boundary nodes were stored when the grid / mask last changed, and nothing
// else writes them there (the periodic node copies that follow overwrite the periodic planes' |S|S_ij every step anyway)
struct LesDeriveBoundary {
  VfsDev d; int dynamic_only;
  VFS_HD void operator()(int i, int j, int k) const {
    const long n = d.idx(i, j, k);
    if (!dynamic_only) { const double z[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}; les_derive_store(d, n, z, 0.); return; }
    const double u0 = d.s[S_U0][n], u1 = d.s[S_U1][n], u2 = d.s[S_U2][n];
    d.s[S_LU0][n] = u0 * d.s[S_CSI0][n] + u1 * d.s[S_CSI1][n] + u2 * d.s[S_CSI2][n];
    d.s[S_LU1][n] = u0 * d.s[S_ETA0][n] + u1 * d.s[S_ETA1][n] + u2 * d.s[S_ETA2][n];
    d.s[S_LU2][n] = u0 * d.s[S_ZET0][n] + u1 * d.s[S_ZET1][n] + u2 * d.s[S_ZET2][n];
  }
};

// accessor adaptor: the cell's own metrics fetched up front (their global-load latency then overlaps the
// shared-memory test filter that does not need them), everything else forwarded
template <class Acc> struct PreMet {
  const Acc &A; double m[9], ajv;
  VFS_HD PreMet(const Acc &a) : A(a) {
#pragma unroll
    for (int s = 0; s < 9; s++) m[s] = a.met(s);
    ajv = a.aj();
  }
  VFS_HD double met(int s) const { return m[s]; }
  VFS_HD double aj() const { return ajv; }
  VFS_HD double u(int a, int di, int dj, int dk) const { return A.u(a, di, dj, dk); }
  VFS_HD double nv(int di, int dj, int dk) const { return A.nv(di, dj, dk); }
  VFS_HD unsigned char nearv(const VfsDev &dd, long pp) const { return A.nearv(dd, pp); }
};
// (1,4,1)^2-weighted sums of w and w u_a over the nine nodes of the cell's row/column neighbourhood in the plane
// at k offset dk: the i-j part of the Simpson test filter (rhs2.c:499-523, weights get_weight les.c:31-40)
template <class Acc0, class Acc> VFS_HD void les1_plane_sum(const Acc0 &A0, const Acc &A, int dk, double *o) {
  o[0] = o[1] = o[2] = o[3] = 0;
#pragma unroll
  for (int q = -1; q <= 1; q++)
#pragma unroll
    for (int pp = -1; pp <= 1; pp++) {
      const double w = (A0.nv(pp, q, dk) > 0.1) ? 0. : A0.iaj(pp, q, dk);
      const double sw = ((q == 0 ? 4. : 1.) * (pp == 0 ? 4. : 1.)) * w;
      o[0] += sw;
#pragma unroll
      for (int a = 0; a < 3; a++) o[1 + a] += sw * A.u(a, pp, q, dk);
    }
}
// k window of those plane sums carried by a marching thread: planes k-1 and k
struct Les1Win { double f[2][4]; };
// les.c:199-246: grad u, |S| and the test-filtered velocity (+ the per-node derived quantities).
// win != null (marching kernels): the test filter is evaluated plane by plane — the nine-point sums of planes
// k-1 and k come from the previous steps, only plane k+1 is gathered (45 shared-memory reads per step instead
// of 135; the k-outer summation order of the reference becomes k-last, a rounding-level difference).
template <class Acc> VFS_HD void les1_core(const VfsDev &d, const Acc &A0, int i, int j, int kg, long p, Les1Win *win = nullptr, bool first = false) {
  double wn[4];
  if (win && !d.testfilter_ik) {
    if (first) { les1_plane_sum(A0, A0, -1, win->f[0]); les1_plane_sum(A0, A0, 0, win->f[1]); }
    les1_plane_sum(A0, A0, 1, wn);
  }
  const bool skip = A0.nv(0, 0, 0) > 1.1;      // skipped by the reference: its zero-initialised work vectors keep 0 here
  double uf[3] = {0, 0, 0};
  if (win && !d.testfilter_ik) {
    const double ws = win->f[0][0] + 4. * win->f[1][0] + wn[0];
#pragma unroll
    for (int a = 0; a < 3; a++) uf[a] = (win->f[0][1 + a] + 4. * win->f[1][1 + a] + wn[1 + a]) / ws;
#pragma unroll
    for (int a = 0; a < 4; a++) { win->f[0][a] = win->f[1][a]; win->f[1][a] = wn[a]; }
  }
  if (skip) {
    const double z[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    d.s[S_SABS][p] = 0;
    for (int a = 0; a < 3; a++) d.s[S_UF0 + a][p] = 0;
    les_derive_store(d, p, z, 0.);
    return;
  }
  const PreMet<Acc> A(A0);
  if (d.testfilter_ik) {
#pragma unroll
    for (int a = 0; a < 3; a++)
      uf[a] = ((A.u(a, -1, 0, -1) + A.u(a, -1, 0, 1) + A.u(a, 1, 0, -1) + A.u(a, 1, 0, 1)) + 4. * (A.u(a, 0, 0, -1) + A.u(a, -1, 0, 0) + A.u(a, 0, 0, 1) + A.u(a, 1, 0, 0)) + 16. * A.u(a, 0, 0, 0)) / 36.;
  } else if (!win) {
    double ws = 0, vs[3] = {0, 0, 0};
#pragma unroll
    for (int r = -1; r <= 1; r++)
#pragma unroll
      for (int q = -1; q <= 1; q++)
#pragma unroll
        for (int pp = -1; pp <= 1; pp++) {
          const double w = (A0.nv(pp, q, r) > 0.1) ? 0. : A0.iaj(pp, q, r);
          const double sw = simpson_w(r, q, pp) * w;
          ws += sw;
#pragma unroll
          for (int a = 0; a < 3; a++) vs[a] += sw * A.u(a, pp, q, r);
        }
    for (int a = 0; a < 3; a++) uf[a] = vs[a] / ws;
  }
  for (int a = 0; a < 3; a++) d.s[S_UF0 + a][p] = uf[a];
  double g[3][3];
  grad_center_auto(d, A, i, j, kg, p, g);
  const double S = sabs_of(g);
  d.s[S_SABS][p] = S;
  les_derive_store_m(d, p, g, S, A.m, A0.nv(0, 0, 0), A0.iaj(0, 0, 0), A.u(0, 0, 0, 0), A.u(1, 0, 0, 0), A.u(2, 0, 0, 0));
}
struct LesPass1 {
  VfsDev d;
  VFS_HD void operator()(int i, int j, int k) const {
    const long p = d.idx(i, j, k);
    GlobalAccS<S_U0> A = {d, p};
    les1_core(d, A, i, j, d.kglob(k), p);
  }
};

// clark: the velocity gradient of every cell kept for pass 2 (lSx, lSy, lSz of les.c:181-232; zero at solid cells and,
// unless a periodic node copy overwrites it, at the boundary nodes).  A separate small kernel so that the marching
// pass-1 kernels keep their register budget; the mixed model runs the one-thread-per-cell pass 1 and 2 anyway.
struct LesGradStore {
  VfsDev d; int boundary;
  VFS_HD void operator()(int i, int j, int k) const {
    const long p = d.idx(i, j, k);
    double g[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    if (!boundary && !(d.s[S_NV][p] > 1.1)) grad_center(d, S_U0, i, j, d.kglob(k), p, g);
    for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) d.s[S_GR0 + 3 * a + b][p] = g[a][b];
  }
};

// les.c:354-439 per-node part of pass 2: weight w and the products entering the test filters,
// v[0] = w, v[1..9] = U_a u_b (a-major), v[10..15] = |S| S_ij (xx,xy,xz,yy,yz,zz)
#define VFS_LES2_NV 16
VFS_HD void les2_products(const VfsDev &d, long n, double *v) {
  v[0] = d.s[S_LW][n];
  const double u[3] = {d.s[S_U0][n], d.s[S_U1][n], d.s[S_U2][n]};
  const double U[3] = {d.s[S_LU0][n], d.s[S_LU1][n], d.s[S_LU2][n]};
  for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) v[1 + 3 * a + b] = U[a] * u[b];
  for (int a = 0; a < 6; a++) v[10 + a] = d.s[S_LSS0 + a][n];
}

// les.c:441-669 after the filters: fs[0] = sum of Simpson weights (or 36 for testfilter_ik),
// fs[1..15] = filtered sums of v[1..15]; sum_weight = sum of w*coef (les.c:441-468)
// clark (mixed model, les.c:497-556, 656): fT = filtered sums of the six Clark-tensor components T_ab (xx,xy,xz,yy,yz,zz)
// of the stencil nodes, h2 = the cell's squared grid lengths; null without the model
VFS_HD void les2_finish(const VfsDev &d, int i, int j, int kg, long p, const double *fs, double sum_weight, const double *fT = nullptr, const double *h2 = nullptr) {
  const double ajc = d.s[S_AJ][p];
  const double csi[3] = {d.s[S_CSI0][p], d.s[S_CSI1][p], d.s[S_CSI2][p]};
  const double eta[3] = {d.s[S_ETA0][p], d.s[S_ETA1][p], d.s[S_ETA2][p]};
  const double zet[3] = {d.s[S_ZET0][p], d.s[S_ZET1][p], d.s[S_ZET2][p]};
  const double fdiv = fs[0];
  const double filter = VFS_CBRT(1. / ajc);
  const double test_filter = d.testfilter_ik ? 1.709975946676697 /* pow(5, 1./3.) */ * filter : VFS_CBRT(sum_weight);
  const double _u[3] = {d.s[S_UF0][p], d.s[S_UF1][p], d.s[S_UF2][p]};
  const double _U[3] = {_u[0] * csi[0] + _u[1] * csi[1] + _u[2] * csi[2], _u[0] * eta[0] + _u[1] * eta[1] + _u[2] * eta[2], _u[0] * zet[0] + _u[1] * zet[1] + _u[2] * zet[2]};
  double gh[3][3];
  grad_center(d, S_UF0, i, j, kg, p, gh);
  const double Sh[3][3] = {{0.5 * (gh[0][0] + gh[0][0]), 0.5 * (gh[0][1] + gh[1][0]), 0.5 * (gh[0][2] + gh[2][0])},
                           {0.5 * (gh[0][1] + gh[1][0]), 0.5 * (gh[1][1] + gh[1][1]), 0.5 * (gh[1][2] + gh[2][1])},
                           {0.5 * (gh[0][2] + gh[2][0]), 0.5 * (gh[1][2] + gh[2][1]), 0.5 * (gh[2][2] + gh[2][2])}};
  const double S_hat = sabs_of(gh);
  double Lij[3][3], SSh[3][3];
  // valsum / wsum (rhs2.c:522) evaluated as valsum * (1/wsum): 15 FP64 divisions -> 1 on the device
  // (<= 1 ulp apart per filtered value); the host emulation keeps the literal division
#if defined(__CUDA_ARCH__)
  const double finv = 1. / fdiv;
#define VFS_FDIV(x) ((x) * finv)
#else
#define VFS_FDIV(x) ((x) / fdiv)
#endif
  for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) Lij[a][b] = VFS_FDIV(fs[1 + 3 * a + b]) - _U[a] * _u[b];
  SSh[0][0] = VFS_FDIV(fs[10]); SSh[0][1] = SSh[1][0] = VFS_FDIV(fs[11]); SSh[0][2] = SSh[2][0] = VFS_FDIV(fs[12]);
  SSh[1][1] = VFS_FDIV(fs[13]); SSh[1][2] = SSh[2][1] = VFS_FDIV(fs[14]); SSh[2][2] = VFS_FDIV(fs[15]);
  double Nij[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  if (fT) {
    // N^c = 4 R - T^ with R the Clark tensor of the test-filtered velocity gradient, rotated like M (les.c:517-556)
    const int ix[3][3] = {{0, 1, 2}, {1, 3, 4}, {2, 4, 5}};
    double Nc[3][3];
    for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) {
      const double R = (gh[a][0] * gh[b][0] * h2[0] + gh[a][1] * gh[b][1] * h2[1] + gh[a][2] * gh[b][2] * h2[2]) / 12.;
      Nc[a][b] = 4.0 * R - VFS_FDIV(fT[ix[a][b]]);
    }
    for (int a = 0; a < 3; a++) {
      Nij[a][0] = Nc[a][0] * csi[0] + Nc[a][1] * csi[1] + Nc[a][2] * csi[2];
      Nij[a][1] = Nc[a][0] * eta[0] + Nc[a][1] * eta[1] + Nc[a][2] * eta[2];
      Nij[a][2] = Nc[a][0] * zet[0] + Nc[a][1] * zet[1] + Nc[a][2] * zet[2];
    }
  }
#undef VFS_FDIV
  // covariant metric tensor G (les.c:607-622)
  const double a11 = csi[0], a12 = csi[1], a13 = csi[2], a21 = eta[0], a22 = eta[1], a23 = eta[2], a31 = zet[0], a32 = zet[1], a33 = zet[2];
  const double det = a11 * (a33 * a22 - a32 * a23) - a21 * (a33 * a12 - a32 * a13) + a31 * (a23 * a12 - a22 * a13);
#if defined(__CUDA_ARCH__)
  const double dinv = 1. / det;       // 9 divisions -> 1 (see VFS_FDIV above)
#define VFS_DDIV(x) ((x) * dinv)
#else
#define VFS_DDIV(x) ((x) / det)
#endif
  const double xcsi = VFS_DDIV(a33 * a22 - a32 * a23), xeta = -VFS_DDIV(a33 * a12 - a32 * a13), xzet = VFS_DDIV(a23 * a12 - a22 * a13);
  const double ycsi = -VFS_DDIV(a33 * a21 - a31 * a23), yeta = VFS_DDIV(a33 * a11 - a31 * a13), yzet = -VFS_DDIV(a23 * a11 - a21 * a13);
  const double zcsi = VFS_DDIV(a32 * a21 - a31 * a22), zeta = -VFS_DDIV(a32 * a11 - a31 * a12), zzet = VFS_DDIV(a22 * a11 - a21 * a12);
#undef VFS_DDIV
  double G[3][3];
  G[0][0] = xcsi * xcsi + ycsi * ycsi + zcsi * zcsi;
  G[1][1] = xeta * xeta + yeta * yeta + zeta * zeta;
  G[2][2] = xzet * xzet + yzet * yzet + zzet * zzet;
  G[0][1] = G[1][0] = xeta * xcsi + yeta * ycsi + zeta * zcsi;
  G[0][2] = G[2][0] = xzet * xcsi + yzet * ycsi + zzet * zcsi;
  G[1][2] = G[2][1] = xeta * xzet + yeta * yzet + zeta * zzet;
  double Mc[3][3], M[3][3];
  const double tf2 = test_filter * test_filter, f2 = filter * filter;   // == pow(x, 2.) (correctly rounded)
  for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) Mc[a][b] = -tf2 * S_hat * Sh[a][b] + f2 * SSh[a][b];
  for (int a = 0; a < 3; a++) {
    M[a][0] = Mc[a][0] * csi[0] + Mc[a][1] * csi[1] + Mc[a][2] * csi[2];
    M[a][1] = Mc[a][0] * eta[0] + Mc[a][1] * eta[1] + Mc[a][2] * eta[2];
    M[a][2] = Mc[a][0] * zet[0] + Mc[a][1] * zet[1] + Mc[a][2] * zet[2];
  }
  double num = 0, den = 0;
  for (int q = 0; q < 3; q++) for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) {
    num += Lij[b][a] * M[a][q] * G[b][q];
    if (fT) num -= Nij[b][a] * M[a][q] * G[b][q];
  }
  for (int m = 0; m < 3; m++) for (int n = 0; n < 3; n++) for (int l = 0; l < 3; l++) den += M[n][m] * M[n][l] * G[m][l];
  d.s[S_LM][p] = num; d.s[S_MM][p] = den;
}

// ---- marching-kernel form of the same algebra -------------------------------------------------------
// Everything in les.c:441-468,607-626 that depends on the grid and the nvert mask only — the Simpson
// weight sum (the divisor of every test filter, rhs2.c:522), filter = (1/aj)^(1/3), test_filter =
// (sum coef*w)^(1/3) and the covariant metric tensor G — is computed once per grid/mask upload
// instead of once per cell per step: two cbrt, three divisions, ~150 multiply-adds and two of the
// seventeen filters leave the per-step kernel.  Sums are accumulated in the reference's order.
struct LesGeo {
  VfsDev d;
  VFS_HD void operator()(int i, int j, int k) const {
    const long p = d.idx(i, j, k);
    double fdiv = 0, sum_weight = 0;
    for (int r = -1; r <= 1; r++) for (int q = -1; q <= 1; q++) for (int pp = -1; pp <= 1; pp++) {
      const long n = p + r * d.sk + q * d.sj + pp;
      const double w = (d.s[S_NV][n] > 0.1) ? 0. : d.s[S_IAJ][n];
      sum_weight += w * (0.125 * (r == 0 ? 2. : 1.) * (q == 0 ? 2. : 1.) * (pp == 0 ? 2. : 1.));
      fdiv += simpson_w(r, q, pp) * w;
    }
    const double filter = VFS_CBRT(1. / d.s[S_AJ][p]);
    const double test_filter = VFS_CBRT(sum_weight);
    d.s[S_LFINV][p] = 1. / fdiv; d.s[S_LTF2][p] = test_filter * test_filter; d.s[S_LF2][p] = filter * filter;
  }
};

// les.c:470-669 given the 15 filtered sums f[0..8] = sum sw U_a u_b (a-major), f[9..14] = sum sw |S|S_ij
// and the precomputed factors above.  With A = [csi;eta;zet], the reference rotates the symmetric
// Cartesian tensor M^c to M = M^c A^T and contracts with the covariant metric tensor G = (A A^T)^-1:
//   MM = sum M_nm M_nl G_ml  = tr(A M^cT M^c A^T (A A^T)^-1) = |M^c|_F^2          (A^T (A A^T)^-1 A = I)
//   LM = sum L_ba M_aq G_bq  = tr(L M^c A^T (A A^T)^-1)      = tr(L M^c A^-1)     (A^T (A A^T)^-1 = A^-1)
// so neither M nor G is formed.  A^-1 = adj(A) / det(A) (rhs2.c:595-611) is built from the centre metrics the
// kernel holds anyway — nine cofactors, the reference's determinant expression and ONE division applied to the
// contracted sum — rather than read as nine more per-node operands (the marching kernel is bound by its
// L2 -> SM operand traffic, profiles/r01p).  Agrees with the reference to rounding (tests: <= 1e-12 on Cs, nu_t).
template <class Ops> VFS_HD void les2_finish_geo(const VfsDev &d, const Ops &O, int i, int j, int kg, long p, const double *f) {
  const double finv = O.geo(0), tf2 = O.geo(1), f2 = O.geo(2);
  double gh[3][3];
  grad_center_auto(d, O, i, j, kg, p, gh);
  const double S_hat = sabs_of(gh);
  const double tS = -tf2 * S_hat;
  // symmetric M^c: xx, xy, xz, yy, yz, zz
  const double mc0 = tS * (0.5 * (gh[0][0] + gh[0][0])) + f2 * (f[9] * finv), mc1 = tS * (0.5 * (gh[0][1] + gh[1][0])) + f2 * (f[10] * finv);
  const double mc2 = tS * (0.5 * (gh[0][2] + gh[2][0])) + f2 * (f[11] * finv), mc3 = tS * (0.5 * (gh[1][1] + gh[1][1])) + f2 * (f[12] * finv);
  const double mc4 = tS * (0.5 * (gh[1][2] + gh[2][1])) + f2 * (f[13] * finv), mc5 = tS * (0.5 * (gh[2][2] + gh[2][2])) + f2 * (f[14] * finv);
  const double den = (mc0 * mc0 + mc3 * mc3 + mc5 * mc5) + 2. * (mc1 * mc1 + mc2 * mc2 + mc4 * mc4);
  const double Mc[3][3] = {{mc0, mc1, mc2}, {mc1, mc3, mc4}, {mc2, mc4, mc5}};
  const double _u[3] = {O.u(0, 0, 0, 0), O.u(1, 0, 0, 0), O.u(2, 0, 0, 0)};
  const double a11 = O.met(0), a12 = O.met(1), a13 = O.met(2), a21 = O.met(3), a22 = O.met(4), a23 = O.met(5), a31 = O.met(6), a32 = O.met(7), a33 = O.met(8);
  const double det = a11 * (a33 * a22 - a32 * a23) - a21 * (a33 * a12 - a32 * a13) + a31 * (a23 * a12 - a22 * a13);
  // adj[c][b] = det * A^-1[c][b]: (x,y,z)_c by (csi,eta,zet)_b
  const double adj[3][3] = {{(a33 * a22 - a32 * a23), -(a33 * a12 - a32 * a13), (a23 * a12 - a22 * a13)},
                            {-(a33 * a21 - a31 * a23), (a33 * a11 - a31 * a13), -(a23 * a11 - a21 * a13)},
                            {(a32 * a21 - a31 * a22), -(a32 * a11 - a31 * a12), (a22 * a11 - a21 * a12)}};
  const double A[3][3] = {{a11, a12, a13}, {a21, a22, a23}, {a31, a32, a33}};
  double num = 0;
#pragma unroll
  for (int b = 0; b < 3; b++) {
    const double Ub = _u[0] * A[b][0] + _u[1] * A[b][1] + _u[2] * A[b][2];     // filtered contravariant velocity
    const double L0 = f[3 * b] * finv - Ub * _u[0], L1 = f[3 * b + 1] * finv - Ub * _u[1], L2 = f[3 * b + 2] * finv - Ub * _u[2];
#pragma unroll
    for (int c = 0; c < 3; c++) num += (L0 * Mc[0][c] + L1 * Mc[1][c] + L2 * Mc[2][c]) * adj[c][b];
  }
  d.s[S_LM][p] = num / det; d.s[S_MM][p] = den;
}

// les.c:308-669: Germano identity contracted with the covariant metric tensor -> LM, MM
// (straightforward one-thread-per-cell form; the shared-memory tiled form is in vfs_fused_kernels.h)
struct LesPass2 {
  VfsDev d;
  VFS_HD void operator()(int i, int j, int k) const {
    const int kg = k + d.kofs;
    const long p = d.idx(i, j, k);
    if (d.s[S_NV][p] > 1.1) { d.s[S_LM][p] = 0; d.s[S_MM][p] = 0; return; }
    double fs[VFS_LES2_NV], sum_weight = 0;
    for (int a = 0; a < VFS_LES2_NV; a++) fs[a] = 0;
    double fT[6] = {0, 0, 0, 0, 0, 0}, h2[3] = {0, 0, 0};
    if (d.clark) {                                     // the cell's own grid lengths weigh every stencil node's gradient (les.c:347-348, 420-428)
      double m9[9], dx, dy, dz;
      for (int a = 0; a < 9; a++) m9[a] = d.s[S_CSI0 + a][p];
      grid_lengths(d.s[S_AJ][p], m9, dx, dy, dz);
      h2[0] = dx * dx; h2[1] = dy * dy; h2[2] = dz * dz;
    }
    for (int r = -1; r <= 1; r++) for (int q = -1; q <= 1; q++) for (int pp = -1; pp <= 1; pp++) {
      const long n = p + r * d.sk + q * d.sj + pp;
      double v[VFS_LES2_NV];
      les2_products(d, n, v);
      if (d.clark) {
        double g[3][3];
        for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) g[a][b] = d.s[S_GR0 + 3 * a + b][n];
        const double cw = d.testfilter_ik ? (q != 0 ? 0. : (r == 0 ? 4. : 1.) * (pp == 0 ? 4. : 1.)) : simpson_w(r, q, pp) * v[0];
        int t = 0;
        for (int a = 0; a < 3; a++) for (int b = a; b < 3; b++, t++)
          fT[t] += cw * ((g[a][0] * g[b][0] * h2[0] + g[a][1] * g[b][1] * h2[1] + g[a][2] * g[b][2] * h2[2]) / 12.);
      }
      sum_weight += v[0] * (0.125 * (r == 0 ? 2. : 1.) * (q == 0 ? 2. : 1.) * (pp == 0 ? 2. : 1.));   // coef table les.c:441-451
      if (d.testfilter_ik) {
        if (q != 0) continue;
        const double c = (r == 0 ? 4. : 1.) * (pp == 0 ? 4. : 1.);   // 1,4,16 (rhs2.c:440)
        for (int a = 1; a < VFS_LES2_NV; a++) fs[a] += c * v[a];
      } else {
        const double sw = simpson_w(r, q, pp) * v[0];
        fs[0] += sw;
        for (int a = 1; a < VFS_LES2_NV; a++) fs[a] += sw * v[a];
      }
    }
    if (d.testfilter_ik) fs[0] = 36.;
    if (d.clark) les2_finish(d, i, j, kg, p, fs, sum_weight, fT, h2);
    else les2_finish(d, i, j, kg, p, fs, sum_weight);
  }
};

// les.c:716-796: Simpson-filter LM, MM (weights zeroed at solid cells and non-periodic domain
// ghosts, with the J==0-only quirk of les.c:756), C = 0.5 LM/(MM + 1e-4), Cs = max(C,0).
// REGULAR = true: the cell is not adjacent to a periodic boundary plane, so no neighbour index is
// remapped to a ghost image (les.c:738-768) and all fetches are at offsets -1..1.
struct GlobalAcc3 {
  const VfsDev &d; long p;
  VFS_HD double lm(int di, int dj, int dk) const { return d.s[S_LM][p + di + dj * d.sj + dk * d.sk]; }
  VFS_HD double mm(int di, int dj, int dk) const { return d.s[S_MM][p + di + dj * d.sj + dk * d.sk]; }
  VFS_HD double nv(int di, int dj, int dk) const { return d.s[S_NV][p + di + dj * d.sj + dk * d.sk]; }
  VFS_HD double iaj(int di, int dj, int dk) const { return d.s[S_IAJ][p + di + dj * d.sj + dk * d.sk]; }
};
template <bool REGULAR, class Acc> VFS_HD void les3_core(const VfsDev &d, const Acc &A, int i, int j, int kg, long p) {
  if (A.nv(0, 0, 0) > 1.1) { d.s[S_CS][p] = 0; return; }
  double LM_avg, MM_avg;
  if (d.testfilter_ik) {
    // integrate_testfilter_simpson defers to the weight-free i-k rule (rhs2.c:504-506)
    double l9 = 0, m9 = 0;
#pragma unroll
    for (int c = -1; c <= 1; c++)
#pragma unroll
      for (int a = -1; a <= 1; a++) {
        int I = i + a, K = kg + c, da = a, dc = c;
        if (!REGULAR) {
          if (d.perx) { if (I == 0) da = a - 2; else if (I == d.mx - 1) da = a + 2; }
          if (d.perz) { if (K == 0) dc = c - 2; else if (K == d.mz - 1) dc = c + 2; }
        }
        const double cf = (c == 0 ? 4. : 1.) * (a == 0 ? 4. : 1.);
        l9 += cf * A.lm(da, 0, dc); m9 += cf * A.mm(da, 0, dc);
      }
    LM_avg = l9 / 36.; MM_avg = m9 / 36.;
  } else {
    double ws = 0, lm = 0, mmv = 0;
#pragma unroll
    for (int c = -1; c <= 1; c++)
#pragma unroll
      for (int b = -1; b <= 1; b++)
#pragma unroll
        for (int a = -1; a <= 1; a++) {
          const int I = i + a, J = j + b, K = kg + c;
          double w = A.iaj(a, b, c);
          if (A.nv(a, b, c) > 1.1) w = 0;
          int da = a, db = b, dc = c;      // fetch offsets after the periodic remap
          if (d.perx) { if (!REGULAR) { if (I == 0) da = a - 2; else if (I == d.mx - 1) da = a + 2; } }
          else if (I == 0 || I == d.mx - 1) w = 0;
          if (d.pery) { if (!REGULAR) { if (J == 0) db = b - 2; else if (J == d.my - 1) db = b + 2; } }
          else if (J == 0) w = 0;
          if (d.perz) { if (!REGULAR) { if (K == 0) dc = c - 2; else if (K == d.mz - 1) dc = c + 2; } }
          else if (K == 0 || K == d.mz - 1) w = 0;
          const double sw = simpson_w(c, b, a) * w;
          ws += sw; lm += sw * A.lm(da, db, dc); mmv += sw * A.mm(da, db, dc);
        }
    LM_avg = lm / ws; MM_avg = mmv / ws;
  }
  const double C = 0.5 * LM_avg / (MM_avg + 1.e-4);
  // PetscMax(a,b) = (a<b)?b:a, PetscMin(a,b) = (a<b)?a:b, written out so that a NaN C (0/0 when every filter weight
  // around the cell is zero) ends like the reference's: max_cs, or 0.001 at an IB node
  double cs = (C < 0) ? 0 : C;                                   // les.c:795
  // clip chain of les.c:967-980 for interior nodes (boundary nodes are zeroed by LesClipBoundary)
  const double nvc = A.nv(0, 0, 0);
  if (nvc > 0.1 && nvc < 1.1) cs = (0.001 < cs) ? cs : 0.001;
  cs = (cs < 0) ? 0 : cs;
  cs = (cs < d.max_cs) ? cs : d.max_cs;
  d.s[S_CS][p] = cs;
}
struct LesPass3 {
  VfsDev d;
  VFS_HD void operator()(int i, int j, int k) const {
    const long p = d.idx(i, j, k);
    GlobalAcc3 A = {d, p};
    les3_core<false>(d, A, i, j, k + d.kofs, p);
  }
};


// ---- homogeneous-direction averaging of LM, MM (les.c:798-965) -----------------------------------------------
// i_homo_filter && k_homo_filter: Cs of EVERY interior cell of a j plane = 0.5 <LM> / (<MM> + les_eps), the averages
// taken over the plane's fluid cells (nvert < 0.1) of all ranks (MPI_Allreduce at :822-824).  One of i / j / k
// alone: the same along lines of that direction, written to the fluid cells only (:840-965).  The sums are the
// only reduction on the path: lines / planes are reduced on the device in a fixed order (results do not depend
// on the launch), the ranks' partial sums are added with ncclAllReduce — so unlike everything else the N-rank
// result equals the 1-rank result only to rounding (sum re-association; tolerance 1e-12 on Cs in the tests).
// mode: 0 = i and k, 1 = i, 2 = j, 3 = k.  `line` indexes the output: mode 0: j; 1: k_local*my + j; 2: k_local*mx + i; 3: j*mx + i.
struct HomoGeom { int mode, nline; };
VFS_HD void homo_cell(const VfsDev &d, long p, double acc[3]) {
  if (d.s[S_NV][p] < 0.1) { acc[0] += d.s[S_LM][p]; acc[1] += d.s[S_MM][p]; acc[2] += 1.0; }
}
// interior cells of this rank: i in [1, mx-1), j in [1, my-1), local k in [k1, k2)
VFS_HD void homo_line_serial(const VfsDev &d, int mode, int line, int k1, int k2, double acc[3]) {
  acc[0] = acc[1] = acc[2] = 0;
  if (mode == 0) { const int j = line; if (j < 1 || j > d.my - 2) return;
    for (int k = k1; k < k2; k++) for (int i = 1; i < d.mx - 1; i++) homo_cell(d, d.idx(i, j, k), acc); }
  else if (mode == 1) { const int k = line / d.my, j = line % d.my; if (k < k1 || k >= k2 || j < 1 || j > d.my - 2) return;
    for (int i = 1; i < d.mx - 1; i++) homo_cell(d, d.idx(i, j, k), acc); }
  else if (mode == 2) { const int k = line / d.mx, i = line % d.mx; if (k < k1 || k >= k2 || i < 1 || i > d.mx - 2) return;
    for (int j = 1; j < d.my - 1; j++) homo_cell(d, d.idx(i, j, k), acc); }
  else { const int j = line / d.mx, i = line % d.mx; if (i < 1 || i > d.mx - 2 || j < 1 || j > d.my - 2) return;
    for (int k = k1; k < k2; k++) homo_cell(d, d.idx(i, j, k), acc); }
}
// Cs from the averaged sums (buf[3*line + 0..2] = sum LM, sum MM, count over all ranks), then the clip chain of :967-980
struct HomoApply {
  VfsDev d; int mode; const double *buf;
  VFS_HD void operator()(int i, int j, int k) const {
    const long p = d.idx(i, j, k);
    const double nvc = d.s[S_NV][p];
    const int line = mode == 0 ? j : (mode == 1 ? k * d.my + j : (mode == 2 ? k * d.mx + i : j * d.mx + i));
    double cs;
    if (mode == 0 || nvc < 0.1) {
      double lm = buf[3 * line], mm = buf[3 * line + 1]; const double cnt = buf[3 * line + 2];
      if (cnt > 0) { lm /= cnt; mm /= cnt; }
      cs = 0.5 * lm / (mm + 1.e-4);
    } else {                        // cells the averages are not written to keep the UNFILTERED pointwise value (les.c:776-795)
      const double C = 0.5 * d.s[S_LM][p] / (d.s[S_MM][p] + 1.e-4);
      cs = (C < 0) ? 0 : C;
    }
    if (nvc > 1.1) cs = 0;
    else {
      if (nvc > 0.1 && nvc < 1.1) cs = (0.001 < cs) ? cs : 0.001;
      cs = (cs < 0) ? 0 : cs;
      cs = (cs < d.max_cs) ? cs : d.max_cs;
    }
    d.s[S_CS][p] = cs;
  }
};
#if defined(__CUDACC__) && !defined(VFS_EMU)
// modes 0 and 1: one block per line, the line's cells strided over the block, fixed-order tree reduction
__global__ void __launch_bounds__(256) k_homo_block(VfsDev d, int mode, int k1, int k2, double *buf) {
  const int line = blockIdx.x;
  double acc[3] = {0, 0, 0};
  const int ni = d.mx - 2;
  if (mode == 0) {
    const int j = line;
    if (j >= 1 && j <= d.my - 2) { const long n = (long)ni * (k2 - k1);
      for (long t = threadIdx.x; t < n; t += 256) homo_cell(d, d.idx(1 + (int)(t % ni), j, k1 + (int)(t / ni)), acc); }
  } else {
    const int k = line / d.my, j = line % d.my;
    if (k >= k1 && k < k2 && j >= 1 && j <= d.my - 2) for (int t = threadIdx.x; t < ni; t += 256) homo_cell(d, d.idx(1 + t, j, k), acc);
  }
  __shared__ double sm[3][256];
  for (int c = 0; c < 3; c++) sm[c][threadIdx.x] = acc[c];
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) { if (threadIdx.x < s) for (int c = 0; c < 3; c++) sm[c][threadIdx.x] += sm[c][threadIdx.x + s]; __syncthreads(); }
  if (threadIdx.x < 3) buf[3 * line + threadIdx.x] = sm[threadIdx.x][0];
}
// modes 2 and 3: one thread per line (consecutive lines = consecutive i: coalesced), serial along the line
__global__ void __launch_bounds__(256) k_homo_thread(VfsDev d, int mode, int nline, int k1, int k2, double *buf) {
  const int line = blockIdx.x * 256 + threadIdx.x;
  if (line >= nline) return;
  double acc[3];
  homo_line_serial(d, mode, line, k1, k2, acc);
  buf[3 * line] = acc[0]; buf[3 * line + 1] = acc[1]; buf[3 * line + 2] = acc[2];
}
#endif

// les.c:967-980: Cs = 0 on the domain-boundary nodes (the interior part of the clip chain is in les3_core)
struct LesClipBoundary {
  VfsDev d;
  VFS_HD void operator()(int i, int j, int k) const { d.s[S_CS][d.idx(i, j, k)] = 0; }
};

// nu_t of one interior cell from its Cs and |S| (les.c:1206-1211)
VFS_HD double nut_value(const VfsDev &d, long p, double cs, double Sabs) {
  const double *nv = d.s[S_NV];
  const double filter = VFS_CBRT(1. / d.s[S_AJ][p]);
  double v = cs * (filter * filter) * Sabs;
  if (d.wallfunction == 2 && nv[p] + nv[p + 1] + nv[p - 1] + nv[p + d.sj] + nv[p - d.sj] + nv[p + d.sk] + nv[p - d.sk] > 0.1) v = 0;
  return v;
}
// the same with Delta^2 = filter * filter taken from LesGeo's S_LF2 (the identical product, stored): no cbrt, no aj read
VFS_HD double nut_value_geo(const VfsDev &d, long p, double cs, double Sabs) {
  const double *nv = d.s[S_NV];
  double v = cs * d.s[S_LF2][p] * Sabs;
  if (d.wallfunction == 2 && nv[p] + nv[p + 1] + nv[p - 1] + nv[p + d.sj] + nv[p - d.sj] + nv[p + d.sk] + nv[p - d.sk] > 0.1) v = 0;
  return v;
}
// les.c:1185-1211: nu_t = Cs * Delta^2 * |S|.  FROM_S: |S| was stored by pass 1 of vfs_les_cs from
// the same ucat (identical arithmetic), so it is read back instead of being recomputed.
template <bool FROM_S> struct NuT {
  VfsDev d;
  VFS_HD void operator()(int i, int j, int k) const {
    const int kg = k + d.kofs;
    const long p = d.idx(i, j, k);
    const double *nv = d.s[S_NV];
    if (nv[p] > 1.1) { d.s[S_NUT][p] = 0; return; }
    double Sabs;
    if (FROM_S) Sabs = d.s[S_SABS][p];
    else { double g[3][3]; grad_center(d, S_U0, i, j, kg, p, g); Sabs = sabs_of(g); }
    d.s[S_NUT][p] = nut_value(d, p, d.s[S_CS][p], Sabs);
  }
};

#endif
