// vfs_rhs_kernels.h — Formfunction_2 (Source/momentum.c:454-2013) and the residual assembly of
// FormFunction_SNES (Source/momentum.c:2237-2336); difference helpers follow Compute_du_i/j/k
// (Source/k-omega.c:56-311, global `solid = 0.1`, k-omega.c:19).
//
// Staged formulation (round 1): face-flux kernel -> Fp kernel -> projection/assembly kernel with
// the face fluxes and Fp held in HBM work arrays, ghost refreshes in between exactly where the
// reference has DALocalToLocal.  Face metrics are recomputed from centre metrics (NEWMETRIC).
#ifndef VFS_RHS_KERNELS_H
#define VFS_RHS_KERNELS_H
#include "vfs_common.h"
#include "vfs_c2c_kernels.h"

#define VFS_SOLID 0.1

// Accessor over the global padded arrays: offsets (di,dj,dk) are relative to the face's node p;
// met/iaj/nut/uc give the centre metrics (scalar S_CSI0+s), 1/aj, nu_t at p (side 0) or p + e_D
// (side 1) and the contravariant flux component D at p + off*e_D.  The marching kernels
// (vfs_march_kernels.h) supply accessors with the same interface over TMA-staged shared-memory
// planes and exchange buffers, so every form runs the identical arithmetic below.
struct GlobalAcc {
  const VfsDev &d; long p;
  VFS_HD double u(int a, int di, int dj, int dk) const { return d.s[S_U0 + a][p + di + dj * d.sj + dk * d.sk]; }
  VFS_HD double nv(int di, int dj, int dk) const { return d.s[S_NV][p + di + dj * d.sj + dk * d.sk]; }
  template <int D> VFS_HD long sn() const { return D == 0 ? 1 : (D == 1 ? d.sj : d.sk); }
  template <int D> VFS_HD double met(int s, int side) const { return d.s[S_CSI0 + s][p + side * sn<D>()]; }
  template <int D> VFS_HD double iaj(int side) const { return d.s[S_IAJ][p + side * sn<D>()]; }
  template <int D> VFS_HD double nut(int side) const { return d.s[S_NUT][p + side * sn<D>()]; }
  template <int D> VFS_HD double uc(int off) const { return d.s[S_UC0 + D][p + off * sn<D>()]; }
  VFS_HD double wm() const { return d.s[S_WM][p]; }        // wall-model SGS viscosity of the j = 0 face (vfs_wm_kernels.h)
};

// tangential difference of component a along unit direction T at the face between node offset
// (0,0,0) and its neighbour in direction D  (k-omega.c:56-311, `solid` = 0.1)
template <int D, int T, class Acc> VFS_HD double dtan(const Acc &A, int a, const double solid = VFS_SOLID) {
  constexpr int ni = (D == 0), nj = (D == 1), nk = (D == 2);      // pn = p + n
  constexpr int ti = (T == 0), tj = (T == 1), tk = (T == 2);
  // The reference picks one of three stencils (k-omega.c:56-311):
  //   row +t solid : (u_n + u_0 - u_{n-t} - u_{-t}) / 2
  //   row -t solid : (u_{n+t} + u_t - u_n - u_0) / 2
  //   otherwise    : (u_{n+t} + u_t - u_{n-t} - u_{-t}) / 4
  // i.e. (P - M) * c with P, M pair sums chosen by two predicates.  All operands are fetched
  // unconditionally (independent loads, one memory round trip) and chosen by selects; pair sums
  // re-associate the reference's left-to-right sum (rounding-level difference, ~1e-16 relative).
  const double s0 = A.u(a, ni, nj, nk) + A.u(a, 0, 0, 0);
  const double sp = A.u(a, ni + ti, nj + tj, nk + tk) + A.u(a, ti, tj, tk);
  const double sm = A.u(a, ni - ti, nj - tj, nk - tk) + A.u(a, -ti, -tj, -tk);
  const bool hi = A.nv(ti, tj, tk) > solid || A.nv(ni + ti, nj + tj, nk + tk) > solid;
  const bool lo = !hi && (A.nv(-ti, -tj, -tk) > solid || A.nv(ni - ti, nj - tj, nk - tk) > solid);
  return ((hi ? s0 : sp) - (lo ? s0 : sm)) * ((hi || lo) ? 0.5 : 0.25);
}

// One face of family D (0/1/2 = i-/j-/k-face) between node p and p + e_D, stored at p ("upper
// integer node", momentum.c:508-509).  c = index of p along D (global), m = node count along D.
// REGULAR = true compiles out the domain-end / periodic-end special cases (faces 1..m-3 only).
// Metrics, nu_t and ucont come through the accessor (at p and p + e_D).
// X = true adds the rarely used variants, one-thread-per-face staged path only (never instantiated by the marching
// kernels, whose register budget they would break): WENO3 convection (`inviscid`, levelset_weno == 5; momentum.c:754-770,
// level.c:1352), the skew-symmetric form (`skew`: half the divergence form + the advective half in adv[], :789-800) and
// the Clark mixed-model term of the viscous flux (`clark`, :904-923).
// Calculate_dxdydz (rhs2.c:683-699): |sum of the three cell-edge vectors| per Cartesian direction, edge vector m = unit
// normal of metric family m (column m of the inverse metric matrix, normalised) times vol / |metric m|
VFS_HD void grid_lengths(double ajc, const double *m, double &dx, double &dy, double &dz) {
  const double a11 = m[0], a12 = m[1], a13 = m[2], a21 = m[3], a22 = m[4], a23 = m[5], a31 = m[6], a32 = m[7], a33 = m[8];
  const double det = a11 * (a33 * a22 - a32 * a23) - a21 * (a33 * a12 - a32 * a13) + a31 * (a23 * a12 - a22 * a13);
  double G[3][3];
  G[0][0] = (a33 * a22 - a32 * a23) / det; G[0][1] = -(a33 * a12 - a32 * a13) / det; G[0][2] = (a23 * a12 - a22 * a13) / det;
  G[1][0] = -(a33 * a21 - a31 * a23) / det; G[1][1] = (a33 * a11 - a31 * a13) / det; G[1][2] = -(a23 * a11 - a21 * a13) / det;
  G[2][0] = (a32 * a21 - a31 * a22) / det; G[2][1] = -(a32 * a11 - a31 * a12) / det; G[2][2] = (a22 * a11 - a21 * a12) / det;
  const double vol = 1. / ajc;
  double e[3][3];     // e[q][x]: edge vector of family q
  for (int q = 0; q < 3; q++) {
    const double nx = G[0][q], ny = G[1][q], nz = G[2][q];
    const double sum = sqrt(nx * nx + ny * ny + nz * nz);
    const double area = sqrt(m[3 * q] * m[3 * q] + m[3 * q + 1] * m[3 * q + 1] + m[3 * q + 2] * m[3 * q + 2]);
    const double L = vol / area;
    e[q][0] = L * (nx / sum); e[q][1] = L * (ny / sum); e[q][2] = L * (nz / sum);
  }
  dx = fabs(e[0][0] + e[1][0] + e[2][0]); dy = fabs(e[0][1] + e[1][1] + e[2][1]); dz = fabs(e[0][2] + e[1][2] + e[2][2]);
}
VFS_HD double weno3(double f0, double f1, double f2, double f3, double wavespeed) {      // level.c:1352-1388
  const double fL = wavespeed > 0 ? f0 : f3, fC = wavespeed > 0 ? f1 : f2, fR = wavespeed > 0 ? f2 : f1;
  const double d0 = 2. / 3., d1 = 1. / 3., eps = 1.e-6;
  const double beta0 = (fC - fR) * (fC - fR), beta1 = (fL - fC) * (fL - fC);          // pow(x, 2.) is exact x*x
  const double alpha0 = d0 / ((eps + beta0) * (eps + beta0)), alpha1 = d1 / ((eps + beta1) * (eps + beta1));
  const double sumalpha = alpha0 + alpha1;
  const double w0 = alpha0 / sumalpha, w1 = alpha1 / sumalpha;
  const double u0 = (fC * 0.5 + fR * 0.5), u1 = (-fL * 0.5 + fC * 1.5);
  return w0 * u0 + w1 * u1;
}
template <int D, bool REGULAR, class Acc, bool X = false>
VFS_HD void face_flux_core(const VfsDev &d, const Acc &A, int c, double fc[3], double fv[3], double *adv = nullptr) {
  constexpr int ni = (D == 0), nj = (D == 1), nk = (D == 2);
  const int m = (D == 0 ? d.mx : (D == 1 ? d.my : d.mz));
  const int per = (D == 0 ? d.perx : (D == 1 ? d.pery : d.perz));
  const double nvp = A.nv(0, 0, 0), nvn = A.nv(ni, nj, nk);

  // face metrics (metrics.c:589-592 and j/k twins): 0.5*centre + 0.5*centre, harmonic mean for aj
#define VFS_F3(s0) mk3(0.5 * A.template met<D>(s0, 0) + 0.5 * A.template met<D>(s0, 1), 0.5 * A.template met<D>(s0 + 1, 0) + 0.5 * A.template met<D>(s0 + 1, 1), \
                       0.5 * A.template met<D>(s0 + 2, 0) + 0.5 * A.template met<D>(s0 + 2, 1))
  const V3 cs = VFS_F3(0), et = VFS_F3(3), ze = VFS_F3(6);
#undef VFS_F3
  // 2/(1/aj_p + 1/aj_n) with the stored 1/aj (S_IAJ = 1./aj, the same correctly-rounded quotient): 1 division, not 3
  const double ajc = 2. / (A.template iaj<D>(0) + A.template iaj<D>(1));
  const V3 n = (D == 0 ? cs : (D == 1 ? et : ze));

  // du[a][b] = d u_a / d xi_b  (b: csi, eta, zet)
  double du[3][3];
#pragma unroll
  for (int a = 0; a < 3; a++) {
    const double dn = A.u(a, ni, nj, nk) - A.u(a, 0, 0, 0);
    if (D == 0) { du[a][0] = dn; du[a][1] = dtan<D, 1>(A, a); du[a][2] = dtan<D, 2>(A, a); }
    else if (D == 1) { du[a][0] = dtan<D, 0>(A, a); du[a][1] = dn; du[a][2] = dtan<D, 2>(A, a); }
    else { du[a][0] = dtan<D, 0>(A, a); du[a][1] = dtan<D, 1>(A, a); du[a][2] = dn; }
  }
  const double g1 = cs.x * n.x + cs.y * n.y + cs.z * n.z;
  const double g2 = et.x * n.x + et.y * n.y + et.z * n.z;
  const double g3 = ze.x * n.x + ze.y * n.y + ze.z * n.z;
  // r[a][b] = du_a/dx_b * J   (momentum.c:686-696)
  double r[3][3];
#pragma unroll
  for (int a = 0; a < 3; a++) {
    r[a][0] = du[a][0] * cs.x + du[a][1] * et.x + du[a][2] * ze.x;
    r[a][1] = du[a][0] * cs.y + du[a][1] * et.y + du[a][2] * ze.y;
    r[a][2] = du[a][0] * cs.z + du[a][1] * et.z + du[a][2] * ze.z;
  }

  // ---- convective flux (momentum.c:700-813) ----
  // offsets along D of the outer stencil nodes; regular faces use fixed offsets and a select
  // (`coll`: the stencil collapses to the two face nodes), so no accessor is indexed dynamically
  int oL = -1, oR = 2;
  bool coll = false;
  if (!REGULAR && (c == 0 || c == m - 2)) {
    if (per && c == m - 2) oR = 4;                       // index m+2
    else if (per && c == 0) oL = -3;                     // index -3
    else oL = 0, oR = 1;
  } else if (A.nv(-ni, -nj, -nk) + A.nv(2 * ni, 2 * nj, 2 * nk) > 0.1) { oL = 0, oR = 1; coll = true; }
  if (d.second_order) { oL = 0, oR = 1; coll = true; }

  const double uc0 = A.template uc<D>(0);
  double ucon = uc0;
  if (!REGULAR) {
    if (per && c == 0) ucon = A.template uc<D>(-2);
    if (D == 2 && c == m - 2 && d.bc[5] == 4 && (int)nvp == 0) ucon = A.template uc<D>(-1);
  }
  const double up = -0.5 * (ucon + fabs(ucon));
  const double um = -0.5 * (ucon - fabs(ucon));
  if (d.immersed && (REGULAR || c != m - 2) && nvp > 0.1) {
#pragma unroll
    for (int a = 0; a < 3; a++) {
      const double u0 = A.u(a, 0, 0, 0), u1 = A.u(a, ni, nj, nk), u2 = A.u(a, 2 * ni, 2 * nj, 2 * nk);
      fc[a] = um * (0.125 * (-u2 - 2. * u1 + 3. * u0) + u1) + up * (0.125 * (-u0 - 2. * u0 + 3. * u1) + u0);
    }
  } else if (d.immersed && (REGULAR || c != 0) && nvn > 0.1) {
#pragma unroll
    for (int a = 0; a < 3; a++) {
      const double u0 = A.u(a, 0, 0, 0), u1 = A.u(a, ni, nj, nk), um1 = A.u(a, -ni, -nj, -nk);
      fc[a] = um * (0.125 * (-u1 - 2. * u1 + 3. * u0) + u1) + up * (0.125 * (-um1 - 2. * u0 + 3. * u1) + u0);
    }
  } else if (X && d.weno) {
    for (int a = 0; a < 3; a++)
      fc[a] = -uc0 * weno3(A.u(a, oL * ni, oL * nj, oL * nk), A.u(a, 0, 0, 0), A.u(a, ni, nj, nk), A.u(a, oR * ni, oR * nj, oR * nk), uc0);
  } else if (d.second_order) {
#pragma unroll
    for (int a = 0; a < 3; a++) fc[a] = -uc0 * 0.5 * (A.u(a, 0, 0, 0) + A.u(a, ni, nj, nk));
  } else {
#pragma unroll
    for (int a = 0; a < 3; a++) {
      const double u0 = A.u(a, 0, 0, 0), u1 = A.u(a, ni, nj, nk);
      double uL, uR;
      if (REGULAR) { uL = A.u(a, -ni, -nj, -nk); uR = A.u(a, 2 * ni, 2 * nj, 2 * nk); uL = coll ? u0 : uL; uR = coll ? u1 : uR; }
      else { uL = A.u(a, oL * ni, oL * nj, oL * nk); uR = A.u(a, oR * ni, oR * nj, oR * nk); }
      fc[a] = -uc0 * 0.0625 * (-uL + 9. * u0 + 9. * u1 - uR);
    }
  }
  if (X && d.skew) {
    // denom: 3, or 1 at a non-periodic domain end and with second_order; the nvert collapse of the stencil resets it
    // in the k direction only (momentum.c:734 / 996 against 1258)
    double denom = 3.;
    if ((c == 0 || c == m - 2) && !per) denom = 1.;
    else if (D == 2 && !(c == 0 || c == m - 2) && coll) denom = 1.;
    if (d.second_order) denom = 1.;
    for (int a = 0; a < 3; a++) {
      fc[a] *= 0.5;
      const double d4 = (A.u(a, oR * ni, oR * nj, oR * nk) - A.u(a, oL * ni, oL * nj, oL * nk)) * (1. / denom);
      adv[a] = -0.5 * uc0 * (9. / 8. * du[a][D] - 1. / 8. * d4);
    }
  }
  if (nvp + nvn > 0.1 && (d.immersed == 3 || !d.immersed)) {
    fc[0] = fc[1] = fc[2] = 0;
    if (X && d.skew) adv[0] = adv[1] = adv[2] = 0;
  }

  // ---- viscous + SGS flux (momentum.c:856-902) ----
  const double nu = 1. / d.ren;
  fv[0] = fv[1] = fv[2] = 0;
  if (d.les) {
    double nu_t;
    if ((!REGULAR && c == 0 && !per) || nvp > 0.1) nu_t = A.template nut<D>(1);
    else if ((!REGULAR && c == m - 2 && !per) || nvn > 0.1) nu_t = A.template nut<D>(0);
    else nu_t = 0.5 * (A.template nut<D>(0) + A.template nut<D>(1));
    if constexpr (!REGULAR && D == 1) { if (c == 0 && d.visc_wm) nu_t = A.wm(); }     // momentum.c:1139-1154
#pragma unroll
    for (int a = 0; a < 3; a++)
      fv[a] = (g1 * du[a][0] + g2 * du[a][1] + g3 * du[a][2] + r[0][a] * n.x + r[1][a] * n.y + r[2][a] * n.z) * ajc * nu_t;
  }
  if (d.laplacian) {
#pragma unroll
    for (int a = 0; a < 3; a++)
      fv[a] += (g1 * du[a][0] + g2 * du[a][1] + g3 * du[a][2] + 0. * n.x + 0. * n.y + 0. * n.z) * ajc * nu;
  } else {
#pragma unroll
    for (int a = 0; a < 3; a++)
      fv[a] += (g1 * du[a][0] + g2 * du[a][1] + g3 * du[a][2] + r[0][a] * n.x + r[1][a] * n.y + r[2][a] * n.z) * ajc * nu;
  }
  if (X && d.clark) {
    // grid lengths from the CENTRE metrics of node p with the face Jacobian (Calculate_dxdydz, rhs2.c:683-699)
    double m9[9];
    for (int q = 0; q < 9; q++) m9[q] = A.template met<D>(q, 0);
    double dx, dy, dz;
    grid_lengths(ajc, m9, dx, dy, dz);
    const double h2[3] = {dx * dx, dy * dy, dz * dz};
    double g[3][3];                       // g[a][b] = du_a/dx_b (Compute_du_dxyz)
    for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) g[a][b] = r[a][b] * ajc;
    for (int a = 0; a < 3; a++) {
      double t[3];
      for (int b = 0; b < 3; b++) t[b] = (g[a][0] * g[b][0] * h2[0] + g[a][1] * g[b][1] * h2[1] + g[a][2] * g[b][2] * h2[2]);
      fv[a] -= (t[0] * n.x + t[1] * n.y + t[2] * n.z) / 12.;
    }
  }
}

template <int D> struct FaceFlux {
  VfsDev d;
  VFS_HD void operator()(int i, int j, int k) const {
    const int c = (D == 0 ? i : (D == 1 ? j : k + d.kofs));
    const long p = d.idx(i, j, k);
    GlobalAcc A = {d, p};
    double fc[3], fv[3];
    face_flux_core<D, false>(d, A, c, fc, fv);
    const int sc = S_FC1 + 3 * D, sv = S_FV1 + 3 * D;
    for (int a = 0; a < 3; a++) { d.s[sc + a][p] = fc[a]; d.s[sv + a][p] = fv[a]; }
  }
};

// the variants' form of the same kernel (weno / skew / clark, see face_flux_core): also stores the advective half Adv1-3
template <int D> struct FaceFluxX {
  VfsDev d;
  VFS_HD void operator()(int i, int j, int k) const {
    const int c = (D == 0 ? i : (D == 1 ? j : k + d.kofs));
    const long p = d.idx(i, j, k);
    GlobalAcc A = {d, p};
    double fc[3], fv[3], adv[3] = {0, 0, 0};
    face_flux_core<D, false, GlobalAcc, true>(d, A, c, fc, fv, adv);
    const int sc = S_FC1 + 3 * D, sv = S_FV1 + 3 * D, sa = S_ADV1 + 3 * D;
    for (int a = 0; a < 3; a++) { d.s[sc + a][p] = fc[a]; d.s[sv + a][p] = fv[a]; }
    if (d.skew) for (int a = 0; a < 3; a++) d.s[sa + a][p] = adv[a];
  }
};


// ---- legacy Convection / Viscous (Source/rhs.c:751-1069 and :1071-1582) -------------
// The explicit solvers' (RungeKutta / timeadvancing1.c FormFunctionSNES) Cartesian right-hand-side
// pieces: QUICK flux-difference convection and the (nu + nu_t) viscous term with the inline stencil
// switch at `solid = 0.5` (rhs.c:1105).  Neither handles periodic directions (the reference code does
// not); face fluxes land in the same work arrays as Formfunction_2's (Div1-3 / Visc1-3 twins), the
// results in two free work triplets exposed as the public fields VFS_CONV / VFS_VISC.
template <int D, class Acc> VFS_HD void legacy_conv_flux(const VfsDev &d, const Acc &A, int c, double f[3]) {
  constexpr int ni = (D == 0), nj = (D == 1), nk = (D == 2);
  const int m = (D == 0 ? d.mx : (D == 1 ? d.my : d.mz));
  const double coef = 0.125;
  const double ucon = A.template uc<D>(0) * 0.5;
  const double up = ucon + fabs(ucon), um = ucon - fabs(ucon);
  const double nvm = A.nv(-ni, -nj, -nk), nvq = A.nv(ni, nj, nk);
  int br;          // 0: interior stencil, 1: low side collapsed, 2: high side collapsed, 3: untouched (rhs.c:886-935, 1003-1040)
  if (c > 0 && c < m - 2 && nvq < 0.1 && nvm < 0.1) br = 0;
  else if (D == 2 ? (c < m - 2 && (c == 0 || nvm > 0.1)) : (c == 0 || nvm > 0.1)) br = 1;
  else if (D == 2 ? (c > 0 && (c == m - 2 || nvq > 0.1)) : (c == m - 2 || nvq > 0.1)) br = 2;
  else br = 3;
#pragma unroll
  for (int a = 0; a < 3; a++) {
    const double u0 = A.u(a, 0, 0, 0), u1 = A.u(a, ni, nj, nk);
    const double uR = br == 2 ? u1 : A.u(a, 2 * ni, 2 * nj, 2 * nk);
    const double uL = br == 1 ? u0 : A.u(a, -ni, -nj, -nk);
    const double v = um * (coef * (-uR - 2. * u1 + 3. * u0) + u1) + up * (coef * (-uL - 2. * u0 + 3. * u1) + u0);
    f[a] = br == 3 ? 0. : v;
  }
}
template <int D, class Acc> VFS_HD void legacy_visc_flux(const VfsDev &d, const Acc &A, double f[3]) {
  constexpr int ni = (D == 0), nj = (D == 1), nk = (D == 2);
  const double solid = 0.5;                                    // rhs.c:1105
#define VFS_F3(s0) mk3(0.5 * A.template met<D>(s0, 0) + 0.5 * A.template met<D>(s0, 1), 0.5 * A.template met<D>(s0 + 1, 0) + 0.5 * A.template met<D>(s0 + 1, 1), \
                       0.5 * A.template met<D>(s0 + 2, 0) + 0.5 * A.template met<D>(s0 + 2, 1))
  const V3 cs = VFS_F3(0), et = VFS_F3(3), ze = VFS_F3(6);
#undef VFS_F3
  const double ajc = 2. / (A.template iaj<D>(0) + A.template iaj<D>(1));
  const V3 n = (D == 0 ? cs : (D == 1 ? et : ze));
  double du[3][3];
#pragma unroll
  for (int a = 0; a < 3; a++) {
    const double dn = A.u(a, ni, nj, nk) - A.u(a, 0, 0, 0);
    if (D == 0) { du[a][0] = dn; du[a][1] = dtan<D, 1>(A, a, solid); du[a][2] = dtan<D, 2>(A, a, solid); }
    else if (D == 1) { du[a][0] = dtan<D, 0>(A, a, solid); du[a][1] = dn; du[a][2] = dtan<D, 2>(A, a, solid); }
    else { du[a][0] = dtan<D, 0>(A, a, solid); du[a][1] = dtan<D, 1>(A, a, solid); du[a][2] = dn; }
  }
  const double g1 = cs.x * n.x + cs.y * n.y + cs.z * n.z;
  const double g2 = et.x * n.x + et.y * n.y + et.z * n.z;
  const double g3 = ze.x * n.x + ze.y * n.y + ze.z * n.z;
  double r[3][3];
#pragma unroll
  for (int a = 0; a < 3; a++) {
    r[a][0] = du[a][0] * cs.x + du[a][1] * et.x + du[a][2] * ze.x;
    r[a][1] = du[a][0] * cs.y + du[a][1] * et.y + du[a][2] * ze.y;
    r[a][2] = du[a][0] * cs.z + du[a][1] * et.z + du[a][2] * ze.z;
  }
  const double nu = 1. / d.ren;
  const double nu_t = d.les ? 0.5 * (A.template nut<D>(0) + A.template nut<D>(1)) : 0.;      // rhs.c:1220-1226
#pragma unroll
  for (int a = 0; a < 3; a++)
    f[a] = (g1 * du[a][0] + g2 * du[a][1] + g3 * du[a][2] + r[0][a] * n.x + r[1][a] * n.y + r[2][a] * n.z) * ajc * (nu + nu_t);
}
// VISC = false: Convection's Fp1-3 into S_FC1.., VISC = true: Viscous's into S_FV1..
template <int D, bool VISC> struct LegacyFlux {
  VfsDev d;
  VFS_HD void operator()(int i, int j, int k) const {
    const long p = d.idx(i, j, k);
    GlobalAcc A = {d, p};
    double f[3];
    if (VISC) legacy_visc_flux<D>(d, A, f);
    else legacy_conv_flux<D>(d, A, (D == 0 ? i : (D == 1 ? j : k + d.kofs)), f);
    const int s0 = (VISC ? S_FV1 : S_FC1) + 3 * D;
    for (int a = 0; a < 3; a++) d.s[s0 + a][p] = f[a];
  }
};
// flux difference at the cell centres (rhs.c:1044-1061, 1487-1493)
struct LegacyDiv {
  VfsDev d; int sf, so;
  VFS_HD void operator()(int i, int j, int k) const {
    const long p = d.idx(i, j, k);
    for (int a = 0; a < 3; a++)
      d.s[so + a][p] = d.s[sf + a][p] - d.s[sf + a][p - 1] + d.s[sf + 3 + a][p] - d.s[sf + 3 + a][p - d.sj] + d.s[sf + 6 + a][p] - d.s[sf + 6 + a][p - d.sk];
  }
};


// ---- Pressure_Gradient (Source/momentum.c:203-439) ------------------------------------
// dP_D = (dP/dcsi g_D1 + dP/deta g_D2 + dP/dzeta g_D3) * J_face on the D-face between p and p + e_D, with
// g_Dm = (metric m at the face) . (metric D at the face); the difference along D is two-point (the periodic image
// at index m+1 for the last face, :347,372,397), the tangential ones are the four-point averages that fall back to
// one-sided pairs next to an IB node ((int)(nvert + 0.5) == 1) or a non-periodic domain end (:349-395).  Face
// metrics are rebuilt from the centre metrics (NEWMETRIC), as in the flux kernels.
struct PressureGradient {
  VfsDev d; double kforce;
  VFS_HD bool ib(long q) const { return (int)(d.s[S_NV][q] + 0.5) == 1; }
  // tangential difference along T (stride t, index c of m, periodicity per) on the face between p and p + n
  VFS_HD double tdiff(long p, long n, long t, int c, int m, int per) const {
    const double *P = d.s[S_P];
    if (ib(p + t) || ib(p + t + n) || (c == m - 2 && !per)) return (P[p] - P[p - t] + P[p + n] - P[p + n - t]) * 0.5;
    if (ib(p - t) || ib(p - t + n) || (c == 1 && !per)) return (P[p + t] - P[p] + P[p + t + n] - P[p + n]) * 0.5;
    return (P[p + t] - P[p - t] + P[p + n + t] - P[p + n - t]) * 0.25;
  }
  VFS_HD double ndiff(long p, long n, int c, int m, int per) const {
    const double *P = d.s[S_P];
    return (c == m - 2 && per) ? P[p + 3 * n] - P[p] : P[p + n] - P[p];
  }
  template <int D> VFS_HD void face(long p, V3 &cs, V3 &et, V3 &ze, double &aj) const {
    GlobalAcc A = {d, p};
#define VFS_F3(s0) mk3(0.5 * A.template met<D>(s0, 0) + 0.5 * A.template met<D>(s0, 1), 0.5 * A.template met<D>(s0 + 1, 0) + 0.5 * A.template met<D>(s0 + 1, 1), \
                       0.5 * A.template met<D>(s0 + 2, 0) + 0.5 * A.template met<D>(s0 + 2, 1))
    cs = VFS_F3(0); et = VFS_F3(3); ze = VFS_F3(6);
#undef VFS_F3
    aj = 2. / (A.template iaj<D>(0) + A.template iaj<D>(1));
  }
  VFS_HD void operator()(int i, int j, int k) const {
    const int kg = k + d.kofs;
    const long p = d.idx(i, j, k), si = 1, sj = d.sj, sk = d.sk;
    const double *nv = d.s[S_NV];
    V3 cs, et, ze; double aj;
    {
      face<0>(p, cs, et, ze, aj);
      const double dc = ndiff(p, si, i, d.mx, d.perx), de = tdiff(p, si, sj, j, d.my, d.pery), dz = tdiff(p, si, sk, kg, d.mz, d.perz);
      double v = (dc * dot3(cs, cs) + de * dot3(et, cs) + dz * dot3(ze, cs)) * aj / 1.0;
      if (nv[p] + nv[p + si] > 0.1 || (!d.perx && i == d.mx - 2)) v = 0;
      d.s[S_DP0][p] = v;
    }
    {
      face<1>(p, cs, et, ze, aj);
      const double dc = tdiff(p, sj, si, i, d.mx, d.perx), de = ndiff(p, sj, j, d.my, d.pery), dz = tdiff(p, sj, sk, kg, d.mz, d.perz);
      double v = (dc * dot3(cs, et) + de * dot3(et, et) + dz * dot3(ze, et)) * aj / 1.0;
      if (nv[p] + nv[p + sj] > 0.1 || (!d.pery && j == d.my - 2)) v = 0;
      d.s[S_DP1][p] = v;
    }
    {
      face<2>(p, cs, et, ze, aj);
      const double dc = tdiff(p, sk, si, i, d.mx, d.perx), de = tdiff(p, sk, sj, j, d.my, d.pery);
      double dz = ndiff(p, sk, kg, d.mz, d.perz);
      if (d.perz) dz += kforce * (1. / aj / sqrt(dot3(ze, ze)));          // momentum.c:399-411
      double v = (dc * dot3(cs, ze) + de * dot3(et, ze) + dz * dot3(ze, ze)) * aj / 1.0;
      if (nv[p] + nv[p + sk] > 0.1 || (!d.perz && kg == d.mz - 2)) v = 0;
      d.s[S_DP2][p] = v;
    }
  }
};

// ---- UpdatePressure / Projection (Source/poisson.c:3137-3296, 2700-3040) -------------------------
// P += Phi at fluid cells, 0 at solid cells (:3191-3193)
struct UpdateP {
  VfsDev d;
  VFS_HD void operator()(int i, int j, int k) const {
    const long p = d.idx(i, j, k);
    const double nv = d.s[S_NV][p];
    if (nv < 0.1) d.s[S_P][p] += d.s[S_PHI][p];
    if (nv > 1.1) d.s[S_P][p] = 0;
  }
};
// Ucont -= dt * st * (contravariant gradient of Phi) on the three faces of cell (i,j,k) (poisson.c:2790-2960).  Same face
// geometry as PressureGradient; the tangential differences fall back to one-sided pairs next to a non-periodic domain
// end or a masked pair of cells (nvert sum > poisson_threshold) and to zero when both sides are blocked.
struct ProjectionCorr {
  VfsDev d; double st, thr;           // `* dt * st / coeff / r`: time_coeff() and the density factors are 1 on this path
  VFS_HD double tdiff(long p, long n, long t, int c, int m, int per) const {
    const double *F = d.s[S_PHI], *nv = d.s[S_NV];
    if ((c == m - 2 && !per) || nv[p + t] + nv[p + t + n] > thr) {
      if (nv[p - t] + nv[p - t + n] < thr && c != 1) return (F[p] + F[p + n] - F[p - t] - F[p - t + n]) * 0.5;
      return 0.;
    }
    if ((c == 1 && !per) || nv[p - t] + nv[p - t + n] > thr) {
      if (nv[p + t] + nv[p + t + n] < thr) return (F[p + t] + F[p + t + n] - F[p] - F[p + n]) * 0.5;
      return 0.;
    }
    return (F[p + t] + F[p + t + n] - F[p - t] - F[p - t + n]) * 0.25;
  }
  VFS_HD double ndiff(long p, long n, int c, int m, int per) const {
    const double *F = d.s[S_PHI];
    return (c == m - 2 && per) ? F[p + 3 * n] - F[p] : F[p + n] - F[p];
  }
  VFS_HD void operator()(int i, int j, int k) const {
    const int kg = k + d.kofs;
    const long p = d.idx(i, j, k), si = 1, sj = d.sj, sk = d.sk;
    const double *nv = d.s[S_NV];
    PressureGradient G = {d, 0.};
    V3 cs, et, ze; double aj;
    if (i < d.mx - 2 || d.perx) {
      const double dc = ndiff(p, si, i, d.mx, d.perx), de = tdiff(p, si, sj, j, d.my, d.pery), dz = tdiff(p, si, sk, kg, d.mz, d.perz);
      if (nv[p] + nv[p + si] < thr) {
        G.face<0>(p, cs, et, ze, aj);
        d.s[S_UC0][p] -= (dc * dot3(cs, cs) * aj + de * dot3(et, cs) * aj + dz * dot3(ze, cs) * aj) * d.dt * st;
      }
    }
    if (j < d.my - 2 || d.pery) {
      const double dc = tdiff(p, sj, si, i, d.mx, d.perx), de = ndiff(p, sj, j, d.my, d.pery), dz = tdiff(p, sj, sk, kg, d.mz, d.perz);
      if (nv[p] + nv[p + sj] < thr || (int)(nv[p] + nv[p + sj]) == 5) {
        G.face<1>(p, cs, et, ze, aj);
        d.s[S_UC1][p] -= (dc * dot3(cs, et) * aj + de * dot3(et, et) * aj + dz * dot3(ze, et) * aj) * d.dt * st;
      }
    }
    if (kg < d.mz - 2 || d.perz) {
      const double dc = tdiff(p, sk, si, i, d.mx, d.perx), de = tdiff(p, sk, sj, j, d.my, d.pery), dz = ndiff(p, sk, kg, d.mz, d.perz);
      if (nv[p] + nv[p + sk] < thr) {
        G.face<2>(p, cs, et, ze, aj);
        d.s[S_UC2][p] -= (dc * dot3(cs, ze) * aj + de * dot3(et, ze) * aj + dz * dot3(ze, ze) * aj) * d.dt * st;
      }
    }
  }
};
// the component-wise periodic copies that follow (poisson.c:2997-3025; the same rule as IB_BC's, momentum.c:2210-2221)
struct PeriodicCompCopy {
  VfsDev d;
  VFS_HD void operator()(int i, int j, int k) const {
    const int mx = d.mx, my = d.my, mz = d.mz, kg = k + d.kofs;
    const long p = d.idx(i, j, k);
    long q = p; bool fi = false, fj = false, fk = false;
    if (d.perx && i == 0) q += -2, fi = true; else if (d.perx && i == mx - 1) q += 2, fi = true;
    if (d.pery && j == 0) q += -2 * d.sj, fj = true; else if (d.pery && j == my - 1) q += 2 * d.sj, fj = true;
    if (d.perz && kg == 0) q += -2 * d.sk, fk = true; else if (d.perz && kg == mz - 1) q += 2 * d.sk, fk = true;
    if (fi) d.s[S_UC0][p] = d.s[S_UC0][q];
    if (fj) d.s[S_UC1][p] = d.s[S_UC1][q];
    if (fk) d.s[S_UC2][p] = d.s[S_UC2][q];
  }
};

// ---- body-fitted cylinder diagnostics of Formfunction_2 (momentum.c:570-579, 822-849; bctype[0] == 11, bctype[1] == 1) ----
// Per i-face of the wall plane i = mx-2: face area A = |icsi|, |icsi.x|, |icsi.z| and the pressure / viscous force
// components along x and z with the inward unit normal (the first column of the inverse metric matrix, normalised,
// rhs2.c:649-680) — seven numbers per face, written to out[q * nface + (k * my + j)], zero where the loop does not run.
// The host adds them up in the reference's loop order (k, then j).  Face gradients as in face_flux_core.
struct CylinderForce {
  VfsDev d; double *out; long nface;
  VFS_HD void operator()(int i, int j, int k) const {
    const long p = d.idx(i, j, k);
    GlobalAcc A = {d, p};
#define VFS_F3(s0) mk3(0.5 * A.template met<0>(s0, 0) + 0.5 * A.template met<0>(s0, 1), 0.5 * A.template met<0>(s0 + 1, 0) + 0.5 * A.template met<0>(s0 + 1, 1), \
                       0.5 * A.template met<0>(s0 + 2, 0) + 0.5 * A.template met<0>(s0 + 2, 1))
    const V3 cs = VFS_F3(0), et = VFS_F3(3), ze = VFS_F3(6);
#undef VFS_F3
    const double ajc = 2. / (A.template iaj<0>(0) + A.template iaj<0>(1));
    double du[3][3];
    for (int a = 0; a < 3; a += 2) { du[a][0] = A.u(a, 1, 0, 0) - A.u(a, 0, 0, 0); du[a][1] = dtan<0, 1>(A, a); du[a][2] = dtan<0, 2>(A, a); }
    const double du_dx = (du[0][0] * cs.x + du[0][1] * et.x + du[0][2] * ze.x) * ajc, du_dz = (du[0][0] * cs.z + du[0][1] * et.z + du[0][2] * ze.z) * ajc;
    const double dw_dx = (du[2][0] * cs.x + du[2][1] * et.x + du[2][2] * ze.x) * ajc, dw_dz = (du[2][0] * cs.z + du[2][1] * et.z + du[2][2] * ze.z) * ajc;
    const double Sxx = 0.5 * (du_dx + du_dx), Sxz = 0.5 * (du_dz + dw_dx), Szz = 0.5 * (dw_dz + dw_dz);
    const double Ai = sqrt(cs.x * cs.x + cs.y * cs.y + cs.z * cs.z);
    const double a11 = cs.x, a12 = cs.y, a13 = cs.z, a21 = et.x, a22 = et.y, a23 = et.z, a31 = ze.x, a32 = ze.y, a33 = ze.z;
    const double det = a11 * (a33 * a22 - a32 * a23) - a21 * (a33 * a12 - a32 * a13) + a31 * (a23 * a12 - a22 * a13);
    double nx = (a33 * a22 - a32 * a23) / det, ny = -(a33 * a21 - a31 * a23) / det, nz = (a32 * a21 - a31 * a22) / det;
    const double sum = sqrt(nx * nx + ny * ny + nz * nz);
    nx /= sum; ny /= sum; nz /= sum;
    nx *= -1; ny *= -1; nz *= -1;
    const double P = 2 * d.s[S_P][p - 1] - d.s[S_P][p];
    const long f = (long)k * d.my + j;
    out[0 * nface + f] = Ai; out[1 * nface + f] = fabs(cs.x); out[2 * nface + f] = fabs(cs.z);
    out[3 * nface + f] = -P * nx * Ai; out[4 * nface + f] = -P * nz * Ai;
    out[5 * nface + f] = 2.0 * (Sxx * nx + 0 * ny + Sxz * nz) * Ai / d.ren;
    out[6 * nface + f] = 2.0 * (Sxz * nx + 0 * ny + Szz * nz) * Ai / d.ren;
  }
};

// momentum.c:1548-1678: flux divergence with 4th-order correction + viscous divergence -> Fp of the cell at
// node p = (i, j, k); kg = global k of the cell (for a ghost plane across the periodic seam: the plane it images)
template <bool X = false> VFS_HD void fp_cell_value(const VfsDev &d, int i, int j, int kg, long p, double out[3]) {
  const long st[3] = {1, d.sj, d.sk};
  const int cc[3] = {i, j, kg}, mm[3] = {d.mx, d.my, d.mz}, pp[3] = {d.perx, d.pery, d.perz};
  const double *nv = d.s[S_NV];
  double div[3] = {0, 0, 0}, div4[3] = {0, 0, 0}, vis[3] = {0, 0, 0};
  // accumulate in the reference's order: i-, j-, k-family
  for (int a = 0; a < 3; a++) {
    div[a] = (d.s[S_FC1 + a][p] - d.s[S_FC1 + a][p - 1] + d.s[S_FC2 + a][p] - d.s[S_FC2 + a][p - d.sj] + d.s[S_FC3 + a][p] - d.s[S_FC3 + a][p - d.sk]);
    vis[a] = (d.s[S_FV1 + a][p] - d.s[S_FV1 + a][p - 1] + d.s[S_FV2 + a][p] - d.s[S_FV2 + a][p - d.sj] + d.s[S_FV3 + a][p] - d.s[S_FV3 + a][p - d.sk]);
  }
  if (X && d.inviscid) {                      // momentum.c:1624, 1663: second-order divergence only, no viscous term, no advective half
    for (int a = 0; a < 3; a++) out[a] = div[a];
  } else if (X && d.skew) {
    // momentum.c:1626-1651: the divergence half as below, plus the advective half averaged to the cell with the same
    // 9/8, -1/8 weights over the same outer faces (with second_order the outer faces collapse onto the cell's own)
    double adv[3][3];
    for (int D = 0; D < 3; D++) {
      const int c = cc[D], m = mm[D], per = pp[D];
      const long s = st[D];
      long pR = p + s, pL = p - 2 * s;
      double den = 3.;
      if (c == 1) { if (per) pL = p - 4 * s; else pR = p, pL = p - s, den = 1.; }
      else if (c == 2 || c == m - 3) { if (!per) pR = p, pL = p - s, den = 1.; }
      else if (c == m - 2) { if (per) pR = p + 3 * s; else pR = p, pL = p - s, den = 1.; }
      if (nv[p - s] + nv[p] + nv[p + s] > 0.1) pR = p, pL = p - s, den = 1.;
      const double inv = 1. / den;
      if (d.second_order) pR = p, pL = p - s;
      else for (int a = 0; a < 3; a++) div4[a] += (d.s[S_FC1 + 3 * D + a][pR] - d.s[S_FC1 + 3 * D + a][pL]) * inv;
      for (int a = 0; a < 3; a++) {
        const double *A = d.s[S_ADV1 + 3 * D + a];
        adv[D][a] = 9. / 8. * 0.5 * (A[p] + A[p - s]) - 1. / 8. * 0.5 * (A[pR] + A[pL]);
      }
    }
    for (int a = 0; a < 3; a++) {
      double v = d.second_order ? div[a] : (9. / 8.) * div[a] + (-1. / 8.) * div4[a];
      v += adv[0][a]; v += adv[1][a]; v += adv[2][a];
      out[a] = v + vis[a];
    }
  } else if (!d.second_order) {
    for (int D = 0; D < 3; D++) {
      const int c = cc[D], m = mm[D], per = pp[D];
      const long s = st[D];
      long pR = p + s, pL = p - 2 * s;
      double den = 3.;
      if (c == 1) { if (per) pL = p - 4 * s; else pR = p, pL = p - s, den = 1.; }
      else if (c == 2 || c == m - 3) { if (!per) pR = p, pL = p - s, den = 1.; }
      else if (c == m - 2) { if (per) pR = p + 3 * s; else pR = p, pL = p - s, den = 1.; }
      if (nv[p - s] + nv[p] + nv[p + s] > 0.1) pR = p, pL = p - s, den = 1.;
      const double inv = 1. / den;
      for (int a = 0; a < 3; a++) div4[a] += (d.s[S_FC1 + 3 * D + a][pR] - d.s[S_FC1 + 3 * D + a][pL]) * inv;
    }
    for (int a = 0; a < 3; a++) out[a] = (9. / 8.) * div[a] + (-1. / 8.) * div4[a] + vis[a];
  } else {
    for (int a = 0; a < 3; a++) out[a] = div[a] + vis[a];
  }
}
struct FpCell {
  VfsDev d;
  VFS_HD void operator()(int i, int j, int k) const {
    const long p = d.idx(i, j, k);
    double f[3];
    fp_cell_value(d, i, j, k + d.kofs, p, f);
    for (int a = 0; a < 3; a++) d.s[S_FP0 + a][p] = f[a];
  }
};
struct FpCellX {      // with the inviscid / skew variants (staged path only)
  VfsDev d;
  VFS_HD void operator()(int i, int j, int k) const {
    const long p = d.idx(i, j, k);
    double f[3];
    fp_cell_value<true>(d, i, j, k + d.kofs, p, f);
    for (int a = 0; a < 3; a++) d.s[S_FP0 + a][p] = f[a];
  }
};


// FpCell for TWO cells (i0, i0 + 1), i0 even: every operand row is read with 16-byte loads (the padded index of an
// even i is even, rows are 128-byte aligned).  Fast path: both cells in the regular interior (3 <= c <= m-4 in every
// direction, so none of the end-of-domain / periodic-seam cases of momentum.c:1565-1637 applies) and no nvert != 0
// within +-2 (VfsDev::near) — the 4th-order differences then use the fixed offsets +1 / -2 and the arithmetic is
// fp_cell_value's, operand for operand.  Everything else falls back to fp_cell_value per cell.
#if defined(__CUDA_ARCH__)
#define VFS_LD2(ptr) (*reinterpret_cast<const double2 *>(ptr))
#endif
struct FpCell2 {
  VfsDev d;
  VFS_HD void operator()(int ii, int j, int k) const {
    const int i0 = 2 * ii, kg = k + d.kofs;
    const long p = d.idx(i0, j, k);
#if defined(__CUDA_ARCH__)
    const bool regular = i0 >= 4 && i0 + 1 <= d.mx - 4 && j >= 3 && j <= d.my - 4 && kg >= 3 && kg <= d.mz - 4 && d.near[p] == 0 && d.near[p + 1] == 0;
    if (regular) {
      const long sj = d.sj, sk = d.sk;
      double o0[3], o1[3];
#pragma unroll
      for (int a = 0; a < 3; a++) {
        const double *c1 = d.s[S_FC1 + a] + p, *c2 = d.s[S_FC2 + a] + p, *c3 = d.s[S_FC3 + a] + p;
        const double *v1 = d.s[S_FV1 + a] + p, *v2 = d.s[S_FV2 + a] + p, *v3 = d.s[S_FV3 + a] + p;
        const double2 am = VFS_LD2(c1 - 2), a0 = VFS_LD2(c1), ap = VFS_LD2(c1 + 2);          // FC1 at i0-2 .. i0+3
        const double2 bm2 = VFS_LD2(c2 - 2 * sj), bm1 = VFS_LD2(c2 - sj), b0 = VFS_LD2(c2), bp1 = VFS_LD2(c2 + sj);
        const double2 cm2 = VFS_LD2(c3 - 2 * sk), cm1 = VFS_LD2(c3 - sk), c0 = VFS_LD2(c3), cp1 = VFS_LD2(c3 + sk);
        const double2 vm = VFS_LD2(v1 - 2), v0 = VFS_LD2(v1);
        const double2 wm1 = VFS_LD2(v2 - sj), w0 = VFS_LD2(v2), xm1 = VFS_LD2(v3 - sk), x0 = VFS_LD2(v3);
        // cell i0
        {
          const double div = (a0.x - am.y + b0.x - bm1.x + c0.x - cm1.x);
          const double vis = (v0.x - vm.y + w0.x - wm1.x + x0.x - xm1.x);
          double out;
          if (!d.second_order) {
            const double inv = 1. / 3.;
            double div4 = 0;
            div4 += (a0.y - am.x) * inv; div4 += (bp1.x - bm2.x) * inv; div4 += (cp1.x - cm2.x) * inv;
            out = (9. / 8.) * div + (-1. / 8.) * div4 + vis;
          } else out = div + vis;
          o0[a] = out;
        }
        // cell i0 + 1
        {
          const double div = (a0.y - a0.x + b0.y - bm1.y + c0.y - cm1.y);
          const double vis = (v0.y - v0.x + w0.y - wm1.y + x0.y - xm1.y);
          double out;
          if (!d.second_order) {
            const double inv = 1. / 3.;
            double div4 = 0;
            div4 += (ap.x - am.y) * inv; div4 += (bp1.y - bm2.y) * inv; div4 += (cp1.y - cm2.y) * inv;
            out = (9. / 8.) * div + (-1. / 8.) * div4 + vis;
          } else out = div + vis;
          o1[a] = out;
        }
      }
#pragma unroll
      for (int a = 0; a < 3; a++) *reinterpret_cast<double2 *>(d.s[S_FP0 + a] + p) = make_double2(o0[a], o1[a]);
      return;
    }
#endif
    for (int q = 0; q < 2; q++) {
      const int i = i0 + q;
      if (i < 1 || i > d.mx - 2) continue;
      double f[3];
      fp_cell_value(d, i, j, kg, p + q, f);
      for (int a = 0; a < 3; a++) d.s[S_FP0 + a][p + q] = f[a];
    }
  }
};

// projection of Fp on the face area vectors (momentum.c:1733-1735) for one node; returns the
// three contravariant components (before scale and masks)
VFS_HD V3 project_fp(const VfsDev &d, long p) {
  const double f0 = d.s[S_FP0][p], f1 = d.s[S_FP1][p], f2 = d.s[S_FP2][p];
  const double ia = d.s[S_IAJ][p];          // 1/aj; face Jacobian = 2/(1/aj_p + 1/aj_q)
  V3 r;
  {
    long q = p + 1;
    double iaj = 2. / (ia + d.s[S_IAJ][q]);
    r.x = (0.5 * (d.s[S_CSI0][p] * f0 + d.s[S_CSI1][p] * f1 + d.s[S_CSI2][p] * f2) +
           0.5 * (d.s[S_CSI0][q] * d.s[S_FP0][q] + d.s[S_CSI1][q] * d.s[S_FP1][q] + d.s[S_CSI2][q] * d.s[S_FP2][q])) * iaj;
  }
  {
    long q = p + d.sj;
    double jaj = 2. / (ia + d.s[S_IAJ][q]);
    r.y = (0.5 * (d.s[S_ETA0][p] * f0 + d.s[S_ETA1][p] * f1 + d.s[S_ETA2][p] * f2) +
           0.5 * (d.s[S_ETA0][q] * d.s[S_FP0][q] + d.s[S_ETA1][q] * d.s[S_FP1][q] + d.s[S_ETA2][q] * d.s[S_FP2][q])) * jaj;
  }
  {
    long q = p + d.sk;
    double kaj = 2. / (ia + d.s[S_IAJ][q]);
    r.z = (0.5 * (d.s[S_ZET0][p] * f0 + d.s[S_ZET1][p] * f1 + d.s[S_ZET2][p] * f2) +
           0.5 * (d.s[S_ZET0][q] * d.s[S_FP0][q] + d.s[S_ZET1][q] * d.s[S_FP1][q] + d.s[S_ZET2][q] * d.s[S_FP2][q])) * kaj;
  }
  return r;
}

// masks of momentum.c:1833-1841 + boundary planes :1866-1938.  bit a set => component a zeroed.
VFS_HD int rhs_mask(const VfsDev &d, int i, int j, int kg, long p) {
  const int mx = d.mx, my = d.my, mz = d.mz;
  if (i == 0 || i == mx - 1 || j == 0 || j == my - 1 || kg == 0 || kg == mz - 1) return 7;
  const double *nv = d.s[S_NV];
  int m = 0;
  if (nv[p] + nv[p + 1] > 0.1 || (!d.perx && i == mx - 2)) m |= 1;
  if (nv[p] + nv[p + d.sj] > 0.1 || (!d.pery && j == my - 2)) m |= 2;
  if (nv[p] + nv[p + d.sk] > 0.1 || (!d.perz && kg == mz - 2)) m |= 4;
  return m;
}

// Formfunction_2 tail: rhs[field s0] += scale * projection, then masks (in place)
struct ProjectAdd {
  VfsDev d; int s0; double scale;
  VFS_HD void operator()(int i, int j, int k) const {
    const int kg = k + d.kofs;
    const long p = d.idx(i, j, k);
    const int m = rhs_mask(d, i, j, kg, p);
    V3 r = mk3(0, 0, 0);
    if (m != 7) r = project_fp(d, p);
    d.s[s0][p] = (m & 1) ? 0. : d.s[s0][p] + scale * r.x;
    d.s[s0 + 1][p] = (m & 2) ? 0. : d.s[s0 + 1][p] + scale * r.y;
    d.s[s0 + 2][p] = (m & 4) ? 0. : d.s[s0 + 2][p] + scale * r.z;
  }
};

// FormFunction_SNES assembly (momentum.c:2297-2331) fused with the projection:
//   Rhs = mask( (-U + U_o)/dt + 0.5 R(U) ) + 0.5 RHS_o - dP [+ F_eul]        (time_coeff()==1)
//   Rhs = mask( (-1.5U + 2U_o - 0.5U_rm1)/dt + R(U) ) - dP [+ F_eul]         (BDF2)
// mask() = the Formfunction_2 zeroing, applied before the last three terms.
// inputs of the assembly for component a at node p, fetched together so the loads overlap
struct SnesIn { double uc, uco, ucm, ro, dp, fe; };
VFS_HD SnesIn snes_inputs(const VfsDev &d, int a, long p) {
  SnesIn v;
  v.uc = d.s[S_UC0 + a][p]; v.uco = d.s[S_UCO0 + a][p]; v.dp = d.s[S_DP0 + a][p];
  v.ucm = d.bdf2 ? d.s[S_UCM0 + a][p] : 0.; v.ro = d.bdf2 ? 0. : d.s[S_RO0 + a][p];
  v.fe = d.has_feul ? d.s[S_FE0 + a][p] : 0.;
  return v;
}
VFS_HD double snes_combine(const VfsDev &d, const SnesIn &in, bool masked, double r) {
  const double dt = d.dt;
  double v;
  if (masked) v = 0.;
  else if (!d.bdf2) {
    v = (-1. / dt) * in.uc;
    v += (1. / dt) * in.uco;
    v += 0.5 * r;
  } else {
    v = (-1.5 / dt) * in.uc;
    v += (2. / dt) * in.uco;
    v += (-0.5 / dt) * in.ucm;
    v += 1.0 * r;
  }
  if (!d.bdf2) v += 0.5 * in.ro;
  v += -1. * in.dp;
  if (d.has_feul) v += 1. * in.fe;
  return v;
}
VFS_HD double snes_assemble(const VfsDev &d, int a, long p, bool masked, double r) { return snes_combine(d, snes_inputs(d, a, p), masked, r); }
struct ProjectSNES {
  VfsDev d;
  VFS_HD void operator()(int i, int j, int k) const {
    const int kg = k + d.kofs;
    const long p = d.idx(i, j, k);
    const int m = rhs_mask(d, i, j, kg, p);
    V3 r = mk3(0, 0, 0);
    if (m != 7) r = project_fp(d, p);
    const double rr[3] = {r.x, r.y, r.z};
    for (int a = 0; a < 3; a++) d.s[S_R0 + a][p] = snes_assemble(d, a, p, (m >> a) & 1, rr[a]);
  }
};

#endif
