// vfs_wm_kernels.h — Cabot wall model used by Formfunction_2 at the j = 0 faces when
// `viscosity_wallmodel` is set (Source/momentum.c:1139-1154; Source/wallfunction.c:25-33 wall_function_
// freesurface, :240-262 f_Cabot/df_Cabot, :277-329 nu_t/pre_integrate, :331-376 integrate_F, :395-410
// find_utau_Cabot).  u_tau solves u = u_tau^2 * int_0^y dy/(nu + nu_t) by Newton iteration with a
// central-difference derivative; the integral comes from a table of 500 001 entries (dy+ = 2, 24
// Simpson-3/8 panels each) built once on the device, plus linear interpolation.
#ifndef VFS_WM_KERNELS_H
#define VFS_WM_KERNELS_H
#include "vfs_common.h"
#include "vfs_c2c_kernels.h"

#define VFS_WM_INTERVAL 2
#define VFS_WM_MAXYP 1000000
#define VFS_WM_NYP (VFS_WM_MAXYP / VFS_WM_INTERVAL)

// nu_t / nu of the mixing-length law (wallfunction.c:277-280); pow(x, 2.0) == x*x (correctly rounded)
VFS_HD double wm_nut_ratio(double yplus) {
  const double e = 1. - exp(-yplus / 19.);
  return 0.41 * yplus * (e * e);
}
// Simpson-3/8 over [ya, ya+ydiff] with N panels, exactly as the two loops of wallfunction.c:309-325,356-372
VFS_HD double wm_simpson(double ya, double ydiff, int N) {
  const double dy = ydiff / (double)N;
  double val = 0, ybegin = ya;
  double Eprev = 1. / (1. + wm_nut_ratio(ya + dy * 0));
  for (int k = 0; k < N; k++) {
    const double Enext = 1. / (1. + wm_nut_ratio(ya + dy * (k + 1)));
    const double F1 = 1. / (1. + wm_nut_ratio(ybegin + dy * 1. / 3.));
    const double F2 = 1. / (1. + wm_nut_ratio(ybegin + dy * 2. / 3.));
    val += dy / 3. * (3 * Eprev + 9 * F1 + 9 * F2 + 3 * Enext) / 8.;
    ybegin += dy;
    Eprev = Enext;
  }
  return val;
}
// table, step 1: increment of interval i (1..NYP) into buf[i]; step 2 (one thread): running sum in index order
struct WmTableIntervals {
  double *buf;
  VFS_HD void operator()(int i, int, int) const {
    if (i == 0) { buf[0] = 0.; return; }
    buf[i] = wm_simpson((double)(i - 1) * VFS_WM_INTERVAL, (double)i * VFS_WM_INTERVAL - (double)(i - 1) * VFS_WM_INTERVAL, 24);
  }
};
struct WmTableScan {
  double *buf;
  VFS_HD void operator()(int, int, int) const {
    double acc = 0.;
    for (int i = 1; i <= VFS_WM_NYP; i++) { acc = acc + buf[i]; buf[i] = acc; }
  }
};

// wallfunction.c:331-376
VFS_HD double wm_integrate_F(const double *buf, double nu, double utau, double yb) {
  const double yb_plus = yb * utau / nu;
  if (yb_plus <= (double)VFS_WM_MAXYP) {
    int ib = (int)(yb_plus / (double)VFS_WM_INTERVAL);
    ib = ib < 0 ? 0 : (ib > VFS_WM_NYP - 1 ? VFS_WM_NYP - 1 : ib);      // (a diverged iterate must not index outside the table)
    const double int_b = (buf[ib + 1] - buf[ib]) / (double)VFS_WM_INTERVAL * (yb_plus - (double)ib * VFS_WM_INTERVAL) + buf[ib];
    return (int_b - 0) / utau;
  }
  double val = buf[VFS_WM_NYP];
  val += wm_simpson((double)VFS_WM_MAXYP, yb_plus - (double)VFS_WM_MAXYP, 4);
  return val / utau;
}
VFS_HD double wm_f(const double *buf, double nu, double u, double y, double utau) { return utau * utau * wm_integrate_F(buf, nu, utau, y) - u; }
VFS_HD double wm_find_utau(const double *buf, double nu, double u, double y, double guess) {
  double x = guess, x0 = guess;
  for (int it = 0; it < 30; it++) {
    const double eps = 1.e-7;
    const double df = (wm_f(buf, nu, u, y, x0 + eps) - wm_f(buf, nu, u, y, x0 - eps)) / (2 * eps);
    x = x0 - wm_f(buf, nu, u, y, x0) / df;
    if (fabs(x0 - x) < 1.e-10) break;
    x0 = x;
  }
  return x;
}

// momentum.c:1139-1154 for the face between nodes (i,0,k) and (i,1,k): u_tau -> lUstar at the first cell,
// SGS viscosity override of that face -> S_WM at node (i,0,k).  Launched over the j = 0 plane.
struct WallModelPlane {
  VfsDev d; const double *buf;
  VFS_HD void operator()(int i, int j, int k) const {
    const long p0 = d.idx(i, j, k), p = p0 + d.sj;
    const double ex = d.s[S_ETA0][p], ey = d.s[S_ETA1][p], ez = d.s[S_ETA2][p];
    const double area = sqrt(ex * ex + ey * ey + ez * ez);
    const double sb = 0.5 / d.s[S_AJ][p] / area;
    const V3 Ub = ld3(d, S_U0, p);
    V3 n = cov_column(d, p, 1);                       // Calculate_normal (rhs2.c:614-647): x_eta, y_eta, z_eta, normalised
    const double sum = sqrt(n.x * n.x + n.y * n.y + n.z * n.z);
    n.x /= sum, n.y /= sum, n.z /= sum;
    const double un = Ub.x * n.x + Ub.y * n.y + Ub.z * n.z;
    const double ut = Ub.x - un * n.x, vt = Ub.y - un * n.y, wt = Ub.z - un * n.z;
    const double ut_mag = sqrt(ut * ut + vt * vt + wt * wt);
    const double nu = 1. / d.ren;
    const double ustar = wm_find_utau(buf, nu, ut_mag, sb, 0.01);
    d.s[S_USTAR][p] = ustar;
    double nu_t = ustar * ustar / (Ub.z / sb) - 1. / d.ren;
    if (nu_t < 0.0) nu_t = 0.;
    d.s[S_WM][p0] = nu_t;
  }
};
#endif
