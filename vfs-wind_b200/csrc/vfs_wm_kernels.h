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

// ---- wall-function boundary types -1 (smooth, Cabot) and -2 (rough log law) on the i and j sides -------------
// Contra2Cart_2 (rhs.c:311-440): the first interior cell next to such a side gets a modelled velocity: the
// tangential part of the second cell's velocity rescaled to the wall law at the first cell's height, plus (smooth
// law only) the normal part scaled by sb/sc (wallfunction.c:34-59 wall_function, :87-111 wall_function_roughness_
// loglaw).  u_tau goes to lUstar.  Neighbours are read from the snapshot taken before the rules (S_FP0..2: the
// reference reads its lUcat copy), so cells that are first cells of two walls do not see each other's update.
VFS_HD double wm_u_loglaw_rough(double y, double utau, double ks) { return utau * (1. / 0.41 * log(y / ks) + 8.5); }   // wallfunction.c:227-238
VFS_HD double wm_find_utau_rough(double u, double y, double guess, double ks) {                                      // wallfunction.c:429-444
  double x = guess, x0 = guess;
  for (int it = 0; it < 30; it++) {
    const double eps = 1.e-7;
    const double df = ((wm_u_loglaw_rough(y, x0 + eps, ks) - u) - (wm_u_loglaw_rough(y, x0 - eps, ks) - u)) / (2 * eps);
    x = x0 - (wm_u_loglaw_rough(y, x0, ks) - u) / df;
    if (fabs(x0 - x) < 1.e-10) break;
    x0 = x;
  }
  return x;
}
struct C2CWallFn {
  VfsDev d; const double *buf;
  // one wall: D = 0 (i sides) / 1 (j sides), type -1 / -2, `far` = the side at m-2 (normal flipped, neighbour at -1)
  VFS_HD void apply(long p, int D, int type, bool far) const {
    const int sm = S_CSI0 + 3 * D;
    const double ax = d.s[sm][p], ay = d.s[sm + 1][p], az = d.s[sm + 2][p];
    const double area = sqrt(ax * ax + ay * ay + az * az);
    const long nb = p + (far ? -1 : 1) * (D == 0 ? 1 : d.sj);
    const double sb = 0.5 / d.s[S_AJ][p] / area;
    const double sc = 2 * sb + 0.5 / d.s[S_AJ][nb] / area;
    const V3 Uc = ld3(d, S_FP0, nb);
    V3 n = cov_column(d, p, D);                       // Calculate_normal (rhs2.c:614-647)
    const double sum = sqrt(n.x * n.x + n.y * n.y + n.z * n.z);
    n.x /= sum, n.y /= sum, n.z /= sum;
    if (far) { n.x *= -1, n.y *= -1, n.z *= -1; }
    const double nu = 1. / d.ren;
    const double un = Uc.x * n.x + Uc.y * n.y + Uc.z * n.z;
    double ut = Uc.x - un * n.x, vt = Uc.y - un * n.y, wt = Uc.z - un * n.z;
    const double ut_mag = sqrt(ut * ut + vt * vt + wt * wt);
    double ustar, mod;
    if (type == -1) { ustar = wm_find_utau(buf, nu, ut_mag, sc, 0.01); mod = ustar * ustar * wm_integrate_F(buf, nu, ustar, sb); }
    else { ustar = wm_find_utau_rough(ut_mag, sc, 0.01, d.roughness); mod = wm_u_loglaw_rough(sb, ustar, d.roughness); }
    if (ut_mag > 1.e-10) { ut *= mod / ut_mag; vt *= mod / ut_mag; wt *= mod / ut_mag; }
    else ut = vt = wt = 0;
    V3 Ub = mk3(ut, vt, wt);
    if (type == -1) { Ub.x = ut + sb / sc * un * n.x; Ub.y = vt + sb / sc * un * n.y; Ub.z = wt + sb / sc * un * n.z; }
    d.s[S_USTAR][p] = ustar;
    st3(d, S_U0, p, Ub);
  }
  VFS_HD void operator()(int i, int j, int k) const {
    const int kg = d.kglob(k);
    if (kg < 1 || kg > d.mz - 2) return;
    const long p = d.idx(i, j, k);
    if ((int)(d.s[S_NV][p] + 0.1) == 3) return;       // rhs.c:306-309: solid cells are zeroed and skipped
    const int *bc = d.bc;
    for (int t = -1; t >= -2; t--)                      // rhs.c:311-376: smooth, then rough, i sides
      if ((bc[0] == t && i == 1) || (bc[1] == t && i == d.mx - 2)) apply(p, 0, t, i != 1);
    for (int t = -1; t >= -2; t--)                      // rhs.c:378-440: j sides
      if ((bc[2] == t && j == 1) || (bc[3] == t && j == d.my - 2)) apply(p, 1, t, j != 1);
  }
};
// IB_BC (momentum.c:2048-2074): at the first time step of a run without immersed bodies the first cells next to
// a wall-function side (k sides included) become IB nodes, nvert = 1
struct IbBcMarkWall {
  VfsDev d;
  VFS_HD void operator()(int i, int j, int k) const {
    const int kg = k + d.kofs;
    const int *bc = d.bc;
    const bool m = ((bc[0] == -1 || bc[0] == -2) && i == 1) || ((bc[1] == -1 || bc[1] == -2) && i == d.mx - 2) ||
                   ((bc[2] == -1 || bc[2] == -2) && j == 1) || ((bc[3] == -1 || bc[3] == -2) && j == d.my - 2) ||
                   ((bc[4] == -1 || bc[4] == -2) && kg == 1) || ((bc[5] == -1 || bc[5] == -2) && kg == d.mz - 2);
    if (m) d.s[S_NV][d.idx(i, j, k)] = 1;
  }
};
// IB_BC (momentum.c:2169-2189): no flux through the wall face of a wall-function first cell
struct IbBcWallFn {
  VfsDev d;
  VFS_HD void operator()(int i, int j, int k) const {
    const long p = d.idx(i, j, k);
    const int *bc = d.bc;
    if (((bc[0] == -1 || bc[0] == -2) && i == 1)) d.s[S_UC0][p - 1] = 0;
    else if (((bc[1] == -1 || bc[1] == -2) && i == d.mx - 2)) d.s[S_UC0][p] = 0;
    if (((bc[2] == -1 || bc[2] == -2) && j == 1)) d.s[S_UC1][p - d.sj] = 0;
    else if (((bc[3] == -1 || bc[3] == -2) && j == d.my - 2)) d.s[S_UC1][p] = 0;
  }
};
#endif
