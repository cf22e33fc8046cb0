// vfs_halo_kernels.h — layout conversion, DA-wrap ghost fill and periodic node copies.
//
// Replaces, on one rank, what PETSc's DAGlobalToLocal/DALocalToLocal do for the reference
// (ghost width 3, box stencil, Source/init.c:131-160) plus the explicit "if(periodic) ... a=-2 /
// a=mx+1" node copies that follow almost every exchange (e.g. Source/rhs.c:129-156,254-287,
// Source/momentum.c:638-666,1506-1546,1687-1713).
#ifndef VFS_HALO_KERNELS_H
#define VFS_HALO_KERNELS_H
#include "vfs_common.h"

#define VFS_MAXGRP 18
struct Grp { int n; int sid[VFS_MAXGRP]; };

// host AoS [nzl][my][mx][dof]  ->  padded SoA scalars s0..s0+dof-1
struct UnpackAoS {
  VfsDev d; const double *src; int s0, dof;
  VFS_HD void operator()(int i, int j, int k) const {
    long q = (((long)k * d.my + j) * d.mx + i) * dof, p = d.idx(i, j, k);
    for (int c = 0; c < dof; c++) d.s[s0 + c][p] = src[q + c];
  }
};
struct PackAoS {
  VfsDev d; double *dst; int s0, dof;
  VFS_HD void operator()(int i, int j, int k) const {
    long q = (((long)k * d.my + j) * d.mx + i) * dof, p = d.idx(i, j, k);
    for (int c = 0; c < dof; c++) dst[q + c] = d.s[s0 + c][p];
  }
};

// X -> Ucont with the wall-normal flux zeroing of FormFunction_SNES (momentum.c:2264-2289)
struct UnpackX {
  VfsDev d; const double *src;
  VFS_HD void operator()(int i, int j, int k) const {
    const int mx = d.mx, my = d.my, mz = d.mz, kg = k + d.kofs;
    long q = (((long)k * my + j) * mx + i) * 3, p = d.idx(i, j, k);
    double x = src[q], y = src[q + 1], z = src[q + 2];
    const bool jin = (j != 0 && j != my - 1), kin = (kg != 0 && kg != mz - 1), iin = (i != 0 && i != mx - 1);
    if ((i == 0 && d.bc[0] == 1) || (i == mx - 2 && d.bc[1] == 1)) x = 0;
    if (d.bc[0] == 10 && i == 0 && jin && kin) x = 0;
    if (d.bc[1] == 10 && i == mx - 2 && jin && kin) x = 0;
    if ((j == 0 && d.bc[2] == 1) || (j == my - 2 && d.bc[3] == 1)) y = 0;
    if (j == my - 2 && (d.bc[3] == 2 || d.bc[3] == 12)) y = 0;
    if (j == 0 && d.bc[2] == 12) y = 0;
    if (d.bc[2] == 10 && j == 0 && iin && kin) y = 0;
    if ((d.bc[3] == 10 || d.bc[3] == -10) && j == my - 2 && iin && kin) y = 0;
    if ((kg == 0 && d.bc[4] == 1) || (kg == mz - 2 && d.bc[5] == 1)) z = 0;
    d.s[S_UC0][p] = x; d.s[S_UC1][p] = y; d.s[S_UC2][p] = z;
  }
};

// DA-wrap ghost fill in one direction.  Launched over ii in [0,2G) x full padded extent of the
// other two directions as given by the launch box (box coordinates are logical indices except
// in `dir`, where the box coordinate is the ghost counter ii).
struct WrapFill {
  VfsDev d; Grp g; int dir;
  VFS_HD void operator()(int a, int b, int c) const {
    int i = a, j = b, k = c; long src;
    if (dir == 0) { i = a < VFS_G ? a - VFS_G : d.mx + (a - VFS_G); src = d.idx(i < 0 ? i + d.mx : i - d.mx, j, k); }
    else if (dir == 1) { j = b < VFS_G ? b - VFS_G : d.my + (b - VFS_G); src = d.idx(i, j < 0 ? j + d.my : j - d.my, k); }
    else { k = c < VFS_G ? c - VFS_G : d.mz + (c - VFS_G); src = d.idx(i, j, k < 0 ? k + d.mz : k - d.mz); }
    long p = d.idx(i, j, k);
    for (int n = 0; n < g.n; n++) d.s[g.sid[n]][p] = d.s[g.sid[n]][src];
  }
};

// "if(flag) f[k][j][i] = f[c][b][a]" with a=-2 / mx+1 etc.  Sources are always ghost nodes, so
// the copy is race-free in place.  Launched on the two boundary planes of each periodic direction.
struct NodeCopy {
  VfsDev d; Grp g; int kw;      // kw: evaluate ghost planes across the periodic seam as the planes they image (VfsDev::kglob)
  VFS_HD void operator()(int i, int j, int k) const {
    const int kg = kw ? d.kglob(k) : k + d.kofs;
    int a = i, b = j, c = k, flag = 0;
    if (d.perx) { if (i == 0) a = -2, flag = 1; else if (i == d.mx - 1) a = d.mx + 1, flag = 1; }
    if (d.pery) { if (j == 0) b = -2, flag = 1; else if (j == d.my - 1) b = d.my + 1, flag = 1; }
    if (d.perz) { if (kg == 0) c = k - 2, flag = 1; else if (kg == d.mz - 1) c = k + 2, flag = 1; }
    if (!flag) return;
    long p = d.idx(i, j, k), q = d.idx(a, b, c);
    for (int n = 0; n < g.n; n++) d.s[g.sid[n]][p] = d.s[g.sid[n]][q];
  }
};

// near-solid byte mask (VfsDev::near): 1 where any node of the 5x5x5 cube around the node has nvert != 0.
// Evaluated on the owned nodes grown by 2 (the cube then stays inside the G = 4 ghost frame, whose
// nvert values are the wrap / neighbour-rank images); everything outside keeps the initial 1.
struct NearSolid {
  VfsDev d; unsigned char *out;
  VFS_HD void operator()(int i, int j, int k) const {
    const long p = d.idx(i, j, k);
    const double *nv = d.s[S_NV];
    bool any = false;
    for (int c = -2; c <= 2; c++) for (int b = -2; b <= 2; b++) for (int a = -2; a <= 2; a++) any = any || (nv[p + c * d.sk + b * d.sj + a] != 0.);
    out[p] = any ? 1 : 0;
  }
};

struct FillScalar {
  VfsDev d; int sid; double v;
  VFS_HD void operator()(int i, int j, int k) const { d.s[sid][d.idx(i, j, k)] = v; }
};

#endif
