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

// Single-rank ghost refresh in ONE launch.  mode 1 = DALocalToLocal (wrap fill of every periodic direction);
// mode 3 = wrap fill, periodic node copies (NodeCopy), wrap fill again — the sequence that follows most
// updates in the reference (e.g. rhs.c:251-291).  Each destination node (a ghost, or for mode 3 also a node of
// a periodic boundary plane) takes its final value straight from the interior node the sequence would have
// propagated it from: per periodic direction, ghost x -> x -+ m, then (mode 3) 0 -> m-2 and m-1 -> 1.  Sources
// are interior in every periodic direction and destinations are not, so the in-place update is race-free.
// Visited as three slabs (one per direction: its 2G ghost planes + 2 boundary planes, full extent of the other
// directions); a node in two slabs is written twice with the same value.
struct RefreshSlabs { long n[3]; int ext[3]; int depth; };       // nodes per slab; extent (incl. ghosts if periodic) per direction; ghost layers refreshed
struct RefreshFused {
  VfsDev d; Grp g; int mode; RefreshSlabs S;
  VFS_HD static int map1(int x, int m) { return x < 0 ? x + m : (x >= m ? x - m : x); }
  VFS_HD void node(int i, int j, int k) const {
    int a = i, b = j, c = k;
    if (d.perx) { a = map1(a, d.mx); if (mode == 3) a = a == 0 ? d.mx - 2 : (a == d.mx - 1 ? 1 : a); }
    if (d.pery) { b = map1(b, d.my); if (mode == 3) b = b == 0 ? d.my - 2 : (b == d.my - 1 ? 1 : b); }
    if (d.perz) { c = map1(c, d.mz); if (mode == 3) c = c == 0 ? d.mz - 2 : (c == d.mz - 1 ? 1 : c); }
    if (a == i && b == j && c == k) return;
    const long p = d.idx(i, j, k), q = d.idx(a, b, c);
    for (int n = 0; n < g.n; n++) d.s[g.sid[n]][p] = d.s[g.sid[n]][q];
  }
  // shell coordinate s in [0, 2 depth + 2) of a direction with m nodes -> -depth..0, m-1..m+depth-1
  VFS_HD int shell(int s, int m) const { return s <= S.depth ? s - S.depth : m - 1 + (s - S.depth - 1); }
  VFS_HD void operator()(long tl) const {          // 32-bit index arithmetic (the slabs have far fewer than 2^31 nodes)
    const int lo[3] = {d.perx ? -S.depth : 0, d.pery ? -S.depth : 0, d.perz ? -S.depth : 0};
    const unsigned W = 2 * S.depth + 2, e0 = (unsigned)S.ext[0], e1 = (unsigned)S.ext[1];
    unsigned t = (unsigned)tl;
    const unsigned n0 = (unsigned)S.n[0], n1 = (unsigned)S.n[1];
    if (t < n0) {                                       // i slab: shell x ext[1] x ext[2], shell fastest
      const unsigned r = t / W, s = t - r * W, q = r / e1;
      node(shell((int)s, d.mx), lo[1] + (int)(r - q * e1), lo[2] + (int)q);
    } else if (t < n0 + n1) {                           // j slab
      t -= n0;
      const unsigned r = t / e0, x = t - r * e0, q = r / W;
      node(lo[0] + (int)x, shell((int)(r - q * W), d.my), lo[2] + (int)q);
    } else {                                            // k slab
      t -= n0 + n1;
      const unsigned r = t / e0, x = t - r * e0, q = r / e1;
      node(lo[0] + (int)x, lo[1] + (int)(r - q * e1), shell((int)q, d.mz));
    }
  }
};

// The same copies for the 18 face-flux work scalars (momentum.c:1506-1546), restricted to what FpCell
// consumes: the fluxes of face family D are only ever read along direction D at the cell's own other two
// indices (momentum.c:1565-1668), so family D needs the copies of the D-boundary planes only.
struct NodeCopyFlux {
  VfsDev d; int parts;      // bit 0: the i- and j-plane copies, bit 1: the k-plane copies (which read k ghost planes)
  VFS_HD void operator()(int i, int j, int k) const {
    const int kg = k + d.kofs;
    const long p = d.idx(i, j, k);
    if (!(parts & 1)) goto kpart;
    if (d.perx && (i == 0 || i == d.mx - 1)) {
      const long q = d.idx(i == 0 ? -2 : d.mx + 1, j, k);
      for (int a = 0; a < 3; a++) { d.s[S_FC1 + a][p] = d.s[S_FC1 + a][q]; d.s[S_FV1 + a][p] = d.s[S_FV1 + a][q]; }
    }
    if (d.pery && (j == 0 || j == d.my - 1)) {
      const long q = d.idx(i, j == 0 ? -2 : d.my + 1, k);
      for (int a = 0; a < 3; a++) { d.s[S_FC2 + a][p] = d.s[S_FC2 + a][q]; d.s[S_FV2 + a][p] = d.s[S_FV2 + a][q]; }
    }
  kpart:
    if ((parts & 2) && d.perz && (kg == 0 || kg == d.mz - 1)) {
      const long q = d.idx(i, j, kg == 0 ? k - 2 : k + 2);
      for (int a = 0; a < 3; a++) { d.s[S_FC3 + a][p] = d.s[S_FC3 + a][q]; d.s[S_FV3 + a][p] = d.s[S_FV3 + a][q]; }
    }
  }
};

// the same periodic boundary-node copies for the advective half Adv1-3 of the skew-symmetric form (momentum.c:1540-1544)
struct NodeCopyAdv {
  VfsDev d;
  VFS_HD void operator()(int i, int j, int k) const {
    const int kg = k + d.kofs;
    const long p = d.idx(i, j, k);
    if (d.perx && (i == 0 || i == d.mx - 1)) { const long q = d.idx(i == 0 ? -2 : d.mx + 1, j, k); for (int a = 0; a < 3; a++) d.s[S_ADV1 + a][p] = d.s[S_ADV1 + a][q]; }
    if (d.pery && (j == 0 || j == d.my - 1)) { const long q = d.idx(i, j == 0 ? -2 : d.my + 1, k); for (int a = 0; a < 3; a++) d.s[S_ADV2 + a][p] = d.s[S_ADV2 + a][q]; }
    if (d.perz && (kg == 0 || kg == d.mz - 1)) { const long q = d.idx(i, j, kg == 0 ? k - 2 : k + 2); for (int a = 0; a < 3; a++) d.s[S_ADV3 + a][p] = d.s[S_ADV3 + a][q]; }
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
