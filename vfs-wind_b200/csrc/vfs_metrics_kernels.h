// vfs_metrics_kernels.h — FormMetrics on the device (reference: Source/metrics.c:12-1354,
// compiled there with `#define NEWMETRIC`, metrics.c:10).
//
// Only the centre metrics csi/eta/zet (area vectors) and aj = 1/J are stored.  With NEWMETRIC
// the face metrics are exactly 0.5*a + 0.5*b of the two adjacent centre values and the face
// Jacobian is 2/(1/a + 1/b) (metrics.c:589-592,697-700,805-808), so the flux kernels recompute
// them on the fly bit-exactly instead of streaming 30 more doubles per cell from HBM.
#ifndef VFS_METRICS_KERNELS_H
#define VFS_METRICS_KERNELS_H
#include "vfs_common.h"

// metrics.c:194-259: Jacobian and centre metrics from the 8 corner nodes of cell (i,j,k);
// node (i,j,k) is the cell's upper corner.
struct MetricsCenter {
  VfsDev d;
  VFS_HD void operator()(int i, int j, int k) const {
    const double *X = d.s[S_X], *Y = d.s[S_Y], *Z = d.s[S_Z];
    long p = d.idx(i, j, k);
    const long I = 1, J = d.sj, K = d.sk;
    // corner offsets: c000 = (k,j,i) ... naming c[dk][dj][di] with 1 meaning "-1"
    const long o000 = p, o010 = p - J, o100 = p - K, o110 = p - K - J;
    const long o001 = p - I, o011 = p - J - I, o101 = p - K - I, o111 = p - K - J - I;
#define VFS_D3(A, res_c, res_e, res_z)                                                         \
    double res_c = 0.25 * (A[o000] + A[o010] + A[o100] + A[o110] - A[o001] - A[o011] - A[o101] - A[o111]); \
    double res_e = 0.25 * (A[o000] + A[o001] + A[o100] + A[o101] - A[o010] - A[o011] - A[o110] - A[o111]); \
    double res_z = 0.25 * (A[o000] + A[o010] + A[o001] + A[o011] - A[o100] - A[o110] - A[o101] - A[o111]);
    VFS_D3(X, dxdc, dxde, dxdz)
    VFS_D3(Y, dydc, dyde, dydz)
    VFS_D3(Z, dzdc, dzde, dzdz)
#undef VFS_D3
    double det = dxdc * (dyde * dzdz - dzde * dydz) - dydc * (dxde * dzdz - dzde * dxdz) + dzdc * (dxde * dydz - dyde * dxdz);
    d.s[S_AJ][p] = 1. / det;
    d.s[S_CSI0][p] = dyde * dzdz - dzde * dydz;
    d.s[S_CSI1][p] = -dxde * dzdz + dzde * dxdz;
    d.s[S_CSI2][p] = dxde * dydz - dyde * dxdz;
    d.s[S_ETA0][p] = dydz * dzdc - dzdz * dydc;
    d.s[S_ETA1][p] = -dxdz * dzdc + dzdz * dxdc;
    d.s[S_ETA2][p] = dxdz * dydc - dydz * dxdc;
    d.s[S_ZET0][p] = dydc * dzde - dzdc * dyde;
    d.s[S_ZET1][p] = -dxdc * dzde + dzdc * dxde;
    d.s[S_ZET2][p] = dxdc * dyde - dydc * dxde;
  }
};

// metrics.c:262-342: domain ghost cells mirror the first interior cell, x planes first, then y,
// then z (later passes overwrite edges/corners, as in the reference).  Launched on one plane.
struct MetricsMirror {
  VfsDev d; int dir, side;   // side 0: low plane copies from +1, side 1: high plane copies from -1
  VFS_HD void operator()(int i, int j, int k) const {
    long p = d.idx(i, j, k);
    long q = p + (side ? -1 : 1) * (dir == 0 ? 1 : (dir == 1 ? d.sj : d.sk));
    for (int s = S_CSI0; s <= S_AJ; s++) d.s[s][p] = d.s[s][q];
  }
};

#endif
