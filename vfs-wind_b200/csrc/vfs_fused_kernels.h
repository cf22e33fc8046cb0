// vfs_fused_kernels.h — fused, shared-memory-tiled residual kernel (performance path).
// Placeholder until the k-marching kernel lands: the staged kernels of vfs_rhs_kernels.h are used.
#ifndef VFS_FUSED_KERNELS_H
#define VFS_FUSED_KERNELS_H
#include "vfs_common.h"
static inline bool fused_rhs_applicable(const VfsDev &) { return false; }
template <class S> static inline int launch_fused_rhs(S, const VfsDev &, int, int, double, long *) { return -3; }
#endif
