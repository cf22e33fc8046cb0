// vfs_fused_kernels.h — shared-memory-tiled k-marching kernels (the performance path).
//
// A thread block owns an (i,j) tile and marches along k; per-node quantities that the one-thread-
// per-cell kernels of vfs_les_kernels.h / vfs_rhs_kernels.h recompute for each of a cell's 27
// neighbours are computed ONCE per node per plane into a ring of shared-memory planes.  The
// arithmetic is the same device functions in the same summation order as the staged kernels, so
// both forms are bitwise identical (tests/test_gpu_parity.py::test_tiled_equals_staged).
// CUDA only: the host emulation (tests/emu, -DVFS_EMU) always uses the staged kernels.
#ifndef VFS_FUSED_KERNELS_H
#define VFS_FUSED_KERNELS_H
#include "vfs_common.h"
#include "vfs_les_kernels.h"

#ifndef VFS_EMU
#include <cuda_runtime.h>

// ---- LES pass 2 (les.c:308-669) ------------------------------------------------------------------
// ring of 3 planes x 16 per-node products x (TX+2)x(TY+2) nodes  (130.6 KB for 32x8)
template <int TX, int TY>
__global__ void __launch_bounds__(TX *TY) k_les2_tile(VfsDev d, int kbeg, int kend, int kchunk) {
  extern __shared__ double sm[];
  constexpr int NXP = TX + 2, NYP = TY + 2, NN = NXP * NYP, NV = VFS_LES2_NV;
  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * TX + tx;
  const int i0 = 1 + blockIdx.x * TX, j0 = 1 + blockIdx.y * TY;
  const int ka = kbeg + blockIdx.z * kchunk;
  const int kb = min(kend, ka + kchunk);
  if (ka >= kb) return;
  const int i = i0 + tx, j = j0 + ty;
  const bool active = (i <= d.mx - 2) && (j <= d.my - 2);

  auto fill = [&](int kk) {
    const int slot = (kk - (ka - 1)) % 3;
    double *base = sm + (size_t)slot * NV * NN;
    for (int n = tid; n < NN; n += TX * TY) {
      const int jj = n / NXP, ii = n - jj * NXP;
      const int gi = i0 - 1 + ii, gj = j0 - 1 + jj;
      double v[NV];
      if (gi <= d.mx - 1 && gj <= d.my - 1) les2_products(d, d.idx(gi, gj, kk), v);
      else { for (int a = 0; a < NV; a++) v[a] = 0.; }
#pragma unroll
      for (int a = 0; a < NV; a++) base[a * NN + n] = v[a];
    }
  };
  fill(ka - 1);
  fill(ka);
  for (int k = ka; k < kb; k++) {
    fill(k + 1);
    __syncthreads();
    if (active) {
      const long p = d.idx(i, j, k);
      if (d.s[S_NV][p] > 1.1) { d.s[S_LM][p] = 0; d.s[S_MM][p] = 0; }
      else {
        double fs[NV], sum_weight = 0;
#pragma unroll
        for (int a = 0; a < NV; a++) fs[a] = 0;
#pragma unroll
        for (int r = -1; r <= 1; r++) {
          const double *pl = sm + (size_t)((k + r - (ka - 1)) % 3) * NV * NN;
#pragma unroll
          for (int q = -1; q <= 1; q++) {
#pragma unroll
            for (int pp = -1; pp <= 1; pp++) {
              const int n = (ty + 1 + q) * NXP + (tx + 1 + pp);
              const double w = pl[n];
              sum_weight += w * (0.125 * (r == 0 ? 2. : 1.) * (q == 0 ? 2. : 1.) * (pp == 0 ? 2. : 1.));
              const double sw = simpson_w(r, q, pp) * w;
              fs[0] += sw;
#pragma unroll
              for (int a = 1; a < NV; a++) fs[a] += sw * pl[a * NN + n];
            }
          }
        }
        les2_finish(d, i, j, k + d.kofs, p, fs, sum_weight);
      }
    }
    __syncthreads();
  }
}

static inline int launch_les2_tile(cudaStream_t st, const VfsDev &d, int k0, int k1, long *launches) {
  constexpr int TX = 32, TY = 8;
  if (k1 <= k0) return 0;
  const size_t smem = (size_t)3 * VFS_LES2_NV * (TX + 2) * (TY + 2) * sizeof(double);
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(k_les2_tile<TX, TY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -2;
    attr_set = true;
  }
  const int kchunk = 32;
  dim3 grd((d.mx - 2 + TX - 1) / TX, (d.my - 2 + TY - 1) / TY, (k1 - k0 + kchunk - 1) / kchunk), blk(TX, TY, 1);
  k_les2_tile<TX, TY><<<grd, blk, smem, st>>>(d, k0, k1, kchunk);
  (*launches)++;
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}
#endif  // !VFS_EMU

// fused residual kernel: not built yet (the staged FaceFlux/FpCell/Project kernels are used)
static inline bool fused_rhs_applicable(const VfsDev &) { return false; }
template <class S> static inline int launch_fused_rhs(S, const VfsDev &, int, int, double, long *) { return -3; }
#endif
