// vfs_fused_kernels.h — shared-memory-tiled k-marching kernels (the performance path).
//
// A thread block owns an (i,j) tile and marches along k.  The per-node planes a cell's 27-point
// neighbourhood needs are staged into a ring of shared-memory planes by TMA
// (cp.async.bulk.tensor.4d, one box of (TX+2)x(TY+2) doubles per scalar per plane, completion on
// an mbarrier), STAGES-2 planes ahead of the plane being computed, so HBM/L2 latency is hidden by
// the prefetch distance instead of by occupancy.  Out-of-range box coordinates (tile overhang) are
// zero-filled by the TMA unit.  The arithmetic is the same device functions, in the same summation
// order, as the one-thread-per-cell kernels, so both forms are bitwise identical
// (tests/test_gpu_parity.py::test_tiled_equals_staged).
// CUDA only: the host emulation (tests/emu, -DVFS_EMU) always uses the staged kernels.
#ifndef VFS_FUSED_KERNELS_H
#define VFS_FUSED_KERNELS_H
#include "vfs_common.h"
#include "vfs_les_kernels.h"
#include "vfs_rhs_kernels.h"

#ifndef VFS_EMU
#include <cuda.h>
#include <cuda_runtime.h>

#define VFS_TILE_TX 32
#define VFS_TILE_TY 8

// ---- TMA / mbarrier primitives (raw PTX) -------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "VFS_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra VFS_DONE_%=;\n"
      "bra VFS_WAIT_%=;\n"
      "VFS_DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// one (box_x, box_y, 1, 1) tile of scalar `w` at padded coordinates (x, y, z) -> shared memory
__device__ __forceinline__ void tma_load_tile(void *dst, const CUtensorMap *map, int x, int y, int z, int w, unsigned long long *bar) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(smem_u32(dst)),
               "l"((unsigned long long)map), "r"(x), "r"(y), "r"(z), "r"(w), "r"(smem_u32(bar))
               : "memory");
}

// host: 4-D tensor map over the whole scalar pool [S_COUNT][nzt][ny][pitch], box (bx, by, 1, 1)
static inline int vfs_make_tensor_map(CUtensorMap *map, void *pool, const VfsDev &d, long scalar_len, int bx, int by) {
  typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void *fn = 0;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) return -1;
  cuuint64_t gdim[4] = {(cuuint64_t)d.pitch, (cuuint64_t)d.ny, (cuuint64_t)d.nzt, (cuuint64_t)S_COUNT};
  cuuint64_t gstr[3] = {(cuuint64_t)d.pitch * 8, (cuuint64_t)d.sk * 8, (cuuint64_t)scalar_len * 8};
  cuuint32_t box[4] = {(cuuint32_t)bx, (cuuint32_t)by, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = ((encode_fn)fn)(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, pool, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -2;
}

// ---- generic ring of TMA-staged planes ----------------------------------------------------------------
// NS scalars per plane, each tile padded to a 128-byte multiple (TMA destination alignment).
template <int TX, int TY, int NS, int STAGES, int HX = 2, int HY = 2> struct PlaneRing {
  static constexpr int NXP = TX + HX, NYP = TY + HY, NN = NXP * NYP;
  static constexpr int TILE_B = ((NN * 8 + 127) / 128) * 128;
  static constexpr int TILE_D = TILE_B / 8;                 // doubles per padded scalar tile
  static constexpr int PLANE_D = TILE_D * NS;
  static constexpr size_t BYTES = (size_t)STAGES * PLANE_D * 8 + STAGES * 8;
};

// ---- LES pass 2 (les.c:308-669) --------------------------------------------------------------------------
// staged scalars per node: ucat(3), w, U(3), |S|S_ij(6)  (vfs_les_kernels.h: les_derive_store)
template <int TX, int TY, int STAGES>
__global__ void __launch_bounds__(TX *TY, 2) k_les2_tma(const __grid_constant__ CUtensorMap tmap, VfsDev d, int kbeg, int kend, int kchunk) {
  typedef PlaneRing<TX, TY, 13, STAGES> R;
  extern __shared__ __align__(128) unsigned char smraw[];
  double *sm = reinterpret_cast<double *>(smraw);
  unsigned long long *bars = reinterpret_cast<unsigned long long *>(smraw + (size_t)STAGES * R::PLANE_D * 8);
  constexpr int NXP = R::NXP, TD = R::TILE_D, NV = VFS_LES2_NV;
  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * TX + tx;
  const int i0 = 1 + blockIdx.x * TX, j0 = 1 + blockIdx.y * TY;
  const int ka = kbeg + blockIdx.z * kchunk;
  const int kb = min(kend, ka + kchunk);
  if (ka >= kb) return;
  const int i = i0 + tx, j = j0 + ty;
  const bool active = (i <= d.mx - 2) && (j <= d.my - 2);
  const int sid[13] = {S_U0, S_U1, S_U2, S_LW, S_LU0, S_LU1, S_LU2, S_LSS0, S_LSS1, S_LSS2, S_LSS3, S_LSS4, S_LSS5};

  if (tid == 0) {
    for (int s = 0; s < STAGES; s++) mbar_init(&bars[s], 1);
    fence_mbar_init();
  }
  __syncthreads();
  // plane kk lives in slot (kk - (ka-1)) % STAGES; its barrier completes phase ((kk-(ka-1))/STAGES)&1
  auto issue = [&](int kk) {
    const int n = kk - (ka - 1), slot = n % STAGES;
    double *dst = sm + (size_t)slot * R::PLANE_D;
    mbar_expect_tx(&bars[slot], 13 * R::NN * 8);
#pragma unroll
    for (int s = 0; s < 13; s++) tma_load_tile(dst + s * TD, &tmap, i0 - 1 + VFS_G, j0 - 1 + VFS_G, kk + VFS_G, sid[s], &bars[slot]);
  };
  auto wait_plane = [&](int kk) {
    const int n = kk - (ka - 1);
    mbar_wait(&bars[n % STAGES], (n / STAGES) & 1);
  };
  if (tid == 0) {
    for (int kk = ka - 1; kk < ka - 1 + STAGES && kk <= kb; kk++) issue(kk);
  }
  wait_plane(ka - 1);
  wait_plane(ka);
  for (int k = ka; k < kb; k++) {
    wait_plane(k + 1);
    if (active) {
      const long p = d.idx(i, j, k);
      if (d.s[S_NV][p] > 1.1) { d.s[S_LM][p] = 0; d.s[S_MM][p] = 0; }
      else {
        double fs[NV], sum_weight = 0;
#pragma unroll
        for (int a = 0; a < NV; a++) fs[a] = 0;
#pragma unroll
        for (int r = -1; r <= 1; r++) {
          const double *pl = sm + (size_t)((k + r - (ka - 1)) % STAGES) * R::PLANE_D;
#pragma unroll
          for (int q = -1; q <= 1; q++) {
#pragma unroll
            for (int pp = -1; pp <= 1; pp++) {
              const int n = (ty + 1 + q) * NXP + (tx + 1 + pp);
              const double w = pl[3 * TD + n];
              sum_weight += w * (0.125 * (r == 0 ? 2. : 1.) * (q == 0 ? 2. : 1.) * (pp == 0 ? 2. : 1.));
              const double sw = simpson_w(r, q, pp) * w;
              fs[0] += sw;
              const double u0 = pl[n], u1 = pl[TD + n], u2 = pl[2 * TD + n];
              const double U0 = pl[4 * TD + n], U1 = pl[5 * TD + n], U2 = pl[6 * TD + n];
              fs[1] += sw * (U0 * u0); fs[2] += sw * (U0 * u1); fs[3] += sw * (U0 * u2);
              fs[4] += sw * (U1 * u0); fs[5] += sw * (U1 * u1); fs[6] += sw * (U1 * u2);
              fs[7] += sw * (U2 * u0); fs[8] += sw * (U2 * u1); fs[9] += sw * (U2 * u2);
#pragma unroll
              for (int a = 0; a < 6; a++) fs[10 + a] += sw * pl[(7 + a) * TD + n];
            }
          }
        }
        les2_finish(d, i, j, k + d.kofs, p, fs, sum_weight);
      }
    }
    __syncthreads();                         // everyone is done reading plane k-1: its slot is free
    if (tid == 0) {
      const int kn = k - 1 + STAGES;         // next plane for that slot
      if (kn <= kb) { fence_proxy_async(); issue(kn); }
    }
  }
}

static inline int launch_les2_tma(cudaStream_t st, const CUtensorMap &tmap, const VfsDev &d, int k0, int k1, long *launches) {
  constexpr int TX = VFS_TILE_TX, TY = VFS_TILE_TY, STAGES = 3;   // 3 x 36.6 KB: two blocks per SM
  if (k1 <= k0) return 0;
  const size_t smem = PlaneRing<TX, TY, 13, STAGES>::BYTES;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(k_les2_tma<TX, TY, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -2;
    attr_set = true;
  }
  const int kchunk = 64;
  dim3 grd((d.mx - 2 + TX - 1) / TX, (d.my - 2 + TY - 1) / TY, (k1 - k0 + kchunk - 1) / kchunk), blk(TX, TY, 1);
  k_les2_tma<TX, TY, STAGES><<<grd, blk, smem, st>>>(tmap, d, k0, k1, kchunk);
  (*launches)++;
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}
// ---- face fluxes of Formfunction_2 (momentum.c:669-1451), regular faces ---------------------------------
// One thread per node computes its i-, j- and k-face fluxes (Fc, Fv: 18 doubles) from ucat/nvert
// planes P-1..P+2 staged by TMA: box (TX+4) x (TY+3) with origin (i0-1, j0-1), i.e. node offsets
// -1..TX+2 in i (4th-order stencil of the i-face) and -1..TY+1 in j.  Faces with index 0 or m-2 along
// their normal (domain-end / periodic-end stencils) are left to the staged FaceFlux<D> kernels, which
// the host runs on those thin slabs only.
#define VFS_FLUX_HX 4
#define VFS_FLUX_HY 3
template <int TX, int TY, int STAGES> struct SmemAcc {
  typedef PlaneRing<TX, TY, 4, STAGES, VFS_FLUX_HX, VFS_FLUX_HY> R;
  const double *sm; int n0, x, y;      // n0: ring index of the node's own plane; (x,y): tile position incl. halo
  __device__ __forceinline__ double get(int s, int di, int dj, int dk) const {
    return sm[(size_t)((n0 + dk) % STAGES) * R::PLANE_D + s * R::TILE_D + (y + dj) * R::NXP + (x + di)];
  }
  __device__ __forceinline__ double u(int a, int di, int dj, int dk) const { return get(a, di, dj, dk); }
  __device__ __forceinline__ double nv(int di, int dj, int dk) const { return get(3, di, dj, dk); }
};

template <int TX, int TY, int STAGES>
__global__ void __launch_bounds__(TX *TY, 2) k_flux_tma(const __grid_constant__ CUtensorMap tmap, VfsDev d, int kbeg, int kend, int kchunk) {
  typedef PlaneRing<TX, TY, 4, STAGES, VFS_FLUX_HX, VFS_FLUX_HY> R;
  extern __shared__ __align__(128) unsigned char smraw[];
  double *sm = reinterpret_cast<double *>(smraw);
  unsigned long long *bars = reinterpret_cast<unsigned long long *>(smraw + (size_t)STAGES * R::PLANE_D * 8);
  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * TX + tx;
  const int i0 = 1 + blockIdx.x * TX, j0 = 1 + blockIdx.y * TY;
  const int ka = kbeg + blockIdx.z * kchunk;
  const int kb = min(kend, ka + kchunk);
  if (ka >= kb) return;
  const int i = i0 + tx, j = j0 + ty;
  const bool active = (i <= d.mx - 2) && (j <= d.my - 2);
  const int sid[4] = {S_U0, S_U1, S_U2, S_NV};
  if (tid == 0) {
    for (int s = 0; s < STAGES; s++) mbar_init(&bars[s], 1);
    fence_mbar_init();
  }
  __syncthreads();
  // ring index n = kk - (ka-1): plane kk in slot n % STAGES, barrier phase (n / STAGES) & 1
  auto issue = [&](int kk) {
    const int n = kk - (ka - 1), slot = n % STAGES;
    double *dst = sm + (size_t)slot * R::PLANE_D;
    mbar_expect_tx(&bars[slot], 4 * R::NN * 8);
#pragma unroll
    for (int s = 0; s < 4; s++) tma_load_tile(dst + s * R::TILE_D, &tmap, i0 - 1 + VFS_G, j0 - 1 + VFS_G, kk + VFS_G, sid[s], &bars[slot]);
  };
  auto wait_plane = [&](int kk) {
    const int n = kk - (ka - 1);
    mbar_wait(&bars[n % STAGES], (n / STAGES) & 1);
  };
  const int klast = kb - 1 + 2;                   // last plane any step needs
  if (tid == 0) {
    for (int kk = ka - 1; kk < ka - 1 + STAGES && kk <= klast; kk++) issue(kk);
  }
  wait_plane(ka - 1); wait_plane(ka); wait_plane(ka + 1);
  for (int k = ka; k < kb; k++) {
    wait_plane(k + 2);
    if (active) {
      const int kg = k + d.kofs;
      const long p = d.idx(i, j, k);
      SmemAcc<TX, TY, STAGES> A = {sm, k - (ka - 1), tx + 1, ty + 1};
      double fc[3], fv[3];
      if (i <= d.mx - 3) {
        face_flux_core<0, true>(d, A, p, i, fc, fv);
#pragma unroll
        for (int a = 0; a < 3; a++) { d.s[S_FC1 + a][p] = fc[a]; d.s[S_FV1 + a][p] = fv[a]; }
      }
      if (j <= d.my - 3) {
        face_flux_core<1, true>(d, A, p, j, fc, fv);
#pragma unroll
        for (int a = 0; a < 3; a++) { d.s[S_FC2 + a][p] = fc[a]; d.s[S_FV2 + a][p] = fv[a]; }
      }
      if (kg <= d.mz - 3) {
        face_flux_core<2, true>(d, A, p, kg, fc, fv);
#pragma unroll
        for (int a = 0; a < 3; a++) { d.s[S_FC3 + a][p] = fc[a]; d.s[S_FV3 + a][p] = fv[a]; }
      }
    }
    __syncthreads();                         // plane k-1 is no longer needed by anyone
    if (tid == 0) {
      const int kn = k - 1 + STAGES;
      if (kn <= klast) { fence_proxy_async(); issue(kn); }
    }
  }
}

#define VFS_FLUX_STAGES 5
static inline int launch_flux_tma(cudaStream_t st, const CUtensorMap &tmap, const VfsDev &d, int k0, int k1, long *launches) {
  constexpr int TX = VFS_TILE_TX, TY = VFS_TILE_TY, STAGES = VFS_FLUX_STAGES;
  if (k1 <= k0) return 0;
  const size_t smem = PlaneRing<TX, TY, 4, STAGES, VFS_FLUX_HX, VFS_FLUX_HY>::BYTES;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(k_flux_tma<TX, TY, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -2;
    attr_set = true;
  }
  const int kchunk = 64;
  dim3 grd((d.mx - 2 + TX - 1) / TX, (d.my - 2 + TY - 1) / TY, (k1 - k0 + kchunk - 1) / kchunk), blk(TX, TY, 1);
  k_flux_tma<TX, TY, STAGES><<<grd, blk, smem, st>>>(tmap, d, k0, k1, kchunk);
  (*launches)++;
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}
#endif  // !VFS_EMU

// fused residual kernel: not built yet (the staged FaceFlux/FpCell/Project kernels are used)
static inline bool fused_rhs_applicable(const VfsDev &) { return false; }
template <class S> static inline int launch_fused_rhs(S, const VfsDev &, int, int, double, long *) { return -3; }
#endif
